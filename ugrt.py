"""Import shim: the package directory is named after the reference (`uniformgrid-raytracing_b200/`),
which is not a Python identifier.  `import ugrt` registers it as `uniformgrid_raytracing_b200`."""
import importlib.util
import os
import sys

_NAME = "uniformgrid_raytracing_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "uniformgrid-raytracing_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

pkg = sys.modules[_NAME]
