"""The drop-in boundary: libugrt.so loads without a GPU, exports every function include/ugrt.h declares, its host-only
entry points (camera.h mirror, direction lattice, loader) agree with the oracle, and it FAILS LOUDLY - no CPU fallback -
when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parity
from oracle import harness
from uniformgrid_raytracing_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ugrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z_0-9]*\s*\**\s*\b([A-Za-z_][A-Za-z_0-9]*)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return sorted(set(names))


def test_header_symbols_are_exported(ug):
    lib = C.CDLL(ug.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 40 and {"ug_create", "ug_render_frame", "ug_render_frame_host", "init_model", "build_frustum_grid", "check_for_shadows",
                                 "writePPM", "fillCoordinatesData", "processData"} <= set(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ugrt.h but not exported: {missing}"


def test_no_oracle_in_product(ug):
    """The product library must not link against (or dlopen) anything under oracle/."""
    blob = open(ug.LIB_PATH, "rb").read()
    assert b"libugoracle" not in blob and b"libugref" not in blob and b"orc_" not in blob


def test_gl_presenter_entry_points_check_their_arguments(ug):
    """present_gl.cpp (main.cu:139/262/347): exported, and a null context is refused before anything touches CUDA
    or GL (there is no GL on the build or the GPU image, so the mapped path itself is not exercised)."""
    L = ug.lib()
    assert L.ug_gl_register_pbo(None, 1) == 1      # UG_ERR_ARG
    assert L.ug_gl_map_pbo(None) == 1 and L.ug_gl_unmap_pbo(None) == 1 and L.ug_gl_unregister_pbo(None) == 1


def test_pdl_kernels_wait_first():
    """Programmatic dependent launch (k_common.cuh, -DUG_PDL=1 builds): a kernel that is ever launched through ug_launch may
    start while its predecessor still runs, so ug_pdl_enter<group>() (griddepcontrol.wait) has to be the first statement of its
    body - only __shared__ / constexpr declarations, comments and preprocessor lines may stand in front of it."""
    import glob
    import re
    src_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "uniformgrid-raytracing_b200", "csrc")
    text = {f: open(f).read() for f in glob.glob(os.path.join(src_dir, "*.cu"))}
    launched = set()
    for body in text.values():
        launched |= set(re.findall(r"ug_launch\(\s*UG_PDL_[A-Z]+\s*,\s*(k_[a-z_0-9]+)", body))
    assert {"k_sweep_setup", "k_sweep_pass", "k_trace", "k_regroup", "k_shadow", "k_shade"} <= launched
    for name in sorted(launched):
        defs = [(f, m) for f, body in text.items() for m in re.finditer(r"__global__[^;{]*?\b" + name + r"\s*\(", body)]
        assert len(defs) == 1, f"{name}: {len(defs)} definitions"
        f, m = defs[0]
        rest = text[f][m.end():]
        rest = rest[rest.index("{") + 1:]
        first = None
        in_comment = False
        for ln in rest.split("\n"):
            t = ln.strip()
            if in_comment:
                in_comment = "*/" not in t
                continue
            if t.startswith("/*"):
                in_comment = "*/" not in t
                continue
            if not t or t.startswith("//") or t.startswith("#") or t.startswith("__shared__") or t.startswith("constexpr") or t.startswith("*"):
                continue
            first = t
            break
        assert re.fullmatch(r"ug_pdl_enter<UG_PDL_[A-Z]+>\(\);", first or ""), f"{name} ({os.path.basename(f)}): first statement is {first!r}"


def test_create_fails_loudly_without_gpu(ug):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ug.UgError):
        ug.Renderer(0)
    ctx = C.c_void_p()
    assert ug.lib().ug_create(C.byref(ctx), 0) == 2 and not ctx.value      # UG_ERR_CUDA, no context


@pytest.mark.parametrize("w,h,eye,look,up", [(512, 512, (0, -30, 12), (0, 0, 2), (0, 0, 1)), (1920, 1080, (14.5, -14, 13), (14.5, 13, 2), (0, 0, 1)),
                                              (64, 64, (3, 15, 5), (13, 13, 3), (0, 0, 1))])
def test_camera_matches_oracle(ug, oracle, w, h, eye, look, up):
    """camera.h mirror: modelview/projection/MVP, Gribb-Hartmann planes, near corners, and the 5x5 direction lattice."""
    oracle.configure(w, h, 45, 0)
    want = oracle.camera_setup(eye, look, up, 0.1, 100.0)
    cam = ug.make_camera(eye, look, up, 0.1, 100.0, 45.0, float(np.float32(w) / np.float32(h)))
    got = ug.camcoords(cam)
    assert np.array_equal(parity.bits(got), parity.bits(want))
    assert np.array_equal(parity.bits(ug.direction_lattice(got)), parity.bits(oracle.dirtex()))
    want_l = oracle.camera_setup(eye, look, up, 0.1, 100.0, square=True)     # light cameras: aspect 1
    got_l = ug.camcoords(ug.make_camera(eye, look, up, 0.1, 100.0, 45.0, 1.0))
    assert np.array_equal(parity.bits(got_l), parity.bits(want_l))


def test_frame_params_layout(ug):
    """ctypes mirrors of the PODs in include/ugrt.h have the sizes the C compiler gives them."""
    import subprocess, tempfile, textwrap
    prog = textwrap.dedent('''
        #include <stdio.h>
        #include "ugrt.h"
        int main(void){ printf("%zu %zu %zu %zu\\n", sizeof(ug_camera), sizeof(ug_light), sizeof(ug_frame_params), sizeof(ug_stats)); return 0; }''')
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)   # the header is plain C
        sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    assert sizes == [C.sizeof(ug.Camera), C.sizeof(ug.LightS), C.sizeof(ug.FrameParams), C.sizeof(ug.Stats)]


def test_strip_bounds_cover_the_frame(ug):
    import strips_helpers as strips
    for h, world in [(1024, 1), (1024, 8), (1080, 4), (1080, 8), (64, 3), (4096, 8)]:
        nby = h // 8
        b = [strips.strip_bounds(nby, k, world) for k in range(world)]
        assert b[0][0] == 0 and b[-1][1] == nby and all(b[k][1] == b[k + 1][0] for k in range(world - 1))
        assert sum(strips.strip_nbytes(640, h, world)) == 3 * 640 * h


def test_acosf_restatement_matches_libm(ug, oracle):
    """The spherical light grid (a14/a15) bins by acosf.  The kernels use a restatement of glibc 2.39's float algorithm
    (csrc/k_common.cuh ug_acosf: pure float + - * / sqrtf, so the device reproduces it under -fmad=false); the
    host-compiled reference calls libm (here through the oracle).  Pin one against the other on the host: 8 M bit
    patterns spread over [-1, 1], every float next to the branch points, and the out-of-domain / NaN cases.
    (An exhaustive run over all 2^32 patterns was done once when the restatement was written: 0 mismatches.)"""
    import ctypes as C
    import numpy as np
    rng = np.random.default_rng(7)
    pats = [rng.integers(0, 0x3F800001, 4_000_000, dtype=np.uint32)]          # [0, 1]
    pats.append(pats[0] | np.uint32(0x80000000))                               # [-1, 0]
    for centre in (0x3F000000, 0x3F800000, 0x23000000, 0x00000000, 0x7F800000):
        near = (np.arange(-4096, 4097, dtype=np.int64) + centre).clip(0, 0x7FFFFFFF).astype(np.uint32)
        pats += [near, near | np.uint32(0x80000000)]
    x = np.ascontiguousarray(np.concatenate(pats).view(np.float32))
    got, want = np.empty_like(x), np.empty_like(x)
    fp = C.POINTER(C.c_float)
    ug.lib().ug_acosf_host(x.ctypes.data_as(fp), got.ctypes.data_as(fp), x.size)
    fn = oracle.lib.orc_acosf_array
    fn.restype, fn.argtypes = None, [fp, fp, C.c_long]
    fn(x.ctypes.data_as(fp), want.ctypes.data_as(fp), x.size)
    nan = np.isnan(want)
    assert np.array_equal(nan, np.isnan(got))
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))
    assert nan.sum() > 1000 and (~nan).sum() > 8_000_000


def test_shadow_item_table_bound(ug):
    """The shadow pass sizes its work-item table as N + 17 C entries; that holds only if no light cell ever asks for more than
    rays + 16 items (a spherical light grid with long lists once overflowed an N + C table silently).  Check the policy the
    kernels use (host view of the same function) over the whole range of cell shapes, and that every item's rays and triangles
    tile the cell exactly."""
    import ctypes as C
    import numpy as np
    fn = ug.lib().ug_shadow_item_shape
    out = (C.c_uint * 4)()
    rng = np.random.default_rng(5)
    rays = np.unique(np.concatenate([np.arange(0, 70), 2 ** np.arange(6, 27), 2 ** np.arange(6, 27) - 1, rng.integers(1, 1 << 26, 400)]))
    tris = np.unique(np.concatenate([np.arange(0, 70), [127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097],
                                     2 ** np.arange(12, 21), rng.integers(1, 1 << 20, 300)]))
    for r in rays:
        for t in tris:
            fn(int(r), int(t), out)
            rpi, nseg, seglen, items = out[0], out[1], out[2], out[3]
            if r == 0 or t == 0:
                assert items == 0
                continue
            assert items == -(-int(r) // rpi) * nseg
            assert items <= int(r) + 16, (r, t, rpi, nseg)
            assert 1 <= rpi <= 32 and nseg >= 1 and seglen % 32 == 0 and nseg * seglen >= t and (nseg - 1) * seglen < t + 32 * nseg
