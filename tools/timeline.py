"""Development aid: how the branches of the frame graph interleave on the GPU (UG_TIMELINE markers).
usage: UG_TIMELINE=1 python tools/timeline.py [config] [bounces]"""
import ctypes as C, os, sys
os.environ.setdefault("UG_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ugrt  # noqa
import uniformgrid_raytracing_b200 as ug
from uniformgrid_raytracing_b200 import scenes

NAMES = {0: "frame start", 1: "cam grid done", 2: "primary rays done", 3: "shade done", 4: "combine done", 5: "gather done", 6: "ready handshake done", 10: "light grid start", 11: "light grid done",
         12: "regroup start", 13: "regroup done", 14: "shadow done", 20: "ugrid start", 21: "ugrid done", 22: "refl rays start", 23: "refl rays done"}
name = sys.argv[1] if len(sys.argv) > 1 else "B"
bounces = int(sys.argv[2]) if len(sys.argv) > 2 else 1
strip = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1)
s = scenes.SCENES[name]()
r = ug.Renderer()
r.set_scene(s)
L = ug.lib()
L.ug_timeline.restype = C.c_int
L.ug_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p = r.params(s, bounces=bounces, strip=strip)
acc = []
for it in range(8):
    flush.fill_(it)
    torch.cuda.synchronize()
    r.render(p)
    out = (C.c_double * 64)()
    assert L.ug_timeline(r.ctx, out) == 0
    acc.append(np.array(out[:]))
t = np.median(np.stack(acc[3:]), axis=0) / 1000.0
for k in sorted(NAMES, key=lambda k: t[k]):
    if t[k] or k == 0: print(f"{t[k]:8.1f} us  {NAMES[k]}")
