#!/usr/bin/env python
"""Print the phase time stamps the fused grid-build kernel leaves in counters[8..] (block 0's view)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ugrt  # noqa: F401
import uniformgrid_raytracing_b200 as ug
from uniformgrid_raytracing_b200 import scenes
s = scenes.SCENES[sys.argv[1] if len(sys.argv) > 1 else "B"]()
r = ug.Renderer(0); r.set_scene(s)
for i in range(5): r.render(r.params(s, frame=i))
r.synchronize()
# counters live right behind... read through a raw cudaMemcpy using torch: get pointer via stats? use ug_device_ptr on a debug id
ptr, n = r.device_ptr(ug.BUF_CAM_COUNTERS)
buf = np.zeros(32, np.uint32)
r._ck(r.L.ug_read(r.ctx, ug.BUF_CAM_COUNTERS, 0, buf.ctypes.data_as(C.c_void_p), buf.nbytes), "read")
st = buf[8:20].astype(np.int64)
names = ["start", "p0 end", "sync1 done", "p1 end", "sync2 done", "offsets done", "base0 done", "scatter0 done", "hist1 done", "sync3 done", "base1 done", "scatter1 done"]
t0 = st[0]
for k, nm in enumerate(names):
    print(f"{nm:16s} +{(st[k]-t0) & 0xffffffff:8d} ns")
print("P", buf[0], "stats", r.stats().ms_frame)
