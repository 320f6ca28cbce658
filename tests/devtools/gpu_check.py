#!/usr/bin/env python
"""First-contact GPU check: per-stage parity against the CPU oracle with diagnostics (not a pytest)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from parity import ug, harness
from uniformgrid_raytracing_b200 import scenes

which = sys.argv[1:] or ["tiny", "A"]
orc = harness.load("oracle")
r = ug.Renderer(0)
rc_total = 0
def run(scene, **kw):
    global rc_total
    print(f"== {scene.name}: {scene.num_tris} tris {scene.width}x{scene.height} lights={len(scene.lights)} {kw}")
    cpu = harness.run_frame(orc, scene, threads=0)
    r.set_scene(scene)
    t0 = time.time()
    gpu = parity.gpu_frame(r, scene, **kw)
    st = r.stats()
    print(f"  gpu frame {st.ms_frame:.3f} ms, launches {st.kernel_launches}, P={st.cam_pairs} lightP={list(st.light_pairs)[:len(scene.lights)]} cap={st.pair_capacity} ovf={st.overflowed}")
    cpu = dict(cpu); cpu["sizelist"] = orc_sizelist(scene)
    bad = parity.compare(gpu, cpu)
    if bad:
        rc_total += 1
        for k, v in bad.items():
            if isinstance(v, int):
                a, b = np.asarray(gpu[k]).reshape(-1), np.asarray(cpu[k]).reshape(-1)
                idx = np.flatnonzero(parity.bits(a) != parity.bits(b))[:8]
                print(f"  !! {k}: first diffs at {idx.tolist()} gpu={a[idx].tolist()} cpu={b[idx].tolist()}")
    print("  max rel err t:", parity.rel_err_t(gpu["t"], cpu["t"]))
    return gpu, cpu
def orc_sizelist(scene):
    orc.configure(scene.width, scene.height, scene.fovy, 0)
    orc.camera_setup(scene.eye, scene.look, scene.up, scene.near, scene.far)
    return orc.grid_count(np.ascontiguousarray(scene.faces, np.int32), np.ascontiguousarray(scene.verts, np.float32))[0]

if "tiny" in which:
    for kind in ("one", "tie", "behind", "near", "outside", "ka"):
        run(scenes.scene_tiny(kind))
    run(scenes.scene_random(1, 2000, 128, 64, nlights=2))
    run(scenes.scene_random(2, 500, 96, 160, nlights=1), stepwise=True)
if "A" in which:
    g, c = run(scenes.scene_a())
    run(scenes.scene_a(), flags=ug.FLAG_FULL_BUFFERS | ug.FLAG_NO_GRAPH)
    run(scenes.scene_a(), stepwise=True)
    harness.write_ppm(os.path.join(ROOT, "gpurun_out", "a_gpu.ppm"), g["rgb"], 512, 512)
if "B" in which:
    run(scenes.scene_b())
if "C" in which:
    run(scenes.scene_c())
print("FAILED STAGES" if rc_total else "ALL OK", rc_total)
sys.exit(1 if rc_total else 0)
