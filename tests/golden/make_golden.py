#!/usr/bin/env python
"""Generate the committed golden vectors from the REFERENCE build (oracle/_ref/libugref.so = the reference's own
kernel sources compiled as host C++, see oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

For each scene: sha256 of every stage buffer, plus the full arrays of the small scenes, in tests/golden/<name>.npz.
tests/test_golden.py checks the oracle (CPU) and the CUDA path (GPU) against these files; nothing at test time
reads /root/reference."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import ugrt  # noqa: E402,F401
from oracle import harness  # noqa: E402
from uniformgrid_raytracing_b200 import scenes  # noqa: E402

STAGES = ["camcoords", "dirtex", "cam_values", "cam_span", "cam_offset", "t", "id", "normal", "dir", "light0_camcoords", "light0_cell",
          "light0_values", "light0_span", "light0_offset", "light1_cell", "light1_values", "light1_span", "light1_offset", "shadowed",
          "rgb_unshadowed", "material_id", "rgb"]


def golden_scenes():
    out = {f"tiny_{k}": scenes.scene_tiny(k) for k in ("one", "tie", "behind", "near", "outside", "ka")}
    out["rand_11"] = scenes.scene_random(11, 800, 96, 64, nlights=1)
    out["a_256"] = scenes.scene_a(256, 256)
    # "sph_" = the light path of the committed display() (spherical ray map + spherical light grid, main.cu:186-192)
    out["sph_tiny_ka"] = scenes.scene_tiny("ka")
    out["sph_rand_11"] = scenes.scene_random(11, 800, 96, 64, nlights=1)
    out["sph_a_256"] = scenes.scene_a(256, 256)
    # "mode+mode+scene": spot = Shader::spotlight_shade (display() from its third frame on), slabN = NUM_SLABS N, sph as above
    out["sph+spot+a_256"] = scenes.scene_a(256, 256)          # the committed display(), frame 3 onwards
    out["spot+rand_11"] = scenes.scene_random(11, 800, 96, 64, nlights=1)
    out["slab4+tiny_ka"] = scenes.scene_tiny("ka")
    out["slab4+a_256"] = scenes.scene_a(256, 256)
    out["sph+slab3+rand_11"] = scenes.scene_random(11, 800, 96, 64, nlights=1)
    # the BASELINE configs at their stated sizes (primary + shadow rays: the part of the frame the reference defines)
    out["b_full"] = scenes.scene_b()          # 80 000 triangles, 1024x1024, 1 light
    out["c_full"] = scenes.scene_c()          # 174 000 triangles, 1920x1080, 2 lights, frame 0 of the dynamic scene
    out["d_full"] = scenes.scene_d()          # 282 000 triangles, 4096x4096, 1 light
    return out


def modes_of(name):
    """{'spherical': bool, 'spot': bool, 'slabs': int} encoded in a fixture name (see golden_scenes)."""
    toks = name.split("+")[:-1] if "+" in name else (["sph"] if name.startswith("sph_") else [])
    slabs = [int(t[4:]) for t in toks if t.startswith("slab")]
    return {"spherical": "sph" in toks, "spot": "spot" in toks, "slabs": slabs[0] if slabs else 1}


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    harness.build("ref")
    ref = harness.load("ref")
    only = set(sys.argv[1:])
    for name, scene in golden_scenes().items():
        if only and name not in only:
            continue
        res = harness.run_frame(ref, scene, threads=0, **modes_of(name))
        ref.set_slabs(1)
        data = {"stages": np.array([s for s in STAGES if s in res]), "sha256": np.array([digest(res[s]) for s in STAGES if s in res]),
                "P": np.array([res["cam_P"], res.get("light0_P", 0)], dtype=np.int64)}
        small = scene.num_pixels <= 96 * 64
        big = scene.num_pixels > 512 * 512
        keep = STAGES if small else (["camcoords", "dirtex"] if big else ["camcoords", "dirtex", "cam_span", "light0_span"])
        for s in keep:
            if s in res:
                data["full_" + s] = np.asarray(res[s])
        if not small:  # a strided sample of the per-pixel buffers
            idx = np.arange(0, scene.num_pixels, 4999 if big else 97)
            data["sample_idx"] = idx
            for s in ("t", "id", "shadowed", "light0_cell"):
                data["sample_" + s] = np.asarray(res[s])[idx]
            data["sample_rgb"] = np.asarray(res["rgb"]).reshape(-1, 3)[idx]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, "P", data["P"].tolist(), "hits", int((res["id"] >= 0).sum()), "shadowed", int(res["shadowed"].sum()))


if __name__ == "__main__":
    main()
