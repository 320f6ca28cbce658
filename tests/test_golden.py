"""Committed golden vectors (tests/golden/*.npz, generated from the reference build by tests/golden/make_golden.py):
the oracle must reproduce them on the CPU, the CUDA path on the GPU.  Nothing here reads /root/reference."""
import glob
import hashlib
import os

import numpy as np
import pytest

import parity
from oracle import harness
from uniformgrid_raytracing_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def modes_of(name):
    """Modes encoded in a fixture name: "sph_x" or "mode+mode+x" with sph (spherical light grid), spot (spotlight_shade), slabN."""
    toks = name.split("+")[:-1] if "+" in name else (["sph"] if name.startswith("sph_") else [])
    slabs = [int(t[4:]) for t in toks if t.startswith("slab")]
    return {"spherical": "sph" in toks, "spot": "spot" in toks, "slabs": slabs[0] if slabs else 1}


def scene_for(name):
    if "+" in name:
        return scene_for(name.split("+")[-1])
    if name.startswith("sph_"):
        return scene_for(name[4:])
    if name.startswith("tiny_"):
        return scenes.scene_tiny(name[5:])
    if name == "rand_11":
        return scenes.scene_random(11, 800, 96, 64, nlights=1)
    if name == "a_256":
        return scenes.scene_a(256, 256)
    if name.endswith("_full"):          # the BASELINE configs at their stated sizes
        return {"b": scenes.scene_b, "c": scenes.scene_c, "d": scenes.scene_d}[name[0]]()
    raise KeyError(name)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check(result, gold, stages=None):
    names, shas = [str(s) for s in gold["stages"]], [str(s) for s in gold["sha256"]]
    for n, h in zip(names, shas):
        if stages is not None and n not in stages:
            continue
        assert n in result, n
        if "full_" + n in gold:
            a, b = np.asarray(result[n]).reshape(-1), gold["full_" + n].reshape(-1)
            assert a.shape == b.shape and np.array_equal(parity.bits(a), parity.bits(b)), f"stage {n} differs from the golden vector"
        assert digest(result[n]) == h, f"stage {n}: sha256 differs from the golden vector"
    if "sample_idx" in gold:
        idx = gold["sample_idx"]
        for s in ("t", "id", "shadowed", "light0_cell"):
            assert np.array_equal(parity.bits(np.asarray(result[s])[idx]), parity.bits(gold["sample_" + s]))
        assert np.array_equal(np.asarray(result["rgb"]).reshape(-1, 3)[idx], gold["sample_rgb"])


def test_golden_files_present():
    assert len(FILES) >= 19


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    name = os.path.basename(path)[:-4]
    gold = np.load(path)
    scene = scene_for(name)
    try:
        res = harness.run_frame(oracle, scene, threads=0, **modes_of(name))
    finally:
        oracle.set_slabs(1)
    assert [res["cam_P"], res.get("light0_P", 0)] == gold["P"].tolist()
    check(res, gold)


GPU_STAGES = ["cam_values", "cam_span", "cam_offset", "t", "id", "normal", "dir", "light0_cell", "light0_values", "light0_span", "light0_offset",
              "light1_cell", "light1_values", "light1_span", "light1_offset", "shadowed", "rgb"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_cuda_reproduces_golden(renderer, ug, path):
    name = os.path.basename(path)[:-4]
    gold = np.load(path)
    scene = scene_for(name)
    renderer.set_scene(scene)
    m = modes_of(name)
    flags = ug.FLAG_FULL_BUFFERS | (ug.FLAG_SPHERICAL_LIGHTS if m["spherical"] else 0) | (ug.FLAG_SPOT_SHADE if m["spot"] else 0)
    res = parity.gpu_frame(renderer, scene, flags=flags, slabs=m["slabs"])
    assert int(renderer.stats().cam_pairs) == int(gold["P"][0])
    check(res, gold, stages=GPU_STAGES)
