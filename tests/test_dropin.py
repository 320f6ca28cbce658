"""The drop-in layer (csrc/ref_compat.cpp): the reference's own entry points - init_model, updateLightPosition,
fillCoordinatesData, setDirectionTexture, build_frustum_grid, getRayGridMapping, getEffectiveRayGridMapping,
build_secondary_frustum_grid, processData, check_for_shadows, writePPM (per_app_funcs.h:8-66, per_frame_funcs.h:6-159) -
executed from a plain C program that follows display() (main.cu:59-302) the way INTEGRATION.md section 2 binds it.
The PPM files it writes are compared with the oracle's frames (spherical light grid; Lambert shader on the first frame,
spot shader afterwards, as display() does)."""
import os
import subprocess

import numpy as np
import pytest

import parity  # noqa: F401
from oracle import harness
from uniformgrid_raytracing_b200 import scenes
import uniformgrid_raytracing_b200 as ugpkg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not os.path.exists(ugpkg.LIB_PATH):
        ugpkg.build()
    out = str(tmp_path_factory.mktemp("dropin") / "display_dropin")
    libdir = os.path.dirname(ugpkg.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-D_DEFAULT_SOURCE", "-O1", "-Wall", "-Werror", os.path.join(HERE, "c", "display_dropin.c"), "-I", os.path.join(ROOT, "include"),
           "-L", libdir, "-l:" + os.path.basename(ugpkg.LIB_PATH), "-Wl,-rpath," + libdir, "-lm", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr       # every reference-named symbol resolves against libugrt.so
    return out


def write_params(path, s):
    l = s.lights[0]
    vals = [s.width, s.height, s.fovy, s.near, s.far, *s.eye, *s.look, *s.up, *l.pos, *l.look, *l.up]
    with open(path, "w") as f:
        f.write(" ".join(repr(float(v)) for v in vals) + "\n")


def read_p3(path, s):
    tok = open(path).read().split()
    assert tok[:4] == ["P3", str(s.width), str(s.height), "255"]
    return np.asarray(tok[4:], dtype=np.int64).astype(np.uint8)


def test_usage_errors_like_the_reference(exe, tmp_path):
    """per_app_funcs.h:10-14, 18-27: no model -> usage + exit(-1); a directory without a frame count -> usage + exit(-1).
    Both exits happen before any CUDA call, so this runs without a GPU."""
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and "USAGE: ./exec {file/dir_name}.obj [number_frames]" in r.stderr
    d = tmp_path / "frames"
    d.mkdir()
    r = subprocess.run([exe, str(d), "mat"], capture_output=True, text=True)
    assert r.returncode == 255 and "USAGE: ./exec {dir_name}.ply [desc_file] [number_frames]" in r.stderr


@pytest.mark.gpu
def test_display_sequence_static_model(exe, oracle, tmp_path):
    """init_model(obj, material) + three display() calls at the reference's 1024x1024."""
    s = scenes.scene_a(1024, 1024)
    obj, mat = scenes.write_obj(s, str(tmp_path / "a.obj"))
    write_params(str(tmp_path / "params.txt"), s)
    r = subprocess.run([exe, obj, mat, "3", str(tmp_path / "params.txt"), str(tmp_path / "img")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    V, F, M = ugpkg.parse_obj(obj)
    k = scenes.Scene(**{**s.__dict__, "verts": V})
    lambert = harness.run_frame(oracle, k, threads=0, spherical=True, spot=False)["rgb"]
    spot = harness.run_frame(oracle, k, threads=0, spherical=True, spot=True)["rgb"]
    assert np.count_nonzero(lambert != spot) > 1000
    assert np.array_equal(read_p3(str(tmp_path / "img-1.ppm"), s), lambert)      # frame_cnt < 2: simpleShade
    assert np.array_equal(read_p3(str(tmp_path / "img-2.ppm"), s), spot)         # then spotlight_shade
    assert np.array_equal(read_p3(str(tmp_path / "img-3.ppm"), s), spot)


@pytest.mark.gpu
def test_display_sequence_keyframe_directory(exe, oracle, tmp_path):
    """init_model(dir, material, frames) - the S_ISDIR branch of per_app_funcs.h:16-30 - then Model::tmp_model per frame."""
    s = scenes.scene_random(23, 400, 128, 96, 1)
    d = tmp_path / "anim"
    d.mkdir()
    _, mat = scenes.write_obj(s, str(d / "f_0.obj"))
    rng = np.random.default_rng(7)
    for i in (1, 2):
        k = scenes.Scene(**{**s.__dict__, "verts": (s.verts + rng.normal(scale=0.2, size=s.verts.shape)).astype(np.float32)})
        scenes.write_obj(k, str(d / f"f_{i}.obj"))
    write_params(str(tmp_path / "params.txt"), s)
    r = subprocess.run([exe, str(d), mat, "3", str(tmp_path / "params.txt"), str(tmp_path / "k")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for i in range(3):
        V, F, M = ugpkg.parse_obj(str(d / f"f_{i}.obj"))
        k = scenes.Scene(**{**s.__dict__, "verts": V})
        want = harness.run_frame(oracle, k, threads=0, spherical=True, spot=i >= 1)["rgb"]
        assert np.array_equal(read_p3(str(tmp_path / f"k-{i + 1}.ppm"), s), want), f"keyframe {i}"
