"""Pin the oracle: oracle/ug_oracle.c (plain-C restatement) must reproduce, bit for bit, what the reference's own
kernel sources compute when compiled as host C++ (oracle/_ref/libugref.so, built from /root/reference by
oracle/Makefile) - every stage of display(): camera block, direction lattice, grid lists, hit ids, t, shadow-ray
cells, shadow bits, RGB.  The reference ships no tests or golden vectors of its own."""
import os

import numpy as np
import pytest

import parity  # noqa: F401  (sets sys.path)
from oracle import harness
from uniformgrid_raytracing_b200 import scenes


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(parity.bits(a.reshape(-1)), parity.bits(b.reshape(-1)))


def check_frame(ref, oracle, scene, frame=0, spherical=False, spot=False, slabs=1):
    want = harness.run_frame(ref, scene, frame=frame, threads=0, spherical=spherical, spot=spot, slabs=slabs)
    got = harness.run_frame(oracle, scene, frame=frame, threads=0, spherical=spherical, spot=spot, slabs=slabs)
    assert set(want) == set(got)
    bad = [k for k in want if not same(want[k], got[k])]
    assert not bad, f"{scene.name}: oracle differs from the reference build in {bad}"
    return want


@pytest.mark.parametrize("kind", ["one", "tie", "behind", "near", "outside", "ka"])
def test_micro_scenes(ref, oracle, kind):
    out = check_frame(ref, oracle, scenes.scene_tiny(kind))
    if kind == "tie":       # coincident triangles: the lower id takes every hit (trace_kernel.cu:38 + stable sort)
        assert (out["id"] == 0).sum() > 0 and (out["id"] == 1).sum() == 0
    if kind == "behind":    # |t| mirrors geometry behind the eye into the image (trace_kernel.cu:35)
        assert (out["id"] == 1).sum() > 0
    if kind == "outside":   # a triangle outside the frustum is clamped into border cells, never culled
        assert out["cam_P"] > 16 and (out["id"] == 1).sum() == 0
    if kind == "ka":        # the occluder casts a shadow on the floor
        assert out["shadowed"].sum() > 0


@pytest.mark.parametrize("seed,ntris,w,h,nl", [(1, 2000, 128, 64, 2), (2, 500, 96, 160, 1), (3, 6000, 256, 256, 1), (4, 50, 64, 64, 0)])
def test_random_soups(ref, oracle, seed, ntris, w, h, nl):
    check_frame(ref, oracle, scenes.scene_random(seed, ntris, w, h, nlights=nl))


@pytest.mark.parametrize("seed", [77, 78])
def test_hostile_geometry(ref, oracle, seed):
    """Triangles through the eye plane / behind the camera / degenerate / duplicated: whatever the reference computes
    (garbage boxes included) the oracle must compute too."""
    out = check_frame(ref, oracle, scenes.scene_hostile(seed))
    assert out["cam_P"] > 5000 and (out["id"] == 5).sum() == 0     # the duplicate (id 5) never beats id 4


@pytest.mark.parametrize("case", ["ka", "a", "soup1", "soup2", "hostile"])
def test_spherical_light_grid(ref, oracle, case):
    """a14/a15 - the light path the committed display() runs (main.cu:186-192): mapSort_Effective_kernel
    (misc_kernel.cu:255-296) + DS_spherical_Kernel/Fillkernel (grid_kernel.cu:481-854) with getEffective_x/y
    (:395-479, including the `a + b*c*d` dot product of :468) executed VERBATIM by the reference build vs the restatement."""
    s = {"ka": lambda: scenes.scene_tiny("ka"), "a": lambda: scenes.scene_a(256, 256),
         "soup1": lambda: scenes.scene_random(1, 2000, 128, 64, nlights=2), "soup2": lambda: scenes.scene_random(3, 6000, 256, 256, nlights=1),
         "hostile": lambda: scenes.scene_hostile(77)}[case]()
    out = check_frame(ref, oracle, s, spherical=True)
    per = harness.run_frame(oracle, s, threads=0)
    assert not same(out["light0_cell"], per["light0_cell"])          # it really is a different binning ...
    assert same(out["id"], per["id"]) and same(out["t"], per["t"])   # ... of the same primary hits
    if case == "a":
        assert out["shadowed"].sum() > 1000
        # the sizeList of the spherical count kernel alone (DS_spherical_Kernel), light camera resident
        for lib in (ref, oracle):
            lib.configure(s.width, s.height, s.fovy, 0)
            lib.camera_setup(s.lights[0].pos, s.lights[0].look, s.lights[0].up, s.near, s.far, square=True)
        pi = np.float32(np.pi)
        a = ref.grid_count(np.ascontiguousarray(s.faces), np.ascontiguousarray(s.verts), spherical=(pi, pi))
        b = oracle.grid_count(np.ascontiguousarray(s.faces), np.ascontiguousarray(s.verts), spherical=(pi, pi))
        assert same(a[0], b[0]) and same(a[1], b[1]) and int(a[0].sum()) == out["light0_P"]


@pytest.mark.parametrize("case", ["ka", "a", "soup1", "hostile"])
def test_spot_shade(ref, oracle, case):
    """Shader::spotlight_shade -> spot_shade (shader_kernel.cu:275-345), the shader display() uses from its third frame on
    (main.cu:205-219), run verbatim by the reference build vs the restatement - with the committed display()'s spherical
    light path in front of it.  The two cone angles the reference dumps per pixel are compared on hit pixels as well."""
    s = {"ka": lambda: scenes.scene_tiny("ka"), "a": lambda: scenes.scene_a(256, 256),
         "soup1": lambda: scenes.scene_random(1, 2000, 128, 64, nlights=2), "hostile": lambda: scenes.scene_hostile(77)}[case]()
    out = check_frame(ref, oracle, s, spherical=True, spot=True)
    hit = np.repeat(out["id"] >= 0, 2)
    assert same(ref.last_spot_angles[hit], oracle.last_spot_angles[hit])
    plain = harness.run_frame(oracle, s, threads=0, spherical=True)
    if case in ("ka", "a"):
        assert not same(out["rgb"], plain["rgb"])       # part of the image lies outside the cone and is dimmed to a quarter
    assert same(out["shadowed"], plain["shadowed"])


@pytest.mark.parametrize("case,slabs,sph", [("ka", 4, False), ("a", 4, False), ("a", 3, True), ("soup1", 8, False), ("soup2", 2, True), ("hostile", 4, False)])
def test_depth_slabs(ref, oracle, case, slabs, sph):
    """NUM_SLABS > 1 (main.cu.h:18 compiles 1, the code for more is all there): SlabKernel (grid_kernel.cu:334-352) and the
    slab part of the DSFillkernel key (:322) run VERBATIM in the reference build, isWithin / intersectTriUV verbatim inside
    the per-tile slab loop of rckernel_alpha (trace_kernel.cu:132-229), the light kernel's slab loop (light_kernel.cu:105).
    What that code computes includes its quirk: a ray that finished in slab s is reset by isWithin in slab s+1 unless the
    whole tile finished - so hits are lost relative to one slab; shadows are an OR over all slabs and keep their meaning."""
    s = {"ka": lambda: scenes.scene_tiny("ka"), "a": lambda: scenes.scene_a(256, 256),
         "soup1": lambda: scenes.scene_random(1, 2000, 128, 64, nlights=2), "soup2": lambda: scenes.scene_random(3, 6000, 256, 256, nlights=1),
         "hostile": lambda: scenes.scene_hostile(77)}[case]()
    try:
        out = check_frame(ref, oracle, s, spherical=sph, slabs=slabs)
        one = harness.run_frame(oracle, s, threads=0, spherical=sph)
    finally:
        ref.set_slabs(1)
        oracle.set_slabs(1)
    C = (s.width // 8) * (s.height // 8)
    assert out["cam_span"].size == C * slabs and out["cam_P"] == one["cam_P"]
    # the slabs partition every cell's list: the same triangles per cell, grouped by slab
    assert np.array_equal(out["cam_span"].reshape(C, slabs).sum(axis=1), one["cam_span"])
    assert np.array_equal(np.sort(out["cam_values"]), np.sort(one["cam_values"]))
    if case == "a":
        assert (out["cam_span"].reshape(C, slabs) > 0).sum(axis=0).min() > 0           # every slab is used
        assert 0 < (out["id"] >= 0).sum() <= (one["id"] >= 0).sum()


@pytest.mark.parametrize("case", ["8x8", "8lights", "nan", "allmiss", "wide"])
def test_edge_sizes(ref, oracle, case):
    """One-cell screen, the maximum number of lights, NaN / inf vertices, a frame without a hit, a 256x1-cell screen."""
    if case == "8x8":
        s = scenes.scene_tiny("ka", 8, 8)
    elif case == "8lights":
        s = scenes.scene_random(21, 800, 96, 64, nlights=8)
    elif case == "nan":
        s = scenes.scene_random(22, 600, 96, 96, nlights=1)
        s.verts = s.verts.copy()
        s.verts[7, 1] = np.nan
        s.verts[100] = np.inf
    elif case == "allmiss":
        s = scenes.scene_tiny("one", 64, 64)
        s.verts = (s.verts + np.array([500.0, 0.0, 0.0], np.float32)).astype(np.float32)
    else:
        s = scenes.scene_tiny("one", 2048, 8)
    out = check_frame(ref, oracle, s)
    if case == "allmiss":
        assert (out["id"] >= 0).sum() == 0 and out["rgb"].max() == 0
    check_frame(ref, oracle, s, spherical=True, spot=True)


def test_config_a_reduced(ref, oracle):
    out = check_frame(ref, oracle, scenes.scene_a(256, 256))
    assert out["cam_P"] > 16370 and out["shadowed"].sum() > 1000


def test_dynamic_transform(ref, oracle):
    """copy_data_transform (transformation_kernel.cu:4-18) on a vertex sub-range, two different frames."""
    s = scenes.scene_random(7, 1200, 128, 128, nlights=1)
    s.verts = (s.verts * np.float32(0.5) + np.array([12.0, 5.0, 4.0], np.float32)).astype(np.float32)
    s.eye, s.look = (12.0, -12.0, 6.0), (14.0, 12.0, 4.0)
    s.dyn_offset, s.dyn_size = 300, 2100
    a = check_frame(ref, oracle, s, frame=0)
    b = check_frame(ref, oracle, s, frame=3)
    assert not np.array_equal(a["verts"], b["verts"])


def test_grid_is_sorted_and_complete(oracle):
    """Structural invariants of the cell lists (what the reference's sort+compact+histogram guarantee)."""
    s = scenes.scene_random(5, 3000, 128, 128)
    out = harness.run_frame(oracle, s, stages="grid")
    span, off, vals = out["cam_span"], out["cam_offset"], out["cam_values"]
    assert span.sum() == out["cam_P"] == vals.size
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(span)[:-1]]).astype(np.uint32))   # exclusive scan over ALL cells
    for c in np.flatnonzero(span > 1)[:200]:
        seg = vals[off[c]:off[c] + span[c]]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)          # ascending triangle ids, no duplicates


def test_grid_lists_are_conservative(oracle):
    """The grid result equals a brute-force nearest hit over ALL triangles wherever the brute-force hit exists
     - checked on a sample of pixels of a well-formed scene."""
    s = scenes.scene_a(128, 128)
    out = harness.run_frame(oracle, s, stages="trace")
    # brute force: one giant cell holding every triangle, traced with the same code
    T, C = s.num_tris, (s.width // 8) * (s.height // 8)
    grid = {"P": T * C, "values": np.tile(np.arange(T, dtype=np.uint32), C), "span": np.full(C, T, np.uint32),
            "offset": (np.arange(C, dtype=np.uint32) * T).astype(np.uint32)}
    oracle.configure(s.width, s.height, s.fovy, 0)
    oracle.camera_setup(s.eye, s.look, s.up, s.near, s.far)
    brute = oracle.trace(grid, np.ascontiguousarray(s.faces), np.ascontiguousarray(s.verts))
    assert np.array_equal(brute["id"], out["id"])
    assert np.array_equal(parity.bits(brute["t"]), parity.bits(out["t"]))


def test_loader_matches_obj_parser(ref, oracle, ug, tmp_path):
    """OBJ/MTL semantics of obj_parser/: first three corners of a face, v/vt/vn forms, negative (relative) indices,
    prefix match of usemtl against material names that keep their trailing newline, -1 without usemtl."""
    mtl = tmp_path / "m.mtl"
    scenes.write_material_file(str(mtl), [scenes.Material("floor2", (.1, .2, .3), (.4, .5, .6)), scenes.Material("floor", (.2, .2, .2), (.7, .8, .9)),
                                          scenes.Material("wall", (.3, .3, .3), (.15, .25, .35), reflect=0.5)])
    obj = tmp_path / "q.obj"
    obj.write_text(f"# quirks\nmtllib {mtl}\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1.25e0\nf 1 2 3\nusemtl floor\nf 1/1/1 2/2/2 3/3/3 4/4/4\n"
                   "usemtl wall\nf -1 -2 -3\nusemtl flo\nf 1//1 3//2 5//3\nusemtl nosuch\nf 2 3 4\ng grp\ns off\nf 5 4 3\n")
    want = ref.load_obj(str(obj))
    for got in (oracle.load_obj(str(obj)), ug.parse_obj(str(obj))):
        for a, b in zip(want, got):
            assert same(a, b)
    V, F, M = want
    assert F.tolist()[1] == [0, 1, 2] and F.tolist()[2] == [4, 3, 2]      # quad truncated; relative indices
    assert M.tolist() == [-1, 0, 2, 0, -1, -1]                            # 'floor' and 'flo' prefix-match 'floor2' (first in the list)
    wm = ref.load_materials(str(mtl))
    assert same(wm, oracle.load_materials(str(mtl))) and same(wm, ug.parse_materials(str(mtl)))
    assert wm.shape == (3, 6) and abs(wm[2, 5] - 0.35) < 1e-6


def test_generated_scene_roundtrip(ref, ug, tmp_path):
    """Scenes written by the generator come back identical through both loaders (triangle id = file order)."""
    s = scenes.scene_a(64, 64)
    obj, mtl = scenes.write_obj(s, str(tmp_path / "a.obj"))
    Vr, Fr, Mr = ref.load_obj(obj)
    Vu, Fu, Mu = ug.parse_obj(obj)
    assert same(Vr, Vu) and same(Fr, Fu) and same(Mr, Mu)
    assert same(Fr, s.faces) and same(Mr, s.face_mat)
    assert np.allclose(Vr, s.verts, rtol=0, atol=0)                      # %.9g round-trips float32 exactly
    assert same(ref.load_materials(mtl), harness.material_table(s))
