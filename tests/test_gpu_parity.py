"""Parity of the CUDA path (through the C-ABI, libugrt.so) against the CPU oracle.

Bars from BASELINE.json: cell/triangle index lists, hit-triangle ids and shadow bits BIT-EXACT (ties at equal t ->
lower triangle id); hit t to 1e-5 relative; RGB within 1/255 on >= 99.9 % of pixels.  The CUDA path follows the
reference's arithmetic operation for operation (no FMA contraction, IEEE div/sqrt, x86 float->int rule), so these
tests demand bit-exact t and RGB as well and report the looser bars only in the assertion messages."""
import numpy as np
import pytest

import parity
from oracle import harness
from uniformgrid_raytracing_b200 import scenes

pytestmark = pytest.mark.gpu

T_REL_TOL = 1e-5        # north_star tolerance for hit t
RGB_TOL, RGB_FRAC = 1, 0.999


def assert_frame(gpu, cpu, what):
    bad = parity.compare(gpu, cpu, verbose=False)
    # the stated bars, for the message
    rel = parity.rel_err_t(gpu["t"], cpu["t"])
    close = np.mean(np.abs(gpu["rgb"].astype(np.int16) - cpu["rgb"].astype(np.int16)).reshape(-1, 3).max(axis=1) <= RGB_TOL)
    assert rel <= T_REL_TOL and close >= RGB_FRAC, f"{what}: t rel err {rel}, rgb within 1/255 on {close:.5f}"
    assert not bad, f"{what}: stages not bit-exact: {bad}"


def cpu_frame(oracle, scene, frame=0, spherical=False, spot=False, slabs=1):
    try:
        out = harness.run_frame(oracle, scene, frame=frame, threads=0, spherical=spherical, spot=spot, slabs=slabs)
    finally:
        oracle.set_slabs(1)
    oracle.configure(scene.width, scene.height, scene.fovy, 0)
    oracle.camera_setup(scene.eye, scene.look, scene.up, scene.near, scene.far)
    out["sizelist"] = oracle.grid_count(np.ascontiguousarray(scene.faces, np.int32), np.ascontiguousarray(out["verts"], np.float32))[0]
    return out


@pytest.mark.parametrize("kind", ["one", "tie", "behind", "near", "outside", "ka"])
def test_micro_scenes(renderer, ug, oracle, kind):
    s = scenes.scene_tiny(kind)
    renderer.set_scene(s)
    assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, s), s.name)


@pytest.mark.parametrize("seed,ntris,w,h,nl", [(1, 2000, 128, 64, 2), (2, 500, 96, 160, 1), (3, 6000, 256, 256, 1), (4, 50, 64, 64, 0),
                                              (5, 20000, 320, 200, 3)])
def test_random_soups(renderer, ug, oracle, seed, ntris, w, h, nl):
    s = scenes.scene_random(seed, ntris, w, h, nlights=nl)
    renderer.set_scene(s)
    assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, s), s.name)


@pytest.mark.parametrize("seed", [77, 78, 79])
def test_hostile_geometry(renderer, ug, oracle, seed):
    """No clipping, no culling: triangles through the eye / light plane, behind the camera, degenerate,
    duplicated - garbage boxes included, every stage must still match the oracle bit for bit."""
    s = scenes.scene_hostile(seed)
    renderer.set_scene(s)
    assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, s), s.name)


def test_config_a(renderer, ug, oracle):
    """BASELINE config A: 16 370-triangle icosphere on a plane, 512x512, primary + shadow, 1 light - every stage."""
    s = scenes.scene_a()
    renderer.set_scene(s)
    cpu = cpu_frame(oracle, s)
    assert_frame(parity.gpu_frame(renderer, s), cpu, "config A (graph)")
    assert_frame(parity.gpu_frame(renderer, s, flags=ug.FLAG_FULL_BUFFERS | ug.FLAG_NO_GRAPH), cpu, "config A (direct launches)")
    assert_frame(parity.gpu_frame(renderer, s, stepwise=True), cpu, "config A (per_frame_funcs call sequence)")


@pytest.mark.parametrize("case", ["ka", "soup2l", "soup", "hostile", "a", "c"])
def test_spherical_light_grid(renderer, ug, oracle, case):
    """a14/a15: the light path of the committed display() (main.cu:186-192) - spherical shadow-ray map
    (misc_kernel.cu:255-296) and spherical light grid (grid_kernel.cu:395-479, 481-854; frustum_grid.h:368-532).  Needs
    acosf to agree with the host's libm bit for bit: light cells, light-grid lists and shadow bits are all compared."""
    s = {"ka": lambda: scenes.scene_tiny("ka"), "soup2l": lambda: scenes.scene_random(1, 2000, 128, 64, nlights=2),
         "soup": lambda: scenes.scene_random(5, 20000, 320, 200, nlights=3), "hostile": lambda: scenes.scene_hostile(77),
         "a": lambda: scenes.scene_a(), "c": lambda: scenes.scene_c(480, 272)}[case]()
    renderer.set_scene(s)
    fl = ug.FLAG_FULL_BUFFERS | ug.FLAG_SPHERICAL_LIGHTS
    cpu = cpu_frame(oracle, s, spherical=True)
    assert_frame(parity.gpu_frame(renderer, s, flags=fl), cpu, f"{s.name} spherical (graph)")
    if case in ("ka", "a"):
        assert_frame(parity.gpu_frame(renderer, s, flags=fl | ug.FLAG_NO_GRAPH), cpu, f"{s.name} spherical (direct launches)")
        assert_frame(parity.gpu_frame(renderer, s, flags=fl, stepwise=True), cpu, f"{s.name} spherical (per_frame_funcs call sequence)")
        # and back: the same context renders the perspective variant again
        assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, s), f"{s.name} perspective after spherical")


@pytest.mark.parametrize("case", ["ka", "soup2l", "hostile", "a", "c"])
def test_spot_shade(renderer, ug, oracle, case):
    """Shader::spotlight_shade -> spot_shade (shader_kernel.cu:275-345; display() from its third frame on, main.cu:205-219),
    behind the spherical light path and behind the perspective one: RGB bit-exact, cone border pixels included."""
    s = {"ka": lambda: scenes.scene_tiny("ka"), "soup2l": lambda: scenes.scene_random(1, 2000, 128, 64, nlights=2),
         "hostile": lambda: scenes.scene_hostile(77), "a": lambda: scenes.scene_a(), "c": lambda: scenes.scene_c(480, 272)}[case]()
    renderer.set_scene(s)
    for sph in (True, False):
        fl = ug.FLAG_FULL_BUFFERS | ug.FLAG_SPOT_SHADE | (ug.FLAG_SPHERICAL_LIGHTS if sph else 0)
        cpu = cpu_frame(oracle, s, spherical=sph, spot=True)
        assert_frame(parity.gpu_frame(renderer, s, flags=fl), cpu, f"{s.name} spot shade, spherical={sph}")
        if case in ("ka", "a"):
            assert_frame(parity.gpu_frame(renderer, s, flags=fl, stepwise=True), cpu, f"{s.name} spot shade, call sequence, spherical={sph}")
    if case == "a":
        plain = parity.gpu_frame(renderer, s)
        assert not np.array_equal(plain["rgb"], cpu["rgb"])     # the flag is part of the graph key: the plain shader is back


@pytest.mark.parametrize("case,slabs,sph", [("ka", 4, False), ("soup2l", 8, False), ("hostile", 4, False), ("a", 4, False), ("a", 3, True), ("c", 2, True),
                                            ("b", 4, False)])
def test_depth_slabs(renderer, ug, oracle, case, slabs, sph):
    """NUM_SLABS > 1: slab-keyed cell lists of every grid (SlabKernel, grid_kernel.cu:334-352, :322), primary rays
    through rckernel_alpha's per-tile slab loop (trace_kernel.cu:132-229, its reset-by-isWithin quirk included), shadow rays
    over the union of a light cell's slabs (light_kernel.cu:105-115) - all stages bit-exact against the oracle."""
    s = {"ka": lambda: scenes.scene_tiny("ka"), "soup2l": lambda: scenes.scene_random(1, 2000, 128, 64, nlights=2),
         "hostile": lambda: scenes.scene_hostile(77), "a": lambda: scenes.scene_a(), "c": lambda: scenes.scene_c(480, 272),
         "b": lambda: scenes.scene_b()}[case]()
    renderer.set_scene(s)
    fl = ug.FLAG_FULL_BUFFERS | (ug.FLAG_SPHERICAL_LIGHTS if sph else 0)
    cpu = cpu_frame(oracle, s, spherical=sph, slabs=slabs)
    cpu.pop("sizelist", None)
    C = (s.width // 8) * (s.height // 8)
    gpu = parity.gpu_frame(renderer, s, flags=fl, slabs=slabs)
    assert gpu["cam_span"].size == C * slabs
    assert_frame(gpu, cpu, f"{s.name} {slabs} slabs (graph)")
    if case in ("ka", "a"):
        assert_frame(parity.gpu_frame(renderer, s, flags=fl | ug.FLAG_NO_GRAPH, slabs=slabs), cpu, f"{s.name} {slabs} slabs (direct launches)")
        assert_frame(parity.gpu_frame(renderer, s, flags=fl, stepwise=True, slabs=slabs), cpu, f"{s.name} {slabs} slabs (per_frame_funcs call sequence)")
        assert_frame(parity.gpu_frame(renderer, s, flags=fl), cpu_frame(oracle, s, spherical=sph), f"{s.name} back to one slab")


def test_spherical_build_variants(renderer, ug, oracle, monkeypatch):
    """Both build variants (one cooperative kernel / kernel chain) of the spherical light grid give the same lists."""
    s = scenes.scene_a(256, 256)
    cpu = cpu_frame(oracle, s, spherical=True)
    for env in ({"UG_NO_FUSED": "1"}, {"UG_FUSED_GRID": "1", "UG_FUSED_REGROUP": "1"}, {"UG_NO_SWEEP": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = ug.Renderer(0)
        try:
            r.set_scene(s)
            assert_frame(parity.gpu_frame(r, s, flags=ug.FLAG_FULL_BUFFERS | ug.FLAG_SPHERICAL_LIGHTS), cpu, f"spherical {env}")
        finally:
            r.close()
        for k in env:
            monkeypatch.delenv(k)


def test_config_b(renderer, ug, oracle):
    """BASELINE config B geometry (80 000 triangles, 1024x1024, 1 light): every stage."""
    s = scenes.scene_b()
    renderer.set_scene(s)
    assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, s), "config B")


def test_config_c_dynamic_two_lights(renderer, ug, oracle):
    """BASELINE config C at reduced resolution: per-frame vertex transform + rebuild of the camera grid and two light grids."""
    s = scenes.scene_c(480, 272)
    renderer.set_scene(s)
    for frame in (0, 2):
        assert_frame(parity.gpu_frame(renderer, s, frame=frame), cpu_frame(oracle, s, frame=frame), f"config C frame {frame}")


def test_idempotent_and_graph_replay(renderer, ug):
    """Rendering the same frame again (graph replay, self-resetting counters) gives the same bytes."""
    s = scenes.scene_random(9, 5000, 256, 128, nlights=2)
    renderer.set_scene(s)
    a = parity.gpu_frame(renderer, s, flags=0)
    for _ in range(3):
        b = parity.gpu_frame(renderer, s, flags=0)
        assert not parity.compare(b, a, verbose=False)


def test_strips_equal_full_frame(renderer, ug, oracle):
    """Image-strip sharding: every strip's lists, hits, shadow bits and pixels equal the corresponding
    rows of the single-context frame; the strips concatenate to the full image byte for byte."""
    s = scenes.scene_a(256, 264)        # 33 cell rows: uneven strips
    renderer.set_scene(s)
    full = parity.gpu_frame(renderer, s, flags=0)
    cpu = cpu_frame(oracle, s)
    assert not parity.compare(full, cpu, keys=["t", "id", "shadowed", "rgb"], verbose=False)
    nbx, nby = s.width // 8, s.height // 8
    for world in (2, 3, 8):
        rows_done, img = 0, []
        for rank in range(world):
            part = parity.gpu_frame(renderer, s, flags=0, strip=(rank, world))
            lo, hi = nby * rank // world, nby * (rank + 1) // world
            npix = s.width * (hi - lo) * 8
            sl = slice(s.width * lo * 8, s.width * hi * 8)
            assert part["t"].size == npix
            for k in ("t", "id", "shadowed"):
                assert np.array_equal(parity.bits(part[k]), parity.bits(full[k][sl])), (world, rank, k)
            assert np.array_equal(part["rgb"], full["rgb"][3 * sl.start:3 * sl.stop])
            # the strip's cell lists are the global lists of the owned cells
            span_full = full["cam_span"].reshape(nbx, nby)[:, lo:hi]
            assert np.array_equal(part["cam_span"].reshape(nbx, hi - lo), span_full)
            off_f, off_p = full["cam_offset"].reshape(nbx, nby), part["cam_offset"].reshape(nbx, hi - lo)
            for gx in range(0, nbx, 5):
                for gy in range(lo, hi):
                    n = int(span_full[gx, gy - lo])
                    a = full["cam_values"][off_f[gx, gy]:off_f[gx, gy] + n]
                    b = part["cam_values"][off_p[gx, gy - lo]:off_p[gx, gy - lo] + n]
                    assert np.array_equal(a, b)
            img.append(part["rgb"])
            rows_done += hi - lo
        assert rows_done == nby and np.array_equal(np.concatenate(img), full["rgb"])
    renderer.render(renderer.params(s, strip=(0, 1)))


def test_host_path_and_ppm(renderer, ug, oracle, tmp_path):
    """ug_render_frame_host (pinned H2D constants, D2H image) and the P3 writer (per_app_funcs.h:39-66)."""
    s = scenes.scene_random(12, 3000, 128, 96, nlights=1)
    renderer.set_scene(s)
    cpu = cpu_frame(oracle, s)
    out = np.zeros(3 * s.num_pixels, np.uint8)
    renderer.render_host(renderer.params(s), out)
    assert np.array_equal(out, cpu["rgb"])
    path = str(tmp_path / "f.ppm")
    renderer.write_ppm(path)
    tok = open(path).read().split()
    assert tok[:4] == ["P3", str(s.width), str(s.height), "255"]
    assert np.array_equal(np.asarray(tok[4:], dtype=np.int64).astype(np.uint8), cpu["rgb"])


def test_model_load_through_obj(renderer, ug, oracle, tmp_path):
    """init_model path: OBJ + material file -> Model -> frame; same result as uploading the arrays."""
    s = scenes.scene_a(128, 128)
    obj, mtl = scenes.write_obj(s, str(tmp_path / "a.obj"))
    renderer.load(obj, mtl)
    renderer.scene = s
    assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, s), "config A via OBJ loader")


def test_keyframed_loader_tmp_model(renderer, ug, oracle, tmp_path):
    """Model::tmp_model (scene.h:70-120): f_<i>.obj keyframes reloaded into the same topology every frame; each frame
    must equal the oracle run on that keyframe's vertices (as parsed: the OBJ text is %.6f)."""
    s = scenes.scene_random(21, 300, 128, 96, 2)
    obj, mat = scenes.write_obj(s, str(tmp_path / "f_0.obj"))
    rng = np.random.default_rng(5)
    for i in (1, 2, 3):
        k = scenes.Scene(**{**s.__dict__, "verts": (s.verts + rng.normal(scale=0.15, size=s.verts.shape)).astype(np.float32)})
        scenes.write_obj(k, str(tmp_path / f"f_{i}.obj"))
    renderer.load(obj, mat)
    renderer.scene = s
    for i in (0, 1, 2, 3, 1):
        renderer.tmp_model(str(tmp_path), i)
        V, F, M = ug.parse_obj(str(tmp_path / f"f_{i}.obj"))
        k = scenes.Scene(**{**s.__dict__, "verts": V})
        b = renderer.bounds()
        assert np.array_equal(b[:3], V.min(0)) and np.array_equal(b[3:], V.max(0))
        assert_frame(parity.gpu_frame(renderer, s), cpu_frame(oracle, k), f"keyframe {i}")
    with pytest.raises(RuntimeError):
        renderer.tmp_model(str(tmp_path), 9)          # no such keyframe: loud
    with pytest.raises(RuntimeError):
        renderer.update_vertices(np.zeros((len(s.verts) + 1, 3), np.float32))   # more vertices than the model holds


def test_binary_ppm(renderer, ug, tmp_path):
    s = scenes.scene_random(22, 200, 128, 96, 1)
    renderer.set_scene(s)
    renderer.render(renderer.params(s))
    rgb = renderer.read(ug.BUF_RGB).reshape(s.height, s.width, 3)
    path = tmp_path / "o.ppm"
    renderer.write_ppm_binary(str(path))
    raw = path.read_bytes()
    head = b"P6\n%d %d\n255\n" % (s.width, s.height)
    assert raw.startswith(head) and len(raw) == len(head) + rgb.size
    assert np.array_equal(np.frombuffer(raw[len(head):], np.uint8).reshape(s.height, s.width, 3), rgb[::-1])   # top row first


def test_frames_in_flight(ug, oracle):
    """ug_render_frame_host_async / ug_wait_frame: the read-back of frame i runs on its own stream while frame i+1 renders (the
    first kernel that rewrites the image waits for it inside the graph).  Six frames of a dynamic scene with two frames in
    flight and two pinned buffers: every image must be the frame the synchronous call returns (and the oracle's)."""
    import torch
    s = scenes.scene_c(240, 136)
    r = ug.Renderer()
    try:
        r.set_scene(s)
        want = []
        host = torch.empty(3 * s.num_pixels, dtype=torch.uint8).pin_memory()
        for f in range(6):
            r.render_host(r.params(s, frame=f), host.numpy())
            want.append(host.numpy().copy())
        assert np.array_equal(want[3], harness.run_frame(oracle, s, frame=3, threads=0)["rgb"])
        bufs = [torch.empty(3 * s.num_pixels, dtype=torch.uint8).pin_memory() for _ in range(2)]
        got = {}
        for f in range(6):
            r.render_host_async(r.params(s, frame=f), bufs[f & 1].numpy())
            if f >= 1:
                r.wait_frame(1)
                got[f - 1] = bufs[(f - 1) & 1].numpy().copy()
        r.wait_frame(0)
        got[5] = bufs[1].numpy().copy()
        for f in range(6):
            assert np.array_equal(got[f], want[f]), f"frame {f}"
        with pytest.raises(RuntimeError):
            r.render_host_async(r.params(s, frame=0), np.zeros(3 * s.num_pixels, np.uint8))   # pageable destination: refused
        r.render_host(r.params(s, frame=2), host.numpy())                                    # the synchronous call still works afterwards
        assert np.array_equal(host.numpy(), want[2])
    finally:
        r.close()


def test_png(renderer, ug, tmp_path):
    """ug_write_png (what display() gets from writePPM + `convert -flip`): decoded here with zlib + the PNG chunk
    layout by hand - signature, IHDR, one IDAT of stored deflate blocks, IEND, every CRC - and compared with the frame.  The
    image is wider than one 65535-byte deflate block per few rows, so several blocks are written."""
    import struct
    import zlib
    s = scenes.scene_random(23, 200, 320, 136, 1)
    renderer.set_scene(s)
    renderer.render(renderer.params(s))
    rgb = renderer.read(ug.BUF_RGB).reshape(s.height, s.width, 3)
    path = tmp_path / "o.png"
    renderer.write_png(str(path))
    raw = path.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        data = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + data) & 0xffffffff, typ
        chunks.append((typ, data))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    assert struct.unpack(">IIBBBBB", chunks[0][1]) == (s.width, s.height, 8, 2, 0, 0, 0)
    lines = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(s.height, 1 + 3 * s.width)
    assert np.all(lines[:, 0] == 0)                                                            # filter type 0 on every scanline
    assert np.array_equal(lines[:, 1:].reshape(s.height, s.width, 3), rgb[::-1])               # top row first


def test_pair_arena_overflow_regrows(ug, oracle):
    """P is data dependent: a scene whose pair count exceeds the initial arena must be re-run with a larger one."""
    # 24 huge triangles at 1024x1024: each covers most of the 16 384 cells -> P far above 16*T + C + 2^20
    rng = np.random.default_rng(3)
    T = 96
    V = np.zeros((T, 3, 3), np.float32)
    for k in range(T):
        y = 20.0 + 0.25 * k
        V[k] = [[-60 - rng.random(), y, -60], [60 + rng.random(), y, -60], [0, y, 70 + rng.random()]]
    s = scenes.scene_random(1, 10, 1024, 1024, nlights=1)
    s.verts, s.faces, s.face_mat = V.reshape(-1, 3), np.arange(3 * T, dtype=np.int32).reshape(-1, 3), np.zeros(T, np.int32)
    r = ug.Renderer(0)
    try:
        r.set_scene(s)
        out = np.zeros(3 * s.num_pixels, np.uint8)
        r.render_host(r.params(s), out)
        st = r.stats()
        assert st.cam_pairs > 16 * T + 16384 + (1 << 20) and st.overflowed == 1
        cpu = cpu_frame(oracle, s)
        assert st.cam_pairs == cpu["cam_P"] and np.array_equal(out, cpu["rgb"])
        assert np.array_equal(r.read(ug.BUF_ID), cpu["id"])
    finally:
        r.close()


def test_async_frame_overflow_is_reported_and_repaired(ug, oracle):
    """ug_render_frame cannot re-run itself: the overflow shows in the stats, ug_get_stats grows the arena, the second
    attempt is complete."""
    T = 96
    V = np.zeros((T, 3, 3), np.float32)
    for k in range(T):
        y = 20.0 + 0.25 * k
        V[k] = [[-60, y, -60], [60, y, -60], [0, y, 70]]
    s = scenes.scene_random(1, 10, 1024, 1024, nlights=1)
    s.verts, s.faces, s.face_mat = V.reshape(-1, 3), np.arange(3 * T, dtype=np.int32).reshape(-1, 3), np.zeros(T, np.int32)
    r = ug.Renderer(0)
    try:
        r.set_scene(s)
        r.render(r.params(s))
        assert r.stats().overflowed == 1
        r.render(r.params(s))
        st = r.stats()
        cpu = cpu_frame(oracle, s)
        assert st.overflowed == 0 and st.cam_pairs == cpu["cam_P"]
        assert np.array_equal(r.read(ug.BUF_RGB), cpu["rgb"]) and np.array_equal(r.read(ug.BUF_ID), cpu["id"])
    finally:
        r.close()


@pytest.mark.parametrize("case", ["8x8", "8lights", "nan", "allmiss", "1tri_wide"])
def test_edge_sizes(renderer, ug, oracle, case):
    """Smallest screen (one cell), the maximum number of lights, a NaN vertex (NaN-propagating min/max, x86 float->int rule),
    a frame without a single hit, one triangle on a very wide strip-shaped screen."""
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
    renderer.set_scene(s)
    cpu = cpu_frame(oracle, s)
    assert_frame(parity.gpu_frame(renderer, s), cpu, s.name + " " + case)
    if case == "allmiss":
        assert (cpu["id"] >= 0).sum() == 0 and cpu["rgb"].max() == 0


def reflective(scene, ks, bounces):
    for m, k in zip(scene.materials, ks):
        m.reflect = k
    scene.bounces = bounces
    return scene


@pytest.mark.parametrize("make,bounces", [(lambda: reflective(scenes.scene_random(31, 1500, 160, 96, nlights=1), (0.5, 0.25), 2), 2),
                                          (lambda: reflective(scenes.scene_random(32, 4000, 128, 128, nlights=2), (0.75, 0.0), 1), 1),
                                          (lambda: reflective(scenes.scene_a(128, 128), (0.5, 0.3), 2), 2),
                                          (lambda: reflective(scenes.scene_tiny("ka", 96, 96), (0.6, 0.6), 3), 3)])
def test_reflection_row_r(renderer, ug, oracle, make, bounces):
    """Row R (authored spec, parity unpinned: the reference has no reflection code): uniform grid + 3-D DDA + scan
    compaction on the GPU against the oracle's brute-force nearest hit over ALL triangles, bit-exact RGB."""
    s = make()
    renderer.set_scene(s)
    cpu = harness.run_frame(oracle, s, threads=0, reflect=True)
    gpu = parity.gpu_frame(renderer, s, flags=0, bounces=bounces)
    assert np.count_nonzero(cpu["rgb_reflect"] != cpu["rgb"]) > 50                    # the bounce changes pixels
    assert not parity.compare(gpu, cpu, keys=["t", "id", "shadowed"], verbose=False)  # the primary pass is untouched
    diff = np.flatnonzero(gpu["rgb"] != cpu["rgb_reflect"])
    assert diff.size == 0, f"{diff.size} bytes differ, first at {diff[:8]}"
    again = parity.gpu_frame(renderer, s, flags=0, bounces=bounces)                   # replay
    assert np.array_equal(again["rgb"], gpu["rgb"])
    plain = parity.gpu_frame(renderer, s, flags=0, bounces=0)                         # bounces off -> the reference frame
    assert np.array_equal(plain["rgb"], cpu["rgb"])


@pytest.mark.parametrize("env", [{"UG_NO_FUSED": "1"}, {"UG_FUSED_GRID": "1", "UG_FUSED_REGROUP": "1", "UG_FUSED_UGRID": "1"}, {"UG_NO_SWEEP": "1"}, {}, {"UG_SHADOW_DIRECT": "1"},
                                 {"UG_TRACE_ITEM": "64"}, {"UG_TRACE_ITEM": "128"}, {"UG_NO_SPLIT_MAP": "1"}, {"UG_DIR_RECOMPUTE": "1"}, {"UG_DIR_RECOMPUTE": "2"},
                                 {"UG_DIR_RECOMPUTE": "3"}, {"UG_NO_HOIST_NORMALS": "1"}, {"UG_LATE_PUBLISH": "1"}, {"UG_NO_PDL": "1"}])
def test_build_variants(ug, oracle, env, monkeypatch):
    """Every view grid can be built three ways - the three one-sweep launches of k_gridsweep.cu (default), the 8-bit LSD
    kernel chain, the one-launch cooperative kernel - and regroup / uniform grid two ways; which one a frame uses is a
    policy (context.cu), so each is forced here: two lights + two reflection bounces.  UG_SHADOW_DIRECT selects the shadow
    pass that leaves the rays in screen order (k_shadowdirect.cu, TMA-staged records) instead of regrouping them;
    UG_TRACE_ITEM the two work-item sizes of k_trace (chosen per frame from the last pair count otherwise); UG_NO_SPLIT_MAP
    keeps the ray -> light cell map on the light grid's branch of the frame graph; UG_DIR_RECOMPUTE says which kernels of
    a frame evaluate the primary ray direction again from the lattice instead of reading d_primary_ray_direction (bit 0
    k_shade, bit 1 the ray -> light cell map; measured slower than the stored buffer, so the default is 0); UG_NO_HOIST_NORMALS makes k_shade evaluate
    the shading normal per pixel instead of reading the per-triangle one k_shade_normals made for the frame; UG_LATE_PUBLISH
    keeps the counters' publish kernel behind the shading instead of beside it; UG_NO_PDL launches the frame's kernel chain the plain
    way instead of with programmatic dependent launches (k_common.cuh; read per frame)."""
    for k in ("UG_NO_FUSED", "UG_FUSED_GRID", "UG_FUSED_REGROUP", "UG_FUSED_UGRID", "UG_NO_SWEEP", "UG_SHADOW_DIRECT", "UG_TRACE_ITEM", "UG_NO_SPLIT_MAP",
              "UG_DIR_RECOMPUTE", "UG_NO_HOIST_NORMALS", "UG_LATE_PUBLISH", "UG_NO_PDL", "UG_PDL_MASK"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    r = ug.Renderer()      # the switches are read by ug_create
    try:
        s = reflective(scenes.scene_random(31, 900, 192, 128, 2), (0.5, 0.3), 2)
        r.set_scene(s)
        cpu = harness.run_frame(oracle, s, threads=0, reflect=True)
        gpu = parity.gpu_frame(r, s, bounces=2)
        for k in ("cam_values", "cam_span", "cam_offset", "t", "id", "shadowed", "light0_values", "light1_values", "light0_cell", "light1_cell"):
            assert np.array_equal(gpu[k], cpu[k]), k
        assert np.array_equal(gpu["rgb"], cpu["rgb_reflect"])
        assert np.array_equal(parity.gpu_frame(r, s, bounces=0)["rgb"], cpu["rgb"])
    finally:
        r.close()


def test_full_size_properties(renderer, ug):
    """Size-independent properties at a BASELINE full size (config C, 174 000 triangles, 1920x1080, 2 lights, dynamic):
    span sums to P, offset is the exclusive scan over ALL cells, lists ascend inside a cell, every hit id is in the
    pixel's own cell list, t > 0 exactly where id >= 0, shadow bits only 0/1, unshadowed pixels are not darker."""
    s = scenes.scene_c()
    renderer.set_scene(s)
    g = parity.gpu_frame(renderer, s, frame=1, flags=0)
    st = renderer.stats()
    span, off, vals = g["cam_span"], g["cam_offset"], g["cam_values"]
    assert span.sum() == st.cam_pairs == vals.size
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(span, dtype=np.uint64)[:-1]]).astype(np.uint32))
    starts = np.zeros(vals.size, bool)
    starts[off[span > 0]] = True
    assert np.all((np.diff(vals.astype(np.int64)) > 0) | starts[1:])            # ascending ids inside every cell
    hit = g["id"] >= 0
    assert np.array_equal(hit, g["t"] > 0) and np.all(g["id"][~hit] == -2) and np.all(g["t"][~hit] == -1.0)
    assert set(np.unique(g["shadowed"]).tolist()) <= {0, 1}
    nby = s.height // 8
    pix = np.flatnonzero(hit)[:: max(1, hit.sum() // 4000)]
    for p in pix:
        py, px = divmod(int(p), s.width)
        c = (px // 8) * nby + py // 8
        assert g["id"][p] in vals[off[c]:off[c] + span[c]]
    for li in range(2):
        ls, lo_ = g[f"light{li}_span"], g[f"light{li}_offset"]
        assert ls.sum() == st.light_pairs[li] and lo_[0] == 0
        assert g[f"light{li}_cell"].max() <= (s.width // 8) * nby
    renderer.render(renderer.params(s, frame=1, flags=0))
    assert np.array_equal(renderer.read(ug.BUF_RGB), g["rgb"])                   # same frame again -> same bytes


# ------------------------------------------------------------------------------------------------ BASELINE configs at their stated sizes
def test_config_c_full_size(renderer, ug, oracle):
    """BASELINE config C as stated: 174 000 triangles, 1920x1080, 2 lights, per-frame vertex transform - every stage of
    frames 0 and 3 bit-exact against the oracle (camera grid + two light grids rebuilt per frame)."""
    s = scenes.scene_c()
    renderer.set_scene(s)
    for frame in (0, 3):
        assert_frame(parity.gpu_frame(renderer, s, frame=frame), cpu_frame(oracle, s, frame=frame), f"config C 1920x1080 frame {frame}")


def test_config_d_full_size(renderer, ug, oracle):
    """BASELINE config D as stated: 282 000 triangles, 4096x4096, 1 light - every stage bit-exact against the oracle,
    both as one frame and (pixels only) as the 8 image strips an 8-GPU run cuts it into."""
    s = scenes.scene_d()
    renderer.set_scene(s)
    cpu = cpu_frame(oracle, s)
    assert_frame(parity.gpu_frame(renderer, s), cpu, "config D 4096x4096")
    nby = s.height // 8
    for world in (8,):
        img = []
        for rank in range(world):
            renderer.render(renderer.params(s, strip=(rank, world)))
            lo, hi = nby * rank // world, nby * (rank + 1) // world
            sl = slice(s.width * lo * 8, s.width * hi * 8)
            assert np.array_equal(renderer.read(ug.BUF_ID), cpu["id"][sl]), f"strip {rank}/{world}: hit ids"
            assert np.array_equal(renderer.read(ug.BUF_SHADOWED), cpu["shadowed"][sl]), f"strip {rank}/{world}: shadow bits"
            img.append(renderer.read(ug.BUF_RGB))
        assert np.array_equal(np.concatenate(img), cpu["rgb"])
    renderer.render(renderer.params(s, strip=(0, 1)))


def test_config_b_reflection_full_size(renderer, ug, oracle):
    """BASELINE config B as stated, reflection bounce included: 80 000 triangles, 1024x1024, 1 light, 1 bounce.  Row R
    is an authored spec (parity unpinned); the oracle's nearest hit is a brute force over ALL triangles for each of the
    ~234 K reflection rays, the GPU walks its uniform grid - RGB must agree byte for byte."""
    s = scenes.scene_b()
    assert s.bounces == 1 and any(m.reflect > 0 for m in s.materials)
    renderer.set_scene(s)
    cpu = harness.run_frame(oracle, s, threads=0, reflect=True)
    gpu = parity.gpu_frame(renderer, s, flags=0, bounces=s.bounces)
    assert renderer.stats().reflect_rays > 100000
    assert not parity.compare(gpu, cpu, keys=["cam_values", "cam_span", "t", "id", "shadowed", "light0_values", "light0_cell"], verbose=False)
    diff = np.flatnonzero(gpu["rgb"] != cpu["rgb_reflect"])
    assert diff.size == 0, f"{diff.size} bytes differ, first at {diff[:8]}"
    assert np.count_nonzero(cpu["rgb_reflect"] != cpu["rgb"]) > 10000


def test_config_e_subsample(ug, oracle):
    """BASELINE config E as stated: 1 000 000-triangle soup, 8192x8192, 4 lights.  The five grids (camera + 4 lights,
    143 M + 4 x 20 M pairs) are compared in full; the oracle then traces primary and shadow rays for 1/8 of the cells
    (every pixel of a pseudo-random eighth of the 8x8 tiles) and those 8.4 M pixels must agree bit for bit in hit id,
    t, light cell, shadow bit and RGB.  The two reflection bounces run on the GPU only (the oracle's brute force over
    1 M triangles per ray is out of reach): ray counts must be plausible and the frame must replay to the same bytes."""
    s = scenes.scene_e()
    r = ug.Renderer(0)
    try:
        r.set_scene(s)
        r.render(r.params(s, flags=0))
        if r.stats().overflowed:            # P is data dependent: the stats call grew the arena, render again
            r.render(r.params(s, flags=0))
        st = r.stats()
        assert st.overflowed == 0
        W, H, N = s.width, s.height, s.num_pixels
        nbx, nby = W // 8, H // 8
        faces = np.ascontiguousarray(s.faces, np.int32)
        verts = np.ascontiguousarray(s.verts, np.float32)
        oracle.configure(W, H, s.fovy, 0)
        oracle.set_threads(0)
        oracle.set_slabs(1)
        lights = np.asarray([l.pos for l in s.lights], np.float32).reshape(-1, 3)
        oracle.set_lights(lights[:2])
        oracle.camera_setup(s.eye, s.look, s.up, s.near, s.far)
        grid = oracle.build_grid(faces, verts)
        assert grid["P"] == st.cam_pairs
        assert np.array_equal(r.read(ug.BUF_CAM_SPAN), grid["span"]) and np.array_equal(r.read(ug.BUF_CAM_OFFSET), grid["offset"])
        assert np.array_equal(r.read(ug.BUF_CAM_VALUES), grid["values"])
        # 1/8 of the cells
        gx, gy = np.divmod(np.arange(nbx * nby, dtype=np.int64), nby)
        keep_cell = ((gx * 7 + gy * 13) % 8) == 0
        sub = dict(grid)
        sub["span"] = np.where(keep_cell, grid["span"], 0).astype(np.uint32)
        prim = oracle.trace(sub, faces, verts)
        py, px = np.divmod(np.arange(N, dtype=np.int64), W)
        keep_pix = keep_cell[(px // 8) * nby + py // 8]
        del py, px, gx, gy
        g_t, g_id = r.read(ug.BUF_T), r.read(ug.BUF_ID)
        assert np.array_equal(g_id[keep_pix], prim["id"][keep_pix])
        assert np.array_equal(parity.bits(g_t[keep_pix]), parity.bits(prim["t"][keep_pix]))
        assert (prim["id"][keep_pix] >= 0).sum() > 1000000
        shadowed = prim["shadowed"]
        cam_pos = np.asarray(s.eye, np.float32)
        for li, light in enumerate(s.lights):
            oracle.camera_setup(light.pos, light.look, light.up, s.near, s.far, square=True)
            m = oracle.map_rays(prim["t"], prim["dir"], cam_pos, s.near)
            g_cell = r.read(ug.BUF_LIGHT_CELL, li)
            assert np.array_equal(g_cell[keep_pix], m[N:][keep_pix]), f"light {li}: ray -> light cell"
            m[N:][~keep_pix] = nbx * nby                                   # outside the light frustum: no shadow ray
            lgrid = oracle.build_grid(faces, verts)
            assert lgrid["P"] == st.light_pairs[li]
            assert np.array_equal(r.read(ug.BUF_LIGHT_SPAN, li), lgrid["span"])
            assert np.array_equal(r.read(ug.BUF_LIGHT_VALUES, li), lgrid["values"]), f"light {li}: triangle list"
            oracle.shadow(lgrid, faces, verts, prim["t"], prim["dir"], shadowed, m, cam_pos)
            del lgrid, m, g_cell
        g_sh = r.read(ug.BUF_SHADOWED)
        assert np.array_equal(g_sh[keep_pix], shadowed[keep_pix])
        assert shadowed[keep_pix].sum() > 10000
        ids = prim["id"].copy()
        img = oracle.shade(prim["normal"], prim["t"], prim["dir"], ids, cam_pos, s.face_mat, harness.material_table(s))
        oracle.add_shadows(img, shadowed)
        g_rgb = r.read(ug.BUF_RGB).reshape(-1, 3)
        assert np.array_equal(g_rgb[keep_pix], img.reshape(-1, 3)[keep_pix])
        del prim, img, ids, g_t, g_id, g_sh
        # the two reflection bounces (row R) at E's size: GPU only
        r.render(r.params(s, flags=0, bounces=s.bounces))
        if r.stats().overflowed:
            r.render(r.params(s, flags=0, bounces=s.bounces))
        st2 = r.stats()
        assert st2.overflowed == 0 and st2.reflect_rays > 1000000
        a = r.read(ug.BUF_RGB)
        assert np.count_nonzero(a.reshape(-1, 3) != g_rgb) > 100000        # the bounces change pixels
        r.render(r.params(s, flags=0, bounces=s.bounces))
        assert np.array_equal(r.read(ug.BUF_RGB), a)
    finally:
        r.close()


def _flat_far_scene():
    """Every vertex on the axis-aligned plane y = 48: the scene's bounding box has no extent along y, and the pad that
    would give it one is lost to rounding that far from the origin."""
    s = scenes.scene_random(41, 1200, 128, 96, nlights=1)
    V = s.verts.copy()
    V[:, 1] = 48.0
    V[:, 0] *= 3.0
    V[:, 2] *= 3.0
    s.verts = V
    return reflective(s, (0.5, 0.5), 2)


@pytest.mark.parametrize("make", [_flat_far_scene])
def test_reflection_degenerate_bounding_box(ug, oracle, make):
    """The uniform grid of row R must stay inside its cell arena when the scene's bounding box is flat:
    the resolution is clamped, the frame still equals the oracle's brute force, and the context renders an ordinary
    scene correctly afterwards."""
    s = make()
    r = ug.Renderer(0)
    try:
        r.set_scene(s)
        cpu = harness.run_frame(oracle, s, threads=0, reflect=True)
        gpu = parity.gpu_frame(r, s, flags=0, bounces=2)
        assert not parity.compare(gpu, cpu, keys=["t", "id", "shadowed"], verbose=False)
        assert np.array_equal(gpu["rgb"], cpu["rgb_reflect"])
        s2 = reflective(scenes.scene_random(31, 1500, 160, 96, nlights=1), (0.5, 0.25), 2)
        r.set_scene(s2)
        cpu2 = harness.run_frame(oracle, s2, threads=0, reflect=True)
        assert np.array_equal(parity.gpu_frame(r, s2, flags=0, bounces=2)["rgb"], cpu2["rgb_reflect"])
    finally:
        r.close()


def test_unequal_strips(renderer, ug, oracle):
    """ug_set_strip_bounds: strips of unequal height (what a host does to balance the ranks of an N-GPU frame) - every strip's
    hits, shadow bits and pixels are the corresponding rows of the oracle's frame, and the equal split is back afterwards."""
    s = scenes.scene_a(256, 264)        # 33 cell rows
    renderer.set_scene(s)
    cpu = cpu_frame(oracle, s)
    bounds = [0, 1, 7, 8, 30, 33]
    renderer.set_screen(s.width, s.height, s.fovy)
    img = []
    for rank in range(5):
        renderer.set_strip_bounds(bounds, rank)
        part = parity.gpu_frame(renderer, s, flags=0, strip=(rank, 5))
        sl = slice(s.width * bounds[rank] * 8, s.width * bounds[rank + 1] * 8)
        assert part["t"].size == sl.stop - sl.start
        for k in ("t", "id", "shadowed"):
            assert np.array_equal(parity.bits(part[k]), parity.bits(cpu[k][sl])), (rank, k)
        img.append(part["rgb"])
    assert np.array_equal(np.concatenate(img), cpu["rgb"])
    with pytest.raises(RuntimeError):
        renderer.set_strip_bounds([0, 5, 5, 33], 0)           # an empty strip
    with pytest.raises(RuntimeError):
        renderer.set_strip_bounds([0, 5, 32], 0)              # does not end at the last row
    part = parity.gpu_frame(renderer, s, flags=0, strip=(1, 3))   # another split: equal strips again
    assert np.array_equal(part["rgb"], cpu["rgb"][3 * s.width * 88:3 * s.width * 176])
    renderer.render(renderer.params(s, strip=(0, 1)))
