"""Procedural scenes for the five BASELINE.json configs.

The reference ships no scene files, so every input is synthetic: geometry is generated here with numpy
(PRNG = splitmix64-seeded numpy Generator), written as OBJ + one material file that is valid both as
the `mtllib` MTL (obj_parser/obj_parser.cpp:177-294) and as the positional 6-float material description
read by Model::some_material (scene.h:370-439), and loaded back through the loader under test.

All geometry is kept in front of the eye plane of the camera and of every light and inside NDC z in
(0,1), because the reference has neither clipping nor culling; triangles
outside the frustum sideways are kept on purpose - they exercise the border-cell clamp.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Tuple

import numpy as np


@dataclasses.dataclass
class Material:
    name: str
    ka: Tuple[float, float, float]
    kd: Tuple[float, float, float]
    reflect: float = 0.0


@dataclasses.dataclass
class Light:
    pos: Tuple[float, float, float]
    look: Tuple[float, float, float]
    up: Tuple[float, float, float] = (0.0, 1.0, 0.0)


@dataclasses.dataclass
class Scene:
    """One BASELINE config: geometry + frame parameters (main.cu:82-90,158-164 made data)."""
    name: str
    verts: np.ndarray          # float32 [V,3]
    faces: np.ndarray          # int32 [T,3], 0-based
    face_mat: np.ndarray       # int32 [T]
    materials: List[Material]
    width: int
    height: int
    eye: Tuple[float, float, float]
    look: Tuple[float, float, float]
    up: Tuple[float, float, float]
    near: float = 0.1
    far: float = 100.0
    fovy: int = 45
    lights: List[Light] = dataclasses.field(default_factory=list)
    bounces: int = 0
    dyn_offset: int = 0        # Model::init_orig_list(size, offset), scene.h:336-368
    dyn_size: int = 0
    rot_step: float = 0.05     # per_frame_funcs.h:15 lightRotFactor += 0.05

    @property
    def num_tris(self) -> int:
        return int(self.faces.shape[0])

    @property
    def num_pixels(self) -> int:
        return self.width * self.height

    def rays_per_frame(self) -> int:
        """primary + one shadow ray per pixel per light (reflection rays are counted live)."""
        return self.num_pixels * (1 + len(self.lights))


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


# ----------------------------------------------------------------------------- primitives
def icosphere(subdiv: int, radius: float, centre) -> Tuple[np.ndarray, np.ndarray]:
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = list(f)
    for _ in range(subdiv):
        cache = {}
        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]
        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    V = np.asarray(verts) * radius + np.asarray(centre, dtype=np.float64)
    return V.astype(np.float32), np.asarray(faces, dtype=np.int32)


def grid_plane(nx: int, ny: int, x0, x1, y0, y1, zfun=None) -> Tuple[np.ndarray, np.ndarray]:
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    Z = np.zeros_like(X) if zfun is None else zfun(X, Y)
    V = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    a = (j * (nx + 1) + i).ravel()
    b = a + 1
    c = a + (nx + 1)
    d = c + 1
    F = np.empty((2 * nx * ny, 3), dtype=np.int32)
    F[0::2] = np.stack([a, b, d], axis=1)
    F[1::2] = np.stack([a, d, c], axis=1)
    return V, F


def quad_patch(p0, pu, pv, nu: int, nv: int) -> Tuple[np.ndarray, np.ndarray]:
    """Tessellated parallelogram p0 + s*pu + t*pv, s,t in [0,1]."""
    V, F = grid_plane(nu, nv, 0.0, 1.0, 0.0, 1.0)
    p0, pu, pv = (np.asarray(a, dtype=np.float64) for a in (p0, pu, pv))
    P = p0[None, :] + V[:, 0:1].astype(np.float64) * pu[None, :] + V[:, 1:2].astype(np.float64) * pv[None, :]
    return P.astype(np.float32), F


def cylinder(cx, cy, z0, z1, r, sides: int, rings: int, r_top=None) -> Tuple[np.ndarray, np.ndarray]:
    """Open cylinder / cone frustum, `sides` x `rings` quads."""
    r_top = r if r_top is None else r_top
    ang = np.linspace(0.0, 2.0 * np.pi, sides, endpoint=False)
    zs = np.linspace(z0, z1, rings + 1)
    rr = np.linspace(r, r_top, rings + 1)
    V = np.array([[cx + rr[k] * np.cos(a), cy + rr[k] * np.sin(a), zs[k]] for k in range(rings + 1) for a in ang],
                 dtype=np.float32)
    F = []
    for k in range(rings):
        for s in range(sides):
            a = k * sides + s
            b = k * sides + (s + 1) % sides
            c = a + sides
            d = b + sides
            F += [(a, b, d), (a, d, c)]
    return V, np.asarray(F, dtype=np.int32)


def box(lo, hi, n: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    parts = [quad_patch(lo, ex, ey, n, n), quad_patch(lo + ez, ex, ey, n, n), quad_patch(lo, ex, ez, n, n),
             quad_patch(lo + ey, ex, ez, n, n), quad_patch(lo, ey, ez, n, n), quad_patch(lo + ex, ey, ez, n, n)]
    return merge(parts)


def merge(parts):
    Vs, Fs, off = [], [], 0
    for V, F in parts:
        Vs.append(V)
        Fs.append(F + off)
        off += V.shape[0]
    return np.concatenate(Vs).astype(np.float32), np.concatenate(Fs).astype(np.int32)


class _Builder:
    def __init__(self):
        self.parts, self.mats = [], []

    def add(self, VF, mat: int):
        self.parts.append(VF)
        self.mats.append(np.full(VF[1].shape[0], mat, dtype=np.int32))

    @property
    def ntris(self):
        return sum(p[1].shape[0] for p in self.parts)

    def build(self):
        V, F = merge(self.parts)
        return V, F, np.concatenate(self.mats)


# ----------------------------------------------------------------------------- the five configs
def scene_a(width: int = 512, height: int = 512) -> Scene:
    """Config A: icosphere(4) r=4 at (0,0,5) on a 75x75-quad plane [-20,20]^2 = 16 370 tris; 1 light."""
    b = _Builder()
    b.add(grid_plane(75, 75, -20.0, 20.0, -20.0, 20.0), 0)
    b.add(icosphere(4, 4.0, (0.0, 0.0, 5.0)), 1)
    V, F, M = b.build()
    mats = [Material("floor", (0.2, 0.2, 0.2), (0.55, 0.6, 0.5)), Material("ball", (0.2, 0.1, 0.1), (0.9, 0.35, 0.25))]
    return Scene("A", V, F, M, mats, width, height, eye=(0.0, -30.0, 12.0), look=(0.0, 0.0, 2.0), up=(0.0, 0.0, 1.0),
                 lights=[Light((12.0, -12.0, 25.0), (0.0, 0.0, 2.0), (0.0, 1.0, 0.0))])


def scene_b(width: int = 1024, height: int = 1024, target_tris: int = 80000) -> Scene:
    """Config B: "cathedral" hall 40x20x15 open towards the camera, two rows of 16-gon columns with
    arches (half-tori approximated by bent cylinders), tessellated to exactly `target_tris`."""
    b = _Builder()
    L, Wd, Hh = 40.0, 20.0, 15.0
    # columns + arches first, then walls, then the floor takes the remainder
    ncol = 8
    for row, y in enumerate((-5.0, 5.0)):
        for k in range(ncol):
            x = 3.0 + k * (L - 6.0) / (ncol - 1)
            b.add(cylinder(x, y, 0.0, 9.0, 0.7, 16, 24), 2)
            b.add(cylinder(x, y, 9.0, 9.8, 0.7, 16, 2, r_top=1.1), 2)
    # arches between consecutive columns: poly-line tubes
    for y in (-5.0, 5.0):
        for k in range(ncol - 1):
            x0 = 3.0 + k * (L - 6.0) / (ncol - 1)
            x1 = 3.0 + (k + 1) * (L - 6.0) / (ncol - 1)
            b.add(_arch(x0, x1, y, 9.8, 2.2, 0.35, 12, 8), 2)
    wall_n = 40
    b.add(quad_patch((0, -Wd / 2, 0), (L, 0, 0), (0, 0, Hh), 2 * wall_n, wall_n), 1)      # y = -10 wall
    b.add(quad_patch((0, Wd / 2, 0), (L, 0, 0), (0, 0, Hh), 2 * wall_n, wall_n), 1)       # y = +10 wall
    b.add(quad_patch((L, -Wd / 2, 0), (0, Wd, 0), (0, 0, Hh), wall_n, wall_n), 1)          # far wall
    b.add(quad_patch((0, -Wd / 2, Hh), (L, 0, 0), (0, Wd, 0), 2 * wall_n, wall_n), 1)     # ceiling
    rest = target_tris - b.ntris
    assert rest > 0 and rest % 2 == 0, rest
    nq = rest // 2
    # floor: nx*ny quads + a strip to make the count exact
    nx = int((2 * nq) ** 0.5)
    ny = nq // nx
    b.add(quad_patch((0, -Wd / 2, 0), (L, 0, 0), (0, Wd, 0), nx, ny), 0)
    left = nq - nx * ny
    if left:
        b.add(quad_patch((-2.0, -Wd / 2, 0), (2.0, 0, 0), (0, Wd, 0), 1, left), 0)
    V, F, M = b.build()
    assert F.shape[0] == target_tris, F.shape
    mats = [Material("floor", (0.2, 0.2, 0.2), (0.6, 0.6, 0.65), reflect=0.5),
            Material("wall", (0.2, 0.2, 0.2), (0.75, 0.7, 0.6)),
            Material("stone", (0.2, 0.2, 0.2), (0.8, 0.8, 0.75))]
    return Scene("B", V, F, M, mats, width, height, eye=(-6.0, 0.5, 6.0), look=(20.0, 0.0, 5.0), up=(0.0, 0.0, 1.0),
                 lights=[Light((-5.0, -6.0, 13.5), (22.0, 2.0, 2.0), (0.0, 0.0, 1.0))], bounces=1)


def _arch(x0, x1, y, z, rise, r, sides, segs):
    """Tube along a half-ellipse from (x0,y,z) to (x1,y,z) rising by `rise`."""
    th = np.linspace(np.pi, 0.0, segs + 1)
    cx, a = 0.5 * (x0 + x1), 0.5 * (x1 - x0)
    path = np.stack([cx + a * np.cos(th), np.full_like(th, y), z + rise * np.sin(th)], axis=1)
    ang = np.linspace(0.0, 2.0 * np.pi, sides, endpoint=False)
    V = []
    for k in range(segs + 1):
        tan = path[min(k + 1, segs)] - path[max(k - 1, 0)]
        tan /= np.linalg.norm(tan)
        n1 = np.array([0.0, 1.0, 0.0])
        n2 = np.cross(tan, n1)
        for aa in ang:
            V.append(path[k] + r * (np.cos(aa) * n1 + np.sin(aa) * n2))
    F = []
    for k in range(segs):
        for s in range(sides):
            a_ = k * sides + s
            b_ = k * sides + (s + 1) % sides
            F += [(a_, b_, b_ + sides), (a_, b_ + sides, a_ + sides)]
    return np.asarray(V, dtype=np.float32), np.asarray(F, dtype=np.int32)


def scene_c(width: int = 1920, height: int = 1080, target_tris: int = 174000) -> Scene:
    """Config C: dynamic scene - 256x256-quad terrain (131 072 tris) + instanced low-poly trees whose
    vertex range is the dynamic sub-range rotated each frame by copy_data_transform."""
    rng = _rng(0xC)
    b = _Builder()
    terr = lambda X, Y: 0.6 * np.sin(0.21 * X) * np.cos(0.17 * Y) + 0.25 * np.sin(0.53 * X + 0.4 * Y)
    b.add(grid_plane(256, 256, -14.0, 46.0, -16.0, 44.0, terr), 0)
    static_v = b.parts[0][0].shape[0]
    rest = target_tris - b.ntris                       # 42 928
    # one tree = trunk (8-gon, 2 rings = 32) + 3 cones (12-gon, 3 rings = 72 each) = 248 tris
    per = 32 + 3 * 72
    ntree = rest // per
    tree_parts = []
    for _ in range(ntree):
        x, y = rng.uniform(1.5, 22.5), rng.uniform(0.5, 21.5)     # inside the transform's source box
        h = rng.uniform(3.0, 7.5)
        tree_parts.append((cylinder(x, y, 0.3, 0.35 * h, 0.18, 8, 2), 2))
        for lvl in range(3):
            z0 = 0.3 * h + lvl * 0.22 * h
            tree_parts.append((cylinder(x, y, z0, z0 + 0.32 * h, 0.9 - 0.22 * lvl, 12, 3, r_top=0.02), 1))
    left = rest - ntree * per
    if left:
        assert left % 2 == 0
        tree_parts.append((quad_patch((11.0, 10.0, 8.0), (2.0, 0, 0), (0, 2.0, 0), 1, left // 2), 1))
    for vf, m in tree_parts:
        b.add(vf, m)
    V, F, M = b.build()
    assert F.shape[0] == target_tris
    mats = [Material("ground", (0.2, 0.2, 0.2), (0.45, 0.6, 0.35)), Material("leaf", (0.1, 0.2, 0.1), (0.25, 0.7, 0.3)),
            Material("bark", (0.2, 0.15, 0.1), (0.55, 0.4, 0.25))]
    return Scene("C", V, F, M, mats, width, height, eye=(14.5, -14.0, 13.0), look=(14.5, 13.0, 2.0), up=(0.0, 0.0, 1.0),
                 lights=[Light((34.0, -10.0, 30.0), (14.5, 13.0, 2.0), (0.0, 0.0, 1.0)),
                         Light((-8.0, -8.0, 32.0), (14.5, 13.0, 2.0), (0.0, 0.0, 1.0))],
                 dyn_offset=static_v, dyn_size=V.shape[0] - static_v)


def scene_d(width: int = 4096, height: int = 4096, target_tris: int = 282000) -> Scene:
    """Config D: "conference" - room box + 30 x (table + 6 chairs) of tessellated boxes / cylinders."""
    b = _Builder()
    L, Wd, Hh = 36.0, 24.0, 9.0
    for i in range(30):
        tx, ty = 4.0 + (i % 6) * 5.5, -9.0 + (i // 6) * 4.5
        b.add(box((tx - 1.6, ty - 0.9, 1.45), (tx + 1.6, ty + 0.9, 1.6), 10), 1)                   # table top
        b.add(cylinder(tx, ty, 0.0, 1.45, 0.25, 24, 16), 2)                                          # pedestal
        for c in range(6):
            cx, cy = tx - 1.2 + (c % 3) * 1.2, ty + (1.5 if c < 3 else -1.5)
            b.add(box((cx - 0.35, cy - 0.35, 0.75), (cx + 0.35, cy + 0.35, 0.85), 3), 3)            # seat
            b.add(box((cx - 0.35, cy + (0.3 if c < 3 else -0.35), 0.85), (cx + 0.35, cy + (0.35 if c < 3 else -0.3), 1.7), 3), 3)
            b.add(cylinder(cx, cy, 0.0, 0.75, 0.06, 8, 4), 2)
    n = 64
    b.add(quad_patch((0, -Wd / 2, 0), (L, 0, 0), (0, 0, Hh), 2 * n, n), 4)
    b.add(quad_patch((0, Wd / 2, 0), (L, 0, 0), (0, 0, Hh), 2 * n, n), 4)
    b.add(quad_patch((L, -Wd / 2, 0), (0, Wd, 0), (0, 0, Hh), n, n), 4)
    b.add(quad_patch((0, -Wd / 2, Hh), (L, 0, 0), (0, Wd, 0), 2 * n, n), 4)
    rest = target_tris - b.ntris
    assert rest > 0 and rest % 2 == 0, rest
    nq = rest // 2
    nx = int((1.5 * nq) ** 0.5)
    ny = nq // nx
    b.add(quad_patch((0, -Wd / 2, 0), (L, 0, 0), (0, Wd, 0), nx, ny), 0)
    left = nq - nx * ny
    if left:
        b.add(quad_patch((-1.5, -Wd / 2, 0), (1.5, 0, 0), (0, Wd, 0), 1, left), 0)
    V, F, M = b.build()
    assert F.shape[0] == target_tris
    mats = [Material("carpet", (0.2, 0.2, 0.2), (0.35, 0.4, 0.6)), Material("wood", (0.2, 0.2, 0.2), (0.7, 0.5, 0.3)),
            Material("steel", (0.2, 0.2, 0.2), (0.7, 0.7, 0.75)), Material("fabric", (0.2, 0.2, 0.2), (0.65, 0.25, 0.25)),
            Material("plaster", (0.2, 0.2, 0.2), (0.85, 0.85, 0.8))]
    return Scene("D", V, F, M, mats, width, height, eye=(-5.0, -1.0, 5.5), look=(18.0, 0.0, 1.5), up=(0.0, 0.0, 1.0),
                 lights=[Light((-4.0, 4.0, 8.6), (20.0, -1.0, 0.5), (0.0, 0.0, 1.0))])


def scene_e(width: int = 8192, height: int = 8192, target_tris: int = 1000000, nlights: int = 4) -> Scene:
    """Config E: random triangle soup, centroids uniform in a box in front of the camera at depth [5,80]."""
    rng = _rng(0xE)
    depth = rng.uniform(5.0, 80.0, target_tris)
    half = depth * np.tan(np.radians(22.5)) * 1.05          # a little wider than the frustum: border clamp
    cx = rng.uniform(-1.0, 1.0, target_tris) * np.minimum(half, 20.0)
    cz = rng.uniform(-1.0, 1.0, target_tris) * np.minimum(half, 20.0)
    cen = np.stack([cx, depth, cz], axis=1)
    edge = np.exp(rng.uniform(np.log(0.05), np.log(0.5), target_tris))
    d1 = rng.normal(size=(target_tris, 3))
    d2 = rng.normal(size=(target_tris, 3))
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    V = np.empty((target_tris, 3, 3))
    V[:, 0] = cen - 0.33 * edge[:, None] * (d1 + d2)
    V[:, 1] = V[:, 0] + edge[:, None] * d1
    V[:, 2] = V[:, 0] + edge[:, None] * d2
    V = V.reshape(-1, 3).astype(np.float32)
    F = np.arange(3 * target_tris, dtype=np.int32).reshape(-1, 3)
    M = (np.arange(target_tris) % 3).astype(np.int32)
    mats = [Material("red", (0.2, 0.1, 0.1), (0.85, 0.3, 0.3), reflect=0.3),
            Material("green", (0.1, 0.2, 0.1), (0.3, 0.8, 0.35), reflect=0.3),
            Material("blue", (0.1, 0.1, 0.2), (0.3, 0.4, 0.9), reflect=0.3)]
    lights = []
    for k in range(nlights):
        a = 2.0 * np.pi * k / max(nlights, 1)
        lights.append(Light((30.0 * np.cos(a), -25.0, 30.0 * np.sin(a)), (0.0, 42.0, 0.0), (0.0, 0.0, 1.0) if abs(np.cos(a)) > 0.5 else (1.0, 0.0, 0.0)))
    return Scene("E", V, F, M, mats, width, height, eye=(0.0, 0.0, 0.0), look=(0.0, 1.0, 0.0), up=(0.0, 0.0, 1.0),
                 lights=lights, bounces=2)


def scene_tiny(kind: str = "ka", width: int = 64, height: int = 64) -> Scene:
    """Known-answer micro scenes."""
    mats = [Material("m0", (0.2, 0.2, 0.2), (0.8, 0.6, 0.4)), Material("m1", (0.2, 0.2, 0.2), (0.3, 0.6, 0.9))]
    eye, look, up = (0.0, -10.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0)
    if kind == "one":            # a single triangle in front of the camera
        V = np.array([[-2, 0, -2], [2, 0, -2], [0, 0, 2]], dtype=np.float32)
        F = np.array([[0, 1, 2]], dtype=np.int32)
    elif kind == "tie":          # two coincident triangles: lower id must win every pixel
        V = np.array([[-2, 0, -2], [2, 0, -2], [0, 0, 2], [-2, 0, -2], [2, 0, -2], [0, 0, 2]], dtype=np.float32)
        F = np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32)
    elif kind == "behind":       # triangle behind the camera: |t| mirrors it into the image
        V = np.array([[-2, 0, -2], [2, 0, -2], [0, 0, 2], [-1, -20, -1], [1, -20, -1], [0, -20, 1]], dtype=np.float32)
        F = np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32)
    elif kind == "near":         # a triangle closer than NDC z = 0 turns its pixels into misses
        V = np.array([[-2, 0, -2], [2, 0, -2], [0, 0, 2], [-0.02, -9.95, -0.02], [0.02, -9.95, -0.02], [0, -9.95, 0.02]],
                     dtype=np.float32)
        F = np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32)
    elif kind == "outside":      # fully outside the frustum: clamped into border cells
        V = np.array([[-2, 0, -2], [2, 0, -2], [0, 0, 2], [30, 0, 30], [34, 0, 30], [32, 0, 34]], dtype=np.float32)
        F = np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32)
    elif kind == "ka":           # occluder above a floor quad: shadow on the floor
        V = np.array([[-4, -4, -2], [4, -4, -2], [4, 4, -2], [-4, 4, -2], [-1, -1, 0], [1, -1, 0], [0, 1, 0.5]],
                     dtype=np.float32)
        F = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], dtype=np.int32)
        eye, look = (0.0, -9.0, 5.0), (0.0, 0.0, -1.0)
    else:
        raise ValueError(kind)
    M = (np.arange(F.shape[0]) % 2).astype(np.int32)
    return Scene("tiny-" + kind, V, F, M, mats, width, height, eye=eye, look=look, up=up,
                 lights=[Light((3.0, -3.0, 9.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0))])


def scene_random(seed: int, ntris: int, width: int, height: int, nlights: int = 1) -> Scene:
    """Seeded soup used by property tests: small triangles in a slab in front of the camera."""
    rng = _rng(seed)
    cen = np.stack([rng.uniform(-6, 6, ntris), rng.uniform(4, 30, ntris), rng.uniform(-6, 6, ntris)], axis=1)
    size = np.exp(rng.uniform(np.log(0.1), np.log(2.5), ntris))[:, None]
    V = (cen[:, None, :] + size[:, None, :] * rng.normal(size=(ntris, 3, 3)) * 0.5).reshape(-1, 3).astype(np.float32)
    F = np.arange(3 * ntris, dtype=np.int32).reshape(-1, 3)
    M = (np.arange(ntris) % 2).astype(np.int32)
    mats = [Material("m0", (0.2, 0.2, 0.2), (0.8, 0.6, 0.4)), Material("m1", (0.2, 0.2, 0.2), (0.3, 0.6, 0.9))]
    lights = [Light((10.0 - 20.0 * k, -12.0, 14.0), (0.0, 17.0, 0.0), (0.0, 0.0, 1.0)) for k in range(nlights)]
    return Scene(f"rand-{seed}", V, F, M, mats, width, height, eye=(0.0, -8.0, 1.0), look=(0.0, 10.0, 0.0),
                 up=(0.0, 0.0, 1.0), lights=lights)


def scene_hostile(seed: int = 77, ntris: int = 600, width: int = 128, height: int = 128) -> Scene:
    """Parity stress: geometry the reference has no protection against - triangles straddling the eye
    plane and the light's plane (no near clipping: w <= 0 vertices give flipped / huge NDC boxes), triangles behind the
    camera (mirrored into the image by |t|), zero-area and needle triangles, exact duplicates (tie rule), triangles far
    outside the frustum (border clamp)."""
    rng = _rng(seed)
    cen = np.stack([rng.uniform(-8, 8, ntris), rng.uniform(-14, 25, ntris), rng.uniform(-8, 8, ntris)], axis=1)
    size = np.exp(rng.uniform(np.log(0.05), np.log(9.0), ntris))[:, None, None]
    V = cen[:, None, :] + size * rng.normal(size=(ntris, 3, 3)) * 0.5
    V[5] = V[4]                                   # exact duplicate of triangle 4
    V[10, 2] = V[10, 1]                           # zero-area (two equal vertices)
    V[11, 2] = 0.5 * (V[11, 0] + V[11, 1])        # zero-area (collinear)
    V[12] = np.array([[0.0, -8.0, 1.0]] * 3) + np.array([[0, 0, 0], [1e-3, 30, 0], [0, 30, 1e-3]])   # needle through the eye
    V = V.reshape(-1, 3).astype(np.float32)
    F = np.arange(3 * ntris, dtype=np.int32).reshape(-1, 3)
    M = (np.arange(ntris) % 2).astype(np.int32)
    mats = [Material("m0", (0.2, 0.2, 0.2), (0.8, 0.6, 0.4)), Material("m1", (0.2, 0.2, 0.2), (0.3, 0.6, 0.9))]
    return Scene(f"hostile-{seed}", V, F, M, mats, width, height, eye=(0.0, -8.0, 1.0), look=(0.0, 10.0, 0.0), up=(0.0, 0.0, 1.0),
                 lights=[Light((6.0, 2.0, 9.0), (0.0, 8.0, 0.0), (0.0, 0.0, 1.0)), Light((-3.0, 14.0, 3.0), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0))])


SCENES = {"A": scene_a, "B": scene_b, "C": scene_c, "D": scene_d, "E": scene_e}


# ----------------------------------------------------------------------------- file output
def write_material_file(path: str, materials: List[Material]) -> None:
    """One file, two readers: obj_parse_mtl_file (obj_parser.cpp:177-294) and the positional
    Model::some_material (scene.h:400-427: 3 tokens, 3 floats, 1 token, 3 floats, 11 tokens, texture|NA)."""
    with open(path, "w") as f:
        for m in materials:
            f.write(f"newmtl {m.name}\n")
            f.write("Ka %.6f %.6f %.6f\n" % tuple(m.ka))
            f.write("Kd %.6f %.6f %.6f\n" % tuple(m.kd))
            f.write("Ks 1.000000 1.000000 1.000000\n")
            f.write("Ns 10.000000\n")
            f.write("r %.6f\n" % m.reflect)
            f.write("illum 2\n")
            f.write("map_Ka NA\n")


def write_obj(scene: Scene, obj_path: str, mtl_path: str | None = None) -> Tuple[str, str]:
    """Write OBJ (+ material file).  `mtllib` is resolved relative to the CWD by the reference parser
    (obj_parser.cpp:388-390), so an absolute path is written."""
    obj_path = os.path.abspath(obj_path)
    mtl_path = os.path.abspath(mtl_path or os.path.splitext(obj_path)[0] + ".mtl")
    write_material_file(mtl_path, scene.materials)
    with open(obj_path, "w") as f:
        f.write(f"# {scene.name}: {scene.num_tris} triangles\n")
        f.write(f"mtllib {mtl_path}\n")
        np.savetxt(f, scene.verts.astype(np.float64), fmt="v %.9g %.9g %.9g")
        cur = None
        order = np.arange(scene.num_tris)
        # faces keep their order (triangle id = file order); emit usemtl on every change
        change = np.flatnonzero(np.diff(scene.face_mat, prepend=-12345))
        bounds = list(change) + [scene.num_tris]
        for s, e in zip(bounds[:-1], bounds[1:]):
            f.write(f"usemtl {scene.materials[int(scene.face_mat[s])].name}\n")
            np.savetxt(f, scene.faces[s:e] + 1, fmt="f %d %d %d")
        del cur, order
    return obj_path, mtl_path
