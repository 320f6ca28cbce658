"""oracle/harness.py -- TEST INFRASTRUCTURE ONLY.

ctypes driver shared by the two CPU checkers:
  * `load("oracle")`  -> oracle/libugoracle.so   (ug_oracle.c, the plain-C restatement, prefix orc_)
  * `load("ref")`     -> oracle/_ref/libugref.so (the reference's own kernel sources compiled as host
                         C++ behind ref/ref_shim.h, prefix ref_)
Both export the same functions, so `run_frame` restates the reference's display() (main.cu:59-302)
once and runs it on either.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time
from typing import Dict, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {"oracle": ("libugoracle.so", "orc_"), "ref": (os.path.join("_ref", "libugref.so"), "ref_")}

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u = C.POINTER(C.c_uint)
_b = C.POINTER(C.c_ubyte)


def build(which: str = "all") -> None:
    """Compile the checkers (oracle/Makefile).  `ref` is only rebuilt where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, which], check=True)


def available(kind: str) -> bool:
    return os.path.exists(os.path.join(HERE, _LIBS[kind][0]))


class CpuLib:
    def __init__(self, kind: str):
        path, self.prefix = _LIBS[kind]
        self.kind = kind
        self.lib = C.CDLL(os.path.join(HERE, path))
        p = self.prefix
        sig = {
            "configure": (None, [C.c_int] * 4),
            "set_threads": (None, [C.c_int]),
            "set_slabs": (None, [C.c_int]),
            "get_threads": (C.c_int, []),
            "camera_setup": (None, [_f, C.c_float, C.c_float, C.c_int, _f]),
            "set_camcoords": (None, [_f]),
            "get_dirtex": (None, [_f]),
            "set_lights": (None, [_f, C.c_int]),
            "transform": (None, [_f, _f, C.c_int, C.c_int, C.c_float]),
            "grid_count": (None, [_i, _f, C.c_int, _u, _f]),
            "build_grid": (C.c_long, [_i, _f, C.c_int, _u, _u, _u, _u, C.c_long]),
            "grid_count_spherical": (None, [_i, _f, C.c_int, _u, _f, C.c_float, C.c_float]),
            "build_grid_spherical": (C.c_long, [_i, _f, C.c_int, _u, _u, _u, _u, C.c_long, C.c_float, C.c_float]),
            "map_rays_spherical": (None, [_f, _f, _u, _f, C.c_float, C.c_float]),
            "trace": (None, [_u, _u, _u, _f, _f, _f, _i, _i, _f, _i]),
            "map_rays": (None, [_f, _f, _u, _f, C.c_float]),
            "shadow": (None, [_u, _f, _i, _u, _u, _f, _f, _i, _u, _f]),
            "shade": (None, [_b, _f, _f, _f, _i, _f, _i, _f, C.c_int]),
            "spot_shade": (None, [_b, _f, _f, _f, _i, _f, _i, _f, C.c_int, _f]),
            "add_shadows": (None, [_b, _i]),
            "load_obj": (C.c_int, [C.c_char_p, _i, _i, _f, _i, _i]),
            "load_materials": (C.c_int, [C.c_char_p, _i, _f]),
        }
        if kind == "oracle":   # row R has no reference code: only the restatement carries the (authored) spec
            sig["reflect"] = (None, [_f, _i, C.c_int, _f, _f, _i, _b, _f, _i, _f, C.c_int, _f, C.c_int])
        for name, (res, args) in sig.items():
            fn = getattr(self.lib, p + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, "_" + name, fn)

    # ---- thin typed wrappers -------------------------------------------------------------
    @staticmethod
    def _p(a: np.ndarray, t):
        assert a.flags["C_CONTIGUOUS"], "need a contiguous array"
        return a.ctypes.data_as(t)

    def configure(self, W, H, fovy=45, tex_q8=0):
        self.W, self.H, self.nbx, self.nby = W, H, W // 8, H // 8
        self._configure(W, H, fovy, tex_q8)

    def set_threads(self, n: int):
        self._set_threads(n)

    def set_slabs(self, n: int):
        """NUM_SLABS (main.cu.h:18; the committed value is 1): depth slabs of every grid."""
        self.slabs = max(int(n), 1)
        self._set_slabs(self.slabs)

    def get_threads(self) -> int:
        return self._get_threads()

    def camera_setup(self, eye, look, up, near, far, square=False) -> np.ndarray:
        pose = np.asarray(list(eye) + list(look) + list(up), dtype=np.float32)
        out = np.zeros(64, dtype=np.float32)
        self._camera_setup(self._p(pose, _f), near, far, int(square), self._p(out, _f))
        return out

    def set_camcoords(self, cc: np.ndarray):
        cc = np.ascontiguousarray(cc, dtype=np.float32)
        self._set_camcoords(self._p(cc, _f))

    def dirtex(self) -> np.ndarray:
        out = np.zeros(100, dtype=np.float32)
        self._get_dirtex(self._p(out, _f))
        return out

    def set_lights(self, lights: np.ndarray):
        lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1)
        self._set_lights(self._p(lights, _f), lights.size // 3)

    def transform(self, verts: np.ndarray, orig: np.ndarray, size: int, offset: int, rot: float):
        self._transform(self._p(verts, _f), self._p(orig, _f), size, offset, rot)

    def grid_count(self, faces, verts, spherical=None):
        T = faces.shape[0]
        size = np.zeros(T, dtype=np.uint32)
        z = np.zeros(T, dtype=np.float32)
        if spherical is None:
            self._grid_count(self._p(faces, _i), self._p(verts, _f), T, self._p(size, _u), self._p(z, _f))
        else:
            self._grid_count_spherical(self._p(faces, _i), self._p(verts, _f), T, self._p(size, _u), self._p(z, _f),
                                       float(spherical[0]), float(spherical[1]))
        return size, z

    def build_grid(self, faces, verts, spherical=None) -> Dict[str, np.ndarray]:
        """`spherical` = (xM, yM) selects buildSphericalGrid (frustum_grid.h:368-532) instead of buildGrid (:210-366)."""
        T = faces.shape[0]
        Cn = self.nbx * self.nby * getattr(self, "slabs", 1)
        span = np.zeros(Cn, dtype=np.uint32)
        offset = np.zeros(Cn, dtype=np.uint32)
        dummy = np.zeros(1, dtype=np.uint32)
        fn, extra = (self._build_grid, ()) if spherical is None else (self._build_grid_spherical, tuple(float(x) for x in spherical))
        P = fn(self._p(faces, _i), self._p(verts, _f), T, self._p(span, _u), self._p(offset, _u),
               self._p(dummy, _u), self._p(dummy, _u), 0, *extra)
        keys = np.zeros(max(P, 1), dtype=np.uint32)
        vals = np.zeros(max(P, 1), dtype=np.uint32)
        P2 = fn(self._p(faces, _i), self._p(verts, _f), T, self._p(span, _u), self._p(offset, _u),
                self._p(keys, _u), self._p(vals, _u), P, *extra)
        assert P2 == P
        return {"P": int(P), "keys": keys[:P], "values": vals[:P], "span": span, "offset": offset}

    def trace(self, grid, faces, verts) -> Dict[str, np.ndarray]:
        N = self.W * self.H
        out = {"normal": np.zeros(3 * N, np.float32), "t": np.zeros(N, np.float32), "dir": np.zeros(3 * N, np.float32),
               "shadowed": np.zeros(N, np.int32), "id": np.zeros(N, np.int32)}
        vals = grid["values"] if grid["P"] else np.zeros(1, np.uint32)
        self._trace(self._p(vals, _u), self._p(grid["span"], _u), self._p(grid["offset"], _u), self._p(out["normal"], _f),
                    self._p(out["t"], _f), self._p(out["dir"], _f), self._p(out["shadowed"], _i), self._p(out["id"], _i),
                    self._p(verts, _f), self._p(faces, _i))
        return out

    def map_rays(self, t, d, cam_pos, near, spherical=None) -> np.ndarray:
        """`spherical` = (xM, yM) selects mapSort_Effective_kernel (misc_kernel.cu:255-296) instead of mapSortkernel (:63-129)."""
        N = self.W * self.H
        m = np.zeros(2 * N, dtype=np.uint32)
        cam_pos = np.ascontiguousarray(cam_pos, dtype=np.float32)
        if spherical is None:
            self._map_rays(self._p(t, _f), self._p(d, _f), self._p(m, _u), self._p(cam_pos, _f), near)
        else:
            self._map_rays_spherical(self._p(t, _f), self._p(d, _f), self._p(m, _u), self._p(cam_pos, _f), float(spherical[0]), float(spherical[1]))
        return m

    def shadow(self, grid, faces, verts, t, d, shadowed, m, cam_pos):
        cam_pos = np.ascontiguousarray(cam_pos, dtype=np.float32)
        vals = grid["values"] if grid["P"] else np.zeros(1, np.uint32)
        self._shadow(self._p(vals, _u), self._p(verts, _f), self._p(faces, _i), self._p(grid["span"], _u),
                     self._p(grid["offset"], _u), self._p(t, _f), self._p(d, _f), self._p(shadowed, _i), self._p(m, _u),
                     self._p(cam_pos, _f))

    def shade(self, normal, t, d, ids, cam_pos, mat_idx, mat_list, spot: bool = False) -> np.ndarray:
        """`ids` (triangle ids, -2 on miss) is overwritten with material indices like the reference does.
        `spot`: Shader::spotlight_shade (shader.h:88-113) instead of simpleShade (:70-86)."""
        N = self.W * self.H
        img = np.zeros(3 * N, dtype=np.uint8)
        cam_pos = np.ascontiguousarray(cam_pos, dtype=np.float32)
        padded = np.concatenate([np.full(2, -1, np.int32), np.ascontiguousarray(mat_idx, np.int32)])  # mat_idx[-2] readable
        base = padded.ctypes.data + 8
        mat_list = np.ascontiguousarray(mat_list, dtype=np.float32)
        if spot:
            dump = np.zeros(2 * N, dtype=np.float32)
            self._spot_shade(self._p(img, _b), self._p(normal, _f), self._p(t, _f), self._p(d, _f), self._p(ids, _i),
                             self._p(cam_pos, _f), C.cast(base, _i), self._p(mat_list, _f), mat_list.size // 6, self._p(dump, _f))
            self.last_spot_angles = dump
            return img
        self._shade(self._p(img, _b), self._p(normal, _f), self._p(t, _f), self._p(d, _f), self._p(ids, _i),
                    self._p(cam_pos, _f), C.cast(base, _i), self._p(mat_list, _f), mat_list.size // 6)
        return img

    def add_shadows(self, img, shadowed):
        self._add_shadows(self._p(img, _b), self._p(shadowed, _i))

    def reflect(self, verts, faces, t, d, ids, img, cam_pos, mat_idx, mat_list, mat_reflect, bounces: int):
        """Row R (authored spec): reflection bounces on top of the shaded + shadowed frame `img` (modified in place)."""
        cam_pos = np.ascontiguousarray(cam_pos, dtype=np.float32)
        mat_idx = np.ascontiguousarray(mat_idx, np.int32)
        mat_list = np.ascontiguousarray(mat_list, np.float32)
        mat_reflect = np.ascontiguousarray(mat_reflect, np.float32)
        self._reflect(self._p(verts, _f), self._p(faces, _i), faces.shape[0], self._p(t, _f), self._p(d, _f), self._p(ids, _i),
                      self._p(img, _b), self._p(cam_pos, _f), self._p(mat_idx, _i), self._p(mat_list, _f), mat_list.size // 6,
                      self._p(mat_reflect, _f), bounces)

    def load_obj(self, path: str):
        nv, nf = C.c_int(0), C.c_int(0)
        rc = self._load_obj(path.encode(), C.byref(nv), C.byref(nf), None, None, None)
        if rc != 0:
            raise OSError(f"cannot load {path}")
        V = np.zeros(3 * nv.value, np.float32)
        F = np.zeros(3 * nf.value, np.int32)
        M = np.zeros(nf.value, np.int32)
        self._load_obj(path.encode(), C.byref(nv), C.byref(nf), self._p(V, _f), self._p(F, _i), self._p(M, _i))
        return V.reshape(-1, 3), F.reshape(-1, 3), M

    def load_materials(self, path: str) -> np.ndarray:
        nm = C.c_int(0)
        if self._load_materials(path.encode(), C.byref(nm), None) != 0:
            raise OSError(f"cannot load {path}")
        out = np.zeros(6 * nm.value, np.float32)
        self._load_materials(path.encode(), C.byref(nm), self._p(out, _f))
        return out.reshape(-1, 6)


_cache: Dict[str, CpuLib] = {}


def load(kind: str) -> CpuLib:
    if kind not in _cache:
        _cache[kind] = CpuLib(kind)
    return _cache[kind]


def material_table(scene) -> np.ndarray:
    """What Model::some_material (scene.h:400-427) would read from the generated material file: the
    %.6f-printed values parsed back as floats."""
    rows = []
    for m in scene.materials:
        rows.append([float("%.6f" % x) for x in (*m.ka, *m.kd)])
    return np.asarray(rows, dtype=np.float32)


def reflect_table(scene) -> np.ndarray:
    """Reflectivity per material as the MTL `r` line carries it (%.6f round trip)."""
    return np.asarray([float("%.6f" % m.reflect) for m in scene.materials], dtype=np.float32)


def run_frame(lib: CpuLib, scene, frame: int = 0, threads: int = 0, stages: str = "all",
              timings: Optional[dict] = None, reflect: bool = False, spherical: bool = False,
              spot: bool = False, slabs: int = 1) -> Dict[str, np.ndarray]:
    """One display() (main.cu:59-302) on the CPU: [transform ->] camera grid -> primary rays -> per
    light (light camera, ray->cell map, light grid, shadow rays) -> shade -> shadow modulation.
    Returns every stage's output for bit-exact comparison.  `spherical` runs the per-light steps the way the committed
    display() does (main.cu:186-192: spherical ray map + spherical grid with x_max = y_max = M_PI) instead of the
    perspective variant north-star names; `spot` shades with spotlight_shade, as display() does from its third frame on
    (main.cu:205-219); `slabs` = NUM_SLABS (main.cu.h:18), the depth slabs of every grid."""
    sph = (np.float32(np.pi), np.float32(np.pi)) if spherical else None
    lib.configure(scene.width, scene.height, scene.fovy, 0)
    lib.set_threads(threads)
    lib.set_slabs(slabs)
    faces = np.ascontiguousarray(scene.faces, dtype=np.int32)
    verts = np.ascontiguousarray(scene.verts, dtype=np.float32).copy()
    out: Dict[str, np.ndarray] = {}
    tm = timings if timings is not None else {}

    def tick(name, t0):
        tm[name] = tm.get(name, 0.0) + (time.perf_counter() - t0)

    t0 = time.perf_counter()
    if scene.dyn_size:
        orig = np.ascontiguousarray(scene.verts[scene.dyn_offset:scene.dyn_offset + scene.dyn_size]).copy()
        rot = np.float32(np.float32(scene.rot_step) * np.float32(frame + 1))
        lib.transform(verts.reshape(-1), orig.reshape(-1), scene.dyn_size, scene.dyn_offset, float(rot))
    tick("transform", t0)
    out["verts"] = verts

    lights = np.asarray([l.pos for l in scene.lights], dtype=np.float32).reshape(-1, 3)
    lib.set_lights(lights[:2])
    cam_cc = lib.camera_setup(scene.eye, scene.look, scene.up, scene.near, scene.far)
    out["camcoords"] = cam_cc
    out["dirtex"] = lib.dirtex()
    cam_pos = np.asarray(scene.eye, dtype=np.float32)

    t0 = time.perf_counter()
    grid = lib.build_grid(faces, verts)
    tick("grid", t0)
    out.update({"cam_P": grid["P"], "cam_values": grid["values"], "cam_span": grid["span"], "cam_offset": grid["offset"]})
    if stages == "grid":
        return out

    t0 = time.perf_counter()
    prim = lib.trace(grid, faces, verts)
    tick("trace", t0)
    out.update({"t": prim["t"], "id": prim["id"].copy(), "normal": prim["normal"], "dir": prim["dir"]})
    shadowed = prim["shadowed"]
    if stages == "trace":
        return out

    for li, light in enumerate(scene.lights):
        cc = lib.camera_setup(light.pos, light.look, light.up, scene.near, scene.far, square=True)
        out[f"light{li}_camcoords"] = cc
        t0 = time.perf_counter()
        m = lib.map_rays(prim["t"], prim["dir"], cam_pos, scene.near, spherical=sph)
        tick("map", t0)
        t0 = time.perf_counter()
        lgrid = lib.build_grid(faces, verts, spherical=sph)
        tick("grid", t0)
        t0 = time.perf_counter()
        lib.shadow(lgrid, faces, verts, prim["t"], prim["dir"], shadowed, m, cam_pos)
        tick("shadow", t0)
        out[f"light{li}_cell"] = m[scene.num_pixels:].copy()
        out.update({f"light{li}_P": lgrid["P"], f"light{li}_values": lgrid["values"], f"light{li}_span": lgrid["span"],
                    f"light{li}_offset": lgrid["offset"]})
    out["shadowed"] = shadowed.copy()

    t0 = time.perf_counter()
    mat_list = material_table(scene)
    ids = prim["id"].copy()
    img = lib.shade(prim["normal"], prim["t"], prim["dir"], ids, cam_pos, scene.face_mat, mat_list, spot=spot)
    out["rgb_unshadowed"] = img.copy()
    lib.add_shadows(img, shadowed)
    tick("shade", t0)
    out["material_id"] = ids
    out["rgb"] = img
    if getattr(scene, "bounces", 0) > 0 and lib.kind == "oracle" and reflect:
        t0 = time.perf_counter()
        img2 = img.copy()
        lib.reflect(verts.reshape(-1), faces, prim["t"], prim["dir"], prim["id"], img2, cam_pos, scene.face_mat, mat_list,
                    reflect_table(scene), scene.bounces)
        tick("reflect", t0)
        out["rgb_reflect"] = img2
    return out


def write_ppm(path: str, rgb: np.ndarray, W: int, H: int, flip: bool = True) -> None:
    """Binary PPM for eyeballing (row 0 of the buffer is the bottom of the view)."""
    img = rgb.reshape(H, W, 3)
    if flip:
        img = img[::-1]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (W, H))
        f.write(np.ascontiguousarray(img).tobytes())
