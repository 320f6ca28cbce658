"""B200-native per-frame hot path of sushruta/uniformgrid-raytracing.

Python here is only a ctypes veneer over the C-ABI in include/ugrt.h (libugrt.so, hand-written sm_100a
CUDA).  There is no Python or CPU implementation of any stage: if the library is missing or no GPU is
usable, construction of `Renderer` raises.  See DESIGN.md / INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from . import scenes  # noqa: F401  (scene generators for the BASELINE configs)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UGRT_LIB") or os.path.join(PKG_DIR, "libugrt.so")   # UGRT_LIB: development aid (A/B builds)

UG_OK = 0
UG_MAX_LIGHTS = 8
UG_TILE = 8
(FLAG_FULL_BUFFERS, FLAG_NO_GRAPH, FLAG_STAGE_EVENTS, FLAG_GATHER_FRAMES, FLAG_SPHERICAL_LIGHTS, FLAG_SPOT_SHADE,
 FLAG_GATHER_DEFERRED) = 1, 2, 4, 8, 16, 32, 64
LIGHT_GRID_PERSPECTIVE, LIGHT_GRID_SPHERICAL = 0, 1

(BUF_VERTEXLIST, BUF_CAM_VALUES, BUF_CAM_SPAN, BUF_CAM_OFFSET, BUF_T, BUF_ID, BUF_NORMAL, BUF_DIR, BUF_SHADOWED, BUF_RGB,
 BUF_LIGHT_VALUES, BUF_LIGHT_SPAN, BUF_LIGHT_OFFSET, BUF_LIGHT_CELL, BUF_SIZELIST, BUF_RGB_FULL, BUF_CAM_COUNTERS) = range(17)

_BUF_DTYPE = {BUF_VERTEXLIST: np.float32, BUF_CAM_VALUES: np.uint32, BUF_CAM_SPAN: np.uint32, BUF_CAM_OFFSET: np.uint32,
              BUF_T: np.float32, BUF_ID: np.int32, BUF_NORMAL: np.float32, BUF_DIR: np.float32, BUF_SHADOWED: np.int32,
              BUF_RGB: np.uint8, BUF_LIGHT_VALUES: np.uint32, BUF_LIGHT_SPAN: np.uint32, BUF_LIGHT_OFFSET: np.uint32,
              BUF_LIGHT_CELL: np.uint32, BUF_SIZELIST: np.uint32, BUF_RGB_FULL: np.uint8, BUF_CAM_COUNTERS: np.uint32}


class Camera(C.Structure):  # ug_camera == camera.h:7-47
    _fields_ = [("cx", C.c_float), ("cy", C.c_float), ("cz", C.c_float), ("lx", C.c_float), ("ly", C.c_float), ("lz", C.c_float),
                ("ux", C.c_float), ("uy", C.c_float), ("uz", C.c_float), ("nearPlane", C.c_float), ("farPlane", C.c_float),
                ("worldori", C.c_float * 4), ("modelview_matrix", C.c_float * 16), ("projection_matrix", C.c_float * 16),
                ("mvp_matrix", C.c_float * 16), ("frustum_plane_eq", (C.c_float * 6) * 6), ("frustumcorner", (C.c_float * 3) * 8),
                ("fovy", C.c_float), ("aspect", C.c_float)]


class LightS(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("look", C.c_float * 3), ("up", C.c_float * 3)]


class FrameParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fovy", C.c_float), ("nearPlane", C.c_float), ("farPlane", C.c_float),
                ("eye", C.c_float * 3), ("look", C.c_float * 3), ("up", C.c_float * 3), ("num_lights", C.c_int),
                ("lights", LightS * UG_MAX_LIGHTS), ("transform", C.c_int), ("rot_factor", C.c_float),
                ("strip_index", C.c_int), ("strip_count", C.c_int), ("flags", C.c_int), ("bounces", C.c_int), ("num_slabs", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("num_faces", C.c_int), ("num_vertices", C.c_int), ("num_materials", C.c_int), ("width", C.c_int),
                ("height", C.c_int), ("num_cells", C.c_int), ("num_pixels", C.c_int), ("cam_pairs", C.c_uint),
                ("light_pairs", C.c_uint * UG_MAX_LIGHTS), ("pair_capacity", C.c_uint), ("overflowed", C.c_int),
                ("kernel_launches", C.c_int), ("ms_frame", C.c_float), ("ms_transform", C.c_float), ("ms_cam_grid", C.c_float),
                ("ms_trace", C.c_float), ("ms_light_grid", C.c_float), ("ms_regroup", C.c_float), ("ms_shadow", C.c_float),
                ("ms_shade", C.c_float), ("reflect_rays", C.c_uint), ("reflect_pairs", C.c_uint), ("ms_own", C.c_float)]


def build(verbose: bool = False) -> str:
    """Compile libugrt.so in-tree for sm_100a (csrc/build.sh)."""
    r = subprocess.run(["bash", os.path.join(PKG_DIR, "csrc", "build.sh")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libugrt.so build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout, r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """The C-ABI library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU path)")
        L = C.CDLL(LIB_PATH)
        vp, f, i = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
        sig = {
            "ug_create": (C.c_int, [C.POINTER(vp), C.c_int]), "ug_destroy": (None, [vp]), "ug_last_error": (C.c_char_p, [vp]),
            "ug_version": (C.c_char_p, []), "ug_frame_const_bytes": (C.c_size_t, []), "ug_set_stream": (C.c_int, [vp, vp]), "ug_comm_unique_id": (C.c_int, [vp]), "ug_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]), "ug_comm_flush": (C.c_int, [vp]), "ug_set_rgb_buffer": (C.c_int, [vp, vp, C.c_size_t]), "ug_gl_register_pbo": (C.c_int, [vp, C.c_uint]), "ug_gl_map_pbo": (C.c_int, [vp]), "ug_gl_unmap_pbo": (C.c_int, [vp]), "ug_gl_unregister_pbo": (C.c_int, [vp]), "ug_synchronize": (C.c_int, [vp]),
            "ug_model_load": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
            "ug_model_set": (C.c_int, [vp, f, C.c_int, i, C.c_int, i, f, C.c_int]),
            "ug_obj_parse": (C.c_int, [C.c_char_p, i, i, f, i, i]), "ug_material_parse": (C.c_int, [C.c_char_p, i, f]),
            "ug_model_init_orig_list": (C.c_int, [vp, C.c_int, C.c_int]), "ug_model_set_reflect": (C.c_int, [vp, f, C.c_int]), "ug_model_rotate_bunny": (C.c_int, [vp, C.c_float]),
            "ug_set_screen": (C.c_int, [vp, C.c_int, C.c_int, C.c_float]), "ug_set_strip": (C.c_int, [vp, C.c_int, C.c_int]), "ug_set_strip_bounds": (C.c_int, [vp, i, C.c_int, C.c_int]), "ug_set_slabs": (C.c_int, [vp, C.c_int]),
            "ug_update_light_position": (C.c_int, [vp, f, C.c_int]),
            "ug_fill_coordinates_data": (C.c_int, [vp, C.POINTER(Camera)]), "ug_build_frustum_grid": (C.c_int, [vp]),
            "ug_trace": (C.c_int, [vp]), "ug_shadow_pass": (C.c_int, [vp, C.c_int, C.POINTER(Camera)]),
            "ug_simple_shade": (C.c_int, [vp]), "ug_spotlight_shade": (C.c_int, [vp]), "ug_add_shadows": (C.c_int, [vp]),
            "ug_set_light_grid_mode": (C.c_int, [vp, C.c_int, C.c_float, C.c_float]), "ug_acosf_host": (None, [f, f, C.c_size_t]), "ug_shadow_item_shape": (None, [C.c_uint, C.c_uint, C.POINTER(C.c_uint)]),
            "ug_render_frame": (C.c_int, [vp, C.POINTER(FrameParams)]),
            "ug_render_frame_host": (C.c_int, [vp, C.POINTER(FrameParams), C.c_void_p]),
            "ug_render_frame_host_async": (C.c_int, [vp, C.POINTER(FrameParams), C.c_void_p]), "ug_wait_frame": (C.c_int, [vp, C.c_int]),
            "ug_get_stats": (C.c_int, [vp, C.POINTER(Stats)]), "ug_read": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
            "ug_device_ptr": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]),
            "ug_write_ppm": (C.c_int, [vp, C.c_char_p]), "ug_write_ppm_binary": (C.c_int, [vp, C.c_char_p]), "ug_write_png": (C.c_int, [vp, C.c_char_p]),
            "ug_model_tmp_model": (C.c_int, [vp, C.c_char_p, C.c_int]), "ug_model_update_vertices": (C.c_int, [vp, f, C.c_int]),
            "ug_model_bounds": (C.c_int, [vp, f]), "ug_obj_parse_vertices": (C.c_int, [C.c_char_p, i, f, C.c_int]), "ug_stage_times": (C.c_int, [vp, f, C.c_int]),
            "ug_camera_adjustCameraAndPosition": (None, [C.POINTER(Camera)]), "ug_camera_pack": (None, [C.POINTER(Camera), f]),
            "ug_direction_lattice": (None, [f, f]), "ug_fp32_issue_peak": (C.c_int, [vp, C.POINTER(C.c_double)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if lib().ug_comm_unique_id(C.cast(buf, C.c_void_p)) != UG_OK:
        raise RuntimeError("ug_comm_unique_id failed (libnccl not available?)")
    return bytes(buf.raw)


def make_camera(eye, look, up, near, far, fovy, aspect) -> Camera:
    """Camera setters + adjustCameraAndPosition/getGLMatrices/getFrustumProperties (camera.h), host only."""
    c = Camera()
    c.cx, c.cy, c.cz = eye
    c.lx, c.ly, c.lz = look
    c.ux, c.uy, c.uz = up
    c.nearPlane, c.farPlane, c.fovy, c.aspect = near, far, fovy, aspect
    lib().ug_camera_adjustCameraAndPosition(C.byref(c))
    return c


def camcoords(cam: Camera) -> np.ndarray:
    out = np.zeros(64, np.float32)
    lib().ug_camera_pack(C.byref(cam), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def direction_lattice(cc: np.ndarray) -> np.ndarray:
    cc = np.ascontiguousarray(cc, np.float32)
    out = np.zeros(100, np.float32)
    lib().ug_direction_lattice(cc.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def parse_obj(path: str):
    """Host-only OBJ/MTL parse through the library's loader boundary (no GPU needed)."""
    L = lib()
    nv, nf = C.c_int(0), C.c_int(0)
    if L.ug_obj_parse(path.encode(), C.byref(nv), C.byref(nf), None, None, None) != UG_OK:
        raise OSError(f"cannot parse {path}")
    V = np.zeros(3 * nv.value, np.float32)
    F = np.zeros(3 * nf.value, np.int32)
    M = np.zeros(nf.value, np.int32)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.ug_obj_parse(path.encode(), C.byref(nv), C.byref(nf), V.ctypes.data_as(fp), F.ctypes.data_as(ip), M.ctypes.data_as(ip))
    return V.reshape(-1, 3), F.reshape(-1, 3), M


def parse_materials(path: str) -> np.ndarray:
    L = lib()
    nm = C.c_int(0)
    if L.ug_material_parse(path.encode(), C.byref(nm), None) != UG_OK:
        raise OSError(f"cannot parse {path}")
    out = np.zeros(6 * nm.value, np.float32)
    L.ug_material_parse(path.encode(), C.byref(nm), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out.reshape(-1, 6)


class UgError(RuntimeError):
    pass


class Renderer:
    """One ug_context (one GPU).  Mirrors the reference's frame driver: load a Model once, then render frames."""

    def __init__(self, device: int = 0):
        self.L = lib()
        self.ctx = C.c_void_p()
        rc = self.L.ug_create(C.byref(self.ctx), device)
        if rc != UG_OK:
            raise UgError(f"ug_create(device={device}) failed with {rc}: a CUDA device is required, there is no CPU fallback")
        self.scene = None

    def close(self):
        if self.ctx:
            self.L.ug_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != UG_OK:
            raise UgError(f"{what} failed ({rc}): {self.L.ug_last_error(self.ctx).decode()}")

    # ---- Model
    def load(self, obj_path: str, material_file: Optional[str]):
        self._ck(self.L.ug_model_load(self.ctx, obj_path.encode(), material_file.encode() if material_file else None), "ug_model_load")

    def set_model(self, verts, faces, mat_idx, mat_list):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        faces = np.ascontiguousarray(faces, np.int32).reshape(-1)
        mat_idx = np.ascontiguousarray(mat_idx, np.int32)
        mat_list = np.ascontiguousarray(mat_list, np.float32).reshape(-1)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        self._ck(self.L.ug_model_set(self.ctx, verts.ctypes.data_as(fp), verts.size // 3, faces.ctypes.data_as(ip), faces.size // 3,
                                     mat_idx.ctypes.data_as(ip), mat_list.ctypes.data_as(fp), mat_list.size // 6), "ug_model_set")

    def set_scene(self, scene, mat_list=None):
        """Upload a scenes.Scene (arrays as Model::load_model would leave them)."""
        if mat_list is None:
            mat_list = np.asarray([[float("%.6f" % x) for x in (*m.ka, *m.kd)] for m in scene.materials], np.float32)
        self.set_model(scene.verts, scene.faces, scene.face_mat, mat_list)
        if scene.dyn_size:
            self._ck(self.L.ug_model_init_orig_list(self.ctx, scene.dyn_size, scene.dyn_offset), "ug_model_init_orig_list")
        if any(m.reflect > 0 for m in scene.materials):
            refl = np.asarray([float("%.6f" % m.reflect) for m in scene.materials], np.float32)
            self._ck(self.L.ug_model_set_reflect(self.ctx, refl.ctypes.data_as(C.POINTER(C.c_float)), refl.size), "ug_model_set_reflect")
        self.scene = scene

    # ---- keyframed scenes (Model::tmp_model, scene.h:70-120)
    def tmp_model(self, path: str, i: int):
        self._ck(self.L.ug_model_tmp_model(self.ctx, path.encode(), int(i)), "ug_model_tmp_model")

    def update_vertices(self, verts):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1)
        self._ck(self.L.ug_model_update_vertices(self.ctx, verts.ctypes.data_as(C.POINTER(C.c_float)), verts.size // 3), "ug_model_update_vertices")

    def bounds(self) -> np.ndarray:
        out = np.zeros(6, np.float32)
        self._ck(self.L.ug_model_bounds(self.ctx, out.ctypes.data_as(C.POINTER(C.c_float))), "ug_model_bounds")
        return out

    # ---- frames
    def params(self, scene=None, frame: int = 0, flags: int = 0, strip=(0, 1), bounces: int = 0, slabs: int = 1) -> FrameParams:
        s = scene or self.scene
        p = FrameParams()
        p.width, p.height, p.fovy, p.nearPlane, p.farPlane = s.width, s.height, float(s.fovy), s.near, s.far
        p.eye[:], p.look[:], p.up[:] = s.eye, s.look, s.up
        p.num_lights = len(s.lights)
        for k, l in enumerate(s.lights):
            p.lights[k].pos[:], p.lights[k].look[:], p.lights[k].up[:] = l.pos, l.look, l.up
        p.transform = 1 if s.dyn_size else 0
        p.rot_factor = float(np.float32(np.float32(s.rot_step) * np.float32(frame + 1)))
        p.strip_index, p.strip_count = strip
        p.flags = flags
        p.bounces = bounces
        p.num_slabs = slabs
        return p

    def render(self, p: FrameParams):
        self._last_lights = int(p.num_lights)
        self._ck(self.L.ug_render_frame(self.ctx, C.byref(p)), "ug_render_frame")

    def render_host(self, p: FrameParams, out: np.ndarray):
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
        self._ck(self.L.ug_render_frame_host(self.ctx, C.byref(p), out.ctypes.data_as(C.c_void_p)), "ug_render_frame_host")

    def render_host_async(self, p: FrameParams, out: np.ndarray):
        """`out` must be pinned; valid after wait_frame()."""
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
        self._ck(self.L.ug_render_frame_host_async(self.ctx, C.byref(p), out.ctypes.data_as(C.c_void_p)), "ug_render_frame_host_async")

    def wait_frame(self, frames_back: int = 0):
        self._ck(self.L.ug_wait_frame(self.ctx, int(frames_back)), "ug_wait_frame")

    def set_strip_bounds(self, row_bounds, index: int):
        """Strips of unequal height (cell-row boundaries 0 ... height/8); frames rendered with strip=(index, len(row_bounds)-1) keep them."""
        b = np.ascontiguousarray(row_bounds, np.int32)
        self._ck(self.L.ug_set_strip_bounds(self.ctx, b.ctypes.data_as(C.POINTER(C.c_int)), b.size - 1, int(index)), "ug_set_strip_bounds")

    def set_screen(self, width: int, height: int, fovy: float):
        self._ck(self.L.ug_set_screen(self.ctx, width, height, float(fovy)), "ug_set_screen")

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.ug_set_stream(self.ctx, C.c_void_p(cuda_stream)), "ug_set_stream")

    def set_rgb_buffer(self, device_ptr: int, nbytes: int):
        self._ck(self.L.ug_set_rgb_buffer(self.ctx, C.c_void_p(device_ptr), nbytes), "ug_set_rgb_buffer")

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        """Join the NCCL communicator of the strip gather (unique_id: 128 bytes from unique_id() on one rank)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.L.ug_comm_init(self.ctx, rank, world, C.cast(buf, C.c_void_p)), "ug_comm_init")

    def comm_flush(self):
        """Collective: completes the last FLAG_GATHER_DEFERRED frame on rank 0 and synchronises."""
        self._ck(self.L.ug_comm_flush(self.ctx), "ug_comm_flush")

    def synchronize(self):
        self._ck(self.L.ug_synchronize(self.ctx), "ug_synchronize")

    def stats(self) -> Stats:
        st = Stats()
        self._ck(self.L.ug_get_stats(self.ctx, C.byref(st)), "ug_get_stats")
        return st

    def stage_times(self):
        """{stage name: ms} of the last FLAG_STAGE_EVENTS frame."""
        buf = (C.c_float * 40)()
        n = self.L.ug_stage_times(self.ctx, buf, 40)
        names = ["transform", "cam_grid", "trace"]
        for l in range(getattr(self, "_last_lights", 0)):
            names += [f"light{l}_grid", f"light{l}_regroup", f"light{l}_shadow"]
        names += ["shade", "reflect_grid", "reflect_rays", "reflect_combine"]
        return {names[k]: buf[k] for k in range(min(max(n, 0), len(names)))}

    def fp32_issue_peak(self) -> float:
        """Measured FMUL + FADD lane-ops per second of this device (k_peak.cu)."""
        v = C.c_double(0.0)
        self._ck(self.L.ug_fp32_issue_peak(self.ctx, C.byref(v)), "ug_fp32_issue_peak")
        return float(v.value)

    def read(self, buffer: int, index: int = 0) -> np.ndarray:
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._ck(self.L.ug_device_ptr(self.ctx, buffer, index, C.byref(ptr), C.byref(nbytes)), f"ug_device_ptr({buffer})")
        dt = np.dtype(_BUF_DTYPE[buffer])
        out = np.zeros(nbytes.value // dt.itemsize, dt)
        if out.size:
            self._ck(self.L.ug_read(self.ctx, buffer, index, out.ctypes.data_as(C.c_void_p), out.nbytes), f"ug_read({buffer})")
        return out

    def device_ptr(self, buffer: int, index: int = 0):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._ck(self.L.ug_device_ptr(self.ctx, buffer, index, C.byref(ptr), C.byref(nbytes)), f"ug_device_ptr({buffer})")
        return ptr.value, nbytes.value

    def write_ppm_binary(self, path: str):
        self._ck(self.L.ug_write_ppm_binary(self.ctx, path.encode()), "ug_write_ppm_binary")

    def write_png(self, path: str):
        self._ck(self.L.ug_write_png(self.ctx, path.encode()), "ug_write_png")

    def write_ppm(self, path: str):
        self._ck(self.L.ug_write_ppm(self.ctx, path.encode()), "ug_write_ppm")

    # ---- the reference's call-by-call sequence (display(), main.cu:59-302)
    def render_stepwise(self, scene=None, spherical: bool = False, spot: bool = False, slabs: int = 1):
        """The per_frame_funcs.h call sequence of display(), one C-ABI call per reference function; `spherical` selects the
        getEffectiveRayGridMapping / buildSphericalGrid variant with x_max = y_max = M_PI (main.cu:186-192)."""
        s = scene or self.scene
        L = self.L
        pi = float(np.float32(np.pi))
        self._ck(L.ug_set_light_grid_mode(self.ctx, LIGHT_GRID_SPHERICAL if spherical else LIGHT_GRID_PERSPECTIVE, pi, pi), "ug_set_light_grid_mode")
        self._ck(L.ug_set_screen(self.ctx, s.width, s.height, float(s.fovy)), "ug_set_screen")
        self._ck(L.ug_set_slabs(self.ctx, slabs), "ug_set_slabs")
        self._ck(L.ug_set_strip(self.ctx, 0, 1), "ug_set_strip")
        lights = np.asarray([l.pos for l in s.lights], np.float32).reshape(-1)
        if len(s.lights):
            self._ck(L.ug_update_light_position(self.ctx, lights.ctypes.data_as(C.POINTER(C.c_float)), len(s.lights)), "updateLightPosition")
        cam = make_camera(s.eye, s.look, s.up, s.near, s.far, float(s.fovy), float(np.float32(s.width) / np.float32(s.height)))
        self._ck(L.ug_fill_coordinates_data(self.ctx, C.byref(cam)), "fillCoordinatesData")
        self._ck(L.ug_build_frustum_grid(self.ctx), "build_frustum_grid")
        self._ck(L.ug_trace(self.ctx), "trace")
        for k, l in enumerate(s.lights):
            lc = make_camera(l.pos, l.look, l.up, s.near, s.far, float(s.fovy), 1.0)
            self._ck(L.ug_shadow_pass(self.ctx, k, C.byref(lc)), "shadow pass")
        if spot:
            self._ck(L.ug_spotlight_shade(self.ctx), "spotlight_shade")
        else:
            self._ck(L.ug_simple_shade(self.ctx), "simpleShade")
        self._ck(L.ug_add_shadows(self.ctx), "add_shadows")
