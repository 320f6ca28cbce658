"""Shared parity helpers: run one frame on the GPU through the C-ABI and on the CPU oracle, compare stage by stage.

Bars (BASELINE.json north_star): cell/triangle index lists, hit ids and shadow bits bit-exact; t to 1e-5
relative (we require bit-exact and report both); RGB within 1/255 on >= 99.9 % of pixels (we require exact).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ugrt  # noqa: E402,F401
import uniformgrid_raytracing_b200 as ug  # noqa: E402
from oracle import harness  # noqa: E402


def bits(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def gpu_frame(r: "ug.Renderer", scene, frame=0, flags=ug.FLAG_FULL_BUFFERS, strip=(0, 1), stepwise=False, bounces=0, slabs=1):
    """Render through the C-ABI and read every stage buffer back."""
    if stepwise:
        r.render_stepwise(scene, spherical=bool(flags & ug.FLAG_SPHERICAL_LIGHTS), spot=bool(flags & ug.FLAG_SPOT_SHADE), slabs=slabs)
    else:
        r.render(r.params(scene, frame=frame, flags=flags, strip=strip, bounces=bounces, slabs=slabs))
    out = {"verts": r.read(ug.BUF_VERTEXLIST).reshape(-1, 3), "cam_values": r.read(ug.BUF_CAM_VALUES), "cam_span": r.read(ug.BUF_CAM_SPAN),
           "cam_offset": r.read(ug.BUF_CAM_OFFSET), "t": r.read(ug.BUF_T), "id": r.read(ug.BUF_ID), "shadowed": r.read(ug.BUF_SHADOWED),
           "rgb": r.read(ug.BUF_RGB)}
    if strip == (0, 1):
        out["sizelist"] = r.read(ug.BUF_SIZELIST)
    if flags & ug.FLAG_FULL_BUFFERS or stepwise:
        out["normal"] = r.read(ug.BUF_NORMAL)
        out["dir"] = r.read(ug.BUF_DIR)
    for li in range(len(scene.lights)):
        out[f"light{li}_values"] = r.read(ug.BUF_LIGHT_VALUES, li)
        out[f"light{li}_span"] = r.read(ug.BUF_LIGHT_SPAN, li)
        out[f"light{li}_offset"] = r.read(ug.BUF_LIGHT_OFFSET, li)
        out[f"light{li}_cell"] = r.read(ug.BUF_LIGHT_CELL, li)
    return out


def compare(gpu: dict, cpu: dict, keys=None, verbose=True) -> dict:
    """Bit-exact comparison of every buffer present in both; returns {key: number of differing elements}."""
    bad = {}
    for k in (keys or gpu.keys()):
        if k not in cpu or k not in gpu:
            continue
        a, b = np.asarray(gpu[k]).reshape(-1), np.asarray(cpu[k]).reshape(-1)
        if a.shape != b.shape:
            bad[k] = f"shape {a.shape} vs {b.shape}"
        else:
            n = int(np.count_nonzero(bits(a) != bits(b)))
            if n:
                bad[k] = n
        if verbose:
            print(f"  {k:18s} {str(a.shape):14s} {'OK' if k not in bad else 'DIFF ' + str(bad[k])}")
    return bad


def rel_err_t(gpu_t, cpu_t) -> float:
    m = (cpu_t > 0) & (gpu_t > 0)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(gpu_t[m] - cpu_t[m]) / cpu_t[m]))
