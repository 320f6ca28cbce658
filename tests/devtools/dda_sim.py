"""Development tool (lives under tests/ because it drives the oracle, which only tests, smoke() and the bench's CPU leg may do): reflection-ray work statistics (cells stepped / triangles tested per ray) on the CPU for
different uniform-grid resolutions.  Uses the oracle only to obtain the primary hits of the scene."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ugrt  # noqa
from oracle import harness
from uniformgrid_raytracing_b200 import scenes

def rays_of(scene):
    cache = f"build/rays_{scene.name}_{scene.width}.npz"
    if os.path.exists(cache):
        z = np.load(cache); return z["O"], z["r"]
    lib = harness.load("oracle")
    out = harness.run_frame(lib, scene, frame=0, threads=16)
    t, d, tid = out["t"], out["dir"].reshape(-1, 3), out["id"]
    k = np.array([m.reflect for m in scene.materials], np.float32)
    V = out["verts"].reshape(-1, 3); F = scene.faces
    ok = (tid >= 0) & (t > 0)
    ok &= k[scene.face_mat[np.clip(tid, 0, None)]] > 0
    idx = np.nonzero(ok)[0]
    eye = np.asarray(scene.eye, np.float32)
    P = eye + t[idx, None] * d[idx]
    tri = tid[idx]
    e1 = V[F[tri, 1]] - V[F[tri, 0]]; e2 = V[F[tri, 2]] - V[F[tri, 0]]
    n = np.cross(e1, e2); n /= np.linalg.norm(n, axis=1, keepdims=True)
    flip = (n * d[idx]).sum(1) > 0; n[flip] *= -1
    O = (P + 1e-3 * n).astype(np.float32)
    r = (d[idx] - 2 * (d[idx] * n).sum(1, keepdims=True) * n).astype(np.float32)
    np.savez(cache, O=O, r=r)
    return O, r

def sim(scene, O, r, n):
    L = C.CDLL("build/libddasim.so")
    L.dda_sim.restype = C.c_long
    C.c_int.in_dll(L, "g_cull").value = int(os.environ.get("CULL", "0"))
    V = np.ascontiguousarray(scene.verts, np.float32); F = np.ascontiguousarray(scene.faces, np.int32)
    R = len(O)
    steps = np.zeros(R, np.int32); tests = np.zeros(R, np.int32); ne = np.zeros(R, np.int32)
    nn = np.asarray(n, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    P = L.dda_sim(p(V), len(V), p(F), len(F), p(np.ascontiguousarray(O)), p(np.ascontiguousarray(r)), R, p(nn), p(steps), p(tests), p(ne))
    sk = (C.c_long * 4).in_dll(L, "g_skip_steps"); bl = (C.c_long * 4).in_dll(L, "g_skip_blocks")
    global SKIP
    SKIP = {1 << k: (sk[k] / max(R, 1), bl[k] / max(R, 1)) for k in (1, 2, 3)}
    return P, steps & 0xffff, tests, ne, steps >> 16

def res_for(scene, lam):
    lo = scene.verts.min(0); hi = scene.verts.max(0); ext = (hi - lo) * 1.002
    target = lam * scene.num_tris
    h = (ext.prod() / target) ** (1 / 3)
    return np.maximum(1, np.ceil(ext / h)).astype(int)

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "B"
    s = scenes.SCENES[name]()
    O, r = rays_of(s)
    print(name, "rays", len(O), "tris", s.num_tris)
    for lam in [float(x) for x in (sys.argv[2:] or ["1", "4", "16", "64"])]:
        n = res_for(s, lam)
        P, st, te, ne, nc = sim(s, O, r, n)
        # crude cost model in thread-instructions: 30/step + 50/test
        cost = 30 * st + 50 * te + 30 * ne + 60 * nc
        # warp cost: max over each group of 32 consecutive rays
        R = len(O) // 32 * 32
        wc = cost[:R].reshape(-1, 32).max(1).sum() * 32
        print("   empty-space skipping, coarse block edge -> (fine steps inside empty blocks, empty blocks entered) per ray -> steps of a two-level walk:",
              {b: (round(v[0], 1), round(v[1], 1), round(st.mean() - v[0] + v[1], 1)) for b, v in SKIP.items()})
        print(f"lam={lam:5.1f} n={n} pairs={P} steps/ray={st.mean():7.1f} nonemptycells={nc.mean():5.1f} boxtests={ne.mean():6.1f} tests/ray={te.mean():7.1f} p99tests={np.percentile(te,99):.0f} cost/ray={cost.mean():8.0f} warpcost/ray={wc/R:8.0f}")
