#!/usr/bin/env python
"""Development aid: on ONE GPU, render every strip of an N-strip split of a config and print the per-strip stage times
(what the slowest rank of an N-GPU run would spend where).  python tools/strip_balance.py [--config D] [--strips 8]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugrt  # noqa: E402,F401
import uniformgrid_raytracing_b200 as ug  # noqa: E402
from uniformgrid_raytracing_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="D")
    ap.add_argument("--strips", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    scene = scenes.scene_a() if a.config == "A" else scenes.SCENES[a.config]()
    r = ug.Renderer(0)
    r.set_scene(scene)
    rows = []
    for n, idx in [(1, 0)] + [(a.strips, i) for i in range(a.strips)]:
        p = r.params(scene, strip=(idx, n))
        for _ in range(3):
            r.render(p)
        r.synchronize()
        ms = []
        for i in range(a.reps * 2):
            r.render(p)
            ms.append(r.stats().ms_frame)
        stage = {}
        for i in range(a.reps):
            r.render(r.params(scene, strip=(idx, n), flags=ug.FLAG_STAGE_EVENTS))
            for k, v in r.stage_times().items():
                stage[k] = stage.get(k, 0.0) + v / a.reps
        st = r.stats()
        row = {"strips": n, "strip": idx, "ms_frame_graph": round(sorted(ms)[len(ms) // 2], 4), "cam_pairs": int(st.cam_pairs),
               "stage_ms": {k: round(v, 4) for k, v in stage.items()}}
        rows.append(row)
        print(json.dumps(row), flush=True)
    full = rows[0]["ms_frame_graph"]
    worst = max(x["ms_frame_graph"] for x in rows[1:])
    print(json.dumps({"full_frame_ms": full, "slowest_strip_ms": worst, "speedup_bound": round(full / worst, 3),
                      "efficiency_bound": round(full / worst / a.strips, 3)}))
    r.close()


if __name__ == "__main__":
    main()
