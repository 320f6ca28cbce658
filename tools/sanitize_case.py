#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer: micro scene + soup with 2 lights, strips, reflection, overflow regrow."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ugrt  # noqa: F401
import uniformgrid_raytracing_b200 as ug
from uniformgrid_raytracing_b200 import scenes
r = ug.Renderer(0)
for s in (scenes.scene_tiny("ka"), scenes.scene_random(5, 3000, 128, 96, nlights=2)):
    s.materials[0].reflect = 0.5; s.bounces = 2
    r.set_scene(s)
    for strip in ((0, 1), (1, 3)):
        r.render(r.params(s, strip=strip, bounces=2))
        r.render(r.params(s, strip=strip, bounces=2, flags=ug.FLAG_FULL_BUFFERS | ug.FLAG_NO_GRAPH))
    out = np.zeros(3 * s.num_pixels, np.uint8)
    r.render_host(r.params(s), out)
    r.render_stepwise(s)
    print(s.name, "ok", int(out.sum()))
r.close()
