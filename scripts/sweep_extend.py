"""Sweep of the lane kernel's (steps, tests) per main-loop iteration on configs[1] geometry.  Not a bench."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ptb200 as ptb
from ptb200 import procedural as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 707
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = ptb.Scene.create(P.heightfield_scene(n))
ptb.set_option("time_stages", 1)
ref = None
for variant, steps, tests in [(0, 3, 1), (1, 4, 2), (3, 4, 2)]:
    ptb.set_option("extend_variant", variant); ptb.set_option("extend_steps", steps); ptb.set_option("extend_tests", tests)
    best = None
    for rep in range(3):
        rgb, a, st = s.render_tile(1920, 1080, spp, 4, seed=1)
        if best is None or st["extend_seconds"] < best["extend_seconds"]:
            best = st
    if ref is None:
        ref = rgb.copy()
    same = np.array_equal(rgb.view(np.uint32), ref.view(np.uint32))
    print(f"variant={variant} steps={steps} tests={tests}: extend {best['extend_seconds']*1e3:8.2f} ms  "
          f"{best['rays']/best['extend_seconds']/1e6:8.1f} Mrays/s  shade+rest {best['shade_seconds']*1e3:6.2f} ms  image_identical={same}")
