"""ncu target: one 1080p x 4 spp wave of the C5 scene (49 instances of the C2 mesh); not a bench."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200 as ptb
from ptb200 import procedural as P
s = ptb.Scene.create(P.instanced_heightfield_scene(707, 7))
for i in range(2):
    rgb, a, st = s.render_tile(1920, 1080, 4, 4, seed=1 + i)
print("rays", st["rays"], "paths", st["paths"], "launches", st["kernel_launches"])
