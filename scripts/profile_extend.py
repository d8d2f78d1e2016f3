"""Short, profiler-friendly run of the hot path on configs[1] geometry: one 1080p wave (4 spp, depth 4).
Used under ncu (launch list / --set full on extend_kernel); prints nothing that is a bench value."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200 as ptb
from ptb200 import procedural as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 707
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if os.environ.get('PTB_VARIANT'):
    ptb.set_option('extend_variant', int(os.environ['PTB_VARIANT']))
s = ptb.Scene.create(P.heightfield_scene(n))
for i in range(2):
    rgb, a, st = s.render_tile(1920, 1080, spp, 4, seed=1 + i)
print("rays", st["rays"], "paths", st["paths"], "launches", st["kernel_launches"])
