"""Quick extend timing on configs[1] geometry (time_stages); not a bench.
usage: quick_extend.py [option=v1,v2,...]   e.g. extend_setup_lanes=4,8,12,16"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200 as ptb
from ptb200 import procedural as P
s = ptb.Scene.create(P.heightfield_scene(int(os.environ.get('PTB_N', '707'))))
ptb.set_option("time_stages", 1)
W, H, SPP = (int(os.environ.get(k, d)) for k, d in (('PTB_W', '1920'), ('PTB_H', '1080'), ('PTB_SPP', '8')))
TILE = tuple(int(v) for v in os.environ['PTB_TILE'].split(',')) if os.environ.get('PTB_TILE') else None
sweep = [(None, None)]
if len(sys.argv) > 1:
    name, vals = sys.argv[1].split("=")
    sweep = [(name, int(v)) for v in vals.split(",")]
for name, v in sweep:
    if name:
        ptb.set_option(name, v)
    best = None
    for rep in range(4):
        rgb, a, st = s.render_tile(W, H, SPP, 4, tile=TILE, seed=1)
        if best is None or st["extend_seconds"] < best["extend_seconds"]:
            best = st
    print(f"{name}={v} regs {ptb.lib().ptb_extend_registers()} extend {best['extend_seconds']*1e3:.2f} ms "
          f"{best['rays']/best['extend_seconds']/1e6:.1f} Mrays/s total {best['gpu_seconds']*1e3:.2f} ms "
          f"rays/path {best['rays']/best['paths']:.3f}")
