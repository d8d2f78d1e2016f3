"""C4-sized frames (3840x2160) through the frame driver on one GPU: comb tiles vs rectangles, tiles in flight.
usage: c4_probe.py [spp]   (not a bench)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200 as ptb
from ptb200 import cluster, procedural as P
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
scene = ptb.Scene.create(P.heightfield_scene(707), 0)
group = cluster.make_group(0)
W, H = 3840, 2160
CASES = [(0, (0, 0), 4), (0, (0, 0), 6), (0, (0, 0), 8), (2, (0, 0), 6), (2, (0, 0), 8), (0, (0, 0), 8), (2, (0, 0), 8)]
if os.environ.get('WAVE'): ptb.set_option('wave_paths', int(os.environ['WAVE']))
for comb, tile, k in CASES:
    ptb.set_option("frame_comb_tiles", comb)
    group.render_frame(scene, W, H, 4, 4, output=ptb.OUT_NONE, tile=tile, tiles_in_flight=k)
    t0 = time.perf_counter()
    _, st = group.render_frame(scene, W, H, spp, 4, output=ptb.OUT_NONE, tile=tile, tiles_in_flight=k, seed=3)
    wall = time.perf_counter() - t0
    print(f"comb={comb} tile={tile} in_flight={k}: {st['n_tiles']} tiles, {st['gpu_seconds']:.3f} s gpu, {wall:.3f} s wall, "
          f"{st['rays'] / st['gpu_seconds'] / 1e6:.0f} Mrays/s, {st['kernel_launches']} launches", flush=True)
