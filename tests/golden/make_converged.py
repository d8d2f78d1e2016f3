"""Mints the converged reference images (statistical image parity) from the UNMODIFIED reference library
(oracle/_ref/libptref.so): 16 384 spp each, as BASELINE.json's north_star states the bar.

    python tests/golden/make_converged.py [--threads T] [A] [B] [B16]

  cornell_converged_A.npz    renderer::trace (LIB/core/renderer.cpp:437-643), depth 4                (config C1)
  cornell_converged_B.npz    worker::trace_iter restated over the reference library, depth 8
  cornell_converged_B16.npz  the same at depth 16 with Russian roulette                              (config C3)
Each: linear running-mean radiance 64x64 + the per-pixel standard deviation of ONE sample, measured from
the spread of independent 64-spp batches.  About 15 minutes of CPU per file on 8 threads.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reflib  # noqa: E402

SPECS = {"A": (0, 4, 256), "B": (1, 8, 256), "B16": (1, 16, 256)}  # mode, depth, batches of 64 spp


def main():
    argv = sys.argv[1:]
    threads = 0
    if "--threads" in argv:
        i = argv.index("--threads")
        threads = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if a in SPECS]
    gltf = os.path.join(HERE, "scenes", "cornell-box", "cornell.gltf")
    ref = reflib.RefScene.from_gltf(gltf)
    for name in (args or list(SPECS)):
        mode, depth, batches = SPECS[name]
        acc = np.zeros((64, 64, 3), np.float64)
        acc2 = np.zeros((64, 64, 3), np.float64)
        total = 0.0
        for b in range(batches):
            rgb64, _, _, secs = ref.render_linear(64, 64, 64, depth, mode=mode, threads=threads)
            acc += rgb64
            acc2 += rgb64.astype(np.float64) ** 2
            total += secs
            if (b + 1) % 16 == 0:
                print(f"{name}: batch {b + 1}/{batches}, {total:.0f} s so far", flush=True)
        mean = acc / batches
        var_batch = np.maximum(acc2 / batches - mean ** 2, 0) * batches / (batches - 1)  # variance of a 64-spp mean
        np.savez_compressed(os.path.join(HERE, f"cornell_converged_{name}.npz"), mean=mean.astype(np.float32),
                            sigma_per_sample=np.sqrt(var_batch * 64).astype(np.float32),
                            spp=np.array(64 * batches), depth=np.array(depth), mode=np.array(mode))
        print(f"{name}: wrote {64 * batches} spp, depth {depth}, {total:.0f} s", flush=True)


if __name__ == "__main__":
    main()
