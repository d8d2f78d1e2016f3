#!/usr/bin/env python
"""torchrun check of the process-per-GPU frame path (ptb_group), run by tests/test_gpu_frame.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/group_check.py

Rank 0 builds the scene; it is replicated over NCCL as one blob; every rank renders the tiles it steals straight
into rank 0's frame (CUDA IPC); rank 0 compares the frame with the single-GPU render, bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import ptb200 as ptb
    from ptb200 import cluster, procedural as P
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ptb.set_option("group_timeout_ms", 60000)
    scene = None
    t0 = time.perf_counter()
    if rank == 0:
        scene = ptb.Scene.create(P.heightfield_scene(120), local)
    scene = cluster.replicate_scene(scene, local)
    t_rep = time.perf_counter() - t0
    group = cluster.make_group(local)
    ok = True
    for (W, H, spp, depth, tile) in ((320, 200, 8, 4, (0, 0)), (320, 200, 8, 4, (64, 32)), (257, 131, 5, 6, (48, 24))):
        frame, st = group.render_frame(scene, W, H, spp, depth, seed=3, tile=tile)
        rgba8, _ = group.render_frame(scene, W, H, spp, depth, seed=3, tile=tile, output=ptb.OUT_RGBA8)
        if rank == 0:
            rgb, alpha, st1 = scene.render_tile(W, H, spp, depth, seed=3)
            same = (np.array_equal(frame[..., :3].view(np.uint32), rgb.view(np.uint32)) and
                    np.array_equal(frame[..., 3].view(np.uint32), alpha.view(np.uint32)) and st["rays"] == st1["rays"] and
                    np.array_equal(rgba8, ptb.tonemap_rgba8(rgb, alpha)))
            ok = ok and same
            print(f"{W}x{H} tile {tile}: {st['n_tiles']} tiles over {world} ranks {st['tiles_per_rank']}, "
                  f"{st['gpu_seconds'] * 1e3:.2f} ms GPU, identical: {same}", flush=True)
    # every replica answers rays like the built scene
    rng = np.random.default_rng(1)
    od = np.concatenate([rng.uniform(-5, 5, (5000, 3)) * (1, 0.3, 1) + (0, 2, 0), rng.normal(size=(5000, 3))], 1).astype(np.float32)
    h = torch.from_numpy(scene.trace_rays(od).view(np.uint8).reshape(len(od), -1).copy()).cuda()
    gathered = [torch.empty_like(h) for _ in range(world)]
    dist.all_gather(gathered, h)
    if rank == 0:
        same = all(torch.equal(g, gathered[0]) for g in gathered)
        ok = ok and same
        print(f"replicas agree on 5000 rays: {same}; replication took {t_rep:.2f} s", flush=True)
        print(f"group frame == single-GPU frame: {ok}", flush=True)
    group.barrier()
    group.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
