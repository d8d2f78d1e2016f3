"""Two or more GPUs (torchrun): geometry-sharded closest hit with the merge done over peer memory
(cluster.ShardMergeContext) against the unsharded search on rank 0's own GPU — must be bit-identical.
Also times the exchange.  usage: torchrun --nproc-per-node N scripts/shard_merge_check.py"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ptb200 as ptb  # noqa: E402
from ptb200 import procedural  # noqa: E402

cluster = importlib.import_module("distributed-path-tracer_b200.cluster")


def scene_description():
    rng = np.random.default_rng(21)
    a, b = procedural.heightfield_mesh(200, 1.0, 3), procedural.heightfield_mesh(4, 0.7, 4)
    insts = []
    for i in range(37):
        ang = rng.uniform(0, 6.28)
        sc = rng.uniform(0.4, 1.6, 3)
        c, s_ = np.cos(ang), np.sin(ang)
        basis = np.array([c * sc[0], 0, -s_ * sc[0], 0, sc[1], 0, s_ * sc[2], 0, c * sc[2]], np.float32)
        insts.append((rng.uniform(-5, 5, 3) * (1, 0.2, 1), basis, i % 2, 1))
    insts.append(((0, 0.1, 0), np.eye(3, dtype=np.float32).ravel(), 0, 2))
    insts.append(insts[4])  # exact duplicate: distance ties must go to the lower scene index
    mats = [dict(albedo=(0.7, 0.7, 0.7), roughness=1.0, metallic=0.0)] * 2
    cam = procedural.look_at((0, 6, 14), (0, 0, 0))
    return ptb.SceneDescription([a, b], [(0, 0), (1, 1)], insts, mats, (cam[0], cam[1], 0.8))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    desc = scene_description()
    n = 2_000_000
    rng = np.random.default_rng(5)
    o = (rng.uniform(-7, 7, (n, 3)) * (1, 0.4, 1)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    od = np.concatenate([o, d], 1)
    rays = torch.from_numpy(od).cuda()
    shard_desc, imap = cluster.shard_instances(desc, rank, world)
    shard = ptb.Scene.create(shard_desc, device=torch.cuda.current_device())
    ctx = cluster.ShardMergeContext(shard, imap, n)
    hits = ctx.trace(rays)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        hits = ctx.trace(rays, hits)
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    got = hits.cpu().numpy().view(ptb.HIT_DTYPE).reshape(-1)
    # shadow query: any-hit against the shard, OR-merged into every rank's buffer from inside the kernel
    occ = ctx.occlusion(rays)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        occ = ctx.occlusion(rays)
    torch.cuda.synchronize()
    dist.barrier()
    dt_occ = (time.perf_counter() - t0) / reps
    occ = occ.cpu().numpy().astype(bool)
    ok = True
    if rank == 0:
        with ptb.Scene.create(desc, device=torch.cuda.current_device()) as full:
            want = full.trace_rays(od)
            t1 = time.perf_counter()
            full.trace_rays(od)
            t_full = time.perf_counter() - t1
        for f in ("instance", "surface", "triangle"):
            ok &= bool(np.array_equal(got[f], want[f]))
        ok &= bool(np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)))
        ok &= bool(np.array_equal(got["bary"].view(np.uint32), want["bary"].view(np.uint32)))
        hit = want["instance"] != 0xFFFFFFFF
        ok_occ = bool(np.array_equal(occ, hit))
        print(f"world {world}: sharded+merged == unsharded: {ok}; {hit.mean():.3f} of {n} rays hit; "
              f"sharded trace + peer-memory merge {dt*1e3:.2f} ms per call ({n/dt/1e6:.0f} Mrays/s), "
              f"unsharded host-API call {t_full*1e3:.2f} ms", flush=True)
        print(f"world {world}: sharded shadow query == unsharded: {ok_occ}; any-hit trace + OR-merge over NVLink "
              f"{dt_occ*1e3:.2f} ms per call ({n/dt_occ/1e6:.0f} Mrays/s)", flush=True)
        ok &= ok_occ
    else:
        # every rank holds the merged answer: compare with rank 0's through an all-reduce below
        pass
    # all ranks must hold the SAME merged occlusion bytes
    occ_dev = torch.from_numpy(occ.astype(np.int32)).cuda()
    lo, hi = occ_dev.clone(), occ_dev.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok &= bool(torch.equal(lo, hi))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    shard.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
