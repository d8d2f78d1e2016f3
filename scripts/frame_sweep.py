#!/usr/bin/env python
"""Strong-scaling sweep of the frame driver (ptb_group) under torchrun: one set-up, many (tile size, tiles in
flight, queue depth) settings on the C2 frame; prints one line per setting and the NVLink traffic of the frame
return (nvidia-smi nvlink counters of GPU 0 before / after).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/frame_sweep.py
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nvlink_bytes(index):
    """Sum of the NVLink data counters (KiB → bytes) of one GPU: (rx, tx), or None when unavailable."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
    except Exception:
        return None
    rx = tx = 0
    seen = False
    for line in out.splitlines():
        line = line.strip()
        try:
            if "Data Rx:" in line:
                rx += int(line.split("Data Rx:")[1].split()[0]) * 1024
                seen = True
            elif "Data Tx:" in line:
                tx += int(line.split("Data Tx:")[1].split()[0]) * 1024
                seen = True
        except ValueError:  # "N/A": the counters are not exposed in this container
            return None
    return (rx, tx) if seen else None


def main():
    import torch
    import torch.distributed as dist
    import ptb200 as ptb
    from ptb200 import cluster, procedural as P
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n_grid = int(os.environ.get("SWEEP_NGRID", "707"))
    scene = ptb.Scene.create(P.heightfield_scene(n_grid), local) if rank == 0 else None
    scene = cluster.replicate_scene(scene, local)
    group = cluster.make_group(local)
    W, H, spp, depth = int(os.environ.get("SWEEP_W", "1920")), int(os.environ.get("SWEEP_H", "1080")), 64, 4
    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    steps = int(os.environ.get("SWEEP_STEPS", "6"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(tile, streams, depth_q, to_host, spin=0, rpl=None):
        ptb.set_option("frame_queue_depth", depth_q)
        ptb.set_option("frame_spin_wait", spin)
        if rpl is not None:
            ptb.set_option("extend_rays_per_lane", rpl)
        out = pinned.data_ptr() if (to_host and rank == 0) else None
        kw = dict(seed=1, tile=tile, tiles_in_flight=streams, output=ptb.OUT_RGBA32F if to_host else ptb.OUT_NONE)
        for i in range(2):
            group.render_frame(scene, W, H, spp, depth, out=out, **kw)
        gpu = wall = 0.0
        rays = 0
        st = None
        for i in range(steps):
            barrier()
            t0 = time.perf_counter()
            _, st = group.render_frame(scene, W, H, spp, depth, out=out, **kw)
            torch.cuda.synchronize()
            w = (time.perf_counter() - t0) * 1e3
            barrier()
            wall += cluster.all_max([w], dev)[0]
            if rank == 0:
                gpu += st["gpu_seconds"] * 1e3
                rays += st["rays"]
        if rank == 0:
            per = np.array(st["gpu_seconds_per_rank"]) * 1e3
            print(json.dumps(dict(world=world, tile=list(tile), n_tiles=st["n_tiles"], streams=streams, queue_depth=depth_q,
                                  spin=spin, rays_per_lane=rpl, to_host=to_host, gpu_ms=round(gpu / steps, 3), wall_ms=round(wall / steps, 3),
                                  mrays_s=round(rays / (gpu * 1e-3) / 1e6, 1), frames_s_wall=round(1e3 / (wall / steps), 2),
                                  tiles_per_rank=st["tiles_per_rank"], last_gpu_ms_per_rank=[round(float(x), 2) for x in per])),
                  flush=True)

    if os.environ.get("SWEEP_SETTINGS"):
        # explicit list: [[tile_w, tile_h, streams, queue_depth, spin, rays_per_lane], ...]
        #            or {"tile": [w, h], "streams": n, "opts": {"frame_comb_tiles": 0, ...}, "host": false}
        for item in json.loads(os.environ["SWEEP_SETTINGS"]):
            if isinstance(item, dict):
                for k, v in item.get("opts", {}).items():
                    ptb.set_option(k, v)
                if rank == 0:
                    print(json.dumps({"opts": item.get("opts", {})}), flush=True)
                run(tuple(item.get("tile", (0, 0))), item.get("streams", 0), 1, bool(item.get("host", False)))
                continue
            tw, th, streams, dq, spin, rpl = item
            run((tw, th), streams, dq, False, spin=spin, rpl=rpl)
        group.barrier()
        group.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    settings = [((0, 0), 6, 1)]
    if os.environ.get("SWEEP_FULL", "1") != "0":
        for tile in ((0, 0), (128, 64), (192, 96), (256, 128), (96, 48), (320, 160)):
            for streams in (4, 6, 8):
                settings.append((tile, streams, 1))
        settings += [((0, 0), 3, 2), ((0, 0), 4, 2), ((0, 0), 6, 2), ((192, 96), 4, 2)]
    seen = set()
    for tile, streams, dq in settings:
        key = (tuple(ptb.frame_tiles(W, H, spp, world, tile)[0][2:]), streams, dq)
        if key in seen:
            continue
        seen.add(key)
        run(tile, streams, dq, False)
    run((0, 0), 6, 1, False, spin=1)
    run((0, 0), 4, 1, False, spin=1)
    run((0, 0), 6, 1, True)

    # NVLink traffic of the frame return: GPU 0's counters around 20 frames
    barrier()
    before = nvlink_bytes(0) if rank == 0 else None
    for i in range(20):
        group.render_frame(scene, W, H, spp, depth, seed=i, output=ptb.OUT_NONE)
    barrier()
    if rank == 0:
        after = nvlink_bytes(0)
        if before and after:
            rx = (after[0] - before[0]) / 20
            tx = (after[1] - before[1]) / 20
            print(json.dumps(dict(nvlink_gpu0_rx_bytes_per_frame=rx, nvlink_gpu0_tx_bytes_per_frame=tx,
                                  frame_bytes=W * H * 16, expected_rx=W * H * 16 * (world - 1) / world)), flush=True)
        else:
            print(json.dumps(dict(nvlink="counters unavailable")), flush=True)
    group.barrier()
    group.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
