#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ptb|reference] [--config c2|c1|c3|c5]
                    [--scaling strong|weak]

Metric (BASELINE.json): Mrays/s (primary + bounce + shadow rays, counted by the extend/shadow kernels)
on the named configuration, and 1080p@64spp frames/s (`config.frames_per_s_1080p64`, measured).

A "step" is one pass of the hot path over one frame of synthetic input: configs[1] of BASELINE.json — the
synthetic 999 698-triangle procedural heightfield, 1920x1080, 64 spp, depth 4 (reference library default),
integrator = renderer::trace semantics.
  * N = 1: that frame on one GPU;
  * N > 1 (torchrun, one rank per GPU): THE SAME frame cut into tiles that the N ranks steal from one
    shared-memory counter → "scaling": "strong" (fixed total work; what "1080p@64spp frames/s at 1/2/4/8"
    asks).  The scene is built once on rank 0 and its flattened blob broadcast over NCCL.  There is no
    collective in the data path: every rank's accumulate kernel stores its pixels into rank 0's frame over
    NVLink, so the frame return is INSIDE the timed region of `value`.  The weak-scaling figure (64 spp per
    GPU), BASELINE's C4 (3840x2160, 1024 spp) and C5 (49 instances of the 1 M-triangle mesh, 1080p x 64 spp) ride along
    as extra objects `weak`, `c4` and `c5`; at N = 1 the Cornell-box configurations C1 and C3 do (`c1`, `c3`);
    `--scaling weak` swaps the roles.
`value` is timed with CUDA events on the library's streams (first tile start to last tile end, slowest rank)
with the scene resident in HBM, barrier + synchronize on both sides.  `e2e` is host wall-clock around the
public call (ptb_group_render_frame): request in, finished float frame copied into pinned host memory.

`--impl reference` times the reference's own CPU renderer (oracle/_ref, the unmodified library compiled
from /root/reference by oracle/Makefile; the plain-C port oracle/pt_oracle.c if that is absent) on the
box's host cores on a bounded sample of the same workload.  This file is the only place outside tests/
and __graft_entry__.smoke() that touches oracle/, and only for that leg and the `cpu_baseline` object.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (description, full_w, full_h, spp, depth, integrator)
    "c1": ("C1 cornell-box 256x256 16spp depth4 LIB", 256, 256, 16, 4, 0),
    "c2": ("C2 heightfield 999698 tris 1920x1080 64spp depth4 LIB", 1920, 1080, 64, 4, 0),
    "c3": ("C3 cornell-box 1920x1080 256spp depth16 APP_RR", 1920, 1080, 256, 16, 1),
    "c2sun": ("C2 heightfield 999698 tris + a sun (one shadow ray per shade event) 1920x1080 64spp depth4 LIB", 1920, 1080, 64, 4, 0),
    "c5": ("C5 49 instances x 999698-tri heightfield (49M instanced tris) 1920x1080 64spp depth4 LIB", 1920, 1080, 64, 4, 0),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_description(cfg_name, n_grid):
    import ptb200 as ptb
    from ptb200 import procedural as P
    if cfg_name == "c2":
        return P.heightfield_scene(n_grid)
    if cfg_name == "c2sun":
        d = P.heightfield_scene(n_grid)
        a = np.float32(0.9)  # the light's -z axis tilted 0.9 rad off the vertical: long shadows over the terrain
        basis = np.array([1, 0, 0, 0, np.cos(a), -np.sin(a), 0, np.sin(a), np.cos(a)], np.float32)
        d.sun = (basis, np.array([3, 3, 3], np.float32), 0.004732)  # sun_light.hpp:9-10 angular radius
        return d
    if cfg_name == "c5":
        return P.instanced_heightfield_scene(n_grid, 7)
    return ptb.load_gltf_description(P.cornell_gltf_path())


# ------------------------------------------------------------------ reference arm / CPU baseline ----

def cpu_reference(cfg_name, n_grid, depth, integrator, sample_seconds=15.0, steps=1, warmup=0):
    """Times the reference CPU renderer on a bounded sample of the workload.  → dict for `cpu_baseline`
    plus per-step values.  kind = "reference" (oracle/_ref) or "port" (oracle/pt_oracle.c)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reflib
    desc = build_description(cfg_name, n_grid)
    flat = reflib.FlatScene(desc.meshes, desc.surfaces, desc.instances, desc.materials, desc.camera, desc.sun,
                            desc.environment_factor, desc.transparent_background)
    cores = os.cpu_count() or 1
    full_w, full_h = CONFIGS[cfg_name][1], CONFIGS[cfg_name][2]
    t0 = time.perf_counter()
    if reflib.available():
        kind = "reference"
        scene = reflib.RefScene.from_flat(flat)
        threads = reflib.hardware_threads()
    else:
        import portlib
        kind = "port"
        scene = portlib.PortScene(flat)
        threads = cores
    build_s = time.perf_counter() - t0
    # rays per path of this workload (restated integrator that counts scene-level intersect calls)
    pw, ph = max(16, full_w // 8), max(9, full_h // 8)
    mode_count = 1 if integrator == 1 else (2 if kind == "reference" else 0)
    _, _, rays, _ = scene.render_linear(pw, ph, 2, depth, mode=mode_count, threads=threads)
    rays_per_path = rays / float(pw * ph * 2)
    # calibrate the sample so that one step is about sample_seconds of CPU wall time
    sw, sh = max(32, full_w // 4), max(18, full_h // 4)

    def one(spp):
        if kind == "reference" and integrator == 0:
            _, secs = scene.render_png(sw, sh, spp, depth, threads=0)  # the reference's own render()
        else:
            _, _, _, secs = scene.render_linear(sw, sh, spp, depth, mode=(1 if integrator == 1 else 0), threads=threads)
        return secs

    probe = one(1)
    spp = int(max(1, min(64, round(sample_seconds / max(probe, 1e-3)))))
    values, secs_all = [], []
    for i in range(warmup + steps):
        secs = one(spp)
        if i >= warmup:
            secs_all.append(secs)
            values.append(sw * sh * spp * rays_per_path / secs / 1e6)
    sample = (f"{sw}x{sh} x {spp} spp of the same scene and camera, depth {depth}, "
              f"{'renderer::render()' if (kind == 'reference' and integrator == 0) else 'trace_iter restated over the library' if kind == 'reference' else 'plain-C port'}"
              f" on {threads} threads; rays/path {rays_per_path:.3f} counted on a {pw}x{ph}x2 pass; "
              f"scene build {build_s:.1f} s not timed")
    return dict(value=float(np.mean(values)), unit="Mrays/s", cores=int(threads), kind=kind, sample=sample,
                ms_per_step=float(np.mean(secs_all) * 1e3), paths_per_s=float(sw * sh * spp / np.mean(secs_all)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    desc_name, full_w, full_h, spp, depth, integ = CONFIGS[args.config]
    # every step is a bounded sample; the whole run (steps + one warm-up) is kept to about 2.5 minutes of CPU time
    wu = min(args.warmup, 1)
    per_step = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + wu)))
    r = cpu_reference(args.config, args.n_grid, depth, integ, sample_seconds=per_step, steps=args.steps, warmup=wu)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc_name, "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ product arm ----

def _dist_setup():
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    return world, rank, torch.cuda.current_device()


def run_ptb(args):
    """Product arm.  The frame goes through libptb's frame driver (ptb_group: one process per GPU, tiles stolen from
    a shared-memory counter, every GPU's accumulate kernel storing into rank 0's frame over NVLink).  If that path
    cannot be set up on this box (no peer access / CUDA IPC), all ranks fall back together to the first scheduler
    (cluster.render_frame + NCCL reduce) and say so in `config`."""
    import torch
    import torch.distributed as dist
    import ptb200 as ptb
    from ptb200 import cluster

    world, rank, device_index = _dist_setup()
    dev = torch.device("cuda", device_index)
    if args.gather == "nccl":
        return run_ptb_legacy(args, world, rank, device_index, why="--gather nccl")

    desc_name, full_w, full_h, spp0, depth, integ = CONFIGS[args.config]
    if args.spp:
        spp0 = args.spp
    ptb.set_option("wave_paths", args.wave_paths)
    ptb.set_option("group_timeout_ms", 180000)
    ptb.set_option("frame_queue_depth", args.queue_depth)

    # ---- the scene: built ONCE on rank 0, its flattened HBM image broadcast over NCCL
    t0 = time.perf_counter()
    scene = ptb.Scene.create(build_description(args.config, args.n_grid), device_index) if rank == 0 else None
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    scene = cluster.replicate_scene(scene, device_index)
    t_replicate = time.perf_counter() - t0
    info = scene.info()

    ok = 1
    group = None
    try:
        group = cluster.make_group(device_index)
        group.render_frame(scene, 64, 32, 1, 1, output=ptb.OUT_NONE)  # maps the frame on every rank (peer / IPC)
    except ptb.PtbError as e:
        sys.stderr.write(f"bench.py rank {rank}: frame driver unavailable ({e}); falling back\n")
        ok = 0
    if world > 1:
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    if not ok:
        if group is not None:
            group.close()
        scene.close()
        return run_ptb_legacy(args, world, rank, device_index, why="peer access / CUDA IPC unavailable")

    pinned = torch.empty((full_h, full_w, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)
    tile = tuple(args.tile)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    leg = {"scene": scene, "depth": depth, "integ": integ}  # what the frames below render (the extra legs swap it)

    def frame(seed, spp, to_host, w=full_w, h=full_h):
        return group.render_frame(leg["scene"], w, h, spp, leg["depth"],
                                  out=(pinned.data_ptr() if (to_host and rank == 0) else None), seed=seed,
                                  integrator=leg["integ"], tile=tile, tiles_in_flight=args.streams,
                                  output=ptb.OUT_RGBA32F if to_host else ptb.OUT_NONE)

    def timed(spp, steps, warmup, to_host=False, w=full_w, h=full_h):
        """→ dict(value, ms_per_step, rays, paths, launches, ...) on rank 0 (None elsewhere): device-timed when the
        frame stays in HBM, host-timed through the public call when it is copied out (to_host)."""
        for i in range(warmup):
            frame(1000 + i, spp, to_host, w, h)
        ms_total = wall_total = 0.0
        rays = paths = launches = 0
        per_rank_tiles = None
        for i in range(steps):
            flush.fill_(i & 0xFF)  # L2 flush between timed iterations
            barrier()
            t0 = time.perf_counter()
            _, st = frame(1 + i, spp, to_host, w, h)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            barrier()
            wall_total += cluster.all_max([wall], dev)[0]
            if rank == 0:
                ms_total += st["gpu_seconds"] * 1e3  # CUDA events on the library's streams, slowest rank
                rays += st["rays"]
                paths += st["paths"]
                launches += st["kernel_launches"]
                per_rank_tiles = st["tiles_per_rank"]
        if rank != 0:
            return None
        t_ms = wall_total if to_host else ms_total
        return dict(value=rays / (t_ms * 1e-3) / 1e6, ms_per_step=t_ms / steps, wall_ms_per_step=wall_total / steps,
                    rays=rays, paths=paths, launches=launches, n_tiles=st["n_tiles"], tiles_per_rank=per_rank_tiles)

    weak = args.scaling == "weak"
    spp_main = spp0 * world if weak else spp0

    # ---- timed region: kernel throughput, the finished frame stays in HBM on rank 0 (the tile return is inside)
    sampler = ClockSampler(device_index)
    for i in range(args.warmup):
        frame(1000 + i, spp_main, True)
    if rank == 0:
        sampler.start()
    main = timed(spp_main, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    # ---- end to end: request in, finished float frame in pinned host memory on rank 0
    e2e = timed(spp_main, args.steps, 0, to_host=True)

    # ---- the other scaling mode and the multi-GPU configurations BASELINE.json names, as extra legs
    extra = {}
    if world > 1 and not args.no_extra_legs:
        other_spp = spp0 if weak else spp0 * world
        r = timed(other_spp, max(2, min(args.steps, 5)), 1)
        if rank == 0:
            extra["strong" if weak else "weak"] = {
                "value": r["value"], "unit": "Mrays/s", "ms_per_step": r["ms_per_step"], "spp_total": other_spp,
                "what": (f"the fixed {spp0}-spp frame over {world} GPUs" if weak
                         else f"{other_spp} spp in total ({spp0} per GPU): per-GPU work fixed")}
        if args.config == "c2":
            # C4: 3840x2160, 1024 spp, tile-sharded with work stealing; the frame return is part of the kernels
            c4_w, c4_h, c4_spp = 3840, 2160, args.c4_spp
            timed(4, 1, 0, w=c4_w, h=c4_h)  # maps the 4K frame on every rank (the buffers are sized by the frame itself)
            r = timed(c4_spp, 1, 0, w=c4_w, h=c4_h)
            if rank == 0:
                extra["c4"] = {"workload": f"C4 {c4_w}x{c4_h} {c4_spp}spp depth{depth} over {world} GPUs",
                               "value": r["value"], "unit": "Mrays/s", "s_per_frame": r["ms_per_step"] * 1e-3,
                               "wall_s_per_frame": r["wall_ms_per_step"] * 1e-3, "n_tiles": r["n_tiles"],
                               "tiles_per_rank": r["tiles_per_rank"], "paths": r["paths"]}

        if args.config == "c2" and not args.no_c5_leg:
            # C5 (BASELINE configs[4]): 49 instances of the C2 mesh + the lights, 1080p x 64 spp, scene replicated like
            # the main one (built once on rank 0, blob broadcast), the same frame driver
            t0 = time.perf_counter()
            scene5 = ptb.Scene.create(build_description("c5", args.n_grid), device_index) if rank == 0 else None
            scene5 = cluster.replicate_scene(scene5, device_index)
            t_scene5 = time.perf_counter() - t0
            leg["scene"] = scene5
            try:
                r = timed(spp0, max(2, min(args.steps, 5)), 2)
            finally:
                leg["scene"] = scene
            if rank == 0:
                extra["c5"] = {"workload": f"{CONFIGS['c5'][0]} over {world} GPUs", "value": r["value"], "unit": "Mrays/s",
                               "ms_per_step": r["ms_per_step"], "frames_per_s": 1e3 / r["ms_per_step"],
                               "instances": int(scene5.info()["n_instances"]), "scene_build_and_replicate_s": t_scene5,
                               "tiles_per_rank": r["tiles_per_rank"]}
            barrier()
            scene5.close()

    if world == 1 and args.config == "c2" and not args.no_extra_legs:
        # the reference's own CPU-runnable case and the indirect-lighting case (BASELINE configs[0] and [2]: the bundled
        # Cornell box; C3 with worker::trace_iter's Russian roulette at depth 16) ride along on one GPU
        for name in ("c1", "c3"):
            _, w_, h_, spp_, depth_, integ_ = CONFIGS[name]
            try:  # (an extra leg must not cost the line its headline)
                sc = ptb.Scene.create(build_description(name, args.n_grid), device_index)
                leg.update(scene=sc, depth=depth_, integ=integ_)
                try:
                    r = timed(spp_, 3, 2, w=w_, h=h_)
                finally:
                    leg.update(scene=scene, depth=depth, integ=integ)
                extra[name] = {"workload": CONFIGS[name][0], "value": r["value"], "unit": "Mrays/s",
                               "ms_per_step": r["ms_per_step"], "rays_per_path": r["rays"] / max(1, r["paths"])}
                sc.close()
            except Exception as e:  # noqa: BLE001
                extra[name] = {"workload": CONFIGS[name][0], "error": str(e)[:200]}

    # ---- roofline of the dominant kernel (extend): separate, untimed-for-`value` passes on rank 0
    hbm_peak, peak_src = load_peaks()
    roofline = None
    n_tiles_main = main["n_tiles"] if rank == 0 else 0
    if rank == 0:
        # the kernel's roofline is a property of the kernel on this workload, not of how many GPUs share the frame:
        # always the one-GPU tiling (large waves), every world-th tile when the frame is shared
        tiles = ptb.frame_tiles(full_w, full_h, spp_main, 1, tile)
        sample_tiles = tiles[:: max(1, world)]
        buf = torch.zeros(max(t[2] * t[3] for t in tiles) * 4, dtype=torch.float32, device=dev)

        def one_tile(t):
            return scene.render_tile_dev(buf.data_ptr(), full_w, full_h, spp_main, depth, tile=t, seed=1, integrator=integ,
                                         stream=torch.cuda.current_stream().cuda_stream, want_stats=True)
        ptb.set_option("time_stages", 1)
        st_t = [one_tile(t) for t in sample_tiles]
        ptb.set_option("time_stages", 0)
        ptb.set_option("count_visits", 1)
        st_c = [one_tile(t) for t in sample_tiles]
        ptb.set_option("count_visits", 0)
        ext_s = sum(s["extend_seconds"] for s in st_t)
        shade_s = sum(s["shade_seconds"] for s in st_t)
        ext_launches = sum(s["extend_launches"] for s in st_t)
        n_rays = sum(s["rays"] for s in st_c)
        nb, nl, nt = (sum(s[k] for s in st_c) for k in ("node_visits", "leaf_visits", "tri_tests"))
        alg_bytes = 32 * n_rays + 8 * nb + 8 * nl + (4 + 48) * nt + 32 * n_rays
        achieved = alg_bytes / ext_s / 1e9
        roofline = {"bound": "hbm", "kernel": "extend_lanes_kernel", "achieved": achieved, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "bytes_per_ray": alg_bytes / max(n_rays, 1), "rays_per_launch": n_rays / max(ext_launches, 1),
                    "avg_launch_ms": ext_s / max(ext_launches, 1) * 1e3,
                    "extend_share_of_step": ext_s / max(ext_s + shade_s, 1e-12),
                    "visits_per_ray": {"branch": nb / max(n_rays, 1), "leaf": nl / max(n_rays, 1),
                                       "tri": nt / max(n_rays, 1)},
                    "extend_Mrays_per_s": n_rays / ext_s / 1e6,
                    "how": "CUDA events around every extend launch of the frame's tiles rendered one at a time on rank 0"}
        prof = os.path.join(ROOT, "profiles", "extend_traffic.json")
        if os.path.exists(prof):
            try:
                # measured DRAM bytes per ray of the ncu capture x the rays one launch processes here
                tj = json.load(open(prof))
                roofline["traffic"] = tj["dram_bytes_per_ray"] * roofline["rays_per_launch"]
                roofline["traffic_source"] = ("profiles/extend_traffic.json (" + str(tj.get("source", "ncu --set full")) +
                                              "): dram bytes/ray x rays_per_launch")
                roofline["issue"] = {"slots_busy_pct": tj.get("issue_slots_busy_pct_per_launch"),
                                     "active_lanes_per_instruction": tj.get("active_lanes_per_instruction_per_launch"),
                                     "long_scoreboard_cycles_per_instruction":
                                         tj.get("stall_cycles_per_instruction_long_scoreboard_per_launch"),
                                     "source": tj.get("summary", "profiles/")}
            except Exception:
                pass

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = cpu_reference(args.config, args.n_grid, depth, integ, sample_seconds=args.cpu_seconds)
        cpu = {"value": c["value"], "unit": "Mrays/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]}

    if rank == 0:
        frames_per_s = 1e3 / main["ms_per_step"] * (world if weak else 1)
        tl = ptb.frame_tile_layout(full_w, full_h, spp_main, world, tile, tiles_in_flight=args.streams).tolist()
        if tl[0][4]:  # comb tiles: granules of gx x gy pixels spread over the whole frame, all tiles equal
            tile_desc = (f"{tl[0][2]}x{tl[0][3]} pixels each, as {tl[0][4]}x{tl[0][6]}-pixel granules "
                         f"{tl[0][5]}x{tl[0][7]} apart (comb tiles: every tile samples the whole frame)")
        else:
            tile_desc = f"{tl[0][2]}x{tl[0][3]}" + (f" first, {tl[-len(tl) // 3][2]}x{tl[-len(tl) // 3][3]} last (guided)"
                                                    if tl[0][2:4] != tl[-len(tl) // 3][2:4] else "")
        line = {
            "metric": "Mrays/s", "value": main["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc_name + (f", {spp_main} spp total ({spp0}/GPU)" if (weak and world > 1) else ""),
                       "triangles": int(info["n_triangles"]), "kd_nodes": int(info["n_kd_nodes"]),
                       "kd_leaf_refs": int(info["n_leaf_refs"]), "scene_bytes": int(info["device_bytes"]),
                       "kd_build_s": t_build, "scene_replicate_s": t_replicate if world > 1 else None,
                       "scene_replication": "built once on rank 0; flattened blob broadcast over NCCL" if world > 1 else None,
                       "tiles": f"{n_tiles_main} tiles of {tile_desc}, stolen from one shared counter, "
                                + (f"{args.streams} in flight per GPU" if args.streams else
                                   (f"{n_tiles_main // world} per GPU (library's choice of tiles in flight: 3..8 of ~5 M paths)"
                                    if tl[0][4] else "8 in flight per GPU")),
                       "tiles_per_rank": main["tiles_per_rank"],
                       "frame_return": "accumulate kernels store into rank 0's frame over NVLink (CUDA IPC), inside the timed region",
                       "parallelism": f"tiles x{world}", "wave_paths": args.wave_paths,
                       "frames_per_s_1080p64": (frames_per_s if args.config in ("c2", "c5") else None),
                       "frames_per_s_measured": not weak or world == 1,
                       "wall_ms_per_step": main["wall_ms_per_step"],
                       "rays_per_path": main["rays"] / max(main["paths"], 1),
                       "l2": "256 MiB buffer written between timed iterations; scene + path state exceed L2"},
            "e2e": {"value": e2e["value"], "unit": "Mrays/s", "h2d_bytes_per_step": int(ctypes.sizeof(ptb.FrameReq)),
                    "d2h_bytes_per_step": int(full_w * full_h * 16), "ms_per_step": e2e["ms_per_step"],
                    "frames_per_s": 1e3 / e2e["ms_per_step"],
                    "what": "ptb_group_render_frame: request in, float RGBA frame in pinned host memory on rank 0"},
            "gpu_launches": int(main["launches"]),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    group.barrier()
    group.close()
    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ptb_legacy(args, world, rank, device_index, why):
    """The first scheduler (round 1): tile claims through the torch.distributed store, one NCCL reduce as the gather.
    Weak scaling only.  Selected with --gather nccl, or automatically when the frame driver cannot map rank 0's frame."""
    import torch
    import torch.distributed as dist
    import ptb200 as ptb
    from ptb200 import cluster
    dev = torch.device("cuda", device_index)

    desc_name, full_w, full_h, spp0, depth, integ = CONFIGS[args.config]
    if args.spp:
        spp0 = args.spp
    spp = spp0 * world  # weak scaling: 64 spp per GPU
    t_build = time.perf_counter()
    desc = build_description(args.config, args.n_grid)
    scene = ptb.Scene.create(desc, device_index)
    info = scene.info()
    t_build = time.perf_counter() - t_build
    ptb.set_option("wave_paths", args.wave_paths)

    # tiles per GPU: as many as asked for, but a tile keeps at least ~2 M paths (a small frame is not shredded)
    paths_per_gpu = full_w * full_h * spp0
    tiles_per_gpu = max(1, min(args.tiles_per_gpu, max(8, paths_per_gpu // (2 << 20))))
    cols, rows = cluster.tile_grid_for(world, tiles_per_gpu)
    tiles = cluster.make_tiles(full_w, full_h, cols, rows)
    frame = torch.zeros((full_h, full_w, 4), dtype=torch.float32, device=dev)
    n_workers = max(1, args.streams or 6)
    main_stream = torch.cuda.current_stream()
    streams = [main_stream] + [torch.cuda.Stream(device=dev) for _ in range(n_workers - 1)]
    max_px = max(t[2] * t[3] for t in tiles)
    tile_bufs = [torch.zeros(max_px * 4, dtype=torch.float32, device=dev) for _ in range(n_workers)]
    pinned = torch.empty((full_h, full_w, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)
    stream = main_stream

    def render_tile(tile, out_frame, seed, worker=0, stats=True):
        x0, y0, w, h = tile
        torch.cuda.set_device(dev)  # worker threads start without a current device
        with torch.cuda.stream(streams[worker]):
            st = scene.render_tile_dev(tile_bufs[worker].data_ptr(), full_w, full_h, spp, depth, tile=tile, seed=seed,
                                       integrator=integ, stream=streams[worker].cuda_stream, want_stats=stats)
            out_frame[y0:y0 + h, x0:x0 + w, :] = tile_bufs[worker][: w * h * 4].view(h, w, 4)
        return st

    epoch = [0]

    def step(seed, gather):
        frame.zero_()
        epoch[0] += 1
        for s_ in streams[1:]:
            s_.wait_stream(main_stream)  # tiles start after whatever precedes the step on the main stream
        r = cluster.render_frame(full_w, full_h, tiles, lambda t, f, wk: render_tile(t, f, seed, wk), frame,
                                 epoch=epoch[0], gather=False, local_workers=n_workers)
        for s_ in streams[1:]:
            main_stream.wait_stream(s_)  # the step ends on the main stream (that is where the events are)
        if gather and world > 1:
            dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(args.warmup):
        step(1000 + i, gather=True)
    barrier()

    # ---- timed region: kernel throughput, results stay in HBM
    sampler = ClockSampler(device_index)
    if rank == 0:
        sampler.start()
    rays = paths = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_total = 0.0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed iterations
        barrier()
        ev0.record(stream)
        r = step(1 + i, gather=False)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        ms_total += cluster.all_max([ms], dev)[0]
        rays += r["rays"]
        paths += r["paths"]
    rays_all, paths_all = cluster.all_sum([rays, paths], dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = rays_all / (ms_total * 1e-3) / 1e6

    # ---- end to end: public host call, request in, finished frame in pinned host memory on rank 0
    e2e_ms = 0.0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        barrier()
        t0 = time.perf_counter()
        r = step(1 + i, gather=True)
        if rank == 0:
            pinned.copy_(frame, non_blocking=True)
        barrier()
        e2e_ms += cluster.all_max([(time.perf_counter() - t0) * 1e3], dev)[0]
    e2e_value = rays_all / (e2e_ms * 1e-3) / 1e6
    h2d = int(ctypes.sizeof(ptb.TileReq) * len(tiles))
    d2h = int(full_w * full_h * 16)
    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc_name + (f", {spp} spp total ({spp0}/GPU)" if world > 1 else ""),
                       "fallback": f"first scheduler (torch.distributed store + NCCL reduce): {why}",
                       "triangles": int(info["n_triangles"]), "kd_build_s": info["build_seconds"],
                       "tiles": f"{cols}x{rows} work-stolen, {n_workers} in flight per GPU",
                       "parallelism": f"tiles x{world}", "wave_paths": args.wave_paths,
                       "rays_per_path": rays_all / max(paths_all, 1),
                       "l2": "256 MiB buffer written between timed iterations; scene + path state exceed L2"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches_total(tiles, spp, depth, args.wave_paths, args.steps)),
            "clocks": clocks, "roofline": None, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def launches_total(tiles, spp, depth, wave_paths, steps):
    """Kernels of libptb launched inside the timed region: per wave 1 raygen + depth x (extend + shade)
    + 1 accumulate (same arithmetic as render.cu; ptb_render_stats.kernel_launches reports it per call)."""
    total = 0
    for (_, _, w, h) in tiles:
        padded = ((w + 7) // 8) * ((h + 3) // 4) * 32
        wave_samples = max(1, min(spp, wave_paths // padded))
        waves = (spp + wave_samples - 1) // wave_samples
        total += waves * (2 + 2 * depth)
    return total * steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ptb", choices=["ptb", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the SAME frame (64 spp) over N GPUs — BASELINE's frames/s; weak: 64 spp per GPU")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="peer: libptb's frame driver (tile stores into rank 0's frame over NVLink); nccl: round-1 scheduler")
    ap.add_argument("--tile", type=int, nargs=2, default=[0, 0], metavar=("W", "H"), help="tile size (0 0: library's choice)")
    ap.add_argument("--queue-depth", type=int, default=1, help="tiles queued per stream (1 or 2)")
    ap.add_argument("--c4-spp", type=int, default=1024, help="samples of the C4 leg (3840x2160) at N > 1")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the weak-scaling / C4 / C5 legs at N > 1")
    ap.add_argument("--no-c5-leg", action="store_true", help="skip the C5 (49-instance scene) leg at N > 1")
    ap.add_argument("--n-grid", type=int, default=707, help="heightfield grid (707 → 999 698 triangles)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU")
    ap.add_argument("--tiles-per-gpu", type=int, default=32,
                    help="tiles per GPU (work-stolen); more tiles = a shorter tail when ranks finish unevenly")
    ap.add_argument("--wave-paths", type=int, default=8 << 20)
    ap.add_argument("--streams", type=int, default=0, help="tiles in flight per GPU (host threads / CUDA streams); 0 = library default (8)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of one reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ptb(args)


if __name__ == "__main__":
    sys.exit(main())
