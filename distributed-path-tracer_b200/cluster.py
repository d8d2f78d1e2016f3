"""Tile-sharded rendering across GPUs: one process per GPU, ``torch.distributed`` for the plumbing.

The product path is in libptb (``csrc/frame.cu``: ``ptb_group`` / ``ptb_ctx``): tiles stolen from a shared-memory
counter, every GPU's accumulate kernel storing into rank 0's frame over NVLink.  This module adds the two things
that need a process group — :func:`replicate_scene` (rank 0 builds, the flattened scene blob is broadcast over
NCCL) and :func:`make_group` (agreeing on the rendezvous name) — and keeps the first, renderer-agnostic
scheduler (:func:`render_frame`: tile claims through the ``torch.distributed`` store, ``reduce(SUM)`` as the
gather), which the CPU tests drive on gloo and ``bench.py --gather nccl`` can still select.

The reference distributes by GEOMETRY (every worker holds a subset of the primitives and every ray is
meant to visit all workers, ``src/processors/worker/intersection_worker.cpp:78-110``; the transport was
never written).  With 180 GB of HBM the scene is simply replicated and the IMAGE is sharded instead:

* every rank builds/loads the same scene (or rank 0 broadcasts the description: :func:`broadcast_description`);
* the frame is cut into tiles; ranks claim tiles from one shared counter (work stealing through the
  ``torch.distributed`` store's atomic ``add`` — a tile takes milliseconds to seconds, a claim ~0.1 ms);
* a tile never needs another rank's data, so the render itself has NO collective;
* the only exchange step is the framebuffer return: tiles are disjoint, so a ``reduce(SUM)`` of the
  zero-initialised per-rank frames onto rank 0 IS the gather (NCCL over NVLink; gloo on CPU in the tests).

The module is renderer-agnostic: ``render_tile_fn(tile, out_frame)`` does the work, which lets the CPU
test-suite drive the scheduling / merge logic with world_size 2 on gloo and a fake renderer.
"""
from __future__ import annotations

import threading
import time
from typing import Callable, List, Sequence, Tuple

import numpy as np

Tile = Tuple[int, int, int, int]  # x0, y0, w, h


def make_tiles(full_w: int, full_h: int, cols: int, rows: int) -> List[Tile]:
    """Row-major grid of cols x rows tiles covering the frame exactly (edges absorb the remainder)."""
    if cols < 1 or rows < 1 or cols > full_w or rows > full_h:
        raise ValueError("bad tile grid")
    xs = [full_w * i // cols for i in range(cols + 1)]
    ys = [full_h * j // rows for j in range(rows + 1)]
    return [(xs[i], ys[j], xs[i + 1] - xs[i], ys[j + 1] - ys[j]) for j in range(rows) for i in range(cols)]


def tile_grid_for(world_size: int, tiles_per_rank: int = 8) -> Tuple[int, int]:
    """cols x rows with cols*rows == tiles_per_rank * world_size, as square as powers of two allow."""
    n = max(1, tiles_per_rank * world_size)
    cols = 1
    while cols * cols < n:
        cols *= 2
    rows = max(1, n // cols)
    if cols * rows < n:
        rows += 1
    return cols, rows


class TileCounter:
    """Work-stealing tile queue: one shared counter, atomic fetch-add.

    With ``torch.distributed`` initialised the counter lives in the default process group's store
    (``store.add`` is atomic across ranks); without it (single process) it is a plain integer."""

    def __init__(self, n_tiles: int, key: str = "ptb_tile_counter", group_store=None):
        self.n = n_tiles
        self.key = key
        self.store = group_store
        self.local = 0
        self._lock = threading.Lock()

    def next(self) -> int:
        """Index of the next unclaimed tile, or -1 when the frame is exhausted (thread-safe)."""
        with self._lock:
            if self.store is None:
                i = self.local
                self.local += 1
            else:
                i = self.store.add(self.key, 1) - 1
        return i if i < self.n else -1


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _default_store():
    import torch.distributed as dist
    try:
        return dist.distributed_c10d._get_default_store()
    except Exception:  # pragma: no cover - very old / very new torch
        return None


def render_frame(full_w: int, full_h: int, tiles: Sequence[Tile],
                 render_tile_fn: Callable[[Tile, "object", int], dict | None], frame, *, epoch: int = 0,
                 static_assignment: bool = False, gather: bool = True, local_workers: int = 1,
                 sync_fn: Callable[[], None] | None = None):
    """Renders the tiles this rank claims into ``frame`` and gathers the full frame on rank 0.

    frame: a zero-initialised torch tensor [full_h, full_w, C] (device memory under NCCL, CPU under gloo);
           tiles this rank did not render stay zero.
    render_tile_fn(tile, frame, worker) renders one tile into frame[y0:y0+h, x0:x0+w] and may return a stats
           dict (its "rays" and "paths" entries are summed).  ``worker`` in [0, local_workers) identifies the
           host thread / CUDA stream: with local_workers > 1 several tiles of this rank are in flight at once
           (each on its own stream), which hides the tail of one tile's kernels behind the next tile's.
    epoch: distinguishes successive frames in the store (a new counter key per frame; the previous frame's key
           is deleted once this frame has been reduced).
    sync_fn: called before the gather: must make the CURRENT stream wait for whatever streams the workers rendered
           on (the reduce is issued on the current stream).  Required with gather and local_workers > 1 on a device.
    Returns dict(rays, paths, tiles=[indices this rank rendered], seconds_render, seconds_gather).
    """
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    store = _default_store() if (dist and not static_assignment) else None
    stealing = not (static_assignment or (dist and store is None))
    counter = TileCounter(len(tiles), key=f"ptb_tiles_{epoch}", group_store=store)
    mine = list(range(rank, len(tiles), world))  # round-robin when there is no stealing
    mine_pos = TileCounter(len(mine))
    done, totals, errors = [], {"rays": 0, "paths": 0}, []
    lock = threading.Lock()

    def work(worker: int):
        try:
            while True:
                if stealing:
                    i = counter.next()
                else:
                    j = mine_pos.next()
                    i = mine[j] if j >= 0 else -1
                if i < 0:
                    break
                st = render_tile_fn(tiles[i], frame, worker)
                with lock:
                    done.append(i)
                    if st:
                        totals["rays"] += int(st.get("rays", 0))
                        totals["paths"] += int(st.get("paths", 0))
        except BaseException as e:  # surfaced on the calling thread
            errors.append(e)

    t0 = time.perf_counter()
    if local_workers <= 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(w,)) for w in range(local_workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    t1 = time.perf_counter()
    if dist and gather and world > 1:
        if sync_fn is not None:
            sync_fn()
        elif local_workers > 1 and getattr(frame, "is_cuda", False):
            raise ValueError("render_frame(gather=True, local_workers > 1) on a CUDA frame needs sync_fn: the workers' "
                             "streams must be joined to the current stream before the reduce reads their tiles")
        # disjoint tiles + zero elsewhere: SUM onto rank 0 is the gather of the framebuffer
        dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
        if store is not None and rank == 0 and epoch > 0:
            try:
                store.delete_key(f"ptb_tiles_{epoch - 1}")  # everybody has left that frame (they are in this reduce)
            except Exception:
                pass  # not every store implements delete_key
    t2 = time.perf_counter()
    return dict(rays=totals["rays"], paths=totals["paths"], tiles=sorted(done), seconds_render=t1 - t0,
                seconds_gather=t2 - t1)


def all_sum(values: Sequence[float], device=None) -> List[float]:
    """Sum of per-rank scalars over all ranks (ray / path counters)."""
    dist = _dist()
    if not dist or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.cpu()]


def all_max(values: Sequence[float], device=None) -> List[float]:
    """Max over ranks (timings are reported as the slowest rank's)."""
    dist = _dist()
    if not dist or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def broadcast_description(desc, src: int = 0):
    """Scene replication: rank ``src`` sends the flat description, every rank then builds the same
    KD trees locally (deterministic) and uploads them.  Array payloads go through ``broadcast`` as
    byte tensors (NCCL when the group is NCCL); the small structure goes through ``broadcast_object_list``."""
    dist = _dist()
    if not dist or dist.get_world_size() == 1:
        return desc
    import torch
    from . import SceneDescription
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if rank == src:
        arrays, meta_meshes = [], []
        for m in desc.meshes:
            entry = {}
            for k in ("positions", "normals", "tangents", "uvs", "indices"):
                entry[k] = (m[k].shape, str(m[k].dtype))
                arrays.append(m[k])
            meta_meshes.append(entry)
        meta = dict(meshes=meta_meshes, surfaces=desc.surfaces.tolist(),
                    instances=[(o.tolist(), b.tolist(), f, c) for (o, b, f, c) in desc.instances],
                    materials=desc.materials, camera=(desc.camera[0].tolist(), desc.camera[1].tolist(), desc.camera[2]),
                    sun=None if desc.sun is None else tuple(np.asarray(x).tolist() if i < 2 else x
                                                            for i, x in enumerate(desc.sun)),
                    environment_factor=desc.environment_factor, transparent_background=desc.transparent_background,
                    kd_use_sah=desc.kd_use_sah, kd_max_depth=desc.kd_max_depth,
                    environment_texture=getattr(desc, "environment_texture", None),
                    textures=[(t["pixels"].shape, str(t["pixels"].dtype), bool(t.get("srgb", False)))
                              for t in desc.textures])
        arrays += [np.ascontiguousarray(t["pixels"]) for t in desc.textures]
        box = [meta]
    else:
        arrays, box = None, [None]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    shapes = [v for m in meta["meshes"] for v in (m[k] for k in ("positions", "normals", "tangents", "uvs", "indices"))]
    shapes += [(s, d) for (s, d, _) in meta["textures"]]
    out_arrays = []
    for i, (shape, dtype) in enumerate(shapes):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(arrays[i]).view(np.uint8).reshape(-1).copy()).to(dev)
        else:
            t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if nbytes:
            dist.broadcast(t, src=src)
        out_arrays.append(t.cpu().numpy().view(dtype).reshape(shape))
    if rank == src:
        return desc
    it = iter(out_arrays)
    meshes = [{k: next(it) for k in ("positions", "normals", "tangents", "uvs", "indices")} for _ in meta["meshes"]]
    textures = [dict(pixels=next(it), srgb=srgb) for (_, _, srgb) in meta["textures"]]
    return SceneDescription(meshes, np.array(meta["surfaces"], np.uint32).reshape(-1, 2), meta["instances"],
                            meta["materials"], meta["camera"], meta["sun"], meta["environment_factor"],
                            meta["transparent_background"], meta["kd_use_sah"], meta["kd_max_depth"], textures,
                            environment_texture=meta.get("environment_texture"))


# ---------------------------------------------------------------------------------------------------------
# Geometry-sharded closest hit: the reference's own distribution axis.
#
# ``intersection_worker.cpp:69-147``: every worker holds a subset of the geometry, every ray visits every
# worker, and the per-ray results are merged — closest hit = the smallest distance (``:85-92``), shadow ray =
# OR of the workers' answers (``:126-133``).  Here a shard is a subset of the scene's INSTANCES (the unit
# at which the reference's two-level model/mesh structure can be cut without touching a KD tree); meshes
# that no kept instance uses are dropped from the shard, so a shard's HBM footprint is its own geometry only.
#
# The merge is exact: ``renderer::intersect`` scans the instances in order and keeps a candidate only when
# it is strictly nearer (``renderer.cpp:663-669``), i.e. the result is the minimum of (distance, instance
# index) in lexicographic order.  Distances of hits are non-negative floats, whose bit patterns order like
# the numbers, so one integer MIN over the key  t_bits << 32 | instance << 12 | surface  reproduces the
# unsharded answer bit for bit whatever the partition; the winner's triangle and barycentrics follow with
# one SUM (exactly one rank owns the winning instance, everybody else contributes zeros).

HIT_MISS_KEY = np.int64(0x7FFFFFFFFFFFFFFF)
_SURFACE_BITS = 12  # kernels.hpp: HIT_SURFACE_BITS


def shard_instances(desc, rank: int, world_size: int):
    """→ (shard description, global instance index of every kept instance).  Instance i goes to rank
    i % world_size; unused meshes / surfaces are dropped and the rest renumbered."""
    from . import SceneDescription
    keep = [i for i in range(len(desc.instances)) if i % world_size == rank]
    surf_map, mesh_map, surfaces, meshes, instances = {}, {}, [], [], []
    for i in keep:
        o, b, first, count = desc.instances[i]
        new_first = len(surfaces)
        for s in range(first, first + count):
            mesh, mat = (int(v) for v in desc.surfaces[s])
            if mesh not in mesh_map:
                mesh_map[mesh] = len(meshes)
                meshes.append(desc.meshes[mesh])
            surfaces.append((mesh_map[mesh], mat))
            surf_map[s] = len(surfaces) - 1
        instances.append((o, b, new_first, count))
    shard = SceneDescription(meshes, np.array(surfaces, np.uint32).reshape(-1, 2), instances, desc.materials,
                             desc.camera, desc.sun, desc.environment_factor, desc.transparent_background,
                             desc.kd_use_sah, desc.kd_max_depth, desc.textures,
                             environment_texture=getattr(desc, "environment_texture", None))
    return shard, np.array(keep, np.uint32)


def hit_keys(hits, instance_map) -> np.ndarray:
    """Merge key per ray (int64): distance bits, then global instance, then surface; HIT_MISS_KEY for a miss."""
    inst = hits["instance"]
    miss = inst == np.uint32(0xFFFFFFFF)
    glob = np.asarray(instance_map, np.uint32)[np.where(miss, 0, inst)] if len(instance_map) else inst
    key = (hits["t"].view(np.uint32).astype(np.int64) << 32) | (glob.astype(np.int64) << _SURFACE_BITS) | \
        hits["surface"].astype(np.int64)
    return np.where(miss, HIT_MISS_KEY, key)


def merge_closest_hits(hits, instance_map, device=None):
    """All ranks pass their shard's hits for the SAME rays; every rank gets the merged hits (global instance
    indices) — the closest-hit merge of ``intersection_worker.cpp:69-110`` as two all-reduces."""
    from . import HIT_DTYPE
    dist = _dist()
    key = hit_keys(hits, instance_map)
    payload = np.empty((len(hits), 4), np.int32)
    payload[:, 0] = hits["triangle"].view(np.int32)
    payload[:, 1:] = hits["bary"].view(np.int32)
    best = key
    if dist and dist.get_world_size() > 1:
        import torch
        tk = torch.from_numpy(key.copy())
        tk = tk.to(device) if device is not None else tk
        dist.all_reduce(tk, op=dist.ReduceOp.MIN)
        best = tk.cpu().numpy()
        mine = (key == best) & (best != HIT_MISS_KEY)
        tp = torch.from_numpy(np.where(mine[:, None], payload, 0).astype(np.int32))
        tp = tp.to(device) if device is not None else tp
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
        payload = tp.cpu().numpy()
    return unpack_merged(best, payload, HIT_DTYPE)


def unpack_merged(best_key, payload, hit_dtype):
    """Merged keys + winner payload → hit records laid out like ptb_trace_rays' (misses: ids = 0xFFFFFFFF, t = -1)."""
    out = np.zeros(len(best_key), hit_dtype)
    miss = best_key == HIT_MISS_KEY
    low = (best_key & 0xFFFFFFFF).astype(np.uint32)
    out["instance"] = np.where(miss, np.uint32(0xFFFFFFFF), low >> np.uint32(_SURFACE_BITS))
    out["surface"] = np.where(miss, np.uint32(0xFFFFFFFF), low & np.uint32((1 << _SURFACE_BITS) - 1))
    out["t"] = np.where(miss, np.float32(-1.0), (best_key >> 32).astype(np.uint32).view(np.float32))
    out["triangle"] = np.where(miss, np.uint32(0xFFFFFFFF), payload[:, 0].view(np.uint32))
    out["bary"] = np.where(miss[:, None], np.float32(0), payload[:, 1:].view(np.float32))
    return out


def merge_occlusion(occluded, device=None):
    """Shadow rays: a ray is occluded when ANY shard reports a hit (``intersection_worker.cpp:126-133``)."""
    dist = _dist()
    occ = np.ascontiguousarray(occluded, np.uint8)
    if not dist or dist.get_world_size() == 1:
        return occ.astype(bool)
    import torch
    t = torch.from_numpy(occ.copy())
    t = t.to(device) if device is not None else t
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy().astype(bool)


def trace_rays_sharded(shard_scene, instance_map, origin_dir, device=None):
    """Closest hits of `origin_dir` (identical on every rank) against the union of all ranks' shards."""
    return merge_closest_hits(shard_scene.trace_rays(origin_dir), instance_map, device=device)


class ShardMergeContext:
    """Device-resident geometry-shard merge: the exchange step runs as 64-bit atomics and plain stores into the
    other GPUs' memory over NVLink (``ptb_shard_*_dev``, include/ptb.h), not through NCCL.  The buffers are
    ``torch.distributed._symmetric_memory`` allocations, so every rank knows every rank's device address;
    ``torch.distributed`` (NCCL) is only the rendezvous."""

    def __init__(self, shard_scene, instance_map, capacity: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.torch, self.dist, self.C = torch, dist, C
        self.scene = shard_scene
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.capacity = int(capacity)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.keys = symm.empty(self.capacity, dtype=torch.int64, device=dev)
        self.payload = symm.empty(self.capacity * 4, dtype=torch.int32, device=dev)
        self.h_keys = symm.rendezvous(self.keys, self.group)
        self.h_payload = symm.rendezvous(self.payload, self.group)
        self.key_ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.h_keys.buffer_ptrs])
        self.payload_ptrs = (C.c_void_p * self.world)(*[int(p) for p in self.h_payload.buffer_ptrs])
        self.instance_map = torch.from_numpy(np.ascontiguousarray(instance_map, np.uint32).view(np.int32)).to(dev)
        if self.instance_map.numel() == 0:  # a shard may be empty; the kernel never dereferences the map then
            self.instance_map = torch.zeros(1, dtype=torch.int32, device=dev)

    def trace(self, rays_dev, hits_dev=None):
        """rays_dev: float32 CUDA tensor [n, 6], identical on every rank.  → uint8 CUDA tensor [n, 28] holding
        ptb_hit records (view it with HIT_DTYPE on the host): the closest hits against ALL ranks' shards."""
        from . import lib, _check, HIT_DTYPE
        torch, C = self.torch, self.C
        n = int(rays_dev.shape[0])
        if n > self.capacity:
            raise ValueError("more rays than the merge buffers hold")
        rays_dev = rays_dev.contiguous()
        if hits_dev is None:
            hits_dev = torch.empty((n, HIT_DTYPE.itemsize), dtype=torch.uint8, device=rays_dev.device)
        st = torch.cuda.current_stream().cuda_stream
        L = lib()
        _check(L.ptb_shard_reset_dev(C.c_void_p(self.keys.data_ptr()), n, C.c_void_p(st)))
        self.h_keys.barrier(channel=0)       # everybody's key buffer is reset before anybody merges into it
        _check(L.ptb_shard_trace_dev(self.scene.h, C.c_void_p(rays_dev.data_ptr()), n,
                                     C.c_void_p(self.instance_map.data_ptr()), self.key_ptrs, self.world, C.c_void_p(st)))
        self.h_keys.barrier(channel=1)       # all minima have landed
        _check(L.ptb_shard_publish_dev(self.scene.h, n, C.c_void_p(self.keys.data_ptr()), self.payload_ptrs,
                                       self.world, C.c_void_p(st)))
        self.h_payload.barrier(channel=0)    # the winners' payloads have landed
        _check(L.ptb_shard_unpack_dev(C.c_void_p(self.keys.data_ptr()), C.c_void_p(self.payload.data_ptr()), n,
                                      C.c_void_p(hits_dev.data_ptr()), C.c_void_p(st)))
        return hits_dev

    def occlusion(self, rays_dev):
        """Shadow rays (float32 CUDA tensor [n, 6], identical on every rank) → uint8 CUDA tensor [n]: 1 where ANY
        rank's shard occludes the ray.  The OR is done by the any-hit kernel itself, into every rank's buffer over
        NVLink (``ptb_shard_occlusion_dev``); the key buffer doubles as the occlusion buffer."""
        from . import lib, _check
        torch, C = self.torch, self.C
        n = int(rays_dev.shape[0])
        if n > self.capacity * 8:
            raise ValueError("more rays than the merge buffers hold")
        rays_dev = rays_dev.contiguous()
        occ = self.keys.view(torch.uint8)
        occ[: (n + 3) // 4 * 4].zero_()
        st = torch.cuda.current_stream().cuda_stream
        self.h_keys.barrier(channel=0)       # everybody's buffer is zero before anybody ORs into it
        _check(lib().ptb_shard_occlusion_dev(self.scene.h, C.c_void_p(rays_dev.data_ptr()), n, self.key_ptrs, self.world,
                                             C.c_void_p(st)))
        self.h_keys.barrier(channel=1)       # all ORs have landed
        return occ[:n].clone()


# ---------------------------------------------------------------------------------------------------------
# Process-group glue for the frame driver of libptb (ptb_group): scene replication and the rendezvous name.

class _DeviceBytes:
    """A raw device allocation as something ``torch.as_tensor`` understands (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_bytes_tensor(ptr: int, nbytes: int, device):
    """uint8 CUDA tensor aliasing [ptr, ptr + nbytes) — no copy; the memory stays owned by libptb."""
    import torch
    return torch.as_tensor(_DeviceBytes(ptr, nbytes), device=device)


def replicate_scene(scene, device_index: int, src: int = 0, chunk_bytes: int = 256 << 20):
    """Rank ``src`` passes its built :class:`Scene`, the others pass None.  → every rank's replica.

    The scene is built ONCE: the header (a few hundred bytes of plain data) travels with ``broadcast_object_list``,
    the flattened HBM blob with ``broadcast`` over NCCL straight from rank src's blob into the other ranks'
    (``ptb_scene_export_header`` / ``ptb_scene_import`` / ``ptb_scene_blob``); nobody rebuilds a tree."""
    dist = _dist()
    if not dist or dist.get_world_size() == 1:
        return scene
    import torch
    from . import Scene
    rank = dist.get_rank()
    dev = torch.device("cuda", device_index)
    box = [scene.export_header() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    if rank != src:
        scene = Scene.import_header(box[0], device_index)  # allocates the blob, leaves it to be filled
    ptr, nbytes = scene.blob()
    blob = device_bytes_tensor(ptr, nbytes, dev)
    for off in range(0, nbytes, chunk_bytes):
        dist.broadcast(blob[off:off + chunk_bytes], src=src)
    torch.cuda.synchronize(dev)
    return scene


def make_group(device_index: int):
    """A :class:`Group` over the default process group's ranks (one node): rank 0 picks the shared-memory name."""
    import os
    import uuid
    from . import Group
    dist = _dist()
    if not dist or dist.get_world_size() == 1:
        return Group(f"solo_{os.getpid()}_{uuid.uuid4().hex[:8]}", 0, 1, device_index)
    box = [f"{os.getpid()}_{uuid.uuid4().hex[:12]}" if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return Group(box[0], dist.get_rank(), dist.get_world_size(), device_index)
