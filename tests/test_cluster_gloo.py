"""CPU: the multi-rank tile path (work stealing through the store, disjoint-tile reduce, scene broadcast)
with world_size 2 on gloo and a fake tile renderer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, static, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import importlib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cluster = importlib.import_module("distributed-path-tracer_b200.cluster")
    procedural = importlib.import_module("distributed-path-tracer_b200.procedural")
    W, H = 64, 40
    tiles = cluster.make_tiles(W, H, *cluster.tile_grid_for(world, 4))
    frame = torch.zeros((H, W, 4), dtype=torch.float32)

    def fake_render(tile, out, worker):
        x0, y0, w, h = tile
        ys, xs = np.mgrid[y0:y0 + h, x0:x0 + w]
        val = np.stack([xs, ys, xs * 1000 + ys, np.ones_like(xs)], -1).astype(np.float32)
        out[y0:y0 + h, x0:x0 + w] = torch.from_numpy(val)
        return {"rays": w * h * 3, "paths": w * h}

    for epoch in range(3):
        frame.zero_()
        r = cluster.render_frame(W, H, tiles, fake_render, frame, epoch=epoch, static_assignment=static,
                                 local_workers=1 + epoch % 2)
        rays, paths = cluster.all_sum([r["rays"], r["paths"]])
        n_tiles = cluster.all_sum([len(r["tiles"])])[0]
        assert (rays, paths, n_tiles) == (W * H * 3, W * H, len(tiles))
        if rank == 0:
            ys, xs = np.mgrid[0:H, 0:W]
            want = np.stack([xs, ys, xs * 1000 + ys, np.ones_like(xs)], -1).astype(np.float32)
            assert np.array_equal(frame.numpy(), want), "gathered frame is wrong"
    assert cluster.all_max([float(rank)])[0] == world - 1
    # scene replication
    desc = procedural.heightfield_scene(6) if rank == 0 else None
    got = cluster.broadcast_description(desc, src=0)
    ref = procedural.heightfield_scene(6)
    for a, b in zip(got.meshes, ref.meshes):
        for k in a:
            assert np.array_equal(a[k], b[k])
    assert got.instances[1][2:] == ref.instances[1][2:] and got.materials == ref.materials
    _geometry_sharding(rank, world, cluster, procedural)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def _instanced_scene(procedural, ptb_mod):
    rng = np.random.default_rng(11)
    a, b = procedural.heightfield_mesh(6, 1.0, 3), procedural.heightfield_mesh(3, 0.7, 4)
    insts = []
    for i in range(9):
        ang = rng.uniform(0, 6.28)
        sc = rng.uniform(0.5, 1.6, 3)
        c, s_ = np.cos(ang), np.sin(ang)
        basis = np.array([c * sc[0], 0, -s_ * sc[0], 0, sc[1], 0, s_ * sc[2], 0, c * sc[2]], np.float32)
        insts.append((rng.uniform(-4, 4, 3) * (1, 0.2, 1), basis, i % 2, 1))
    insts.append(((0, 0.1, 0), np.eye(3, dtype=np.float32).ravel(), 0, 2))  # two surfaces, overlaps instance 0..8
    insts.append(insts[2])  # an exact duplicate: equal distances, the lower instance index must win
    mats = [dict(albedo=(0.7, 0.7, 0.7), roughness=1.0, metallic=0.0)] * 2
    cam = procedural.look_at((0, 6, 14), (0, 0, 0))
    return ptb_mod.SceneDescription([a, b], [(0, 0), (1, 1)], insts, mats, (cam[0], cam[1], 0.8))


def _geometry_sharding(rank, world, cluster, procedural):
    """Closest-hit merge across geometry shards (intersection_worker.cpp:69-147) == the unsharded search, bit for
    bit; the per-shard searches are done by the plain-C oracle (this is a CPU test)."""
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import portlib
    import reflib
    ptb_mod = importlib.import_module("distributed-path-tracer_b200")
    desc = _instanced_scene(procedural, ptb_mod)
    rng = np.random.default_rng(3)
    n = 20_000
    o = (rng.uniform(-6, 6, (n, 3)) * (1, 0.4, 1)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    od = np.concatenate([o, d], 1)

    def flat(ds):
        return reflib.FlatScene(ds.meshes, ds.surfaces, ds.instances, ds.materials, ds.camera)

    want = portlib.PortScene(flat(desc)).trace_rays(od)
    shard, imap = cluster.shard_instances(desc, rank, world)
    assert list(imap) == [i for i in range(len(desc.instances)) if i % world == rank]
    local = portlib.PortScene(flat(shard)).trace_rays(od)
    got = cluster.merge_closest_hits(local, imap)
    for f in ("instance", "surface", "triangle"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    assert np.array_equal(got["bary"].view(np.uint32), want["bary"].view(np.uint32))
    hit = want["instance"] != 0xFFFFFFFF
    assert 0.02 < hit.mean() < 0.95 and (want["instance"][hit] == 9).any(), hit.mean()
    assert not (want["instance"] == 10).any()  # the duplicate never wins a tie against instance 2
    occ = cluster.merge_occlusion(local["instance"] != 0xFFFFFFFF)
    assert np.array_equal(occ, hit)


@pytest.mark.parametrize("static", [False, True])
def test_two_ranks_gloo(tmp_path, static):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, static, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_tile_helpers():
    import importlib
    sys.path.insert(0, ROOT)
    cluster = importlib.import_module("distributed-path-tracer_b200.cluster")
    for world in (1, 2, 4, 8):
        cols, rows = cluster.tile_grid_for(world, 8)
        assert cols * rows == 8 * world
        tiles = cluster.make_tiles(1920, 1080, cols, rows)
        cover = np.zeros((1080, 1920), np.int32)
        for (x0, y0, w, h) in tiles:
            cover[y0:y0 + h, x0:x0 + w] += 1
        assert (cover == 1).all()
    tiles = cluster.make_tiles(101, 67, 7, 5)  # ragged
    assert sum(w * h for (_, _, w, h) in tiles) == 101 * 67
    c = cluster.TileCounter(3)
    assert [c.next() for _ in range(5)] == [0, 1, 2, -1, -1]
    with pytest.raises(ValueError):
        cluster.make_tiles(4, 4, 8, 1)
