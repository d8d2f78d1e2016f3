"""GPU (B200): the multi-GPU frame driver (ptb_ctx / ptb_group), scene replication, caller-owned chaining
state and thread safety of the host entry points — all through the C ABI.  A frame rendered by the frame
driver, on any number of GPUs and with any tile size, must equal the single-call render BIT FOR BIT."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cornell(ptb, procedural):
    s = ptb.Scene.load_gltf(procedural.cornell_gltf_path(), device=0)
    yield s
    s.close()


def _transparent_scene(ptb, procedural):
    d = procedural.heightfield_scene(16)
    d.transparent_background = True
    d.camera = (np.array([0, 3, 12], np.float32), d.camera[1], 0.9)  # part of the frame sees the sky
    return ptb.Scene.create(d)


@pytest.mark.parametrize("transparent", [False, True])
def test_sample_ranges_chain_with_caller_owned_state(ptb, procedural, cornell, transparent):
    """8 + 8 samples chained through the caller's buffers == 16 samples at once, bit for bit — with ANOTHER tile and
    another scene rendered on the same device and stream in between (the library keeps no chaining state), with and
    without the transparent-background claim mask."""
    s = _transparent_scene(ptb, procedural) if transparent else cornell
    W, Hh = 64, 36
    try:
        want_rgb, want_a, _ = s.render_tile(W, Hh, 16, 4, seed=9)
        rgb = np.zeros((Hh, W, 3), np.float32)
        alpha = np.zeros((Hh, W), np.float32)
        mask = np.zeros((Hh, W), np.uint8)
        s.render_tile(W, Hh, 8, 4, seed=9, state=(rgb, alpha, mask))
        # something else on the same device / default stream, larger than the chained tile (workspace regrows)
        cornell.render_tile(200, 120, 2, 3, seed=1)
        s.render_tile(W, Hh, 4, 4, tile=(8, 4, 40, 20), seed=77)
        s.render_tile(W, Hh, 8, 4, seed=9, first_sample=8, state=(rgb, alpha, mask))
        assert np.array_equal(H.bits(rgb), H.bits(want_rgb))
        assert np.array_equal(H.bits(alpha), H.bits(want_a))
        if transparent:
            assert (alpha == 0).any() and (alpha == 1).any() and mask.any() and not mask.all()
            # chaining a transparent-background scene without the mask is refused, not silently wrong
            with pytest.raises(ptb.PtbError) as e:
                s.render_tile(W, Hh, 8, 4, seed=9, first_sample=8)
            assert e.value.status == ptb.PTB_E_INVALID and "claim_mask" in str(e.value)
    finally:
        if transparent:
            s.close()


def test_sample_ranges_chain_on_the_device(ptb, procedural):
    """The same through ptb_render_tile_dev: rgba_dev is in/out, the claim mask is a caller-owned device buffer."""
    import torch
    s = _transparent_scene(ptb, procedural)
    try:
        W, Hh = 64, 36
        want_rgb, want_a, _ = s.render_tile(W, Hh, 12, 4, seed=5)
        rgba = torch.full((Hh * W * 4,), 7.0, dtype=torch.float32, device="cuda")  # garbage: the first call must not read it
        mask = torch.full((Hh * W,), 3, dtype=torch.uint8, device="cuda")
        other = torch.zeros(128 * 128 * 4, dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        s.render_tile_dev(rgba.data_ptr(), W, Hh, 5, 4, seed=5, stream=st, claim_mask_dev=mask.data_ptr())
        s.render_tile_dev(other.data_ptr(), 128, 128, 3, 4, seed=2, stream=st)  # in between, same stream
        s.render_tile_dev(rgba.data_ptr(), W, Hh, 7, 4, seed=5, first_sample=5, stream=st, claim_mask_dev=mask.data_ptr())
        torch.cuda.synchronize()
        got = rgba.cpu().numpy().reshape(Hh, W, 4)
        assert np.array_equal(H.bits(got[..., :3]), H.bits(want_rgb))
        assert np.array_equal(H.bits(got[..., 3]), H.bits(want_a))
    finally:
        s.close()


def test_host_entry_is_thread_safe(cornell):
    """Several host threads calling ptb_render_tile on the same device share one workspace: every thread must get
    ITS tile (bit-identical to the sequential render), whatever the interleaving."""
    tiles = [(0, 0, 96, 64), (0, 0, 40, 40), (20, 10, 60, 50), (50, 30, 46, 34), (3, 5, 80, 20), (0, 0, 17, 9)]
    want = [cornell.render_tile(96, 64, 6, 4, tile=t, seed=11)[0] for t in tiles]
    got = [[None] * len(tiles) for _ in range(3)]
    errors = []

    def work(k):
        try:
            for rep in range(3):
                got[rep][k] = cornell.render_tile(96, 64, 6, 4, tile=tiles[k], seed=11)[0]
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(tiles))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for rep in range(3):
        for k in range(len(tiles)):
            assert np.array_equal(H.bits(got[rep][k]), H.bits(want[k])), (rep, k)


def test_scene_replicas_are_exact(ptb, cornell):
    """ptb_scene_clone (device-to-device copy of the blob) and export_header / import: the replica answers every ray
    and renders every pixel exactly like the scene it came from; replicas refuse ptb_scene_dump_kd."""
    r = H.load("cornell_rays.npz")
    rep = cornell.clone(0)
    try:
        for key in ("cam", "rnd", "bounce"):
            H.assert_hits_equal(rep.trace_rays(r[key + "_rays"]), r[key + "_hits"], "replica:" + key)
        a, _, _ = cornell.render_tile(64, 48, 4, 4, seed=2)
        b, _, _ = rep.render_tile(64, 48, 4, 4, seed=2)
        assert np.array_equal(H.bits(a), H.bits(b))
        with pytest.raises(ptb.PtbError):
            rep.dump_kd(0)
        i0, i1 = cornell.info(), rep.info()
        for k in ("n_triangles", "n_kd_nodes", "n_leaf_refs", "device_bytes", "n_instances"):
            assert i0[k] == i1[k]
    finally:
        rep.close()
    hdr = cornell.export_header()
    blob, nbytes = cornell.blob()
    rep2 = ptb.Scene.import_header(hdr, 0, src_blob=blob, src_device=0)
    try:
        H.assert_hits_equal(rep2.trace_rays(r["rnd_rays"]), r["rnd_hits"], "imported")
        assert rep2.blob()[1] == nbytes
    finally:
        rep2.close()
    with pytest.raises(ptb.PtbError):
        ptb.Scene.import_header(hdr[:-8], 0)
    with pytest.raises(ptb.PtbError):
        ptb.Scene.import_header(b"\0" * len(hdr), 0)


def _frame_cases():
    # (W, H, spp, depth, integrator, tile, in flight)
    return [(96, 64, 8, 4, 0, (0, 0), 0), (96, 64, 8, 4, 0, (32, 16), 3), (100, 70, 5, 6, 1, (48, 24), 2),
            (64, 64, 3, 4, 0, (64, 64), 1)]


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
def test_context_frame_is_bit_identical_to_the_single_call(ptb, procedural, cornell, n_gpus):
    """ptb_render_frame on n GPUs (tiles stolen from one counter, every GPU's accumulate kernel storing into GPU 0's
    frame over NVLink) == ptb_render_tile of the whole frame on one GPU, bit for bit; RGBA8 output == the tonemap of
    the float frame; statistics add up."""
    if ptb.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    ctx = ptb.Context(n_gpus)
    try:
        ctx.load_gltf(procedural.cornell_gltf_path())
        for (W, Hh, spp, depth, integ, tile, inflight) in _frame_cases():
            want_rgb, want_a, st1 = cornell.render_tile(W, Hh, spp, depth, seed=21, integrator=integ)
            frame, st = ctx.render_frame(W, Hh, spp, depth, seed=21, integrator=integ, tile=tile, tiles_in_flight=inflight)
            assert frame.shape == (Hh, W, 4)
            assert np.array_equal(H.bits(frame[..., :3]), H.bits(want_rgb)), (W, Hh, tile)
            assert np.array_equal(H.bits(frame[..., 3]), H.bits(want_a))
            assert st["paths"] == W * Hh * spp == st1["paths"] and st["rays"] == st1["rays"]
            assert st["n_ranks"] == n_gpus and sum(st["tiles_per_rank"]) == st["n_tiles"]
            assert st["gpu_seconds"] > 0 and st["wall_seconds"] >= st["gpu_seconds"] * 0.5
            rgba8, _ = ctx.render_frame(W, Hh, spp, depth, seed=21, integrator=integ, tile=tile, output=ptb.OUT_RGBA8)
            assert np.array_equal(rgba8, ptb.tonemap_rgba8(want_rgb, want_a))
        # pinned output buffer (no staging copy) and a frame that stays on the device
        pin = ptb.PinnedBuffer((64, 96, 4), np.float32)
        want_rgb, _, _ = cornell.render_tile(96, 64, 8, 4, seed=21)
        ctx.render_frame(96, 64, 8, 4, seed=21, out=pin.array)
        assert np.array_equal(H.bits(pin.array[..., :3]), H.bits(want_rgb))
        pin.close()
        none, st = ctx.render_frame(96, 64, 8, 4, seed=21, output=ptb.OUT_NONE)
        assert none is None and st["rays"] > 0
        # the replicas ARE the scene: every GPU answers the golden rays
        r = H.load("cornell_rays.npz")
        for i in range(n_gpus):
            H.assert_hits_equal(ctx.scene(i).trace_rays(r["rnd_rays"]), r["rnd_hits"], f"replica on GPU {i}")
        # errors come back as status codes, and the context survives them
        with pytest.raises(ptb.PtbError):
            ctx.render_frame(96, 64, 8, 300)
        frame, _ = ctx.render_frame(96, 64, 8, 4, seed=21)
        assert np.array_equal(H.bits(frame[..., :3]), H.bits(want_rgb))
    finally:
        ctx.close()


def test_comb_tiles_render_the_same_frame(ptb, procedural, cornell):
    """The library's tiling for several ranks (comb tiles: every tile holds pixel granules spread over the whole frame,
    include/ptb.h ptb_frame_tile_layout) forced onto whatever GPUs this box has: the tiles cover every pixel once, and
    the frame is the single call's frame bit for bit — odd frame sizes, both integrators, transparent background."""
    n = min(2, ptb.device_count())
    ptb.set_option("frame_comb_tiles", 2)
    try:
        for (W, Hh, inflight) in [(96, 64, 3), (250, 131, 4), (333, 77, 2)]:
            lay = ptb.frame_tile_layout(W, Hh, 4, n, tiles_in_flight=inflight)
            assert len(lay) % (n * inflight) == 0 and (lay[:, 4] > 0).all() and (lay[:, 6] > 0).all()
            cover = np.zeros((Hh, W), np.int32)
            for row in lay:
                xs, ys = ptb.tile_pixels(row)
                cover[np.ix_(ys, xs)] += 1
            assert (cover == 1).all()
        ctx = ptb.Context(n)
        try:
            ctx.load_gltf(procedural.cornell_gltf_path())
            for (W, Hh, spp, depth, integ, inflight) in [(96, 64, 8, 4, 0, 3), (250, 131, 5, 6, 1, 4), (333, 77, 3, 4, 0, 2)]:
                want_rgb, want_a, st1 = cornell.render_tile(W, Hh, spp, depth, seed=5, integrator=integ)
                frame, st = ctx.render_frame(W, Hh, spp, depth, seed=5, integrator=integ, tiles_in_flight=inflight)
                assert st["n_tiles"] % (n * inflight) == 0
                assert np.array_equal(H.bits(frame[..., :3]), H.bits(want_rgb)), (W, Hh)
                assert np.array_equal(H.bits(frame[..., 3]), H.bits(want_a))
                assert st["rays"] == st1["rays"] and st["paths"] == st1["paths"]
        finally:
            ctx.close()
        d = procedural.heightfield_scene(16)
        d.transparent_background = True
        d.camera = (np.array([0, 3, 12], np.float32), d.camera[1], 0.9)
        ctx = ptb.Context(n)
        try:
            ctx.set_scene(d)
            with ptb.Scene.create(d) as s:
                want_rgb, want_a, _ = s.render_tile(150, 90, 16, 4, seed=1)
            frame, st = ctx.render_frame(150, 90, 16, 4, seed=1, tiles_in_flight=3)
            assert np.array_equal(H.bits(frame[..., :3]), H.bits(want_rgb)) and np.array_equal(H.bits(frame[..., 3]), H.bits(want_a))
        finally:
            ctx.close()
    finally:
        ptb.set_option("frame_comb_tiles", 1)


def test_context_transparent_background_and_worker_request(ptb, procedural, tmp_path):
    """Transparent-background claim logic through the frame driver, and the worker_info request over a context."""
    import json
    import shutil
    n = min(2, ptb.device_count())
    d = procedural.heightfield_scene(16)
    d.transparent_background = True
    d.camera = (np.array([0, 3, 12], np.float32), d.camera[1], 0.9)
    ctx = ptb.Context(n)
    try:
        ctx.set_scene(d)
        with ptb.Scene.create(d) as s:
            want_rgb, want_a, _ = s.render_tile(64, 36, 16, 4, seed=1)
        frame, _ = ctx.render_frame(64, 36, 16, 4, seed=1, tile=(32, 16))
        assert np.array_equal(H.bits(frame[..., :3]), H.bits(want_rgb)) and np.array_equal(H.bits(frame[..., 3]), H.bits(want_a))
        src = os.path.dirname(procedural.cornell_gltf_path())
        scene_dir = tmp_path / "scenes" / "cornell"
        scene_dir.mkdir(parents=True)
        shutil.copy(os.path.join(src, "cornell.gltf"), scene_dir / "scene.gltf")
        shutil.copy(os.path.join(src, "cornell.bin"), scene_dir / "cornell.bin")
        gl = json.load(open(scene_dir / "scene.gltf"))
        info = {"scene_info": {"work": {m["name"]: list(range(len(m["primitives"]))) for m in gl["meshes"]}},
                "worker_id": "1", "num_workers": 1, "samples": 8, "bounces": 6, "X": 48, "Y": 32}
        one, st1 = ptb.worker_run(info, str(scene_dir))
        many, st = ctx.worker_run(info, str(scene_dir), png_path=str(tmp_path / "t.png"))
        assert np.array_equal(one, many) and st["rays"] == st1["rays"]
    finally:
        ctx.close()


@pytest.mark.parametrize("world", [2, 8])
def test_process_group_frame_under_torchrun(ptb, world):
    """One process per GPU (torchrun): scene built on rank 0 and broadcast over NCCL as one blob, tiles stolen through
    shared memory, the frame written into rank 0's GPU through CUDA IPC — bit-identical to one GPU
    (scripts/group_check.py)."""
    if ptb.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
                        os.path.join(ROOT, "scripts", "group_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "group frame == single-GPU frame: True" in r.stdout
