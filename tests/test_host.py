"""CPU: host logic of the product (KD builder, glTF loader, PNG, C-ABI surface) — no compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers as H


def test_abi_exports_every_declared_symbol(ptb):
    header = open(ptb.HEADER_PATH).read()
    declared = set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", header))
    lib = ptb.lib()
    assert declared, "no declarations found in include/ptb.h"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libptb.so does not export {name}"
    assert declared == set(ptb.EXPORTS), declared ^ set(ptb.EXPORTS)
    want = int(re.search(r"#define\s+PTB_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.ptb_abi_version() == want


def test_graft_entry_build_runs():
    """The driver's build check: __graft_entry__.build() must compile everything and exit cleanly."""
    import __graft_entry__ as g
    g.build()
    assert g.abi_version_of_header() >= 2


def test_struct_layouts_match_header(ptb):
    assert ctypes.sizeof(ptb.TileReq) == 64  # 14 x u32 with the u64 seed at offset 32, + the claim-mask pointer
    assert ptb.TileReq.seed.offset == 32 and ptb.TileReq.claim_mask.offset == 56
    assert ctypes.sizeof(ptb.FrameReq) == 48 and ptb.FrameReq.seed.offset == 16
    assert ctypes.sizeof(ptb.FrameStats) == 3 * 8 + 2 * 4 + 2 * 8 + 16 * 8 + 16 * 8
    assert ptb.HIT_DTYPE.itemsize == 28
    assert ctypes.sizeof(ptb.MaterialDesc) == 4 * (3 + 1 + 1 + 1 + 3 + 1 + 1 + 6)
    assert ctypes.sizeof(ptb.InstanceDesc) == 4 * (3 + 9 + 2)


def test_no_gpu_means_error_not_fallback(ptb, procedural):
    """Without a CUDA device scene creation must FAIL with PTB_E_CUDA (there is no CPU path)."""
    if ptb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ptb.PtbError) as e:
        ptb.Scene.create(procedural.heightfield_scene(4))
    assert e.value.status == ptb.PTB_E_CUDA


def test_invalid_descriptions_are_rejected(ptb, procedural):
    d = procedural.heightfield_scene(4)
    d.meshes[0]["indices"][0, 0] = 10 ** 6
    with pytest.raises(ptb.PtbError) as e:
        ptb.Scene.create(d)
    assert e.value.status == ptb.PTB_E_INVALID
    with pytest.raises(ptb.PtbError):
        ptb.set_option("no_such_option", 1)
    d = procedural.heightfield_scene(4)
    d.environment_texture = 3  # there is no texture 3
    with pytest.raises(ptb.PtbError) as e:
        ptb.Scene.create(d)
    assert e.value.status == ptb.PTB_E_INVALID and "environment" in str(e.value)
    with pytest.raises(ptb.PtbError) as e:
        ptb.load_gltf_description("/nonexistent/scene.gltf")
    assert e.value.status == ptb.PTB_E_IO


def test_kd_builder_reproduces_reference_trees_cornell(ptb):
    z = H.load("cornell_scene.npz")
    kd = H.load("cornell_kd.npz")
    for m in range(int(z["n_meshes"])):
        words, aabb = ptb.host_build_kd(z[f"mesh{m}_positions"], z[f"mesh{m}_indices"])
        assert np.array_equal(words, kd[f"mesh{m}"]), f"mesh {m}"
        assert np.array_equal(H.bits(aabb), H.bits(kd["mesh_aabbs"][m]))


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_kd_builder_heightfield_any_thread_count(ptb, procedural, threads):
    m = procedural.heightfield_mesh(40)
    words, _ = ptb.host_build_kd(m["positions"], m["indices"], threads=threads)
    assert np.array_equal(words, H.load("heightfield40_kd.npz")["mesh0"])


def test_kd_tree_disk_cache(ptb, procedural, tmp_path, monkeypatch):
    """PTB_KD_CACHE: the second build of the same mesh comes from disk and is the same tree; another mesh, another
    depth limit or a damaged file never is."""
    monkeypatch.setenv("PTB_KD_CACHE", str(tmp_path))
    want = H.load("heightfield40_kd.npz")["mesh0"]
    m = procedural.heightfield_mesh(40)
    first, _ = ptb.host_build_kd(m["positions"], m["indices"])
    files = sorted(tmp_path.glob("kd_*.bin"))
    assert len(files) == 1 and np.array_equal(first, want)
    stamp = files[0].stat().st_mtime_ns
    again, _ = ptb.host_build_kd(m["positions"], m["indices"])
    assert np.array_equal(again, want) and files[0].stat().st_mtime_ns == stamp  # served, not rewritten
    # a different depth limit and a different mesh get their own entries
    shallow, _ = ptb.host_build_kd(m["positions"], m["indices"], max_depth=6)
    moved = m["positions"].copy()
    moved[0, 1] += 0.25
    other, _ = ptb.host_build_kd(moved, m["indices"])
    assert len(list(tmp_path.glob("kd_*.bin"))) == 3
    assert not np.array_equal(shallow, want) and not np.array_equal(other, want)
    # a damaged file is ignored and replaced
    raw = bytearray(files[0].read_bytes())
    raw[len(raw) // 2] ^= 0xFF
    files[0].write_bytes(bytes(raw))
    healed, _ = ptb.host_build_kd(m["positions"], m["indices"])
    assert np.array_equal(healed, want)
    assert files[0].read_bytes() != bytes(raw)
    monkeypatch.setenv("PTB_KD_CACHE", str(tmp_path / "does" / "not" / "exist"))
    nocache, _ = ptb.host_build_kd(m["positions"], m["indices"])  # unwritable directory: builds, does not fail
    assert np.array_equal(nocache, want)


def test_kd_builder_against_c_oracle_on_awkward_meshes(ptb, portlib, reflib):
    """Degenerate / axis-aligned / duplicated / all-negative geometry: ties in the SAH sweep, the
    FLT_MIN quirk of aabb::clear, empty children."""
    rng = np.random.default_rng(3)
    cases = []
    # all-negative coordinates (max of the box starts at FLT_MIN, the smallest positive float)
    p = -rng.random((60, 3)).astype(np.float32) - 1
    cases.append((p, rng.integers(0, 60, (80, 3))))
    # axis-aligned quads stacked on a lattice: many equal event positions, flat triangles
    g = np.stack(np.meshgrid(np.arange(5), np.arange(5), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    idx = rng.integers(0, len(g), (150, 3))
    cases.append((g, idx))
    # exact duplicates and zero-area triangles
    p = rng.random((30, 3)).astype(np.float32)
    idx = np.concatenate([rng.integers(0, 30, (40, 3)), np.array([[1, 1, 1], [2, 2, 5], [7, 8, 9], [7, 8, 9]])])
    cases.append((p, idx))
    # single triangle and empty mesh
    cases.append((p, np.array([[0, 1, 2]])))
    cases.append((p, np.zeros((0, 3), np.int64)))
    for pos, idx in cases:
        idx = np.ascontiguousarray(idx, np.uint32)
        for use_sah in (True, False):
            # the median builder splits every node down to the depth limit (2^depth leaves): keep it shallow
            for depth in ((25, 6) if use_sah else (7,)):
                words, aabb = ptb.host_build_kd(pos, idx, use_sah=use_sah, max_depth=depth)
                nv = len(pos)
                mesh = dict(positions=pos, normals=np.zeros((nv, 3), np.float32), tangents=np.zeros((nv, 3), np.float32),
                            uvs=np.zeros((nv, 2), np.float32), indices=idx)
                flat = reflib.FlatScene([mesh], [(0, 0)], [((0, 0, 0), (1, 0, 0, 0, 1, 0, 0, 0, 1), 0, 1)],
                                        [dict()], ((0, 0, 5), (1, 0, 0, 0, 1, 0, 0, 0, 1), 0.7),
                                        kd_use_sah=use_sah, kd_max_depth=depth)
                port = portlib.PortScene(flat)
                assert np.array_equal(words, port.dump_kd(0)), (len(idx), use_sah, depth)
                assert np.array_equal(H.bits(aabb), H.bits(port.mesh_aabb(0)))


def test_gltf_loader_reproduces_reference_scene(ptb, procedural):
    """Vertices, scrambled tangents, transforms, materials, camera and the renderer::intersect visiting
    order, bit for bit against the flat export of the reference's loader."""
    got = ptb.load_gltf_description(procedural.cornell_gltf_path())
    z = H.load("cornell_scene.npz")
    want = H.scene_parts_from_npz(z)
    assert len(got.meshes) == len(want["meshes"])
    for a, b in zip(got.meshes, want["meshes"]):
        for k in H.MESH_KEYS:
            assert np.array_equal(H.bits(a[k]), H.bits(b[k])), k
    assert np.array_equal(got.surfaces, want["surfaces"])
    for a, b in zip(got.instances, want["instances"]):
        assert np.array_equal(H.bits(a[0]), H.bits(b[0])) and np.array_equal(H.bits(a[1]), H.bits(b[1]))
        assert a[2:] == b[2:]
    for a, b in zip(got.materials, want["materials"]):
        for k in ("albedo", "opacity", "roughness", "metallic", "emissive", "ior"):
            assert np.array_equal(np.float32(a[k]), np.float32(b[k])), k
        assert a["albedo_tex"] == ptb.NO_TEXTURE
    assert np.array_equal(H.bits(got.camera[0]), H.bits(want["camera"][0]))
    assert np.array_equal(H.bits(got.camera[1]), H.bits(want["camera"][1]))
    assert np.float32(got.camera[2]) == np.float32(want["camera"][2])
    assert got.sun is None


def test_gltf_and_png_loader_on_jack_of_blades(ptb, reflib):
    """Only where the reference tree is mounted: the textured fixture (17 PNG textures, KHR_lights_punctual sun,
    BLEND materials) loaded by gltf.cpp + png.cpp against the reference's cgltf + stb_image load."""
    path = "/root/reference/path-tracer-core/scenes/jack-of-blades/jack-of-blades.gltf"
    if not (os.path.exists(path) and reflib.available()):
        pytest.skip("reference tree or oracle/_ref not present")
    want = reflib.RefScene.from_gltf(path).export_flat()
    got = ptb.load_gltf_description(path)
    assert len(got.meshes) == len(want.meshes) == 7 and len(got.textures) == len(want.textures) == 17
    for a, b in zip(got.meshes, want.meshes):
        for k in H.MESH_KEYS:
            assert np.array_equal(H.bits(a[k]), H.bits(b[k])), k
    assert np.array_equal(got.surfaces, want.surfaces)
    for a, b in zip(got.instances, want.instances):
        assert np.array_equal(H.bits(a[0]), H.bits(b[0])) and np.array_equal(H.bits(a[1]), H.bits(b[1])) and a[2:] == b[2:]
    for a, b in zip(got.materials, want.materials):
        for k in b:
            assert np.all(np.float32(a[k]) == np.float32(b[k])), k
    for a, b in zip(got.textures, want.textures):
        assert a["srgb"] == b["srgb"] and np.array_equal(a["pixels"], b["pixels"])
    assert got.sun is not None and np.allclose(got.sun[1], want.sun[1]) and np.array_equal(
        np.float32(got.sun[0]), np.float32(want.sun[0]))


def test_gltf_loader_errors(ptb, tmp_path):
    bad = tmp_path / "bad.gltf"
    bad.write_text("{ not json")
    with pytest.raises(ptb.PtbError) as e:
        ptb.load_gltf_description(str(bad))
    assert e.value.status == ptb.PTB_E_IO
    nocam = tmp_path / "nocam.gltf"
    nocam.write_text('{"asset":{"version":"2.0"},"scenes":[{"nodes":[]}],"nodes":[]}')
    with pytest.raises(ptb.PtbError) as e:
        ptb.load_gltf_description(str(nocam))
    assert "camera" in str(e.value)  # renderer.cpp:73-74


def test_gltf_loader_rejects_malformed_accessors(ptb, procedural, tmp_path):
    """A malformed or hostile file must come back as PTB_E_IO, never as an out-of-bounds read: negative / fractional /
    huge offsets, strides and counts, an accessor longer than its bufferView, an index accessor without a bufferView."""
    import json
    import shutil
    src = os.path.dirname(procedural.cornell_gltf_path())
    good = json.load(open(os.path.join(src, "cornell.gltf")))
    shutil.copy(os.path.join(src, "cornell.bin"), tmp_path / "cornell.bin")
    prim = good["meshes"][0]["primitives"][0]
    pos_acc, idx_acc = prim["attributes"]["POSITION"], prim["indices"]

    def broken(edit):
        g = json.loads(json.dumps(good))
        edit(g)
        path = tmp_path / "m.gltf"
        path.write_text(json.dumps(g))
        with pytest.raises(ptb.PtbError) as e:
            ptb.load_gltf_description(str(path))
        assert e.value.status == ptb.PTB_E_IO, str(e.value)

    ptb.load_gltf_description(os.path.join(src, "cornell.gltf"))  # the unedited file loads
    broken(lambda g: g["accessors"][pos_acc].update(byteOffset=-16))
    broken(lambda g: g["accessors"][pos_acc].update(byteOffset=1.5))
    broken(lambda g: g["accessors"][pos_acc].update(count=2 ** 40))
    broken(lambda g: g["accessors"][pos_acc].update(count=2 ** 31))  # (count - 1) * stride would wrap 32 bits
    broken(lambda g: g["bufferViews"][g["accessors"][pos_acc]["bufferView"]].update(byteStride=2 ** 33))
    broken(lambda g: g["bufferViews"][g["accessors"][pos_acc]["bufferView"]].update(byteOffset=2 ** 31))
    broken(lambda g: g["bufferViews"][g["accessors"][pos_acc]["bufferView"]].update(byteLength=8))
    broken(lambda g: g["bufferViews"][g["accessors"][idx_acc]["bufferView"]].update(buffer=7))
    broken(lambda g: g["accessors"][idx_acc].pop("bufferView"))
    broken(lambda g: prim_of(g).update(indices=10 ** 6))


def prim_of(g):
    return g["meshes"][0]["primitives"][0]


def test_png_round_trip(ptb, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    path = str(tmp_path / "x.png")
    ptb.write_png(path, img)
    assert np.array_equal(np.asarray(Image.open(path)), img)


def test_frame_tilings_cover_every_pixel_once(ptb):
    """ptb_frame_tile_layout: rectangles for one rank, comb tiles for several — every pixel in exactly one tile, comb
    tiles equal to within 1 % (that is their point: all tiles cost the same, none is small)."""
    for (W, Hh, world, k) in [(1920, 1080, 1, 8), (1920, 1080, 2, 8), (1920, 1080, 8, 8), (3840, 2160, 8, 8), (1921, 1083, 4, 6),
                              (640, 480, 2, 4), (100, 37, 2, 8)]:
        lay = ptb.frame_tile_layout(W, Hh, 64, world, tiles_in_flight=k)
        cover = np.zeros((Hh, W), np.int32)
        sizes = []
        for row in lay:
            xs, ys = ptb.tile_pixels(row)
            cover[np.ix_(ys, xs)] += 1
            sizes.append(len(xs) * len(ys))
        assert (cover == 1).all(), (W, Hh, world)
        auto = ptb.frame_tile_layout(W, Hh, 64, world)  # in flight chosen by the library: 3..8 tiles of ~4 M paths per GPU
        assert [tuple(r[:4]) for r in auto.tolist()] == ptb.frame_tiles(W, Hh, 64, world)
        if world > 1 and W >= 640:
            assert len(lay) % (world * k) == 0 and (lay[:, 4] > 0).all()  # every stream of every rank: the same count
            assert max(sizes) <= 1.02 * np.mean(sizes), (W, Hh, world, max(sizes) / np.mean(sizes))
        # the plain list is the same tiles without their layout


def test_tile_request_validation_needs_no_gpu(ptb):
    """NULL handles are refused before anything touches CUDA."""
    req = ptb.TileReq(8, 8, 0, 0, 8, 8, 1, 1, 1, 0, 0, 0, 0)
    rgb = np.zeros((8, 8, 3), np.float32)
    st = ptb.lib().ptb_render_tile(None, ctypes.byref(req), rgb.ctypes.data_as(ptb.f32p), None, None)
    assert st == ptb.PTB_E_INVALID
    assert b"NULL" in ptb.lib().ptb_last_error()


def test_worker_request_errors_need_no_gpu(ptb, tmp_path):
    """worker_info parsing and the scene look-up fail before anything touches CUDA."""
    with pytest.raises(ptb.PtbError) as e:
        ptb.worker_run("{ not json", str(tmp_path))
    assert e.value.status == ptb.PTB_E_INVALID
    with pytest.raises(ptb.PtbError) as e:
        ptb.worker_run({"samples": 1, "bounces": 2, "X": 8, "Y": 8}, str(tmp_path / "missing"))
    assert e.value.status == ptb.PTB_E_IO
    with pytest.raises(ptb.PtbError) as e:
        ptb.worker_run({"samples": 1, "bounces": 999, "X": 8, "Y": 8}, str(tmp_path))
    assert e.value.status == ptb.PTB_E_INVALID


# ---- the multi-GPU group's host logic: rendezvous, work-stealing counter, barrier (two processes, no GPU) ----

def _group_rank(name, rank, world, n_tiles, frames, q):
    import importlib
    ptb = importlib.import_module("distributed-path-tracer_b200")
    try:
        q.put((rank, ptb.group_selftest_host(name, rank, world, n_tiles, frames, work_us=200)))
    except Exception as e:  # pragma: no cover - surfaced by the parent
        q.put((rank, e))


def test_group_tile_counter_and_barrier_two_processes(ptb):
    """Two processes rendezvous through shared memory and steal tiles from one counter: over both ranks every
    tile of every frame is claimed exactly once, both ranks get work, frames do not bleed into each other."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"test_{os.getpid()}"
    procs = [ctx.Process(target=_group_rank, args=(name, r, 2, 64, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert isinstance(got[r], np.ndarray), got[r]
    total = got[0].astype(int) + got[1].astype(int)
    assert total.shape == (5, 64) and np.all(total == 1)
    assert got[0].sum() > 0 and got[1].sum() > 0
    assert not os.path.exists("/dev/shm/ptb_" + name)  # rank 0 unlinked the name after the rendezvous


def test_group_times_out_instead_of_hanging(ptb):
    """A rank that never arrives: the others get PTB_E_NCCL after the time-out, nobody hangs."""
    ptb.set_option("group_timeout_ms", 300)
    try:
        name = f"lonely_{os.getpid()}"
        with pytest.raises(ptb.PtbError) as e:
            ptb.group_selftest_host(name, 0, 2, 8, 1)
        assert e.value.status == ptb.PTB_E_NCCL and "timed out" in str(e.value)
    finally:
        ptb.set_option("group_timeout_ms", 120000)
        try:
            os.unlink("/dev/shm/ptb_" + name)
        except OSError:
            pass
