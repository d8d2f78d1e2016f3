"""GPU (B200): BASELINE.json's configurations at FULL size against the plain-C oracle (oracle/pt_oracle.c, itself
pinned bit for bit to the unmodified reference library on the golden fixtures).

  C2  999 698-triangle heightfield (the headline scene): the KD tree the product builds is word-identical to the
      oracle's (LIB/core/mesh.cpp:131-247), and > 200 000 camera / bounce / edge-case rays return bit-identical hits
      (LIB/core/mesh.cpp:300-405, LIB/geometry/triangle.cpp:120-190, LIB/scene/model.cpp:20-72).
  C5  49 instances of that mesh + the light quads (49 M instanced triangles): the same, through the instance loop
      with its conservative culling (LIB/core/renderer.cpp:645-675).
  C3  depth 16 + Russian roulette: tests/test_gpu_parity.py::test_image_statistics_against_converged_reference[B16].
The oracle needs ~10 s per scene on the host for the tree; the rays take a few seconds."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _ray_set(scene, w, h, rng, n_random, box):
    """Camera rays of a w x h frame, one diffuse bounce ray from every camera hit, random rays in `box`, and the
    edge-case set of test_large_mesh_against_c_oracle (axis-parallel, tiny / huge / zero directions, far origins)."""
    ys, xs = np.mgrid[0:h, 0:w]
    cam = scene.camera_rays(w, h, xs.ravel(), ys.ravel(), rng.random((w * h, 2), dtype=np.float32))
    hits = scene.trace_rays(cam)
    ok = hits["instance"] != 0xFFFFFFFF
    d = cam[ok, 3:] / np.linalg.norm(cam[ok, 3:], axis=1, keepdims=True)
    p = cam[ok, :3] + d * hits["t"][ok, None]
    bd = rng.normal(size=p.shape).astype(np.float32)
    bd[:, 1] = np.abs(bd[:, 1])  # upper hemisphere of a terrain: most leave, many graze it
    bounce = np.concatenate([p + bd * np.float32(1e-4), bd], 1).astype(np.float32)
    lo, hi = np.array(box[0], np.float32), np.array(box[1], np.float32)
    o = (rng.random((n_random, 3), dtype=np.float32) * (hi - lo) + lo).astype(np.float32)
    dd = rng.normal(size=(n_random, 3)).astype(np.float32)
    k = n_random // 40
    dd[:k, 0] = 0
    dd[k:2 * k, 1] = 0
    dd[2 * k:3 * k] = (0, -1, 0)
    dd[3 * k:4 * k] *= np.float32(1e-3)
    o[4 * k:5 * k] *= np.float32(40)
    dd[5 * k:5 * k + 10] = 0  # degenerate direction: NaN everywhere, a miss on both sides
    rnd = np.concatenate([o, dd], 1)
    return dict(cam=cam, bounce=bounce, rnd=rnd)


def test_c2_full_size_tree_and_hits_against_the_oracle(ptb, procedural, portlib, reflib):
    desc = procedural.heightfield_scene(707)
    rng = np.random.default_rng(707)
    port = portlib.PortScene(reflib.FlatScene(desc.meshes, desc.surfaces, desc.instances, desc.materials, desc.camera))
    with ptb.Scene.create(desc) as s:
        assert s.info()["n_triangles"] == 999_698 + 8
        for m in range(2):
            got, want = s.dump_kd(m), port.dump_kd(m)
            assert len(got) == len(want) and np.array_equal(got, want), f"C2 mesh {m}: KD tree differs from the oracle's"
        rays = _ray_set(s, 480, 270, rng, 120_000, ((-6, -1, -6), (6, 4, 9)))
        total = 0
        for key, od in rays.items():
            H.assert_hits_equal(s.trace_rays(od), port.trace_rays(od), "C2 full size:" + key)
            occ = s.trace_occlusion(od)
            assert np.array_equal(occ, port.trace_rays(od)["instance"] != 0xFFFFFFFF), key
            total += len(od)
        assert total > 200_000
        frac = (s.trace_rays(rays["cam"])["instance"] != 0xFFFFFFFF).mean()
        assert frac > 0.99  # the terrain fills the frame


def test_c5_full_size_instances_against_the_oracle(ptb, procedural, portlib, reflib):
    desc = procedural.instanced_heightfield_scene(707, 7)
    assert len(desc.instances) == 50
    rng = np.random.default_rng(505)
    port = portlib.PortScene(reflib.FlatScene(desc.meshes, desc.surfaces, desc.instances, desc.materials, desc.camera))
    with ptb.Scene.create(desc) as s:
        info = s.info()
        assert info["n_instances"] == 50 and info["n_triangles"] == 999_698 + 8
        assert np.array_equal(s.dump_kd(0), port.dump_kd(0))
        rays = _ray_set(s, 480, 270, rng, 120_000, ((-38, -1, -38), (38, 6, 42)))
        total = 0
        for key, od in rays.items():
            got, want = s.trace_rays(od), port.trace_rays(od)
            H.assert_hits_equal(got, want, "C5 full size:" + key)
            total += len(od)
        assert total > 200_000
        hit_inst = port.trace_rays(rays["cam"])["instance"]
        assert len(np.unique(hit_inst[hit_inst != 0xFFFFFFFF])) > 20  # the camera sees most of the 7 x 7 tiles
