"""Mints the golden fixtures in this directory from the UNMODIFIED reference library
(oracle/_ref/libptref.so, built by `make -C oracle ref` from /root/reference).

    python tests/golden/make_golden.py            # everything (about 5 minutes of CPU)
    python tests/golden/make_golden.py --fast     # skip the converged images

The reference's own tests hold no golden vectors for this path, so these files ARE the
pin: they record what the reference computes, and both the plain-C oracle (oracle/pt_oracle.c) and the
CUDA path are checked against them.  Files:
  cornell_scene.npz       flat export of scenes/cornell-box in renderer::intersect visiting order
  cornell_kd.npz          depth-first record stream of every mesh's KD tree + mesh AABBs
  cornell_rays.npz        camera / random / bounce ray sets with the reference's closest hits and attributes
  heightfield40_kd.npz    KD tree of the n=40 procedural heightfield (3 200 triangles)
  heightfield40_rays.npz  ray sets + hits on that scene
  sun_scene_rays.npz      a small two-instance scene with scaled/rotated transforms: rays + hits
  tonemap.npz             linear rgb(a) → RGBA8 through tonemap_approx_aces + image::write
  cornell_converged_*.npz linear running-mean radiance, 64x64: mode A (renderer::trace, depth 4, 4096 spp)
                          and mode B (restated worker::trace_iter over the reference library, depth 8, 2048 spp)
                          + the per-pixel sample standard deviation measured from independent 64-spp batches
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reflib  # noqa: E402
import ptb200 as ptb  # noqa: E402  (procedural generators only; no GPU use)
from ptb200 import procedural as P  # noqa: E402

FAST = "--fast" in sys.argv
rng = np.random.default_rng(20261018)


def flat_to_npz(flat):
    d = {}
    for i, m in enumerate(flat.meshes):
        for k in ("positions", "normals", "tangents", "uvs", "indices"):
            d[f"mesh{i}_{k}"] = m[k]
    d["n_meshes"] = np.array(len(flat.meshes))
    d["surfaces"] = flat.surfaces
    d["inst_origin"] = np.array([i[0] for i in flat.instances], np.float32)
    d["inst_basis"] = np.array([i[1] for i in flat.instances], np.float32)
    d["inst_range"] = np.array([(i[2], i[3]) for i in flat.instances], np.uint32)
    d["materials"] = np.array([[*m["albedo"], m["opacity"], m["roughness"], m["metallic"], *m["emissive"], m["ior"],
                                float(m.get("shadow_catcher", 0))] for m in flat.materials], np.float32)
    d["material_tex"] = np.array([[m.get(k + "_tex", 0xFFFFFFFF) for k in
                                   ("normal", "albedo", "opacity", "roughness", "metallic", "emissive")]
                                  for m in flat.materials], np.uint32)
    d["n_textures"] = np.array(len(flat.textures))
    for i, t in enumerate(flat.textures):
        d[f"tex{i}_pixels"] = np.ascontiguousarray(t["pixels"])
        d[f"tex{i}_srgb"] = np.array(int(bool(t.get("srgb", False))))
    d["camera_origin"], d["camera_basis"] = flat.camera[0], flat.camera[1]
    d["camera_yfov"] = np.array(flat.camera[2], np.float32)
    if flat.sun is not None:
        d["sun_basis"] = np.array(flat.sun[0], np.float32)
        d["sun_energy"] = np.array(flat.sun[1], np.float32)
        d["sun_angular_radius"] = np.array(flat.sun[2], np.float32)
    d["environment_factor"] = np.array(flat.environment_factor, np.float32)
    d["transparent_background"] = np.array(int(flat.transparent_background))
    return d


def ray_sets(ref, w, h, n_random, box_lo, box_hi):
    ys, xs = np.mgrid[0:h, 0:w]
    px, py = xs.ravel().astype(np.uint32), ys.ravel().astype(np.uint32)
    aa = rng.random((w * h, 2), dtype=np.float32)
    aa[: w * h // 2] = 0.5  # half pixel centres, half jittered
    cam = ref.camera_rays(w, h, px, py, aa)
    hits_cam, attrs_cam = ref.trace_rays(cam, attrs=True)
    o = rng.uniform(box_lo, box_hi, (n_random, 3)).astype(np.float32)
    d = rng.normal(size=(n_random, 3)).astype(np.float32)
    d[: n_random // 8] *= np.float32(37.5)      # un-normalised directions: the ray ctor must normalise
    d[n_random // 8: n_random // 4, rng.integers(0, 3)] = 0.0  # axis-parallel components (inf / nan slabs)
    rnd = np.concatenate([o, d], 1)
    hits_rnd, attrs_rnd = ref.trace_rays(rnd, attrs=True)
    hit = hits_cam["instance"] != 0xFFFFFFFF
    pos = attrs_cam[hit, 0:3]
    dirs = rng.normal(size=pos.shape).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    bounce = np.concatenate([pos + dirs * np.float32(1e-4), dirs], 1).astype(np.float32)
    hits_b, attrs_b = ref.trace_rays(bounce, attrs=True)
    return dict(cam_px=px, cam_py=py, cam_aa=aa, cam_res=np.array([w, h], np.uint32), cam_rays=cam,
                cam_hits=hits_cam, cam_attrs=attrs_cam, rnd_rays=rnd, rnd_hits=hits_rnd, rnd_attrs=attrs_rnd,
                bounce_rays=bounce, bounce_hits=hits_b, bounce_attrs=attrs_b,
                visits_cam=np.array(list(ref.count_visits(cam).values()), np.uint64))


def main():
    gltf = os.path.join(HERE, "scenes", "cornell-box", "cornell.gltf")
    ref = reflib.RefScene.from_gltf(gltf)
    flat = ref.export_flat()
    np.savez_compressed(os.path.join(HERE, "cornell_scene.npz"), **flat_to_npz(flat),
                        model_aabbs=flat.model_aabbs)
    kd = {f"mesh{i}": ref.dump_kd(i) for i in range(len(flat.meshes))}
    kd["mesh_aabbs"] = np.array(flat.mesh_aabbs, np.float32)
    np.savez_compressed(os.path.join(HERE, "cornell_kd.npz"), **kd)
    np.savez_compressed(os.path.join(HERE, "cornell_rays.npz"),
                        **ray_sets(ref, 64, 64, 8192, (-3.5, -1.5, -3.5), (3.5, 6.0, 12.0)))

    # --- heightfield n=40
    sc = P.heightfield_scene(40)
    fs = reflib.FlatScene(sc.meshes, sc.surfaces, sc.instances, sc.materials, sc.camera)
    hf = reflib.RefScene.from_flat(fs)
    np.savez_compressed(os.path.join(HERE, "heightfield40_kd.npz"), mesh0=hf.dump_kd(0), mesh1=hf.dump_kd(1))
    np.savez_compressed(os.path.join(HERE, "heightfield40_rays.npz"),
                        **ray_sets(hf, 96, 54, 4096, (-6, -1, -6), (6, 5, 9)))

    # --- transformed instances + sun (scale, rotation, instancing of one mesh)
    base = P.heightfield_mesh(12, 1.0, 99)
    quad = P.quad_lights_mesh(height=0.0, half=4.0, centres=((0.0, 0.0),))
    quad["normals"][:] = (0, 1, 0)
    def rot_y(a, s):
        c, sn = np.cos(a), np.sin(a)
        return np.array([c * s[0], 0, -sn * s[0], 0, s[1], 0, sn * s[2], 0, c * s[2]], np.float32)
    insts = [((0.0, -0.2, 0.0), P.IDENTITY, 1, 1),
             ((-1.5, 0.3, 0.5), rot_y(0.7, (1.0, 2.0, 1.0)), 0, 1),
             ((1.6, 0.1, -0.4), rot_y(-1.1, (0.6, 0.6, 1.7)), 0, 1),
             ((0.0, 1.4, -2.5), rot_y(2.0, (1.3, 0.5, 0.8)), 0, 1)]
    mats = [dict(albedo=(0.7, 0.5, 0.3), opacity=1.0, roughness=0.6, metallic=0.0, emissive=(0, 0, 0), ior=1.33),
            dict(albedo=(0.6, 0.6, 0.6), opacity=1.0, roughness=0.9, metallic=0.0, emissive=(0, 0, 0), ior=1.33)]
    cam_o, cam_b = P.look_at((0.5, 3.0, 6.0), (0, 0.3, 0))
    sun_o, sun_b = P.look_at((1.0, 2.0, 0.7), (0, 0, 0))  # light comes from +basis.z
    sun_scene = reflib.FlatScene([base, quad], [(0, 0), (1, 1)], insts, mats, (cam_o, cam_b, 0.7),
                                 sun=(sun_b, (3.0, 2.8, 2.5), 0.004732), environment_factor=(0.3, 0.4, 0.6))
    sr = reflib.RefScene.from_flat(sun_scene)
    d = flat_to_npz(sun_scene)
    d.update(ray_sets(sr, 64, 48, 4096, (-4, -1, -4), (4, 4, 7)))
    np.savez_compressed(os.path.join(HERE, "sun_scene_rays.npz"), **d)

    # --- jack-of-blades: the reference's organic fixture (58 740 triangles, 7 meshes, 12-14x leaf duplication).
    # Geometry only travels (2 MB); its 20 MB of textures stay in the reference tree, texture parity is covered by
    # the synthetic scene below and, where /root/reference exists, by tests/test_host.py on the real files.
    jack_gltf = "/root/reference/path-tracer-core/scenes/jack-of-blades/jack-of-blades.gltf"
    if os.path.exists(jack_gltf):
        jf = reflib.RefScene.from_gltf(jack_gltf).export_flat()
        for m in jf.materials:
            for k in ("normal", "albedo", "opacity", "roughness", "metallic", "emissive"):
                m[k + "_tex"] = 0xFFFFFFFF
        jf.textures = []
        jr = reflib.RefScene.from_flat(jf)
        d = flat_to_npz(jf)
        d.update(ray_sets(jr, 96, 54, 8192, (-3, -1, -3), (3, 5, 6)))
        d["kd_words"] = np.array([len(jr.dump_kd(i)) for i in range(len(jf.meshes))], np.uint64)
        d["kd_crc"] = np.array([int(np.bitwise_xor.reduce(jr.dump_kd(i) * np.arange(1, len(jr.dump_kd(i)) + 1, dtype=np.uint32)))
                                for i in range(len(jf.meshes))], np.uint64)
        np.savez_compressed(os.path.join(HERE, "jack_geometry_rays.npz"), **d)

    # --- synthetic textured scene: every material slot, sRGB + linear + float textures, alpha-tested opacity,
    # normal mapping, a scaled/rotated second instance, a sun
    trng = np.random.default_rng(77)
    def blobs(h, w, c, lo=0, hi=256):
        base = trng.integers(lo, hi, (h // 4 + 1, w // 4 + 1, c)).astype(np.float32)
        img = np.kron(base, np.ones((4, 4, 1), np.float32))[:h, :w]
        return np.clip(img + trng.normal(0, 6, (h, w, c)), 0, 255).astype(np.uint8)
    albedo = blobs(32, 32, 4, 40, 256)
    albedo[..., 3] = np.where(trng.random((32, 32)) < 0.25, 90, 255).astype(np.uint8)   # holes: opacity 0.35
    nrm = blobs(16, 16, 3, 96, 160)
    nrm[..., 2] = 255
    rough_metal = blobs(8, 8, 3, 30, 220)
    emis = (trng.random((8, 8, 3)) ** 6).astype(np.float32) * np.float32(0.6)                 # float texture
    textures = [dict(pixels=albedo, srgb=True), dict(pixels=nrm, srgb=False), dict(pixels=rough_metal, srgb=False),
                dict(pixels=emis, srgb=False), dict(pixels=blobs(20, 12, 3, 0, 256), srgb=True)]
    ground = P.heightfield_mesh(10, 2.0, 5)
    ground["positions"][:, 1] *= 0.35
    NT = 0xFFFFFFFF
    tmats = [dict(albedo=(0.9, 0.9, 0.9), opacity=1.0, roughness=0.9, metallic=0.8, emissive=(0.0, 0.0, 0.0), ior=1.33,
                  normal_tex=1, albedo_tex=0, opacity_tex=0, roughness_tex=2, metallic_tex=2, emissive_tex=NT),
             dict(albedo=(0.7, 0.8, 1.0), opacity=0.6, roughness=0.4, metallic=0.1, emissive=(1.0, 0.9, 0.8), ior=1.33,
                  normal_tex=NT, albedo_tex=4, opacity_tex=NT, roughness_tex=NT, metallic_tex=NT, emissive_tex=3)]
    tinst = [((0.0, 0.0, 0.0), P.IDENTITY, 0, 1),
             ((0.3, 0.9, -0.6), rot_y(0.9, (0.45, 1.6, 0.45)), 1, 1)]
    tcam_o, tcam_b = P.look_at((0.4, 2.2, 3.4), (0.05, 0.2, 0))
    tsun_o, tsun_b = P.look_at((0.8, 2.0, 1.1), (0, 0, 0))
    tscene = reflib.FlatScene([ground], [(0, 0), (0, 1)], tinst, tmats, (tcam_o, tcam_b, 0.75),
                              sun=(tsun_b, (2.5, 2.3, 2.0), 0.004732), environment_factor=(0.5, 0.6, 0.8),
                              textures=textures)
    tr = reflib.RefScene.from_flat(tscene)
    d = flat_to_npz(tscene)
    d.update(ray_sets(tr, 64, 48, 4096, (-2.5, -0.5, -2.5), (2.5, 3, 3.5)))
    for name, mode, depth in (("A", 0, 4), ("B", 1, 6)):
        acc = np.zeros((48, 64, 3), np.float64); acc2 = np.zeros((48, 64, 3), np.float64); nb = 24
        for b in range(nb):
            rgb64, _, _, _ = tr.render_linear(64, 48, 64, depth, mode=mode)
            acc += rgb64; acc2 += rgb64.astype(np.float64) ** 2
        mean = acc / nb
        d[f"converged_{name}_mean"] = mean.astype(np.float32)
        d[f"converged_{name}_sigma"] = np.sqrt(np.maximum(acc2 / nb - mean ** 2, 0) * nb / (nb - 1) * 64).astype(np.float32)
        d[f"converged_{name}_spp"] = np.array(64 * nb)
        d[f"converged_{name}_depth"] = np.array(depth)
    np.savez_compressed(os.path.join(HERE, "textured_scene_rays.npz"), **d)

    # --- tonemap
    rgb = np.concatenate([rng.random((4096, 3), dtype=np.float32) * np.float32(4.0),
                          np.linspace(0, 12, 4096 * 3, dtype=np.float32).reshape(-1, 3),
                          np.zeros((4, 3), np.float32)])
    alpha = rng.random(len(rgb), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "tonemap.npz"), rgb=rgb, alpha=alpha,
                        rgba8=reflib.tonemap_rgba8(rgb, alpha))

    if FAST:
        return
    # --- converged linear images (statistical parity)
    for name, mode, depth, batches in (("A", 0, 4, 64), ("B", 1, 8, 32)):
        acc = np.zeros((64, 64, 3), np.float64)
        acc2 = np.zeros((64, 64, 3), np.float64)
        for b in range(batches):
            rgb64, _, _, secs = ref.render_linear(64, 64, 64, depth, mode=mode)
            acc += rgb64
            acc2 += rgb64.astype(np.float64) ** 2
            print(f"mode {name} batch {b + 1}/{batches} {secs:.1f}s", flush=True)
        mean = acc / batches
        var_batch = np.maximum(acc2 / batches - mean ** 2, 0) * batches / (batches - 1)  # variance of a 64-spp mean
        np.savez_compressed(os.path.join(HERE, f"cornell_converged_{name}.npz"), mean=mean.astype(np.float32),
                            sigma_per_sample=np.sqrt(var_batch * 64).astype(np.float32),
                            spp=np.array(64 * batches), depth=np.array(depth), mode=np.array(mode))


if __name__ == "__main__":
    main()
