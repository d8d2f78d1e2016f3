"""Shared helpers of the test-suite (fixture loading, comparisons)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESH_KEYS = ("positions", "normals", "tangents", "uvs", "indices")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def scene_parts_from_npz(z):
    meshes = [{k: z[f"mesh{i}_{k}"] for k in MESH_KEYS} for i in range(int(z["n_meshes"]))]
    inst = [(z["inst_origin"][i], z["inst_basis"][i], int(z["inst_range"][i][0]), int(z["inst_range"][i][1]))
            for i in range(len(z["inst_origin"]))]
    mats = [dict(albedo=tuple(m[0:3]), opacity=float(m[3]), roughness=float(m[4]), metallic=float(m[5]),
                 emissive=tuple(m[6:9]), ior=float(m[9]), shadow_catcher=int(m[10])) for m in z["materials"]]
    if "material_tex" in z:
        for m, ids in zip(mats, z["material_tex"]):
            for k, v in zip(("normal", "albedo", "opacity", "roughness", "metallic", "emissive"), ids):
                m[k + "_tex"] = int(v)
    textures = [dict(pixels=z[f"tex{i}_pixels"], srgb=bool(z[f"tex{i}_srgb"]))
                for i in range(int(z["n_textures"]))] if "n_textures" in z else []
    sun = (z["sun_basis"], z["sun_energy"], float(z["sun_angular_radius"])) if "sun_basis" in z else None
    cam = (z["camera_origin"], z["camera_basis"], float(z["camera_yfov"]))
    return dict(meshes=meshes, surfaces=z["surfaces"], instances=inst, materials=mats, camera=cam, sun=sun,
                environment_factor=tuple(float(x) for x in z["environment_factor"]),
                transparent_background=bool(z["transparent_background"]), textures=textures)


def make_flat(cls, parts):
    """cls: reflib.FlatScene or ptb.SceneDescription (same constructor)."""
    return cls(parts["meshes"], parts["surfaces"], parts["instances"], parts["materials"], parts["camera"],
               parts["sun"], parts["environment_factor"], parts["transparent_background"],
               textures=parts.get("textures", ()), environment_texture=parts.get("environment_texture"))


def environment_scene_parts(procedural):
    """A small heightfield under an equirectangular float environment map (renderer.hpp:28 `environment`): the camera
    looks along the horizon, so about half of the primary rays and most bounce rays end in the map."""
    mesh = procedural.heightfield_mesh(12, 2.0, 7)
    h, w = 32, 64
    v, u = np.mgrid[0:h, 0:w]
    env = np.zeros((h, w, 3), np.float32)
    env[..., 0] = 0.2 + 0.8 * u / (w - 1)
    env[..., 1] = 0.1 + 0.9 * v / (h - 1)
    env[..., 2] = 0.5
    env[4:9, 40:48] = (6.0, 5.0, 3.0)  # a bright patch: an image-based light
    mats = [dict(albedo=(0.75, 0.7, 0.6), opacity=1.0, roughness=0.6, metallic=0.1, emissive=(0, 0, 0), ior=1.33,
                 shadow_catcher=0)]
    cam = procedural.look_at((0.2, 1.1, 3.2), (0.0, 0.6, 0.0))
    return dict(meshes=[mesh], surfaces=np.array([[0, 0]], np.uint32),
                instances=[((0, 0, 0), np.eye(3, dtype=np.float32).ravel(), 0, 1)], materials=mats,
                camera=(cam[0], cam[1], 0.9), sun=None, environment_factor=(0.9, 1.0, 1.1),
                transparent_background=False, textures=[dict(pixels=env, srgb=False)], environment_texture=0)


def block_mean_agreement(a, b, scale=5.0, rel=0.01):
    """Two noisy renders of the same 64x48 image: do their means agree within the pooled standard error
    (estimated from 8x8 blocks)?  → (ok, difference of the means, standard error)"""
    def block_means(x):
        return x.reshape(6, 8, 8, 8, 3).mean((1, 3)).reshape(-1, 3)
    diff = block_means(a) - block_means(b)
    se = diff.std(0) / np.sqrt(len(diff)) + 1e-4
    ok = bool(np.all(np.abs(diff.mean(0)) < scale * se + rel * np.abs(b.mean((0, 1)))))
    return ok, diff.mean(0), se


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_hits_equal(got, want, what=""):
    """Bit-exact on ids, distance and barycentrics."""
    for f in ("instance", "surface", "triangle"):
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"{what}: {f} differs at {bad[:5]} ({bad.size} rays)"
    assert np.array_equal(bits(got["t"]), bits(want["t"])), f"{what}: distances are not bit-identical"
    assert np.array_equal(bits(got["bary"]), bits(want["bary"])), f"{what}: barycentrics are not bit-identical"


def mean_z(rgb, conv, spp):
    """z-score per channel of the image mean against the converged reference image."""
    se = conv["sigma_per_sample"].astype(np.float64) / np.sqrt(float(spp))
    n = rgb.shape[0] * rgb.shape[1]
    return (rgb.astype(np.float64).mean((0, 1)) - conv["mean"].astype(np.float64).mean((0, 1))) / (
        np.sqrt((se ** 2).sum((0, 1))) / n)
