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
               textures=parts.get("textures", ()))


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
