"""Deterministic procedural scenes for the benchmark configurations of BASELINE.json.

``heightfield_scene`` is the "synthetic 1M-triangle procedural mesh" (config 2/4): the generator
that SURVEY.md §8(d) fixes — grid n=707 → 2·n² = 999 698 triangles, x,z ∈ [-5,5],
y = 0.5·sin(3x)·cos(2z) + 0.02·u with u from the LCG s ← s·1664525 + 1013904223 (seed 12345,
u = (s >> 8) / 2²⁴) — plus a few emissive quads above it and a fixed camera looking down at the terrain (~43°,
chosen so that the mesh fills the 16:9 frame: 99.6 % of the primary rays hit it).
The arrays are produced once here and handed, identical, to libptb and (in tests / the CPU
baseline) to the reference library, so both sides trace exactly the same geometry.
"""
from __future__ import annotations

import os

import numpy as np

from . import SceneDescription


def _lcg_uniform(count: int, seed: int = 12345) -> np.ndarray:
    """u_k = (s_k >> 8) / 2^24 for s_k = s_{k-1}*1664525 + 1013904223 mod 2^32, vectorised by doubling."""
    a, c = np.uint64(1664525), np.uint64(1013904223)
    mask = np.uint64(0xFFFFFFFF)
    out = np.empty(count, np.uint64)
    if count == 0:
        return out.astype(np.float32)
    out[0] = (np.uint64(seed) * a + c) & mask
    have = 1
    # s_{k+m} = A_m * s_k + C_m ; (A_2m, C_2m) = (A_m^2, A_m*C_m + C_m)
    A, Cc = a, c
    while have < count:
        take = min(have, count - have)
        out[have:have + take] = (out[:take] * A + Cc) & mask
        have += take
        Cc = (A * Cc + Cc) & mask
        A = (A * A) & mask
    return ((out >> np.uint64(8)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def look_at(eye, target, up=(0.0, 1.0, 0.0)):
    """Camera transform (origin, basis[9] column-major x,y,z); the camera looks down its local -Z."""
    eye = np.asarray(eye, np.float64)
    back = eye - np.asarray(target, np.float64)
    back /= np.linalg.norm(back)
    right = np.cross(np.asarray(up, np.float64), back)
    right /= np.linalg.norm(right)
    upv = np.cross(back, right)
    return eye.astype(np.float32), np.concatenate([right, upv, back]).astype(np.float32)


IDENTITY = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)


def heightfield_mesh(n: int = 707, extent: float = 5.0, seed: int = 12345):
    """(n+1)^2 vertices, 2*n^2 triangles, row-major over (z, x)."""
    lin = np.linspace(-extent, extent, n + 1, dtype=np.float64)
    x, z = np.meshgrid(lin, lin)  # x varies fastest
    u = _lcg_uniform((n + 1) * (n + 1), seed).astype(np.float64).reshape(n + 1, n + 1)
    y = 0.5 * np.sin(3.0 * x) * np.cos(2.0 * z) + 0.02 * u
    pos = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32)
    dydx = 1.5 * np.cos(3.0 * x) * np.cos(2.0 * z)
    dydz = -1.0 * np.sin(3.0 * x) * np.sin(2.0 * z)
    nrm = np.stack([-dydx, np.ones_like(x), -dydz], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    tan = np.stack([np.ones_like(x), dydx, np.zeros_like(x)], -1)
    tan /= np.linalg.norm(tan, axis=-1, keepdims=True)
    uv = np.stack([(x + extent) / (2 * extent), (z + extent) / (2 * extent)], -1)
    j, i = np.meshgrid(np.arange(n), np.arange(n))  # i: row (z), j: column (x)
    v00 = (i * (n + 1) + j).astype(np.uint32)
    v01, v10, v11 = v00 + 1, v00 + (n + 1), v00 + (n + 2)
    tris = np.stack([np.stack([v00, v10, v01], -1), np.stack([v01, v10, v11], -1)], 2).reshape(-1, 3)
    return dict(positions=pos, normals=nrm.reshape(-1, 3).astype(np.float32),
                tangents=tan.reshape(-1, 3).astype(np.float32), uvs=uv.reshape(-1, 2).astype(np.float32),
                indices=np.ascontiguousarray(tris, np.uint32))


def quad_lights_mesh(height: float = 3.0, half: float = 0.75, centres=((-2.5, -2.5), (2.5, -2.5), (-2.5, 2.5), (2.5, 2.5))):
    """Downward-facing emissive quads above the terrain (two triangles each)."""
    pos, idx = [], []
    for k, (cx, cz) in enumerate(centres):
        b = 4 * k
        pos += [(cx - half, height, cz - half), (cx + half, height, cz - half),
                (cx + half, height, cz + half), (cx - half, height, cz + half)]
        idx += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    pos = np.array(pos, np.float32)
    nv = len(pos)
    return dict(positions=pos, normals=np.tile(np.array([[0, -1, 0]], np.float32), (nv, 1)),
                tangents=np.tile(np.array([[1, 0, 0]], np.float32), (nv, 1)),
                uvs=np.zeros((nv, 2), np.float32), indices=np.array(idx, np.uint32))


def heightfield_scene(n: int = 707, seed: int = 12345) -> SceneDescription:
    """Config 2/4 of BASELINE.json: ~2*n^2 diffuse triangles + 8 emissive ones, white environment."""
    terrain = heightfield_mesh(n, 5.0, seed)
    lights = quad_lights_mesh()
    materials = [
        dict(albedo=(0.8, 0.8, 0.8), opacity=1.0, roughness=1.0, metallic=0.0, emissive=(0, 0, 0), ior=1.33),
        dict(albedo=(0.8, 0.8, 0.8), opacity=1.0, roughness=1.0, metallic=0.0, emissive=(1, 1, 1), ior=1.33),
    ]
    # Looking down at ~43 degrees; the terrain fills 99.3 % of a 16:9 frame (measured on primary rays).
    # The eye is deliberately NOT at x = 0: the reference's traversal takes only the far child when a ray
    # starts exactly on a split plane (split_dist = 0 < tmin, mesh.cpp:354-360), and the SAH root plane of
    # this mesh is x = 0, so with the eye on it every ray heading to +x is reported as a miss — by the
    # reference and, bit for bit, by this implementation (seen with n = 707; see DESIGN.md "reference quirks").
    cam_o, cam_b = look_at((0.31, 3.4, 3.9), (0.13, 0.0, 0.2))
    return SceneDescription(
        meshes=[terrain, lights], surfaces=[(0, 0), (1, 1)],
        instances=[((0, 0, 0), IDENTITY, 0, 1), ((0, 0, 0), IDENTITY, 1, 1)],
        materials=materials, camera=(cam_o, cam_b, 0.7), sun=None, environment_factor=(1.0, 1.0, 1.0))


def instanced_heightfield_scene(n: int = 707, grid: int = 7, seed: int = 12345) -> SceneDescription:
    """Config 5: grid x grid instances of one heightfield tile (grid=7, n=707 → 49 M instanced triangles).

    Instances share one surface range — the reference's own model/transform mechanism."""
    terrain = heightfield_mesh(n, 5.0, seed)
    lights = quad_lights_mesh(height=3.0)
    materials = [
        dict(albedo=(0.8, 0.8, 0.8), opacity=1.0, roughness=1.0, metallic=0.0, emissive=(0, 0, 0), ior=1.33),
        dict(albedo=(0.8, 0.8, 0.8), opacity=1.0, roughness=1.0, metallic=0.0, emissive=(1, 1, 1), ior=1.33),
    ]
    instances = []
    half = (grid - 1) / 2.0
    for gz in range(grid):
        for gx in range(grid):
            instances.append((((gx - half) * 10.0, 0.0, (gz - half) * 10.0), IDENTITY, 0, 1))
    instances.append(((0, 0, 0), IDENTITY, 1, 1))
    span = 10.0 * grid
    cam_o, cam_b = look_at((0.0, 0.12 * span + 4.0, 0.5 * span + 3.0), (0.0, 0.0, 0.0))
    return SceneDescription(meshes=[terrain, lights], surfaces=[(0, 0), (1, 1)], instances=instances,
                            materials=materials, camera=(cam_o, cam_b, 0.8), sun=None,
                            environment_factor=(1.0, 1.0, 1.0))


def cornell_gltf_path() -> str:
    """The reference's bundled Cornell box (scenes/cornell-box), kept as a fixture."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(root, "tests", "golden", "scenes", "cornell-box", "cornell.gltf")
