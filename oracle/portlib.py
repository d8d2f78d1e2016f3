"""ctypes binding of oracle/libpt_oracle.so (the plain-C restatement, oracle/pt_oracle.c) — TEST
INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing of the product does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from reflib import FlatScene, SceneDesc, HIT_DTYPE, f32p, u32p, _fp, _up  # same C structs (include/ptb.h)

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libpt_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "pt_oracle.c")
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "port"], check=True, capture_output=True)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO_PATH)
        L.po_scene_create.restype = C.c_void_p
        L.po_scene_create.argtypes = [C.POINTER(SceneDesc)]
        L.po_scene_free.argtypes = [C.c_void_p]
        L.po_dump_kd.restype = C.c_int
        L.po_dump_kd.argtypes = [C.c_void_p, C.c_uint32, u32p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.po_mesh_aabb.argtypes = [C.c_void_p, C.c_uint32, f32p]
        L.po_trace_rays.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_void_p, f32p]
        L.po_count_visits.argtypes = [C.c_void_p, f32p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.po_camera_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u32p, u32p, f32p, C.c_uint64, f32p]
        L.po_render_linear.argtypes = [C.c_void_p] + [C.c_uint32] * 8 + [C.c_int, C.c_int, C.c_uint64, C.c_int,
                                                                        f32p, f32p, C.POINTER(C.c_uint64),
                                                                        C.POINTER(C.c_double)]
        L.po_tonemap_rgba8.argtypes = [f32p, f32p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


class PortScene:
    def __init__(self, flat: FlatScene):
        d, self._keep = flat.to_c()
        self.h = C.c_void_p(lib().po_scene_create(C.byref(d)))
        self.n_meshes = len(flat.meshes)

    def close(self):
        if self.h:
            lib().po_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dump_kd(self, mesh):
        n = C.c_uint64()
        lib().po_dump_kd(self.h, mesh, None, 0, n)
        out = np.empty(n.value, np.uint32)
        assert lib().po_dump_kd(self.h, mesh, _up(out), n.value, n) == 0
        return out

    def mesh_aabb(self, mesh):
        out = np.empty(6, np.float32)
        lib().po_mesh_aabb(self.h, mesh, _fp(out))
        return out

    def trace_rays(self, origin_dir, attrs=False):
        od = np.ascontiguousarray(origin_dir, np.float32).reshape(-1, 6)
        hits = np.zeros(len(od), HIT_DTYPE)
        at = np.zeros((len(od), 14), np.float32) if attrs else None
        lib().po_trace_rays(self.h, _fp(od), len(od), hits.ctypes.data, _fp(at) if attrs else None)
        return (hits, at) if attrs else hits

    def count_visits(self, origin_dir):
        od = np.ascontiguousarray(origin_dir, np.float32).reshape(-1, 6)
        c = (C.c_uint64 * 6)()
        lib().po_count_visits(self.h, _fp(od), len(od), c)
        return dict(zip(("model_tests", "surface_tests", "branch_visits", "leaf_visits", "tri_tests",
                         "stack_pushes"), [int(x) for x in c]))

    def camera_rays(self, w, h, px, py, aa):
        px = np.ascontiguousarray(px, np.uint32)
        py = np.ascontiguousarray(py, np.uint32)
        aa = np.ascontiguousarray(aa, np.float32).reshape(-1, 2)
        od = np.empty((len(px), 6), np.float32)
        lib().po_camera_rays(self.h, w, h, _up(px), _up(py), _fp(aa), len(px), _fp(od))
        return od

    def render_linear(self, full_w, full_h, spp, depth, mode=0, tile=None, first_sample_unjittered=False, seed=1,
                      threads=0):
        x0, y0, w, h = tile if tile else (0, 0, full_w, full_h)
        if threads <= 0:
            threads = os.cpu_count() or 1
        rgb = np.empty((h, w, 3), np.float32)
        alpha = np.empty((h, w), np.float32)
        rays, secs = C.c_uint64(), C.c_double()
        lib().po_render_linear(self.h, full_w, full_h, x0, y0, w, h, spp, depth, mode, int(first_sample_unjittered),
                               seed, threads, _fp(rgb), _fp(alpha), rays, secs)
        return rgb, alpha, rays.value, secs.value


def tonemap_rgba8(rgb, alpha=None):
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
    a = np.ascontiguousarray(alpha, np.float32).reshape(-1) if alpha is not None else None
    out = np.empty((len(rgb), 4), np.uint8)
    lib().po_tonemap_rgba8(_fp(rgb), _fp(a) if a is not None else None, len(rgb), out.ctypes.data)
    return out
