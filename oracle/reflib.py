"""ctypes binding of oracle/_ref/libptref.so — TEST INFRASTRUCTURE ONLY.

The shared object is the UNMODIFIED reference library
(vmanam0451/distributed-path-tracer, path-tracer-core/path_tracer_lib) compiled by
oracle/Makefile plus oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "_ref", "libptref.so")

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)


class MeshDesc(C.Structure):
    _fields_ = [("positions", f32p), ("normals", f32p), ("tangents", f32p), ("uvs", f32p),
                ("n_vertices", C.c_uint32), ("indices", u32p), ("n_triangles", C.c_uint32)]


class SurfaceDesc(C.Structure):
    _fields_ = [("mesh", C.c_uint32), ("material", C.c_uint32)]


class InstanceDesc(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("basis", C.c_float * 9),
                ("first_surface", C.c_uint32), ("n_surfaces", C.c_uint32)]


class TextureDesc(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32),
                ("channels", C.c_uint32), ("is_float", C.c_uint32), ("srgb", C.c_uint32)]


class MaterialDesc(C.Structure):
    _fields_ = [("albedo", C.c_float * 3), ("opacity", C.c_float), ("roughness", C.c_float),
                ("metallic", C.c_float), ("emissive", C.c_float * 3), ("ior", C.c_float),
                ("shadow_catcher", C.c_uint32),
                ("normal_tex", C.c_uint32), ("albedo_tex", C.c_uint32), ("opacity_tex", C.c_uint32),
                ("roughness_tex", C.c_uint32), ("metallic_tex", C.c_uint32), ("emissive_tex", C.c_uint32)]


class CameraDesc(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("basis", C.c_float * 9), ("yfov", C.c_float)]


class SunDesc(C.Structure):
    _fields_ = [("enabled", C.c_uint32), ("basis", C.c_float * 9), ("energy", C.c_float * 3),
                ("angular_radius", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("meshes", C.POINTER(MeshDesc)), ("n_meshes", C.c_uint32),
                ("surfaces", C.POINTER(SurfaceDesc)), ("n_surfaces", C.c_uint32),
                ("instances", C.POINTER(InstanceDesc)), ("n_instances", C.c_uint32),
                ("materials", C.POINTER(MaterialDesc)), ("n_materials", C.c_uint32),
                ("textures", C.POINTER(TextureDesc)), ("n_textures", C.c_uint32),
                ("camera", CameraDesc), ("sun", SunDesc),
                ("environment_factor", C.c_float * 3), ("transparent_background", C.c_uint32),
                ("kd_use_sah", C.c_uint32), ("kd_max_depth", C.c_uint32), ("environment_tex_plus1", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("instance", C.c_uint32), ("surface", C.c_uint32), ("triangle", C.c_uint32),
                ("t", C.c_float), ("bary", C.c_float * 3)]


HIT_DTYPE = np.dtype([("instance", "<u4"), ("surface", "<u4"), ("triangle", "<u4"),
                      ("t", "<f4"), ("bary", "<f4", (3,))])
assert HIT_DTYPE.itemsize == C.sizeof(Hit) == 28


def _fp(a):
    return a.ctypes.data_as(f32p)


def _up(a):
    return a.ctypes.data_as(u32p)


class FlatScene:
    """Plain-numpy scene description shared by the reference harness and libptb.

    meshes: list of dicts(positions[nv,3], normals[nv,3], tangents[nv,3], uvs[nv,2], indices[nt,3])
    surfaces: uint32[ns,2] (mesh, material); instances: list of (origin[3], basis[9], first, count)
    materials: list of dicts(albedo, opacity, roughness, metallic, emissive, ior, shadow_catcher)
    camera: (origin[3], basis[9], yfov); sun: None | (basis[9], energy[3], angular_radius)
    """

    def __init__(self, meshes, surfaces, instances, materials, camera, sun=None,
                 environment_factor=(1.0, 1.0, 1.0), transparent_background=False,
                 kd_use_sah=True, kd_max_depth=25, textures=(), environment_texture=None):
        self.environment_texture = None if environment_texture is None else int(environment_texture)
        self.meshes = [
            {k: np.ascontiguousarray(m[k], dtype=np.uint32 if k == "indices" else np.float32)
             for k in ("positions", "normals", "tangents", "uvs", "indices")} for m in meshes]
        self.surfaces = np.ascontiguousarray(surfaces, dtype=np.uint32).reshape(-1, 2)
        self.instances = [(np.asarray(o, np.float32), np.asarray(b, np.float32), int(f), int(c))
                          for (o, b, f, c) in instances]
        self.materials = [dict(m) for m in materials]
        self.camera = (np.asarray(camera[0], np.float32), np.asarray(camera[1], np.float32), float(camera[2]))
        self.sun = sun
        self.environment_factor = tuple(float(x) for x in environment_factor)
        self.transparent_background = bool(transparent_background)
        self.kd_use_sah = bool(kd_use_sah)
        self.kd_max_depth = int(kd_max_depth)
        self.textures = list(textures)  # dicts(pixels ndarray[h,w,c], srgb)

    @property
    def n_triangles(self):
        return sum(len(m["indices"]) for m in self.meshes)

    def to_c(self):
        """Returns (SceneDesc, keepalive) — keepalive must outlive every use of the desc."""
        keep = []
        md = (MeshDesc * len(self.meshes))()
        for i, m in enumerate(self.meshes):
            md[i].positions = _fp(m["positions"])
            md[i].normals = _fp(m["normals"])
            md[i].tangents = _fp(m["tangents"])
            md[i].uvs = _fp(m["uvs"])
            md[i].n_vertices = len(m["positions"])
            md[i].indices = _up(m["indices"])
            md[i].n_triangles = len(m["indices"])
        sd = (SurfaceDesc * len(self.surfaces))()
        for i, (me, ma) in enumerate(self.surfaces):
            sd[i].mesh, sd[i].material = int(me), int(ma)
        idesc = (InstanceDesc * len(self.instances))()
        for i, (o, b, f, c) in enumerate(self.instances):
            idesc[i].origin[:] = [float(x) for x in o]
            idesc[i].basis[:] = [float(x) for x in b]
            idesc[i].first_surface, idesc[i].n_surfaces = f, c
        mat = (MaterialDesc * len(self.materials))()
        for i, m in enumerate(self.materials):
            mat[i].albedo[:] = [float(x) for x in m.get("albedo", (1, 1, 1))]
            mat[i].opacity = float(m.get("opacity", 1.0))
            mat[i].roughness = float(m.get("roughness", 1.0))
            mat[i].metallic = float(m.get("metallic", 1.0))
            mat[i].emissive[:] = [float(x) for x in m.get("emissive", (0, 0, 0))]
            mat[i].ior = float(m.get("ior", 1.33))
            mat[i].shadow_catcher = int(m.get("shadow_catcher", 0))
            for slot in ("normal", "albedo", "opacity", "roughness", "metallic", "emissive"):
                setattr(mat[i], slot + "_tex", int(m.get(slot + "_tex", 0xFFFFFFFF)))
        tex = (TextureDesc * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            px = np.ascontiguousarray(t["pixels"])
            keep.append(px)
            tex[i].pixels = px.ctypes.data
            tex[i].height, tex[i].width = px.shape[0], px.shape[1]
            tex[i].channels = px.shape[2] if px.ndim == 3 else 1
            tex[i].is_float = 1 if px.dtype == np.float32 else 0
            tex[i].srgb = int(bool(t.get("srgb", False)))
        d = SceneDesc()
        d.meshes, d.n_meshes = md, len(self.meshes)
        d.surfaces, d.n_surfaces = sd, len(self.surfaces)
        d.instances, d.n_instances = idesc, len(self.instances)
        d.materials, d.n_materials = mat, len(self.materials)
        d.textures, d.n_textures = tex, len(self.textures)
        d.camera.origin[:] = [float(x) for x in self.camera[0]]
        d.camera.basis[:] = [float(x) for x in self.camera[1]]
        d.camera.yfov = self.camera[2]
        if self.sun is not None:
            d.sun.enabled = 1
            d.sun.basis[:] = [float(x) for x in self.sun[0]]
            d.sun.energy[:] = [float(x) for x in self.sun[1]]
            d.sun.angular_radius = float(self.sun[2])
        d.environment_factor[:] = list(self.environment_factor)
        d.transparent_background = int(self.transparent_background)
        d.kd_use_sah = int(self.kd_use_sah)
        d.kd_max_depth = self.kd_max_depth
        d.environment_tex_plus1 = 0 if self.environment_texture is None else self.environment_texture + 1
        keep += [md, sd, idesc, mat, tex, self]
        return d, keep


_lib = None


def available() -> bool:
    return os.path.exists(SO_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{SO_PATH} is missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(SO_PATH)
        L.ref_scene_from_gltf.restype = C.c_void_p
        L.ref_scene_from_gltf.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
        L.ref_scene_from_desc.restype = C.c_void_p
        L.ref_scene_from_desc.argtypes = [C.POINTER(SceneDesc)]
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_export_counts.argtypes = [C.c_void_p, u32p, u32p, u32p, u32p]
        L.ref_export_mesh_counts.argtypes = [C.c_void_p, C.c_uint32, u32p, u32p]
        L.ref_export_mesh.argtypes = [C.c_void_p, C.c_uint32, f32p, f32p, f32p, f32p, u32p, f32p]
        L.ref_export_surfaces.argtypes = [C.c_void_p, C.POINTER(SurfaceDesc)]
        L.ref_export_instances.argtypes = [C.c_void_p, C.POINTER(InstanceDesc), f32p]
        L.ref_export_materials.argtypes = [C.c_void_p, C.POINTER(MaterialDesc), u32p]
        L.ref_export_texture_count.restype = C.c_uint32
        L.ref_export_texture_count.argtypes = [C.c_void_p]
        L.ref_export_texture_info.argtypes = [C.c_void_p, C.c_uint32, u32p]
        L.ref_export_texture_data.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_export_globals.argtypes = [C.c_void_p, C.POINTER(CameraDesc), C.POINTER(SunDesc), f32p, u32p]
        L.ref_dump_kd.restype = C.c_int
        L.ref_dump_kd.argtypes = [C.c_void_p, C.c_uint32, u32p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ref_trace_rays.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_void_p, f32p, C.c_int]
        L.ref_camera_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u32p, u32p, f32p, C.c_uint64, f32p]
        L.ref_count_visits.argtypes = [C.c_void_p, f32p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ref_render_linear.argtypes = [C.c_void_p] + [C.c_uint32] * 8 + [C.c_int] * 3 + [
            f32p, f32p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.ref_render_png.restype = C.c_uint64
        L.ref_render_png.argtypes = [C.c_void_p] + [C.c_uint32] * 5 + [C.c_void_p, C.c_uint64,
                                                                         C.POINTER(C.c_double)]
        L.ref_tonemap_rgba8.argtypes = [f32p, f32p, C.c_uint64, C.c_void_p]
        L.ref_hardware_threads.restype = C.c_uint32
        _lib = L
    return _lib


class RefScene:
    """A scene held by the reference library."""

    def __init__(self, handle, keep=None):
        if not handle:
            raise RuntimeError("reference failed to build the scene")
        self.h = C.c_void_p(handle)
        self._keep = keep

    @classmethod
    def from_gltf(cls, path, camera_index=0, sun_light_index=0):
        return cls(lib().ref_scene_from_gltf(os.fsencode(path), camera_index, sun_light_index))

    @classmethod
    def from_flat(cls, flat: FlatScene):
        d, keep = flat.to_c()
        return cls(lib().ref_scene_from_desc(C.byref(d)), keep)

    def close(self):
        if self.h:
            lib().ref_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- flat export in renderer::intersect visiting order
    def export_flat(self) -> FlatScene:
        L = lib()
        nm, ns, ni, nmat = (C.c_uint32() for _ in range(4))
        L.ref_export_counts(self.h, nm, ns, ni, nmat)
        meshes = []
        for m in range(nm.value):
            nv, nt = C.c_uint32(), C.c_uint32()
            L.ref_export_mesh_counts(self.h, m, nv, nt)
            pos = np.empty((nv.value, 3), np.float32)
            nrm = np.empty((nv.value, 3), np.float32)
            tan = np.empty((nv.value, 3), np.float32)
            uv = np.empty((nv.value, 2), np.float32)
            idx = np.empty((nt.value, 3), np.uint32)
            aabb = np.empty(6, np.float32)
            L.ref_export_mesh(self.h, m, _fp(pos), _fp(nrm), _fp(tan), _fp(uv), _up(idx), _fp(aabb))
            meshes.append(dict(positions=pos, normals=nrm, tangents=tan, uvs=uv, indices=idx, aabb=aabb))
        sd = (SurfaceDesc * ns.value)()
        L.ref_export_surfaces(self.h, sd)
        surfaces = np.array([(s.mesh, s.material) for s in sd], np.uint32).reshape(-1, 2)
        idesc = (InstanceDesc * ni.value)()
        maabb = np.empty((ni.value, 6), np.float32)
        L.ref_export_instances(self.h, idesc, _fp(maabb))
        instances = [(np.array(list(i.origin), np.float32), np.array(list(i.basis), np.float32),
                      i.first_surface, i.n_surfaces) for i in idesc]
        mat = (MaterialDesc * nmat.value)()
        mask = np.zeros(nmat.value, np.uint32)
        L.ref_export_materials(self.h, mat, _up(mask))
        materials = [dict(albedo=tuple(m.albedo), opacity=m.opacity, roughness=m.roughness,
                          metallic=m.metallic, emissive=tuple(m.emissive), ior=m.ior,
                          shadow_catcher=m.shadow_catcher, normal_tex=m.normal_tex, albedo_tex=m.albedo_tex,
                          opacity_tex=m.opacity_tex, roughness_tex=m.roughness_tex, metallic_tex=m.metallic_tex,
                          emissive_tex=m.emissive_tex) for m in mat]
        textures = []
        for t in range(L.ref_export_texture_count(self.h)):
            info = np.zeros(5, np.uint32)
            L.ref_export_texture_info(self.h, t, _up(info))
            w, h, c, is_float, srgb = (int(x) for x in info)
            px = np.empty((h, w, c), np.float32 if is_float else np.uint8)
            L.ref_export_texture_data(self.h, t, px.ctypes.data)
            textures.append(dict(pixels=px, srgb=bool(srgb)))
        cam, sun = CameraDesc(), SunDesc()
        env = np.empty(3, np.float32)
        tr = C.c_uint32()
        L.ref_export_globals(self.h, cam, sun, _fp(env), tr)
        flat = FlatScene(meshes, surfaces, instances, materials,
                         (list(cam.origin), list(cam.basis), cam.yfov),
                         (list(sun.basis), list(sun.energy), sun.angular_radius) if sun.enabled else None,
                         tuple(env), bool(tr.value), textures=textures)
        flat.texture_masks = mask
        flat.model_aabbs = maabb
        flat.mesh_aabbs = [m["aabb"] for m in meshes]
        return flat

    def dump_kd(self, mesh: int) -> np.ndarray:
        n = C.c_uint64()
        lib().ref_dump_kd(self.h, mesh, None, 0, n)
        out = np.empty(n.value, np.uint32)
        rc = lib().ref_dump_kd(self.h, mesh, _up(out), n.value, n)
        assert rc == 0
        return out

    def trace_rays(self, origin_dir: np.ndarray, attrs=False, threads=0):
        od = np.ascontiguousarray(origin_dir, np.float32).reshape(-1, 6)
        hits = np.zeros(len(od), HIT_DTYPE)
        at = np.zeros((len(od), 14), np.float32) if attrs else None
        lib().ref_trace_rays(self.h, _fp(od), len(od), hits.ctypes.data, _fp(at) if attrs else None, threads)
        return (hits, at) if attrs else hits

    def camera_rays(self, w, h, px, py, aa):
        px = np.ascontiguousarray(px, np.uint32)
        py = np.ascontiguousarray(py, np.uint32)
        aa = np.ascontiguousarray(aa, np.float32).reshape(-1, 2)
        od = np.empty((len(px), 6), np.float32)
        lib().ref_camera_rays(self.h, w, h, _up(px), _up(py), _fp(aa), len(px), _fp(od))
        return od

    def count_visits(self, origin_dir):
        od = np.ascontiguousarray(origin_dir, np.float32).reshape(-1, 6)
        c = (C.c_uint64 * 6)()
        lib().ref_count_visits(self.h, _fp(od), len(od), c)
        return dict(zip(("model_tests", "surface_tests", "branch_visits", "leaf_visits", "tri_tests",
                         "stack_pushes"), [int(x) for x in c]))

    def render_linear(self, full_w, full_h, spp, depth, mode=0, tile=None, first_sample_unjittered=False,
                      threads=0):
        x0, y0, w, h = tile if tile else (0, 0, full_w, full_h)
        rgb = np.empty((h, w, 3), np.float32)
        alpha = np.empty((h, w), np.float32)
        rays, secs = C.c_uint64(), C.c_double()
        lib().ref_render_linear(self.h, full_w, full_h, x0, y0, w, h, spp, depth, mode,
                                int(first_sample_unjittered), threads, _fp(rgb), _fp(alpha), rays, secs)
        return rgb, alpha, rays.value, secs.value

    def render_png(self, w, h, spp, depth, threads=0):
        secs = C.c_double()
        cap = w * h * 4 + (1 << 20)
        buf = (C.c_uint8 * cap)()
        n = lib().ref_render_png(self.h, w, h, spp, depth, threads, buf, cap, secs)
        return bytes(buf[:n]), secs.value


def tonemap_rgba8(rgb, alpha=None):
    rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
    out = np.empty((len(rgb), 4), np.uint8)
    a = np.ascontiguousarray(alpha, np.float32).reshape(-1) if alpha is not None else None
    lib().ref_tonemap_rgba8(_fp(rgb), _fp(a) if a is not None else None, len(rgb), out.ctypes.data)
    return out


def hardware_threads():
    return lib().ref_hardware_threads()
