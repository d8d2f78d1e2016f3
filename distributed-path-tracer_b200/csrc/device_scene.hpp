// device_scene.hpp — the scene as it lies in HBM (included by nvcc-compiled files only).
//
// Layout (every array one contiguous cudaMalloc, 256-byte aligned; sized for a
// 180 GB part: 64-bit counts on the host, 32-bit indices per mesh on the device):
//   instances[]  one per reference entity-with-model, in renderer::intersect
//                visiting order (LIB/core/renderer.cpp:646-671); carries the
//                inverse transform the reference recomputes per ray
//                (LIB/scene/model.cpp:22-25) precomputed with the same float
//                ops, the forward transform, the normal matrix and the model AABB
//   surfaces[]   {mesh, material} pairs (scene::model::surface)
//   meshes[]     per unique mesh: AABB + offsets into the arrays below
//   kd_nodes[]   8-byte nodes, all meshes back to back (kd_build.hpp)
//   kd_refs[]    leaf → triangle references (u32, index into tri_* of the mesh)
//   tri[]        three float4 per unique triangle (one 48-byte record): a, a-b, a-c.  The two edge
//                differences are what triangle::intersect forms first
//                (LIB/geometry/triangle.cpp:136-140); they are single float
//                subtractions, so precomputing them changes no bit.
//                .w lanes carry the three vertex indices of the triangle.
//   vtx_pos/nrm/tan/uv   vertex attributes for shading
//   materials[]  factors + texture ids; textures[] descriptors + texel pool
// LIB = path-tracer-core/path_tracer_lib/path_tracer in the reference repository.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "vecmath.hpp"

namespace ptb {

struct DInstance {
    Xform inv;     // world → local
    Xform fwd;     // local → world
    M3 normal_mat; // transpose(inverse(fwd.basis)), LIB/core/renderer.cpp:698
    float aabb_min[3], aabb_max[3]; // scene::model::aabb (local space)
    uint32_t first_surface, n_surfaces;
    uint32_t same_box; // 1: a single surface whose mesh box is bit for bit the model box (its slab test need not be repeated)
};

struct DSurface {
    uint32_t mesh;
    uint32_t material;
};

struct DMesh {
    float aabb_min[3], aabb_max[3]; // core::mesh::aabb
    uint32_t node_base;             // root node index in kd_nodes (child indices are mesh-relative)
    uint32_t ref_base;              // added to a leaf's first reference
    uint32_t tri_base;              // added to a reference to index tri_*
    uint32_t vtx_base;              // added to a vertex index
    uint32_t n_triangles;
    uint32_t pair_base;             // the mesh's root pair in kd_pairs (child pair indices are mesh-relative)
};

struct DMaterial {
    float albedo[3];
    float opacity;
    float roughness;
    float metallic;
    float emissive[3];
    float ior;
    uint32_t shadow_catcher;
    uint32_t normal_tex, albedo_tex, opacity_tex, roughness_tex, metallic_tex, emissive_tex;
    uint32_t any_tex; // 0 when every slot is PTB_NO_TEXTURE
    uint32_t pad[2];
};

struct DTexture {
    unsigned long long offset; // byte offset into the texel pool
    uint32_t width, height, channels;
    uint32_t is_float, srgb;
    uint32_t pad;
};

struct DCamera {
    Xform xf;
    float tan_half_fov; // camera::set_fov, LIB/scene/camera.cpp:27-30
};

struct DSun {
    uint32_t enabled;
    V3 direction; // basis * (0,0,1), LIB/core/renderer.cpp:499
    V3 energy;
    float angular_radius;
};

// Kernel argument (by value).
struct DScene {
    const DInstance* instances;
    const float4* inst_sphere; // conservative world-space bound per instance: centre xyz, radius (< 0: never hit)
    const DSurface* surfaces;
    const DMesh* meshes;
    const uint2* kd_nodes;
    const uint4* kd_pairs; // the same trees as sibling pairs (16 B: left record, right record), see scene.cu
    const uint32_t* kd_refs;
    const float4* tri; // 3 float4 per triangle: a, a-b, a-c (.w = the three vertex indices)
    const float* vtx_pos; // 3 per vertex
    const float* vtx_nrm; // 3 per vertex
    const float* vtx_tan; // 3 per vertex
    const float* vtx_uv;  // 2 per vertex
    const DMaterial* materials;
    const DTexture* textures;
    const unsigned char* texels;
    uint32_t n_instances;
    uint32_t n_pairs, n_refs, n_tris; // array sizes, for the instrumented kernel's bounds checks
    DCamera camera;
    DSun sun;
    V3 environment;
    uint32_t environment_tex; // 0xFFFFFFFF: none; else the equirectangular map a missing ray samples
    uint32_t transparent_background;
};

} // namespace ptb
