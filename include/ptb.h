/*
 * ptb.h — C ABI of the B200-native path-tracing hot path (libptb.so).
 *
 * This is the drop-in boundary for the one hot path of
 * vmanam0451/distributed-path-tracer: KD-tree traversal, ray/triangle
 * intersection and the Monte-Carlo integrator.  The reference has no FFI seam
 * for this path (it is statically linked, path-tracer-core/CMakeLists.txt:25,41-44),
 * so every entry point below names the reference C++ interface it replaces.
 * All paths in the comments are relative to the reference repository root;
 * LIB/ = path-tracer-core/path_tracer_lib/path_tracer/, APP/ = path-tracer-core/src/.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary
 *   - every function returns a ptb_status; nothing throws, nothing aborts
 *   - ptb_last_error() returns a thread-local, human-readable message
 *   - scene handles are immutable after creation and may be shared by threads
 *   - output buffers are caller-owned HOST memory unless the name ends in _dev
 *   - the library FAILS (PTB_E_CUDA) when no CUDA device is usable: there is
 *     no CPU fallback of any kind
 */
#ifndef PTB_H
#define PTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_ABI_VERSION 4

typedef enum ptb_status {
    PTB_OK = 0,
    PTB_E_INVALID = 1, /* bad argument / malformed scene description        */
    PTB_E_CUDA = 2,    /* CUDA runtime error, or no usable device            */
    PTB_E_NCCL = 3,    /* multi-GPU exchange failed: peer access, CUDA IPC,
                          shared-memory rendezvous or a group time-out       */
    PTB_E_OOM = 4,     /* host or device allocation failed                   */
    PTB_E_IO = 5       /* file missing / unreadable / malformed glTF         */
} ptb_status;

/* ---------------------------------------------------------------- scene -- */

/* One triangle mesh = one glTF primitive = one reference core::mesh
 * (LIB/core/mesh.hpp:13-37; vertex layout LIB/core/vertex.hpp:7-12). */
typedef struct ptb_mesh_desc {
    const float* positions;  /* n_vertices * 3                               */
    const float* normals;    /* n_vertices * 3                               */
    const float* tangents;   /* n_vertices * 3 (the reference reads xyz only,
                                LIB/core/renderer.cpp:215-218,248-250)        */
    const float* uvs;        /* n_vertices * 2                               */
    uint32_t n_vertices;
    const uint32_t* indices; /* n_triangles * 3                              */
    uint32_t n_triangles;
} ptb_mesh_desc;

/* scene::model::surface = {mesh, material} (LIB/scene/model.hpp:15-18). */
typedef struct ptb_surface_desc {
    uint32_t mesh;
    uint32_t material;
} ptb_surface_desc;

/* One entity carrying a scene::model, with its GLOBAL transform
 * (LIB/scene/transform.hpp:14-15: origin + column-major 3x3 basis, columns
 * x,y,z).  Instances are intersected in array order with strict '<' on world
 * distance, i.e. the first one wins a tie — the array order therefore has to
 * be the visiting order of renderer::intersect (LIB/core/renderer.cpp:645-671).
 * Several instances may name the same surface range (instancing). */
typedef struct ptb_instance_desc {
    float origin[3];
    float basis[9]; /* x.x x.y x.z  y.x y.y y.z  z.x z.y z.z                 */
    uint32_t first_surface;
    uint32_t n_surfaces;
} ptb_instance_desc;

/* An 8-bit or float texture (LIB/image/image.hpp, image_texture.cpp:21-62).
 * `srgb` makes reads of channels 0..2 decode with pow(v, 2.2)
 * (LIB/image/image.cpp:124-141). */
typedef struct ptb_texture_desc {
    const void* pixels; /* row-major, `channels` interleaved                 */
    uint32_t width, height, channels;
    uint32_t is_float; /* 0: uint8 (value/255), 1: float32                   */
    uint32_t srgb;
} ptb_texture_desc;

#define PTB_NO_TEXTURE 0xFFFFFFFFu

/* core::material (LIB/core/material.hpp:9-41, getters material.cpp:6-53). */
typedef struct ptb_material_desc {
    float albedo[3];
    float opacity;
    float roughness;
    float metallic;
    float emissive[3];
    float ior; /* 1.33 in the reference (material.hpp:16)                    */
    uint32_t shadow_catcher;
    uint32_t normal_tex, albedo_tex, opacity_tex, roughness_tex, metallic_tex,
        emissive_tex; /* PTB_NO_TEXTURE when absent                          */
} ptb_material_desc;

/* scene::camera on an entity (LIB/scene/camera.cpp:10-30): global transform
 * and vertical field of view in radians. */
typedef struct ptb_camera_desc {
    float origin[3];
    float basis[9];
    float yfov;
} ptb_camera_desc;

/* scene::sun_light (LIB/scene/sun_light.hpp:7-11) with the global basis of
 * its entity (LIB/core/renderer.cpp:499). */
typedef struct ptb_sun_desc {
    uint32_t enabled;
    float basis[9];
    float energy[3];
    float angular_radius;
} ptb_sun_desc;

typedef struct ptb_scene_desc {
    const ptb_mesh_desc* meshes;
    uint32_t n_meshes;
    const ptb_surface_desc* surfaces;
    uint32_t n_surfaces;
    const ptb_instance_desc* instances;
    uint32_t n_instances;
    const ptb_material_desc* materials;
    uint32_t n_materials;
    const ptb_texture_desc* textures;
    uint32_t n_textures;
    ptb_camera_desc camera;
    ptb_sun_desc sun;
    float environment_factor[3];    /* renderer.hpp:29                       */
    uint32_t transparent_background; /* renderer.hpp:30                      */
    uint32_t kd_use_sah;  /* mesh::build_kd_tree(use_sah=true, ...)          */
    uint32_t kd_max_depth; /* ... max_depth=25  (LIB/core/mesh.hpp:34); 0 → 25 */
    uint32_t environment_tex_plus1; /* renderer.hpp:28 `environment`: 1 + index of the equirectangular
                                       texture a missing ray samples (renderer.cpp:446-448,
                                       worker.cpp:308-311); 0 = none (environment_factor alone) */
} ptb_scene_desc;

typedef struct ptb_scene ptb_scene;

/* Replaces renderer::get_mesh + mesh::recalculate_aabb + mesh::build_kd_tree
 * (LIB/core/renderer.cpp:177-263, LIB/core/mesh.cpp:254-298): builds, on the
 * host, the identical SAH KD-tree per mesh, flattens it and uploads triangles,
 * nodes, leaf references, vertex attributes, materials and textures to HBM on
 * `device`. */
ptb_status ptb_scene_create(const ptb_scene_desc* desc, int device, ptb_scene** out);

/* Replaces renderer::load_gltf (LIB/core/renderer.cpp:61-99, process_node
 * :101-174): camera and sun light are matched by name, `camera_index` and
 * `sun_light_index` as renderer.hpp:31-32 (0xFFFFFFFF = renderer::no_sun_light). */
ptb_status ptb_scene_load_gltf(const char* path, uint32_t camera_index,
                               uint32_t sun_light_index, int device, ptb_scene** out);

void ptb_scene_destroy(ptb_scene* scene);

typedef struct ptb_scene_info {
    uint32_t n_instances, n_surfaces, n_meshes, n_materials, n_textures;
    uint64_t n_triangles;   /* unique triangles over all meshes              */
    uint64_t n_kd_nodes;    /* flattened nodes (branches + leaves)           */
    uint64_t n_kd_branches;
    uint64_t n_kd_leaves;
    uint64_t n_leaf_refs;   /* triangle references held by leaves            */
    uint32_t kd_max_depth_reached;
    uint64_t device_bytes;  /* HBM held by the scene                         */
    double build_seconds;   /* host KD build                                 */
    double upload_seconds;
} ptb_scene_info;

ptb_status ptb_scene_get_info(const ptb_scene* scene, ptb_scene_info* out);

/* Test hook for row K of the scope table: serialises the host-built tree of
 * one mesh in depth-first order so that it can be compared with the
 * reference's pointer tree.  Record stream (uint32 words):
 *   branch: 0x80000000|axis, split-bits, has_left, has_right  then left, right
 *   leaf:   count, idx[count]
 * Call with words == NULL to obtain the required length. */
ptb_status ptb_scene_dump_kd(const ptb_scene* scene, uint32_t mesh, uint32_t* words,
                             uint64_t capacity, uint64_t* n_words);

/* ------------------------------------------------------------- hot path -- */

typedef struct ptb_hit {
    uint32_t instance; /* 0xFFFFFFFF = miss                                   */
    uint32_t surface;  /* ordinal inside the instance's surface list          */
    uint32_t triangle; /* index into the mesh's triangle list                 */
    float t;           /* world-space distance (LIB/scene/model.cpp:62-63)    */
    float bary[3];     /* (alpha, beta, gamma), triangle.cpp:185-189          */
} ptb_hit;

#define PTB_MISS 0xFFFFFFFFu

/* Closest hit for an explicit ray set: replaces renderer::intersect's search
 * (LIB/core/renderer.cpp:645-675 → LIB/scene/model.cpp:20-72 →
 * LIB/core/mesh.cpp:300-405 → LIB/geometry/triangle.cpp:120-190).
 * `origin_dir` is n * 6 floats (origin xyz, direction xyz); directions are
 * normalised on the way in exactly as geometry::ray's constructor does
 * (LIB/geometry/ray.cpp:6-8).  Triangle ids are bit-exact with the reference. */
ptb_status ptb_trace_rays(const ptb_scene* scene, const float* origin_dir, uint64_t n,
                          ptb_hit* hits_out);

/* Full intersect_result for an explicit ray set (renderer.cpp:688-724): world
 * position(3), uv(2), interpolated normal(3), tangent(3), shading normal(3)
 * = 14 floats per ray, zeros on a miss.  `attrs_out` may be NULL. */
ptb_status ptb_trace_rays_attrs(const ptb_scene* scene, const float* origin_dir, uint64_t n,
                                ptb_hit* hits_out, float* attrs_out);

typedef enum ptb_integrator {
    /* core::renderer::trace, fixed depth, no Russian roulette
     * (LIB/core/renderer.cpp:437-643) */
    PTB_INTEGRATOR_LIB = 0,
    /* processors::worker::trace_iter, throughput clamp + Russian roulette
     * (APP/processors/worker/worker.cpp:285-514) */
    PTB_INTEGRATOR_APP_RR = 1
} ptb_integrator;

/* The worker's render request.  The reference request is
 * {samples, bounces, X, Y} (APP/models/work_info.hpp:27-30) over the whole
 * frame; tile rectangle, seed and first sample index are extensions that let
 * tiles and sample ranges be sharded across GPUs. */
typedef struct ptb_tile_req {
    uint32_t full_w, full_h; /* X, Y                                          */
    uint32_t x0, y0, w, h;   /* tile rectangle inside the frame               */
    uint32_t spp;            /* samples                                       */
    uint32_t max_depth;      /* bounces                                       */
    uint64_t seed;           /* Philox key                                    */
    uint32_t first_sample;   /* index of this call's first sample (0 normally)*/
    uint32_t integrator;     /* ptb_integrator                                */
    uint32_t first_sample_unjittered; /* APP/processors/worker/worker.cpp:125-129 */
    uint32_t reserved;
    /* Transparent-background scenes only (renderer.cpp:374-393 keeps a per-pixel "claimed" flag next to the
     * running mean): w*h bytes, caller-owned, in the memory space of the output (host for ptb_render_tile,
     * device for ptb_render_tile_dev).  Written by every call; read when first_sample != 0.  May be NULL for a
     * call that renders all its samples at once (first_sample == 0 and no later call continues it).  Together
     * with the output buffer — which is IN/OUT when first_sample != 0 — this is the whole state that chains
     * sample ranges: the library keeps none of it. */
    void* claim_mask;
} ptb_tile_req;

typedef struct ptb_render_stats {
    uint64_t paths;         /* camera paths started                           */
    uint64_t rays;          /* scene-level intersect calls: primary + bounce
                               + shadow, counted by the extend/shadow kernels */
    uint64_t kernel_launches;
    double gpu_seconds;     /* CUDA-event time of the whole render            */
    double extend_seconds;  /* closest-hit + shadow kernels                   */
    double shade_seconds;   /* ray-gen + shade + accumulate kernels           */
    uint64_t extend_launches;
    uint64_t node_visits, leaf_visits, tri_tests; /* only when the scene was
                               created with counters enabled (ptb_set_option) */
} ptb_render_stats;

/* Shadow query for an explicit ray set: replaces the SECOND intersect call of a shade event,
 * `!intersect(shadow_ray).hit` (LIB/core/renderer.cpp:505-511; the staged worker:
 * APP/processors/worker/intersection_worker.cpp:49-67).  occluded_out[i] = 1 when ray i hits anything.  Runs the
 * any-hit instantiation of the extend kernel, the one the wavefront resolves its sun shadow rays with: it stops at
 * the first accepted triangle of the first leaf that holds one, and equals (ptb_trace_rays(...).instance != PTB_MISS)
 * ray for ray.  stats_out may be NULL (rays, kernel time, visit counters with "count_visits"). */
ptb_status ptb_trace_occlusion(const ptb_scene* scene, const float* origin_dir, uint64_t n, uint8_t* occluded_out,
                               ptb_render_stats* stats_out);

/* Replaces renderer::render (LIB/core/renderer.cpp:334-428) /
 * worker::run's pipeline (APP/processors/worker/worker.cpp:25-105) for one
 * tile: linear running-mean radiance, row-major w*h*3, and the alpha channel
 * (w*h, may be NULL when first_sample == 0). Synchronous; safe to call from several
 * host threads.  With first_sample != 0 the buffers are IN/OUT: they must hold the
 * running mean of samples [0, first_sample) as an earlier call returned it. */
ptb_status ptb_render_tile(const ptb_scene* scene, const ptb_tile_req* req, float* rgb_out,
                           float* alpha_out, ptb_render_stats* stats_out);

/* Same, leaving the result in device memory: `rgba_dev` is a device pointer
 * to w*h float4 (rgb = mean radiance, a = alpha; IN/OUT when first_sample != 0).  `stream` is a cudaStream_t
 * (NULL = the default stream).  Asynchronous with respect to the host except
 * for the per-bounce queue-size read-back. */
ptb_status ptb_render_tile_dev(const ptb_scene* scene, const ptb_tile_req* req, void* rgba_dev,
                               void* stream, ptb_render_stats* stats_out);

/* tonemap_approx_aces (LIB/core/utils.hpp:29-36) + image::write's sRGB
 * encode and rounding (LIB/image/image.cpp:143-154): n pixels of linear rgb
 * (+ alpha, may be NULL → 1) to RGBA8, on the GPU. */
ptb_status ptb_tonemap_rgba8(const float* rgb, const float* alpha, uint64_t n_pixels,
                             uint8_t* rgba8_out);

/* image::save_to_memory_png (LIB/image/image.cpp:111-122): RGBA8 → PNG file. */
ptb_status ptb_write_png(const char* path, const uint8_t* rgba8, uint32_t w, uint32_t h);

/* The Lambda worker's request/response (my_handler → processors::worker::run,
 * APP/main.cpp:9-31, APP/processors/worker/worker.cpp:25-105) without the S3
 * hops.  `worker_info_json` is the payload the preprocessor sends
 * (models::worker_info, APP/models/work_info.hpp:17-31: scene_info.work = mesh
 * name → primitive indices assigned to this worker, samples, bounces, X, Y, …;
 * a payload without samples/bounces/X/Y — what PRE/app.py:119-127 really emits —
 * gets the worker's defaults 50 / 10 / 640 / 480, worker.hpp:20-24).
 * `scene_dir` is a local mirror of s3://scene_bucket/scene_root and must hold
 * scene.gltf with its buffers and textures.  Renders with the worker's
 * integrator (PTB_INTEGRATOR_APP_RR, first sample unjittered), tonemaps as
 * worker::generate_final_image does, writes the RGBA8 PNG the worker would
 * upload as "test.png" to `png_path` (may be NULL) and/or copies the RGBA8
 * pixels to `rgba8_out` (may be NULL; X*Y*4 bytes). */
ptb_status ptb_worker_run(const char* worker_info_json, const char* scene_dir, int device,
                          const char* png_path, uint8_t* rgba8_out, uint32_t* width_out,
                          uint32_t* height_out, ptb_render_stats* stats_out);

/* ------------------------------------------------------------ multi-GPU -- */
/*
 * Image tiles shard across GPUs, the scene is replicated.  What the reference does with a
 * thread pool over scanlines and a barrier per sample (LIB/core/renderer.cpp:354-407), and its worker with
 * stage threads (APP/processors/worker/worker.cpp:25-105), is here:
 *   - one RANK per GPU: a host thread of one process (ptb_ctx) or one process per GPU (ptb_group);
 *   - tiles claimed from ONE shared counter (work stealing: std::atomic fetch-add, in the process heap or in a
 *     POSIX shared-memory segment), several tiles in flight per GPU on separate streams;
 *   - NO collective in the data path: the frame lives on rank 0's GPU and every rank's accumulate kernel
 *     stores its finished pixels straight into it over NVLink (peer access in one process, CUDA IPC across
 *     processes) — the tile return is fused into the kernel that produces the pixels;
 *   - the scene is built ONCE and its flattened HBM image copied GPU → GPU (ptb_scene_clone), or broadcast
 *     by the caller between ptb_scene_export_header / ptb_scene_import (e.g. ncclBroadcast into ptb_scene_blob).
 * A frame rendered on N GPUs is bit-identical to the same frame on one (tiles compose exactly).
 */

/* The flattened scene is ONE device allocation ("blob") plus a small plain-data header describing it. */
ptb_status ptb_scene_blob(const ptb_scene* scene, void** blob_dev, uint64_t* blob_bytes);
/* Call with header == NULL to obtain the required size. */
ptb_status ptb_scene_export_header(const ptb_scene* scene, void* header, uint64_t capacity, uint64_t* n_bytes);
/* A replica on `device` from a header.  src_blob_dev != NULL: the blob is copied from that device pointer
 * (cudaMemcpyPeer from src_device).  src_blob_dev == NULL: the blob is left for the caller to fill (obtain it
 * with ptb_scene_blob, e.g. as the destination of an ncclBroadcast) before the scene is used.  Replicas hold
 * no host copy of the trees (ptb_scene_dump_kd fails on them). */
ptb_status ptb_scene_import(const void* header, uint64_t n_bytes, int device, const void* src_blob_dev, int src_device,
                            ptb_scene** out);
/* = export + import with a device-to-device copy: no rebuild, no host round trip. */
ptb_status ptb_scene_clone(const ptb_scene* scene, int device, ptb_scene** out);

#define PTB_OUT_NONE 0u    /* the frame stays in HBM on rank 0 (throughput measurements) */
#define PTB_OUT_RGBA32F 1u /* full_w*full_h*4 floats: linear running-mean rgb + alpha     */
#define PTB_OUT_RGBA8 2u   /* full_w*full_h*4 bytes: tonemap_approx_aces + sRGB, what worker::generate_final_image
                              hands to the PNG encoder (APP/processors/worker/worker.cpp:172-191) */

/* The worker's request over the whole frame ({samples, bounces, X, Y}, APP/models/work_info.hpp:27-30) plus how
 * to cut it. */
typedef struct ptb_frame_req {
    uint32_t full_w, full_h;
    uint32_t spp, max_depth;
    uint64_t seed;
    uint32_t integrator;              /* ptb_integrator */
    uint32_t first_sample_unjittered;
    uint32_t tile_w, tile_h;          /* the unit of work stealing; 0 = chosen by the library        */
    uint32_t tiles_in_flight;         /* streams (host threads) per GPU; 0 = default (8)            */
    uint32_t output;                  /* PTB_OUT_*                                                  */
} ptb_frame_req;

typedef struct ptb_frame_stats {
    uint64_t paths, rays, kernel_launches; /* summed over all ranks                                 */
    uint32_t n_tiles, n_ranks;
    double gpu_seconds;   /* CUDA-event time on the slowest rank, first tile start to last tile end  */
    double wall_seconds;  /* host clock around the whole call on rank 0, output copy included        */
    uint64_t tiles_per_rank[16];
    double gpu_seconds_per_rank[16];
} ptb_frame_stats;

/* The tiles a frame request is cut into for `world` ranks, in claim order: (x0, y0, w, h) quadruples.  With
 * tile_w / tile_h in the request: a uniform row-major grid.  With 0: the library's choice — for several ranks big
 * tiles first and small tiles last (guided scheduling).  Call with xywh == NULL to obtain the count. */
ptb_status ptb_frame_tiles(const ptb_frame_req* req, int world, uint32_t* xywh, uint64_t capacity, uint32_t* n_tiles);

/* The same with the tiles' pixel layout: eight values per tile, (x0, y0, w, h, gx, sx, gy, sy).  For several ranks
 * the library's choice is COMB tiles: tile pixel (x, y), 0 <= x < w, 0 <= y < h, is frame pixel
 *   (x0 + (x / gx) * sx + x % gx,  y0 + (y / gy) * sy + y % gy)
 * i.e. granules of gx x gy pixels that lie sx / sy pixels apart — every tile samples the whole image, so all tiles
 * cost the same and each can be as large as a GPU's share of the frame divided by its streams (a frame cut into
 * rectangles needs SMALL tiles to balance sky against terrain, and small launches are slow).  gx = gy = 0: a plain
 * rectangle.  A frame's pixels never depend on the tiling (they are a function of seed, pixel and sample). */
ptb_status ptb_frame_tile_layout(const ptb_frame_req* req, int world, uint32_t* layout, uint64_t capacity, uint32_t* n_tiles);

/* One process per GPU (torchrun, MPI, ...): collective over `world` processes of ONE node that pass the same
 * `name` (a POSIX shared-memory object "/ptb_<name>": tile counters, barrier, IPC handle of the frame).
 * Rank 0 owns the frame.  Every call below is COLLECTIVE: all ranks call it with the same arguments.
 * A rank that fails or does not arrive within the time-out (option "group_timeout_ms") makes every rank
 * return PTB_E_NCCL instead of hanging. */
typedef struct ptb_group ptb_group;
ptb_status ptb_group_create(const char* name, int rank, int world, int device, ptb_group** out);
void ptb_group_destroy(ptb_group* group);
ptb_status ptb_group_barrier(ptb_group* group);
/* `scene` is this rank's replica.  out_host (rank 0 only; others pass NULL) receives the frame in the format
 * req->output names; pinned memory is written directly, pageable memory through a pinned staging buffer.
 * stats_out may be NULL; it is filled on rank 0. */
ptb_status ptb_group_render_frame(ptb_group* group, const ptb_scene* scene, const ptb_frame_req* req,
                                  void* out_host, ptb_frame_stats* stats_out);

/* One process, n_gpus GPUs, one host thread per GPU: what replaces worker::run for a multi-GPU box.
 * devices == NULL means 0 .. n_gpus-1. */
typedef struct ptb_ctx ptb_ctx;
ptb_status ptb_ctx_create(int n_gpus, const int* devices, ptb_ctx** out);
void ptb_ctx_destroy(ptb_ctx* ctx);
/* Builds the scene once (host KD build + upload to the first device) and replicates its blob GPU → GPU. */
ptb_status ptb_ctx_set_scene(ptb_ctx* ctx, const ptb_scene_desc* desc);
ptb_status ptb_ctx_load_gltf(ptb_ctx* ctx, const char* path, uint32_t camera_index, uint32_t sun_light_index);
/* The replica on the i-th GPU of the context (owned by the context). */
const ptb_scene* ptb_ctx_scene(const ptb_ctx* ctx, int i);
ptb_status ptb_render_frame(ptb_ctx* ctx, const ptb_frame_req* req, void* out_host, ptb_frame_stats* stats_out);

/* ptb_worker_run (the Lambda worker's request → RGBA8 / PNG) with the frame tile-sharded over ctx's GPUs. */
ptb_status ptb_worker_run_ctx(ptb_ctx* ctx, const char* worker_info_json, const char* scene_dir,
                              const char* png_path, uint8_t* rgba8_out, uint32_t* width_out,
                              uint32_t* height_out, ptb_frame_stats* stats_out);

/* Pinned host memory for frame outputs (a plain cudaHostAlloc): lets ptb_render_frame copy without staging. */
ptb_status ptb_host_alloc(uint64_t bytes, void** out);
void ptb_host_free(void* p);

/* ------------------------------------------------- host-only / test hooks -- */

/* The build step of ptb_scene_create alone, on the host, without CUDA:
 * mesh::recalculate_aabb + mesh::build_kd_tree (LIB/core/mesh.cpp:254-298)
 * serialised as ptb_scene_dump_kd does.  Lets the tree-identity tests run on
 * a machine without a GPU.  aabb6_out (min xyz, max xyz) may be NULL. */
ptb_status ptb_host_build_kd(const float* positions, uint32_t n_vertices, const uint32_t* indices,
                             uint32_t n_triangles, uint32_t use_sah, uint32_t max_depth, int threads,
                             uint32_t* words, uint64_t capacity, uint64_t* n_words, float* aabb6_out);

/* renderer::load_gltf's parsing half, host-only: glTF → scene description
 * (instances in renderer::intersect visiting order). */
typedef struct ptb_desc ptb_desc;
ptb_status ptb_desc_load_gltf(const char* path, uint32_t camera_index, uint32_t sun_light_index,
                              ptb_desc** out);
const ptb_scene_desc* ptb_desc_get(const ptb_desc* desc);
void ptb_desc_free(ptb_desc* desc);

/* Primary rays exactly as renderer::render builds them (renderer.cpp:360-370 →
 * camera::get_ray) for caller-supplied pixel coordinates and jitter (aa = 2
 * floats per ray): the device ray-gen function, exposed so that it can be
 * compared bit for bit.  origin_dir_out: n * 6 floats. */
ptb_status ptb_camera_rays(const ptb_scene* scene, uint32_t w, uint32_t h, const uint32_t* px,
                           const uint32_t* py, const float* aa, uint64_t n, float* origin_dir_out);

/* ptb_trace_rays plus CUDA-event timing of the extend kernel and, with the
 * "count_visits" option, the node / leaf / triangle visit counters. */
ptb_status ptb_trace_rays_stats(const ptb_scene* scene, const float* origin_dir, uint64_t n,
                                ptb_hit* hits_out, ptb_render_stats* stats_out);

/* ptb_trace_rays with rays (n * 6 floats) and hits in DEVICE memory, asynchronous on `stream`
 * (a cudaStream_t; NULL = the default stream).  No host copy. */
ptb_status ptb_trace_rays_dev(const ptb_scene* scene, const float* origin_dir_dev, uint64_t n, ptb_hit* hits_dev,
                              void* stream);

/* Geometry-sharded closest hit — the reference's own distribution model
 * (APP/processors/worker/intersection_worker.cpp:69-147: every worker holds part of the geometry, every ray
 * visits every worker, the nearest result wins).  One process per GPU; each holds a SHARD scene (a subset
 * of the instances) and, in peer-mapped (symmetric) memory, a key buffer (n uint64) and a payload buffer
 * (n * 16 bytes) whose device addresses on every rank are known to all ranks.
 *   1. ptb_shard_reset_dev          own key buffer := "miss"                       -- then a barrier
 *   2. ptb_shard_trace_dev          trace the rays against the shard, and min-merge  key = t_bits << 32 |
 *                                   global instance << 12 | surface  into EVERY rank's key buffer with 64-bit
 *                                   atomics over NVLink (the exchange step; no NCCL)  -- then a barrier
 *   3. ptb_shard_publish_dev        the rank whose key won stores triangle + barycentrics into every
 *                                   rank's payload buffer                           -- then a barrier
 *   4. ptb_shard_unpack_dev         merged keys + payload -> ptb_hit records (global instance indices)
 * The result is bit-identical to ptb_trace_rays on the unsharded scene: hit distances are non-negative floats
 * (their bits order like the numbers) and renderer::intersect keeps the first instance in scene order on
 * equal distances (renderer.cpp:663-669).  instance_map_dev: shard-local -> global instance index.
 * peer_keys / peer_payload: HOST arrays of `world` device pointers (rank order).  Steps 2 and 3 must use the
 * same stream. */
ptb_status ptb_shard_reset_dev(uint64_t* keys_dev, uint64_t n, void* stream);
ptb_status ptb_shard_trace_dev(const ptb_scene* shard, const float* origin_dir_dev, uint64_t n,
                               const uint32_t* instance_map_dev, void* const* peer_keys, int world, void* stream);
ptb_status ptb_shard_publish_dev(const ptb_scene* shard, uint64_t n, const uint64_t* best_keys_dev,
                                 void* const* peer_payload, int world, void* stream);
ptb_status ptb_shard_unpack_dev(const uint64_t* best_keys_dev, const void* payload_dev, uint64_t n, ptb_hit* hits_dev,
                                void* stream);

/* Geometry-sharded SHADOW query — intersection_worker.cpp:114-147: a shadow ray is lit only when NO worker's geometry
 * occludes it.  Every rank traces the same rays against its shard with the any-hit kernel, and an occluded ray ORs
 * its byte (1) into EVERY rank's occlusion buffer with a 32-bit atomic over NVLink, from inside the kernel.
 * peer_occluded: HOST array of `world` device pointers (rank order) to buffers of n bytes rounded up to a multiple
 * of 4, zeroed by their owners before anybody traces (memset + barrier); after a barrier behind the call every
 * rank's buffer holds the merged answer, equal to ptb_trace_occlusion on the unsharded scene. */
ptb_status ptb_shard_occlusion_dev(const ptb_scene* shard, const float* origin_dir_dev, uint64_t n, void* const* peer_occluded,
                                   int world, void* stream);

/* Host-only exercise of ptb_group's rendezvous, work-stealing counter and barrier (no CUDA): `frames` rounds in
 * which the `world` processes that pass the same `name` claim n_tiles tiles each (work_us microseconds of fake work
 * per tile); mine_out[e * n_tiles + i] = 1 where THIS rank claimed tile i of round e.  Over all ranks every tile of
 * every round is claimed exactly once.  A rank that never arrives → PTB_E_NCCL after "group_timeout_ms". */
ptb_status ptb_group_selftest_host(const char* name, int rank, int world, uint32_t n_tiles, uint32_t frames,
                                   uint32_t work_us, uint8_t* mine_out);

/* Registers per thread of the extend kernel / of its any-hit (shadow) instantiation as loaded
 * (cudaFuncGetAttributes). */
int ptb_extend_registers(void);
int ptb_shadow_registers(void);

/* The extend kernel's set-up computes the reciprocal ray direction of the slab tests
 * (LIB/geometry/aabb.cpp:41-67) through a refined reciprocal and three FMAs, the
 * fast path of an IEEE division (the split-plane distance (split - o) / d of a node
 * step, LIB/core/mesh.cpp:336-337, is a plain IEEE division).  This runs n random operand
 * pairs inside the guarded exponent window through that shortcut on the GPU and returns how
 * many differ from the correctly rounded quotient (must be 0; ~0ull on a CUDA error). */
uint64_t ptb_selftest_division(uint64_t n, uint64_t seed);

/* ----------------------------------------------------------------- misc -- */

/* Tuning options (none changes a result bit; all are for A/B measurements):
 *   "wave_paths"            paths per wavefront (default 8 Mi)
 *   "count_visits"          0/1: instrumented extend kernel (node / leaf / triangle visit counters)
 *   "time_stages"           0/1: CUDA events around every extend / shade launch (extend_seconds, shade_seconds)
 *   "extend_variant"        1: lane state machine with ray replacement (default).  Only in a library built with
 *                           PTB_BUILD_EXPERIMENTS=1 (csrc/experiments/): 0: first one-thread-per-ray kernel;
 *                           3: warp-cooperative leaf tests; 4: several ray contexts per lane, traversal state in
 *                           shared memory
 *   "extend_contexts"       rays per lane of variant 4 (2..4)
 *   "extend_steps", "extend_tests"   tree levels (2..8; instantiated: 4, 6, 8; default 0 = 6 for scenes whose meshes
 *                           average 8192+ node pairs, else 4) / triangle-test slots (1..2) offered per loop iteration
 *   "extend_defer"          1 (default): a lane registers a leaf and keeps descending while its triangles are tested
 *   "extend_dense"          1 (default, needs extend_defer): a test slot spreads the registered leaves' triangles over
 *                           all 32 lanes (a lane without a leaf tests a neighbour's triangle), and the first two
 *                           instances a ray enters are resolved 32 wide when a warp refills its pool of rays
 *   "extend_dense_min2"     lanes that must still hold a leaf for an iteration's second test slot to run (default 6)
 *   "extend_setup_lanes"    waiting lanes that trigger the set-up section (1..32, default 8)
 *   "extend_sm_ranges"      0/1: every SM starts on its own contiguous range of the ray queue (default 0)
 *   "path_order"            1: a wave's samples of one 8x4 pixel block are adjacent in the queue (default);
 *                           0: sample planes
 *   "extend_blocks_per_sm", "shade_blocks_per_sm"   caps on the persistent grids
 *   "extend_rays_per_lane"  extend blocks beyond ceil(rays / (128 x this)) leave at once, so that a small launch
 *                           does not hold the whole machine (default 8; 0 = every block stays)
 *   "frame_tiles_in_flight" multi-GPU frame: streams (host threads) per GPU (default 8)
 *   "frame_guided_tiles"    library-chosen tiling for several ranks: 1 = big tiles first, small last (default)
 *   "frame_queue_depth"     multi-GPU frame: tiles queued per stream (1 or 2)
 *   "frame_spin_wait"       multi-GPU frame: 1 = workers spin on their tile's event instead of sleeping (default 0)
 *   "group_timeout_ms"      multi-GPU: longest wait at a barrier / rendezvous before PTB_E_NCCL (default 120 000)
 * Unknown names or out-of-range values → PTB_E_INVALID.
 * Environment: PTB_OPTIONS="name=value,name=value" applies options when the library is loaded;
 * PTB_KD_CACHE=<directory> caches flattened KD trees on disk (keyed by a hash of everything a tree depends
 * on, checksummed; a damaged or foreign file is ignored and rebuilt). */
ptb_status ptb_set_option(const char* name, int64_t value);

const char* ptb_last_error(void);
int ptb_abi_version(void);
int ptb_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PTB_H */
