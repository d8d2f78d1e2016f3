// render.hpp — entry points of render.cu used by the C ABI layer (api.cu).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "ptb.h"

struct ptb_scene;

namespace ptb {

struct Options {
    int64_t wave_paths = 8ll << 20;     // paths per wavefront (rounded to whole samples of the tile)
    int64_t count_visits = 0;           // instrumented extend kernel (node / leaf / triangle counters)
    int64_t extend_blocks_per_sm = 16;  // persistent extend grid = SMs x this
    int64_t shade_blocks_per_sm = 32;
    int64_t extend_variant = 1;         // 0: one thread per ray, 1: lane state machine with ray replacement
    int64_t extend_steps = 0, extend_tests = 2; // work offered per main-loop iteration of the lane kernel (steps 0: by tree size)
    int64_t extend_setup_lanes = 8;             // lanes that must be waiting before the set-up section runs
    int64_t extend_dense_min2 = 6;              // DENSE: pending lanes the second test slot of an iteration asks for
    int64_t extend_dense = 1;                   // 1 (with extend_defer): leaf tests spread over the whole warp
    int64_t extend_defer = 1;                   // 1: leaves are registered and tested while the lane keeps descending
    int64_t path_order = 1;                     // 1: samples of an 8x4 block adjacent in the queue, 0: sample planes
    int64_t extend_contexts = 2;                // rays per lane of the context kernel (variant 4)
    int64_t extend_rays_per_lane = 8;           // extend blocks beyond ceil(rays / (128 x this)) exit at once (0 = off)
    int64_t extend_sm_ranges = 0;               // every SM starts on its own contiguous part of the ray queue
    int64_t time_stages = 0;            // CUDA-event pair around every extend / shade launch (perturbs the total)
    int64_t group_timeout_ms = 120000;  // multi-GPU: longest wait at a barrier / rendezvous before PTB_E_NCCL
    int64_t frame_tiles_in_flight = 8;  // multi-GPU frame: host threads (streams) per GPU
    int64_t frame_comb_tiles = 1;       // library-chosen tiling: 1 = comb tiles (frame.cu) for several ranks, 2 = always (tests), 0 = never
    int64_t frame_comb_rounds = 0;      // comb tiles per stream (0: as many as keep a tile under 32 M paths)
    int64_t frame_guided_tiles = 1;     // library-chosen tiling for several ranks: big tiles first, small tiles last
    int64_t frame_spin_wait = 0;        // workers wait for their tiles by spinning instead of sleeping on a blocking event
    int64_t frame_queue_depth = 1;      // tiles queued per stream (1: claim after the previous tile finished, 2: one ahead)
};
extern Options g_options;

void render_tile_dev(const ptb_scene* s, const ptb_tile_req& req, float4* rgba_dev, cudaStream_t st,
                     ptb_render_stats* stats);
void render_tile_host(const ptb_scene* s, const ptb_tile_req& req, float* rgb_out, float* alpha_out,
                      ptb_render_stats* stats);
// frame driver: one tile accumulated in place at `base` (pitch pixels per row, own or peer-mapped memory),
// asynchronous on `st`; rays / paths / launches add up in the stream's workspace until they are read
// comb: granule / stride in x and y (kernels.hpp: WaveGeom::comb_*), all 0 for a plain rectangle; req.w / req.h are
// then the pixels the tile covers, req.x0 / req.y0 its first granule
struct TileComb {
    uint32_t gx = 0, sx = 0, gy = 0, sy = 0;
};
void render_tile_into(const ptb_scene* s, const ptb_tile_req& req, const TileComb& comb, float4* base, uint32_t pitch,
                      cudaStream_t st);
// sizes the stream's workspace for w x h tiles before a frame starts (no allocation inside the frame)
void reserve_tile_workspace(const ptb_scene* s, cudaStream_t st, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth);
void stream_counters_reset(int device, cudaStream_t st);
void stream_counters_read(int device, cudaStream_t st, uint64_t* rays, uint64_t* paths, uint64_t* launches);
void trace_rays_host(const ptb_scene* s, const float* origin_dir, uint64_t n, ptb_hit* hits_out, float* attrs_out,
                     ptb_render_stats* stats);
void trace_occlusion_host(const ptb_scene* s, const float* origin_dir, uint64_t n, uint8_t* occluded_out,
                          ptb_render_stats* stats);
void trace_rays_dev(const ptb_scene* s, const float* rays_dev, uint64_t n, ptb_hit* hits_dev, cudaStream_t st);
void shard_reset_dev(uint64_t* keys_dev, uint64_t n, cudaStream_t st);
void shard_trace_dev(const ptb_scene* s, const float* rays_dev, uint64_t n, const uint32_t* instance_map_dev,
                     void* const* peer_keys, int world, cudaStream_t st);
void shard_occlusion_dev(const ptb_scene* s, const float* rays_dev, uint64_t n, void* const* peer_occluded, int world,
                         cudaStream_t st);
void shard_publish_dev(const ptb_scene* s, uint64_t n, const uint64_t* best_keys_dev, void* const* peer_payload, int world,
                       cudaStream_t st);
void shard_unpack_dev(const uint64_t* best_keys_dev, const void* payload_dev, uint64_t n, ptb_hit* hits_dev,
                      cudaStream_t st);
void camera_rays_host(const ptb_scene* s, uint32_t w, uint32_t h, const uint32_t* px, const uint32_t* py,
                      const float* aa, uint64_t n, float* origin_dir);
void tonemap_host(const float* rgb, const float* alpha, uint64_t n, uint8_t* rgba8);

} // namespace ptb
