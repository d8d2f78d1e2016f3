// kernels.hpp — launch interface between the wavefront driver (render.cu) and kernels.cu.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "device_scene.hpp"

namespace ptb {

// Hit record written by extend and read by shade: 16 bytes.
//   x = instance << 12 | surface ordinal   (0xFFFFFFFF: miss)
//   y = triangle index inside the mesh
//   z = beta bits, w = gamma bits          (alpha = 1 - beta - gamma)
constexpr uint32_t HIT_MISS = 0xFFFFFFFFu;
constexpr uint32_t HIT_SURFACE_BITS = 12;

struct WaveGeom {
    uint32_t full_w, full_h;       // frame resolution (camera aspect, ndc)
    uint32_t x0, y0, w, h;         // tile
    uint32_t blocks_x, blocks_y;   // tile in 8x4-pixel blocks
    uint32_t sblocks_x;            // tile width in super-blocks of 8x8 blocks (64x32 pixels)
    uint32_t padded_pixels;        // slots: super-blocks * 64 * 32
    uint32_t wave_samples;         // samples of every pixel in this wave
    uint32_t first_sample;         // global index of the wave's first sample
    uint32_t block_major;          // path order: 1 = the wave's samples of an 8x4 block adjacent, 0 = sample planes
    // COMB tiles (frame.cu): the tile's columns / rows are granules of comb_gx / comb_gy pixels that lie comb_sx /
    // comb_sy pixels apart in the frame (0: a plain rectangle).  Tile pixel (x, y) is frame pixel
    // (x0 + comb_x(x), y0 + comb_y(y)); only ray generation, the RNG's pixel id and the accumulate store care.
    uint32_t comb_gx, comb_sx, comb_gy, comb_sy;
    __host__ __device__ uint32_t comb_x(uint32_t x) const { return comb_gx ? (x / comb_gx) * comb_sx + x % comb_gx : x; }
    __host__ __device__ uint32_t comb_y(uint32_t y) const { return comb_gy ? (y / comb_gy) * comb_sy + y % comb_gy : y; }
};

struct RenderParams {
    uint32_t seed_lo, seed_hi;
    uint32_t max_depth;
    uint32_t integrator; // 0: LIB (renderer::trace), 1: APP_RR (worker::trace_iter)
    uint32_t first_sample_unjittered;
};

struct PathBuffers { // one of the two ping-pong sets, each array `capacity` float4
    float4* ray_o;
    float4* ray_d;
    float4* thr;
    float4* rad;
};

struct DeviceCounters {
    unsigned long long node_visits, leaf_visits, tri_tests, rays;
    unsigned long long bound_errors; // instrumented kernel only (count_visits): indices / stack heights out of range
};

constexpr uint32_t QHEAD_STRIDE = 256; // work heads per extend launch: one per SM range (>= SM count)

struct LaunchCfg {
    int sm_count;
    int extend_blocks_per_sm;
    int shade_blocks_per_sm;
    bool count_visits;
    int extend_variant; // 0: one thread per ray (kernels.cu), 1: lane state machine (extend.cu)
    int extend_steps, extend_tests; // node steps / triangle tests offered per main-loop iteration (variant 1)
    int extend_setup_lanes;         // waiting lanes that trigger the set-up section (variant 1)
    int extend_defer;               // 1: leaves are registered and tested while the lane keeps descending (extend.cu: DEFER)
    int extend_sm_ranges;           // 1: every SM works through its own contiguous part of the queue first
    int extend_contexts;            // rays per lane of the context kernel (variant 4): 2..4
    int extend_rays_per_lane;       // blocks beyond ceil(n / (128 x this)) leave at once (0: all blocks stay)
    int extend_dense_min2;          // DENSE: a 2nd test slot runs only with this many lanes still holding a leaf
    int extend_dense;               // 1 (with extend_defer): leaf tests spread over the whole warp (extend.cu: DENSE)
};

// qcount[i] = number of live paths entering iteration i; qhead[i] = extend's work head for iteration i.
void launch_raygen(const DScene& S, const WaveGeom& g, const RenderParams& rp, const PathBuffers& out,
                   float4* sample_out, uint32_t* qcount0, const LaunchCfg& cfg, cudaStream_t st);
// scenes with a sun: shade event → shadow queue (sh_o / sh_d, compacted; shadow_slot[k] = position or 0xFFFFFFFF)
void launch_shadow_gen(const DScene& S, const WaveGeom& g, const RenderParams& rp, const float4* ray_o,
                       const float4* ray_d, const uint4* hits, float4* sh_o, float4* sh_d, uint32_t* shadow_slot,
                       const uint32_t* n_ptr, uint32_t* n_shadow, const LaunchCfg& cfg, cudaStream_t st);
// shadow_slot / occluded: what launch_shadow_gen and launch_extend_anyhit produced (null without a sun)
void launch_shade(const DScene& S, const WaveGeom& g, const RenderParams& rp, const PathBuffers& in,
                  const uint4* hits, const PathBuffers& out, float4* sample_out, const uint32_t* n_ptr,
                  uint32_t* n_next, const uint32_t* shadow_slot, const uint8_t* occluded, const LaunchCfg& cfg,
                  cudaStream_t st);
// dst: the tile's pixel (0,0) inside a buffer with `pitch` pixels per row (own or peer-mapped memory);
// fresh: this wave starts the running mean (dst / claimed are not read)
void launch_accumulate(const WaveGeom& g, const float4* sample_out, float4* dst, uint32_t pitch, uint8_t* claimed,
                       bool transparent, bool fresh, cudaStream_t st);
void launch_tonemap(const float* rgb, const float* alpha, uint64_t n, uint8_t* rgba8, cudaStream_t st);
void launch_split_rgba(const float4* rgba, uint64_t n, float* rgb, float* alpha, cudaStream_t st);
void launch_join_rgba(const float* rgb, const float* alpha, uint64_t n, float4* rgba, cudaStream_t st);
void launch_tonemap_rgba(const float4* rgba, uint64_t n, uint8_t* rgba8, cudaStream_t st);

// explicit ray sets (ptb_trace_rays): normalise directions like geometry::ray's constructor, then export
void launch_prep_rays(const float* origin_dir, uint64_t n, float4* ray_o, float4* ray_d, cudaStream_t st);
void launch_export_hits(const DScene& S, const uint4* hits, const float* t, uint64_t n, void* hits_out /*ptb_hit*/,
                        float* attrs_out /*14 per ray or null*/, cudaStream_t st);
void launch_camera_rays(const DScene& S, uint32_t w, uint32_t h, const uint32_t* px, const uint32_t* py,
                        const float* aa, uint64_t n, float* origin_dir, cudaStream_t st);

// geometry-shard merge over peer memory: every rank's key / payload buffer (symmetric allocations), by value
constexpr int SHARD_MAX_WORLD = 16;
struct ShardPeers {
    int world;
    unsigned long long* keys[SHARD_MAX_WORLD];
    uint4* payload[SHARD_MAX_WORLD];
};
// what the fused extend kernel needs to publish its results (device memory, read at result-write time only)
struct MergeArgs {
    ShardPeers peers;
    const uint32_t* instance_map;   // shard-local → global instance index
    unsigned long long* local_keys; // this rank's own key per ray (ptb_shard_publish_dev compares against it)
};
constexpr unsigned long long MERGE_MISS_KEY = 0x7FFFFFFFFFFFFFFFull;

void launch_shard_keys(const uint4* hits, const float* t, uint64_t n, const uint32_t* instance_map,
                       unsigned long long* local_keys, const ShardPeers& peers, cudaStream_t st);
void launch_shard_payload(const uint4* hits, const unsigned long long* local_keys, const unsigned long long* best_keys,
                          uint64_t n, const ShardPeers& peers, cudaStream_t st);
void launch_shard_unpack(const unsigned long long* best_keys, const uint4* payload, uint64_t n, void* hits_out,
                         cudaStream_t st);
void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t st);

// extend.cu — the lane-state-machine closest-hit kernel
void launch_extend_lanes(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                         const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                         cudaStream_t st);
// the same kernel with the geometry-shard exchange fused into its result write: every finished ray's key goes
// straight into all ranks' key buffers (64-bit atomicMin over NVLink) — compute and collective in one kernel
void launch_extend_lanes_merge(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                               const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const MergeArgs* merge_dev,
                               const LaunchCfg& cfg, cudaStream_t st);
// the shadow kernel with the geometry-shard exchange fused in: an occluded ray ORs its byte into every rank's
// occlusion buffer (merge_dev->peers.keys[r] reinterpreted as bytes) over NVLink
void launch_extend_anyhit_merge(const DScene& S, const float4* ray_o, const float4* ray_d, uint8_t* occluded,
                                const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const MergeArgs* merge_dev,
                                const LaunchCfg& cfg, cudaStream_t st);
// the SHADOW kernel: the any-hit instantiation of the same kernel; occluded[k] = 1 when ray k hits anything
void launch_extend_anyhit(const DScene& S, const float4* ray_o, const float4* ray_d, uint8_t* occluded,
                          const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                          cudaStream_t st);
int extend_lanes_regs_per_thread(bool defer, bool dense);
// the dense kernel's per-stream scratch, allocated ahead of a frame (a cudaMalloc inside it would stall every tile in flight)
void extend_reserve_scratch(const LaunchCfg& cfg, cudaStream_t st);
int extend_anyhit_regs_per_thread();

// ---- experiments (csrc/experiments/, built only with PTB_BUILD_EXPERIMENTS=1; option extend_variant) ----
// extend_simple.cu — the first kernel: one thread per ray, shared-memory stack
void launch_extend(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                   const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                   cudaStream_t st);
// extend_coop.cu — lane state machine + warp-cooperative leaf tests
void launch_extend_coop(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                        const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                        cudaStream_t st);
int extend_coop_regs_per_thread();
// extend_ctx.cu — several ray contexts per lane, traversal state in shared memory
void launch_extend_ctx(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                       const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                       cudaStream_t st);
int extend_ctx_regs_per_thread(int contexts);
unsigned long long division_selftest(uint64_t n, uint64_t seed);

int extend_regs_per_thread();

} // namespace ptb
