// render.cu — the wavefront driver: what renderer::render's sample/row loops
// (LIB/core/renderer.cpp:354-407) and the worker's queue pipeline
// (APP/processors/worker/worker.cpp:46-68) become on one GPU.
//
// One wave = wave_samples consecutive samples of every pixel of the tile.  For
// each wave: raygen, then at most max_depth iterations of
//     extend (closest hit for every live path) → shade (terminate or emit the next ray)
// with the queue size of iteration i held in device memory (qcount[i]), so the
// host enqueues a whole wave without waiting; accumulate then folds the wave's
// samples into the running mean in sample order.  Extra iterations (stochastic
// opacity, shadow-catcher pass-through do not consume a bounce) are driven by
// one host read-back per extra iteration and only happen for such scenes.

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "errors.hpp"
#include "kernels.hpp"
#include "render.hpp"
#include "scene.hpp"

namespace ptb {

Options g_options;

namespace {

constexpr uint32_t MAX_EXTRA_ITERS = 256; // pass-through events per path beyond max_depth before giving up

struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void ensure(size_t need) {
        if (need <= bytes) return;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        PTB_CUDA(cudaMalloc(&p, need));
        bytes = need;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

// Per-device scratch, reused across calls (allocation is not part of the hot path).
struct Workspace {
    DeviceBuf path[2][4]; // ping-pong × {ray_o, ray_d, thr, rad}
    DeviceBuf hits, t, sample_out, counters, qcount, accum, claimed, io_a, io_b, io_c;
    DeviceBuf sh_o, sh_d, sh_slot, sh_occluded; // shadow queue of a shade event (scenes with a sun)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> stage_ev; // pool for per-launch timing (option time_stages)
    std::mutex lock;
    uint64_t launches = 0, paths = 0;  // since the last stream_counters_reset (frame driver)
    cudaEvent_t stage_event(size_t i) {
        while (stage_ev.size() <= i) {
            cudaEvent_t e;
            PTB_CUDA(cudaEventCreate(&e));
            stage_ev.push_back(e);
        }
        return stage_ev[i];
    }
    void events() {
        for (auto& e : ev)
            if (!e) PTB_CUDA(cudaEventCreate(&e));
    }
};

// One workspace per (device, stream): calls on different streams (e.g. two host threads rendering
// different tiles) run concurrently on the GPU, each with its own path buffers.
Workspace& workspace(int device, cudaStream_t stream = nullptr) {
    static std::mutex m;
    static std::map<std::pair<int, cudaStream_t>, Workspace*> ws;
    std::lock_guard<std::mutex> g(m);
    Workspace*& w = ws[{device, stream}];
    if (!w) w = new Workspace;
    return *w;
}

PathBuffers path_set(Workspace& w, int i) {
    return PathBuffers{(float4*)w.path[i][0].p, (float4*)w.path[i][1].p, (float4*)w.path[i][2].p,
                       (float4*)w.path[i][3].p};
}

void run_extend(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg, cudaStream_t st) {
#ifdef PTB_BUILD_EXPERIMENTS
    if (cfg.extend_variant == 0) return launch_extend(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters, cfg, st);
    if (cfg.extend_variant == 3) return launch_extend_coop(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters, cfg, st);
    if (cfg.extend_variant == 4) return launch_extend_ctx(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters, cfg, st);
#endif
    launch_extend_lanes(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters, cfg, st);
}

LaunchCfg launch_cfg(const ptb_scene* s) {
    LaunchCfg c;
    c.sm_count = s->sm_count;
    c.extend_blocks_per_sm = (int)g_options.extend_blocks_per_sm;
    c.shade_blocks_per_sm = (int)g_options.shade_blocks_per_sm;
    c.count_visits = g_options.count_visits != 0;
    c.extend_variant = (int)g_options.extend_variant;
    // step slots per loop iteration: 6 pay on deep trees (C2: 58 node steps per ray, +1.3 %), and cost 5 % where a ray
    // takes a handful of steps per mesh (Cornell: 3.5) — chosen by the average tree size unless the option says otherwise
    c.extend_steps = g_options.extend_steps > 0 ? (int)g_options.extend_steps
                                                : (s->d.n_pairs / std::max<uint32_t>(1u, s->info.n_meshes) >= 8192u ? 6 : 4);
    c.extend_tests = (int)g_options.extend_tests;
    c.extend_setup_lanes = (int)g_options.extend_setup_lanes;
    c.extend_defer = (int)g_options.extend_defer;
    c.extend_dense = (int)g_options.extend_dense;
    c.extend_dense_min2 = (int)g_options.extend_dense_min2;
    c.extend_sm_ranges = (int)g_options.extend_sm_ranges;
    c.extend_contexts = (int)g_options.extend_contexts;
    c.extend_rays_per_lane = (int)g_options.extend_rays_per_lane;
    return c;
}

struct TileTarget {
    float4* base;
    uint32_t pitch;   // pixels per row of the destination
    uint8_t* claimed; // caller-owned claim mask (w*h bytes) or null
};

// Paths of one wave of a w x h tile and the buffers that hold them.
struct TileSizing {
    WaveGeom g{};
    uint64_t wave_samples = 1, cap = 0;
};

TileSizing size_tile(uint32_t w, uint32_t h, uint32_t spp) {
    TileSizing z;
    WaveGeom& g = z.g;
    g.w = w;
    g.h = h;
    g.blocks_x = (w + 7) / 8;
    g.blocks_y = (h + 3) / 4;
    g.sblocks_x = (g.blocks_x + 7) / 8;
    g.block_major = g_options.path_order != 0;
    const uint64_t padded = uint64_t(g.sblocks_x) * ((g.blocks_y + 7) / 8) * 64 * 32;
    if (padded >= (1ull << 31)) throw Error(PTB_E_INVALID, "tile too large");
    g.padded_pixels = (uint32_t)padded;
    uint64_t wave_samples = std::max<uint64_t>(1, (uint64_t)g_options.wave_paths / padded);
    wave_samples = std::min<uint64_t>(wave_samples, std::max<uint32_t>(spp, 1));
    while (wave_samples > 1 && wave_samples * padded >= (1ull << 32)) wave_samples--;
    z.wave_samples = wave_samples;
    z.cap = wave_samples * padded;
    return z;
}

// Every device buffer a tile of this size needs (a cudaMalloc synchronises the whole device: inside a frame with
// several tiles in flight it would stall all of them, so the frame driver sizes the workspaces up front).
void ensure_tile_buffers(Workspace& w, const TileSizing& z, uint32_t max_depth, bool sun, bool transparent, uint32_t tw,
                         uint32_t th) {
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 4; j++) w.path[i][j].ensure(z.cap * sizeof(float4));
    w.hits.ensure(z.cap * sizeof(uint4));
    w.sample_out.ensure(z.cap * sizeof(float4));
    const size_t n_counters = size_t(max_depth) + MAX_EXTRA_ITERS + 2;
    w.qcount.ensure(n_counters * (1 + QHEAD_STRIDE) * sizeof(uint32_t) * (sun ? 2 : 1));
    if (sun) {
        w.sh_o.ensure(z.cap * sizeof(float4));
        w.sh_d.ensure(z.cap * sizeof(float4));
        w.sh_slot.ensure(z.cap * sizeof(uint32_t));
        w.sh_occluded.ensure(z.cap);
    }
    if (transparent) w.claimed.ensure(size_t(tw) * th);
}

// The wavefront loop for one tile.  The caller holds w.lock.
//   dst        where the running mean lives: pixel (0,0) of the tile inside a buffer with dst.pitch pixels per
//              row (the caller's tile buffer, or a full frame — possibly peer-mapped memory of another GPU)
//   dst.claimed  transparent-background claim mask of the tile (w*h bytes) or null → workspace scratch
//   stats      non-null: counters are reset, the stream is synchronised at the end and the totals returned;
//              null: nothing is reset or read (the frame driver reads the workspace's counters once per frame)
void render_tile_locked(Workspace& w, const ptb_scene* s, const ptb_tile_req& req, const TileTarget& dst, cudaStream_t st,
                        ptb_render_stats* stats, const TileComb& comb = TileComb{}) {
    w.events();

    const TileSizing sizing = size_tile(req.w, req.h, req.spp);
    WaveGeom g = sizing.g;
    g.full_w = req.full_w; g.full_h = req.full_h;
    g.x0 = req.x0; g.y0 = req.y0;
    g.comb_gx = comb.gx; g.comb_sx = comb.sx; g.comb_gy = comb.gy; g.comb_sy = comb.sy;
    const uint64_t wave_samples = sizing.wave_samples;

    RenderParams rp{};
    rp.seed_lo = (uint32_t)req.seed;
    rp.seed_hi = (uint32_t)(req.seed >> 32);
    rp.max_depth = req.max_depth;
    rp.integrator = req.integrator;
    rp.first_sample_unjittered = req.first_sample_unjittered;

    const uint32_t n_iters_fixed = req.max_depth;
    const size_t n_counters = size_t(n_iters_fixed) + MAX_EXTRA_ITERS + 2;
    const bool sun = s->d.sun.enabled != 0;
    // per iteration: queue size + extend's work heads; with a sun the same again for the shadow queue
    ensure_tile_buffers(w, sizing, req.max_depth, sun, false, req.w, req.h);
    if (!w.counters.p) {
        w.counters.ensure(sizeof(DeviceCounters));
        PTB_CUDA(cudaMemsetAsync(w.counters.p, 0, sizeof(DeviceCounters), st));
    }
    const bool transparent = s->d.transparent_background != 0;
    uint8_t* claimed = dst.claimed;
    if (transparent && !claimed) {
        // scratch: enough for a call that starts its own running mean; chained calls must own the mask
        if (req.first_sample != 0)
            throw Error(PTB_E_INVALID, "first_sample != 0 on a transparent-background scene needs ptb_tile_req.claim_mask "
                                       "(the state that chains sample ranges is caller-owned)");
        w.claimed.ensure(size_t(req.w) * req.h);
        claimed = (uint8_t*)w.claimed.p;
    }

    uint32_t* qcount = (uint32_t*)w.qcount.p;
    uint32_t* qhead = qcount + n_counters;
    uint32_t* qshadow = qhead + n_counters * QHEAD_STRIDE; // shadow-queue sizes and heads (sun only)
    uint32_t* qshadow_head = qshadow + n_counters;
    DeviceCounters* counters = (DeviceCounters*)w.counters.p;
    const LaunchCfg cfg = launch_cfg(s);
    uint64_t launches = 0, extend_launches = 0, paths = 0;
    const bool time_stages = g_options.time_stages != 0 && stats != nullptr;
    size_t n_stage_ev = 0;            // events used: [2k], [2k+1] bracket launch k
    std::vector<uint8_t> stage_kind;  // 0 extend / shadow, 1 shade (+raygen/accumulate)
    auto stage_begin = [&](uint8_t kind) {
        if (!time_stages) return;
        PTB_CUDA(cudaEventRecord(w.stage_event(n_stage_ev++), st));
        stage_kind.push_back(kind);
    };
    auto stage_end = [&]() {
        if (!time_stages) return;
        PTB_CUDA(cudaEventRecord(w.stage_event(n_stage_ev++), st));
    };
    // one iteration of the wavefront: closest hit for every live path, [sun: shadow rays generated, resolved by
    // the any-hit kernel,] shade (terminate or emit the next ray)
    auto iteration = [&](uint32_t it, const PathBuffers& in, const PathBuffers& out, const WaveGeom& g,
                         const RenderParams& rp) {
        stage_begin(0);
        run_extend(s->d, in.ray_o, in.ray_d, (uint4*)w.hits.p, nullptr, &qcount[it], &qhead[size_t(it) * QHEAD_STRIDE],
                   counters, cfg, st);
        stage_end();
        launches++;
        extend_launches++;
        if (sun) {
            stage_begin(1);
            launch_shadow_gen(s->d, g, rp, in.ray_o, in.ray_d, (const uint4*)w.hits.p, (float4*)w.sh_o.p, (float4*)w.sh_d.p,
                              (uint32_t*)w.sh_slot.p, &qcount[it], &qshadow[it], cfg, st);
            stage_end();
            stage_begin(0);
            launch_extend_anyhit(s->d, (const float4*)w.sh_o.p, (const float4*)w.sh_d.p, (uint8_t*)w.sh_occluded.p,
                                 &qshadow[it], &qshadow_head[size_t(it) * QHEAD_STRIDE], counters, cfg, st);
            stage_end();
            launches += 2;
            extend_launches++;
        }
        stage_begin(1);
        launch_shade(s->d, g, rp, in, (const uint4*)w.hits.p, out, (float4*)w.sample_out.p, &qcount[it], &qcount[it + 1],
                     (const uint32_t*)w.sh_slot.p, (const uint8_t*)w.sh_occluded.p, cfg, st);
        stage_end();
        launches++;
    };

    // May a path need more shade events than max_depth?  Only through stochastic
    // opacity or shadow-catcher pass-through.
    const bool may_pass_through = s->has_pass_through;

    if (stats) PTB_CUDA(cudaMemsetAsync(counters, 0, sizeof(DeviceCounters), st));
    if (stats) PTB_CUDA(cudaEventRecord(w.ev[0], st));

    if (req.spp == 0 && req.first_sample == 0) { // renderer::render with sample_count 0 leaves the cleared image
        PTB_CUDA(cudaMemset2DAsync(dst.base, size_t(dst.pitch) * sizeof(float4), 0, size_t(req.w) * sizeof(float4), req.h, st));
    }
    for (uint32_t s0 = 0; s0 < req.spp; s0 += (uint32_t)wave_samples) {
        g.wave_samples = (uint32_t)std::min<uint64_t>(wave_samples, req.spp - s0);
        g.first_sample = req.first_sample + s0;
        paths += uint64_t(req.w) * req.h * g.wave_samples;
        PTB_CUDA(cudaMemsetAsync(qcount, 0, n_counters * (1 + QHEAD_STRIDE) * sizeof(uint32_t) * (sun ? 2 : 1), st));
        stage_begin(1);
        launch_raygen(s->d, g, rp, path_set(w, 0), (float4*)w.sample_out.p, &qcount[0], cfg, st);
        stage_end();
        launches++;
        int cur = 0;
        uint32_t it = 0;
        for (; it < n_iters_fixed; it++) {
            iteration(it, path_set(w, cur), path_set(w, cur ^ 1), g, rp);
            cur ^= 1;
        }
        if (may_pass_through && n_iters_fixed > 0) {
            for (uint32_t extra = 0; extra < MAX_EXTRA_ITERS; extra++, it++) {
                uint32_t live = 0;
                PTB_CUDA(cudaMemcpyAsync(&live, &qcount[it], sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                PTB_CUDA(cudaStreamSynchronize(st));
                if (live == 0) break;
                iteration(it, path_set(w, cur), path_set(w, cur ^ 1), g, rp);
                cur ^= 1;
            }
        }
        stage_begin(1);
        // the wave that starts the running mean writes without reading: a one-wave tile goes to its
        // destination (the frame, possibly on another GPU) with this single store per pixel
        launch_accumulate(g, (const float4*)w.sample_out.p, dst.base, dst.pitch, claimed, transparent,
                          g.first_sample == 0, st);
        stage_end();
        launches++;
    }
    w.launches += launches;
    w.paths += paths;
    PTB_CUDA(cudaGetLastError());

    if (stats) {
        PTB_CUDA(cudaEventRecord(w.ev[1], st));
        DeviceCounters hc{};
        PTB_CUDA(cudaMemcpyAsync(&hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
        PTB_CUDA(cudaStreamSynchronize(st));
        if (hc.bound_errors)
            throw Error(PTB_E_CUDA, "instrumented extend kernel: " + std::to_string(hc.bound_errors) +
                                        " index / stack bound violations");
        float ms = 0;
        PTB_CUDA(cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]));
        std::memset(stats, 0, sizeof(*stats));
        stats->paths = paths;
        stats->rays = hc.rays;
        stats->kernel_launches = launches;
        stats->extend_launches = extend_launches;
        stats->gpu_seconds = ms * 1e-3;
        stats->node_visits = hc.node_visits;
        stats->leaf_visits = hc.leaf_visits;
        stats->tri_tests = hc.tri_tests;
        for (size_t k = 0; k < stage_kind.size(); k++) {
            float sms = 0;
            PTB_CUDA(cudaEventElapsedTime(&sms, w.stage_ev[2 * k], w.stage_ev[2 * k + 1]));
            (stage_kind[k] == 0 ? stats->extend_seconds : stats->shade_seconds) += sms * 1e-3;
        }
    }
}

void check_tile_req(const ptb_scene* s, const ptb_tile_req& req) {
    if (!s) throw Error(PTB_E_INVALID, "scene is NULL");
    if (req.w == 0 || req.h == 0 || req.full_w == 0 || req.full_h == 0) throw Error(PTB_E_INVALID, "empty tile or frame");
    if (uint64_t(req.x0) + req.w > req.full_w || uint64_t(req.y0) + req.h > req.full_h)
        throw Error(PTB_E_INVALID, "tile exceeds the frame");
    if (req.max_depth > 255) throw Error(PTB_E_INVALID, "max_depth above 255 (the reference's bounce_count is uint8_t)");
    if (req.integrator > 1) throw Error(PTB_E_INVALID, "unknown integrator");
    if (uint64_t(req.first_sample) + req.spp >= (1ull << 32)) throw Error(PTB_E_INVALID, "sample index overflow");
}

} // namespace

// ---- entry points used by the frame driver (frame.cu) -----------------------------------------------------------

void render_tile_into(const ptb_scene* s, const ptb_tile_req& req, const TileComb& comb, float4* base, uint32_t pitch,
                      cudaStream_t st) {
    if (comb.gx || comb.gy) {
        // the last pixel of the comb must lie inside the frame (check_tile_req knows rectangles only)
        WaveGeom g{};
        g.comb_gx = comb.gx; g.comb_sx = comb.sx; g.comb_gy = comb.gy; g.comb_sy = comb.sy;
        if (!s) throw Error(PTB_E_INVALID, "scene is NULL");
        if (req.w == 0 || req.h == 0 || (comb.gx && comb.sx < comb.gx) || (comb.gy && comb.sy < comb.gy) ||
            uint64_t(req.x0) + g.comb_x(req.w - 1) >= req.full_w || uint64_t(req.y0) + g.comb_y(req.h - 1) >= req.full_h)
            throw Error(PTB_E_INVALID, "comb tile exceeds the frame");
        ptb_tile_req plain = req;
        plain.x0 = plain.y0 = 0;
        plain.w = plain.h = 1;
        check_tile_req(s, plain);
    } else {
        check_tile_req(s, req);
    }
    if (!base) throw Error(PTB_E_INVALID, "destination is NULL");
    Workspace& w = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    render_tile_locked(w, s, req, TileTarget{base, pitch, nullptr}, st, nullptr, comb);
}

void reserve_tile_workspace(const ptb_scene* s, cudaStream_t st, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth) {
    Workspace& ws = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(ws.lock);
    // (w, h) is the frame's LARGEST tile, but a smaller tile may fit more samples into a wave and so come closer to
    // wave_paths than the largest one does: size for the bound every tile obeys — a wave holds whole sample planes of
    // at most wave_paths paths (or one plane if that is already larger), never more than spp planes.  A buffer that
    // grows inside the frame costs a cudaFree, which waits for every tile in flight on the device.
    TileSizing z = size_tile(w, h, spp);
    const uint64_t planes = std::max<uint64_t>(1, std::max<uint32_t>(spp, 1));
    z.cap = std::max<uint64_t>(z.cap, std::min<uint64_t>(std::max<uint64_t>((uint64_t)g_options.wave_paths, z.g.padded_pixels),
                                                         planes * z.g.padded_pixels));
    ensure_tile_buffers(ws, z, max_depth, s->d.sun.enabled != 0, s->d.transparent_background != 0, w, h);
    ws.counters.ensure(sizeof(DeviceCounters));
    extend_reserve_scratch(launch_cfg(s), st);
}

void stream_counters_reset(int device, cudaStream_t st) {
    Workspace& w = workspace(device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    w.counters.ensure(sizeof(DeviceCounters));
    PTB_CUDA(cudaMemsetAsync(w.counters.p, 0, sizeof(DeviceCounters), st));
    w.launches = w.paths = 0;
}

void stream_counters_read(int device, cudaStream_t st, uint64_t* rays, uint64_t* paths, uint64_t* launches) {
    Workspace& w = workspace(device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    DeviceCounters hc{};
    if (w.counters.p) {
        PTB_CUDA(cudaMemcpyAsync(&hc, w.counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        PTB_CUDA(cudaStreamSynchronize(st));
    }
    if (hc.bound_errors)
        throw Error(PTB_E_CUDA, "instrumented extend kernel: " + std::to_string(hc.bound_errors) +
                                    " index / stack bound violations");
    *rays = hc.rays;
    *paths = w.paths;
    *launches = w.launches;
}

// ---- C ABI entry points ---------------------------------------------------------------------------------------

void render_tile_dev(const ptb_scene* s, const ptb_tile_req& req, float4* rgba_dev, cudaStream_t st,
                     ptb_render_stats* stats) {
    check_tile_req(s, req);
    if (!rgba_dev) throw Error(PTB_E_INVALID, "output is NULL");
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    // the running mean (rgba_dev) and the claim mask (req.claim_mask) are the caller's: nothing that chains
    // sample ranges lives in the library
    render_tile_locked(w, s, req, TileTarget{rgba_dev, req.w, static_cast<uint8_t*>(req.claim_mask)}, st, stats);
}

void render_tile_host(const ptb_scene* s, const ptb_tile_req& req, float* rgb_out, float* alpha_out,
                      ptb_render_stats* stats) {
    check_tile_req(s, req);
    if (!rgb_out) throw Error(PTB_E_INVALID, "output is NULL");
    PTB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = nullptr;
    Workspace& w = workspace(s->device, st);
    // ONE lock from the first allocation to the last copy: two host threads on the same device share this
    // workspace and must not free / overwrite each other's buffers
    std::lock_guard<std::mutex> guard(w.lock);
    const size_t npix = size_t(req.w) * req.h;
    const bool transparent = s->d.transparent_background != 0;
    const bool chained = req.first_sample != 0;
    if (chained && !alpha_out) throw Error(PTB_E_INVALID, "first_sample != 0 needs alpha_out (the running mean is in/out)");
    if (chained && transparent && !req.claim_mask)
        throw Error(PTB_E_INVALID, "first_sample != 0 on a transparent-background scene needs ptb_tile_req.claim_mask");
    w.accum.ensure(npix * sizeof(float4));
    w.io_a.ensure(npix * 3 * sizeof(float));
    w.io_b.ensure(npix * sizeof(float));
    uint8_t* mask_dev = nullptr;
    if (transparent && req.claim_mask) {
        w.io_c.ensure(npix);
        mask_dev = (uint8_t*)w.io_c.p;
    }
    if (chained) { // the caller's running mean (and mask) continue: host → device
        PTB_CUDA(cudaMemcpyAsync(w.io_a.p, rgb_out, npix * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
        PTB_CUDA(cudaMemcpyAsync(w.io_b.p, alpha_out, npix * sizeof(float), cudaMemcpyHostToDevice, st));
        launch_join_rgba((const float*)w.io_a.p, (const float*)w.io_b.p, npix, (float4*)w.accum.p, st);
        if (mask_dev) PTB_CUDA(cudaMemcpyAsync(mask_dev, req.claim_mask, npix, cudaMemcpyHostToDevice, st));
    }
    render_tile_locked(w, s, req, TileTarget{(float4*)w.accum.p, req.w, mask_dev}, st, stats);
    launch_split_rgba((const float4*)w.accum.p, npix, (float*)w.io_a.p, alpha_out ? (float*)w.io_b.p : nullptr, st);
    PTB_CUDA(cudaMemcpyAsync(rgb_out, w.io_a.p, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (alpha_out) PTB_CUDA(cudaMemcpyAsync(alpha_out, w.io_b.p, npix * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (mask_dev) PTB_CUDA(cudaMemcpyAsync(req.claim_mask, mask_dev, npix, cudaMemcpyDeviceToHost, st));
    PTB_CUDA(cudaStreamSynchronize(st));
}

void trace_rays_host(const ptb_scene* s, const float* origin_dir, uint64_t n, ptb_hit* hits_out, float* attrs_out,
                     ptb_render_stats* stats) {
    if (!s) throw Error(PTB_E_INVALID, "scene is NULL");
    if (n == 0) return;
    if (!origin_dir || !hits_out) throw Error(PTB_E_INVALID, "rays or hits_out is NULL");
    if (n >= (1ull << 31)) throw Error(PTB_E_INVALID, "too many rays in one call (max 2^31 - 1)");
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device);
    std::lock_guard<std::mutex> guard(w.lock);
    w.events();
    cudaStream_t st = nullptr;
    w.io_a.ensure(n * 6 * sizeof(float));
    w.io_b.ensure(n * sizeof(ptb_hit));
    if (attrs_out) w.io_c.ensure(n * 14 * sizeof(float));
    w.path[0][0].ensure(n * sizeof(float4));
    w.path[0][1].ensure(n * sizeof(float4));
    w.hits.ensure(n * sizeof(uint4));
    w.t.ensure(n * sizeof(float));
    w.qcount.ensure((1 + QHEAD_STRIDE) * sizeof(uint32_t));
    w.counters.ensure(sizeof(DeviceCounters));
    uint32_t* qc = (uint32_t*)w.qcount.p;
    const uint32_t init[1] = {(uint32_t)n};
    PTB_CUDA(cudaMemsetAsync(qc, 0, (1 + QHEAD_STRIDE) * sizeof(uint32_t), st));
    PTB_CUDA(cudaMemcpyAsync(qc, init, sizeof(init), cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaMemsetAsync(w.counters.p, 0, sizeof(DeviceCounters), st));
    PTB_CUDA(cudaMemcpyAsync(w.io_a.p, origin_dir, n * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    launch_prep_rays((const float*)w.io_a.p, n, (float4*)w.path[0][0].p, (float4*)w.path[0][1].p, st);
    PTB_CUDA(cudaEventRecord(w.ev[0], st));
    run_extend(s->d, (const float4*)w.path[0][0].p, (const float4*)w.path[0][1].p, (uint4*)w.hits.p, (float*)w.t.p,
                  &qc[0], &qc[1], (DeviceCounters*)w.counters.p, launch_cfg(s), st);
    PTB_CUDA(cudaEventRecord(w.ev[1], st));
    launch_export_hits(s->d, (const uint4*)w.hits.p, (const float*)w.t.p, n, w.io_b.p,
                       attrs_out ? (float*)w.io_c.p : nullptr, st);
    PTB_CUDA(cudaMemcpyAsync(hits_out, w.io_b.p, n * sizeof(ptb_hit), cudaMemcpyDeviceToHost, st));
    if (attrs_out) PTB_CUDA(cudaMemcpyAsync(attrs_out, w.io_c.p, n * 14 * sizeof(float), cudaMemcpyDeviceToHost, st));
    DeviceCounters hc{};
    PTB_CUDA(cudaMemcpyAsync(&hc, w.counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    PTB_CUDA(cudaStreamSynchronize(st));
    PTB_CUDA(cudaGetLastError());
    if (hc.bound_errors)
        throw Error(PTB_E_CUDA, "instrumented extend kernel: " + std::to_string(hc.bound_errors) +
                                    " index / stack bound violations");
    if (stats) {
        float ms = 0;
        PTB_CUDA(cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]));
        std::memset(stats, 0, sizeof(*stats));
        stats->rays = hc.rays;
        stats->kernel_launches = 3;
        stats->extend_launches = 1;
        stats->gpu_seconds = stats->extend_seconds = ms * 1e-3;
        stats->node_visits = hc.node_visits;
        stats->leaf_visits = hc.leaf_visits;
        stats->tri_tests = hc.tri_tests;
    }
}

// Shadow query for an explicit ray set: the any-hit kernel the wavefront uses for sun shadow rays.
void trace_occlusion_host(const ptb_scene* s, const float* origin_dir, uint64_t n, uint8_t* occluded_out,
                          ptb_render_stats* stats) {
    if (!s) throw Error(PTB_E_INVALID, "scene is NULL");
    if (n == 0) return;
    if (!origin_dir || !occluded_out) throw Error(PTB_E_INVALID, "rays or occluded_out is NULL");
    if (n >= (1ull << 31)) throw Error(PTB_E_INVALID, "too many rays in one call (max 2^31 - 1)");
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device);
    std::lock_guard<std::mutex> guard(w.lock);
    w.events();
    cudaStream_t st = nullptr;
    w.io_a.ensure(n * 6 * sizeof(float));
    w.io_b.ensure(n);
    w.path[0][0].ensure(n * sizeof(float4));
    w.path[0][1].ensure(n * sizeof(float4));
    w.qcount.ensure((1 + QHEAD_STRIDE) * sizeof(uint32_t));
    w.counters.ensure(sizeof(DeviceCounters));
    uint32_t* qc = (uint32_t*)w.qcount.p;
    const uint32_t init[1] = {(uint32_t)n};
    PTB_CUDA(cudaMemsetAsync(qc, 0, (1 + QHEAD_STRIDE) * sizeof(uint32_t), st));
    PTB_CUDA(cudaMemcpyAsync(qc, init, sizeof(init), cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaMemsetAsync(w.counters.p, 0, sizeof(DeviceCounters), st));
    PTB_CUDA(cudaMemcpyAsync(w.io_a.p, origin_dir, n * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
    launch_prep_rays((const float*)w.io_a.p, n, (float4*)w.path[0][0].p, (float4*)w.path[0][1].p, st);
    PTB_CUDA(cudaEventRecord(w.ev[0], st));
    launch_extend_anyhit(s->d, (const float4*)w.path[0][0].p, (const float4*)w.path[0][1].p, (uint8_t*)w.io_b.p, &qc[0], &qc[1],
                         (DeviceCounters*)w.counters.p, launch_cfg(s), st);
    PTB_CUDA(cudaEventRecord(w.ev[1], st));
    PTB_CUDA(cudaMemcpyAsync(occluded_out, w.io_b.p, n, cudaMemcpyDeviceToHost, st));
    DeviceCounters hc{};
    PTB_CUDA(cudaMemcpyAsync(&hc, w.counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    PTB_CUDA(cudaStreamSynchronize(st));
    PTB_CUDA(cudaGetLastError());
    if (hc.bound_errors)
        throw Error(PTB_E_CUDA, "instrumented shadow kernel: " + std::to_string(hc.bound_errors) +
                                    " index / stack bound violations");
    if (stats) {
        float ms = 0;
        PTB_CUDA(cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]));
        std::memset(stats, 0, sizeof(*stats));
        stats->rays = hc.rays;
        stats->kernel_launches = 2;
        stats->extend_launches = 1;
        stats->gpu_seconds = stats->extend_seconds = ms * 1e-3;
        stats->node_visits = hc.node_visits;
        stats->leaf_visits = hc.leaf_visits;
        stats->tri_tests = hc.tri_tests;
    }
}

// ---- device-resident closest hit and the geometry-shard merge ------------------------------------------------

namespace {

// prep + extend for rays that are already in device memory; hit records stay in the workspace (w.hits, w.t)
void trace_into_workspace(const ptb_scene* s, Workspace& w, const float* rays_dev, uint64_t n, cudaStream_t st,
                          const MergeArgs* merge_dev = nullptr) {
    w.path[0][0].ensure(n * sizeof(float4));
    w.path[0][1].ensure(n * sizeof(float4));
    w.hits.ensure(n * sizeof(uint4));
    w.t.ensure(n * sizeof(float));
    w.qcount.ensure((1 + QHEAD_STRIDE) * sizeof(uint32_t));
    w.counters.ensure(sizeof(DeviceCounters));
    uint32_t* qc = (uint32_t*)w.qcount.p;
    const uint32_t n32 = (uint32_t)n;
    PTB_CUDA(cudaMemsetAsync(qc, 0, (1 + QHEAD_STRIDE) * sizeof(uint32_t), st));
    PTB_CUDA(cudaMemcpyAsync(qc, &n32, sizeof(n32), cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaStreamSynchronize(st)); // n32 lives on this stack frame
    PTB_CUDA(cudaMemsetAsync(w.counters.p, 0, sizeof(DeviceCounters), st));
    launch_prep_rays(rays_dev, n, (float4*)w.path[0][0].p, (float4*)w.path[0][1].p, st);
    if (merge_dev)
        launch_extend_lanes_merge(s->d, (const float4*)w.path[0][0].p, (const float4*)w.path[0][1].p, (uint4*)w.hits.p,
                                  (float*)w.t.p, &qc[0], &qc[1], (DeviceCounters*)w.counters.p, merge_dev, launch_cfg(s),
                                  st);
    else
        run_extend(s->d, (const float4*)w.path[0][0].p, (const float4*)w.path[0][1].p, (uint4*)w.hits.p, (float*)w.t.p,
                   &qc[0], &qc[1], (DeviceCounters*)w.counters.p, launch_cfg(s), st);
}

void check_rays(const ptb_scene* s, const void* a, const void* b, uint64_t n) {
    if (!s) throw Error(PTB_E_INVALID, "scene is NULL");
    if (n && (!a || !b)) throw Error(PTB_E_INVALID, "NULL device pointer");
    if (n >= (1ull << 31)) throw Error(PTB_E_INVALID, "too many rays in one call (max 2^31 - 1)");
}

ShardPeers make_peers(void* const* keys, void* const* payload, int world) {
    if (world < 1 || world > SHARD_MAX_WORLD) throw Error(PTB_E_INVALID, "world size out of range (1..16)");
    ShardPeers p{};
    p.world = world;
    for (int r = 0; r < world; r++) {
        if ((keys && !keys[r]) || (payload && !payload[r])) throw Error(PTB_E_INVALID, "NULL peer buffer");
        p.keys[r] = keys ? static_cast<unsigned long long*>(keys[r]) : nullptr;
        p.payload[r] = payload ? static_cast<uint4*>(payload[r]) : nullptr;
    }
    return p;
}

} // namespace

void trace_rays_dev(const ptb_scene* s, const float* rays_dev, uint64_t n, ptb_hit* hits_dev, cudaStream_t st) {
    check_rays(s, rays_dev, hits_dev, n);
    if (n == 0) return;
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    trace_into_workspace(s, w, rays_dev, n, st);
    launch_export_hits(s->d, (const uint4*)w.hits.p, (const float*)w.t.p, n, hits_dev, nullptr, st);
    PTB_CUDA(cudaGetLastError());
}

void shard_reset_dev(uint64_t* keys_dev, uint64_t n, cudaStream_t st) {
    if (n && !keys_dev) throw Error(PTB_E_INVALID, "NULL device pointer");
    if (n) launch_fill_u64(reinterpret_cast<unsigned long long*>(keys_dev), n, 0x7FFFFFFFFFFFFFFFull, st);
    PTB_CUDA(cudaGetLastError());
}

void shard_trace_dev(const ptb_scene* s, const float* rays_dev, uint64_t n, const uint32_t* instance_map_dev,
                     void* const* peer_keys, int world, cudaStream_t st) {
    check_rays(s, rays_dev, instance_map_dev, n);
    const ShardPeers peers = make_peers(peer_keys, nullptr, world);
    if (n == 0) return;
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    w.io_c.ensure(n * sizeof(unsigned long long));
    if (g_options.extend_variant == 1 && !g_options.count_visits) {
        // default kernel: the exchange is fused into extend's result write (one kernel: compute + collective)
        MergeArgs args{peers, instance_map_dev, (unsigned long long*)w.io_c.p};
        w.io_b.ensure(sizeof(MergeArgs));
        PTB_CUDA(cudaMemcpyAsync(w.io_b.p, &args, sizeof(args), cudaMemcpyHostToDevice, st));
        PTB_CUDA(cudaStreamSynchronize(st)); // args lives on this stack frame
        trace_into_workspace(s, w, rays_dev, n, st, (const MergeArgs*)w.io_b.p);
    } else {
        trace_into_workspace(s, w, rays_dev, n, st);
        launch_shard_keys((const uint4*)w.hits.p, (const float*)w.t.p, n, instance_map_dev,
                          (unsigned long long*)w.io_c.p, peers, st);
    }
    PTB_CUDA(cudaGetLastError());
}

void shard_occlusion_dev(const ptb_scene* s, const float* rays_dev, uint64_t n, void* const* peer_occluded, int world,
                         cudaStream_t st) {
    check_rays(s, rays_dev, peer_occluded, n);
    const ShardPeers peers = make_peers(peer_occluded, nullptr, world);
    if (n == 0) return;
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    w.path[0][0].ensure(n * sizeof(float4));
    w.path[0][1].ensure(n * sizeof(float4));
    w.sh_occluded.ensure(n);
    w.qcount.ensure((1 + QHEAD_STRIDE) * sizeof(uint32_t));
    w.counters.ensure(sizeof(DeviceCounters));
    MergeArgs args{peers, nullptr, nullptr};
    w.io_b.ensure(sizeof(MergeArgs));
    uint32_t* qc = (uint32_t*)w.qcount.p;
    const uint32_t n32 = (uint32_t)n;
    PTB_CUDA(cudaMemsetAsync(qc, 0, (1 + QHEAD_STRIDE) * sizeof(uint32_t), st));
    PTB_CUDA(cudaMemcpyAsync(qc, &n32, sizeof(n32), cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaMemcpyAsync(w.io_b.p, &args, sizeof(args), cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaStreamSynchronize(st)); // n32 and args live on this stack frame
    PTB_CUDA(cudaMemsetAsync(w.counters.p, 0, sizeof(DeviceCounters), st));
    launch_prep_rays(rays_dev, n, (float4*)w.path[0][0].p, (float4*)w.path[0][1].p, st);
    launch_extend_anyhit_merge(s->d, (const float4*)w.path[0][0].p, (const float4*)w.path[0][1].p, (uint8_t*)w.sh_occluded.p,
                               &qc[0], &qc[1], (DeviceCounters*)w.counters.p, (const MergeArgs*)w.io_b.p, launch_cfg(s), st);
    PTB_CUDA(cudaGetLastError());
}

void shard_publish_dev(const ptb_scene* s, uint64_t n, const uint64_t* best_keys_dev, void* const* peer_payload, int world,
                       cudaStream_t st) {
    check_rays(s, best_keys_dev, peer_payload, n);
    const ShardPeers peers = make_peers(nullptr, peer_payload, world);
    if (n == 0) return;
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device, st);
    std::lock_guard<std::mutex> guard(w.lock);
    if (w.hits.bytes < n * sizeof(uint4) || w.io_c.bytes < n * sizeof(unsigned long long))
        throw Error(PTB_E_INVALID, "ptb_shard_publish_dev without a matching ptb_shard_trace_dev on this stream");
    launch_shard_payload((const uint4*)w.hits.p, (const unsigned long long*)w.io_c.p,
                         reinterpret_cast<const unsigned long long*>(best_keys_dev), n, peers, st);
    PTB_CUDA(cudaGetLastError());
}

void shard_unpack_dev(const uint64_t* best_keys_dev, const void* payload_dev, uint64_t n, ptb_hit* hits_dev,
                      cudaStream_t st) {
    if (n && (!best_keys_dev || !payload_dev || !hits_dev)) throw Error(PTB_E_INVALID, "NULL device pointer");
    if (n)
        launch_shard_unpack(reinterpret_cast<const unsigned long long*>(best_keys_dev), static_cast<const uint4*>(payload_dev),
                            n, hits_dev, st);
    PTB_CUDA(cudaGetLastError());
}

void camera_rays_host(const ptb_scene* s, uint32_t wd, uint32_t ht, const uint32_t* px, const uint32_t* py,
                      const float* aa, uint64_t n, float* origin_dir) {
    if (!s) throw Error(PTB_E_INVALID, "scene is NULL");
    if (n == 0) return;
    PTB_CUDA(cudaSetDevice(s->device));
    Workspace& w = workspace(s->device);
    std::lock_guard<std::mutex> guard(w.lock);
    cudaStream_t st = nullptr;
    w.io_a.ensure(n * 6 * sizeof(float));
    w.io_b.ensure(n * 2 * sizeof(uint32_t));
    w.io_c.ensure(n * 2 * sizeof(float));
    uint32_t* dpx = (uint32_t*)w.io_b.p;
    uint32_t* dpy = dpx + n;
    PTB_CUDA(cudaMemcpyAsync(dpx, px, n * 4, cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaMemcpyAsync(dpy, py, n * 4, cudaMemcpyHostToDevice, st));
    PTB_CUDA(cudaMemcpyAsync(w.io_c.p, aa, n * 8, cudaMemcpyHostToDevice, st));
    launch_camera_rays(s->d, wd, ht, dpx, dpy, (const float*)w.io_c.p, n, (float*)w.io_a.p, st);
    PTB_CUDA(cudaMemcpyAsync(origin_dir, w.io_a.p, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    PTB_CUDA(cudaStreamSynchronize(st));
    PTB_CUDA(cudaGetLastError());
}

void tonemap_host(const float* rgb, const float* alpha, uint64_t n, uint8_t* rgba8) {
    if (n == 0) return;
    if (!rgb || !rgba8) throw Error(PTB_E_INVALID, "rgb or rgba8_out is NULL");
    int dev = 0;
    PTB_CUDA(cudaGetDevice(&dev));
    Workspace& w = workspace(dev);
    std::lock_guard<std::mutex> guard(w.lock);
    cudaStream_t st = nullptr;
    w.io_a.ensure(n * 3 * sizeof(float));
    w.io_b.ensure(n * sizeof(float));
    w.io_c.ensure(n * 4);
    PTB_CUDA(cudaMemcpyAsync(w.io_a.p, rgb, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (alpha) PTB_CUDA(cudaMemcpyAsync(w.io_b.p, alpha, n * sizeof(float), cudaMemcpyHostToDevice, st));
    launch_tonemap((const float*)w.io_a.p, alpha ? (const float*)w.io_b.p : nullptr, n, (uint8_t*)w.io_c.p, st);
    PTB_CUDA(cudaMemcpyAsync(rgba8, w.io_c.p, n * 4, cudaMemcpyDeviceToHost, st));
    PTB_CUDA(cudaStreamSynchronize(st));
    PTB_CUDA(cudaGetLastError());
}

} // namespace ptb
