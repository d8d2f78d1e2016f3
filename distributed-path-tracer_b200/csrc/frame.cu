// frame.cu — the multi-GPU frame driver: ptb_group (one process per GPU), ptb_ctx (one process, one host
// thread per GPU) and the engine both share.
//
// What it replaces in the reference: renderer::render's thread pool over scanlines with a barrier per sample
// (LIB/core/renderer.cpp:354-407) and worker::run's stage threads + monitor (APP/processors/worker/worker.cpp:25-105).
// The reference's cross-worker transport (SNS/SQS, provisioned by PRE/app.py:110-112) was never written; on one
// NVSwitch box none is needed:
//
//   * a RANK is a GPU.  The image is cut into tiles; ranks claim tiles from ONE shared counter with an atomic
//     fetch-add (work stealing).  The counter, the barrier and the per-rank statistics live in a GroupShared
//     block — on the heap when the ranks are threads of one process, in a POSIX shared-memory object when
//     they are processes (torchrun).
//   * the frame lives on rank 0's GPU.  Every rank's accumulate kernel stores its finished pixels straight
//     into it (peer access within a process, a CUDA IPC mapping across processes): the tile return rides on
//     the kernel's own stores over NVLink, there is no collective and no staging copy in the data path.
//     Two frames alternate, so rank 0 copies frame e to the host while everybody renders frame e + 1.
//   * each rank keeps several tiles in flight (one host thread + stream each), so the tail of one tile's
//     kernels overlaps the head of the next; workers sleep on blocking events instead of spinning (8 ranks x
//     6 workers would otherwise oversubscribe the host).
//   * every wait has a time-out: a rank that dies makes the others return PTB_E_NCCL, nobody hangs.
//
// Results do not depend on the number of ranks, the tile size or who rendered what: a tile's pixels depend only
// on (seed, pixel, sample) — the frame is bit-identical to the one-GPU frame.

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "errors.hpp"
#include "frame.hpp"
#include "kernels.hpp"
#include "render.hpp"
#include "scene.hpp"

namespace ptb {

namespace {

constexpr int GROUP_MAX_WORLD = 16;
constexpr uint32_t GROUP_MAGIC = 0x50544247u; // "PTBG"
constexpr int MAX_IN_FLIGHT = 16;

struct RankStats {
    uint64_t rays, paths, launches, tiles;
    double gpu_seconds;
};

// Shared by all ranks of a group: process heap (threads) or a shared-memory object (processes).
struct GroupShared {
    std::atomic<uint32_t> magic;      // set last by rank 0: the block is initialised
    std::atomic<uint32_t> arrived;    // barrier: ranks that have arrived in this generation
    std::atomic<uint32_t> generation; // barrier: bumped by the last rank to arrive
    std::atomic<uint32_t> failed;     // a rank gave up: every wait ends with an error
    std::atomic<uint32_t> counter[2]; // next unclaimed tile of the even / odd frame
    // frame of rank 0 (written by rank 0 between two barriers, read by the others after the second)
    uint32_t frame_w, frame_h;
    int32_t root_device;
    cudaIpcMemHandle_t frame_handle; // processes
    void* frame_ptr;                 // threads of one process
    RankStats rank[GROUP_MAX_WORLD];
};

struct Tile {
    uint32_t x0, y0, w, h;
    TileComb comb; // all 0: a rectangle
};

double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

} // namespace ptb

using namespace ptb;

// One rank of a group: the worker pool of one GPU.
struct ptb_group {
    int rank = 0, world = 1, device = 0;
    bool processes = false;
    GroupShared* sh = nullptr;
    bool owns_shared = false; // thread mode: rank 0's group frees the heap block (ptb_ctx does, in fact)
    std::string shm_name;
    size_t shm_bytes = 0;

    // the frame as this rank's device sees it (two frames back to back)
    float4* frame_view = nullptr;
    uint32_t frame_w = 0, frame_h = 0;
    bool view_is_ipc = false;
    float4* frame_owned = nullptr; // rank 0: the allocation
    uint8_t* rgba8_dev = nullptr;  // rank 0: tonemapped frame
    size_t rgba8_bytes = 0;
    void* staging = nullptr; // rank 0: pinned staging for pageable outputs
    size_t staging_bytes = 0;
    uint64_t epoch = 0;

    cudaStream_t ctl = nullptr;
    cudaEvent_t ev_start = nullptr, ev_end = nullptr;

    // worker pool
    struct Worker {
        std::thread th;
        cudaStream_t st = nullptr;
        cudaEvent_t done[2] = {nullptr, nullptr}; // blocking-sync events, ring of tiles in flight on this stream
        cudaEvent_t done_spin[2] = {nullptr, nullptr}; // the same, waited for by spinning (option frame_spin_wait)
        cudaEvent_t fin = nullptr;
        uint64_t rays = 0, paths = 0, launches = 0, tiles = 0;
    };
    std::vector<Worker*> workers;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    uint64_t job_seq = 0;
    int job_workers = 0; // workers that take part in the current job
    int pending = 0;
    bool quit = false;
    // the current job
    const ptb_scene* job_scene = nullptr;
    ptb_frame_req job_req{};
    const std::vector<Tile>* job_tiles = nullptr;
    std::atomic<uint32_t>* job_counter = nullptr;
    float4* job_frame = nullptr;
    int job_depth = 1;
    bool job_spin = false;
    uint32_t job_max_w = 0, job_max_h = 0; // the largest tile of the frame: every worker's buffers are sized for it
    std::string job_error;
    ptb_status job_status = PTB_OK;
};

namespace ptb {

namespace {

[[noreturn]] void group_fail(ptb_group* g, const std::string& what) {
    if (g->sh) g->sh->failed.store(1, std::memory_order_release);
    throw Error(PTB_E_NCCL, what);
}

// Sense-reversing barrier over the shared block, with a time-out.
void group_barrier(ptb_group* g) {
    GroupShared* sh = g->sh;
    if (g->world == 1) return;
    if (sh->failed.load(std::memory_order_acquire)) throw Error(PTB_E_NCCL, "group: another rank failed");
    const uint32_t gen = sh->generation.load(std::memory_order_acquire);
    if (sh->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)g->world) {
        sh->arrived.store(0, std::memory_order_relaxed);
        sh->generation.store(gen + 1, std::memory_order_release);
        return;
    }
    const double t0 = now_seconds();
    const double limit = double(g_options.group_timeout_ms) * 1e-3;
    for (uint32_t spin = 0;; spin++) {
        if (sh->generation.load(std::memory_order_acquire) != gen) return;
        if (sh->failed.load(std::memory_order_acquire)) throw Error(PTB_E_NCCL, "group: another rank failed");
        if (spin < 4096) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        } else {
            if ((spin & 63u) == 0 && now_seconds() - t0 > limit)
                group_fail(g, "group barrier timed out after " + std::to_string(g_options.group_timeout_ms) + " ms (rank " +
                                  std::to_string(g->rank) + ")");
            usleep(20);
        }
    }
}

// ---- worker pool --------------------------------------------------------------------------------------------------

void worker_main(ptb_group* g, ptb_group::Worker* w, int index) {
    cudaSetDevice(g->device);
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(g->m);
            g->cv_job.wait(lk, [&] { return g->quit || g->job_seq != seen; });
            if (g->quit) return;
            seen = g->job_seq;
            if (index >= g->job_workers) continue; // not part of this job
        }
        try {
            const ptb_frame_req& fr = g->job_req;
            const std::vector<Tile>& tiles = *g->job_tiles;
            const uint32_t n_tiles = (uint32_t)tiles.size();
            reserve_tile_workspace(g->job_scene, w->st, g->job_max_w, g->job_max_h, fr.spp, fr.max_depth);
            PTB_CUDA(cudaStreamWaitEvent(w->st, g->ev_start, 0));
            stream_counters_reset(g->device, w->st);
            bool outstanding[2] = {false, false};
            int slot = 0;
            w->tiles = 0;
            cudaEvent_t* done = g->job_spin ? w->done_spin : w->done;
            for (;;) {
                if (outstanding[slot]) { // keep at most job_depth tiles queued on this stream
                    PTB_CUDA(cudaEventSynchronize(done[slot]));
                    outstanding[slot] = false;
                }
                if (g->sh->failed.load(std::memory_order_acquire)) throw Error(PTB_E_NCCL, "group: another rank failed");
                const uint32_t i = g->job_counter->fetch_add(1, std::memory_order_relaxed); // the work-stealing claim
                if (i >= n_tiles) break;
                const Tile& t = tiles[i];
                ptb_tile_req tr{};
                tr.full_w = fr.full_w;
                tr.full_h = fr.full_h;
                tr.x0 = t.x0;
                tr.y0 = t.y0;
                tr.w = t.w;
                tr.h = t.h;
                tr.spp = fr.spp;
                tr.max_depth = fr.max_depth;
                tr.seed = fr.seed;
                tr.first_sample = 0;
                tr.integrator = fr.integrator;
                tr.first_sample_unjittered = fr.first_sample_unjittered;
                // straight into the frame on rank 0's GPU: base = the tile's first pixel, pitch = the frame's width
                render_tile_into(g->job_scene, tr, t.comb, g->job_frame + size_t(t.y0) * fr.full_w + t.x0, fr.full_w, w->st);
                PTB_CUDA(cudaEventRecord(done[slot], w->st));
                outstanding[slot] = true;
                slot = (slot + 1) % g->job_depth;
                w->tiles++;
            }
            stream_counters_read(g->device, w->st, &w->rays, &w->paths, &w->launches); // synchronises the stream
            PTB_CUDA(cudaEventRecord(w->fin, w->st));
        } catch (const Error& e) {
            std::lock_guard<std::mutex> lk(g->m);
            if (g->job_status == PTB_OK) {
                g->job_status = e.code;
                g->job_error = e.what();
            }
            if (g->sh) g->sh->failed.store(1, std::memory_order_release);
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> lk(g->m);
            if (g->job_status == PTB_OK) {
                g->job_status = PTB_E_INVALID;
                g->job_error = e.what();
            }
            if (g->sh) g->sh->failed.store(1, std::memory_order_release);
        }
        {
            std::lock_guard<std::mutex> lk(g->m);
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

void ensure_workers(ptb_group* g, int n) {
    while ((int)g->workers.size() < n) {
        auto* w = new ptb_group::Worker;
        PTB_CUDA(cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking));
        for (auto& e : w->done) PTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventBlockingSync | cudaEventDisableTiming));
        for (auto& e : w->done_spin) PTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        PTB_CUDA(cudaEventCreateWithFlags(&w->fin, cudaEventDisableTiming));
        const int index = (int)g->workers.size();
        g->workers.push_back(w);
        w->th = std::thread(worker_main, g, w, index);
    }
}

// ---- the frame on rank 0 --------------------------------------------------------------------------------------------

void release_view(ptb_group* g) {
    if (g->frame_view && g->view_is_ipc) cudaIpcCloseMemHandle(g->frame_view);
    g->frame_view = nullptr;
    g->view_is_ipc = false;
}

// Collective.  After it every rank holds a device pointer to rank 0's two w x h frames.
void ensure_frame(ptb_group* g, uint32_t w, uint32_t h) {
    if (g->frame_view && g->frame_w == w && g->frame_h == h) return;
    GroupShared* sh = g->sh;
    const size_t bytes = size_t(w) * h * sizeof(float4) * 2;
    if (g->rank != 0) release_view(g);
    group_barrier(g); // nobody maps the old frame any more
    if (g->rank == 0) {
        g->frame_view = nullptr;
        if (g->frame_owned) PTB_CUDA(cudaFree(g->frame_owned));
        g->frame_owned = nullptr;
        PTB_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->frame_owned), bytes));
        PTB_CUDA(cudaMemset(g->frame_owned, 0, bytes));
        g->frame_view = g->frame_owned;
        sh->frame_w = w;
        sh->frame_h = h;
        sh->root_device = g->device;
        if (g->processes && g->world > 1) {
            PTB_CUDA(cudaIpcGetMemHandle(&sh->frame_handle, g->frame_owned));
        } else {
            sh->frame_ptr = g->frame_owned;
        }
    }
    group_barrier(g); // the new frame is published
    if (g->rank != 0) {
        if (g->processes) {
            void* p = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&p, sh->frame_handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                group_fail(g, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) + " (rank " + std::to_string(g->rank) +
                                  ")");
            }
            g->frame_view = static_cast<float4*>(p);
            g->view_is_ipc = true;
        } else {
            if (sh->root_device != g->device) {
                int can = 0;
                PTB_CUDA(cudaDeviceCanAccessPeer(&can, g->device, sh->root_device));
                if (!can)
                    group_fail(g, "GPU " + std::to_string(g->device) + " cannot access GPU " + std::to_string(sh->root_device) +
                                      " as a peer");
                const cudaError_t e = cudaDeviceEnablePeerAccess(sh->root_device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    group_fail(g, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
            }
            g->frame_view = static_cast<float4*>(sh->frame_ptr);
        }
    }
    g->frame_w = w;
    g->frame_h = h;
}

// Pixels one comb phase covers along an axis: granules c, c + C, c + 2C, ... of G pixels (the frame's last granule
// may be partial).
uint32_t comb_extent(uint32_t full, uint32_t G, uint32_t C, uint32_t c) {
    const uint32_t n_gran = (full + G - 1) / G;
    if (c >= n_gran) return 0;
    const uint32_t mine = (n_gran - 1 - c) / C + 1;
    const uint32_t last = c + (mine - 1) * C;
    return (mine - 1) * G + (last == n_gran - 1 ? full - last * G : G);
}

// COMB tiles, the library's choice for several ranks.  A fixed frame cut over N GPUs leaves each of them little work
// (1080p x 64 spp over 8 B200s: 26 ms), and a GPU is only fast on LARGE launches — while rectangular tiles of very
// different cost (sky against terrain) need to be SMALL to balance the ranks at the end of the frame.  A comb tile is
// both large and average: tile (c, p) of C x P holds the pixel granules (c + i C, p + j P) of the whole frame, so
// every tile samples the entire image and all tiles cost the same (to the granule count: < 1 %).  N x in-flight x rounds
// tiles: every stream of every rank renders `rounds` tiles of full launch size, and all finish together.  Which rank
// takes which tile still does not matter (shared counter), and the pixels do not depend on the tiling at all.
std::vector<Tile> make_comb_tiles(const ptb_frame_req& r, uint32_t T) {
    const uint64_t npix = uint64_t(r.full_w) * r.full_h;
    static const uint32_t GX[] = {64, 32, 16, 8}, GY[] = {32, 16, 8, 4}; // whole 8 x 4 pixel blocks (a warp's bundle of primary rays)
    // pass 1: the smallest "largest tile" any comb achieves; pass 2: among the combs within 1 % of it, the one with
    // the largest granules and the squarest shape (rays of neighbouring blocks share tree nodes in L1 / L2)
    uint64_t min_worst = ~0ull, best_penalty = ~0ull;
    uint32_t bC = 0, bP = 0, bgx = 0, bgy = 0;
    for (int pass = 0; pass < 2; pass++)
        for (uint32_t C = 1; C <= T; C++) {
            if (T % C) continue;
            const uint32_t P = T / C;
            for (uint32_t gx : GX)
                for (uint32_t gy : GY) {
                    if (uint64_t(C) * gx > r.full_w || uint64_t(P) * gy > r.full_h) continue; // every phase needs a granule
                    // (phase 0 is never smaller than another phase)
                    const uint64_t worst = uint64_t(comb_extent(r.full_w, gx, C, 0)) * comb_extent(r.full_h, gy, P, 0);
                    if (pass == 0) {
                        min_worst = std::min(min_worst, worst);
                    } else if (worst <= min_worst + min_worst / 100) {
                        const uint64_t penalty = (64 / gx) * (32 / gy) * 16 + (C > P ? C / P : P / C);
                        if (penalty < best_penalty) {
                            best_penalty = penalty;
                            bC = C; bP = P; bgx = gx; bgy = gy;
                        }
                    }
                }
        }
    std::vector<Tile> tiles;
    // no even comb, or tiles too small to be worth a launch of their own (a tiny frame): rectangles
    if (!bC || min_worst * T > npix + npix / 8) return tiles;
    if (g_options.frame_comb_tiles != 2 && min_worst * std::max<uint32_t>(r.spp, 1) < (64u << 10)) return tiles;
    for (uint32_t p = 0; p < bP; p++)
        for (uint32_t c = 0; c < bC; c++) {
            Tile t{c * bgx, p * bgy, comb_extent(r.full_w, bgx, bC, c), comb_extent(r.full_h, bgy, bP, p), TileComb{}};
            t.comb.gx = bgx; t.comb.sx = bC * bgx; t.comb.gy = bgy; t.comb.sy = bP * bgy;
            if (t.w && t.h) tiles.push_back(t);
        }
    return tiles;
}

// Tile list, in claim order.  With a tile size from the caller: a uniform grid.  Chosen by the library: a base tile
// so that every rank has ~24 tiles to steal from while a tile keeps enough paths to fill a launch (multiples of
// the 32 x 16 pixel granule of the path order) — and, for several ranks, GUIDED scheduling: the first ~60 % of the
// rows are cut into tiles of 2 x 2 base tiles, which run closer to the GPU's large-wave rate, and only the rest
// into base tiles, which are what balances the ranks at the end of the frame (big units first, small units last).
// → the frame's tiles in claim order, and the number of tiles each rank keeps in flight (host threads / streams).
std::vector<Tile> make_tiles(const ptb_frame_req& r, int world, int* workers_out = nullptr) {
    uint32_t tw = r.tile_w, th = r.tile_h;
    const bool chosen = (tw == 0 || th == 0);
    int in_flight = r.tiles_in_flight ? (int)r.tiles_in_flight : (int)g_options.frame_tiles_in_flight;
    in_flight = std::max(1, std::min(in_flight, MAX_IN_FLIGHT));
    if (workers_out) *workers_out = in_flight;
    if (chosen && ((world > 1 && g_options.frame_comb_tiles == 1) || g_options.frame_comb_tiles == 2)) {
        // Comb tiles cost the same, so every stream gets the same number of them and size is free to choose: a tile
        // of ~5 M paths runs close to the GPU's large-launch rate and 3+ of them in flight keep the SMs busy across
        // the tiles' own tails (C2 over 8 GPUs, v10 kernel: 26.9 / 27.6 / 27.3 ms per frame with 3 / 4 / 5 in flight;
        // v9: 28.4 / 28.8 / 29.7 ms with 4 / 6 / 8).  A tile
        // of more than 32 M paths is cut further (rounds), which lets a slow GPU shed work to the others.
        const double share = double(r.full_w) * r.full_h * std::max<uint32_t>(r.spp, 1) / world; // paths per GPU
        int k = in_flight;
        if (!r.tiles_in_flight) k = (int)std::max(3.0, std::min(double(in_flight), std::floor(share / double(5 << 20) + 0.5)));
        int64_t rounds = g_options.frame_comb_rounds;
        if (rounds <= 0) rounds = std::max<int64_t>(1, std::min<int64_t>(64, (int64_t)std::ceil(share / k / double(32 << 20))));
        std::vector<Tile> comb = make_comb_tiles(r, uint32_t(world) * uint32_t(k) * uint32_t(rounds));
        if (!comb.empty()) {
            if (workers_out) *workers_out = k;
            return comb;
        }
    }
    if (chosen) {
        const uint64_t want = uint64_t(world) * 24;
        tw = 32;
        th = 16;
        for (uint32_t k = 2; k <= 256; k++) {
            const uint32_t cw = 32 * k, ch = 16 * k;
            const uint64_t n = uint64_t((r.full_w + cw - 1) / cw) * ((r.full_h + ch - 1) / ch);
            if (n < want) break;
            tw = cw;
            th = ch;
        }
        // a tile should carry at least ~128 k paths
        while (uint64_t(tw) * th * std::max<uint32_t>(r.spp, 1) < (128u << 10) && (tw < r.full_w || th < r.full_h)) {
            tw *= 2;
            th *= 2;
        }
    }
    tw = std::min(tw, r.full_w);
    th = std::min(th, r.full_h);
    std::vector<Tile> tiles;
    uint32_t y0 = 0;
    if (chosen && world > 1 && g_options.frame_guided_tiles) {
        const uint32_t bw = std::min(2 * tw, r.full_w), bh = 2 * th;
        const uint32_t big_rows = uint32_t(uint64_t(r.full_h) * 6 / 10) / bh * bh;
        for (; y0 < big_rows; y0 += bh)
            for (uint32_t x = 0; x < r.full_w; x += bw) tiles.push_back(Tile{x, y0, std::min(bw, r.full_w - x), bh, TileComb{}});
    }
    for (uint32_t y = y0; y < r.full_h; y += th)
        for (uint32_t x = 0; x < r.full_w; x += tw) tiles.push_back(Tile{x, y, std::min(tw, r.full_w - x), std::min(th, r.full_h - y), TileComb{}});
    return tiles;
}

} // namespace

void frame_tiles(const ptb_frame_req& req, int world, uint32_t* out, uint64_t capacity, uint32_t* n_tiles, bool with_comb) {
    if (req.full_w == 0 || req.full_h == 0 || world < 1) throw Error(PTB_E_INVALID, "empty frame or bad world size");
    if (!n_tiles) throw Error(PTB_E_INVALID, "n_tiles is NULL");
    const std::vector<Tile> tiles = make_tiles(req, world);
    *n_tiles = (uint32_t)tiles.size();
    if (out) {
        if (capacity < tiles.size()) throw Error(PTB_E_INVALID, "capacity too small");
        const size_t stride = with_comb ? 8 : 4;
        for (size_t i = 0; i < tiles.size(); i++) {
            uint32_t* o = out + stride * i;
            o[0] = tiles[i].x0;
            o[1] = tiles[i].y0;
            o[2] = tiles[i].w;
            o[3] = tiles[i].h;
            if (with_comb) {
                o[4] = tiles[i].comb.gx;
                o[5] = tiles[i].comb.sx;
                o[6] = tiles[i].comb.gy;
                o[7] = tiles[i].comb.sy;
            }
        }
    }
}

// ---- group life cycle ------------------------------------------------------------------------------------------------

static void group_init_device(ptb_group* g) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        throw Error(PTB_E_CUDA, "no CUDA device is usable; libptb has no CPU fallback");
    if (g->device < 0 || g->device >= n_dev) throw Error(PTB_E_INVALID, "device ordinal out of range");
    PTB_CUDA(cudaSetDevice(g->device));
    PTB_CUDA(cudaStreamCreateWithFlags(&g->ctl, cudaStreamNonBlocking));
    PTB_CUDA(cudaEventCreate(&g->ev_start));
    PTB_CUDA(cudaEventCreate(&g->ev_end));
}

ptb_group* group_create_threads(void* shared_block, int rank, int world, int device) {
    auto* g = new ptb_group;
    g->rank = rank;
    g->world = world;
    g->device = device;
    g->processes = false;
    g->sh = static_cast<GroupShared*>(shared_block);
    try {
        group_init_device(g);
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

void* group_shared_alloc() {
    auto* sh = new GroupShared;
    std::memset(static_cast<void*>(sh), 0, sizeof(GroupShared));
    sh->magic.store(GROUP_MAGIC);
    return sh;
}
void group_shared_free(void* p) { delete static_cast<GroupShared*>(p); }
// thread mode only, with every rank idle: lets a context be used again after a failed frame
void group_shared_clear_failure(void* p) {
    auto* sh = static_cast<GroupShared*>(p);
    if (sh->failed.load()) {
        sh->failed.store(0);
        sh->arrived.store(0);
        sh->counter[0].store(0);
        sh->counter[1].store(0);
    }
}

ptb_group* group_create_processes(const char* name, int rank, int world, int device) {
    if (!name || !*name) throw Error(PTB_E_INVALID, "group name is empty");
    if (world < 1 || world > GROUP_MAX_WORLD || rank < 0 || rank >= world) throw Error(PTB_E_INVALID, "bad rank / world (1..16)");
    auto* g = new ptb_group;
    g->rank = rank;
    g->world = world;
    g->device = device;
    g->processes = true;
    g->shm_name = std::string("/ptb_") + name;
    g->shm_bytes = (sizeof(GroupShared) + 4095) & ~size_t(4095);
    try {
        if (device >= 0) group_init_device(g); // device < 0: host-only rendezvous (group_selftest_host)
        const int fd = shm_open(g->shm_name.c_str(), O_CREAT | O_RDWR, 0600);
        if (fd < 0) throw Error(PTB_E_NCCL, "shm_open(" + g->shm_name + ") failed: " + std::strerror(errno));
        if (ftruncate(fd, (off_t)g->shm_bytes) != 0) {
            close(fd);
            throw Error(PTB_E_NCCL, std::string("ftruncate on the group's shared memory failed: ") + std::strerror(errno));
        }
        void* p = mmap(nullptr, g->shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) throw Error(PTB_E_NCCL, std::string("mmap of the group's shared memory failed: ") + std::strerror(errno));
        g->sh = static_cast<GroupShared*>(p); // a fresh object is zero-filled: counters, barrier and flags start at 0
        if (rank == 0) {
            g->sh->magic.store(GROUP_MAGIC, std::memory_order_release);
        } else {
            const double t0 = now_seconds();
            while (g->sh->magic.load(std::memory_order_acquire) != GROUP_MAGIC) {
                if (now_seconds() - t0 > double(g_options.group_timeout_ms) * 1e-3)
                    throw Error(PTB_E_NCCL, "group rendezvous timed out waiting for rank 0");
                usleep(100);
            }
        }
        group_barrier(g);                               // everybody has mapped the object ...
        if (rank == 0) shm_unlink(g->shm_name.c_str()); // ... so its name can go: nothing is left behind in /dev/shm
    } catch (...) {
        group_destroy(g);
        throw;
    }
    return g;
}

void group_destroy(ptb_group* g) {
    if (!g) return;
    if (g->device < 0) { // host-only group: nothing of CUDA was touched
        if (g->processes && g->sh) munmap(g->sh, g->shm_bytes);
        delete g;
        return;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    {
        std::lock_guard<std::mutex> lk(g->m);
        g->quit = true;
    }
    g->cv_job.notify_all();
    for (auto* w : g->workers) {
        if (w->th.joinable()) w->th.join();
        if (w->st) cudaStreamDestroy(w->st);
        for (auto& e : w->done)
            if (e) cudaEventDestroy(e);
        for (auto& e : w->done_spin)
            if (e) cudaEventDestroy(e);
        if (w->fin) cudaEventDestroy(w->fin);
        delete w;
    }
    g->workers.clear();
    if (g->rank != 0) release_view(g);
    if (g->frame_owned) cudaFree(g->frame_owned);
    if (g->rgba8_dev) cudaFree(g->rgba8_dev);
    if (g->staging) cudaFreeHost(g->staging);
    if (g->ctl) cudaStreamDestroy(g->ctl);
    if (g->ev_start) cudaEventDestroy(g->ev_start);
    if (g->ev_end) cudaEventDestroy(g->ev_end);
    if (g->processes && g->sh) munmap(g->sh, g->shm_bytes);
    cudaSetDevice(prev);
    delete g;
}

// Host-only exercise of the rendezvous, the work-stealing counter and the barrier (no CUDA): `frames` rounds in
// which the ranks claim n_tiles tiles each; mine_out[e * n_tiles + i] = 1 where this rank claimed tile i of round e.
void group_selftest_host(const char* name, int rank, int world, uint32_t n_tiles, uint32_t frames, uint32_t work_us,
                         uint8_t* mine_out) {
    if (!mine_out && n_tiles * frames) throw Error(PTB_E_INVALID, "mine_out is NULL");
    ptb_group* g = group_create_processes(name, rank, world, -1);
    try {
        std::memset(mine_out, 0, size_t(n_tiles) * frames);
        for (uint32_t e = 0; e < frames; e++) {
            std::atomic<uint32_t>& counter = g->sh->counter[e & 1u];
            for (;;) {
                const uint32_t i = counter.fetch_add(1, std::memory_order_relaxed);
                if (i >= n_tiles) break;
                mine_out[size_t(e) * n_tiles + i] = 1;
                if (work_us) usleep(work_us);
            }
            if (rank == 0) g->sh->counter[(e + 1) & 1u].store(0, std::memory_order_release);
            group_barrier(g);
        }
    } catch (...) {
        group_destroy(g);
        throw;
    }
    group_destroy(g);
}

void group_barrier_public(ptb_group* g) {
    if (!g) throw Error(PTB_E_INVALID, "group is NULL");
    group_barrier(g);
}

// ---- one frame --------------------------------------------------------------------------------------------------------

void group_render_frame(ptb_group* g, const ptb_scene* scene, const ptb_frame_req& req, void* out_host, ptb_frame_stats* stats) {
    if (!g || !scene) throw Error(PTB_E_INVALID, "group or scene is NULL");
    if (scene->device != g->device) throw Error(PTB_E_INVALID, "the scene replica lives on another device than the group's rank");
    if (req.full_w == 0 || req.full_h == 0) throw Error(PTB_E_INVALID, "empty frame");
    if (uint64_t(req.full_w) * req.full_h >= (1ull << 31)) throw Error(PTB_E_INVALID, "frame too large");
    if (req.max_depth > 255) throw Error(PTB_E_INVALID, "max_depth above 255 (the reference's bounce_count is uint8_t)");
    if (req.integrator > 1) throw Error(PTB_E_INVALID, "unknown integrator");
    if (req.output > PTB_OUT_RGBA8) throw Error(PTB_E_INVALID, "unknown output format");
    if (req.output != PTB_OUT_NONE && g->rank == 0 && !out_host) throw Error(PTB_E_INVALID, "out_host is NULL on rank 0");
    const double wall0 = now_seconds();
    PTB_CUDA(cudaSetDevice(g->device));
    GroupShared* sh = g->sh;

    int n_workers = 1;
    const std::vector<Tile> tiles = make_tiles(req, g->world, &n_workers);
    ensure_frame(g, req.full_w, req.full_h);
    const uint64_t e = g->epoch++;
    const size_t npix = size_t(req.full_w) * req.full_h;
    float4* frame = g->frame_view + (e & 1u) * npix;
    ensure_workers(g, n_workers);

    // post the job
    PTB_CUDA(cudaEventRecord(g->ev_start, g->ctl));
    {
        // Every worker's buffers are sized HERE, before any tile runs (a frame of a new size only): a cudaFree /
        // cudaMalloc waits for the kernels in flight on the device, so a worker that grew its buffers while the others
        // were already rendering stalled for a whole tile per buffer (C4 over 2 GPUs: 10-12 s instead of 6.5 s).
        uint32_t max_w = 0, max_h = 0;
        for (const Tile& t : tiles) {
            max_w = std::max(max_w, t.w);
            max_h = std::max(max_h, t.h);
        }
        for (int i = 0; i < n_workers; i++)
            reserve_tile_workspace(scene, g->workers[i]->st, max_w, max_h, req.spp, req.max_depth);
    }
    {
        std::lock_guard<std::mutex> lk(g->m);
        g->job_scene = scene;
        g->job_req = req;
        g->job_tiles = &tiles;
        g->job_counter = &sh->counter[e & 1u];
        g->job_frame = frame;
        g->job_max_w = g->job_max_h = 0;
        for (const Tile& t : tiles) { // (the padded size is monotone in both dimensions)
            g->job_max_w = std::max(g->job_max_w, t.w);
            g->job_max_h = std::max(g->job_max_h, t.h);
        }
        g->job_depth = g_options.frame_queue_depth >= 2 ? 2 : 1;
        g->job_spin = g_options.frame_spin_wait != 0;
        g->job_status = PTB_OK;
        g->job_error.clear();
        g->job_workers = n_workers;
        g->pending = n_workers;
        g->job_seq++;
    }
    g->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(g->m);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    if (g->job_status != PTB_OK) {
        sh->failed.store(1, std::memory_order_release);
        throw Error(g->job_status, g->job_error);
    }
    RankStats mine{};
    for (int i = 0; i < n_workers; i++) {
        ptb_group::Worker* w = g->workers[i];
        PTB_CUDA(cudaStreamWaitEvent(g->ctl, w->fin, 0));
        mine.rays += w->rays;
        mine.paths += w->paths;
        mine.launches += w->launches;
        mine.tiles += w->tiles;
    }
    PTB_CUDA(cudaEventRecord(g->ev_end, g->ctl));
    PTB_CUDA(cudaEventSynchronize(g->ev_end)); // every store of this rank into the frame has landed
    float ms = 0;
    PTB_CUDA(cudaEventElapsedTime(&ms, g->ev_start, g->ev_end));
    mine.gpu_seconds = ms * 1e-3;
    sh->rank[g->rank] = mine;
    // the counter of the NEXT frame was last used two frames ago, and everybody has left that frame
    if (g->rank == 0) sh->counter[(e + 1) & 1u].store(0, std::memory_order_release);
    group_barrier(g); // the frame is complete on rank 0's GPU

    if (g->rank != 0) return;
    if (req.output == PTB_OUT_RGBA32F) {
        const size_t bytes = npix * sizeof(float4);
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, out_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (pinned) {
            PTB_CUDA(cudaMemcpyAsync(out_host, frame, bytes, cudaMemcpyDeviceToHost, g->ctl));
            PTB_CUDA(cudaStreamSynchronize(g->ctl));
        } else {
            if (g->staging_bytes < bytes) {
                if (g->staging) cudaFreeHost(g->staging);
                g->staging = nullptr;
                g->staging_bytes = 0;
                PTB_CUDA(cudaHostAlloc(&g->staging, bytes, cudaHostAllocDefault));
                g->staging_bytes = bytes;
            }
            PTB_CUDA(cudaMemcpyAsync(g->staging, frame, bytes, cudaMemcpyDeviceToHost, g->ctl));
            PTB_CUDA(cudaStreamSynchronize(g->ctl));
            std::memcpy(out_host, g->staging, bytes);
        }
    } else if (req.output == PTB_OUT_RGBA8) {
        const size_t bytes = npix * 4;
        if (g->rgba8_bytes < bytes) {
            if (g->rgba8_dev) cudaFree(g->rgba8_dev);
            g->rgba8_dev = nullptr;
            g->rgba8_bytes = 0;
            PTB_CUDA(cudaMalloc(reinterpret_cast<void**>(&g->rgba8_dev), bytes));
            g->rgba8_bytes = bytes;
        }
        launch_tonemap_rgba(frame, npix, g->rgba8_dev, g->ctl);
        PTB_CUDA(cudaMemcpyAsync(out_host, g->rgba8_dev, bytes, cudaMemcpyDeviceToHost, g->ctl));
        PTB_CUDA(cudaStreamSynchronize(g->ctl));
    }
    PTB_CUDA(cudaGetLastError());
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->n_tiles = (uint32_t)tiles.size();
        stats->n_ranks = (uint32_t)g->world;
        for (int r = 0; r < g->world; r++) {
            const RankStats& rs = sh->rank[r];
            stats->paths += rs.paths;
            stats->rays += rs.rays;
            stats->kernel_launches += rs.launches;
            stats->tiles_per_rank[r] = rs.tiles;
            stats->gpu_seconds_per_rank[r] = rs.gpu_seconds;
            stats->gpu_seconds = std::max(stats->gpu_seconds, rs.gpu_seconds);
        }
        stats->wall_seconds = now_seconds() - wall0;
    }
}

const float4* group_frame_dev(const ptb_group* g) {
    if (!g || !g->frame_view || g->epoch == 0) return nullptr;
    return g->frame_view + ((g->epoch - 1) & 1u) * size_t(g->frame_w) * g->frame_h;
}

} // namespace ptb

// ---- ptb_ctx: one process, one controller thread per GPU ---------------------------------------------------------------

struct ptb_ctx {
    std::vector<int> devices;
    void* shared = nullptr;
    std::vector<ptb_group*> groups;
    std::vector<ptb_scene*> scenes;
    // controller threads for ranks 1..n-1 (the calling thread drives rank 0)
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    uint64_t job_seq = 0;
    int pending = 0;
    bool quit = false;
    ptb_frame_req req{};
    std::vector<ptb_status> status;
    std::vector<std::string> errors;
};

namespace ptb {

namespace {

void ctx_controller(ptb_ctx* c, int i) {
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(c->m);
            c->cv_job.wait(lk, [&] { return c->quit || c->job_seq != seen; });
            if (c->quit) return;
            seen = c->job_seq;
        }
        ptb_status st = PTB_OK;
        std::string err;
        try {
            group_render_frame(c->groups[i], c->scenes[i], c->req, nullptr, nullptr);
        } catch (const Error& e) {
            st = e.code;
            err = e.what();
        } catch (const std::exception& e) {
            st = PTB_E_INVALID;
            err = e.what();
        }
        {
            std::lock_guard<std::mutex> lk(c->m);
            c->status[i] = st;
            c->errors[i] = err;
            if (--c->pending == 0) c->cv_done.notify_all();
        }
    }
}

void ctx_drop_scenes(ptb_ctx* c) {
    for (ptb_scene*& s : c->scenes) {
        destroy_scene(s);
        s = nullptr;
    }
}

} // namespace

ptb_ctx* ctx_create(int n_gpus, const int* devices) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        throw Error(PTB_E_CUDA, "no CUDA device is usable; libptb has no CPU fallback");
    if (n_gpus < 1 || n_gpus > GROUP_MAX_WORLD) throw Error(PTB_E_INVALID, "n_gpus out of range (1..16)");
    auto* c = new ptb_ctx;
    try {
        for (int i = 0; i < n_gpus; i++) {
            const int d = devices ? devices[i] : i;
            if (d < 0 || d >= n_dev) throw Error(PTB_E_INVALID, "device ordinal out of range (" + std::to_string(n_dev) + " visible)");
            for (int p : c->devices)
                if (p == d) throw Error(PTB_E_INVALID, "a device is listed twice");
            c->devices.push_back(d);
        }
        c->shared = group_shared_alloc();
        c->scenes.assign(n_gpus, nullptr);
        c->status.assign(n_gpus, PTB_OK);
        c->errors.assign(n_gpus, std::string());
        for (int i = 0; i < n_gpus; i++) c->groups.push_back(group_create_threads(c->shared, i, n_gpus, c->devices[i]));
        for (int i = 1; i < n_gpus; i++) c->threads.emplace_back(ctx_controller, c, i);
    } catch (...) {
        ctx_destroy(c);
        throw;
    }
    return c;
}

void ctx_destroy(ptb_ctx* c) {
    if (!c) return;
    {
        std::lock_guard<std::mutex> lk(c->m);
        c->quit = true;
    }
    c->cv_job.notify_all();
    for (auto& t : c->threads)
        if (t.joinable()) t.join();
    for (ptb_group* g : c->groups) group_destroy(g);
    ctx_drop_scenes(c);
    if (c->shared) group_shared_free(c->shared);
    delete c;
}

void ctx_set_scene(ptb_ctx* c, const ptb_scene_desc& desc) {
    if (!c) throw Error(PTB_E_INVALID, "ctx is NULL");
    ctx_drop_scenes(c);
    c->scenes[0] = create_scene(desc, c->devices[0]); // the one host build
    // GPU → GPU replication of the flattened blob, all destinations at once
    std::vector<std::thread> copies;
    std::vector<std::string> errs(c->devices.size());
    for (size_t i = 1; i < c->devices.size(); i++)
        copies.emplace_back([c, i, &errs] {
            try {
                c->scenes[i] = clone_scene(c->scenes[0], c->devices[i]);
            } catch (const std::exception& e) {
                errs[i] = e.what();
            }
        });
    for (auto& t : copies) t.join();
    for (const std::string& e : errs)
        if (!e.empty()) {
            ctx_drop_scenes(c);
            throw Error(PTB_E_CUDA, "scene replication failed: " + e);
        }
}

const ptb_scene* ctx_scene(const ptb_ctx* c, int i) {
    if (!c || i < 0 || i >= (int)c->scenes.size()) return nullptr;
    return c->scenes[i];
}

int ctx_size(const ptb_ctx* c) { return c ? (int)c->devices.size() : 0; }

void ctx_render_frame(ptb_ctx* c, const ptb_frame_req& req, void* out_host, ptb_frame_stats* stats) {
    if (!c) throw Error(PTB_E_INVALID, "ctx is NULL");
    if (!c->scenes[0]) throw Error(PTB_E_INVALID, "the context has no scene (ptb_ctx_set_scene / ptb_ctx_load_gltf)");
    const int n = (int)c->devices.size();
    group_shared_clear_failure(c->shared);
    {
        std::lock_guard<std::mutex> lk(c->m);
        c->req = req;
        c->pending = n - 1;
        for (int i = 0; i < n; i++) {
            c->status[i] = PTB_OK;
            c->errors[i].clear();
        }
        c->job_seq++;
    }
    c->cv_job.notify_all();
    ptb_status st0 = PTB_OK;
    std::string err0;
    try {
        group_render_frame(c->groups[0], c->scenes[0], req, out_host, stats);
    } catch (const Error& e) {
        st0 = e.code;
        err0 = e.what();
    } catch (const std::exception& e) {
        st0 = PTB_E_INVALID;
        err0 = e.what();
    }
    {
        std::unique_lock<std::mutex> lk(c->m);
        c->cv_done.wait(lk, [&] { return c->pending == 0; });
    }
    // report the root cause: a rank's own failure rather than "another rank failed"
    ptb_status st = st0;
    std::string err = err0;
    for (int i = 1; i < n; i++)
        if (c->status[i] != PTB_OK && (st == PTB_OK || err.find("another rank failed") != std::string::npos)) {
            st = c->status[i];
            err = "GPU " + std::to_string(c->devices[i]) + ": " + c->errors[i];
        }
    if (st != PTB_OK) throw Error(st, err);
}

} // namespace ptb
