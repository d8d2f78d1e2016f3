// extend_ctx.cu — closest-hit kernel, fourth design: several ray CONTEXTS per lane, state in shared memory.
//
// Same arithmetic, same order of node visits and triangle tests per ray as extend.cu / trace_device.cuh
// (reference: LIB/core/renderer.cpp:645-675, LIB/scene/model.cpp:20-72, LIB/core/mesh.cpp:300-405,
// LIB/geometry/triangle.cpp:120-190).  What changes is, once more, the mapping onto the warp.
//
// In extend.cu a lane owns ONE ray whose state lives in registers, and the warp's loop offers node steps
// to the lanes that descend and triangle tests to the lanes that sit in a leaf: ncu shows 13 of 32 lanes
// active (profiles/r01_v5_extend_ncu_summary.txt) — a lane can only use the section its single ray is in —
// and the kernel is latency bound (throughput still grows ~10 % per extra resident block at 7 blocks/SM).
// Here a lane owns K rays.  Their traversal state lives in shared memory, laid out [field][context][thread]
// (bank = thread, conflict-free), the rarely touched per-ray results in local memory.  Every section of the
// loop picks, per lane, a context that can use it, loads the handful of fields the section needs, works,
// and stores what changed:
//     SET-UP   a waiting context: fold / write the result / take a new ray / instance + mesh boxes
//     DESCEND  a context at a branch (or with a pending pop): STEPS node steps, then leaf entry
//     TEST     a context inside a leaf: two triangles, fetched together
// No per-ray value stays in a register between sections, so registers stop limiting residency, a lane is
// busy whenever ANY of its K rays can use the section, and the two triangle tests of a section are
// independent instruction streams.
#include <algorithm>

#include "extend_common.cuh"
#include "kernels.hpp"

namespace ptb {

namespace {

constexpr int T_THREADS = 128;
constexpr uint32_t T_BATCH = 128; // rays a warp takes from the global head at once

// context states (4 bits each in the lane's state word); CS_FETCH..CS_DONE wait for the set-up section
enum : uint32_t { CS_FETCH = 0, CS_SETUP = 1, CS_DONE = 2, CS_TRAV = 3, CS_POP = 4, CS_LEAF = 5 };

// hot fields, shared memory
enum : int {
    H_OX, H_OY, H_OZ, H_DX, H_DY, H_DZ, H_YX, H_YY, H_YZ, // ray in instance space, refined reciprocals of d
    H_TMIN, H_TMAX, H_NDX, H_NDY, H_SP, H_SLOW,           // segment, current record, stack height, exact-division flag
    H_TRIS,                                               // triangle base of the current mesh (pair / reference indices are absolute)
    H_LPOS, H_LEND, H_LT, H_LB, H_LG, H_LTRI,             // leaf cursor and the best hit in the current leaf
    H_COUNT
};

// cold fields, local memory
enum : int {
    C_K, C_NEXT_INST, C_SURF, C_NSURF, C_FIRST_SURF,
    C_IT, C_IB, C_IG, C_ITRI, C_ISURF, // best over the surfaces of the current instance (local distance)
    C_NT, C_NB, C_NG, C_NTRI, C_NIS,   // nearest over the instances (world distance)
    C_COUNT = 16
};

__device__ __forceinline__ uint32_t st_get(uint32_t w, int k) { return (w >> (4 * k)) & 15u; }
__device__ __forceinline__ uint32_t st_set(uint32_t w, int k, uint32_t s) {
    return (w & ~(15u << (4 * k))) | (s << (4 * k));
}

} // namespace

template <bool COUNT, int K, int STEPS, int MINB>
__global__ void __launch_bounds__(T_THREADS, MINB)
    extend_ctx_kernel(DScene S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                      uint4* __restrict__ hits, float* __restrict__ t_out, const uint32_t* __restrict__ n_ptr,
                      uint32_t* __restrict__ head, DeviceCounters* __restrict__ counters, int setup_lanes) {
    extern __shared__ uint32_t sh[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
#define HOT(f, kb) sh[(f) * (K * T_THREADS) + (kb)]
#define HOTF(f, kb) __uint_as_float(HOT(f, kb))
    uint4 stk[K * KD_STACK_DEPTH]; // pending child records + their segments, per context
    uint32_t cold[K * C_COUNT];

    const uint32_t n = *n_ptr;
    uint32_t pool_next = 0, pool_end = 0; // warp-uniform
    bool drained = (n == 0);              // warp-uniform: the global queue has nothing left
    uint32_t stw = 0;                     // K context states, all CS_FETCH
    unsigned long long c_nodes = 0, c_leaves = 0, c_tris = 0, c_rays = 0;

    for (;;) {
        __syncwarp();
        const bool fetch_possible = !(drained && pool_next == pool_end);
        // ---- which context waits for the set-up section?  results first, then empty contexts
        int kw = -1;
#pragma unroll
        for (int k = K - 1; k >= 0; k--)
            if (st_get(stw, k) == CS_FETCH && fetch_possible) kw = k;
#pragma unroll
        for (int k = K - 1; k >= 0; k--) {
            const uint32_t s = st_get(stw, k);
            if (s == CS_SETUP || s == CS_DONE) kw = k;
        }
        const unsigned m_wait = __ballot_sync(0xFFFFFFFFu, kw >= 0);
        if (!fetch_possible && __all_sync(0xFFFFFFFFu, stw == 0)) break;

        if (__popc(m_wait) >= (fetch_possible ? setup_lanes : 1)) {
            // ================================================================== SET-UP
            uint32_t sw = CS_TRAV; // anything that is not a waiting state
            uint32_t kb = threadIdx.x, cb = 0;
            uint32_t k = 0, next_inst = 0, surf = 0, n_surf = 0, first_surf = 0;
            float it = -1, nt = -1;
            V3 o{0, 0, 0}, d{0, 0, 1}, y{0, 0, 1};
            bool slowdiv = false;
            if (kw >= 0) {
                sw = st_get(stw, kw);
                kb = kw * T_THREADS + threadIdx.x;
                cb = kw * C_COUNT;
                if (sw != CS_FETCH) {
                    k = cold[cb + C_K];
                    next_inst = cold[cb + C_NEXT_INST];
                    surf = cold[cb + C_SURF];
                    n_surf = cold[cb + C_NSURF];
                    first_surf = cold[cb + C_FIRST_SURF];
                    it = __uint_as_float(cold[cb + C_IT]);
                    nt = __uint_as_float(cold[cb + C_NT]);
                    o = V3{HOTF(H_OX, kb), HOTF(H_OY, kb), HOTF(H_OZ, kb)};
                    d = V3{HOTF(H_DX, kb), HOTF(H_DY, kb), HOTF(H_DZ, kb)};
                    y = V3{HOTF(H_YX, kb), HOTF(H_YY, kb), HOTF(H_YZ, kb)};
                    slowdiv = HOT(H_SLOW, kb) != 0;
                    if (sw == CS_DONE) { // the traversal of a mesh is over: on to the next surface
                        surf++;
                        sw = CS_SETUP;
                    }
                }
            }
            // ---- A: the current instance is exhausted: local → world distance, keep the nearest
            // (model.cpp:52-63, renderer.cpp:663-669); after the last instance the ray is finished
            if (sw == CS_SETUP && surf >= n_surf) {
                if (next_inst > 0 && it >= 0) {
                    const DInstance& I = S.instances[next_inst - 1];
                    const V3 hit_vec = d * it;
                    const float tw = length(mul(I.fwd.basis, hit_vec));
                    if (tw >= 0 && (tw < nt || !(nt >= 0))) {
                        nt = tw;
                        cold[cb + C_NB] = cold[cb + C_IB];
                        cold[cb + C_NG] = cold[cb + C_IG];
                        cold[cb + C_NTRI] = cold[cb + C_ITRI];
                        cold[cb + C_NIS] = ((next_inst - 1) << HIT_SURFACE_BITS) | cold[cb + C_ISURF];
                    }
                    it = -1.0f;
                }
                if (next_inst >= S.n_instances) {
                    uint4 rec;
                    rec.x = (nt >= 0) ? cold[cb + C_NIS] : HIT_MISS;
                    rec.y = cold[cb + C_NTRI];
                    rec.z = cold[cb + C_NB];
                    rec.w = cold[cb + C_NG];
                    __stcs(hits + k, rec);
                    if (t_out) __stcs(t_out + k, (nt >= 0) ? nt : -1.0f);
                    c_rays++;
                    sw = CS_FETCH;
                }
            }
            __syncwarp();
            // ---- B: hand the pool's rays to the empty contexts
            const unsigned m_fetch = __ballot_sync(0xFFFFFFFFu, kw >= 0 && sw == CS_FETCH);
            if (m_fetch && fetch_possible) {
                if (pool_next == pool_end) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(head, T_BATCH);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n) {
                        drained = true;
                    } else {
                        pool_next = base;
                        pool_end = min(base + T_BATCH, n);
                        if (pool_end == n) drained = true;
                    }
                }
                const uint32_t avail = pool_end - pool_next;
                const uint32_t rank = __popc(m_fetch & lt_mask);
                if (kw >= 0 && sw == CS_FETCH && rank < avail) {
                    k = pool_next + rank;
                    next_inst = 0;
                    surf = 0;
                    n_surf = 0;
                    it = -1.0f;
                    nt = -1.0f;
                    cold[cb + C_NTRI] = 0;
                    cold[cb + C_NB] = 0;
                    cold[cb + C_NG] = 0;
                    sw = CS_SETUP;
                }
                pool_next += min((uint32_t)__popc(m_fetch), avail);
            }
            __syncwarp();
            // ---- C: model::intersect's entry for the next instance: world → local ray, model box (model.cpp:22-33)
            if (sw == CS_SETUP && surf >= n_surf && next_inst < S.n_instances) {
                const float4 o4 = __ldcs(ray_o + k), d4 = __ldcs(ray_d + k); // streaming: keep L2 for the scene
                const V3 ow{o4.x, o4.y, o4.z}, dw{d4.x, d4.y, d4.z};
                // Skip instances whose conservative world-space sphere a REGULAR ray (all direction components
                // inside the division window: no zero, inf, NaN or denormal) clearly misses: the reference's
                // local-space slab test would reject them too, so no result changes (scene.cu).
                const bool regular = in_div_window(dw.x) && in_div_window(dw.y) && in_div_window(dw.z);
                n_surf = 0;
                surf = 0;
                while (next_inst < S.n_instances) {
                    if (regular && sphere_missed(__ldg(S.inst_sphere + next_inst), ow, dw)) {
                        next_inst++;
                        continue;
                    }
                    const DInstance& I = S.instances[next_inst];
                    next_inst++;
                    o = apply(I.inv, ow);
                    d = normalize(mul(I.inv.basis, dw));
                    y = V3{rcp_refined(d.x), rcp_refined(d.y), rcp_refined(d.z)};
                    slowdiv = !(in_div_window(d.x) && in_div_window(d.y) && in_div_window(d.z));
                    float nr, fr;
                    if (slab_test_inv(I.aabb_min, I.aabb_max, o, inv_dir(d, y, slowdiv), nr, fr)) {
                        first_surf = I.first_surface;
                        n_surf = I.n_surfaces;
                        break;
                    }
                }
                HOT(H_OX, kb) = __float_as_uint(o.x);
                HOT(H_OY, kb) = __float_as_uint(o.y);
                HOT(H_OZ, kb) = __float_as_uint(o.z);
                HOT(H_DX, kb) = __float_as_uint(d.x);
                HOT(H_DY, kb) = __float_as_uint(d.y);
                HOT(H_DZ, kb) = __float_as_uint(d.z);
                HOT(H_YX, kb) = __float_as_uint(y.x);
                HOT(H_YY, kb) = __float_as_uint(y.y);
                HOT(H_YZ, kb) = __float_as_uint(y.z);
                HOT(H_SLOW, kb) = slowdiv ? 1u : 0u;
                // no instance left: phase A of the next visit writes the result
            }
            __syncwarp();
            // ---- D: mesh::intersect's entry: slab test against the mesh box (mesh.cpp:301-303)
            if (sw == CS_SETUP && surf < n_surf) {
                const V3 inv = inv_dir(d, y, slowdiv);
                do {
                    const DMesh& M = S.meshes[S.surfaces[first_surf + surf].mesh];
                    float nr, fr;
                    if (slab_test_inv(M.aabb_min, M.aabb_max, o, inv, nr, fr)) {
                        const uint2 root = __ldg(reinterpret_cast<const uint2*>(S.kd_pairs + M.pair_base));
                        HOT(H_TRIS, kb) = M.tri_base;
                        HOT(H_NDX, kb) = root.x;
                        HOT(H_NDY, kb) = root.y;
                        HOT(H_TMIN, kb) = __float_as_uint(nr);
                        HOT(H_TMAX, kb) = __float_as_uint(fr);
                        HOT(H_SP, kb) = 0;
                        sw = CS_TRAV;
                        break;
                    }
                    surf++;
                } while (surf < n_surf);
                // every surface missed: phase A of the next visit moves on to the next instance
            }
            if (kw >= 0) {
                cold[cb + C_K] = k;
                cold[cb + C_NEXT_INST] = next_inst;
                cold[cb + C_SURF] = surf;
                cold[cb + C_NSURF] = n_surf;
                cold[cb + C_FIRST_SURF] = first_surf;
                cold[cb + C_IT] = __float_as_uint(it);
                cold[cb + C_NT] = __float_as_uint(nt);
                stw = st_set(stw, kw, sw);
            }
            __syncwarp();
        }

        // ====================================================================== DESCEND (mesh.cpp:309-379)
        {
            int kt = -1;
#pragma unroll
            for (int k = K - 1; k >= 0; k--) {
                const uint32_t s = st_get(stw, k);
                if (s == CS_TRAV || s == CS_POP) kt = k;
            }
            if (kt >= 0) {
                const uint32_t kb = kt * T_THREADS + threadIdx.x;
                uint32_t st = st_get(stw, kt);
                const V3 o{HOTF(H_OX, kb), HOTF(H_OY, kb), HOTF(H_OZ, kb)};
                const V3 d{HOTF(H_DX, kb), HOTF(H_DY, kb), HOTF(H_DZ, kb)};
                const V3 y{HOTF(H_YX, kb), HOTF(H_YY, kb), HOTF(H_YZ, kb)};
                float tmin = HOTF(H_TMIN, kb), tmax = HOTF(H_TMAX, kb);
                uint2 nd = make_uint2(HOT(H_NDX, kb), HOT(H_NDY, kb));
                int sp = (int)HOT(H_SP, kb);
                const bool slowdiv = HOT(H_SLOW, kb) != 0;
                uint4* const stack = stk + kt * KD_STACK_DEPTH;
#pragma unroll
                for (int s = 0; s < STEPS; s++) {
                    // next pending subtree, or this mesh is finished without a hit (mesh.cpp:309-311,404)
                    if (st == CS_POP) {
                        if (sp == 0) {
                            st = CS_DONE;
                        } else {
                            sp--;
                            const uint4 e = stack[sp];
                            nd = make_uint2(e.x, e.y);
                            tmin = __uint_as_float(e.z);
                            tmax = __uint_as_float(e.w);
                            st = CS_TRAV;
                        }
                    }
                    if (st == CS_TRAV && (nd.y & 3u) != 3u) {
                        if (COUNT) c_nodes++;
                        // both children in one aligned 16-byte load, in flight during the arithmetic below
                        const uint4 ch = __ldg(S.kd_pairs + (nd.y >> 2));
                        const uint32_t axis = nd.y & 3u;
                        const float split = __uint_as_float(nd.x);
                        float oa, da, ya;
                        select_axis(axis, o, d, y, oa, da, ya);
                        const float num = split - oa;
                        float split_dist = div_with_rcp(num, da, ya);
                        if (slowdiv || !in_div_window(num)) split_dist = num / da; // rare: exact division
                        const bool left_first = oa < split;
                        const uint2 first = left_first ? make_uint2(ch.x, ch.y) : make_uint2(ch.z, ch.w);
                        const uint2 second = left_first ? make_uint2(ch.z, ch.w) : make_uint2(ch.x, ch.y);
                        // same comparisons, same order as mesh.cpp:354-369 (a NaN distance takes the "both" branch)
                        const bool near_only = (split_dist < 0) || (split_dist > tmax);
                        const bool far_only = !near_only && (split_dist < tmin);
                        const bool both = !near_only && !far_only;
                        if (both && second.y != KD_ABSENT) {
                            stack[sp] = make_uint4(second.x, second.y, __float_as_uint(split_dist), __float_as_uint(tmax));
                            sp++;
                        }
                        tmax = both ? split_dist : tmax;
                        nd = far_only ? second : first;
                        if (nd.y == KD_ABSENT) st = CS_POP;
                    }
                }
                // arrival at a leaf (mesh.cpp:376-379)
                if (st == CS_TRAV && (nd.y & 3u) == 3u) {
                    if (COUNT) c_leaves++;
                    const uint32_t cnt = nd.y >> 2;
                    if (cnt) {
                        HOT(H_LPOS, kb) = nd.x;
                        HOT(H_LEND, kb) = nd.x + cnt;
                        HOT(H_LT, kb) = __float_as_uint(-1.0f);
                        st = CS_LEAF;
                    } else {
                        st = CS_POP;
                    }
                }
                HOT(H_TMIN, kb) = __float_as_uint(tmin);
                HOT(H_TMAX, kb) = __float_as_uint(tmax);
                HOT(H_NDX, kb) = nd.x;
                HOT(H_NDY, kb) = nd.y;
                HOT(H_SP, kb) = (uint32_t)sp;
                stw = st_set(stw, kt, st);
            }
        }
        __syncwarp();

        // ====================================================================== TEST (mesh.cpp:381-401)
        {
            int kl = -1;
#pragma unroll
            for (int k = K - 1; k >= 0; k--)
                if (st_get(stw, k) == CS_LEAF) kl = k;
            if (kl >= 0) {
                const uint32_t kb = kl * T_THREADS + threadIdx.x;
                const V3 o{HOTF(H_OX, kb), HOTF(H_OY, kb), HOTF(H_OZ, kb)};
                const V3 d{HOTF(H_DX, kb), HOTF(H_DY, kb), HOTF(H_DZ, kb)};
                const float tmax = HOTF(H_TMAX, kb);
                uint32_t lpos = HOT(H_LPOS, kb);
                const uint32_t lend = HOT(H_LEND, kb);
                float lt = HOTF(H_LT, kb), lb = HOTF(H_LB, kb), lg = HOTF(H_LG, kb);
                uint32_t ltri = HOT(H_LTRI, kb);
                const uint32_t* __restrict__ refs = S.kd_refs;
                const float4* __restrict__ tris = S.tri + size_t(HOT(H_TRIS, kb)) * 3;
                // two triangles per visit, fetched together: two independent load → test chains
                const bool two = lpos + 1 < lend;
                const uint32_t r0 = __ldg(refs + lpos);
                const uint32_t r1 = two ? __ldg(refs + lpos + 1) : r0;
                const float4* t0 = tris + size_t(r0) * 3;
                const float4* t1 = tris + size_t(r1) * 3;
                const float4 a0 = __ldg(t0), ab0 = __ldg(t0 + 1), ac0 = __ldg(t0 + 2);
                const float4 a1 = __ldg(t1), ab1 = __ldg(t1 + 1), ac1 = __ldg(t1 + 2);
                float b0, g0, b1, g1;
                const float d0 = tri_test(V3{a0.x, a0.y, a0.z}, V3{ab0.x, ab0.y, ab0.z}, V3{ac0.x, ac0.y, ac0.z}, o, d, b0, g0);
                const float d1 = tri_test(V3{a1.x, a1.y, a1.z}, V3{ab1.x, ab1.y, ab1.z}, V3{ac1.x, ac1.y, ac1.z}, o, d, b1, g1);
                if (COUNT) c_tris += two ? 2 : 1;
                if (d0 >= 0 && d0 <= tmax && (d0 < lt || !(lt >= 0))) {
                    lt = d0;
                    lb = b0;
                    lg = g0;
                    ltri = r0;
                }
                if (two && d1 >= 0 && d1 <= tmax && (d1 < lt || !(lt >= 0))) {
                    lt = d1;
                    lb = b1;
                    lg = g1;
                    ltri = r1;
                }
                lpos += two ? 2u : 1u;
                if (lpos == lend) {
                    if (lt >= 0) {
                        // "return at the first leaf that yields a hit"; fold into the instance's best (model.cpp:45-49)
                        const uint32_t cb = kl * C_COUNT;
                        const float it = __uint_as_float(cold[cb + C_IT]);
                        if (lt < it || !(it >= 0)) {
                            cold[cb + C_IT] = __float_as_uint(lt);
                            cold[cb + C_IB] = __float_as_uint(lb);
                            cold[cb + C_IG] = __float_as_uint(lg);
                            cold[cb + C_ITRI] = ltri;
                            cold[cb + C_ISURF] = cold[cb + C_SURF];
                        }
                        stw = st_set(stw, kl, CS_DONE);
                    } else {
                        stw = st_set(stw, kl, CS_POP);
                    }
                }
                HOT(H_LPOS, kb) = lpos;
                HOT(H_LT, kb) = __float_as_uint(lt);
                HOT(H_LB, kb) = __float_as_uint(lb);
                HOT(H_LG, kb) = __float_as_uint(lg);
                HOT(H_LTRI, kb) = ltri;
            }
        }
    }
#undef HOT
#undef HOTF

    for (int off = 16; off; off >>= 1) c_rays += __shfl_xor_sync(0xFFFFFFFFu, c_rays, off);
    if (lane == 0 && c_rays) atomicAdd(&counters->rays, c_rays);
    if (COUNT) {
        for (int off = 16; off; off >>= 1) {
            c_nodes += __shfl_xor_sync(0xFFFFFFFFu, c_nodes, off);
            c_leaves += __shfl_xor_sync(0xFFFFFFFFu, c_leaves, off);
            c_tris += __shfl_xor_sync(0xFFFFFFFFu, c_tris, off);
        }
        if (lane == 0) {
            atomicAdd(&counters->node_visits, c_nodes);
            atomicAdd(&counters->leaf_visits, c_leaves);
            atomicAdd(&counters->tri_tests, c_tris);
        }
    }
}

namespace {

using CtxFn = void (*)(DScene, const float4*, const float4*, uint4*, float*, const uint32_t*, uint32_t*, DeviceCounters*, int);

struct CtxKernel {
    CtxFn fn;
    int contexts;
    int min_blocks;
};

template <bool COUNT>
CtxKernel pick_ctx(int contexts) {
    if (contexts <= 2) return {extend_ctx_kernel<COUNT, 2, 4, 8>, 2, 8};
    if (contexts == 3) return {extend_ctx_kernel<COUNT, 3, 4, 6>, 3, 6};
    return {extend_ctx_kernel<COUNT, 4, 4, 4>, 4, 4};
}

} // namespace

void launch_extend_ctx(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                       const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                       cudaStream_t st) {
    const CtxKernel kk = cfg.count_visits ? pick_ctx<true>(cfg.extend_contexts) : pick_ctx<false>(cfg.extend_contexts);
    const size_t smem = size_t(H_COUNT) * kk.contexts * T_THREADS * sizeof(uint32_t);
    cudaFuncSetAttribute(kk.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kk.fn, T_THREADS, smem) != cudaSuccess || per_sm <= 0)
        per_sm = kk.min_blocks;
    const int grid = cfg.sm_count * std::min(per_sm, cfg.extend_blocks_per_sm);
    kk.fn<<<grid, T_THREADS, smem, st>>>(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters,
                                          std::max(1, std::min(32, cfg.extend_setup_lanes)));
}

int extend_ctx_regs_per_thread(int contexts) {
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, pick_ctx<false>(contexts).fn) != cudaSuccess) return -1;
    return a.numRegs;
}

} // namespace ptb
