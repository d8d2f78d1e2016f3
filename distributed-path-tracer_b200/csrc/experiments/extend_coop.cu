// extend_coop.cu — closest-hit kernel, third design: warp-cooperative leaf tests.
//
// Same arithmetic as extend.cu / trace_device.cuh (reference: LIB/core/mesh.cpp:300-405,
// LIB/geometry/triangle.cpp:120-190, LIB/scene/model.cpp:20-72, LIB/core/renderer.cpp:645-675).
// What changes again is the mapping onto the warp.  In extend.cu a lane that reaches a leaf tests the
// leaf's triangles itself, one or two per main-loop iteration, and meanwhile takes no part in the node
// steps; ncu shows ~14 of 32 lanes active in every section (profiles/r01_v4_extend_ncu_summary.txt).
// Here a lane that reaches a leaf only REGISTERS the leaf (first reference, count) and waits; once per
// iteration the warp turns the pending (ray, triangle) pairs into a dense batch: pair j goes to lane j,
// which fetches the owner's ray by shuffle, tests the triangle, and a segmented minimum hands the
// leaf's closest hit back to the owner ("nearest, first in leaf order on ties", mesh.cpp:381-389).
// Lanes therefore spend their time in node steps, and triangle tests run 32 wide.
#include <algorithm>

#include "kernels.hpp"
#include "trace_device.cuh"

namespace ptb {

namespace {

constexpr int C_THREADS = 128;
constexpr int C_MIN_BLOCKS = 6;
constexpr uint32_t C_BATCH = 128;      // rays a warp takes from the global head at once
constexpr int C_SETUP_MIN_LANES = 8;
constexpr int C_STEPS = 4;             // node steps offered per main-loop iteration
constexpr int C_TEST_MIN_PAIRS = 16;   // run a (partial) test batch when at least this many pairs are pending

enum : int { ST_FETCH = 0, ST_SETUP = 1, ST_TRAV = 2, ST_WAIT = 3, ST_POP = 4 };

__device__ __forceinline__ float rcp_refined(float b) {
    float y0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(b));
    const float e = __fmaf_rn(-b, y0, 1.0f);
    return __fmaf_rn(y0, e, y0);
}
__device__ __forceinline__ float div_with_rcp(float a, float b, float y) {
    const float q0 = __fmul_rn(a, y);
    const float r0 = __fmaf_rn(-b, q0, a);
    return __fmaf_rn(r0, y, q0);
}
__device__ __forceinline__ bool in_div_window(float x) {
    const float ax = fabsf(x);
    return ax > 8.673617e-19f /* 2^-60 */ && ax < 1.1529215e18f /* 2^60 */;
}
__device__ __forceinline__ void select_axis(uint32_t axis, const V3& o, const V3& d, const V3& y, float& oa, float& da,
                                            float& ya) {
    asm("{\n\t"
        ".reg .pred p0, p1;\n\t"
        "setp.eq.u32 p0, %3, 0;\n\t"
        "setp.eq.u32 p1, %3, 1;\n\t"
        "selp.f32 %0, %5, %6, p1;\n\t"
        "selp.f32 %0, %4, %0, p0;\n\t"
        "selp.f32 %1, %8, %9, p1;\n\t"
        "selp.f32 %1, %7, %1, p0;\n\t"
        "selp.f32 %2, %11, %12, p1;\n\t"
        "selp.f32 %2, %10, %2, p0;\n\t"
        "}"
        : "=&f"(oa), "=&f"(da), "=&f"(ya)
        : "r"(axis), "f"(o.x), "f"(o.y), "f"(o.z), "f"(d.x), "f"(d.y), "f"(d.z), "f"(y.x), "f"(y.y), "f"(y.z));
}

} // namespace

template <bool COUNT>
__global__ void __launch_bounds__(C_THREADS, C_MIN_BLOCKS)
    extend_coop_kernel(DScene S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                       uint4* __restrict__ hits, float* __restrict__ t_out, const uint32_t* __restrict__ n_ptr,
                       uint32_t* __restrict__ head, DeviceCounters* __restrict__ counters) {
    __shared__ int s_owner[C_THREADS / 32][32]; // pair slot → owner lane (scatter + max-scan)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t stk_node[KD_STACK_DEPTH];
    float stk_tmin[KD_STACK_DEPTH], stk_tmax[KD_STACK_DEPTH];

    const uint32_t n = *n_ptr;
    uint32_t pool_next = 0, pool_end = 0;
    bool drained = (n == 0);

    int state = ST_FETCH;
    uint32_t k = 0;
    V3 o{0, 0, 0}, d{0, 0, 1}, y{0, 0, 1};
    bool slowdiv = false;
    uint32_t next_inst = 0;
    uint32_t surf = 0, n_surf = 0, first_surf = 0;
    uint32_t node_base = 0, ref_base = 0, tri_base = 0; // of the current mesh
    uint32_t node = 0;
    uint2 nd = make_uint2(0, 3);
    float tmin = 0, tmax = 0;
    int sp = 0;
    uint32_t ref_pos = 0, rem = 0; // pending leaf: next reference (absolute index into kd_refs), references left
    float lt = -1, lb = 0, lg = 0; // best in the pending leaf so far
    uint32_t ltri = 0;
    float it = -1, ib = 0, ig = 0;
    uint32_t itri = 0, isurf = 0;
    float nt = -1, nb = 0, ng = 0;
    uint32_t ntri = 0, nis = 0;
    unsigned long long c_nodes = 0, c_leaves = 0, c_tris = 0, c_rays = 0;

    for (;;) {
        __syncwarp();
        const unsigned m_wait = __ballot_sync(0xFFFFFFFFu, state <= ST_SETUP);
        const int n_wait = __popc(m_wait);
        if (n_wait >= C_SETUP_MIN_LANES) {
            const unsigned m_fetch = __ballot_sync(0xFFFFFFFFu, state == ST_FETCH);
            const unsigned m_setup = m_wait & ~m_fetch;
            const bool fetch_possible = !(drained && pool_next == pool_end);
            if (m_wait == 0xFFFFFFFFu && m_setup == 0 && !fetch_possible) break;
            if (m_fetch && fetch_possible) {
                if (pool_next == pool_end) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(head, C_BATCH);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= n) {
                        drained = true;
                    } else {
                        pool_next = base;
                        pool_end = min(base + C_BATCH, n);
                        if (pool_end == n) drained = true;
                    }
                }
                const uint32_t avail = pool_end - pool_next;
                const uint32_t rank = __popc(m_fetch & lt_mask);
                if (state == ST_FETCH && rank < avail) {
                    k = pool_next + rank;
                    next_inst = 0;
                    surf = 0;
                    n_surf = 0;
                    it = -1.0f;
                    nt = -1.0f;
                    state = ST_SETUP;
                }
                pool_next += min((uint32_t)__popc(m_fetch), avail);
            }
            if (state == ST_SETUP) {
                for (;;) {
                    if (surf < n_surf) {
                        const DMesh& M = S.meshes[S.surfaces[first_surf + surf].mesh];
                        float nr, fr;
                        if (slab_test(M.aabb_min, M.aabb_max, o, d, nr, fr)) {
                            node_base = M.node_base;
                            ref_base = M.ref_base;
                            tri_base = M.tri_base;
                            node = 0;
                            nd = __ldg(S.kd_nodes + node_base);
                            tmin = nr;
                            tmax = fr;
                            sp = 0;
                            state = ST_TRAV;
                            break;
                        }
                        surf++;
                        continue;
                    }
                    if (next_inst > 0 && it >= 0) {
                        const DInstance& I = S.instances[next_inst - 1];
                        const V3 hit_vec = d * it;
                        const float tw = length(mul(I.fwd.basis, hit_vec));
                        if (tw >= 0 && (tw < nt || !(nt >= 0))) {
                            nt = tw;
                            nb = ib;
                            ng = ig;
                            ntri = itri;
                            nis = ((next_inst - 1) << HIT_SURFACE_BITS) | isurf;
                        }
                        it = -1.0f;
                    }
                    if (next_inst >= S.n_instances) {
                        uint4 rec;
                        rec.x = (nt >= 0) ? nis : HIT_MISS;
                        rec.y = ntri;
                        rec.z = __float_as_uint(nb);
                        rec.w = __float_as_uint(ng);
                        __stcs(hits + k, rec);
                        if (t_out) __stcs(t_out + k, (nt >= 0) ? nt : -1.0f);
                        c_rays++;
                        state = ST_FETCH;
                        break;
                    }
                    const float4 o4 = __ldcs(ray_o + k), d4 = __ldcs(ray_d + k);
                    const V3 ow{o4.x, o4.y, o4.z}, dw{d4.x, d4.y, d4.z};
                    if (in_div_window(dw.x) && in_div_window(dw.y) && in_div_window(dw.z)) {
                        while (next_inst < S.n_instances) { // conservative instance culling, see extend.cu / scene.cu
                            const float4 sp4 = __ldg(S.inst_sphere + next_inst);
                            const V3 oc = V3{sp4.x, sp4.y, sp4.z} - ow;
                            const float tproj = dot(oc, dw), oc2 = dot(oc, oc), r2 = sp4.w * sp4.w;
                            const bool miss = sp4.w < 0 || (oc2 - tproj * tproj > r2) || (tproj < 0 && oc2 > r2);
                            if (!miss) break;
                            next_inst++;
                        }
                        if (next_inst >= S.n_instances) continue;
                    }
                    const DInstance& I = S.instances[next_inst];
                    next_inst++;
                    o = apply(I.inv, ow);
                    d = normalize(mul(I.inv.basis, dw));
                    float nr, fr;
                    n_surf = 0;
                    surf = 0;
                    it = -1.0f;
                    if (!slab_test(I.aabb_min, I.aabb_max, o, d, nr, fr)) continue;
                    first_surf = I.first_surface;
                    n_surf = I.n_surfaces;
                    y = V3{rcp_refined(d.x), rcp_refined(d.y), rcp_refined(d.z)};
                    slowdiv = !(in_div_window(d.x) && in_div_window(d.y) && in_div_window(d.z));
                }
            }
        }

        // ---- node steps (mesh.cpp:333-369)
#pragma unroll
        for (int s = 0; s < C_STEPS; s++) {
            if (state == ST_TRAV && (nd.y & 3u) != 3u) {
                if (COUNT) c_nodes++;
                const uint32_t axis = nd.y & 3u;
                const float split = __uint_as_float(nd.x);
                float oa, da, ya;
                select_axis(axis, o, d, y, oa, da, ya);
                const float num = split - oa;
                float split_dist = div_with_rcp(num, da, ya);
                if (slowdiv || !in_div_window(num)) split_dist = num / da;
                const uint32_t has_l = (nd.y >> 2) & 1u, has_r = (nd.y >> 3) & 1u;
                const uint32_t li = nd.y >> 4, ri = li + has_l;
                const uint32_t lnode = has_l ? li : NO_NODE, rnode = has_r ? ri : NO_NODE;
                const bool left_first = oa < split;
                const uint32_t first = left_first ? lnode : rnode;
                const uint32_t second = left_first ? rnode : lnode;
                const bool near_only = (split_dist < 0) || (split_dist > tmax);
                const bool far_only = !near_only && (split_dist < tmin);
                const bool both = !near_only && !far_only;
                if (both && second != NO_NODE) {
                    stk_node[sp] = second;
                    stk_tmin[sp] = split_dist;
                    stk_tmax[sp] = tmax;
                    sp++;
                }
                tmax = both ? split_dist : tmax;
                node = far_only ? second : first;
                if (node == NO_NODE)
                    state = ST_POP;
                else
                    nd = __ldg(S.kd_nodes + node_base + node);
            }
        }
        __syncwarp();

        // ---- arrival at a leaf: register it (mesh.cpp:376-379)
        if (state == ST_TRAV && (nd.y & 3u) == 3u) {
            if (COUNT) c_leaves++;
            rem = nd.y >> 2;
            ref_pos = ref_base + nd.x;
            lt = -1.0f;
            state = rem ? ST_WAIT : ST_POP;
        }

        // ---- cooperative triangle tests (mesh.cpp:381-401): dense batches of (owner lane, reference) pairs
        for (;;) {
            __syncwarp();
            const uint32_t my_rem = (state == ST_WAIT) ? rem : 0u;
            // exclusive prefix sum of the pending counts over the lanes
            uint32_t incl = my_rem;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                if (lane >= (uint32_t)off) incl += v;
            }
            const uint32_t pre = incl - my_rem;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total == 0) break;
            if (total < (uint32_t)C_TEST_MIN_PAIRS) {
                // a thin batch is only worth running when hardly any lane could use a node step instead
                const int n_trav = __popc(__ballot_sync(0xFFFFFFFFu, state == ST_TRAV));
                if (n_trav >= 8) break;
            }
            // pair slot j (< 32) belongs to the owner whose [pre, pre + rem) contains j
            s_owner[warp][lane] = -1;
            __syncwarp();
            if (my_rem && pre < 32u) s_owner[warp][pre] = (int)lane;
            __syncwarp();
            int owner = s_owner[warp][lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, owner, off);
                if (lane >= (uint32_t)off) owner = max(owner, v);
            }
            const bool active = lane < min(total, 32u);
            const int src = active ? owner : (int)lane;
            // the owner's ray, leaf interval and mesh, by shuffle
            const float oox = __shfl_sync(0xFFFFFFFFu, o.x, src), ooy = __shfl_sync(0xFFFFFFFFu, o.y, src),
                        ooz = __shfl_sync(0xFFFFFFFFu, o.z, src);
            const float odx = __shfl_sync(0xFFFFFFFFu, d.x, src), ody = __shfl_sync(0xFFFFFFFFu, d.y, src),
                        odz = __shfl_sync(0xFFFFFFFFu, d.z, src);
            const float otmax = __shfl_sync(0xFFFFFFFFu, tmax, src);
            const uint32_t opre = __shfl_sync(0xFFFFFFFFu, pre, src);
            const uint32_t oref = __shfl_sync(0xFFFFFFFFu, ref_pos, src);
            const uint32_t otri_base = __shfl_sync(0xFFFFFFFFu, tri_base, src);
            float dist = -1.0f, beta = 0.0f, gamma = 0.0f;
            uint32_t tri = 0;
            if (active) {
                tri = __ldg(S.kd_refs + oref + (lane - opre));
                const float4* t3 = S.tri + size_t(otri_base + tri) * 3;
                const float4 a = __ldg(t3), ab = __ldg(t3 + 1), ac = __ldg(t3 + 2);
                if (COUNT) c_tris++;
                dist = tri_test(V3{a.x, a.y, a.z}, V3{ab.x, ab.y, ab.z}, V3{ac.x, ac.y, ac.z}, V3{oox, ooy, ooz},
                                V3{odx, ody, odz}, beta, gamma);
                if (!(dist >= 0 && dist <= otmax)) dist = -1.0f; // "has_hit && distance <= max_dist"
            }
            // segmented minimum per owner (segments are contiguous runs of slots): nearest distance; the ballot
            // below then picks the lowest slot (= leaf order) among equal distances
            const uint32_t key = (dist >= 0) ? __float_as_uint(dist) : 0xFFFFFFFFu; // non-negative floats order as uints
            const int seg_id = active ? owner : -1 - (int)lane;
            uint32_t red = key;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t other = __shfl_down_sync(0xFFFFFFFFu, red, off);
                const int other_seg = __shfl_down_sync(0xFFFFFFFFu, seg_id, off);
                if (lane + off < 32u && other_seg == seg_id) red = min(red, other);
            }
            // the first slot of a segment (slot opre of its owner) now holds the segment minimum
            const uint32_t seg_min = __shfl_sync(0xFFFFFFFFu, red, active ? (int)opre : (int)lane);
            const unsigned winners = __ballot_sync(0xFFFFFFFFu, active && key != 0xFFFFFFFFu && key == seg_min);
            // back on the owner side: my slots are [pre, pre + take)
            const uint32_t take = (my_rem && pre < 32u) ? min(my_rem, 32u - pre) : 0u;
            const unsigned my_slots = take ? (((take >= 32u) ? 0xFFFFFFFFu : ((1u << take) - 1u)) << pre) : 0u;
            const unsigned my_win = winners & my_slots;
            const int wl = my_win ? (__ffs(my_win) - 1) : (int)lane;
            const float wd = __shfl_sync(0xFFFFFFFFu, dist, wl), wb = __shfl_sync(0xFFFFFFFFu, beta, wl),
                        wg = __shfl_sync(0xFFFFFFFFu, gamma, wl);
            const uint32_t wt = __shfl_sync(0xFFFFFFFFu, tri, wl);
            if (take) {
                if (my_win && (wd < lt || !(lt >= 0))) { // strict '<': an earlier batch of the same leaf wins ties
                    lt = wd;
                    lb = wb;
                    lg = wg;
                    ltri = wt;
                }
                rem -= take;
                ref_pos += take;
                if (rem == 0) {
                    if (lt >= 0) {
                        // "return at the first leaf that yields a hit"; fold into the instance's best (model.cpp:45-49)
                        if (lt < it || !(it >= 0)) {
                            it = lt;
                            ib = lb;
                            ig = lg;
                            itri = ltri;
                            isurf = surf;
                        }
                        surf++;
                        state = ST_SETUP;
                    } else {
                        state = ST_POP;
                    }
                }
            }
        }
        __syncwarp();

        // ---- POP (mesh.cpp:309-311,404)
        if (state == ST_POP) {
            if (sp == 0) {
                surf++;
                state = ST_SETUP;
            } else {
                sp--;
                node = stk_node[sp];
                tmin = stk_tmin[sp];
                tmax = stk_tmax[sp];
                nd = __ldg(S.kd_nodes + node_base + node);
                state = ST_TRAV;
            }
        }
    }

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

void launch_extend_coop(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                        const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                        cudaStream_t st) {
    int per_sm = 0;
    cudaError_t e = cfg.count_visits
                        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extend_coop_kernel<true>, C_THREADS, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extend_coop_kernel<false>, C_THREADS, 0);
    if (e != cudaSuccess || per_sm <= 0) per_sm = C_MIN_BLOCKS;
    const int grid = cfg.sm_count * std::min(per_sm, cfg.extend_blocks_per_sm);
    if (cfg.count_visits)
        extend_coop_kernel<true><<<grid, C_THREADS, 0, st>>>(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters);
    else
        extend_coop_kernel<false><<<grid, C_THREADS, 0, st>>>(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters);
}

int extend_coop_regs_per_thread() {
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, extend_coop_kernel<false>) != cudaSuccess) return -1;
    return a.numRegs;
}

} // namespace ptb
