// experiments/extend_simple.cu — the FIRST closest-hit kernel: one thread per ray, the reference's while-while loop
// (trace_device.cuh: scene_closest), traversal stack in shared memory.  ncu showed 3.9-8.9 of 32 lanes active
// (profiles/r01_v1_extend_ncu_summary.txt); kept for A/B measurements only: built when PTB_BUILD_EXPERIMENTS=1
// (build.py), selected with option extend_variant = 0.
#include "kernels.hpp"
#include "trace_device.cuh"

namespace ptb {

namespace {

constexpr int EXT_THREADS = 128;
constexpr size_t STACK_SMEM_BYTES = size_t(KD_STACK_DEPTH) * 3 * sizeof(uint32_t); // per thread

__device__ __forceinline__ KdStack make_stack(uint32_t* smem) {
    KdStack s;
    s.base = smem + threadIdx.x;
    s.stride = blockDim.x;
    return s;
}

template <bool COUNT>
__global__ void __launch_bounds__(EXT_THREADS)
    extend_kernel(DScene S, const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                  uint4* __restrict__ hits, float* __restrict__ t_out, const uint32_t* __restrict__ n_ptr,
                  uint32_t* __restrict__ head, DeviceCounters* __restrict__ counters) {
    extern __shared__ uint32_t smem[];
    const KdStack stack = make_stack(smem);
    const uint32_t n = *n_ptr;
    const int lane = threadIdx.x & 31;
    TraceCounters cnt{0, 0, 0, 0};
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(head, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base >= n) break;
        const uint32_t k = base + lane;
        if (k < n) {
            const float4 o4 = ray_o[k], d4 = ray_d[k];
            const SceneHit h = scene_closest<COUNT>(S, V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, stack, cnt);
            uint4 rec;
            rec.x = (h.t >= 0) ? ((h.instance << HIT_SURFACE_BITS) | h.surface) : HIT_MISS;
            rec.y = h.tri;
            rec.z = __float_as_uint(h.beta);
            rec.w = __float_as_uint(h.gamma);
            hits[k] = rec;
            if (t_out) t_out[k] = h.t;
            cnt.rays++;
        }
    }
    // one atomic per warp for the ray count (always), visit counters only when asked for
    unsigned long long r = cnt.rays;
    for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xFFFFFFFFu, r, o);
    if (lane == 0 && r) atomicAdd(&counters->rays, r);
    if (COUNT) {
        unsigned long long a = cnt.node_visits, b = cnt.leaf_visits, c = cnt.tri_tests;
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
            b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
            c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        }
        if (lane == 0) {
            atomicAdd(&counters->node_visits, a);
            atomicAdd(&counters->leaf_visits, b);
            atomicAdd(&counters->tri_tests, c);
        }
    }
}

} // namespace

void launch_extend(const DScene& S, const float4* ray_o, const float4* ray_d, uint4* hits, float* t_out,
                   const uint32_t* n_ptr, uint32_t* head, DeviceCounters* counters, const LaunchCfg& cfg,
                   cudaStream_t st) {
    const size_t smem = STACK_SMEM_BYTES * EXT_THREADS;
    int per_sm = 0; // persistent grid: no more blocks than are resident at once
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extend_kernel<false>, EXT_THREADS, smem) != cudaSuccess ||
        per_sm <= 0)
        per_sm = 4;
    const int grid = cfg.sm_count * (per_sm < cfg.extend_blocks_per_sm ? per_sm : cfg.extend_blocks_per_sm);
    if (cfg.count_visits)
        extend_kernel<true><<<grid, EXT_THREADS, smem, st>>>(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters);
    else
        extend_kernel<false><<<grid, EXT_THREADS, smem, st>>>(S, ray_o, ray_d, hits, t_out, n_ptr, head, counters);
}

int extend_regs_per_thread() {
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, extend_kernel<false>) != cudaSuccess) return -1;
    return a.numRegs;
}

} // namespace ptb
