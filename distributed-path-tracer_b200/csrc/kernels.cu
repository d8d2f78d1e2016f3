// kernels.cu — the wavefront kernels (sm_100a) and their launch wrappers.
//
//   raygen      one thread per (pixel, sample) of the wave   LIB/core/renderer.cpp:359-370, APP worker.cpp:117-146
//   extend      persistent warps pulling 32-ray batches from an atomic head;
//               closest hit per ray (trace_device.cuh)        LIB/core/renderer.cpp:645-675
//   shadow_gen  (scenes with a sun) one thread per live path: the cone-jittered
//               shadow ray of this shade event, compacted into the shadow queue;
//               the any-hit instantiation of the extend kernel (extend.cu)
//               resolves the queue before shade runs            LIB/core/renderer.cpp:498-511,
//                                                               APP intersection_worker.cpp:22-40,49-67
//   shade       one thread per live path: attributes, material, emission,
//               direct light from the resolved shadow ray, BSDF sampling,
//               throughput, Russian roulette; survivors are compacted into the
//               next queue with one atomicAdd per warp        LIB/core/renderer.cpp:437-643, APP worker.cpp:285-514
//   accumulate  one thread per pixel: the reference's sequential running
//               mean over the wave's samples                  LIB/core/renderer.cpp:373-399
//   tonemap     ACES approximation + sRGB encode + RGBA8      LIB/core/utils.hpp:29-36, LIB/image/image.cpp:143-154
// (LIB = path-tracer-core/path_tracer_lib/path_tracer, APP = path-tracer-core/src/processors/worker)
//
// Path state travels densely with the queue (64 B per live path, ping-pong):
//   ray_o  = (origin.xyz, path id within the wave)
//   ray_d  = (direction.xyz, packed: bounces left | shade events | flags)
//   thr    = (throughput.rgb, -)        rad = (radiance.rgb, alpha)
// so extend and shade read and write fully coalesced float4 streams and no
// kernel gathers by index.  Queue sizes live in device memory, one counter per
// iteration, so a whole wave is enqueued without host round trips.
//
// Built with --fmad=false: the closest-hit arithmetic must round like the
// reference's x86-64 SSE2 code.

#include "kernels.hpp"
#include "shade_device.cuh"
#include "trace_device.cuh"

#include "ptb.h"

namespace ptb {

namespace {

constexpr int SHADE_THREADS = 128;
#ifndef PTB_SHADE_MIN_BLOCKS
#define PTB_SHADE_MIN_BLOCKS 8
#endif
constexpr int SHADE_MIN_BLOCKS = PTB_SHADE_MIN_BLOCKS;

// flags packed in ray_d.w
constexpr uint32_t F_BOUNCE_MASK = 0xFFu;  // bounces remaining
constexpr uint32_t F_EVENT_SHIFT = 8;      // shade events so far (RNG counter), 16 bits
constexpr uint32_t F_EVENT_MASK = 0xFFFFu;
constexpr uint32_t F_PRIMARY = 1u << 24;   // no opaque surface interaction yet (alpha bookkeeping of renderer::trace)

// Slots walk the tile in 8x4 pixel blocks (a warp's primary rays form a compact bundle), the blocks in
// super-blocks of 8x8 (64x32 pixels), so that a stretch of the queue is a compact patch of the image and the
// rays an SM works on at one time share nodes and triangles in its L1.
__device__ __forceinline__ bool slot_to_pixel(const WaveGeom& g, uint32_t q, uint32_t& x, uint32_t& y) {
    const uint32_t blk = q >> 5, in = q & 31u;
    const uint32_t sb = blk >> 6, inb = blk & 63u;
    const uint32_t sby = sb / g.sblocks_x, sbx = sb - sby * g.sblocks_x;
    const uint32_t bx = sbx * 8 + (inb & 7u), by = sby * 8 + (inb >> 3);
    x = bx * 8 + (in & 7u);
    y = by * 4 + (in >> 3);
    return x < g.w && y < g.h;
}

__device__ __forceinline__ uint32_t pixel_to_slot(const WaveGeom& g, uint32_t x, uint32_t y) {
    const uint32_t bx = x >> 3, by = y >> 2;
    const uint32_t sb = (by >> 3) * g.sblocks_x + (bx >> 3), inb = ((by & 7u) << 3) | (bx & 7u);
    return (((sb << 6) | inb) << 5) + ((y & 3u) << 3) + (x & 7u);
}

// Path p of the wave ↔ (sample s of the wave, slot q): block-major, so the wave's samples of one 8x4 block
// are adjacent in the queue (p = ((q / 32) * wave_samples + s) * 32 + q % 32).
__device__ __forceinline__ void path_to_sample_slot(const WaveGeom& g, uint32_t p, uint32_t& s, uint32_t& q) {
    if (!g.block_major) { // sample planes: p = s * padded_pixels + q
        s = p / g.padded_pixels;
        q = p - s * g.padded_pixels;
        return;
    }
    const uint32_t bs = p >> 5, blk = bs / g.wave_samples;
    s = bs - blk * g.wave_samples;
    q = (blk << 5) | (p & 31u);
}

__device__ __forceinline__ size_t sample_slot_to_path(const WaveGeom& g, uint32_t s, uint32_t q) {
    if (!g.block_major) return size_t(s) * g.padded_pixels + q;
    return ((size_t(q >> 5) * g.wave_samples + s) << 5) + (q & 31u);
}

// Queue positions for the threads of a block that emit an item: ONE atomicAdd on the queue's counter per block and
// round instead of one per warp.  8.3 M paths are 260 k atomics on one address otherwise, and raygen — which does
// little else — waited for them: 0.46 ms per wave where its stores take 0.12.  Every thread of the block must call it
// (three barriers); → this thread's position if `emit`.
template <int THREADS>
__device__ __forceinline__ uint32_t block_queue_position(bool emit, uint32_t* __restrict__ counter) {
    __shared__ uint32_t warp_count[THREADS / 32];
    __shared__ uint32_t block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, emit);
    __syncthreads(); // (the previous round's readers are done with warp_count / block_base)
    if (lane == 0) warp_count[warp] = (uint32_t)__popc(mask);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) {
            const uint32_t c = warp_count[w];
            warp_count[w] = total; // exclusive prefix
            total += c;
        }
        block_base = total ? atomicAdd(counter, total) : 0u;
    }
    __syncthreads();
    return block_base + warp_count[warp] + __popc(mask & ((1u << lane) - 1u));
}

// ------------------------------------------------------------- raygen ------

__global__ void __launch_bounds__(256)
    raygen_kernel(DScene S, WaveGeom g, RenderParams rp, PathBuffers out, float4* __restrict__ sample_out,
                  uint32_t* __restrict__ qcount) {
    const uint32_t n = g.padded_pixels * g.wave_samples; // multiple of 32
    for (uint32_t p0 = blockIdx.x * blockDim.x; p0 < n; p0 += gridDim.x * blockDim.x) { // block-uniform trip count
        const uint32_t p = p0 + threadIdx.x;
        uint32_t s = 0, q = 0, x = 0, y = 0;
        bool valid = false;
        if (p < n) {
            path_to_sample_slot(g, p, s, q);
            valid = slot_to_pixel(g, q, x, y);
        }
        const uint32_t k = block_queue_position<256>(valid, qcount);
        if (!valid) continue;
        const uint32_t gx = g.x0 + g.comb_x(x), gy = g.y0 + g.comb_y(y);
        const uint32_t sample = g.first_sample + s;
        float aax = 0.0f, aay = 0.0f;
        if (!(sample == 0 && rp.first_sample_unjittered)) {
            const Philox ph{rp.seed_lo, rp.seed_hi};
            const uint4 r = ph(gy * g.full_w + gx, sample, 0xFFFFFFFFu, 0u);
            aax = u01(r.x);
            aay = u01(r.y);
        }
        const Ray ray = camera_ray(S.camera, gx, gy, aax, aay, g.full_w, g.full_h);
        __stcs(out.ray_o + k, make_float4(ray.o.x, ray.o.y, ray.o.z, __uint_as_float(p)));
        __stcs(out.ray_d + k,
               make_float4(ray.d.x, ray.d.y, ray.d.z, __uint_as_float((rp.max_depth & F_BOUNCE_MASK) | F_PRIMARY)));
        __stcs(out.thr + k, make_float4(1.0f, 1.0f, 1.0f, 0.0f));
        // worker::trace_iter starts alpha at the background value (worker.cpp:299); renderer::trace has no such state
        const float alpha0 = (rp.integrator == 1 && S.transparent_background) ? 0.0f : 1.0f;
        __stcs(out.rad + k, make_float4(0.0f, 0.0f, 0.0f, alpha0));
        // bounce_count == 0: trace returns fvec4::future, trace_iter returns (0, alpha0)
        __stcs(sample_out + p, make_float4(0.0f, 0.0f, 0.0f, alpha0));
    }
}

// -------------------------------------------------------------- shade ------

struct PathState {
    V3 o, d;
    V3 thr;
    V3 rad;
    float alpha;
    uint32_t p;     // path id within the wave
    uint32_t flags; // see F_*
};

// math::is_approx (LIB/math/math.inl:49-52)
__device__ __forceinline__ bool is_approx(float a, float b) { return a == b || fabsf(a - b) < kEpsilon; }

// The shadow ray of a shade event and its answer.  The reference traces it in the middle of the shade function
// (renderer.cpp:505-511; the staged worker: intersection_worker.cpp:22-40 builds it, :49-67 resolves it).  Here
// the shade code runs twice over a path when the scene has a sun: once to GENERATE the ray (everything up to the
// point where the reference calls intersect(), same instructions, same random numbers, then stop), and — after
// the any-hit kernel has resolved the whole queue — once more to USE the answer.
struct ShadowIO {
    V3 o, d;      // GENERATE: the shadow ray (origin, normalised direction)
    bool emit;    // GENERATE: this event casts a shadow ray
    bool visible; // USE: nothing between the hit point and the sun
};
enum : int { SHADOW_NONE = 0, SHADOW_GENERATE = 1, SHADOW_USE = 2 };

// Shades one path.  Returns true when the path continues (state updated in
// place), false when it ended (result holds what the pixel sample receives).
template <bool APP_RR, bool HAS_SUN, int SHADOW>
__device__ __forceinline__ bool shade_path(const DScene& S, const WaveGeom& g, const RenderParams& rp, PathState& st,
                                           const uint4 hit, ShadowIO& shadow, float4& result) {
    const uint32_t bounce = st.flags & F_BOUNCE_MASK;
    const uint32_t event = (st.flags >> F_EVENT_SHIFT) & F_EVENT_MASK;
    const bool primary = (st.flags & F_PRIMARY) != 0;
    const float bg_alpha = S.transparent_background ? 0.0f : 1.0f;

    if (hit.x == HIT_MISS) { // renderer.cpp:443-451, worker.cpp:307-317
        V3 env = S.environment;
        if (S.environment_tex != 0xFFFFFFFFu) { // equirectangular_proj, LIB/core/utils.hpp:22-27
            const float eu = atan2f(st.d.z, st.d.x) * 0.1591F + 0.5F, ev = asinf(st.d.y) * 0.3183F + 0.5F;
            const float4 c = tex_sample(S, S.environment_tex, eu, ev);
            env = V3{c.x, c.y, c.z} * S.environment;
        }
        st.rad = st.rad + st.thr * env;
        const float a = APP_RR ? bg_alpha : (primary ? bg_alpha : 1.0f);
        result = make_float4(st.rad.x, st.rad.y, st.rad.z, a);
        return false;
    }

    const uint32_t instance = hit.x >> HIT_SURFACE_BITS, surface = hit.x & ((1u << HIT_SURFACE_BITS) - 1u);
    const HitAttrs at = hit_attributes(S, instance, surface, hit.y, __uint_as_float(hit.z), __uint_as_float(hit.w));
    const MatSample m = material_sample(S, at.material, at.u, at.v);

    // random numbers of this shade event: a = (opacity, lobe, u1, u2), b = (sun phi, sun theta, roulette, -)
    uint32_t s, q;
    path_to_sample_slot(g, st.p, s, q);
    uint32_t x, y;
    slot_to_pixel(g, q, x, y);
    const uint32_t pixel_id = (g.y0 + g.comb_y(y)) * g.full_w + (g.x0 + g.comb_x(x));
    const uint32_t sample = g.first_sample + s;
    const Philox ph{rp.seed_lo, rp.seed_hi};
    const uint4 ra = ph(pixel_id, sample, event, 0u);
    const uint32_t next_event = (event + 1u) & F_EVENT_MASK;

    if (APP_RR) {
        st.alpha = 1.0f;                         // worker.cpp:320
        st.rad = st.rad + st.thr * m.emissive;   // worker.cpp:331 — before the opacity test
    }

    // stochastic opacity: continue along the same direction, same bounce (renderer.cpp:466-472, worker.cpp:334-341)
    if (!is_approx(m.opacity, 1.0f) && u01(ra.x) > m.opacity) {
        st.o = at.position + st.d * kEpsilon;
        st.d = normalize(st.d); // geometry::ray's constructor normalises again
        st.flags = bounce | (next_event << F_EVENT_SHIFT) | (st.flags & F_PRIMARY);
        return true;
    }

    const V3 normal = shading_normal(at, m.normal_map);
    const V3 outcoming = -st.d;
    if (!APP_RR) st.alpha = 1.0f;

    if (dot(normal, outcoming) <= 0) { // renderer.cpp:478-479 returns black; worker.cpp:348-350 keeps what it has
        result = make_float4(st.rad.x, st.rad.y, st.rad.z, 1.0f);
        return false;
    }

    uint4 rb = make_uint4(0, 0, 0, 0);
    if (HAS_SUN || APP_RR) rb = ph(pixel_id, sample, event, 1u);

    // sun direction for this event, shared by the shadow-catcher test and direct lighting
    V3 direct_incoming{0.0f, 0.0f, 0.0f};
    bool sun_visible = false, sun_above = false;
    if (HAS_SUN) {
        direct_incoming = rand_cone_vec(u01(rb.x), cosf(u01(rb.y) * S.sun.angular_radius), S.sun.direction);
        sun_above = dot(normal, direct_incoming) > 0;
        if (sun_above) {
            if (SHADOW == SHADOW_GENERATE) {
                shadow.o = at.position + direct_incoming * kEpsilon;
                shadow.d = normalize(direct_incoming); // geometry::ray's constructor
                shadow.emit = true;
                return false;
            }
            sun_visible = shadow.visible; // !intersect(shadow ray).hit, resolved by the any-hit kernel
        }
    }
    if (SHADOW == SHADOW_GENERATE) return false;

    if (APP_RR && m.shadow_catcher && bounce == rp.max_depth) { // worker.cpp:353-389
        if (!(HAS_SUN && sun_visible)) {
            result = make_float4(0.0f, 0.0f, 0.0f, 1.0f); // fvec4::future
            return false;
        }
        st.o = at.position + st.d * kEpsilon;
        st.d = normalize(st.d);
        st.flags = bounce | (next_event << F_EVENT_SHIFT) | (st.flags & F_PRIMARY);
        return true;
    }

    const float roughness = rmax(m.roughness, 0.05F);
    float specular_probability = fresnel_schlick(outcoming, reflect(-outcoming, normal), m.ior);
    specular_probability = rmax(specular_probability, m.metallic);
    const bool specular_sample = u01(ra.y) < specular_probability;

    V3 direct_out{0.0f, 0.0f, 0.0f};
    if (HAS_SUN && sun_above) {
        if (sun_visible) {
            if (!APP_RR && m.shadow_catcher && bounce == rp.max_depth) { // renderer.cpp:513-519
                st.o = at.position + st.d * kEpsilon;
                st.d = normalize(st.d);
                st.flags = bounce | (next_event << F_EVENT_SHIFT) | F_PRIMARY;
                return true;
            }
            V3 brdf;
            float pdf;
            eval_brdf(normal, outcoming, direct_incoming, m.albedo, roughness, m.metallic, specular_probability, brdf,
                      pdf);
            pdf = rlerp(1.0f, 1.0f, specular_probability); // "100% chance of hitting the sun"
            direct_out = brdf * S.sun.energy / rmax(pdf, kEpsilon);
            direct_out = clamp3(direct_out, V3{0.0f, 0.0f, 0.0f}, S.sun.energy);
        } else if (!APP_RR && m.shadow_catcher && bounce == rp.max_depth) { // renderer.cpp:560-561
            result = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
            return false;
        }
    }
    if (APP_RR)
        st.rad = st.rad + st.thr * direct_out; // worker.cpp:448
    else
        st.rad = st.rad + st.thr * (direct_out + m.emissive); // renderer.cpp:642, unrolled

    const V3 incoming = specular_sample ? importance_ggx(u01(ra.z), u01(ra.w), normal, outcoming, roughness)
                                        : importance_lambert(u01(ra.z), u01(ra.w), normal);
    if (!(dot(normal, incoming) > 0)) { // renderer.cpp:578, worker.cpp:504-507
        result = make_float4(st.rad.x, st.rad.y, st.rad.z, st.alpha);
        return false;
    }
    V3 brdf;
    float pdf;
    eval_brdf(normal, outcoming, incoming, m.albedo, roughness, m.metallic, specular_probability, brdf, pdf);
    const V3 k = brdf / rmax(pdf, kEpsilon);
    if (APP_RR) {
        st.thr = clamp3(st.thr * k, V3{0.0f, 0.0f, 0.0f}, V3{10.0f, 10.0f, 10.0f}); // worker.cpp:485-488
    } else {
        // clamp(k * Lin, 0, Lin) with Lin >= 0 is min(max(k,0),1) * Lin (renderer.cpp:617-620)
        st.thr = st.thr * clamp3(k, V3{0.0f, 0.0f, 0.0f}, V3{1.0f, 1.0f, 1.0f});
    }
    st.o = at.position + incoming * kEpsilon;
    st.d = normalize(incoming);
    if (APP_RR && (int)bounce < (int)rp.max_depth - 2) { // worker.cpp:497-503
        const float pr = rmax(st.thr.x, rmax(st.thr.y, st.thr.z));
        if (u01(rb.z) > pr) {
            result = make_float4(st.rad.x, st.rad.y, st.rad.z, st.alpha);
            return false;
        }
        st.thr = st.thr / pr;
    }
    const uint32_t left = bounce - 1u;
    if (left == 0) { // trace(0, ...) contributes nothing; trace_iter's loop ends
        result = make_float4(st.rad.x, st.rad.y, st.rad.z, st.alpha);
        return false;
    }
    st.flags = left | (next_event << F_EVENT_SHIFT);
    return true;
}

// Scenes with a sun: the shadow ray of every live path's shade event, compacted into the shadow queue.
// shadow_slot[k] = the path's position in that queue (0xFFFFFFFF: this event casts none).
template <bool APP_RR>
__global__ void __launch_bounds__(SHADE_THREADS, SHADE_MIN_BLOCKS)
    shadow_gen_kernel(DScene S, WaveGeom g, RenderParams rp, const float4* __restrict__ ray_o,
                      const float4* __restrict__ ray_d, const uint4* __restrict__ hits, float4* __restrict__ sh_o,
                      float4* __restrict__ sh_d, uint32_t* __restrict__ shadow_slot, const uint32_t* __restrict__ n_ptr,
                      uint32_t* __restrict__ n_shadow) {
    const uint32_t n = *n_ptr;
    const int lane = threadIdx.x & 31;
    const uint32_t n_warp = (n + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_warp; k += gridDim.x * blockDim.x) {
        ShadowIO sh;
        sh.emit = false;
        sh.visible = false;
        if (k < n) {
            const float4 o4 = __ldg(ray_o + k), d4 = __ldg(ray_d + k);
            PathState st;
            st.o = V3{o4.x, o4.y, o4.z};
            st.d = V3{d4.x, d4.y, d4.z};
            st.thr = V3{1.0f, 1.0f, 1.0f};
            st.rad = V3{0.0f, 0.0f, 0.0f};
            st.alpha = 1.0f;
            st.p = __float_as_uint(o4.w);
            st.flags = __float_as_uint(d4.w);
            float4 unused;
            shade_path<APP_RR, true, SHADOW_GENERATE>(S, g, rp, st, __ldg(hits + k), sh, unused);
        }
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, sh.emit);
        uint32_t base = 0;
        if (lane == 0 && mask) base = atomicAdd(n_shadow, (uint32_t)__popc(mask));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (k < n) {
            uint32_t slot = 0xFFFFFFFFu;
            if (sh.emit) {
                slot = base + __popc(mask & ((1u << lane) - 1u));
                __stcs(sh_o + slot, make_float4(sh.o.x, sh.o.y, sh.o.z, 0.0f));
                __stcs(sh_d + slot, make_float4(sh.d.x, sh.d.y, sh.d.z, 0.0f));
            }
            shadow_slot[k] = slot;
        }
    }
}

template <bool APP_RR, bool HAS_SUN>
__global__ void __launch_bounds__(SHADE_THREADS, SHADE_MIN_BLOCKS)
    shade_kernel(DScene S, WaveGeom g, RenderParams rp, PathBuffers in, const uint4* __restrict__ hits,
                 PathBuffers out, float4* __restrict__ sample_out, const uint32_t* __restrict__ n_ptr,
                 uint32_t* __restrict__ n_next, const uint32_t* __restrict__ shadow_slot,
                 const uint8_t* __restrict__ occluded) {
    const uint32_t n = *n_ptr;
    const int lane = threadIdx.x & 31;
    const uint32_t n_warp = (n + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_warp; k += gridDim.x * blockDim.x) {
        bool alive = false;
        PathState st;
        if (k < n) {
            const float4 o4 = __ldcs(in.ray_o + k), d4 = __ldcs(in.ray_d + k), t4 = __ldcs(in.thr + k),
                         r4 = __ldcs(in.rad + k);
            st.o = V3{o4.x, o4.y, o4.z};
            st.d = V3{d4.x, d4.y, d4.z};
            st.thr = V3{t4.x, t4.y, t4.z};
            st.rad = V3{r4.x, r4.y, r4.z};
            st.alpha = r4.w;
            st.p = __float_as_uint(o4.w);
            st.flags = __float_as_uint(d4.w);
            ShadowIO sh;
            sh.emit = false;
            sh.visible = false;
            if (HAS_SUN) {
                const uint32_t slot = __ldcs(shadow_slot + k);
                if (slot != 0xFFFFFFFFu) sh.visible = __ldcs(occluded + slot) == 0;
            }
            float4 result;
            alive = shade_path<APP_RR, HAS_SUN, HAS_SUN ? SHADOW_USE : SHADOW_NONE>(S, g, rp, st, __ldcs(hits + k), sh, result);
            if (!alive) __stcs(sample_out + st.p, result);
        }
        // survivors, compacted: one atomicAdd per warp (per block, as in raygen, measured slower here: the barriers
        // cost this divergent kernel more than the atomics do — 0.62 against 0.57 ms per 8.3 M paths)
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, alive);
        uint32_t base = 0;
        if (lane == 0 && mask) base = atomicAdd(n_next, (uint32_t)__popc(mask));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (alive) {
            const uint32_t k2 = base + __popc(mask & ((1u << lane) - 1u));
            __stcs(out.ray_o + k2, make_float4(st.o.x, st.o.y, st.o.z, __uint_as_float(st.p)));
            __stcs(out.ray_d + k2, make_float4(st.d.x, st.d.y, st.d.z, __uint_as_float(st.flags)));
            __stcs(out.thr + k2, make_float4(st.thr.x, st.thr.y, st.thr.z, 0.0f));
            __stcs(out.rad + k2, make_float4(st.rad.x, st.rad.y, st.rad.z, st.alpha));
        }
    }
}

// ---------------------------------------------------------- accumulate -----

// The reference blends sample after sample, in order (renderer.cpp:373-399):
// `sample` is the global sample index, so waves can be chained.  The destination is addressed with a row
// pitch, so a tile can be accumulated in place inside a full frame — which may live in ANOTHER GPU's memory
// (peer-mapped over NVLink): the tile return of the multi-GPU frame is this kernel's store, not a collective
// that follows it.  `fresh`: this wave starts the running mean (nothing is read from dst / claimed).
__global__ void __launch_bounds__(256)
    accumulate_kernel(WaveGeom g, const float4* __restrict__ sample_out, float4* __restrict__ dst, uint32_t pitch,
                      uint8_t* __restrict__ claimed, bool transparent, bool fresh) {
    const uint32_t npix = g.w * g.h;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const uint32_t x = i % g.w, y = i / g.w;
        const uint32_t q = pixel_to_slot(g, x, y);
        float4* out = dst + size_t(g.comb_y(y)) * pitch + g.comb_x(x);
        float4 px = fresh ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : *out;
        bool cl = (transparent && !fresh) ? (claimed[i] != 0) : false;
        for (uint32_t s = 0; s < g.wave_samples; s++) {
            const float4 d = __ldcs(sample_out + sample_slot_to_path(g, s, q));
            const uint32_t sample = g.first_sample + s;
            if (transparent) {
                if (d.w > 0.5 && !cl) {
                    px.x = d.x; px.y = d.y; px.z = d.z;
                    px.w = float(1u / (sample + 1u)); // integer division, as written in the reference
                    cl = true;
                    continue;
                } else if (d.w < 0.5 && cl) {
                    px.w = px.w * float(sample) + d.w;
                    px.w = px.w / float(sample + 1u);
                    continue;
                } else if (d.w < 0.5) {
                    continue;
                }
            }
            const float fs = float(sample), fs1 = float(sample + 1u);
            px.x = (px.x * fs + d.x) / fs1;
            px.y = (px.y * fs + d.y) / fs1;
            px.z = (px.z * fs + d.z) / fs1;
            px.w = (px.w * fs + d.w) / fs1;
        }
        *out = px;
        if (transparent) claimed[i] = cl ? 1 : 0;
    }
}

// ------------------------------------------------------------- tonemap -----

__device__ __forceinline__ float aces(float hdr) { // tonemap_approx_aces, per component
    const float a = 2.51F, b = 0.03F, c = 2.43F, d = 0.59F, e = 0.14F;
    const float v = (hdr * (a * hdr + b)) / (hdr * (c * hdr + d) + e);
    return rclamp(v, 0.0f, 1.0f);
}

__global__ void tonemap_kernel(const float* __restrict__ rgb, const float* __restrict__ alpha, uint64_t n,
                               uint8_t* __restrict__ rgba8) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uchar4 o;
        // image::write: sRGB encode pow(v, 1/2.2F), then uint8(v * 255 + 0.5F)
        o.x = (unsigned char)(powf(aces(rgb[3 * i]), 1 / 2.2F) * 255 + 0.5F);
        o.y = (unsigned char)(powf(aces(rgb[3 * i + 1]), 1 / 2.2F) * 255 + 0.5F);
        o.z = (unsigned char)(powf(aces(rgb[3 * i + 2]), 1 / 2.2F) * 255 + 0.5F);
        o.w = (unsigned char)((alpha ? alpha[i] : 1.0f) * 255 + 0.5F);
        reinterpret_cast<uchar4*>(rgba8)[i] = o;
    }
}

__global__ void split_rgba_kernel(const float4* __restrict__ rgba, uint64_t n, float* __restrict__ rgb,
                                  float* __restrict__ alpha) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = rgba[i];
        rgb[3 * i] = v.x;
        rgb[3 * i + 1] = v.y;
        rgb[3 * i + 2] = v.z;
        if (alpha) alpha[i] = v.w;
    }
}

__global__ void join_rgba_kernel(const float* __restrict__ rgb, const float* __restrict__ alpha, uint64_t n,
                                 float4* __restrict__ rgba) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        rgba[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], alpha ? alpha[i] : 1.0f);
}

// frame → RGBA8 in one pass (tonemap_approx_aces + image::write's encode, as tonemap_kernel)
__global__ void tonemap_rgba_kernel(const float4* __restrict__ rgba, uint64_t n, uint8_t* __restrict__ rgba8) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = rgba[i];
        uchar4 o;
        o.x = (unsigned char)(powf(aces(v.x), 1 / 2.2F) * 255 + 0.5F);
        o.y = (unsigned char)(powf(aces(v.y), 1 / 2.2F) * 255 + 0.5F);
        o.z = (unsigned char)(powf(aces(v.z), 1 / 2.2F) * 255 + 0.5F);
        o.w = (unsigned char)(v.w * 255 + 0.5F);
        reinterpret_cast<uchar4*>(rgba8)[i] = o;
    }
}

// ------------------------------------------------- explicit ray sets -------

__global__ void prep_rays_kernel(const float* __restrict__ od, uint64_t n, float4* __restrict__ ray_o,
                                 float4* __restrict__ ray_d) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const V3 d = normalize(V3{od[6 * i + 3], od[6 * i + 4], od[6 * i + 5]}); // geometry::ray::ray, ray.cpp:6-8
        ray_o[i] = make_float4(od[6 * i], od[6 * i + 1], od[6 * i + 2], 0.0f);
        ray_d[i] = make_float4(d.x, d.y, d.z, 0.0f);
    }
}

__global__ void export_hits_kernel(DScene S, const uint4* __restrict__ hits, const float* __restrict__ t, uint64_t n,
                                   ptb_hit* __restrict__ out, float* __restrict__ attrs) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 h = hits[i];
        ptb_hit o;
        if (h.x == HIT_MISS) {
            o.instance = o.surface = o.triangle = PTB_MISS;
            o.t = -1.0f;
            o.bary[0] = o.bary[1] = o.bary[2] = 0.0f;
        } else {
            const float beta = __uint_as_float(h.z), gamma = __uint_as_float(h.w);
            o.instance = h.x >> HIT_SURFACE_BITS;
            o.surface = h.x & ((1u << HIT_SURFACE_BITS) - 1u);
            o.triangle = h.y;
            o.t = t[i];
            o.bary[0] = 1 - beta - gamma;
            o.bary[1] = beta;
            o.bary[2] = gamma;
        }
        out[i] = o;
        if (attrs) {
            float* a = attrs + 14 * i;
            if (h.x == HIT_MISS) {
                for (int j = 0; j < 14; j++) a[j] = 0.0f;
            } else {
                const HitAttrs at = hit_attributes(S, o.instance, o.surface, o.triangle, o.bary[1], o.bary[2]);
                const MatSample m = material_sample(S, at.material, at.u, at.v);
                const V3 sn = shading_normal(at, m.normal_map);
                a[0] = at.position.x; a[1] = at.position.y; a[2] = at.position.z;
                a[3] = at.u; a[4] = at.v;
                a[5] = at.normal.x; a[6] = at.normal.y; a[7] = at.normal.z;
                a[8] = at.tangent.x; a[9] = at.tangent.y; a[10] = at.tangent.z;
                a[11] = sn.x; a[12] = sn.y; a[13] = sn.z;
            }
        }
    }
}

__global__ void camera_rays_kernel(DScene S, uint32_t w, uint32_t h, const uint32_t* __restrict__ px,
                                   const uint32_t* __restrict__ py, const float* __restrict__ aa, uint64_t n,
                                   float* __restrict__ od) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const Ray r = camera_ray(S.camera, px[i], py[i], aa[2 * i], aa[2 * i + 1], w, h);
        od[6 * i] = r.o.x; od[6 * i + 1] = r.o.y; od[6 * i + 2] = r.o.z;
        od[6 * i + 3] = r.d.x; od[6 * i + 4] = r.d.y; od[6 * i + 5] = r.d.z;
    }
}

inline int grid_for(uint64_t n, int threads, int cap) {
    uint64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (uint64_t)cap) b = cap;
    return (int)b;
}

} // namespace

// ------------------------------------------------------------ launchers ----

void launch_raygen(const DScene& S, const WaveGeom& g, const RenderParams& rp, const PathBuffers& out,
                   float4* sample_out, uint32_t* qcount0, const LaunchCfg& cfg, cudaStream_t st) {
    const uint64_t n = uint64_t(g.padded_pixels) * g.wave_samples;
    raygen_kernel<<<grid_for(n, 256, cfg.sm_count * 8), 256, 0, st>>>(S, g, rp, out, sample_out, qcount0);
}

// Persistent grid-stride kernels (shade, shadow_gen): exactly the blocks that are resident at once, capped by the
// option.  One block more than fits (8 requested, 7 resident at 72 registers) runs as a second wave on an almost empty
// machine and the launch takes nearly twice as long as its work (shade: 27.9 against 27.0 ms per two C2 waves).
template <typename F>
int resident_grid(F fn, const LaunchCfg& cfg) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, SHADE_THREADS, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
    return cfg.sm_count * std::max(1, std::min(per_sm, cfg.shade_blocks_per_sm));
}

void launch_shadow_gen(const DScene& S, const WaveGeom& g, const RenderParams& rp, const float4* ray_o,
                       const float4* ray_d, const uint4* hits, float4* sh_o, float4* sh_d, uint32_t* shadow_slot,
                       const uint32_t* n_ptr, uint32_t* n_shadow, const LaunchCfg& cfg, cudaStream_t st) {
    if (rp.integrator == 1)
        shadow_gen_kernel<true><<<resident_grid(shadow_gen_kernel<true>, cfg), SHADE_THREADS, 0, st>>>(
            S, g, rp, ray_o, ray_d, hits, sh_o, sh_d, shadow_slot, n_ptr, n_shadow);
    else
        shadow_gen_kernel<false><<<resident_grid(shadow_gen_kernel<false>, cfg), SHADE_THREADS, 0, st>>>(
            S, g, rp, ray_o, ray_d, hits, sh_o, sh_d, shadow_slot, n_ptr, n_shadow);
}

void launch_shade(const DScene& S, const WaveGeom& g, const RenderParams& rp, const PathBuffers& in,
                  const uint4* hits, const PathBuffers& out, float4* sample_out, const uint32_t* n_ptr,
                  uint32_t* n_next, const uint32_t* shadow_slot, const uint8_t* occluded, const LaunchCfg& cfg,
                  cudaStream_t st) {
    const bool sun = S.sun.enabled != 0;
    if (rp.integrator == 1) {
        if (sun)
            shade_kernel<true, true><<<resident_grid(shade_kernel<true, true>, cfg), SHADE_THREADS, 0, st>>>(
                S, g, rp, in, hits, out, sample_out, n_ptr, n_next, shadow_slot, occluded);
        else
            shade_kernel<true, false><<<resident_grid(shade_kernel<true, false>, cfg), SHADE_THREADS, 0, st>>>(
                S, g, rp, in, hits, out, sample_out, n_ptr, n_next, nullptr, nullptr);
    } else {
        if (sun)
            shade_kernel<false, true><<<resident_grid(shade_kernel<false, true>, cfg), SHADE_THREADS, 0, st>>>(
                S, g, rp, in, hits, out, sample_out, n_ptr, n_next, shadow_slot, occluded);
        else
            shade_kernel<false, false><<<resident_grid(shade_kernel<false, false>, cfg), SHADE_THREADS, 0, st>>>(
                S, g, rp, in, hits, out, sample_out, n_ptr, n_next, nullptr, nullptr);
    }
}

void launch_accumulate(const WaveGeom& g, const float4* sample_out, float4* dst, uint32_t pitch, uint8_t* claimed,
                       bool transparent, bool fresh, cudaStream_t st) {
    const uint64_t n = uint64_t(g.w) * g.h;
    accumulate_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(g, sample_out, dst, pitch, claimed, transparent, fresh);
}

void launch_tonemap(const float* rgb, const float* alpha, uint64_t n, uint8_t* rgba8, cudaStream_t st) {
    tonemap_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(rgb, alpha, n, rgba8);
}

void launch_split_rgba(const float4* rgba, uint64_t n, float* rgb, float* alpha, cudaStream_t st) {
    split_rgba_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(rgba, n, rgb, alpha);
}

void launch_join_rgba(const float* rgb, const float* alpha, uint64_t n, float4* rgba, cudaStream_t st) {
    join_rgba_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(rgb, alpha, n, rgba);
}

void launch_tonemap_rgba(const float4* rgba, uint64_t n, uint8_t* rgba8, cudaStream_t st) {
    tonemap_rgba_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(rgba, n, rgba8);
}

void launch_prep_rays(const float* origin_dir, uint64_t n, float4* ray_o, float4* ray_d, cudaStream_t st) {
    prep_rays_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(origin_dir, n, ray_o, ray_d);
}

void launch_export_hits(const DScene& S, const uint4* hits, const float* t, uint64_t n, void* hits_out, float* attrs_out,
                        cudaStream_t st) {
    export_hits_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(S, hits, t, n, static_cast<ptb_hit*>(hits_out),
                                                                   attrs_out);
}

// ---- geometry-shard merge over peer memory (cluster.py: trace_rays_sharded_dev) -------------------------------
// key = distance bits << 32 | global instance << 12 | surface; a miss is the largest key.  Distances of hits are
// non-negative floats, whose bits order like the numbers, and renderer::intersect keeps the first instance in
// scene order on equal distances (strict <, renderer.cpp:663-669): the integer minimum of the keys over all
// shards IS the unsharded answer.
__global__ void shard_keys_kernel(const uint4* __restrict__ hits, const float* __restrict__ t, uint64_t n,
                                  const uint32_t* __restrict__ instance_map, unsigned long long* __restrict__ local_keys,
                                  ShardPeers peers) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 h = hits[i];
        unsigned long long key = MERGE_MISS_KEY;
        if (h.x != HIT_MISS) {
            const uint32_t inst = instance_map[h.x >> HIT_SURFACE_BITS], surf = h.x & ((1u << HIT_SURFACE_BITS) - 1u);
            key = ((unsigned long long)__float_as_uint(t[i]) << 32) | ((unsigned long long)inst << HIT_SURFACE_BITS) | surf;
            // the exchange step: one 64-bit minimum per peer, straight into that GPU's memory over NVLink
            for (int r = 0; r < peers.world; r++) atomicMin_system(peers.keys[r] + i, key);
        }
        local_keys[i] = key;
    }
}

__global__ void shard_payload_kernel(const uint4* __restrict__ hits, const unsigned long long* __restrict__ local_keys,
                                     const unsigned long long* __restrict__ best_keys, uint64_t n, ShardPeers peers) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = local_keys[i];
        if (key == MERGE_MISS_KEY || key != best_keys[i]) continue;
        // exactly one shard owns the winning instance: plain stores, no race
        const uint4 h = hits[i];
        const uint4 pay = make_uint4(h.y, h.z, h.w, 0u); // triangle, beta, gamma
        for (int r = 0; r < peers.world; r++) peers.payload[r][i] = pay;
    }
}

__global__ void shard_unpack_kernel(const unsigned long long* __restrict__ best_keys, const uint4* __restrict__ payload,
                                    uint64_t n, ptb_hit* __restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = best_keys[i];
        ptb_hit o;
        if (key == MERGE_MISS_KEY) {
            o.instance = o.surface = o.triangle = PTB_MISS;
            o.t = -1.0f;
            o.bary[0] = o.bary[1] = o.bary[2] = 0.0f;
        } else {
            const uint32_t low = (uint32_t)key;
            const uint4 pay = payload[i];
            const float beta = __uint_as_float(pay.y), gamma = __uint_as_float(pay.z);
            o.instance = low >> HIT_SURFACE_BITS;
            o.surface = low & ((1u << HIT_SURFACE_BITS) - 1u);
            o.triangle = pay.x;
            o.t = __uint_as_float((uint32_t)(key >> 32));
            o.bary[0] = 1 - beta - gamma;
            o.bary[1] = beta;
            o.bary[2] = gamma;
        }
        out[i] = o;
    }
}

__global__ void fill_u64_kernel(unsigned long long* p, uint64_t n, unsigned long long v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

void launch_shard_keys(const uint4* hits, const float* t, uint64_t n, const uint32_t* instance_map,
                       unsigned long long* local_keys, const ShardPeers& peers, cudaStream_t st) {
    shard_keys_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(hits, t, n, instance_map, local_keys, peers);
}
void launch_shard_payload(const uint4* hits, const unsigned long long* local_keys, const unsigned long long* best_keys,
                          uint64_t n, const ShardPeers& peers, cudaStream_t st) {
    shard_payload_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(hits, local_keys, best_keys, n, peers);
}
void launch_shard_unpack(const unsigned long long* best_keys, const uint4* payload, uint64_t n, void* hits_out,
                         cudaStream_t st) {
    shard_unpack_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(best_keys, payload, n, static_cast<ptb_hit*>(hits_out));
}
void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t st) {
    fill_u64_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(p, n, v);
}

void launch_camera_rays(const DScene& S, uint32_t w, uint32_t h, const uint32_t* px, const uint32_t* py, const float* aa,
                        uint64_t n, float* origin_dir, cudaStream_t st) {
    camera_rays_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, st>>>(S, w, h, px, py, aa, n, origin_dir);
}

} // namespace ptb
