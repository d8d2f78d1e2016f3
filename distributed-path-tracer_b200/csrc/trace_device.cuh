// trace_device.cuh — closest-hit search on the device: instance loop → slab
// tests → short-stack KD traversal → Cramer's-rule triangle test.
//
// Each function states the reference function whose float arithmetic it
// reproduces operation for operation (compiled with --fmad=false, IEEE
// division and square root; LIB = path-tracer-core/path_tracer_lib/path_tracer):
//   slab_test        geometry::aabb::intersect        LIB/geometry/aabb.cpp:41-67
//   tri_test         geometry::triangle::intersect    LIB/geometry/triangle.cpp:120-190
//   mesh_closest     core::mesh::intersect            LIB/core/mesh.cpp:300-405
//   scene_closest    scene::model::intersect          LIB/scene/model.cpp:20-72
//                    + renderer::intersect's loop     LIB/core/renderer.cpp:645-671
// What differs is everything that is not arithmetic: flat 8-byte nodes instead
// of a pointer tree, a per-thread stack in shared memory laid out
// [level][word][thread] (bank-conflict free), triangles as three float4 loads.
#pragma once

#include "device_scene.hpp"

namespace ptb {

constexpr int KD_STACK_DEPTH = 25; // mesh::build_kd_tree max_depth (LIB/core/mesh.hpp:34): one pending entry per level
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;

struct TraceCounters {
    unsigned long long node_visits, leaf_visits, tri_tests, rays;
};

struct MeshHit {
    float t; // < 0: none (mesh::intersection::has_hit, mesh.cpp:250-252)
    float beta, gamma;
    uint32_t tri;
};

struct SceneHit {
    float t; // world distance; < 0: miss
    float beta, gamma;
    uint32_t tri;
    uint32_t instance;
    uint32_t surface; // ordinal inside the instance
};

// Per-thread view of the shared-memory traversal stack.
struct KdStack {
    uint32_t* base; // &smem[threadIdx.x]
    uint32_t stride; // blockDim.x
    __device__ __forceinline__ void push(int sp, uint32_t node, float tmin, float tmax) const {
        uint32_t* p = base + (sp * 3) * stride;
        p[0] = node;
        p[stride] = __float_as_uint(tmin);
        p[2 * stride] = __float_as_uint(tmax);
    }
    __device__ __forceinline__ void pop(int sp, uint32_t& node, float& tmin, float& tmax) const {
        const uint32_t* p = base + (sp * 3) * stride;
        node = p[0];
        tmin = __uint_as_float(p[stride]);
        tmax = __uint_as_float(p[2 * stride]);
    }
};

__device__ __forceinline__ float comp(V3 v, uint32_t axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// aabb::intersect + intersection::has_hit (far >= 0), given inv = 1 / d (componentwise, correctly rounded).
__device__ __forceinline__ bool slab_test_inv(const float* bmin, const float* bmax, V3 o, V3 inv, float& near_out,
                                              float& far_out) {
    if (bmin[0] > bmax[0] || bmin[1] > bmax[1] || bmin[2] > bmax[2])
        return false;
    V3 t0 = (V3{bmin[0], bmin[1], bmin[2]} - o) * inv;
    V3 t1 = (V3{bmax[0], bmax[1], bmax[2]} - o) * inv;
    V3 nd = V3{rmin(t0.x, t1.x), rmin(t0.y, t1.y), rmin(t0.z, t1.z)};
    V3 fd = V3{rmax(t0.x, t1.x), rmax(t0.y, t1.y), rmax(t0.z, t1.z)};
    float nr = rmax(rmax(nd.x, nd.y), nd.z);
    float fr = rmin(rmin(fd.x, fd.y), fd.z);
    if (nr > fr)
        return false;
    near_out = nr;
    far_out = fr;
    return fr >= 0;
}

__device__ __forceinline__ bool slab_test(const float* bmin, const float* bmax, V3 o, V3 d, float& near_out,
                                          float& far_out) {
    return slab_test_inv(bmin, bmax, o, V3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z}, near_out, far_out);
}

// triangle::intersect.  Returns the distance, or -1 where the reference returns {-1}.
__device__ __forceinline__ float tri_test(V3 a, V3 ab, V3 ac, V3 o, V3 d, float& beta_out, float& gamma_out) {
    // m = [a-b, a-c, dir] (columns), v = a - origin
    V3 v = a - o;
    float c1 = ac.y * d.z - d.y * ac.z;
    float c2 = ab.y * d.z - d.y * ab.z;
    float c3 = ab.y * ac.z - ac.y * ab.z;
    float c4 = v.y * d.z - d.y * v.z;
    float c5 = ab.y * v.z - v.y * ab.z;
    float c6 = ac.y * v.z - v.y * ac.z;
    float inv_det = 1 / (ab.x * c1 - ac.x * c2 + d.x * c3);
    float beta = inv_det * (v.x * c1 - ac.x * c4 - d.x * c6);
    if (beta < 0 - kEpsilon || beta > 1 + kEpsilon)
        return -1.0f;
    float gamma = inv_det * (ab.x * c4 - v.x * c2 + d.x * c5);
    if (gamma < 0 - kEpsilon || gamma + beta > 1 + kEpsilon)
        return -1.0f;
    float dist = inv_det * (ab.x * c6 - ac.x * c5 + v.x * c3);
    beta_out = beta;
    gamma_out = gamma;
    return dist;
}

// mesh::intersect: front-to-back traversal that returns at the first leaf which yields a hit.
template <bool COUNT>
__device__ __forceinline__ MeshHit mesh_closest(const DScene& S, const DMesh& M, V3 o, V3 d, const KdStack& stack,
                                                TraceCounters& cnt) {
    MeshHit none{-1.0f, 0.0f, 0.0f, 0u};
    float tmin, tmax;
    if (!slab_test(M.aabb_min, M.aabb_max, o, d, tmin, tmax))
        return none;
    const uint2* __restrict__ nodes = S.kd_nodes + M.node_base;
    const uint32_t* __restrict__ refs = S.kd_refs + M.ref_base;
    int sp = 0;
    uint32_t node = 0;
    for (;;) {
        uint2 n = __ldg(nodes + node);
        // "Explore down the tree until we reach a leaf" (mesh.cpp:314-370)
        while ((n.y & 3u) != 3u) {
            if (COUNT) cnt.node_visits++;
            const uint32_t axis = n.y & 3u;
            const float split = __uint_as_float(n.x);
            const float oa = comp(o, axis);
            const float split_dist = (split - oa) / comp(d, axis);
            const uint32_t has_l = (n.y >> 2) & 1u, has_r = (n.y >> 3) & 1u;
            const uint32_t li = n.y >> 4, ri = li + has_l;
            const bool left_first = oa < split;
            const uint32_t first = left_first ? (has_l ? li : NO_NODE) : (has_r ? ri : NO_NODE);
            const uint32_t second = left_first ? (has_r ? ri : NO_NODE) : (has_l ? li : NO_NODE);
            if (split_dist < 0 || split_dist > tmax) {
                node = first;
            } else if (split_dist < tmin) {
                node = second;
            } else {
                if (second != NO_NODE) {
                    stack.push(sp, second, split_dist, tmax);
                    sp++;
                }
                node = first;
                tmax = split_dist;
            }
            if (node == NO_NODE)
                break;
            n = __ldg(nodes + node);
        }
        if (node != NO_NODE) {
            // leaf (mesh.cpp:376-401)
            if (COUNT) cnt.leaf_visits++;
            const uint32_t count = n.y >> 2;
            const uint32_t* r = refs + n.x;
            MeshHit best = none;
            for (uint32_t i = 0; i < count; i++) {
                const uint32_t tri = __ldg(r + i);
                const float4* t3 = S.tri + size_t(M.tri_base + tri) * 3;
                const float4 a = __ldg(t3), ab = __ldg(t3 + 1), ac = __ldg(t3 + 2);
                if (COUNT) cnt.tri_tests++;
                float beta, gamma;
                float dist = tri_test(V3{a.x, a.y, a.z}, V3{ab.x, ab.y, ab.z}, V3{ac.x, ac.y, ac.z}, o, d, beta, gamma);
                if (dist >= 0 && dist <= tmax && (dist < best.t || !(best.t >= 0))) {
                    best.t = dist;
                    best.beta = beta;
                    best.gamma = gamma;
                    best.tri = tri;
                }
            }
            if (best.t >= 0)
                return best;
        }
        if (sp == 0)
            return none;
        sp--;
        stack.pop(sp, node, tmin, tmax);
    }
}

// renderer::intersect's search over models, model::intersect inside.
template <bool COUNT>
__device__ __forceinline__ SceneHit scene_closest(const DScene& S, V3 o, V3 d, const KdStack& stack,
                                                  TraceCounters& cnt) {
    SceneHit nearest{-1.0f, 0.0f, 0.0f, 0u, 0xFFFFFFFFu, 0u};
    for (uint32_t i = 0; i < S.n_instances; i++) {
        const DInstance& I = S.instances[i];
        // ray::transform (LIB/geometry/ray.cpp:10-15): the constructor re-normalises the direction
        const V3 ol = apply(I.inv, o);
        const V3 dl = normalize(mul(I.inv.basis, d));
        float nr, fr;
        if (!slab_test(I.aabb_min, I.aabb_max, ol, dl, nr, fr))
            continue;
        MeshHit best{-1.0f, 0.0f, 0.0f, 0u};
        uint32_t best_surface = 0;
        for (uint32_t s = 0; s < I.n_surfaces; s++) {
            const DMesh& M = S.meshes[S.surfaces[I.first_surface + s].mesh];
            MeshHit h = mesh_closest<COUNT>(S, M, ol, dl, stack, cnt);
            if (!(h.t >= 0))
                continue;
            if (h.t < best.t || !(best.t >= 0)) {
                best = h;
                best_surface = s;
            }
        }
        if (!(best.t >= 0))
            continue;
        // local → world distance (model.cpp:62-63)
        const V3 hit_vec = dl * best.t;
        const float tw = length(mul(I.fwd.basis, hit_vec));
        if (!(tw >= 0))
            continue;
        if (tw < nearest.t || !(nearest.t >= 0)) {
            nearest.t = tw;
            nearest.beta = best.beta;
            nearest.gamma = best.gamma;
            nearest.tri = best.tri;
            nearest.instance = i;
            nearest.surface = best_surface;
        }
    }
    return nearest;
}

} // namespace ptb
