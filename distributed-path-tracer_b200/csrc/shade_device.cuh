// shade_device.cuh — hit attributes, materials, BSDF kit and RNG on the device.
//
// Reference functions restated (LIB = path-tracer-core/path_tracer_lib/path_tracer):
//   hit_attributes     renderer::intersect, attribute part   LIB/core/renderer.cpp:688-724
//   shading_normal     intersect_result::get_normal          LIB/core/renderer.cpp:430-435
//   mat_*              core::material getters                LIB/core/material.cpp:6-53
//   tex_sample         image_texture::sample (bilinear,wrap) LIB/image/image_texture.cpp:21-62
//   fresnel_schlick    pbr.cpp:13-25      importance_lambert  pbr.cpp:71-77
//   importance_ggx     pbr.cpp:79-91      geometry_smith*     pbr.cpp:95-114
//   distribution_*     pbr.cpp:118-140    pdf_specular        pbr.cpp:172-184
//   rand_cone_vec      LIB/util/rand_cone_vec.cpp:8-35
//   camera_ray         renderer.cpp:365-370 + camera::get_ray LIB/scene/camera.cpp:10-21
// The reference draws from an unseeded thread_local mt19937 (LIB/core/utils.hpp:8-13);
// here every draw is Philox4x32-10 keyed by the request seed with counter
// (global pixel, sample, shade-event, block), so images agree statistically,
// not bitwise.  Position/normal interpolation keeps the reference's operation
// order; the transcendental-heavy BSDF code uses CUDA's IEEE-accurate float
// functions (no fast-math).
#pragma once

#include "device_scene.hpp"

namespace ptb {

// ------------------------------------------------------------------ RNG ----

struct Philox {
    uint32_t key0, key1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ k0;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ k1;
            c3 = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// uniform in [0,1) with 24 bits, like uniform_real_distribution<float>(0,1)
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }

// ------------------------------------------------------------- camera ------

struct Ray {
    V3 o, d;
};

__device__ __forceinline__ Ray camera_ray(const DCamera& cam, uint32_t px, uint32_t py, float aax, float aay,
                                          uint32_t res_x, uint32_t res_y) {
    float ndcx = ((float(px) + aax) / float(res_x)) * 2 - 1.0f;
    float ndcy = ((float(py) + aay) / float(res_y)) * 2 - 1.0f;
    ndcy = -ndcy;
    const float ratio = float(res_x) / float(res_y);
    float dx = cam.tan_half_fov * ndcx;
    const float dy = cam.tan_half_fov * ndcy;
    dx *= ratio;
    // ray(zero, (dx,dy,-1)) normalises; ray::transform normalises again
    const V3 d0 = normalize(V3{dx, dy, -1.0f});
    Ray r;
    r.o = apply(cam.xf, V3{0.0f, 0.0f, 0.0f});
    r.d = normalize(mul(cam.xf.basis, d0));
    return r;
}

// ------------------------------------------------------ hit attributes -----

struct HitAttrs {
    V3 position;
    float u, v;
    V3 normal;
    V3 tangent;
    uint32_t material;
};

__device__ __forceinline__ V3 ld3(const float* p, uint32_t i) { return V3{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

__device__ __forceinline__ HitAttrs hit_attributes(const DScene& S, uint32_t instance, uint32_t surface, uint32_t tri,
                                                   float beta, float gamma) {
    const DInstance& I = S.instances[instance];
    const DSurface sf = S.surfaces[I.first_surface + surface];
    const DMesh& M = S.meshes[sf.mesh];
    const float alpha = 1 - beta - gamma; // triangle.cpp:185
    const uint32_t t = M.tri_base + tri;
    const uint32_t i0 = M.vtx_base + __float_as_uint(__ldg(&S.tri[size_t(t) * 3].w));
    const uint32_t i1 = M.vtx_base + __float_as_uint(__ldg(&S.tri[size_t(t) * 3 + 1].w));
    const uint32_t i2 = M.vtx_base + __float_as_uint(__ldg(&S.tri[size_t(t) * 3 + 2].w));
    HitAttrs h;
    h.material = sf.material;
    h.position = apply(I.fwd, ld3(S.vtx_pos, i0) * alpha + ld3(S.vtx_pos, i1) * beta + ld3(S.vtx_pos, i2) * gamma);
    h.u = S.vtx_uv[2 * i0] * alpha + S.vtx_uv[2 * i1] * beta + S.vtx_uv[2 * i2] * gamma;
    h.v = S.vtx_uv[2 * i0 + 1] * alpha + S.vtx_uv[2 * i1 + 1] * beta + S.vtx_uv[2 * i2 + 1] * gamma;
    h.normal =
        normalize(mul(I.normal_mat, ld3(S.vtx_nrm, i0) * alpha + ld3(S.vtx_nrm, i1) * beta + ld3(S.vtx_nrm, i2) * gamma));
    h.tangent =
        normalize(mul(I.normal_mat, ld3(S.vtx_tan, i0) * alpha + ld3(S.vtx_tan, i1) * beta + ld3(S.vtx_tan, i2) * gamma));
    return h;
}

// ------------------------------------------------------------ textures -----

__device__ __forceinline__ float tex_read(const DScene& S, const DTexture& T, uint32_t x, uint32_t y, uint32_t c) {
    const unsigned long long index = ((unsigned long long)y * T.width + x) * T.channels + c;
    float value;
    if (T.is_float)
        value = reinterpret_cast<const float*>(S.texels + T.offset)[index];
    else
        value = S.texels[T.offset + index] / 255.0F;
    if (T.srgb && c < 3)
        value = powf(value, 2.2F); // image::read, LIB/image/image.cpp:137-138
    return value;
}

__device__ __forceinline__ float4 tex_pixel(const DScene& S, const DTexture& T, uint32_t x, uint32_t y) {
    float4 c = make_float4(1.0f, 1.0f, 1.0f, 1.0f); // read_pixel: missing channels stay 1
    if (T.channels >= 4) c.w = tex_read(S, T, x, y, 3);
    if (T.channels >= 3) c.z = tex_read(S, T, x, y, 2);
    if (T.channels >= 2) c.y = tex_read(S, T, x, y, 1);
    if (T.channels >= 1) c.x = tex_read(S, T, x, y, 0);
    return c;
}

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float w) {
    return make_float4(rlerp(a.x, b.x, w), rlerp(a.y, b.y, w), rlerp(a.z, b.z, w), rlerp(a.w, b.w, w));
}

// image_texture::sample: bilinear with wrap-around; floor/ceil go through an
// unsigned conversion and `mod` exactly as uvec2(...) % size does.
__device__ __forceinline__ float4 tex_sample(const DScene& S, uint32_t tex, float u, float v) {
    const DTexture& T = S.textures[tex];
    const float cx = u * T.width - 0.5F, cy = (1 - v) * T.height - 0.5F;
    const float fx = floorf(cx), fy = floorf(cy), gx = ceilf(cx), gy = ceilf(cy);
    // uvec2(float): float → uint32 conversion; mod(x, size) = (size + x % size) % size for integers
    auto wrap = [](float f, uint32_t size) {
        const uint32_t x = (uint32_t)(long long)f; // negative values wrap like the C++ conversion on x86-64
        return (size + (x % size)) % size;
    };
    const uint32_t x0 = wrap(fx, T.width), x1 = wrap(gx, T.width), y0 = wrap(fy, T.height), y1 = wrap(gy, T.height);
    const float dx = cx - fx, dy = cy - fy;
    const float4 t = lerp4(tex_pixel(S, T, x0, y0), tex_pixel(S, T, x1, y0), dx);
    const float4 b = lerp4(tex_pixel(S, T, x0, y1), tex_pixel(S, T, x1, y1), dx);
    return lerp4(t, b, dy);
}

struct MatSample {
    V3 albedo;
    float opacity, roughness, metallic;
    V3 emissive; // already times 10 (renderer.cpp:462)
    float ior;
    V3 normal_map; // tangent-space normal
    bool shadow_catcher;
};

__device__ __forceinline__ MatSample material_sample(const DScene& S, uint32_t mat, float u, float v) {
    const DMaterial& m = S.materials[mat];
    MatSample r;
    r.albedo = V3{m.albedo[0], m.albedo[1], m.albedo[2]};
    r.opacity = m.opacity;
    r.roughness = m.roughness;
    r.metallic = m.metallic;
    r.emissive = V3{m.emissive[0], m.emissive[1], m.emissive[2]};
    r.ior = m.ior;
    r.normal_map = V3{0.0f, 0.0f, 1.0f}; // fvec3::backward
    r.shadow_catcher = m.shadow_catcher != 0;
    if (m.any_tex) {
        const uint32_t none = 0xFFFFFFFFu;
        if (m.normal_tex != none) {
            const float4 c = tex_sample(S, m.normal_tex, u, v);
            r.normal_map = V3{c.x, c.y, c.z} * 2.0f - V3{1.0f, 1.0f, 1.0f};
        }
        if (m.albedo_tex != none) {
            const float4 c = tex_sample(S, m.albedo_tex, u, v);
            r.albedo = r.albedo * V3{c.x, c.y, c.z};
        }
        if (m.opacity_tex != none) r.opacity *= tex_sample(S, m.opacity_tex, u, v).w;
        if (m.roughness_tex != none) r.roughness *= tex_sample(S, m.roughness_tex, u, v).y;
        if (m.metallic_tex != none) r.metallic *= tex_sample(S, m.metallic_tex, u, v).z;
        if (m.emissive_tex != none) {
            const float4 c = tex_sample(S, m.emissive_tex, u, v);
            r.emissive = r.emissive * V3{c.x, c.y, c.z};
        }
    }
    r.emissive = r.emissive * 10.0f;
    return r;
}

// intersect_result::get_normal: TBN * tangent-space normal
__device__ __forceinline__ V3 shading_normal(const HitAttrs& h, V3 nm) {
    const V3 binormal = cross(h.normal, h.tangent);
    const M3 tbn{h.tangent, binormal, h.normal};
    return mul(tbn, nm);
}

// --------------------------------------------------------------- BSDF ------

constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float pow5(float x) {
    const float x2 = x * x;
    return x2 * x2 * x;
}

__device__ __forceinline__ float fresnel_schlick(V3 outcoming, V3 incoming, float ior) {
    const V3 halfway = normalize(outcoming + incoming);
    const float cos_theta = dot(outcoming, halfway);
    float f0 = (ior - 1) / (ior + 1);
    f0 *= f0;
    return rlerp(f0, 1.0f, pow5(1 - cos_theta));
}

__device__ __forceinline__ V3 rand_cone_vec(float rnd, float cos_theta, V3 normal) {
    const float phi = rnd * 2 * kPi;
    const float sin_theta = sqrtf(1 - cos_theta * cos_theta);
    float sp, cp;
    sincosf(phi, &sp, &cp);
    const V3 cone{cp * sin_theta, sp * sin_theta, cos_theta};
    V3 helper{0.0f, 0.0f, 0.0f};
    const float inv_sqrt3 = 0.57735026918962576f;
    if (fabsf(normal.x) < inv_sqrt3)
        helper.x = 1;
    else if (fabsf(normal.y) < inv_sqrt3)
        helper.y = 1;
    else
        helper.z = 1;
    const V3 tangent = normalize(cross(normal, helper));
    const V3 binormal = cross(normal, tangent);
    return mul(M3{tangent, binormal, normal}, cone);
}

__device__ __forceinline__ V3 importance_lambert(float r0, float r1, V3 normal) {
    const float theta = acosf(2 * r0 - 1) * 0.5F;
    return rand_cone_vec(r1, cosf(theta), normal);
}

__device__ __forceinline__ V3 importance_ggx(float r0, float r1, V3 normal, V3 outcoming, float roughness) {
    roughness *= roughness;
    roughness *= roughness;
    const float cos_theta = sqrtf((1 - r0) / (1 + (roughness - 1) * r0));
    const V3 halfway = rand_cone_vec(r1, cos_theta, normal);
    return reflect(-outcoming, halfway);
}

__device__ __forceinline__ float geometry_smith_g1(V3 normal, V3 light_dir, float k) {
    const float cos_theta = dot(normal, light_dir);
    return cos_theta / rmax(rlerp(k, 1.0f, cos_theta), kEpsilon);
}

__device__ __forceinline__ float geometry_smith(V3 normal, V3 outcoming, V3 incoming, float roughness) {
    const float r = roughness + 1;
    const float k = (r * r) / 8;
    return geometry_smith_g1(normal, outcoming, k) * geometry_smith_g1(normal, incoming, k);
}

__device__ __forceinline__ float pdf_diffuse(V3 normal, V3 incoming) { return dot(normal, incoming) / kPi; }

__device__ __forceinline__ float distribution_ggx(V3 normal, V3 outcoming, V3 incoming, float roughness) {
    roughness *= roughness;
    roughness *= roughness;
    const V3 halfway = normalize(outcoming + incoming);
    const float cos_phi = dot(normal, halfway);
    const float denom = rlerp(1.0f, roughness, cos_phi * cos_phi);
    const float cos_theta = dot(normal, incoming);
    return cos_theta * roughness / rmax(kPi * denom * denom, kEpsilon);
}

__device__ __forceinline__ float pdf_specular(V3 normal, V3 outcoming, V3 incoming, float roughness) {
    const float dist = distribution_ggx(normal, outcoming, incoming, roughness);
    const float geo = geometry_smith(normal, outcoming, incoming, roughness);
    const float n_dot_o = dot(normal, outcoming);
    const float n_dot_i = dot(normal, incoming);
    return (dist * geo) / rmax(4 * n_dot_o * n_dot_i, kEpsilon);
}

// brdf and mixture "pdf" for one incoming direction (renderer.cpp:581-606, identical at :523-552
// except that the direct-light branch overrides both pdfs with 1)
__device__ __forceinline__ void eval_brdf(V3 normal, V3 outcoming, V3 incoming, V3 albedo, float roughness,
                                          float metallic, float specular_probability, V3& brdf, float& pdf) {
    const float diffuse_pdf = pdf_diffuse(normal, incoming);
    V3 diffuse_brdf = diffuse_pdf * albedo;
    const float specular_pdf = pdf_specular(normal, outcoming, incoming, roughness);
    const V3 specular_brdf{specular_pdf, specular_pdf, specular_pdf};
    V3 fresnel = rlerp(V3{0.04F, 0.04F, 0.04F}, albedo, metallic);
    {
        const V3 halfway = normalize(outcoming + incoming);
        const float cos_theta = dot(outcoming, halfway);
        fresnel = rlerp(fresnel, V3{1.0f, 1.0f, 1.0f}, pow5(1 - cos_theta));
    }
    diffuse_brdf = rlerp(diffuse_brdf, V3{0.0f, 0.0f, 0.0f}, metallic);
    brdf = rlerp(diffuse_brdf, specular_brdf, fresnel);
    pdf = rlerp(diffuse_pdf, specular_pdf, specular_probability);
}

__device__ __forceinline__ V3 clamp3(V3 x, V3 lo, V3 hi) {
    return V3{rclamp(x.x, lo.x, hi.x), rclamp(x.y, lo.y, hi.y), rclamp(x.z, lo.z, hi.z)};
}

} // namespace ptb
