// ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin C-callable harness over the UNMODIFIED reference library
// (/root/reference/path-tracer-core/path_tracer_lib), compiled from the
// reference sources where they lie by oracle/Makefile into oracle/_ref/libptref.so.
// Nothing of the product links or loads this; only tests/, __graft_entry__.smoke()
// and bench.py's CPU-baseline / --impl reference legs do.
//
// This translation unit is compiled with -fno-access-control so that the
// private renderer::trace / renderer::intersect (LIB/core/renderer.hpp:38-58)
// can be driven directly.  It contains no copy of reference code except the
// restatement of worker::trace_iter (APP/processors/worker/worker.cpp:285-514),
// which cannot be compiled here (its translation unit needs the AWS SDK, spdlog,
// nlohmann_json and concurrentqueue); that restatement calls the reference
// library for everything underneath (intersect, pbr::*, rand_cone_vec, material).
//
// LIB/ = path-tracer-core/path_tracer_lib/path_tracer/, APP/ = path-tracer-core/src/.

#include "path_tracer/core/renderer.hpp"

#include "path_tracer/core/pbr.hpp"
#include "path_tracer/core/utils.hpp"
#include "path_tracer/geometry/ray.hpp"
#include "path_tracer/image/image.hpp"
#include "path_tracer/image/image_texture.hpp"
#include "path_tracer/scene/model.hpp"
#include "path_tracer/util/rand_cone_vec.hpp"

#include <atomic>
#include <chrono>
#include <limits>
#include <map>
#include <optional>
#include <thread>

#include "ptb.h"

using namespace math;

namespace {

struct ref_scene {
    core::renderer r;
    // flat views, in the visiting order of renderer::intersect
    std::vector<scene::entity*> instances;        // entities that carry a model
    std::vector<uint32_t> inst_first_surface;     // into `surfaces`
    std::vector<const scene::model::surface*> surfaces;
    std::vector<core::mesh*> meshes;              // unique, first-appearance order
    std::vector<core::material*> materials;       // unique, first-appearance order
    std::map<const core::mesh*, uint32_t> mesh_id;
    std::map<const core::material*, uint32_t> material_id;
    std::map<const scene::model::surface*, std::pair<uint32_t, uint32_t>> surface_id; // → (instance, ordinal)
    std::vector<std::shared_ptr<scene::entity>> keep_alive;
    std::vector<const ::image::texture*> textures; // unique, first-appearance order over the materials' slots
    std::map<const ::image::texture*, uint32_t> texture_id;
};

// Enumerate entities in exactly the order renderer::intersect visits them
// (LIB/core/renderer.cpp:646-671: push all roots in unordered_map order, pop,
// push children, test the popped entity's model).
void index_scene(ref_scene& s) {
    std::stack<scene::entity*> stack;
    for (const auto& [_, entity] : s.r.entities)
        stack.push(entity.get());
    while (!stack.empty()) {
        scene::entity* e = stack.top();
        stack.pop();
        for (const auto& child : e->get_children())
            stack.push(child.get());
        auto model = e->get_component<scene::model>();
        if (!model)
            continue;
        uint32_t inst = static_cast<uint32_t>(s.instances.size());
        s.instances.push_back(e);
        s.inst_first_surface.push_back(static_cast<uint32_t>(s.surfaces.size()));
        uint32_t ordinal = 0;
        for (const auto& surf : model->surfaces) {
            s.surfaces.push_back(&surf);
            s.surface_id[&surf] = {inst, ordinal++};
            if (!s.mesh_id.count(surf.mesh.get())) {
                s.mesh_id[surf.mesh.get()] = static_cast<uint32_t>(s.meshes.size());
                s.meshes.push_back(surf.mesh.get());
            }
            if (!s.material_id.count(surf.material.get())) {
                s.material_id[surf.material.get()] = static_cast<uint32_t>(s.materials.size());
                s.materials.push_back(surf.material.get());
                const core::material& m = *surf.material;
                // occlusion_tex is loaded by the reference but never sampled by the integrators: not exported
                for (const auto& t : {m.normal_tex, m.albedo_tex, m.opacity_tex, m.roughness_tex, m.metallic_tex,
                                      m.emissive_tex}) {
                    if (t && !s.texture_id.count(t.get())) {
                        s.texture_id[t.get()] = static_cast<uint32_t>(s.textures.size());
                        s.textures.push_back(t.get());
                    }
                }
            }
        }
    }
}

scene::transform make_transform(const float* origin, const float* basis) {
    return scene::transform(fvec3(origin[0], origin[1], origin[2]),
                            fmat3(basis[0], basis[1], basis[2], basis[3], basis[4], basis[5], basis[6],
                                  basis[7], basis[8]));
}

void put_transform(const scene::transform& t, float* origin, float* basis) {
    origin[0] = t.origin.x; origin[1] = t.origin.y; origin[2] = t.origin.z;
    basis[0] = t.basis.x.x; basis[1] = t.basis.x.y; basis[2] = t.basis.x.z;
    basis[3] = t.basis.y.x; basis[4] = t.basis.y.y; basis[5] = t.basis.y.z;
    basis[6] = t.basis.z.x; basis[7] = t.basis.z.y; basis[8] = t.basis.z.z;
}

struct silence_cout {
    std::ios::iostate old;
    silence_cout() : old(std::cout.rdstate()) { std::cout.setstate(std::ios::failbit); }
    ~silence_cout() { std::cout.clear(old); }
};

thread_local uint64_t tl_rays = 0;

// ---- restated integrators over the reference library -----------------------

// worker::trace_iter, APP/processors/worker/worker.cpp:285-514 (mode B).
fvec4 trace_iter_app(const core::renderer& r, uint8_t initial_bounce, const geometry::ray& initial_ray) {
    using namespace core;
    geometry::ray current_ray = initial_ray;
    uint8_t bounce_remaining = initial_bounce;
    fvec3 accumulated_color = fvec3::zero;
    fvec3 throughput = fvec3::one;
    float alpha = r.transparent_background ? 0.0f : 1.0f;

    while (bounce_remaining > 0) {
        tl_rays++;
        auto result = r.intersect(current_ray);
        if (!result.hit) {
            if (r.environment) // worker.cpp:308-311
                accumulated_color += throughput * (fvec3(r.environment->sample(equirectangular_proj(current_ray.get_dir()))) *
                                                   r.environment_factor);
            else
                accumulated_color += throughput * r.environment_factor; // worker.cpp:313
            alpha = r.transparent_background ? 0.0f : 1.0f;
            break;
        }
        alpha = 1.0f;

        fvec3 albedo = result.material->get_albedo(result.tex_coord);
        float opacity = result.material->get_opacity(result.tex_coord);
        float roughness = result.material->get_roughness(result.tex_coord);
        float metallic = result.material->get_metallic(result.tex_coord);
        fvec3 emissive = result.material->get_emissive(result.tex_coord) * 10;
        float ior = result.material->ior;

        accumulated_color += throughput * emissive; // worker.cpp:331

        if (!math::is_approx(opacity, 1) && core::rand() > opacity) { // worker.cpp:334-341
            current_ray = geometry::ray(result.position + current_ray.get_dir() * math::epsilon,
                                        current_ray.get_dir());
            continue;
        }

        fvec3 normal = result.get_normal();
        fvec3 outcoming = -current_ray.get_dir();
        if (math::dot(normal, outcoming) <= 0) // worker.cpp:348-350
            break;

        if (result.material->shadow_catcher && bounce_remaining == initial_bounce) { // worker.cpp:353-389
            bool in_shadow = true;
            if (r.sun_light) {
                fvec3 direct_incoming = r.sun_light->get_global_transform().basis * fvec3::backward;
                direct_incoming = util::rand_cone_vec(
                    core::rand(),
                    math::cos(core::rand() * r.sun_light->get_component<scene::sun_light>()->angular_radius),
                    direct_incoming);
                if (math::dot(normal, direct_incoming) > 0) {
                    geometry::ray shadow_ray(result.position + direct_incoming * math::epsilon, direct_incoming);
                    tl_rays++;
                    if (!r.intersect(shadow_ray).hit)
                        in_shadow = false;
                }
            }
            if (in_shadow)
                return fvec4::future;
            current_ray = geometry::ray(result.position + current_ray.get_dir() * math::epsilon,
                                        current_ray.get_dir());
            continue;
        }

        roughness = math::max(roughness, 0.05F);
        float specular_probability = core::pbr::fresnel(outcoming, core::reflect(-outcoming, normal), ior);
        specular_probability = math::max(specular_probability, metallic);
        bool specular_sample = core::rand() < specular_probability;

        if (r.sun_light) { // worker.cpp:400-451
            fvec3 direct_incoming = r.sun_light->get_global_transform().basis * fvec3::backward;
            direct_incoming = util::rand_cone_vec(
                core::rand(),
                math::cos(core::rand() * r.sun_light->get_component<scene::sun_light>()->angular_radius),
                direct_incoming);
            if (math::dot(normal, direct_incoming) > 0) {
                geometry::ray direct_ray(result.position + direct_incoming * math::epsilon, direct_incoming);
                tl_rays++;
                if (!r.intersect(direct_ray).hit) {
                    float diffuse_pdf = pbr::pdf_diffuse(normal, direct_incoming);
                    fvec3 diffuse_brdf = diffuse_pdf * albedo;
                    float specular_pdf = pbr::pdf_specular(normal, outcoming, direct_incoming, roughness);
                    fvec3 specular_brdf(specular_pdf);
                    fvec3 fresnel = lerp(fvec3(0.04F), albedo, metallic);
                    {
                        fvec3 halfway = normalize(outcoming + direct_incoming);
                        float cos_theta = dot(outcoming, halfway);
                        fresnel = lerp(fresnel, fvec3::one, math::pow(1 - cos_theta, 5));
                    }
                    diffuse_brdf = lerp(diffuse_brdf, fvec3::zero, metallic);
                    fvec3 brdf = lerp(diffuse_brdf, specular_brdf, fresnel);
                    diffuse_pdf = 1;
                    specular_pdf = 1;
                    float pdf = lerp(diffuse_pdf, specular_pdf, specular_probability);
                    fvec3 direct_in = r.sun_light->get_component<scene::sun_light>()->energy;
                    fvec3 direct_out = brdf * direct_in / math::max(pdf, math::epsilon);
                    direct_out = math::clamp(direct_out, fvec3::zero, direct_in);
                    accumulated_color += throughput * direct_out;
                }
            }
        }

        fvec2 rand_val(core::rand(), core::rand());
        fvec3 indirect_incoming = specular_sample
                                      ? pbr::importance_specular(rand_val, normal, outcoming, roughness)
                                      : pbr::importance_diffuse(rand_val, normal, outcoming);

        if (math::dot(normal, indirect_incoming) > 0) { // worker.cpp:460-503
            float diffuse_pdf = pbr::pdf_diffuse(normal, indirect_incoming);
            fvec3 diffuse_brdf = diffuse_pdf * albedo;
            float specular_pdf = pbr::pdf_specular(normal, outcoming, indirect_incoming, roughness);
            fvec3 specular_brdf(specular_pdf);
            fvec3 fresnel = lerp(fvec3(0.04F), albedo, metallic);
            {
                fvec3 halfway = normalize(outcoming + indirect_incoming);
                float cos_theta = dot(outcoming, halfway);
                fresnel = lerp(fresnel, fvec3::one, math::pow(1 - cos_theta, 5));
            }
            diffuse_brdf = lerp(diffuse_brdf, fvec3::zero, metallic);
            fvec3 brdf = lerp(diffuse_brdf, specular_brdf, fresnel);
            float pdf = lerp(diffuse_pdf, specular_pdf, specular_probability);
            throughput *= brdf / math::max(pdf, math::epsilon);
            throughput = math::clamp(throughput, fvec3::zero, fvec3(10.0f));
            current_ray = geometry::ray(result.position + indirect_incoming * math::epsilon, indirect_incoming);
            if (bounce_remaining < initial_bounce - 2) {
                float p = math::max(throughput.x, math::max(throughput.y, throughput.z));
                if (core::rand() > p)
                    break;
                throughput /= p;
            }
        } else {
            break;
        }
        bounce_remaining--;
    }
    return fvec4(accumulated_color, alpha);
}

// ---- the STAGED form of the app integrator, restated separately -----------------------------------------------------
// The worker does not call trace_iter on its queue pipeline: a ray travels INTERSECT → [DIRECT_LIGHTING →] SHADING →
// (INTERSECT ... | ACCUMULATE) as a models::cloud_ray (APP/models/cloud_ray.hpp:43-58).  This is that state machine for
// one path and one worker (num_workers = 1: the min / or merges of intersection_worker.cpp:69-147 are identities),
// written from the stage functions, NOT from trace_iter above — a second, independently restated oracle for row I-B:
//   generate_rays                            APP/processors/worker/worker.cpp:117-146
//   process_object_intersections             APP/processors/worker/intersection_worker.cpp:10-47
//   process_object_intersection_results      intersection_worker.cpp:69-112   (which stage comes next)
//   process_direct_lighting_intersections    intersection_worker.cpp:49-67
//   process_shading                          APP/processors/worker/shading_worker.cpp:10-201
// intersect_min_result (APP/scene/intersect.cpp:9-78) returns {hit, distance, position, tbn * normal map}: the same
// closest hit and the same shading normal as intersect() + get_normal(), which is what the library exposes here.
struct staged_ray { // the fields of models::cloud_ray the stages use
    geometry::ray ray;
    std::optional<geometry::ray> direct_light_ray;
    float object_intersect_distance = 0;
    bool direct_light_intersect_result = false;
    fvec3 color = fvec3::zero;
    float alpha = 0;
    fvec3 scale = fvec3::one;
    uint8_t bounce = 0;
    int stage = 0; // 0 INTERSECT, 1 DIRECT_LIGHTING, 2 SHADING, 3 ACCUMULATE
};

static void stage_object_intersection(const core::renderer& r, staged_ray& ray) { // intersection_worker.cpp:21-40, 94-108
    tl_rays++;
    auto result = r.intersect(ray.ray);
    ray.object_intersect_distance = result.hit ? 0.0f : std::numeric_limits<float>::max();
    if (result.hit && r.sun_light) {
        fvec3 direct_incoming = r.sun_light->get_global_transform().basis * fvec3::backward;
        direct_incoming = util::rand_cone_vec(
            core::rand(), math::cos(core::rand() * r.sun_light->get_component<scene::sun_light>()->angular_radius),
            direct_incoming);
        geometry::ray direct_ray(result.position + direct_incoming * math::epsilon, direct_incoming);
        if (math::dot(result.get_normal(), direct_incoming) > 0)
            ray.direct_light_ray = direct_ray;
        else
            ray.direct_light_ray = {};
    }
    // process_object_intersection_results with one worker: the only result is the best one
    if (ray.object_intersect_distance == std::numeric_limits<float>::max())
        ray.stage = 2;
    else
        ray.stage = ray.direct_light_ray.has_value() ? 1 : 2;
}

static void stage_direct_lighting(const core::renderer& r, staged_ray& ray) { // intersection_worker.cpp:58-66, 139-143
    bool hit = false;
    if (ray.direct_light_ray.has_value()) {
        tl_rays++;
        hit = r.intersect(ray.direct_light_ray.value()).hit;
    }
    ray.direct_light_intersect_result = hit;
    ray.stage = 2;
}

static void stage_shading(const core::renderer& r, staged_ray& ray, uint8_t bounce_count) { // shading_worker.cpp:22-199
    using namespace core;
    geometry::ray& current_ray = ray.ray;
    fvec3& accumulated_color = ray.color;
    fvec3& throughput = ray.scale;
    float& alpha = ray.alpha;

    auto result = r.intersect(current_ray); // (the stage intersects again; not counted as a second ray of the path)
    if (!result.hit) {
        if (r.environment)
            accumulated_color += throughput * (fvec3(r.environment->sample(equirectangular_proj(current_ray.get_dir()))) *
                                               r.environment_factor);
        else
            accumulated_color += throughput * r.environment_factor;
        alpha = r.transparent_background ? 0.0f : 1.0f;
        ray.stage = 3;
        return;
    }
    alpha = 1.0f;
    fvec3 albedo = result.material->get_albedo(result.tex_coord);
    float opacity = result.material->get_opacity(result.tex_coord);
    float roughness = result.material->get_roughness(result.tex_coord);
    float metallic = result.material->get_metallic(result.tex_coord);
    fvec3 emissive = result.material->get_emissive(result.tex_coord) * 10;
    float ior = result.material->ior;
    accumulated_color += throughput * emissive;

    if (!math::is_approx(opacity, 1) && core::rand() > opacity) {
        current_ray = geometry::ray(result.position + current_ray.get_dir() * math::epsilon, current_ray.get_dir());
        ray.stage = 0;
        return;
    }
    fvec3 normal = result.get_normal();
    fvec3 outcoming = -current_ray.get_dir();
    if (math::dot(normal, outcoming) <= 0) {
        ray.stage = 3;
        return;
    }
    if (result.material->shadow_catcher && ray.bounce == bounce_count) { // shading_worker.cpp:71-105
        bool in_shadow = true;
        if (r.sun_light && ray.direct_light_ray.has_value()) {
            fvec3 direct_incoming = ray.direct_light_ray.value().get_dir();
            if (math::dot(normal, direct_incoming) > 0 && !ray.direct_light_intersect_result) in_shadow = false;
        }
        if (in_shadow) {
            ray.color = fvec3::zero;
            ray.alpha = 1;
            ray.stage = 3;
        } else {
            current_ray = geometry::ray(result.position + current_ray.get_dir() * math::epsilon, current_ray.get_dir());
            ray.stage = 0;
        }
        return;
    }
    roughness = math::max(roughness, 0.05F);
    float specular_probability = core::pbr::fresnel(outcoming, core::reflect(-outcoming, normal), ior);
    specular_probability = math::max(specular_probability, metallic);
    bool specular_sample = core::rand() < specular_probability;

    if (r.sun_light && ray.direct_light_ray.has_value()) { // shading_worker.cpp:112-146
        fvec3 direct_incoming = ray.direct_light_ray.value().get_dir();
        if (math::dot(normal, direct_incoming) > 0 && !ray.direct_light_intersect_result) {
            float diffuse_pdf = pbr::pdf_diffuse(normal, direct_incoming);
            fvec3 diffuse_brdf = diffuse_pdf * albedo;
            float specular_pdf = pbr::pdf_specular(normal, outcoming, direct_incoming, roughness);
            fvec3 specular_brdf(specular_pdf);
            fvec3 fresnel = lerp(fvec3(0.04F), albedo, metallic);
            fvec3 halfway = normalize(outcoming + direct_incoming);
            fresnel = lerp(fresnel, fvec3::one, math::pow(1 - dot(outcoming, halfway), 5));
            diffuse_brdf = lerp(diffuse_brdf, fvec3::zero, metallic);
            fvec3 brdf = lerp(diffuse_brdf, specular_brdf, fresnel);
            float pdf = lerp(1.0f, 1.0f, specular_probability);
            fvec3 direct_in = r.sun_light->get_component<scene::sun_light>()->energy;
            fvec3 direct_out = math::clamp(brdf * direct_in / math::max(pdf, math::epsilon), fvec3::zero, direct_in);
            accumulated_color += throughput * direct_out;
        }
    }
    fvec2 rand_val(core::rand(), core::rand());
    fvec3 indirect_incoming = specular_sample ? pbr::importance_specular(rand_val, normal, outcoming, roughness)
                                              : pbr::importance_diffuse(rand_val, normal, outcoming);
    if (math::dot(normal, indirect_incoming) > 0) { // shading_worker.cpp:153-193
        float diffuse_pdf = pbr::pdf_diffuse(normal, indirect_incoming);
        fvec3 diffuse_brdf = diffuse_pdf * albedo;
        float specular_pdf = pbr::pdf_specular(normal, outcoming, indirect_incoming, roughness);
        fvec3 specular_brdf(specular_pdf);
        fvec3 fresnel = lerp(fvec3(0.04F), albedo, metallic);
        fvec3 halfway = normalize(outcoming + indirect_incoming);
        fresnel = lerp(fresnel, fvec3::one, math::pow(1 - dot(outcoming, halfway), 5));
        diffuse_brdf = lerp(diffuse_brdf, fvec3::zero, metallic);
        fvec3 brdf = lerp(diffuse_brdf, specular_brdf, fresnel);
        float pdf = lerp(diffuse_pdf, specular_pdf, specular_probability);
        throughput *= brdf / math::max(pdf, math::epsilon);
        throughput = math::clamp(throughput, fvec3::zero, fvec3(10.0f));
        current_ray = geometry::ray(result.position + indirect_incoming * math::epsilon, indirect_incoming);
        if (ray.bounce < bounce_count - 2) {
            float p = math::max(throughput.x, math::max(throughput.y, throughput.z));
            if (core::rand() > p) {
                ray.stage = 3;
                return;
            }
            throughput /= p;
        }
        ray.bounce -= 1;
        ray.stage = ray.bounce > 0 ? 0 : 3;
    } else {
        ray.stage = 3;
    }
}

fvec4 trace_staged_app(const core::renderer& r, uint8_t bounce_count, const geometry::ray& initial_ray) {
    staged_ray ray;
    ray.ray = initial_ray;
    ray.bounce = bounce_count;
    ray.stage = bounce_count > 0 ? 0 : 3;
    ray.alpha = r.transparent_background ? 0.0f : 1.0f; // (cloud_ray leaves it uninitialised; only a 0-bounce request shows it)
    while (ray.stage != 3) {
        if (ray.stage == 0) stage_object_intersection(r, ray);
        else if (ray.stage == 1) stage_direct_lighting(r, ray);
        else stage_shading(r, ray, bounce_count);
    }
    return fvec4(ray.color, ray.alpha);
}

// Iterative equivalent of renderer::trace (LIB/core/renderer.cpp:437-643), used
// (a) to count rays and (b) to check the claim the GPU integrator rests on:
// since indirect_in >= 0 and k = brdf/max(pdf,eps) >= 0,
// clamp(k*Lin, 0, Lin) == min(k,1)*Lin per channel, so the recursion unrolls
// into a throughput product.  Shadow-catcher first-bounce special cases
// (renderer.cpp:513-519,560-561) are kept.
fvec4 trace_iter_lib(const core::renderer& r, uint8_t bounce_count, const geometry::ray& initial_ray) {
    using namespace core;
    geometry::ray ray = initial_ray;
    fvec3 radiance = fvec3::zero;
    fvec3 throughput = fvec3::one;
    float alpha = 1;
    bool primary = true; // alpha is decided by the first non-pass-through event
    uint8_t bounce = bounce_count;
    while (bounce > 0) {
        tl_rays++;
        auto result = r.intersect(ray);
        if (!result.hit) {
            if (primary)
                alpha = r.transparent_background ? 0 : 1;
            if (r.environment) // renderer.cpp:446-448
                radiance += throughput * (fvec3(r.environment->sample(equirectangular_proj(ray.get_dir()))) *
                                          r.environment_factor);
            else
                radiance += throughput * r.environment_factor;
            break;
        }
        fvec3 albedo = result.material->get_albedo(result.tex_coord);
        float opacity = result.material->get_opacity(result.tex_coord);
        float roughness = result.material->get_roughness(result.tex_coord);
        float metallic = result.material->get_metallic(result.tex_coord);
        fvec3 emissive = result.material->get_emissive(result.tex_coord) * 10;
        float ior = result.material->ior;

        if (!math::is_approx(opacity, 1) && core::rand() > opacity) {
            ray = geometry::ray(result.position + ray.get_dir() * math::epsilon, ray.get_dir());
            continue; // same bounce, still "primary" for alpha purposes
        }
        fvec3 normal = result.get_normal();
        fvec3 outcoming = -ray.get_dir();
        primary = false;
        if (math::dot(normal, outcoming) <= 0)
            break;
        roughness = math::max(roughness, 0.05F);
        float specular_probability = pbr::fresnel(outcoming, core::reflect(-outcoming, normal), ior);
        specular_probability = math::max(specular_probability, metallic);
        bool specular_sample = core::rand() < specular_probability;

        fvec3 direct_out;
        if (r.sun_light) {
            fvec3 direct_incoming = r.sun_light->get_global_transform().basis * fvec3::backward;
            direct_incoming = util::rand_cone_vec(
                core::rand(),
                math::cos(core::rand() * r.sun_light->get_component<scene::sun_light>()->angular_radius),
                direct_incoming);
            if (math::dot(normal, direct_incoming) > 0) {
                geometry::ray direct_ray(result.position + direct_incoming * math::epsilon, direct_incoming);
                tl_rays++;
                if (!r.intersect(direct_ray).hit) {
                    if (result.material->shadow_catcher && bounce == bounce_count) {
                        ray = geometry::ray(result.position + ray.get_dir() * math::epsilon, ray.get_dir());
                        primary = true;
                        continue;
                    }
                    float diffuse_pdf = pbr::pdf_diffuse(normal, direct_incoming);
                    fvec3 diffuse_brdf = diffuse_pdf * albedo;
                    float specular_pdf = pbr::pdf_specular(normal, outcoming, direct_incoming, roughness);
                    fvec3 specular_brdf(specular_pdf);
                    fvec3 fresnel = lerp(fvec3(0.04F), albedo, metallic);
                    {
                        fvec3 halfway = normalize(outcoming + direct_incoming);
                        float cos_theta = dot(outcoming, halfway);
                        fresnel = lerp(fresnel, fvec3::one, math::pow(1 - cos_theta, 5));
                    }
                    diffuse_brdf = lerp(diffuse_brdf, fvec3::zero, metallic);
                    fvec3 brdf = lerp(diffuse_brdf, specular_brdf, fresnel);
                    float pdf = lerp(1.0f, 1.0f, specular_probability);
                    fvec3 direct_in = r.sun_light->get_component<scene::sun_light>()->energy;
                    direct_out = brdf * direct_in / math::max(pdf, math::epsilon);
                    direct_out = math::clamp(direct_out, fvec3::zero, direct_in);
                } else if (result.material->shadow_catcher && bounce == bounce_count) {
                    return fvec4(radiance, 1); // black so far, opaque
                }
            }
        }
        radiance += throughput * (direct_out + emissive);

        fvec2 rnd(core::rand(), core::rand());
        fvec3 indirect_incoming = specular_sample ? pbr::importance_specular(rnd, normal, outcoming, roughness)
                                                  : pbr::importance_diffuse(rnd, normal, outcoming);
        if (!(math::dot(normal, indirect_incoming) > 0))
            break;
        float diffuse_pdf = pbr::pdf_diffuse(normal, indirect_incoming);
        fvec3 diffuse_brdf = diffuse_pdf * albedo;
        float specular_pdf = pbr::pdf_specular(normal, outcoming, indirect_incoming, roughness);
        fvec3 specular_brdf(specular_pdf);
        fvec3 fresnel = lerp(fvec3(0.04F), albedo, metallic);
        {
            fvec3 halfway = normalize(outcoming + indirect_incoming);
            float cos_theta = dot(outcoming, halfway);
            fresnel = lerp(fresnel, fvec3::one, math::pow(1 - cos_theta, 5));
        }
        diffuse_brdf = lerp(diffuse_brdf, fvec3::zero, metallic);
        fvec3 brdf = lerp(diffuse_brdf, specular_brdf, fresnel);
        float pdf = lerp(diffuse_pdf, specular_pdf, specular_probability);
        fvec3 k = brdf / math::max(pdf, math::epsilon);
        throughput *= math::clamp(k, fvec3::zero, fvec3::one);
        ray = geometry::ray(result.position + indirect_incoming * math::epsilon, indirect_incoming);
        bounce--;
    }
    return fvec4(radiance, alpha);
}

void count_mesh_visits(const core::mesh& mesh, const geometry::ray& ray, uint64_t* c);

} // namespace

extern "C" {

// ---- construction -----------------------------------------------------------

void* ref_scene_from_gltf(const char* path, uint32_t camera_index, uint32_t sun_light_index) {
    silence_cout quiet;
    auto* s = new ref_scene;
    try {
        s->r.camera_index = camera_index;
        s->r.sun_light_index = sun_light_index;
        s->r.load_gltf(path);
        index_scene(*s);
    } catch (const std::exception& e) {
        std::cerr << "ref_scene_from_gltf: " << e.what() << std::endl;
        delete s;
        return nullptr;
    }
    return s;
}

// Builds reference entities/models/meshes from a flat description
// (recipe: SURVEY.md appendix B).  Instances become children of one root
// entity, attached in reverse so that renderer::intersect visits them in
// array order.  Textures become image::image_texture objects over a copy of the pixels.
void* ref_scene_from_desc(const ptb_scene_desc* d) {
    silence_cout quiet;
    auto* s = new ref_scene;
    std::vector<std::shared_ptr<core::mesh>> meshes;
    for (uint32_t m = 0; m < d->n_meshes; m++) {
        const ptb_mesh_desc& md = d->meshes[m];
        auto mesh = std::make_shared<core::mesh>();
        mesh->vertices.resize(md.n_vertices);
        for (uint32_t v = 0; v < md.n_vertices; v++) {
            core::vertex& vx = mesh->vertices[v];
            vx.position = fvec3(md.positions[3 * v], md.positions[3 * v + 1], md.positions[3 * v + 2]);
            vx.normal = fvec3(md.normals[3 * v], md.normals[3 * v + 1], md.normals[3 * v + 2]);
            vx.tangent = fvec3(md.tangents[3 * v], md.tangents[3 * v + 1], md.tangents[3 * v + 2]);
            vx.tex_coord = fvec2(md.uvs[2 * v], md.uvs[2 * v + 1]);
        }
        mesh->triangles.resize(md.n_triangles);
        for (uint32_t t = 0; t < md.n_triangles; t++)
            mesh->triangles[t] = uvec3(md.indices[3 * t], md.indices[3 * t + 1], md.indices[3 * t + 2]);
        mesh->recalculate_aabb();
        mesh->build_kd_tree(d->kd_use_sah != 0, d->kd_max_depth ? d->kd_max_depth : 25);
        meshes.push_back(mesh);
    }
    std::vector<std::shared_ptr<::image::texture>> textures;
    for (uint32_t t = 0; t < d->n_textures; t++) {
        const ptb_texture_desc& td = d->textures[t];
        auto img = std::make_shared<::image::image>(uvec2(td.width, td.height), td.channels, td.is_float != 0,
                                                    td.srgb != 0);
        const size_t bytes = size_t(td.width) * td.height * td.channels * (td.is_float ? 4 : 1);
        memcpy(img->data.data(), td.pixels, bytes); // private member; this TU is built with -fno-access-control
        textures.push_back(std::make_shared<::image::image_texture>(img));
    }
    auto tex = [&](uint32_t id) -> std::shared_ptr<::image::texture> {
        return id == PTB_NO_TEXTURE ? nullptr : textures[id];
    };
    std::vector<std::shared_ptr<core::material>> materials;
    for (uint32_t m = 0; m < d->n_materials; m++) {
        const ptb_material_desc& md = d->materials[m];
        auto mat = std::make_shared<core::material>();
        mat->normal_tex = tex(md.normal_tex);
        mat->albedo_tex = tex(md.albedo_tex);
        mat->opacity_tex = tex(md.opacity_tex);
        mat->roughness_tex = tex(md.roughness_tex);
        mat->metallic_tex = tex(md.metallic_tex);
        mat->emissive_tex = tex(md.emissive_tex);
        mat->albedo_fac = fvec3(md.albedo[0], md.albedo[1], md.albedo[2]);
        mat->opacity_fac = md.opacity;
        mat->roughness_fac = md.roughness;
        mat->metallic_fac = md.metallic;
        mat->emissive_fac = fvec3(md.emissive[0], md.emissive[1], md.emissive[2]);
        mat->ior = md.ior;
        mat->shadow_catcher = md.shadow_catcher != 0;
        materials.push_back(mat);
    }
    auto root = std::make_shared<scene::entity>();
    root->set_name("root");
    s->keep_alive.push_back(root);
    std::vector<std::shared_ptr<scene::entity>> inst(d->n_instances);
    for (uint32_t i = 0; i < d->n_instances; i++) {
        const ptb_instance_desc& id = d->instances[i];
        auto e = std::make_shared<scene::entity>();
        e->set_name("instance" + std::to_string(i));
        e->set_local_transform(make_transform(id.origin, id.basis));
        auto model = e->add_component<scene::model>();
        for (uint32_t k = 0; k < id.n_surfaces; k++) {
            const ptb_surface_desc& sd = d->surfaces[id.first_surface + k];
            model->surfaces.push_back({meshes[sd.mesh], materials[sd.material]});
        }
        model->recalculate_aabb();
        inst[i] = e;
        s->keep_alive.push_back(e);
    }
    for (uint32_t i = d->n_instances; i-- > 0;)
        inst[i]->set_parent(root);
    s->r.entities["root"] = root;

    auto cam = std::make_shared<scene::entity>();
    cam->set_name("camera");
    cam->set_local_transform(make_transform(d->camera.origin, d->camera.basis));
    cam->add_component<scene::camera>()->set_fov(d->camera.yfov);
    s->r.camera = cam;
    s->keep_alive.push_back(cam);

    if (d->sun.enabled) {
        auto sun = std::make_shared<scene::entity>();
        sun->set_name("sun");
        float zero[3] = {0, 0, 0};
        sun->set_local_transform(make_transform(zero, d->sun.basis));
        auto comp = sun->add_component<scene::sun_light>();
        comp->energy = fvec3(d->sun.energy[0], d->sun.energy[1], d->sun.energy[2]);
        comp->angular_radius = d->sun.angular_radius;
        s->r.sun_light = sun;
        s->keep_alive.push_back(sun);
    }
    s->r.environment_factor =
        fvec3(d->environment_factor[0], d->environment_factor[1], d->environment_factor[2]);
    s->r.transparent_background = d->transparent_background != 0;
    if (d->environment_tex_plus1) s->r.environment = textures[d->environment_tex_plus1 - 1]; // renderer.hpp:28
    index_scene(*s);
    return s;
}

void ref_scene_free(void* h) {
    silence_cout quiet;
    delete static_cast<ref_scene*>(h);
}

// ---- flat export (visiting order) -------------------------------------------

void ref_export_counts(void* h, uint32_t* n_meshes, uint32_t* n_surfaces, uint32_t* n_instances,
                       uint32_t* n_materials) {
    auto* s = static_cast<ref_scene*>(h);
    *n_meshes = s->meshes.size();
    *n_surfaces = s->surfaces.size();
    *n_instances = s->instances.size();
    *n_materials = s->materials.size();
}

void ref_export_mesh_counts(void* h, uint32_t mesh, uint32_t* nv, uint32_t* nt) {
    auto* s = static_cast<ref_scene*>(h);
    *nv = s->meshes[mesh]->vertices.size();
    *nt = s->meshes[mesh]->triangles.size();
}

void ref_export_mesh(void* h, uint32_t mesh, float* pos, float* nrm, float* tan, float* uv, uint32_t* idx,
                     float* aabb6) {
    auto* s = static_cast<ref_scene*>(h);
    const core::mesh& m = *s->meshes[mesh];
    for (size_t v = 0; v < m.vertices.size(); v++) {
        const core::vertex& vx = m.vertices[v];
        pos[3 * v] = vx.position.x; pos[3 * v + 1] = vx.position.y; pos[3 * v + 2] = vx.position.z;
        nrm[3 * v] = vx.normal.x; nrm[3 * v + 1] = vx.normal.y; nrm[3 * v + 2] = vx.normal.z;
        tan[3 * v] = vx.tangent.x; tan[3 * v + 1] = vx.tangent.y; tan[3 * v + 2] = vx.tangent.z;
        uv[2 * v] = vx.tex_coord.x; uv[2 * v + 1] = vx.tex_coord.y;
    }
    for (size_t t = 0; t < m.triangles.size(); t++) {
        idx[3 * t] = m.triangles[t].x; idx[3 * t + 1] = m.triangles[t].y; idx[3 * t + 2] = m.triangles[t].z;
    }
    if (aabb6) {
        // NOTE: build_kd_tree std::move()s mesh.aabb, which for a POD-like struct is a copy: still valid.
        aabb6[0] = m.aabb.min.x; aabb6[1] = m.aabb.min.y; aabb6[2] = m.aabb.min.z;
        aabb6[3] = m.aabb.max.x; aabb6[4] = m.aabb.max.y; aabb6[5] = m.aabb.max.z;
    }
}

void ref_export_surfaces(void* h, ptb_surface_desc* out) {
    auto* s = static_cast<ref_scene*>(h);
    for (size_t i = 0; i < s->surfaces.size(); i++) {
        out[i].mesh = s->mesh_id[s->surfaces[i]->mesh.get()];
        out[i].material = s->material_id[s->surfaces[i]->material.get()];
    }
}

void ref_export_instances(void* h, ptb_instance_desc* out, float* model_aabb6) {
    auto* s = static_cast<ref_scene*>(h);
    for (size_t i = 0; i < s->instances.size(); i++) {
        put_transform(s->instances[i]->get_global_transform(), out[i].origin, out[i].basis);
        out[i].first_surface = s->inst_first_surface[i];
        auto model = s->instances[i]->get_component<scene::model>();
        out[i].n_surfaces = model->surfaces.size();
        if (model_aabb6) {
            float* a = model_aabb6 + 6 * i;
            a[0] = model->aabb.min.x; a[1] = model->aabb.min.y; a[2] = model->aabb.min.z;
            a[3] = model->aabb.max.x; a[4] = model->aabb.max.y; a[5] = model->aabb.max.z;
        }
    }
}

// Returns, per material, which texture slots are populated as a bit mask
// (normal, albedo, opacity, roughness, metallic, emissive = bits 0..5).
void ref_export_materials(void* h, ptb_material_desc* out, uint32_t* tex_mask) {
    auto* s = static_cast<ref_scene*>(h);
    for (size_t i = 0; i < s->materials.size(); i++) {
        const core::material& m = *s->materials[i];
        ptb_material_desc& o = out[i];
        o.albedo[0] = m.albedo_fac.x; o.albedo[1] = m.albedo_fac.y; o.albedo[2] = m.albedo_fac.z;
        o.opacity = m.opacity_fac;
        o.roughness = m.roughness_fac;
        o.metallic = m.metallic_fac;
        o.emissive[0] = m.emissive_fac.x; o.emissive[1] = m.emissive_fac.y; o.emissive[2] = m.emissive_fac.z;
        o.ior = m.ior;
        o.shadow_catcher = m.shadow_catcher;
        auto tid = [&](const std::shared_ptr<::image::texture>& t) -> uint32_t {
            return t ? s->texture_id[t.get()] : PTB_NO_TEXTURE;
        };
        o.normal_tex = tid(m.normal_tex);
        o.albedo_tex = tid(m.albedo_tex);
        o.opacity_tex = tid(m.opacity_tex);
        o.roughness_tex = tid(m.roughness_tex);
        o.metallic_tex = tid(m.metallic_tex);
        o.emissive_tex = tid(m.emissive_tex);
        if (tex_mask)
            tex_mask[i] = (m.normal_tex ? 1u : 0u) | (m.albedo_tex ? 2u : 0u) | (m.opacity_tex ? 4u : 0u) |
                          (m.roughness_tex ? 8u : 0u) | (m.metallic_tex ? 16u : 0u) |
                          (m.emissive_tex ? 32u : 0u);
    }
}

void ref_export_globals(void* h, ptb_camera_desc* cam, ptb_sun_desc* sun, float* env3, uint32_t* transparent) {
    auto* s = static_cast<ref_scene*>(h);
    put_transform(s->r.camera->get_global_transform(), cam->origin, cam->basis);
    cam->yfov = s->r.camera->get_component<scene::camera>()->get_fov();
    memset(sun, 0, sizeof(*sun));
    if (s->r.sun_light) {
        float o[3];
        sun->enabled = 1;
        put_transform(s->r.sun_light->get_global_transform(), o, sun->basis);
        auto c = s->r.sun_light->get_component<scene::sun_light>();
        sun->energy[0] = c->energy.x; sun->energy[1] = c->energy.y; sun->energy[2] = c->energy.z;
        sun->angular_radius = c->angular_radius;
    }
    env3[0] = s->r.environment_factor.x; env3[1] = s->r.environment_factor.y; env3[2] = s->r.environment_factor.z;
    *transparent = s->r.transparent_background;
}

uint32_t ref_export_texture_count(void* h) { return static_cast<uint32_t>(static_cast<ref_scene*>(h)->textures.size()); }

// info5 = width, height, channels, is_float, srgb
void ref_export_texture_info(void* h, uint32_t id, uint32_t* info5) {
    auto* s = static_cast<ref_scene*>(h);
    auto* it = static_cast<const ::image::image_texture*>(s->textures[id]);
    const ::image::image& img = *it->img;
    info5[0] = img.size.x; info5[1] = img.size.y; info5[2] = img.channel_count; info5[3] = img.hdr; info5[4] = img.srgb;
}

void ref_export_texture_data(void* h, uint32_t id, void* out) {
    auto* s = static_cast<ref_scene*>(h);
    auto* it = static_cast<const ::image::image_texture*>(s->textures[id]);
    memcpy(out, it->img->data.data(), it->img->data.size());
}

// ---- KD tree serialisation (same record stream as ptb_scene_dump_kd) ----------

static void dump_node(const core::kd_tree_node* node, std::vector<uint32_t>& out) {
    if (auto b = dynamic_cast<const core::kd_tree_branch*>(node)) {
        uint32_t bits;
        memcpy(&bits, &b->split, 4);
        out.push_back(0x80000000u | b->axis);
        out.push_back(bits);
        out.push_back(b->left ? 1 : 0);
        out.push_back(b->right ? 1 : 0);
        if (b->left) dump_node(b->left.get(), out);
        if (b->right) dump_node(b->right.get(), out);
    } else {
        auto l = static_cast<const core::kd_tree_leaf*>(node);
        out.push_back(static_cast<uint32_t>(l->indices.size()));
        for (uint32_t i : l->indices) out.push_back(i);
    }
}

int ref_dump_kd(void* h, uint32_t mesh, uint32_t* words, uint64_t capacity, uint64_t* n_words) {
    auto* s = static_cast<ref_scene*>(h);
    std::vector<uint32_t> out;
    dump_node(s->meshes[mesh]->kd_tree.get(), out);
    *n_words = out.size();
    if (words) {
        if (capacity < out.size()) return 1;
        memcpy(words, out.data(), out.size() * 4);
    }
    return 0;
}

// ---- closest hits -------------------------------------------------------------

// Scene-level search exactly as renderer::intersect does it (renderer.cpp:645-671),
// through the public scene::model::intersect so that the winning surface and
// triangle are visible; attrs (14 floats/ray) come from the private
// renderer::intersect + intersect_result::get_normal.
void ref_trace_rays(void* h, const float* od, uint64_t n, ptb_hit* hits, float* attrs, int threads) {
    auto* s = static_cast<ref_scene*>(h);
    if (threads <= 0) threads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> next{0};
    // warm the lazily cached global transforms before going parallel (entity.cpp:72-85)
    for (auto* e : s->instances) e->get_global_transform();
    auto work = [&]() {
        const uint64_t chunk = 1024;
        for (;;) {
            uint64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            uint64_t e = std::min(n, b + chunk);
            for (uint64_t i = b; i < e; i++) {
                geometry::ray ray(fvec3(od[6 * i], od[6 * i + 1], od[6 * i + 2]),
                                  fvec3(od[6 * i + 3], od[6 * i + 4], od[6 * i + 5]));
                scene::model::intersection nearest;
                for (auto* ent : s->instances) {
                    auto hit = ent->get_component<scene::model>()->intersect(ray);
                    if (!hit.has_hit()) continue;
                    if (hit.distance < nearest.distance || !nearest.has_hit()) nearest = hit;
                }
                ptb_hit& o = hits[i];
                if (!nearest.has_hit()) {
                    o.instance = o.surface = o.triangle = PTB_MISS;
                    o.t = -1;
                    o.bary[0] = o.bary[1] = o.bary[2] = 0;
                } else {
                    auto id = s->surface_id[nearest.surface];
                    o.instance = id.first;
                    o.surface = id.second;
                    o.triangle = nearest.triangle_index;
                    o.t = nearest.distance;
                    o.bary[0] = nearest.barycentric.x;
                    o.bary[1] = nearest.barycentric.y;
                    o.bary[2] = nearest.barycentric.z;
                }
                if (attrs) {
                    float* a = attrs + 14 * i;
                    auto res = s->r.intersect(ray);
                    if (!res.hit) {
                        for (int k = 0; k < 14; k++) a[k] = 0;
                    } else {
                        fvec3 sn = res.get_normal();
                        a[0] = res.position.x; a[1] = res.position.y; a[2] = res.position.z;
                        a[3] = res.tex_coord.x; a[4] = res.tex_coord.y;
                        a[5] = res.normal.x; a[6] = res.normal.y; a[7] = res.normal.z;
                        a[8] = res.tangent.x; a[9] = res.tangent.y; a[10] = res.tangent.z;
                        a[11] = sn.x; a[12] = sn.y; a[13] = sn.z;
                    }
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

// Primary rays exactly as renderer::render builds them (renderer.cpp:360-370),
// with a caller-supplied jitter (aa) per ray: out = n*6 floats.
void ref_camera_rays(void* h, uint32_t w, uint32_t hgt, const uint32_t* px, const uint32_t* py, const float* aa,
                     uint64_t n, float* od) {
    auto* s = static_cast<ref_scene*>(h);
    uvec2 resolution(w, hgt);
    auto cam = s->r.camera->get_component<scene::camera>();
    for (uint64_t i = 0; i < n; i++) {
        uvec2 pixel(px[i], py[i]);
        fvec2 aa_offset(aa[2 * i], aa[2 * i + 1]);
        fvec2 ndc = ((fvec2(pixel) + aa_offset) / resolution) * 2 - fvec2::one;
        ndc.y = -ndc.y;
        float ratio = static_cast<float>(resolution.x) / resolution.y;
        geometry::ray ray = cam->get_ray(ndc, ratio);
        fvec3 d = ray.get_dir();
        od[6 * i] = ray.origin.x; od[6 * i + 1] = ray.origin.y; od[6 * i + 2] = ray.origin.z;
        od[6 * i + 3] = d.x; od[6 * i + 4] = d.y; od[6 * i + 5] = d.z;
    }
}

// Reference-algorithm visit counts for a ray set (the figures the roofline's
// "algorithmic bytes per ray" are built from):
// c[0] model tests, c[1] surface (mesh) tests, c[2] branch visits, c[3] leaf
// visits, c[4] triangle tests, c[5] stack pushes — mesh.cpp:300-405 semantics.
void ref_count_visits(void* h, const float* od, uint64_t n, uint64_t* c6) {
    auto* s = static_cast<ref_scene*>(h);
    for (int k = 0; k < 6; k++) c6[k] = 0;
    for (uint64_t i = 0; i < n; i++) {
        geometry::ray ray(fvec3(od[6 * i], od[6 * i + 1], od[6 * i + 2]),
                          fvec3(od[6 * i + 3], od[6 * i + 4], od[6 * i + 5]));
        for (auto* ent : s->instances) {
            auto model = ent->get_component<scene::model>();
            c6[0]++;
            auto view_ray = ray.transform(ent->get_global_transform().inverse());
            if (!model->aabb.intersect(view_ray).has_hit()) continue;
            for (const auto& surf : model->surfaces) {
                c6[1]++;
                count_mesh_visits(*surf.mesh, view_ray, c6);
            }
        }
    }
}

// ---- rendering ------------------------------------------------------------------

// Linear (pre-tonemap) running-mean radiance of a tile, built like
// renderer::render's inner loop (renderer.cpp:357-400) but keeping float data.
// mode 0: the reference's own recursive renderer::trace
// mode 1: restated worker::trace_iter (APP_RR)
// mode 2: iterative equivalent of mode 0 (counts rays)
// mode 3: the staged form of the app integrator (shading_worker.cpp / intersection_worker.cpp), restated separately
// rays_out: scene-level intersect calls (modes 1 and 2 only; 0 for mode 0).
void ref_render_linear(void* h, uint32_t full_w, uint32_t full_h, uint32_t x0, uint32_t y0, uint32_t w,
                       uint32_t hgt, uint32_t spp, uint32_t depth, int mode, int first_sample_unjittered,
                       int threads, float* rgb, float* alpha, uint64_t* rays_out, double* seconds_out) {
    auto* s = static_cast<ref_scene*>(h);
    if (threads <= 0) threads = std::max(1u, std::thread::hardware_concurrency());
    s->r.bounce_count = depth;
    s->r.resolution = uvec2(full_w, full_h);
    for (auto* e : s->instances) e->get_global_transform();
    s->r.camera->get_global_transform();
    if (s->r.sun_light) s->r.sun_light->get_global_transform();
    struct pixel { fvec3 color; float alpha; bool claimed; };
    std::vector<pixel> pixels(size_t(w) * hgt, pixel{fvec3::zero, 0, false});
    std::atomic<uint32_t> next_row{0};
    std::atomic<uint64_t> rays{0};
    uvec2 resolution(full_w, full_h);
    auto cam = s->r.camera->get_component<scene::camera>();
    const bool transparent = s->r.transparent_background;
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        tl_rays = 0;
        for (;;) {
            uint32_t row = next_row.fetch_add(1);
            if (row >= hgt) break;
            for (uint32_t col = 0; col < w; col++) {
                pixel& p = pixels[size_t(row) * w + col];
                for (uint32_t sample = 0; sample < spp; sample++) {
                    uvec2 px(x0 + col, y0 + row);
                    fvec2 aa_offset = (sample == 0 && first_sample_unjittered) ? fvec2(0, 0)
                                                                                 : fvec2(core::rand(), core::rand());
                    fvec2 ndc = ((fvec2(px) + aa_offset) / resolution) * 2 - fvec2::one;
                    ndc.y = -ndc.y;
                    float ratio = static_cast<float>(resolution.x) / resolution.y;
                    geometry::ray ray = cam->get_ray(ndc, ratio);
                    fvec4 data = mode == 0   ? s->r.trace(depth, ray)
                                 : mode == 1 ? trace_iter_app(s->r, depth, ray)
                                 : mode == 3 ? trace_staged_app(s->r, depth, ray)
                                             : trace_iter_lib(s->r, depth, ray);
                    if (transparent) { // renderer.cpp:374-393
                        if (data.w > 0.5 && !p.claimed) {
                            p.color = fvec3(data);
                            p.alpha = 1 / (sample + 1);
                            p.claimed = true;
                            continue;
                        } else if (data.w < 0.5 && p.claimed) {
                            p.alpha = p.alpha * sample + data.w;
                            p.alpha /= sample + 1;
                            continue;
                        } else if (data.w < 0.5) {
                            continue;
                        }
                    }
                    p.color = p.color * sample + fvec3(data); // renderer.cpp:396-399
                    p.color /= sample + 1;
                    p.alpha = p.alpha * sample + data.w;
                    p.alpha /= sample + 1;
                }
            }
        }
        rays += tl_rays;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < pixels.size(); i++) {
        rgb[3 * i] = pixels[i].color.x; rgb[3 * i + 1] = pixels[i].color.y; rgb[3 * i + 2] = pixels[i].color.z;
        if (alpha) alpha[i] = pixels[i].alpha;
    }
    if (rays_out) *rays_out = rays.load();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
}

// The reference's own entry point, untouched: renderer::render()
// (renderer.cpp:334-428), timed with steady_clock around the call only.
// Returns the PNG size; png_out may be NULL.
uint64_t ref_render_png(void* h, uint32_t w, uint32_t hgt, uint32_t spp, uint32_t depth, uint32_t threads,
                        uint8_t* png_out, uint64_t capacity, double* seconds_out) {
    auto* s = static_cast<ref_scene*>(h);
    silence_cout quiet;
    s->r.resolution = uvec2(w, hgt);
    s->r.sample_count = spp;
    s->r.bounce_count = depth;
    s->r.thread_count = threads;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<uint8_t> png = s->r.render();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (png_out && capacity >= png.size()) memcpy(png_out, png.data(), png.size());
    return png.size();
}

// tonemap_approx_aces + image::write, via the reference functions themselves.
void ref_tonemap_rgba8(const float* rgb, const float* alpha, uint64_t n, uint8_t* out) {
    auto img = std::make_shared<::image::image>(uvec2(static_cast<uint32_t>(n), 1), 4, false, true);
    for (uint64_t i = 0; i < n; i++) {
        fvec3 c = core::tonemap_approx_aces(fvec3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
        uvec2 px(static_cast<uint32_t>(i), 0);
        img->write(px, 0, c.x);
        img->write(px, 1, c.y);
        img->write(px, 2, c.z);
        img->write(px, 3, alpha ? alpha[i] : 1.0f);
    }
    // image::data is private; this TU is built with -fno-access-control
    memcpy(out, img->data.data(), n * 4);
}

uint32_t ref_hardware_threads() { return std::thread::hardware_concurrency(); }

} // extern "C"

namespace {
// Counting twin of mesh::intersect's control flow (mesh.cpp:300-405); results
// are not needed, only how many nodes / triangles the reference touches.
void count_mesh_visits(const core::mesh& mesh, const geometry::ray& ray, uint64_t* c) {
    auto result = mesh.aabb.intersect(ray);
    if (!result.has_hit()) return;
    std::stack<std::tuple<const core::kd_tree_node*, float, float>> stack;
    stack.push({mesh.kd_tree.get(), result.near, result.far});
    while (!stack.empty()) {
        auto [node, min_dist, max_dist] = stack.top();
        stack.pop();
        while (node && typeid(*node) == typeid(core::kd_tree_branch)) {
            c[2]++;
            auto branch = static_cast<const core::kd_tree_branch*>(node);
            float split_dist = (branch->split - ray.origin[branch->axis]) / ray.get_dir()[branch->axis];
            const core::kd_tree_node *first, *second;
            if (ray.origin[branch->axis] < branch->split) {
                first = branch->left.get(); second = branch->right.get();
            } else {
                first = branch->right.get(); second = branch->left.get();
            }
            if (split_dist < 0 || split_dist > max_dist) node = first;
            else if (split_dist < min_dist) node = second;
            else {
                if (second) { stack.push({second, split_dist, max_dist}); c[5]++; }
                node = first;
                max_dist = split_dist;
            }
        }
        if (!node) continue;
        c[3]++;
        auto leaf = static_cast<const core::kd_tree_leaf*>(node);
        geometry::triangle::intersection nearest_hit;
        for (uint32_t i = 0; i < leaf->triangles.size(); i++) {
            c[4]++;
            auto hit = leaf->triangles[i].intersect(ray);
            if (hit.has_hit() && hit.distance <= max_dist &&
                (hit.distance < nearest_hit.distance || !nearest_hit.has_hit()))
                nearest_hit = hit;
        }
        if (nearest_hit.has_hit()) return;
    }
}
} // namespace
