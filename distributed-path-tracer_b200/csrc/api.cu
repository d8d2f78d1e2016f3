// api.cu — the extern "C" boundary declared in include/ptb.h.  Catches every
// exception, maps it to a ptb_status and a thread-local message.

#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include <fstream>
#include <functional>

#include "errors.hpp"
#include "frame.hpp"
#include "json.hpp"
#include "kd_build.hpp"
#include "kernels.hpp"
#include "render.hpp"
#include "scene.hpp"

namespace {

thread_local std::string tl_error;

template <typename F>
ptb_status guarded(F&& f) {
    try {
        f();
        tl_error.clear();
        return PTB_OK;
    } catch (const ptb::Error& e) {
        tl_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        tl_error = "host allocation failed";
        return PTB_E_OOM;
    } catch (const std::exception& e) {
        tl_error = e.what();
        return PTB_E_INVALID;
    } catch (...) {
        tl_error = "unknown error";
        return PTB_E_INVALID;
    }
}

} // namespace

struct ptb_desc {
    ptb::OwnedScene owned;
    ptb_scene_desc view;
};

extern "C" {

ptb_status ptb_scene_create(const ptb_scene_desc* desc, int device, ptb_scene** out) {
    return guarded([&] {
        if (!desc || !out) throw ptb::Error(PTB_E_INVALID, "desc or out is NULL");
        *out = nullptr;
        *out = ptb::create_scene(*desc, device);
    });
}

ptb_status ptb_scene_load_gltf(const char* path, uint32_t camera_index, uint32_t sun_light_index, int device,
                               ptb_scene** out) {
    return guarded([&] {
        if (!path || !out) throw ptb::Error(PTB_E_INVALID, "path or out is NULL");
        *out = nullptr;
        ptb::OwnedScene owned;
        ptb::load_gltf(path, camera_index, sun_light_index, owned);
        *out = ptb::create_scene(owned.view(), device);
    });
}

void ptb_scene_destroy(ptb_scene* scene) {
    try {
        ptb::destroy_scene(scene);
    } catch (...) {
    }
}

ptb_status ptb_scene_get_info(const ptb_scene* scene, ptb_scene_info* out) {
    return guarded([&] {
        if (!scene || !out) throw ptb::Error(PTB_E_INVALID, "scene or out is NULL");
        *out = scene->info;
    });
}

ptb_status ptb_scene_dump_kd(const ptb_scene* scene, uint32_t mesh, uint32_t* words, uint64_t capacity,
                             uint64_t* n_words) {
    return guarded([&] {
        if (!scene || !n_words) throw ptb::Error(PTB_E_INVALID, "scene or n_words is NULL");
        if (scene->replica) throw ptb::Error(PTB_E_INVALID, "a replica holds no host copy of the trees; dump the built scene");
        if (mesh >= scene->trees.size()) throw ptb::Error(PTB_E_INVALID, "mesh out of range");
        std::vector<uint32_t> w;
        ptb::dump_kd_tree(scene->trees[mesh], w);
        *n_words = w.size();
        if (words) {
            if (capacity < w.size()) throw ptb::Error(PTB_E_INVALID, "capacity too small");
            std::memcpy(words, w.data(), w.size() * sizeof(uint32_t));
        }
    });
}

// Host-only twin of the build step of ptb_scene_create, for tests that have no GPU.
ptb_status ptb_host_build_kd(const float* positions, uint32_t n_vertices, const uint32_t* indices,
                             uint32_t n_triangles, uint32_t use_sah, uint32_t max_depth, int threads, uint32_t* words,
                             uint64_t capacity, uint64_t* n_words, float* aabb6_out) {
    return guarded([&] {
        if (!n_words) throw ptb::Error(PTB_E_INVALID, "n_words is NULL");
        if ((n_vertices && !positions) || (n_triangles && !indices)) throw ptb::Error(PTB_E_INVALID, "NULL input");
        for (size_t i = 0; i < size_t(n_triangles) * 3; i++)
            if (indices[i] >= n_vertices) throw ptb::Error(PTB_E_INVALID, "triangle index out of range");
        if (max_depth == 0) max_depth = 25;
        const ptb::Aabb box = ptb::mesh_aabb(positions, n_vertices);
        if (aabb6_out) {
            for (int a = 0; a < 3; a++) {
                aabb6_out[a] = box.min[a];
                aabb6_out[3 + a] = box.max[a];
            }
        }
        ptb::KdTree tree;
        ptb::build_kd_tree_cached(positions, n_vertices, indices, n_triangles, box, use_sah != 0, max_depth, threads, tree);
        std::vector<uint32_t> w;
        ptb::dump_kd_tree(tree, w);
        *n_words = w.size();
        if (words) {
            if (capacity < w.size()) throw ptb::Error(PTB_E_INVALID, "capacity too small");
            std::memcpy(words, w.data(), w.size() * sizeof(uint32_t));
        }
    });
}

// Host-only: parse a glTF file into a scene description without touching CUDA.
ptb_status ptb_desc_load_gltf(const char* path, uint32_t camera_index, uint32_t sun_light_index, ptb_desc** out) {
    return guarded([&] {
        if (!path || !out) throw ptb::Error(PTB_E_INVALID, "path or out is NULL");
        *out = nullptr;
        auto* d = new ptb_desc;
        try {
            ptb::load_gltf(path, camera_index, sun_light_index, d->owned);
            d->view = d->owned.view();
        } catch (...) {
            delete d;
            throw;
        }
        *out = d;
    });
}

const ptb_scene_desc* ptb_desc_get(const ptb_desc* d) { return d ? &d->view : nullptr; }

void ptb_desc_free(ptb_desc* d) { delete d; }

ptb_status ptb_trace_rays(const ptb_scene* scene, const float* origin_dir, uint64_t n, ptb_hit* hits_out) {
    return guarded([&] { ptb::trace_rays_host(scene, origin_dir, n, hits_out, nullptr, nullptr); });
}

ptb_status ptb_trace_rays_attrs(const ptb_scene* scene, const float* origin_dir, uint64_t n, ptb_hit* hits_out,
                                float* attrs_out) {
    return guarded([&] { ptb::trace_rays_host(scene, origin_dir, n, hits_out, attrs_out, nullptr); });
}

ptb_status ptb_trace_rays_stats(const ptb_scene* scene, const float* origin_dir, uint64_t n, ptb_hit* hits_out,
                                ptb_render_stats* stats_out) {
    return guarded([&] { ptb::trace_rays_host(scene, origin_dir, n, hits_out, nullptr, stats_out); });
}

ptb_status ptb_trace_occlusion(const ptb_scene* scene, const float* origin_dir, uint64_t n, uint8_t* occluded_out,
                               ptb_render_stats* stats_out) {
    return guarded([&] { ptb::trace_occlusion_host(scene, origin_dir, n, occluded_out, stats_out); });
}

ptb_status ptb_camera_rays(const ptb_scene* scene, uint32_t w, uint32_t h, const uint32_t* px, const uint32_t* py,
                           const float* aa, uint64_t n, float* origin_dir_out) {
    return guarded([&] {
        if (n && (!px || !py || !aa || !origin_dir_out)) throw ptb::Error(PTB_E_INVALID, "NULL argument");
        ptb::camera_rays_host(scene, w, h, px, py, aa, n, origin_dir_out);
    });
}

ptb_status ptb_render_tile(const ptb_scene* scene, const ptb_tile_req* req, float* rgb_out, float* alpha_out,
                           ptb_render_stats* stats_out) {
    return guarded([&] {
        if (!req) throw ptb::Error(PTB_E_INVALID, "req is NULL");
        ptb::render_tile_host(scene, *req, rgb_out, alpha_out, stats_out);
    });
}

ptb_status ptb_render_tile_dev(const ptb_scene* scene, const ptb_tile_req* req, void* rgba_dev, void* stream,
                               ptb_render_stats* stats_out) {
    return guarded([&] {
        if (!req) throw ptb::Error(PTB_E_INVALID, "req is NULL");
        ptb::render_tile_dev(scene, *req, static_cast<float4*>(rgba_dev), static_cast<cudaStream_t>(stream), stats_out);
    });
}

ptb_status ptb_trace_rays_dev(const ptb_scene* scene, const float* origin_dir_dev, uint64_t n, ptb_hit* hits_dev,
                              void* stream) {
    return guarded([&] { ptb::trace_rays_dev(scene, origin_dir_dev, n, hits_dev, static_cast<cudaStream_t>(stream)); });
}

ptb_status ptb_shard_reset_dev(uint64_t* keys_dev, uint64_t n, void* stream) {
    return guarded([&] { ptb::shard_reset_dev(keys_dev, n, static_cast<cudaStream_t>(stream)); });
}

ptb_status ptb_shard_trace_dev(const ptb_scene* shard, const float* origin_dir_dev, uint64_t n,
                               const uint32_t* instance_map_dev, void* const* peer_keys, int world, void* stream) {
    return guarded([&] {
        ptb::shard_trace_dev(shard, origin_dir_dev, n, instance_map_dev, peer_keys, world, static_cast<cudaStream_t>(stream));
    });
}

ptb_status ptb_shard_occlusion_dev(const ptb_scene* shard, const float* origin_dir_dev, uint64_t n, void* const* peer_occluded,
                                   int world, void* stream) {
    return guarded([&] {
        ptb::shard_occlusion_dev(shard, origin_dir_dev, n, peer_occluded, world, static_cast<cudaStream_t>(stream));
    });
}

ptb_status ptb_shard_publish_dev(const ptb_scene* shard, uint64_t n, const uint64_t* best_keys_dev,
                                 void* const* peer_payload, int world, void* stream) {
    return guarded([&] {
        ptb::shard_publish_dev(shard, n, best_keys_dev, peer_payload, world, static_cast<cudaStream_t>(stream));
    });
}

ptb_status ptb_shard_unpack_dev(const uint64_t* best_keys_dev, const void* payload_dev, uint64_t n, ptb_hit* hits_dev,
                                void* stream) {
    return guarded([&] { ptb::shard_unpack_dev(best_keys_dev, payload_dev, n, hits_dev, static_cast<cudaStream_t>(stream)); });
}

ptb_status ptb_tonemap_rgba8(const float* rgb, const float* alpha, uint64_t n_pixels, uint8_t* rgba8_out) {
    return guarded([&] { ptb::tonemap_host(rgb, alpha, n_pixels, rgba8_out); });
}

ptb_status ptb_write_png(const char* path, const uint8_t* rgba8, uint32_t w, uint32_t h) {
    return guarded([&] {
        if (!path || !rgba8 || !w || !h) throw ptb::Error(PTB_E_INVALID, "bad argument");
        ptb::write_png_rgba8(path, rgba8, w, h);
    });
}

ptb_status ptb_set_option(const char* name, int64_t value) {
    return guarded([&] {
        if (!name) throw ptb::Error(PTB_E_INVALID, "name is NULL");
        const std::string n(name);
        if (n == "wave_paths") {
            if (value < 1024) throw ptb::Error(PTB_E_INVALID, "wave_paths must be >= 1024");
            ptb::g_options.wave_paths = value;
        } else if (n == "count_visits") {
            ptb::g_options.count_visits = value ? 1 : 0;
        } else if (n == "extend_variant") {
            if (value < 0 || value > 4 || value == 2) throw ptb::Error(PTB_E_INVALID, "extend_variant must be 0, 1, 3 or 4");
#ifndef PTB_BUILD_EXPERIMENTS
            if (value != 1)
                throw ptb::Error(PTB_E_INVALID, "extend_variant " + std::to_string(value) +
                                                    " is an experiment: rebuild with PTB_BUILD_EXPERIMENTS=1");
#endif
            ptb::g_options.extend_variant = value;
        } else if (n == "extend_steps") {
            if (value != 0 && (value < 2 || value > 8)) throw ptb::Error(PTB_E_INVALID, "extend_steps must be 0 (by tree size) or 2..8");
            ptb::g_options.extend_steps = value;
        } else if (n == "extend_setup_lanes") {
            if (value < 1 || value > 32) throw ptb::Error(PTB_E_INVALID, "extend_setup_lanes must be 1..32");
            ptb::g_options.extend_setup_lanes = value;
        } else if (n == "frame_comb_tiles") {
            if (value < 0 || value > 2) throw ptb::Error(PTB_E_INVALID, "frame_comb_tiles must be 0, 1 or 2");
            ptb::g_options.frame_comb_tiles = value;
        } else if (n == "frame_comb_rounds") {
            if (value < 0 || value > 64) throw ptb::Error(PTB_E_INVALID, "frame_comb_rounds must be 0..64");
            ptb::g_options.frame_comb_rounds = value;
        } else if (n == "extend_dense_min2") {
            if (value < 0 || value > 32) throw ptb::Error(PTB_E_INVALID, "extend_dense_min2 must be 0..32");
            ptb::g_options.extend_dense_min2 = value;
        } else if (n == "extend_dense") {
            if (value < 0 || value > 1) throw ptb::Error(PTB_E_INVALID, "extend_dense must be 0 or 1");
            ptb::g_options.extend_dense = value;
        } else if (n == "extend_defer") {
            if (value < 0 || value > 1) throw ptb::Error(PTB_E_INVALID, "extend_defer must be 0 or 1");
            ptb::g_options.extend_defer = value;
        } else if (n == "path_order") {
            ptb::g_options.path_order = value != 0;
        } else if (n == "extend_contexts") {
            if (value < 2 || value > 4) throw ptb::Error(PTB_E_INVALID, "extend_contexts must be 2..4");
            ptb::g_options.extend_contexts = value;
        } else if (n == "extend_rays_per_lane") {
            if (value < 0 || value > 4096) throw ptb::Error(PTB_E_INVALID, "extend_rays_per_lane must be 0..4096");
            ptb::g_options.extend_rays_per_lane = value;
        } else if (n == "extend_sm_ranges") {
            ptb::g_options.extend_sm_ranges = value != 0;
        } else if (n == "extend_tests") {
            if (value < 1 || value > 2) throw ptb::Error(PTB_E_INVALID, "extend_tests must be 1 or 2");
            ptb::g_options.extend_tests = value;
        } else if (n == "time_stages") {
            ptb::g_options.time_stages = value ? 1 : 0;
        } else if (n == "group_timeout_ms") {
            if (value < 100) throw ptb::Error(PTB_E_INVALID, "group_timeout_ms must be >= 100");
            ptb::g_options.group_timeout_ms = value;
        } else if (n == "frame_tiles_in_flight") {
            if (value < 1 || value > 16) throw ptb::Error(PTB_E_INVALID, "frame_tiles_in_flight must be 1..16");
            ptb::g_options.frame_tiles_in_flight = value;
        } else if (n == "frame_guided_tiles") {
            ptb::g_options.frame_guided_tiles = value ? 1 : 0;
        } else if (n == "frame_spin_wait") {
            ptb::g_options.frame_spin_wait = value ? 1 : 0;
        } else if (n == "frame_queue_depth") {
            if (value < 1 || value > 2) throw ptb::Error(PTB_E_INVALID, "frame_queue_depth must be 1 or 2");
            ptb::g_options.frame_queue_depth = value;
        } else if (n == "extend_blocks_per_sm") {
            if (value < 1 || value > 32) throw ptb::Error(PTB_E_INVALID, "extend_blocks_per_sm out of range");
            ptb::g_options.extend_blocks_per_sm = value;
        } else if (n == "shade_blocks_per_sm") {
            if (value < 1 || value > 32) throw ptb::Error(PTB_E_INVALID, "shade_blocks_per_sm out of range");
            ptb::g_options.shade_blocks_per_sm = value;
        } else {
            throw ptb::Error(PTB_E_INVALID, "unknown option: " + n);
        }
    });
}

// PTB_OPTIONS="name=value,name=value": tuning options applied when the library is loaded (same names and
// checks as ptb_set_option; a bad entry is reported on stderr and skipped).
namespace {
const int g_env_options_applied = [] {
    const char* env = std::getenv("PTB_OPTIONS");
    if (!env) return 0;
    const std::string text(env);
    size_t pos = 0;
    while (pos < text.size()) {
        size_t end = text.find(',', pos);
        if (end == std::string::npos) end = text.size();
        const std::string item = text.substr(pos, end - pos);
        const size_t eq = item.find('=');
        if (eq != std::string::npos) {
            const std::string name = item.substr(0, eq);
            if (ptb_set_option(name.c_str(), std::atoll(item.c_str() + eq + 1)) != PTB_OK)
                std::fprintf(stderr, "libptb: PTB_OPTIONS entry '%s' rejected: %s\n", item.c_str(), ptb_last_error());
        }
        pos = end + 1;
    }
    return 1;
}();
} // namespace

} // extern "C"

// The Lambda worker's entry (my_handler → worker::run, APP/main.cpp:9-31, APP/processors/worker/worker.cpp:25-105)
// without the S3 hops: the request is the worker_info JSON the preprocessor sends, the scene is read from a local
// mirror of s3://scene_bucket/scene_root.
namespace {

struct WorkerRequest {
    ptb::OwnedScene owned;
    uint32_t samples = 50, bounces = 10, X = 640, Y = 480;
    uint64_t seed = 0;
};

void parse_worker_request(const char* worker_info_json, const char* scene_dir, WorkerRequest& out) {
    if (!worker_info_json || !scene_dir) throw ptb::Error(PTB_E_INVALID, "worker_info_json or scene_dir is NULL");
    ptb::Json info;
    try {
        const std::string text(worker_info_json);
        info = ptb::JsonParser(text).parse();
    } catch (const std::exception& e) {
        throw ptb::Error(PTB_E_INVALID, std::string("worker_info: ") + e.what());
    }
    if (info.kind != ptb::Json::Object) throw ptb::Error(PTB_E_INVALID, "worker_info: not a JSON object");
    // models::worker_info (APP/models/work_info.hpp:17-31); the preprocessor omits samples/bounces/X/Y
    // (PRE/app.py:119-127), in which case the worker's own defaults apply (worker.hpp:20-24)
    ptb::WorkFilter work;
    if (const ptb::Json* si = info.find("scene_info"))
        if (const ptb::Json* w = si->find("work"))
            for (const auto& kv : w->obj) {
                std::vector<int>& v = work[kv.first];
                for (const ptb::Json& p : kv.second.arr) v.push_back(static_cast<int>(p.number(-1)));
            }
    out.samples = static_cast<uint32_t>(info.get("samples", 50.0));
    out.bounces = static_cast<uint32_t>(info.get("bounces", 10.0));
    out.X = static_cast<uint32_t>(static_cast<float>(info.get("X", 640.0)));
    out.Y = static_cast<uint32_t>(static_cast<float>(info.get("Y", 480.0)));
    if (!out.X || !out.Y || out.bounces > 255) throw ptb::Error(PTB_E_INVALID, "worker_info: bad X / Y / bounces");
    out.seed = std::hash<std::string>{}(info.get("worker_id", std::string("0")));
    const std::string gltf = std::string(scene_dir) + "/scene.gltf"; // worker::download_gltf_file: scene_root + "scene.gltf"
    if (!std::ifstream(gltf).good()) throw ptb::Error(PTB_E_IO, "cannot open " + gltf);
    ptb::load_gltf(gltf, 0, 0, out.owned, info.has("scene_info") ? &work : nullptr);
}

} // namespace

extern "C" {

ptb_status ptb_worker_run(const char* worker_info_json, const char* scene_dir, int device, const char* png_path,
                          uint8_t* rgba8_out, uint32_t* width_out, uint32_t* height_out, ptb_render_stats* stats_out) {
    return guarded([&] {
        WorkerRequest wr;
        parse_worker_request(worker_info_json, scene_dir, wr);
        const uint32_t X = wr.X, Y = wr.Y;
        ptb_scene* scene = ptb::create_scene(wr.owned.view(), device);
        try {
            ptb_tile_req req{};
            req.full_w = req.w = X;
            req.full_h = req.h = Y;
            req.spp = wr.samples;
            req.max_depth = wr.bounces;
            req.seed = wr.seed;
            req.integrator = PTB_INTEGRATOR_APP_RR;
            req.first_sample_unjittered = 1; // worker::generate_rays, worker.cpp:125-129
            std::vector<float> rgb(size_t(X) * Y * 3), alpha(size_t(X) * Y);
            ptb::render_tile_host(scene, req, rgb.data(), alpha.data(), stats_out);
            std::vector<uint8_t> rgba8(size_t(X) * Y * 4);
            ptb::tonemap_host(rgb.data(), alpha.data(), size_t(X) * Y, rgba8.data()); // worker::generate_final_image
            if (png_path) ptb::write_png_rgba8(png_path, rgba8.data(), X, Y);
            if (rgba8_out) std::memcpy(rgba8_out, rgba8.data(), rgba8.size());
            if (width_out) *width_out = X;
            if (height_out) *height_out = Y;
        } catch (...) {
            ptb::destroy_scene(scene);
            throw;
        }
        ptb::destroy_scene(scene);
    });
}

// worker::run on every GPU of the box: same request, the frame tile-sharded over the context's GPUs.
ptb_status ptb_worker_run_ctx(ptb_ctx* ctx, const char* worker_info_json, const char* scene_dir, const char* png_path,
                              uint8_t* rgba8_out, uint32_t* width_out, uint32_t* height_out, ptb_frame_stats* stats_out) {
    return guarded([&] {
        if (!ctx) throw ptb::Error(PTB_E_INVALID, "ctx is NULL");
        WorkerRequest wr;
        parse_worker_request(worker_info_json, scene_dir, wr);
        ptb::ctx_set_scene(ctx, wr.owned.view());
        ptb_frame_req fr{};
        fr.full_w = wr.X;
        fr.full_h = wr.Y;
        fr.spp = wr.samples;
        fr.max_depth = wr.bounces;
        fr.seed = wr.seed;
        fr.integrator = PTB_INTEGRATOR_APP_RR;
        fr.first_sample_unjittered = 1;
        fr.output = PTB_OUT_RGBA8; // worker::generate_final_image's tonemap + encode, on rank 0's GPU
        std::vector<uint8_t> rgba8(size_t(wr.X) * wr.Y * 4);
        ptb::ctx_render_frame(ctx, fr, rgba8.data(), stats_out);
        if (png_path) ptb::write_png_rgba8(png_path, rgba8.data(), wr.X, wr.Y);
        if (rgba8_out) std::memcpy(rgba8_out, rgba8.data(), rgba8.size());
        if (width_out) *width_out = wr.X;
        if (height_out) *height_out = wr.Y;
    });
}

// ---- multi-GPU --------------------------------------------------------------------------------------------------------

ptb_status ptb_scene_blob(const ptb_scene* scene, void** blob_dev, uint64_t* blob_bytes) {
    return guarded([&] {
        if (!scene || !blob_dev || !blob_bytes) throw ptb::Error(PTB_E_INVALID, "NULL argument");
        *blob_dev = scene->blob;
        *blob_bytes = scene->blob_bytes;
    });
}

ptb_status ptb_scene_export_header(const ptb_scene* scene, void* header, uint64_t capacity, uint64_t* n_bytes) {
    return guarded([&] {
        if (!scene || !n_bytes) throw ptb::Error(PTB_E_INVALID, "scene or n_bytes is NULL");
        *n_bytes = ptb::scene_header_bytes();
        if (header) {
            if (capacity < *n_bytes) throw ptb::Error(PTB_E_INVALID, "capacity too small");
            ptb::export_scene_header(scene, header);
        }
    });
}

ptb_status ptb_scene_import(const void* header, uint64_t n_bytes, int device, const void* src_blob_dev, int src_device,
                            ptb_scene** out) {
    return guarded([&] {
        if (!out) throw ptb::Error(PTB_E_INVALID, "out is NULL");
        *out = nullptr;
        *out = ptb::import_scene(header, n_bytes, device, src_blob_dev, src_device);
    });
}

ptb_status ptb_scene_clone(const ptb_scene* scene, int device, ptb_scene** out) {
    return guarded([&] {
        if (!out) throw ptb::Error(PTB_E_INVALID, "out is NULL");
        *out = nullptr;
        *out = ptb::clone_scene(scene, device);
    });
}

ptb_status ptb_frame_tiles(const ptb_frame_req* req, int world, uint32_t* xywh, uint64_t capacity, uint32_t* n_tiles) {
    return guarded([&] {
        if (!req) throw ptb::Error(PTB_E_INVALID, "req is NULL");
        ptb::frame_tiles(*req, world, xywh, capacity, n_tiles, false);
    });
}

ptb_status ptb_frame_tile_layout(const ptb_frame_req* req, int world, uint32_t* layout, uint64_t capacity, uint32_t* n_tiles) {
    return guarded([&] {
        if (!req) throw ptb::Error(PTB_E_INVALID, "req is NULL");
        ptb::frame_tiles(*req, world, layout, capacity, n_tiles, true);
    });
}

ptb_status ptb_group_create(const char* name, int rank, int world, int device, ptb_group** out) {
    return guarded([&] {
        if (!out) throw ptb::Error(PTB_E_INVALID, "out is NULL");
        *out = nullptr;
        *out = ptb::group_create_processes(name, rank, world, device);
    });
}

void ptb_group_destroy(ptb_group* group) {
    try {
        ptb::group_destroy(group);
    } catch (...) {
    }
}

ptb_status ptb_group_selftest_host(const char* name, int rank, int world, uint32_t n_tiles, uint32_t frames,
                                   uint32_t work_us, uint8_t* mine_out) {
    return guarded([&] { ptb::group_selftest_host(name, rank, world, n_tiles, frames, work_us, mine_out); });
}

ptb_status ptb_group_barrier(ptb_group* group) {
    return guarded([&] { ptb::group_barrier_public(group); });
}

ptb_status ptb_group_render_frame(ptb_group* group, const ptb_scene* scene, const ptb_frame_req* req, void* out_host,
                                  ptb_frame_stats* stats_out) {
    return guarded([&] {
        if (!req) throw ptb::Error(PTB_E_INVALID, "req is NULL");
        ptb::group_render_frame(group, scene, *req, out_host, stats_out);
    });
}

ptb_status ptb_ctx_create(int n_gpus, const int* devices, ptb_ctx** out) {
    return guarded([&] {
        if (!out) throw ptb::Error(PTB_E_INVALID, "out is NULL");
        *out = nullptr;
        *out = ptb::ctx_create(n_gpus, devices);
    });
}

void ptb_ctx_destroy(ptb_ctx* ctx) {
    try {
        ptb::ctx_destroy(ctx);
    } catch (...) {
    }
}

ptb_status ptb_ctx_set_scene(ptb_ctx* ctx, const ptb_scene_desc* desc) {
    return guarded([&] {
        if (!desc) throw ptb::Error(PTB_E_INVALID, "desc is NULL");
        ptb::ctx_set_scene(ctx, *desc);
    });
}

ptb_status ptb_ctx_load_gltf(ptb_ctx* ctx, const char* path, uint32_t camera_index, uint32_t sun_light_index) {
    return guarded([&] {
        if (!path) throw ptb::Error(PTB_E_INVALID, "path is NULL");
        ptb::OwnedScene owned;
        ptb::load_gltf(path, camera_index, sun_light_index, owned);
        ptb::ctx_set_scene(ctx, owned.view());
    });
}

const ptb_scene* ptb_ctx_scene(const ptb_ctx* ctx, int i) { return ptb::ctx_scene(ctx, i); }

ptb_status ptb_render_frame(ptb_ctx* ctx, const ptb_frame_req* req, void* out_host, ptb_frame_stats* stats_out) {
    return guarded([&] {
        if (!req) throw ptb::Error(PTB_E_INVALID, "req is NULL");
        ptb::ctx_render_frame(ctx, *req, out_host, stats_out);
    });
}

ptb_status ptb_host_alloc(uint64_t bytes, void** out) {
    return guarded([&] {
        if (!out) throw ptb::Error(PTB_E_INVALID, "out is NULL");
        *out = nullptr;
        PTB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    });
}

void ptb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

const char* ptb_last_error(void) { return tl_error.c_str(); }
int ptb_abi_version(void) { return PTB_ABI_VERSION; }
int ptb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
int ptb_extend_registers(void) {
#ifdef PTB_BUILD_EXPERIMENTS
    if (ptb::g_options.extend_variant == 3) return ptb::extend_coop_regs_per_thread();
    if (ptb::g_options.extend_variant == 4) return ptb::extend_ctx_regs_per_thread((int)ptb::g_options.extend_contexts);
    if (ptb::g_options.extend_variant == 0) return ptb::extend_regs_per_thread();
#endif
    return ptb::extend_lanes_regs_per_thread(ptb::g_options.extend_defer != 0, ptb::g_options.extend_dense != 0);
}
int ptb_shadow_registers(void) { return ptb::extend_anyhit_regs_per_thread(); }
uint64_t ptb_selftest_division(uint64_t n, uint64_t seed) { return ptb::division_selftest(n, seed); }

} // extern "C"
