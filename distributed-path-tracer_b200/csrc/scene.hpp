// scene.hpp — the host-side handle behind ptb_scene.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "device_scene.hpp"
#include "kd_build.hpp"
#include "ptb.h"

struct ptb_scene {
    int device = 0;
    ptb::DScene d{};            // device pointers + small by-value globals
    std::vector<void*> allocs;  // every cudaMalloc of this scene
    std::vector<ptb::KdTree> trees; // host copies, kept for ptb_scene_dump_kd
    ptb_scene_info info{};
    int sm_count = 148;
    bool has_pass_through = false; // some material can continue a path without using a bounce
};

namespace ptb {

// Owned, self-contained scene description (what the glTF loader produces).
struct OwnedMesh {
    std::vector<float> positions, normals, tangents, uvs;
    std::vector<uint32_t> indices;
};
struct OwnedTexture {
    std::vector<uint8_t> pixels;
    uint32_t width = 0, height = 0, channels = 0, is_float = 0, srgb = 0;
};
struct OwnedScene {
    std::vector<OwnedMesh> meshes;
    std::vector<ptb_surface_desc> surfaces;
    std::vector<ptb_instance_desc> instances;
    std::vector<ptb_material_desc> materials;
    std::vector<OwnedTexture> textures;
    ptb_camera_desc camera{};
    ptb_sun_desc sun{};
    float environment_factor[3] = {1, 1, 1};
    uint32_t transparent_background = 0;

    // views (valid while *this is alive and unchanged)
    std::vector<ptb_mesh_desc> mesh_views;
    std::vector<ptb_texture_desc> texture_views;
    ptb_scene_desc view();
};

// gltf.cpp — replaces renderer::load_gltf (LIB/core/renderer.cpp:61-331). Throws ptb::Error.
// `work` (may be null) is the worker's primitive assignment, mesh name → primitive indices to keep
// (distributed_scene::process_node, APP/scene/load_gltf.cpp:93-100): a mesh that is not named keeps nothing.
using WorkFilter = std::map<std::string, std::vector<int>>;
void load_gltf(const std::string& path, uint32_t camera_index, uint32_t sun_light_index, OwnedScene& out,
               const WorkFilter* work = nullptr);

// png.cpp
void write_png_rgba8(const std::string& path, const uint8_t* rgba8, uint32_t w, uint32_t h);
// decode an 8-bit PNG (grey, grey+alpha, RGB, RGBA, palette); throws on anything else
void read_png(const std::string& path, OwnedTexture& out);

// scene.cu
ptb_scene* create_scene(const ptb_scene_desc& desc, int device); // throws
void destroy_scene(ptb_scene* s);

} // namespace ptb
