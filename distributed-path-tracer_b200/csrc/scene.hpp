// scene.hpp — the host-side handle behind ptb_scene.
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "device_scene.hpp"
#include "kd_build.hpp"
#include "ptb.h"

constexpr int PTB_SCENE_ARRAYS = 15; // arrays inside the blob, in the order of DScene's pointer members

struct ptb_scene {
    int device = 0;
    ptb::DScene d{};            // device pointers (into the blob) + small by-value globals
    void* blob = nullptr;       // the ONE device allocation that holds every array of the scene
    uint64_t blob_bytes = 0;
    uint64_t offsets[PTB_SCENE_ARRAYS] = {}; // byte offset of each array inside the blob (256-byte aligned)
    bool replica = false;       // imported / cloned: no host trees
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
// replication: header (plain data) + blob (one device allocation)
uint64_t scene_header_bytes();
void export_scene_header(const ptb_scene* s, void* out);
ptb_scene* import_scene(const void* header, uint64_t n_bytes, int device, const void* src_blob_dev, int src_device);
ptb_scene* clone_scene(const ptb_scene* s, int device);

} // namespace ptb
