// scene.cu — scene construction: validates a ptb_scene_desc, rebuilds every
// mesh's KD-tree on the host (kd_build.cpp), flattens everything into the HBM
// layout of device_scene.hpp and uploads it.
//
// Replaces, for the hot path, what renderer::get_mesh / mesh::recalculate_aabb /
// mesh::build_kd_tree / model::recalculate_aabb set up as a shared_ptr object
// graph (LIB/core/renderer.cpp:177-263, LIB/core/mesh.cpp:254-298,
// LIB/scene/model.cpp:13-18; LIB = path-tracer-core/path_tracer_lib/path_tracer).
// Host float code here is compiled without FMA contraction.

#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "errors.hpp"
#include "kernels.hpp"
#include "scene.hpp"

namespace ptb {

namespace {

// The scene's arrays live back to back (256-byte aligned) in ONE device allocation, the blob: replicating a
// scene onto another GPU is a single device-to-device copy (or one broadcast), no rebuild and no fix-up
// beyond re-basing the fifteen pointers of DScene.
struct BlobPlan {
    struct Part {
        const void* src;
        size_t bytes;
    };
    std::vector<Part> parts;
    template <typename T>
    void add(const std::vector<T>& v) {
        parts.push_back(Part{v.data(), v.size() * sizeof(T)});
    }
};

void set_scene_pointers(ptb_scene* s) {
    char* b = static_cast<char*>(s->blob);
    const uint64_t* o = s->offsets;
    DScene& d = s->d;
    d.instances = reinterpret_cast<const DInstance*>(b + o[0]);
    d.inst_sphere = reinterpret_cast<const float4*>(b + o[1]);
    d.surfaces = reinterpret_cast<const DSurface*>(b + o[2]);
    d.meshes = reinterpret_cast<const DMesh*>(b + o[3]);
    d.kd_nodes = reinterpret_cast<const uint2*>(b + o[4]);
    d.kd_pairs = reinterpret_cast<const uint4*>(b + o[5]);
    d.kd_refs = reinterpret_cast<const uint32_t*>(b + o[6]);
    d.tri = reinterpret_cast<const float4*>(b + o[7]);
    d.vtx_pos = reinterpret_cast<const float*>(b + o[8]);
    d.vtx_nrm = reinterpret_cast<const float*>(b + o[9]);
    d.vtx_tan = reinterpret_cast<const float*>(b + o[10]);
    d.vtx_uv = reinterpret_cast<const float*>(b + o[11]);
    d.materials = reinterpret_cast<const DMaterial*>(b + o[12]);
    d.textures = reinterpret_cast<const DTexture*>(b + o[13]);
    d.texels = reinterpret_cast<const unsigned char*>(b + o[14]);
}

void upload_blob(ptb_scene* s, const BlobPlan& plan) {
    if (plan.parts.size() != size_t(PTB_SCENE_ARRAYS)) throw Error(PTB_E_INVALID, "internal: blob plan");
    uint64_t total = 0;
    for (int i = 0; i < PTB_SCENE_ARRAYS; i++) {
        s->offsets[i] = total;
        total += (std::max<uint64_t>(plan.parts[i].bytes, 256) + 255) & ~uint64_t(255);
    }
    PTB_CUDA(cudaMalloc(&s->blob, total));
    s->blob_bytes = total;
    s->info.device_bytes = total;
    for (int i = 0; i < PTB_SCENE_ARRAYS; i++)
        if (plan.parts[i].bytes)
            PTB_CUDA(cudaMemcpy(static_cast<char*>(s->blob) + s->offsets[i], plan.parts[i].src, plan.parts[i].bytes,
                                cudaMemcpyHostToDevice));
    set_scene_pointers(s);
}

constexpr uint32_t SCENE_HEADER_MAGIC = 0x42545053u; // "SPTB"
constexpr uint32_t SCENE_HEADER_VERSION = 1;

struct SceneHeader {
    uint32_t magic, version;
    uint64_t blob_bytes;
    uint64_t offsets[PTB_SCENE_ARRAYS];
    DScene d; // pointer members are meaningless outside the exporting process: re-based on import
    ptb_scene_info info;
    uint32_t has_pass_through, pad;
};

Xform xform_from(const float* origin, const float* basis) {
    Xform t;
    t.origin = V3{origin[0], origin[1], origin[2]};
    t.basis = M3{V3{basis[0], basis[1], basis[2]}, V3{basis[3], basis[4], basis[5]}, V3{basis[6], basis[7], basis[8]}};
    return t;
}

void require(bool ok, const char* msg) {
    if (!ok) throw Error(PTB_E_INVALID, msg);
}

// The traversal kernel's view of a tree: SIBLING PAIRS.  One 16-byte element holds the records of both
// children of a branch (left in .xy, right in .zw), so a step fetches both with a single aligned load that
// does not depend on the step's arithmetic, and the record of the far child can go onto the traversal
// stack (a pop needs no further load).  A record is
//   branch  x = split plane (float bits)      y = axis (0..2) | child pair index << 2   (index into kd_pairs)
//   leaf    x = first reference (index into kd_refs)   y = 3 | count << 2
//   absent  x = 0                             y = 3           (the reference's null child, mesh.cpp:372)
// Indices are ABSOLUTE (all meshes share the arrays), so the kernel carries no per-mesh base for them.
// The first pair of a mesh (DMesh::pair_base) carries the root in .xy.  Same tree, same order of children: nothing about the
// traversal's decisions changes.  Depth-first, left subtree first, so a descent to the left stays in a line.
// (Round 2 tried blocks of three pairs — a branch on an even level next to both of its children's pairs, so that two
// tree levels could be requested per round trip: the two-level step and the 50 % larger pair array cost 8 % on C2.)
constexpr uint32_t KD_ABSENT_Y = KD_LEAF_TAG;

void append_sibling_pairs(const KdTree& tree, uint32_t ref_base, std::vector<uint4>& pairs) {
    const size_t base = pairs.size();
    pairs.push_back(make_uint4(0, KD_ABSENT_Y, 0, KD_ABSENT_Y));
    if (tree.nodes.empty()) return;
    struct Item {
        uint32_t src;  // index in tree.nodes
        uint32_t pair; // destination pair (mesh-relative)
        uint32_t half; // 0: .xy, 1: .zw
    };
    std::vector<Item> todo;
    todo.push_back(Item{0, 0, 0});
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        const KdNode& n = tree.nodes[it.src];
        uint32_t x = n.w0, y = n.w1;
        if ((n.w1 & 3u) != KD_LEAF_TAG) {
            const uint32_t has_l = (n.w1 >> 2) & 1u, has_r = (n.w1 >> 3) & 1u, first = n.w1 >> 4;
            const uint32_t child_pair = static_cast<uint32_t>(pairs.size());
            pairs.push_back(make_uint4(0, KD_ABSENT_Y, 0, KD_ABSENT_Y));
            y = (n.w1 & 3u) | (child_pair << 2);
            const uint32_t rel = child_pair - static_cast<uint32_t>(base);
            if (has_r) todo.push_back(Item{first + has_l, rel, 1});
            if (has_l) todo.push_back(Item{first, rel, 0}); // popped next: left subtree first
        } else {
            x = n.w0 + ref_base;
        }
        uint4& dst = pairs[base + it.pair];
        if (it.half == 0) {
            dst.x = x;
            dst.y = y;
        } else {
            dst.z = x;
            dst.w = y;
        }
    }
}

bool finite3(const float* p, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (!std::isfinite(p[i])) return false;
    return true;
}

} // namespace

ptb_scene_desc OwnedScene::view() {
    mesh_views.resize(meshes.size());
    for (size_t i = 0; i < meshes.size(); i++) {
        mesh_views[i].positions = meshes[i].positions.data();
        mesh_views[i].normals = meshes[i].normals.data();
        mesh_views[i].tangents = meshes[i].tangents.data();
        mesh_views[i].uvs = meshes[i].uvs.data();
        mesh_views[i].n_vertices = static_cast<uint32_t>(meshes[i].positions.size() / 3);
        mesh_views[i].indices = meshes[i].indices.data();
        mesh_views[i].n_triangles = static_cast<uint32_t>(meshes[i].indices.size() / 3);
    }
    texture_views.resize(textures.size());
    for (size_t i = 0; i < textures.size(); i++) {
        texture_views[i].pixels = textures[i].pixels.data();
        texture_views[i].width = textures[i].width;
        texture_views[i].height = textures[i].height;
        texture_views[i].channels = textures[i].channels;
        texture_views[i].is_float = textures[i].is_float;
        texture_views[i].srgb = textures[i].srgb;
    }
    ptb_scene_desc d{};
    d.meshes = mesh_views.data();
    d.n_meshes = static_cast<uint32_t>(meshes.size());
    d.surfaces = surfaces.data();
    d.n_surfaces = static_cast<uint32_t>(surfaces.size());
    d.instances = instances.data();
    d.n_instances = static_cast<uint32_t>(instances.size());
    d.materials = materials.data();
    d.n_materials = static_cast<uint32_t>(materials.size());
    d.textures = texture_views.data();
    d.n_textures = static_cast<uint32_t>(textures.size());
    d.camera = camera;
    d.sun = sun;
    for (int i = 0; i < 3; i++) d.environment_factor[i] = environment_factor[i];
    d.transparent_background = transparent_background;
    d.kd_use_sah = 1;
    d.kd_max_depth = 25;
    return d;
}

void destroy_scene(ptb_scene* s) {
    if (!s) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(s->device);
    if (s->blob) cudaFree(s->blob);
    cudaSetDevice(prev);
    delete s;
}

ptb_scene* create_scene(const ptb_scene_desc& desc, int device) {
    // ---- validation (nothing below may read out of bounds) ----
    require(desc.n_meshes == 0 || desc.meshes, "meshes is NULL");
    require(desc.n_surfaces == 0 || desc.surfaces, "surfaces is NULL");
    require(desc.n_instances == 0 || desc.instances, "instances is NULL");
    require(desc.n_materials == 0 || desc.materials, "materials is NULL");
    require(desc.n_textures == 0 || desc.textures, "textures is NULL");
    require(desc.environment_tex_plus1 <= desc.n_textures, "environment texture id out of range");
    require(desc.n_instances < (1u << (32 - HIT_SURFACE_BITS)), "too many instances (max 2^20 - 1)");
    const uint32_t max_depth = desc.kd_max_depth ? desc.kd_max_depth : 25;
    require(max_depth <= 25, "kd_max_depth above 25 is not supported (traversal stack depth)");
    for (uint32_t m = 0; m < desc.n_meshes; m++) {
        const ptb_mesh_desc& md = desc.meshes[m];
        require(md.n_vertices == 0 || (md.positions && md.normals && md.tangents && md.uvs),
                "mesh attribute pointer is NULL");
        require(md.n_triangles == 0 || md.indices, "mesh indices is NULL");
        for (size_t i = 0; i < size_t(md.n_triangles) * 3; i++)
            require(md.indices[i] < md.n_vertices, "triangle index out of range");
        require(finite3(md.positions, size_t(md.n_vertices) * 3), "non-finite vertex position");
    }
    for (uint32_t s = 0; s < desc.n_surfaces; s++) {
        require(desc.surfaces[s].mesh < desc.n_meshes, "surface.mesh out of range");
        require(desc.surfaces[s].material < desc.n_materials, "surface.material out of range");
    }
    for (uint32_t i = 0; i < desc.n_instances; i++) {
        const ptb_instance_desc& id = desc.instances[i];
        require(uint64_t(id.first_surface) + id.n_surfaces <= desc.n_surfaces, "instance surface range out of bounds");
        require(id.n_surfaces < (1u << HIT_SURFACE_BITS), "too many surfaces in one instance (max 4095)");
    }
    for (uint32_t m = 0; m < desc.n_materials; m++) {
        const ptb_material_desc& md = desc.materials[m];
        const uint32_t slots[6] = {md.normal_tex, md.albedo_tex,   md.opacity_tex,
                                   md.roughness_tex, md.metallic_tex, md.emissive_tex};
        for (uint32_t t : slots) require(t == PTB_NO_TEXTURE || t < desc.n_textures, "material texture id out of range");
    }
    for (uint32_t t = 0; t < desc.n_textures; t++) {
        const ptb_texture_desc& td = desc.textures[t];
        require(td.pixels && td.width && td.height && td.channels >= 1 && td.channels <= 4, "bad texture");
    }

    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        throw Error(PTB_E_CUDA, "no CUDA device is usable; libptb has no CPU fallback");
    require(device >= 0 && device < n_dev, "device ordinal out of range");
    PTB_CUDA(cudaSetDevice(device));

    ptb_scene* s = new ptb_scene;
    try {
        s->device = device;
        cudaDeviceProp prop{};
        PTB_CUDA(cudaGetDeviceProperties(&prop, device));
        s->sm_count = prop.multiProcessorCount;

        // ---- meshes: AABB, KD tree, triangle records ----
        auto t0 = std::chrono::steady_clock::now();
        std::vector<DMesh> meshes(desc.n_meshes);
        std::vector<uint2> nodes;
        std::vector<uint4> pairs;
        std::vector<uint32_t> refs;
        std::vector<float4> tri;
        std::vector<float> vpos, vnrm, vtan, vuv;
        s->trees.resize(desc.n_meshes);
        for (uint32_t m = 0; m < desc.n_meshes; m++) {
            const ptb_mesh_desc& md = desc.meshes[m];
            const Aabb box = mesh_aabb(md.positions, md.n_vertices);
            KdTree& tree = s->trees[m];
            build_kd_tree_cached(md.positions, md.n_vertices, md.indices, md.n_triangles, box, desc.kd_use_sah != 0,
                                 max_depth, 0, tree);
            require(nodes.size() + tree.nodes.size() < (1ull << 32), "KD nodes exceed 32-bit indexing");
            require(refs.size() + tree.refs.size() < (1ull << 32), "KD leaf references exceed 32-bit indexing");
            require(tri.size() / 3 + md.n_triangles < (1ull << 30), "triangles exceed 30-bit indexing");
            DMesh& dm = meshes[m];
            for (int a = 0; a < 3; a++) {
                dm.aabb_min[a] = box.min[a];
                dm.aabb_max[a] = box.max[a];
            }
            dm.node_base = static_cast<uint32_t>(nodes.size());
            dm.ref_base = static_cast<uint32_t>(refs.size());
            dm.tri_base = static_cast<uint32_t>(tri.size() / 3);
            dm.vtx_base = static_cast<uint32_t>(vpos.size() / 3);
            dm.n_triangles = md.n_triangles;
            dm.pair_base = static_cast<uint32_t>(pairs.size());
            for (const KdNode& n : tree.nodes) nodes.push_back(make_uint2(n.w0, n.w1));
            append_sibling_pairs(tree, dm.ref_base, pairs);
            require(pairs.size() < (1ull << 29), "KD sibling pairs exceed 29-bit indexing");
            refs.insert(refs.end(), tree.refs.begin(), tree.refs.end());
            for (uint32_t t = 0; t < md.n_triangles; t++) {
                const uint32_t i0 = md.indices[3 * t], i1 = md.indices[3 * t + 1], i2 = md.indices[3 * t + 2];
                const V3 a{md.positions[3 * i0], md.positions[3 * i0 + 1], md.positions[3 * i0 + 2]};
                const V3 b{md.positions[3 * i1], md.positions[3 * i1 + 1], md.positions[3 * i1 + 2]};
                const V3 c{md.positions[3 * i2], md.positions[3 * i2 + 1], md.positions[3 * i2 + 2]};
                const V3 ab = a - b, ac = a - c; // the first two columns of triangle::intersect's matrix
                float w0, w1, w2;
                std::memcpy(&w0, &i0, 4);
                std::memcpy(&w1, &i1, 4);
                std::memcpy(&w2, &i2, 4);
                tri.push_back(make_float4(a.x, a.y, a.z, w0));
                tri.push_back(make_float4(ab.x, ab.y, ab.z, w1));
                tri.push_back(make_float4(ac.x, ac.y, ac.z, w2));
            }
            vpos.insert(vpos.end(), md.positions, md.positions + size_t(md.n_vertices) * 3);
            vnrm.insert(vnrm.end(), md.normals, md.normals + size_t(md.n_vertices) * 3);
            vtan.insert(vtan.end(), md.tangents, md.tangents + size_t(md.n_vertices) * 3);
            vuv.insert(vuv.end(), md.uvs, md.uvs + size_t(md.n_vertices) * 2);
            s->info.n_triangles += md.n_triangles;
            s->info.n_kd_nodes += tree.nodes.size();
            s->info.n_kd_branches += tree.n_branches;
            s->info.n_kd_leaves += tree.n_leaves;
            s->info.n_leaf_refs += tree.refs.size();
            s->info.kd_max_depth_reached = std::max(s->info.kd_max_depth_reached, tree.max_depth_reached);
        }
        s->info.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        // ---- instances ----
        std::vector<DInstance> instances(desc.n_instances);
        std::vector<float4> spheres(desc.n_instances);
        for (uint32_t i = 0; i < desc.n_instances; i++) {
            const ptb_instance_desc& id = desc.instances[i];
            DInstance& di = instances[i];
            di.fwd = xform_from(id.origin, id.basis);
            di.inv = inverse(di.fwd);                         // transform::inverse, per ray in the reference
            di.normal_mat = transpose(inverse(di.fwd.basis)); // renderer.cpp:698
            // model::recalculate_aabb: clear() then add(min), add(max) of every surface's mesh box
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {FLT_MIN, FLT_MIN, FLT_MIN};
            for (uint32_t k = 0; k < id.n_surfaces; k++) {
                const DMesh& dm = meshes[desc.surfaces[id.first_surface + k].mesh];
                for (int a = 0; a < 3; a++) {
                    mn[a] = rmin(mn[a], dm.aabb_min[a]);
                    mx[a] = rmax(mx[a], dm.aabb_min[a]);
                    mn[a] = rmin(mn[a], dm.aabb_max[a]);
                    mx[a] = rmax(mx[a], dm.aabb_max[a]);
                }
            }
            for (int a = 0; a < 3; a++) {
                di.aabb_min[a] = mn[a];
                di.aabb_max[a] = mx[a];
            }
            di.first_surface = id.first_surface;
            di.n_surfaces = id.n_surfaces;
            di.same_box = 0;
            if (id.n_surfaces == 1) {
                const DMesh& only = meshes[desc.surfaces[id.first_surface].mesh];
                di.same_box = std::memcmp(only.aabb_min, di.aabb_min, sizeof(di.aabb_min)) == 0 &&
                              std::memcmp(only.aabb_max, di.aabb_max, sizeof(di.aabb_max)) == 0;
            }
            // Conservative world-space bounding sphere of the model box (double precision, 2 % + absolute slack):
            // a regular ray that misses it misses the box in local space by far more than any rounding of the
            // reference's slab test, so extend may skip the instance without changing a single result.
            if (mn[0] > mx[0] || mn[1] > mx[1] || mn[2] > mx[2]) {
                spheres[i] = make_float4(0, 0, 0, -1.0f); // aabb::intersect rejects min > max outright
            } else {
                double c[3], r2 = 0;
                for (int a = 0; a < 3; a++) c[a] = 0.5 * (double(mn[a]) + double(mx[a]));
                const M3& B = di.fwd.basis;
                const double bx[3] = {B.x.x, B.x.y, B.x.z}, by[3] = {B.y.x, B.y.y, B.y.z}, bz[3] = {B.z.x, B.z.y, B.z.z};
                double cw[3];
                const double o3[3] = {di.fwd.origin.x, di.fwd.origin.y, di.fwd.origin.z};
                for (int a = 0; a < 3; a++) cw[a] = bx[a] * c[0] + by[a] * c[1] + bz[a] * c[2] + o3[a];
                for (int sx = -1; sx <= 1; sx += 2)
                    for (int sy = -1; sy <= 1; sy += 2)
                        for (int sz = -1; sz <= 1; sz += 2) {
                            const double h[3] = {sx * 0.5 * (double(mx[0]) - mn[0]), sy * 0.5 * (double(mx[1]) - mn[1]),
                                                 sz * 0.5 * (double(mx[2]) - mn[2])};
                            double d2 = 0;
                            for (int a = 0; a < 3; a++) {
                                const double v = bx[a] * h[0] + by[a] * h[1] + bz[a] * h[2];
                                d2 += v * v;
                            }
                            r2 = std::max(r2, d2);
                        }
                const double scale = std::sqrt(cw[0] * cw[0] + cw[1] * cw[1] + cw[2] * cw[2]) + std::sqrt(r2);
                const double r = std::sqrt(r2) * 1.02 + 1e-3 * scale + 1e-4;
                spheres[i] = make_float4(float(cw[0]), float(cw[1]), float(cw[2]),
                                         std::isfinite(r) ? float(r) : 3.0e38f);
            }
        }
        std::vector<DSurface> surfaces(desc.n_surfaces);
        for (uint32_t k = 0; k < desc.n_surfaces; k++) surfaces[k] = DSurface{desc.surfaces[k].mesh, desc.surfaces[k].material};

        // ---- materials and textures ----
        std::vector<DMaterial> materials(desc.n_materials);
        for (uint32_t m = 0; m < desc.n_materials; m++) {
            const ptb_material_desc& md = desc.materials[m];
            DMaterial& dm = materials[m];
            std::memset(&dm, 0, sizeof(dm));
            for (int c = 0; c < 3; c++) {
                dm.albedo[c] = md.albedo[c];
                dm.emissive[c] = md.emissive[c];
            }
            dm.opacity = md.opacity;
            dm.roughness = md.roughness;
            dm.metallic = md.metallic;
            dm.ior = md.ior;
            dm.shadow_catcher = md.shadow_catcher;
            dm.normal_tex = md.normal_tex;
            dm.albedo_tex = md.albedo_tex;
            dm.opacity_tex = md.opacity_tex;
            dm.roughness_tex = md.roughness_tex;
            dm.metallic_tex = md.metallic_tex;
            dm.emissive_tex = md.emissive_tex;
            if (md.shadow_catcher || md.opacity_tex != PTB_NO_TEXTURE ||
                !(md.opacity == 1.0f || std::fabs(md.opacity - 1.0f) < kEpsilon))
                s->has_pass_through = true;
            dm.any_tex = (md.normal_tex & md.albedo_tex & md.opacity_tex & md.roughness_tex & md.metallic_tex &
                          md.emissive_tex) != PTB_NO_TEXTURE;
        }
        std::vector<DTexture> textures(desc.n_textures);
        std::vector<unsigned char> texels;
        for (uint32_t t = 0; t < desc.n_textures; t++) {
            const ptb_texture_desc& td = desc.textures[t];
            const size_t bytes = size_t(td.width) * td.height * td.channels * (td.is_float ? 4 : 1);
            while (texels.size() % 16) texels.push_back(0);
            textures[t] = DTexture{(unsigned long long)texels.size(), td.width, td.height, td.channels, td.is_float, td.srgb, 0};
            const unsigned char* src = static_cast<const unsigned char*>(td.pixels);
            texels.insert(texels.end(), src, src + bytes);
        }

        // ---- upload ----
        auto t1 = std::chrono::steady_clock::now();
        DScene& d = s->d;
        BlobPlan plan; // order = DScene's pointer members = set_scene_pointers
        plan.add(instances);
        plan.add(spheres);
        plan.add(surfaces);
        plan.add(meshes);
        plan.add(nodes);
        plan.add(pairs);
        plan.add(refs);
        plan.add(tri);
        plan.add(vpos);
        plan.add(vnrm);
        plan.add(vtan);
        plan.add(vuv);
        plan.add(materials);
        plan.add(textures);
        plan.add(texels);
        upload_blob(s, plan);
        d.n_instances = desc.n_instances;
        d.n_pairs = static_cast<uint32_t>(pairs.size());
        d.n_refs = static_cast<uint32_t>(refs.size());
        d.n_tris = static_cast<uint32_t>(tri.size() / 3);
        d.camera.xf = xform_from(desc.camera.origin, desc.camera.basis);
        d.camera.tan_half_fov = std::tan(desc.camera.yfov * 0.5F); // camera::set_fov (glibc tanf, as the reference)
        d.sun.enabled = desc.sun.enabled ? 1 : 0;
        if (desc.sun.enabled) {
            const float zero[3] = {0, 0, 0};
            const Xform sx = xform_from(zero, desc.sun.basis);
            d.sun.direction = mul(sx.basis, V3{0.0f, 0.0f, 1.0f}); // basis * fvec3::backward
            d.sun.energy = V3{desc.sun.energy[0], desc.sun.energy[1], desc.sun.energy[2]};
            d.sun.angular_radius = desc.sun.angular_radius;
        } else {
            d.sun.direction = V3{0, 0, 1};
            d.sun.energy = V3{0, 0, 0};
            d.sun.angular_radius = 0;
        }
        d.environment = V3{desc.environment_factor[0], desc.environment_factor[1], desc.environment_factor[2]};
        d.environment_tex = desc.environment_tex_plus1 ? desc.environment_tex_plus1 - 1 : 0xFFFFFFFFu;
        d.transparent_background = desc.transparent_background ? 1 : 0;
        PTB_CUDA(cudaDeviceSynchronize());
        s->info.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        s->info.n_instances = desc.n_instances;
        s->info.n_surfaces = desc.n_surfaces;
        s->info.n_meshes = desc.n_meshes;
        s->info.n_materials = desc.n_materials;
        s->info.n_textures = desc.n_textures;
    } catch (...) {
        destroy_scene(s);
        throw;
    }
    return s;
}

// ---- replication ------------------------------------------------------------------------------------------------

uint64_t scene_header_bytes() { return sizeof(SceneHeader); }

void export_scene_header(const ptb_scene* s, void* out) {
    SceneHeader h{};
    h.magic = SCENE_HEADER_MAGIC;
    h.version = SCENE_HEADER_VERSION;
    h.blob_bytes = s->blob_bytes;
    for (int i = 0; i < PTB_SCENE_ARRAYS; i++) h.offsets[i] = s->offsets[i];
    h.d = s->d;
    h.info = s->info;
    h.has_pass_through = s->has_pass_through ? 1 : 0;
    std::memcpy(out, &h, sizeof(h));
}

ptb_scene* import_scene(const void* header, uint64_t n_bytes, int device, const void* src_blob_dev, int src_device) {
    require(header && n_bytes == sizeof(SceneHeader), "scene header: wrong size");
    SceneHeader h;
    std::memcpy(&h, header, sizeof(h));
    require(h.magic == SCENE_HEADER_MAGIC && h.version == SCENE_HEADER_VERSION, "scene header: bad magic / version");
    uint64_t prev_end = 0;
    for (int i = 0; i < PTB_SCENE_ARRAYS; i++) {
        require(h.offsets[i] >= prev_end && h.offsets[i] % 256 == 0 && h.offsets[i] < h.blob_bytes, "scene header: bad offsets");
        prev_end = h.offsets[i];
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        throw Error(PTB_E_CUDA, "no CUDA device is usable; libptb has no CPU fallback");
    require(device >= 0 && device < n_dev, "device ordinal out of range");
    PTB_CUDA(cudaSetDevice(device));
    ptb_scene* s = new ptb_scene;
    try {
        auto t0 = std::chrono::steady_clock::now();
        s->device = device;
        s->replica = true;
        cudaDeviceProp prop{};
        PTB_CUDA(cudaGetDeviceProperties(&prop, device));
        s->sm_count = prop.multiProcessorCount;
        s->d = h.d;
        s->info = h.info;
        s->has_pass_through = h.has_pass_through != 0;
        s->blob_bytes = h.blob_bytes;
        for (int i = 0; i < PTB_SCENE_ARRAYS; i++) s->offsets[i] = h.offsets[i];
        PTB_CUDA(cudaMalloc(&s->blob, s->blob_bytes));
        set_scene_pointers(s);
        if (src_blob_dev) {
            if (src_device == device)
                PTB_CUDA(cudaMemcpy(s->blob, src_blob_dev, s->blob_bytes, cudaMemcpyDeviceToDevice));
            else
                PTB_CUDA(cudaMemcpyPeer(s->blob, device, src_blob_dev, src_device, s->blob_bytes));
            PTB_CUDA(cudaDeviceSynchronize());
        }
        s->info.build_seconds = 0; // nothing was built here
        s->info.upload_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } catch (...) {
        destroy_scene(s);
        throw;
    }
    return s;
}

ptb_scene* clone_scene(const ptb_scene* s, int device) {
    require(s != nullptr, "scene is NULL");
    SceneHeader h;
    export_scene_header(s, &h);
    return import_scene(&h, sizeof(h), device, s->blob, s->device);
}

} // namespace ptb
