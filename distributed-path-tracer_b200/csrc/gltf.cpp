// gltf.cpp — glTF 2.0 → flat scene description.
//
// Host-side replacement of renderer::load_gltf / process_node / get_mesh /
// get_material (LIB/core/renderer.cpp:61-331; LIB =
// path-tracer-core/path_tracer_lib/path_tracer), written against the glTF
// specification with its own JSON reader instead of cgltf.  What it must
// reproduce is the reference's observable behaviour, quirks included, because
// bit-exact hit ids need the same vertices, transforms and visiting order:
//   * node transforms come from translation/rotation/scale only (:113-128);
//     a `matrix` property is ignored
//   * entity name = camera name / light name / node name (:106-111); root
//     entities live in an unordered_map keyed by name (:171), so duplicates
//     overwrite and the visiting order of renderer::intersect (:646-671) is the
//     reverse of that map's iteration order, children after their parent
//   * camera and sun light are attached by NAME equality (:145,154)
//   * TANGENT (VEC4) accessors are unpacked into a count*3 buffer and read with
//     stride 3 (:215-218,248-250 with cgltf_accessor_unpack_floats' element
//     rule), so tangents are scrambled but deterministic
//   * every primitive gets its own material object (:265-331); textures are
//     cached by path only, so the first user decides the sRGB flag (:33-51)
// Compiled with -ffp-contract=off.

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <unordered_map>

#include "errors.hpp"
#include "json.hpp"
#include "scene.hpp"
#include "vecmath.hpp"

namespace ptb {

namespace {

std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(PTB_E_IO, "cannot open " + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string dir_of(const std::string& path) {
    const size_t p = path.find_last_of('/');
    return p == std::string::npos ? std::string(".") : path.substr(0, p);
}

std::string replace_all(std::string s, const std::string& from, const std::string& to) {
    size_t p = 0;
    while ((p = s.find(from, p)) != std::string::npos) {
        s.replace(p, from.size(), to);
        p += to.size();
    }
    return s;
}

int components_of(const std::string& type) {
    if (type == "SCALAR") return 1;
    if (type == "VEC2") return 2;
    if (type == "VEC3") return 3;
    if (type == "VEC4") return 4;
    if (type == "MAT2") return 4;
    if (type == "MAT3") return 9;
    if (type == "MAT4") return 16;
    throw Error(PTB_E_IO, "glTF: unknown accessor type " + type);
}

int component_size(int ct) {
    switch (ct) {
    case 5120: case 5121: return 1;
    case 5122: case 5123: return 2;
    case 5125: case 5126: return 4;
    }
    throw Error(PTB_E_IO, "glTF: unknown componentType");
}

struct Gltf {
    Json root;
    std::string dir;
    std::vector<std::string> buffers; // loaded lazily
    std::vector<bool> buffer_loaded;

    const Json& arr(const char* key) const {
        static const Json empty = [] {
            Json j;
            j.kind = Json::Array;
            return j;
        }();
        const Json* j = root.find(key);
        return j ? *j : empty;
    }

    const std::string& buffer(size_t idx) {
        if (buffers.size() <= idx) {
            buffers.resize(arr("buffers").size());
            buffer_loaded.resize(buffers.size(), false);
        }
        if (idx >= buffers.size()) throw Error(PTB_E_IO, "glTF: buffer index out of range");
        if (!buffer_loaded[idx]) {
            const Json& b = arr("buffers").at(idx);
            const std::string uri = b.get("uri", "");
            if (uri.empty()) throw Error(PTB_E_IO, "glTF: buffer without uri (GLB is not supported)");
            if (uri.rfind("data:", 0) == 0) throw Error(PTB_E_IO, "glTF: data: URIs are not supported");
            buffers[idx] = read_file(dir + "/" + replace_all(uri, "%20", " "));
            buffer_loaded[idx] = true;
        }
        return buffers[idx];
    }

    struct View {
        const unsigned char* data;
        size_t size, stride;
        int ncomp, ctype;
        bool normalized;
        size_t count;
    };

    // A JSON number that must be a byte count / element count: non-negative, integral, below 2^32 (a malformed or
    // hostile file must produce an error, not undefined behaviour in the double -> size_t cast or a wrapped sum).
    static size_t checked_size(const Json& o, const char* key, double fallback) {
        const double v = o.get(key, fallback);
        if (!(v >= 0.0) || v >= 4294967296.0 || v != std::floor(v)) throw Error(PTB_E_IO, std::string("glTF: bad ") + key);
        return static_cast<size_t>(v);
    }

    View accessor(long long idx) {
        if (idx < 0 || static_cast<size_t>(idx) >= arr("accessors").size()) throw Error(PTB_E_IO, "glTF: accessor index out of range");
        const Json& a = arr("accessors").at(static_cast<size_t>(idx));
        View v{};
        v.ncomp = components_of(a.get("type", "SCALAR"));
        v.ctype = static_cast<int>(a.get("componentType", 5126));
        v.normalized = a.has("normalized") && a.at("normalized").b;
        v.count = checked_size(a, "count", 0);
        const long long bv = a.index("bufferView");
        if (bv < 0) {
            v.data = nullptr;
            return v;
        }
        if (static_cast<size_t>(bv) >= arr("bufferViews").size()) throw Error(PTB_E_IO, "glTF: bufferView index out of range");
        const Json& view = arr("bufferViews").at(static_cast<size_t>(bv));
        const std::string& buf = buffer(checked_size(view, "buffer", 0));
        const size_t view_off = checked_size(view, "byteOffset", 0), acc_off = checked_size(a, "byteOffset", 0);
        const size_t elem = static_cast<size_t>(v.ncomp) * component_size(v.ctype);
        if (elem == 0) throw Error(PTB_E_IO, "glTF: bad accessor type / componentType");
        v.stride = checked_size(view, "byteStride", 0);
        if (v.stride == 0) v.stride = elem;
        // the bytes the accessor touches, relative to the view: every term is below 2^32, so 64-bit sums cannot wrap
        const uint64_t span = v.count ? uint64_t(acc_off) + uint64_t(v.count - 1) * v.stride + elem : 0;
        const uint64_t view_len = view.has("byteLength") ? checked_size(view, "byteLength", 0) : uint64_t(buf.size());
        if (span > view_len || uint64_t(view_off) + span > buf.size() || uint64_t(view_off) + view_len > buf.size())
            throw Error(PTB_E_IO, "glTF: accessor reads past the end of its buffer view");
        v.data = reinterpret_cast<const unsigned char*>(buf.data()) + view_off + acc_off;
        v.size = elem;
        return v;
    }
};

float read_component(const unsigned char* p, int ctype, bool normalized) {
    switch (ctype) {
    case 5126: { float f; std::memcpy(&f, p, 4); return f; }
    case 5125: { uint32_t u; std::memcpy(&u, p, 4); return normalized ? 0.0f : static_cast<float>(u); }
    case 5123: { uint16_t u; std::memcpy(&u, p, 2); return normalized ? u / 65535.0f : static_cast<float>(u); }
    case 5122: { int16_t u; std::memcpy(&u, p, 2); return normalized ? u / 32767.0f : static_cast<float>(u); }
    case 5121: { uint8_t u = *p; return normalized ? u / 255.0f : static_cast<float>(u); }
    case 5120: { int8_t u = static_cast<int8_t>(*p); return normalized ? u / 127.0f : static_cast<float>(u); }
    }
    return 0.0f;
}

// cgltf_accessor_unpack_floats(accessor, out, requested): whole elements only.
void unpack_floats(const Gltf::View& v, std::vector<float>& out, size_t requested) {
    out.assign(requested, 0.0f);
    const size_t available = v.count * static_cast<size_t>(v.ncomp);
    const size_t float_count = std::min(available, requested);
    const size_t elements = float_count / static_cast<size_t>(v.ncomp);
    if (!v.data) return;
    const int cs = component_size(v.ctype);
    for (size_t e = 0; e < elements; e++)
        for (int c = 0; c < v.ncomp; c++)
            out[e * v.ncomp + c] = read_component(v.data + e * v.stride + size_t(c) * cs, v.ctype, v.normalized);
}

struct Entity {
    std::string name;
    Xform local;
    int parent = -1;
    std::vector<int> children;
    int mesh = -1; // glTF mesh index
    bool has_model = false;
    std::vector<ptb_surface_desc> surfaces; // filled while the node is processed, like the reference
};

// quat::to_basis (LIB/math/quat.cpp:95-113) then scale (LIB/scene/transform.cpp:21-31)
M3 basis_from(float w, float x, float y, float z, V3 scale) {
    M3 b{V3{1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w)},
         V3{2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w)},
         V3{2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)}};
    b.x = b.x * scale.x;
    b.y = b.y * scale.y;
    b.z = b.z * scale.z;
    return b;
}

float fnum(const Json& j, size_t i, float fallback) {
    return (j.kind == Json::Array && i < j.arr.size()) ? static_cast<float>(j.arr[i].number(fallback)) : fallback;
}

} // namespace

void load_gltf(const std::string& path, uint32_t camera_index, uint32_t sun_light_index, OwnedScene& out,
               const WorkFilter* work) {
    Gltf g;
    g.dir = dir_of(path);
    try {
        const std::string text = read_file(path);
        g.root = JsonParser(text).parse();
    } catch (const Error&) {
        throw;
    } catch (const std::exception& e) {
        throw Error(PTB_E_IO, std::string("glTF: ") + e.what());
    }
    try {
        const Json& cameras = g.arr("cameras");
        if (cameras.size() < size_t(camera_index) + 1)
            throw Error(PTB_E_IO, "Scene does not contain camera #" + std::to_string(camera_index) + ".");
        const Json& cam = cameras.at(camera_index);
        const std::string camera_name = cam.get("name", "");

        const Json* lights = nullptr;
        if (const Json* ext = g.root.find("extensions"))
            if (const Json* lp = ext->find("KHR_lights_punctual")) lights = lp->find("lights");
        const Json* sun = nullptr;
        if (sun_light_index != 0xFFFFFFFFu && lights && lights->size() >= size_t(sun_light_index) + 1 &&
            lights->at(sun_light_index).get("type", "") == "directional")
            sun = &lights->at(sun_light_index);
        const std::string sun_name = sun ? sun->get("name", "") : std::string();

        const Json& nodes = g.arr("nodes");
        const Json& scenes = g.arr("scenes");
        if (scenes.size() == 0) throw Error(PTB_E_IO, "glTF: no scenes");

        // ---- meshes / materials / textures ----
        const Json& gl_meshes = g.arr("meshes");
        const Json& gl_materials = g.arr("materials");
        const Json& gl_textures = g.arr("textures");
        const Json& gl_images = g.arr("images");
        std::unordered_map<std::string, uint32_t> texture_cache; // path → texture id (srgb of the first load)
        std::unordered_map<uint64_t, uint32_t> mesh_cache;      // (gltf mesh, primitive) → mesh id

        auto texture_for = [&](const Json& mat, const char* slot, const Json* parent_obj, bool srgb) -> uint32_t {
            const Json* holder = parent_obj ? parent_obj->find(slot) : mat.find(slot);
            if (!holder) return PTB_NO_TEXTURE;
            const long long ti = holder->index("index");
            if (ti < 0) return PTB_NO_TEXTURE;
            const long long src = gl_textures.at(static_cast<size_t>(ti)).index("source");
            if (src < 0) return PTB_NO_TEXTURE;
            const std::string uri = gl_images.at(static_cast<size_t>(src)).get("uri", "");
            if (uri.empty()) return PTB_NO_TEXTURE;
            const std::string full = replace_all(g.dir + "/" + uri, "%20", " ");
            auto it = texture_cache.find(full);
            if (it != texture_cache.end()) return it->second;
            OwnedTexture t;
            read_png(full, t);
            t.srgb = srgb ? 1 : 0;
            const uint32_t id = static_cast<uint32_t>(out.textures.size());
            out.textures.push_back(std::move(t));
            texture_cache[full] = id;
            return id;
        };

        auto make_material = [&](const Json& prim) -> uint32_t { // renderer::get_material
            ptb_material_desc m{};
            m.albedo[0] = m.albedo[1] = m.albedo[2] = 1;
            m.opacity = 1;
            m.roughness = 1;
            m.metallic = 1;
            m.emissive[0] = m.emissive[1] = m.emissive[2] = 1; // material.hpp:15 — only seen without a glTF material
            m.ior = 1.33F;
            m.normal_tex = m.albedo_tex = m.opacity_tex = m.roughness_tex = m.metallic_tex = m.emissive_tex =
                PTB_NO_TEXTURE;
            const long long mi = prim.index("material");
            if (mi >= 0) {
                const Json& gm = gl_materials.at(static_cast<size_t>(mi));
                const Json* pbr = gm.find("pbrMetallicRoughness");
                float base[4] = {1, 1, 1, 1};
                float rough = 1, metal = 1;
                if (pbr) {
                    if (const Json* bc = pbr->find("baseColorFactor"))
                        for (int c = 0; c < 4; c++) base[c] = fnum(*bc, c, 1);
                    rough = static_cast<float>(pbr->get("roughnessFactor", 1.0));
                    metal = static_cast<float>(pbr->get("metallicFactor", 1.0));
                }
                float em[3] = {0, 0, 0};
                if (const Json* ef = gm.find("emissiveFactor"))
                    for (int c = 0; c < 3; c++) em[c] = fnum(*ef, c, 0);
                for (int c = 0; c < 3; c++) {
                    m.albedo[c] = base[c];
                    m.emissive[c] = em[c];
                }
                m.opacity = base[3];
                m.roughness = rough;
                m.metallic = metal;
                const bool opaque = gm.get("alphaMode", "OPAQUE") == "OPAQUE";
                // load order as in the reference: normal, albedo(+opacity), occlusion, roughness/metallic, emissive
                m.normal_tex = texture_for(gm, "normalTexture", nullptr, false);
                m.albedo_tex = texture_for(gm, "baseColorTexture", pbr, true);
                if (m.albedo_tex != PTB_NO_TEXTURE && !opaque) m.opacity_tex = m.albedo_tex;
                (void)texture_for(gm, "occlusionTexture", nullptr, false); // loaded (affects the cache), never sampled
                const uint32_t rm = texture_for(gm, "metallicRoughnessTexture", pbr, false);
                m.roughness_tex = rm;
                m.metallic_tex = rm;
                m.emissive_tex = texture_for(gm, "emissiveTexture", nullptr, true);
                const std::string name = gm.get("name", "");
                if (name.find("shadow") != std::string::npos && name.find("catcher") != std::string::npos)
                    m.shadow_catcher = 1;
            }
            out.materials.push_back(m);
            return static_cast<uint32_t>(out.materials.size() - 1);
        };

        auto make_mesh = [&](size_t mesh_idx, size_t prim_idx, const Json& prim) -> uint32_t { // renderer::get_mesh
            const uint64_t key = (uint64_t(mesh_idx) << 32) | prim_idx;
            auto it = mesh_cache.find(key);
            if (it != mesh_cache.end()) return it->second; // same data → same tree; the reference rebuilds it
            OwnedMesh m;
            size_t vertex_count = 0;
            const Json& attrs = prim.at("attributes");
            for (const auto& kv : attrs.obj) {
                const Gltf::View v = g.accessor(static_cast<long long>(kv.second.number(-1)));
                if (kv.first == "POSITION") {
                    unpack_floats(v, m.positions, v.count * 3);
                    vertex_count = v.count;
                } else if (kv.first.rfind("TEXCOORD", 0) == 0) {
                    unpack_floats(v, m.uvs, v.count * 2);
                } else if (kv.first == "NORMAL") {
                    unpack_floats(v, m.normals, v.count * 3);
                } else if (kv.first == "TANGENT") {
                    unpack_floats(v, m.tangents, v.count * 3);
                }
            }
            m.positions.resize(vertex_count * 3, 0.0f);
            m.normals.resize(vertex_count * 3, 0.0f); // the reference reads out of bounds when one is absent
            m.tangents.resize(vertex_count * 3, 0.0f);
            m.uvs.resize(vertex_count * 2, 0.0f);
            const long long ii = prim.index("indices");
            if (ii < 0) throw Error(PTB_E_IO, "glTF: non-indexed primitives are not supported (nor by the reference)");
            const Gltf::View iv = g.accessor(ii);
            if (!iv.data) throw Error(PTB_E_IO, "glTF: index accessor without a bufferView");
            m.indices.resize((iv.count / 3) * 3);
            for (size_t i = 0; i < m.indices.size(); i++) {
                const unsigned char* p = iv.data + i * iv.stride;
                uint32_t idx = 0;
                if (iv.ctype == 5125) std::memcpy(&idx, p, 4);
                else if (iv.ctype == 5123) { uint16_t u; std::memcpy(&u, p, 2); idx = u; }
                else if (iv.ctype == 5121) idx = *p;
                else throw Error(PTB_E_IO, "glTF: bad index component type");
                if (idx >= vertex_count) throw Error(PTB_E_IO, "glTF: vertex index out of range");
                m.indices[i] = idx;
            }
            out.meshes.push_back(std::move(m));
            const uint32_t id = static_cast<uint32_t>(out.meshes.size() - 1);
            mesh_cache[key] = id;
            return id;
        };

        // ---- entity tree (process_node) ----
        std::vector<Entity> ents;
        std::unordered_map<std::string, int> roots; // renderer::entities
        int camera_entity = -1, sun_entity = -1;
        std::function<int(size_t, int, int)> process = [&](size_t node_idx, int parent, int depth) -> int {
            if (depth > 256) throw Error(PTB_E_IO, "glTF: node hierarchy too deep");
            const Json& n = nodes.at(node_idx);
            Entity e;
            long long light_idx = -1;
            if (const Json* ext = n.find("extensions"))
                if (const Json* lp = ext->find("KHR_lights_punctual")) light_idx = lp->index("light");
            if (n.index("camera") >= 0)
                e.name = cameras.at(static_cast<size_t>(n.index("camera"))).get("name", "");
            else if (light_idx >= 0 && lights)
                e.name = lights->at(static_cast<size_t>(light_idx)).get("name", "");
            else
                e.name = n.get("name", "");
            float qw = 0, qx = 0, qy = 0, qz = 0; // math::quat() is all zeros; to_basis still yields identity
            if (const Json* r = n.find("rotation")) {
                qx = fnum(*r, 0, 0); qy = fnum(*r, 1, 0); qz = fnum(*r, 2, 0); qw = fnum(*r, 3, 1);
            }
            V3 scale{1, 1, 1}, translation{0, 0, 0};
            if (const Json* s = n.find("scale")) scale = V3{fnum(*s, 0, 1), fnum(*s, 1, 1), fnum(*s, 2, 1)};
            if (const Json* t = n.find("translation")) translation = V3{fnum(*t, 0, 0), fnum(*t, 1, 0), fnum(*t, 2, 0)};
            e.local = Xform{translation, basis_from(qw, qx, qy, qz, scale)};
            e.parent = parent;
            e.mesh = static_cast<int>(n.index("mesh"));
            e.has_model = e.mesh >= 0;
            if (e.has_model) { // renderer.cpp:132-143 — primitives before children
                const Json& gmesh = gl_meshes.at(static_cast<size_t>(e.mesh));
                const Json& prims = gmesh.at("primitives");
                const std::vector<int>* keep = nullptr;
                static const std::vector<int> nothing;
                if (work) { // APP/scene/load_gltf.cpp:93-100
                    auto it = work->find(gmesh.get("name", ""));
                    keep = it == work->end() ? &nothing : &it->second;
                }
                for (size_t p = 0; p < prims.size(); p++) {
                    if (keep && std::find(keep->begin(), keep->end(), static_cast<int>(p)) == keep->end()) continue;
                    ptb_surface_desc sd{};
                    sd.mesh = make_mesh(static_cast<size_t>(e.mesh), p, prims.at(p));
                    sd.material = make_material(prims.at(p));
                    e.surfaces.push_back(sd);
                }
            }
            const int me = static_cast<int>(ents.size());
            ents.push_back(e);
            if (ents[me].name == camera_name) camera_entity = me;
            if (sun && ents[me].name == sun_name) sun_entity = me;
            if (const Json* ch = n.find("children"))
                for (size_t c = 0; c < ch->size(); c++) {
                    const int child = process(static_cast<size_t>(ch->at(c).number(0)), me, depth + 1);
                    ents[me].children.push_back(child);
                }
            if (parent < 0) roots[ents[me].name] = me; // duplicates overwrite
            return me;
        };
        const Json& scene_nodes = scenes.at(0).has("nodes") ? scenes.at(0).at("nodes") : g.arr("__none__");
        for (size_t i = 0; i < scene_nodes.size(); i++) process(static_cast<size_t>(scene_nodes.at(i).number(0)), -1, 0);
        if (camera_entity < 0) throw Error(PTB_E_IO, "Scene is missing a camera.");

        std::function<Xform(int)> global = [&](int e) -> Xform { // entity::get_global_transform
            if (ents[e].parent >= 0) return compose(global(ents[e].parent), ents[e].local);
            return ents[e].local;
        };

        // ---- visiting order of renderer::intersect ----
        std::vector<int> order;
        {
            std::vector<int> stack;
            for (const auto& kv : roots) stack.push_back(kv.second);
            while (!stack.empty()) {
                const int e = stack.back();
                stack.pop_back();
                for (int c : ents[e].children) stack.push_back(c);
                if (ents[e].has_model) order.push_back(e);
            }
        }

        for (int e : order) {
            ptb_instance_desc inst{};
            const Xform gx = global(e);
            inst.origin[0] = gx.origin.x; inst.origin[1] = gx.origin.y; inst.origin[2] = gx.origin.z;
            const M3& b = gx.basis;
            const float bb[9] = {b.x.x, b.x.y, b.x.z, b.y.x, b.y.y, b.y.z, b.z.x, b.z.y, b.z.z};
            std::memcpy(inst.basis, bb, sizeof(bb));
            inst.first_surface = static_cast<uint32_t>(out.surfaces.size());
            inst.n_surfaces = static_cast<uint32_t>(ents[e].surfaces.size());
            for (const ptb_surface_desc& sd : ents[e].surfaces) out.surfaces.push_back(sd);
            out.instances.push_back(inst);
        }

        // Renumber meshes, materials and textures by first use in visiting order (materials were created in
        // node-processing order so that the texture cache sees the reference's load order); objects that no
        // visited surface uses (an overwritten duplicate root entity, the occlusion texture) are dropped.
        {
            const uint32_t none = PTB_NO_TEXTURE;
            std::vector<uint32_t> mesh_map(out.meshes.size(), none), mat_map(out.materials.size(), none),
                tex_map(out.textures.size(), none);
            std::vector<OwnedMesh> meshes2;
            std::vector<ptb_material_desc> mats2;
            std::vector<OwnedTexture> tex2;
            for (ptb_surface_desc& sd : out.surfaces) {
                if (mesh_map[sd.mesh] == none) {
                    mesh_map[sd.mesh] = static_cast<uint32_t>(meshes2.size());
                    meshes2.push_back(std::move(out.meshes[sd.mesh]));
                }
                if (mat_map[sd.material] == none) {
                    mat_map[sd.material] = static_cast<uint32_t>(mats2.size());
                    ptb_material_desc m = out.materials[sd.material];
                    uint32_t* slots[6] = {&m.normal_tex, &m.albedo_tex, &m.opacity_tex,
                                          &m.roughness_tex, &m.metallic_tex, &m.emissive_tex};
                    for (uint32_t* slot : slots) {
                        if (*slot == none) continue;
                        if (tex_map[*slot] == none) {
                            tex_map[*slot] = static_cast<uint32_t>(tex2.size());
                            tex2.push_back(out.textures[*slot]); // copy: a texture may be shared
                        }
                        *slot = tex_map[*slot];
                    }
                    mats2.push_back(m);
                }
                sd.mesh = mesh_map[sd.mesh];
                sd.material = mat_map[sd.material];
            }
            out.meshes = std::move(meshes2);
            out.materials = std::move(mats2);
            out.textures = std::move(tex2);
        }

        // ---- camera, sun ----
        {
            const Xform cx = global(camera_entity);
            out.camera.origin[0] = cx.origin.x; out.camera.origin[1] = cx.origin.y; out.camera.origin[2] = cx.origin.z;
            const M3& b = cx.basis;
            const float bb[9] = {b.x.x, b.x.y, b.x.z, b.y.x, b.y.y, b.y.z, b.z.x, b.z.y, b.z.z};
            std::memcpy(out.camera.basis, bb, sizeof(bb));
            const Json* persp = cam.find("perspective");
            out.camera.yfov = persp ? static_cast<float>(persp->get("yfov", 0.0)) : 0.0f;
        }
        std::memset(&out.sun, 0, sizeof(out.sun));
        if (sun && sun_entity >= 0) {
            const Xform sx = global(sun_entity);
            const M3& b = sx.basis;
            const float bb[9] = {b.x.x, b.x.y, b.x.z, b.y.x, b.y.y, b.y.z, b.z.x, b.z.y, b.z.z};
            out.sun.enabled = 1;
            std::memcpy(out.sun.basis, bb, sizeof(bb));
            float color[3] = {1, 1, 1};
            if (const Json* c = sun->find("color"))
                for (int k = 0; k < 3; k++) color[k] = fnum(*c, k, 1);
            const float intensity = static_cast<float>(sun->get("intensity", 1.0));
            for (int k = 0; k < 3; k++) out.sun.energy[k] = color[k] * intensity; // renderer.cpp:159
            out.sun.angular_radius = 0.004732f;                                   // sun_light.hpp:10
        }
        out.environment_factor[0] = out.environment_factor[1] = out.environment_factor[2] = 1.0f;
        out.transparent_background = 0;
    } catch (const Error&) {
        throw;
    } catch (const std::exception& e) {
        throw Error(PTB_E_IO, std::string("glTF: ") + e.what());
    }
}

} // namespace ptb
