// kd_build.cpp — see kd_build.hpp.  Compiled with -ffp-contract=off: every
// float operation below has to round exactly as the reference's does.
#include "kd_build.hpp"

#include <algorithm>
#include <functional>
#include <string>
#include <cstdlib>
#include <cstdio>
#include <atomic>
#include <cfloat>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <thread>
#include <utility>

namespace ptb {

namespace {

constexpr float kEps = 0.0001f; // math::epsilon, LIB/math/math.hpp:16

// math::min / math::max (LIB/math/math.inl:169-187) are "b < a ? b : a".
inline float rmin(float a, float b) { return b < a ? b : a; }
inline float rmax(float a, float b) { return b > a ? b : a; }

// aabb::get_surface_area, LIB/geometry/aabb.cpp:34-39.
inline float surface_area(const Aabb& b) {
    float wx = b.max[0] - b.min[0], wy = b.max[1] - b.min[1], wz = b.max[2] - b.min[2];
    return (wx * wy + wy * wz + wx * wz) * 2;
}

struct BuildNode {
    int axis = -1; // -1: leaf
    float split = 0;
    std::unique_ptr<BuildNode> left, right;
    std::vector<uint32_t> tris; // leaf only, parent order preserved
};

struct Event {
    float pos;
    bool start;
};

struct Builder {
    // per-triangle bounds, [axis][triangle]
    std::vector<float> lo[3], hi[3];
    bool use_sah;

    struct Split {
        bool found = false;
        float cost = 0;
        float split = 0;
    };

    // One axis of the sweep of init_node_sah (LIB/core/mesh.cpp:151-211).  Returns
    // the first candidate, in sweep order, whose cost is the axis minimum below
    // `limit` (strict '<', as the running best_cost comparison at :205).
    Split sweep_axis(const Aabb& box, const std::vector<uint32_t>& tris, int axis, float limit,
                     std::vector<Event>& ev) const {
        ev.clear();
        const float* l = lo[axis].data();
        const float* h = hi[axis].data();
        for (uint32_t t : tris) { // :154-160, start then end per triangle
            ev.push_back({l[t], true});
            ev.push_back({h[t], false});
        }
        // :162-163 — same algorithm (libstdc++ std::sort) on the same sequence with
        // the same comparison, hence the same permutation, including which of
        // several equal-position events ends up last (it matters at :172-176).
        std::sort(ev.begin(), ev.end(), [](const Event& x, const Event& y) { return x.pos < y.pos; });

        Split best;
        best.cost = limit;
        float split = 0;
        uint32_t lcount = 0;
        uint32_t rcount = static_cast<uint32_t>(tris.size());
        const size_t n = ev.size();
        for (size_t i = 0; i <= n; i++) {
            if (i == 0) {
                split = ev.front().pos - kEps;
            } else if (i == n) {
                rcount--; // "Last event is always END" — taken literally, wraps if it is not
                split = ev.back().pos + kEps;
            } else {
                const Event& prev = ev[i - 1];
                const Event& next = ev[i];
                if (prev.start)
                    lcount++;
                else
                    rcount--;
                if (prev.pos == next.pos)
                    continue;
                split = (prev.pos + next.pos) * 0.5F;
            }
            if (split <= box.min[axis])
                continue;
            if (split >= box.max[axis])
                break;
            Aabb lb = box, rb = box; // split_aabb, :21-33
            lb.max[axis] = split;
            rb.min[axis] = split;
            float cost = lcount * surface_area(lb) + rcount * surface_area(rb);
            if (cost < best.cost) {
                best.cost = cost;
                best.split = split;
                best.found = true;
            }
        }
        return best;
    }

    // Decide what one node becomes.  Returns false for a leaf.
    bool choose_split(const Aabb& box, const std::vector<uint32_t>& tris, uint32_t depth, bool axes_parallel,
                      int& axis_out, float& split_out) const {
        if (depth == 0)
            return false;
        if (!use_sah) { // init_node_median, :96-101 (never makes a leaf above depth 0)
            float w[3] = {box.max[0] - box.min[0], box.max[1] - box.min[1], box.max[2] - box.min[2]};
            int a = 0; // std::max_element: first of the largest
            if (w[1] > w[a]) a = 1;
            if (w[2] > w[a]) a = 2;
            axis_out = a;
            split_out = box.min[a] + w[a] * 0.5F;
            return true;
        }
        float base_cost = tris.size() * surface_area(box); // :143
        Split per_axis[3];
        if (axes_parallel) {
            std::thread th[2];
            for (int a = 0; a < 2; a++)
                th[a] = std::thread([&, a]() {
                    std::vector<Event> ev;
                    ev.reserve(tris.size() * 2);
                    per_axis[a + 1] = sweep_axis(box, tris, a + 1, base_cost, ev);
                });
            std::vector<Event> ev;
            ev.reserve(tris.size() * 2);
            per_axis[0] = sweep_axis(box, tris, 0, base_cost, ev);
            th[0].join();
            th[1].join();
        } else {
            thread_local std::vector<Event> ev;
            ev.reserve(tris.size() * 2);
            for (int a = 0; a < 3; a++)
                per_axis[a] = sweep_axis(box, tris, a, base_cost, ev);
        }
        // The reference keeps one running best over (axis, sweep) order with
        // strict '<': the earliest global minimum wins.
        float best_cost = base_cost;
        bool found = false;
        for (int a = 0; a < 3; a++) {
            if (per_axis[a].found && per_axis[a].cost < best_cost) {
                best_cost = per_axis[a].cost;
                axis_out = a;
                split_out = per_axis[a].split;
                found = true;
            }
        }
        return found; // :213 best_cost < base_cost
    }

    // split_triangles, :35-80: a vertex below the plane sends the triangle left,
    // a vertex at or above it sends it right; both can hold.
    void partition(const std::vector<uint32_t>& tris, int axis, float split, std::vector<uint32_t>& l,
                   std::vector<uint32_t>& r) const {
        const float* lw = lo[axis].data();
        const float* hg = hi[axis].data();
        l.reserve(tris.size());
        r.reserve(tris.size());
        for (uint32_t t : tris) {
            if (lw[t] < split) l.push_back(t);
            if (hg[t] >= split) r.push_back(t);
        }
    }

    struct Pending {
        BuildNode* node;
        Aabb box;
        std::vector<uint32_t> tris;
        uint32_t depth;
    };

    // Expand one node; children that need further work are appended to `out`.
    void expand(Pending&& p, bool axes_parallel, std::vector<Pending>& out) const {
        int axis = 0;
        float split = 0;
        if (!choose_split(p.box, p.tris, p.depth, axes_parallel, axis, split)) {
            p.node->axis = -1;
            p.node->tris = std::move(p.tris);
            p.node->tris.shrink_to_fit();
            return;
        }
        p.node->axis = axis;
        p.node->split = split;
        std::vector<uint32_t> l, r;
        partition(p.tris, axis, split, l, r);
        std::vector<uint32_t>().swap(p.tris);
        if (!l.empty()) { // :227-233 — an empty side stays a null child
            p.node->left = std::make_unique<BuildNode>();
            Aabb lb = p.box;
            lb.max[axis] = split;
            out.push_back({p.node->left.get(), lb, std::move(l), p.depth - 1});
        }
        if (!r.empty()) {
            p.node->right = std::make_unique<BuildNode>();
            Aabb rb = p.box;
            rb.min[axis] = split;
            out.push_back({p.node->right.get(), rb, std::move(r), p.depth - 1});
        }
    }

    void build_subtree(Pending&& root) const {
        std::vector<Pending> stack;
        stack.push_back(std::move(root));
        std::vector<Pending> kids;
        while (!stack.empty()) {
            Pending p = std::move(stack.back());
            stack.pop_back();
            kids.clear();
            expand(std::move(p), false, kids);
            for (auto& k : kids) stack.push_back(std::move(k));
        }
    }
};

void flatten(const BuildNode* n, uint32_t idx, uint32_t depth, KdTree& out) {
    if (n->axis < 0) {
        out.nodes[idx].w0 = static_cast<uint32_t>(out.refs.size());
        out.nodes[idx].w1 = KD_LEAF_TAG | (static_cast<uint32_t>(n->tris.size()) << 2);
        out.refs.insert(out.refs.end(), n->tris.begin(), n->tris.end());
        out.n_leaves++;
        out.max_depth_reached = std::max(out.max_depth_reached, depth);
        return;
    }
    uint32_t has_l = n->left ? 1 : 0, has_r = n->right ? 1 : 0;
    uint64_t first = out.nodes.size();
    if (first + 2 >= (1ull << 28))
        throw std::runtime_error("KD tree exceeds 2^28 nodes");
    out.nodes.resize(first + has_l + has_r);
    uint32_t bits;
    std::memcpy(&bits, &n->split, 4);
    out.nodes[idx].w0 = bits;
    out.nodes[idx].w1 = static_cast<uint32_t>(n->axis) | (has_l << 2) | (has_r << 3) |
                        (static_cast<uint32_t>(first) << 4);
    out.n_branches++;
    if (has_l) flatten(n->left.get(), static_cast<uint32_t>(first), depth + 1, out);
    if (has_r) flatten(n->right.get(), static_cast<uint32_t>(first + has_l), depth + 1, out);
}

void dump_rec(const KdTree& t, uint32_t idx, std::vector<uint32_t>& w) {
    const KdNode& n = t.nodes[idx];
    if ((n.w1 & 3u) == KD_LEAF_TAG) {
        uint32_t count = n.w1 >> 2;
        w.push_back(count);
        for (uint32_t i = 0; i < count; i++) w.push_back(t.refs[n.w0 + i]);
        return;
    }
    uint32_t has_l = (n.w1 >> 2) & 1, has_r = (n.w1 >> 3) & 1, first = n.w1 >> 4;
    w.push_back(0x80000000u | (n.w1 & 3u));
    w.push_back(n.w0);
    w.push_back(has_l);
    w.push_back(has_r);
    if (has_l) dump_rec(t, first, w);
    if (has_r) dump_rec(t, first + has_l, w);
}

} // namespace

Aabb mesh_aabb(const float* positions, uint32_t n_vertices) {
    Aabb b;
    for (int a = 0; a < 3; a++) {
        b.min[a] = FLT_MAX;
        b.max[a] = FLT_MIN; // sic: smallest positive float (aabb.cpp:31)
    }
    for (uint32_t v = 0; v < n_vertices; v++)
        for (int a = 0; a < 3; a++) {
            float p = positions[3 * v + a];
            b.min[a] = rmin(b.min[a], p); // aabb::add, aabb.cpp:19-22
            b.max[a] = rmax(b.max[a], p);
        }
    for (int a = 0; a < 3; a++) {
        b.min[a] = b.min[a] - kEps;
        b.max[a] = b.max[a] + kEps;
    }
    return b;
}

void build_kd_tree(const float* positions, const uint32_t* indices, uint32_t n_triangles, const Aabb& root_box,
                   bool use_sah, uint32_t max_depth, int threads, KdTree& out) {
    if (threads <= 0) threads = std::max(1u, std::thread::hardware_concurrency());
    Builder b;
    b.use_sah = use_sah;
    for (int a = 0; a < 3; a++) {
        b.lo[a].resize(n_triangles);
        b.hi[a].resize(n_triangles);
    }
    for (uint32_t t = 0; t < n_triangles; t++) {
        const float* pa = positions + 3 * size_t(indices[3 * t]);
        const float* pb = positions + 3 * size_t(indices[3 * t + 1]);
        const float* pc = positions + 3 * size_t(indices[3 * t + 2]);
        for (int a = 0; a < 3; a++) {
            b.lo[a][t] = rmin(rmin(pa[a], pb[a]), pc[a]); // math::min(a,b,c), mesh.cpp:155
            b.hi[a][t] = rmax(rmax(pa[a], pb[a]), pc[a]);
        }
    }

    auto root = std::make_unique<BuildNode>();
    std::vector<uint32_t> all(n_triangles);
    for (uint32_t t = 0; t < n_triangles; t++) all[t] = t;

    // Phase 1: breadth-first over the top of the tree (the three axis sweeps
    // of a big node run on three threads) until there is enough independent
    // work; phase 2: the open subtrees run in parallel, each single-threaded.
    // Every node's outcome depends only on its own (box, triangle list, depth),
    // so the schedule cannot change the tree.
    std::vector<Builder::Pending> open;
    open.push_back({root.get(), root_box, std::move(all), max_depth});
    const size_t want = threads > 1 ? size_t(threads) * 8 : 1;
    const size_t small = 4096;
    while (threads > 1 && open.size() < want) {
        size_t big = open.size();
        size_t big_n = small;
        for (size_t i = 0; i < open.size(); i++)
            if (open[i].tris.size() > big_n) {
                big = i;
                big_n = open[i].tris.size();
            }
        if (big == open.size()) break; // nothing large left to split
        Builder::Pending p = std::move(open[big]);
        open.erase(open.begin() + big);
        b.expand(std::move(p), true, open);
    }
    // largest first, so the tail of the parallel phase is made of small jobs
    std::sort(open.begin(), open.end(),
              [](const Builder::Pending& x, const Builder::Pending& y) { return x.tris.size() > y.tris.size(); });
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= open.size()) break;
            b.build_subtree(std::move(open[i]));
        }
    };
    std::vector<std::thread> pool;
    int nthreads = std::min<size_t>(threads, std::max<size_t>(1, open.size()));
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();

    out = KdTree();
    out.nodes.reserve(size_t(n_triangles) * 3 + 16);
    out.refs.reserve(size_t(n_triangles) * 8 + 16);
    out.nodes.resize(1);
    flatten(root.get(), 0, 0, out);
    out.nodes.shrink_to_fit();
    out.refs.shrink_to_fit();
}

void dump_kd_tree(const KdTree& tree, std::vector<uint32_t>& words) {
    words.clear();
    if (!tree.nodes.empty()) dump_rec(tree, 0, words);
}

} // namespace ptb

// ---- on-disk cache -------------------------------------------------------------------------------------

namespace ptb {

namespace {

constexpr uint64_t KD_CACHE_MAGIC = 0x3130444B42545050ull; // "PPTBKD01"

struct Hash2 {
    uint64_t a = 0xcbf29ce484222325ull, b = 0x9E3779B97F4A7C15ull;
    void feed(const void* data, size_t bytes) {
        const unsigned char* p = static_cast<const unsigned char*>(data);
        // a: FNV-1a over 8-byte words (tail bytewise); b: multiply-xorshift mix of the same words
        size_t i = 0;
        for (; i + 8 <= bytes; i += 8) {
            uint64_t w;
            std::memcpy(&w, p + i, 8);
            a = (a ^ w) * 0x100000001b3ull;
            b += w * 0xBF58476D1CE4E5B9ull;
            b = (b ^ (b >> 29)) * 0x94D049BB133111EBull;
        }
        for (; i < bytes; i++) {
            a = (a ^ p[i]) * 0x100000001b3ull;
            b = (b ^ (uint64_t(p[i]) + 0x9E37u)) * 0x94D049BB133111EBull;
        }
    }
    template <typename T>
    void feed_value(const T& v) { feed(&v, sizeof(T)); }
};

struct CacheHeader {
    uint64_t magic, key_a, key_b;
    uint64_t n_nodes, n_refs, n_branches, n_leaves;
    uint32_t max_depth_reached, reserved;
};

uint64_t payload_checksum(const KdTree& t) {
    Hash2 h;
    h.feed(t.nodes.data(), t.nodes.size() * sizeof(KdNode));
    h.feed(t.refs.data(), t.refs.size() * sizeof(uint32_t));
    return h.a ^ h.b;
}

bool load_tree(const std::string& path, uint64_t ka, uint64_t kb, KdTree& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = false;
    CacheHeader h{};
    if (std::fread(&h, sizeof(h), 1, f) == 1 && h.magic == KD_CACHE_MAGIC && h.key_a == ka && h.key_b == kb &&
        h.n_nodes < (1ull << 32) && h.n_refs < (1ull << 32)) {
        KdTree t;
        t.nodes.resize(h.n_nodes);
        t.refs.resize(h.n_refs);
        uint64_t sum = 0;
        if (std::fread(t.nodes.data(), sizeof(KdNode), h.n_nodes, f) == h.n_nodes &&
            std::fread(t.refs.data(), sizeof(uint32_t), h.n_refs, f) == h.n_refs &&
            std::fread(&sum, sizeof(sum), 1, f) == 1 && sum == payload_checksum(t)) {
            t.n_branches = h.n_branches;
            t.n_leaves = h.n_leaves;
            t.max_depth_reached = h.max_depth_reached;
            out = std::move(t);
            ok = true;
        }
    }
    std::fclose(f);
    return ok;
}

void store_tree(const std::string& path, uint64_t ka, uint64_t kb, const KdTree& t) {
    const std::string tmp = path + ".tmp" + std::to_string((unsigned long long)std::hash<std::thread::id>()(std::this_thread::get_id()));
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return; // an unwritable cache directory is not an error: the tree is simply not cached
    CacheHeader h{KD_CACHE_MAGIC, ka, kb, t.nodes.size(), t.refs.size(), t.n_branches, t.n_leaves, t.max_depth_reached, 0};
    const uint64_t sum = payload_checksum(t);
    const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 &&
                    std::fwrite(t.nodes.data(), sizeof(KdNode), t.nodes.size(), f) == t.nodes.size() &&
                    std::fwrite(t.refs.data(), sizeof(uint32_t), t.refs.size(), f) == t.refs.size() &&
                    std::fwrite(&sum, sizeof(sum), 1, f) == 1;
    const bool closed = std::fclose(f) == 0;
    if (ok && closed)
        std::rename(tmp.c_str(), path.c_str()); // atomic: readers see the old file, no file, or the whole new file
    else
        std::remove(tmp.c_str());
}

} // namespace

bool build_kd_tree_cached(const float* positions, uint32_t n_vertices, const uint32_t* indices, uint32_t n_triangles,
                          const Aabb& root, bool use_sah, uint32_t max_depth, int threads, KdTree& out) {
    const char* dir = std::getenv("PTB_KD_CACHE");
    if (!dir || !*dir) {
        build_kd_tree(positions, indices, n_triangles, root, use_sah, max_depth, threads, out);
        return false;
    }
    Hash2 h;
    const uint32_t version = 1, sah = use_sah ? 1u : 0u;
    h.feed_value(version);
    h.feed_value(n_vertices);
    h.feed_value(n_triangles);
    h.feed_value(sah);
    h.feed_value(max_depth);
    h.feed(&root, sizeof(root));
    h.feed(positions, size_t(n_vertices) * 3 * sizeof(float));
    h.feed(indices, size_t(n_triangles) * 3 * sizeof(uint32_t));
    char name[64];
    std::snprintf(name, sizeof(name), "/kd_%016llx%016llx.bin", (unsigned long long)h.a, (unsigned long long)h.b);
    const std::string path = std::string(dir) + name;
    if (load_tree(path, h.a, h.b, out)) return true;
    build_kd_tree(positions, indices, n_triangles, root, use_sah, max_depth, threads, out);
    store_tree(path, h.a, h.b, out);
    return false;
}

} // namespace ptb
