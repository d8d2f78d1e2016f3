// kd_build.hpp — host-side KD-tree construction for one triangle mesh.
//
// Produces, from scratch, the SAME tree that the reference's recursive
// builder produces (kd_tree_builder::init_node_sah / init_node_median,
// LIB/core/mesh.cpp:82-247, LIB = path-tracer-core/path_tracer_lib/path_tracer),
// because closest-hit triangle ids can only be bit-exact on an identical tree.
// The construction itself is organised differently: it works on triangle
// indices and per-triangle bounds instead of copying triangle vectors, runs
// independent subtrees on a thread pool, and emits a flattened, 8-byte-per-node
// array that is what the GPU traverses.
#pragma once

#include <cstdint>
#include <vector>

namespace ptb {

struct Aabb {
    float min[3];
    float max[3];
};

// One flattened node, 8 bytes, 8 per 64-byte line.
//   branch: w0 = split plane (float bits)
//           w1 = axis (bits 0..1, values 0..2) | has_left << 2 | has_right << 3 | first_child << 4
//           children are adjacent: left at first_child, right at first_child + has_left
//   leaf:   w0 = first reference (index into refs[])
//           w1 = 3 | count << 2
struct KdNode {
    uint32_t w0;
    uint32_t w1;
};

constexpr uint32_t KD_LEAF_TAG = 3u;

struct KdTree {
    std::vector<KdNode> nodes;  // nodes[0] is the root
    std::vector<uint32_t> refs; // leaf triangle references (mesh triangle indices)
    uint64_t n_branches = 0, n_leaves = 0;
    uint32_t max_depth_reached = 0; // number of branch levels above the deepest leaf
};

// mesh::recalculate_aabb (LIB/core/mesh.cpp:254-261) including the
// aabb::clear() quirk (LIB/geometry/aabb.cpp:29-32: max starts at FLT_MIN, the
// smallest POSITIVE float) and the +-epsilon padding.
Aabb mesh_aabb(const float* positions, uint32_t n_vertices);

// positions: n_vertices*3, indices: n_triangles*3.  `root` is the box returned
// by mesh_aabb.  threads <= 0 → hardware concurrency.
void build_kd_tree(const float* positions, const uint32_t* indices, uint32_t n_triangles, const Aabb& root,
                   bool use_sah, uint32_t max_depth, int threads, KdTree& out);

// build_kd_tree behind an on-disk cache of the flattened tree.  The cache directory is the
// environment variable PTB_KD_CACHE (unset or empty: no cache).  A file is named by two independent 64-bit
// hashes of everything the tree depends on (vertex positions, indices, root box, builder kind, depth limit,
// format version) and is verified on load (sizes, magic, trailing checksum); a damaged or foreign file is
// ignored and rebuilt.  Returns true when the tree came from the cache.
bool build_kd_tree_cached(const float* positions, uint32_t n_vertices, const uint32_t* indices, uint32_t n_triangles,
                          const Aabb& root, bool use_sah, uint32_t max_depth, int threads, KdTree& out);

// Depth-first record stream documented at ptb_scene_dump_kd (include/ptb.h).
void dump_kd_tree(const KdTree& tree, std::vector<uint32_t>& words);

} // namespace ptb
