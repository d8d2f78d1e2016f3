/*
 * pt_oracle.c — TEST INFRASTRUCTURE ONLY.  Plain-C, scalar, CPU restatement of
 * the reference algorithm for the hot path of vmanam0451/distributed-path-tracer
 * (KD build, KD traversal, ray/triangle intersection, Monte-Carlo integrator).
 *
 * Nothing of the product links, loads or calls this file: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * do, and only as the checker.  It is pinned against golden vectors minted from
 * the UNMODIFIED reference library (tests/golden/, see make_golden.py): KD trees
 * word for word, closest hits bit for bit, images statistically.
 *
 * Every function cites the reference code it follows.  Paths are relative to the
 * reference repository: LIB/ = path-tracer-core/path_tracer_lib/path_tracer/,
 * APP/ = path-tracer-core/src/.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (oracle/Makefile) — float operations
 * must not be fused, the reference ships x86-64 SSE2 code without FMA.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/ptb.h"

#define EPS 0.0001f /* math::epsilon, LIB/math/math.hpp:16 */

typedef struct { float x, y, z; } v3;
typedef struct { v3 x, y, z; } m3;       /* columns, LIB/math/mat3.inl:13-29 */
typedef struct { v3 origin; m3 basis; } xform; /* LIB/scene/transform.hpp:14-15 */

/* ---- LIB/math: operation order matters ---------------------------------- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 divv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 smul(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float fminr(float a, float b) { return b < a ? b : a; } /* math.inl:169-187 */
static inline float fmaxr(float a, float b) { return b > a ? b : a; }
static inline float clampr(float x, float lo, float hi) { return fminr(fmaxr(x, lo), hi); }
static inline float lerpf(float a, float b, float w) { return a + (b - a) * w; } /* math.inl:164-167 */
static inline v3 lerp3s(v3 a, v3 b, float w) { return V(lerpf(a.x, b.x, w), lerpf(a.y, b.y, w), lerpf(a.z, b.z, w)); }
static inline v3 lerp3v(v3 a, v3 b, v3 w) { return V(lerpf(a.x, b.x, w.x), lerpf(a.y, b.y, w.y), lerpf(a.z, b.z, w.z)); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* vec3.inl:235-238 */
static inline v3 cross(v3 l, v3 r) { /* vec3.inl:221-228 */
    return V((l.y * r.z) - (l.z * r.y), (l.z * r.x) - (l.x * r.z), (l.x * r.y) - (l.y * r.x));
}
static inline float length(v3 v) { return sqrtf(dot(v, v)); }
static inline v3 normalize(v3 v) { return muls(v, 1 / length(v)); } /* vec3.inl:250-253 */
static inline v3 reflect(v3 i, v3 n) { return sub(i, smul(2 * dot(n, i), n)); } /* LIB/core/utils.hpp:38-40 */
static inline v3 mat_vec(const m3* m, v3 v) { /* mat3.inl:219-224: transpose, row dots */
    return V(dot(V(m->x.x, m->y.x, m->z.x), v), dot(V(m->x.y, m->y.y, m->z.y), v), dot(V(m->x.z, m->y.z, m->z.z), v));
}
static inline m3 transpose(m3 m) {
    m3 r = {V(m.x.x, m.y.x, m.z.x), V(m.x.y, m.y.y, m.z.y), V(m.x.z, m.y.z, m.z.z)};
    return r;
}
static m3 mat_inverse(m3 m) { /* mat3.inl:245-263 */
    float det1 = +(m.y.y * m.z.z - m.z.y * m.y.z);
    float det2 = -(m.x.y * m.z.z - m.z.y * m.x.z);
    float det3 = +(m.x.y * m.y.z - m.y.y * m.x.z);
    float det = m.x.x * det1 + m.y.x * det2 + m.z.x * det3;
    float r = 1 / det;
    m3 a = {V(det1, det2, det3),
            V(-(m.y.x * m.z.z - m.z.x * m.y.z), +(m.x.x * m.z.z - m.z.x * m.x.z), -(m.x.x * m.y.z - m.y.x * m.x.z)),
            V(+(m.y.x * m.z.y - m.z.x * m.y.y), -(m.x.x * m.z.y - m.z.x * m.x.y), +(m.x.x * m.y.y - m.y.x * m.x.y))};
    m3 o = {muls(a.x, r), muls(a.y, r), muls(a.z, r)};
    return o;
}
static inline v3 xf_apply(const xform* t, v3 v) { return add(mat_vec(&t->basis, v), t->origin); } /* transform.cpp:117-119 */
static xform xf_inverse(const xform* t) { /* transform.cpp:33-36 */
    xform r;
    r.basis = mat_inverse(t->basis);
    r.origin = mat_vec(&r.basis, neg(t->origin));
    return r;
}
static xform xf_from(const float* o, const float* b) {
    xform t;
    t.origin = V(o[0], o[1], o[2]);
    t.basis.x = V(b[0], b[1], b[2]);
    t.basis.y = V(b[3], b[4], b[5]);
    t.basis.z = V(b[6], b[7], b[8]);
    return t;
}

/* ---- scene ----------------------------------------------------------------- */
typedef struct { float min[3], max[3]; } aabb;

typedef struct kd_node { /* LIB/core/kd_tree.hpp:10-31 */
    int axis;            /* -1: leaf */
    float split;
    struct kd_node *left, *right;
    uint32_t* tris;      /* leaf: triangle indices (leaf->indices) */
    uint32_t count;
} kd_node;

typedef struct {
    uint32_t n_vertices, n_triangles;
    const float *positions, *normals, *tangents, *uvs; /* copies owned by the scene */
    const uint32_t* indices;
    aabb box;
    kd_node* root;
} po_mesh;

typedef struct {
    xform fwd, inv;
    m3 normal_mat;
    aabb box; /* scene::model::aabb */
    uint32_t first_surface, n_surfaces;
} po_instance;

typedef struct po_scene {
    uint32_t n_meshes, n_surfaces, n_instances, n_materials;
    po_mesh* meshes;
    ptb_surface_desc* surfaces;
    po_instance* instances;
    ptb_material_desc* materials;
    uint32_t n_textures;
    ptb_texture_desc* textures; /* pixels are owned copies */
    xform camera;
    float tan_half_fov;
    int sun_enabled;
    v3 sun_dir, sun_energy;
    float sun_radius;
    v3 environment;
    uint32_t environment_tex; /* PTB_NO_TEXTURE: none */
    int transparent;
} po_scene;

/* ---- std::sort restated ------------------------------------------------------
 * init_node_sah sorts its (position, is_start) events with std::sort and a
 * comparison on the position only (LIB/core/mesh.cpp:162-163).  std::sort is not
 * stable, and which of several equal-position events ends up LAST decides the
 * final candidate plane's counts (mesh.cpp:172-176), so the tree depends on the
 * exact algorithm.  This is libstdc++'s (GCC 13, bits/stl_algo.h, the toolchain
 * the reference is built with here): introsort — median-of-3 quicksort until
 * 16-element runs, depth limit 2*floor(log2 n) with a heapsort fallback, then
 * one insertion-sort pass. */
typedef struct { float pos; int start; } event;
#define EV_LT(a, b) ((a).pos < (b).pos)

static void ev_swap(event* a, event* b) { event t = *a; *a = *b; *b = t; }

static void ev_move_median_to_first(event* result, event* a, event* b, event* c) {
    if (EV_LT(*a, *b)) {
        if (EV_LT(*b, *c)) ev_swap(result, b);
        else if (EV_LT(*a, *c)) ev_swap(result, c);
        else ev_swap(result, a);
    } else if (EV_LT(*a, *c)) ev_swap(result, a);
    else if (EV_LT(*b, *c)) ev_swap(result, c);
    else ev_swap(result, b);
}
static event* ev_unguarded_partition(event* first, event* last, event* pivot) {
    for (;;) {
        while (EV_LT(*first, *pivot)) ++first;
        --last;
        while (EV_LT(*pivot, *last)) --last;
        if (!(first < last)) return first;
        ev_swap(first, last);
        ++first;
    }
}
static void ev_push_heap(event* first, long hole, long top, event value) {
    long parent = (hole - 1) / 2;
    while (hole > top && EV_LT(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void ev_adjust_heap(event* first, long hole, long len, event value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (EV_LT(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    ev_push_heap(first, hole, top, value);
}
static void ev_heapsort(event* first, event* last) { /* __partial_sort(first, last, last) */
    long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            event value = first[parent];
            ev_adjust_heap(first, parent, len, value);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        event value = *last;
        *last = *first;
        ev_adjust_heap(first, 0, last - first, value);
    }
}
static void ev_introsort_loop(event* first, event* last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) {
            ev_heapsort(first, last);
            return;
        }
        --depth_limit;
        event* mid = first + (last - first) / 2;
        ev_move_median_to_first(first, first + 1, mid, last - 1);
        event* cut = ev_unguarded_partition(first + 1, last, first);
        ev_introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}
static void ev_unguarded_linear_insert(event* last) {
    event val = *last;
    event* next = last - 1;
    while (EV_LT(val, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}
static void ev_insertion_sort(event* first, event* last) {
    if (first == last) return;
    for (event* i = first + 1; i != last; ++i) {
        if (EV_LT(*i, *first)) {
            event val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(event));
            *first = val;
        } else {
            ev_unguarded_linear_insert(i);
        }
    }
}
static void ev_sort(event* first, event* last) {
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((n >> (lg + 1)) > 0) lg++;
    ev_introsort_loop(first, last, lg * 2);
    if (last - first > 16) {
        ev_insertion_sort(first, first + 16);
        for (event* i = first + 16; i != last; ++i) ev_unguarded_linear_insert(i);
    } else {
        ev_insertion_sort(first, last);
    }
}

/* ---- KD build: kd_tree_builder, LIB/core/mesh.cpp:10-247 ------------------- */
static float surface_area(const aabb* b) { /* aabb::get_surface_area, LIB/geometry/aabb.cpp:34-39 */
    float wx = b->max[0] - b->min[0], wy = b->max[1] - b->min[1], wz = b->max[2] - b->min[2];
    return (wx * wy + wy * wz + wx * wz) * 2;
}

static kd_node* make_leaf(uint32_t* tris, uint32_t count) { /* init_leaf, mesh.cpp:10-19 */
    kd_node* n = (kd_node*)calloc(1, sizeof(kd_node));
    n->axis = -1;
    n->tris = tris;
    n->count = count;
    return n;
}

static inline float tri_coord(const po_mesh* m, uint32_t tri, int vertex, int axis) {
    return m->positions[3 * (size_t)m->indices[3 * (size_t)tri + vertex] + axis];
}

/* init_node_sah (mesh.cpp:131-247) / init_node_median (mesh.cpp:82-128); takes ownership of tris. */
static kd_node* build_node(const po_mesh* m, aabb box, uint32_t* tris, uint32_t count, uint32_t depth, int use_sah) {
    if (depth == 0) return make_leaf(tris, count);
    int best_axis = 0;
    float best_split = 0;
    int do_split = 0;
    if (!use_sah) {
        float w[3] = {box.max[0] - box.min[0], box.max[1] - box.min[1], box.max[2] - box.min[2]};
        best_axis = 0;
        if (w[1] > w[best_axis]) best_axis = 1;
        if (w[2] > w[best_axis]) best_axis = 2;
        best_split = box.min[best_axis] + w[best_axis] * 0.5F;
        do_split = 1;
    } else {
        float base_cost = count * surface_area(&box); /* :143 */
        float best_cost = base_cost;
        event* ev = (event*)malloc(sizeof(event) * 2 * (size_t)(count ? count : 1));
        for (int axis = 0; axis < 3; axis++) {
            size_t ne = 0;
            for (uint32_t i = 0; i < count; i++) { /* :154-160 */
                float a = tri_coord(m, tris[i], 0, axis), b = tri_coord(m, tris[i], 1, axis), c = tri_coord(m, tris[i], 2, axis);
                ev[ne].pos = fminr(fminr(a, b), c); ev[ne].start = 1; ne++;
                ev[ne].pos = fmaxr(fmaxr(a, b), c); ev[ne].start = 0; ne++;
            }
            ev_sort(ev, ev + ne); /* :162-163 */
            float split = 0;
            uint32_t lcount = 0, rcount = count;
            for (size_t i = 0; i <= ne; i++) { /* :169-210 */
                if (ne == 0) break; /* the reference would read bounds.front() of an empty vector */
                if (i == 0) {
                    split = ev[0].pos - EPS;
                } else if (i == ne) {
                    rcount--; /* "Last event is always END" */
                    split = ev[ne - 1].pos + EPS;
                } else {
                    if (ev[i - 1].start) lcount++; else rcount--;
                    if (ev[i - 1].pos == ev[i].pos) continue;
                    split = (ev[i - 1].pos + ev[i].pos) * 0.5F;
                }
                if (split <= box.min[axis]) continue;
                if (split >= box.max[axis]) break;
                aabb l = box, r = box; /* split_aabb, :21-33 */
                l.max[axis] = split;
                r.min[axis] = split;
                float cost = lcount * surface_area(&l) + rcount * surface_area(&r);
                if (cost < best_cost) {
                    best_cost = cost;
                    best_axis = axis;
                    best_split = split;
                }
            }
        }
        free(ev);
        do_split = best_cost < base_cost; /* :213 */
    }
    if (!do_split) return make_leaf(tris, count);
    /* split_triangles, :35-80 */
    uint32_t* l = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(count ? count : 1));
    uint32_t* r = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(count ? count : 1));
    uint32_t nl = 0, nr = 0;
    for (uint32_t i = 0; i < count; i++) {
        int lassign = 0, rassign = 0;
        for (int v = 0; v < 3; v++) {
            if (tri_coord(m, tris[i], v, best_axis) < best_split) lassign = 1; else rassign = 1;
        }
        if (lassign) l[nl++] = tris[i];
        if (rassign) r[nr++] = tris[i];
    }
    free(tris);
    kd_node* n = (kd_node*)calloc(1, sizeof(kd_node));
    n->axis = best_axis;
    n->split = best_split;
    if (nl > 0) { /* :227-233 */
        aabb lb = box;
        lb.max[best_axis] = best_split;
        n->left = build_node(m, lb, l, nl, depth - 1, use_sah);
    } else free(l);
    if (nr > 0) {
        aabb rb = box;
        rb.min[best_axis] = best_split;
        n->right = build_node(m, rb, r, nr, depth - 1, use_sah);
    } else free(r);
    return n;
}

static void free_node(kd_node* n) {
    if (!n) return;
    free_node(n->left);
    free_node(n->right);
    free(n->tris);
    free(n);
}

static aabb mesh_aabb(const float* pos, uint32_t nv) { /* mesh::recalculate_aabb, mesh.cpp:254-261 */
    aabb b;
    for (int a = 0; a < 3; a++) { b.min[a] = FLT_MAX; b.max[a] = FLT_MIN; } /* aabb::clear, aabb.cpp:29-32 (sic) */
    for (uint32_t v = 0; v < nv; v++)
        for (int a = 0; a < 3; a++) {
            b.min[a] = fminr(b.min[a], pos[3 * (size_t)v + a]);
            b.max[a] = fmaxr(b.max[a], pos[3 * (size_t)v + a]);
        }
    for (int a = 0; a < 3; a++) { b.min[a] = b.min[a] - EPS; b.max[a] = b.max[a] + EPS; }
    return b;
}

static void* dup_mem(const void* p, size_t bytes) {
    void* r = malloc(bytes ? bytes : 1);
    if (bytes) memcpy(r, p, bytes);
    return r;
}

po_scene* po_scene_create(const ptb_scene_desc* d) {
    po_scene* s = (po_scene*)calloc(1, sizeof(po_scene));
    s->n_meshes = d->n_meshes; s->n_surfaces = d->n_surfaces; s->n_instances = d->n_instances; s->n_materials = d->n_materials;
    s->meshes = (po_mesh*)calloc(d->n_meshes ? d->n_meshes : 1, sizeof(po_mesh));
    uint32_t max_depth = d->kd_max_depth ? d->kd_max_depth : 25;
    for (uint32_t i = 0; i < d->n_meshes; i++) {
        const ptb_mesh_desc* md = &d->meshes[i];
        po_mesh* m = &s->meshes[i];
        m->n_vertices = md->n_vertices; m->n_triangles = md->n_triangles;
        m->positions = (const float*)dup_mem(md->positions, sizeof(float) * 3 * (size_t)md->n_vertices);
        m->normals = (const float*)dup_mem(md->normals, sizeof(float) * 3 * (size_t)md->n_vertices);
        m->tangents = (const float*)dup_mem(md->tangents, sizeof(float) * 3 * (size_t)md->n_vertices);
        m->uvs = (const float*)dup_mem(md->uvs, sizeof(float) * 2 * (size_t)md->n_vertices);
        m->indices = (const uint32_t*)dup_mem(md->indices, sizeof(uint32_t) * 3 * (size_t)md->n_triangles);
        m->box = mesh_aabb(m->positions, m->n_vertices);
        uint32_t* all = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(md->n_triangles ? md->n_triangles : 1));
        for (uint32_t t = 0; t < md->n_triangles; t++) all[t] = t; /* std::iota, mesh.cpp:278-279 */
        m->root = build_node(m, m->box, all, md->n_triangles, max_depth, d->kd_use_sah != 0);
    }
    s->surfaces = (ptb_surface_desc*)dup_mem(d->surfaces, sizeof(ptb_surface_desc) * d->n_surfaces);
    s->materials = (ptb_material_desc*)dup_mem(d->materials, sizeof(ptb_material_desc) * d->n_materials);
    s->n_textures = d->n_textures;
    s->textures = (ptb_texture_desc*)dup_mem(d->textures, sizeof(ptb_texture_desc) * d->n_textures);
    for (uint32_t t = 0; t < d->n_textures; t++) {
        const ptb_texture_desc* td = &d->textures[t];
        s->textures[t].pixels = dup_mem(td->pixels, (size_t)td->width * td->height * td->channels * (td->is_float ? 4 : 1));
    }
    s->instances = (po_instance*)calloc(d->n_instances ? d->n_instances : 1, sizeof(po_instance));
    for (uint32_t i = 0; i < d->n_instances; i++) {
        const ptb_instance_desc* id = &d->instances[i];
        po_instance* in = &s->instances[i];
        in->fwd = xf_from(id->origin, id->basis);
        in->inv = xf_inverse(&in->fwd);                          /* model.cpp:22-25 */
        in->normal_mat = transpose(mat_inverse(in->fwd.basis));  /* LIB/core/renderer.cpp:698 */
        in->first_surface = id->first_surface; in->n_surfaces = id->n_surfaces;
        for (int a = 0; a < 3; a++) { in->box.min[a] = FLT_MAX; in->box.max[a] = FLT_MIN; } /* model::recalculate_aabb */
        for (uint32_t k = 0; k < id->n_surfaces; k++) {
            const aabb* mb = &s->meshes[s->surfaces[id->first_surface + k].mesh].box;
            for (int a = 0; a < 3; a++) { /* aabb::add(aabb): add(min), add(max) */
                in->box.min[a] = fminr(in->box.min[a], mb->min[a]); in->box.max[a] = fmaxr(in->box.max[a], mb->min[a]);
                in->box.min[a] = fminr(in->box.min[a], mb->max[a]); in->box.max[a] = fmaxr(in->box.max[a], mb->max[a]);
            }
        }
    }
    s->camera = xf_from(d->camera.origin, d->camera.basis);
    s->tan_half_fov = tanf(d->camera.yfov * 0.5F); /* camera::set_fov, LIB/scene/camera.cpp:27-30 */
    s->sun_enabled = d->sun.enabled != 0;
    if (s->sun_enabled) {
        float zero[3] = {0, 0, 0};
        xform sx = xf_from(zero, d->sun.basis);
        s->sun_dir = mat_vec(&sx.basis, V(0, 0, 1)); /* basis * fvec3::backward, renderer.cpp:499 */
        s->sun_energy = V(d->sun.energy[0], d->sun.energy[1], d->sun.energy[2]);
        s->sun_radius = d->sun.angular_radius;
    }
    s->environment = V(d->environment_factor[0], d->environment_factor[1], d->environment_factor[2]);
    s->environment_tex = d->environment_tex_plus1 ? d->environment_tex_plus1 - 1 : PTB_NO_TEXTURE;
    s->transparent = d->transparent_background != 0;
    return s;
}

void po_scene_free(po_scene* s) {
    if (!s) return;
    for (uint32_t i = 0; i < s->n_meshes; i++) {
        po_mesh* m = &s->meshes[i];
        free((void*)m->positions); free((void*)m->normals); free((void*)m->tangents); free((void*)m->uvs);
        free((void*)m->indices);
        free_node(m->root);
    }
    for (uint32_t t = 0; t < s->n_textures; t++) free((void*)s->textures[t].pixels);
    free(s->textures);
    free(s->meshes); free(s->surfaces); free(s->materials); free(s->instances);
    free(s);
}

/* same record stream as ptb_scene_dump_kd / ref_dump_kd */
static void dump_rec(const kd_node* n, uint32_t* w, uint64_t cap, uint64_t* pos) {
#define PUT(v) do { if (w && *pos < cap) w[*pos] = (v); (*pos)++; } while (0)
    if (n->axis < 0) {
        PUT(n->count);
        for (uint32_t i = 0; i < n->count; i++) PUT(n->tris[i]);
        return;
    }
    uint32_t bits;
    memcpy(&bits, &n->split, 4);
    PUT(0x80000000u | (uint32_t)n->axis);
    PUT(bits);
    PUT(n->left ? 1u : 0u);
    PUT(n->right ? 1u : 0u);
    if (n->left) dump_rec(n->left, w, cap, pos);
    if (n->right) dump_rec(n->right, w, cap, pos);
#undef PUT
}
int po_dump_kd(const po_scene* s, uint32_t mesh, uint32_t* words, uint64_t capacity, uint64_t* n_words) {
    uint64_t pos = 0;
    dump_rec(s->meshes[mesh].root, words, capacity, &pos);
    *n_words = pos;
    return (words && pos > capacity) ? 1 : 0;
}
void po_mesh_aabb(const po_scene* s, uint32_t mesh, float* out6) {
    for (int a = 0; a < 3; a++) { out6[a] = s->meshes[mesh].box.min[a]; out6[3 + a] = s->meshes[mesh].box.max[a]; }
}

/* ---- intersection ------------------------------------------------------------ */
typedef struct { v3 o, d; } ray;
static ray make_ray(v3 o, v3 d) { ray r = {o, normalize(d)}; return r; } /* geometry::ray::ray, LIB/geometry/ray.cpp:6-8 */

/* aabb::intersect + has_hit, LIB/geometry/aabb.cpp:41-67, :8-10 */
static int slab(const aabb* b, const ray* r, float* near_o, float* far_o) {
    if (b->min[0] > b->max[0] || b->min[1] > b->max[1] || b->min[2] > b->max[2]) return 0;
    v3 inv = divv(V(1, 1, 1), r->d);
    v3 t0 = mulv(sub(V(b->min[0], b->min[1], b->min[2]), r->o), inv);
    v3 t1 = mulv(sub(V(b->max[0], b->max[1], b->max[2]), r->o), inv);
    v3 nd = V(fminr(t0.x, t1.x), fminr(t0.y, t1.y), fminr(t0.z, t1.z));
    v3 fd = V(fmaxr(t0.x, t1.x), fmaxr(t0.y, t1.y), fmaxr(t0.z, t1.z));
    float nr = fmaxr(fmaxr(nd.x, nd.y), nd.z);
    float fr = fminr(fminr(fd.x, fd.y), fd.z);
    if (nr > fr) return 0;
    *near_o = nr;
    *far_o = fr;
    return fr >= 0;
}

/* triangle::intersect, LIB/geometry/triangle.cpp:120-190 */
static float tri_intersect(v3 a, v3 b, v3 c, const ray* r, v3* bary) {
    v3 mx = sub(a, b), my = sub(a, c), mz = r->d;
    v3 v = sub(a, r->o);
    float c1 = my.y * mz.z - mz.y * my.z;
    float c2 = mx.y * mz.z - mz.y * mx.z;
    float c3 = mx.y * my.z - my.y * mx.z;
    float c4 = v.y * mz.z - mz.y * v.z;
    float c5 = mx.y * v.z - v.y * mx.z;
    float c6 = my.y * v.z - v.y * my.z;
    float inv_det = 1 / (mx.x * c1 - my.x * c2 + mz.x * c3);
    float beta = inv_det * (v.x * c1 - my.x * c4 - mz.x * c6);
    if (beta < 0 - EPS || beta > 1 + EPS) return -1;
    float gamma = inv_det * (mx.x * c4 - v.x * c2 + mz.x * c5);
    if (gamma < 0 - EPS || gamma + beta > 1 + EPS) return -1;
    float dist = inv_det * (mx.x * c6 - my.x * c5 + v.x * c3);
    float alpha = 1 - beta - gamma;
    *bary = V(alpha, beta, gamma);
    return dist;
}

typedef struct { float t; v3 bary; uint32_t tri; } mesh_hit;
typedef struct { uint64_t branch, leaf, tri, push, model, surface; } visit_counts;

/* mesh::intersect, LIB/core/mesh.cpp:300-405 */
static mesh_hit mesh_intersect(const po_mesh* m, const ray* r, visit_counts* vc) {
    mesh_hit none = {-1, {0, 0, 0}, 0};
    float nr, fr;
    if (!slab(&m->box, r, &nr, &fr)) return none;
    struct { const kd_node* node; float tmin, tmax; } stack[64];
    int sp = 0;
    stack[sp].node = m->root; stack[sp].tmin = nr; stack[sp].tmax = fr; sp++;
    while (sp > 0) {
        sp--;
        const kd_node* node = stack[sp].node;
        float min_dist = stack[sp].tmin, max_dist = stack[sp].tmax;
        while (node && node->axis >= 0) {
            if (vc) vc->branch++;
            float o = node->axis == 0 ? r->o.x : (node->axis == 1 ? r->o.y : r->o.z);
            float d = node->axis == 0 ? r->d.x : (node->axis == 1 ? r->d.y : r->d.z);
            float split_dist = (node->split - o) / d;
            const kd_node *first, *second;
            if (o < node->split) { first = node->left; second = node->right; }
            else { first = node->right; second = node->left; }
            if (split_dist < 0 || split_dist > max_dist) node = first;
            else if (split_dist < min_dist) node = second;
            else {
                if (second) { stack[sp].node = second; stack[sp].tmin = split_dist; stack[sp].tmax = max_dist; sp++; if (vc) vc->push++; }
                node = first;
                max_dist = split_dist;
            }
        }
        if (!node) continue;
        if (vc) vc->leaf++;
        mesh_hit best = none;
        for (uint32_t i = 0; i < node->count; i++) {
            uint32_t t = node->tris[i];
            const float* pa = m->positions + 3 * (size_t)m->indices[3 * (size_t)t];
            const float* pb = m->positions + 3 * (size_t)m->indices[3 * (size_t)t + 1];
            const float* pc = m->positions + 3 * (size_t)m->indices[3 * (size_t)t + 2];
            v3 bary;
            if (vc) vc->tri++;
            float dist = tri_intersect(V(pa[0], pa[1], pa[2]), V(pb[0], pb[1], pb[2]), V(pc[0], pc[1], pc[2]), r, &bary);
            if (dist >= 0 && dist <= max_dist && (dist < best.t || !(best.t >= 0))) {
                best.t = dist; best.bary = bary; best.tri = t;
            }
        }
        if (!(best.t >= 0)) continue;
        return best;
    }
    return none;
}

typedef struct { float t; v3 bary; uint32_t tri, instance, surface; } scene_hit;

/* renderer::intersect's loop (LIB/core/renderer.cpp:645-671) over model::intersect (LIB/scene/model.cpp:20-72) */
static scene_hit scene_intersect(const po_scene* s, const ray* r, visit_counts* vc) {
    scene_hit nearest = {-1, {0, 0, 0}, 0, PTB_MISS, 0};
    for (uint32_t i = 0; i < s->n_instances; i++) {
        const po_instance* in = &s->instances[i];
        if (vc) vc->model++;
        ray view = make_ray(xf_apply(&in->inv, r->o), mat_vec(&in->inv.basis, r->d)); /* ray::transform, ray.cpp:10-15 */
        float nr, fr;
        if (!slab(&in->box, &view, &nr, &fr)) continue;
        mesh_hit best = {-1, {0, 0, 0}, 0};
        uint32_t best_surface = 0;
        for (uint32_t k = 0; k < in->n_surfaces; k++) {
            if (vc) vc->surface++;
            mesh_hit h = mesh_intersect(&s->meshes[s->surfaces[in->first_surface + k].mesh], &view, vc);
            if (!(h.t >= 0)) continue;
            if (h.t < best.t || !(best.t >= 0)) { best = h; best_surface = k; }
        }
        if (!(best.t >= 0)) continue;
        v3 hit_vec = muls(view.d, best.t); /* model.cpp:62-63 */
        float tw = length(mat_vec(&in->fwd.basis, hit_vec));
        if (!(tw >= 0)) continue;
        if (tw < nearest.t || !(nearest.t >= 0)) {
            nearest.t = tw; nearest.bary = best.bary; nearest.tri = best.tri; nearest.instance = i; nearest.surface = best_surface;
        }
    }
    return nearest;
}

typedef struct { v3 position; float u, v; v3 normal, tangent; uint32_t material; } attrs_t;

/* renderer::intersect, attribute part, LIB/core/renderer.cpp:688-724 */
static attrs_t hit_attrs(const po_scene* s, const scene_hit* h) {
    const po_instance* in = &s->instances[h->instance];
    const ptb_surface_desc* sf = &s->surfaces[in->first_surface + h->surface];
    const po_mesh* m = &s->meshes[sf->mesh];
    uint32_t i0 = m->indices[3 * (size_t)h->tri], i1 = m->indices[3 * (size_t)h->tri + 1], i2 = m->indices[3 * (size_t)h->tri + 2];
#define P3(arr, i) V((arr)[3 * (size_t)(i)], (arr)[3 * (size_t)(i) + 1], (arr)[3 * (size_t)(i) + 2])
#define BLEND(arr) add(add(muls(P3(arr, i0), h->bary.x), muls(P3(arr, i1), h->bary.y)), muls(P3(arr, i2), h->bary.z))
    attrs_t a;
    a.material = sf->material;
    a.position = xf_apply(&in->fwd, BLEND(m->positions));
    a.u = m->uvs[2 * (size_t)i0] * h->bary.x + m->uvs[2 * (size_t)i1] * h->bary.y + m->uvs[2 * (size_t)i2] * h->bary.z;
    a.v = m->uvs[2 * (size_t)i0 + 1] * h->bary.x + m->uvs[2 * (size_t)i1 + 1] * h->bary.y + m->uvs[2 * (size_t)i2 + 1] * h->bary.z;
    a.normal = normalize(mat_vec(&in->normal_mat, BLEND(m->normals)));
    a.tangent = normalize(mat_vec(&in->normal_mat, BLEND(m->tangents)));
#undef BLEND
#undef P3
    return a;
}

/* image::read, LIB/image/image.cpp:124-141 */
static float tex_read(const ptb_texture_desc* T, uint32_t x, uint32_t y, uint32_t c) {
    uint32_t index = y * T->width + x;
    index = index * T->channels + c;
    float value;
    if (T->is_float) value = ((const float*)T->pixels)[index];
    else value = ((const uint8_t*)T->pixels)[index] / 255.0F;
    if (T->srgb && c < 3) value = powf(value, 2.2F);
    return value;
}
typedef struct { float x, y, z, w; } v4;
/* image_texture::read_pixel, LIB/image/image_texture.cpp:47-62 (missing channels stay 1) */
static v4 tex_pixel(const ptb_texture_desc* T, uint32_t x, uint32_t y) {
    v4 c = {1, 1, 1, 1};
    if (T->channels >= 4) c.w = tex_read(T, x, y, 3);
    if (T->channels >= 3) c.z = tex_read(T, x, y, 2);
    if (T->channels >= 2) c.y = tex_read(T, x, y, 1);
    if (T->channels >= 1) c.x = tex_read(T, x, y, 0);
    return c;
}
static v4 lerp4(v4 a, v4 b, float w) {
    v4 r = {lerpf(a.x, b.x, w), lerpf(a.y, b.y, w), lerpf(a.z, b.z, w), lerpf(a.w, b.w, w)};
    return r;
}
/* image_texture::sample, LIB/image/image_texture.cpp:21-45: bilinear, wrap-around */
static v4 tex_sample(const po_scene* s, uint32_t id, float u, float v) {
    const ptb_texture_desc* T = &s->textures[id];
    float cx = u * T->width - 0.5F, cy = (1 - v) * T->height - 0.5F;
    float fx = floorf(cx), fy = floorf(cy), gx = ceilf(cx), gy = ceilf(cy);
    /* uvec2(float) then math::mod(uvec2, size) = (size + x % size) % size, LIB/math/math.inl:190-196 */
#define WRAP(f, size) (((size) + ((uint32_t)(int64_t)(f)) % (size)) % (size))
    uint32_t x0 = WRAP(fx, T->width), x1 = WRAP(gx, T->width), y0 = WRAP(fy, T->height), y1 = WRAP(gy, T->height);
#undef WRAP
    float dx = cx - fx, dy = cy - fy; /* math::fract */
    v4 t = lerp4(tex_pixel(T, x0, y0), tex_pixel(T, x1, y0), dx);
    v4 b = lerp4(tex_pixel(T, x0, y1), tex_pixel(T, x1, y1), dx);
    return lerp4(t, b, dy);
}

/* material::get_normal, LIB/core/material.cpp:6-11 */
static v3 material_normal(const po_scene* s, uint32_t mat, float u, float v) {
    const ptb_material_desc* m = &s->materials[mat];
    if (m->normal_tex != PTB_NO_TEXTURE) {
        v4 c = tex_sample(s, m->normal_tex, u, v);
        return sub(muls(V(c.x, c.y, c.z), 2), V(1, 1, 1));
    }
    return V(0, 0, 1);
}

/* intersect_result::get_normal, LIB/core/renderer.cpp:430-435 */
static v3 shading_normal_s(const po_scene* s, const attrs_t* a) {
    v3 binormal = cross(a->normal, a->tangent);
    m3 tbn = {a->tangent, binormal, a->normal};
    return mat_vec(&tbn, material_normal(s, a->material, a->u, a->v));
}

void po_trace_rays(const po_scene* s, const float* od, uint64_t n, ptb_hit* hits, float* attrs) {
    for (uint64_t i = 0; i < n; i++) {
        ray r = make_ray(V(od[6 * i], od[6 * i + 1], od[6 * i + 2]), V(od[6 * i + 3], od[6 * i + 4], od[6 * i + 5]));
        scene_hit h = scene_intersect(s, &r, NULL);
        ptb_hit* o = &hits[i];
        if (!(h.t >= 0)) {
            o->instance = o->surface = o->triangle = PTB_MISS;
            o->t = -1;
            o->bary[0] = o->bary[1] = o->bary[2] = 0;
        } else {
            o->instance = h.instance; o->surface = h.surface; o->triangle = h.tri; o->t = h.t;
            o->bary[0] = h.bary.x; o->bary[1] = h.bary.y; o->bary[2] = h.bary.z;
        }
        if (attrs) {
            float* a = attrs + 14 * i;
            if (!(h.t >= 0)) { for (int k = 0; k < 14; k++) a[k] = 0; continue; }
            attrs_t at = hit_attrs(s, &h);
            v3 sn = shading_normal_s(s, &at);
            a[0] = at.position.x; a[1] = at.position.y; a[2] = at.position.z; a[3] = at.u; a[4] = at.v;
            a[5] = at.normal.x; a[6] = at.normal.y; a[7] = at.normal.z;
            a[8] = at.tangent.x; a[9] = at.tangent.y; a[10] = at.tangent.z;
            a[11] = sn.x; a[12] = sn.y; a[13] = sn.z;
        }
    }
}

void po_count_visits(const po_scene* s, const float* od, uint64_t n, uint64_t* c6) {
    visit_counts vc;
    memset(&vc, 0, sizeof(vc));
    for (uint64_t i = 0; i < n; i++) {
        ray r = make_ray(V(od[6 * i], od[6 * i + 1], od[6 * i + 2]), V(od[6 * i + 3], od[6 * i + 4], od[6 * i + 5]));
        scene_intersect(s, &r, &vc);
    }
    c6[0] = vc.model; c6[1] = vc.surface; c6[2] = vc.branch; c6[3] = vc.leaf; c6[4] = vc.tri; c6[5] = vc.push;
}

/* renderer::render's ray set-up (LIB/core/renderer.cpp:365-370) + camera::get_ray (LIB/scene/camera.cpp:10-21) */
static ray camera_ray(const po_scene* s, uint32_t px, uint32_t py, float aax, float aay, uint32_t rw, uint32_t rh) {
    float ndcx = (((float)px + aax) / (float)rw) * 2 - 1.0f;
    float ndcy = (((float)py + aay) / (float)rh) * 2 - 1.0f;
    ndcy = -ndcy;
    float ratio = (float)rw / (float)rh;
    float dx = s->tan_half_fov * ndcx, dy = s->tan_half_fov * ndcy;
    dx *= ratio;
    ray local = make_ray(V(0, 0, 0), V(dx, dy, -1));
    return make_ray(xf_apply(&s->camera, local.o), mat_vec(&s->camera.basis, local.d));
}

void po_camera_rays(const po_scene* s, uint32_t w, uint32_t h, const uint32_t* px, const uint32_t* py, const float* aa,
                    uint64_t n, float* od) {
    for (uint64_t i = 0; i < n; i++) {
        ray r = camera_ray(s, px[i], py[i], aa[2 * i], aa[2 * i + 1], w, h);
        od[6 * i] = r.o.x; od[6 * i + 1] = r.o.y; od[6 * i + 2] = r.o.z;
        od[6 * i + 3] = r.d.x; od[6 * i + 4] = r.d.y; od[6 * i + 5] = r.d.z;
    }
}

/* ---- BSDF kit: LIB/core/pbr.cpp, LIB/util/rand_cone_vec.cpp ------------------- */
#define PI_D 3.14159265358979323846 /* math::pi is a double (LIB/math/math.hpp:18) */
#define SQRT3_D 1.7320508075688772

static float fresnel_schlick(v3 o, v3 i, float ior) { /* pbr.cpp:13-25 */
    v3 h = normalize(add(o, i));
    float cos_theta = dot(o, h);
    float f0 = (ior - 1) / (ior + 1);
    f0 *= f0;
    return lerpf(f0, 1, (float)pow(1 - cos_theta, 5)); /* math::pow<float,int>: std::pow in double, narrowed */
}
static v3 rand_cone_vec(float rnd, float cos_theta, v3 normal) { /* rand_cone_vec.cpp:8-35 */
    float phi = (float)(rnd * 2 * PI_D);
    float sin_theta = sqrtf(1 - cos_theta * cos_theta);
    v3 cone = V(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
    v3 helper = V(0, 0, 0);
    if (fabsf(normal.x) < (1 / SQRT3_D)) helper.x = 1;
    else if (fabsf(normal.y) < (1 / SQRT3_D)) helper.y = 1;
    else helper.z = 1;
    v3 tangent = normalize(cross(normal, helper));
    v3 binormal = cross(normal, tangent);
    m3 tbn = {tangent, binormal, normal};
    return mat_vec(&tbn, cone);
}
static v3 importance_lambert(float r0, float r1, v3 n) { /* pbr.cpp:71-77 */
    float theta = acosf(2 * r0 - 1) * 0.5F;
    return rand_cone_vec(r1, cosf(theta), n);
}
static v3 importance_ggx(float r0, float r1, v3 n, v3 o, float roughness) { /* pbr.cpp:79-91 */
    roughness *= roughness;
    roughness *= roughness;
    float cos_theta = sqrtf((1 - r0) / (1 + (roughness - 1) * r0));
    v3 h = rand_cone_vec(r1, cos_theta, n);
    return reflect(neg(o), h);
}
static float smith_g1(v3 n, v3 l, float k) { /* pbr.cpp:95-102 */
    float cos_theta = dot(n, l);
    return cos_theta / fmaxr(lerpf(k, 1, cos_theta), EPS);
}
static float geometry_smith(v3 n, v3 o, v3 i, float roughness) { /* pbr.cpp:104-114 */
    float r = roughness + 1;
    float k = (r * r) / 8;
    return smith_g1(n, o, k) * smith_g1(n, i, k);
}
static float pdf_diffuse(v3 n, v3 i) { return (float)(dot(n, i) / PI_D); } /* pbr.cpp:118-123 */
static float distribution_ggx(v3 n, v3 o, v3 i, float roughness) { /* pbr.cpp:125-140 */
    roughness *= roughness;
    roughness *= roughness;
    v3 h = normalize(add(o, i));
    float cos_phi = dot(n, h);
    float denom = lerpf(1, roughness, cos_phi * cos_phi);
    float cos_theta = dot(n, i);
    double dd = PI_D * denom * denom;
    double mx = (double)EPS > dd ? (double)EPS : dd; /* math::max(double, float) */
    return (float)(cos_theta * roughness / mx);
}
static float pdf_specular(v3 n, v3 o, v3 i, float roughness) { /* pbr.cpp:172-184 */
    float dist = distribution_ggx(n, o, i, roughness);
    float geo = geometry_smith(n, o, i, roughness);
    float n_dot_o = dot(n, o), n_dot_i = dot(n, i);
    return (dist * geo) / fmaxr(4 * n_dot_o * n_dot_i, EPS);
}
static void eval_brdf(v3 n, v3 o, v3 i, v3 albedo, float roughness, float metallic, float sp, v3* brdf, float* pdf) {
    /* LIB/core/renderer.cpp:581-606 (and :523-552) */
    float diffuse_pdf = pdf_diffuse(n, i);
    v3 diffuse_brdf = smul(diffuse_pdf, albedo);
    float specular_pdf = pdf_specular(n, o, i, roughness);
    v3 specular_brdf = V(specular_pdf, specular_pdf, specular_pdf);
    v3 fresnel = lerp3s(V(0.04F, 0.04F, 0.04F), albedo, metallic);
    {
        v3 h = normalize(add(o, i));
        float cos_theta = dot(o, h);
        fresnel = lerp3s(fresnel, V(1, 1, 1), (float)pow(1 - cos_theta, 5));
    }
    diffuse_brdf = lerp3s(diffuse_brdf, V(0, 0, 0), metallic);
    *brdf = lerp3v(diffuse_brdf, specular_brdf, fresnel);
    *pdf = lerpf(diffuse_pdf, specular_pdf, sp);
}
static v3 clamp3(v3 x, v3 lo, v3 hi) { return V(clampr(x.x, lo.x, hi.x), clampr(x.y, lo.y, hi.y), clampr(x.z, lo.z, hi.z)); }
static int is_approx(float a, float b) { return a == b || fabsf(a - b) < EPS; } /* math.inl:49-52 */

/* ---- RNG: the reference uses an unseeded thread_local mt19937 (LIB/core/utils.hpp:8-13);
 * any good uniform [0,1) generator is statistically equivalent.  xoshiro128+ here. */
typedef struct { uint32_t s[4]; uint64_t rays; } rng_t;
static inline uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
static inline float rnd(rng_t* g) {
    uint32_t r = g->s[0] + g->s[3], t = g->s[1] << 9;
    g->s[2] ^= g->s[0]; g->s[3] ^= g->s[1]; g->s[1] ^= g->s[2]; g->s[0] ^= g->s[3];
    g->s[2] ^= t; g->s[3] = rotl(g->s[3], 11);
    return (r >> 8) * (1.0f / 16777216.0f);
}
static void rng_seed(rng_t* g, uint64_t seed) {
    for (int i = 0; i < 4; i++) { /* splitmix64 */
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        g->s[i] = (uint32_t)((z ^ (z >> 31)) >> 16);
    }
    if (!(g->s[0] | g->s[1] | g->s[2] | g->s[3])) g->s[0] = 1;
    g->rays = 0;
}

typedef struct { v3 rgb; float alpha; } rgba_t;
typedef struct { v3 albedo; float opacity, roughness, metallic; v3 emissive; float ior; int shadow_catcher; } mat_t;
static mat_t material_of(const po_scene* s, uint32_t id, float u, float v) { /* LIB/core/material.cpp:13-53 */
    const ptb_material_desc* m = &s->materials[id];
    mat_t r;
    r.albedo = V(m->albedo[0], m->albedo[1], m->albedo[2]);
    if (m->albedo_tex != PTB_NO_TEXTURE) { v4 c = tex_sample(s, m->albedo_tex, u, v); r.albedo = mulv(r.albedo, V(c.x, c.y, c.z)); }
    r.opacity = m->opacity;
    if (m->opacity_tex != PTB_NO_TEXTURE) r.opacity *= tex_sample(s, m->opacity_tex, u, v).w;
    r.roughness = m->roughness;
    if (m->roughness_tex != PTB_NO_TEXTURE) r.roughness *= tex_sample(s, m->roughness_tex, u, v).y;
    r.metallic = m->metallic;
    if (m->metallic_tex != PTB_NO_TEXTURE) r.metallic *= tex_sample(s, m->metallic_tex, u, v).z;
    r.emissive = V(m->emissive[0], m->emissive[1], m->emissive[2]);
    if (m->emissive_tex != PTB_NO_TEXTURE) { v4 c = tex_sample(s, m->emissive_tex, u, v); r.emissive = mulv(r.emissive, V(c.x, c.y, c.z)); }
    r.emissive = muls(r.emissive, 10); /* renderer.cpp:462 */
    r.ior = m->ior; r.shadow_catcher = m->shadow_catcher != 0;
    return r;
}

static v3 sun_sample(const po_scene* s, rng_t* g) { /* renderer.cpp:499-501 */
    float r0 = rnd(g), r1 = rnd(g);
    return rand_cone_vec(r0, cosf(r1 * s->sun_radius), s->sun_dir);
}
static v3 direct_term(const po_scene* s, v3 n, v3 o, v3 l, const mat_t* m, float roughness, float sp) { /* renderer.cpp:521-556 */
    v3 brdf;
    float pdf;
    eval_brdf(n, o, l, m->albedo, roughness, m->metallic, sp, &brdf, &pdf);
    pdf = lerpf(1, 1, sp);
    v3 out = divs(mulv(brdf, s->sun_energy), fmaxr(pdf, EPS));
    return clamp3(out, V(0, 0, 0), s->sun_energy);
}

/* What a ray that hits nothing receives: renderer.cpp:446-450 / worker.cpp:308-313 with equirectangular_proj
 * (LIB/core/utils.hpp:22-27). */
static v3 environment_of(const po_scene* s, v3 dir) {
    if (s->environment_tex == PTB_NO_TEXTURE) return s->environment;
    float u = atan2f(dir.z, dir.x) * 0.1591F + 0.5F, v = asinf(dir.y) * 0.3183F + 0.5F;
    v4 c = tex_sample(s, s->environment_tex, u, v);
    return mulv(V(c.x, c.y, c.z), s->environment);
}

/* core::renderer::trace, LIB/core/renderer.cpp:437-643 — recursive, as written. */
static rgba_t trace_lib(const po_scene* s, uint32_t bounce, uint32_t bounce_count, ray r, rng_t* g) {
    rgba_t future = {{0, 0, 0}, 1}; /* fvec4::future */
    if (bounce == 0) return future;
    g->rays++;
    scene_hit h = scene_intersect(s, &r, NULL);
    if (!(h.t >= 0)) { rgba_t e = {environment_of(s, r.d), s->transparent ? 0.0f : 1.0f}; return e; }
    attrs_t at = hit_attrs(s, &h);
    mat_t m = material_of(s, at.material, at.u, at.v);
    if (!is_approx(m.opacity, 1) && rnd(g) > m.opacity)
        return trace_lib(s, bounce, bounce_count, make_ray(add(at.position, muls(r.d, EPS)), r.d), g);
    v3 normal = shading_normal_s(s, &at);
    v3 outcoming = neg(r.d);
    if (dot(normal, outcoming) <= 0) return future;
    float roughness = fmaxr(m.roughness, 0.05F);
    float sp = fmaxr(fresnel_schlick(outcoming, reflect(neg(outcoming), normal), m.ior), m.metallic);
    int specular_sample = rnd(g) < sp;
    v3 direct_out = V(0, 0, 0);
    if (s->sun_enabled) {
        v3 l = sun_sample(s, g);
        if (dot(normal, l) > 0) {
            ray dr = make_ray(add(at.position, muls(l, EPS)), l);
            g->rays++;
            scene_hit dh = scene_intersect(s, &dr, NULL);
            if (!(dh.t >= 0)) {
                if (m.shadow_catcher && bounce == bounce_count)
                    return trace_lib(s, bounce, bounce_count, make_ray(add(at.position, muls(r.d, EPS)), r.d), g);
                direct_out = direct_term(s, normal, outcoming, l, &m, roughness, sp);
            } else if (m.shadow_catcher && bounce == bounce_count) {
                return future;
            }
        }
    }
    v3 indirect_out = V(0, 0, 0);
    float r0 = rnd(g), r1 = rnd(g);
    v3 inc = specular_sample ? importance_ggx(r0, r1, normal, outcoming, roughness) : importance_lambert(r0, r1, normal);
    if (dot(normal, inc) > 0) {
        v3 brdf;
        float pdf;
        eval_brdf(normal, outcoming, inc, m.albedo, roughness, m.metallic, sp, &brdf, &pdf);
        rgba_t in = trace_lib(s, bounce - 1, bounce_count, make_ray(add(at.position, muls(inc, EPS)), inc), g);
        indirect_out = divs(mulv(brdf, in.rgb), fmaxr(pdf, EPS));
        indirect_out = clamp3(indirect_out, V(0, 0, 0), in.rgb);
    }
    rgba_t out = {add(add(direct_out, indirect_out), m.emissive), 1};
    return out;
}

/* processors::worker::trace_iter, APP/processors/worker/worker.cpp:285-514 */
static rgba_t trace_app(const po_scene* s, uint32_t initial_bounce, ray cur, rng_t* g) {
    uint32_t bounce_remaining = initial_bounce;
    v3 acc = V(0, 0, 0), thr = V(1, 1, 1);
    float alpha = s->transparent ? 0.0f : 1.0f;
    rgba_t future = {{0, 0, 0}, 1};
    while (bounce_remaining > 0) {
        g->rays++;
        scene_hit h = scene_intersect(s, &cur, NULL);
        if (!(h.t >= 0)) {
            acc = add(acc, mulv(thr, environment_of(s, cur.d)));
            alpha = s->transparent ? 0.0f : 1.0f;
            break;
        }
        alpha = 1.0f;
        attrs_t at = hit_attrs(s, &h);
        mat_t m = material_of(s, at.material, at.u, at.v);
        acc = add(acc, mulv(thr, m.emissive));
        if (!is_approx(m.opacity, 1) && rnd(g) > m.opacity) {
            cur = make_ray(add(at.position, muls(cur.d, EPS)), cur.d);
            continue;
        }
        v3 normal = shading_normal_s(s, &at);
        v3 outcoming = neg(cur.d);
        if (dot(normal, outcoming) <= 0) break;
        if (m.shadow_catcher && bounce_remaining == initial_bounce) {
            int in_shadow = 1;
            if (s->sun_enabled) {
                v3 l = sun_sample(s, g);
                if (dot(normal, l) > 0) {
                    ray sr = make_ray(add(at.position, muls(l, EPS)), l);
                    g->rays++;
                    scene_hit sh = scene_intersect(s, &sr, NULL);
                    if (!(sh.t >= 0)) in_shadow = 0;
                }
            }
            if (in_shadow) return future;
            cur = make_ray(add(at.position, muls(cur.d, EPS)), cur.d);
            continue;
        }
        float roughness = fmaxr(m.roughness, 0.05F);
        float sp = fmaxr(fresnel_schlick(outcoming, reflect(neg(outcoming), normal), m.ior), m.metallic);
        int specular_sample = rnd(g) < sp;
        if (s->sun_enabled) {
            v3 l = sun_sample(s, g);
            if (dot(normal, l) > 0) {
                ray dr = make_ray(add(at.position, muls(l, EPS)), l);
                g->rays++;
                scene_hit dh = scene_intersect(s, &dr, NULL);
                if (!(dh.t >= 0)) acc = add(acc, mulv(thr, direct_term(s, normal, outcoming, l, &m, roughness, sp)));
            }
        }
        float r0 = rnd(g), r1 = rnd(g);
        v3 inc = specular_sample ? importance_ggx(r0, r1, normal, outcoming, roughness) : importance_lambert(r0, r1, normal);
        if (dot(normal, inc) > 0) {
            v3 brdf;
            float pdf;
            eval_brdf(normal, outcoming, inc, m.albedo, roughness, m.metallic, sp, &brdf, &pdf);
            thr = mulv(thr, divs(brdf, fmaxr(pdf, EPS)));
            thr = clamp3(thr, V(0, 0, 0), V(10.0f, 10.0f, 10.0f));
            cur = make_ray(add(at.position, muls(inc, EPS)), inc);
            if ((int)bounce_remaining < (int)initial_bounce - 2) {
                float p = fmaxr(thr.x, fmaxr(thr.y, thr.z));
                if (rnd(g) > p) break;
                thr = divs(thr, p);
            }
        } else {
            break;
        }
        bounce_remaining--;
    }
    rgba_t out = {acc, alpha};
    return out;
}

/* ---- render loop: LIB/core/renderer.cpp:354-407 (rows on threads, samples sequential per pixel) */
typedef struct {
    const po_scene* s;
    uint32_t full_w, full_h, x0, y0, w, h, spp, depth;
    int mode, first_unjittered;
    uint64_t seed;
    float *rgb, *alpha;
    volatile uint32_t* next_row;
    uint64_t rays;
    int tid;
} job_t;

static void* render_rows(void* arg) {
    job_t* j = (job_t*)arg;
    rng_t g;
    uint64_t rays = 0;
    for (;;) {
        uint32_t row = __sync_fetch_and_add(j->next_row, 1);
        if (row >= j->h) break;
        /* one stream per image row (frame coordinates), not per thread: the image is a function of the seed alone,
           whatever the thread count and however the rows fall to the threads */
        rng_seed(&g, j->seed * 0x100000001B3ull + (uint64_t)(j->y0 + row));
        for (uint32_t col = 0; col < j->w; col++) {
            v3 color = V(0, 0, 0);
            float alpha = 0;
            int claimed = 0;
            for (uint32_t sample = 0; sample < j->spp; sample++) {
                float aax = 0, aay = 0;
                if (!(sample == 0 && j->first_unjittered)) { aax = rnd(&g); aay = rnd(&g); }
                ray r = camera_ray(j->s, j->x0 + col, j->y0 + row, aax, aay, j->full_w, j->full_h);
                rgba_t d = j->mode == 1 ? trace_app(j->s, j->depth, r, &g) : trace_lib(j->s, j->depth, j->depth, r, &g);
                if (j->s->transparent) { /* renderer.cpp:374-393 */
                    if (d.alpha > 0.5 && !claimed) { color = d.rgb; alpha = (float)(1 / (sample + 1)); claimed = 1; continue; }
                    else if (d.alpha < 0.5 && claimed) { alpha = alpha * sample + d.alpha; alpha /= sample + 1; continue; }
                    else if (d.alpha < 0.5) continue;
                }
                color = add(muls(color, (float)sample), d.rgb); /* renderer.cpp:396-399 */
                color = divs(color, (float)(sample + 1));
                alpha = alpha * sample + d.alpha;
                alpha /= sample + 1;
            }
            size_t i = (size_t)row * j->w + col;
            j->rgb[3 * i] = color.x; j->rgb[3 * i + 1] = color.y; j->rgb[3 * i + 2] = color.z;
            if (j->alpha) j->alpha[i] = alpha;
        }
        rays += g.rays;
    }
    j->rays = rays;
    return NULL;
}

void po_render_linear(const po_scene* s, uint32_t full_w, uint32_t full_h, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h,
                      uint32_t spp, uint32_t depth, int mode, int first_unjittered, uint64_t seed, int threads, float* rgb,
                      float* alpha, uint64_t* rays_out, double* seconds_out) {
    if (threads <= 0) threads = 1;
    if (threads > 256) threads = 256;
    volatile uint32_t next_row = 0;
    job_t jobs[256];
    pthread_t tids[256];
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        job_t j = {s, full_w, full_h, x0, y0, w, h, spp, depth, mode, first_unjittered, seed, rgb, alpha, &next_row, 0, t};
        jobs[t] = j;
        if (t > 0) pthread_create(&tids[t], NULL, render_rows, &jobs[t]);
    }
    render_rows(&jobs[0]);
    uint64_t rays = jobs[0].rays;
    for (int t = 1; t < threads; t++) { pthread_join(tids[t], NULL); rays += jobs[t].rays; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (rays_out) *rays_out = rays;
    if (seconds_out) *seconds_out = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* tonemap_approx_aces (LIB/core/utils.hpp:29-36) + image::write (LIB/image/image.cpp:143-154) */
static float aces(float hdr) {
    const float a = 2.51F, b = 0.03F, c = 2.43F, d = 0.59F, e = 0.14F;
    return clampr((hdr * (a * hdr + b)) / (hdr * (c * hdr + d) + e), 0, 1);
}
void po_tonemap_rgba8(const float* rgb, const float* alpha, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) out[4 * i + c] = (uint8_t)(powf(aces(rgb[3 * i + c]), 1 / 2.2F) * 255 + 0.5F);
        out[4 * i + 3] = (uint8_t)((alpha ? alpha[i] : 1.0f) * 255 + 0.5F);
    }
}
