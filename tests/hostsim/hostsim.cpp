// tests/hostsim/hostsim.cpp — DEBUGGING AID, NOT A PRODUCT PATH.
//
// The build container has no GPU.  To find logic errors before spending GPU minutes, this file
// compiles the per-item bodies of the CUDA kernels (lightmetrica-v3_b200/csrc/*.cuh, all plain
// `inline` functions when __CUDACC__ is undefined) with g++ and drives them with sequential loops
// that stand in for the kernel grids (std::sort stands in for the CUB radix sort).  It is built
// only by tests/test_hostsim.py into tests/hostsim/_build/, is never loaded by lmcu/, bench.py
// or the plugin, and liblmcu.so contains none of it: the product has no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../lightmetrica-v3_b200/csrc/prep.hpp"
#include "../../lightmetrica-v3_b200/csrc/bdpt.cuh"

using namespace lmcu;

namespace {

struct SimScene {
    PreparedScene ps;
    std::vector<double> mesh_p, mesh_n;
    std::vector<WideNode> nodes;
    std::vector<TriF32> t32;
    std::vector<TriF64> t64;
    std::vector<TriShade> shade;
    std::vector<TriUV> uv;
    std::vector<TriObj> tobj;            // two-level layout
    std::vector<InstDev> inst;
    bool two_level = false;
    bool fp32 = false;
    float bounds[6] = { 0, 0, 0, 0, 0, 0 };
    DevScene dev() const {
        DevScene d{};
        d.nodes = nodes.data(); d.t32 = fp32 ? t32.data() : nullptr; d.t64 = fp32 || two_level ? nullptr : t64.data();
        if (two_level) {
            d.tobj = tobj.data(); d.inst = inst.data(); d.mesh_p = mesh_p.data(); d.mesh_n = mesh_n.data();
            d.mesh_t = ps.mesh_t.empty() ? nullptr : ps.mesh_t.data();
        }
        d.shade = shade.data(); d.prims = ps.prims.data(); d.mats = ps.mats.data(); d.lights = ps.lights.data();
        d.emit = ps.emit.data(); d.cdf = ps.cdf.data(); d.env_tex = ps.env_tex.data();
        d.num_prims = (int)ps.prims.size(); d.num_lights = (int)ps.lights.size(); d.env_light = ps.env_light;
        d.num_tris = ps.nt; d.cam = ps.cam;
        d.uv = uv.empty() ? nullptr : uv.data(); d.textures = ps.textures.data(); d.texels = ps.texels.data(); d.mix = ps.mix.data();
        return d;
    }
};

std::string g_err;

}

template <typename TriT>
static int intersect_t(const SimScene& s, const LmcuRay* rays, long long n, LmcuHit* hits, double* npr, double* tpr) {
    const DevScene d = s.dev();
    uint2 stack[kStackSize];
    unsigned long long cn = 0, ct = 0;
    int overflow = 0;
    for (long long i = 0; i < n; i++) {
        RayD r;
        r.o = d3(rays[i].o[0], rays[i].o[1], rays[i].o[2]); r.d = d3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        r.tmin = rays[i].tmin; r.tmax = rays[i].tmax;
        HitD h; h.t = r.tmax; h.u = h.v = 0; h.tri = -1;
        TraverseCounters tc; tc.nodes = tc.tris = 0;
        if (!traverse<false, true, TriT>(d.nodes, scene_tris<TriT>(d), r, h, stack, &tc, d.inst)) overflow = 1;
        cn += tc.nodes; ct += tc.tris;
        int32_t prim = -1, face = -1;
        if (h.tri >= 0) {
            int32_t pi;
            tri_ids(d, (uint32_t)h.tri, pi, face);
            prim = d.prims[pi].node_index;
        }
        hits[i] = LmcuHit{ h.t, h.u, h.v, prim, face };
    }
    if (npr) *npr = (double)cn / (double)n;
    if (tpr) *tpr = (double)ct / (double)n;
    return overflow;
}


// EXPERIMENT (builder quality probe, HOSTSIM_BUILDER=sah): a top-down binned-SAH binary tree in the BinNode format of
// build.cuh, to measure what tree quality is worth in node visits per ray before writing a device builder for it.
namespace {
struct SahCtx {
    std::vector<BinNode>* bin; std::vector<uint32_t>* perm; const float* lo; const float* hi; const double* centre; uint32_t n; uint32_t next = 0;
};
uint32_t sah_build(SahCtx& c, uint32_t b, uint32_t e) {
    if (e - b == 1) {
        const uint32_t id = c.n - 1 + b, g = (*c.perm)[b];
        BinNode leaf;
        for (int a = 0; a < 3; a++) { leaf.lo[a] = c.lo[(size_t)g * 3 + a]; leaf.hi[a] = c.hi[(size_t)g * 3 + a]; }
        leaf.left = leaf.right = 0xffffffffu; leaf.first = b; leaf.count = 1;
        (*c.bin)[id] = leaf;
        return id;
    }
    const uint32_t id = c.next++;
    double cl[3] = { 1e300, 1e300, 1e300 }, ch[3] = { -1e300, -1e300, -1e300 };
    for (uint32_t i = b; i < e; i++) for (int a = 0; a < 3; a++) {
        const double v = c.centre[(size_t)(*c.perm)[i] * 3 + a];
        cl[a] = std::min(cl[a], v); ch[a] = std::max(ch[a], v);
    }
    const int NB = 16;
    int best_axis = -1, best_bin = -1;
    double best_cost = 1e300;
    for (int a = 0; a < 3; a++) {
        if (!(ch[a] > cl[a])) continue;
        float blo[NB][3], bhi[NB][3]; uint32_t cnt[NB];
        for (int k = 0; k < NB; k++) { cnt[k] = 0; for (int d = 0; d < 3; d++) { blo[k][d] = 3e38f; bhi[k][d] = -3e38f; } }
        const double sc = NB / (ch[a] - cl[a]);
        for (uint32_t i = b; i < e; i++) {
            const uint32_t g = (*c.perm)[i];
            int k = (int)((c.centre[(size_t)g * 3 + a] - cl[a]) * sc); k = k < 0 ? 0 : (k > NB - 1 ? NB - 1 : k);
            cnt[k]++;
            for (int d = 0; d < 3; d++) { blo[k][d] = std::min(blo[k][d], c.lo[(size_t)g * 3 + d]); bhi[k][d] = std::max(bhi[k][d], c.hi[(size_t)g * 3 + d]); }
        }
        float rlo[NB][3], rhi[NB][3]; uint32_t rc[NB];
        float l[3] = { 3e38f, 3e38f, 3e38f }, h[3] = { -3e38f, -3e38f, -3e38f }; uint32_t n = 0;
        for (int k = NB - 1; k >= 1; k--) {
            for (int d = 0; d < 3; d++) { l[d] = std::min(l[d], blo[k][d]); h[d] = std::max(h[d], bhi[k][d]); }
            n += cnt[k];
            for (int d = 0; d < 3; d++) { rlo[k][d] = l[d]; rhi[k][d] = h[d]; } rc[k] = n;
        }
        for (int d = 0; d < 3; d++) { l[d] = 3e38f; h[d] = -3e38f; } n = 0;
        for (int k = 0; k < NB - 1; k++) {
            for (int d = 0; d < 3; d++) { l[d] = std::min(l[d], blo[k][d]); h[d] = std::max(h[d], bhi[k][d]); }
            n += cnt[k];
            if (n == 0 || rc[k + 1] == 0) continue;
            const double cost = (double)box_area(l, h) * n + (double)box_area(rlo[k + 1], rhi[k + 1]) * rc[k + 1];
            if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
        }
    }
    uint32_t mid;
    if (best_axis < 0) mid = (b + e) / 2;
    else {
        const double sc = NB / (ch[best_axis] - cl[best_axis]);
        auto it = std::partition(c.perm->begin() + b, c.perm->begin() + e, [&](uint32_t g) {
            int k = (int)((c.centre[(size_t)g * 3 + best_axis] - cl[best_axis]) * sc); k = k < 0 ? 0 : (k > NB - 1 ? NB - 1 : k);
            return k <= best_bin;
        });
        mid = (uint32_t)(it - c.perm->begin());
        if (mid == b || mid == e) mid = (b + e) / 2;
    }
    const uint32_t L = sah_build(c, b, mid), R = sah_build(c, mid, e);
    BinNode nd;
    for (int a = 0; a < 3; a++) { nd.lo[a] = fminf((*c.bin)[L].lo[a], (*c.bin)[R].lo[a]); nd.hi[a] = fmaxf((*c.bin)[L].hi[a], (*c.bin)[R].hi[a]); }
    nd.left = L; nd.right = R; nd.first = b; nd.count = e - b;
    (*c.bin)[id] = nd;
    return id;
}
}


// morton -> sort -> hierarchy -> fit -> collapse over n items with float bounds lo / hi and centres (the middle of
// lmcu_accel_build, shared by the flattened tree, the per-mesh trees and the tree over the instances)
static bool build_tree_host(uint32_t nt, const std::vector<float>& tri_lo, const std::vector<float>& tri_hi, const std::vector<double>& centre,
                            int max_leaf, std::vector<WideNode>& wide, std::vector<uint32_t>& tri_order) {
    float slo[3] = { 3e38f, 3e38f, 3e38f }, shi[3] = { -3e38f, -3e38f, -3e38f };
    for (uint32_t g = 0; g < nt; g++)
        for (int a = 0; a < 3; a++) { slo[a] = fminf(slo[a], tri_lo[(size_t)g * 3 + a]); shi[a] = fmaxf(shi[a], tri_hi[(size_t)g * 3 + a]); }
    // k_morton + sort
    const double dlo[3] = { slo[0], slo[1], slo[2] }, dhi[3] = { shi[0], shi[1], shi[2] };
    std::vector<uint64_t> keys(nt);
    std::vector<uint32_t> vals(nt);
    for (uint32_t g = 0; g < nt; g++) { keys[g] = morton_body(centre.data(), g, dlo, dhi, tri_lo.data(), tri_hi.data(), 1.0 / 16.0); vals[g] = g; }
    std::vector<uint32_t> order(nt);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> keys_sorted(nt);
    std::vector<uint32_t> vals_sorted(nt);
    for (uint32_t i = 0; i < nt; i++) { keys_sorted[i] = keys[order[i]]; vals_sorted[i] = vals[order[i]]; }
    // k_hierarchy
    std::vector<BinNode> bin((size_t)2 * nt);
    std::vector<uint32_t> parent((size_t)2 * nt, 0xffffffffu), arrive(nt, 0);
    for (int i = 0; i + 1 < (int)nt; i++) hierarchy_body(i, keys_sorted.data(), (int)nt, bin.data(), parent.data());
    const char* builder = getenv("HOSTSIM_BUILDER");
    if (builder && std::string(builder) == "sah" && nt > 1) {
        std::iota(vals_sorted.begin(), vals_sorted.end(), 0u);
        SahCtx c{ &bin, &vals_sorted, tri_lo.data(), tri_hi.data(), centre.data(), nt };
        sah_build(c, 0, nt);
    }
    else
    // k_fit (sequential stand-in for the atomic-arrival climb)
    for (int j = 0; j < (int)nt; j++) {
        const uint32_t g = vals_sorted[j];
        BinNode leaf;
        for (int a = 0; a < 3; a++) { leaf.lo[a] = tri_lo[(size_t)g * 3 + a]; leaf.hi[a] = tri_hi[(size_t)g * 3 + a]; }
        leaf.left = leaf.right = 0xffffffffu; leaf.first = (uint32_t)j; leaf.count = 1;
        const uint32_t id = (uint32_t)(nt - 1 + j);
        bin[id] = leaf;
        if (nt == 1) break;
        uint32_t cur = parent[id];
        while (true) {
            if (arrive[cur]++ == 0) break;
            BinNode& nd = bin[cur];
            for (int a = 0; a < 3; a++) {
                nd.lo[a] = fminf(bin[nd.left].lo[a], bin[nd.right].lo[a]);
                nd.hi[a] = fmaxf(bin[nd.left].hi[a], bin[nd.right].hi[a]);
            }
            if (cur == 0) break;
            cur = parent[cur];
        }
    }
    // k_collapse, level by level
    wide.assign(nt, WideNode{});
    tri_order.assign(nt, 0u);
    std::vector<CollapseItem> queue[2];
    queue[0].resize(nt); queue[1].resize(nt);
    uint32_t counters[3] = { 1u, 0u, 0u };
    queue[0][0] = CollapseItem{ 0u, 0u };
    uint32_t count = 1;
    int cur = 0;
    while (count > 0) {
        CollapseCtx c;
        c.bin = bin.data(); c.sorted_tri = vals_sorted.data(); c.n = (int)nt; c.max_leaf = max_leaf; c.wide = wide.data();
        c.width = getenv("HOSTSIM_WIDTH") ? atoi(getenv("HOSTSIM_WIDTH")) : 8;
        c.tri_order = tri_order.data(); c.node_counter = &counters[0]; c.tri_counter = &counters[1];
        c.next_queue = queue[cur ^ 1].data(); c.next_count = &counters[2];
        for (uint32_t i = 0; i < count; i++) collapse_body(c, queue[cur][i]);
        count = counters[2];
        counters[2] = 0;
        cur ^= 1;
    }
    wide.resize(counters[0]);
    return counters[1] == nt;
}

extern "C" {

const char* hostsim_last_error() { return g_err.c_str(); }

// The two-level layout (LmcuBuildOptions::two_level), as lmcu_accel_build lays it out: [top-level tree | one tree per mesh in use]
static int build_two_level(SimScene* s, int max_leaf) {
    const std::vector<PrimDev>& prims = s->ps.prims;
    const uint32_t np = (uint32_t)prims.size(), nt = s->ps.nt;
    // top level: world-space box per primitive
    std::vector<float> ilo((size_t)np * 3, 3e38f), ihi((size_t)np * 3, -3e38f);
    for (uint32_t g = 0; g < nt; g++) {
        float lo[3], hi[3];
        const uint32_t pi = world_bounds_body(g, prims.data(), (int)np, s->mesh_p.data(), lo, hi);
        for (int a = 0; a < 3; a++) { ilo[pi * 3 + a] = fminf(ilo[pi * 3 + a], lo[a]); ihi[pi * 3 + a] = fmaxf(ihi[pi * 3 + a], hi[a]); }
    }
    std::vector<double> icentre((size_t)np * 3);
    for (size_t i = 0; i < (size_t)np * 3; i++) icentre[i] = ((double)ilo[i] + (double)ihi[i]) * 0.5;
    for (int a = 0; a < 3; a++) { s->bounds[a] = 3e38f; s->bounds[3 + a] = -3e38f; }
    for (uint32_t i = 0; i < np; i++) for (int a = 0; a < 3; a++) {
        s->bounds[a] = fminf(s->bounds[a], ilo[i * 3 + a]); s->bounds[3 + a] = fmaxf(s->bounds[3 + a], ihi[i * 3 + a]);
    }
    std::vector<WideNode> top;
    std::vector<uint32_t> inst_order;
    if (!build_tree_host(np, ilo, ihi, icentre, max_leaf, top, inst_order)) return 3;
    s->nodes = top;
    // bottom levels
    const size_t nm = s->ps.face_offset.size() - 1;
    std::vector<uint32_t> root(nm, 0xffffffffu);
    for (uint32_t i = 0; i < np; i++) {
        const int m = prims[i].mesh;
        if (root[m] != 0xffffffffu) continue;
        const uint32_t nf = prims[i].num_faces;
        const double* p = s->mesh_p.data() + (size_t)s->ps.face_offset[m] * 9;
        std::vector<float> lo((size_t)nf * 3), hi((size_t)nf * 3);
        std::vector<double> centre((size_t)nf * 3);
        for (uint32_t k = 0; k < nf; k++) obj_bounds_body(k, p, lo.data(), hi.data(), centre.data());
        std::vector<WideNode> wide;
        std::vector<uint32_t> order;
        if (!build_tree_host(nf, lo, hi, centre, max_leaf, wide, order)) return 3;
        const uint32_t node_off = (uint32_t)s->nodes.size(), item_off = (uint32_t)s->tobj.size();
        for (WideNode& w : wide) relocate_body(w, node_off, item_off);
        s->nodes.insert(s->nodes.end(), wide.begin(), wide.end());
        s->tobj.resize(item_off + nf);
        for (uint32_t k = 0; k < nf; k++) emit_obj_body(k, order.data(), p, m, s->tobj.data() + item_off);
        root[m] = node_off;
    }
    s->inst.resize(np);
    for (uint32_t k = 0; k < np; k++) {
        const uint32_t pi = inst_order[k];
        if (!make_inst(prims[pi], pi, root[prims[pi].mesh], s->inst[k])) return 4;
    }
    s->two_level = true;
    return 0;
}

int hostsim_scene_create2(const LmcuSceneDesc* d, int max_leaf, int two_level, void** out, long long* num_nodes) {
    auto* s = new SimScene;
    const int rc = prepare_scene(d, s->ps, g_err);
    if (rc) { delete s; return rc; }
    const uint32_t nt = s->ps.nt;
    const size_t total_faces = s->ps.face_offset.back();
    s->mesh_p.assign(total_faces * 9, 0.0);
    s->mesh_n.assign(total_faces * 9, 0.0);
    for (int i = 0; i < d->num_meshes; i++) {
        const size_t cnt = (size_t)d->meshes[i].num_faces * 9;
        memcpy(s->mesh_p.data() + (size_t)s->ps.face_offset[i] * 9, d->meshes[i].p, cnt * sizeof(double));
        if (d->meshes[i].n) memcpy(s->mesh_n.data() + (size_t)s->ps.face_offset[i] * 9, d->meshes[i].n, cnt * sizeof(double));
    }
    if (want_two_level(two_level, s->ps.prims, s->ps.face_offset, nt)) {
        const int rc2 = build_two_level(s, max_leaf);
        if (rc2) { g_err = "two-level build failed"; delete s; return rc2; }
        if (num_nodes) *num_nodes = (long long)s->nodes.size();
        *out = s;
        return 0;
    }
    // k_flatten
    std::vector<double> wp((size_t)nt * 9), centre((size_t)nt * 3);
    std::vector<float> tri_lo((size_t)nt * 3), tri_hi((size_t)nt * 3);
    bool exact = true;
    float slo[3] = { 3e38f, 3e38f, 3e38f }, shi[3] = { -3e38f, -3e38f, -3e38f };
    for (uint32_t g = 0; g < nt; g++) {
        exact = flatten_body(g, s->ps.prims.data(), (int)s->ps.prims.size(), s->mesh_p.data(), wp.data(), tri_lo.data(),
                             tri_hi.data(), centre.data()) && exact;
        for (int a = 0; a < 3; a++) { slo[a] = fminf(slo[a], tri_lo[(size_t)g * 3 + a]); shi[a] = fmaxf(shi[a], tri_hi[(size_t)g * 3 + a]); }
    }
    for (int a = 0; a < 3; a++) { s->bounds[a] = slo[a]; s->bounds[3 + a] = shi[a]; }
    std::vector<WideNode> wide;
    std::vector<uint32_t> tri_order;
    if (!build_tree_host(nt, tri_lo, tri_hi, centre, max_leaf, wide, tri_order)) { g_err = "collapse lost triangles"; delete s; return 3; }
    s->nodes = wide;
    s->fp32 = exact;
    if (exact) s->t32.resize(nt); else s->t64.resize(nt);
    s->shade.resize(nt);
    if (!s->ps.mesh_t.empty()) s->uv.resize(nt);
    for (uint32_t k = 0; k < nt; k++) {
        emit_body(k, tri_order.data(), s->ps.prims.data(), (int)s->ps.prims.size(), wp.data(), s->mesh_p.data(), s->mesh_n.data(),
                  exact ? s->t32.data() : nullptr, exact ? nullptr : s->t64.data(), s->shade.data(),
                  s->uv.empty() ? nullptr : s->ps.mesh_t.data(), s->uv.empty() ? nullptr : s->uv.data());
    }
    if (num_nodes) *num_nodes = (long long)s->nodes.size();
    *out = s;
    return 0;
}

int hostsim_scene_create(const LmcuSceneDesc* d, int max_leaf, void** out, long long* num_nodes) {
    return hostsim_scene_create2(d, max_leaf, -1, out, num_nodes);
}

int hostsim_two_level(void* p) { return static_cast<SimScene*>(p)->two_level ? 1 : 0; }

void hostsim_scene_destroy(void* p) { delete static_cast<SimScene*>(p); }

int hostsim_fp32(void* p) { return static_cast<SimScene*>(p)->fp32 ? 1 : 0; }

int hostsim_intersect(void* p, const LmcuRay* rays, long long n, LmcuHit* hits, double* npr, double* tpr) {
    const SimScene& s = *static_cast<SimScene*>(p);
    if (s.two_level) return intersect_t<TriObj>(s, rays, n, hits, npr, tpr);
    return s.fp32 ? intersect_t<TriF32>(s, rays, n, hits, npr, tpr) : intersect_t<TriF64>(s, rays, n, hits, npr, tpr);
}

int hostsim_raycast(void* p, const LmcuRenderParams* rp_, float* film) {
    const SimScene& s = *static_cast<SimScene*>(p);
    const DevScene d = s.dev();
    RaycastParams rp{};
    rp.width = rp_->width; rp.height = rp_->height;
    rp.use_constant_color = rp_->use_constant_color; rp.visualize_normal = rp_->visualize_normal; rp.has_color = rp_->has_color;
    for (int k = 0; k < 3; k++) { rp.bg[k] = (float)rp_->bg_color[k]; rp.color[k] = (float)rp_->color[k]; }
    uint2 stack[kStackSize];
    for (int y = 0; y < rp.height; y++) {
        for (int x = 0; x < rp.width; x++) {
            if (s.two_level) raycast_pixel<TriObj>(d, rp, x, y, film, stack);
            else if (s.fp32) raycast_pixel<TriF32>(d, rp, x, y, film, stack); else raycast_pixel<TriF64>(d, rp, x, y, film, stack);
        }
    }
    return 0;
}

int hostsim_ao(void* p, const LmcuRenderParams* rp, float* film) {
    const SimScene& s = *static_cast<SimScene*>(p);
    const DevScene d = s.dev();
    uint2 stack[kStackSize];
    uint32_t rays = 0;
    memset(film, 0, sizeof(float) * 3 * rp->width * rp->height);
    for (int y = 0; y < rp->height; y++) {
        for (int x = 0; x < rp->width; x++) {
            if (s.two_level) ao_pixel<TriObj>(d, rp->width, rp->height, x, y, rp->spp, rp->sample_offset, rp->seed, rp->ao_distance, film, stack, &rays);
            else if (s.fp32) ao_pixel<TriF32>(d, rp->width, rp->height, x, y, rp->spp, rp->sample_offset, rp->seed, rp->ao_distance, film, stack, &rays);
            else ao_pixel<TriF64>(d, rp->width, rp->height, x, y, rp->spp, rp->sample_offset, rp->seed, rp->ao_distance, film, stack, &rays);
        }
    }
    return 0;
}

// One path at a time through the wavefront stage bodies (generate -> extend -> shade -> shadow).
int hostsim_pt(void* p, const LmcuRenderParams* rp, float* film, long long* rays_out) {
    const SimScene& s = *static_cast<SimScene*>(p);
    const DevScene d = s.dev();
    PtParams pp{};
    pp.width = rp->width; pp.height = rp->height; pp.max_verts = rp->max_verts; pp.sampling_mode = rp->sampling_mode;
    pp.seed = rp->seed; pp.tiled = (rp->width % 8 == 0 && rp->height % 4 == 0) ? 1u : 0u;
    pp.sample_offset = (uint64_t)rp->sample_offset;
    const bool image = rp->primary_ray_sampling_mode == LMCU_PRIMARY_IMAGE;
    pp.image_mode = image ? 1u : 0u;
    pp.total_paths = image ? (uint64_t)rp->num_samples : (uint64_t)rp->width * rp->height * (uint64_t)rp->spp;
    pp.capacity = 1;
    pp.chunk = 1;
    if (!image) { uint32_t c = 16; while (c > 1u && (rp->spp % c) != 0) c >>= 1; pp.chunk = c; }
    const size_t npix = (size_t)rp->width * rp->height;
    std::vector<double> acc(npix * 3, 0.0);
    float4 ro, rd, tp, aux; uint2 info;
    PathQueue q; q.ray_o = &ro; q.ray_d = &rd; q.thr_pix = &tp; q.aux = &aux; q.info = &info;
    uint2 stack[kStackSize];
    long long rays = 0;
    for (uint64_t g = 0; g < pp.total_paths; g++) {
        const GenOut go = pt_generate_body(d, pp, g, q, 0);
        if (go.shadow) {        // image mode: NEE connection of the camera vertex
            HitD hs;
            if (s.two_level) pt_trace<true, false, TriObj>(d, go.sh_o, go.sh_d, hs, stack, nullptr); else if (s.fp32) pt_trace<true, false, TriF32>(d, go.sh_o, go.sh_d, hs, stack, nullptr); else pt_trace<true, false, TriF64>(d, go.sh_o, go.sh_d, hs, stack, nullptr);
            rays++;
            const uint32_t px = f2u(go.sh_c.w);
            if (hs.tri < 0) { acc[px * 3 + 0] += go.sh_c.x; acc[px * 3 + 1] += go.sh_c.y; acc[px * 3 + 2] += go.sh_c.z; }
        }
        while (true) {
            HitD h;
            if (s.two_level) pt_trace<false, false, TriObj>(d, ro, rd, h, stack, nullptr); else if (s.fp32) pt_trace<false, false, TriF32>(d, ro, rd, h, stack, nullptr); else pt_trace<false, false, TriF64>(d, ro, rd, h, stack, nullptr);
            rays++;
            const float4 hit = make_float4((float)h.t, (float)h.u, (float)h.v, u2f((uint32_t)h.tri));
            const ShadeOut o = pt_shade_body<true>(d, pp, ro, rd, tp, aux, info, hit);
            const uint32_t pixel = f2u(tp.w);
            acc[pixel * 3 + 0] += o.emit.x; acc[pixel * 3 + 1] += o.emit.y; acc[pixel * 3 + 2] += o.emit.z;
            if (o.shadow) {
                HitD hs;
                if (s.two_level) pt_trace<true, false, TriObj>(d, o.sh_o, o.sh_d, hs, stack, nullptr); else if (s.fp32) pt_trace<true, false, TriF32>(d, o.sh_o, o.sh_d, hs, stack, nullptr); else pt_trace<true, false, TriF64>(d, o.sh_o, o.sh_d, hs, stack, nullptr);
                rays++;
                if (hs.tri < 0) { acc[pixel * 3 + 0] += o.sh_c.x; acc[pixel * 3 + 1] += o.sh_c.y; acc[pixel * 3 + 2] += o.sh_c.z; }
            }
            if (!o.alive) break;
            ro = o.o; rd = o.d; tp = o.thr_pix; aux = o.aux; info = o.info;
        }
    }
    const double sc = !rp->rescale ? 1.0 : image ? (double)npix / (double)rp->num_samples : 1.0 / (double)rp->spp;
    for (size_t i = 0; i < npix * 3; i++) film[i] = (float)(acc[i] * sc);
    if (rays_out) *rays_out = rays;
    return 0;
}


// Light tracer, one light path at a time through lt_generate_body / lt_shade_body (renderer_lt.cpp).
int hostsim_lt(void* p, const LmcuRenderParams* rp, float* film, long long* rays_out) {
    const SimScene& s = *static_cast<SimScene*>(p);
    DevScene d = s.dev();
    {
        float r2 = 0.f;
        for (int a = 0; a < 3; a++) {
            d.bound_center[a] = 0.5f * (s.bounds[a] + s.bounds[3 + a]);
            const float e = s.bounds[3 + a] - d.bound_center[a];
            r2 += e * e;
        }
        d.bound_radius = std::sqrt(r2) * 1.01f;
    }
    PtParams pp{};
    pp.width = rp->width; pp.height = rp->height; pp.max_verts = rp->max_verts; pp.seed = rp->seed;
    pp.sample_offset = (uint64_t)rp->sample_offset;
    pp.image_mode = 1u; pp.light_tracing = 1u;
    pp.total_paths = (uint64_t)rp->num_samples;
    const size_t npix = (size_t)rp->width * rp->height;
    std::vector<double> acc(npix * 3, 0.0);
    float4 ro, rd, tp, aux; uint2 info;
    PathQueue q; q.ray_o = &ro; q.ray_d = &rd; q.thr_pix = &tp; q.aux = &aux; q.info = &info;
    uint2 stack[kStackSize];
    long long rays = 0;
    for (uint64_t g = 0; g < pp.total_paths; g++) {
        lt_generate_body(d, pp, g, q, 0);
        while (true) {
            HitD h;
            if (s.two_level) pt_trace<false, false, TriObj>(d, ro, rd, h, stack, nullptr); else if (s.fp32) pt_trace<false, false, TriF32>(d, ro, rd, h, stack, nullptr); else pt_trace<false, false, TriF64>(d, ro, rd, h, stack, nullptr);
            rays++;
            const float4 hit = make_float4((float)h.t, (float)h.u, (float)h.v, u2f((uint32_t)h.tri));
            const ShadeOut o = lt_shade_body<true>(d, pp, ro, rd, tp, info, hit);
            if (o.shadow) {
                HitD hs;
                if (s.two_level) pt_trace<true, false, TriObj>(d, o.sh_o, o.sh_d, hs, stack, nullptr); else if (s.fp32) pt_trace<true, false, TriF32>(d, o.sh_o, o.sh_d, hs, stack, nullptr); else pt_trace<true, false, TriF64>(d, o.sh_o, o.sh_d, hs, stack, nullptr);
                rays++;
                const uint32_t px = f2u(o.sh_c.w);
                if (hs.tri < 0) { acc[px * 3 + 0] += o.sh_c.x; acc[px * 3 + 1] += o.sh_c.y; acc[px * 3 + 2] += o.sh_c.z; }
            }
            if (!o.alive) break;
            ro = o.o; rd = o.d; tp = o.thr_pix; aux = o.aux; info = o.info;
        }
    }
    const double sc = rp->rescale ? (double)npix / (double)rp->num_samples : 1.0;
    for (size_t i = 0; i < npix * 3; i++) film[i] = (float)(acc[i] * sc);
    if (rays_out) *rays_out = rays;
    return 0;
}

// Bidirectional path tracer, one sample at a time through bdpt_sample (renderer_bdpt.cpp:304-407).
int hostsim_bdpt(void* p, const LmcuRenderParams* rp, float* film, long long* rays_out) {
    const SimScene& s = *static_cast<SimScene*>(p);
    DevScene d = s.dev();
    {
        float r2 = 0.f;
        for (int a = 0; a < 3; a++) {
            d.bound_center[a] = 0.5f * (s.bounds[a] + s.bounds[3 + a]);
            const float e = s.bounds[3 + a] - d.bound_center[a];
            r2 += e * e;
        }
        d.bound_radius = std::sqrt(r2) * 1.01f;
    }
    PtParams pp{};
    pp.width = rp->width; pp.height = rp->height; pp.max_verts = rp->max_verts; pp.seed = rp->seed;
    pp.sample_offset = (uint64_t)rp->sample_offset;
    pp.image_mode = 1u;
    pp.bdpt_mode = (uint32_t)rp->bdpt_strategies;
    pp.total_paths = (uint64_t)rp->num_samples;
    if (rp->max_verts > kBdptMaxVerts) { g_err = "max_verts too large"; return 1; }
    const int min_verts = rp->min_verts > 0 ? rp->min_verts : 2;
    const size_t npix = (size_t)rp->width * rp->height;
    std::vector<double> acc(npix * 3, 0.0);
    auto splat = [&](uint32_t px, F3 C) { acc[px * 3 + 0] += C.x; acc[px * 3 + 1] += C.y; acc[px * 3 + 2] += C.z; };
    uint2 stack[kStackSize];
    BdptCounters bc;
    bc.rays = 0; bc.overflow = 0;
    long long total_rays = 0;
    for (uint64_t g = 0; g < pp.total_paths; g++) {
        bc.rays = 0;
        if (s.two_level) bdpt_sample<TriObj, true>(d, pp, min_verts, g, stack, &bc, splat);
        else if (s.fp32) bdpt_sample<TriF32, true>(d, pp, min_verts, g, stack, &bc, splat); else bdpt_sample<TriF64, true>(d, pp, min_verts, g, stack, &bc, splat);
        total_rays += bc.rays;
    }
    if (bc.overflow) { g_err = "bvh traversal stack overflow"; return 1; }
    const double sc = rp->rescale ? (double)npix / (double)rp->num_samples : 1.0;
    for (size_t i = 0; i < npix * 3; i++) film[i] = (float)(acc[i] * sc);
    if (rays_out) *rays_out = total_rays;
    return 0;
}

}
