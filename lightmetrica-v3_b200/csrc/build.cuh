// build.cuh — per-item bodies of the parallel BVH builder.
//
// Replaces /root/reference/src/accel/accel_sahbvh.cpp:127-266 (flatten + SAH build).  Pipeline
// (one kernel per stage, see lmcu.cu):
//   flatten     world-space FP64 vertices per triangle, glm operation order (accel_sahbvh.cpp:145-150)
//   morton      60-bit Morton code of the bound centre (the reference also splits on bound centres)
//   radix sort  (CUB, library call — not on the ray hot path)
//   hierarchy   binary radix tree over sorted codes (Karras, HPG 2012)
//   fit         bottom-up AABBs + subtree triangle counts (atomic arrival flags)
//   collapse    top-down, level-synchronous greedy collapse to 8-wide nodes by surface area, octant
//               slot assignment, 8-bit quantisation (Ylitie et al., HPG 2017)
//   emit        triangles + shading normals written in leaf order (TriF32 48 B / TriF64 80 B)
#pragma once
#include "bvh.cuh"

namespace lmcu {

struct PrimDev {
    double M[16];            // global transform, column-major
    double NM[9];            // normal matrix = mat3(transpose(inverse(M))), column-major (math.h:742-745)
    double J;                // |det(mat3(M))|
    int32_t node_index, mesh, material, light;
    uint32_t tri_offset;     // first global triangle id of this primitive
    uint32_t num_faces;
    uint32_t mesh_face_offset;   // first face of the mesh in the concatenated mesh arrays
    int32_t has_normals;
    float inv_area;          // area lights: 1 / sum of local triangle areas (light_area.cpp:68-73)
    float det_sign;          // sign of det(mat3(M)): keeps geometric normals consistent under mirroring
    int32_t pad_[2];
};

struct BinNode {             // binary radix-tree node (internal nodes [0,N-1), leaves [N-1, 2N-1))
    float lo[3], hi[3];
    uint32_t left, right;    // node ids (leaf j is id N-1+j)
    uint32_t first, count;   // range of sorted leaves covered by the subtree
};

LMCU_HD uint32_t find_prim(const PrimDev* prims, int num_prims, uint32_t g) {
    int lo = 0, hi = num_prims - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (prims[mid].tri_offset <= g) lo = mid; else hi = mid - 1;
    }
    return (uint32_t)lo;
}

// glm mat4 * vec4(p,1): (m0*x + m1*y) + (m2*z + m3*1)   (glm/detail/type_mat4x4.inl:526-537)
LMCU_HD D3 transform_point(const double* M, D3 p) {
    D3 r;
    r.x = addrn(addrn(mulrn(M[0], p.x), mulrn(M[4], p.y)), addrn(mulrn(M[8], p.z), M[12]));
    r.y = addrn(addrn(mulrn(M[1], p.x), mulrn(M[5], p.y)), addrn(mulrn(M[9], p.z), M[13]));
    r.z = addrn(addrn(mulrn(M[2], p.x), mulrn(M[6], p.y)), addrn(mulrn(M[10], p.z), M[14]));
    return r;
}

LMCU_HD bool f32_exact(double v) { return (double)(float)v == v; }

// One triangle: world vertices (FP64), float bounds rounded outward, exact-float flag.
// Returns false if any coordinate is not float-representable.
LMCU_HD bool flatten_body(uint32_t g, const PrimDev* prims, int num_prims, const double* mesh_p,
                          double* wp, float* tri_lo, float* tri_hi, double* centre) {
    const uint32_t pi = find_prim(prims, num_prims, g);
    const PrimDev& pr = prims[pi];
    const uint32_t face = g - pr.tri_offset;
    const double* src = mesh_p + (size_t)(pr.mesh_face_offset + face) * 9;
    double lo[3] = { kInf, kInf, kInf }, hi[3] = { -kInf, -kInf, -kInf };   // empty Bound, math.h:83-84
    bool exact = true;
    for (int k = 0; k < 3; k++) {
        const D3 w = transform_point(pr.M, d3(src[3 * k], src[3 * k + 1], src[3 * k + 2]));
        double* dst = wp + (size_t)g * 9 + 3 * k;
        dst[0] = w.x; dst[1] = w.y; dst[2] = w.z;
        exact = exact && f32_exact(w.x) && f32_exact(w.y) && f32_exact(w.z);
        lo[0] = fmin(lo[0], w.x); lo[1] = fmin(lo[1], w.y); lo[2] = fmin(lo[2], w.z);
        hi[0] = fmax(hi[0], w.x); hi[1] = fmax(hi[1], w.y); hi[2] = fmax(hi[2], w.z);
    }
    for (int a = 0; a < 3; a++) {
        tri_lo[(size_t)g * 3 + a] = f32_down(lo[a]);
        tri_hi[(size_t)g * 3 + a] = f32_up(hi[a]);
        centre[(size_t)g * 3 + a] = (lo[a] + hi[a]) * 0.5;      // Bound::center(), math.h:103
    }
    return exact;
}

LMCU_HD uint64_t expand_bits21(uint64_t v) {   // spread the low 20 bits: bit i -> bit 3i
    v &= 0xfffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// `big_frac` > 0: a triangle whose bound diagonal exceeds big_frac x the scene diagonal gets key bit 60 set, so that the
// radix tree keeps all such triangles in ONE subtree under the root.  Sorted by centre alone, a wall or ground polygon
// lands in a leaf next to millimetre-sized neighbours and inflates the bounds of every ancestor on its path: most rays
// then walk that whole path for nothing.  Measured (tools/big_probe.py): C2 4.89 -> 3.16 nodes per closest-hit ray and
// 4.35 -> 1.96 per shadow ray, 1331 -> 1599 Mpaths/s; C4 22.1 -> 20.2 nodes per ray, 369 -> 391 Mpaths/s; thresholds 1/4 .. 1/64
// give the same trees on both.
LMCU_HD uint64_t morton_body(const double* centre, uint32_t g, const double* slo, const double* shi,
                             const float* tri_lo = nullptr, const float* tri_hi = nullptr, double big_frac = 0.0) {
    uint64_t code = 0;
    if (big_frac > 0.0 && tri_lo) {
        double d2 = 0.0, s2 = 0.0;
        for (int a = 0; a < 3; a++) {
            const double e = (double)tri_hi[(size_t)g * 3 + a] - (double)tri_lo[(size_t)g * 3 + a], se = shi[a] - slo[a];
            d2 += e * e; s2 += se * se;
        }
        if (d2 > big_frac * big_frac * s2) code = 1ull << 60;
    }
    // one CUBIC grid over the scene (cell edge = longest extent / 2^20): on a per-axis grid a flat scene (a terrain) would be
    // split along its short axis as often as along the long ones
    const double ext = fmax(fmax(shi[0] - slo[0], shi[1] - slo[1]), shi[2] - slo[2]);
    for (int a = 0; a < 3; a++) {
        double u = ext > 0 ? (centre[(size_t)g * 3 + a] - slo[a]) / ext : 0.0;
        u = fmin(fmax(u, 0.0), 1.0);
        uint64_t q = (uint64_t)(u * 1048575.0);
        code |= expand_bits21(q) << (2 - a);
    }
    return code;
}

// -- Karras 2012 ------------------------------------------------------------------------------------
LMCU_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
LMCU_HD int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

LMCU_HD int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(a ^ b);
}

// Internal node i of the radix tree over n sorted keys.
LMCU_HD void hierarchy_body(int i, const uint64_t* keys, int n, BinNode* nodes, uint32_t* parent) {
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) {
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    }
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t left = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    const uint32_t right = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    nodes[i].left = left;
    nodes[i].right = right;
    nodes[i].first = (uint32_t)lo;
    nodes[i].count = (uint32_t)(hi - lo + 1);
    parent[left] = (uint32_t)i;
    parent[right] = (uint32_t)i;
}

LMCU_HD float box_area(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

// -- collapse one wide node ---------------------------------------------------------------------------
struct CollapseItem { uint32_t bin_node, wide_index; };

struct CollapseCtx {
    const BinNode* bin;          // 2N-1 nodes
    const uint32_t* sorted_tri;  // sorted leaf position -> global triangle id
    int n;                       // number of triangles
    int max_leaf;                // triangles per leaf slot (1..3)
    int width;                   // children per wide node (2..8; 8 = the shipped layout; fewer leaves slots empty: probe only)
    WideNode* wide;
    uint32_t* tri_order;         // leaf-order position -> global triangle id
    uint32_t* node_counter;      // next free wide node
    uint32_t* tri_counter;       // next free leaf-order position
    CollapseItem* next_queue;
    uint32_t* next_count;
};

LMCU_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const uint32_t o = *p; *p += v; return o;
#endif
}

LMCU_HD void collapse_body(const CollapseCtx& c, CollapseItem item) {
    const BinNode* bin = c.bin;
    uint32_t ch[8];
    int nch;
    const BinNode& root = bin[item.bin_node];
    const bool root_is_leaf_cluster = item.bin_node >= (uint32_t)(c.n - 1) || (int)root.count <= c.max_leaf;
    if (root_is_leaf_cluster) {           // only for scenes with <= max_leaf triangles in total
        ch[0] = item.bin_node; nch = 1;
    } else {
        ch[0] = root.left; ch[1] = root.right; nch = 2;
        // (Opening by the largest area REDUCTION instead of the largest area: within 1 % on C2, C3, C4 — hostsim counters.)
        // Area-greedy opening leaves 3.7 of 8 slots occupied on C2, yet a variant that kept subtrees of <= 16..24
        // triangles whole as dense bottom nodes cut the node count by 22 % and RAISED nodes/ray by 10 % (hostsim
        // counters, round 1) — the large-area-first order is what keeps the visited set small.
        const int W = c.width >= 2 && c.width <= 8 ? c.width : 8;
        while (nch < W) {                 // open the child with the largest surface area
            int best = -1;
            float best_area = -1.f;
            for (int k = 0; k < nch; k++) {
                const BinNode& b = bin[ch[k]];
                if (ch[k] >= (uint32_t)(c.n - 1) || (int)b.count <= c.max_leaf) continue;
                const float a = box_area(b.lo, b.hi);
                if (a > best_area) { best_area = a; best = k; }
            }
            if (best < 0) break;
            const BinNode& b = bin[ch[best]];
            ch[best] = b.left;
            ch[nch++] = b.right;
        }
        // Bottom nodes: every child is a leaf cluster now and slots are left over.  Split the clusters (largest area
        // first) down to single triangles while slots remain: same node, tighter leaf boxes, fewer FP64 triangle tests.
        while (nch < W) {
            int best = -1;
            float best_area = -1.f;
            for (int k = 0; k < nch; k++) {
                if (ch[k] >= (uint32_t)(c.n - 1)) continue;                 // a single triangle
                const BinNode& b = bin[ch[k]];
                const float a = box_area(b.lo, b.hi);
                if (a > best_area) { best_area = a; best = k; }
            }
            if (best < 0) break;
            const BinNode& b = bin[ch[best]];
            ch[best] = b.left;
            ch[nch++] = b.right;
        }
    }
    // node bounds
    float lo[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, hi[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
    for (int k = 0; k < nch; k++) {
        const BinNode& b = bin[ch[k]];
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], b.lo[a]); hi[a] = fmaxf(hi[a], b.hi[a]); }
    }
    // octant slot assignment: slot s should hold the child that comes first for ray direction signs s
    int slot_of[8];
    int child_in_slot[8];
    for (int k = 0; k < 8; k++) { slot_of[k] = -1; child_in_slot[k] = -1; }
    const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
    for (int it = 0; it < nch; it++) {
        float best = 3.0e38f;
        int bc = -1, bs = -1;
        for (int k = 0; k < nch; k++) {
            if (slot_of[k] >= 0) continue;
            const BinNode& b = bin[ch[k]];
            const float ox = 0.5f * (b.lo[0] + b.hi[0]) - cx, oy = 0.5f * (b.lo[1] + b.hi[1]) - cy,
                        oz = 0.5f * (b.lo[2] + b.hi[2]) - cz;
            for (int s = 0; s < 8; s++) {
                if (child_in_slot[s] >= 0) continue;
                const float cost = ((s & 4) ? -ox : ox) + ((s & 2) ? -oy : oy) + ((s & 1) ? -oz : oz);
                if (cost < best) { best = cost; bc = k; bs = s; }
            }
        }
        slot_of[bc] = bs;
        child_in_slot[bs] = bc;
    }
    // count internal children / triangles, reserve contiguous ranges
    uint32_t n_internal = 0, n_tris = 0;
    for (int s = 0; s < 8; s++) {
        const int k = child_in_slot[s];
        if (k < 0) continue;
        const BinNode& b = bin[ch[k]];
        const bool leaf = ch[k] >= (uint32_t)(c.n - 1) || (int)b.count <= c.max_leaf;
        if (leaf) n_tris += (ch[k] >= (uint32_t)(c.n - 1)) ? 1u : b.count; else n_internal++;
    }
    const uint32_t child_base = n_internal ? atomic_add_u32(c.node_counter, n_internal) : 0u;
    const uint32_t tri_base = n_tris ? atomic_add_u32(c.tri_counter, n_tris) : 0u;
    const uint32_t q_base = n_internal ? atomic_add_u32(c.next_count, n_internal) : 0u;
    // quantisation grid
    WideNode w;
    memset(&w, 0, sizeof(w));
    w.px = lo[0]; w.py = lo[1]; w.pz = lo[2];
    uint8_t eb[3];
    double scale[3];
    for (int a = 0; a < 3; a++) {
        const double ext = (double)hi[a] - (double)lo[a];
        int e = 1;   // biased exponent of the grid step
        if (ext > 0) {
            int ex;
            frexp(ext / 255.0, &ex);        // ext/255 = m * 2^ex, m in [0.5,1)  ->  2^ex >= ext/255
            e = ex + 127;
            if (e < 1) e = 1;
            if (e > 254) e = 254;
        }
        eb[a] = (uint8_t)e;
        scale[a] = ldexp(1.0, e - 127);
    }
    w.ex = eb[0]; w.ey = eb[1]; w.ez = eb[2];
    w.child_base = child_base;
    w.tri_base = tri_base;
    uint32_t k_internal = 0, k_tri = 0;
    for (int s = 0; s < 8; s++) {
        const int k = child_in_slot[s];
        if (k < 0) continue;
        const BinNode& b = bin[ch[k]];
        const bool is_leaf_node = ch[k] >= (uint32_t)(c.n - 1);
        const bool leaf = is_leaf_node || (int)b.count <= c.max_leaf;
        uint8_t ql[3], qh[3];
        for (int a = 0; a < 3; a++) {
            const double l = floor(((double)b.lo[a] - (double)lo[a]) / scale[a]);
            const double h = ceil(((double)b.hi[a] - (double)lo[a]) / scale[a]);
            ql[a] = (uint8_t)fmin(fmax(l, 0.0), 255.0);
            qh[a] = (uint8_t)fmin(fmax(h, 0.0), 255.0);
        }
        w.qlox[s] = ql[0]; w.qloy[s] = ql[1]; w.qloz[s] = ql[2];
        w.qhix[s] = qh[0]; w.qhiy[s] = qh[1]; w.qhiz[s] = qh[2];
        if (leaf) {
            const uint32_t cnt = is_leaf_node ? 1u : b.count;
            const uint32_t first = is_leaf_node ? (ch[k] - (uint32_t)(c.n - 1)) : b.first;
            const uint32_t unary = cnt == 1 ? 1u : (cnt == 2 ? 3u : 7u);
            w.meta[s] = (uint8_t)((unary << 5) | k_tri);
            for (uint32_t t = 0; t < cnt; t++) {
                c.tri_order[tri_base + k_tri + t] = c.sorted_tri[first + t];
            }
            k_tri += cnt;
        } else {
            w.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            w.imask |= (uint8_t)(1u << s);
            CollapseItem ni;
            ni.bin_node = ch[k];
            ni.wide_index = child_base + k_internal;
            c.next_queue[q_base + k_internal] = ni;
            k_internal++;
        }
    }
    c.wide[item.wide_index] = w;
}

// -- emit ----------------------------------------------------------------------------------------------
// glm mat3 * vec3 = m0*x + m1*y + m2*z (left to right)
LMCU_HD F3 normal_transform(const double* NM, double x, double y, double z) {
    return f3((float)(NM[0] * x + NM[3] * y + NM[6] * z),
              (float)(NM[1] * x + NM[4] * y + NM[7] * z),
              (float)(NM[2] * x + NM[5] * y + NM[8] * z));
}

// world-space (unnormalised) vertex normals of one face; `mf` = offset of the face in the concatenated mesh arrays.  Meshes
// without normals shade with the geometric normal of the object-space triangle.
LMCU_HD void shade_normals(const PrimDev& pr, const double* mesh_p, const double* mesh_n, size_t mf, F3* n) {
    if (pr.has_normals) {
        for (int v = 0; v < 3; v++) {
            n[v] = normal_transform(pr.NM, mesh_n[mf + 3 * v], mesh_n[mf + 3 * v + 1], mesh_n[mf + 3 * v + 2]);
        }
    } else {
        const double* p = mesh_p + mf;
        const double e1[3] = { p[3] - p[0], p[4] - p[1], p[5] - p[2] };
        const double e2[3] = { p[6] - p[0], p[7] - p[1], p[8] - p[2] };
        const double gx = e1[1] * e2[2] - e1[2] * e2[1], gy = e1[2] * e2[0] - e1[0] * e2[2],
                     gz = e1[0] * e2[1] - e1[1] * e2[0];
        const double inv = 1.0 / sqrt(gx * gx + gy * gy + gz * gz);
        n[0] = n[1] = n[2] = normal_transform(pr.NM, gx * inv, gy * inv, gz * inv);
    }
}

// world-space (unnormalised) vertex normals; n1.w = sign of det(M) (geometric normal orientation),
// n2.w = material index, n3.w = light index or -1 (int bit patterns): shading needs no PrimDev lookup
struct alignas(16) TriShade { float4 n1, n2, n3; };

// per-vertex texture coordinates (Mesh::Point::t), leaf order; only built when a material reads a texture
struct alignas(8) TriUV { float2 t1, t2, t3; };

LMCU_HD void emit_body(uint32_t k, const uint32_t* tri_order, const PrimDev* prims, int num_prims,
                       const double* wp, const double* mesh_p, const double* mesh_n,
                       TriF32* t32, TriF64* t64, TriShade* shade, const float* mesh_t = nullptr, TriUV* uv = nullptr) {
    const uint32_t g = tri_order[k];
    const uint32_t pi = find_prim(prims, num_prims, g);
    const PrimDev& pr = prims[pi];
    const uint32_t face = g - pr.tri_offset;
    const double* w = wp + (size_t)g * 9;
    if (t32) {
        t32[k].a = make_float4((float)w[0], (float)w[1], (float)w[2], u2f(pi));
        t32[k].b = make_float4((float)w[3], (float)w[4], (float)w[5], u2f(face));
        t32[k].c = make_float4((float)w[6], (float)w[7], (float)w[8], 0.f);
    } else {
        for (int i = 0; i < 9; i++) t64[k].p[i] = w[i];
        t64[k].prim = (int32_t)pi;
        t64[k].face = (int32_t)face;
    }
    F3 n[3];
    shade_normals(pr, mesh_p, mesh_n, (size_t)(pr.mesh_face_offset + face) * 9, n);
    shade[k].n1 = make_float4(n[0].x, n[0].y, n[0].z, pr.det_sign);
    shade[k].n2 = make_float4(n[1].x, n[1].y, n[1].z, u2f((uint32_t)pr.material));
    shade[k].n3 = make_float4(n[2].x, n[2].y, n[2].z, u2f((uint32_t)pr.light));
    if (uv) {
        const float* t = mesh_t + (size_t)(pr.mesh_face_offset + face) * 6;
        TriUV o;
        o.t1.x = t[0]; o.t1.y = t[1]; o.t2.x = t[2]; o.t2.y = t[3]; o.t3.x = t[4]; o.t3.y = t[5];
        uv[k] = o;
    }
}

// -- two-level layout (bvh.cuh TriObj / InstDev) ---------------------------------------------------------------------------
// Bottom level: bounds of face k of ONE mesh in object space (`p` = the mesh's first face), float rounded outward.
LMCU_HD void obj_bounds_body(uint32_t k, const double* p, float* tri_lo, float* tri_hi, double* centre) {
    const double* src = p + (size_t)k * 9;
    double lo[3] = { kInf, kInf, kInf }, hi[3] = { -kInf, -kInf, -kInf };
    for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) { lo[a] = fmin(lo[a], src[3 * v + a]); hi[a] = fmax(hi[a], src[3 * v + a]); }
    for (int a = 0; a < 3; a++) {
        tri_lo[(size_t)k * 3 + a] = f32_down(lo[a]);
        tri_hi[(size_t)k * 3 + a] = f32_up(hi[a]);
        centre[(size_t)k * 3 + a] = (lo[a] + hi[a]) * 0.5;
    }
}

// Top level: world-space bounds of global triangle g — the vertices the leaf test will see (transform_point), so an
// instance's box (the union over its faces) encloses them exactly.
LMCU_HD uint32_t world_bounds_body(uint32_t g, const PrimDev* prims, int num_prims, const double* mesh_p, float* lo, float* hi) {
    const uint32_t pi = find_prim(prims, num_prims, g);
    const PrimDev& pr = prims[pi];
    const double* src = mesh_p + (size_t)(pr.mesh_face_offset + (g - pr.tri_offset)) * 9;
    double l[3] = { kInf, kInf, kInf }, h[3] = { -kInf, -kInf, -kInf };
    for (int k = 0; k < 3; k++) {
        const D3 w = transform_point(pr.M, d3(src[3 * k], src[3 * k + 1], src[3 * k + 2]));
        l[0] = fmin(l[0], w.x); l[1] = fmin(l[1], w.y); l[2] = fmin(l[2], w.z);
        h[0] = fmax(h[0], w.x); h[1] = fmax(h[1], w.y); h[2] = fmax(h[2], w.z);
    }
    for (int a = 0; a < 3; a++) { lo[a] = f32_down(l[a]); hi[a] = f32_up(h[a]); }
    return pi;
}

// Bottom-level triangle k (leaf order) of a mesh: object-space vertices + the mesh-local face
LMCU_HD void emit_obj_body(uint32_t k, const uint32_t* order, const double* p, int32_t mesh, TriObj* out) {
    const uint32_t face = order[k];
    for (int i = 0; i < 9; i++) out[k].p[i] = p[(size_t)face * 9 + i];
    out[k].mesh = mesh;
    out[k].face = (int32_t)face;
}

// A tree built on its own (nodes and leaf items counted from 0) moved to its place in the shared arrays
LMCU_HD void relocate_body(WideNode& w, uint32_t node_off, uint32_t item_off) {
    w.child_base += node_off;
    w.tri_base += item_off;
}

}  // namespace lmcu
