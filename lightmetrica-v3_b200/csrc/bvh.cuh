// bvh.cuh — compressed 8-wide BVH: node/triangle layout in HBM and the closest-hit / any-hit
// traversal used by every kernel (accel batches, raycast, ao, pt extend + shadow).
//
// Replaces /root/reference/src/accel/accel_sahbvh.cpp:58-75 (Tri::intersect), :269-299
// (Accel_SAHBVH::intersect) and include/lm/math.h:143-158 (Bound::isect_range).
//
// Layout (all 16-byte aligned, fetched as uint4 / float4 / double2 vectors):
//   WideNode   80 B   8 children, AABBs quantised to 8 bits on a per-node power-of-two grid,
//                     children stored in octant order so slot^octant gives front-to-back order
//                     (compressed wide BVH, Ylitie, Karras & Laine, HPG 2017)
//   TriF32     48 B   world-space p1,p2,p3 as float when every coordinate is float-representable
//   TriF64     80 B   world-space p1,p2,p3 as double otherwise (e.g. rotated instances)
//
// Exactness contract: leaf tests run the reference's Möller–Trumbore in FP64
// with the reference's operation order and no FMA contraction, so (t,u,v) of a tested triangle are
// bit-identical to accel::sahbvh's.  Node tests run in FP32 on conservatively padded slabs, so they
// can only visit MORE boxes than an exact test would — never skip a triangle the FP64 test accepts.
#pragma once
#include "common.cuh"

namespace lmcu {

struct alignas(16) WideNode {
    float px, py, pz;            // origin of the quantisation grid (<= node min)
    uint8_t ex, ey, ez;          // biased float exponents: grid step = 2^(e-127)
    uint8_t imask;               // bit i set <=> slot i holds an internal node
    uint32_t child_base;         // index of the first internal child (children are contiguous)
    uint32_t tri_base;           // index of the first triangle referenced by leaf slots
    uint8_t meta[8];             // internal: 0b001'11sss (sss=slot); leaf: unary count<<5 | first tri (<24); 0 = empty
    uint8_t qlox[8], qloy[8], qloz[8];
    uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

struct alignas(16) TriF32 { float4 a, b, c; static constexpr bool kTwoLevel = false; };   // a=(p1,prim) b=(p2,face) c=(p3,0); prim/face are int bit patterns
struct alignas(16) TriF64 { double p[9]; int32_t prim, face; static constexpr bool kTwoLevel = false; };
static_assert(sizeof(TriF32) == 48 && sizeof(TriF64) == 80, "triangle records");

// Two-level layout (instanced scenes; the analogue of plugin/accel_embree/accel_embree_instanced.cpp:96-165): ONE tree per
// mesh over its OBJECT-space triangles (TriObj, bottom level) and one tree over the instances (InstDev, top level), all in one
// node array — the top level's root is node 0, its leaf slots index InstDev records instead of triangles.  Node tests of a
// bottom level run on the ray taken to object space (t is the same parameter in both spaces: the direction is transformed,
// not normalised).  The LEAF test stays in world space: the triangle's vertices are transformed on the fly with the operations
// the flattened build applies once (transform_point: glm's mat4 * vec4 order), so the Möller–Trumbore operands — and
// (t, u, v) — are bit-identical to the flattened tree's and to accel::sahbvh's.
struct alignas(16) TriObj { double p[9]; int32_t mesh, face; static constexpr bool kTwoLevel = true; };
struct alignas(16) InstDev {
    double M[12];                // object -> world, columns 0..3 of the instance's global transform (rows 0..2)
    double Minv[12];             // world -> object
    uint32_t root;               // root node of the mesh's tree
    uint32_t tri_offset;         // global id of the instance's face 0 (PrimDev::tri_offset): a hit reports tri_offset + face
    uint32_t prim, pad_;
};
static_assert(sizeof(TriObj) == 80 && sizeof(InstDev) == 208, "two-level records");

struct RayD {                 // FP64 ray as the reference sees it
    D3 o, d;
    double tmin, tmax;
};

struct HitD {                 // closest-hit state; tri < 0 = no hit yet
    double t, u, v;
    int32_t tri;
};

// -- integer helpers ----------------------------------------------------------------------------
LMCU_HD int bfind(uint32_t v) {      // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz(v);
#else
    return 31 - __builtin_clz(v);
#endif
}
LMCU_HD int popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
LMCU_HD uint32_t sign_extend_s8x4(uint32_t v) {   // each byte: 0x80 set -> 0xff else 0x00
    return ((v >> 7) & 0x01010101u) * 0xffu;
}

// -- Möller–Trumbore, FP64, reference operation order (accel_sahbvh.cpp:58-75) -------------------
LMCU_HD bool tri_intersect(const RayD& r, D3 p1, D3 p2, D3 p3, double tl, double th,
                           double& t_out, double& u_out, double& v_out) {
    const D3 e1 = dsub3(p2, p1);          // Tri ctor, accel_sahbvh.cpp:43-44
    const D3 e2 = dsub3(p3, p1);
    const D3 p = dcross(r.d, e2);
    const D3 tv = dsub3(r.o, p1);
    const D3 q = dcross(tv, e1);
    const double d = ddot(e1, p);
    const double ad = fabs(d);
    const double u = mul_sign_of(ddot(tv, p), d);
    const double v = mul_sign_of(ddot(r.d, q), d);
    if (ad < 1e-8 || u < 0.0 || v < 0.0 || addrn(u, v) > ad) {
        return false;
    }
    const double t = divrn(ddot(e2, q), d);
    if (t < tl || th < t) {
        return false;
    }
    t_out = t;
    u_out = divrn(u, ad);
    v_out = divrn(v, ad);
    return true;
}

LMCU_HD void load_tri(const TriF32* tris, uint32_t i, D3& p1, D3& p2, D3& p3) {
    const float4 a = tris[i].a, b = tris[i].b, c = tris[i].c;
    p1 = d3(a.x, a.y, a.z); p2 = d3(b.x, b.y, b.z); p3 = d3(c.x, c.y, c.z);
}
LMCU_HD void load_tri(const TriF64* tris, uint32_t i, D3& p1, D3& p2, D3& p3) {
    const double2* q = reinterpret_cast<const double2*>(tris + i);
    const double2 a = q[0], b = q[1], c = q[2], d = q[3];
    const double e = tris[i].p[8];
    p1 = d3(a.x, a.y, b.x); p2 = d3(b.y, c.x, c.y); p3 = d3(d.x, d.y, e);
}
LMCU_HD void load_tri(const TriObj* tris, uint32_t i, D3& p1, D3& p2, D3& p3) {
    load_tri(reinterpret_cast<const TriF64*>(tris), i, p1, p2, p3);
}

// glm mat4 * vec4(p,1) on the 3x4 part: (m0*x + m1*y) + (m2*z + m3*1)   (glm/detail/type_mat4x4.inl:526-537; build.cuh
// transform_point is the same on a 16-entry matrix)
LMCU_HD D3 transform_point34(const double* M, D3 p) {
    D3 r;
    r.x = addrn(addrn(mulrn(M[0], p.x), mulrn(M[3], p.y)), addrn(mulrn(M[6], p.z), M[9]));
    r.y = addrn(addrn(mulrn(M[1], p.x), mulrn(M[4], p.y)), addrn(mulrn(M[7], p.z), M[10]));
    r.z = addrn(addrn(mulrn(M[2], p.x), mulrn(M[5], p.y)), addrn(mulrn(M[8], p.z), M[11]));
    return r;
}

// -- traversal ------------------------------------------------------------------------------------
constexpr int kStackSize = 48;   // node-group entries; depth of the wide tree bounds it

struct TraverseCounters { uint32_t nodes, tris; };

// Conservative FP32 view of an FP64 ray.
struct RayF {
    float ox, oy, oz;
    float idx, idy, idz;         // 1/d, |d| clamped away from 0
    float aox, aoy, aoz;         // |o|
    uint32_t octinv4;            // (7 - octant) replicated in 4 bytes
};

// FP32 view of a ray whose components are given as floats (the pt queues hold float rays: no FP64 round trip)
LMCU_HD RayF make_rayf_f(float ox, float oy, float oz, float dx, float dy, float dz) {
    RayF f;
    f.ox = ox; f.oy = oy; f.oz = oz;
    const float tiny = 1e-20f;
    if (fabsf(dx) < tiny) dx = copysignf(tiny, dx);
    if (fabsf(dy) < tiny) dy = copysignf(tiny, dy);
    if (fabsf(dz) < tiny) dz = copysignf(tiny, dz);
    f.idx = 1.0f / dx; f.idy = 1.0f / dy; f.idz = 1.0f / dz;
    f.aox = fabsf(f.ox); f.aoy = fabsf(f.oy); f.aoz = fabsf(f.oz);
    const uint32_t oct = (dx < 0.f ? 4u : 0u) | (dy < 0.f ? 2u : 0u) | (dz < 0.f ? 1u : 0u);
    f.octinv4 = (7u - oct) * 0x01010101u;
    return f;
}

LMCU_HD RayF make_rayf(const RayD& r) {
    return make_rayf_f((float)r.o.x, (float)r.o.y, (float)r.o.z, (float)r.d.x, (float)r.d.y, (float)r.d.z);
}

// Byte j of `w` as the float 32768 + byte, built with ONE byte permute and no integer->float
// conversion: the bytes land in mantissa bits 8..15 of 0x47000000 (= 2^15, ulp 2^-8).  I2F.U8 runs on
// the XU pipe at 16 lanes/clk/SM on sm_100a — the first ncu capture of this kernel showed the XU pipe
// 60 % busy with the 48 conversions per node while the FMA pipe idled at 13 %
// (profiles/r01_ncu_summary_v2.md) — PRMT runs on the ALU pipe at 64 lanes/clk/SM.
#if defined(__CUDACC__)
// Lives in the constant bank so that ptxas keeps the SELECTOR as PRMT's immediate and reads this operand
// from c[][] (with a literal here it spends one extra IMAD.MOV per permute to materialise the selector).
__constant__ uint32_t kBiasBits = 0x47000000u;
#endif
template <int J>
LMCU_HD float byte_as_biased_float(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, kBiasBits, 0x7404u | ((uint32_t)J << 4)));
#else
    return u2f(0x47000000u | (((w >> (8 * J)) & 0xffu) << 8));
#endif
}
constexpr float kByteBias = 32768.0f;

// Intersects the 8 quantised child boxes of one node; returns the CWBVH hit mask
// (bits 31..24: internal children in traversal priority, bits 23..0: triangles).
// (Measured and dropped: sm_100a's packed fma.rn.f32x2 — one FFMA2 forming the near and the far plane distance of an axis,
// 24 FFMA2 for the 48 FFMA of a node step, no extra moves in the SASS, 704 -> 680 instructions in k_pt_extend_q — is
// 0.5 % SLOWER on C2 (1909 against 1919 Mpaths/s, profiles/r02_ab_ffma2.log): an FFMA2 takes the issue port for two cycles,
// so the kernel's binding resource, issue slots, is not relieved.)
LMCU_HD uint32_t node_hitmask(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4,
                              const RayF& f, float tmin_dn, float tmax_up) {
    const float px = u2f(n0.x), py = u2f(n0.y), pz = u2f(n0.z);
    const uint32_t e = n0.w;
    const float sx = u2f((e & 0xffu) << 23), sy = u2f(((e >> 8) & 0xffu) << 23), sz = u2f(((e >> 16) & 0xffu) << 23);
    const float ax = sx * f.idx, ay = sy * f.idy, az = sz * f.idz;
    const float rx = px - f.ox, ry = py - f.oy, rz = pz - f.oz;
    const float bx = rx * f.idx, by = ry * f.idy, bz = rz * f.idz;
    // Slab planes are evaluated as t = (32768 + q) * a + (b - 32768 a) with q the quantised coordinate.
    // Padding that covers (a) rounding o,d from FP64 to FP32, (b) the FP32 rounding of the slab
    // arithmetic including cancellation in (p - o): 8 ulp32 of the largest magnitude involved, where
    // the grid term counts the 255-step extent plus the 32768-step bias whose rounding enters b' once
    // (2^-24 * 32768 |a| <= 8448 * 2^-21 |a| with room to spare).
    const float kPad = 4.76837158e-7f;   // 2^-21
    const float slx = (fabsf(rx) + 8448.0f * sx + f.aox) * kPad * fabsf(f.idx);
    const float sly = (fabsf(ry) + 8448.0f * sy + f.aoy) * kPad * fabsf(f.idy);
    const float slz = (fabsf(rz) + 8448.0f * sz + f.aoz) * kPad * fabsf(f.idz);
    const float cx = fmaf(-kByteBias, ax, bx), cy = fmaf(-kByteBias, ay, by), cz = fmaf(-kByteBias, az, bz);
    const float bnx = cx - slx, bfx = cx + slx;
    const float bny = cy - sly, bfy = cy + sly;
    const float bnz = cz - slz, bfz = cz + slz;
    const bool negx = f.idx < 0.f, negy = f.idy < 0.f, negz = f.idz < 0.f;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (f.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z, qlz = half ? n3.y : n3.x;
        const uint32_t qhx = half ? n3.w : n3.z, qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
        const uint32_t nx4 = negx ? qhx : qlx, fx4 = negx ? qlx : qhx;
        const uint32_t ny4 = negy ? qhy : qly, fy4 = negy ? qly : qhy;
        const uint32_t nz4 = negz ? qhz : qlz, fz4 = negz ? qlz : qhz;
#define LMCU_CHILD(J)                                                                                   \
        {                                                                                               \
            const float tnx = fmaf(byte_as_biased_float<J>(nx4), ax, bnx);                              \
            const float tny = fmaf(byte_as_biased_float<J>(ny4), ay, bny);                              \
            const float tnz = fmaf(byte_as_biased_float<J>(nz4), az, bnz);                              \
            const float tfx = fmaf(byte_as_biased_float<J>(fx4), ax, bfx);                              \
            const float tfy = fmaf(byte_as_biased_float<J>(fy4), ay, bfy);                              \
            const float tfz = fmaf(byte_as_biased_float<J>(fz4), az, bfz);                              \
            const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin_dn));                             \
            const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax_up));                             \
            if (cmin <= cmax) {                                                                         \
                hitmask |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu);     \
            }                                                                                           \
        }
        LMCU_CHILD(0) LMCU_CHILD(1) LMCU_CHILD(2) LMCU_CHILD(3)
#undef LMCU_CHILD
    }
    return hitmask;
}

// FP32 view of ray r in the object space of an instance: o' = Minv (o, 1), d' = Minv (d, 0) formed in FP64, then rounded
// (the rounding is what node_hitmask's padding covers; d' is NOT normalised, so t means the same in both spaces).
LMCU_HD RayF make_rayf_obj(const RayD& r, const double* Mi) {
    const double ox = Mi[0] * r.o.x + Mi[3] * r.o.y + Mi[6] * r.o.z + Mi[9];
    const double oy = Mi[1] * r.o.x + Mi[4] * r.o.y + Mi[7] * r.o.z + Mi[10];
    const double oz = Mi[2] * r.o.x + Mi[5] * r.o.y + Mi[8] * r.o.z + Mi[11];
    const double dx = Mi[0] * r.d.x + Mi[3] * r.d.y + Mi[6] * r.d.z;
    const double dy = Mi[1] * r.d.x + Mi[4] * r.d.y + Mi[7] * r.d.z;
    const double dz = Mi[2] * r.d.x + Mi[5] * r.d.y + Mi[8] * r.d.z;
    return make_rayf_f((float)ox, (float)oy, (float)oz, (float)dx, (float)dy, (float)dz);
}

constexpr uint32_t kNoInst = 0xffffffffu;

// Stack entries of the two-level traversal, told apart by .y: > 0x00ffffff a node group (as in the one-level tree),
// 1..0x00ffffff the instances of a top-level leaf still to be entered, 0 the marker below which the top level continues
// (popping it takes the ray back to world space).

// Closest hit (ANY=false) or any hit (ANY=true) of ray r in [r.tmin, r.tmax].
// hit.t must be initialised to r.tmax and hit.tri to -1.  Returns false on stack overflow.
// TriObj (two-level): `inst` is the top level's leaf-ordered instance table; hit.tri = InstDev::tri_offset + face.
template <bool ANY, bool COUNT, typename TriT>
LMCU_HD bool traverse(const WideNode* __restrict__ nodes, const TriT* __restrict__ tris,
                      const RayD& r, HitD& hit, uint2* stack, TraverseCounters* ctr, const InstDev* __restrict__ inst = nullptr) {
    RayF f = make_rayf(r);
    const float tmin_dn = f32_down(r.tmin);
    float tmax_up = f32_up(hit.t);
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);   // the root occupies virtual slot 7 of group 0
    uint2 tgroup = make_uint2(0u, 0u);
    uint32_t cinst = kNoInst;
    const uint4* nodes4 = reinterpret_cast<const uint4*>(nodes);
    while (true) {
        if (ngroup.y > 0x00ffffffu) {
            const uint32_t hits = ngroup.y;
            const uint32_t imask = ngroup.y & 0xffu;
            const int child_bit = bfind(hits);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << child_bit);
            if (ngroup.y > 0x00ffffffu) {
                if (sp >= kStackSize) return false;
                stack[sp++] = ngroup;
            }
            const uint32_t slot = (uint32_t)(child_bit - 24) ^ (f.octinv4 & 0xffu);
            const uint32_t ni = base + popc(imask & ~(0xffffffffu << slot));
            const uint4 n0 = nodes4[ni * 5 + 0], n1 = nodes4[ni * 5 + 1], n2 = nodes4[ni * 5 + 2],
                        n3 = nodes4[ni * 5 + 3], n4 = nodes4[ni * 5 + 4];
            if (COUNT) ctr->nodes++;
            const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, f, tmin_dn, tmax_up);
            ngroup.x = n1.x;
            ngroup.y = (hm & 0xff000000u) | (n0.w >> 24);
            tgroup.x = n1.y;
            tgroup.y = hm & 0x00ffffffu;
        }
        while (tgroup.y != 0) {
            const int ti = bfind(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const uint32_t tri = tgroup.x + (uint32_t)ti;
            D3 p1, p2, p3;
            int32_t id = (int32_t)tri;
            if constexpr (TriT::kTwoLevel) {
                if (cinst == kNoInst) {
                    // a top-level leaf item: park what is left of this node, then walk the instance's tree in its object space
                    if (sp + 3 > kStackSize) return false;
                    if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
                    if (tgroup.y != 0u) stack[sp++] = tgroup;
                    stack[sp++] = make_uint2(0u, 0u);
                    cinst = tri;
                    f = make_rayf_obj(r, inst[cinst].Minv);
                    ngroup = make_uint2(inst[cinst].root, 0x80000000u);
                    tgroup = make_uint2(0u, 0u);
                    break;
                }
                load_tri(tris, tri, p1, p2, p3);
                const double* M = inst[cinst].M;
                p1 = transform_point34(M, p1); p2 = transform_point34(M, p2); p3 = transform_point34(M, p3);
                id = (int32_t)(inst[cinst].tri_offset + (uint32_t)tris[tri].face);
            } else {
                load_tri(tris, tri, p1, p2, p3);
            }
            if (COUNT) ctr->tris++;
            double t, u, v;
            if (tri_intersect(r, p1, p2, p3, r.tmin, hit.t, t, u, v)) {
                hit.t = t; hit.u = u; hit.v = v; hit.tri = id;
                if (ANY) return true;
                tmax_up = f32_up(t);
            }
        }
        if (ngroup.y <= 0x00ffffffu && tgroup.y == 0u) {
            if (sp == 0) break;
            const uint2 e = stack[--sp];
            if (TriT::kTwoLevel && e.y <= 0x00ffffffu) {
                if (e.y == 0u) { cinst = kNoInst; f = make_rayf(r); }      // back in the top level
                else tgroup = e;
            } else {
                ngroup = e;
            }
        }
    }
    return true;
}

}  // namespace lmcu
