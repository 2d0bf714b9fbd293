// lmcu.cu — liblmcu.so: kernels (sm_100a) + host orchestration + the C ABI of include/lmcu.h.
//
// There is NO CPU path in this library: without a CUDA device every entry returns
// LMCU_ERR_NO_DEVICE.  Kernel launch geometry: persistent grids sized as a multiple of the SM
// count (148 on B200); traversal kernels fetch work one warp (32 rays) at a time from a global
// cursor; the path-tracer queues are compacted with warp ballots + one atomic per CTA.
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/lmcu.h"
#include "kernels.cuh"
#include "trace.cuh"
#include "trace2.cuh"
#include "bdpt.cuh"

#ifndef LMCU_SHADE_BLOCK
#define LMCU_SHADE_BLOCK 64    // threads per CTA of the shade kernel: small CTAs slip into the register space a single
                               // traversal CTA of the other wavefront frees (256 / 128 / 64 threads: 1810 / 1826 / 1833 Mpaths/s on C2)
#endif
#ifndef LMCU_SHADE_MINB
#define LMCU_SHADE_MINB 12 // resident CTAs per SM the shade kernel is compiled for (80 registers)
#endif
#ifndef LMCU_TRAV_BLOCK
#define LMCU_TRAV_BLOCK 128    // threads per CTA of the queue-traversal kernels (the loop is per warp; this only sets the granularity
                               // at which the block scheduler hands SM resources to the kernels of different wavefronts)
#endif
#ifndef LMCU_MINB_TWO
#define LMCU_MINB_TWO 5  // ... the two-level instantiations (TriObj): 96 registers, no spills (at 80 they spill 140 bytes)
#endif
#ifndef LMCU_MINB
#define LMCU_MINB 6      // resident CTAs per SM the queue kernels are compiled for (register budget: 80 registers)
#endif
#include "prep.hpp"

using namespace lmcu;

// The path / shadow queues and the hit records are written once and read once per iteration (2.4 GB per wavefront iteration):
// streaming loads / stores (ld.global.cs / st.global.cs: evict-first) keep them from pushing BVH nodes, triangles and film
// lines out of L1 / L2.  LMCU_STREAM_IO=0 builds the plain accesses (A/B).
#ifndef LMCU_STREAM_IO
#define LMCU_STREAM_IO 1
#endif
#if LMCU_STREAM_IO
#define Q_LD(p) __ldcs(p)
#define Q_ST(p, v) __stcs(p, v)
#else
#define Q_LD(p) (*(p))
#define Q_ST(p, v) (*(p) = (v))
#endif

// ================================================================================================
// error handling
// ================================================================================================
static thread_local std::string g_error;

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            return fail(LMCU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
        }                                                                                          \
    } while (0)

// ================================================================================================
// kernels: builder
// ================================================================================================
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    // ordered-int trick, valid for mixed signs
    if (v >= 0.f) atomicMin((int*)a, __float_as_int(v)); else atomicMax((unsigned int*)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.f) atomicMax((int*)a, __float_as_int(v)); else atomicMin((unsigned int*)a, __float_as_uint(v));
}

// scene_bounds: [0..2] min, [3..5] max (floats, outward); flags[0] &= exact
__global__ void k_flatten(const PrimDev* prims, int num_prims, const double* mesh_p, uint32_t nt, double* wp,
                          float* tri_lo, float* tri_hi, double* centre, float* scene_bounds, int* flags) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool exact = true;
    float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
    if (g < nt) {
        exact = flatten_body(g, prims, num_prims, mesh_p, wp, tri_lo, tri_hi, centre);
        for (int a = 0; a < 3; a++) { lo[a] = tri_lo[(size_t)g * 3 + a]; hi[a] = tri_hi[(size_t)g * 3 + a]; }
    }
    // warp reduce, one atomic per warp
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    const bool all_exact = __all_sync(0xffffffffu, exact);
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; a++) {
            if (lo[a] <= hi[a]) { atomic_min_f(scene_bounds + a, lo[a]); atomic_max_f(scene_bounds + 3 + a, hi[a]); }
        }
        if (!all_exact) atomicAnd(flags, 0);
    }
}

__global__ void k_morton(const double* centre, uint32_t nt, const float* scene_bounds, uint64_t* keys, uint32_t* vals,
                         const float* tri_lo, const float* tri_hi, double big_frac) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nt) return;
    const double slo[3] = { scene_bounds[0], scene_bounds[1], scene_bounds[2] };
    const double shi[3] = { scene_bounds[3], scene_bounds[4], scene_bounds[5] };
    keys[g] = morton_body(centre, g, slo, shi, tri_lo, tri_hi, big_frac);
    vals[g] = g;
}

__global__ void k_hierarchy(const uint64_t* keys, int n, BinNode* nodes, uint32_t* parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) hierarchy_body(i, keys, n, nodes, parent);
}

// leaves: bounds from the sorted triangles; then climb, the second arrival at a node fits it
__global__ void k_fit(BinNode* nodes, const uint32_t* parent, const uint32_t* sorted_tri, const float* tri_lo,
                      const float* tri_hi, int n, uint32_t* arrive) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t g = sorted_tri[j];
    BinNode leaf;
    for (int a = 0; a < 3; a++) { leaf.lo[a] = tri_lo[(size_t)g * 3 + a]; leaf.hi[a] = tri_hi[(size_t)g * 3 + a]; }
    leaf.left = leaf.right = 0xffffffffu;
    leaf.first = (uint32_t)j; leaf.count = 1;
    uint32_t id = (uint32_t)(n - 1 + j);
    nodes[id] = leaf;
    if (n == 1) return;
    __threadfence();
    uint32_t cur = parent[id];
    while (true) {
        const uint32_t old = atomicAdd(&arrive[cur], 1u);
        if (old == 0) return;                 // first arrival: the sibling subtree is not ready yet
        __threadfence();
        BinNode* nd = nodes + cur;
        const uint32_t l = nd->left, r = nd->right;
        for (int a = 0; a < 3; a++) {
            const float llo = __ldcg(&nodes[l].lo[a]), rlo = __ldcg(&nodes[r].lo[a]);
            const float lhi = __ldcg(&nodes[l].hi[a]), rhi = __ldcg(&nodes[r].hi[a]);
            nd->lo[a] = fminf(llo, rlo);
            nd->hi[a] = fmaxf(lhi, rhi);
        }
        __threadfence();
        if (cur == 0) return;                 // root
        cur = parent[cur];
    }
}

__global__ void k_collapse(CollapseCtx c, const CollapseItem* queue, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) collapse_body(c, queue[i]);
}

// Surface-area cost of the finished wide tree: the expected number of wide-node visits and of triangle tests for a random
// line through the root box (each node / leaf slot weighted by its box area over the root's).  out[0] += area of this
// node's box (union of its dequantised child boxes), out[1] += sum over leaf slots of area * triangle count, out[2] = root area.
__global__ void k_sah_cost(const WideNode* __restrict__ nodes, uint32_t num_nodes, double* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    double a_node = 0.0, a_leaf = 0.0;
    if (i < num_nodes) {
        const WideNode n = nodes[i];
        const float sx = __uint_as_float((uint32_t)n.ex << 23), sy = __uint_as_float((uint32_t)n.ey << 23), sz = __uint_as_float((uint32_t)n.ez << 23);
        float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
        for (int s = 0; s < 8; s++) {
            if (n.meta[s] == 0) continue;
            const float l[3] = { n.px + n.qlox[s] * sx, n.py + n.qloy[s] * sy, n.pz + n.qloz[s] * sz };
            const float h[3] = { n.px + n.qhix[s] * sx, n.py + n.qhiy[s] * sy, n.pz + n.qhiz[s] * sz };
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], l[a]); hi[a] = fmaxf(hi[a], h[a]); }
            if (!((n.imask >> s) & 1u)) a_leaf += (double)box_area(l, h) * (double)__popc((uint32_t)n.meta[s] >> 5);
        }
        a_node = lo[0] <= hi[0] ? (double)box_area(lo, hi) : 0.0;
        if (i == 0) out[2] = a_node;
    }
    for (int o = 16; o > 0; o >>= 1) { a_node += __shfl_xor_sync(0xffffffffu, a_node, o); a_leaf += __shfl_xor_sync(0xffffffffu, a_leaf, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a_node); atomicAdd(out + 1, a_leaf); }
}

__global__ void k_emit(const uint32_t* tri_order, uint32_t nt, const PrimDev* prims, int num_prims, const double* wp,
                       const double* mesh_p, const double* mesh_n, TriF32* t32, TriF64* t64, TriShade* shade,
                       const float* mesh_t, TriUV* uv) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nt) emit_body(k, tri_order, prims, num_prims, wp, mesh_p, mesh_n, t32, t64, shade, mesh_t, uv);
}

// -- two-level layout (bvh.cuh TriObj / InstDev, build.cuh) ------------------------------------------------------------------
__global__ void k_obj_bounds(const double* p, uint32_t n, float* tri_lo, float* tri_hi, double* centre, float* bounds6) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
    if (k < n) {
        obj_bounds_body(k, p, tri_lo, tri_hi, centre);
        for (int a = 0; a < 3; a++) { lo[a] = tri_lo[(size_t)k * 3 + a]; hi[a] = tri_hi[(size_t)k * 3 + a]; }
    }
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int a = 0; a < 3; a++) if (lo[a] <= hi[a]) { atomic_min_f(bounds6 + a, lo[a]); atomic_max_f(bounds6 + 3 + a, hi[a]); }
    }
}

// World-space box of every primitive: the union over its faces of the vertices the leaf test will form.  A warp's 32
// consecutive triangles nearly always belong to one primitive: then one lane issues the atomics for all of them.
__global__ void k_inst_bounds(const PrimDev* prims, int num_prims, const double* mesh_p, uint32_t nt, float* inst_lo, float* inst_hi) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
    uint32_t pi = 0xffffffffu;
    if (g < nt) pi = world_bounds_body(g, prims, num_prims, mesh_p, lo, hi);
    const uint32_t p0 = __shfl_sync(0xffffffffu, pi, 0);
    if (__all_sync(0xffffffffu, pi == p0)) {
        if (p0 == 0xffffffffu) return;
        for (int a = 0; a < 3; a++) {
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
        }
        if ((threadIdx.x & 31) != 0) return;
    }
    if (pi == 0xffffffffu) return;
    for (int a = 0; a < 3; a++) { atomic_min_f(inst_lo + (size_t)pi * 3 + a, lo[a]); atomic_max_f(inst_hi + (size_t)pi * 3 + a, hi[a]); }
}

__global__ void k_inst_centres(const float* inst_lo, const float* inst_hi, uint32_t np, double* centre, float* bounds6) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    for (int a = 0; a < 3; a++) {
        const float l = inst_lo[(size_t)i * 3 + a], h = inst_hi[(size_t)i * 3 + a];
        centre[(size_t)i * 3 + a] = ((double)l + (double)h) * 0.5;
        atomic_min_f(bounds6 + a, l); atomic_max_f(bounds6 + 3 + a, h);
    }
}

__global__ void k_emit_obj(const uint32_t* order, const double* p, uint32_t n, int32_t mesh, TriObj* out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) emit_obj_body(k, order, p, mesh, out);
}

__global__ void k_relocate(WideNode* nodes, uint32_t n, uint32_t node_off, uint32_t item_off) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) relocate_body(nodes[i], node_off, item_off);
}

// ================================================================================================
// kernels: accel batches
// ================================================================================================
// Persistent threads: each warp claims 32 consecutive rays from a global cursor.
__device__ __forceinline__ unsigned long long warp_fetch(unsigned long long* cursor) {
    unsigned long long base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(cursor, 32ull);
    return __shfl_sync(0xffffffffu, base, 0);
}

// The scalar lm::Accel::intersect contract (include/lm/accel.h:70): ONE ray, read from and answered into a pinned host
// slot the calling thread owns (mapped into the device address space: no cudaMemcpy on either side).
template <typename TriT>
__global__ void k_intersect_one(DevScene s, const LmcuRay* __restrict__ ray, LmcuHit* __restrict__ hit,
                                volatile unsigned long long* done, unsigned long long ticket) {
    if (threadIdx.x != 0) return;
    uint2 stack[kStackSize];
    RayD r;
    r.o = d3(ray->o[0], ray->o[1], ray->o[2]); r.d = d3(ray->d[0], ray->d[1], ray->d[2]);
    r.tmin = ray->tmin; r.tmax = ray->tmax;
    HitD h; h.t = r.tmax; h.u = h.v = 0; h.tri = -1;
    const bool ok = traverse<false, false, TriT>(s.nodes, scene_tris<TriT>(s), r, h, stack, nullptr, s.inst);
    LmcuHit o;
    o.t = h.t; o.u = h.u; o.v = h.v; o.primitive = ok ? -1 : -2; o.face = -1;     // -2: traversal stack overflow
    if (ok && h.tri >= 0) {
        int32_t pi;
        tri_ids(s, (uint32_t)h.tri, pi, o.face);
        o.primitive = s.prims[pi].node_index;
    }
    *hit = o;
    __threadfence_system();          // the hit record is in host memory before the ticket the host thread spins on
    *done = ticket;
}

// L2 read bandwidth probe (the denominator of the roofline of an L2-resident traversal): every CTA streams the same
// `bytes`-sized buffer (far smaller than L2) `reps` times with 16-byte loads.
__global__ void __launch_bounds__(256) k_probe_l2(const uint4* __restrict__ buf, size_t n16, int reps, uint32_t* sink) {
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++) {
        // n16 is a power of two (host-checked): the wrap is a mask, not a division
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x + (size_t)r * 7919u, k = 0; k < n16; k += (size_t)gridDim.x * blockDim.x) {
            const size_t j = (i + k) & (n16 - 1);
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + j));
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// ================================================================================================
// kernels: renderers
// ================================================================================================
template <typename TriT>
__global__ void __launch_bounds__(128) k_raycast(DevScene s, RaycastParams p, float* film, unsigned long long* cursor) {
    uint2 stack[kStackSize];
    const unsigned long long n = (unsigned long long)p.width * p.height;
    const bool tiled = (p.width % 8 == 0) && (p.height % 4 == 0);
    while (true) {
        const unsigned long long base = warp_fetch(cursor);
        if (base >= n) break;
        const unsigned long long i = base + (threadIdx.x & 31);
        if (i >= n) continue;
        int x, y;
        if (tiled) {
            const uint32_t tiles_x = (uint32_t)p.width / 8u, tile = (uint32_t)(i >> 5), idx = (uint32_t)i & 31u;
            x = (int)((tile % tiles_x) * 8u + (idx & 7u)); y = (int)((tile / tiles_x) * 4u + (idx >> 3));
        } else {
            x = (int)(i % p.width); y = (int)(i / p.width);
        }
        raycast_pixel<TriT>(s, p, x, y, film, stack);
    }
}

template <typename TriT>
__global__ void __launch_bounds__(128) k_ao(DevScene s, int width, int height, long long spp, unsigned long long sample_offset,
                                            uint32_t seed, double ao_distance, float* film, unsigned long long* cursor,
                                            unsigned long long* rays) {
    uint2 stack[kStackSize];
    const unsigned long long n = (unsigned long long)width * height;
    const bool tiled = (width % 8 == 0) && (height % 4 == 0);
    uint32_t local_rays = 0;
    while (true) {
        const unsigned long long base = warp_fetch(cursor);
        if (base >= n) break;
        const unsigned long long i = base + (threadIdx.x & 31);
        if (i >= n) continue;
        int x, y;
        if (tiled) {
            const uint32_t tiles_x = (uint32_t)width / 8u, tile = (uint32_t)(i >> 5), idx = (uint32_t)i & 31u;
            x = (int)((tile % tiles_x) * 8u + (idx & 7u)); y = (int)((tile / tiles_x) * 4u + (idx >> 3));
        } else {
            x = (int)(i % width); y = (int)(i / width);
        }
        ao_pixel<TriT>(s, width, height, x, y, spp, sample_offset, seed, ao_distance, film, stack, &local_rays);
    }
    atomicAdd(rays, (unsigned long long)local_rays);
}

// -- bdpt (bdpt.cuh) as a wavefront ---------------------------------------------------------------------------------------
// One thread per sample doing everything (the form tests/hostsim runs) executes at 3 of 32 lanes on the device: the (s, t)
// loops, their early exits and the traversals inside them diverge completely (ncu, profiles/r02_ncu_summary.md).  Here:
//   the random walks  as a wavefront over all subpaths: k_bdpt_gen, then per vertex index k_pt_extend_q -> k_bdpt_vertex,
//                     vertices + their alpha prefix and densities (bdpt_vertex_data) into SoA arrays in HBM;
//   k_bdpt_lists      per strategy, the samples whose subpaths are long enough for it;
//   k_bdpt_connect    one thread per (strategy, sample), strategy-major: every lane of a warp runs the SAME (s, t) on
//                     consecutive samples — uniform loops, coalesced loads — and appends a visibility ray carrying the
//                     weighted contribution to the shadow queue (s = 0 strategies splat at once);
//   k_pt_shadow_q     the path tracer's any-hit queue kernel: visible => splat.
struct BdptStore {
    float4* va;        // [2][mv][cap]  p.xyz, kind | comp << 8 | infinite << 16 | degenerated << 17
    float4* vb;        //               n.xyz, material
    float4* vc;        //               gn.xyz, light
    float4* vd;        //               alpha of the first k + 1 vertices, component pdf of vertex k
    float2* ve;        //               F (sign bit: the vertex is specular), R  (bdpt_subpath_pdfs)
    float2* vt;        //               texture coordinates (null when no material reads a texture)
    int32_t* nverts;   // [2][cap]
    uint32_t* eye_pixel;   // [cap]
    uint32_t cap;
    int32_t mv;
    __host__ __device__ size_t at(int side, int k, uint32_t i) const { return (size_t)(side * mv + k) * cap + i; }
};

struct BdptSoAAcc {
    BdptStore st;
    uint32_t i;
    __device__ __forceinline__ BVert load(int side, int k) const {
        const size_t o = st.at(side, k, i);
        const float4 a = st.va[o], b = st.vb[o], c = st.vc[o];
        BVert v;
        v.p = f3(a.x, a.y, a.z); v.n = f3(b.x, b.y, b.z); v.gn = f3(c.x, c.y, c.z);
        const uint32_t bits = __float_as_uint(a.w);
        v.kind = (uint8_t)(bits & 0xffu); v.comp = (uint8_t)((bits >> 8) & 0xffu);
        v.infinite = (uint8_t)((bits >> 16) & 1u); v.degenerated = (uint8_t)((bits >> 17) & 1u);
        v.mat = (int32_t)__float_as_uint(b.w); v.light = (int32_t)__float_as_uint(c.w);
        v.tu = 0.f; v.tv = 0.f;
        if (st.vt) { const float2 t = st.vt[o]; v.tu = t.x; v.tv = t.y; }
        return v;
    }
    __device__ __forceinline__ BVert Y(int k) const { return load(1, k); }
    __device__ __forceinline__ BVert Z(int k) const { return load(0, k); }
    __device__ __forceinline__ F3 alpha(int side, int l) const {
        if (l == 0) return f3s(1.f);
        const float4 d = st.vd[st.at(side, l - 1, i)];
        return f3(d.x, d.y, d.z);
    }
    __device__ __forceinline__ F3 aL(int l) const { return alpha(1, l); }
    __device__ __forceinline__ F3 aE(int l) const { return alpha(0, l); }
    __device__ __forceinline__ float pcE(int k) const { return st.vd[st.at(0, k, i)].w; }
    __device__ __forceinline__ float yF(int k) const { return fabsf(st.ve[st.at(1, k, i)].x); }
    __device__ __forceinline__ float yR(int k) const { return st.ve[st.at(1, k, i)].y; }
    __device__ __forceinline__ float zF(int k) const { return fabsf(st.ve[st.at(0, k, i)].x); }
    __device__ __forceinline__ float zR(int k) const { return st.ve[st.at(0, k, i)].y; }
    __device__ __forceinline__ bool ySpec(int k) const { return signbit(st.ve[st.at(1, k, i)].x); }
    __device__ __forceinline__ bool zSpec(int k) const { return signbit(st.ve[st.at(0, k, i)].x); }
};

__device__ __forceinline__ void bdpt_store_vertex(const BdptStore& st, int side, int k, uint32_t g, const BVert& v) {
    const size_t o = st.at(side, k, g);
    const uint32_t bits = (uint32_t)v.kind | ((uint32_t)v.comp << 8) | ((uint32_t)v.infinite << 16) | ((uint32_t)v.degenerated << 17);
    st.va[o] = make_float4(v.p.x, v.p.y, v.p.z, __uint_as_float(bits));
    st.vb[o] = make_float4(v.n.x, v.n.y, v.n.z, __uint_as_float((uint32_t)v.mat));
    st.vc[o] = make_float4(v.gn.x, v.gn.y, v.gn.z, __uint_as_float((uint32_t)v.light));
    if (st.vt) st.vt[o] = make_float2(v.tu, v.tv);
}

// The random walks as a wavefront over ALL subpaths of the pass (eye subpaths first, then the light subpaths): vertex 0 + the
// primary rays, then per vertex index k: the path tracer's extend queue kernel (k_pt_extend_q) -> k_bdpt_vertex (vertex k from
// the hit, its component, the next direction -> the next queue, ballot-compacted) -> k_bdpt_flip.  A walk that ends simply
// stops producing rays.  (A first version walked each subpath in one thread with the per-thread traversal: 5 of 32 lanes.)
__global__ void k_bdpt_reset(PtCounters* ctr, uint32_t nrays) {
    ctr->shadow_count = 0; ctr->fetch_shadow = 0; ctr->count[0] = nrays; ctr->count[1] = 0; ctr->fetch_extend = 0;
}

__global__ void k_bdpt_flip(PtCounters* ctr, int cur) {
    ctr->rays += ctr->count[cur]; ctr->extend_rays += ctr->count[cur];
    ctr->count[cur] = 0; ctr->fetch_extend = 0;
}

__global__ void __launch_bounds__(256) k_bdpt_gen(DevScene s, PtParams p, BdptStore st, float4* __restrict__ ro, float4* __restrict__ rd,
                                                  uint32_t* __restrict__ id) {
    const uint32_t n = (uint32_t)p.total_paths, total = 2u * n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int side = i >= n ? 1 : 0;
        const uint32_t g = i - (side ? n : 0u);
        BVert v;
        F3 o, d;
        bdpt_first_vertex(s, p, p.sample_offset + g, side != 0, v, o, d);
        bdpt_store_vertex(st, side, 0, g, v);
        st.nverts[(size_t)side * st.cap + g] = 1;
        {   // alpha[1], pc[0], F[0], spec[0] (bdpt_alphas / bdpt_subpath_pdfs for vertex 0)
            const bool conn0 = bv_connectable(s, v);
            const float pa = conn0 ? bv_pdf_position(s, v) : 1.f;
            const size_t o0 = st.at(side, 0, g);
            st.vd[o0] = conn0 ? make_float4(1.0f / pa, 1.0f / pa, 1.0f / pa, 1.f) : make_float4(0.f, 0.f, 0.f, 1.f);
            st.ve[o0] = make_float2(bv_specular<false>(s, v) ? -pa : pa, 0.f);
            if (!side) st.eye_pixel[g] = 0u;
        }
        // single-strategy modes walk one side only (bdpt_eye_max / bdpt_light_max): its rays fill the queue from slot 0
        if (p.bdpt_mode != 0u && (side != 0) != (p.bdpt_mode == 3u)) continue;
        const uint32_t slot = p.bdpt_mode == 3u ? g : i;
        ro[slot] = make_float4(o.x, o.y, o.z, (float)kEps);
        rd[slot] = make_float4(d.x, d.y, d.z, (float)kInf);
        id[slot] = i;
    }
}

template <bool COMP>
__global__ void __launch_bounds__(128) k_bdpt_vertex(DevScene s, PtParams p, BdptStore st, int k, const float4* __restrict__ rd_in,
                                                     const uint32_t* __restrict__ id_in, const float4* __restrict__ hits,
                                                     float4* __restrict__ ro_out, float4* __restrict__ rd_out, uint32_t* __restrict__ id_out,
                                                     PtCounters* ctr, int cur) {
    const uint32_t nq = ctr->count[cur], n = (uint32_t)p.total_paths, lane = threadIdx.x & 31u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < nq; base += gridDim.x * blockDim.x) {          // warp-uniform trip count
        const uint32_t j = base + threadIdx.x;
        bool push = false;
        uint32_t i = 0;
        BVert v;
        F3 wo;
        if (j < nq) {
            i = id_in[j];
            const int side = i >= n ? 1 : 0;
            const uint32_t g = i - (side ? n : 0u);
            const float4 dd = Q_LD(rd_in + j), hit = Q_LD(hits + j);
            if (bdpt_hit_vertex<COMP>(s, p, p.sample_offset + g, side != 0, k, __float_as_int(hit.w), hit.y, hit.z, f3(dd.x, dd.y, dd.z), v)) {
                bdpt_store_vertex(st, side, k, g, v);
                st.nverts[(size_t)side * st.cap + g] = k + 1;
                BdptSoAAcc acc;
                acc.st = st; acc.i = g;
                const BVert prev = acc.load(side, k - 1);
                BVert prev2;
                if (k >= 2) prev2 = acc.load(side, k - 2);
                // alpha[k + 1], pc[k], F[k], spec[k], R[k - 2]: everything the connection kernel needs from this vertex
                const BdptVertexData vd = bdpt_vertex_data<COMP>(s, side != 0, k, v, prev, k >= 2 ? &prev2 : nullptr, acc.alpha(side, k));
                const size_t ok = st.at(side, k, g);
                st.vd[ok] = make_float4(vd.alpha_next.x, vd.alpha_next.y, vd.alpha_next.z, vd.pc);
                st.ve[ok] = make_float2(vd.spec ? -vd.F : vd.F, 0.f);
                if (k >= 2) st.ve[st.at(side, k - 2, g)].y = vd.R_prev2;
                if (!side && k == 1) st.eye_pixel[g] = bdpt_eye_pixel(s, p, prev, v);
                if (!v.infinite && k + 1 < p.max_verts) push = bdpt_next_dir<COMP>(s, p, p.sample_offset + g, side != 0, k, v, prev, wo);
            }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, push);
        if (b) {
            uint32_t pos = 0;
            if (lane == (uint32_t)(__ffs(b) - 1)) pos = atomicAdd(&ctr->count[cur ^ 1], (uint32_t)__popc(b));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(b) - 1) + (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (push) {
                Q_ST(ro_out + pos, make_float4(v.p.x, v.p.y, v.p.z, (float)kEps));
                Q_ST(rd_out + pos, make_float4(wo.x, wo.y, wo.z, (float)kInf));
                id_out[pos] = i;
            }
        }
    }
}

// Work lists of the connection kernel: strategy (s, t) applies to a sample only if its subpaths are long enough — in an open
// scene most are not (C2: 17 of the 63 possible rays per sample at max_verts 10), and a connection kernel launched over every
// (strategy, sample) ran at 6.8 of 32 lanes.  One thread per (strategy, sample) appends the sample to the strategy's list
// (ballot-ranked, one atomic per warp).
__global__ void __launch_bounds__(256) k_bdpt_lists(PtParams p, BdptStore st, const int2* __restrict__ pairs, uint32_t* __restrict__ list,
                                                    uint32_t* __restrict__ count) {
    const int2 pr = pairs[blockIdx.y];
    const uint32_t total = (uint32_t)p.total_paths, lane = threadIdx.x & 31u;
    uint32_t* mine = list + (size_t)blockIdx.y * st.cap;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        bool ok = i < total && pr.y <= st.nverts[i] && pr.x <= st.nverts[(size_t)st.cap + i];
        if (ok && p.bdpt_mode != 0u) ok = bdpt_strategy_on(p.bdpt_mode, pr.x, pr.y, [&](int j) { return signbit(st.ve[st.at(0, j, i)].x); });
        const uint32_t b = __ballot_sync(0xffffffffu, ok);
        if (b) {
            uint32_t pos = 0;
            if (lane == (uint32_t)(__ffs(b) - 1)) pos = atomicAdd(count + blockIdx.y, (uint32_t)__popc(b));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(b) - 1) + (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (ok) mine[pos] = i;
        }
    }
}

template <bool COMP>
__global__ void __launch_bounds__(128) k_bdpt_connect(DevScene s, PtParams p, BdptStore st, const int2* __restrict__ pairs,
                                                      const uint32_t* __restrict__ list, const uint32_t* __restrict__ count,
                                                      ShadowQueue sq, float* film, PtCounters* ctr) {
    const int2 pr = pairs[blockIdx.y];
    const int sL = pr.x, t = pr.y;
    const uint32_t total = count[blockIdx.y], lane = threadIdx.x & 31u;
    const uint32_t* mine = list + (size_t)blockIdx.y * st.cap;
    for (uint32_t base = blockIdx.x * blockDim.x; base < total; base += gridDim.x * blockDim.x) {      // warp-uniform trip count
        const uint32_t j = base + threadIdx.x;
        bool push = false;
        BdptConn c;
        if (j < total) {
            BdptSoAAcc acc;
            acc.st = st; acc.i = mine[j];
            if (bdpt_connect<COMP>(s, p, sL, t, acc, st.eye_pixel[acc.i], c)) {
                if (c.shadow) push = true;
                else {
                    float* px = film + (size_t)c.pixel * 3;
                    atomicAdd(px + 0, c.C.x); atomicAdd(px + 1, c.C.y); atomicAdd(px + 2, c.C.z);
                }
            }
        }
        const uint32_t b = __ballot_sync(0xffffffffu, push);
        if (b) {
            uint32_t pos = 0;
            if (lane == (uint32_t)(__ffs(b) - 1)) pos = atomicAdd(&ctr->shadow_count, (uint32_t)__popc(b));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(b) - 1) + (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (push) {
                Q_ST(sq.ray_o + pos, c.sh_o); Q_ST(sq.ray_d + pos, c.sh_d);
                Q_ST(sq.contrib + pos, make_float4(c.C.x, c.C.y, c.C.z, __uint_as_float(c.pixel)));
            }
        }
    }
}

// -- wavefront pt ----------------------------------------------------------------------------------
// advance: single thread; sizes this iteration's refill (new camera paths up to capacity), publishes the new count
// and resets the per-iteration cursors
__global__ void k_pt_advance(PtParams p, PtCounters* ctr, int cur) {
    const uint32_t m = ctr->count[cur];
    const unsigned long long remaining = p.total_paths - ctr->next_path;
    const uint32_t room = p.capacity - m;
    const uint32_t n_new = remaining < room ? (uint32_t)remaining : room;
    ctr->gen_slot = m; ctr->gen_count = n_new; ctr->gen_first = p.path_base + ctr->next_path;
    ctr->count[cur] = m + n_new;
    ctr->next_path += n_new;
    ctr->count[cur ^ 1] = 0;
    ctr->shadow_count = 0;
    ctr->fetch_extend = 0;
    ctr->fetch_shadow = 0;
    ctr->rays += m + n_new;
    ctr->extend_rays += m + n_new;
    ctr->iterations += 1;
}

// generate: refill the current queue with new camera paths (contiguous -> coherent warps).  In image mode the camera
// vertex's own NEE connection goes to the shadow queue (ballot-ranked, one atomic per warp).
__global__ void k_pt_generate(DevScene s, PtParams p, PathQueue q, ShadowQueue sq, PtCounters* ctr) {
    const uint32_t m = ctr->gen_slot, n_new = ctr->gen_count;
    const unsigned long long first = ctr->gen_first;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_new; base += stride) {     // warp-uniform trip count
        const uint32_t i = base + threadIdx.x;
        GenOut o;
        o.shadow = false;
        if (i < n_new) o = pt_generate_body(s, p, first + i, q, m + i);
        if (p.image_mode) {
            const uint32_t b = __ballot_sync(0xffffffffu, o.shadow);
            if (b) {
                uint32_t pos = 0;
                if (lane == 0) pos = atomicAdd(&ctr->shadow_count, (uint32_t)__popc(b));
                pos = __shfl_sync(0xffffffffu, pos, 0) + (uint32_t)__popc(b & ((1u << lane) - 1u));
                if (o.shadow) { sq.ray_o[pos] = o.sh_o; sq.ray_d[pos] = o.sh_d; sq.contrib[pos] = o.sh_c; }
            }
        }
    }
}

// light tracer: refill the current queue with new light paths (renderer_lt.cpp:60-66)
__global__ void k_lt_generate(DevScene s, PtParams p, PathQueue q, PtCounters* ctr) {
    const uint32_t m = ctr->gen_slot, n_new = ctr->gen_count;
    const unsigned long long first = ctr->gen_first;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_new; i += gridDim.x * blockDim.x) lt_generate_body(s, p, first + i, q, m + i);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Two compactions in one pass (next-path queue and shadow queue): the per-warp counts travel packed
// 16+16 bits through ONE shared-memory scan.  `smem` is one of two alternating 34-word buffers: a warp can run at
// most one tile ahead of the slowest (it meets it again at the first barrier of the next tile), so the buffer of the
// previous tile needs no trailing barrier — 2 barriers per tile.
__device__ __forceinline__ void block_compact2(bool p0, bool p1, uint32_t* counter0, uint32_t* counter1,
                                               uint32_t* smem /* [34] */, uint32_t& pos0, uint32_t& pos1) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
    if (lane == 0) smem[warp] = (uint32_t)__popc(b0) | ((uint32_t)__popc(b1) << 16);
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = lane < nwarps ? smem[lane] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
            smem[32] = (total & 0xffffu) ? atomicAdd(counter0, total & 0xffffu) : 0u;
            smem[33] = (total >> 16) ? atomicAdd(counter1, total >> 16) : 0u;
        }
        __syncwarp();
        if (lane < nwarps) smem[lane] = incl - v;
    }
    __syncthreads();
    const uint32_t ex = smem[warp], lt = (1u << lane) - 1u;
    pos0 = smem[32] + (ex & 0xffffu) + (uint32_t)__popc(b0 & lt);
    pos1 = smem[33] + (ex >> 16) + (uint32_t)__popc(b1 & lt);
}

// light tracer: one vertex per path — Russian roulette, component, camera connection -> shadow queue, next direction ->
// next queue.  Same two-queue compaction as k_pt_shade; nothing is added to the film here (an emitter is never "hit").
template <bool COMPOSITE>
__global__ void __launch_bounds__(LMCU_SHADE_BLOCK, LMCU_SHADE_MINB) k_lt_shade(DevScene s, PtParams p, PathQueue qin, PathQueue qout, ShadowQueue sq,
                                                  const float4* hits, PtCounters* ctr, int cur) {
    __shared__ uint32_t smem[2][34];
    const uint32_t n = ctr->count[cur];
    uint32_t tile = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x, tile ^= 1u) {
        const uint32_t i = base + threadIdx.x;
        ShadeOut o;
        o.alive = false; o.shadow = false;
        if (i < n) o = lt_shade_body<COMPOSITE>(s, p, qin.ray_o[i], qin.ray_d[i], qin.thr_pix[i], qin.info[i], hits[i]);
        uint32_t pos, spos;
        block_compact2(o.alive, o.shadow, &ctr->count[cur ^ 1], &ctr->shadow_count, smem[tile], pos, spos);
        if (o.alive) {
            qout.ray_o[pos] = o.o; qout.ray_d[pos] = o.d; qout.thr_pix[pos] = o.thr_pix; qout.aux[pos] = o.aux;
            qout.info[pos] = o.info;
        }
        if (o.shadow) {
            sq.ray_o[spos] = o.sh_o; sq.ray_d[spos] = o.sh_d; sq.contrib[spos] = o.sh_c;
        }
    }
}

template <bool COMPOSITE, bool INST = COMPOSITE>
__global__ void __launch_bounds__(LMCU_SHADE_BLOCK, LMCU_SHADE_MINB) k_pt_shade(DevScene s, PtParams p, PathQueue qin, PathQueue qout, ShadowQueue sq,
                                                  const float4* hits, float* film, PtCounters* ctr, int cur) {
    __shared__ uint32_t smem[2][34];
    const uint32_t n = ctr->count[cur];
    uint32_t tile = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x, tile ^= 1u) {
        const uint32_t i = base + threadIdx.x;
        // (Measured and dropped, profiles/r02_ab_shade_pf.log: reading the NEXT tile's hit triangle here and pulling its
        // geometry + shading records toward L1 after the shading below — shade 27 % -> 37 % of the step, C2 1920 -> 1768.)
        {   // the queues stream from HBM: pull the next grid-stride tile of this CTA into L2 while this one is shaded
            const uint32_t nx = i + gridDim.x * blockDim.x;
            if (nx < n && (threadIdx.x & 7u) == 0u) {          // one prefetch per 128-byte line of float4s
                prefetch_l2(qin.ray_o + nx); prefetch_l2(qin.ray_d + nx); prefetch_l2(qin.thr_pix + nx);
                prefetch_l2(qin.aux + nx); prefetch_l2(hits + nx);
                if ((threadIdx.x & 15u) == 0u) prefetch_l2(qin.info + nx);
            }
        }
        ShadeOut o;
        o.alive = false; o.shadow = false;
        if (i < n) {
            const float4 tp = Q_LD(qin.thr_pix + i);
            o = pt_shade_body<COMPOSITE, INST>(s, p, Q_LD(qin.ray_o + i), Q_LD(qin.ray_d + i), tp, Q_LD(qin.aux + i), Q_LD(qin.info + i), Q_LD(hits + i));
            if (!is_zero(o.emit)) {
                float* px = film + (size_t)f2u(tp.w) * 3;
                atomicAdd(px + 0, o.emit.x); atomicAdd(px + 1, o.emit.y); atomicAdd(px + 2, o.emit.z);
            }
        }
        uint32_t pos, spos;
        block_compact2(o.alive, o.shadow, &ctr->count[cur ^ 1], &ctr->shadow_count, smem[tile], pos, spos);
        if (o.alive) {
            Q_ST(qout.ray_o + pos, o.o); Q_ST(qout.ray_d + pos, o.d); Q_ST(qout.thr_pix + pos, o.thr_pix); Q_ST(qout.aux + pos, o.aux);
            Q_ST(qout.info + pos, o.info);
        }
        if (o.shadow) {
            Q_ST(sq.ray_o + spos, o.sh_o); Q_ST(sq.ray_d + spos, o.sh_d); Q_ST(sq.contrib + spos, o.sh_c);
        }
    }
}

// -- queue kernels on the warp-persistent traversal (trace.cuh) -----------------------------------------
template <typename TriT> constexpr int trav_minb() { return TriT::kTwoLevel ? LMCU_MINB_TWO : LMCU_MINB; }

struct BatchIO {                       // accel batches: LmcuRay in, LmcuHit out (include/lmcu.h)
    static constexpr bool kExactT = true;
    DevScene s;
    const LmcuRay* rays;
    LmcuHit* hits;
    __device__ __forceinline__ bool load(uint32_t i, RayD& r, RayF& f, float& tmin_dn, float& tmax_up) const {
        const double2* src = reinterpret_cast<const double2*>(rays + i);
        const double2 a = src[0], b = src[1], c = src[2], d = src[3];
        r.o = d3(a.x, a.y, b.x); r.d = d3(b.y, c.x, c.y); r.tmin = d.x; r.tmax = d.y;
        f = make_rayf(r);
        tmin_dn = f32_down(r.tmin); tmax_up = f32_up(r.tmax);
        return true;
    }
    __device__ __forceinline__ void store(uint32_t i, double t, double u_raw, double v_raw, double ad, int32_t tri) const {
        int32_t prim = -1, face = -1;
        double u = 0.0, v = 0.0;
        if (tri >= 0) {
            int32_t pi;
            tri_ids(s, (uint32_t)tri, pi, face);
            prim = s.prims[pi].node_index;
            u = divrn(u_raw, ad); v = divrn(v_raw, ad);         // accel_sahbvh.cpp:74, the same IEEE quotients
        }
        double2* dst = reinterpret_cast<double2*>(hits + i);
        dst[0] = make_double2(t, u);
        double2 o2; o2.x = v;
        const int2 pf = make_int2(prim, face);
        memcpy(&o2.y, &pf, 8);
        dst[1] = o2;
    }
};

template <typename TriT, bool COUNT>
__global__ void __launch_bounds__(LMCU_TRAV_BLOCK, trav_minb<TriT>()) k_intersect_q(DevScene s, const LmcuRay* __restrict__ rays, uint32_t n,
                                                     LmcuHit* __restrict__ hits, uint32_t* cursor,
                                                     unsigned long long* counters, uint32_t* overflow) {
    BatchIO io; io.s = s; io.rays = rays; io.hits = hits;
    TraverseCounters tc; tc.nodes = tc.tris = 0;
    trace_queue<false, COUNT, TriT>(s.nodes, scene_tris<TriT>(s), n, cursor, io, overflow, &tc, s.fetch_threshold, s.inst);
    if (COUNT) { atomicAdd(counters, (unsigned long long)tc.nodes); atomicAdd(counters + 1, (unsigned long long)tc.tris); }
}

struct ExtendIO {                      // pt extend: path queue in, float4 hit record out
    static constexpr bool kExactT = false;
    const float4* ray_o; const float4* ray_d;
    float4* hits;
    __device__ __forceinline__ bool load(uint32_t i, RayD& r, RayF& f, float& tmin_dn, float& tmax_up) const {
        const float4 o = Q_LD(ray_o + i), d = Q_LD(ray_d + i);
        r.o = d3(o.x, o.y, o.z); r.d = d3(d.x, d.y, d.z); r.tmin = o.w; r.tmax = d.w;
        f = make_rayf_f(o.x, o.y, o.z, d.x, d.y, d.z);          // the queue holds float rays: exact in both views
        tmin_dn = o.w; tmax_up = d.w;
        return true;
    }
    __device__ __forceinline__ void store(uint32_t i, double t, double u_raw, double v_raw, double ad, int32_t tri) const {
        // the path tracer shades in FP32: the barycentrics are formed in FP32 here (two FP64 divisions cost ~60
        // instructions at the 2-3 lanes that finish per iteration, 14 % of this kernel's issue slots)
        const float inv = tri >= 0 ? __frcp_rn((float)ad) : 0.f;
        Q_ST(hits + i, make_float4((float)t, (float)u_raw * inv, (float)v_raw * inv, __int_as_float(tri)));
    }
};

template <typename TriT, bool COUNT>
__global__ void __launch_bounds__(LMCU_TRAV_BLOCK, trav_minb<TriT>()) k_pt_extend_q(DevScene s, PathQueue q, float4* hits, PtCounters* ctr, int cur) {
    ExtendIO io; io.ray_o = q.ray_o; io.ray_d = q.ray_d; io.hits = hits;
    TraverseCounters tc; tc.nodes = tc.tris = 0;
    trace_queue<false, COUNT, TriT>(s.nodes, scene_tris<TriT>(s), ctr->count[cur], &ctr->fetch_extend, io, &ctr->overflow, &tc, s.fetch_threshold, s.inst);
    if (COUNT) { atomicAdd(&ctr->nodes_extend, (unsigned long long)tc.nodes); atomicAdd(&ctr->tris_extend, (unsigned long long)tc.tris); }
}

struct ShadowIO {                      // pt shadow: visible (include/lm/scene.h:343-364) => splat the NEE contribution
    static constexpr bool kExactT = false;
    const float4* ray_o; const float4* ray_d; const float4* contrib;
    float* film;
    __device__ __forceinline__ bool load(uint32_t i, RayD& r, RayF& f, float& tmin_dn, float& tmax_up) const {
        const float4 o = Q_LD(ray_o + i), d = Q_LD(ray_d + i);
        r.o = d3(o.x, o.y, o.z); r.d = d3(d.x, d.y, d.z); r.tmin = o.w; r.tmax = d.w;
        f = make_rayf_f(o.x, o.y, o.z, d.x, d.y, d.z);          // the queue holds float rays: exact in both views
        tmin_dn = o.w; tmax_up = d.w;
        return true;
    }
    __device__ __forceinline__ void store(uint32_t i, double, double, double, double, int32_t tri) const {
        if (tri < 0) {
            const float4 c = Q_LD(contrib + i);
            float* px = film + (size_t)__float_as_uint(c.w) * 3;
            atomicAdd(px + 0, c.x); atomicAdd(px + 1, c.y); atomicAdd(px + 2, c.z);
        }
    }
};

template <typename TriT, bool COUNT>
__global__ void __launch_bounds__(LMCU_TRAV_BLOCK, trav_minb<TriT>()) k_pt_shadow_q(DevScene s, ShadowQueue sq, float* film, PtCounters* ctr) {
    const uint32_t n = ctr->shadow_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr->rays += n; ctr->shadow_rays += n; }
    ShadowIO io; io.ray_o = sq.ray_o; io.ray_d = sq.ray_d; io.contrib = sq.contrib; io.film = film;
    TraverseCounters tc; tc.nodes = tc.tris = 0;
    trace_queue<true, COUNT, TriT>(s.nodes, scene_tris<TriT>(s), n, &ctr->fetch_shadow, io, &ctr->overflow, &tc, s.fetch_threshold, s.inst);
    if (COUNT) { atomicAdd(&ctr->nodes_shadow, (unsigned long long)tc.nodes); atomicAdd(&ctr->tris_shadow, (unsigned long long)tc.tris); }
}

// -- the same three queue kernels on the shared-memory ray pool (trace2.cuh): K = 2 rays per lane ------------------------
constexpr int kPoolK = 2;

template <typename IO>
__device__ __forceinline__ void pool_pointers(float4* smem, uint2* arena, float4*& pool, uint2*& stack) {
    const uint32_t w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    pool = smem + (size_t)w * (pool_bytes_per_warp<IO, kPoolK>() / 16);
    stack = arena + (size_t)(blockIdx.x * wpb + w) * ((size_t)kStackSize * 32 * kPoolK);
}

template <typename TriT, bool COUNT>
__global__ void __launch_bounds__(LMCU_TRAV_BLOCK, 5) k_intersect_p(DevScene s, const LmcuRay* __restrict__ rays, uint32_t n,
                                                     LmcuHit* __restrict__ hits, uint32_t* cursor,
                                                     unsigned long long* counters, uint32_t* overflow, uint2* arena) {
    extern __shared__ float4 pool_smem[];
    float4* pool; uint2* stack;
    pool_pointers<BatchIO>(pool_smem, arena, pool, stack);
    BatchIO io; io.s = s; io.rays = rays; io.hits = hits;
    TraverseCounters tc; tc.nodes = tc.tris = 0;
    trace_pool<false, COUNT, kPoolK, TriT>(s.nodes, scene_tris<TriT>(s), n, cursor, io, overflow, &tc, s.pool_refill, pool, stack);
    if (COUNT) { atomicAdd(counters, (unsigned long long)tc.nodes); atomicAdd(counters + 1, (unsigned long long)tc.tris); }
}

template <typename TriT, bool COUNT>
__global__ void __launch_bounds__(LMCU_TRAV_BLOCK, LMCU_MINB) k_pt_extend_p(DevScene s, PathQueue q, float4* hits, PtCounters* ctr, int cur, uint2* arena) {
    extern __shared__ float4 pool_smem[];
    float4* pool; uint2* stack;
    pool_pointers<ExtendIO>(pool_smem, arena, pool, stack);
    ExtendIO io; io.ray_o = q.ray_o; io.ray_d = q.ray_d; io.hits = hits;
    TraverseCounters tc; tc.nodes = tc.tris = 0;
    trace_pool<false, COUNT, kPoolK, TriT>(s.nodes, scene_tris<TriT>(s), ctr->count[cur], &ctr->fetch_extend, io, &ctr->overflow, &tc, s.pool_refill, pool, stack);
    if (COUNT) { atomicAdd(&ctr->nodes_extend, (unsigned long long)tc.nodes); atomicAdd(&ctr->tris_extend, (unsigned long long)tc.tris); }
}

template <typename TriT, bool COUNT>
__global__ void __launch_bounds__(LMCU_TRAV_BLOCK, LMCU_MINB) k_pt_shadow_p(DevScene s, ShadowQueue sq, float* film, PtCounters* ctr, uint2* arena) {
    extern __shared__ float4 pool_smem[];
    float4* pool; uint2* stack;
    pool_pointers<ShadowIO>(pool_smem, arena, pool, stack);
    const uint32_t n = ctr->shadow_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr->rays += n; ctr->shadow_rays += n; }
    ShadowIO io; io.ray_o = sq.ray_o; io.ray_d = sq.ray_d; io.contrib = sq.contrib; io.film = film;
    TraverseCounters tc; tc.nodes = tc.tris = 0;
    trace_pool<true, COUNT, kPoolK, TriT>(s.nodes, scene_tris<TriT>(s), n, &ctr->fetch_shadow, io, &ctr->overflow, &tc, s.pool_refill, pool, stack);
    if (COUNT) { atomicAdd(&ctr->nodes_shadow, (unsigned long long)tc.nodes); atomicAdd(&ctr->tris_shadow, (unsigned long long)tc.tris); }
}

__global__ void k_fill(float* film, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) film[i] = v;
}

__global__ void k_scale(float* film, size_t n, float s) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) film[i] *= s;
}

// Order-independent 64-bit digest of an array of 8-byte words (mesh positions / normals): sum of a splitmix64 mix of
// (word, index).  Identifies the geometry an accel blob was built from (lmcu_accel_save / lmcu_accel_load).
__global__ void k_digest(const unsigned long long* __restrict__ words, size_t n, unsigned long long* out) {
    unsigned long long h = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = words[i] ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1));
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        h += z ^ (z >> 31);
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}

// The reference's film is Float = double (film_bitmap.cpp:20-55).  Here the kernels splat with fp32 atomics, and a running
// fp32 sum of thousands of samples would start dropping small contributions (ulp(2^16) = 0.008).  So the fp32 film only
// ever holds the last few dozen samples per pixel: this kernel MOVES what has accumulated into an FP64 film and may run at
// any time beside the render kernels — atomicExch hands every splat to exactly one fold (a concurrent splat lands either
// before the exchange and is taken now, or after it and is taken by the next fold).
__global__ void k_film_fold(float* __restrict__ film, double* __restrict__ acc, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = atomicExch(film + i, 0.f);
        if (v != 0.f) atomicAdd(acc + i, (double)v);
    }
}

// film = float(acc * scale): the single rescale of renderer_pt.cpp:307-312 in FP64 (scale = 1 leaves sums, sample split)
__global__ void k_film_finish(float* __restrict__ film, const double* __restrict__ acc, size_t n, double scale) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) film[i] = (float)(acc[i] * scale);
}

// root = (root + sum over peers) * scale; the peer pointers are NVLink peer-mapped (P2P / CUDA IPC), so the
// loads below travel over NVLink while the FADDs and the store run here: reduce + film::rescale in one pass.
struct PeerFilms { const float* p[LMCU_MAX_PEERS]; };

__global__ void __launch_bounds__(256) k_film_reduce_peers(float* __restrict__ root, PeerFilms peers, int num_peers, size_t n, float scale) {
    const size_t n4 = n / 4;
    float4* root4 = reinterpret_cast<float4*>(root);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = root4[i];
        for (int k = 0; k < num_peers; k++) {
            const float4 v = __ldcv(reinterpret_cast<const float4*>(peers.p[k]) + i);   // volatile-cached: peer memory is not coherent in L1
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        root4[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const size_t i = n4 * 4 + threadIdx.x;
        float acc = root[i];
        for (int k = 0; k < num_peers; k++) acc += __ldcv(peers.p[k] + i);
        root[i] = acc * scale;
    }
}

// Slice-parallel form of the same combine: every GPU runs this kernel on ITS slice [begin, end) of the film, reading that
// slice of all films (its own + the peers' over NVLink) and storing the scaled sum into `out` — the root's film, a peer
// store over NVLink for every rank but the root.  All links carry traffic at once and no GPU takes more than (G-1)/G of
// a film inbound, instead of G-1 whole films converging on the root: reduce-scatter + gather-to-root + film::rescale
// in one pass.  Only this rank touches slice [begin, end) of `out`, so reading it as an input first is safe.
struct FilmSet { const float* p[LMCU_MAX_PEERS + 1]; };

__global__ void __launch_bounds__(256) k_film_reduce_slice(FilmSet films, int num_films, float* __restrict__ out, size_t begin, size_t end, float scale) {
    const size_t b4 = begin / 4, e4 = end / 4;      // begin is a multiple of 4 (host-checked)
    for (size_t i = b4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < e4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < num_films; k++) {
            const float4 v = __ldcv(reinterpret_cast<const float4*>(films.p[k]) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    }
    if (blockIdx.x == 0 && e4 * 4 + threadIdx.x < end) {
        const size_t i = e4 * 4 + threadIdx.x;
        float acc = 0.f;
        for (int k = 0; k < num_films; k++) acc += __ldcv(films.p[k] + i);
        out[i] = acc * scale;
    }
}

// ================================================================================================
// host side
// ================================================================================================
// Scratch allocations (path pools, films) are recycled across scenes of a process: cudaMalloc / cudaFree of
// the 1.7 GB of path queues cost far more than a BVH build, and a host that re-creates its scene per render
// (the Lightmetrica plugin does) would pay it every time.  Freed blocks of every size — the builder's small temporaries too: next to gigabytes of live path
// pools a cudaFree / cudaMalloc pair costs 0.1-0.2 ms and now and then hundreds (tools/build_probe.py) — are parked here per
// (device, size), at most kCacheBytes per process, and handed out again on the next equal-sized request.
namespace {
struct BlockCache {
    struct Parked { void* p; uint64_t seq; };
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, Parked> blocks;
    std::map<int, size_t> bytes;           // parked bytes per device
    uint64_t clock = 0;
    size_t cap = size_t(24) << 30;         // per DEVICE; LMCU_CACHE_GB overrides it (0 = park nothing)
    static constexpr size_t kMaxBlocks = 8192;
    BlockCache() { if (const char* e = getenv("LMCU_CACHE_GB")) cap = (size_t)(atof(e) * double(size_t(1) << 30)); }
    void* take(int device, size_t n) {
        std::lock_guard<std::mutex> lk(mu);
        auto it = blocks.find({ device, n });
        if (it == blocks.end()) return nullptr;
        void* p = it->second.p;
        blocks.erase(it);
        bytes[device] -= n;
        return p;
    }
    // gives every parked block of `device` back to the driver (an allocation failed: the memory is needed now)
    void flush(int device) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto it = blocks.begin(); it != blocks.end();) {
            if (it->first.first == device && !exported.count(it->second.p)) {
                cudaFree(it->second.p); bytes[device] -= it->first.second; it = blocks.erase(it);
            }
            else ++it;
        }
    }
    // A block that was exported over CUDA IPC may still be mapped by a peer process — freeing it under that mapping is
    // undefined behaviour — so it is parked whatever its size and the cap, for the life of the process.
    std::set<void*> exported;
    void mark_exported(void* p) { std::lock_guard<std::mutex> lk(mu); exported.insert(p); }
    // Over the cap the blocks parked LONGEST ago make room (a process that built a 50 M-triangle scene once would otherwise
    // sit on its temporaries for good while the path pools of every later scene go back and forth to the driver).
    bool park(int device, size_t n, void* p) {
        std::lock_guard<std::mutex> lk(mu);
        const bool keep = exported.count(p) != 0;
        if (!keep) {
            if (n > cap) return false;
            while (bytes[device] + n > cap || blocks.size() >= kMaxBlocks) {
                auto victim = blocks.end();
                for (auto it = blocks.begin(); it != blocks.end(); ++it) {
                    if (it->first.first != device || exported.count(it->second.p)) continue;
                    if (victim == blocks.end() || it->second.seq < victim->second.seq) victim = it;
                }
                if (victim == blocks.end()) return false;
                cudaFree(victim->second.p);
                bytes[device] -= victim->first.second;
                blocks.erase(victim);
            }
        }
        blocks.insert({ { device, n }, Parked{ p, ++clock } });
        bytes[device] += n;
        return true;
    }
};
// Never destroyed: a host may still release scenes from its own static destructors at process exit (the reference's
// UserContext singleton owns the asset tree and outlives every library loaded after lm::init()).
BlockCache& g_cache = *new BlockCache;
}  // namespace

template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    int device = -1;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        cudaGetDevice(&device);
        p = (T*)g_cache.take(device, count * sizeof(T));
        if (p) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaErrorMemoryAllocation) {            // out of memory with blocks parked in the cache: release them, retry once
            cudaGetLastError();
            g_cache.flush(device);
            e = cudaMalloc((void**)&p, count * sizeof(T));
        }
        return e;
    }
    void release() {
        if (p && !g_cache.park(device, n * sizeof(T), p)) cudaFree(p);
        p = nullptr; n = 0;
    }
    ~DBuf() { release(); }
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
};

struct LmcuScene {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    // description (device)
    DBuf<double> mesh_p, mesh_n;
    DBuf<PrimDev> prims;
    DBuf<MaterialDev> mats;
    DBuf<LightDev> lights;
    DBuf<EmitTri> emit;
    DBuf<float> cdf;
    DBuf<float4> env_tex;
    DBuf<float> mesh_t;                  // texture coordinates per mesh face (only when a material reads a texture)
    DBuf<TextureDev> textures;
    DBuf<float4> texels;
    DBuf<MixEntry> mix;
    std::vector<PrimDev> h_prims;
    uint32_t nt = 0;
    int num_prims = 0, num_lights = 0, env_light = -1;
    CameraDev cam{};
    bool has_camera = false;
    // accel
    DBuf<WideNode> nodes;
    DBuf<TriF32> t32;
    DBuf<TriF64> t64;
    DBuf<TriShade> shade;
    DBuf<TriUV> uv;
    DBuf<TriObj> tobj;                   // two-level layout (bvh.cuh): object-space triangles of the mesh trees, leaf order
    DBuf<InstDev> inst;                  // ... instance records in the top level's leaf order
    bool two_level = false;
    int tri_kind() const { return two_level ? 2 : fp32 ? 1 : 0; }      // LMCU_DISPATCH: TriObj / TriF32 / TriF64 instantiations
    int trav_ctas(bool wavefront = false) const {
        return (two_level ? std::min(ctas_per_sm, LMCU_MINB_TWO) : ctas_per_sm) + (wavefront ? spare_ctas : 0);
    }   // CTAs per SM of the persistent traversal grids
    bool general() const { return composite || two_level; }            // the <true> instantiations of the shade / bdpt kernels: textured or
                                                                       // mixture materials, or surface points resolved through instances
    bool pool_on() const { return pool_mode && !two_level; }           // the shared-memory ray pool (trace2.cuh) walks one-level trees only
    std::vector<uint32_t> face_offset;   // per mesh, first face in mesh_p / mesh_n / mesh_t
    bool built = false, fp32 = false;
    float bounds[6] = { 0, 0, 0, 0, 0, 0 };   // scene AABB (float, outward): Light::set_scene_bound for the light tracer
    uint64_t blob_desc_digest = 0;       // accel-only scene restored from a blob: the digest that blob carried
    bool composite = false;              // a material is textured or a mixture: the shade kernel's <true> instantiation
    // scratch / outputs
    DBuf<unsigned long long> cursor;     // [0] work cursor, [1..2] counters, [3] rays, [4..5] digests, [8..] per-launch cursors
    static constexpr int kCursorRing = 256;  // accel batches: every launch takes the next cursor word, so that batches in flight
    std::atomic<unsigned> cursor_next{ 0 };  // on different streams (lmcu_accel_intersect_device) never share one
    DBuf<uint32_t> flags;                // [0] stack overflow
    DBuf<float> film;
    DBuf<double> film_acc;               // pt: FP64 sums the fp32 film is folded into (k_film_fold)
    struct BdptPools {                   // bdpt: subpath SoA (BdptStore), strategy table, shadow queue, counters
        DBuf<float4> va, vb, vc, vd, sq[3];
        DBuf<float4> q_o[2], q_d[2], hits;        // the random walks' ray queue (ping-pong) and hit records, 2 x cap entries
        DBuf<uint32_t> q_id[2];
        DBuf<float2> ve, vt;
        DBuf<int32_t> nverts;
        DBuf<uint32_t> eye_pixel;
        DBuf<int2> pairs;
        DBuf<uint32_t> pair_list, pair_count;      // per strategy: the samples whose subpaths are long enough for it
        DBuf<PtCounters> ctr;
        uint32_t cap = 0; int mv = 0; size_t sq_cap = 0;
    } bdpt;
    int film_w = 0, film_h = 0;
    // pt pools: one set per wavefront (LMCU_WAVES staggered wavefronts, each on its own stream)
    struct PtWave {
        DBuf<float4> pq[2][4];
        DBuf<uint2> pinfo[2];
        DBuf<float4> hits;
        DBuf<float4> sq[3];
        DBuf<PtCounters> ctr;
        DBuf<uint2> arena;                   // pool mode: traversal stacks [warp][depth][slot] of this wavefront's kernels
        cudaStream_t stream = nullptr;       // wave 0 runs on the scene's stream, the others on their own
        cudaEvent_t done = nullptr;
        // two groups of iterations are in flight per wavefront; per slot: per-kernel timing events, a "group complete"
        // event and the counters read back after the group (PINNED: a D2H copy into pageable memory blocks the host
        // until the stream drains, which would serialise the wavefronts)
        std::vector<cudaEvent_t> ring[2];
        cudaEvent_t grp[2] = { nullptr, nullptr };
        PtCounters* h = nullptr;             // [2]
        int cur = 0;                         // ping-pong queue index of the next iteration to enqueue
        double since_fold = 0.0;             // samples per pixel this wavefront may have added to the fp32 film since its last fold
        int cur_after[2] = { 0, 0 };         // ... as it stands after the group of a slot
        bool finished = false;
    };
    static constexpr int kMaxWaves = 4;
    PtWave wave[kMaxWaves];
    int pt_waves = 0;
    int waves = 2;                           // wavefronts of the current render (LmcuRenderParams::wavefronts, LMCU_WAVES)
    int default_waves = 2;
    cudaEvent_t fork = nullptr;
    uint32_t pt_capacity = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int fetch_threshold = 12;            // LMCU_FETCH
    uint32_t path_chunk = 64;            // LMCU_CHUNK: samples of one pixel on consecutive path indices (kernels.cuh path_pixel)
    int pool_mode = 0;                   // LMCU_TRACE=pool: the queue kernels run on the shared-memory ray pool (trace2.cuh)
    int pool_refill = 16;                // LMCU_POOL_FETCH
    DBuf<uint2> batch_arena;             // pool mode: traversal stacks of the accel-batch kernel
    int shade_ctas_per_sm = LMCU_SHADE_MINB;     // LMCU_SHADE_GRID: CTAs per SM in the shade kernel's grid-stride grid
    int ctas_per_sm = LMCU_MINB;                 // LMCU_CTAS: resident 128-thread CTAs per SM for the persistent traversal grids
    int spare_ctas = 2;                          // LMCU_SPARE_CTAS: the wavefront loops launch this many MORE traversal CTAs per SM than are
                                                 // resident; they start as the first ones run out of rays and fill the tail of a launch
                                                 // next to the other wavefront's kernels (grids of 6 / 7 / 8 / 10 / 12 per SM: 1930 / 1945 /
                                                 // 1946 / 1945 / 1944 Mpaths/s on C2; a batch launch on its own loses 1-2 % to them:
                                                 // profiles/r02_ab_knobs.log)

    DevScene dev() const {
        DevScene d{};
        d.nodes = nodes.p; d.t32 = t32.p; d.t64 = t64.p; d.shade = shade.p; d.prims = prims.p; d.mats = mats.p;
        d.lights = lights.p; d.emit = emit.p; d.cdf = cdf.p; d.env_tex = env_tex.p;
        if (two_level) { d.tobj = tobj.p; d.inst = inst.p; d.mesh_p = mesh_p.p; d.mesh_n = mesh_n.p; d.mesh_t = mesh_t.p; }
        d.uv = uv.p; d.textures = textures.p; d.texels = texels.p; d.mix = mix.p;
        d.num_prims = num_prims; d.num_lights = num_lights; d.env_light = env_light; d.fetch_threshold = fetch_threshold; d.num_tris = nt; d.cam = cam;
        d.pool_refill = pool_refill;
        {   // light_envconst.cpp:37-43: sphere around the scene AABB, radius x 1.01
            float r2 = 0.f;
            for (int a = 0; a < 3; a++) {
                d.bound_center[a] = 0.5f * (bounds[a] + bounds[3 + a]);
                const float e = bounds[3 + a] - d.bound_center[a];
                r2 += e * e;
            }
            d.bound_radius = std::sqrt(r2) * 1.01f;
        }
        return d;
    }
};

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

extern "C" {

const char* lmcu_last_error(void) { return g_error.c_str(); }
const char* lmcu_version(void) { return "lmcu 0.1 (sm_100a)"; }

int lmcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lmcu_scene_create(int device, LmcuScene** out) {
    if (!out) return fail(LMCU_ERR_INVALID_ARGUMENT, "out is null");
    const int n = lmcu_device_count();
    if (n == 0) return fail(LMCU_ERR_NO_DEVICE, "no CUDA device: liblmcu has no CPU fallback");
    if (device < 0 || device >= n) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad device index");
    CU(cudaSetDevice(device));
    struct Guard { LmcuScene* s; ~Guard() { if (s) lmcu_scene_destroy(s); } } guard{ new LmcuScene };   // no leak on an error return
    LmcuScene* s = guard.s;
    s->device = device;
    CU(cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, device));
    CU(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    CU(cudaEventCreate(&s->ev0));
    CU(cudaEventCreate(&s->ev1));
    if (const char* e = getenv("LMCU_FETCH")) s->fetch_threshold = atoi(e);
    if (const char* e = getenv("LMCU_TRACE")) s->pool_mode = std::string(e) == "pool" ? 1 : 0;
    if (const char* e = getenv("LMCU_POOL_FETCH")) s->pool_refill = std::max(1, atoi(e));
    if (const char* e = getenv("LMCU_CHUNK")) { uint32_t c = 1; while (c * 2 <= (uint32_t)std::max(1, atoi(e)) && c < 4096) c *= 2; s->path_chunk = c; }
    if (const char* e = getenv("LMCU_CTAS")) s->ctas_per_sm = std::max(1, atoi(e));
    if (const char* e = getenv("LMCU_SPARE_CTAS")) s->spare_ctas = std::max(0, atoi(e));
    if (const char* e = getenv("LMCU_SHADE_GRID")) s->shade_ctas_per_sm = std::max(1, atoi(e));
    if (const char* e = getenv("LMCU_WAVES")) s->default_waves = std::min((int)LmcuScene::kMaxWaves, std::max(1, atoi(e)));
    CU(s->cursor.alloc(8 + LmcuScene::kCursorRing));
    CU(s->flags.alloc(4));
    CU(cudaMemset(s->cursor.p, 0, (8 + LmcuScene::kCursorRing) * sizeof(unsigned long long)));
    CU(cudaMemset(s->flags.p, 0, 4 * sizeof(uint32_t)));
    // cudaMemset is asynchronous and runs on the legacy default stream, which the scene's non-blocking streams do not wait for
    CU(cudaStreamSynchronize(0));
    guard.s = nullptr;
    *out = s;
    return LMCU_OK;
}

void lmcu_scene_destroy(LmcuScene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    for (auto& w : s->wave) {
        for (int k = 0; k < 2; k++) {
            for (cudaEvent_t e : w.ring[k]) cudaEventDestroy(e);
            if (w.grp[k]) cudaEventDestroy(w.grp[k]);
        }
        if (w.done) cudaEventDestroy(w.done);
        if (w.h) cudaFreeHost(w.h);
        if (w.stream) cudaStreamDestroy(w.stream);
    }
    if (s->fork) cudaEventDestroy(s->fork);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

int lmcu_scene_set_stream(LmcuScene* s, void* stream) {
    if (!s) return fail(LMCU_ERR_INVALID_ARGUMENT, "null scene");
    s->stream = stream ? (cudaStream_t)stream : s->own_stream;
    return LMCU_OK;
}

// -- upload -----------------------------------------------------------------------------------------
int lmcu_scene_upload(LmcuScene* s, const LmcuSceneDesc* d) {
    if (!s || !d) return fail(LMCU_ERR_INVALID_ARGUMENT, "null argument");
    CU(cudaSetDevice(s->device));
    PreparedScene ps;
    std::string err;
    const int rc = prepare_scene(d, ps, err);
    if (rc) return fail(rc, err);
    const size_t total_faces = ps.face_offset.back();
    CU(s->mesh_p.alloc(total_faces * 9));
    CU(s->mesh_n.alloc(total_faces * 9));
    for (int i = 0; i < d->num_meshes; i++) {
        const size_t bytes = (size_t)d->meshes[i].num_faces * 9 * sizeof(double);
        CU(cudaMemcpy(s->mesh_p.p + (size_t)ps.face_offset[i] * 9, d->meshes[i].p, bytes, cudaMemcpyHostToDevice));
        if (d->meshes[i].n) {
            CU(cudaMemcpy(s->mesh_n.p + (size_t)ps.face_offset[i] * 9, d->meshes[i].n, bytes, cudaMemcpyHostToDevice));
        } else {
            // never read by the kernels (has_normals = 0), but part of the description digest of lmcu_accel_save / load
            CU(cudaMemset(s->mesh_n.p + (size_t)ps.face_offset[i] * 9, 0, bytes));
        }
    }
    s->h_prims = ps.prims;
    s->face_offset = ps.face_offset;
    s->nt = ps.nt;
    s->num_prims = (int)ps.prims.size();
    s->num_lights = (int)ps.lights.size();
    s->env_light = ps.env_light;
    s->cam = ps.cam;
    s->has_camera = ps.has_camera;
    s->composite = false;
    for (const MaterialDev& m : ps.mats) if (!mat_is_plain(m)) s->composite = true;
    CU(s->mats.alloc(ps.mats.size()));
    if (!ps.mats.empty()) CU(cudaMemcpy(s->mats.p, ps.mats.data(), ps.mats.size() * sizeof(MaterialDev), cudaMemcpyHostToDevice));
    CU(s->lights.alloc(ps.lights.size()));
    if (!ps.lights.empty()) CU(cudaMemcpy(s->lights.p, ps.lights.data(), ps.lights.size() * sizeof(LightDev), cudaMemcpyHostToDevice));
    CU(s->emit.alloc(ps.emit.size()));
    if (!ps.emit.empty()) CU(cudaMemcpy(s->emit.p, ps.emit.data(), ps.emit.size() * sizeof(EmitTri), cudaMemcpyHostToDevice));
    CU(s->cdf.alloc(ps.cdf.size()));
    if (!ps.cdf.empty()) CU(cudaMemcpy(s->cdf.p, ps.cdf.data(), ps.cdf.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(s->env_tex.alloc(ps.env_tex.size()));
    if (!ps.env_tex.empty()) CU(cudaMemcpy(s->env_tex.p, ps.env_tex.data(), ps.env_tex.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(s->mesh_t.alloc(ps.mesh_t.size()));
    if (!ps.mesh_t.empty()) CU(cudaMemcpy(s->mesh_t.p, ps.mesh_t.data(), ps.mesh_t.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(s->textures.alloc(ps.textures.size()));
    if (!ps.textures.empty()) CU(cudaMemcpy(s->textures.p, ps.textures.data(), ps.textures.size() * sizeof(TextureDev), cudaMemcpyHostToDevice));
    CU(s->texels.alloc(ps.texels.size()));
    if (!ps.texels.empty()) CU(cudaMemcpy(s->texels.p, ps.texels.data(), ps.texels.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(s->mix.alloc(ps.mix.size()));
    if (!ps.mix.empty()) CU(cudaMemcpy(s->mix.p, ps.mix.data(), ps.mix.size() * sizeof(MixEntry), cudaMemcpyHostToDevice));
    CU(s->prims.alloc(ps.prims.size()));
    CU(cudaMemcpy(s->prims.p, ps.prims.data(), ps.prims.size() * sizeof(PrimDev), cudaMemcpyHostToDevice));
    // a cudaMemcpy from pageable memory may return once the data is staged, and the memsets above are asynchronous: both sit on
    // the legacy default stream, which the scene's non-blocking streams do not wait for
    CU(cudaStreamSynchronize(0));
    s->built = false;
    return LMCU_OK;
}

// -- build ------------------------------------------------------------------------------------------
namespace {
struct SyncOnExit { cudaStream_t st; ~SyncOnExit() { cudaStreamSynchronize(st); cudaGetLastError(); } };
}

// morton -> radix sort -> hierarchy -> fit -> collapse over n items with float boxes lo / hi and centres, all on the device;
// bounds6 = the items' common box.  Out: `wide` (n entries allocated, *num_nodes used, root = entry 0) and `order`
// (leaf-order position -> item).  Shared by the flattened tree, the per-mesh trees and the tree over the instances.
static int build_tree(cudaStream_t st, uint32_t nt, const float* tri_lo, const float* tri_hi, const double* centre, const float* bounds6,
                      int max_leaf, double big_frac, DBuf<WideNode>& wide, DBuf<uint32_t>& tri_order, uint32_t* num_nodes) {
    DBuf<uint64_t> keys, keys_sorted;
    DBuf<uint32_t> vals, vals_sorted, parent, arrive, counters;
    DBuf<BinNode> bin;
    DBuf<CollapseItem> queue[2];
    // declared after the temporaries => destroyed before them: on an early error return kernels may still be running on
    // the stream, and a block parked in the cache while a kernel writes it could be handed to the next allocation
    SyncOnExit sync_on_exit{ st };
    CU(keys.alloc(nt)); CU(keys_sorted.alloc(nt)); CU(vals.alloc(nt)); CU(vals_sorted.alloc(nt));
    k_morton<<<grid_for(nt, 256), 256, 0, st>>>(centre, nt, bounds6, keys.p, vals.p, tri_lo, tri_hi, big_frac);
    CU(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)nt, 0, 63, st));
        DBuf<unsigned char> tmp;
        CU(tmp.alloc(tmp_bytes));
        CU(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)nt, 0, 63, st));
        CU(cudaStreamSynchronize(st));
    }
    CU(bin.alloc((size_t)2 * nt)); CU(parent.alloc((size_t)2 * nt)); CU(arrive.alloc(nt));
    CU(cudaMemsetAsync(arrive.p, 0, (size_t)nt * sizeof(uint32_t), st));
    CU(cudaMemsetAsync(parent.p, 0xff, (size_t)2 * nt * sizeof(uint32_t), st));
    if (nt > 1) {
        k_hierarchy<<<grid_for(nt - 1, 256), 256, 0, st>>>(keys_sorted.p, (int)nt, bin.p, parent.p);
        CU(cudaGetLastError());
    }
    k_fit<<<grid_for(nt, 256), 256, 0, st>>>(bin.p, parent.p, vals_sorted.p, tri_lo, tri_hi, (int)nt, arrive.p);
    CU(cudaGetLastError());
    // collapse, level by level
    CU(wide.alloc(nt)); CU(tri_order.alloc(nt)); CU(counters.alloc(4));
    CU(queue[0].alloc(nt)); CU(queue[1].alloc(nt));
    const uint32_t cinit[4] = { 1u, 0u, 0u, 0u };   // node_counter (root = 0 taken), tri_counter, next_count, -
    CU(cudaMemcpyAsync(counters.p, cinit, sizeof(cinit), cudaMemcpyHostToDevice, st));
    const CollapseItem root_item = { 0u, 0u };      // binary root is internal node 0 (or leaf 0 when nt == 1)
    CU(cudaMemcpyAsync(queue[0].p, &root_item, sizeof(root_item), cudaMemcpyHostToDevice, st));
    uint32_t count = 1;
    int cur = 0;
    while (count > 0) {
        CollapseCtx c;
        c.bin = bin.p; c.sorted_tri = vals_sorted.p; c.n = (int)nt; c.max_leaf = max_leaf; c.wide = wide.p;
        c.width = 8;
        c.tri_order = tri_order.p; c.node_counter = counters.p; c.tri_counter = counters.p + 1;
        c.next_queue = queue[cur ^ 1].p; c.next_count = counters.p + 2;
        k_collapse<<<grid_for(count, 128), 128, 0, st>>>(c, queue[cur].p, count);
        CU(cudaGetLastError());
        uint32_t h[3];
        CU(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        count = h[2];
        const uint32_t zero = 0;
        CU(cudaMemcpyAsync(counters.p + 2, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
        cur ^= 1;
        if (count == 0) {
            if (h[1] != nt) return fail(LMCU_ERR_CUDA, "bvh collapse lost triangles");
            *num_nodes = h[0];
        }
    }
    return LMCU_OK;
}

// The two-level layout (LmcuBuildOptions::two_level): node array = [tree over the primitives | one tree per mesh in use],
// object-space triangles per mesh in leaf order, instance records in the top level's leaf order.
static int build_two_level(LmcuScene* s, cudaStream_t st, int max_leaf, double big_frac) {
    const uint32_t np = (uint32_t)s->num_prims, nt = s->nt;
    const float binit[6] = { 3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f };
    DBuf<float> ilo, ihi, bounds;
    DBuf<double> icentre;
    DBuf<WideNode> top;
    DBuf<uint32_t> iorder;
    struct MeshTree { DBuf<WideNode> wide; DBuf<uint32_t> order; uint32_t num_nodes = 0, nf = 0, node_off = 0, item_off = 0; bool used = false; };
    std::vector<MeshTree> trees(s->face_offset.empty() ? 0 : s->face_offset.size() - 1);
    SyncOnExit sync_on_exit{ st };
    CU(ilo.alloc((size_t)np * 3)); CU(ihi.alloc((size_t)np * 3)); CU(icentre.alloc((size_t)np * 3)); CU(bounds.alloc(6));
    k_fill<<<grid_for((size_t)np * 3, 256), 256, 0, st>>>(ilo.p, (size_t)np * 3, 3e38f);
    k_fill<<<grid_for((size_t)np * 3, 256), 256, 0, st>>>(ihi.p, (size_t)np * 3, -3e38f);
    CU(cudaMemcpyAsync(bounds.p, binit, sizeof(binit), cudaMemcpyHostToDevice, st));
    k_inst_bounds<<<grid_for(nt, 256), 256, 0, st>>>(s->prims.p, s->num_prims, s->mesh_p.p, nt, ilo.p, ihi.p);
    k_inst_centres<<<grid_for(np, 256), 256, 0, st>>>(ilo.p, ihi.p, np, icentre.p, bounds.p);
    CU(cudaGetLastError());
    uint32_t top_nodes = 0;
    int rc = build_tree(st, np, ilo.p, ihi.p, icentre.p, bounds.p, max_leaf, big_frac, top, iorder, &top_nodes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(s->bounds, bounds.p, sizeof(s->bounds), cudaMemcpyDeviceToHost, st));
    std::vector<uint32_t> h_order(np);
    CU(cudaMemcpyAsync(h_order.data(), iorder.p, (size_t)np * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // bottom levels: one tree per mesh some primitive refers to
    uint32_t total_nodes = top_nodes, total_items = 0;
    for (const PrimDev& pr : s->h_prims) {
        MeshTree& t = trees[pr.mesh];
        if (t.used) continue;
        t.used = true; t.nf = pr.num_faces;
        const double* p = s->mesh_p.p + (size_t)pr.mesh_face_offset * 9;
        DBuf<float> lo, hi, mb;
        DBuf<double> centre;
        CU(lo.alloc((size_t)t.nf * 3)); CU(hi.alloc((size_t)t.nf * 3)); CU(centre.alloc((size_t)t.nf * 3)); CU(mb.alloc(6));
        CU(cudaMemcpyAsync(mb.p, binit, sizeof(binit), cudaMemcpyHostToDevice, st));
        k_obj_bounds<<<grid_for(t.nf, 256), 256, 0, st>>>(p, t.nf, lo.p, hi.p, centre.p, mb.p);
        CU(cudaGetLastError());
        rc = build_tree(st, t.nf, lo.p, hi.p, centre.p, mb.p, max_leaf, big_frac, t.wide, t.order, &t.num_nodes);
        if (rc) return rc;
        t.node_off = total_nodes; t.item_off = total_items;
        total_nodes += t.num_nodes; total_items += t.nf;
    }
    CU(s->nodes.alloc(total_nodes));
    CU(s->tobj.alloc(total_items));
    CU(cudaMemcpyAsync(s->nodes.p, top.p, (size_t)top_nodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
    for (const PrimDev& pr : s->h_prims) {
        MeshTree& t = trees[pr.mesh];
        if (!t.used) continue;
        t.used = false;                                   // placed once
        k_relocate<<<grid_for(t.num_nodes, 256), 256, 0, st>>>(t.wide.p, t.num_nodes, t.node_off, t.item_off);
        CU(cudaMemcpyAsync(s->nodes.p + t.node_off, t.wide.p, (size_t)t.num_nodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
        k_emit_obj<<<grid_for(t.nf, 256), 256, 0, st>>>(t.order.p, s->mesh_p.p + (size_t)pr.mesh_face_offset * 9, t.nf, pr.mesh, s->tobj.p + t.item_off);
        CU(cudaGetLastError());
    }
    std::vector<InstDev> h_inst(np);
    for (uint32_t k = 0; k < np; k++) {
        const uint32_t pi = h_order[k];
        if (pi >= np || !make_inst(s->h_prims[pi], pi, trees[s->h_prims[pi].mesh].node_off, h_inst[k]))
            return fail(LMCU_ERR_CUDA, "two-level build: bad instance record");
    }
    CU(s->inst.alloc(np));
    CU(cudaMemcpyAsync(s->inst.p, h_inst.data(), (size_t)np * sizeof(InstDev), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return LMCU_OK;
}

int lmcu_accel_build(LmcuScene* s, const LmcuBuildOptions* opts, LmcuBuildStats* stats) {
    if (!s) return fail(LMCU_ERR_INVALID_ARGUMENT, "null scene");
    if (s->nt == 0) return fail(LMCU_ERR_UNSUPPORTED, "Missing primitives.");
    CU(cudaSetDevice(s->device));
    const uint32_t nt = s->nt;
    int max_leaf = opts && opts->max_leaf_tris ? opts->max_leaf_tris : 2;
    if (const char* e = getenv("LMCU_LEAF")) max_leaf = atoi(e);      // tuning only
    if (max_leaf < 1 || max_leaf > 3) return fail(LMCU_ERR_INVALID_ARGUMENT, "max_leaf_tris must be in 1..3");
    int two_level_opt = opts ? opts->two_level : 0;
    if (const char* e = getenv("LMCU_TWO_LEVEL")) two_level_opt = atoi(e);
    double big_frac = 1.0 / 16.0;                                    // LMCU_BIG: denominator (0 = off), tuning only
    if (const char* e = getenv("LMCU_BIG")) big_frac = atof(e) > 0 ? 1.0 / atof(e) : 0.0;
    cudaStream_t st = s->stream;
    CU(cudaEventRecord(s->ev0, st));
    s->built = false;
    {
        // the device is only asked for its free memory when the answer can matter (cudaMemGetInfo costs ~10 ms next to
        // gigabytes of live allocations, and now and then hundreds: tools/build_probe.py)
        const int cand = two_level_candidate(two_level_opt, s->h_prims, s->face_offset, nt);
        s->two_level = cand > 0;
        if (cand == 0) {
            size_t free_b = 0, total_b = 0;
            if (const char* e = getenv("LMCU_BUILD_BUDGET_GB")) free_b = (size_t)(atof(e) * double(size_t(1) << 30));   // the memory the automatic layout choice plans with
            else CU(cudaMemGetInfo(&free_b, &total_b));
            s->two_level = two_level_by_budget(nt, (double)free_b);
        }
    }
    s->t32.release(); s->t64.release(); s->tobj.release(); s->inst.release(); s->shade.release(); s->uv.release();

    if (s->two_level) {
        const int rc = build_two_level(s, st, max_leaf, big_frac);
        if (rc) return rc;
        s->fp32 = false;
    } else {
        DBuf<double> wp, centre;
        DBuf<float> tri_lo, tri_hi, bounds;
        DBuf<int> exact;
        DBuf<uint32_t> tri_order;
        DBuf<WideNode> wide;
        SyncOnExit sync_on_exit{ st };
        CU(wp.alloc((size_t)nt * 9)); CU(centre.alloc((size_t)nt * 3));
        CU(tri_lo.alloc((size_t)nt * 3)); CU(tri_hi.alloc((size_t)nt * 3));
        CU(bounds.alloc(6)); CU(exact.alloc(1));
        const float binit[6] = { 3e38f, 3e38f, 3e38f, -3e38f, -3e38f, -3e38f };
        const int one = 1;
        CU(cudaMemcpyAsync(bounds.p, binit, sizeof(binit), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(exact.p, &one, sizeof(int), cudaMemcpyHostToDevice, st));
        k_flatten<<<grid_for(nt, 256), 256, 0, st>>>(s->prims.p, s->num_prims, s->mesh_p.p, nt, wp.p, tri_lo.p, tri_hi.p,
                                                     centre.p, bounds.p, exact.p);
        CU(cudaGetLastError());
        uint32_t num_nodes = 0;
        const int rc = build_tree(st, nt, tri_lo.p, tri_hi.p, centre.p, bounds.p, max_leaf, big_frac, wide, tri_order, &num_nodes);
        if (rc) return rc;
        // shrink the node array to its final size
        CU(s->nodes.alloc(num_nodes));
        CU(cudaMemcpyAsync(s->nodes.p, wide.p, (size_t)num_nodes * sizeof(WideNode), cudaMemcpyDeviceToDevice, st));
        int h_exact = 0;
        CU(cudaMemcpyAsync(s->bounds, bounds.p, sizeof(s->bounds), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&h_exact, exact.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // 48-byte float records when every vertex is float-representable AND the triangle array is too large to stay in
        // cache (C3: 480 MB instead of 800 MB streamed from HBM).  A small scene keeps the 80-byte FP64 records: they are
        // served by L1 / L2 either way and spare the 9 F2F.F64.F32 per triangle test, the largest stall site of the extend
        // kernel (profiles/r01_ncu_summary_v6.md; C2 pt 1322 -> 1340 Mpaths/s).
        s->fp32 = h_exact != 0 && (size_t)nt * sizeof(TriF64) > (size_t(48) << 20);
        if (const char* e = getenv("LMCU_TRI64")) s->fp32 = h_exact != 0 && atoi(e) == 0;   // tuning only: 1 = FP64 records, 0 = float records
        if (s->fp32) CU(s->t32.alloc(nt)); else CU(s->t64.alloc(nt));
        CU(s->shade.alloc(nt));
        CU(s->uv.alloc(s->mesh_t.p ? nt : 0));
        k_emit<<<grid_for(nt, 256), 256, 0, st>>>(tri_order.p, nt, s->prims.p, s->num_prims, wp.p, s->mesh_p.p, s->mesh_n.p,
                                                  s->t32.p, s->t64.p, s->shade.p, s->mesh_t.p, s->uv.p);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(s->ev1, st));
    double h_cost[3] = { 0, 0, 0 };
    if (stats && !s->two_level) {
        DBuf<double> cost;
        CU(cost.alloc(3));
        CU(cudaMemsetAsync(cost.p, 0, 3 * sizeof(double), st));
        k_sah_cost<<<grid_for(s->nodes.n, 256), 256, 0, st>>>(s->nodes.p, (uint32_t)s->nodes.n, cost.p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(h_cost, cost.p, sizeof(h_cost), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));      // before `cost` goes back to the block cache
    }
    CU(cudaStreamSynchronize(st));
    s->built = true;
    if (stats) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        stats->elapsed = ms * 1e-3;
        stats->num_triangles = nt;
        stats->num_wide_nodes = (int64_t)s->nodes.n;
        stats->node_bytes = (int64_t)(s->nodes.n * sizeof(WideNode));
        stats->triangle_bytes = s->two_level ? (int64_t)(s->tobj.n * sizeof(TriObj) + s->inst.n * sizeof(InstDev))
                                             : (int64_t)nt * (int64_t)(s->fp32 ? sizeof(TriF32) : sizeof(TriF64));
        stats->triangles_fp32 = s->fp32 ? 1 : 0;
        stats->two_level = s->two_level ? 1 : 0;
        // expected wide-node visits + triangle tests of a random line through the scene box (one-level trees only)
        stats->sah_cost = h_cost[2] > 0 ? (h_cost[0] + h_cost[1]) / h_cost[2] : 0.0;
    }
    return LMCU_OK;
}

// -- save / load ----------------------------------------------------------------------------------------
namespace {
struct AccelBlobHeader {
    char magic[8];                // "LMCUBVH1"
    uint32_t version, fp32, nt, num_nodes, num_prims, has_uv;
    uint64_t payload_bytes, checksum;
    uint64_t desc_digest;         // digest of the description the BVH was built from (mesh data + primitive table)
    float bounds[6];              // scene AABB
    uint32_t pad_[2];
};
constexpr uint32_t kBlobVersion = 3;
inline uint64_t fnv1a(const unsigned char* p, size_t n, uint64_t h = 1469598103934665603ull) {
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}
// payload checksum: 8 bytes per step (a byte-wise FNV over the 4 GB blob of a 50 M-triangle scene takes seconds)
inline uint64_t checksum64(const unsigned char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 1099511628211ull; h ^= h >> 29; }
    return fnv1a(p + i, n - i, h);
}
inline size_t blob_payload(const AccelBlobHeader& h) {
    return (size_t)h.num_nodes * sizeof(WideNode) + (size_t)h.nt * (h.fp32 ? sizeof(TriF32) : sizeof(TriF64)) +
           (size_t)h.nt * sizeof(TriShade) + (h.has_uv ? (size_t)h.nt * sizeof(TriUV) : 0) + (size_t)h.num_prims * sizeof(PrimDev);
}
}  // namespace

// Digest of the uploaded description: mesh positions and normals (on the device since lmcu_scene_upload) + the primitive
// table (transforms, mesh / material / light indices).  Two scenes with equal counts but other geometry or transforms differ.
static int desc_digest(LmcuScene* s, uint64_t* out) {
    if (!s->mesh_p.p) { *out = s->blob_desc_digest; return LMCU_OK; }
    CU(cudaMemsetAsync(s->cursor.p + 4, 0, 2 * sizeof(unsigned long long), s->stream));
    k_digest<<<s->num_sms * 8, 256, 0, s->stream>>>((const unsigned long long*)s->mesh_p.p, s->mesh_p.n, s->cursor.p + 4);
    k_digest<<<s->num_sms * 8, 256, 0, s->stream>>>((const unsigned long long*)s->mesh_n.p, s->mesh_n.n, s->cursor.p + 5);
    CU(cudaGetLastError());
    unsigned long long d[2];
    CU(cudaMemcpyAsync(d, s->cursor.p + 4, sizeof(d), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    uint64_t h = fnv1a((const unsigned char*)s->h_prims.data(), s->h_prims.size() * sizeof(PrimDev));
    h = (h ^ d[0]) * 1099511628211ull;
    h = (h ^ d[1]) * 1099511628211ull;
    *out = h ? h : 1;
    return LMCU_OK;
}

int lmcu_accel_save_size(LmcuScene* s, int64_t* bytes) {
    if (!s || !bytes) return fail(LMCU_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->built) return fail(LMCU_ERR_INVALID_ARGUMENT, "accel not built");
    if (s->two_level) return fail(LMCU_ERR_UNSUPPORTED, "a two-level accel is rebuilt from the scene, not serialised (LmcuBuildOptions::two_level = -1 builds the flattened tree)");
    AccelBlobHeader h{};
    h.fp32 = s->fp32; h.nt = s->nt; h.num_nodes = (uint32_t)s->nodes.n; h.num_prims = (uint32_t)s->num_prims; h.has_uv = s->uv.p ? 1 : 0;
    *bytes = (int64_t)(sizeof(AccelBlobHeader) + blob_payload(h));
    return LMCU_OK;
}

int lmcu_accel_save(LmcuScene* s, void* buffer, int64_t bytes) {
    int64_t need = 0;
    const int rc = lmcu_accel_save_size(s, &need);
    if (rc) return rc;
    if (!buffer || bytes < need) return fail(LMCU_ERR_INVALID_ARGUMENT, "buffer too small for the accel blob");
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->stream));
    AccelBlobHeader h{};
    memcpy(h.magic, "LMCUBVH1", 8);
    h.version = kBlobVersion; h.fp32 = s->fp32; h.nt = s->nt; h.num_nodes = (uint32_t)s->nodes.n;
    h.num_prims = (uint32_t)s->num_prims; h.has_uv = s->uv.p ? 1 : 0;
    h.payload_bytes = blob_payload(h);
    memcpy(h.bounds, s->bounds, sizeof(h.bounds));
    const int rd = desc_digest(s, &h.desc_digest);
    if (rd) return rd;
    unsigned char* base = static_cast<unsigned char*>(buffer) + sizeof(AccelBlobHeader);
    unsigned char* p = base;
    const auto pull = [&](const void* d, size_t n) -> cudaError_t {
        const cudaError_t e = n ? cudaMemcpy(p, d, n, cudaMemcpyDeviceToHost) : cudaSuccess;
        p += n;
        return e;
    };
    CU(pull(s->nodes.p, (size_t)h.num_nodes * sizeof(WideNode)));
    if (h.fp32) CU(pull(s->t32.p, (size_t)h.nt * sizeof(TriF32))); else CU(pull(s->t64.p, (size_t)h.nt * sizeof(TriF64)));
    CU(pull(s->shade.p, (size_t)h.nt * sizeof(TriShade)));
    if (h.has_uv) CU(pull(s->uv.p, (size_t)h.nt * sizeof(TriUV)));
    memcpy(p, s->h_prims.data(), (size_t)h.num_prims * sizeof(PrimDev));
    h.checksum = checksum64(base, h.payload_bytes);
    memcpy(buffer, &h, sizeof(h));
    return LMCU_OK;
}

// Every index a traversal or shade kernel will follow is range-checked here, on the host, before the blob reaches the
// device: a stale or foreign blob must be refused, not dereferenced.
static const char* validate_blob(const AccelBlobHeader& h, const unsigned char* base, int num_materials, int num_lights, bool described) {
    const WideNode* nodes = reinterpret_cast<const WideNode*>(base);
    const unsigned char* tris = base + (size_t)h.num_nodes * sizeof(WideNode);
    const size_t tri_bytes = (size_t)h.nt * (h.fp32 ? sizeof(TriF32) : sizeof(TriF64));
    const TriShade* shade = reinterpret_cast<const TriShade*>(tris + tri_bytes);
    const unsigned char* after_shade = reinterpret_cast<const unsigned char*>(shade + h.nt) + (h.has_uv ? (size_t)h.nt * sizeof(TriUV) : 0);
    std::vector<PrimDev> prims(h.num_prims);
    memcpy(prims.data(), after_shade, (size_t)h.num_prims * sizeof(PrimDev));
    for (uint32_t i = 0; i < h.num_nodes; i++) {
        WideNode n;
        memcpy(&n, nodes + i, sizeof(n));
        uint32_t n_internal = 0;
        for (int slot = 0; slot < 8; slot++) {
            const uint32_t m = n.meta[slot];
            const bool internal = (n.imask >> slot) & 1u;
            if (internal) {
                if (m != ((1u << 5) | (24u + (uint32_t)slot))) return "accel blob: an internal child slot has a foreign meta byte";
                n_internal++;
            } else if (m != 0) {
                const uint32_t unary = m >> 5, first = m & 0x1fu;
                const uint32_t cnt = unary == 1 ? 1u : unary == 3 ? 2u : unary == 7 ? 3u : 0u;
                if (cnt == 0 || first + cnt > 24u) return "accel blob: a leaf slot has a foreign meta byte";
                if ((uint64_t)n.tri_base + first + cnt > h.nt) return "accel blob: a leaf references triangles past the end";
            }
        }
        if (n_internal && ((uint64_t)n.child_base + n_internal > h.num_nodes || n.child_base <= i)) {
            return "accel blob: a node references children outside the node array";   // children always follow their parent
        }
    }
    for (uint32_t k = 0; k < h.nt; k++) {
        int32_t prim, face;
        if (h.fp32) {
            TriF32 t; memcpy(&t, tris + (size_t)k * sizeof(TriF32), sizeof(t));
            memcpy(&prim, &t.a.w, 4); memcpy(&face, &t.b.w, 4);
        } else {
            TriF64 t; memcpy(&t, tris + (size_t)k * sizeof(TriF64), sizeof(t));
            prim = t.prim; face = t.face;
        }
        if (prim < 0 || (uint32_t)prim >= h.num_prims) return "accel blob: a triangle references a primitive outside the table";
        if (face < 0 || (uint32_t)face >= prims[prim].num_faces) return "accel blob: a triangle references a face outside its mesh";
        if (described) {
            TriShade sh; memcpy(&sh, shade + k, sizeof(sh));
            int32_t mat, light;
            memcpy(&mat, &sh.n2.w, 4); memcpy(&light, &sh.n3.w, 4);
            if (mat < 0 || mat >= num_materials) return "accel blob: a triangle references a material outside the uploaded table";
            if (light < -1 || light >= num_lights) return "accel blob: a triangle references a light outside the uploaded table";
        }
    }
    return nullptr;
}

int lmcu_accel_load(LmcuScene* s, const void* buffer, int64_t bytes) {
    if (!s || !buffer) return fail(LMCU_ERR_INVALID_ARGUMENT, "null argument");
    if (bytes < (int64_t)sizeof(AccelBlobHeader)) return fail(LMCU_ERR_INVALID_ARGUMENT, "accel blob: truncated header");
    AccelBlobHeader h;
    memcpy(&h, buffer, sizeof(h));
    if (memcmp(h.magic, "LMCUBVH1", 8) != 0 || h.version != kBlobVersion) return fail(LMCU_ERR_INVALID_ARGUMENT, "accel blob: bad magic or version");
    if (h.nt == 0 || h.num_nodes == 0 || h.num_prims == 0 || h.payload_bytes != blob_payload(h) || bytes < (int64_t)(sizeof(h) + h.payload_bytes)) {
        return fail(LMCU_ERR_INVALID_ARGUMENT, "accel blob: inconsistent sizes");
    }
    const unsigned char* base = static_cast<const unsigned char*>(buffer) + sizeof(h);
    if (checksum64(base, h.payload_bytes) != h.checksum) return fail(LMCU_ERR_INVALID_ARGUMENT, "accel blob: checksum mismatch");
    const bool described = s->nt != 0;
    CU(cudaSetDevice(s->device));
    if (described) {
        if (s->nt != h.nt || (uint32_t)s->num_prims != h.num_prims || (s->mesh_t.p != nullptr) != (h.has_uv != 0)) {
            return fail(LMCU_ERR_INVALID_ARGUMENT, "accel blob does not belong to the uploaded scene");
        }
        // same counts are not enough: world-space triangles come from the blob, materials and primitives from the upload
        uint64_t mine = 0;
        const int rd = desc_digest(s, &mine);
        if (rd) return rd;
        const unsigned char* blob_prims = base + h.payload_bytes - (size_t)h.num_prims * sizeof(PrimDev);
        if (mine != h.desc_digest || memcmp(blob_prims, s->h_prims.data(), (size_t)h.num_prims * sizeof(PrimDev)) != 0) {
            return fail(LMCU_ERR_INVALID_ARGUMENT, "accel blob was built from other geometry or transforms than the uploaded scene");
        }
    }
    if (const char* why = validate_blob(h, base, (int)s->mats.n, s->num_lights, described)) return fail(LMCU_ERR_INVALID_ARGUMENT, why);
    CU(cudaStreamSynchronize(s->stream));
    const unsigned char* p = base;
    const auto push = [&](void* d, size_t n) -> cudaError_t {
        const cudaError_t e = n ? cudaMemcpy(d, p, n, cudaMemcpyHostToDevice) : cudaSuccess;
        p += n;
        return e;
    };
    s->built = false;
    memcpy(s->bounds, h.bounds, sizeof(s->bounds));
    CU(s->nodes.alloc(h.num_nodes));
    CU(push(s->nodes.p, (size_t)h.num_nodes * sizeof(WideNode)));
    s->fp32 = h.fp32 != 0;
    s->two_level = false; s->tobj.release(); s->inst.release();
    s->t32.release(); s->t64.release();
    if (s->fp32) { CU(s->t32.alloc(h.nt)); CU(push(s->t32.p, (size_t)h.nt * sizeof(TriF32))); }
    else { CU(s->t64.alloc(h.nt)); CU(push(s->t64.p, (size_t)h.nt * sizeof(TriF64))); }
    CU(s->shade.alloc(h.nt));
    CU(push(s->shade.p, (size_t)h.nt * sizeof(TriShade)));
    CU(s->uv.alloc(h.has_uv ? h.nt : 0));
    if (h.has_uv) CU(push(s->uv.p, (size_t)h.nt * sizeof(TriUV)));
    if (!described) {
        s->h_prims.resize(h.num_prims);
        memcpy(s->h_prims.data(), p, (size_t)h.num_prims * sizeof(PrimDev));
        CU(s->prims.alloc(h.num_prims));
        CU(cudaMemcpy(s->prims.p, s->h_prims.data(), (size_t)h.num_prims * sizeof(PrimDev), cudaMemcpyHostToDevice));
        s->nt = h.nt;
        s->num_prims = (int)h.num_prims;
        s->blob_desc_digest = h.desc_digest;
    }
    CU(cudaStreamSynchronize(0));           // the pageable copies above (legacy default stream) have landed
    s->built = true;
    return LMCU_OK;
}

// -- intersect ------------------------------------------------------------------------------------------
// Picks the kernel instantiation for (triangle record: LmcuScene::tri_kind(), counters).
#define LMCU_DISPATCH(KERNEL, KIND, COUNT, GRID, BLOCK, STREAM, ...)                                             \
    do {                                                                                                         \
        if ((KIND) == 2) {                                                                                       \
            if (COUNT) KERNEL<TriObj, true><<<GRID, BLOCK, 0, STREAM>>>(__VA_ARGS__);                            \
            else KERNEL<TriObj, false><<<GRID, BLOCK, 0, STREAM>>>(__VA_ARGS__);                                 \
        } else if ((KIND) == 1) {                                                                                \
            if (COUNT) KERNEL<TriF32, true><<<GRID, BLOCK, 0, STREAM>>>(__VA_ARGS__);                            \
            else KERNEL<TriF32, false><<<GRID, BLOCK, 0, STREAM>>>(__VA_ARGS__);                                 \
        } else {                                                                                                 \
            if (COUNT) KERNEL<TriF64, true><<<GRID, BLOCK, 0, STREAM>>>(__VA_ARGS__);                            \
            else KERNEL<TriF64, false><<<GRID, BLOCK, 0, STREAM>>>(__VA_ARGS__);                                 \
        }                                                                                                        \
    } while (0)

#define LMCU_DISPATCH_SMEM(KERNEL, FP32, COUNT, GRID, BLOCK, SMEM, STREAM, ...)                                  \
    do {                                                                                                         \
        if (FP32) {                                                                                              \
            if (COUNT) KERNEL<TriF32, true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                         \
            else KERNEL<TriF32, false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                              \
        } else {                                                                                                 \
            if (COUNT) KERNEL<TriF64, true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                         \
            else KERNEL<TriF64, false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__);                              \
        }                                                                                                        \
    } while (0)

static int launch_intersect(LmcuScene* s, const LmcuRay* d_rays, int64_t n, LmcuHit* d_hits, bool count, cudaStream_t st) {
    const DevScene d = s->dev();
    if (count) CU(cudaMemsetAsync(s->cursor.p + 1, 0, 2 * sizeof(unsigned long long), st));
    const int64_t kChunk = 1ll << 30;            // the queue cursor is 32-bit
    for (int64_t off = 0; off < n; off += kChunk) {
        const uint32_t m = (uint32_t)std::min<int64_t>(kChunk, n - off);
        unsigned long long* cur = s->cursor.p + 8 + (s->cursor_next.fetch_add(1) % LmcuScene::kCursorRing);
        CU(cudaMemsetAsync(cur, 0, sizeof(unsigned long long), st));
        if (s->pool_on() && st == s->stream) {      // one arena per scene: only for launches ordered by the scene's stream
            const unsigned pblocks = (unsigned)std::min<int64_t>((m + 2 * LMCU_TRAV_BLOCK - 1) / (2 * LMCU_TRAV_BLOCK), (int64_t)s->num_sms * 5);
            const size_t need = (size_t)s->num_sms * 5 * (LMCU_TRAV_BLOCK / 32) * kStackSize * 32 * kPoolK;
            if (s->batch_arena.n != need) CU(s->batch_arena.alloc(need));
            const size_t smem = pool_bytes_per_warp<BatchIO, kPoolK>() * (LMCU_TRAV_BLOCK / 32);
            LMCU_DISPATCH_SMEM(k_intersect_p, s->fp32, count, pblocks, LMCU_TRAV_BLOCK, smem, st, d, d_rays + off, m, d_hits + off,
                               (uint32_t*)cur, s->cursor.p + 1, s->flags.p, s->batch_arena.p);
            CU(cudaGetLastError());
            continue;
        }
        const unsigned blocks = (unsigned)std::min<int64_t>((m + LMCU_TRAV_BLOCK - 1) / LMCU_TRAV_BLOCK, (int64_t)s->num_sms * s->trav_ctas());
        LMCU_DISPATCH(k_intersect_q, s->tri_kind(), count, blocks, LMCU_TRAV_BLOCK, st, d, d_rays + off, m, d_hits + off,
                      (uint32_t*)cur, s->cursor.p + 1, s->flags.p);
        CU(cudaGetLastError());
    }
    return LMCU_OK;
}

static int check_overflow(LmcuScene* s) {
    uint32_t f = 0;
    CU(cudaMemcpy(&f, s->flags.p, sizeof(f), cudaMemcpyDeviceToHost));
    if (f) {
        CU(cudaMemset(s->flags.p, 0, sizeof(uint32_t)));
        return fail(LMCU_ERR_CUDA, "bvh traversal stack overflow");
    }
    return LMCU_OK;
}

int lmcu_accel_intersect_device(LmcuScene* s, const void* d_rays, int64_t n, void* d_hits, void* stream) {
    if (!s || !s->built) return fail(LMCU_ERR_UNSUPPORTED, "Missing acceleration structure. Build first.");
    if (n <= 0) return LMCU_OK;
    CU(cudaSetDevice(s->device));
    return launch_intersect(s, (const LmcuRay*)d_rays, n, (LmcuHit*)d_hits, false, stream ? (cudaStream_t)stream : s->stream);
}

static int intersect_host(LmcuScene* s, const LmcuRay* rays, int64_t n, LmcuHit* hits, bool count, double* npr, double* tpr) {
    if (!s || !s->built) return fail(LMCU_ERR_UNSUPPORTED, "Missing acceleration structure. Build first.");
    if (n <= 0) return LMCU_OK;
    if (!rays || !hits) return fail(LMCU_ERR_INVALID_ARGUMENT, "null buffer");
    CU(cudaSetDevice(s->device));
    DBuf<LmcuRay> d_rays;
    DBuf<LmcuHit> d_hits;
    SyncOnExit sync_on_exit{ s->stream };   // destroyed before the buffers: an early error return never parks a block a kernel still writes
    CU(d_rays.alloc(n)); CU(d_hits.alloc(n));
    CU(cudaMemcpyAsync(d_rays.p, rays, (size_t)n * sizeof(LmcuRay), cudaMemcpyHostToDevice, s->stream));
    const int rc = launch_intersect(s, d_rays.p, n, d_hits.p, count, s->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(hits, d_hits.p, (size_t)n * sizeof(LmcuHit), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (count) {
        unsigned long long c[2];
        CU(cudaMemcpy(c, s->cursor.p + 1, sizeof(c), cudaMemcpyDeviceToHost));
        if (npr) *npr = (double)c[0] / (double)n;
        if (tpr) *tpr = (double)c[1] / (double)n;
    }
    return check_overflow(s);
}

int lmcu_accel_intersect(LmcuScene* s, const LmcuRay* rays, int64_t n, LmcuHit* hits) {
    return intersect_host(s, rays, n, hits, false, nullptr, nullptr);
}

int lmcu_accel_intersect_counted(LmcuScene* s, const LmcuRay* rays, int64_t n, LmcuHit* hits, double* npr, double* tpr) {
    return intersect_host(s, rays, n, hits, true, npr, tpr);
}

// -- scalar intersect -----------------------------------------------------------------------------------
// One pinned, device-mapped (ray, hit) slot and one stream per calling THREAD and device: callers such as scene::default
// under the reference's CPU renderers call Accel::intersect concurrently from every OpenMP thread (src/scene.cpp:320).
// No lock, no allocation, no memcpy per call: write the ray, launch, synchronise the thread's own stream, read the hit.
namespace {
struct ScalarSlot {
    int device = -1;
    cudaStream_t st = nullptr;
    LmcuRay* h_ray = nullptr; LmcuHit* h_hit = nullptr;
    LmcuRay* d_ray = nullptr; LmcuHit* d_hit = nullptr;
    volatile unsigned long long* h_done = nullptr; unsigned long long* d_done = nullptr;
    unsigned long long ticket = 0;
};
struct ScalarSlots {
    std::vector<ScalarSlot> v;
    ~ScalarSlots() {                       // thread exit; errors are ignored (the context may already be gone)
        for (auto& s : v) { if (s.h_ray) cudaFreeHost(s.h_ray); if (s.st) cudaStreamDestroy(s.st); }
        cudaGetLastError();
    }
};
thread_local ScalarSlots t_slots;
}  // namespace

int lmcu_accel_intersect1(LmcuScene* s, const LmcuRay* ray, LmcuHit* hit) {
    if (!s || !s->built) return fail(LMCU_ERR_UNSUPPORTED, "Missing acceleration structure. Build first.");
    if (!ray || !hit) return fail(LMCU_ERR_INVALID_ARGUMENT, "null argument");
    ScalarSlot* slot = nullptr;
    for (auto& c : t_slots.v) if (c.device == s->device) slot = &c;
    CU(cudaSetDevice(s->device));
    if (!slot) {
        ScalarSlot c;
        c.device = s->device;
        CU(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        void* h = nullptr;
        CU(cudaHostAlloc(&h, 128, cudaHostAllocMapped));
        memset(h, 0, 128);
        c.h_ray = (LmcuRay*)h; c.h_hit = (LmcuHit*)((char*)h + 64); c.h_done = (volatile unsigned long long*)((char*)h + 96);
        void* d = nullptr;
        CU(cudaHostGetDevicePointer(&d, h, 0));
        c.d_ray = (LmcuRay*)d; c.d_hit = (LmcuHit*)((char*)d + 64); c.d_done = (unsigned long long*)((char*)d + 96);
        t_slots.v.push_back(c);
        slot = &t_slots.v.back();
    }
    *slot->h_ray = *ray;
    const DevScene d = s->dev();
    const unsigned long long ticket = ++slot->ticket;
    if (s->two_level) k_intersect_one<TriObj><<<1, 32, 0, slot->st>>>(d, slot->d_ray, slot->d_hit, slot->d_done, ticket);
    else if (s->fp32) k_intersect_one<TriF32><<<1, 32, 0, slot->st>>>(d, slot->d_ray, slot->d_hit, slot->d_done, ticket);
    else k_intersect_one<TriF64><<<1, 32, 0, slot->st>>>(d, slot->d_ray, slot->d_hit, slot->d_done, ticket);
    CU(cudaGetLastError());
    // Spin on the ticket the kernel writes into the pinned slot (a stream synchronise costs several microseconds of driver
    // time per ray); every few thousand spins ask the stream whether the launch failed instead of completing.
    for (unsigned spins = 0; *slot->h_done != ticket; spins++) {
        if ((spins & 0xfffu) == 0xfffu) {
            const cudaError_t q = cudaStreamQuery(slot->st);
            if (q != cudaSuccess && q != cudaErrorNotReady) CU(q);
            if (q == cudaSuccess && *slot->h_done != ticket) return fail(LMCU_ERR_CUDA, "scalar intersect kernel finished without an answer");
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    *hit = *slot->h_hit;
    if (hit->primitive == -2) return fail(LMCU_ERR_CUDA, "bvh traversal stack overflow");
    return LMCU_OK;
}

int lmcu_cache_trim(int device) {
    if (device < 0 || device >= lmcu_device_count()) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad device index");
    CU(cudaSetDevice(device));
    CU(cudaDeviceSynchronize());
    g_cache.flush(device);
    return LMCU_OK;
}

int lmcu_cache_stats(int device, int64_t* parked_bytes, int64_t* parked_blocks) {
    if (device < 0 || device >= lmcu_device_count()) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad device index");
    std::lock_guard<std::mutex> lk(g_cache.mu);
    int64_t nb = 0;
    for (const auto& b : g_cache.blocks) if (b.first.first == device) nb++;
    if (parked_bytes) *parked_bytes = (int64_t)g_cache.bytes[device];
    if (parked_blocks) *parked_blocks = nb;
    return LMCU_OK;
}

int lmcu_probe_l2_bandwidth(int device, int64_t bytes, double* gb_per_s) {
    if (!gb_per_s || bytes < 4096) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad argument");
    if (lmcu_device_count() == 0) return fail(LMCU_ERR_NO_DEVICE, "no CUDA device: liblmcu has no CPU fallback");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DBuf<uint4> buf;
    DBuf<uint32_t> sink;
    size_t n16 = 256;
    while (n16 * 2 * 16 <= (size_t)bytes) n16 *= 2;          // largest power of two of 16-byte words within `bytes`
    CU(buf.alloc(n16)); CU(sink.alloc(1));
    CU(cudaMemset(buf.p, 1, n16 * 16));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const unsigned blocks = (unsigned)sms * 8;
    const size_t per_rep = ((n16 + (size_t)blocks * 256 - 1) / ((size_t)blocks * 256)) * (size_t)blocks * 256;   // loads issued per rep
    double best = 0.0;
    for (int it = 0; it < 4; it++) {
        const int reps = it == 0 ? 2 : 64;
        CU(cudaEventRecord(e0, 0));
        k_probe_l2<<<blocks, 256>>>(buf.p, n16, reps, sink.p);
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0) best = std::max(best, (double)per_rep * 16.0 * reps / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gb_per_s = best;
    return LMCU_OK;
}

// -- film ---------------------------------------------------------------------------------------------
static int ensure_film(LmcuScene* s, int w, int h, bool with_acc = false) {
    if (w <= 0 || h <= 0) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad film size");
    if (s->film_w != w || s->film_h != h || !s->film.p) {
        CU(s->film.alloc((size_t)w * h * 3));
        s->film_w = w; s->film_h = h;
    }
    CU(cudaMemsetAsync(s->film.p, 0, s->film.n * sizeof(float), s->stream));   // film_->clear()
    if (with_acc) {
        if (s->film_acc.n != s->film.n) CU(s->film_acc.alloc(s->film.n));
        CU(cudaMemsetAsync(s->film_acc.p, 0, s->film_acc.n * sizeof(double), s->stream));
    }
    return LMCU_OK;
}

int lmcu_film_device_ptr(LmcuScene* s, void** d_film, int64_t* num_floats) {
    if (!s || !s->film.p) return fail(LMCU_ERR_UNSUPPORTED, "no film rendered yet");
    if (d_film) *d_film = s->film.p;
    if (num_floats) *num_floats = (int64_t)s->film.n;
    return LMCU_OK;
}

int lmcu_film_rescale(LmcuScene* s, double v) {
    if (!s || !s->film.p) return fail(LMCU_ERR_UNSUPPORTED, "no film rendered yet");
    CU(cudaSetDevice(s->device));
    k_scale<<<s->num_sms * 4, 256, 0, s->stream>>>(s->film.p, s->film.n, (float)v);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    return LMCU_OK;
}

int lmcu_film_download(LmcuScene* s, float* film_rgb) {
    if (!s || !s->film.p) return fail(LMCU_ERR_UNSUPPORTED, "no film rendered yet");
    CU(cudaSetDevice(s->device));
    CU(cudaMemcpyAsync(film_rgb, s->film.p, s->film.n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return LMCU_OK;
}

int lmcu_film_ipc_export(LmcuScene* s, void* handle) {
    if (!s || !s->film.p) return fail(LMCU_ERR_UNSUPPORTED, "no film rendered yet");
    if (!handle) return fail(LMCU_ERR_INVALID_ARGUMENT, "null handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == LMCU_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->film.p));
    g_cache.mark_exported(s->film.p);
    memcpy(handle, &h, sizeof(h));
    return LMCU_OK;
}

int lmcu_ipc_open(int device, const void* handle, void** d_ptr) {
    if (!handle || !d_ptr) return fail(LMCU_ERR_INVALID_ARGUMENT, "null argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return LMCU_OK;
}

int lmcu_ipc_close(int device, void* d_ptr) {
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(d_ptr));
    return LMCU_OK;
}

int lmcu_enable_peer_access(int device, int peer_device) {
    if (device == peer_device) return LMCU_OK;
    CU(cudaSetDevice(device));
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) return fail(LMCU_ERR_UNSUPPORTED, "devices are not P2P-capable");
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return LMCU_OK; }
    CU(e);
    return LMCU_OK;
}

int lmcu_film_reduce_peers(LmcuScene* s, void* const* peer_films, int num_peers, double scale) {
    if (!s || !s->film.p) return fail(LMCU_ERR_UNSUPPORTED, "no film rendered yet");
    if (num_peers < 0 || num_peers > LMCU_MAX_PEERS) return fail(LMCU_ERR_INVALID_ARGUMENT, "too many peers");
    if (num_peers > 0 && !peer_films) return fail(LMCU_ERR_INVALID_ARGUMENT, "null peer list");
    CU(cudaSetDevice(s->device));
    PeerFilms pf{};
    for (int k = 0; k < num_peers; k++) {
        if (!peer_films[k]) return fail(LMCU_ERR_INVALID_ARGUMENT, "null peer film");
        pf.p[k] = (const float*)peer_films[k];
    }
    k_film_reduce_peers<<<s->num_sms * 8, 256, 0, s->stream>>>(s->film.p, pf, num_peers, s->film.n, (float)scale);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    return LMCU_OK;
}

int lmcu_film_reduce_slice(LmcuScene* s, void* const* films, int num_films, void* out, int64_t begin, int64_t end, double scale) {
    if (!s) return fail(LMCU_ERR_INVALID_ARGUMENT, "null scene");
    if (num_films < 1 || num_films > LMCU_MAX_PEERS + 1 || !films || !out) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad film list");
    if (begin < 0 || end < begin || (begin & 3)) return fail(LMCU_ERR_INVALID_ARGUMENT, "slice must start at a multiple of 4 floats");
    CU(cudaSetDevice(s->device));
    FilmSet fs{};
    for (int k = 0; k < num_films; k++) {
        if (!films[k]) return fail(LMCU_ERR_INVALID_ARGUMENT, "null film");
        fs.p[k] = (const float*)films[k];
    }
    if (end > begin) {
        const size_t quads = (size_t)(end - begin) / 4 + 1;
        const unsigned blocks = (unsigned)std::min<size_t>((quads + 255) / 256, (size_t)s->num_sms * 8);
        k_film_reduce_slice<<<blocks, 256, 0, s->stream>>>(fs, num_films, (float*)out, (size_t)begin, (size_t)end, (float)scale);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(s->stream));
    return LMCU_OK;
}

static int require_renderable(LmcuScene* s, bool need_light) {
    if (!s) return fail(LMCU_ERR_INVALID_ARGUMENT, "null scene");
    if (s->nt == 0) return fail(LMCU_ERR_UNSUPPORTED, "Missing primitives. Use lm::primitive() function to add primitives.");
    if (!s->has_camera) return fail(LMCU_ERR_UNSUPPORTED, "Missing camera primitive.");
    if (need_light && s->num_lights == 0) return fail(LMCU_ERR_UNSUPPORTED, "No light in the scene.");
    if (!s->built) return fail(LMCU_ERR_UNSUPPORTED, "Missing acceleration structure.");
    return LMCU_OK;
}

// -- raycast ------------------------------------------------------------------------------------------
int lmcu_render_raycast(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats) {
    int rc = require_renderable(s, false);
    if (rc) return rc;
    if (!p) return fail(LMCU_ERR_INVALID_ARGUMENT, "null params");
    CU(cudaSetDevice(s->device));
    rc = ensure_film(s, p->width, p->height);
    if (rc) return rc;
    RaycastParams rp{};
    rp.width = p->width; rp.height = p->height;
    rp.use_constant_color = p->use_constant_color; rp.visualize_normal = p->visualize_normal; rp.has_color = p->has_color;
    for (int k = 0; k < 3; k++) { rp.bg[k] = (float)p->bg_color[k]; rp.color[k] = (float)p->color[k]; }
    cudaStream_t st = s->stream;
    CU(cudaMemsetAsync(s->cursor.p, 0, sizeof(unsigned long long), st));
    CU(cudaEventRecord(s->ev0, st));
    const size_t n = (size_t)p->width * p->height;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 127) / 128, (size_t)s->num_sms * 12);
    const DevScene d = s->dev();
    if (s->two_level) k_raycast<TriObj><<<blocks, 128, 0, st>>>(d, rp, s->film.p, s->cursor.p);
    else if (s->fp32) k_raycast<TriF32><<<blocks, 128, 0, st>>>(d, rp, s->film.p, s->cursor.p);
    else k_raycast<TriF64><<<blocks, 128, 0, st>>>(d, rp, s->film.p, s->cursor.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(s->ev1, st));
    if (film_rgb) CU(cudaMemcpyAsync(film_rgb, s->film.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        memset(stats, 0, sizeof(*stats));
        stats->elapsed = ms * 1e-3; stats->processed = 1; stats->paths = (int64_t)n; stats->rays = (int64_t)n;
        stats->kernel_launches = 1;
    }
    return check_overflow(s);
}

// -- ao -------------------------------------------------------------------------------------------------
int lmcu_render_ao(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats) {
    int rc = require_renderable(s, false);
    if (rc) return rc;
    if (!p || p->spp <= 0) return fail(LMCU_ERR_INVALID_ARGUMENT, "spp must be positive");
    CU(cudaSetDevice(s->device));
    rc = ensure_film(s, p->width, p->height);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    CU(cudaMemsetAsync(s->cursor.p, 0, 4 * sizeof(unsigned long long), st));
    CU(cudaEventRecord(s->ev0, st));
    const size_t n = (size_t)p->width * p->height;
    if (p->ao_mark_misses) {                       // pixels the kernel never writes (primary ray missed) read back as -1
        k_fill<<<s->num_sms * 4, 256, 0, st>>>(s->film.p, s->film.n, -1.0f);
        CU(cudaGetLastError());
    }
    const unsigned blocks = (unsigned)std::min<size_t>((n + 127) / 128, (size_t)s->num_sms * 12);
    const DevScene d = s->dev();
    const double dist = p->ao_distance > 0 ? p->ao_distance : 0.2;
    if (s->two_level) k_ao<TriObj><<<blocks, 128, 0, st>>>(d, p->width, p->height, p->spp, (unsigned long long)p->sample_offset, p->seed, dist, s->film.p, s->cursor.p, s->cursor.p + 3);
    else if (s->fp32) k_ao<TriF32><<<blocks, 128, 0, st>>>(d, p->width, p->height, p->spp, (unsigned long long)p->sample_offset, p->seed, dist, s->film.p, s->cursor.p, s->cursor.p + 3);
    else k_ao<TriF64><<<blocks, 128, 0, st>>>(d, p->width, p->height, p->spp, (unsigned long long)p->sample_offset, p->seed, dist, s->film.p, s->cursor.p, s->cursor.p + 3);
    CU(cudaGetLastError());
    CU(cudaEventRecord(s->ev1, st));
    if (film_rgb) CU(cudaMemcpyAsync(film_rgb, s->film.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    unsigned long long rays = 0;
    CU(cudaMemcpyAsync(&rays, s->cursor.p + 3, sizeof(rays), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        memset(stats, 0, sizeof(*stats));
        stats->elapsed = ms * 1e-3; stats->processed = p->spp; stats->paths = (int64_t)n; stats->rays = (int64_t)rays;
        stats->kernel_launches = 1;
    }
    return check_overflow(s);
}

// -- pt ---------------------------------------------------------------------------------------------------
// shadow_factor: 1, or 2 in image mode where one iteration can hold a camera-vertex NEE ray per NEW path on top of one
// NEE ray per path in the queue
static int ensure_pt_pools(LmcuScene* s, uint32_t capacity, uint32_t shadow_factor) {
    if (s->pt_capacity == capacity && s->pt_waves == s->waves && s->wave[0].ctr.p &&
        s->wave[0].sq[0].n >= (size_t)capacity * shadow_factor) return LMCU_OK;
    for (int w = 0; w < LmcuScene::kMaxWaves; w++) {
        LmcuScene::PtWave& v = s->wave[w];
        const uint32_t cap = w < s->waves ? capacity : 0u;
        for (int b = 0; b < 2; b++) {
            for (int k = 0; k < 4; k++) CU(v.pq[b][k].alloc(cap));
            CU(v.pinfo[b].alloc(cap));
        }
        CU(v.hits.alloc(cap));
        for (int k = 0; k < 3; k++) CU(v.sq[k].alloc((size_t)cap * shadow_factor));
        CU(v.ctr.alloc(cap ? 1 : 0));
        CU(v.arena.alloc(cap && s->pool_mode ? (size_t)s->num_sms * s->ctas_per_sm * (LMCU_TRAV_BLOCK / 32) * kStackSize * 32 * kPoolK : 0));
        if (cap && w > 0 && !v.stream) CU(cudaStreamCreateWithFlags(&v.stream, cudaStreamNonBlocking));
        if (cap && !v.done) CU(cudaEventCreateWithFlags(&v.done, cudaEventDisableTiming));
        if (cap && !v.h) CU(cudaHostAlloc((void**)&v.h, 2 * sizeof(PtCounters), cudaHostAllocDefault));
    }
    if (!s->fork) CU(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
    s->pt_capacity = capacity;
    s->pt_waves = s->waves;
    return LMCU_OK;
}

// One wavefront pass: traces pp.total_paths paths (sample indices pp.sample_offset ...) to completion, accumulating into
// the film.  Adds its counters and per-kernel times to `acc`.
struct PtAccum {
    // progress hook (LmcuRenderParams::progress): paths started by the passes before this one, paths of the job, start time
    void (*progress)(void*, int64_t, int64_t, double) = nullptr;
    void* progress_user = nullptr;
    int64_t paths_before = 0, paths_total = 0;
    std::chrono::steady_clock::time_point t0;
    int64_t launches = 0, extend_launches = 0;
    double t_extend = 0, t_shade = 0, t_shadow = 0;
    unsigned long long rays = 0, extend_rays = 0, shadow_rays = 0, nodes_extend = 0, tris_extend = 0, nodes_shadow = 0, tris_shadow = 0;
};

static int pt_pass(LmcuScene* s, const PtParams& pp_all, bool count, PtAccum& acc) {
    // K staggered wavefronts: the paths of this pass are cut into K contiguous ranges, each traced to completion by its own
    // generate / extend / shade / shadow loop on its own stream over its own pools, all adding into the one film.  The
    // kernels of a loop depend on each other, so ONE wavefront leaves the machine to one kernel at a time: while it sits
    // in the shade kernel (bound by dependent gathers, issue slots half idle) nothing traverses.  With two loops out of
    // phase the traversal kernels of one fill the shade kernel's bubbles of the other.  Every kernel keeps its full
    // persistent grid (one wave of resident CTAs): the block scheduler hands a freed CTA slot to whichever loop has CTAs
    // pending, so a kernel's drain tail is filled by the other loop's next kernel (grids cut to 1/K measured worse).
    const int K = s->waves;
    const DevScene d = s->dev();
    const unsigned q_blocks = (unsigned)(s->num_sms * s->trav_ctas(true));
    const unsigned shade_blocks = (unsigned)(s->num_sms * s->shade_ctas_per_sm), gen_blocks = (unsigned)s->num_sms * 4;
    const int kGroup = 4;                       // iterations between host checks
    const size_t pool_smem_pt = pool_bytes_per_warp<ExtendIO, kPoolK>() * (LMCU_TRAV_BLOCK / 32);
    const double spp_per_iteration = (double)pp_all.capacity / ((double)pp_all.width * (double)pp_all.height);
    const unsigned fold_blocks = (unsigned)std::min<size_t>((s->film.n + 255) / 256, (size_t)s->num_sms * 8);
    PtParams pp[LmcuScene::kMaxWaves];
    PathQueue q[LmcuScene::kMaxWaves][2];
    ShadowQueue sq[LmcuScene::kMaxWaves];
    if (K > 1) CU(cudaEventRecord(s->fork, s->stream));
    for (int w = 0; w < K; w++) {
        LmcuScene::PtWave& v = s->wave[w];
        cudaStream_t st = w == 0 ? s->stream : v.stream;
        if (w > 0) CU(cudaStreamWaitEvent(st, s->fork, 0));           // the film clear / earlier passes happened on s->stream
        pp[w] = pp_all;
        const uint64_t lo = pp_all.total_paths * (uint64_t)w / (uint64_t)K, hi = pp_all.total_paths * (uint64_t)(w + 1) / (uint64_t)K;
        pp[w].path_base = pp_all.path_base + lo;
        pp[w].total_paths = hi - lo;
        for (int b = 0; b < 2; b++) {
            q[w][b].ray_o = v.pq[b][0].p; q[w][b].ray_d = v.pq[b][1].p; q[w][b].thr_pix = v.pq[b][2].p; q[w][b].aux = v.pq[b][3].p;
            q[w][b].info = v.pinfo[b].p;
        }
        sq[w].ray_o = v.sq[0].p; sq[w].ray_d = v.sq[1].p; sq[w].contrib = v.sq[2].p;
        CU(cudaMemsetAsync(v.ctr.p, 0, sizeof(PtCounters), st));
        v.cur = 0;
        v.finished = false;
        v.since_fold = 0.0;
        memset(v.h, 0, 2 * sizeof(PtCounters));
        for (int k = 0; k < 2; k++) {
            if (v.ring[k].empty()) {
                v.ring[k].resize(kGroup * 4);
                for (auto& e : v.ring[k]) CU(cudaEventCreate(&e));
                CU(cudaEventCreateWithFlags(&v.grp[k], cudaEventDisableTiming));
            }
        }
    }
    // One group = kGroup iterations + the counters' read-back + the slot's completion event, all asynchronous.
    auto enqueue_group = [&](int w, int slot) -> int {
        LmcuScene::PtWave& v = s->wave[w];
        cudaStream_t st = w == 0 ? s->stream : v.stream;
        for (int it = 0; it < kGroup; it++) {
            cudaEvent_t* ev = &v.ring[slot][it * 4];
            const int cur = v.cur;
            k_pt_advance<<<1, 1, 0, st>>>(pp[w], v.ctr.p, cur);
            if (pp[w].light_tracing) k_lt_generate<<<gen_blocks, 256, 0, st>>>(d, pp[w], q[w][cur], v.ctr.p);
            else k_pt_generate<<<gen_blocks, 256, 0, st>>>(d, pp[w], q[w][cur], sq[w], v.ctr.p);
            cudaEventRecord(ev[0], st);
            if (s->pool_on()) LMCU_DISPATCH_SMEM(k_pt_extend_p, s->fp32, count, q_blocks, LMCU_TRAV_BLOCK, pool_smem_pt, st, d, q[w][cur], v.hits.p, v.ctr.p, cur, v.arena.p);
            else LMCU_DISPATCH(k_pt_extend_q, s->tri_kind(), count, q_blocks, LMCU_TRAV_BLOCK, st, d, q[w][cur], v.hits.p, v.ctr.p, cur);
            cudaEventRecord(ev[1], st);
            if (pp[w].light_tracing) {
                if (s->general()) k_lt_shade<true><<<shade_blocks, LMCU_SHADE_BLOCK, 0, st>>>(d, pp[w], q[w][cur], q[w][cur ^ 1], sq[w], v.hits.p, v.ctr.p, cur);
                else k_lt_shade<false><<<shade_blocks, LMCU_SHADE_BLOCK, 0, st>>>(d, pp[w], q[w][cur], q[w][cur ^ 1], sq[w], v.hits.p, v.ctr.p, cur);
            }
            else if (s->composite) k_pt_shade<true><<<shade_blocks, LMCU_SHADE_BLOCK, 0, st>>>(d, pp[w], q[w][cur], q[w][cur ^ 1], sq[w], v.hits.p, s->film.p, v.ctr.p, cur);
            else if (s->two_level) k_pt_shade<false, true><<<shade_blocks, LMCU_SHADE_BLOCK, 0, st>>>(d, pp[w], q[w][cur], q[w][cur ^ 1], sq[w], v.hits.p, s->film.p, v.ctr.p, cur);
            else k_pt_shade<false><<<shade_blocks, LMCU_SHADE_BLOCK, 0, st>>>(d, pp[w], q[w][cur], q[w][cur ^ 1], sq[w], v.hits.p, s->film.p, v.ctr.p, cur);
            cudaEventRecord(ev[2], st);
            if (s->pool_on()) LMCU_DISPATCH_SMEM(k_pt_shadow_p, s->fp32, count, q_blocks, LMCU_TRAV_BLOCK, pool_smem_pt, st, d, sq[w], s->film.p, v.ctr.p, v.arena.p);
            else LMCU_DISPATCH(k_pt_shadow_q, s->tri_kind(), count, q_blocks, LMCU_TRAV_BLOCK, st, d, sq[w], s->film.p, v.ctr.p);
            cudaEventRecord(ev[3], st);
            v.cur ^= 1;
            acc.launches += 5;
            acc.extend_launches++;
            // fold the fp32 film into the FP64 one whenever this wavefront may have added ~32 samples per pixel to it
            v.since_fold += spp_per_iteration;
            if (v.since_fold >= 32.0) {
                k_film_fold<<<fold_blocks, 256, 0, st>>>(s->film.p, s->film_acc.p, s->film.n);
                v.since_fold = 0.0;
                acc.launches++;
            }
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(v.h + slot, v.ctr.p, sizeof(PtCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(v.grp[slot], st));
        v.cur_after[slot] = v.cur;
        return LMCU_OK;
    };
    // Two groups in flight per wavefront: the host reads the counters of group g while group g+1 runs, so no loop ever
    // waits for the host and the loops drift out of phase on their own.  The end of a wavefront is seen one group late:
    // the group already in flight then runs kGroup empty iterations (queues of 0 paths, a few microseconds).
    for (int slot = 0; slot < 2; slot++) {
        for (int w = 0; w < K; w++) {
            if (pp[w].total_paths == 0 && slot == 1) continue;
            const int rc = enqueue_group(w, slot);
            if (rc) return rc;
        }
    }
    int remaining = K;
    std::vector<uint64_t> started(K, 0);
    for (int slot = 0; remaining > 0; slot ^= 1) {
        if (acc.progress) {
            uint64_t sum = 0;
            for (int w = 0; w < K; w++) sum += started[w];
            acc.progress(acc.progress_user, acc.paths_before + (int64_t)sum, acc.paths_total,
                         std::chrono::duration<double>(std::chrono::steady_clock::now() - acc.t0).count());
        }
        for (int w = 0; w < K; w++) {
            LmcuScene::PtWave& v = s->wave[w];
            if (v.finished) continue;
            CU(cudaEventSynchronize(v.grp[slot]));
            const PtCounters& h = v.h[slot];
            for (int it = 0; it < kGroup; it++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, v.ring[slot][it * 4 + 0], v.ring[slot][it * 4 + 1]); acc.t_extend += ms * 1e-3;
                cudaEventElapsedTime(&ms, v.ring[slot][it * 4 + 1], v.ring[slot][it * 4 + 2]); acc.t_shade += ms * 1e-3;
                cudaEventElapsedTime(&ms, v.ring[slot][it * 4 + 2], v.ring[slot][it * 4 + 3]); acc.t_shadow += ms * 1e-3;
            }
            if (h.overflow) return fail(LMCU_ERR_CUDA, "bvh traversal stack overflow");
            started[w] = h.next_path;
            if (h.next_path >= pp[w].total_paths && h.count[v.cur_after[slot]] == 0) {
                // done; the other slot's group (if one was enqueued) only runs empty iterations: wait for it, then join
                v.finished = true;
                remaining--;
                cudaStream_t st = w == 0 ? s->stream : v.stream;
                CU(cudaStreamSynchronize(st));
                acc.rays += h.rays; acc.extend_rays += h.extend_rays; acc.shadow_rays += h.shadow_rays;
                acc.nodes_extend += h.nodes_extend; acc.tris_extend += h.tris_extend;
                acc.nodes_shadow += h.nodes_shadow; acc.tris_shadow += h.tris_shadow;
                if (w > 0) {                                            // later work on s->stream sees this wave's film adds
                    CU(cudaEventRecord(v.done, v.stream));
                    CU(cudaStreamWaitEvent(s->stream, v.done, 0));
                }
            } else {
                const int rc = enqueue_group(w, slot);
                if (rc) return rc;
            }
        }
    }
    return LMCU_OK;
}

static int render_paths(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats, bool light_tracing);

int lmcu_render_pt(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats) {
    return render_paths(s, p, film_rgb, stats, false);
}

// renderer::lt (src/renderer/renderer_lt.cpp): scheduler::spi::* only ("Scheduler is fixed to image space scheduler", :44-47)
int lmcu_render_lt(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats) {
    return render_paths(s, p, film_rgb, stats, true);
}

static int render_paths(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats, bool light_tracing) {
    int rc = require_renderable(s, true);
    if (rc) return rc;
    if (!p) return fail(LMCU_ERR_INVALID_ARGUMENT, "null params");
    const bool image = light_tracing || p->primary_ray_sampling_mode == LMCU_PRIMARY_IMAGE;
    const bool timed = p->scheduler == LMCU_SCHED_TIME;
    if (p->primary_ray_sampling_mode != LMCU_PRIMARY_PIXEL && !image) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad primary_ray_sampling_mode");
    if (light_tracing && !s->has_camera) return fail(LMCU_ERR_UNSUPPORTED, "Missing camera primitive.");
    if (p->scheduler != LMCU_SCHED_SAMPLE && !timed) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad scheduler");
    if (timed) {
        if (!(p->render_time >= 0)) return fail(LMCU_ERR_INVALID_ARGUMENT, "render_time must be >= 0");
    } else if (image) {
        if (p->num_samples <= 0) return fail(LMCU_ERR_INVALID_ARGUMENT, "num_samples must be positive");
    } else if (p->spp <= 0) {
        return fail(LMCU_ERR_INVALID_ARGUMENT, "spp must be positive");
    }
    if (p->max_verts < 2) return fail(LMCU_ERR_INVALID_ARGUMENT, "max_verts must be >= 2");
    if (p->sampling_mode < 0 || p->sampling_mode > 2) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad sampling_mode");
    CU(cudaSetDevice(s->device));
    rc = ensure_film(s, p->width, p->height, true);
    if (rc) return rc;
    if (p->wavefronts < 0 || p->wavefronts > LmcuScene::kMaxWaves) return fail(LMCU_ERR_INVALID_ARGUMENT, "wavefronts must be in 0..4");
    s->waves = p->wavefronts > 0 ? p->wavefronts : s->default_waves;
    // Paths in flight per wavefront: 16 Mi (208 B each: 72 B x 2 queues + 48 B shadow + 16 B hit = 3.4 GB per wavefront; the
    // 180 GB of HBM make large launches cheap and they amortise the persistent kernels' tails: C2 at 1024 spp, two
    // wavefronts of 4 / 8 / 16 / 32 Mi paths: 1709 (256 spp) / 1772 / 1812 / 1826 Mpaths/s), never more than the job
    // needs.  LMCU_PT_POOL overrides it.
    uint32_t capacity = 1u << 24;
    if (!timed) {
        const uint64_t npx = (uint64_t)p->width * p->height;
        const uint64_t job = image ? (uint64_t)p->num_samples : npx * (uint64_t)p->spp;
        const uint64_t per_wave = (job + (uint64_t)s->waves - 1) / (uint64_t)s->waves;
        const uint64_t need = std::max<uint64_t>(1u << 16, (per_wave + 0xffffull) & ~0xffffull);
        if (need < capacity) capacity = (uint32_t)need;
    }
    if (const char* e = getenv("LMCU_PT_POOL")) capacity = (uint32_t)std::max(1024l, atol(e));
    rc = ensure_pt_pools(s, capacity, image ? 2u : 1u);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const uint64_t npix = (uint64_t)p->width * p->height;
    PtParams pp{};
    pp.width = p->width; pp.height = p->height; pp.max_verts = p->max_verts; pp.sampling_mode = p->sampling_mode;
    pp.seed = p->seed; pp.tiled = (p->width % 8 == 0 && p->height % 4 == 0) ? 1u : 0u;
    pp.capacity = capacity;
    pp.image_mode = image ? 1u : 0u;
    pp.light_tracing = light_tracing ? 1u : 0u;
    const bool count = p->count_traversal != 0;
    PtAccum acc;
    acc.progress = p->progress; acc.progress_user = p->progress_user;
    acc.t0 = std::chrono::steady_clock::now();
    // `unit` = what one scheduler step processes: one sample of every pixel (scheduler::spp) or one sample of the image
    // (scheduler::spi); a pass renders `units` of them starting at sample index sample_offset + processed.
    const uint64_t paths_per_unit = image ? 1ull : npix;
    int64_t processed = 0;
    auto run_units = [&](int64_t units) -> int {
        pp.sample_offset = (uint64_t)p->sample_offset + (uint64_t)processed;
        pp.total_paths = paths_per_unit * (uint64_t)units;
        // pixel mode: samples of a pixel are walked in chunks of the largest power of two <= 16 that divides this pass
        pp.chunk = 1u;
        if (!image) { uint32_t c = s->path_chunk; while (c > 1u && (units % c) != 0) c >>= 1; pp.chunk = c; }
        acc.paths_before = (int64_t)(paths_per_unit * (uint64_t)processed);
        acc.paths_total = timed ? 0 : (int64_t)pp.total_paths;
        const int r = pt_pass(s, pp, count, acc);
        if (r == LMCU_OK) processed += units;
        return r;
    };
    CU(cudaEventRecord(s->ev0, st));
    if (!timed) {
        rc = run_units(image ? p->num_samples : p->spp);
        if (rc) return rc;
    } else {
        // scheduler::{spp,spi}::time (src/scheduler.cpp:75-103,162-187): whole passes until the clock passes render_time,
        // checked after each pass.  The reference's pass is 1 spp (or samples_per_iter samples); here a pass is a
        // multiple of that, sized from the measured rate so that the wavefront stays full and the overshoot small.
        const int64_t iter = image ? (p->samples_per_iter > 0 ? p->samples_per_iter : 100000) : 1;   // reference granularity
        const auto t0 = std::chrono::steady_clock::now();
        int64_t units = std::max<int64_t>(1, (int64_t)((uint64_t)capacity * (uint64_t)s->waves / paths_per_unit) / iter) * iter;    // first pass: full queues
        while (true) {
            rc = run_units(units);
            if (rc) return rc;
            const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (elapsed > p->render_time) break;
            const double rate = (double)processed / std::max(elapsed, 1e-6);                          // units per second
            const double want = (p->render_time - elapsed) * rate * 0.5;                              // approach the budget geometrically
            units = std::max<int64_t>(1, (int64_t)(want / (double)iter)) * iter;
        }
    }
    {
        // last fold (every wavefront has joined s->stream), then the film the callers see = float(FP64 sums * scale);
        // renderer_pt.cpp:307-312: scale = 1/processed (pixel mode) or w*h/processed (image mode); 1 when the caller
        // combines films of several GPUs first (rescale = 0)
        const double scale = !p->rescale ? 1.0 : image ? (double)npix / (double)processed : 1.0 / (double)processed;
        const unsigned blocks = (unsigned)std::min<size_t>((s->film.n + 255) / 256, (size_t)s->num_sms * 8);
        k_film_fold<<<blocks, 256, 0, st>>>(s->film.p, s->film_acc.p, s->film.n);
        k_film_finish<<<blocks, 256, 0, st>>>(s->film.p, s->film_acc.p, s->film.n, scale);
        acc.launches += 2;
    }
    CU(cudaEventRecord(s->ev1, st));
    if (film_rgb) CU(cudaMemcpyAsync(film_rgb, s->film.p, s->film.n * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        memset(stats, 0, sizeof(*stats));
        stats->elapsed = ms * 1e-3; stats->processed = processed; stats->paths = (int64_t)(paths_per_unit * (uint64_t)processed);
        stats->rays = (int64_t)acc.rays; stats->kernel_launches = acc.launches;
        stats->extend_seconds = acc.t_extend; stats->shade_seconds = acc.t_shade; stats->shadow_seconds = acc.t_shadow;
        stats->extend_launches = acc.extend_launches;
        stats->extend_rays = (int64_t)acc.extend_rays; stats->shadow_rays = (int64_t)acc.shadow_rays;
        if (count && acc.extend_rays) {
            stats->bvh_nodes_per_ray = (double)acc.nodes_extend / (double)acc.extend_rays;
            stats->bvh_tris_per_ray = (double)acc.tris_extend / (double)acc.extend_rays;
            stats->shadow_nodes_per_ray = acc.shadow_rays ? (double)acc.nodes_shadow / (double)acc.shadow_rays : 0.0;
            stats->shadow_tris_per_ray = acc.shadow_rays ? (double)acc.tris_shadow / (double)acc.shadow_rays : 0.0;
        }
    }
    return LMCU_OK;
}

// renderer::bdpt (src/renderer/renderer_bdpt.cpp:304-407).  Passes of up to 1 Mi samples: random walks (k_bdpt_gen, then
// k_pt_extend_q -> k_bdpt_vertex per vertex index), k_bdpt_lists -> k_bdpt_connect -> k_pt_shadow_q; the fp32 splat film is folded into the FP64 sums after every pass.
int lmcu_render_bdpt(LmcuScene* s, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats) {
    int rc = require_renderable(s, true);
    if (rc) return rc;
    if (!p) return fail(LMCU_ERR_INVALID_ARGUMENT, "null params");
    const bool timed = p->scheduler == LMCU_SCHED_TIME;
    if (p->scheduler != LMCU_SCHED_SAMPLE && !timed) return fail(LMCU_ERR_INVALID_ARGUMENT, "bad scheduler");
    if (timed) {
        if (!(p->render_time >= 0)) return fail(LMCU_ERR_INVALID_ARGUMENT, "render_time must be >= 0");
    } else if (p->num_samples <= 0) {
        return fail(LMCU_ERR_INVALID_ARGUMENT, "num_samples must be positive");
    }
    if (p->max_verts < 2 || p->max_verts > kBdptMaxVerts) return fail(LMCU_ERR_INVALID_ARGUMENT, "bdpt: max_verts must be in 2..20");
    const int min_verts = p->min_verts > 0 ? p->min_verts : 2;
    const int mode = p->bdpt_strategies;
    if (mode < LMCU_BDPT_ALL || mode > LMCU_BDPT_LT_NEE) return fail(LMCU_ERR_INVALID_ARGUMENT, "bdpt: bad bdpt_strategies");
    CU(cudaSetDevice(s->device));
    rc = ensure_film(s, p->width, p->height, true);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    const uint64_t npix = (uint64_t)p->width * p->height;
    PtParams pp{};
    pp.width = p->width; pp.height = p->height; pp.max_verts = p->max_verts; pp.seed = p->seed;
    pp.image_mode = 1u;
    pp.bdpt_mode = (uint32_t)mode;
    const DevScene d = s->dev();
    const unsigned fold_blocks = (unsigned)std::min<size_t>((s->film.n + 255) / 256, (size_t)s->num_sms * 8);
    const auto t0 = std::chrono::steady_clock::now();
    int64_t processed = 0, launches = 0;
    // strategies (s, t) with min_verts <= s + t <= max_verts (renderer_bdpt.cpp:350-361); t = 0 never connects to a pinhole
    std::vector<int2> pairs;
    // the single-strategy renderers (:80-93,150-170,228-233) keep one family; pt_nee_path picks between (1, n-1) and
    // (0, n) per sample (k_bdpt_lists: bdpt_strategy_on)
    for (int n = std::max(2, min_verts); n <= p->max_verts; n++)
        for (int sL = 0; sL < n; sL++) {
            if ((mode == LMCU_BDPT_PT_NAIVE && sL != 0) || (mode == LMCU_BDPT_PT_NEE && sL > 1) || (mode == LMCU_BDPT_LT_NEE && sL != n - 1)) continue;
            pairs.push_back(make_int2(sL, n - sL));
        }
    if (pairs.empty()) return fail(LMCU_ERR_INVALID_ARGUMENT, "bdpt: min_verts exceeds max_verts");
    // samples per pass: the shadow queue holds one entry per (strategy, sample), at most 32 Mi (1.5 GB)
    LmcuScene::BdptPools& B = s->bdpt;
    const uint32_t cap = (uint32_t)std::min<size_t>(size_t(1) << 20, std::max<size_t>(size_t(1) << 14, ((size_t(1) << 25) / pairs.size()) & ~size_t(127)));
    if (B.cap != cap || B.mv != p->max_verts) {
        const size_t nv = (size_t)2 * p->max_verts * cap;
        CU(B.va.alloc(nv)); CU(B.vb.alloc(nv)); CU(B.vc.alloc(nv)); CU(B.vd.alloc(nv)); CU(B.ve.alloc(nv));
        CU(B.vt.alloc(0));
        CU(B.nverts.alloc((size_t)2 * cap)); CU(B.eye_pixel.alloc(cap));
        for (int k = 0; k < 2; k++) { CU(B.q_o[k].alloc((size_t)2 * cap)); CU(B.q_d[k].alloc((size_t)2 * cap)); CU(B.q_id[k].alloc((size_t)2 * cap)); }
        CU(B.hits.alloc((size_t)2 * cap));
        if (!B.ctr.p) CU(B.ctr.alloc(1));
        B.cap = cap; B.mv = p->max_verts;
    }
    if (B.sq_cap < pairs.size() * cap) {
        for (int k = 0; k < 3; k++) CU(B.sq[k].alloc(pairs.size() * cap));
        CU(B.pair_list.alloc(pairs.size() * cap));
        B.sq_cap = pairs.size() * cap;
    }
    if (B.pair_count.n < pairs.size()) CU(B.pair_count.alloc(pairs.size()));
    if (B.pairs.n < pairs.size()) CU(B.pairs.alloc(pairs.size()));
    CU(cudaMemcpyAsync(B.pairs.p, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(B.ctr.p, 0, sizeof(PtCounters), st));
    BdptStore store;
    store.va = B.va.p; store.vb = B.vb.p; store.vc = B.vc.p; store.vd = B.vd.p; store.ve = B.ve.p;
    if (d.uv && B.vt.n != (size_t)2 * p->max_verts * cap) CU(B.vt.alloc((size_t)2 * p->max_verts * cap));
    store.vt = d.uv ? B.vt.p : nullptr;
    store.nverts = B.nverts.p; store.eye_pixel = B.eye_pixel.p; store.cap = cap; store.mv = p->max_verts;
    ShadowQueue sq;
    sq.ray_o = B.sq[0].p; sq.ray_d = B.sq[1].p; sq.contrib = B.sq[2].p;
    const unsigned q_blocks = (unsigned)(s->num_sms * s->trav_ctas(true));
    auto run_pass = [&](int64_t count) -> int {
        for (int64_t done = 0; done < count;) {
            const int64_t n = std::min<int64_t>(cap, count - done);
            pp.sample_offset = (uint64_t)p->sample_offset + (uint64_t)processed;
            pp.total_paths = (uint64_t)n;
            pp.capacity = cap;
            k_bdpt_reset<<<1, 1, 0, st>>>(B.ctr.p, (uint32_t)(mode == LMCU_BDPT_ALL ? 2 * n : n));   // rays k_bdpt_gen enqueues
            const unsigned gen_blocks = (unsigned)std::min<int64_t>((2 * n + 255) / 256, (int64_t)s->num_sms * 8);
            const unsigned vtx_blocks = (unsigned)std::min<int64_t>((2 * n + 127) / 128, (int64_t)s->num_sms * 8);
            k_bdpt_gen<<<gen_blocks, 256, 0, st>>>(d, pp, store, B.q_o[0].p, B.q_d[0].p, B.q_id[0].p);
            int cur = 0;
            for (int k = 1; k < p->max_verts; k++) {
                PathQueue q{};
                q.ray_o = B.q_o[cur].p; q.ray_d = B.q_d[cur].p;
                LMCU_DISPATCH(k_pt_extend_q, s->tri_kind(), false, q_blocks, LMCU_TRAV_BLOCK, st, d, q, B.hits.p, B.ctr.p, cur);
                if (s->general()) k_bdpt_vertex<true><<<vtx_blocks, 128, 0, st>>>(d, pp, store, k, B.q_d[cur].p, B.q_id[cur].p, B.hits.p, B.q_o[cur ^ 1].p, B.q_d[cur ^ 1].p, B.q_id[cur ^ 1].p, B.ctr.p, cur);
                else k_bdpt_vertex<false><<<vtx_blocks, 128, 0, st>>>(d, pp, store, k, B.q_d[cur].p, B.q_id[cur].p, B.hits.p, B.q_o[cur ^ 1].p, B.q_d[cur ^ 1].p, B.q_id[cur ^ 1].p, B.ctr.p, cur);
                k_bdpt_flip<<<1, 1, 0, st>>>(B.ctr.p, cur);
                cur ^= 1;
                launches += 3;
            }
            CU(cudaMemsetAsync(B.pair_count.p, 0, pairs.size() * sizeof(uint32_t), st));
            const dim3 lgrid((unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)s->num_sms), (unsigned)pairs.size());
            k_bdpt_lists<<<lgrid, 256, 0, st>>>(pp, store, B.pairs.p, B.pair_list.p, B.pair_count.p);
            const dim3 cgrid((unsigned)std::min<int64_t>((n + 127) / 128, (int64_t)s->num_sms * 2), (unsigned)pairs.size());
            if (s->general()) k_bdpt_connect<true><<<cgrid, 128, 0, st>>>(d, pp, store, B.pairs.p, B.pair_list.p, B.pair_count.p, sq, s->film.p, B.ctr.p);
            else k_bdpt_connect<false><<<cgrid, 128, 0, st>>>(d, pp, store, B.pairs.p, B.pair_list.p, B.pair_count.p, sq, s->film.p, B.ctr.p);
            LMCU_DISPATCH(k_pt_shadow_q, s->tri_kind(), false, q_blocks, LMCU_TRAV_BLOCK, st, d, sq, s->film.p, B.ctr.p);
            CU(cudaGetLastError());
            k_film_fold<<<fold_blocks, 256, 0, st>>>(s->film.p, s->film_acc.p, s->film.n);
            launches += 7;
            processed += n; done += n;
            if (p->progress) {
                CU(cudaStreamSynchronize(st));
                p->progress(p->progress_user, processed, timed ? 0 : p->num_samples,
                            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            }
        }
        return LMCU_OK;
    };
    CU(cudaEventRecord(s->ev0, st));
    if (!timed) {
        rc = run_pass(p->num_samples);
        if (rc) return rc;
    } else {
        // scheduler::spi::time (src/scheduler.cpp:162-187): whole passes of samples_per_iter until the clock passes render_time
        const int64_t iter = p->samples_per_iter > 0 ? p->samples_per_iter : 100000;
        int64_t units = iter;
        while (true) {
            rc = run_pass(units);
            if (rc) return rc;
            CU(cudaStreamSynchronize(st));
            const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (elapsed > p->render_time) break;
            const double rate = (double)processed / std::max(elapsed, 1e-6);
            units = std::max<int64_t>(1, (int64_t)((p->render_time - elapsed) * rate * 0.5 / (double)iter)) * iter;
        }
    }
    const double scale = !p->rescale ? 1.0 : (double)npix / (double)processed;      // renderer_bdpt.cpp:390-391
    k_film_finish<<<fold_blocks, 256, 0, st>>>(s->film.p, s->film_acc.p, s->film.n, scale);
    launches += 1;
    CU(cudaEventRecord(s->ev1, st));
    if (film_rgb) CU(cudaMemcpyAsync(film_rgb, s->film.p, s->film.n * sizeof(float), cudaMemcpyDeviceToHost, st));
    PtCounters hc;
    CU(cudaMemcpyAsync(&hc, B.ctr.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (stats) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        memset(stats, 0, sizeof(*stats));
        stats->elapsed = ms * 1e-3; stats->processed = processed; stats->paths = processed;
        stats->rays = (int64_t)hc.rays; stats->shadow_rays = (int64_t)hc.shadow_rays; stats->extend_rays = (int64_t)hc.extend_rays;
        stats->kernel_launches = launches;
    }
    if (hc.overflow) return fail(LMCU_ERR_CUDA, "bvh traversal stack overflow");
    return check_overflow(s);
}

}  // extern "C"
