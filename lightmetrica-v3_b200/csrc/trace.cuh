// trace.cuh — warp-persistent BVH traversal over a ray QUEUE with per-lane dynamic ray fetch.
//
// This is the form of bvh.cuh's traversal that the queue kernels (accel batches, pt extend, pt shadow)
// run.  Same node test, same FP64 leaf test, same results; what changes is the scheduling:
//
//   * every lane owns one ray and runs a small state machine; one loop iteration = at most ONE
//     wide-node step (for lanes with no pending triangles) followed by at most ONE FP64 triangle
//     test (for lanes with pending triangles), so node work and leaf work of different lanes
//     proceed in lock-step instead of each lane's leaf loop stalling the other 31;
//   * a lane whose ray terminates stores its result at once and goes idle; as soon as
//     `fetch_threshold` lanes are idle the warp leaves the traversal loop and the idle lanes claim
//     new rays with ONE atomic on the queue cursor (ballot-ranked), so a long ray no longer holds
//     31 finished lanes hostage until the end of the batch;
//   * the two barycentric divisions u/|det|, v/|det| (accel_sahbvh.cpp:74) are applied once per
//     ray, to the final closest hit, not to every intermediate candidate.  The quotients are the
//     same IEEE operations on the same operands, so (t,u,v) stay bit-identical to accel::sahbvh.
//
// IO policy (template parameter), all device functions:
//     bool load(uint32_t i, RayD& r, RayF& f, float& tmin_dn, float& tmax_up)   ray i of the queue, FP64 + FP32 views
//     static constexpr bool kExactT                        true: t is the IEEE quotient of accel_sahbvh.cpp:70 (bit-exact hits)
//     void store(uint32_t i, double t, double u_raw, double v_raw, double ad, int32_t tri)
//                                                          result of ray i (tri < 0 = miss); u = u_raw / ad, v = v_raw / ad
#pragma once
#include "bvh.cuh"

namespace lmcu {

#if defined(__CUDACC__)


// 1/d with relative error ~2^-46: MUFU.RCP64H seed (2^-23) and one Newton step
__device__ __forceinline__ double rcp_newton(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    const double e = fma(-d, x, 1.0);
    return fma(x, e, x);
}

template <bool ANY, bool COUNT, typename TriT, typename IO>
__device__ __forceinline__ void trace_queue(const WideNode* __restrict__ nodes, const TriT* __restrict__ tris,
                                            uint32_t n, uint32_t* cursor, const IO& io, uint32_t* overflow,
                                            TraverseCounters* ctr, int fetch_threshold, const InstDev* __restrict__ inst = nullptr) {
    constexpr bool TWO = TriT::kTwoLevel;   // two-level tree (bvh.cuh): `inst` = the top level's leaf-ordered instance table
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint4* nodes4 = reinterpret_cast<const uint4*>(nodes);
    uint2 stack[kStackSize];
    // (Measured and dropped, profiles/r02_ab_smem_stack.log: the top 4 / 8 entries of every thread's stack in shared memory
    // ([depth][thread], deeper entries spilling to this array): C2 1915 -> 1921 / 1886 Mpaths/s, the extend kernel alone 0.7 %
    // slower — the stack is 69 bytes of local-memory traffic per ray, 91-97 % of it served by L1, profiles/r02_ncu_summary.md.)
#define LMCU_PUSH(E) { stack[sp++] = (E); }
#define LMCU_POP() (stack[--sp])

    // per-lane traversal state
    bool active = false;
    bool pending = false;         // the ray of this lane is finished, its result not yet stored (see the fetch block)
    bool exhausted = false;       // the cursor ran past n: nothing left to fetch
    uint32_t idx = 0;
    RayD r; r.o = d3(0, 0, 0); r.d = d3(0, 0, 1); r.tmin = 0; r.tmax = 0;
    RayF f = make_rayf(r);
    float tmin_dn = 0.f, tmax_up = 0.f;
    double ht = 0, hu = 0, hv = 0, had = 1;   // closest candidate: t, raw u, raw v, |det|
    int32_t htri = -1;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
    int sp = 0;
    uint32_t cinst = kNoInst;     // TWO: the instance whose tree the lane is in (f is then the ray in its object space)

    while (true) {
        // ---- results of the rays that finished since the last visit here.  A lane keeps its finished ray's (t, u, v, |det|,
        // triangle) in its registers and stores them HERE, together with the other lanes that finished, instead of at the
        // moment it finishes: inside the traversal loop 2-3 lanes finish per iteration, so the IO policy's store — two exact
        // FP64 divisions, a triangle and a primitive-table fetch for accel batches — ran at 2-3 of 32 lanes almost every
        // iteration; here it runs once per refill for >= fetch_threshold lanes.
        if (pending) {
            io.store(idx, ht, hu, hv, had, htri);
            pending = false;
        }
        // ---- fetch: idle lanes claim the next rays of the queue -------------------------------------
        if (!exhausted) {
            const uint32_t idle = __ballot_sync(0xffffffffu, !active);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (base + (uint32_t)__popc(idle) >= n) exhausted = true;     // warp-uniform
                if (!active) {
                    idx = base + (uint32_t)__popc(idle & lt_mask);
                    if (idx < n && io.load(idx, r, f, tmin_dn, tmax_up)) {
                        ht = r.tmax; htri = -1;
                        ngroup = make_uint2(0u, 0x80000000u);   // the root occupies virtual slot 7 of group 0
                        tgroup = make_uint2(0u, 0u);
                        sp = 0;
                        cinst = kNoInst;
                        active = true;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, active)) {
            if (exhausted) break;
            continue;
        }
        // ---- traverse until enough lanes are idle to make a re-fetch worth its latency ------------------
        while (true) {
            if (active) {
                if constexpr (TWO) {
                    // ---- nothing in hand: pop.  A marker takes the ray back to the top level (world space) and the entry
                    // under it is popped with it; a top-level leaf's remaining instances come back as a triangle-group entry
                    bool to_world = false;
                    if (tgroup.y == 0u && ngroup.y <= 0x00ffffffu && sp > 0) {
                        uint2 e = LMCU_POP();
                        if (e.y == 0u) {
                            cinst = kNoInst;
                            to_world = true;
                            if (sp > 0) e = LMCU_POP();
                        }
                        if (e.y > 0x00ffffffu) ngroup = e;
                        else if (e.y != 0u) tgroup = e;
                    }
                    // ---- a lane in the top level that holds leaf items (from the node step of the last iteration, or popped
                    // just now) enters the first of them: what is left of the node is parked, the ray goes to the instance's
                    // object space, and the node step below starts on the mesh tree's root in this same iteration
                    if (cinst == kNoInst && tgroup.y != 0u) {
                        const int ei = bfind(tgroup.y);
                        tgroup.y &= ~(1u << ei);
                        const uint32_t in = tgroup.x + (uint32_t)ei;
                        if (sp + 3 > kStackSize) { atomicExch(overflow, 1u); sp = kStackSize - 3; }
                        if (ngroup.y > 0x00ffffffu) LMCU_PUSH(ngroup)
                        if (tgroup.y != 0u) LMCU_PUSH(tgroup)
                        LMCU_PUSH(make_uint2(0u, 0u))
                        cinst = in;
                        f = make_rayf_obj(r, inst[in].Minv);
                        ngroup = make_uint2(inst[in].root, 0x80000000u);
                        tgroup = make_uint2(0u, 0u);
                        to_world = false;                     // straight into the next instance of the same leaf
                    }
                    if (to_world) f = make_rayf(r);
                }
                if (tgroup.y == 0u) {
                    if constexpr (!TWO) {
                        if (ngroup.y <= 0x00ffffffu) {
                            if (sp > 0) ngroup = LMCU_POP();
                        }
                    }
                    if (ngroup.y > 0x00ffffffu) {
                        // ---- one wide-node step
                        const uint32_t hits = ngroup.y;
                        const uint32_t imask = hits & 0xffu;
                        const int child_bit = bfind(hits);
                        const uint32_t base = ngroup.x;
                        ngroup.y &= ~(1u << child_bit);
                        if (ngroup.y > 0x00ffffffu) {
                            if (sp >= kStackSize) { atomicExch(overflow, 1u); sp = kStackSize - 1; }
                            LMCU_PUSH(ngroup)
                        }
                        const uint32_t slot = (uint32_t)(child_bit - 24) ^ (f.octinv4 & 0xffu);
                        const uint32_t ni = base + popc(imask & ~(0xffffffffu << slot));
                        const uint4 n0 = __ldg(nodes4 + ni * 5 + 0), n1 = __ldg(nodes4 + ni * 5 + 1),
                                    n2 = __ldg(nodes4 + ni * 5 + 2), n3 = __ldg(nodes4 + ni * 5 + 3),
                                    n4 = __ldg(nodes4 + ni * 5 + 4);
                        if (COUNT) ctr->nodes++;
                        const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, f, tmin_dn, tmax_up);
                        ngroup.x = n1.x;
                        ngroup.y = (hm & 0xff000000u) | (n0.w >> 24);
                        tgroup.x = n1.y;
                        tgroup.y = hm & 0x00ffffffu;
                    } else {
                        // ---- nothing left: this ray is done; its result is stored at the next refill (top of the outer loop)
                        pending = true;
                        active = false;
                    }
                }
                // ---- leaf work: FP64 Möller–Trumbore, reference operation order (accel_sahbvh.cpp:58-75)
                if (active && tgroup.y != 0u && (!TWO || cinst != kNoInst)) {
                    const int ti = bfind(tgroup.y);
                    tgroup.y &= ~(1u << ti);
                    const uint32_t tri = tgroup.x + (uint32_t)ti;
                    D3 p1, p2, p3;
                    int32_t id = (int32_t)tri;
                    load_tri(tris, tri, p1, p2, p3);
                    if constexpr (TWO) {
                        // the vertices in world space, by the flattened build's operations: the test below sees the same operands
                        const double* M = inst[cinst].M;
                        p1 = transform_point34(M, p1); p2 = transform_point34(M, p2); p3 = transform_point34(M, p3);
                        id = (int32_t)(inst[cinst].tri_offset + (uint32_t)tris[tri].face);
                    }
                    if (COUNT) ctr->tris++;
                    const D3 e1 = dsub3(p2, p1), e2 = dsub3(p3, p1);
                    const D3 p = dcross(r.d, e2);
                    const D3 tv = dsub3(r.o, p1);
                    const D3 q = dcross(tv, e1);
                    const double d = ddot(e1, p);
                    const double ad = fabs(d);
                    const double u = mul_sign_of(ddot(tv, p), d);
                    const double v = mul_sign_of(ddot(r.d, q), d);
                    if (!(ad < 1e-8 || u < 0.0 || v < 0.0 || addrn(u, v) > ad)) {
                        // accel batches: the reference's IEEE quotient.  Path-tracer queues (their hit record is FP32): a
                        // Newton-refined reciprocal, relative error 2^-46 — the exact division is ~30 instructions that
                        // run at the 3-4 lanes whose triangle passed the barycentric test.
                        const double t = IO::kExactT ? divrn(ddot(e2, q), d) : mulrn(ddot(e2, q), rcp_newton(d));
                        if (!(t < r.tmin || ht < t)) {          // inclusive bounds; a later equal-t triangle wins
                            ht = t; hu = u; hv = v; had = ad; htri = id;
                            tmax_up = f32_up(t);
                            if (ANY) {
                                pending = true;
                                active = false;
                            }
                        }
                    }
                }
            }
            const uint32_t act = __ballot_sync(0xffffffffu, active);
            if (act == 0u) break;
            if (!exhausted && 32 - __popc(act) >= fetch_threshold) break;     // enough idle lanes: refill them
        }
    }
}


#undef LMCU_PUSH
#undef LMCU_POP

#endif  // __CUDACC__

}  // namespace lmcu

// Variants measured and dropped (profiles/r01_ncu_summary_v2.md, _v6.md, gpurun logs of round 1): an L1 prefetch of the
// child a lane descends into next, issued at the end of the node step (-3 %: 8 more instructions per step for a node that
// hits L1 80 % of the time anyway), and of a node's first triangles when the node arrives (-4 %); an L2 prefetch of the queue 16 K / 64 K rays ahead of the cursor (+0.3 %, noise); the
// FP64 triangle test of the
// path-tracer queues with fused multiply-adds (14 fewer FP64 instructions, but the DFMA chains are longer than the
// independent DMULs they replace: extend +4 % time, 1296 vs 1321 Mpaths/s on C2); warp-private chunks of 32..128
// rays claimed with one atomic, the next chunk claimed in advance and prefetched toward L2 (-3 % on C2 pt, 1195 vs 1232
// Mpaths/s at 256 spp, for every refill threshold 4..16: the atomic's latency is already covered by the scheduler's other
// warps, whole-chunk tails at the end of each launch are not; chunk 128 lost 30 % on a 2 M-ray batch); 3 triangle tests
// per iteration (-6 % Mpaths/s on C2), a CWBVH-style accumulated "lost iterations" refetch credit (never paid
// back on rays this short), and a phase-partitioned loop that postpones FP64 triangle tests until >= 12..24
// lanes hold some (triangle-step occupancy 9.5 -> 12.8 lanes, but the speculative node steps cost as much as
// the batching saved: -2 % on C2, -20 % on C3 incoherent rays).
