// trace2.cuh — BVH traversal over a ray queue with a shared-memory RAY POOL: every lane owns K rays and every loop
// iteration of the warp is ONE phase — either a wide-node step or an FP64 triangle test — executed by all lanes that hold
// a ray ready for that phase.
//
// Why: the register-resident loop of trace.cuh runs a node step for the lanes without pending triangles followed by a
// triangle test for the lanes with pending triangles.  Both are long straight-line blocks (254 and ~150 instructions) and
// each runs with the OTHER kind of lane masked off: ncu measures 15.7 of 32 lanes active on C2 (17 in node steps, 12 in
// triangle tests) with the kernel bound by instruction issue, not by memory (profiles/r02_*).  A warp instruction costs an
// issue slot whatever its occupancy, so the lever is lanes per instruction.  With K rays per lane a lane almost always
// holds ONE ray that is ready for the phase the warp votes for (with p = fraction of rays waiting for a node step ~0.6:
// 1 - (1-p)^K = 84 % at K = 2, against p itself = 60 % with one ray per lane; triangle phase 64 % against 40 %).
//
// The rays' traversal state lives in shared memory, SoA in 16-byte chunks indexed [chunk][slot], slot = lane + 32 k: a
// lane only ever touches its own K slots, so every access is a conflict-free LDS.128 / STS.128 and nothing is exchanged
// between lanes.  A phase step loads the chunks it needs (node step: 3, triangle test: 6), runs exactly the arithmetic of
// trace.cuh (same node test, same FP64 Möller–Trumbore in the reference's operation order, same per-ray order of triangle
// tests => same ties, same (t,u,v) bit for bit) and stores back what changed.  The traversal stacks sit in a global
// arena laid out [warp][depth][slot] (coalesced pushes and pops, L1-resident like the local-memory stacks of trace.cuh).
//
// IO policy: the one of trace.cuh (load / store / kExactT).
#pragma once
#include "trace.cuh"

namespace lmcu {

#if defined(__CUDACC__)

// 16-byte chunks of state per ray: 9 (144 B) for the path-tracer queues, 10 for accel batches (c9)
template <typename IO> __host__ __device__ constexpr int pool_chunks() { return IO::kExactT ? 10 : 9; }
//  c0: o.x o.y (f64)   c1: o.z d.x   c2: d.y d.z         FP64 ray, as the reference sees it (triangle test)
//  c3: ox oy oz (f32), tmax_up                            FP32 view (node test) + float upper bound of the search interval
//  c4: 1/dx 1/dy 1/dz (f32), tmin_dn
//  c5: ht, |det| (f64)                                    closest candidate
//  c6: u_raw, v_raw (f64)
//  c7: ngroup.x ngroup.y tgroup.x tgroup.y
//  c8: htri, queue index, sp | octinv << 8, -
//  c9: tmin (f64), -                                       lower bound as the reference sees it (accel batches only: the
//                                                          path tracer's tmin IS the float tmin_dn)

template <typename IO, int K> __host__ __device__ constexpr size_t pool_bytes_per_warp() { return (size_t)pool_chunks<IO>() * 32 * K * 16; }

enum : uint32_t { SLOT_EMPTY = 0u, SLOT_NODE = 1u, SLOT_TRI = 2u };

template <bool ANY, bool COUNT, int K, typename TriT, typename IO>
__device__ __forceinline__ void trace_pool(const WideNode* __restrict__ nodes, const TriT* __restrict__ tris,
                                           uint32_t n, uint32_t* cursor, const IO& io, uint32_t* overflow,
                                           TraverseCounters* ctr, int refill_threshold, float4* pool /* this warp's */,
                                           uint2* stack_arena /* this warp's: [kStackSize][32 K] */) {
    constexpr int S = 32 * K;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint4* nodes4 = reinterpret_cast<const uint4*>(nodes);
    double2* pool_d = reinterpret_cast<double2*>(pool);
    uint4* pool_u = reinterpret_cast<uint4*>(pool);
#define CH(c, slot) ((c) * S + (slot))

    uint32_t state = 0;            // 2 bits per own slot
    uint32_t turn = 0;             // which of the own slots is preferred next (alternating hides the node-fetch latency)
    bool exhausted = false;

    while (true) {
        // ---- refill: empty slots claim the next rays of the queue (one atomic per warp) -------------------------------
        if (!exhausted) {
            uint32_t empties[K];
            uint32_t total = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                empties[k] = __ballot_sync(0xffffffffu, ((state >> (2 * k)) & 3u) == SLOT_EMPTY);
                total += (uint32_t)__popc(empties[k]);
            }
            const bool starving = __all_sync(0xffffffffu, state == 0u);
            if (total >= (uint32_t)refill_threshold || starving) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, total);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + total >= n) exhausted = true;                     // warp-uniform
                uint32_t before = 0;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (((state >> (2 * k)) & 3u) == SLOT_EMPTY) {
                        const uint32_t qi = base + before + (uint32_t)__popc(empties[k] & lt_mask);
                        RayD r; RayF f; float tmin_dn, tmax_up;
                        if (qi < n && io.load(qi, r, f, tmin_dn, tmax_up)) {
                            const uint32_t slot = lane + 32u * (uint32_t)k;
                            pool_d[CH(0, slot)] = make_double2(r.o.x, r.o.y);
                            pool_d[CH(1, slot)] = make_double2(r.o.z, r.d.x);
                            pool_d[CH(2, slot)] = make_double2(r.d.y, r.d.z);
                            pool[CH(3, slot)] = make_float4(f.ox, f.oy, f.oz, tmax_up);
                            pool[CH(4, slot)] = make_float4(f.idx, f.idy, f.idz, tmin_dn);
                            pool_d[CH(5, slot)] = make_double2(r.tmax, 1.0);
                            pool_d[CH(6, slot)] = make_double2(0.0, 0.0);
                            pool_u[CH(7, slot)] = make_uint4(0u, 0x80000000u, 0u, 0u);   // the root occupies virtual slot 7 of group 0
                            pool_u[CH(8, slot)] = make_uint4(0xffffffffu, qi, (f.octinv4 & 0xffu) << 8, 0u);
                            if (IO::kExactT) pool_d[CH(9, slot)] = make_double2(r.tmin, 0.0);
                            state |= SLOT_NODE << (2 * k);
                        }
                    }
                    before += (uint32_t)__popc(empties[k]);
                }
            }
        }
        // ---- vote the phase ------------------------------------------------------------------------------------------
        bool has_node = false, has_tri = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t st = (state >> (2 * k)) & 3u;
            has_node |= st == SLOT_NODE;
            has_tri |= st == SLOT_TRI;
        }
        const uint32_t nb = __ballot_sync(0xffffffffu, has_node), tb = __ballot_sync(0xffffffffu, has_tri);
        if ((nb | tb) == 0u) {
            if (exhausted) break;
            continue;
        }
        const bool node_phase = __popc(nb) >= __popc(tb);
        const uint32_t want = node_phase ? SLOT_NODE : SLOT_TRI;
        // own slot for this phase: the preferred one first
        int pick = -1;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int k = (int)((turn + (uint32_t)j) % (uint32_t)K);
            if (pick < 0 && ((state >> (2 * k)) & 3u) == want) pick = k;
        }
        if (pick >= 0) {
            turn = (uint32_t)(pick + 1) % (uint32_t)K;
            const uint32_t slot = lane + 32u * (uint32_t)pick;
            uint4 g = pool_u[CH(7, slot)];                   // ngroup.xy, tgroup.xy
            uint4 m = pool_u[CH(8, slot)];                   // htri, queue index, sp | octinv << 8
            uint32_t sp = m.z & 0xffu;
            uint32_t next_state;
            bool finished = false;
            if (node_phase) {
                const float4 c3 = pool[CH(3, slot)], c4 = pool[CH(4, slot)];
                uint2 ngroup = make_uint2(g.x, g.y);
                if (ngroup.y <= 0x00ffffffu && sp > 0u) {
                    --sp;
                    ngroup = stack_arena[sp * S + slot];
                }
                // (a slot is only in SLOT_NODE with an inner child pending or a non-empty stack)
                RayF f;
                f.ox = c3.x; f.oy = c3.y; f.oz = c3.z;
                f.idx = c4.x; f.idy = c4.y; f.idz = c4.z;
                f.aox = fabsf(c3.x); f.aoy = fabsf(c3.y); f.aoz = fabsf(c3.z);
                f.octinv4 = ((m.z >> 8) & 0xffu) * 0x01010101u;
                const uint32_t hits = ngroup.y;
                const uint32_t imask = hits & 0xffu;
                const int child_bit = bfind(hits);
                const uint32_t base = ngroup.x;
                ngroup.y &= ~(1u << child_bit);
                if (ngroup.y > 0x00ffffffu) {
                    if (sp >= (uint32_t)kStackSize) { atomicExch(overflow, 1u); sp = kStackSize - 1; }
                    stack_arena[sp * S + slot] = ngroup;
                    sp++;
                }
                const uint32_t cslot = (uint32_t)(child_bit - 24) ^ (f.octinv4 & 0xffu);
                const uint32_t ni = base + popc(imask & ~(0xffffffffu << cslot));
                const uint4 n0 = __ldg(nodes4 + ni * 5 + 0), n1 = __ldg(nodes4 + ni * 5 + 1),
                            n2 = __ldg(nodes4 + ni * 5 + 2), n3 = __ldg(nodes4 + ni * 5 + 3),
                            n4 = __ldg(nodes4 + ni * 5 + 4);
                if (COUNT) ctr->nodes++;
                const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, f, c4.w, c3.w);
                g.x = n1.x;
                g.y = (hm & 0xff000000u) | (n0.w >> 24);
                g.z = n1.y;
                g.w = hm & 0x00ffffffu;
            } else {
                // ---- one FP64 Möller–Trumbore, reference operation order (accel_sahbvh.cpp:58-75)
                const double2 c0 = pool_d[CH(0, slot)], c1 = pool_d[CH(1, slot)], c2 = pool_d[CH(2, slot)];
                double2 c5 = pool_d[CH(5, slot)];                        // ht, |det| of the candidate
                RayD r;
                r.o = d3(c0.x, c0.y, c1.x); r.d = d3(c1.y, c2.x, c2.y);
                const int ti = bfind(g.w);
                g.w &= ~(1u << ti);
                const uint32_t tri = g.z + (uint32_t)ti;
                D3 p1, p2, p3;
                load_tri(tris, tri, p1, p2, p3);
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
                    const double t = IO::kExactT ? divrn(ddot(e2, q), d) : mulrn(ddot(e2, q), rcp_newton(d));
                    const double tmin = IO::kExactT ? pool_d[CH(9, slot)].x : (double)pool[CH(4, slot)].w;
                    if (!(t < tmin || c5.x < t)) {                        // inclusive bounds; a later equal-t triangle wins
                        c5.x = t; c5.y = ad;
                        m.x = tri;
                        pool_d[CH(5, slot)] = c5;
                        pool_d[CH(6, slot)] = make_double2(u, v);
                        float4 c3 = pool[CH(3, slot)];
                        c3.w = f32_up(t);
                        pool[CH(3, slot)] = c3;
                        if (ANY) finished = true;
                    }
                }
            }
            // ---- where does this ray go next
            if (!finished) {
                if (g.w != 0u) next_state = SLOT_TRI;
                else if (g.y > 0x00ffffffu || sp > 0u) next_state = SLOT_NODE;
                else finished = true;
            }
            if (finished) {
                const double2 c5 = pool_d[CH(5, slot)], c6 = pool_d[CH(6, slot)];
                io.store(m.y, c5.x, c6.x, c6.y, c5.y, (int32_t)m.x);
                next_state = SLOT_EMPTY;
            } else {
                pool_u[CH(7, slot)] = g;
                m.z = (m.z & 0xffffff00u) | sp;
                pool_u[CH(8, slot)] = m;
            }
            state = (state & ~(3u << (2 * pick))) | (next_state << (2 * pick));
        }
    }
#undef CH
}

#endif  // __CUDACC__

}  // namespace lmcu
