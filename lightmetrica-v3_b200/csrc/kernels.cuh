// kernels.cuh — device-side scene view and the per-item bodies of the render kernels
// (raycast, ao, wavefront pt: generate / extend / shade / shadow).  The __global__ wrappers with
// persistent-thread work fetch and warp-ballot queue compaction live in lmcu.cu.
#pragma once
#include "bvh.cuh"
#include "build.cuh"
#include "shade.cuh"

namespace lmcu {

struct DevScene {
    const WideNode* nodes;
    const TriF32* t32;          // exactly one of t32 / t64 is non-null
    const TriF64* t64;
    const TriObj* tobj;         // two-level tree: object-space triangles of the bottom levels; t32 / t64 / shade / uv are null and
    const InstDev* inst;        // a hit's triangle id is the GLOBAL id tri_offset + face, resolved against prims + mesh arrays
    const double* mesh_p;       // two-level tree only: the uploaded mesh arrays (9 doubles per face, 6 floats of texture coordinates)
    const double* mesh_n;
    const float* mesh_t;
    const TriShade* shade;
    const PrimDev* prims;
    const MaterialDev* mats;
    const LightDev* lights;
    const EmitTri* emit;
    const float* cdf;
    const float4* env_tex;
    const TriUV* uv;            // per-triangle texture coordinates in leaf order; null when no material reads a texture
    const TextureDev* textures;
    const float4* texels;
    const MixEntry* mix;
    int32_t num_prims, num_lights, env_light;
    int32_t fetch_threshold;    // trace.cuh: idle lanes tolerated per iteration before the warp re-fetches
    uint32_t num_tris;
    int32_t pool_refill;        // trace2.cuh: empty pool slots per warp that trigger a re-fetch
    float bound_center[3], bound_radius;   // Light::set_scene_bound (light_envconst.cpp:37-43): sphere around the scene AABB, x 1.01
    CameraDev cam;
};

template <typename TriT> LMCU_HD const TriT* scene_tris(const DevScene& s);
template <> LMCU_HD const TriF32* scene_tris<TriF32>(const DevScene& s) { return s.t32; }
template <> LMCU_HD const TriF64* scene_tris<TriF64>(const DevScene& s) { return s.t64; }
template <> LMCU_HD const TriObj* scene_tris<TriObj>(const DevScene& s) { return s.tobj; }

// -- surface point at a hit (src/scene.cpp:333-344, src/mesh/mesh_raw.cpp:87-97), FP32 ------------------
struct SurfPoint {
    F3 p, n, gn;
    float tu, tv;      // texture coordinates (0,0 without a uv buffer)
    int32_t prim;      // primitive ordinal
    int32_t face;
    int32_t material;  // index into DevScene::mats
    int32_t light;     // index into DevScene::lights if the surface is an area light, else -1
};

LMCU_HD void tri_fetch(const DevScene& s, uint32_t tri, F3& p1, F3& p2, F3& p3, int32_t& prim, int32_t& face) {
    if (s.t32) {
        const float4 a = s.t32[tri].a, b = s.t32[tri].b, c = s.t32[tri].c;
        p1 = f3(a.x, a.y, a.z); p2 = f3(b.x, b.y, b.z); p3 = f3(c.x, c.y, c.z);
        prim = (int32_t)f2u(a.w); face = (int32_t)f2u(b.w);
    } else {
        const TriF64& t = s.t64[tri];
        p1 = f3((float)t.p[0], (float)t.p[1], (float)t.p[2]);
        p2 = f3((float)t.p[3], (float)t.p[4], (float)t.p[5]);
        p3 = f3((float)t.p[6], (float)t.p[7], (float)t.p[8]);
        prim = t.prim; face = t.face;
    }
}

// Two-level tree: the surface point of global triangle id `g`, from the instance's transform and the mesh arrays — the
// operations k_flatten / k_emit apply once per flattened triangle (build.cuh), applied at the hit instead.
LMCU_HD SurfPoint surface_point_inst(const DevScene& s, uint32_t g, float u, float v) {
    SurfPoint sp;
    const uint32_t pi = find_prim(s.prims, s.num_prims, g);
    const PrimDev& pr = s.prims[pi];
    const uint32_t face = g - pr.tri_offset;
    const size_t mf = (size_t)(pr.mesh_face_offset + face) * 9;
    const double* q = s.mesh_p + mf;
    const D3 w1 = transform_point(pr.M, d3(q[0], q[1], q[2])), w2 = transform_point(pr.M, d3(q[3], q[4], q[5])),
             w3 = transform_point(pr.M, d3(q[6], q[7], q[8]));
    const F3 p1 = f3((float)w1.x, (float)w1.y, (float)w1.z), p2 = f3((float)w2.x, (float)w2.y, (float)w2.z),
             p3 = f3((float)w3.x, (float)w3.y, (float)w3.z);
    sp.prim = (int32_t)pi; sp.face = (int32_t)face;
    const float w = 1.0f - u - v;
    sp.p = p1 * w + p2 * u + p3 * v;
    F3 n[3];
    shade_normals(pr, s.mesh_p, s.mesh_n, mf, n);
    sp.n = normalize(n[0] * w + n[1] * u + n[2] * v);
    sp.gn = normalize(cross(p2 - p1, p3 - p1)) * pr.det_sign;
    sp.material = pr.material; sp.light = pr.light;
    sp.tu = 0.f; sp.tv = 0.f;
    if (s.mesh_t) {
        const float* t = s.mesh_t + (size_t)(pr.mesh_face_offset + face) * 6;
        sp.tu = t[0] * w + t[2] * u + t[4] * v;
        sp.tv = t[1] * w + t[3] * u + t[5] * v;
    }
    return sp;
}

// GENERAL = false: the caller's instantiation never sees a two-level scene (the plain-material shade kernels and the one-level
// pixel loops: their register budget has no room for the second path — with it k_pt_shade<false> spills 92 instead of 68 bytes
// and C2 loses 0.8 %, profiles/r02_ab_pre_two_level.log)
template <bool GENERAL = true>
LMCU_HD SurfPoint surface_point(const DevScene& s, uint32_t tri, float u, float v) {
    if constexpr (GENERAL) {
        if (s.tobj) return surface_point_inst(s, tri, u, v);
    }
    SurfPoint sp;
    F3 p1, p2, p3;
    tri_fetch(s, tri, p1, p2, p3, sp.prim, sp.face);
    const float w = 1.0f - u - v;
    sp.p = p1 * w + p2 * u + p3 * v;
    const float4 n1 = s.shade[tri].n1, n2 = s.shade[tri].n2, n3 = s.shade[tri].n3;
    sp.n = normalize(f3(n1.x, n1.y, n1.z) * w + f3(n2.x, n2.y, n2.z) * u + f3(n3.x, n3.y, n3.z) * v);
    sp.gn = normalize(cross(p2 - p1, p3 - p1)) * n1.w;
    sp.material = (int32_t)f2u(n2.w); sp.light = (int32_t)f2u(n3.w);
    sp.tu = 0.f; sp.tv = 0.f;
    if (s.uv) {                                            // math::mix_barycentric of Mesh::Point::t (mesh_raw.cpp:94)
        const TriUV t = s.uv[tri];
        sp.tu = t.t1.x * w + t.t2.x * u + t.t3.x * v;
        sp.tv = t.t1.y * w + t.t2.y * u + t.t3.y * v;
    }
    return sp;
}

// (primitive ordinal, face) of a hit's triangle id
LMCU_HD void tri_ids(const DevScene& s, uint32_t tri, int32_t& prim, int32_t& face) {
    if (s.tobj) {
        prim = (int32_t)find_prim(s.prims, s.num_prims, tri);
        face = (int32_t)(tri - s.prims[prim].tri_offset);
    } else {
        F3 p1, p2, p3;
        tri_fetch(s, tri, p1, p2, p3, prim, face);
    }
}

// -- composite materials: mapKd textures, material::mask, mixture_wavefrontobj, constant-weight mixtures ---------------
// A plain lobe (diffuse | glossy | mirror | mask | glass) with its texture resolved at the surface point
LMCU_HD MaterialDev resolve_lobe(const DevScene& s, const MaterialDev& m, const SurfPoint& sp) {
    MaterialDev o = m;
    if (m.type == MAT_DIFFUSE && m.tex_kd >= 0) {          // material_diffuse.cpp:70,80,90: mapKd has priority over Kd
        const float4 c = tex_fetch(s.textures[m.tex_kd], s.texels, sp.tu, sp.tv);
        o.r = c.x; o.g = c.y; o.b = c.z;
    }
    return o;
}

// The component picked at a vertex (path::sample_component, include/lm/path.h:193-200) in a form eval / pdf / sample
// below can iterate: `n` lobes, each with a weight (multiplies its BSDF) and a selection probability.
struct VComp {
    int32_t kind;        // 0: the material itself (one plain lobe), 1: the diffuse + glossy pair of mixture_obj,
                         // 2: MixEntry range [first, first + n), 3: the pass-through component of mixture_obj
    int32_t first, n;
    bool marginal;       // a sampled direction is weighted by (sum of all lobes' f) / (marginal pdf)
    bool specular;
    float cweight;       // 1 / p(component), multiplies the throughput (renderer_pt.cpp:292-293)
    float alpha;         // mixture_obj: scales eval
    float wsel;          // mixture_obj: probability of sampling the diffuse lobe
    MaterialDev l0, l1;  // kind 0: l0; kind 1: diffuse (Kd resolved), glossy
};

LMCU_HD bool mat_is_plain(const MaterialDev& m) { return m.type <= MAT_MASK && m.tex_kd < 0; }

// Material::sample_component.  uc0 drives the choice (ComponentSampleU::uc[0]).
LMCU_HD VComp pick_component(const DevScene& s, const MaterialDev& mat, const SurfPoint& sp, float uc0) {
    VComp c;
    c.kind = 0; c.first = 0; c.n = 1; c.marginal = false; c.cweight = 1.f; c.alpha = 1.f; c.wsel = 1.f;
    if (mat.type <= MAT_MASK) {
        c.l0 = resolve_lobe(s, mat, sp);
        c.specular = mat_is_specular(mat.type);
    } else if (mat.type == MAT_MIXTURE_OBJ) {              // model_wavefrontobj.cpp:417-424
        F3 kd = f3(mat.r, mat.g, mat.b);
        float alpha = 1.f;
        if (mat.tex_kd >= 0) {
            const float4 t = tex_fetch(s.textures[mat.tex_kd], s.texels, sp.tu, sp.tv);
            kd = f3(t.x, t.y, t.z);
            if (mat.alpha_mask) alpha = t.w;
        }
        if (uc0 < alpha) {
            c.kind = 1; c.n = 2; c.marginal = true; c.specular = false;
            c.cweight = 1.0f / alpha; c.alpha = alpha;
            c.l0 = mat; c.l0.type = MAT_DIFFUSE; c.l0.r = kd.x; c.l0.g = kd.y; c.l0.b = kd.z;
            c.l1 = mat; c.l1.type = MAT_GLOSSY; c.l1.r = mat.ks_r; c.l1.g = mat.ks_g; c.l1.b = mat.ks_b;
            const float wd = max3(kd), wg = fmaxf(mat.ks_r, fmaxf(mat.ks_g, mat.ks_b));       // :431-439
            c.wsel = (wd == 0.f && wg == 0.f) ? 1.f : wd / (wd + wg);
        } else {
            c.kind = 3; c.n = 0; c.specular = true;
            c.cweight = 1.0f / (1.0f - alpha);
        }
    } else if (mat.type == MAT_MIXTURE_RR) {               // material_mixture.cpp:61-65: one entry is the component
        float acc = 0.f;
        int k = 0;
        for (; k < mat.num_entries - 1; k++) {
            acc += s.mix[mat.entry_offset + k].prob;
            if (uc0 < acc) break;
        }
        const MixEntry e = s.mix[mat.entry_offset + k];
        c.kind = 2; c.first = mat.entry_offset + k; c.n = 1;
        c.cweight = 1.0f / e.prob;
        c.specular = mat_is_specular(s.mats[e.material].type);
    } else {                                               // marginalized, :191-195: group 0 or one specular entry
        const int n0 = mat.num_group0;
        float acc = n0 > 0 ? s.mix[mat.entry_offset].group_prob : 0.f;
        if (n0 > 0 && (uc0 < acc || n0 == mat.num_entries)) {
            c.kind = 2; c.first = mat.entry_offset; c.n = n0; c.marginal = true; c.specular = false;
            c.cweight = 1.0f / s.mix[mat.entry_offset].group_prob;
        } else {
            int k = n0;
            for (; k < mat.num_entries - 1; k++) {
                acc += s.mix[mat.entry_offset + k].group_prob;
                if (uc0 < acc) break;
            }
            c.kind = 2; c.first = mat.entry_offset + k; c.n = 1; c.specular = true;
            c.cweight = 1.0f / s.mix[mat.entry_offset + k].group_prob;
        }
    }
    return c;
}

// lobe k of a component: the plain material, its weight and its selection probability
LMCU_HD MaterialDev comp_lobe(const DevScene& s, const VComp& c, const SurfPoint& sp, int k, float& w, float& sel) {
    if (c.kind == 2) {
        const MixEntry e = s.mix[c.first + k];
        w = e.weight; sel = c.marginal ? e.prob : 1.f;
        return resolve_lobe(s, s.mats[e.material], sp);
    }
    if (c.kind == 1) {
        w = c.alpha; sel = k == 0 ? c.wsel : 1.0f - c.wsel;
        return k == 0 ? c.l0 : c.l1;
    }
    w = 1.f; sel = 1.f;
    return c.l0;
}

// Material::eval(comp, eval_delta = true) and Material::pdf_direction(comp, eval_delta = true)
LMCU_HD F3 comp_eval(const DevScene& s, const VComp& c, const SurfPoint& sp, F3 wi, F3 wo) {
    if (c.kind == 0) return mat_eval(c.l0, sp.n, wi, wo);
    F3 f = f3s(0.f);
    for (int k = 0; k < c.n; k++) {
        float w, sel;
        const MaterialDev m = comp_lobe(s, c, sp, k, w, sel);
        f = f + mat_eval(m, sp.n, wi, wo) * w;
    }
    return f;
}
LMCU_HD float comp_pdf(const DevScene& s, const VComp& c, const SurfPoint& sp, F3 wi, F3 wo) {
    if (c.kind == 0) return mat_pdf(c.l0, sp.n, wi, wo);
    float p = 0.f;
    for (int k = 0; k < c.n; k++) {
        float w, sel;
        const MaterialDev m = comp_lobe(s, c, sp, k, w, sel);
        p += sel * mat_pdf(m, sp.n, wi, wo);
    }
    return p;
}

// Material::sample_direction(comp): ud0, ud1 = DirectionSampleU::ud, udc0 = udc[0] (lobe choice inside a marginal
// group, model_wavefrontobj.cpp:441-468, material_mixture.cpp:201-224; Fresnel choice of glass)
LMCU_HD BsdfSample comp_sample(const DevScene& s, const VComp& c, const SurfPoint& sp, F3 wi, float ud0, float ud1, float udc0, bool le = false) {
    if (c.kind == 0) return mat_sample(c.l0, sp.n, sp.gn, wi, ud0, ud1, udc0, le);
    BsdfSample bs;
    if (c.kind == 3) {                                     // model_wavefrontobj.cpp:470-479: weight is Vec3(1), not f/p
        bs.valid = true; bs.specular = true; bs.pdf = 0.f;
        bs.wo = -wi;
        bs.weight = f3s(sn_correction(sp.n, sp.gn, wi, bs.wo, le));
        return bs;
    }
    int k = 0;
    float w = 1.f, sel = 1.f, acc = 0.f;
    MaterialDev m = comp_lobe(s, c, sp, 0, w, sel);
    if (c.marginal) {
        for (; k < c.n - 1; k++) {
            acc += sel;
            if (udc0 < acc) break;
            m = comp_lobe(s, c, sp, k + 1, w, sel);
        }
    }
    bs = mat_sample(m, sp.n, sp.gn, wi, ud0, ud1, udc0, le);
    if (!bs.valid) return bs;
    if (c.marginal) {
        const F3 f = comp_eval(s, c, sp, wi, bs.wo);
        const float p = comp_pdf(s, c, sp, wi, bs.wo);
        if (!(p > 0.f)) { bs.valid = false; return bs; }   // the reference divides 0 by 0 here; the path carries nothing
        bs.weight = f * (sn_correction(sp.n, sp.gn, wi, bs.wo, le) / p);
        bs.pdf = p;
    } else {
        bs.weight = bs.weight * w;
    }
    return bs;
}

// Material::reflectance (raycast shading): mixture_obj shows its diffuse lobe (model_wavefrontobj.cpp:482-485), the
// constant-weight mixtures the weighted sum (material_mixture.cpp:80-86,226-234)
LMCU_HD F3 surface_reflectance(const DevScene& s, const MaterialDev& mat, const SurfPoint& sp) {
    if (mat.type <= MAT_MASK) return mat_reflectance(resolve_lobe(s, mat, sp));
    if (mat.type == MAT_MIXTURE_OBJ) {
        if (mat.tex_kd < 0) return f3(mat.r, mat.g, mat.b);
        const float4 t = tex_fetch(s.textures[mat.tex_kd], s.texels, sp.tu, sp.tv);
        return f3(t.x, t.y, t.z);
    }
    F3 sum = f3s(0.f);
    for (int k = 0; k < mat.num_entries; k++) {
        const MixEntry e = s.mix[mat.entry_offset + k];
        sum = sum + mat_reflectance(resolve_lobe(s, s.mats[e.material], sp)) * e.weight;
    }
    return sum;
}

// Radiance of the environment light seen along direction d (scene -> sky): envconst Le or the image lookup
LMCU_HD F3 light_env_radiance(const DevScene& s, const LightDev& l, F3 d) {
    return l.type == LIGHT_ENV ? env_eval(l, s.env_tex, d) : f3(l.Ke[0], l.Ke[1], l.Ke[2]);
}
// pdf_direct (projected solid angle, without the light-selection probability) of sampling direction d from a
// surface point with shading normal n
// (degenerated = the point is a pinhole camera: no cosine, surface.h:377-390)
LMCU_HD float light_env_pdf(const DevScene& s, const LightDev& l, F3 n, F3 d, bool degenerated = false) {
    if (l.type == LIGHT_ENV) return env_pdf_direct(l, s.cdf, n, d, degenerated);
    const float c = degenerated ? 1.0f : fabsf(dot(n, d));
    return c == 0.f ? 0.f : (1.0f / (4.0f * kPiF)) / c;             // light_envconst.cpp:116-119
}

// -- raycast (src/renderer/renderer_raycast.cpp:70-98) -----------------------------------------------------
struct RaycastParams {
    int32_t width, height;
    int32_t use_constant_color, visualize_normal, has_color, pad_;
    float bg[3], color[3];
};

template <typename TriT>
LMCU_HD void raycast_pixel(const DevScene& s, const RaycastParams& p, int x, int y, float* film, uint2* stack) {
    RayD r;
    camera_primary_ray_d(s.cam, divrn(addrn((double)x, 0.5), (double)p.width), divrn(addrn((double)y, 0.5), (double)p.height), r.o, r.d);
    r.tmin = kEps; r.tmax = kInf;
    HitD h; h.t = r.tmax; h.tri = -1; h.u = h.v = 0;
    traverse<false, false, TriT>(s.nodes, scene_tris<TriT>(s), r, h, stack, nullptr, s.inst);
    F3 c;
    if (h.tri < 0) {
        if (s.env_light >= 0) {
            // miss with tmax = Inf and an environment light -> light endpoint (src/scene.cpp:321-331)
            const LightDev& l = s.lights[s.env_light];
            const F3 dir = f3((float)r.d.x, (float)r.d.y, (float)r.d.z);
            c = p.visualize_normal ? f3s(0.f) : light_env_radiance(s, l, dir);
        } else {
            c = f3(p.bg[0], p.bg[1], p.bg[2]);
        }
    } else {
        const SurfPoint sp = surface_point<TriT::kTwoLevel>(s, (uint32_t)h.tri, (float)h.u, (float)h.v);
        if (p.visualize_normal) {
            c = f3(fabsf(sp.n.x), fabsf(sp.n.y), fabsf(sp.n.z));
        } else {
            c = p.has_color ? f3(p.color[0], p.color[1], p.color[2]) : surface_reflectance(s, s.mats[sp.material], sp);
            if (!p.use_constant_color) {
                const F3 d = f3((float)r.d.x, (float)r.d.y, (float)r.d.z);
                c = c * (0.2f + 0.8f * fabsf(dot(sp.n, -d)));
            }
        }
    }
    float* px = film + ((size_t)y * p.width + x) * 3;
    px[0] = c.x; px[1] = c.y; px[2] = c.z;
}

// -- ambient occlusion (functest/renderer_ao.cpp:26-43) ----------------------------------------------------
template <typename TriT>
LMCU_HD void ao_pixel(const DevScene& s, int width, int height, int x, int y, int64_t spp, uint64_t sample_offset,
                      uint32_t seed, double ao_distance, float* film, uint2* stack, uint32_t* ray_count) {
    RayD r;
    camera_primary_ray_d(s.cam, divrn(addrn((double)x, 0.5), (double)width), divrn(addrn((double)y, 0.5), (double)height), r.o, r.d);
    r.tmin = kEps; r.tmax = kInf;
    HitD h; h.t = r.tmax; h.tri = -1; h.u = h.v = 0;
    traverse<false, false, TriT>(s.nodes, scene_tris<TriT>(s), r, h, stack, nullptr, s.inst);
    *ray_count += 1;
    if (h.tri < 0) return;   // the reference leaves the pixel untouched on a miss
    const SurfPoint sp = surface_point<TriT::kTwoLevel>(s, (uint32_t)h.tri, (float)h.u, (float)h.v);
    const F3 wi = -f3((float)r.d.x, (float)r.d.y, (float)r.d.z);
    const Frame f = frame_twosided(sp.n, wi);
    float V = 0.f;
    const uint32_t pixel = (uint32_t)(y * width + x);
    for (int64_t i = 0; i < spp; i++) {
        const R4 u = rand4(seed, pixel, sample_offset + (uint64_t)i, 0);
        const F3 d = sample_cosine_weighted(u.x, u.y);
        const F3 wo = f.u * d.x + f.v * d.y + f.n * d.z;
        RayD q;
        q.o = d3(sp.p.x, sp.p.y, sp.p.z); q.d = d3(wo.x, wo.y, wo.z);
        q.tmin = kEps; q.tmax = ao_distance;
        HitD hq; hq.t = q.tmax; hq.tri = -1; hq.u = hq.v = 0;
        traverse<true, false, TriT>(s.nodes, scene_tris<TriT>(s), q, hq, stack, nullptr, s.inst);
        V += hq.tri < 0 ? 1.f : 0.f;
    }
    *ray_count += (uint32_t)spp;
    V /= (float)spp;
    float* px = film + ((size_t)y * width + x) * 3;
    px[0] = V; px[1] = V; px[2] = V;
}

// -- wavefront path tracer (src/renderer/renderer_pt.cpp:89-302, pixel mode) --------------------------------
// Path state, SoA, one entry per queue slot.
struct PathQueue {
    float4* ray_o;      // o.xyz, tmin
    float4* ray_d;      // d.xyz, tmax
    float4* thr_pix;    // throughput.rgb, pixel (uint bits)
    float4* aux;        // pdf_bsdf (projected SA) of the sampled direction, shading normal at the previous vertex
    uint2* info;        // x: low 32 bits of the path's sample index, y: num_verts | prev_nee<<8 | sample index bits 32..47 << 16
};

struct ShadowQueue {
    float4* ray_o;      // o.xyz, tmin
    float4* ray_d;      // d.xyz, tmax
    float4* contrib;    // C.rgb, pixel (uint bits)
};

struct PtParams {
    int32_t width, height;
    int32_t max_verts, sampling_mode;
    uint32_t seed, tiled;
    uint64_t sample_offset;
    uint64_t total_paths;
    uint64_t path_base;         // index of this wavefront's first path inside the pass (staggered wavefronts)
    uint32_t capacity;
    uint32_t image_mode;        // primary_ray_sampling_mode = image: the raster window is the whole film (renderer_pt.cpp:104-106)
    uint32_t light_tracing;     // renderer::lt: paths start on the lights and connect to the camera (renderer_lt.cpp)
    uint32_t chunk;             // pixel mode: consecutive path indices walk `chunk` samples of one pixel before the next pixel
    uint32_t bdpt_mode;         // bdpt: LMCU_BDPT_* (lmcu.h) — all strategies with MIS weights, or one unweighted strategy family
};

// RNG counter word 0 of a path's primary sample in image mode (no pixel is known before the raster position is drawn)
constexpr uint32_t kImagePrimaryKey = 0xffffffffu;
// ... and of every draw of a light-tracing path (its samples are not tied to a pixel either)
constexpr uint32_t kLightPathKey = 0xfffffffeu;

struct PtCounters {
    uint32_t count[2];          // entries in path queue A / B
    uint32_t shadow_count;
    uint32_t fetch_extend;      // persistent-thread work fetch cursors
    uint32_t fetch_shadow;
    uint32_t overflow;          // traversal stack overflow flag
    uint32_t gen_slot, gen_count;   // this iteration's refill: first free queue slot, number of new paths (k_pt_advance -> k_pt_generate)
    unsigned long long gen_first;   // ... and the index of the first new path
    unsigned long long next_path;
    unsigned long long rays;
    unsigned long long iterations;
    unsigned long long extend_rays, shadow_rays;
    unsigned long long nodes_extend, tris_extend, nodes_shadow, tris_shadow;   // counters build only
};

// Path index -> (pixel, sample index of the pass).  The reference's index space is sample-major (scheduler.cpp:42-44: pixel =
// index / spp).  Here a pass is walked in blocks of `chunk` samples: path g is sample (g mod chunk) of block g / (npix * chunk)
// at pixel (g / chunk) mod npix — every (pixel, sample) pair exactly once, as before (the RNG is keyed by the pair, so the
// image does not depend on the order).  Why: the paths in flight at any time (16 Mi per wavefront) then cover 1/chunk as
// many pixels, so the film lines their atomics hit stay in L2 instead of sweeping the whole film every iteration (4K film:
// 99.5 MB), and a warp's 32 camera rays belong to 2 pixels instead of 32.
LMCU_HD void path_pixel(const PtParams& p, uint64_t g, uint32_t& pixel, uint32_t& sample, int& x, int& y) {
    const uint64_t npix = (uint64_t)p.width * p.height;
    const uint32_t c = p.chunk ? p.chunk : 1u;
    const uint64_t per_block = npix * c;
    const uint64_t block = g / per_block, r = g % per_block;
    sample = (uint32_t)(block * c + (r % c));
    const uint32_t lin = (uint32_t)(r / c);
    if (p.tiled) {                       // 8x4 pixel tiles: one warp = one tile of coherent primary rays
        const uint32_t tiles_x = (uint32_t)p.width / 8u;
        const uint32_t tile = lin >> 5, idx = lin & 31u;
        x = (int)((tile % tiles_x) * 8u + (idx & 7u));
        y = (int)((tile / tiles_x) * 4u + (idx >> 3));
    } else {
        x = (int)(lin % (uint32_t)p.width);
        y = (int)(lin / (uint32_t)p.width);
    }
    pixel = (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
}

// Geometry term between a surface vertex (normal n1) and a light point (normal n2): surface.h:329-348
// (degenerated1: point 1 is a pinhole camera / point light, its cosine is 1)
LMCU_HD float geometry_term(F3 p1, F3 n1, F3 p2, F3 n2, bool degenerated1 = false) {
    F3 d = p2 - p1;
    const float L2 = dot(d, d);
    d = d * (1.0f / sqrtf(L2));
    return (degenerated1 ? 1.0f : fabsf(dot(n1, d))) * fabsf(dot(n2, -d)) / L2;
}

// path::sample_direct toward a light from point P (include/lm/path.h:493-530, src/scene.cpp:374-387 and the lights'
// sample_direct): direction wo toward the light, radiance Le, pdf pl in projected solid angle at P WITHOUT the
// light-selection probability, shadow-ray length tmax.  `degenerated`: P is the pinhole camera (no cosine at P).
struct LightSample {
    F3 wo, Le;
    float pl, tmax;
    bool delta;          // point / directional: not reachable by BSDF sampling => NEE only, MIS weight 1 (renderer_pt.cpp:176-180)
};

// The sample is drawn in two halves.  pick_light() does everything that does not depend on the shading point — the light,
// the emitter triangle and the point on it, or the sky direction — i.e. every dependent LOAD of the sample (light record ->
// CDF search -> emitter triangle / environment CDFs and texel).  connect_light() then forms direction, distance and pdf at a
// point P.  (Measured and dropped, round 2: calling pick_light() BEFORE the surface point is gathered, so that the two chains
// of dependent loads overlap — shade kernel 0.156 -> 0.163 s per 512-spp C2 render, 1892 -> 1869 Mpaths/s: the picks thrown
// away at terminated and specular vertices and 80 more bytes of spills cost more than the overlap buys.)
struct LightPick {
    F3 Le;               // radiance toward the scene (area: Ke; env: looked up along wo)
    F3 x;                // area / point: the sampled point; directional / env: the direction TOWARD the light
    F3 n;                // area: emitter normal
    float inv_area_j;    // area
    int32_t type;
};

LMCU_HD LightPick pick_light(const DevScene& s, const PtParams& p, uint32_t key, uint64_t sample_index, uint32_t block) {
    const R4 u1 = rand4(p.seed, key, sample_index, block);   // sel, upc, up0, up1
    int li = (int)(u1.x * (float)s.num_lights);
    li = li < 0 ? 0 : (li > s.num_lights - 1 ? s.num_lights - 1 : li);  // src/scene.cpp:374-387
    const LightDev& l = s.lights[li];
    LightPick k;
    k.type = l.type;
    k.Le = f3(l.Ke[0], l.Ke[1], l.Ke[2]);
    k.n = f3s(0.f); k.inv_area_j = 0.f;
    if (l.type == LIGHT_AREA) {                                         // light_area.cpp:47-59
        const int ti = cdf_sample(s.cdf + l.cdf_offset, l.num_tris, u1.y);
        const EmitTri& e = s.emit[l.tri_offset + ti];
        const float sq = safe_sqrt(u1.z);
        const float b1 = 1.0f - sq, b2 = u1.w * sq;
        const F3 a = f3(e.p1.x, e.p1.y, e.p1.z), b = f3(e.p2.x, e.p2.y, e.p2.z), c = f3(e.p3.x, e.p3.y, e.p3.z);
        k.x = a * (1.0f - b1 - b2) + b * b1 + c * b2;
        k.n = f3(e.n.x, e.n.y, e.n.z);
        k.inv_area_j = l.inv_area_j;
    } else if (l.type == LIGHT_POINT) {
        k.x = f3(l.pos[0], l.pos[1], l.pos[2]);
    } else if (l.type == LIGHT_DIRECTIONAL) {
        k.x = -f3(l.pos[0], l.pos[1], l.pos[2]);
    } else {                                                            // light_envconst.cpp:100-119, light_env.cpp:139-162
        const R4 u2 = rand4(p.seed, key, sample_index, block + 1u);
        k.x = l.type == LIGHT_ENV ? env_sample_dir(l, s.cdf, u2) : -sample_uniform_sphere(u2.x, u2.y);
        k.Le = light_env_radiance(s, l, k.x);
        k.n = f3((float)li, 0.f, 0.f);                                  // env: the pdf needs the light record again
    }
    return k;
}

LMCU_HD LightSample connect_light(const DevScene& s, const LightPick& k, F3 P, F3 N, bool degenerated) {
    LightSample o;
    o.Le = k.Le;
    o.pl = 0.f; o.delta = false; o.tmax = 0.f;
    if (k.type == LIGHT_AREA) {                                         // light_area.cpp:148-178
        const F3 dv = k.x - P;
        const float dist = sqrtf(dot(dv, dv));
        o.wo = dv * (1.0f / dist);
        if (dot(-o.wo, k.n) > 0.f) {
            const float G = geometry_term(P, N, k.x, k.n, degenerated);
            o.pl = G == 0.f ? 0.f : k.inv_area_j / G;
        }
        o.tmax = dist * (1.0f - (float)kEps);                           // include/lm/scene.h:349-356
    } else if (k.type == LIGHT_POINT) {                                 // light_point.cpp:96-114, surface.h:329-348
        const F3 dv = k.x - P;
        const float d2 = dot(dv, dv), dist = sqrtf(d2);
        o.wo = dv * (1.0f / dist);
        const float c = degenerated ? 1.0f : fabsf(dot(N, o.wo));       // G = |n.d| / r^2 toward a degenerated point
        o.pl = c == 0.f ? 0.f : d2 / c;
        o.tmax = dist * (1.0f - (float)kEps);
        o.delta = true;
    } else if (k.type == LIGHT_DIRECTIONAL) {                           // light_directional.cpp:93-108
        o.wo = k.x;
        const float c = degenerated ? 1.0f : fabsf(dot(N, o.wo));
        o.pl = c == 0.f ? 0.f : 1.0f / c;
        o.tmax = (float)(kInf - 1.0);
        o.delta = true;
    } else {
        o.wo = k.x;
        o.pl = light_env_pdf(s, s.lights[(int)k.n.x], N, o.wo, degenerated);
        o.tmax = (float)(kInf - 1.0);
    }
    return o;
}

LMCU_HD LightSample sample_light(const DevScene& s, const PtParams& p, uint32_t key, uint64_t sample_index, uint32_t block,
                                 F3 P, F3 N, bool degenerated) {
    return connect_light(s, pick_light(s, p, key, sample_index, block), P, N, degenerated);
}

// Film::raster_to_pixel (include/lm/film.h:166-171)
LMCU_HD uint32_t raster_to_pixel(const PtParams& p, float rpx, float rpy) {
    int x = (int)(rpx * (float)p.width), y = (int)(rpy * (float)p.height);
    x = x < 0 ? 0 : (x > p.width - 1 ? p.width - 1 : x);
    y = y < 0 ? 0 : (y > p.height - 1 ? p.height - 1 : y);
    return (uint32_t)y * (uint32_t)p.width + (uint32_t)x;
}

struct GenOut {
    bool shadow;            // image mode: NEE from the camera vertex produced a shadow ray
    float4 sh_o, sh_d, sh_c;
};

// First path vertex: camera endpoint + direction through the sample window (renderer_pt.cpp:96-113,197-206).  In image
// mode the camera vertex is itself samplable by NEE (renderer_pt.cpp:126-140): a light point is connected straight
// to the pinhole and splatted at ITS raster position (:151-157), and the primary direction carries the camera's
// pdf_direction J(wo) so that a direct hit on an emitter is MIS-weighted against that connection (:248-263).
LMCU_HD GenOut pt_generate_body(const DevScene& s, const PtParams& p, uint64_t g, const PathQueue& q, uint32_t slot) {
    GenOut out;
    out.shadow = false;
    uint32_t pixel, sample, sample_hi = 0; int x, y;
    float rpx, rpy;
    if (p.image_mode) {
        // scheduler::spi (src/scheduler.cpp:131-141): sample g of the whole image; the raster position is uniform in
        // [0,1)^2 and the pixel follows from it
        const R4 u = rand4(p.seed, kImagePrimaryKey, p.sample_offset + g, 0);
        rpx = u.x; rpy = u.y;
        pixel = raster_to_pixel(p, rpx, rpy);
        sample = (uint32_t)g; sample_hi = (uint32_t)(g >> 32) & 0xffffu;
    } else {
        path_pixel(p, g, pixel, sample, x, y);
        const R4 u = rand4(p.seed, pixel, p.sample_offset + sample, 0);
        rpx = ((float)x + u.x) / (float)p.width; rpy = ((float)y + u.y) / (float)p.height;
    }
    const float rx = 2.0f * rpx - 1.0f, ry = 2.0f * rpy - 1.0f;
    const CameraDev& c = s.cam;
    const F3 dl = normalize(f3((float)(c.aspect * c.tf) * rx, (float)c.tf * ry, -1.0f));
    const F3 d = f3((float)c.u[0], (float)c.u[1], (float)c.u[2]) * dl.x + f3((float)c.v[0], (float)c.v[1], (float)c.v[2]) * dl.y
               + f3((float)c.w[0], (float)c.w[1], (float)c.w[2]) * dl.z;
    const F3 cpos = f3((float)c.pos[0], (float)c.pos[1], (float)c.pos[2]);
    const bool camera_nee = p.image_mode && p.sampling_mode != 0 /*naive*/;
    float pdf_dir = 0.f;
    if (camera_nee) {
        float tx, ty;
        pdf_dir = camera_importance(c, d, tx, ty);
        if (s.num_lights > 0) {
            const uint64_t sidx = p.sample_offset + g;
            const LightSample ls = sample_light(s, p, kImagePrimaryKey, sidx, 1u, cpos, f3s(0.f), true);
            float lx, ly;
            const float We = ls.pl > 0.f ? camera_importance(c, ls.wo, lx, ly) : 0.f;     // Camera::eval (camera_pinhole.cpp:233-238)
            if (We > 0.f) {
                float mis = 1.f;
                if (p.sampling_mode == 2 && !ls.delta) {
                    const float p_light = ls.pl / (float)s.num_lights;
                    mis = p_light / (p_light + We);                                         // pdf_direction of the camera = J = We
                }
                const F3 C = ls.Le * (We * mis * (float)s.num_lights / ls.pl);
                if (!is_zero(C)) {
                    out.shadow = true;
                    out.sh_o = make_float4(cpos.x, cpos.y, cpos.z, (float)kEps);
                    out.sh_d = make_float4(ls.wo.x, ls.wo.y, ls.wo.z, ls.tmax);
                    out.sh_c = make_float4(C.x, C.y, C.z, u2f(raster_to_pixel(p, lx, ly)));
                }
            }
        }
    }
    q.ray_o[slot] = make_float4(cpos.x, cpos.y, cpos.z, (float)kEps);
    q.ray_d[slot] = make_float4(d.x, d.y, d.z, (float)kInf);
    q.thr_pix[slot] = make_float4(1.f, 1.f, 1.f, u2f(pixel));
    q.aux[slot] = make_float4(pdf_dir, 0.f, 0.f, 0.f);
    // num_verts = 1; prev_nee + prev_degenerated (bits 8, 9) when the camera vertex was sampled by NEE
    q.info[slot] = make_uint2(sample, 1u | (camera_nee ? 0x300u : 0u) | (sample_hi << 16));
    return out;
}

template <bool ANY, bool COUNT, typename TriT>
LMCU_HD bool pt_trace(const DevScene& s, float4 o, float4 d, HitD& h, uint2* stack, TraverseCounters* tc) {
    RayD r;
    r.o = d3(o.x, o.y, o.z); r.d = d3(d.x, d.y, d.z);
    r.tmin = o.w; r.tmax = d.w;
    h.t = r.tmax; h.tri = -1; h.u = h.v = 0;
    return traverse<ANY, COUNT, TriT>(s.nodes, scene_tris<TriT>(s), r, h, stack, tc, s.inst);
}

struct ShadeOut {
    bool alive;             // path continues: next ray in (o, d), state below
    bool shadow;            // a shadow ray was generated
    float4 o, d, thr_pix, aux;
    uint2 info;
    float4 sh_o, sh_d, sh_c;
    F3 emit;                // direct-hit contribution to add to the film now
};

// One path vertex: everything the reference does between Scene::intersect returning `hit` and the next
// Scene::intersect (renderer_pt.cpp:229-300 for the hit, then :127-209 of the next loop iteration).
// COMPOSITE = the scene holds a textured or mixture material; the <false> instantiation (every BASELINE config) carries
// none of that code or its registers.
// INST: surface points may have to be resolved through instances (two-level layout); the composite instantiation always can
template <bool COMPOSITE, bool INST = COMPOSITE>
LMCU_HD ShadeOut pt_shade_body(const DevScene& s, const PtParams& p, float4 ro, float4 rd, float4 thr_pix, float4 aux,
                               uint2 info, float4 hit) {
    ShadeOut out;
    out.alive = false; out.shadow = false; out.emit = f3s(0.f);
    const uint32_t pixel = f2u(thr_pix.w);
    const uint64_t sample = (uint64_t)info.x | ((uint64_t)(info.y >> 16) << 32);
    uint32_t num_verts = info.y & 0xffu;
    const bool prev_nee = (info.y >> 8) & 1u;
    const bool prev_degen = (info.y >> 9) & 1u;       // the previous vertex is the pinhole camera (image mode)
    F3 thr = f3(thr_pix.x, thr_pix.y, thr_pix.z);
    const F3 rdir = f3(rd.x, rd.y, rd.z);
    const F3 wi = -rdir;
    const int32_t tri = (int32_t)f2u(hit.w);
    const float inv_nl = 1.0f / (float)(s.num_lights > 0 ? s.num_lights : 1);
    const bool direct_ok = p.sampling_mode != 1 /*NEE*/ || !prev_nee;       // renderer_pt.cpp:236-246

    if (tri < 0) {
        // miss: environment light endpoint if any (src/scene.cpp:321-331), then the walk ends (:276-278)
        if (s.env_light >= 0 && direct_ok) {
            const LightDev& l = s.lights[s.env_light];
            float mis = 1.f;
            if (p.sampling_mode == 2 && prev_nee) {
                const F3 np = f3(aux.y, aux.z, aux.w);
                const float pl = inv_nl * light_env_pdf(s, l, np, rdir, prev_degen);
                mis = (aux.x == 0.f && pl == 0.f) ? 0.f : aux.x / (aux.x + pl);
            }
            out.emit = thr * light_env_radiance(s, l, rdir) * mis;
        }
        return out;
    }

    const SurfPoint sp = surface_point<INST>(s, (uint32_t)tri, hit.y, hit.z);
    const MaterialDev& mat = s.mats[sp.material];

    // direct hit on an emitter (renderer_pt.cpp:248-270)
    if (sp.light >= 0 && direct_ok) {
        const LightDev& l = s.lights[sp.light];
        if (dot(wi, sp.n) > 0.f) {                                          // light_area.cpp:185-187
            float mis = 1.f;
            if (p.sampling_mode == 2 && prev_nee) {
                const F3 po = f3(ro.x, ro.y, ro.z), np = f3(aux.y, aux.z, aux.w);
                const float G = geometry_term(po, np, sp.p, sp.n, prev_degen);
                const float pl = G == 0.f ? 0.f : inv_nl * l.inv_area_j / G;   // light_area.cpp:172-178
                mis = (aux.x == 0.f && pl == 0.f) ? 0.f : aux.x / (aux.x + pl);
            }
            out.emit = thr * f3(l.Ke[0], l.Ke[1], l.Ke[2]) * mis;
        }
    }

    const R4 u0 = rand4(p.seed, pixel, p.sample_offset + sample, num_verts * 4u + 0u);   // ud0, ud1, rr, uc0
    // Russian roulette (renderer_pt.cpp:281-287)
    if (num_verts > 5u) {
        const float q = fmaxf(0.2f, 1.0f - max3(thr));
        if (u0.z < q) return out;
        thr = thr * (1.0f / (1.0f - q));
    }
    num_verts++;
    if ((int)num_verts >= p.max_verts) return out;                          // loop bound, renderer_pt.cpp:123

    // component of the material at this vertex (renderer_pt.cpp:292-293; its weight multiplies the throughput).  Plain
    // single-lobe materials keep the direct calls; everything composite goes through VComp.
    const bool plain = !COMPOSITE || mat_is_plain(mat);
    VComp vc;
    float udc0 = u0.w;
    if (!plain) {
        vc = pick_component(s, mat, sp, u0.w);
        thr = thr * vc.cweight;
        udc0 = rand4(p.seed, pixel, p.sample_offset + sample, num_verts * 4u + 3u).x;
    }
    const bool specular = plain ? mat_is_specular(mat.type) : vc.specular;
    const bool nee = p.sampling_mode != 0 /*naive*/ && !specular;           // renderer_pt.cpp:127-140 (num_verts > 1 here)

    // next-event estimation (renderer_pt.cpp:142-192)
    if (nee && s.num_lights > 0) {
        const LightSample ls = sample_light(s, p, pixel, p.sample_offset + sample, num_verts * 4u + 1u, sp.p, sp.n, false);
        const F3 wo = ls.wo, Le = ls.Le;          // wo: direction from the vertex toward the light
        const float pl = ls.pl, tmax = ls.tmax;   // pdf_direct, projected SA, without the selection probability
        const bool delta_light = ls.delta;
        if (pl > 0.f) {
            F3 fs;
            float p_bsdf = 0.f;
            if (plain) fs = mat_eval(mat, sp.n, wi, wo); else { fs = comp_eval(s, vc, sp, wi, wo); p_bsdf = comp_pdf(s, vc, sp, wi, wo); }
            fs = fs * sn_correction(sp.n, sp.gn, wi, wo);
            if (!is_zero(fs)) {
                float mis = 1.f;
                if (p.sampling_mode == 2 && !delta_light) {
                    const float p_light = pl * inv_nl;
                    if (plain) p_bsdf = mat_pdf(mat, sp.n, wi, wo);
                    mis = (p_light == 0.f && p_bsdf == 0.f) ? 0.f : p_light / (p_light + p_bsdf);   // math.h:697-702
                }
                const F3 C = thr * fs * Le * (mis / (pl * inv_nl));
                out.shadow = true;
                out.sh_o = make_float4(sp.p.x, sp.p.y, sp.p.z, (float)kEps);
                out.sh_d = make_float4(wo.x, wo.y, wo.z, tmax);
                out.sh_c = make_float4(C.x, C.y, C.z, u2f(pixel));
            }
        }
    }

    // direction sampling (renderer_pt.cpp:197-209) with the component picked at the hit (:292-293)
    BsdfSample bs;
    if (plain) bs = mat_sample(mat, sp.n, sp.gn, wi, u0.x, u0.y, u0.w); else bs = comp_sample(s, vc, sp, wi, u0.x, u0.y, udc0);
    if (!bs.valid) return out;
    thr = thr * bs.weight;
    if (is_zero(thr)) return out;          // every later contribution is multiplied by zero
    out.alive = true;
    out.o = make_float4(sp.p.x, sp.p.y, sp.p.z, (float)kEps);
    out.d = make_float4(bs.wo.x, bs.wo.y, bs.wo.z, (float)kInf);
    out.thr_pix = make_float4(thr.x, thr.y, thr.z, u2f(pixel));
    out.aux = make_float4(bs.pdf, sp.n.x, sp.n.y, sp.n.z);
    out.info = make_uint2(info.x, num_verts | ((nee ? 1u : 0u) << 8) | (info.y & 0xffff0000u));
    return out;
}

// -- light tracer (src/renderer/renderer_lt.cpp:56-167): paths start on a light, every surface vertex is connected to the
// pinhole and splatted at the raster position of that connection ----------------------------------------------------------

// path::sample_primary_ray(TransDir::LE) (include/lm/path.h:101-118): pick a light uniformly, Light::sample_ray.
// RaySampleU = { up(2), upc(2), ud(2) }: block 0 = (upc0, upc1, up0, up1), block 1 = (ud0, ud1, -, -).
LMCU_HD void lt_generate_body(const DevScene& s, const PtParams& p, uint64_t g, const PathQueue& q, uint32_t slot) {
    const uint64_t sidx = p.sample_offset + g;
    const R4 u1 = rand4(p.seed, kLightPathKey, sidx, 0);
    const R4 u2 = rand4(p.seed, kLightPathKey, sidx, 1);
    int li = (int)(u1.x * (float)s.num_lights);
    li = li < 0 ? 0 : (li > s.num_lights - 1 ? s.num_lights - 1 : li);          // src/scene.cpp:374-387
    const LightDev& l = s.lights[li];
    const float inv_psel = (float)s.num_lights;
    F3 o, d, w;
    const F3 Ke = f3(l.Ke[0], l.Ke[1], l.Ke[2]);
    const F3 bc = f3(s.bound_center[0], s.bound_center[1], s.bound_center[2]);
    const float R = s.bound_radius;
    if (l.type == LIGHT_AREA) {                                                  // light_area.cpp:77-97
        const int ti = cdf_sample(s.cdf + l.cdf_offset, l.num_tris, u1.y);
        const EmitTri& e = s.emit[l.tri_offset + ti];
        const float sq = safe_sqrt(u1.z);
        const float b1 = 1.0f - sq, b2 = u1.w * sq;
        const F3 a = f3(e.p1.x, e.p1.y, e.p1.z), b = f3(e.p2.x, e.p2.y, e.p2.z), c = f3(e.p3.x, e.p3.y, e.p3.z);
        o = a * (1.0f - b1 - b2) + b * b1 + c * b2;
        const F3 nl = f3(e.n.x, e.n.y, e.n.z);
        const F3 dl = sample_cosine_weighted(u2.x, u2.y);
        F3 tu, tv;
        orthonormal_basis(nl, tu, tv);
        d = tu * dl.x + tv * dl.y + nl * dl.z;
        // Le / (pA * pD): pA = invA / J, pD = 1/pi in projected solid angle; Le = Ke on the emitting side only
        w = dot(d, nl) > 0.f ? Ke * (kPiF / l.inv_area_j) : f3s(0.f);
    } else if (l.type == LIGHT_POINT) {                                          // light_point.cpp:36-50
        o = f3(l.pos[0], l.pos[1], l.pos[2]);
        d = sample_uniform_sphere(u2.x, u2.y);
        w = Ke * (4.0f * kPiF);
    } else {
        // directional (light_directional.cpp:50-68), envconst (light_envconst.cpp:47-68), env (light_env.cpp:84-108): a
        // direction (fixed / uniform on the sphere) and a point on the disk of the scene's bounding sphere facing it
        const bool dirl = l.type == LIGHT_DIRECTIONAL;
        d = dirl ? f3(l.pos[0], l.pos[1], l.pos[2]) : sample_uniform_sphere(u2.x, u2.y);
        float dx, dy;
        sample_uniform_disk(u1.z, u1.w, dx, dy);
        F3 tu, tv;
        orthonormal_basis(d, tu, tv);
        o = bc - d * R + (tu * (dx * R) + tv * (dy * R));
        const float inv_pa = kPiF * R * R;
        w = (dirl ? Ke : light_env_radiance(s, l, -d)) * (dirl ? inv_pa : inv_pa * 4.0f * kPiF);
    }
    w = w * inv_psel;
    q.ray_o[slot] = make_float4(o.x, o.y, o.z, (float)kEps);
    q.ray_d[slot] = make_float4(d.x, d.y, d.z, (float)kInf);
    q.thr_pix[slot] = make_float4(w.x, w.y, w.z, 0.f);
    q.aux[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    q.info[slot] = make_uint2((uint32_t)g, 1u | (((uint32_t)(g >> 32) & 0xffffu) << 16));    // num_verts = 1: the light endpoint
}

// One light-path vertex: Russian roulette + component (renderer_lt.cpp:143-156 of the iteration that traced this ray), then
// the next iteration's camera connection (:96-118) and direction sample (:123-128).  A path whose throughput is zero is
// dropped at once (the reference drags it along; it contributes nothing).
template <bool COMPOSITE>
LMCU_HD ShadeOut lt_shade_body(const DevScene& s, const PtParams& p, float4 ro, float4 rd, float4 thr_pix, uint2 info, float4 hit) {
    ShadeOut out;
    out.alive = false; out.shadow = false; out.emit = f3s(0.f);
    const int32_t tri = (int32_t)f2u(hit.w);
    if (tri < 0) return out;                                                // no surface (or the environment): the walk ends
    const uint64_t sample = (uint64_t)info.x | ((uint64_t)(info.y >> 16) << 32);
    const uint64_t sidx = p.sample_offset + sample;
    uint32_t num_verts = info.y & 0xffu;                                    // loop counter of the iteration that traced this ray (1 = primary)
    F3 thr = f3(thr_pix.x, thr_pix.y, thr_pix.z);
    const F3 wi = -f3(rd.x, rd.y, rd.z);
    const SurfPoint sp = surface_point<COMPOSITE>(s, (uint32_t)tri, hit.y, hit.z);
    const MaterialDev& mat = s.mats[sp.material];
    const R4 u0 = rand4(p.seed, kLightPathKey, sidx, num_verts * 4u + 4u);  // ud0, ud1, rr, uc0
    if (num_verts > 5u) {                                                   // renderer_lt.cpp:143-149
        const float q = fmaxf(0.2f, 1.0f - max3(thr));
        if (u0.z < q) return out;
        thr = thr * (1.0f / (1.0f - q));
    }
    const bool plain = !COMPOSITE || mat_is_plain(mat);
    VComp vc;
    float udc0 = u0.w;
    if (!plain) {
        vc = pick_component(s, mat, sp, u0.w);
        thr = thr * vc.cweight;
        udc0 = rand4(p.seed, kLightPathKey, sidx, num_verts * 4u + 7u).x;
    }
    num_verts = num_verts < 2u ? 2u : num_verts + 1u;                       // the loop iteration that starts at this vertex
    if ((int)num_verts >= p.max_verts) return out;                          // for (num_verts = 2; num_verts < max_verts; ...)
    const bool specular = plain ? mat_is_specular(mat.type) : vc.specular;

    // camera connection (path::sample_direct(EL), camera_pinhole.cpp:178-202): We / p = J(wo) * G(surface, pinhole)
    if (!specular) {
        const CameraDev& c = s.cam;
        const F3 cpos = f3((float)c.pos[0], (float)c.pos[1], (float)c.pos[2]);
        const F3 dv = sp.p - cpos;
        const float d2 = dot(dv, dv), dist = sqrtf(d2);
        const F3 wc = dv * (1.0f / dist);                                   // camera -> surface
        float rx, ry;
        const float We = camera_importance(c, wc, rx, ry);
        if (We > 0.f) {
            const F3 wo = -wc;                                              // surface -> camera
            const float G = fabsf(dot(sp.n, wo)) / d2;                      // surface.h:329-348 toward a degenerated point
            F3 fs = plain ? mat_eval(mat, sp.n, wi, wo) : comp_eval(s, vc, sp, wi, wo);
            fs = fs * sn_correction(sp.n, sp.gn, wi, wo, true);
            const F3 C = thr * fs * (We * G);
            if (!is_zero(C) && G > 0.f) {
                out.shadow = true;
                out.sh_o = make_float4(sp.p.x, sp.p.y, sp.p.z, (float)kEps);
                out.sh_d = make_float4(wo.x, wo.y, wo.z, dist * (1.0f - (float)kEps));
                out.sh_c = make_float4(C.x, C.y, C.z, u2f(raster_to_pixel(p, rx, ry)));
            }
        }
    }
    // next direction (path::sample_direction(LE))
    BsdfSample bs;
    if (plain) bs = mat_sample(mat, sp.n, sp.gn, wi, u0.x, u0.y, u0.w, true); else bs = comp_sample(s, vc, sp, wi, u0.x, u0.y, udc0, true);
    if (!bs.valid) return out;
    thr = thr * bs.weight;
    if (is_zero(thr)) return out;
    out.alive = true;
    out.o = make_float4(sp.p.x, sp.p.y, sp.p.z, (float)kEps);
    out.d = make_float4(bs.wo.x, bs.wo.y, bs.wo.z, (float)kInf);
    out.thr_pix = make_float4(thr.x, thr.y, thr.z, 0.f);
    out.aux = make_float4(0.f, 0.f, 0.f, 0.f);
    out.info = make_uint2(info.x, num_verts | (info.y & 0xffff0000u));
    return out;
}

}  // namespace lmcu
