// bdpt.cuh — bidirectional path tracing, one sample per thread (src/renderer/renderer_bdpt.cpp:304-407, include/lm/bidir.h).
//
// A sample = one eye subpath + one light subpath of up to max_verts vertices each (no Russian roulette, bidir.h:487-537),
// every connection (s, t) with min_verts <= s + t <= max_verts (renderer_bdpt.cpp:350-361), each weighted by the power
// heuristic over ALL strategies that could have produced the same full path (bidir.h:446-466).
//
// The reference evaluates everything from the VERTICES of the connected full path (Path::eval_sampling_weight_bidir,
// Path::pdf_bidir); so does this file, through one generic vertex record and the "extended" geometry of surface.h (points
// at infinity, degenerated points).  Two things are done once instead of per strategy:
//   * the subpath weights alpha_L(s), alpha_E(t) (bidir.h:252-294) depend on one subpath only and are prefix products;
//   * p_{s'}(x) / p_s(x) for s' = 0..n is a running product of per-vertex factors fL[i] / fE[i] (density of generating
//     vertex i from the light side / from the eye side); all of them but the four around the connecting edge belong to one
//     subpath and are computed once per sample, so a connection costs 4 density evaluations + O(n) multiply-adds instead
//     of the reference's O(n^2) evaluations.
//
// Scope: every material of the path tracer (diffuse, glossy, glass, mirror, mask, mapKd textures, mixture_wavefrontobj,
// constant-weight mixtures), the five light types, the pinhole camera.
#pragma once
#include "kernels.cuh"

namespace lmcu {

constexpr int kBdptMaxVerts = 20;
constexpr uint32_t kEyePathKey = 0xfffffffdu;     // RNG counter word 0 of the eye subpath's draws (the light subpath uses kLightPathKey)

enum { BV_CAMERA = 0, BV_LIGHT = 1, BV_SURFACE = 2 };

struct BdptCounters { uint32_t rays, overflow; };   // per thread: ray queries issued, traversal stack overflow seen

// lm::Vert (bidir.h:24-38): a scene interaction + the component sampled at it
struct BVert {
    F3 p;                 // position (infinite: a representative point, unused)
    F3 n;                 // shading normal; infinite: PointGeometry::wo, the direction FROM the point at infinity
    F3 gn;
    float tu, tv;         // texture coordinates (composite materials)
    int32_t mat;          // surface: index into DevScene::mats
    int32_t light;        // surface: light index of an emitter, else -1; light endpoint: its light
    uint8_t kind;         // BV_*
    uint8_t comp;         // the component Material::sample_component picked: glass 0 = reflection, 1 = refraction
                          // (material_glass.cpp:124-136); mixture_wavefrontobj 0 = diffuse + glossy, 1 = alpha mask
                          // (model_wavefrontobj.cpp:413-420); mixture_rr the entry; mixture_marginalized the group
    uint8_t infinite, degenerated;
};

// Path::direction (bidir.h:143-157)
LMCU_HD F3 bv_dir(const BVert& a, const BVert& b) {
    if (a.infinite) return a.n;
    if (b.infinite) return -b.n;
    return normalize(b.p - a.p);
}

// surface::geometry_term (surface.h:329-348)
LMCU_HD float bv_G(const BVert& a, const BVert& b) {
    if (a.infinite && b.infinite) return 0.f;
    F3 d;
    float L2;
    if (a.infinite || b.infinite) {
        d = a.infinite ? a.n : -b.n;
        L2 = 1.f;
    } else {
        d = b.p - a.p;
        L2 = dot(d, d);
        d = d * (1.0f / sqrtf(L2));
    }
    const float c1 = (a.degenerated || a.infinite) ? 1.f : fabsf(dot(a.n, d));
    const float c2 = (b.degenerated || b.infinite) ? 1.f : fabsf(dot(b.n, -d));
    return c1 * c2 / L2;
}

// -- composite materials (textures, mixtures): the component by INDEX -----------------------------------------------------
LMCU_HD SurfPoint bv_surf(const BVert& v) {
    SurfPoint sp;
    sp.p = v.p; sp.n = v.n; sp.gn = v.gn; sp.tu = v.tu; sp.tv = v.tv;
    sp.prim = -1; sp.face = -1; sp.material = v.mat; sp.light = v.light;
    return sp;
}

// Material::sample_component: the index pick_component() (kernels.cuh) would choose for uc0
LMCU_HD int comp_pick_index(const DevScene& s, const MaterialDev& mat, const SurfPoint& sp, float uc0) {
    if (mat.type <= MAT_MASK) return 0;
    if (mat.type == MAT_MIXTURE_OBJ) {                                     // model_wavefrontobj.cpp:413-420
        float alpha = 1.f;
        if (mat.tex_kd >= 0 && mat.alpha_mask) alpha = tex_fetch(s.textures[mat.tex_kd], s.texels, sp.tu, sp.tv).w;
        return uc0 < alpha ? 0 : 1;
    }
    if (mat.type == MAT_MIXTURE_RR) {                                      // material_mixture.cpp:61-65
        float acc = 0.f;
        int k = 0;
        for (; k < mat.num_entries - 1; k++) {
            acc += s.mix[mat.entry_offset + k].prob;
            if (uc0 < acc) break;
        }
        return k;
    }
    const int n0 = mat.num_group0;                                         // marginalized, material_mixture.cpp:191-195
    float acc = n0 > 0 ? s.mix[mat.entry_offset].group_prob : 0.f;
    if (n0 > 0 && (uc0 < acc || n0 == mat.num_entries)) return 0;
    int k = n0;
    for (; k < mat.num_entries - 1; k++) {
        acc += s.mix[mat.entry_offset + k].group_prob;
        if (uc0 < acc) break;
    }
    return 1 + (k - n0);
}

// ... and the component record of that index (the same records pick_component() builds); 1 / cweight = pdf_component
LMCU_HD VComp comp_by_index(const DevScene& s, const MaterialDev& mat, const SurfPoint& sp, int comp) {
    VComp c;
    c.kind = 0; c.first = 0; c.n = 1; c.marginal = false; c.cweight = 1.f; c.alpha = 1.f; c.wsel = 1.f;
    if (mat.type <= MAT_MASK) {
        c.l0 = resolve_lobe(s, mat, sp);
        c.specular = mat_is_specular(mat.type);
    } else if (mat.type == MAT_MIXTURE_OBJ) {
        F3 kd = f3(mat.r, mat.g, mat.b);
        float alpha = 1.f;
        if (mat.tex_kd >= 0) {
            const float4 t = tex_fetch(s.textures[mat.tex_kd], s.texels, sp.tu, sp.tv);
            kd = f3(t.x, t.y, t.z);
            if (mat.alpha_mask) alpha = t.w;
        }
        if (comp == 0) {
            c.kind = 1; c.n = 2; c.marginal = true; c.specular = false;
            c.cweight = 1.0f / alpha; c.alpha = alpha;
            c.l0 = mat; c.l0.type = MAT_DIFFUSE; c.l0.r = kd.x; c.l0.g = kd.y; c.l0.b = kd.z;
            c.l1 = mat; c.l1.type = MAT_GLOSSY; c.l1.r = mat.ks_r; c.l1.g = mat.ks_g; c.l1.b = mat.ks_b;
            const float wd = max3(kd), wg = fmaxf(mat.ks_r, fmaxf(mat.ks_g, mat.ks_b));
            c.wsel = (wd == 0.f && wg == 0.f) ? 1.f : wd / (wd + wg);
        } else {
            c.kind = 3; c.n = 0; c.specular = true;
            c.cweight = 1.0f / (1.0f - alpha); c.alpha = alpha;
        }
    } else if (mat.type == MAT_MIXTURE_RR) {
        const MixEntry e = s.mix[mat.entry_offset + comp];
        c.kind = 2; c.first = mat.entry_offset + comp; c.n = 1;
        c.cweight = 1.0f / e.prob;
        c.specular = mat_is_specular(s.mats[e.material].type);
    } else {
        const int n0 = mat.num_group0;
        if (comp == 0) {
            c.kind = 2; c.first = mat.entry_offset; c.n = n0; c.marginal = true; c.specular = false;
            c.cweight = n0 > 0 ? 1.0f / s.mix[mat.entry_offset].group_prob : 0.f;
        } else {
            const int k = n0 + comp - 1;
            c.kind = 2; c.first = mat.entry_offset + k; c.n = 1; c.specular = true;
            c.cweight = 1.0f / s.mix[mat.entry_offset + k].group_prob;
        }
    }
    return c;
}

// path::is_specular_component (path.h:695-704)
template <bool COMP>
LMCU_HD bool bv_specular(const DevScene& s, const BVert& v) {
    if (v.kind == BV_SURFACE) {
        const MaterialDev& m = s.mats[v.mat];
        return (!COMP || mat_is_plain(m)) ? mat_is_specular(m.type) : comp_by_index(s, m, bv_surf(v), v.comp).specular;
    }
    if (v.kind == BV_LIGHT) return s.lights[v.light].type == LIGHT_DIRECTIONAL;
    return false;
}

// path::is_connectable_endpoint (path.h:716-725): pinhole, area and point lights; not the lights at infinity
LMCU_HD bool bv_connectable(const DevScene& s, const BVert& v) {
    if (v.kind == BV_CAMERA) return true;
    const int t = s.lights[v.light].type;
    return t == LIGHT_AREA || t == LIGHT_POINT;
}

// Fresnel reflectance of material::glass for incident direction wi (material_glass.cpp:90-99,125-131)
LMCU_HD float bv_glass_fr(const MaterialDev& m, F3 n, F3 wi) {
    F3 wt; float eta; bool total;
    return glass_fresnel(m, n, wi, wt, eta, total);
}

// path::pdf_component (path.h:228-234): glass picks one of its two delta lobes by Fresnel; mixtures by their pmfs
template <bool COMP>
LMCU_HD float bv_pdf_comp(const DevScene& s, const BVert& v, F3 wi) {
    if (v.kind != BV_SURFACE) return 1.f;
    const MaterialDev& m = s.mats[v.mat];
    if (m.type == MAT_GLASS) {
        const float Fr = bv_glass_fr(m, v.n, wi);
        return v.comp == 0 ? Fr : 1.f - Fr;
    }
    if (!COMP || mat_is_plain(m)) return 1.f;
    const float cw = comp_by_index(s, m, bv_surf(v), v.comp).cweight;
    return cw > 0.f ? 1.0f / cw : 0.f;
}

// path::pdf_direction with eval_delta = false (path.h:457-470): projected solid angle; delta lobes report 1
template <bool COMP>
LMCU_HD float bv_pdf_dir(const DevScene& s, const BVert& v, F3 wi, F3 wo) {
    if (v.kind == BV_CAMERA) {
        float rx, ry;
        return camera_importance(s.cam, wo, rx, ry);                       // camera_pinhole.cpp:157-163
    }
    if (v.kind == BV_LIGHT) {
        const int t = s.lights[v.light].type;
        if (t == LIGHT_AREA) return dot(wo, v.n) <= 0.f ? 0.f : 1.0f / kPiF;   // light_area.cpp:116-121
        if (t == LIGHT_POINT) return 1.0f / (4.0f * kPiF);                     // light_point.cpp:63-65
        return 0.f;                                                            // lights at infinity: joint sampling only
    }
    const MaterialDev& m = s.mats[v.mat];
    if (!COMP || mat_is_plain(m)) return mat_is_specular(m.type) ? 1.f : mat_pdf(m, v.n, wi, wo);
    const SurfPoint sp = bv_surf(v);
    const VComp c = comp_by_index(s, m, sp, v.comp);
    return c.specular ? 1.f : comp_pdf(s, c, sp, wi, wo);
}

// path::eval_contrb_direction (path.h:761-775).  le: TransDir::LE.  Surfaces include the shading-normal correction.
template <bool COMP>
LMCU_HD F3 bv_eval(const DevScene& s, const BVert& v, F3 wi, F3 wo, bool le, bool eval_delta) {
    if (v.kind == BV_CAMERA) {
        float rx, ry;
        return f3s(camera_importance(s.cam, wo, rx, ry));                  // camera_pinhole.cpp:232-237
    }
    if (v.kind == BV_LIGHT) {
        const LightDev& l = s.lights[v.light];
        const F3 Ke = f3(l.Ke[0], l.Ke[1], l.Ke[2]);
        if (l.type == LIGHT_AREA) return dot(wo, v.n) <= 0.f ? f3s(0.f) : Ke;   // light_area.cpp:195-197
        if (l.type == LIGHT_POINT) return Ke;
        if (l.type == LIGHT_DIRECTIONAL) return eval_delta ? f3s(0.f) : Ke;      // light_directional.cpp:124-126
        return light_env_radiance(s, l, -wo);
    }
    const MaterialDev& m = s.mats[v.mat];
    F3 f;
    if (!COMP || mat_is_plain(m)) {
        if (mat_is_specular(m.type)) {
            if (eval_delta) return f3s(0.f);
            if (m.type == MAT_GLASS) {                                     // material_glass.cpp:249-266
                F3 wt; float eta; bool total;
                const float Fr = glass_fresnel(m, v.n, wi, wt, eta, total);
                f = v.comp == 0 ? f3s(Fr) : f3s((1.f - Fr) * (le ? 1.f : eta * eta));
            } else {
                f = f3s(1.f);                                              // mirror, mask
            }
        } else {
            f = mat_eval(m, v.n, wi, wo);
        }
    } else {
        const SurfPoint sp = bv_surf(v);
        const VComp c = comp_by_index(s, m, sp, v.comp);
        if (c.specular) {
            if (eval_delta) return f3s(0.f);
            // delta components with eval_delta = false: the alpha mask evaluates to 1 - alpha (model_wavefrontobj.cpp:497-499),
            // a specular mixture entry to weight x 1 (material_mixture.cpp:100-103,266-274; mirror and mask evaluate to 1)
            f = c.kind == 3 ? f3s(1.0f - c.alpha) : (c.kind == 2 ? f3s(s.mix[c.first].weight) : f3s(1.f));
        } else {
            f = comp_eval(s, c, sp, wi, wo);
        }
    }
    return f * sn_correction(v.n, v.gn, wi, wo, le);
}

// path::pdf_position of a connectable endpoint (path.h:327-344): pinhole 1; area p_sel / (A J); point p_sel
LMCU_HD float bv_pdf_position(const DevScene& s, const BVert& v) {
    if (v.kind == BV_CAMERA) return 1.f;
    const LightDev& l = s.lights[v.light];
    const float psel = 1.0f / (float)s.num_lights;
    return l.type == LIGHT_AREA ? psel * l.inv_area_j : psel;
}

// path::pdf_primary_ray with eval_delta = false (path.h:149-162) of a light endpoint: selection x Light::pdf_ray
LMCU_HD float bv_pdf_primary_ray(const DevScene& s, const BVert& v, F3 wo) {
    if (v.kind == BV_CAMERA) {
        float rx, ry;
        return camera_importance(s.cam, wo, rx, ry);
    }
    const LightDev& l = s.lights[v.light];
    const float psel = 1.0f / (float)s.num_lights;
    const float inv_disk = 1.0f / (kPiF * s.bound_radius * s.bound_radius);
    if (l.type == LIGHT_AREA) return psel * l.inv_area_j * (dot(wo, v.n) <= 0.f ? 0.f : 1.0f / kPiF);
    if (l.type == LIGHT_POINT) return psel / (4.0f * kPiF);
    if (l.type == LIGHT_DIRECTIONAL) return psel * inv_disk;                  // light_directional.cpp:70-73
    return psel * inv_disk / (4.0f * kPiF);                                   // light_envconst.cpp:70-74, light_env.cpp:110-114
}

// Material::sample_direction for the component stored at the vertex (materials' sample_direction; glass:
// material_glass.cpp:140-166).  Returns false when the material rejects the sample (glossy below the surface).
template <bool COMP>
LMCU_HD bool bv_sample_dir(const DevScene& s, const BVert& v, F3 wi, float ud0, float ud1, float udc0, F3& wo) {
    const MaterialDev& m = s.mats[v.mat];
    if (COMP && !mat_is_plain(m)) {
        const SurfPoint sp = bv_surf(v);
        const BsdfSample b = comp_sample(s, comp_by_index(s, m, sp, v.comp), sp, wi, ud0, ud1, udc0, false);
        wo = b.wo;
        return b.valid;
    }
    if (m.type == MAT_GLASS) {
        if (v.comp == 0) { wo = reflection(wi, v.n); return true; }
        F3 wt; float eta; bool total;
        glass_fresnel(m, v.n, wi, wt, eta, total);
        wo = wt;
        return !total;
    }
    const BsdfSample b = mat_sample(m, v.n, v.gn, wi, ud0, ud1, 0.f, false);
    wo = b.wo;
    return b.valid;
}

// path::sample_subpath (bidir.h:487-553) in three steps (the device runs them as separate kernels around the extend queue).
// Random numbers of a subpath: block 0 / 1 = the primary ray (RaySampleU), block 3 + 2k = component of vertex k, 4 + 2k = its
// direction.

// Vertex 0 and the primary ray: Camera_Pinhole::sample_ray (camera_pinhole.cpp:135-141) / path::sample_primary_ray(LE)
// (path.h:101-118: a light chosen uniformly, Light::sample_ray)
LMCU_HD void bdpt_first_vertex(const DevScene& s, const PtParams& p, uint64_t sidx, bool light_side, BVert& v, F3& o, F3& d) {
    const uint32_t key = light_side ? kLightPathKey : kEyePathKey;
    v.mat = -1; v.light = -1; v.comp = 0; v.infinite = 0; v.degenerated = 0;
    v.n = f3s(0.f); v.gn = f3s(0.f); v.tu = 0.f; v.tv = 0.f;
    const R4 u1 = rand4(p.seed, key, sidx, 0);
    const R4 u2 = rand4(p.seed, key, sidx, 1);
    if (!light_side) {
        D3 od, dd;
        camera_primary_ray_d(s.cam, (double)u2.x, (double)u2.y, od, dd);
        v.kind = BV_CAMERA; v.degenerated = 1;
        v.p = f3((float)od.x, (float)od.y, (float)od.z);
        o = v.p; d = f3((float)dd.x, (float)dd.y, (float)dd.z);
        return;
    }
    int li = (int)(u1.x * (float)s.num_lights);
    li = li < 0 ? 0 : (li > s.num_lights - 1 ? s.num_lights - 1 : li);
    const LightDev& l = s.lights[li];
    v.kind = BV_LIGHT; v.light = li;
    if (l.type == LIGHT_AREA) {                                    // light_area.cpp:47-59,77-97
        const int ti = cdf_sample(s.cdf + l.cdf_offset, l.num_tris, u1.y);
        const EmitTri& e = s.emit[l.tri_offset + ti];
        const float sq = safe_sqrt(u1.z);
        const float b1 = 1.0f - sq, b2 = u1.w * sq;
        const F3 a = f3(e.p1.x, e.p1.y, e.p1.z), b = f3(e.p2.x, e.p2.y, e.p2.z), c = f3(e.p3.x, e.p3.y, e.p3.z);
        v.p = a * (1.0f - b1 - b2) + b * b1 + c * b2;
        v.n = f3(e.n.x, e.n.y, e.n.z); v.gn = v.n;
        const F3 dl = sample_cosine_weighted(u2.x, u2.y);
        F3 tu, tv;
        orthonormal_basis(v.n, tu, tv);
        d = tu * dl.x + tv * dl.y + v.n * dl.z;
    } else if (l.type == LIGHT_POINT) {                            // light_point.cpp:38-48
        v.p = f3(l.pos[0], l.pos[1], l.pos[2]); v.degenerated = 1;
        d = sample_uniform_sphere(u2.x, u2.y);
    } else {                                                       // light_directional.cpp:50-68, light_envconst.cpp:47-68
        d = l.type == LIGHT_DIRECTIONAL ? f3(l.pos[0], l.pos[1], l.pos[2]) : sample_uniform_sphere(u2.x, u2.y);
        float dx, dy;
        sample_uniform_disk(u1.z, u1.w, dx, dy);
        F3 tu, tv;
        orthonormal_basis(d, tu, tv);
        const F3 bc = f3(s.bound_center[0], s.bound_center[1], s.bound_center[2]);
        v.p = bc - d * s.bound_radius + (tu * (dx * s.bound_radius) + tv * (dy * s.bound_radius));
        v.n = d; v.infinite = 1;
    }
    o = v.p;
}

// Vertex k from the closest hit of the ray (o, d) (scene.cpp:319-345: a miss is the environment light's endpoint, or
// nothing) + path::sample_component there.  Returns false when the walk ends without a vertex.
template <bool COMP>
LMCU_HD bool bdpt_hit_vertex(const DevScene& s, const PtParams& p, uint64_t sidx, bool light_side, int k, int32_t tri, float hu, float hv,
                             F3 d, BVert& v) {
    v.comp = 0; v.infinite = 0; v.degenerated = 0; v.tu = 0.f; v.tv = 0.f;
    if (tri < 0) {
        if (s.env_light < 0) return false;
        v.kind = BV_LIGHT; v.light = s.env_light; v.mat = -1; v.infinite = 1;
        v.p = f3s(0.f); v.n = -d; v.gn = f3s(0.f);
        return true;
    }
    const SurfPoint sp = surface_point<COMP>(s, (uint32_t)tri, hu, hv);
    v.kind = BV_SURFACE; v.p = sp.p; v.n = sp.n; v.gn = sp.gn; v.tu = sp.tu; v.tv = sp.tv; v.mat = sp.material; v.light = sp.light;
    const MaterialDev& m = s.mats[sp.material];
    const R4 uc = rand4(p.seed, light_side ? kLightPathKey : kEyePathKey, sidx, 3u + 2u * (uint32_t)k);
    if (m.type == MAT_GLASS) v.comp = uc.x < bv_glass_fr(m, sp.n, -d) ? 0 : 1;         // material_glass.cpp:124-136
    else if (COMP && !mat_is_plain(m)) v.comp = (uint8_t)comp_pick_index(s, m, sp, uc.x);
    return true;
}

// Direction out of surface vertex k (path::sample_direction, path.h:414-424); wi is the direction to the previous VERTEX
template <bool COMP>
LMCU_HD bool bdpt_next_dir(const DevScene& s, const PtParams& p, uint64_t sidx, bool light_side, int k, const BVert& v, const BVert& prev, F3& wo) {
    const R4 ud = rand4(p.seed, light_side ? kLightPathKey : kEyePathKey, sidx, 4u + 2u * (uint32_t)k);
    return bv_sample_dir<COMP>(s, v, bv_dir(v, prev), ud.x, ud.y, ud.z, wo);
}

// The whole walk by one thread (tests/hostsim).  Fills vs[0..n), returns n.
template <typename TriT, bool COMP>
LMCU_HD int bdpt_subpath(const DevScene& s, const PtParams& p, uint64_t sidx, bool light_side, int max_verts, BVert* vs,
                         uint2* stack, BdptCounters* ctr) {
    if (max_verts <= 0) return 0;
    int n = 0;
    F3 o, d;
    bdpt_first_vertex(s, p, sidx, light_side, vs[0], o, d);
    n = 1;
    if (max_verts == 1) return n;       // only the endpoint is used (the single-strategy modes)
    while (true) {
        HitD h;
        if (!pt_trace<false, false, TriT>(s, make_float4(o.x, o.y, o.z, (float)kEps), make_float4(d.x, d.y, d.z, (float)kInf), h, stack, nullptr)) ctr->overflow = 1u;
        ctr->rays++;
        BVert v;
        if (!bdpt_hit_vertex<COMP>(s, p, sidx, light_side, n, h.tri, (float)h.u, (float)h.v, d, v)) break;
        vs[n++] = v;
        if (v.infinite || n >= max_verts) break;
        F3 wo;
        if (!bdpt_next_dir<COMP>(s, p, sidx, light_side, n - 1, v, vs[n - 2], wo)) break;
        o = v.p; d = wo;
    }
    return n;
}

// Path::eval_subpath_sampling_weight (bidir.h:252-294) for every prefix length: alpha[l], l = 0..n.  pc[i] = the component
// pdf of vertex i that went into alpha[i + 1] (the caller takes it out again when the vertex becomes an endpoint).
template <bool COMP>
LMCU_HD void bdpt_alphas(const DevScene& s, const BVert* vs, int n, bool le, F3* alpha, float* pc) {
    alpha[0] = f3s(1.f);
    for (int l = 1; l <= n; l++) alpha[l] = f3s(0.f);
    if (n == 0) return;
    int i = 0;
    pc[0] = 1.f;
    if (bv_connectable(s, vs[0])) {
        alpha[1] = f3s(1.0f / bv_pdf_position(s, vs[0]));
    } else {
        if (n < 2) return;
        const F3 d01 = bv_dir(vs[0], vs[1]);
        const F3 f = bv_eval<COMP>(s, vs[0], f3s(0.f), d01, le, false);
        pc[1] = bv_pdf_comp<COMP>(s, vs[1], -d01);
        const float pr = bv_pdf_primary_ray(s, vs[0], d01) * pc[1];
        if (is_zero(f) || !(pr > 0.f)) return;
        alpha[2] = f * (1.0f / pr);
        i = 1;
    }
    for (; i < n - 1; i++) {
        const BVert& v = vs[i];
        const F3 wi = i > 0 ? bv_dir(v, vs[i - 1]) : f3s(0.f);
        const F3 wo = bv_dir(v, vs[i + 1]);
        const F3 f = bv_eval<COMP>(s, v, wi, wo, le, false);
        pc[i + 1] = bv_pdf_comp<COMP>(s, vs[i + 1], -wo);
        const float pd = bv_pdf_dir<COMP>(s, v, wi, wo) * pc[i + 1];
        if (is_zero(f) || !(pd > 0.f)) return;
        alpha[i + 2] = alpha[i + 1] * f * (1.0f / pd);
    }
}

// Density (area measure x component) of generating vertex `cur` from `prev`, whose own previous vertex is `prev2` (null for
// an endpoint): one factor of the products in Path::pdf_bidir (bidir.h:397-431)
template <bool COMP>
LMCU_HD float bdpt_gen_pdf(const DevScene& s, const BVert* prev2, const BVert& prev, const BVert& cur) {
    const F3 wo = bv_dir(prev, cur);
    const F3 wi = prev2 ? bv_dir(prev, *prev2) : f3s(0.f);
    return bv_pdf_dir<COMP>(s, prev, wi, wo) * bv_G(prev, cur) * bv_pdf_comp<COMP>(s, cur, -wo);
}

// Per-subpath halves of the strategy densities (Path::pdf_bidir, bidir.h:388-437).  For vertex i of a subpath:
//   F[i] = density of generating it the way the subpath did (from vertex i-1, given i-2; i = 0: position pdf; a light at
//          infinity is sampled jointly with vertex 1: F[0] = 1, F[1] = pdf_primary_ray x G x component),
//   R[i] = density of generating it from the OTHER end, from vertex i+1 given i+2 (defined for i + 2 < n),
//   spec[i] = the vertex is a delta interaction.
// None of them depends on the connection; the four densities around the connecting edge are evaluated per strategy.
template <bool COMP>
LMCU_HD void bdpt_subpath_pdfs(const DevScene& s, const BVert* vs, int n, float* F, float* R, bool* spec) {
    for (int i = 0; i < n; i++) spec[i] = bv_specular<COMP>(s, vs[i]);
    if (n == 0) return;
    const bool conn0 = bv_connectable(s, vs[0]);
    F[0] = conn0 ? bv_pdf_position(s, vs[0]) : 1.f;
    for (int i = 1; i < n; i++) {
        if (i == 1 && !conn0) {
            const F3 d01 = bv_dir(vs[0], vs[1]);
            F[1] = bv_pdf_primary_ray(s, vs[0], d01) * bv_G(vs[0], vs[1]) * bv_pdf_comp<COMP>(s, vs[1], -d01);
        } else {
            F[i] = bdpt_gen_pdf<COMP>(s, i >= 2 ? &vs[i - 2] : nullptr, vs[i - 1], vs[i]);
        }
    }
    for (int i = 0; i + 2 < n; i++) R[i] = bdpt_gen_pdf<COMP>(s, &vs[i + 2], vs[i + 1], vs[i]);
}

// The same per-vertex data (bdpt_alphas + bdpt_subpath_pdfs), incrementally: everything that becomes known when vertex k >= 1
// of a subpath is created, from v_k, v_{k-1}, v_{k-2} (null for k = 1) and alpha[k] — the wavefront computes it in the
// kernel that creates the vertex.  alpha_next = alpha[k + 1]; R_prev2 = R[k - 2] (k >= 2).
struct BdptVertexData { F3 alpha_next; float pc, F, R_prev2; bool spec; };

template <bool COMP>
LMCU_HD BdptVertexData bdpt_vertex_data(const DevScene& s, bool le, int k, const BVert& v, const BVert& prev, const BVert* prev2, F3 alpha_k) {
    BdptVertexData o;
    o.spec = bv_specular<COMP>(s, v);
    const F3 wo = bv_dir(prev, v);
    const F3 wi = prev2 ? bv_dir(prev, *prev2) : f3s(0.f);
    o.pc = bv_pdf_comp<COMP>(s, v, -wo);
    const F3 f = bv_eval<COMP>(s, prev, wi, wo, le, false);
    const bool joint = k == 1 && !bv_connectable(s, prev);          // a light at infinity: vertex 0 and 1 are sampled jointly
    const float pdir = joint ? bv_pdf_primary_ray(s, prev, wo) : bv_pdf_dir<COMP>(s, prev, wi, wo);
    o.F = pdir * bv_G(prev, v) * o.pc;
    const float pd = pdir * o.pc;
    o.alpha_next = (is_zero(f) || !(pd > 0.f)) ? f3s(0.f) : (joint ? f : alpha_k * f) * (1.0f / pd);
    o.R_prev2 = prev2 ? bdpt_gen_pdf<COMP>(s, &v, prev, *prev2) : 0.f;
    return o;
}

// What one connection (s, t) produces: the MIS-weighted contribution, the pixel it belongs to and — for s >= 1 — the
// visibility ray that decides whether it counts (Scene::visible, scene.h:343-364).
struct BdptConn {
    F3 C;
    uint32_t pixel;
    bool shadow;
    float4 sh_o, sh_d;
};

// Strategy (sL, t) of one sample: path::connect_subpaths (bidir.h:585-637) without the visibility test,
// Path::eval_sampling_weight_bidir (bidir.h:319-343) and Path::eval_mis_weight (bidir.h:446-466).  `a` gives access to the
// two subpaths and their per-vertex data (local arrays on the host, SoA loads on the device):
//   BVert Y(i), Z(j)   light / eye subpath vertices      F3 aL(l), aE(l)   alpha of the first l vertices
//   float pcE(j)       component pdf inside aE(j + 1)     float yF(i), yR(i), zF(j), zR(j); bool ySpec(i), zSpec(j)
// Returns false when the strategy contributes nothing.
template <bool COMP, typename Acc>
LMCU_HD bool bdpt_connect(const DevScene& s, const PtParams& p, int sL, int t, const Acc& a, uint32_t eye_pixel, BdptConn& out) {
    const int n = sL + t;
    out.pixel = eye_pixel;
    out.shadow = false;
    F3 C;
    BVert x0, vL, vLp, vE, vEp;
    vE = a.Z(t - 1);
    if (t >= 2) vEp = a.Z(t - 2);
    if (sL == 0) {
        if (vE.light < 0 || vE.kind == BV_CAMERA) return false;                       // scene->is_light(vL.sp)
        x0 = vE; x0.kind = BV_LIGHT;
        const F3 cst = bv_eval<COMP>(s, x0, f3s(0.f), bv_dir(x0, vEp), true, true);         // Path::eval_connection_term, s = 0
        C = a.aE(t) * a.pcE(t - 1) * cst;                                             // the endpoint's component pdf is 1
    } else {
        vL = a.Y(sL - 1);
        if (sL >= 2) vLp = a.Y(sL - 2);
        if (vL.infinite || vE.infinite) return false;
        const F3 dLE = bv_dir(vL, vE);
        const F3 fsL = bv_eval<COMP>(s, vL, sL >= 2 ? bv_dir(vL, vLp) : f3s(0.f), dLE, true, true);
        const F3 fsE = bv_eval<COMP>(s, vE, t >= 2 ? bv_dir(vE, vEp) : f3s(0.f), -dLE, false, true);
        C = a.aL(sL) * a.aE(t) * fsL * fsE * bv_G(vL, vE);
        if (t == 1) {
            float rx, ry;
            if (camera_importance(s.cam, -dLE, rx, ry) <= 0.f) return false;
            out.pixel = raster_to_pixel(p, rx, ry);
        }
        const F3 dv = vE.p - vL.p;
        const float dist = sqrtf(dot(dv, dv));
        const F3 w = dv * (1.0f / dist);
        out.shadow = true;
        out.sh_o = make_float4(vL.p.x, vL.p.y, vL.p.z, (float)kEps);
        out.sh_d = make_float4(w.x, w.y, w.z, dist * (1.0f - (float)kEps));
    }
    if (is_zero(C) || !(C.x == C.x) || !(C.y == C.y) || !(C.z == C.z)) return false;
    if (p.bdpt_mode != 0u) {            // pt_naive_path / pt_nee_path / lt_nee_path: Path::eval_sampling_weight_bidir alone
        out.C = C;
        return true;
    }
    // -- the densities around the connecting edge; full path x_0 (light) .. x_{n-1} (camera), x_i = Y(i) for i < sL and
    // Z(n-1-i) above.  fL[i] / fE[i] = density of generating x_i from the light / eye side:
    // p_{s'} = prod_{i < s'} fL[i] * prod_{i >= s'} fE[i].
    bool spec0, conn0, degen0;
    float fl_a, fl_b = 0.f, fe_a, fe_b = 0.f, fl0 = 0.f;       // sL >= 1: fL[sL], fL[sL+1], fE[sL-1], fE[sL-2];  sL = 0: fL[1], -, fE[0], -, fL[0]
    if (sL == 0) {
        conn0 = bv_connectable(s, x0); spec0 = bv_specular<COMP>(s, x0); degen0 = x0.degenerated != 0;
        fl0 = conn0 ? bv_pdf_position(s, x0) : 1.f;
        if (conn0) fl_a = bdpt_gen_pdf<COMP>(s, nullptr, x0, vEp);
        else {
            const F3 d01 = bv_dir(x0, vEp);
            fl_a = bv_pdf_primary_ray(s, x0, d01) * bv_G(x0, vEp) * bv_pdf_comp<COMP>(s, vEp, -d01);
        }
        BVert vEpp;
        if (t >= 3) vEpp = a.Z(t - 3);
        fe_a = bdpt_gen_pdf<COMP>(s, t >= 3 ? &vEpp : nullptr, vEp, x0);
    } else {
        const BVert y0 = sL == 1 ? vL : (sL == 2 ? vLp : a.Y(0));
        conn0 = bv_connectable(s, y0); spec0 = a.ySpec(0); degen0 = y0.degenerated != 0;
        fl_a = bdpt_gen_pdf<COMP>(s, sL >= 2 ? &vLp : nullptr, vL, vE);
        if (t >= 2) fl_b = bdpt_gen_pdf<COMP>(s, &vL, vE, vEp);
        fe_a = bdpt_gen_pdf<COMP>(s, t >= 2 ? &vEp : nullptr, vE, vL);
        if (sL >= 2) fe_b = bdpt_gen_pdf<COMP>(s, &vE, vL, vLp);
    }
    auto FL = [&](int i) -> float {
        if (sL == 0) return i == 0 ? fl0 : (i == 1 ? fl_a : a.zR(n - 1 - i));
        return i < sL ? a.yF(i) : (i == sL ? fl_a : (i == sL + 1 ? fl_b : a.zR(n - 1 - i)));
    };
    auto FE = [&](int i) -> float {
        if (sL == 0) return i == 0 ? fe_a : a.zF(n - 1 - i);
        return i >= sL ? a.zF(n - 1 - i) : (i == sL - 1 ? fe_a : (i == sL - 2 ? fe_b : a.yR(i)));
    };
    auto spec = [&](int i) -> bool { return i == 0 ? spec0 : (i < sL ? a.ySpec(i) : a.zSpec(n - 1 - i)); };
    // Path::is_samplable_bidir (bidir.h:199-230)
    auto samplable = [&](int s2) -> bool {
        if (s2 == 0) return !degen0 && !spec0;
        if (s2 == n) return false;                                                    // the pinhole is degenerated
        if (s2 == 1 && !conn0) return false;
        return !spec(s2 - 1) && !spec(s2);
    };
    float inv_w = samplable(sL) ? 1.f : 0.f;
    float r = 1.f;
    for (int s2 = sL + 1; s2 <= n; s2++) {                                            // p_{s2} / p_s, upward
        const float fe = FE(s2 - 1);
        if (!(fe > 0.f)) break;
        r *= FL(s2 - 1) / fe;
        if (!(r > 0.f)) break;
        if (samplable(s2)) inv_w += r * r;
    }
    r = 1.f;
    for (int s2 = sL - 1; s2 >= 0; s2--) {                                            // downward
        const float fl = FL(s2);
        if (!(fl > 0.f)) break;
        r *= FE(s2) / fl;
        if (!(r > 0.f)) break;
        if (samplable(s2)) inv_w += r * r;
    }
    if (!(inv_w > 0.f)) return false;
    out.C = C * (1.0f / inv_w);
    return true;
}

// Which connections (sL, t) a sample evaluates (LmcuRenderParams::bdpt_strategies).  zspec(j) = eye vertex j is a delta
// interaction; only asked for vertices the strategy's eye subpath length implies.
//   1  renderer::pt_naive_path (renderer_bdpt.cpp:80-93): the eye subpath alone, (0, t)
//   2  renderer::pt_nee_path (:150-170): (1, n-1) unless eye vertex n-2 is specular, then (0, n)
//   3  renderer::lt_nee_path (:228-233): (n-1, 1)
template <typename ZSpec>
LMCU_HD bool bdpt_strategy_on(uint32_t mode, int sL, int t, const ZSpec& zspec) {
    switch (mode) {
    case 1u: return sL == 0;
    case 2u: return sL == 1 ? !zspec(t - 1) : (sL == 0 && t >= 2 && zspec(t - 2));
    case 3u: return t == 1;
    default: return true;
    }
}
// longest eye / light subpath any active strategy reads
LMCU_HD int bdpt_eye_max(uint32_t mode, int max_verts) { return mode == 3u ? 1 : max_verts; }
LMCU_HD int bdpt_light_max(uint32_t mode, int max_verts) { return mode == 1u ? 0 : (mode == 2u ? 1 : max_verts); }

// Subpath data in local arrays (the one-thread-per-sample form below)
struct BdptLocalAcc {
    const BVert* ys; const BVert* zs;
    const F3* aL_; const F3* aE_;
    const float* pcE_; const float* yF_; const float* yR_; const float* zF_; const float* zR_;
    const bool* ySpec_; const bool* zSpec_;
    LMCU_HD BVert Y(int i) const { return ys[i]; }
    LMCU_HD BVert Z(int j) const { return zs[j]; }
    LMCU_HD F3 aL(int l) const { return aL_[l]; }
    LMCU_HD F3 aE(int l) const { return aE_[l]; }
    LMCU_HD float pcE(int j) const { return pcE_[j]; }
    LMCU_HD float yF(int i) const { return yF_[i]; }
    LMCU_HD float yR(int i) const { return yR_[i]; }
    LMCU_HD float zF(int j) const { return zF_[j]; }
    LMCU_HD float zR(int j) const { return zR_[j]; }
    LMCU_HD bool ySpec(int i) const { return ySpec_[i]; }
    LMCU_HD bool zSpec(int j) const { return zSpec_[j]; }
};

// raster pixel of every strategy with t >= 2: the eye subpath's own primary ray
LMCU_HD uint32_t bdpt_eye_pixel(const DevScene& s, const PtParams& p, const BVert& z0, const BVert& z1) {
    float rx, ry;
    camera_importance(s.cam, bv_dir(z0, z1), rx, ry);
    return raster_to_pixel(p, rx, ry);
}
LMCU_HD uint32_t bdpt_eye_pixel(const DevScene& s, const PtParams& p, const BVert* zs, int nE) {
    return nE < 2 ? 0u : bdpt_eye_pixel(s, p, zs[0], zs[1]);
}

// One sample by one thread: both subpaths, every connection; splat(pixel, C) receives the weighted contributions.  This is
// the form tests/hostsim runs on the CPU; the device runs the same functions as three kernels (lmcu.cu: k_bdpt_subpaths,
// k_bdpt_connect, k_pt_shadow_q).
template <typename TriT, bool COMP, typename Splat>
LMCU_HD void bdpt_sample(const DevScene& s, const PtParams& p, int min_verts, uint64_t g, uint2* stack, BdptCounters* ctr, Splat& splat) {
    const uint64_t sidx = p.sample_offset + g;
    BVert zs[kBdptMaxVerts], ys[kBdptMaxVerts];
    F3 aE[kBdptMaxVerts + 1], aL[kBdptMaxVerts + 1];
    float pcE[kBdptMaxVerts], pcL[kBdptMaxVerts];
    float zF[kBdptMaxVerts], zR[kBdptMaxVerts], yF[kBdptMaxVerts], yR[kBdptMaxVerts];
    bool zSpec[kBdptMaxVerts], ySpec[kBdptMaxVerts];
    const int nE = bdpt_subpath<TriT, COMP>(s, p, sidx, false, bdpt_eye_max(p.bdpt_mode, p.max_verts), zs, stack, ctr);
    const int nL = bdpt_subpath<TriT, COMP>(s, p, sidx, true, bdpt_light_max(p.bdpt_mode, p.max_verts), ys, stack, ctr);
    bdpt_alphas<COMP>(s, zs, nE, false, aE, pcE);
    bdpt_alphas<COMP>(s, ys, nL, true, aL, pcL);
    bdpt_subpath_pdfs<COMP>(s, zs, nE, zF, zR, zSpec);
    bdpt_subpath_pdfs<COMP>(s, ys, nL, yF, yR, ySpec);
    const uint32_t eye_pixel = bdpt_eye_pixel(s, p, zs, nE);
    BdptLocalAcc acc;
    acc.ys = ys; acc.zs = zs; acc.aL_ = aL; acc.aE_ = aE; acc.pcE_ = pcE;
    acc.yF_ = yF; acc.yR_ = yR; acc.zF_ = zF; acc.zR_ = zR; acc.ySpec_ = ySpec; acc.zSpec_ = zSpec;
    for (int sL = 0; sL <= nL; sL++) {
        for (int t = 1; t <= nE; t++) {                     // t = 0: a light subpath never hits the pinhole (bidir.h:600-606,626-629)
            const int n = sL + t;
            if (n < min_verts || n > p.max_verts || n < 2) continue;
            if (!bdpt_strategy_on(p.bdpt_mode, sL, t, [&](int j) { return zSpec[j]; })) continue;
            BdptConn c;
            if (!bdpt_connect<COMP>(s, p, sL, t, acc, eye_pixel, c)) continue;
            if (c.shadow) {
                HitD h;
                if (!pt_trace<true, false, TriT>(s, c.sh_o, c.sh_d, h, stack, nullptr)) ctr->overflow = 1u;
                ctr->rays++;
                if (h.tri >= 0) continue;
            }
            splat(c.pixel, c.C);
        }
    }
}

}  // namespace lmcu
