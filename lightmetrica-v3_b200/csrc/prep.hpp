// prep.hpp — host-side preparation of an LmcuSceneDesc for upload (pure C++, no CUDA calls):
// validation, per-primitive normal matrices, emitter triangles + area CDFs (light_area.cpp:68-73),
// camera basis in FP64 with glm's operation order (camera_pinhole.cpp:90-95).
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include "../../include/lmcu.h"
#include "kernels.cuh"

namespace lmcu {

struct PreparedScene {
    std::vector<uint32_t> face_offset;       // per mesh, into the concatenated face arrays
    std::vector<PrimDev> prims;
    std::vector<MaterialDev> mats;
    std::vector<LightDev> lights;
    std::vector<EmitTri> emit;
    std::vector<float> cdf;
    std::vector<float4> env_tex;             // light::env texels (rgb, texture::bitmap row order)
    std::vector<TextureDev> textures;        // texture::bitmap assets
    std::vector<float4> texels;              // ... their texels (rgba; alpha = 1 without a channel)
    std::vector<MixEntry> mix;               // entries of the constant-weight mixtures
    std::vector<float> mesh_t;               // [total faces][3][2] texture coordinates; empty when no material reads a texture
    CameraDev cam{};
    bool has_camera = false;
    int env_light = -1;
    uint32_t nt = 0;
};

inline void mat3_normal_matrix(const double* M, double* NM, double* J, float* sign) {
    const double a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], i = M[10];
    const double det = a * (e * i - f * h) - d * (b * i - c * h) + g * (b * f - c * e);
    const double id = 1.0 / det;
    // inverse-transpose = cofactor matrix / det (column-major: NM[col*3+row]); math.h:742-745
    NM[0] = (e * i - f * h) * id; NM[1] = -(d * i - f * g) * id; NM[2] = (d * h - e * g) * id;
    NM[3] = -(b * i - c * h) * id; NM[4] = (a * i - c * g) * id; NM[5] = -(a * h - b * g) * id;
    NM[6] = (b * f - c * e) * id; NM[7] = -(a * f - c * d) * id; NM[8] = (a * e - b * d) * id;
    *J = std::fabs(det);
    *sign = det < 0 ? -1.f : 1.f;
}

// Returns an LMCU_* status; err receives the message.
inline int prepare_scene(const LmcuSceneDesc* d, PreparedScene& o, std::string& err) {
    if (d->num_primitives <= 0 || d->num_meshes <= 0) { err = "Missing primitives."; return LMCU_ERR_UNSUPPORTED; }
    o.face_offset.assign(d->num_meshes + 1, 0);
    for (int i = 0; i < d->num_meshes; i++) {
        if (!d->meshes[i].p) { err = "mesh without positions"; return LMCU_ERR_INVALID_ARGUMENT; }
        o.face_offset[i + 1] = o.face_offset[i] + (uint32_t)d->meshes[i].num_faces;
    }
    o.prims.assign(d->num_primitives, PrimDev{});
    uint64_t nt = 0;
    for (int i = 0; i < d->num_primitives; i++) {
        const LmcuPrimitive& p = d->primitives[i];
        if (p.mesh < 0 || p.mesh >= d->num_meshes) { err = "primitive: bad mesh index"; return LMCU_ERR_INVALID_ARGUMENT; }
        if (p.material < 0 || p.material >= d->num_materials) { err = "primitive: bad material index"; return LMCU_ERR_INVALID_ARGUMENT; }
        if (p.light >= d->num_lights) { err = "primitive: bad light index"; return LMCU_ERR_INVALID_ARGUMENT; }
        PrimDev& q = o.prims[i];
        for (int k = 0; k < 16; k++) q.M[k] = p.M[k];
        mat3_normal_matrix(p.M, q.NM, &q.J, &q.det_sign);
        q.node_index = p.node_index; q.mesh = p.mesh; q.material = p.material; q.light = p.light;
        q.tri_offset = (uint32_t)nt;
        q.num_faces = (uint32_t)d->meshes[p.mesh].num_faces;
        q.mesh_face_offset = o.face_offset[p.mesh];
        q.has_normals = d->meshes[p.mesh].n ? 1 : 0;
        q.inv_area = 0.f;
        nt += q.num_faces;
    }
    if (nt == 0) { err = "Missing primitives."; return LMCU_ERR_UNSUPPORTED; }
    if (nt > 0x7fffffffull) { err = "more than 2^31 triangles"; return LMCU_ERR_UNSUPPORTED; }
    o.nt = (uint32_t)nt;
    // textures (texture::bitmap, src/texture/texture_bitmap.cpp:41-96)
    if (d->num_textures < 0 || (d->num_textures > 0 && !d->textures)) { err = "bad texture table"; return LMCU_ERR_INVALID_ARGUMENT; }
    o.textures.resize(d->num_textures);
    for (int i = 0; i < d->num_textures; i++) {
        const LmcuTexture& t = d->textures[i];
        if (t.w <= 0 || t.h <= 0 || (t.c != 3 && t.c != 4) || !t.data) { err = "texture: need w, h > 0, 3 or 4 channels and texel data"; return LMCU_ERR_INVALID_ARGUMENT; }
        TextureDev& q = o.textures[i];
        q.w = t.w; q.h = t.h; q.offset = (int32_t)o.texels.size(); q.has_alpha = t.c == 4 ? 1 : 0;
        o.texels.reserve(o.texels.size() + (size_t)t.w * t.h);
        for (size_t k = 0; k < (size_t)t.w * t.h; k++) {
            const float* c = t.data + k * t.c;
            o.texels.push_back(make_float4(c[0], c[1], c[2], t.c == 4 ? c[3] : 1.f));
        }
    }
    o.mats.resize(d->num_materials);
    bool reads_texture = false;
    for (int i = 0; i < d->num_materials; i++) {
        const LmcuMaterial& m = d->materials[i];
        if (m.type < 0 || m.type > LMCU_MAT_MIXTURE_MARGINALIZED) { err = "bad material type"; return LMCU_ERR_INVALID_ARGUMENT; }
        MaterialDev& q = o.mats[i];
        memset(&q, 0, sizeof(q));
        q.type = m.type;
        q.r = (float)m.color[0]; q.g = (float)m.color[1]; q.b = (float)m.color[2];
        q.ax = (float)m.ax; q.ay = (float)m.ay; q.Ni = (float)m.Ni;
        q.tex_kd = -1;
        if (m.map_kd != 0 && (m.type == LMCU_MAT_DIFFUSE || m.type == LMCU_MAT_MIXTURE_OBJ)) {
            if (m.map_kd < 0 || m.map_kd > d->num_textures) { err = "material: bad mapKd texture index"; return LMCU_ERR_INVALID_ARGUMENT; }
            q.tex_kd = m.map_kd - 1;
            reads_texture = true;
        }
        if (m.type == LMCU_MAT_MIXTURE_OBJ) {
            q.ks_r = (float)m.Ks[0]; q.ks_g = (float)m.Ks[1]; q.ks_b = (float)m.Ks[2];
            // model_wavefrontobj.cpp:405-412: the mask exists only when the texture has an alpha channel
            q.alpha_mask = (m.alpha_mask && q.tex_kd >= 0 && o.textures[q.tex_kd].has_alpha) ? 1 : 0;
        }
    }
    // constant-weight mixtures (src/material/material_mixture.cpp:51-59,156-186): entries, Dist::norm() pmfs
    for (int i = 0; i < d->num_materials; i++) {
        const LmcuMaterial& m = d->materials[i];
        if (m.type != LMCU_MAT_MIXTURE_RR && m.type != LMCU_MAT_MIXTURE_MARGINALIZED) continue;
        if (m.num_entries <= 0 || !m.entry_material || !m.entry_weight) { err = "mixture material without entries"; return LMCU_ERR_INVALID_ARGUMENT; }
        MaterialDev& q = o.mats[i];
        q.entry_offset = (int32_t)o.mix.size();
        q.num_entries = m.num_entries;
        double sum = 0.0, sum0 = 0.0;
        for (int k = 0; k < m.num_entries; k++) {
            const int e = m.entry_material[k];
            if (e < 0 || e >= d->num_materials) { err = "mixture entry: bad material index"; return LMCU_ERR_INVALID_ARGUMENT; }
            const int et = d->materials[e].type;
            if (et == LMCU_MAT_GLASS || et > LMCU_MAT_MASK) {
                // the reference hands a mixture entry component 0 whatever the entry's own components are
                // (material_mixture.cpp:67-78): only single-component entries are meaningful
                err = "mixture entries must be diffuse, glossy, mirror or mask"; return LMCU_ERR_UNSUPPORTED;
            }
            if (!(m.entry_weight[k] > 0.0)) { err = "mixture entry: weight must be positive"; return LMCU_ERR_INVALID_ARGUMENT; }
            sum += m.entry_weight[k];
            if (!mat_is_specular(et)) sum0 += m.entry_weight[k];
        }
        if (m.type == LMCU_MAT_MIXTURE_RR) {
            for (int k = 0; k < m.num_entries; k++) {
                MixEntry e;
                e.material = m.entry_material[k]; e.weight = (float)m.entry_weight[k];
                e.prob = (float)(m.entry_weight[k] / sum); e.group_prob = e.prob;
                o.mix.push_back(e);
            }
        } else {
            // group 0 = every non-specular entry, then one group per specular entry, in entry order
            int n0 = 0;
            for (int pass = 0; pass < 2; pass++) {
                for (int k = 0; k < m.num_entries; k++) {
                    const bool spec = mat_is_specular(d->materials[m.entry_material[k]].type);
                    if (spec != (pass == 1)) continue;
                    MixEntry e;
                    e.material = m.entry_material[k]; e.weight = (float)m.entry_weight[k];
                    e.prob = spec ? 1.f : (float)(m.entry_weight[k] / sum0);
                    e.group_prob = (float)((spec ? m.entry_weight[k] : sum0) / sum);
                    o.mix.push_back(e);
                    if (!spec) n0++;
                }
            }
            q.num_group0 = n0;
        }
    }
    if (reads_texture) {
        o.mesh_t.assign((size_t)o.face_offset.back() * 6, 0.f);
        for (int i = 0; i < d->num_meshes; i++) {
            if (!d->meshes[i].t) continue;
            const size_t nf = (size_t)d->meshes[i].num_faces;
            for (size_t k = 0; k < nf * 6; k++) o.mesh_t[(size_t)o.face_offset[i] * 6 + k] = (float)d->meshes[i].t[k];
        }
    }
    o.lights.resize(d->num_lights);
    o.env_light = -1;
    for (int i = 0; i < d->num_lights; i++) {
        const LmcuLight& l = d->lights[i];
        LightDev& t = o.lights[i];
        memset(&t, 0, sizeof(t));
        t.type = l.type; t.prim = l.primitive;
        for (int k = 0; k < 3; k++) t.Ke[k] = (float)l.Ke[k];
        if (l.type == LMCU_LIGHT_AREA) {
            if (l.primitive < 0 || l.primitive >= d->num_primitives) { err = "area light: bad primitive"; return LMCU_ERR_INVALID_ARGUMENT; }
            PrimDev& q = o.prims[l.primitive];
            const LmcuMesh& m = d->meshes[q.mesh];
            t.tri_offset = (int32_t)o.emit.size(); t.num_tris = m.num_faces; t.cdf_offset = (int32_t)o.cdf.size();
            std::vector<double> c(1, 0.0);
            for (int f = 0; f < m.num_faces; f++) {
                const double* p = m.p + (size_t)f * 9;
                const double e1[3] = { p[3] - p[0], p[4] - p[1], p[5] - p[2] }, e2[3] = { p[6] - p[0], p[7] - p[1], p[8] - p[2] };
                const double cr[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
                const double len = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                c.push_back(c.back() + len * 0.5);
                EmitTri et;
                float4* dst[3] = { &et.p1, &et.p2, &et.p3 };
                for (int v = 0; v < 3; v++) {
                    const D3 w = transform_point(q.M, d3(p[3 * v], p[3 * v + 1], p[3 * v + 2]));
                    *dst[v] = make_float4((float)w.x, (float)w.y, (float)w.z, 0.f);
                }
                const double gl[3] = { cr[0] / len, cr[1] / len, cr[2] / len };
                const double gw[3] = { q.NM[0] * gl[0] + q.NM[3] * gl[1] + q.NM[6] * gl[2], q.NM[1] * gl[0] + q.NM[4] * gl[1] + q.NM[7] * gl[2],
                                       q.NM[2] * gl[0] + q.NM[5] * gl[1] + q.NM[8] * gl[2] };
                const double gwl = std::sqrt(gw[0] * gw[0] + gw[1] * gw[1] + gw[2] * gw[2]);
                et.n = make_float4((float)(gw[0] / gwl), (float)(gw[1] / gwl), (float)(gw[2] / gwl), 0.f);
                o.emit.push_back(et);
            }
            const double sum = c.back();
            for (double v : c) o.cdf.push_back((float)(v / sum));
            q.inv_area = (float)(1.0 / sum);
            t.inv_area_j = (float)((1.0 / sum) / q.J);
        } else if (l.type == LMCU_LIGHT_ENVCONST) {
            if (o.env_light >= 0) { err = "Environment light is already registered."; return LMCU_ERR_INVALID_ARGUMENT; }
            o.env_light = i;
        } else if (l.type == LMCU_LIGHT_ENV) {
            // light::env (src/light/light_env.cpp:52-77): lat-long image + Dist2 over max(rgb*scale) * sin(theta)
            if (o.env_light >= 0) { err = "Environment light is already registered."; return LMCU_ERR_INVALID_ARGUMENT; }
            if (!l.env_rgb || l.env_w <= 0 || l.env_h <= 0) { err = "light::env without image data"; return LMCU_ERR_INVALID_ARGUMENT; }
            o.env_light = i;
            const int w = l.env_w, h = l.env_h;
            t.env_w = w; t.env_h = h;
            t.env_rot = (float)(l.env_rot * 3.14159265358979323846 / 180.0);      // glm::radians, light_env.cpp:63
            t.env_scale = (float)l.env_scale;
            t.env_tex_offset = (int32_t)o.env_tex.size();
            t.env_cdf_offset = (int32_t)o.cdf.size();
            o.env_tex.reserve(o.env_tex.size() + (size_t)w * h);
            for (size_t k = 0; k < (size_t)w * h; k++) {
                o.env_tex.push_back(make_float4(l.env_rgb[3 * k], l.env_rgb[3 * k + 1], l.env_rgb[3 * k + 2], 0.f));
            }
            // Dist2::init (include/lm/math.h:398-410) in FP64, stored as float CDFs: rows [h][w+1], marginal [h+1]
            std::vector<double> marg(1, 0.0), row((size_t)w + 1);
            o.cdf.reserve(o.cdf.size() + (size_t)h * (w + 1) + h + 1);
            for (int y = 0; y < h; y++) {
                row[0] = 0.0;
                for (int x = 0; x < w; x++) {
                    const size_t k = (size_t)y * w + x;
                    const double mx = std::fmax(std::fmax((double)l.env_rgb[3 * k], (double)l.env_rgb[3 * k + 1]), (double)l.env_rgb[3 * k + 2]) * l.env_scale;
                    // the reference weights texel i by sin(pi * (Float(i)/w + .5) / h) with i the FLAT index (light_env.cpp:71)
                    const double v = mx * std::sin(3.14159265358979323846 * ((double)k / w + .5) / h);
                    row[x + 1] = row[x] + v;
                }
                marg.push_back(marg.back() + row[w]);
                for (int x = 0; x <= w; x++) o.cdf.push_back((float)(row[x] / row[w]));
            }
            for (double v : marg) o.cdf.push_back((float)(v / marg.back()));
        } else if (l.type == LMCU_LIGHT_POINT) {                 // light_point.cpp:38-41
            for (int k = 0; k < 3; k++) t.pos[k] = (float)l.position[k];
        } else if (l.type == LMCU_LIGHT_DIRECTIONAL) {           // light_directional.cpp:37-40: glm::normalize(direction)
            const double len = std::sqrt(l.position[0] * l.position[0] + l.position[1] * l.position[1] + l.position[2] * l.position[2]);
            if (!(len > 0.0)) { err = "light::directional with a zero direction"; return LMCU_ERR_INVALID_ARGUMENT; }
            for (int k = 0; k < 3; k++) t.pos[k] = (float)(l.position[k] / len);
        } else {
            err = "bad light type"; return LMCU_ERR_INVALID_ARGUMENT;
        }
    }
    o.has_camera = d->camera != nullptr;
    if (d->camera) {
        const LmcuCamera& c = *d->camera;
        CameraDev& cam = o.cam;
        const D3 pos = d3(c.position[0], c.position[1], c.position[2]);
        const D3 wv = dsub3(pos, d3(c.center[0], c.center[1], c.center[2]));
        const double wi = divrn(1.0, sqrtrn(ddot(wv, wv)));
        const D3 w = d3(mulrn(wv.x, wi), mulrn(wv.y, wi), mulrn(wv.z, wi));
        const D3 uv = dcross(d3(c.up[0], c.up[1], c.up[2]), w);
        const double ui = divrn(1.0, sqrtrn(ddot(uv, uv)));
        const D3 u = d3(mulrn(uv.x, ui), mulrn(uv.y, ui), mulrn(uv.z, ui));
        const D3 v = dcross(w, u);
        cam.pos[0] = pos.x; cam.pos[1] = pos.y; cam.pos[2] = pos.z;
        cam.u[0] = u.x; cam.u[1] = u.y; cam.u[2] = u.z;
        cam.v[0] = v.x; cam.v[1] = v.y; cam.v[2] = v.z;
        cam.w[0] = w.x; cam.w[1] = w.y; cam.w[2] = w.z;
        cam.tf = std::tan(c.vfov * 3.14159265358979323846 / 180.0 * .5);
        cam.aspect = c.aspect;
    }
    return LMCU_OK;
}

// -- two-level layout: host-side pieces shared by lmcu.cu and tests/hostsim ------------------------------------------------
// The instance record of primitive `prim` whose mesh tree has its root at node `root`.  False when the global transform
// is not an invertible affine map (then the scene is built flattened).
inline bool make_inst(const PrimDev& pr, uint32_t prim, uint32_t root, InstDev& o) {
    const double* M = pr.M;
    if (M[3] != 0.0 || M[7] != 0.0 || M[11] != 0.0 || M[15] != 1.0) return false;
    if (!(pr.J > 0.0) || !std::isfinite(pr.J)) return false;
    for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) o.M[3 * c + r] = M[4 * c + r];
    // inverse of the 3x3 part = transpose of the normal matrix (NM = inverse-transpose, column-major)
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) {
        o.Minv[3 * c + r] = pr.NM[3 * r + c];
        if (!std::isfinite(o.Minv[3 * c + r])) return false;
    }
    for (int r = 0; r < 3; r++) o.Minv[9 + r] = -(o.Minv[r] * M[12] + o.Minv[3 + r] * M[13] + o.Minv[6 + r] * M[14]);
    o.root = root; o.tri_offset = pr.tri_offset; o.prim = prim; o.pad_ = 0;
    return true;
}

// LmcuBuildOptions::two_level resolved against the scene.  The flattened tree is the faster one to walk (C4: 477 against 370
// Mpaths/s — the two-level walk pays an object-space ray per instance entered and a vertex transform per leaf test), so the
// automatic choice keeps it while it fits: two-level only for an instanced scene (flattened triangles >= 4 x the unique ones)
// whose flattened build (~475 bytes per triangle at its peak: 155 resident + the builder's temporaries) would take more than
// half of `budget_bytes` (the device's free memory).
// Step 1 (no device query): -1 = flattened, 1 = two-level, 0 = an instanced scene under the automatic option — the caller then
// decides with two_level_by_budget().
inline int two_level_candidate(int option, const std::vector<PrimDev>& prims, const std::vector<uint32_t>& face_offset, uint32_t nt) {
    if (option < 0) return -1;
    std::vector<char> used(face_offset.size(), 0);
    uint64_t unique = 0;
    InstDev tmp;
    for (size_t i = 0; i < prims.size(); i++) {
        if (!make_inst(prims[i], (uint32_t)i, 0u, tmp)) return -1;
        if (!used[prims[i].mesh]) { used[prims[i].mesh] = 1; unique += prims[i].num_faces; }
    }
    if (option > 0) return 1;
    return (uint64_t)nt >= 4 * unique ? 0 : -1;
}
inline bool two_level_by_budget(uint32_t nt, double budget_bytes) { return budget_bytes > 0.0 && 475.0 * (double)nt > 0.5 * budget_bytes; }

inline bool want_two_level(int option, const std::vector<PrimDev>& prims, const std::vector<uint32_t>& face_offset, uint32_t nt,
                           double budget_bytes = 0.0) {
    const int c = two_level_candidate(option, prims, face_offset, nt);
    return c > 0 || (c == 0 && two_level_by_budget(nt, budget_bytes));
}

}  // namespace lmcu
