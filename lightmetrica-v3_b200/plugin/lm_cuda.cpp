// lm_cuda.cpp — Lightmetrica v3 plugin: the drop-in side of the CUDA hot path.
//
// Registers NEW components through the reference's own registry (include/lm/component.h:816-817):
//     accel::cuda_bvh        lm::Accel      build / intersect            (include/lm/accel.h:38,70)
//     renderer::raycast_cuda lm::Renderer   same JSON keys as renderer::raycast (renderer_raycast.cpp:48-60)
//     renderer::ao_cuda      lm::Renderer   same keys as renderer::ao     (functest/renderer_ao.cpp:18-22)
//     renderer::pt_cuda      lm::Renderer   same keys as renderer::pt     (renderer_pt.cpp:51-77)
//     renderer::lt_cuda      lm::Renderer   same keys as renderer::lt     (renderer_lt.cpp:37-48)
//     film::cuda             lm::Film       the fp32 film the renderers write with one copy
// It is loaded with lm::comp::load_plugin("<dir>/lm_cuda") exactly like accel_nanort / accel_embree.
//
// This file is HOST glue only: it pulls geometry, transforms, materials, lights and the camera out of
// lm::Scene through public interfaces and hands plain arrays to liblmcu.so over the C ABI of
// include/lmcu.h.  All ray tracing happens on the GPU; there is no CPU fallback (a missing device or
// library surfaces as lm::Exception).
#include <lm/lm.h>

#include <cstdio>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "../../include/lmcu.h"

LM_NAMESPACE_BEGIN(LM_NAMESPACE)

namespace {

void check(int rc) {
    if (rc == LMCU_OK) return;
    const std::string msg = lmcu_last_error();
    switch (rc) {
    case LMCU_ERR_INVALID_ARGUMENT: LM_THROW_EXCEPTION(Error::InvalidArgument, "lmcu: {}", msg);
    case LMCU_ERR_UNSUPPORTED: LM_THROW_EXCEPTION(Error::Unsupported, "lmcu: {}", msg);
    default: LM_THROW_EXCEPTION(Error::None, "lmcu: {}", msg);
    }
}

// Everything liblmcu needs, gathered from an lm::Scene.  Owns the arrays the LmcuSceneDesc points to.
struct SceneExport {
    std::vector<std::vector<double>> mesh_p, mesh_n, mesh_t;
    std::vector<LmcuMesh> meshes;
    std::vector<std::vector<float>> tex_data;
    std::vector<LmcuTexture> textures;
    std::map<const Texture*, int> tex_ids;
    std::map<const Material*, int> mat_ids;
    std::vector<std::vector<int32_t>> entry_material;     // per mixture material (kept alive for LmcuMaterial's pointers)
    std::vector<std::vector<double>> entry_weight;
    std::vector<LmcuPrimitive> prims;
    std::vector<LmcuMaterial> mats;
    std::vector<LmcuLight> lights;
    LmcuCamera camera{};
    bool has_camera = false;
    std::map<int, Transform> transforms;   // SceneNode::index -> Transform for Accel::Hit
    LmcuSceneDesc desc{};

    // Material parameters are private in the reference; recover them through the public virtuals
    // (SURVEY.md §7 "hard parts"): key() gives the type, reflectance() gives Kd / Ks, GGX roughness
    // follows from pdf_direction() at three half vectors, the index of refraction from the Fresnel
    // reflectance at normal incidence returned by pdf_component().
    // texture::bitmap keeps its texels behind Texture::buffer() (src/texture/texture_bitmap.cpp:92-94); any other texture
    // is sampled through eval() into a 1x1 image.  Returns 1 + index into textures[].
    int register_texture(Texture* tex) {
        auto it = tex_ids.find(tex);
        if (it != tex_ids.end()) return it->second + 1;
        const auto buf = tex->buffer();
        LmcuTexture t{};
        tex_data.emplace_back();
        auto& data = tex_data.back();
        if (buf.data && (buf.c == 3 || buf.c == 4)) {
            t.w = buf.w; t.h = buf.h; t.c = buf.c;
            data.assign(buf.data, buf.data + size_t(buf.w) * buf.h * buf.c);
        }
        else if (buf.data && buf.c >= 1) {                                  // grey (+alpha) images: expand to rgb(a)
            const bool alpha = buf.c == 2;
            t.w = buf.w; t.h = buf.h; t.c = alpha ? 4 : 3;
            for (size_t i = 0; i < size_t(buf.w) * buf.h; i++) {
                for (int k = 0; k < 3; k++) data.push_back(buf.data[buf.c * i + std::min(k, buf.c - 1 - (alpha ? 1 : 0))]);
                if (alpha) data.push_back(buf.data[buf.c * i + 1]);
            }
        }
        else {
            const Vec3 c = tex->eval(Vec2(.5_f));
            t.w = t.h = 1; t.c = 3;
            data = { float(c.x), float(c.y), float(c.z) };
        }
        tex_ids.emplace(tex, int(textures.size()));
        textures.push_back(t);                                              // data pointers are fixed up in gather()
        return int(textures.size());
    }

    // Material parameters are private in the reference; recover them through the public virtuals
    // (SURVEY.md §7 "hard parts"): key() gives the type, reflectance() gives Kd / Ks, GGX roughness
    // follows from pdf_direction() at three half vectors, the index of refraction from the Fresnel
    // reflectance at normal incidence returned by pdf_component(); composite materials expose their parts through
    // underlying() / foreach_underlying() and their weights through eval(), pdf_component() and sample_direction().
    // Returns the index into mats[].
    int material_index(const Material* m) {
        auto it = mat_ids.find(m);
        if (it != mat_ids.end()) return it->second;
        const LmcuMaterial o = probe_material(m);                           // may register entries first
        mat_ids.emplace(m, int(mats.size()));
        mats.push_back(o);
        return int(mats.size()) - 1;
    }

    LmcuMaterial probe_material(const Material* m) {
        LmcuMaterial o{};
        const auto geom = PointGeometry::make_on_surface(Vec3(0_f), Vec3(0_f, 0_f, 1_f), Vec3(0_f, 0_f, 1_f));
        auto key = m->key();
        if (key == "material::proxy") {                                     // src/material/material_proxy.cpp: pure forwarding
            const Material* ref = nullptr;
            const_cast<Material*>(m)->foreach_underlying([&](Component*& c, bool) { ref = dynamic_cast<const Material*>(c); });
            if (!ref) LM_THROW_EXCEPTION(Error::Unsupported, "material::proxy without a referenced material");
            return probe_material(ref);
        }
        if (key == "material::diffuse") {
            o.type = LMCU_MAT_DIFFUSE;
            auto* tex = dynamic_cast<Texture*>(m->underlying("mapKd"));
            if (tex) {
                o.map_kd = register_texture(tex);
                o.color[0] = o.color[1] = o.color[2] = 1.0;
            }
            else {
                const Vec3 Kd = m->reflectance(geom);
                o.color[0] = Kd.x; o.color[1] = Kd.y; o.color[2] = Kd.z;
            }
        }
        else if (key == "material::glossy") {
            o.type = LMCU_MAT_GLOSSY;
            const Vec3 Ks = m->reflectance(geom);
            o.color[0] = Ks.x; o.color[1] = Ks.y; o.color[2] = Ks.z;
            // pdf(wi=wo=wh) = D(wh)/4 (material_glossy.cpp:92-99), D from :43-46
            const auto D = [&](Vec3 w) { return 4_f * m->pdf_direction(geom, w, w, 0, false); };
            const auto P = 1_f / (Pi * D(geom.n));                          // ax*ay
            const auto th = .5_f;
            const auto s2 = math::sq(std::sin(th)), c2 = math::sq(std::cos(th));
            const auto solve = [&](Vec3 t) {
                const auto w = glm::normalize(t * std::sin(th) + geom.n * std::cos(th));
                const auto q = std::sqrt(1_f / (Pi * P * D(w)));            // (sin^2/a^2 + cos^2)
                return std::sqrt(s2 / (q - c2));
            };
            o.ax = solve(geom.u);
            o.ay = solve(geom.v);
        }
        else if (key == "material::glass") {
            o.type = LMCU_MAT_GLASS;
            const auto R0 = m->pdf_component(0, geom, geom.n);              // Schlick at normal incidence
            const auto r = std::sqrt(R0);
            o.Ni = (1_f + r) / (1_f - r);
        }
        else if (key == "material::mirror") {
            o.type = LMCU_MAT_MIRROR;
        }
        else if (key == "material::mask") {
            o.type = LMCU_MAT_MASK;
        }
        else if (key == "material::mixture_wavefrontobj") {
            // src/model/model_wavefrontobj.cpp:339-515: owns a material::diffuse ("diffuse") and a material::glossy ("glossy")
            o.type = LMCU_MAT_MIXTURE_OBJ;
            const auto* md = dynamic_cast<const Material*>(m->underlying("diffuse"));
            const auto* mg = dynamic_cast<const Material*>(m->underlying("glossy"));
            if (!md || !mg) LM_THROW_EXCEPTION(Error::Unsupported, "material::mixture_wavefrontobj without its underlying lobes");
            const LmcuMaterial d = probe_material(md), g = probe_material(mg);
            for (int k = 0; k < 3; k++) { o.color[k] = d.color[k]; o.Ks[k] = g.color[k]; }
            o.map_kd = d.map_kd;
            o.ax = g.ax; o.ay = g.ay;
            // the alpha mask (mask_tex_, private) shows in pdf_component(0) = alpha at a texel whose alpha is not 1
            o.alpha_mask = 0;
            if (o.map_kd > 0 && textures[o.map_kd - 1].c == 4) {
                const auto& t = textures[o.map_kd - 1];
                const auto& data = tex_data[o.map_kd - 1];
                for (size_t i = 0; i < size_t(t.w) * t.h; i++) {
                    if (data[4 * i + 3] == 1.f) continue;
                    const Vec2 uv((Float(i % t.w) + .5_f) / t.w, (Float(i / t.w) + .5_f) / t.h);
                    const auto gt = PointGeometry::make_on_surface(Vec3(0_f), Vec3(0_f, 0_f, 1_f), Vec3(0_f, 0_f, 1_f), uv);
                    o.alpha_mask = m->pdf_component(0, gt, geom.n) != 1_f ? 1 : 0;
                    break;
                }
            }
        }
        else if (key == "material::constant_weight_mixture_rr" || key == "material::constant_weight_mixture_marginalized") {
            probe_mixture(m, key == "material::constant_weight_mixture_rr", o);
        }
        else {
            LM_THROW_EXCEPTION(Error::Unsupported, "renderer::*_cuda: unsupported material [key='{}']", key);
        }
        return o;
    }

    // Constant-weight mixtures (src/material/material_mixture.cpp): the entries come out of foreach_underlying() in
    // construction order; their weights are private and recovered from the mixture's own answers:
    //   rr            component i IS entry i; sample_direction(comp i).weight = weight_i * (the entry's own sample weight)
    //   marginalized  component 0 = all non-specular entries with eval = sum_i weight_i f_i (solved by least squares over a
    //                 few direction pairs and the 3 colour channels), component g >= 1 = the g-th specular entry with
    //                 sample_direction(comp g).weight = weight
    void probe_mixture(const Material* m, bool rr, LmcuMaterial& o) {
        std::vector<const Material*> entries;
        const_cast<Material*>(m)->foreach_underlying([&](Component*& c, bool) {
            if (auto* e = dynamic_cast<const Material*>(c)) entries.push_back(e);
        });
        if (entries.empty()) LM_THROW_EXCEPTION(Error::Unsupported, "mixture material without entries");
        const auto geom = PointGeometry::make_on_surface(Vec3(0_f), Vec3(0_f, 0_f, 1_f), Vec3(0_f, 0_f, 1_f));
        const Vec3 wi = glm::normalize(Vec3(.3_f, .2_f, 1_f));
        const Material::DirectionSampleU us{ Vec2(.37_f, .61_f), Vec2(.5_f) };
        const auto max_ratio = [](Vec3 a, Vec3 b) {                         // a = weight * b, from the largest channel of b
            int k = 0;
            for (int c = 1; c < 3; c++) if (std::abs(b[c]) > std::abs(b[k])) k = c;
            if (b[k] == 0_f) LM_THROW_EXCEPTION(Error::Unsupported, "mixture entry with a zero sample weight: its weight cannot be recovered");
            return a[k] / b[k];
        };
        const auto sampled_weight = [&](int comp, const Material* entry) {
            const auto sm = m->sample_direction(us, geom, wi, comp, Material::TransDir::EL);
            const auto se = entry->sample_direction(us, geom, wi, 0, Material::TransDir::EL);
            if (!sm || !se) LM_THROW_EXCEPTION(Error::Unsupported, "mixture entry could not be sampled for probing");
            return max_ratio(sm->weight, se->weight);
        };
        std::vector<int32_t> ids;
        std::vector<double> ws(entries.size(), 0.0);
        for (const auto* e : entries) ids.push_back(material_index(e));
        if (rr) {
            o.type = LMCU_MAT_MIXTURE_RR;
            for (size_t i = 0; i < entries.size(); i++) ws[i] = sampled_weight(int(i), entries[i]);
        }
        else {
            o.type = LMCU_MAT_MIXTURE_MARGINALIZED;
            std::vector<size_t> diffuse_like;
            int group = 0;
            for (size_t i = 0; i < entries.size(); i++) {
                if (entries[i]->is_specular_component(0)) ws[i] = sampled_weight(++group, entries[i]);
                else diffuse_like.push_back(i);
            }
            const size_t n = diffuse_like.size();
            if (n == 1) {
                const Vec3 wo = glm::normalize(Vec3(-.2_f, .4_f, 1_f));
                ws[diffuse_like[0]] = max_ratio(m->eval(geom, wi, wo, 0, Material::TransDir::EL, false), entries[diffuse_like[0]]->eval(geom, wi, wo, 0, Material::TransDir::EL, false));
            }
            else if (n > 1) {
                // normal equations A^T A w = A^T b with a tiny ridge (entries with identical lobes are interchangeable)
                std::vector<double> ata(n * n, 0.0), atb(n, 0.0);
                for (int k = 0; k < 12; k++) {
                    const Float a = .4_f + .9_f * k, b = .15_f + .1_f * k;
                    const Vec3 w1 = glm::normalize(Vec3(std::cos(a) * std::sin(b), std::sin(a) * std::sin(b), std::cos(b)));
                    const Vec3 w2 = glm::normalize(Vec3(std::cos(2.1_f * a) * std::sin(1.3_f - b), std::sin(2.1_f * a) * std::sin(1.3_f - b), std::cos(1.3_f - b)));
                    const Vec3 fm = m->eval(geom, w1, w2, 0, Material::TransDir::EL, false);
                    std::vector<Vec3> fe(n);
                    for (size_t i = 0; i < n; i++) fe[i] = entries[diffuse_like[i]]->eval(geom, w1, w2, 0, Material::TransDir::EL, false);
                    for (int c = 0; c < 3; c++) {
                        for (size_t i = 0; i < n; i++) {
                            atb[i] += fe[i][c] * fm[c];
                            for (size_t j = 0; j < n; j++) ata[i * n + j] += fe[i][c] * fe[j][c];
                        }
                    }
                }
                double tr = 0.0;
                for (size_t i = 0; i < n; i++) tr += ata[i * n + i];
                for (size_t i = 0; i < n; i++) ata[i * n + i] += 1e-12 * tr;
                for (size_t c = 0; c < n; c++) {                            // Gaussian elimination with partial pivoting
                    size_t piv = c;
                    for (size_t r = c + 1; r < n; r++) if (std::abs(ata[r * n + c]) > std::abs(ata[piv * n + c])) piv = r;
                    for (size_t j = 0; j < n; j++) std::swap(ata[c * n + j], ata[piv * n + j]);
                    std::swap(atb[c], atb[piv]);
                    for (size_t r = c + 1; r < n; r++) {
                        const double f = ata[r * n + c] / ata[c * n + c];
                        for (size_t j = c; j < n; j++) ata[r * n + j] -= f * ata[c * n + j];
                        atb[r] -= f * atb[c];
                    }
                }
                for (size_t c = n; c-- > 0;) {
                    double v = atb[c];
                    for (size_t j = c + 1; j < n; j++) v -= ata[c * n + j] * ws[diffuse_like[j]];
                    ws[diffuse_like[c]] = v / ata[c * n + c];
                }
            }
        }
        entry_material.push_back(ids);
        entry_weight.push_back(ws);
        o.num_entries = int(ids.size());
        o.entry_material = reinterpret_cast<const int32_t*>(entry_material.size());   // slot + 1; turned into pointers in gather()
    }

    // light::env keeps its image as an underlying texture::bitmap ("envmap", src/light/light_env.cpp:42-47) but
    // rot_ and scale_ are private.  Both follow from ONE deterministic sample_direct() call: the Dist2 the light
    // samples from is invariant to scale_, so the texel (x, y) it picks for u = (.5,.5,.5,.5) can be recomputed
    // here from the raw image with the reference's own arithmetic (light_env.cpp:64-72, math.h:398-431); the
    // azimuth of the returned direction minus that texel's azimuth is rot_, and eval() over the raw texel is scale_.
    std::vector<float> env_rgb;
    void probe_env_light(const Light* light, LmcuLight& l) {
        auto* tex = dynamic_cast<Texture*>(light->underlying("envmap"));
        if (!tex) LM_THROW_EXCEPTION(Error::Unsupported, "light::env without an underlying envmap texture");
        const auto buf = tex->buffer();
        if (!buf.data || buf.c < 3) LM_THROW_EXCEPTION(Error::Unsupported, "light::env: envmap must have >= 3 channels");
        const int w = buf.w, h = buf.h;
        env_rgb.resize(size_t(w) * h * 3);
        for (size_t i = 0; i < size_t(w) * h; i++) for (int k = 0; k < 3; k++) env_rgb[3 * i + k] = buf.data[buf.c * i + k];
        std::vector<Float> ls(size_t(w) * h);
        for (int i = 0; i < w * h; i++) {
            const Float mx = std::max({ Float(env_rgb[3 * i]), Float(env_rgb[3 * i + 1]), Float(env_rgb[3 * i + 2]) });
            ls[i] = mx * std::sin(Pi * (Float(i) / w + .5_f) / h);
        }
        Dist2 dist;
        dist.init(ls, w, h);
        const auto geom = PointGeometry::make_on_surface(Vec3(0_f), Vec3(0_f, 1_f, 0_f), Vec3(0_f, 1_f, 0_f));
        Light::RaySampleU u{ Vec2(.5_f), .5_f, Vec2(.5_f) };
        Float rot = 0_f, scale = 1_f;
        bool found = false;
        // a few (ud) choices in case the sampled direction is exactly tangent to the probe surface (pdf 0 -> no sample)
        for (Float a : { .5_f, .25_f, .75_f, .125_f }) {
            u.ud = Vec2(a, .5_f);
            const auto sL = light->sample_direct(u, geom, Transform(Mat4(1_f)));
            if (!sL) continue;
            const Vec3 d = -sL->wo;
            const int y = dist.m.sample(u.ud[0]);
            const int x = dist.ds[y].sample(u.ud[1]);
            const Float phi_tex = 2_f * Pi * (x + u.up[0]) / w;
            rot = std::atan2(d.x, d.z) - phi_tex;
            rot -= 2_f * Pi * std::floor(rot / (2_f * Pi));
            const Vec3 Le = light->eval(sL->geom, sL->wo, false);
            const size_t i = size_t(y) * w + x;
            int k = 0;
            for (int c = 1; c < 3; c++) if (env_rgb[3 * i + c] > env_rgb[3 * i + k]) k = c;
            if (env_rgb[3 * i + k] > 0.f) { scale = Le[k] / Float(env_rgb[3 * i + k]); found = true; break; }
        }
        if (!found) LM_THROW_EXCEPTION(Error::Unsupported, "light::env: could not recover rot/scale");
        // cross-check on other texels: eval() at the texel-centre direction must return texel * scale
        for (int k = 1; k <= 7; k++) {
            const int x = (w * k) / 8, y = (h * k) / 8;
            const Float t = Pi * (y + .5_f) / h, p = 2_f * Pi * (x + .5_f) / w + rot;
            const Vec3 wo = -Vec3(std::sin(t) * std::sin(p), std::cos(t), std::sin(t) * std::cos(p));
            const Vec3 Le = light->eval(PointGeometry::make_infinite(wo), wo, false);
            const size_t i = size_t(y) * w + x;
            for (int c = 0; c < 3; c++) {
                const Float expect = Float(env_rgb[3 * i + c]) * scale;
                if (std::abs(Le[c] - expect) > 1e-6_f * std::max(1_f, std::abs(expect))) {
                    LM_THROW_EXCEPTION(Error::Unsupported, "light::env: recovered rot/scale fail the cross-check");
                }
            }
        }
        l.type = LMCU_LIGHT_ENV;
        l.env_w = w; l.env_h = h;
        l.env_rot = glm::degrees(rot);
        l.env_scale = scale;
        l.env_rgb = env_rgb.data();
    }

    // -- what identifies the exported scene ----------------------------------------------------------------
    // gather_description() collects everything but the triangles: primitives with their global transforms, materials
    // and lights as probed through the public virtuals, camera, textures, and per mesh its address, its triangle count
    // and a few sampled triangles.  A digest of all that (`sig`) decides whether a device scene uploaded earlier can be
    // used again (lm::Scene has no modification counter); only on a miss does gather_triangles() walk every mesh through
    // Mesh::foreach_triangle — one std::function call per triangle, the expensive part of an export.
    std::vector<const Mesh*> mesh_list;
    uint64_t sig = 1469598103934665603ull;
    bool shading_ = false;
    bool triangles_done = false;
    std::mutex tri_mu;

    void mix(const void* p, size_t n) {
        const auto* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; i++) { sig ^= b[i]; sig *= 1099511628211ull; }
    }
    template <typename T> void mix_value(const T& v) { mix(&v, sizeof(T)); }

    void compute_signature() {
        sig = 1469598103934665603ull;
        mix_value(shading_);
        for (const auto& p : prims) mix_value(p);
        for (const auto& m : mats) {
            mix_value(m.type); mix_value(m.map_kd); mix(m.color, sizeof(m.color)); mix_value(m.ax); mix_value(m.ay); mix_value(m.Ni);
            mix(m.Ks, sizeof(m.Ks)); mix_value(m.alpha_mask); mix_value(m.num_entries);
        }
        for (const auto& v : entry_material) mix(v.data(), v.size() * sizeof(int32_t));
        for (const auto& v : entry_weight) mix(v.data(), v.size() * sizeof(double));
        for (const auto& l : lights) {
            mix_value(l.type); mix_value(l.primitive); mix_value(l.node_index); mix_value(l.env_w); mix_value(l.env_h);
            mix(l.Ke, sizeof(l.Ke)); mix_value(l.env_rot); mix_value(l.env_scale); mix(l.position, sizeof(l.position));
        }
        mix(env_rgb.data(), env_rgb.size() * sizeof(float));
        mix_value(has_camera);
        if (has_camera) mix_value(camera);
        for (size_t i = 0; i < textures.size(); i++) {
            mix_value(textures[i].w); mix_value(textures[i].h); mix_value(textures[i].c);
            mix(tex_data[i].data(), tex_data[i].size() * sizeof(float));
        }
        for (const Mesh* mesh : mesh_list) {
            const int n = mesh->num_triangles();
            mix_value(mesh); mix_value(n);
            for (int k = 0; k < 5 && n > 0; k++) {
                const auto tri = mesh->triangle_at(int((long long)(n - 1) * k / 4));
                for (const auto* v : { &tri.p1, &tri.p2, &tri.p3 }) { mix_value(v->p); mix_value(v->n); mix_value(v->t); }
            }
        }
    }

    void gather_triangles() {
        std::unique_lock<std::mutex> lk(tri_mu);
        if (triangles_done) return;
        for (size_t i = 0; i < mesh_list.size(); i++) {
            const Mesh* mesh = mesh_list[i];
            auto& P = mesh_p[i];
            auto& N = mesh_n[i];
            auto& T = mesh_t[i];
            P.reserve(size_t(mesh->num_triangles()) * 9);
            N.reserve(size_t(mesh->num_triangles()) * 9);
            T.reserve(size_t(mesh->num_triangles()) * 6);
            mesh->foreach_triangle([&](int, const Mesh::Tri& tri) {
                for (const auto* v : { &tri.p1, &tri.p2, &tri.p3 }) {
                    P.push_back(v->p.x); P.push_back(v->p.y); P.push_back(v->p.z);
                    N.push_back(v->n.x); N.push_back(v->n.y); N.push_back(v->n.z);
                    T.push_back(v->t.x); T.push_back(v->t.y);
                }
            });
            meshes[i].num_faces = int(P.size() / 9);
            meshes[i].p = P.data();
            meshes[i].n = N.data();
            meshes[i].t = shading_ ? T.data() : nullptr;
        }
        triangles_done = true;
    }

    void gather(const Scene& scene, bool shading) {
        gather_description(scene, shading);
        gather_triangles();
    }

    void gather_description(const Scene& scene, bool shading) {
        shading_ = shading;
        std::map<const Mesh*, int> mesh_ids;
        scene.traverse_primitive_nodes([&](const SceneNode& node, Mat4 global_transform) {
            if (node.type != SceneNodeType::Primitive) {
                return;
            }
            if (node.primitive.light && !node.primitive.mesh) {
                if (!shading) return;
                // environment light (src/scene.cpp:103-110)
                LmcuLight l{};
                const auto key = node.primitive.light->key();
                l.primitive = -1;
                l.node_index = node.index;
                if (key == "light::envconst") {
                    l.type = LMCU_LIGHT_ENVCONST;
                    const auto Le = node.primitive.light->eval(PointGeometry::make_infinite(Vec3(0_f, 0_f, 1_f)), Vec3(0_f, 0_f, 1_f), false);
                    l.Ke[0] = Le.x; l.Ke[1] = Le.y; l.Ke[2] = Le.z;
                }
                else if (key == "light::env") {
                    probe_env_light(node.primitive.light, l);
                }
                else if (key == "light::point") {
                    // position through sample_position() (light_point.cpp:81-86), Le through eval()
                    l.type = LMCU_LIGHT_POINT;
                    const auto sP = node.primitive.light->sample_position({}, Transform(Mat4(1_f)));
                    const auto Le = node.primitive.light->eval(sP->geom, Vec3(0_f, 0_f, 1_f), false);
                    for (int k = 0; k < 3; k++) { l.position[k] = sP->geom.p[k]; l.Ke[k] = Le[k]; }
                }
                else if (key == "light::directional") {
                    // direction = the wo sample_direct() returns (light_directional.cpp:93-103), Le through eval()
                    l.type = LMCU_LIGHT_DIRECTIONAL;
                    const auto geom = PointGeometry::make_degenerated(Vec3(0_f));
                    const auto sL = node.primitive.light->sample_direct({}, geom, Transform(Mat4(1_f)));
                    const auto Le = node.primitive.light->eval(sL->geom, sL->wo, false);
                    for (int k = 0; k < 3; k++) { l.position[k] = sL->wo[k]; l.Ke[k] = Le[k]; }
                }
                else {
                    LM_THROW_EXCEPTION(Error::Unsupported, "renderer::*_cuda: unsupported light [key='{}']", key);
                }
                lights.push_back(l);
                return;
            }
            if (!node.primitive.mesh) {
                return;
            }
            const Mesh* mesh = node.primitive.mesh;
            auto it = mesh_ids.find(mesh);
            if (it == mesh_ids.end()) {
                it = mesh_ids.emplace(mesh, int(mesh_p.size())).first;
                mesh_p.emplace_back();
                mesh_n.emplace_back();
                mesh_t.emplace_back();
                mesh_list.push_back(mesh);
            }
            LmcuPrimitive p{};
            p.node_index = node.index;
            p.mesh = it->second;
            p.material = 0;
            p.light = -1;
            for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) p.M[c * 4 + r] = global_transform[c][r];
            if (shading) {
                p.material = material_index(node.primitive.material);
                if (node.primitive.light) {
                    const auto key = node.primitive.light->key();
                    if (key != "light::area") {
                        LM_THROW_EXCEPTION(Error::Unsupported, "renderer::*_cuda: unsupported light [key='{}']", key);
                    }
                    LmcuLight l{};
                    l.type = LMCU_LIGHT_AREA;
                    l.primitive = int(prims.size());
                    l.node_index = node.index;
                    // Ke through Light::eval on the emitting side (light_area.cpp:185-187)
                    const auto g = PointGeometry::make_on_surface(Vec3(0_f), Vec3(0_f, 0_f, 1_f), Vec3(0_f, 0_f, 1_f));
                    const auto Ke = node.primitive.light->eval(g, Vec3(0_f, 0_f, 1_f), false);
                    l.Ke[0] = Ke.x; l.Ke[1] = Ke.y; l.Ke[2] = Ke.z;
                    p.light = int(lights.size());
                    lights.push_back(l);
                }
            }
            transforms.emplace(node.index, Transform(global_transform));
            prims.push_back(p);
        });
        if (mats.empty()) {
            mats.push_back(LmcuMaterial{});
        }
        if (shading && scene.camera_node() >= 0) {
            // camera::pinhole exports its parameters (src/camera/camera_pinhole.cpp:66-74)
            const auto* cam = scene.node_at(scene.camera_node()).primitive.camera;
            const Json v = cam->underlying_value("");
            const Vec3 eye = v["eye"], center = v["center"], up = v["up"];
            for (int i = 0; i < 3; i++) { camera.position[i] = eye[i]; camera.center[i] = center[i]; camera.up[i] = up[i]; }
            camera.vfov = v["vfov"];
            camera.aspect = v["aspect"];
            has_camera = true;
        }
        meshes.assign(mesh_p.size(), LmcuMesh{});
        for (size_t i = 0; i < mesh_p.size(); i++) meshes[i].num_faces = mesh_list[i]->num_triangles();   // arrays: gather_triangles()
        for (size_t i = 0; i < textures.size(); i++) textures[i].data = tex_data[i].data();
        for (auto& m : mats) {
            if (m.num_entries > 0) {
                const size_t slot = reinterpret_cast<size_t>(m.entry_material) - 1;
                m.entry_material = entry_material[slot].data();
                m.entry_weight = entry_weight[slot].data();
            }
        }
        desc.num_textures = int(textures.size());
        desc.textures = textures.data();
        desc.num_meshes = int(meshes.size());
        desc.num_primitives = int(prims.size());
        desc.num_materials = int(mats.size());
        desc.num_lights = int(lights.size());
        desc.meshes = meshes.data();
        desc.primitives = prims.data();
        desc.materials = mats.data();
        desc.lights = lights.data();
        desc.camera = has_camera ? &camera : nullptr;
        compute_signature();
    }
};

// One uploaded + built scene on one device.  Shared (through the cache below) by every component of this plugin that
// exports the same lm::Scene content: a renderer asset that is re-created for every render, as the reference's examples
// do with lm::load<Renderer>, finds the scene its predecessor uploaded.
struct DeviceScene {
    LmcuScene* h = nullptr;
    uint64_t sig = 0;
    int device = 0;
    bool shading = false;
    int layout = 0;                        // LmcuBuildOptions::two_level the scene was built with
    std::map<int, Transform> transforms;   // SceneNode::index -> Transform for Accel::Hit
    std::mutex render_mu;                  // one render at a time per device scene (film and path pools live in it)
    ~DeviceScene() { if (h) lmcu_scene_destroy(h); }
    DeviceScene() = default;
    DeviceScene(const DeviceScene&) = delete;
    void upload(SceneExport& ex, int dev, int two_level) {
        ex.gather_triangles();
        if (h) { lmcu_scene_destroy(h); h = nullptr; }
        device = dev; sig = ex.sig; shading = ex.shading_; layout = two_level;
        transforms = ex.transforms;
        check(lmcu_scene_create(device, &h));
        check(lmcu_scene_upload(h, &ex.desc));
        LmcuBuildOptions opts{};
        opts.two_level = two_level;
        check(lmcu_accel_build(h, &opts, nullptr));
    }
};

struct SceneCache {
    std::mutex mu;
    std::list<std::shared_ptr<DeviceScene>> entries;      // most recently used first
    static constexpr size_t kMax = 4;
    long long hits = 0, misses = 0;
    std::shared_ptr<DeviceScene> acquire(SceneExport& ex, int device, bool use_cache, int two_level = 0) {
        if (use_cache) {
            std::unique_lock<std::mutex> lk(mu);
            for (auto it = entries.begin(); it != entries.end(); ++it) {
                if ((*it)->sig == ex.sig && (*it)->device == device && (*it)->shading == ex.shading_ && (*it)->layout == two_level) {
                    auto hit = *it;
                    hits++;
                    entries.erase(it);
                    entries.push_front(hit);
                    return hit;
                }
            }
        }
        auto fresh = std::make_shared<DeviceScene>();
        fresh->upload(ex, device, two_level);
        if (use_cache) {
            std::unique_lock<std::mutex> lk(mu);
            misses++;
            entries.push_front(fresh);
            while (entries.size() > kMax) entries.pop_back();
        }
        return fresh;
    }
    void clear() { std::unique_lock<std::mutex> lk(mu); entries.clear(); }
};
// Never destroyed: liblm's UserContext singleton (constructed by lm::init(), before this plugin is loaded) destroys the asset
// tree — and with it this plugin's components — AFTER the plugin's own statics at process exit.
SceneCache& g_scene_cache = *new SceneCache;

}

// ------------------------------------------------------------------------------------------------

/*
\rst
.. function:: accel::cuda_bvh

   Compressed 8-wide BVH built and traversed on a CUDA device (sm_100a).

   :param int device: CUDA device index. Default: 0.
   :param bool cache: Reuse the device scene of an earlier build of the same scene content. Default: true.
   :param int two_level: 1: instancing-aware layout, one tree per mesh + one over the primitives (as ``accel::embreeinstanced``;
      memory and build time follow the unique triangles, same hits, slower walk); -1: one tree over the flattened scene;
      0 (default): flattened while it fits the device. A two-level accel cannot be saved.

   ``intersect()`` keeps the scalar contract of :cpp:func:`lm::Accel::intersect` (one ray per call,
   thread-safe and lock-free: every calling thread owns a pinned ray/hit slot and a CUDA stream) for callers
   such as ``scene::default`` under the reference's CPU renderers; batched rays and the device-resident
   renderers are the fast path.
\endrst
*/
class Accel_CudaBVH final : public Accel {
private:
    int device_ = 0;
    bool cache_ = true;
    int two_level_ = 0;
    std::shared_ptr<DeviceScene> dev_;
    mutable std::mutex mu_;

public:
    virtual void construct(const Json& prop) override {
        device_ = json::value<int>(prop, "device", 0);
        cache_ = json::value<bool>(prop, "cache", true);
        two_level_ = json::value<int>(prop, "two_level", 0);
    }

    // Component::save / load (include/lm/component.h:210-236), the counterpart of accel::sahbvh archiving nodes_, trs_,
    // indices_ (accel_sahbvh.cpp:122-124): the device BVH travels as the blob of lmcu_accel_save(), so a loaded state
    // answers intersect() without rebuilding (tests/test_gpu_plugin.py::test_save_state_round_trip_with_cuda_components
    // runs lm::save_state_to_file / lm::reset / lm::load_state_from_file over it, func_serial_consistency style).
    virtual void save(OutputArchive& ar) override {
        std::vector<unsigned char> blob;
        std::vector<int> keys;
        std::vector<Transform> values;
        if (dev_ && dev_->h) {
            int64_t n = 0;
            check(lmcu_accel_save_size(dev_->h, &n));
            blob.resize(size_t(n));
            check(lmcu_accel_save(dev_->h, blob.data(), n));
            for (const auto& [k, v] : dev_->transforms) { keys.push_back(k); values.push_back(v); }
        }
        ar(device_, blob, keys, values);
    }

    virtual void load(InputArchive& ar) override {
        std::vector<unsigned char> blob;
        std::vector<int> keys;
        std::vector<Transform> values;
        ar(device_, blob, keys, values);
        dev_.reset();
        if (blob.empty()) return;
        auto fresh = std::make_shared<DeviceScene>();          // restored from the blob: never enters the cache
        fresh->device = device_;
        for (size_t i = 0; i < keys.size() && i < values.size(); i++) fresh->transforms.emplace(keys[i], values[i]);
        check(lmcu_scene_create(device_, &fresh->h));
        check(lmcu_accel_load(fresh->h, blob.data(), int64_t(blob.size())));
        dev_ = fresh;
    }

    virtual void build(const Scene& scene) override {
        LM_INFO("Flattening scene");
        SceneExport ex;
        ex.gather_description(scene, false);
        LM_INFO("Building");
        dev_ = g_scene_cache.acquire(ex, device_, cache_, two_level_);
    }

    // Batched closest hit: n rays in, n hits out, one kernel launch.  Hit::primitive is the scene node
    // index as in the scalar contract; the transform of that node is transforms_.at(primitive).
    void intersect_batch(const LmcuRay* rays, long long n, LmcuHit* hits) const {
        if (!dev_) LM_THROW_EXCEPTION(Error::Unsupported, "accel::cuda_bvh: build() or load() first");
        std::unique_lock<std::mutex> lk(mu_);                  // the batch path stages through the scene's own stream
        check(lmcu_accel_intersect(dev_->h, rays, n, hits));
    }

    virtual std::optional<Hit> intersect(Ray ray, Float tmin, Float tmax) const override {
        if (!dev_) LM_THROW_EXCEPTION(Error::Unsupported, "accel::cuda_bvh: build() or load() first");
        const LmcuRay r{ { ray.o.x, ray.o.y, ray.o.z }, { ray.d.x, ray.d.y, ray.d.z }, tmin, tmax };
        LmcuHit h{};
        check(lmcu_accel_intersect1(dev_->h, &r, &h));         // lock-free: a pinned slot + stream per calling thread
        if (h.primitive < 0) {
            return {};
        }
        return Hit{ h.t, Vec2(h.u, h.v), dev_->transforms.at(h.primitive), h.primitive, h.face };
    }
};

LM_COMP_REG_IMPL(Accel_CudaBVH, "accel::cuda_bvh");

// ------------------------------------------------------------------------------------------------

/*
\rst
.. function:: film::cuda

   Film whose pixels are the fp32 RGB buffer the CUDA renderers produce.

   :param int w: Width of the film.
   :param int h: Height of the film.

   ``renderer::*_cuda`` write their result into it with ONE device-to-host copy; with ``film::bitmap`` the same result
   goes through ``w*h`` ``set_pixel`` calls, each a compare-and-swap loop on a 24-byte atomic
   (src/film/film_bitmap.cpp:41-54,152-154).  Everything else behaves like ``film::bitmap``: pixel (x, y) with y growing
   upward, ``buffer()`` returns ``w*h*3`` Floats, ``save()`` writes .pfm itself and hands the other formats to a
   temporary ``film::bitmap``.
\endrst
*/
class Film_Cuda final : public Film {
private:
    int w_ = 0, h_ = 0;
    std::vector<float> rgb_;                // [h][w][3], row 0 = bottom row (pixel y = 0)
    std::vector<Float> temp_;               // buffer() view, refreshed only after the pixels changed
    bool dirty_ = true;

    static void atomic_add(float* p, float v) {
        auto* a = reinterpret_cast<uint32_t*>(p);
        uint32_t old = __atomic_load_n(a, __ATOMIC_RELAXED);
        for (;;) {
            float f; std::memcpy(&f, &old, 4);
            f += v;
            uint32_t nw; std::memcpy(&nw, &f, 4);
            if (__atomic_compare_exchange_n(a, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
        }
    }

public:
    virtual void save(OutputArchive& ar) override { ar(w_, h_, rgb_); }                       // film::bitmap: film_bitmap.cpp:132-134
    virtual void load(InputArchive& ar) override { ar(w_, h_, rgb_); dirty_ = true; }

    virtual void construct(const Json& prop) override {
        w_ = json::value<int>(prop, "w");
        h_ = json::value<int>(prop, "h");
        rgb_.assign(size_t(w_) * h_ * 3, 0.f);
    }
    // bulk interface of the CUDA renderers
    float* data() { dirty_ = true; return rgb_.data(); }
    size_t num_floats() const { return rgb_.size(); }

    virtual FilmSize size() const override { return { w_, h_ }; }
    virtual long long num_pixels() const override { return (long long)(w_) * h_; }
    virtual void set_pixel(int x, int y, Vec3 v) override {
        float* p = &rgb_[(size_t(y) * w_ + x) * 3];
        p[0] = float(v.x); p[1] = float(v.y); p[2] = float(v.z);
        dirty_ = true;
    }
    virtual bool save(const std::string& outpath) const override {
        const auto dot = outpath.find_last_of('.');
        const std::string ext = dot == std::string::npos ? "" : outpath.substr(dot);
        if (ext == ".pfm") {                                               // film_bitmap.cpp:62-80,197-203 (no flip)
            FILE* f = std::fopen(outpath.c_str(), "wb");
            if (!f) { LM_ERROR("Failed to open [file='{}']", outpath); return false; }
            std::fprintf(f, "PF\n%d %d\n-1\n", w_, h_);
            std::fwrite(rgb_.data(), 4, rgb_.size(), f);
            std::fclose(f);
            return true;
        }
        auto tmp = comp::create<Film>("film::bitmap", "", { {"w", w_}, {"h", h_} });
        if (!tmp) return false;
        for (int y = 0; y < h_; y++) for (int x = 0; x < w_; x++) {
            const float* p = &rgb_[(size_t(y) * w_ + x) * 3];
            tmp->set_pixel(x, y, Vec3(p[0], p[1], p[2]));
        }
        return tmp->save(outpath);
    }
    virtual FilmBuffer buffer() override {
        if (dirty_ || temp_.size() != rgb_.size()) {
            temp_.resize(rgb_.size());
            const float* src = rgb_.data();
            Float* dst = temp_.data();
            const long long n = (long long)rgb_.size();
            #pragma omp parallel for schedule(static)
            for (long long i = 0; i < n; i++) dst[i] = Float(src[i]);
            dirty_ = false;
        }
        return FilmBuffer{ w_, h_, temp_.data() };
    }
    virtual void accum(const Film* film_) override {
        const auto* film = dynamic_cast<const Film_Cuda*>(film_);
        if (!film || film->w_ != w_ || film->h_ != h_) {
            LM_ERROR("Could not accumuate film. Invalid film type or size.");
            return;
        }
        for (size_t i = 0; i < rgb_.size(); i++) rgb_[i] += film->rgb_[i];
        dirty_ = true;
    }
    virtual void splat_pixel(int x, int y, Vec3 v) override {
        float* p = &rgb_[(size_t(y) * w_ + x) * 3];
        atomic_add(p, float(v.x)); atomic_add(p + 1, float(v.y)); atomic_add(p + 2, float(v.z));
        dirty_ = true;
    }
    virtual void update_pixel(int x, int y, const PixelUpdateFunc& update_func) override {
        float* p = &rgb_[(size_t(y) * w_ + x) * 3];
        const Vec3 v = update_func(Vec3(p[0], p[1], p[2]));
        p[0] = float(v.x); p[1] = float(v.y); p[2] = float(v.z);
        dirty_ = true;
    }
    virtual void rescale(Float s) override {
        for (auto& v : rgb_) v = float(v * s);
        dirty_ = true;
    }
    virtual void clear() override {
        rgb_.assign(rgb_.size(), 0.f);
        dirty_ = true;
    }
};

LM_COMP_REG_IMPL(Film_Cuda, "film::cuda");

// ------------------------------------------------------------------------------------------------

// Shared part of the three device-resident renderers.
class Renderer_CudaBase : public Renderer {
protected:
    Scene* scene_;
    Film* film_;
    mutable int device_ = 0;
    bool cache_ = true;
    int two_level_ = 0;                                        // "two_level": LmcuBuildOptions::two_level of the device scene
    mutable std::shared_ptr<DeviceScene> dev_;

public:
    virtual void foreach_underlying(const ComponentVisitor& visit) override {
        comp::visit(visit, scene_);
        comp::visit(visit, film_);
    }

protected:
    // weak references travel as locators and are re-pointed after the whole tree is loaded (include/lm/serial.h:238-300)
    template <typename Archive> void serialize_base(Archive& ar) { ar(scene_, film_, device_, cache_, two_level_); }

    void construct_base(const Json& prop) {
        scene_ = json::comp_ref<Scene>(prop, "scene");
        film_ = json::comp_ref<Film>(prop, "output");
        device_ = json::value<int>(prop, "device", 0);
        cache_ = json::value<bool>(prop, "cache", true);      // false: export, upload and build again on every render()
        two_level_ = json::value<int>(prop, "two_level", 0);   // 1: one tree per mesh + one over the primitives (instanced scenes)
    }

    // The device scene for the CURRENT content of scene_: found in the cache when nothing that the export reads has
    // changed since an earlier render (by this or any other component of the plugin), else exported, uploaded and built.
    void upload() const {
        SceneExport ex;
        ex.gather_description(*scene_, true);
        dev_ = g_scene_cache.acquire(ex, device_, cache_, two_level_);
    }

    // Host destination of the device film: film::cuda's own pixel buffer (one copy), else a staging vector
    float* film_target(std::vector<float>& staging) const {
        if (auto* fc = dynamic_cast<Film_Cuda*>(film_)) return fc->data();
        const auto size = film_->size();
        staging.resize(size_t(size.w) * size.h * 3);
        return staging.data();
    }

    // `skip_marked`: pixels whose value is negative were not written by the renderer (LmcuRenderParams::ao_mark_misses)
    // and keep what the film held
    void write_film(const std::vector<float>& rgb, bool skip_marked) const {
        if (auto* fc = dynamic_cast<Film_Cuda*>(film_)) {
            if (rgb.empty()) return;                                       // already written in place
            float* dst = fc->data();
            for (size_t i = 0; i < rgb.size(); i += 3) {
                if (skip_marked && rgb[i] < 0.f) continue;
                dst[i] = rgb[i]; dst[i + 1] = rgb[i + 1]; dst[i + 2] = rgb[i + 2];
            }
            return;
        }
        // a foreign film takes its pixels one set_pixel at a time (film::bitmap: a CAS loop on a 24-byte atomic per call,
        // src/film/film_bitmap.cpp:41-54): rows go out through the reference's own parallel::foreach, otherwise this
        // write-back of a 4K image costs more than rendering it
        const auto size = film_->size();
        parallel::foreach(size.h, [&](long long y, int) {
            for (int x = 0; x < size.w; x++) {
                const float* c = &rgb[(size_t(y) * size.w + x) * 3];
                if (skip_marked && c[0] < 0.f) continue;
                film_->set_pixel(x, int(y), Vec3(c[0], c[1], c[2]));
            }
        });
    }
};

class Renderer_Raycast_Cuda final : public Renderer_CudaBase {
private:
    Vec3 bg_color_;
    bool use_constant_color_;
    bool visualize_normal_;
    std::optional<Vec3> color_;

public:
    LM_SERIALIZE_IMPL(ar) { serialize_base(ar); ar(bg_color_, use_constant_color_, visualize_normal_, color_); }   // renderer_raycast.cpp:29-31

    virtual void construct(const Json& prop) override {
        construct_base(prop);
        bg_color_ = json::value(prop, "bg_color", Vec3(0_f));
        use_constant_color_ = json::value(prop, "use_constant_color", false);
        visualize_normal_ = json::value(prop, "visualize_normal", false);
        color_ = json::value_or_none<Vec3>(prop, "color");
    }

    virtual Json render() const override {
        scene_->require_primitive();
        scene_->require_camera();
        film_->clear();
        upload();
        const auto size = film_->size();
        LmcuRenderParams p{};
        p.width = size.w; p.height = size.h; p.spp = 1;
        p.use_constant_color = use_constant_color_; p.visualize_normal = visualize_normal_; p.has_color = color_ ? 1 : 0;
        for (int i = 0; i < 3; i++) { p.bg_color[i] = bg_color_[i]; p.color[i] = color_ ? (*color_)[i] : 0_f; }
        std::vector<float> rgb;
        LmcuRenderStats st{};
        {
            std::unique_lock<std::mutex> lk(dev_->render_mu);
            check(lmcu_render_raycast(dev_->h, &p, film_target(rgb), &st));
        }
        write_film(rgb, false);
        return { {"elapsed", st.elapsed} };
    }
};

LM_COMP_REG_IMPL(Renderer_Raycast_Cuda, "renderer::raycast_cuda");

class Renderer_AO_Cuda final : public Renderer_CudaBase {
private:
    long long spp_;
    int rng_seed_ = 42;

public:
    LM_SERIALIZE_IMPL(ar) { serialize_base(ar); ar(spp_, rng_seed_); }

    virtual void construct(const Json& prop) override {
        construct_base(prop);
        spp_ = json::value<long long>(prop, "spp");
    }

    virtual Json render() const override {
        // renderer::ao itself checks nothing and dereferences a missing camera (functest/renderer_ao.cpp:30); the checks of
        // renderer::raycast (renderer_raycast.cpp:63-64) turn that into the reference's own exceptions
        scene_->require_primitive();
        scene_->require_camera();
        upload();
        const auto size = film_->size();
        LmcuRenderParams p{};
        p.width = size.w; p.height = size.h; p.spp = spp_; p.seed = unsigned(rng_seed_); p.ao_distance = .2_f;
        p.ao_mark_misses = 1;                // the reference leaves a pixel whose primary ray misses untouched (:32-34)
        std::vector<float> rgb(size_t(size.w) * size.h * 3);
        LmcuRenderStats st{};
        {
            std::unique_lock<std::mutex> lk(dev_->render_mu);
            check(lmcu_render_ao(dev_->h, &p, rgb.data(), &st));
        }
        write_film(rgb, true);
        return { {"elapsed", st.elapsed} };
    }
};

LM_COMP_REG_IMPL(Renderer_AO_Cuda, "renderer::ao_cuda");

/*
\rst
.. function:: renderer::pt_cuda

   Wavefront path tracer on CUDA devices; same parameters as ``renderer::pt`` plus:

   :param int device: CUDA device index. Default: 0.
   :param int wavefronts: Staggered wavefront loops per GPU (each on its own CUDA stream; their kernels overlap).
                          Default: 0 = the library's default (2). The image does not depend on it.
   :param list devices: Render on several GPUs of this machine by SAMPLE SPLIT: the scene and BVH are replicated,
                        device *g* renders its block of the ``spp`` sample indices of every pixel into a private film,
                        and the first device combines the films with one kernel over NVLink peer memory that also
                        applies the final 1/spp rescale. Default: ``[device]``.
\endrst
*/
class Renderer_PT_Cuda final : public Renderer_CudaBase {
private:
    int max_verts_;
    std::optional<unsigned int> seed_;
    int sampling_mode_;
    int primary_mode_ = LMCU_PRIMARY_PIXEL;     // primary_ray_sampling_mode: pixel -> scheduler::spp::*, image -> scheduler::spi::*
    int scheduler_ = LMCU_SCHED_SAMPLE;         // scheduler: sample | time
    long long spp_ = 0;                         // spp::sample
    long long num_samples_ = 0;                 // spi::sample
    long long samples_per_iter_ = 0;            // spi::time
    Float render_time_ = 0_f;                   // {spp,spi}::time
    std::vector<int> devices_;
    int wavefronts_ = 0;                        // LmcuRenderParams::wavefronts (0 = library default)

public:
    LM_SERIALIZE_IMPL(ar) {                                                                    // renderer_pt.cpp:36-38
        serialize_base(ar);
        ar(max_verts_, seed_, sampling_mode_, primary_mode_, scheduler_, spp_, num_samples_, samples_per_iter_, render_time_, devices_, wavefronts_);
    }

    virtual void construct(const Json& prop) override {
        construct_base(prop);
        max_verts_ = json::value<int>(prop, "max_verts");
        seed_ = json::value_or_none<unsigned int>(prop, "seed");
        {
            const auto s = json::value<std::string>(prop, "sampling_mode", "mis");
            if (s == "naive") sampling_mode_ = LMCU_SAMPLING_NAIVE;
            else if (s == "nee") sampling_mode_ = LMCU_SAMPLING_NEE;
            else sampling_mode_ = LMCU_SAMPLING_MIS;
        }
        {
            // Same keys as renderer::pt (renderer_pt.cpp:63-76) and its schedulers (src/scheduler.cpp:32-35,69-72,127-129,157-160)
            const auto name = json::value<std::string>(prop, "scheduler");
            const auto s = json::value<std::string>(prop, "primary_ray_sampling_mode", "pixel");
            if (s == "pixel") primary_mode_ = LMCU_PRIMARY_PIXEL;
            else if (s == "image") primary_mode_ = LMCU_PRIMARY_IMAGE;
            else LM_THROW_EXCEPTION(Error::InvalidArgument, "renderer::pt_cuda: unknown primary_ray_sampling_mode '{}'", s);
            if (name == "sample") {
                scheduler_ = LMCU_SCHED_SAMPLE;
                if (primary_mode_ == LMCU_PRIMARY_PIXEL) spp_ = json::value<long long>(prop, "spp");
                else num_samples_ = json::value<long long>(prop, "num_samples");
            }
            else if (name == "time") {
                scheduler_ = LMCU_SCHED_TIME;
                render_time_ = json::value<Float>(prop, "render_time");
                samples_per_iter_ = json::value<long long>(prop, "samples_per_iter", 100000);
            }
            else LM_THROW_EXCEPTION(Error::InvalidArgument, "renderer::pt_cuda: unknown scheduler '{}'", name);
        }
        wavefronts_ = json::value<int>(prop, "wavefronts", 0);
        devices_ = json::value<std::vector<int>>(prop, "devices", std::vector<int>{ device_ });
        if (devices_.empty() || devices_.size() > LMCU_MAX_PEERS + 1) {
            LM_THROW_EXCEPTION(Error::InvalidArgument, "renderer::pt_cuda: 'devices' must name 1..{} GPUs", LMCU_MAX_PEERS + 1);
        }
    }

    // progress::update / update_time between wavefront batches (src/scheduler.cpp:45,91,133,173 report per sample)
    struct ProgressCtx { bool timed; int devices; };
    static void report_progress(void* user, int64_t processed_paths, int64_t, double elapsed) {
        const auto* c = static_cast<const ProgressCtx*>(user);
        if (c->timed) progress::update_time(elapsed);
        else progress::update(processed_paths * c->devices);       // every device is as far along as the reporting one
    }

    virtual Json render() const override {
        scene_->require_renderable();
        film_->clear();
        const auto size = film_->size();
        LmcuRenderParams p{};
        p.width = size.w; p.height = size.h; p.max_verts = max_verts_;
        p.sampling_mode = sampling_mode_;
        p.primary_ray_sampling_mode = primary_mode_;
        p.scheduler = scheduler_;
        p.render_time = render_time_;
        p.samples_per_iter = samples_per_iter_;
        p.wavefronts = wavefronts_;
        p.seed = seed_ ? *seed_ : math::rng_seed();
        std::vector<float> rgb;
        float* target = film_target(rgb);
        LmcuRenderStats st{};
        const int G = int(devices_.size());
        const bool image = primary_mode_ == LMCU_PRIMARY_IMAGE;
        const bool timed = scheduler_ == LMCU_SCHED_TIME;
        const long long total = image ? num_samples_ : spp_;      // sample schedulers: the work to split
        // progress context of the reference's schedulers: samples (spp::sample counts pixels x spp, scheduler.cpp:39) or time
        const long long total_paths = image ? num_samples_ : (long long)size.w * size.h * spp_;
        if (timed) progress::start(progress::ProgressMode::Time, 0, render_time_);
        else progress::start(progress::ProgressMode::Samples, total_paths, -1);
        struct ProgressEnd { ~ProgressEnd() { progress::end(); } } progress_end;
        ProgressCtx pctx{ timed, G };
        p.progress = &Renderer_PT_Cuda::report_progress;
        p.progress_user = &pctx;
        if (G == 1) {
            device_ = devices_[0];
            upload();
            p.spp = spp_; p.num_samples = num_samples_; p.rescale = 1;
            std::unique_lock<std::mutex> lk(dev_->render_mu);
            check(lmcu_render_pt(dev_->h, &p, target, &st));
        }
        else {
            // Sample split: one host thread + one replicated device scene per GPU
            if (!timed && total < G) {
                LM_THROW_EXCEPTION(Error::InvalidArgument, "renderer::pt_cuda: spp / num_samples must be >= the number of devices");
            }
            SceneExport ex;
            ex.gather_description(*scene_, true);
            std::vector<std::shared_ptr<DeviceScene>> devs(G);
            std::vector<std::unique_lock<std::mutex>> locks(G);
            std::vector<LmcuRenderStats> stats(G);
            std::vector<std::string> errors(G);
            std::vector<std::thread> threads;
            for (int g = 0; g < G; g++) {
                threads.emplace_back([&, g] {
                    try {
                        devs[g] = g_scene_cache.acquire(ex, devices_[g], cache_, two_level_);
                        locks[g] = std::unique_lock<std::mutex>(devs[g]->render_mu);
                        LmcuRenderParams q = p;
                        q.rescale = 0;
                        if (g != 0) q.progress = nullptr;
                        if (!timed) {
                            const long long base = total / G, rem = total % G;
                            const long long mine = base + (g < rem ? 1 : 0);
                            q.spp = image ? 0 : mine; q.num_samples = image ? mine : 0;
                            q.sample_offset = g * base + std::min<long long>(g, rem);
                        }
                        else {
                            // time budget: every device runs for render_time on its own block of sample indices
                            // and reports how many it finished; the counts are summed like the films
                            q.sample_offset = (long long)g << 40;
                        }
                        check(lmcu_render_pt(devs[g]->h, &q, nullptr, &stats[g]));
                    }
                    catch (const std::exception& e) { errors[g] = e.what(); }
                });
            }
            for (auto& t : threads) t.join();
            for (const auto& e : errors) if (!e.empty()) LM_THROW_EXCEPTION(Error::None, "renderer::pt_cuda: {}", e);
            st.processed = 0;
            for (int g = 0; g < G; g++) {
                st.processed += stats[g].processed;
                st.elapsed = std::max(st.elapsed, stats[g].elapsed);
            }
            // renderer_pt.cpp:307-312 on the combined film.  Slice-parallel combine over NVLink peer memory: device g sums
            // slice g of all G films and stores it, rescaled, into device 0's film (include/lmcu.h, lmcu_film_reduce_slice).
            const double scale = image ? double(size.w) * size.h / double(st.processed) : 1.0 / double(st.processed);
            std::vector<void*> films(G);
            int64_t n = 0;
            for (int g = 0; g < G; g++) check(lmcu_film_device_ptr(devs[g]->h, &films[g], &n));
            for (int a = 0; a < G; a++) for (int b = 0; b < G; b++) if (a != b) check(lmcu_enable_peer_access(devices_[a], devices_[b]));
            threads.clear();
            const int64_t quads = n / 4;
            for (int g = 0; g < G; g++) {
                threads.emplace_back([&, g] {
                    try {
                        const int64_t begin = quads * g / G * 4, end = g == G - 1 ? n : quads * (g + 1) / G * 4;
                        check(lmcu_film_reduce_slice(devs[g]->h, films.data(), G, films[0], begin, end, scale));
                    }
                    catch (const std::exception& e) { errors[g] = e.what(); }
                });
            }
            for (auto& t : threads) t.join();
            for (const auto& e : errors) if (!e.empty()) LM_THROW_EXCEPTION(Error::None, "renderer::pt_cuda: {}", e);
            check(lmcu_film_download(devs[0]->h, target));
        }
        write_film(rgb, false);
        return { {"processed", st.processed}, {"elapsed", st.elapsed} };
    }
};

LM_COMP_REG_IMPL(Renderer_PT_Cuda, "renderer::pt_cuda");

/*
\rst
.. function:: renderer::lt_cuda

   Light tracing with next-event estimation toward the camera on a CUDA device; same parameters as ``renderer::lt``
   (``scene``, ``output``, ``max_verts``, ``seed``, ``scheduler`` = ``sample`` with ``num_samples`` or ``time`` with
   ``render_time`` / ``samples_per_iter``; the scheduler is the image-space one, src/renderer/renderer_lt.cpp:44-47) plus
   ``device``, ``wavefronts`` and ``cache`` as in ``renderer::pt_cuda``.
\endrst
*/
// Shared by the two renderers on the image-space scheduler (scheduler::spi::*): renderer::lt_cuda and renderer::bdpt_cuda
class Renderer_SPI_CudaBase : public Renderer_CudaBase {
protected:
    int max_verts_;
    int min_verts_ = 0;
    std::optional<unsigned int> seed_;
    int scheduler_ = LMCU_SCHED_SAMPLE;
    long long num_samples_ = 0;
    long long samples_per_iter_ = 0;
    Float render_time_ = 0_f;
    int wavefronts_ = 0;

    static void report_progress(void* user, int64_t processed_paths, int64_t, double elapsed) {
        if (*static_cast<const bool*>(user)) progress::update_time(elapsed);
        else progress::update(processed_paths);
    }

public:
    LM_SERIALIZE_IMPL(ar) {                                                                    // renderer_lt.cpp:27-29
        serialize_base(ar);
        ar(max_verts_, min_verts_, seed_, scheduler_, num_samples_, samples_per_iter_, render_time_, wavefronts_);
    }

    // the C-ABI entry of the concrete renderer
    virtual int launch(LmcuScene* h, const LmcuRenderParams* p, float* target, LmcuRenderStats* st) const = 0;

    virtual void construct(const Json& prop) override {
        construct_base(prop);
        max_verts_ = json::value<int>(prop, "max_verts");
        min_verts_ = json::value<int>(prop, "min_verts", 2);                                   // renderer_bdpt.cpp:51 (unused by lt)
        seed_ = json::value_or_none<unsigned int>(prop, "seed");
        const auto name = json::value<std::string>(prop, "scheduler");
        if (name == "sample") {
            scheduler_ = LMCU_SCHED_SAMPLE;
            num_samples_ = json::value<long long>(prop, "num_samples");                        // scheduler::spi::sample, scheduler.cpp:127-129
        }
        else if (name == "time") {
            scheduler_ = LMCU_SCHED_TIME;
            render_time_ = json::value<Float>(prop, "render_time");                            // scheduler::spi::time, :157-160
            samples_per_iter_ = json::value<long long>(prop, "samples_per_iter", 100000);
        }
        else LM_THROW_EXCEPTION(Error::InvalidArgument, "unknown scheduler '{}' (scheduler::spi::sample | time)", name);
        wavefronts_ = json::value<int>(prop, "wavefronts", 0);
    }

    virtual Json render() const override {
        scene_->require_renderable();
        film_->clear();
        const auto size = film_->size();
        LmcuRenderParams p{};
        p.width = size.w; p.height = size.h; p.max_verts = max_verts_; p.min_verts = min_verts_;
        p.primary_ray_sampling_mode = LMCU_PRIMARY_IMAGE;
        p.scheduler = scheduler_;
        p.num_samples = num_samples_;
        p.render_time = render_time_;
        p.samples_per_iter = samples_per_iter_;
        p.wavefronts = wavefronts_;
        p.rescale = 1;
        p.seed = seed_ ? *seed_ : math::rng_seed();
        bool timed = scheduler_ == LMCU_SCHED_TIME;
        if (timed) progress::start(progress::ProgressMode::Time, 0, render_time_);
        else progress::start(progress::ProgressMode::Samples, num_samples_, -1);
        struct ProgressEnd { ~ProgressEnd() { progress::end(); } } progress_end;
        p.progress = &Renderer_SPI_CudaBase::report_progress;
        p.progress_user = &timed;
        upload();
        std::vector<float> rgb;
        float* target = film_target(rgb);
        LmcuRenderStats st{};
        {
            std::unique_lock<std::mutex> lk(dev_->render_mu);
            check(launch(dev_->h, &p, target, &st));
        }
        write_film(rgb, false);
        return { {"processed", st.processed}, {"elapsed", st.elapsed} };
    }
};

class Renderer_LT_Cuda final : public Renderer_SPI_CudaBase {
    virtual int launch(LmcuScene* h, const LmcuRenderParams* p, float* target, LmcuRenderStats* st) const override {
        return lmcu_render_lt(h, p, target, st);
    }
};

LM_COMP_REG_IMPL(Renderer_LT_Cuda, "renderer::lt_cuda");

/*
\rst
.. function:: renderer::bdpt_cuda

   Bidirectional path tracing on the GPU: ``renderer::bdpt`` (src/renderer/renderer_bdpt.cpp:304-407) through
   ``lmcu_render_bdpt``.  Same parameters: ``scene``, ``output``, ``max_verts`` (2..20), ``min_verts`` (default 2), ``seed``,
   ``scheduler`` (``sample`` with ``num_samples`` | ``time`` with ``render_time``, ``samples_per_iter``); plus ``device`` and
   ``cache`` as in ``renderer::pt_cuda``.
\endrst
*/
class Renderer_BDPT_Cuda final : public Renderer_SPI_CudaBase {
    virtual int launch(LmcuScene* h, const LmcuRenderParams* p, float* target, LmcuRenderStats* st) const override {
        return lmcu_render_bdpt(h, p, target, st);
    }
};

LM_COMP_REG_IMPL(Renderer_BDPT_Cuda, "renderer::bdpt_cuda");

/*
\rst
.. function:: renderer::bdptopt_cuda

   ``renderer::bdptopt`` (src/renderer/renderer_bdptopt.cpp) is the estimator of ``renderer::bdpt`` with its MIS weights
   computed incrementally; on the device there is one implementation, so this key runs ``renderer::bdpt_cuda``.

.. function:: renderer::pt_naive_path_cuda
.. function:: renderer::pt_nee_path_cuda
.. function:: renderer::lt_nee_path_cuda

   The reference's single-strategy renderers on the path structure (src/renderer/renderer_bdpt.cpp:63-124, 129-200,
   205-288): the subpaths and connections of ``renderer::bdpt_cuda`` restricted to one strategy family — the eye subpath
   alone, (0, t); next-event estimation, (1, n-1), or (0, n) behind a specular vertex; the light subpath connected to the
   pinhole, (n-1, 1) — each connection unweighted (``LmcuRenderParams::bdpt_strategies``).  Parameters as
   ``renderer::bdpt_cuda``.
\endrst
*/
class Renderer_BDPTOpt_Cuda final : public Renderer_SPI_CudaBase {
    virtual int launch(LmcuScene* h, const LmcuRenderParams* p, float* target, LmcuRenderStats* st) const override {
        return lmcu_render_bdpt(h, p, target, st);
    }
};

LM_COMP_REG_IMPL(Renderer_BDPTOpt_Cuda, "renderer::bdptopt_cuda");

template <int Strategies>
class Renderer_PathStrategy_Cuda final : public Renderer_SPI_CudaBase {
    virtual int launch(LmcuScene* h, const LmcuRenderParams* p, float* target, LmcuRenderStats* st) const override {
        LmcuRenderParams q = *p;
        q.bdpt_strategies = Strategies;
        return lmcu_render_bdpt(h, &q, target, st);
    }
};
using Renderer_PT_Naive_Path_Cuda = Renderer_PathStrategy_Cuda<LMCU_BDPT_PT_NAIVE>;
using Renderer_PT_NEE_Path_Cuda = Renderer_PathStrategy_Cuda<LMCU_BDPT_PT_NEE>;
using Renderer_LT_NEE_Path_Cuda = Renderer_PathStrategy_Cuda<LMCU_BDPT_LT_NEE>;

LM_COMP_REG_IMPL(Renderer_PT_Naive_Path_Cuda, "renderer::pt_naive_path_cuda");
LM_COMP_REG_IMPL(Renderer_PT_NEE_Path_Cuda, "renderer::pt_nee_path_cuda");
LM_COMP_REG_IMPL(Renderer_LT_NEE_Path_Cuda, "renderer::lt_nee_path_cuda");

LM_NAMESPACE_END(LM_NAMESPACE)

// Batched rays for hosts that hold numpy / torch buffers (func_accel_consistency-style notebooks): `accel_loc` is the
// locator of an accel::cuda_bvh asset that Scene::build() has built; rays / hits are the POD records of include/lmcu.h.
// Returns 0, or -1 with the message in `err`.
extern "C" int lm_cuda_accel_intersect_batch(const char* accel_loc, const LmcuRay* rays, long long n, LmcuHit* hits, char* err, int cap) {
    try {
        auto* accel = dynamic_cast<lm::Accel_CudaBVH*>(lm::comp::get<lm::Accel>(accel_loc));
        if (!accel) throw std::runtime_error("not an accel::cuda_bvh asset");
        accel->intersect_batch(rays, n, hits);
        return 0;
    }
    catch (const std::exception& e) {
        if (err && cap > 0) std::snprintf(err, size_t(cap), "%s", e.what());
        return -1;
    }
}

// Test / diagnostics hook: how often a component found its scene already on the device, and how often it had to export,
// upload and build it.  `clear` != 0 also drops the cached device scenes.
extern "C" void lm_cuda_cache_stats(long long* hits, long long* misses, int clear) {
    if (hits) *hits = lm::g_scene_cache.hits;
    if (misses) *misses = lm::g_scene_cache.misses;
    if (clear) lm::g_scene_cache.clear();
}

// Host-logic test hook (tests/test_plugin_probe.py): what the plugin recovers from an lm::Scene through the public
// component interfaces, as JSON, without touching a GPU.  Returns the length written, or -1.
extern "C" int lm_cuda_probe_scene(const char* scene_loc, char* out, int cap) {
    try {
        auto* scene = lm::comp::get<lm::Scene>(scene_loc);
        if (!scene) return -1;
        lm::SceneExport ex;
        ex.gather(*scene, true);
        lm::Json j;
        j["materials"] = lm::Json::array();
        for (const auto& m : ex.mats) {
            lm::Json e{ {"type", m.type}, {"map_kd", m.map_kd}, {"color", { m.color[0], m.color[1], m.color[2] }}, {"ax", m.ax}, {"ay", m.ay},
                        {"Ni", m.Ni}, {"Ks", { m.Ks[0], m.Ks[1], m.Ks[2] }}, {"alpha_mask", m.alpha_mask}, {"entries", lm::Json::array()} };
            for (int k = 0; k < m.num_entries; k++) e["entries"].push_back({ m.entry_material[k], m.entry_weight[k] });
            j["materials"].push_back(e);
        }
        j["textures"] = lm::Json::array();
        for (const auto& t : ex.textures) {
            double sum = 0;
            for (size_t i = 0; i < size_t(t.w) * t.h * t.c; i++) sum += t.data[i];
            j["textures"].push_back({ {"w", t.w}, {"h", t.h}, {"c", t.c}, {"sum", sum} });
        }
        j["primitive_materials"] = lm::Json::array();
        for (const auto& p : ex.prims) j["primitive_materials"].push_back(p.material);
        double tsum = 0;
        for (const auto& t : ex.mesh_t) for (double v : t) tsum += v;
        j["texcoord_sum"] = tsum;
        const std::string text = j.dump();
        if (int(text.size()) + 1 > cap) return -1;
        std::memcpy(out, text.c_str(), text.size() + 1);
        return int(text.size());
    }
    catch (const std::exception& e) {
        std::snprintf(out, size_t(cap), "%s", e.what());
        return -1;
    }
}
