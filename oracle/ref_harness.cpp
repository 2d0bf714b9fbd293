// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// C-ABI driver around the UNMODIFIED reference (oracle/_ref/liblm.so, built from
// /root/reference/src by oracle/Makefile).  It turns an LmcuSceneDesc (include/lmcu.h) into
// Lightmetrica assets through the reference's public user API (lm::load<T>, Scene::add_*),
// builds with the reference's accel::sahbvh (or any registered accel key, e.g. accel::cuda_bvh
// once the plugin is loaded) and exposes batched Accel::intersect and Renderer::render.  Nothing
// of the algorithms lives here: rays go through lm::Accel::intersect (include/lm/accel.h:70)
// and pixels through lm::Renderer::render (include/lm/renderer.h:25).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may load
// the resulting liblmref.so.
#include <lm/lm.h>
#include <omp.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include "../include/lmcu.h"

using namespace lm;
using namespace lm::literals;

LM_NAMESPACE_BEGIN(LM_NAMESPACE)

// Triangle-soup mesh fed from caller memory.  Same contract as mesh::raw
// (src/mesh/mesh_raw.cpp:37-104) without routing millions of numbers through JSON.
class Mesh_Soup final : public Mesh {
private:
    std::vector<Vec3> ps_, ns_;
    std::vector<Vec2> ts_;
    int nf_ = 0;
public:
    LM_SERIALIZE_IMPL(ar) {                    // like mesh::raw (src/mesh/mesh_raw.cpp:45-47): the state round-trips
        ar(ps_, ns_, ts_, nf_);
    }
    virtual void construct(const Json& prop) override {
        nf_ = json::value<int>(prop, "num_faces");
        const auto* p = reinterpret_cast<const double*>(json::value<std::uintptr_t>(prop, "p"));
        const auto* n = reinterpret_cast<const double*>(json::value<std::uintptr_t>(prop, "n"));
        const auto* t = reinterpret_cast<const double*>(json::value<std::uintptr_t>(prop, "t", std::uintptr_t(0)));
        ps_.resize(size_t(nf_) * 3);
        ns_.resize(size_t(nf_) * 3);
        ts_.assign(size_t(nf_) * 3, Vec2());
        if (t) for (size_t i = 0; i < size_t(nf_) * 3; i++) ts_[i] = Vec2(t[2*i], t[2*i+1]);
        for (size_t i = 0; i < size_t(nf_) * 3; i++) {
            ps_[i] = Vec3(p[3*i], p[3*i+1], p[3*i+2]);
        }
        for (size_t f = 0; f < size_t(nf_); f++) {
            for (int k = 0; k < 3; k++) {
                const size_t i = 3*f + k;
                ns_[i] = n ? Vec3(n[3*i], n[3*i+1], n[3*i+2])
                           : math::geometry_normal(ps_[3*f], ps_[3*f+1], ps_[3*f+2]);
            }
        }
    }
    virtual void foreach_triangle(const ProcessTriangleFunc& process_triangle) const override {
        for (int fi = 0; fi < nf_; fi++) {
            process_triangle(fi, triangle_at(fi));
        }
    }
    virtual Tri triangle_at(int face) const override {
        const size_t i = size_t(face) * 3;
        return { { ps_[i], ns_[i], ts_[i] }, { ps_[i+1], ns_[i+1], ts_[i+1] }, { ps_[i+2], ns_[i+2], ts_[i+2] } };
    }
    virtual InterpolatedPoint surface_point(int face, Vec2 uv) const override {
        const size_t i = size_t(face) * 3;
        return {
            math::mix_barycentric(ps_[i], ps_[i+1], ps_[i+2], uv),
            glm::normalize(math::mix_barycentric(ns_[i], ns_[i+1], ns_[i+2], uv)),
            math::geometry_normal(ps_[i], ps_[i+1], ps_[i+2]),
            math::mix_barycentric(ts_[i], ts_[i+1], ts_[i+2], uv)     // mesh_raw.cpp:94
        };
    }
    virtual int num_triangles() const override { return nf_; }
};
LM_COMP_REG_IMPL(Mesh_Soup, "mesh::lmref_soup");

LM_NAMESPACE_END(LM_NAMESPACE)

namespace {

thread_local std::string g_err;
bool g_init = false;
int g_scene_serial = 0;

struct RefScene {
    Scene* scene = nullptr;
    std::string scene_loc;
    std::map<int, int> node_to_desc;   // lm SceneNode::index -> LmcuPrimitive::node_index
    std::string tag;
};

Json vec3(const double* v) { return Json::array({ v[0], v[1], v[2] }); }

template <typename F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
    catch (...) { g_err = "unknown exception"; return 1; }
}

}

extern "C" {

const char* lmref_last_error() { return g_err.c_str(); }

// num_threads <= 0: hardware_concurrency() + num_threads (src/parallel/parallel_openmp.cpp:25-29)
int lmref_init(int num_threads, int verbose) {
    return guarded([&] {
        if (!g_init) {
            lm::init();
            g_init = true;
        }
        lm::parallel::init(lm::parallel::DefaultType, { {"num_threads", num_threads} });
        lm::log::set_severity(verbose ? int(lm::log::LogLevel::Info) : 1000);
    });
}

int lmref_num_threads() { return lm::parallel::num_threads(); }

int lmref_load_plugin(const char* path) {
    return guarded([&] { lm::comp::load_plugin(path); });
}

// Builds assets + scene graph from the description and calls Scene::build() (src/scene.cpp:274).
// One scene lives at a time (the asset group is reset).  build_seconds = wall time of build().
int lmref_scene_create(const LmcuSceneDesc* d, const char* accel_key, void** out, double* build_seconds) {
    return guarded([&] {
        lm::reset();
        auto* rs = new RefScene;
        rs->tag = std::to_string(g_scene_serial++);
        std::vector<Mesh*> meshes;
        for (int i = 0; i < d->num_meshes; i++) {
            const auto& m = d->meshes[i];
            meshes.push_back(lm::load<Mesh>("mesh" + std::to_string(i), "mesh::lmref_soup", {
                {"num_faces", m.num_faces},
                {"p", std::uintptr_t(m.p)},
                {"n", std::uintptr_t(m.n)},
                {"t", std::uintptr_t(m.t)} }));
        }
        std::vector<Texture*> textures;
        for (int i = 0; i < d->num_textures; i++) {
            const auto& t = d->textures[i];
            if (!t.path) LM_THROW_EXCEPTION(Error::InvalidArgument, "texture without a file path");
            textures.push_back(lm::load<Texture>("tex" + std::to_string(i), "texture::bitmap", { {"path", std::string(t.path)} }));
        }
        const auto tex_loc = [&](int map_kd) { return map_kd > 0 ? textures.at(map_kd - 1)->loc() : std::string(); };
        std::vector<Material*> mats;
        for (int i = 0; i < d->num_materials; i++) {
            const auto& m = d->materials[i];
            const auto name = "mat" + std::to_string(i);
            switch (m.type) {
            case LMCU_MAT_DIFFUSE: mats.push_back(lm::load<Material>(name, "material::diffuse", { {"Kd", vec3(m.color)}, {"mapKd", tex_loc(m.map_kd)} })); break;
            case LMCU_MAT_MASK:    mats.push_back(lm::load<Material>(name, "material::mask", {})); break;
            case LMCU_MAT_MIXTURE_OBJ:
                mats.push_back(lm::load<Material>(name, "material::mixture_wavefrontobj", {
                    {"Kd", vec3(m.color)}, {"mapKd", tex_loc(m.map_kd)}, {"Ks", vec3(m.Ks)}, {"ax", m.ax}, {"ay", m.ay},
                    {"no_alpha_mask", !m.alpha_mask} }));
                break;
            case LMCU_MAT_MIXTURE_RR:
            case LMCU_MAT_MIXTURE_MARGINALIZED: {
                // entries refer to materials by index: they must precede the mixture in the table
                Json entries = Json::array();
                for (int k = 0; k < m.num_entries; k++) {
                    if (m.entry_material[k] >= i) LM_THROW_EXCEPTION(Error::InvalidArgument, "mixture entries must precede the mixture");
                    entries.push_back({ {"material", mats.at(m.entry_material[k])->loc()}, {"weight", m.entry_weight[k]} });
                }
                mats.push_back(lm::load<Material>(name, m.type == LMCU_MAT_MIXTURE_RR ? "material::constant_weight_mixture_rr"
                                                                                       : "material::constant_weight_mixture_marginalized", entries));
                break;
            }
            case LMCU_MAT_GLOSSY:  mats.push_back(lm::load<Material>(name, "material::glossy", { {"Ks", vec3(m.color)}, {"ax", m.ax}, {"ay", m.ay} })); break;
            case LMCU_MAT_GLASS:   mats.push_back(lm::load<Material>(name, "material::glass", { {"Ni", m.Ni} })); break;
            case LMCU_MAT_MIRROR:  mats.push_back(lm::load<Material>(name, "material::mirror", {})); break;
            default: LM_THROW_EXCEPTION(Error::InvalidArgument, "bad material type");
            }
        }
        auto* accel = lm::load<Accel>("accel", accel_key, {});
        if (!accel) LM_THROW_EXCEPTION(Error::InvalidArgument, "unknown accel");
        auto* scene = lm::load<Scene>("scene", "scene::default", { {"accel", accel->loc()} });
        rs->scene = scene;
        rs->scene_loc = scene->loc();
        if (d->camera) {
            const auto& c = *d->camera;
            auto* cam = lm::load<Camera>("camera", "camera::pinhole", {
                {"position", vec3(c.position)}, {"center", vec3(c.center)}, {"up", vec3(c.up)},
                {"vfov", c.vfov}, {"aspect", c.aspect} });
            scene->add_primitive({ {"camera", cam->loc()} });
        }
        // Primitives in description order; the scene light table (src/scene.cpp:281-286) then
        // lists area lights in that order, followed by the environment light added last.
        int next_light = 0;
        for (int i = 0; i < d->num_primitives; i++) {
            const auto& p = d->primitives[i];
            Json prop{ {"mesh", meshes.at(p.mesh)->loc()}, {"material", mats.at(p.material)->loc()} };
            if (p.light >= 0) {
                if (p.light != next_light) LM_THROW_EXCEPTION(Error::InvalidArgument, "lights must follow primitive order");
                next_light++;
                const auto& l = d->lights[p.light];
                auto* light = lm::load<Light>("light" + std::to_string(p.light), "light::area", {
                    {"Ke", vec3(l.Ke)}, {"mesh", meshes.at(p.mesh)->loc()} });
                prop["light"] = light->loc();
            }
            Mat4 M;
            for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) M[c][r] = p.M[c*4 + r];
            const int before = scene->num_nodes();
            scene->add_transformed_primitive(M, prop);
            rs->node_to_desc[before + 1] = p.node_index;   // group node, then the primitive node
        }
        for (int i = next_light; i < d->num_lights; i++) {
            const auto& l = d->lights[i];
            Light* light = nullptr;
            if (l.type == LMCU_LIGHT_ENVCONST) {
                light = lm::load<Light>("light" + std::to_string(i), "light::envconst", { {"Le", vec3(l.Ke)} });
            } else if (l.type == LMCU_LIGHT_ENV) {
                light = lm::load<Light>("light" + std::to_string(i), "light::env", {
                    {"path", std::string(l.env_path ? l.env_path : "")}, {"rot", l.env_rot}, {"scale", l.env_scale} });
            } else if (l.type == LMCU_LIGHT_POINT) {
                light = lm::load<Light>("light" + std::to_string(i), "light::point", { {"Le", vec3(l.Ke)}, {"position", vec3(l.position)} });
            } else if (l.type == LMCU_LIGHT_DIRECTIONAL) {
                light = lm::load<Light>("light" + std::to_string(i), "light::directional", { {"Le", vec3(l.Ke)}, {"direction", vec3(l.position)} });
            } else {
                LM_THROW_EXCEPTION(Error::InvalidArgument, "area light without primitive");
            }
            scene->add_primitive({ {"light", light->loc()} });
        }
        const auto t0 = std::chrono::high_resolution_clock::now();
        scene->build();
        const auto t1 = std::chrono::high_resolution_clock::now();
        if (build_seconds) *build_seconds = std::chrono::duration<double>(t1 - t0).count();
        *out = rs;
    });
}

void lmref_scene_destroy(void* s) { delete static_cast<RefScene*>(s); }

// Locator of the lm::Scene asset (for hooks that look the scene up themselves, e.g. the plugin's probe test)
const char* lmref_scene_loc(void* s) { return static_cast<RefScene*>(s)->scene_loc.c_str(); }

// The sequence of functest/func_serial_consistency.py:79-83: lm::save_state_to_file, lm::reset, lm::load_state_from_file.
// The handle then re-resolves its scene by locator (every asset was re-created by the archive).
int lmref_save_reset_load(void* s, const char* path) {
    return guarded([&] {
        auto* rs = static_cast<RefScene*>(s);
        lm::save_state_to_file(path);
        lm::reset();
        rs->scene = nullptr;
        lm::load_state_from_file(path);
        rs->scene = lm::comp::get<Scene>(rs->scene_loc);
        if (!rs->scene) LM_THROW_EXCEPTION(Error::IOError, "scene missing after load_state_from_file");
    });
}

// rays[i] -> Accel::intersect(ray, tmin, tmax) (include/lm/accel.h:70) under OpenMP.
int lmref_accel_intersect(void* s, const LmcuRay* rays, int64_t n, LmcuHit* hits, double* seconds) {
    return guarded([&] {
        auto* rs = static_cast<RefScene*>(s);
        const Accel* accel = rs->scene->accel();
        const auto t0 = std::chrono::high_resolution_clock::now();
        #pragma omp parallel for schedule(dynamic, 4096) num_threads(lm::parallel::num_threads())
        for (int64_t i = 0; i < n; i++) {
            const auto& r = rays[i];
            const auto h = accel->intersect(Ray{ Vec3(r.o[0], r.o[1], r.o[2]), Vec3(r.d[0], r.d[1], r.d[2]) }, r.tmin, r.tmax);
            if (!h) {
                hits[i] = LmcuHit{ r.tmax, 0, 0, -1, -1 };
            } else {
                const auto it = rs->node_to_desc.find(h->primitive);
                hits[i] = LmcuHit{ h->t, h->uv.x, h->uv.y, it == rs->node_to_desc.end() ? h->primitive : it->second, h->face };
            }
        }
        const auto t1 = std::chrono::high_resolution_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    });
}

// renderer_key: "renderer::raycast" | "renderer::ao" | "renderer::pt" | plugin keys.
// film_rgb: width*height*3 doubles (Film::buffer(), src/film/film_bitmap.cpp:212-218).
int lmref_render(void* s, const char* renderer_key, const LmcuRenderParams* p, double* film_rgb,
                 double* elapsed, long long* processed) {
    return guarded([&] {
        auto* rs = static_cast<RefScene*>(s);
        const char* film_key = std::getenv("LMREF_FILM");                  // tests: e.g. the plugin's film::cuda
        auto* film = lm::load<Film>("film", film_key ? film_key : "film::bitmap", { {"w", p->width}, {"h", p->height} });
        if (!film) LM_THROW_EXCEPTION(Error::InvalidArgument, "unknown film");
        if (const char* e = std::getenv("LMREF_FILM_PREFILL")) {           // tests: a film that already holds an image
            const double v = std::atof(e);                                 // (renderer::ao never clears its film)
            for (int y = 0; y < p->height; y++) for (int x = 0; x < p->width; x++) film->set_pixel(x, y, Vec3(v));
        }
        static const char* modes[] = { "naive", "nee", "mis" };
        Json prop{
            {"scene", rs->scene_loc}, {"output", film->loc()},
            {"spp", (long long)p->spp}, {"max_verts", p->max_verts}, {"seed", p->seed},
            {"scheduler", std::string(p->scheduler == LMCU_SCHED_TIME ? "time" : "sample")},
            {"primary_ray_sampling_mode", std::string(p->primary_ray_sampling_mode == LMCU_PRIMARY_IMAGE ? "image" : "pixel")},
            {"num_samples", (long long)p->num_samples}, {"render_time", p->render_time},
            {"sampling_mode", std::string(modes[p->sampling_mode])},   // a bare const char* would hit lm's pointer serializer
            {"bg_color", vec3(p->bg_color)}, {"use_constant_color", bool(p->use_constant_color)},
            {"visualize_normal", bool(p->visualize_normal)} };
        if (p->has_color) prop["color"] = vec3(p->color);
        if (p->samples_per_iter > 0) prop["samples_per_iter"] = (long long)p->samples_per_iter;
        if (const char* e = std::getenv("LMREF_EXTRA_PROPS")) {            // tests: extra JSON keys for plugin renderers
            const Json extra = Json::parse(e);
            for (auto it = extra.begin(); it != extra.end(); ++it) prop[it.key()] = it.value();
        }
        auto* renderer = lm::load<Renderer>("renderer", renderer_key, prop);
        if (!renderer) LM_THROW_EXCEPTION(Error::InvalidArgument, "unknown renderer");
        const auto t0 = std::chrono::high_resolution_clock::now();
        const auto out = renderer->render();
        const auto t1 = std::chrono::high_resolution_clock::now();
        if (elapsed) *elapsed = out.contains("elapsed") ? out["elapsed"].get<double>()
                                                        : std::chrono::duration<double>(t1 - t0).count();
        if (processed) *processed = out.contains("processed") ? out["processed"].get<long long>() : 0;
        const auto buf = film->buffer();
        std::memcpy(film_rgb, buf.data, sizeof(double) * size_t(buf.w) * buf.h * 3);
    });
}

}
