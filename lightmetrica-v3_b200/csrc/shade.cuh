// shade.cuh — FP32 device restatement of the sampling/evaluation functions the renderers call per
// vertex: camera, the four BSDFs, area / constant-environment / image-environment lights, the
// counter-based RNG.  Measures follow the reference: BSDF pdfs are in projected solid angle and
// weights are f/p (doc/path_sampling.rst:269-292).  Citations are relative to /root/reference.
#pragma once
#include "common.cuh"

namespace lmcu {

// -- counter-based RNG: Philox4x32-10 keyed by (seed), counter (pixel, sample, dimension block) ------
struct U4 { uint32_t x, y, z, w; };

LMCU_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

LMCU_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

struct R4 { float x, y, z, w; };   // four uniforms in [0,1)

LMCU_HD float u01(uint32_t v) { return (float)(v >> 8) * 5.9604644775390625e-8f; }   // 2^-24

// sample = global sample index of the pixel (disjoint across GPUs in a sample split)
LMCU_HD R4 rand4(uint32_t seed, uint32_t pixel, uint64_t sample, uint32_t block) {
    U4 c; c.x = pixel; c.y = (uint32_t)sample; c.z = (uint32_t)(sample >> 32); c.w = block;
    const U4 r = philox4x32_10(c, seed, 0x1b873593u);
    R4 o; o.x = u01(r.x); o.y = u01(r.y); o.z = u01(r.z); o.w = u01(r.w);
    return o;
}

// -- geometry helpers -------------------------------------------------------------------------------
// Duff et al. 2017 (include/lm/math.h:540-547)
LMCU_HD void orthonormal_basis(F3 n, F3& u, F3& v) {
    const float s = copysignf(1.0f, n.z);
    const float a = -1.0f / (s + n.z);
    const float b = n.x * n.y * a;
    u = f3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
    v = f3(b, s + n.y * n.y * a, -n.y);
}

struct Frame { F3 n, u, v; };

// PointGeometry::orthonormal_basis_twosided (include/lm/surface.h:190-195)
LMCU_HD Frame frame_twosided(F3 n, F3 wi) {
    Frame f;
    F3 u, v;
    orthonormal_basis(n, u, v);
    const bool up = dot(wi, n) > 0.f;
    f.n = up ? n : -n; f.u = u; f.v = up ? v : -v;
    return f;
}

LMCU_HD float safe_sqrt(float v) { return sqrtf(fmaxf(0.f, v)); }

// math::sample_cosine_weighted (include/lm/math.h:645-652)
LMCU_HD F3 sample_cosine_weighted(float u0, float u1) {
    const float r = safe_sqrt(u0);
    const float t = 2.0f * kPiF * u1;
    float s, c;
#if defined(__CUDA_ARCH__)
    sincosf(t, &s, &c);
#else
    s = sinf(t); c = cosf(t);
#endif
    const float x = r * c, y = r * s;
    return f3(x, y, safe_sqrt(1.0f - x * x - y * y));
}

// math::sample_uniform_sphere (include/lm/math.h:660-667)
LMCU_HD F3 sample_uniform_sphere(float u0, float u1) {
    const float z = 1.0f - 2.0f * u0;
    const float r = safe_sqrt(1.0f - z * z);
    const float t = 2.0f * kPiF * u1;
    return f3(r * cosf(t), r * sinf(t), z);
}

// math::sample_uniform_disk (include/lm/math.h:632-638)
LMCU_HD void sample_uniform_disk(float u0, float u1, float& x, float& y) {
    const float r = safe_sqrt(u0);
    const float t = 2.0f * kPiF * u1;
    x = r * cosf(t); y = r * sinf(t);
}

LMCU_HD F3 reflection(F3 w, F3 n) { return 2.0f * dot(w, n) * n - w; }   // math.h:497-499

// surface::shading_normal_correction (include/lm/surface.h:424-438).  le = false: camera->light transport (TransDir::EL,
// the path tracer): 0 or 1.  le = true: light->camera transport (TransDir::LE, the light tracer): Veach's energy
// compensation |wi.ns||wo.ng| / (|wo.ns||wi.ng|).
LMCU_HD float sn_correction(F3 n, F3 gn, F3 wi, F3 wo, bool le = false) {
    const float wi_ng = dot(wi, gn), wo_ng = dot(wo, gn), wi_ns = dot(wi, n), wo_ns = dot(wo, n);
    if (wi_ng * wi_ns <= 0.f || wo_ng * wo_ns <= 0.f) return 0.f;
    return le ? wi_ns * wo_ng / (wo_ns * wi_ng) : 1.f;
}

LMCU_HD bool opposite(F3 n, F3 w1, F3 w2) { return dot(w1, n) * dot(w2, n) <= 0.f; }   // surface.h:178-180

// -- materials ----------------------------------------------------------------------------------------
enum { MAT_DIFFUSE = 0, MAT_GLOSSY = 1, MAT_GLASS = 2, MAT_MIRROR = 3, MAT_MASK = 4,
       MAT_MIXTURE_OBJ = 5, MAT_MIXTURE_RR = 6, MAT_MIXTURE_MARGINALIZED = 7 };

struct alignas(16) MaterialDev {
    int32_t type;
    float r, g, b;              // Kd / Ks
    float ax, ay, Ni;
    int32_t tex_kd;             // diffuse, mixture_obj: TextureDev index of mapKd, -1 = none
    float ks_r, ks_g, ks_b;     // mixture_obj: Ks of the glossy lobe
    int32_t alpha_mask;         // mixture_obj: 1 = the alpha channel of mapKd masks the surface
    int32_t entry_offset;       // constant-weight mixtures: first MixEntry
    int32_t num_entries;        // ... number of entries; marginalized: the non-specular ones come first
    int32_t num_group0;         // marginalized: entries of the default (non-specular) group
    int32_t pad_;
};
static_assert(sizeof(MaterialDev) == 64, "MaterialDev must be 64 bytes");

// One (material, weight) entry of a constant-weight mixture (src/material/material_mixture.cpp:27-36,126-145)
struct alignas(16) MixEntry {
    int32_t material;           // index into DevScene::mats (diffuse | glossy | mirror | mask)
    float weight;               // multiplies the entry's BSDF
    float prob;                 // rr: pmf of the entry as a component; marginalized: pmf inside its group
    float group_prob;           // marginalized: pmf of the entry's group as a component
};

// texture::bitmap (src/texture/texture_bitmap.cpp:41-96); texels live in one float4 pool (alpha = 1 without a channel)
struct alignas(16) TextureDev { int32_t w, h, offset, has_alpha; };

LMCU_HD bool mat_is_specular(int type) { return type == MAT_GLASS || type == MAT_MIRROR || type == MAT_MASK; }

// material_glossy.cpp:43-46
LMCU_HD float ggx_D(const MaterialDev& m, F3 wh, const Frame& f) {
    const float a = dot(wh, f.u) / m.ax, b = dot(wh, f.v) / m.ay, c = dot(wh, f.n);
    const float s = a * a + b * b + c * c;
    return 1.0f / (kPiF * m.ax * m.ay * s * s);
}
// material_glossy.cpp:49-59
LMCU_HD float ggx_G1(const MaterialDev& m, F3 w, const Frame& f) {
    const float c = dot(w, f.n);
    const float s = fmaxf((float)kEps, safe_sqrt(1.0f - c * c));
    const float cp = dot(w, f.u) / s, cs = dot(w, f.v) / s;
    const float a2 = (cp * m.ax) * (cp * m.ax) + (cs * m.ay) * (cs * m.ay);
    const float sc = s / c;
    return c == 0.f ? 0.f : 2.0f / (1.0f + safe_sqrt(1.0f + a2 * sc * sc));
}

// Material::eval for non-delta lobes (diffuse: material_diffuse.cpp:86-91; glossy: material_glossy.cpp:101-109)
LMCU_HD F3 mat_eval(const MaterialDev& m, F3 n, F3 wi, F3 wo) {
    if (opposite(n, wi, wo)) return f3s(0.f);
    if (m.type == MAT_DIFFUSE) return f3(m.r, m.g, m.b) * (1.0f / kPiF);
    if (m.type == MAT_GLOSSY) {
        const F3 wh = normalize(wi + wo);
        const Frame f = frame_twosided(n, wi);
        const float p5 = powf(1.0f - dot(wo, wh), 5.0f);
        const F3 Ks = f3(m.r, m.g, m.b);
        const F3 Fr = Ks + (f3s(1.f) - Ks) * p5;
        const float k = ggx_D(m, wh, f) * ggx_G1(m, wi, f) * ggx_G1(m, wo, f) / (4.0f * dot(wi, f.n) * dot(wo, f.n));
        return Ks * Fr * k;
    }
    return f3s(0.f);   // delta lobes evaluate to zero with eval_delta=true
}

// Material::pdf_direction, projected solid angle (material_diffuse.cpp:82-84, material_glossy.cpp:92-99)
LMCU_HD float mat_pdf(const MaterialDev& m, F3 n, F3 wi, F3 wo) {
    if (opposite(n, wi, wo)) return 0.f;
    if (m.type == MAT_DIFFUSE) return 1.0f / kPiF;
    if (m.type == MAT_GLOSSY) {
        const F3 wh = normalize(wi + wo);
        const Frame f = frame_twosided(n, wi);
        return ggx_D(m, wh, f) * dot(wh, f.n) / (4.0f * dot(wo, wh) * dot(wo, f.n));
    }
    return 0.f;
}

// Fresnel pick of material::glass (material_glass.cpp:85-136): returns Fr, refracted dir in wt, eta.
LMCU_HD float glass_fresnel(const MaterialDev& m, F3 n, F3 wi, F3& wt, float& eta, bool& total) {
    const bool in = dot(wi, n) > 0.f;
    const F3 nn = in ? n : -n;
    eta = in ? 1.0f / m.Ni : m.Ni;
    const float t = dot(wi, nn);
    const float t2 = 1.0f - eta * eta * (1.0f - t * t);
    total = t2 <= 0.f;
    if (total) { wt = f3s(0.f); return 1.0f; }
    wt = eta * (nn * t - wi) - nn * safe_sqrt(t2);
    const float c = in ? dot(wi, n) : dot(wt, n);
    const float r = (1.0f - m.Ni) / (1.0f + m.Ni);
    const float r2 = r * r;
    return r2 + (1.0f - r2) * powf(1.0f - c, 5.0f);
}

struct BsdfSample {
    F3 wo;
    F3 weight;        // component weight * f/p * shading-normal correction
    float pdf;        // pdf_direction in projected solid angle (0 for delta lobes)
    bool specular;
    bool valid;
};

// path::sample_component + path::sample_direction at a surface vertex (include/lm/path.h:193-200,378-422).
// u: (ud0, ud1, -, uc0)
// le: light->camera transport (TransDir::LE): no eta^2 radiance scaling on refraction (material_glass.cpp:85-87) and the
// LE form of the shading-normal correction.
LMCU_HD BsdfSample mat_sample(const MaterialDev& m, F3 n, F3 gn, F3 wi, float ud0, float ud1, float uc0, bool le = false) {
    BsdfSample s;
    s.valid = true; s.specular = false; s.pdf = 0.f;
    if (m.type == MAT_DIFFUSE) {                       // material_diffuse.cpp:68-76
        const Frame f = frame_twosided(n, wi);
        const F3 d = sample_cosine_weighted(ud0, ud1);
        s.wo = f.u * d.x + f.v * d.y + f.n * d.z;
        s.weight = f3(m.r, m.g, m.b);
        s.pdf = opposite(n, wi, s.wo) ? 0.f : 1.0f / kPiF;
    } else if (m.type == MAT_GLOSSY) {                 // material_glossy.cpp:70-86
        const Frame f = frame_twosided(n, wi);
        const float u1 = ud0 * 2.0f * kPiF, u2 = ud1;
        const F3 wh = normalize(safe_sqrt(u2 / (1.0f - u2)) * (m.ax * cosf(u1) * f.u + m.ay * sinf(u1) * f.v) + f.n);
        s.wo = reflection(wi, wh);
        if (opposite(n, wi, s.wo)) { s.valid = false; return s; }
        const F3 fv = mat_eval(m, n, wi, s.wo);
        s.pdf = mat_pdf(m, n, wi, s.wo);
        s.weight = fv * (1.0f / s.pdf);
    } else if (m.type == MAT_MIRROR) {                 // material_mirror.cpp:41-46
        s.wo = reflection(wi, n);
        s.weight = f3s(1.f);
        s.specular = true;
    } else if (m.type == MAT_MASK) {                   // material_mask.cpp:38-43: straight through
        s.wo = -wi;
        s.weight = f3s(1.f);
        s.specular = true;
    } else {                                           // material_glass.cpp:124-136,162-190
        F3 wt; float eta; bool total;
        const float Fr = glass_fresnel(m, n, wi, wt, eta, total);
        s.specular = true;
        if (uc0 < Fr) {
            s.wo = reflection(wi, n);
            s.weight = f3s(1.f);                       // (1/Fr) * Fr
        } else {
            s.wo = wt;
            s.weight = f3s(le ? 1.f : eta * eta);      // (1/(1-Fr)) * (1-Fr) * refr_correction (eta^2 for EL, 1 for LE)
        }
    }
    s.weight = s.weight * sn_correction(n, gn, wi, s.wo, le);
    return s;
}

// Material::reflectance of a plain lobe (raycast shading): Kd / Ks / 0
LMCU_HD F3 mat_reflectance(const MaterialDev& m) {
    return (m.type == MAT_DIFFUSE || m.type == MAT_GLOSSY) ? f3(m.r, m.g, m.b) : f3s(0.f);
}

// Texture::eval / eval_alpha (texture_bitmap.cpp:68-75,83-90): nearest texel at the fractional part of t
LMCU_HD float4 tex_fetch(const TextureDev& t, const float4* texels, float tu, float tv) {
    const float u = tu - floorf(tu), v = tv - floorf(tv);
    int x = (int)(u * (float)t.w), y = (int)(v * (float)t.h);
    x = x < 0 ? 0 : (x > t.w - 1 ? t.w - 1 : x);
    y = y < 0 ? 0 : (y > t.h - 1 ? t.h - 1 : y);
    return texels[t.offset + y * t.w + x];
}

// -- lights -------------------------------------------------------------------------------------------
enum { LIGHT_AREA = 0, LIGHT_ENVCONST = 1, LIGHT_ENV = 2, LIGHT_POINT = 3, LIGHT_DIRECTIONAL = 4 };

struct alignas(16) LightDev {
    int32_t type;
    int32_t prim;            // area: primitive ordinal
    int32_t tri_offset;      // area: first EmitTri
    int32_t num_tris;        // area: number of emitter triangles
    float Ke[3];             // area: Ke, envconst: Le
    float inv_area_j;        // area: invA / J  (light_area.cpp:39-44)
    int32_t cdf_offset;      // area: first entry of the triangle CDF (num_tris+1 floats)
    int32_t env_w, env_h;    // env
    float env_rot;           // env: radians
    float env_scale;
    int32_t env_cdf_offset;  // env: conditional CDFs [h][w+1] then marginal [h+1]
    int32_t env_tex_offset;  // env: texels (float4) offset
    int32_t pad_;
    float pos[3];            // point: position; directional: normalised direction the light travels in
    float pad2_;
};

struct alignas(16) EmitTri { float4 p1, p2, p3, n; };   // world-space emitter triangle + geometric normal

// Dist::sample (include/lm/math.h:367-370): upper_bound(c, u) - 1, clamped
LMCU_HD int cdf_sample(const float* c, int n, float u) {   // c has n+1 entries, c[0]=0
    int lo = 0, hi = n + 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (c[mid] <= u) lo = mid + 1; else hi = mid;
    }
    int i = lo - 1;
    return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}

// -- light::env (src/light/light_env.cpp:64-72,139-202): lat-long image, Dist2 importance sampling ---------
// (u,v) of a direction d pointing from the scene toward the sky (d = -geomL.wo), light_env.cpp:166-176,191-199
LMCU_HD void env_dir_to_uv(const LightDev& l, F3 d, float& u, float& v) {
    float at = atan2f(d.x, d.z);
    if (at < 0.f) at += 2.0f * kPiF;
    const float t = (at - l.env_rot) * 0.5f / kPiF;
    u = t - floorf(t);
    v = acosf(fminf(1.f, fmaxf(-1.f, d.y))) / kPiF;
}

// Light::eval: nearest texel of texture::bitmap (src/texture/texture_bitmap.cpp:68-75) times scale
LMCU_HD F3 env_eval(const LightDev& l, const float4* tex, F3 d) {
    float u, v;
    env_dir_to_uv(l, d, u, v);
    u = u - floorf(u); v = v - floorf(v);
    int x = (int)(u * (float)l.env_w), y = (int)(v * (float)l.env_h);
    x = x < 0 ? 0 : (x > l.env_w - 1 ? l.env_w - 1 : x);
    y = y < 0 ? 0 : (y > l.env_h - 1 ? l.env_h - 1 : y);
    const float4 t = tex[l.env_tex_offset + y * l.env_w + x];
    return f3(t.x, t.y, t.z) * l.env_scale;
}

// Dist2::pdf (include/lm/math.h:416-419); cdf layout: conditional rows [h][w+1], then marginal [h+1]
LMCU_HD float env_dist_pdf(const LightDev& l, const float* cdf, float u, float v) {
    const int w = l.env_w, h = l.env_h;
    int y = (int)(v * (float)h);
    y = y > h - 1 ? h - 1 : y;
    const int x = (int)(u * (float)w);
    if (y < 0 || x < 0 || x >= w) return 0.f;                       // Dist::pmf is 0 outside its range
    const float* m = cdf + l.env_cdf_offset + (size_t)h * (w + 1);
    const float* row = cdf + l.env_cdf_offset + (size_t)y * (w + 1);
    return (m[y + 1] - m[y]) * (row[x + 1] - row[x]) * (float)w * (float)h;
}

// Light::pdf_direct in projected solid angle at a surface point with shading normal n (light_env.cpp:164-181)
LMCU_HD float env_pdf_direct(const LightDev& l, const float* cdf, F3 n, F3 d, bool degenerated = false) {
    float u, v;
    env_dir_to_uv(l, d, u, v);
    const float st = safe_sqrt(1.0f - d.y * d.y);
    if (st == 0.f) return 0.f;
    const float pdf_sa = env_dist_pdf(l, cdf, u, v) / (2.0f * kPiF * kPiF * st);
    if (degenerated) return pdf_sa;                                 // surface.h:377-390: no cosine at a degenerated point
    const float J = fabsf(dot(n, d));
    return J == 0.f ? 0.f : pdf_sa / J;
}

// Light::sample_direct: returns the direction toward the sky (light_env.cpp:146-153); u = (ud.x, ud.y, up.x, up.y)
LMCU_HD F3 env_sample_dir(const LightDev& l, const float* cdf, R4 u) {
    const int w = l.env_w, h = l.env_h;
    const float* m = cdf + l.env_cdf_offset + (size_t)h * (w + 1);
    const int y = cdf_sample(m, h, u.x);
    const int x = cdf_sample(cdf + l.env_cdf_offset + (size_t)y * (w + 1), w, u.y);
    const float su = ((float)x + u.z) / (float)w, sv = ((float)y + u.w) / (float)h;
    const float t = kPiF * sv, p = 2.0f * kPiF * su + l.env_rot;
    const float st = sinf(t);
    return f3(st * sinf(p), cosf(t), st * cosf(p));
}

// -- camera (src/camera/camera_pinhole.cpp:90-95,102-109,128-132,207-226) ---------------------------------
struct CameraDev {
    double pos[3], u[3], v[3], w[3];
    double tf, aspect;
};

// Camera_Pinhole::raster_position + J (camera_pinhole.cpp:101-109,207-226), FP32: importance = pdf_direction of a
// world direction wo leaving the camera; 0 outside the view frustum.  rp = raster position in [0,1]^2.
LMCU_HD float camera_importance(const CameraDev& c, F3 wo, float& rpx, float& rpy) {
    const float ex = (float)c.u[0] * wo.x + (float)c.u[1] * wo.y + (float)c.u[2] * wo.z;
    const float ey = (float)c.v[0] * wo.x + (float)c.v[1] * wo.y + (float)c.v[2] * wo.z;
    const float ez = (float)c.w[0] * wo.x + (float)c.w[1] * wo.y + (float)c.w[2] * wo.z;
    if (ez >= 0.f) return 0.f;
    const float tf = (float)c.tf, aspect = (float)c.aspect;
    rpx = -ex / ez / tf / aspect * 0.5f + 0.5f;
    rpy = -ey / ez / tf * 0.5f + 0.5f;
    if (rpx < 0.f || rpx > 1.f || rpy < 0.f || rpy > 1.f) return 0.f;
    const float inv = -1.0f / ez;
    return inv * inv * inv / (tf * tf * aspect * 4.0f);
}

// FP64 primary ray with glm's operation order: bit-identical to Camera_Pinhole::primary_ray.
LMCU_HD void camera_primary_ray_d(const CameraDev& c, double rpx, double rpy, D3& o, D3& d) {
    const double rx = subrn(mulrn(2.0, rpx), 1.0), ry = subrn(mulrn(2.0, rpy), 1.0);
    const double vx = mulrn(mulrn(c.aspect, c.tf), rx), vy = mulrn(c.tf, ry), vz = -1.0;
    const double len2 = addrn(addrn(mulrn(vx, vx), mulrn(vy, vy)), mulrn(vz, vz));
    const double inv = divrn(1.0, sqrtrn(len2));                      // glm::inversesqrt
    const double dx = mulrn(vx, inv), dy = mulrn(vy, inv), dz = mulrn(vz, inv);
    o = d3(c.pos[0], c.pos[1], c.pos[2]);
    d = d3(addrn(addrn(mulrn(c.u[0], dx), mulrn(c.v[0], dy)), mulrn(c.w[0], dz)),
           addrn(addrn(mulrn(c.u[1], dx), mulrn(c.v[1], dy)), mulrn(c.w[1], dz)),
           addrn(addrn(mulrn(c.u[2], dx), mulrn(c.v[2], dy)), mulrn(c.w[2], dz)));
}

}  // namespace lmcu
