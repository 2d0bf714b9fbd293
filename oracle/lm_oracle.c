/*
 * oracle/lm_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, FP64 restatement of the reference's ray-tracing hot path, function by function, each
 * citing the file:line of /root/reference it follows.  It is the CHECKER for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the
 * product (lightmetrica-v3_b200/) never does and has no CPU fallback.
 *
 * Parity of this restatement itself is PINNED against the unmodified reference compiled into
 * oracle/_ref (tests/test_oracle_port.py: bit-exact closest hits and raycast images, pt images
 * bit-exact with the reference run single-threaded on the same mt19937 stream) and against the
 * committed golden outputs of the reference (tests/golden/reference_outputs.npz).
 *
 * Compile with -ffp-contract=off: the reference's default build has no FMA contraction.
 */
#include "lm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_INF 1e10   /* include/lm/math.h:54 */
#define ORC_EPS 1e-3   /* include/lm/math.h:55 */
#define ORC_PI 3.14159265358979323846

/* ------------------------------------------------------------------------------------------------
 * glm vector arithmetic with glm's operation order (vendored glm 0.9.9: detail/func_geometric.inl)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double x, y, z; } V3;

static V3 v3(double x, double y, double z) { V3 r = { x, y, z }; return r; }
static V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static V3 vmul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static V3 vscale(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
static V3 vdivs(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
static V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static double vdot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }            /* :54-60 */
static V3 vcross(V3 x, V3 y) {                                                              /* :74-85 */
    return v3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
static V3 vnormalize(V3 v) { return vscale(v, 1.0 / sqrt(vdot(v, v))); }                   /* :88-96 */
static double vlength(V3 v) { return sqrt(vdot(v, v)); }
static double vmaxc(V3 v) { return fmax(fmax(v.x, v.y), v.z); }
static int vis_zero(V3 v) { return v.x == 0.0 && v.y == 0.0 && v.z == 0.0; }              /* math.h:469-471 */
static double safe_sqrt(double v) { return sqrt(fmax(0.0, v)); }                           /* math.h:478-480 */
static double glm_max(double a, double b) { return (a < b) ? b : a; }                      /* glm::max */
static double glm_min(double a, double b) { return (b < a) ? b : a; }                      /* glm::min */

/* mat4 * vec4(p,1): (m0*x + m1*y) + (m2*z + m3*1)   glm/detail/type_mat4x4.inl:526-537 */
static V3 m4_point(const double* M, V3 p) {
    return v3((M[0] * p.x + M[4] * p.y) + (M[8] * p.z + M[12] * 1.0),
              (M[1] * p.x + M[5] * p.y) + (M[9] * p.z + M[13] * 1.0),
              (M[2] * p.x + M[6] * p.y) + (M[10] * p.z + M[14] * 1.0));
}
/* mat3 * vec3: m[0][r]*x + m[1][r]*y + m[2][r]*z   glm/detail/type_mat3x3.inl */
static V3 m3_mul(const double* N, V3 v) {
    return v3(N[0] * v.x + N[3] * v.y + N[6] * v.z, N[1] * v.x + N[4] * v.y + N[7] * v.z, N[2] * v.x + N[5] * v.y + N[8] * v.z);
}

/* ------------------------------------------------------------------------------------------------
 * Rng: std::mt19937 + std::uniform_real_distribution<double> (include/lm/math.h:204-220), as
 * implemented by libstdc++ (bits/random.tcc: generate_canonical<double,53> draws two 32-bit words).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint32_t mt[624]; int idx; } Rng;

static void rng_seed(Rng* r, uint32_t seed) {
    r->mt[0] = seed;
    for (int i = 1; i < 624; i++) r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
    r->idx = 624;
}
static uint32_t rng_u32(Rng* r) {
    if (r->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            const uint32_t y = (r->mt[i] & 0x80000000u) | (r->mt[(i + 1) % 624] & 0x7fffffffu);
            r->mt[i] = r->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
static double rng_u(Rng* r) {
    const double lo = (double)rng_u32(r);
    const double hi = (double)rng_u32(r);
    double ret = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}
void orc_rng_fill(uint32_t seed, double* out, int n) {
    Rng r; rng_seed(&r, seed);
    for (int i = 0; i < n; i++) out[i] = rng_u(&r);
}

/* ------------------------------------------------------------------------------------------------
 * Scene
 * ---------------------------------------------------------------------------------------------- */
typedef struct { V3 min, max; } Bound;                                 /* math.h:82-84 */
static Bound bound_empty(void) { Bound b = { { ORC_INF, ORC_INF, ORC_INF }, { -ORC_INF, -ORC_INF, -ORC_INF } }; return b; }
static Bound bound_merge_p(Bound b, V3 p) {                           /* math.h:166-168 */
    b.min = v3(glm_min(b.min.x, p.x), glm_min(b.min.y, p.y), glm_min(b.min.z, p.z));
    b.max = v3(glm_max(b.max.x, p.x), glm_max(b.max.y, p.y), glm_max(b.max.z, p.z));
    return b;
}
static Bound bound_merge(Bound a, Bound b) {                          /* math.h:176-178 */
    Bound r;
    r.min = v3(glm_min(a.min.x, b.min.x), glm_min(a.min.y, b.min.y), glm_min(a.min.z, b.min.z));
    r.max = v3(glm_max(a.max.x, b.max.x), glm_max(a.max.y, b.max.y), glm_max(a.max.z, b.max.z));
    return r;
}
static double bound_area(Bound b) {                                    /* math.h:109-112 */
    const V3 d = vsub(b.max, b.min);
    return 2.0 * (d.x * d.y + d.y * d.z + d.z * d.x);
}

typedef struct {          /* accel_sahbvh.cpp:26-49 */
    V3 p1, e1, e2;
    Bound b;
    V3 c;
    int prim;             /* index into desc primitives (flattened node) */
    int face;
} Tri;

typedef struct {          /* accel_sahbvh.cpp:78-88 */
    Bound b;
    int leaf, s, e, c1, c2;
} Node;

typedef struct {
    double M[16], NM[9], J;
    int node_index, mesh, material, light;
} Prim;

typedef struct {
    int type, prim;
    V3 Ke;
    double* cdf;          /* area: Dist over local triangle areas (light_area.cpp:68-73), n+1 entries */
    int ntri;
    double invA;
    /* light::env (src/light/light_env.cpp:52-77): lat-long image, rotation (radians), scale, Dist2 */
    int env_w, env_h;
    double env_rot, env_scale;
    float* env_tex;       /* [h][w][3], texture::bitmap order */
    double* env_rows;     /* Dist2::ds: h normalised CDFs of w+1 entries (math.h:398-410) */
    double* env_marg;     /* Dist2::m: h+1 entries */
    V3 pos;               /* light::point position / light::directional normalised direction */
} Light;

struct OrcScene {
    int nt, nn;
    Tri* tris;
    Node* nodes;
    int* indices;
    int nprims, nmats, nlights, nmeshes, env_light;
    Prim* prims;
    LmcuMaterial* mats;
    Light* lights;
    int* mesh_faces;
    double** mesh_p;
    double** mesh_n;      /* always present: geometric normal replicated when the description has none */
    double** mesh_t;      /* [faces][3][2] texture coordinates or NULL */
    int ntex;
    LmcuTexture* tex;     /* texture::bitmap texels (copied) */
    int** mix_mat;        /* per material: entry tables of the constant-weight mixtures (copied) */
    double** mix_w;
    int has_camera;
    V3 cam_pos, cam_u, cam_v, cam_w;
    double cam_tf, cam_aspect;
};

/* Transform ctor (math.h:741-745): normal_M = mat3(transpose(inverse(M))), J = |det(mat3(M))| */
static void make_transform(Prim* p) {
    const double* M = p->M;
    const double a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], i = M[10];
    const double det = a * (e * i - f * h) - d * (b * i - c * h) + g * (b * f - c * e);
    const double id = 1.0 / det;
    double* N = p->NM;
    N[0] = (e * i - f * h) * id; N[1] = -(d * i - f * g) * id; N[2] = (d * h - e * g) * id;
    N[3] = -(b * i - c * h) * id; N[4] = (a * i - c * g) * id; N[5] = -(a * h - b * g) * id;
    N[6] = (b * f - c * e) * id; N[7] = -(a * f - c * d) * id; N[8] = (a * e - b * d) * id;
    p->J = fabs(det);
}

/* -- SAH build (accel_sahbvh.cpp:156-256), single work queue processed in FIFO order ---------------- */
static const Tri* g_sort_tris;
static int g_sort_axis;
static int cmp_centroid(const void* a, const void* b) {
    const double ca = (&g_sort_tris[*(const int*)a].c.x)[g_sort_axis], cb = (&g_sort_tris[*(const int*)b].c.x)[g_sort_axis];
    if (ca < cb) return -1;
    if (cb < ca) return 1;
    return (*(const int*)a > *(const int*)b) - (*(const int*)a < *(const int*)b);   /* deterministic tie order */
}

static void build_bvh(OrcScene* sc) {
    const int nt = sc->nt;
    sc->nodes = (Node*)calloc((size_t)(2 * nt - 1 > 1 ? 2 * nt - 1 : 1), sizeof(Node));
    sc->indices = (int*)malloc(sizeof(int) * (size_t)nt);
    for (int i = 0; i < nt; i++) sc->indices[i] = i;
    for (int i = 0; i < 2 * nt - 1; i++) sc->nodes[i].b = bound_empty();
    typedef struct { int index, start, end; } Entry;
    Entry* q = (Entry*)malloc(sizeof(Entry) * (size_t)(2 * nt));
    int qh = 0, qt = 0, nn = 1;
    q[qt++] = (Entry){ 0, 0, nt };
    double* l = (double*)malloc(sizeof(double) * (size_t)(nt + 1));
    double* r = (double*)malloc(sizeof(double) * (size_t)(nt + 1));
    g_sort_tris = sc->tris;
    while (qh < qt) {
        const Entry en = q[qh++];
        const int s = en.start, e = en.end;
        Node* n = &sc->nodes[en.index];
        for (int i = s; i < e; i++) n->b = bound_merge(n->b, sc->tris[sc->indices[i]].b);     /* :192-195 */
        if (e - s < 2) { n->leaf = 1; n->s = s; n->e = e; continue; }                          /* :219-222 */
        double best = ORC_INF;
        int bi = -1, ba = -1;
        for (int a = 0; a < 3; a++) {                                                          /* :227-246 */
            g_sort_axis = a;
            qsort(sc->indices + s, (size_t)(e - s), sizeof(int), cmp_centroid);
            Bound bl = bound_empty(), br = bound_empty();
            for (int i = 0; i <= e - s; i++) {
                const int j = e - s - i;
                l[i] = bound_area(bl) * i;
                r[j] = bound_area(br) * i;
                if (i < e - s) bl = bound_merge(bl, sc->tris[sc->indices[s + i]].b);
                if (j > 0) br = bound_merge(br, sc->tris[sc->indices[s + j - 1]].b);
            }
            for (int i = 1; i < e - s; i++) {
                const double c = 1.0 + (l[i] + r[i]) / bound_area(n->b);
                if (c < best) { best = c; bi = i; ba = a; }
            }
        }
        if (best > e - s) { n->leaf = 1; n->s = s; n->e = e; continue; }                      /* :247-250 */
        g_sort_axis = ba;
        qsort(sc->indices + s, (size_t)(e - s), sizeof(int), cmp_centroid);
        const int m = s + bi;
        n->c1 = nn++; n->c2 = nn++;
        q[qt++] = (Entry){ n->c1, s, m };
        q[qt++] = (Entry){ n->c2, m, e };
    }
    sc->nn = nn;
    free(q); free(l); free(r);
}

OrcScene* orc_scene_create(const LmcuSceneDesc* d) {
    if (!d || d->num_primitives <= 0) return NULL;
    OrcScene* sc = (OrcScene*)calloc(1, sizeof(OrcScene));
    sc->nmeshes = d->num_meshes;
    sc->mesh_faces = (int*)malloc(sizeof(int) * (size_t)d->num_meshes);
    sc->mesh_p = (double**)malloc(sizeof(double*) * (size_t)d->num_meshes);
    sc->mesh_n = (double**)malloc(sizeof(double*) * (size_t)d->num_meshes);
    sc->mesh_t = (double**)calloc((size_t)d->num_meshes, sizeof(double*));
    for (int i = 0; i < d->num_meshes; i++) {
        if (d->meshes[i].t) {
            const size_t tc = (size_t)d->meshes[i].num_faces * 6;
            sc->mesh_t[i] = (double*)malloc(sizeof(double) * tc);
            memcpy(sc->mesh_t[i], d->meshes[i].t, sizeof(double) * tc);
        }
    }
    sc->ntex = d->num_textures;
    sc->tex = (LmcuTexture*)calloc((size_t)(d->num_textures > 0 ? d->num_textures : 1), sizeof(LmcuTexture));
    for (int i = 0; i < d->num_textures; i++) {
        const size_t cnt = (size_t)d->textures[i].w * d->textures[i].h * d->textures[i].c;
        float* data = (float*)malloc(sizeof(float) * cnt);
        memcpy(data, d->textures[i].data, sizeof(float) * cnt);
        sc->tex[i] = d->textures[i];
        sc->tex[i].data = data; sc->tex[i].path = NULL;
    }
    for (int i = 0; i < d->num_meshes; i++) {
        const size_t cnt = (size_t)d->meshes[i].num_faces * 9;
        sc->mesh_faces[i] = d->meshes[i].num_faces;
        sc->mesh_p[i] = (double*)malloc(sizeof(double) * cnt);
        sc->mesh_n[i] = (double*)malloc(sizeof(double) * cnt);
        memcpy(sc->mesh_p[i], d->meshes[i].p, sizeof(double) * cnt);
        if (d->meshes[i].n) {
            memcpy(sc->mesh_n[i], d->meshes[i].n, sizeof(double) * cnt);
        } else {
            for (int f = 0; f < d->meshes[i].num_faces; f++) {
                const double* p = sc->mesh_p[i] + (size_t)f * 9;
                const V3 a = v3(p[0], p[1], p[2]), b = v3(p[3], p[4], p[5]), c = v3(p[6], p[7], p[8]);
                const V3 gn = vnormalize(vcross(vsub(b, a), vsub(c, a)));
                for (int k = 0; k < 3; k++) { double* q = sc->mesh_n[i] + (size_t)f * 9 + 3 * k; q[0] = gn.x; q[1] = gn.y; q[2] = gn.z; }
            }
        }
    }
    sc->nprims = d->num_primitives;
    sc->prims = (Prim*)calloc((size_t)d->num_primitives, sizeof(Prim));
    int nt = 0;
    for (int i = 0; i < d->num_primitives; i++) {
        const LmcuPrimitive* p = &d->primitives[i];
        memcpy(sc->prims[i].M, p->M, sizeof(double) * 16);
        make_transform(&sc->prims[i]);
        sc->prims[i].node_index = p->node_index; sc->prims[i].mesh = p->mesh;
        sc->prims[i].material = p->material; sc->prims[i].light = p->light;
        nt += d->meshes[p->mesh].num_faces;
    }
    sc->nmats = d->num_materials;
    sc->mats = (LmcuMaterial*)malloc(sizeof(LmcuMaterial) * (size_t)(d->num_materials > 0 ? d->num_materials : 1));
    memcpy(sc->mats, d->materials, sizeof(LmcuMaterial) * (size_t)d->num_materials);
    sc->mix_mat = (int**)calloc((size_t)(d->num_materials > 0 ? d->num_materials : 1), sizeof(int*));
    sc->mix_w = (double**)calloc((size_t)(d->num_materials > 0 ? d->num_materials : 1), sizeof(double*));
    for (int i = 0; i < d->num_materials; i++) {
        const LmcuMaterial* m = &d->materials[i];
        if (m->num_entries <= 0) continue;
        sc->mix_mat[i] = (int*)malloc(sizeof(int) * (size_t)m->num_entries);
        sc->mix_w[i] = (double*)malloc(sizeof(double) * (size_t)m->num_entries);
        for (int k = 0; k < m->num_entries; k++) { sc->mix_mat[i][k] = m->entry_material[k]; sc->mix_w[i][k] = m->entry_weight[k]; }
        sc->mats[i].entry_material = sc->mix_mat[i]; sc->mats[i].entry_weight = sc->mix_w[i];
    }
    sc->nlights = d->num_lights;
    sc->lights = (Light*)calloc((size_t)(d->num_lights > 0 ? d->num_lights : 1), sizeof(Light));
    sc->env_light = -1;
    for (int i = 0; i < d->num_lights; i++) {
        const LmcuLight* l = &d->lights[i];
        Light* o = &sc->lights[i];
        o->type = l->type; o->prim = l->primitive; o->Ke = v3(l->Ke[0], l->Ke[1], l->Ke[2]);
        if (l->type == LMCU_LIGHT_AREA) {                                   /* light_area.cpp:62-75 */
            const int mesh = sc->prims[l->primitive].mesh;
            o->ntri = sc->mesh_faces[mesh];
            o->cdf = (double*)malloc(sizeof(double) * (size_t)(o->ntri + 1));
            o->cdf[0] = 0.0;
            for (int f = 0; f < o->ntri; f++) {
                const double* p = sc->mesh_p[mesh] + (size_t)f * 9;
                const V3 a = v3(p[0], p[1], p[2]), b = v3(p[3], p[4], p[5]), c = v3(p[6], p[7], p[8]);
                const V3 cr = vcross(vsub(b, a), vsub(c, a));
                o->cdf[f + 1] = o->cdf[f] + safe_sqrt(vdot(cr, cr)) * .5;
            }
            o->invA = 1.0 / o->cdf[o->ntri];
            const double sum = o->cdf[o->ntri];
            for (int f = 0; f <= o->ntri; f++) o->cdf[f] /= sum;            /* Dist::norm, math.h:341-346 */
        } else if (l->type == LMCU_LIGHT_POINT) {                            /* light_point.cpp:38-41 */
            o->pos = v3(l->position[0], l->position[1], l->position[2]);
        } else if (l->type == LMCU_LIGHT_DIRECTIONAL) {                      /* light_directional.cpp:37-40 */
            o->pos = vnormalize(v3(l->position[0], l->position[1], l->position[2]));
        } else {
            sc->env_light = i;
            if (l->type == LMCU_LIGHT_ENV) {                                 /* light_env.cpp:63-76 */
                const int w = l->env_w, h = l->env_h;
                o->env_w = w; o->env_h = h;
                o->env_rot = l->env_rot * ORC_PI / 180.0;
                o->env_scale = l->env_scale;
                o->env_tex = (float*)malloc(sizeof(float) * (size_t)w * h * 3);
                memcpy(o->env_tex, l->env_rgb, sizeof(float) * (size_t)w * h * 3);
                o->env_rows = (double*)malloc(sizeof(double) * (size_t)h * (w + 1));
                o->env_marg = (double*)malloc(sizeof(double) * (size_t)(h + 1));
                o->env_marg[0] = 0.0;
                for (int y = 0; y < h; y++) {
                    double* row = o->env_rows + (size_t)y * (w + 1);
                    row[0] = 0.0;
                    for (int x = 0; x < w; x++) {
                        const int k = y * w + x;
                        const double r = (double)l->env_rgb[3 * k] * o->env_scale, g = (double)l->env_rgb[3 * k + 1] * o->env_scale,
                                     b = (double)l->env_rgb[3 * k + 2] * o->env_scale;
                        const double mx = fmax(fmax(r, g), b);                 /* glm::compMax */
                        row[x + 1] = row[x] + mx * sin(ORC_PI * ((double)k / w + .5) / h);   /* Float(i)/w: the FLAT index */
                    }
                    o->env_marg[y + 1] = o->env_marg[y] + row[w];
                    const double sum = row[w];
                    for (int x = 0; x <= w; x++) row[x] /= sum;
                }
                const double msum = o->env_marg[h];
                for (int y = 0; y <= h; y++) o->env_marg[y] /= msum;
            }
        }
    }
    if (d->camera) {                                                         /* camera_pinhole.cpp:86-96 */
        const LmcuCamera* c = d->camera;
        sc->has_camera = 1;
        sc->cam_pos = v3(c->position[0], c->position[1], c->position[2]);
        sc->cam_w = vnormalize(vsub(sc->cam_pos, v3(c->center[0], c->center[1], c->center[2])));
        sc->cam_u = vnormalize(vcross(v3(c->up[0], c->up[1], c->up[2]), sc->cam_w));
        sc->cam_v = vcross(sc->cam_w, sc->cam_u);
        sc->cam_tf = tan(c->vfov * ORC_PI / 180.0 * .5);
        sc->cam_aspect = c->aspect;
    }
    /* flatten (accel_sahbvh.cpp:132-151) */
    sc->nt = nt;
    sc->tris = (Tri*)malloc(sizeof(Tri) * (size_t)nt);
    int k = 0;
    for (int i = 0; i < d->num_primitives; i++) {
        const int mesh = sc->prims[i].mesh;
        for (int f = 0; f < sc->mesh_faces[mesh]; f++) {
            const double* p = sc->mesh_p[mesh] + (size_t)f * 9;
            const V3 p1 = m4_point(sc->prims[i].M, v3(p[0], p[1], p[2]));
            const V3 p2 = m4_point(sc->prims[i].M, v3(p[3], p[4], p[5]));
            const V3 p3 = m4_point(sc->prims[i].M, v3(p[6], p[7], p[8]));
            Tri* t = &sc->tris[k++];
            t->p1 = p1; t->e1 = vsub(p2, p1); t->e2 = vsub(p3, p1);         /* Tri ctor :41-49 */
            t->b = bound_merge_p(bound_merge_p(bound_merge_p(bound_empty(), p1), p2), p3);
            t->c = vscale(vadd(t->b.min, t->b.max), .5);                    /* Bound::center, math.h:103 */
            t->prim = i; t->face = f;
        }
    }
    build_bvh(sc);
    return sc;
}

void orc_scene_destroy(OrcScene* s) {
    if (!s) return;
    for (int i = 0; i < s->nmeshes; i++) { free(s->mesh_p[i]); free(s->mesh_n[i]); free(s->mesh_t[i]); }
    for (int i = 0; i < s->ntex; i++) free((void*)s->tex[i].data);
    for (int i = 0; i < s->nmats; i++) { free(s->mix_mat[i]); free(s->mix_w[i]); }
    free(s->mesh_t); free(s->tex); free(s->mix_mat); free(s->mix_w);
    for (int i = 0; i < s->nlights; i++) {
        free(s->lights[i].cdf); free(s->lights[i].env_tex); free(s->lights[i].env_rows); free(s->lights[i].env_marg);
    }
    free(s->mesh_p); free(s->mesh_n); free(s->mesh_faces); free(s->prims); free(s->mats); free(s->lights);
    free(s->tris); free(s->nodes); free(s->indices); free(s);
}

/* ------------------------------------------------------------------------------------------------
 * Accel
 * ---------------------------------------------------------------------------------------------- */
typedef struct { V3 o, d; } Ray;

/* Bound::isect_range (math.h:143-158): tmin/tmax by value here (Bound::isect copies them, :127-129) */
static int bound_isect(const Bound* b, Ray r, double tmin, double tmax) {
    const double* mn = &b->min.x; const double* mx = &b->max.x;
    const double* o = &r.o.x; const double* d = &r.d.x;
    for (int i = 0; i < 3; i++) {
        const double vd = 1.0 / d[i];
        double t1 = (mn[i] - o[i]) * vd;
        double t2 = (mx[i] - o[i]) * vd;
        if (vd < 0) { const double t = t1; t1 = t2; t2 = t; }
        tmin = glm_max(t1, tmin);
        tmax = glm_min(t2, tmax);
        if (tmax < tmin) return 0;
    }
    return 1;
}

/* Tri::intersect, Möller-Trumbore (accel_sahbvh.cpp:58-75) */
static int tri_isect(const Tri* tr, Ray r, double tl, double th, double* t_, double* u_, double* v_) {
    const V3 p = vcross(r.d, tr->e2);
    const V3 tv = vsub(r.o, tr->p1);
    const V3 q = vcross(tv, tr->e1);
    const double d = vdot(tr->e1, p);
    const double ad = fabs(d);
    const double s = copysign(1.0, d);
    const double u = vdot(tv, p) * s;
    const double v = vdot(r.d, q) * s;
    if (ad < 1e-8 || u < 0.0 || v < 0.0 || u + v > ad) return 0;
    const double t = vdot(tr->e2, q) / d;
    if (t < tl || th < t) return 0;
    *t_ = t; *u_ = u / ad; *v_ = v / ad;
    return 1;
}

typedef struct { double t, u, v; int tri; } AccelHit;

/* Accel_SAHBVH::intersect (accel_sahbvh.cpp:269-299): unordered DFS, c2 visited before c1, later hit wins ties */
static int accel_intersect(const OrcScene* sc, Ray ray, double tmin, double tmax, AccelHit* out) {
    int stack[99];
    int si = 0, mi = -1;
    double mu = 0, mv = 0;
    stack[0] = 0;
    while (si >= 0) {
        const Node* n = &sc->nodes[stack[si--]];
        if (!bound_isect(&n->b, ray, tmin, tmax)) continue;
        if (!n->leaf) { stack[++si] = n->c1; stack[++si] = n->c2; continue; }
        for (int i = n->s; i < n->e; i++) {
            double t, u, v;
            if (tri_isect(&sc->tris[sc->indices[i]], ray, tmin, tmax, &t, &u, &v)) { tmax = t; mu = u; mv = v; mi = i; }
        }
    }
    if (mi < 0) return 0;
    out->t = tmax; out->u = mu; out->v = mv; out->tri = sc->indices[mi];
    return 1;
}

void orc_intersect(const OrcScene* s, const LmcuRay* rays, int64_t n, LmcuHit* hits) {
    #pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
        const Ray r = { v3(rays[i].o[0], rays[i].o[1], rays[i].o[2]), v3(rays[i].d[0], rays[i].d[1], rays[i].d[2]) };
        AccelHit h;
        if (accel_intersect(s, r, rays[i].tmin, rays[i].tmax, &h)) {
            hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v;
            hits[i].primitive = s->prims[s->tris[h.tri].prim].node_index; hits[i].face = s->tris[h.tri].face;
        } else {
            hits[i].t = rays[i].tmax; hits[i].u = hits[i].v = 0; hits[i].primitive = -1; hits[i].face = -1;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Scene interaction (include/lm/surface.h)
 * ---------------------------------------------------------------------------------------------- */
enum { SP_CAMERA = 1, SP_LIGHT = 2, SP_SURFACE = 4 };

typedef struct {
    int type;           /* SceneInteraction::Type, surface.h:213-221 */
    int prim;           /* primitive ordinal (surface / area light) or -1 (camera, env light) */
    int light;          /* light index for light endpoints */
    int infinite, degenerated;
    V3 p, n, gn, wo, u, v;
    double tu, tv;      /* PointGeometry::t */
} SP;

/* math::orthonormal_basis, Duff et al. (math.h:540-547) */
static void orthonormal_basis(V3 n, V3* u, V3* v) {
    const double s = copysign(1.0, n.z);
    const double a = -1.0 / (s + n.z);
    const double b = n.x * n.y * a;
    *u = v3(1.0 + s * n.x * n.x * a, s * b, -s * n.x);
    *v = v3(b, s + n.y * n.y * a, -n.y);
}

static void make_on_surface(SP* sp, V3 p, V3 n, V3 gn) {       /* surface.h:124-136 */
    sp->degenerated = 0; sp->infinite = 0; sp->p = p; sp->n = n; sp->gn = gn;
    orthonormal_basis(n, &sp->u, &sp->v);
}

/* math::mix_barycentric (math.h:573-575) */
static V3 mix_bary(V3 a, V3 b, V3 c, double u, double v) {
    return vadd(vadd(vscale(a, 1.0 - u - v), vscale(b, u)), vscale(c, v));
}

/* Scene_::intersect (src/scene.cpp:319-345) + Mesh::surface_point (src/mesh/mesh_raw.cpp:87-97) */
static int scene_intersect(const OrcScene* sc, Ray ray, double tmin, double tmax, SP* sp) {
    AccelHit h;
    if (!accel_intersect(sc, ray, tmin, tmax, &h)) {
        if (tmax < ORC_INF) return 0;
        if (sc->env_light < 0) return 0;
        memset(sp, 0, sizeof(*sp));
        sp->type = SP_LIGHT; sp->prim = -1; sp->light = sc->env_light;       /* make_light_endpoint(env, make_infinite(-d)) */
        sp->infinite = 1; sp->degenerated = 0; sp->wo = vneg(ray.d);
        return 1;
    }
    const Tri* t = &sc->tris[h.tri];
    const Prim* pr = &sc->prims[t->prim];
    const double* P = sc->mesh_p[pr->mesh] + (size_t)t->face * 9;
    const double* N = sc->mesh_n[pr->mesh] + (size_t)t->face * 9;
    const V3 a = v3(P[0], P[1], P[2]), b = v3(P[3], P[4], P[5]), c = v3(P[6], P[7], P[8]);
    const V3 pl = mix_bary(a, b, c, h.u, h.v);
    const V3 nl = vnormalize(mix_bary(v3(N[0], N[1], N[2]), v3(N[3], N[4], N[5]), v3(N[6], N[7], N[8]), h.u, h.v));
    const V3 gl = vnormalize(vcross(vsub(b, a), vsub(c, a)));                  /* math::geometry_normal, math.h:560-562 */
    memset(sp, 0, sizeof(*sp));
    sp->type = SP_SURFACE; sp->prim = t->prim; sp->light = pr->light;
    make_on_surface(sp, m4_point(pr->M, pl), vnormalize(m3_mul(pr->NM, nl)), vnormalize(m3_mul(pr->NM, gl)));
    if (sc->mesh_t[pr->mesh]) {                                                   /* mix_barycentric of Mesh::Point::t */
        const double* T = sc->mesh_t[pr->mesh] + (size_t)t->face * 6;
        const double w0 = 1.0 - h.u - h.v;
        sp->tu = (T[0] * w0 + T[2] * h.u) + T[4] * h.v;
        sp->tv = (T[1] * w0 + T[3] * h.u) + T[5] * h.v;
    }
    return 1;
}

/* Scene::visible (include/lm/scene.h:343-364) */
static int scene_visible(const OrcScene* sc, const SP* sp1, const SP* sp2) {
    if (sp1->infinite) { const SP* t = sp1; sp1 = sp2; sp2 = t; }
    const V3 wo = sp2->infinite ? vneg(sp2->wo) : vnormalize(vsub(sp2->p, sp1->p));
    const double tmax = sp2->infinite ? ORC_INF - 1.0 : vlength(vsub(sp1->p, sp2->p)) * (1.0 - ORC_EPS);
    SP tmp;
    const Ray r = { sp1->p, wo };
    return !scene_intersect(sc, r, ORC_EPS, tmax, &tmp);
}

static int sp_opposite(const SP* g, V3 w1, V3 w2) { return vdot(w1, g->n) * vdot(w2, g->n) <= 0; }   /* surface.h:178-180 */

static void basis_twosided(const SP* g, V3 wi, V3* n, V3* u, V3* v) {                               /* surface.h:190-195 */
    const int i = vdot(wi, g->n) > 0;
    *n = i ? g->n : vneg(g->n); *u = g->u; *v = i ? g->v : vneg(g->v);
}

/* surface::geometry_term (surface.h:329-348) */
static double geometry_term(const SP* s1, const SP* s2) {
    if (s1->infinite && s2->infinite) return 0.0;
    V3 d; double L2;
    if (s1->infinite || s2->infinite) { d = s1->infinite ? s1->wo : vneg(s2->wo); L2 = 1.0; }
    else { d = vsub(s2->p, s1->p); L2 = vdot(d, d); d = vdivs(d, sqrt(L2)); }
    const double c1 = (s1->degenerated || s1->infinite) ? 1.0 : fabs(vdot(s1->n, d));
    const double c2 = (s2->degenerated || s2->infinite) ? 1.0 : fabs(vdot(s2->n, vneg(d)));
    return c1 * c2 / L2;
}

/* surface::shading_normal_correction, TransDir::EL (surface.h:424-438); to_local*w has z = dot(n, w) */
static double sn_correction(const SP* g, V3 wi, V3 wo) {
    const double wi_ng = vdot(wi, g->gn), wo_ng = vdot(wo, g->gn);
    const double wi_ns = g->n.x * wi.x + g->n.y * wi.y + g->n.z * wi.z;
    const double wo_ns = g->n.x * wo.x + g->n.y * wo.y + g->n.z * wo.z;
    if (wi_ng * wi_ns <= 0.0 || wo_ng * wo_ns <= 0.0) return 0.0;
    return 1.0;
}

/* ------------------------------------------------------------------------------------------------
 * Camera (src/camera/camera_pinhole.cpp)
 * ---------------------------------------------------------------------------------------------- */
static Ray camera_primary_ray(const OrcScene* sc, double rpx, double rpy) {              /* :128-132 */
    rpx = 2.0 * rpx - 1.0; rpy = 2.0 * rpy - 1.0;
    const V3 d = vnormalize(v3(sc->cam_aspect * sc->cam_tf * rpx, sc->cam_tf * rpy, -1.0));
    Ray r;
    r.o = sc->cam_pos;
    r.d = vadd(vadd(vscale(sc->cam_u, d.x), vscale(sc->cam_v, d.y)), vscale(sc->cam_w, d.z));
    return r;
}

static int camera_raster_position(const OrcScene* sc, V3 wo, double* rx, double* ry) {   /* :207-226 */
    const V3 e = v3(sc->cam_u.x * wo.x + sc->cam_u.y * wo.y + sc->cam_u.z * wo.z,
                    sc->cam_v.x * wo.x + sc->cam_v.y * wo.y + sc->cam_v.z * wo.z,
                    sc->cam_w.x * wo.x + sc->cam_w.y * wo.y + sc->cam_w.z * wo.z);
    if (e.z >= 0) return 0;
    *rx = -e.x / e.z / sc->cam_tf / sc->cam_aspect * .5 + .5;
    *ry = -e.y / e.z / sc->cam_tf * .5 + .5;
    if (*rx < 0.0 || *rx > 1.0 || *ry < 0.0 || *ry > 1.0) return 0;
    return 1;
}

/* Camera_Pinhole::J / pdf_direction / eval (camera_pinhole.cpp:101-109,157-163,233-238): importance of direction wo */
static double camera_pdf_direction(const OrcScene* sc, V3 wo) {
    double rx, ry;
    if (!camera_raster_position(sc, wo, &rx, &ry)) return 0.0;
    const double ez = sc->cam_w.x * wo.x + sc->cam_w.y * wo.y + sc->cam_w.z * wo.z;
    const double cos_theta = -ez;
    const double inv = 1.0 / cos_theta;
    const double A = sc->cam_tf * sc->cam_tf * sc->cam_aspect * 4.0;
    return inv * inv * inv / A;
}

/* ------------------------------------------------------------------------------------------------
 * Sampling primitives (include/lm/math.h:632-702)
 * ---------------------------------------------------------------------------------------------- */
static V3 sample_cosine_weighted(double u0, double u1) {
    const double r = safe_sqrt(u0), t = 2.0 * ORC_PI * u1;
    const double x = r * cos(t), y = r * sin(t);
    return v3(x, y, safe_sqrt(1.0 - x * x - y * y));
}
static V3 sample_uniform_sphere(double u0, double u1) {
    const double z = 1.0 - 2.0 * u0, r = safe_sqrt(1.0 - z * z), t = 2.0 * ORC_PI * u1;
    return v3(r * cos(t), r * sin(t), z);
}
static double balance_heuristic(double p1, double p2) { return (p1 == 0.0 && p2 == 0.0) ? 0.0 : p1 / (p1 + p2); }
static V3 reflection(V3 w, V3 n) { return vsub(vscale(n, 2.0 * vdot(w, n)), w); }        /* math.h:497-499 */

/* Dist::sample (math.h:367-370) */
static int dist_sample(const double* c, int n, double u) {
    int lo = 0, hi = n + 1;                      /* upper_bound over n+1 entries */
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid] <= u) lo = mid + 1; else hi = mid; }
    int i = lo - 1;
    if (i < 0) i = 0;
    if (i > n - 1) i = n - 1;
    return i;
}

/* ------------------------------------------------------------------------------------------------
 * Materials (src/material/*.cpp).  wi, wo point away from the surface; pdfs in projected solid angle.
 * ---------------------------------------------------------------------------------------------- */
static double sq(double v) { return v * v; }

static double glossy_D(const LmcuMaterial* m, V3 wh, V3 u, V3 v, V3 n) {                 /* material_glossy.cpp:43-46 */
    return 1.0 / (ORC_PI * m->ax * m->ay * sq(sq(vdot(wh, u) / m->ax) + sq(vdot(wh, v) / m->ay) + sq(vdot(wh, n))));
}
static double glossy_G1(const LmcuMaterial* m, V3 w, V3 u, V3 v, V3 n) {                 /* material_glossy.cpp:50-57 */
    const double c = vdot(w, n);
    const double s = fmax(ORC_EPS, safe_sqrt(1.0 - c * c));
    const double cp = vdot(w, u) / s, cs = vdot(w, v) / s;
    const double a2 = sq(cp * m->ax) + sq(cs * m->ay);
    return c == 0.0 ? 0.0 : 2.0 / (1.0 + safe_sqrt(1.0 + a2 * sq(s / c)));
}

/* glass: refraction + Schlick Fresnel (material_glass.cpp:90-100, math.h:517-524) */
static double glass_fresnel(const LmcuMaterial* m, const SP* g, V3 wi, V3* wt, double* eta) {
    const int in = vdot(wi, g->n) > 0.0;
    const V3 n = in ? g->n : vneg(g->n);
    *eta = in ? 1.0 / m->Ni : m->Ni;
    const double t = vdot(wi, n);
    const double t2 = 1.0 - *eta * *eta * (1.0 - t * t);
    if (t2 <= 0.0) { *wt = v3(0, 0, 0); return 1.0; }
    *wt = vsub(vscale(vsub(vscale(n, t), wi), *eta), vscale(n, safe_sqrt(t2)));
    const double c = in ? vdot(wi, g->n) : vdot(*wt, g->n);
    const double r = (1.0 - m->Ni) / (1.0 + m->Ni);
    const double r2 = r * r;
    return r2 + (1.0 - r2) * pow(1.0 - c, 5.0);
}

static int mat_is_specular(const LmcuMaterial* m) { return m->type == LMCU_MAT_GLASS || m->type == LMCU_MAT_MIRROR; }

/* Material::eval with eval_delta as given */
static V3 mat_eval(const LmcuMaterial* m, const SP* g, V3 wi, V3 wo, int comp, int eval_delta) {
    const V3 K = v3(m->color[0], m->color[1], m->color[2]);
    switch (m->type) {
    case LMCU_MAT_DIFFUSE:                                                               /* material_diffuse.cpp:86-91 */
        return sp_opposite(g, wi, wo) ? v3(0, 0, 0) : vdivs(K, ORC_PI);
    case LMCU_MAT_GLOSSY: {                                                              /* material_glossy.cpp:101-109 */
        if (sp_opposite(g, wi, wo)) return v3(0, 0, 0);
        const V3 wh = vnormalize(vadd(wi, wo));
        V3 n, u, v; basis_twosided(g, wi, &n, &u, &v);
        const double p5 = pow(1.0 - vdot(wo, wh), 5.0);
        const V3 Fr = vadd(K, vscale(vsub(v3(1, 1, 1), K), p5));
        const double k = glossy_D(m, wh, u, v, n) * (glossy_G1(m, wi, u, v, n) * glossy_G1(m, wo, u, v, n)) / (4.0 * vdot(wi, n) * vdot(wo, n));
        return vscale(vmul(K, Fr), k);
    }
    case LMCU_MAT_MIRROR:                                                                /* material_mirror.cpp:50-52 */
        return eval_delta ? v3(0, 0, 0) : v3(1, 1, 1);
    default: {                                                                           /* material_glass.cpp:248-267 */
        if (eval_delta) return v3(0, 0, 0);
        V3 wt; double eta;
        const double Fr = glass_fresnel(m, g, wi, &wt, &eta);
        if (comp == 0) return v3(Fr, Fr, Fr);
        const double c = (1.0 - Fr) * (eta * eta);
        return v3(c, c, c);
    }
    }
}

static double mat_pdf(const LmcuMaterial* m, const SP* g, V3 wi, V3 wo, int eval_delta) {
    switch (m->type) {
    case LMCU_MAT_DIFFUSE: return sp_opposite(g, wi, wo) ? 0.0 : 1.0 / ORC_PI;          /* material_diffuse.cpp:82-84 */
    case LMCU_MAT_GLOSSY: {                                                              /* material_glossy.cpp:92-99 */
        if (sp_opposite(g, wi, wo)) return 0.0;
        const V3 wh = vnormalize(vadd(wi, wo));
        V3 n, u, v; basis_twosided(g, wi, &n, &u, &v);
        return glossy_D(m, wh, u, v, n) * vdot(wh, n) / (4.0 * vdot(wo, wh) * vdot(wo, n));
    }
    case LMCU_MAT_MIRROR: return eval_delta ? 0.0 : 1.0;                                 /* material_mirror.cpp:47-49 */
    default: return eval_delta ? 0.0 : 1.0;                                              /* material_glass.cpp:222-229 */
    }
}

/* Material::sample_component (only glass has two: material_glass.cpp:124-136) */
static int mat_sample_component(const LmcuMaterial* m, const SP* g, V3 wi, double uc0, double* weight) {
    if (m->type != LMCU_MAT_GLASS) { *weight = 1.0; return 0; }
    V3 wt; double eta;
    const double Fr = glass_fresnel(m, g, wi, &wt, &eta);
    const int comp = uc0 < Fr ? 0 : 1;
    *weight = 1.0 / (comp == 0 ? Fr : 1.0 - Fr);
    return comp;
}

/* Material::sample_direction; returns 0 when the sample is invalid */
static int mat_sample_direction(const LmcuMaterial* m, const SP* g, V3 wi, int comp, double ud0, double ud1, V3* wo, V3* weight) {
    const V3 K = v3(m->color[0], m->color[1], m->color[2]);
    switch (m->type) {
    case LMCU_MAT_DIFFUSE: {                                                             /* material_diffuse.cpp:68-76 */
        V3 n, u, v; basis_twosided(g, wi, &n, &u, &v);
        const V3 d = sample_cosine_weighted(ud0, ud1);
        *wo = vadd(vadd(vscale(u, d.x), vscale(v, d.y)), vscale(n, d.z));
        *weight = K;
        return 1;
    }
    case LMCU_MAT_GLOSSY: {                                                              /* material_glossy.cpp:70-86 */
        V3 n, u, v; basis_twosided(g, wi, &n, &u, &v);
        const double u1 = ud0 * 2.0 * ORC_PI, u2 = ud1;
        const V3 t = vadd(vscale(u, m->ax * cos(u1)), vscale(v, m->ay * sin(u1)));
        const V3 wh = vnormalize(vadd(vscale(t, safe_sqrt(u2 / (1.0 - u2))), n));
        *wo = reflection(wi, wh);
        if (sp_opposite(g, wi, *wo)) return 0;
        const V3 f = mat_eval(m, g, wi, *wo, 0, 0);
        const double p = mat_pdf(m, g, wi, *wo, 0);
        *weight = vdivs(f, p);
        return 1;
    }
    case LMCU_MAT_MIRROR:                                                                /* material_mirror.cpp:41-46 */
        *wo = reflection(wi, g->n); *weight = v3(1, 1, 1);
        return 1;
    default: {                                                                           /* material_glass.cpp:162-190 */
        V3 wt; double eta;
        glass_fresnel(m, g, wi, &wt, &eta);
        *wo = comp == 0 ? reflection(wi, g->n) : wt;
        const V3 f = mat_eval(m, g, wi, *wo, comp, 0);
        *weight = vdivs(f, mat_pdf(m, g, wi, *wo, 0));
        return 1;
    }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Composite materials: mapKd textures (src/texture/texture_bitmap.cpp:68-90), material::mask (material_mask.cpp),
 * material::mixture_wavefrontobj (src/model/model_wavefrontobj.cpp:339-515) and the constant-weight mixtures
 * (src/material/material_mixture.cpp).  The cm_* functions are Material::* of whatever sits on the primitive; plain lobes
 * fall through to the mat_* functions above.  `comp` is the component Material::sample_component returned.
 * ---------------------------------------------------------------------------------------------- */
static void tex_lookup(const LmcuTexture* t, double tu, double tv, V3* rgb, double* alpha) {
    const double u = tu - floor(tu), v = tv - floor(tv);
    int x = (int)(u * t->w), y = (int)(v * t->h);
    x = x < 0 ? 0 : (x > t->w - 1 ? t->w - 1 : x);
    y = y < 0 ? 0 : (y > t->h - 1 ? t->h - 1 : y);
    const float* c = t->data + (size_t)t->c * ((size_t)t->w * y + x);
    *rgb = v3(c[0], c[1], c[2]);
    *alpha = t->c == 4 ? (double)c[3] : 0.0;                                             /* Texture::eval_alpha default: 0 */
}

/* a plain lobe with its mapKd resolved at the surface point (material_diffuse.cpp:70,80,90) */
static LmcuMaterial lobe_at(const OrcScene* sc, const LmcuMaterial* m, const SP* g) {
    LmcuMaterial o = *m;
    if (m->type == LMCU_MAT_DIFFUSE && m->map_kd > 0) {
        V3 c; double a;
        tex_lookup(&sc->tex[m->map_kd - 1], g->tu, g->tv, &c, &a);
        o.color[0] = c.x; o.color[1] = c.y; o.color[2] = c.z;
    }
    return o;
}

static int plain_is_specular(int type) { return type == LMCU_MAT_GLASS || type == LMCU_MAT_MIRROR || type == LMCU_MAT_MASK; }

/* the diffuse and glossy lobes mixture_wavefrontobj owns (:393-403), its alpha (:376-378) and diffuse_selection_prob (:431-439) */
static void obj_parts(const OrcScene* sc, const LmcuMaterial* m, const SP* g, LmcuMaterial* d, LmcuMaterial* gl, double* alpha, double* w) {
    *d = *m; d->type = LMCU_MAT_DIFFUSE;
    *gl = *m; gl->type = LMCU_MAT_GLOSSY; gl->color[0] = m->Ks[0]; gl->color[1] = m->Ks[1]; gl->color[2] = m->Ks[2];
    *alpha = 1.0;
    if (m->map_kd > 0) {
        V3 c; double a;
        tex_lookup(&sc->tex[m->map_kd - 1], g->tu, g->tv, &c, &a);
        d->color[0] = c.x; d->color[1] = c.y; d->color[2] = c.z;
        if (m->alpha_mask && sc->tex[m->map_kd - 1].c == 4) *alpha = a;
    }
    const double wd = fmax(fmax(d->color[0], d->color[1]), d->color[2]), wg = fmax(fmax(gl->color[0], gl->color[1]), gl->color[2]);
    *w = (wd == 0.0 && wg == 0.0) ? 1.0 : wd / (wd + wg);
}

/* Dist over n weights (math.h:310-370): cdf by running sums, normalised by the total */
static void dist_build(const double* w, int n, double* cdf) {
    cdf[0] = 0.0;
    for (int i = 0; i < n; i++) cdf[i + 1] = cdf[i] + w[i];
    const double sum = cdf[n];
    for (int i = 0; i <= n; i++) cdf[i] /= sum;
}

#define ORC_MAX_MIX 16
/* groups of constant_weight_mixture_marginalized (:156-186): group 0 = non-specular entries, then one per specular entry */
typedef struct { int ngroups; int count[ORC_MAX_MIX + 1]; int entry[ORC_MAX_MIX + 1][ORC_MAX_MIX]; double gcdf[ORC_MAX_MIX + 2]; } MixGroups;
static void mix_groups(const OrcScene* sc, const LmcuMaterial* m, MixGroups* G) {
    G->ngroups = 1; G->count[0] = 0;
    for (int k = 0; k < m->num_entries && k < ORC_MAX_MIX; k++) {
        const int spec = plain_is_specular(sc->mats[m->entry_material[k]].type);
        if (spec) { G->count[G->ngroups] = 1; G->entry[G->ngroups][0] = k; G->ngroups++; }
        else G->entry[0][G->count[0]++] = k;
    }
    double sums[ORC_MAX_MIX + 1];
    for (int g = 0; g < G->ngroups; g++) {
        sums[g] = 0.0;
        for (int i = 0; i < G->count[g]; i++) sums[g] += m->entry_weight[G->entry[g][i]];
    }
    dist_build(sums, G->ngroups, G->gcdf);
}
static void group_cdf(const LmcuMaterial* m, const MixGroups* G, int g, double* cdf) {
    double w[ORC_MAX_MIX];
    for (int i = 0; i < G->count[g]; i++) w[i] = m->entry_weight[G->entry[g][i]];
    dist_build(w, G->count[g], cdf);
}

static int cm_is_specular(const OrcScene* sc, const LmcuMaterial* m, int comp) {
    switch (m->type) {
    case LMCU_MAT_MIXTURE_OBJ: return comp != 0;
    case LMCU_MAT_MIXTURE_RR: return plain_is_specular(sc->mats[m->entry_material[comp]].type);
    case LMCU_MAT_MIXTURE_MARGINALIZED: return comp != 0;
    default: return plain_is_specular(m->type);
    }
}

static V3 plain_eval(const LmcuMaterial* m, const SP* g, V3 wi, V3 wo, int comp, int eval_delta) {
    if (m->type == LMCU_MAT_MASK) return eval_delta ? v3(0, 0, 0) : v3(1, 1, 1);          /* material_mask.cpp:49-51 */
    return mat_eval(m, g, wi, wo, comp, eval_delta);
}
static double plain_pdf(const LmcuMaterial* m, const SP* g, V3 wi, V3 wo, int eval_delta) {
    if (m->type == LMCU_MAT_MASK) return eval_delta ? 0.0 : 1.0;                           /* material_mask.cpp:45-47 */
    return mat_pdf(m, g, wi, wo, eval_delta);
}
static int plain_sample_direction(const LmcuMaterial* m, const SP* g, V3 wi, int comp, double ud0, double ud1, V3* wo, V3* weight) {
    if (m->type == LMCU_MAT_MASK) { *wo = vneg(wi); *weight = v3(1, 1, 1); return 1; }    /* material_mask.cpp:38-43 */
    return mat_sample_direction(m, g, wi, comp, ud0, ud1, wo, weight);
}

static V3 cm_eval(const OrcScene* sc, const LmcuMaterial* m, const SP* g, V3 wi, V3 wo, int comp, int eval_delta) {
    switch (m->type) {
    case LMCU_MAT_MIXTURE_OBJ: {                                                         /* model_wavefrontobj.cpp:500-511 */
        LmcuMaterial d, gl; double alpha, w;
        obj_parts(sc, m, g, &d, &gl, &alpha, &w);
        if (comp == 0) {
            V3 sum = v3(0, 0, 0);
            sum = vadd(sum, mat_eval(&d, g, wi, wo, 0, eval_delta));
            sum = vadd(sum, mat_eval(&gl, g, wi, wo, 0, eval_delta));
            return vscale(sum, alpha);
        }
        return eval_delta ? v3(0, 0, 0) : v3(1.0 - alpha, 1.0 - alpha, 1.0 - alpha);
    }
    case LMCU_MAT_MIXTURE_RR: {                                                          /* material_mixture.cpp:92-95 */
        const LmcuMaterial e = lobe_at(sc, &sc->mats[m->entry_material[comp]], g);
        return vscale(plain_eval(&e, g, wi, wo, 0, eval_delta), m->entry_weight[comp]);
    }
    case LMCU_MAT_MIXTURE_MARGINALIZED: {                                                /* :248-256 */
        MixGroups G; mix_groups(sc, m, &G);
        V3 r = v3(0, 0, 0);
        for (int i = 0; i < G.count[comp]; i++) {
            const int k = G.entry[comp][i];
            const LmcuMaterial e = lobe_at(sc, &sc->mats[m->entry_material[k]], g);
            r = vadd(r, vscale(plain_eval(&e, g, wi, wo, 0, eval_delta), m->entry_weight[k]));
        }
        return r;
    }
    default: {
        const LmcuMaterial e = lobe_at(sc, m, g);
        return plain_eval(&e, g, wi, wo, comp, eval_delta);
    }
    }
}

static double cm_pdf(const OrcScene* sc, const LmcuMaterial* m, const SP* g, V3 wi, V3 wo, int comp, int eval_delta) {
    switch (m->type) {
    case LMCU_MAT_MIXTURE_OBJ: {                                                         /* model_wavefrontobj.cpp:487-498 */
        if (comp != 0) return eval_delta ? 0.0 : 1.0;
        LmcuMaterial d, gl; double alpha, w;
        obj_parts(sc, m, g, &d, &gl, &alpha, &w);
        double p = 0.0;
        p += w * mat_pdf(&d, g, wi, wo, eval_delta);
        p += (1.0 - w) * mat_pdf(&gl, g, wi, wo, eval_delta);
        return p;
    }
    case LMCU_MAT_MIXTURE_RR:                                                            /* material_mixture.cpp:88-90 */
        return plain_pdf(&sc->mats[m->entry_material[comp]], g, wi, wo, eval_delta);
    case LMCU_MAT_MIXTURE_MARGINALIZED: {                                                /* :236-246 */
        MixGroups G; mix_groups(sc, m, &G);
        double cdf[ORC_MAX_MIX + 1]; group_cdf(m, &G, comp, cdf);
        double p = 0.0;
        for (int i = 0; i < G.count[comp]; i++) {
            const double psel = cdf[i + 1] - cdf[i];
            p += psel * plain_pdf(&sc->mats[m->entry_material[G.entry[comp][i]]], g, wi, wo, eval_delta);
        }
        return p;
    }
    default: return plain_pdf(m, g, wi, wo, eval_delta);
    }
}

/* Material::sample_component; uc0 = ComponentSampleU::uc[0] */
static int cm_sample_component(const OrcScene* sc, const LmcuMaterial* m, const SP* g, V3 wi, double uc0, double* weight) {
    switch (m->type) {
    case LMCU_MAT_MIXTURE_OBJ: {                                                         /* model_wavefrontobj.cpp:417-424 */
        LmcuMaterial d, gl; double alpha, w;
        obj_parts(sc, m, g, &d, &gl, &alpha, &w);
        if (uc0 < alpha) { *weight = 1.0 / alpha; return 0; }
        *weight = 1.0 / (1.0 - alpha);
        return 1;
    }
    case LMCU_MAT_MIXTURE_RR: {                                                          /* material_mixture.cpp:61-65 */
        double cdf[ORC_MAX_MIX + 1]; dist_build(m->entry_weight, m->num_entries, cdf);
        const int c = dist_sample(cdf, m->num_entries, uc0);
        *weight = 1.0 / (cdf[c + 1] - cdf[c]);
        return c;
    }
    case LMCU_MAT_MIXTURE_MARGINALIZED: {                                                /* :191-195 */
        MixGroups G; mix_groups(sc, m, &G);
        const int c = dist_sample(G.gcdf, G.ngroups, uc0);
        *weight = 1.0 / (G.gcdf[c + 1] - G.gcdf[c]);
        return c;
    }
    case LMCU_MAT_MASK: *weight = 1.0; return 0;
    default: return mat_sample_component(m, g, wi, uc0, weight);
    }
}

/* Material::sample_direction; ud = DirectionSampleU::ud, udc0 = udc[0]; returns 0 when the sample is invalid */
static int cm_sample_direction(const OrcScene* sc, const LmcuMaterial* m, const SP* g, V3 wi, int comp, double ud0, double ud1,
                               double udc0, V3* wo, V3* weight) {
    switch (m->type) {
    case LMCU_MAT_MIXTURE_OBJ: {                                                         /* model_wavefrontobj.cpp:441-480 */
        if (comp != 0) { *wo = vneg(wi); *weight = v3(1, 1, 1); return 1; }            /* returns Vec3(1), not f/p (:470-479) */
        LmcuMaterial d, gl; double alpha, w;
        obj_parts(sc, m, g, &d, &gl, &alpha, &w);
        V3 ws;
        if (!mat_sample_direction(udc0 < w ? &d : &gl, g, wi, 0, ud0, ud1, wo, &ws)) return 0;
        *weight = vdivs(cm_eval(sc, m, g, wi, *wo, comp, 0), cm_pdf(sc, m, g, wi, *wo, comp, 0));
        return 1;
    }
    case LMCU_MAT_MIXTURE_RR: {                                                          /* material_mixture.cpp:67-78 */
        const LmcuMaterial e = lobe_at(sc, &sc->mats[m->entry_material[comp]], g);
        V3 ws;
        if (!plain_sample_direction(&e, g, wi, 0, ud0, ud1, wo, &ws)) return 0;
        *weight = vscale(ws, m->entry_weight[comp]);
        return 1;
    }
    case LMCU_MAT_MIXTURE_MARGINALIZED: {                                                /* :201-224 */
        MixGroups G; mix_groups(sc, m, &G);
        double cdf[ORC_MAX_MIX + 1]; group_cdf(m, &G, comp, cdf);
        const int i = dist_sample(cdf, G.count[comp], udc0);
        const LmcuMaterial e = lobe_at(sc, &sc->mats[m->entry_material[G.entry[comp][i]]], g);
        V3 ws;
        if (!plain_sample_direction(&e, g, wi, 0, ud0, ud1, wo, &ws)) return 0;
        *weight = vdivs(cm_eval(sc, m, g, wi, *wo, comp, 0), cm_pdf(sc, m, g, wi, *wo, comp, 0));
        return 1;
    }
    default: {
        const LmcuMaterial e = lobe_at(sc, m, g);
        return plain_sample_direction(&e, g, wi, comp, ud0, ud1, wo, weight);
    }
    }
}

/* Material::reflectance (raycast shading) */
static V3 cm_reflectance(const OrcScene* sc, const LmcuMaterial* m, const SP* g) {
    switch (m->type) {
    case LMCU_MAT_MIXTURE_OBJ: {                                                         /* model_wavefrontobj.cpp:482-485 */
        LmcuMaterial d, gl; double alpha, w;
        obj_parts(sc, m, g, &d, &gl, &alpha, &w);
        return v3(d.color[0], d.color[1], d.color[2]);
    }
    case LMCU_MAT_MIXTURE_RR:
    case LMCU_MAT_MIXTURE_MARGINALIZED: {                                                /* material_mixture.cpp:80-86,226-234 */
        V3 sum = v3(0, 0, 0);
        if (m->type == LMCU_MAT_MIXTURE_RR) {
            for (int k = 0; k < m->num_entries; k++) sum = vadd(sum, vscale(cm_reflectance(sc, &sc->mats[m->entry_material[k]], g), m->entry_weight[k]));
        } else {
            MixGroups G; mix_groups(sc, m, &G);
            for (int gi = 0; gi < G.ngroups; gi++) for (int i = 0; i < G.count[gi]; i++) {
                const int k = G.entry[gi][i];
                sum = vadd(sum, vscale(cm_reflectance(sc, &sc->mats[m->entry_material[k]], g), m->entry_weight[k]));
            }
        }
        return sum;
    }
    case LMCU_MAT_DIFFUSE: case LMCU_MAT_GLOSSY: {
        const LmcuMaterial e = lobe_at(sc, m, g);
        return v3(e.color[0], e.color[1], e.color[2]);
    }
    default: return v3(0, 0, 0);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Lights
 * ---------------------------------------------------------------------------------------------- */
/* light::env lookup coordinates of a direction d = -geomL.wo (light_env.cpp:166-176,191-199) */
static void env_uv(const Light* l, V3 d, double* u, double* v) {
    double at = atan2(d.x, d.z);
    if (at < 0.0) at += 2.0 * ORC_PI;
    const double t = (at - l->env_rot) * .5 / ORC_PI;
    *u = t - floor(t);
    *v = acos(d.y) / ORC_PI;
}

/* Dist::pmf / Dist2::pdf (include/lm/math.h:352-354,416-419) */
static double dist_pmf(const double* c, int n, int i) { return (i < 0 || i + 1 >= n + 1) ? 0.0 : c[i + 1] - c[i]; }
static double env_dist_pdf(const Light* l, double u, double v) {
    const int w = l->env_w, h = l->env_h;
    int y = (int)(v * h);
    if (y > h - 1) y = h - 1;
    return dist_pmf(l->env_marg, h, y) * dist_pmf(l->env_rows + (size_t)y * (w + 1), w, (int)(u * w)) * w * h;
}

/* Light::eval: area (light_area.cpp:185-187), envconst (light_envconst.cpp:139-141), env (light_env.cpp:190-199 +
 * texture_bitmap.cpp:68-75) */
static V3 light_eval(const Light* l, const SP* geomL, V3 wo) {
    if (l->type == LMCU_LIGHT_AREA) return vdot(wo, geomL->n) <= 0.0 ? v3(0, 0, 0) : l->Ke;
    if (l->type == LMCU_LIGHT_ENV) {
        double u, v;
        env_uv(l, vneg(geomL->wo), &u, &v);
        u -= floor(u); v -= floor(v);
        int x = (int)(u * l->env_w), y = (int)(v * l->env_h);
        x = x < 0 ? 0 : (x > l->env_w - 1 ? l->env_w - 1 : x);
        y = y < 0 ? 0 : (y > l->env_h - 1 ? l->env_h - 1 : y);
        const float* t = l->env_tex + ((size_t)y * l->env_w + x) * 3;
        return vscale(v3(t[0], t[1], t[2]), l->env_scale);
    }
    return l->Ke;
}

/* Light::pdf_direct (light_area.cpp:172-178, light_envconst.cpp:116-119, light_env.cpp:164-181 + surface.h:377-390) */
static double light_pdf_direct(const OrcScene* sc, const Light* l, const SP* geom, const SP* geomL, V3 wo) {
    if (l->type == LMCU_LIGHT_AREA) {
        if (vdot(wo, geomL->n) <= 0.0) return 0.0;
        const double G = geometry_term(geom, geomL);
        return G == 0.0 ? 0.0 : (l->invA / sc->prims[l->prim].J) / G;
    }
    if (l->type == LMCU_LIGHT_POINT) {                                                   /* light_point.cpp:111-114 */
        const double G = geometry_term(geom, geomL);
        return G == 0.0 ? 0.0 : 1.0 / G;
    }
    const V3 d = vneg(geomL->wo);
    double pdf_sa = 1.0 / (4.0 * ORC_PI);
    if (l->type == LMCU_LIGHT_DIRECTIONAL) pdf_sa = 1.0;                                 /* light_directional.cpp:105-108, eval_delta = false */
    if (l->type == LMCU_LIGHT_ENV) {
        double u, v;
        env_uv(l, d, &u, &v);
        const double st = safe_sqrt(1.0 - d.y * d.y);
        if (st == 0.0) return 0.0;
        pdf_sa = env_dist_pdf(l, u, v) / (2.0 * ORC_PI * ORC_PI * st);
    }
    if (geom->degenerated) return pdf_sa;
    const double J = fabs(vdot(geom->n, d));
    return J == 0.0 ? 0.0 : pdf_sa / J;
}

/* path::sample_direct, TransDir::LE (include/lm/path.h:493-530); u = {up0,up1, upc0,upc1, ud0,ud1, udc0,udc1} */
static int sample_direct(const OrcScene* sc, const double* u, const SP* sp, SP* spL, V3* wo, V3* weight) {
    const int n = sc->nlights;
    int li = (int)(u[2] * n);                                                            /* src/scene.cpp:374-387 */
    li = li < 0 ? 0 : (li > n - 1 ? n - 1 : li);
    const double p_sel = 1.0 / n;
    const Light* l = &sc->lights[li];
    memset(spL, 0, sizeof(*spL));
    spL->type = SP_LIGHT; spL->light = li; spL->prim = l->prim;
    if (l->type == LMCU_LIGHT_AREA) {                                                    /* light_area.cpp:47-59,148-169 */
        const Prim* pr = &sc->prims[l->prim];
        const int i = dist_sample(l->cdf, l->ntri, u[3]);
        const double s = safe_sqrt(u[0]);
        const double* P = sc->mesh_p[pr->mesh] + (size_t)i * 9;
        const V3 a = v3(P[0], P[1], P[2]), b = v3(P[3], P[4], P[5]), c = v3(P[6], P[7], P[8]);
        const V3 p = mix_bary(a, b, c, 1.0 - s, u[1] * s);
        const V3 gn = vnormalize(vcross(vsub(b, a), vsub(c, a)));
        const V3 gt = vnormalize(m3_mul(pr->NM, gn));
        make_on_surface(spL, m4_point(pr->M, p), gt, gt);
        *wo = vnormalize(vsub(sp->p, spL->p));
    } else if (l->type == LMCU_LIGHT_POINT) {                                            /* light_point.cpp:96-109 */
        *wo = vnormalize(vsub(sp->p, l->pos));
        spL->p = l->pos; spL->degenerated = 1;
    } else if (l->type == LMCU_LIGHT_DIRECTIONAL) {                                      /* light_directional.cpp:93-103 */
        *wo = l->pos;
        spL->infinite = 1; spL->wo = *wo;
    } else if (l->type == LMCU_LIGHT_ENV) {                                              /* light_env.cpp:139-162 */
        const int y = dist_sample(l->env_marg, l->env_h, u[4]);                          /* dist_.sample({ud, up}) */
        const int x = dist_sample(l->env_rows + (size_t)y * (l->env_w + 1), l->env_w, u[5]);
        const double su = (x + u[0]) / l->env_w, sv = (y + u[1]) / l->env_h;
        const double t = ORC_PI * sv, st = sin(t), p = 2.0 * ORC_PI * su + l->env_rot;
        *wo = vneg(v3(st * sin(p), cos(t), st * cos(p)));
        spL->infinite = 1; spL->wo = *wo;
    } else {                                                                             /* light_envconst.cpp:100-115 */
        *wo = sample_uniform_sphere(u[4], u[5]);
        spL->infinite = 1; spL->wo = *wo;
    }
    const double pL = light_pdf_direct(sc, l, sp, spL, *wo);
    if (pL == 0.0) return 0;
    *weight = vdivs(vdivs(light_eval(l, spL, *wo), pL), p_sel);
    return 1;
}

/* path::pdf_direct for a light endpoint (include/lm/path.h:563-583) */
static double pdf_direct(const OrcScene* sc, const SP* sp, const SP* spL, V3 wo) {
    return (1.0 / sc->nlights) * light_pdf_direct(sc, &sc->lights[spL->light], sp, spL, wo);
}

/* ------------------------------------------------------------------------------------------------
 * Film (include/lm/film.h:166-171, src/film/film_bitmap.cpp:236-248)
 * ---------------------------------------------------------------------------------------------- */
static void film_splat(double* film, int w, int h, double rx, double ry, V3 c) {
    int x = (int)(rx * (double)w), y = (int)(ry * (double)h);
    x = x < 0 ? 0 : (x > w - 1 ? w - 1 : x);
    y = y < 0 ? 0 : (y > h - 1 ? h - 1 : y);
    double* px = film + ((size_t)y * w + x) * 3;
    #pragma omp atomic
    px[0] += c.x;
    #pragma omp atomic
    px[1] += c.y;
    #pragma omp atomic
    px[2] += c.z;
}

/* ------------------------------------------------------------------------------------------------
 * Renderers
 * ---------------------------------------------------------------------------------------------- */
/* Renderer_Raycast::render (src/renderer/renderer_raycast.cpp:62-101) */
void orc_render_raycast(const OrcScene* sc, const LmcuRenderParams* p, double* film) {
    const int w = p->width, h = p->height;
    #pragma omp parallel for schedule(dynamic, 16)
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            const Ray ray = camera_primary_ray(sc, (x + .5) / w, (y + .5) / h);
            SP sp;
            V3 C;
            if (!scene_intersect(sc, ray, ORC_EPS, ORC_INF, &sp)) {
                C = v3(p->bg_color[0], p->bg_color[1], p->bg_color[2]);
            } else if (p->visualize_normal) {
                C = v3(fabs(sp.n.x), fabs(sp.n.y), fabs(sp.n.z));
            } else if (sp.infinite) {
                C = sc->lights[sp.light].Ke;
            } else {
                const LmcuMaterial* m = &sc->mats[sc->prims[sp.prim].material];
                if (p->has_color) C = v3(p->color[0], p->color[1], p->color[2]);
                else C = cm_reflectance(sc, m, &sp);
                if (!p->use_constant_color) C = vscale(C, .2 + .8 * fabs(vdot(sp.n, vneg(ray.d))));
            }
            double* px = film + ((size_t)y * w + x) * 3;
            px[0] = C.x; px[1] = C.y; px[2] = C.z;
        }
    }
}

static int resolve_threads(int num_threads) {
#ifdef _OPENMP
    return num_threads > 0 ? num_threads : omp_get_max_threads();
#else
    (void)num_threads;
    return 1;
#endif
}

/* Renderer_AO::render (functest/renderer_ao.cpp:24-45); Rng(42 + threadId) is thread_local */
void orc_render_ao(const OrcScene* sc, const LmcuRenderParams* p, double* film, int num_threads) {
    const int w = p->width, h = p->height, nth = resolve_threads(num_threads);
    const double dist = p->ao_distance > 0 ? p->ao_distance : .2;
    memset(film, 0, sizeof(double) * 3 * (size_t)w * h);
    #pragma omp parallel num_threads(nth)
    {
        Rng rng;
#ifdef _OPENMP
        rng_seed(&rng, (uint32_t)(42 + omp_get_thread_num()));
#else
        rng_seed(&rng, 42u);
#endif
        #pragma omp for schedule(dynamic, 1)
        for (long long index = 0; index < (long long)w * h; index++) {
            const int x = (int)(index % w), y = (int)(index / w);
            const Ray ray = camera_primary_ray(sc, (x + .5) / w, (y + .5) / h);
            SP hit;
            if (!scene_intersect(sc, ray, ORC_EPS, ORC_INF, &hit)) continue;
            double V = 0.0;
            for (long long i = 0; i < p->spp; i++) {
                V3 n, u, v;
                basis_twosided(&hit, vneg(ray.d), &n, &u, &v);
                const double u0 = rng_u(&rng), u1 = rng_u(&rng);
                const V3 d = sample_cosine_weighted(u0, u1);
                const Ray r2 = { hit.p, vadd(vadd(vscale(u, d.x), vscale(v, d.y)), vscale(n, d.z)) };
                SP tmp;
                V += scene_intersect(sc, r2, ORC_EPS, dist, &tmp) ? 0.0 : 1.0;
            }
            V /= (double)p->spp;
            double* px = film + ((size_t)y * w + x) * 3;
            px[0] = px[1] = px[2] = V;
        }
    }
}

/* Renderer_PT::render (src/renderer/renderer_pt.cpp:80-315) with scheduler::spp::sample (src/scheduler.cpp:37-49:
 * index -> pixel = index / spp, sample = index % spp; pixel mode) or scheduler::spi::sample (:131-141: pixel_index 0,
 * the window is the whole image; image mode).  The time schedulers are not restated: they are not deterministic. */
void orc_render_pt(const OrcScene* sc, const LmcuRenderParams* p, double* film, int num_threads) {
    const int w = p->width, h = p->height, nth = resolve_threads(num_threads);
    const int image = p->primary_ray_sampling_mode == LMCU_PRIMARY_IMAGE;
    const long long spp = p->spp, total = image ? (long long)p->num_samples : (long long)w * h * spp;
    const int mode = p->sampling_mode;
    memset(film, 0, sizeof(double) * 3 * (size_t)w * h);
    #pragma omp parallel num_threads(nth)
    {
        Rng rng;
#ifdef _OPENMP
        rng_seed(&rng, p->seed + (uint32_t)omp_get_thread_num());                        /* :91 */
#else
        rng_seed(&rng, p->seed);
#endif
        #pragma omp for schedule(dynamic, 1)
        for (long long index = 0; index < total; index++) {
            const long long pixel_index = image ? 0 : index / spp;
            const int px = (int)(pixel_index % w), py = (int)(pixel_index / w);
            const double dx = 1.0 / w, dy = 1.0 / h;
            double win[4] = { dx * px, dy * py, dx, dy };                                /* :96-107 */
            if (image) { win[0] = 0.0; win[1] = 0.0; win[2] = 1.0; win[3] = 1.0; }
            double u4[8];
            for (int i = 0; i < 4; i++) u4[i] = rng_u(&rng);                             /* sample_position: PositionSampleU */
            for (int i = 0; i < 2; i++) u4[i] = rng_u(&rng);                             /* sample_component: ComponentSampleU */
            SP sp;
            memset(&sp, 0, sizeof(sp));
            sp.type = SP_CAMERA; sp.prim = -1; sp.light = -1; sp.degenerated = 1; sp.p = sc->cam_pos;   /* :112 */
            int comp = 0;
            V3 throughput = v3(1, 1, 1);
            V3 wi = v3(0, 0, 0);
            double rpx = 0, rpy = 0;
            for (int num_verts = 1; num_verts < p->max_verts; num_verts++) {
                const LmcuMaterial* mat = sp.type == SP_SURFACE ? &sc->mats[sc->prims[sp.prim].material] : NULL;
                int samplable_by_nee = 0;                                                /* :127-140 */
                if (mode != LMCU_SAMPLING_NAIVE) {
                    const int is_spec = mat ? cm_is_specular(sc, mat, comp) : 0;
                    samplable_by_nee = image ? !is_spec : (num_verts > 1 && !is_spec);  /* image mode: NEE from the camera too */
                }
                if (samplable_by_nee) {                                                  /* :142-192 */
                    double u[8];
                    for (int i = 0; i < 8; i++) u[i] = rng_u(&rng);                      /* RaySampleU: up, upc, ud, udc */
                    SP spL; V3 woL, wL;
                    double rx = rpx, ry = rpy;                                           /* :151-157 */
                    if (sample_direct(sc, u, &sp, &spL, &woL, &wL) && scene_visible(sc, &sp, &spL) &&
                        (num_verts > 1 || camera_raster_position(sc, vneg(woL), &rx, &ry))) {
                        const V3 wo = vneg(woL);
                        V3 fs;                                                           /* path::eval_contrb_direction, path.h:761-775 */
                        if (mat) fs = vscale(cm_eval(sc, mat, &sp, wi, wo, comp, 1), sn_correction(&sp, wi, wo));
                        else { const double We = camera_pdf_direction(sc, wo); fs = v3(We, We, We); }
                        if (!vis_zero(fs)) {
                            double mis_w = 1.0;
                            /* renderer_pt.cpp:176-180: a specular (directional) or degenerated (point) light endpoint
                             * cannot be reached by BSDF sampling -> NEE only */
                            const int ltype = sc->lights[spL.light].type;
                            const int samplable_by_bsdf = ltype != LMCU_LIGHT_DIRECTIONAL && !spL.degenerated;
                            if (mode == LMCU_SAMPLING_MIS && samplable_by_bsdf) {
                                const double p_light = pdf_direct(sc, &sp, &spL, woL);
                                const double p_bsdf = mat ? cm_pdf(sc, mat, &sp, wi, wo, comp, 1) : camera_pdf_direction(sc, wo);
                                mis_w = balance_heuristic(p_light, p_bsdf);
                            }
                            film_splat(film, w, h, rx, ry, vscale(vmul(vmul(throughput, fs), wL), mis_w));
                        }
                    }
                }
                V3 wo, weight;                                                           /* :197-209 */
                if (num_verts == 1) {
                    /* Vec2(x+w*rng.u(), y+h*rng.u()): g++ evaluates the constructor arguments right to left */
                    const double uy = rng_u(&rng), ux = rng_u(&rng);
                    const double udx = win[0] + win[2] * ux, udy = win[1] + win[3] * uy;
                    rng_u(&rng); rng_u(&rng);                                            /* rng.next<Vec2>() (udc, unused) */
                    wo = camera_primary_ray(sc, udx, udy).d;
                    weight = v3(1, 1, 1);
                } else {
                    double u[4];
                    for (int i = 0; i < 4; i++) u[i] = rng_u(&rng);                      /* DirectionSampleU: ud, udc */
                    if (!cm_sample_direction(sc, mat, &sp, wi, comp, u[0], u[1], u[2], &wo, &weight)) break;
                    weight = vscale(weight, sn_correction(&sp, wi, wo));                 /* include/lm/path.h:415-419 */
                }
                if (num_verts == 1) camera_raster_position(sc, wo, &rpx, &rpy);         /* :214-216 */
                SP hit;                                                                  /* :221-224 */
                const Ray ray = { sp.p, wo };
                if (!scene_intersect(sc, ray, ORC_EPS, ORC_INF, &hit)) break;
                throughput = vmul(throughput, weight);                                   /* :229 */
                const int direct_ok = mode == LMCU_SAMPLING_NEE ? !samplable_by_nee : 1; /* :236-246 */
                if (direct_ok && hit.light >= 0) {                                       /* :248-270 */
                    const V3 woL = vneg(wo);
                    const V3 fs = light_eval(&sc->lights[hit.light], &hit, woL);
                    double mis_w = 1.0;
                    if (mode != LMCU_SAMPLING_NAIVE && samplable_by_nee) {
                        const double pdf_bsdf = mat ? cm_pdf(sc, mat, &sp, wi, wo, comp, 1) : camera_pdf_direction(sc, wo);
                        const double pdf_light = pdf_direct(sc, &sp, &hit, woL);
                        mis_w = balance_heuristic(pdf_bsdf, pdf_light);
                    }
                    film_splat(film, w, h, rpx, rpy, vscale(vmul(throughput, fs), mis_w));
                }
                if (hit.infinite) break;                                                 /* :276-278 */
                if (num_verts > 5) {                                                     /* :281-287 */
                    const double q = glm_max(.2, 1.0 - vmaxc(throughput));
                    if (rng_u(&rng) < q) break;
                    throughput = vdivs(throughput, 1.0 - q);
                }
                const double uc0 = rng_u(&rng);                                          /* :292-293, ComponentSampleU */
                rng_u(&rng);
                double cw;
                const int ncomp = cm_sample_component(sc, &sc->mats[sc->prims[hit.prim].material], &hit, vneg(wo), uc0, &cw);
                throughput = vscale(throughput, cw);
                wi = vneg(wo); sp = hit; comp = ncomp;                                   /* :298-300 */
            }
        }
    }
    if (p->rescale) {                                                                    /* :307-312 */
        const double s = image ? (double)(w * h) / (double)total : 1.0 / (double)spp;
        for (size_t i = 0; i < (size_t)w * h * 3; i++) film[i] *= s;
    }
}
