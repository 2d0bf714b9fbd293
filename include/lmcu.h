/*
 * lmcu.h — C ABI of liblmcu.so, the B200 (sm_100a) ray-tracing hot path for Lightmetrica v3.
 *
 * This is the drop-in boundary: plain pointers and sizes, POD structs, int status codes, no
 * exceptions and no C++/torch types across the line.  The Lightmetrica plugin (plugin/lm_cuda.cpp:
 * accel::cuda_bvh, renderer::{raycast,ao,pt}_cuda) and the Python host (lmcu/) both bind exactly
 * these entry points.  Every entry cites the reference interface it replaces, relative to
 * /root/reference.
 *
 * Conventions shared with the reference:
 *   - all host-side scalars are FP64 (lm::Float is double, include/lm/common.h:279-287);
 *   - Inf = 1e10 is a finite sentinel and the default tmax, Eps = 1e-3 the default tmin
 *     (include/lm/math.h:54-55);
 *   - matrices are column-major 4x4 (glm), vectors xyz;
 *   - film buffers are row-major, row 0 = bottom row (raster y grows upward,
 *     src/film/film_bitmap.cpp:212-218,256-260), 3 channels.
 */
#ifndef LMCU_H
#define LMCU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMCU_INF 1e10
#define LMCU_EPS 1e-3

/* status codes (0 = ok); lmcu_last_error() has the message for the calling thread */
enum {
    LMCU_OK = 0,
    LMCU_ERR_INVALID_ARGUMENT = 1,   /* lm::Error::InvalidArgument */
    LMCU_ERR_UNSUPPORTED = 2,        /* lm::Error::Unsupported */
    LMCU_ERR_CUDA = 3,               /* a CUDA runtime call or kernel failed */
    LMCU_ERR_NO_DEVICE = 4           /* no CUDA device: there is NO CPU fallback */
};

/* ---------------------------------------------------------------------------------------------
 * Ray / hit records.  Replace lm::Ray (include/lm/math.h:67-79) + the (tmin,tmax) arguments of
 * lm::Accel::intersect (include/lm/accel.h:70) and lm::Accel::Hit (include/lm/accel.h:49-55).
 * The 208-byte Hit::global_transform is NOT copied per ray: `primitive` identifies the instance
 * and the caller looks the transform up in its own table.
 * ------------------------------------------------------------------------------------------- */
typedef struct LmcuRay {
    double o[3];
    double d[3];
    double tmin;
    double tmax;
} LmcuRay;                                   /* 64 bytes */

typedef struct LmcuHit {
    double t;                                /* hit distance; tmin <= t <= tmax (inclusive) */
    double u, v;                             /* barycentrics: p = p1(1-u-v) + p2 u + p3 v */
    int32_t primitive;                       /* LmcuPrimitive::node_index of the instance; -1 = miss */
    int32_t face;                            /* mesh-local face index */
} LmcuHit;                                   /* 32 bytes */

/* ---------------------------------------------------------------------------------------------
 * Scene description = what a Lightmetrica Accel/Renderer pulls out of lm::Scene:
 *   meshes      <- Mesh::foreach_triangle (include/lm/mesh.h:64-100), un-indexed triangle soup
 *   primitives  <- Scene::traverse_primitive_nodes (include/lm/scene.h:163-177): one entry per
 *                  primitive node that has a mesh, with its GLOBAL transform (instance groups are
 *                  flattened, src/scene.cpp:216-230), in traversal order
 *   materials   <- material::{diffuse,glossy,glass,mirror,mask} parameters, material::mixture_wavefrontobj and the
 *                  constant-weight mixtures as lists of entries
 *   textures    <- texture::bitmap texels (src/texture/texture_bitmap.cpp:41-96)
 *   lights      <- Scene light table order (src/scene.cpp:281-286): area, envconst, env, point, directional
 *   camera      <- camera::pinhole (src/camera/camera_pinhole.cpp:66-97)
 * ------------------------------------------------------------------------------------------- */
typedef struct LmcuMesh {
    int32_t num_faces;
    const double* p;                         /* [num_faces][3][3] local-space positions */
    const double* n;                         /* [num_faces][3][3] vertex normals, or NULL (use geometric normal) */
    const double* t;                         /* [num_faces][3][2] texture coordinates (Mesh::Point::t), or NULL (0,0) */
} LmcuMesh;

typedef struct LmcuPrimitive {
    int32_t node_index;                      /* value reported back in LmcuHit::primitive (SceneNode::index) */
    int32_t mesh;                            /* index into meshes[] */
    int32_t material;                        /* index into materials[] */
    int32_t light;                           /* index into lights[] if this primitive is an area light, else -1 */
    double M[16];                            /* global transform, column-major */
} LmcuPrimitive;

enum { LMCU_MAT_DIFFUSE = 0, LMCU_MAT_GLOSSY = 1, LMCU_MAT_GLASS = 2, LMCU_MAT_MIRROR = 3,
       LMCU_MAT_MASK = 4,             /* material::mask (src/material/material_mask.cpp): pass-through delta lobe */
       LMCU_MAT_MIXTURE_OBJ = 5,      /* material::mixture_wavefrontobj (src/model/model_wavefrontobj.cpp:339-515):
                                         diffuse(Kd | mapKd) + glossy(Ks, ax, ay) marginalised, alpha mask from mapKd */
       LMCU_MAT_MIXTURE_RR = 6,       /* material::constant_weight_mixture_rr (src/material/material_mixture.cpp:25-107) */
       LMCU_MAT_MIXTURE_MARGINALIZED = 7 };  /* material::constant_weight_mixture_marginalized (:124-258) */

/* texture::bitmap (src/texture/texture_bitmap.cpp:41-96): nearest-texel lookup at (u - floor u, v - floor v) */
typedef struct LmcuTexture {
    int32_t w, h;
    int32_t c;                               /* channels per texel: 3 (rgb) or 4 (rgba: has_alpha()) */
    int32_t pad_;
    const float* data;                       /* [h][w][c] in the row order texture::bitmap holds (Texture::buffer()) */
    const char* path;                        /* image file (only the reference harness reads files) */
} LmcuTexture;

typedef struct LmcuMaterial {
    int32_t type;
    int32_t map_kd;                          /* diffuse, mixture_obj: 1 + index into textures[] of mapKd, 0 = none (use color) */
    double color[3];                         /* diffuse, mixture_obj: Kd; glossy: Ks */
    double ax, ay;                           /* glossy, mixture_obj: anisotropic GGX roughness */
    double Ni;                               /* glass: index of refraction */
    double Ks[3];                            /* mixture_obj: specular reflectance of the glossy lobe */
    int32_t alpha_mask;                      /* mixture_obj: 1 = mapKd's alpha channel masks the surface (!no_alpha_mask and
                                                the texture has_alpha(), model_wavefrontobj.cpp:405-412) */
    int32_t num_entries;                     /* constant-weight mixtures: number of (material, weight) entries */
    const int32_t* entry_material;           /* ... index into materials[] of each entry (diffuse | glossy | mirror | mask) */
    const double* entry_weight;              /* ... weight of each entry */
} LmcuMaterial;

enum { LMCU_LIGHT_AREA = 0, LMCU_LIGHT_ENVCONST = 1, LMCU_LIGHT_ENV = 2,
       LMCU_LIGHT_POINT = 3,          /* light::point (src/light/light_point.cpp): Ke = Le, position */
       LMCU_LIGHT_DIRECTIONAL = 4 };  /* light::directional (src/light/light_directional.cpp): Ke = Le, position = direction */

typedef struct LmcuLight {
    int32_t type;
    int32_t primitive;                       /* area: index into primitives[] (emitter geometry); env: -1 */
    int32_t node_index;                      /* scene node index of the light primitive */
    int32_t env_w, env_h;                    /* env: image size */
    int32_t pad_;
    double Ke[3];                            /* area: Ke; envconst: Le */
    double env_rot;                          /* env: rotation about +y in DEGREES (light_env.cpp:63) */
    double env_scale;                        /* env: luminance scale */
    const float* env_rgb;                    /* env: [env_h][env_w][3], row order as texture::bitmap holds it */
    const char* env_path;                    /* env: image path (only the reference harness reads files) */
    double position[3];                      /* point: position; directional: direction the light travels in */
} LmcuLight;

typedef struct LmcuCamera {
    double position[3];
    double center[3];
    double up[3];
    double vfov;                             /* degrees */
    double aspect;                           /* width / height */
} LmcuCamera;

typedef struct LmcuSceneDesc {
    int32_t num_meshes;
    int32_t num_primitives;
    int32_t num_materials;
    int32_t num_lights;
    int32_t num_textures;
    int32_t pad_;
    const LmcuMesh* meshes;
    const LmcuPrimitive* primitives;
    const LmcuMaterial* materials;
    const LmcuLight* lights;
    const LmcuCamera* camera;                /* may be NULL for accel-only use */
    const LmcuTexture* textures;             /* may be NULL when num_textures == 0 */
} LmcuSceneDesc;

/* ---------------------------------------------------------------------------------------------
 * Renderer parameters = the JSON keys of renderer::raycast (src/renderer/renderer_raycast.cpp:48-60),
 * renderer::ao (functest/renderer_ao.cpp:18-22) and renderer::pt (src/renderer/renderer_pt.cpp:51-77).
 * ------------------------------------------------------------------------------------------- */
enum { LMCU_SAMPLING_NAIVE = 0, LMCU_SAMPLING_NEE = 1, LMCU_SAMPLING_MIS = 2 };
/* primary_ray_sampling_mode (renderer_pt.cpp:65-75,96-107): the window the primary raster position is drawn from */
enum { LMCU_PRIMARY_PIXEL = 0,        /* one pixel per sample index, scheduler::spp::* */
       LMCU_PRIMARY_IMAGE = 1 };      /* the whole image, scheduler::spi::* */
/* scheduler (src/scheduler.cpp): "sample" = fixed spp / num_samples (:37-49,131-141), "time" = whole passes until
 * render_time seconds have elapsed (:75-103,162-187) */
enum { LMCU_SCHED_SAMPLE = 0, LMCU_SCHED_TIME = 1 };
/* The strategies (s light vertices, t eye vertices) lmcu_render_bdpt evaluates per sample.  ALL is renderer::bdpt: every
 * connection, power-heuristic weights.  The other three are the reference's single-strategy renderers on the same path
 * structure (src/renderer/renderer_bdpt.cpp), each connection UNWEIGHTED (Path::eval_sampling_weight_bidir only): */
enum { LMCU_BDPT_ALL = 0,
       LMCU_BDPT_PT_NAIVE = 1,        /* renderer::pt_naive_path (:63-124): (0, t), the eye subpath hits a light */
       LMCU_BDPT_PT_NEE = 2,          /* renderer::pt_nee_path (:129-200): (1, n-1), or (0, n) when eye vertex n-2 is specular */
       LMCU_BDPT_LT_NEE = 3 };        /* renderer::lt_nee_path (:205-288): (n-1, 1), the light subpath connects to the pinhole */

typedef struct LmcuRenderParams {
    int32_t width, height;                   /* film size */
    int32_t max_verts;                       /* pt: max path vertices */
    int32_t sampling_mode;                   /* pt: LMCU_SAMPLING_* (default mis) */
    int64_t spp;                             /* pt/ao: samples per pixel rendered by THIS call */
    int64_t sample_offset;                   /* pt/ao: global index of this call's first sample (sample split) */
    uint32_t seed;                           /* pt/ao: RNG seed */
    int32_t rescale;                         /* pt: 1 = film *= 1/processed (pixel) or w*h/processed (image) at the end
                                                (renderer_pt.cpp:307-312); 0 = leave sums (sample split) */
    int32_t use_constant_color;              /* raycast */
    int32_t visualize_normal;                /* raycast */
    int32_t has_color;                       /* raycast: `color` override present */
    int32_t count_traversal;                 /* pt: 1 = run the counters variant of extend/shadow (slower) */
    double bg_color[3];                      /* raycast */
    double color[3];                         /* raycast */
    double ao_distance;                      /* ao: occlusion ray tmax (0.2 in the reference) */
    int32_t primary_ray_sampling_mode;       /* pt: LMCU_PRIMARY_* (default pixel) */
    int32_t scheduler;                       /* pt: LMCU_SCHED_* (default sample) */
    int64_t num_samples;                     /* pt, image mode + sample scheduler: samples of the whole image by THIS call */
    int64_t samples_per_iter;                /* pt, image mode + time scheduler: samples per pass (0 = 100000, scheduler.cpp:159) */
    double render_time;                      /* pt, time scheduler: seconds; the pass in flight always completes */
    int32_t wavefronts;                      /* pt: staggered wavefront loops running side by side on their own streams
                                                (0 = default 2; 1 = one kernel at a time).  Same image for any value. */
    int32_t ao_mark_misses;                  /* ao: 1 = pixels whose primary ray misses come back as -1 (the reference never
                                                touches them, functest/renderer_ao.cpp:32-34: a host that renders into an
                                                existing film keeps their old value); 0 = they come back 0 */
    /* pt: called on the host thread that runs lmcu_render_pt, between wavefront batches: paths started so far, paths of the
     * job (0 with the time scheduler), seconds since the render began.  The hook of progress::update / update_time
     * (src/scheduler.cpp:45,91,133,173).  NULL = no reports. */
    void (*progress)(void* user, int64_t processed_paths, int64_t total_paths, double elapsed_seconds);
    void* progress_user;
    int32_t min_verts;                       /* bdpt: minimum number of path vertices (0 = 2, renderer_bdpt.cpp:51) */
    int32_t bdpt_strategies;                 /* bdpt: LMCU_BDPT_* — which connections (s, t) a sample evaluates (default all) */
} LmcuRenderParams;

typedef struct LmcuRenderStats {
    double elapsed;                          /* seconds, device time of the render (the "elapsed" Json key) */
    int64_t processed;                       /* the "processed" Json key: spp (pixel mode) or samples (image mode) rendered by this call */
    int64_t paths;                           /* paths traced: width*height*spp (pixel mode) or samples (image mode) */
    int64_t rays;                            /* closest-hit + shadow queries issued (device counter) */
    int64_t kernel_launches;                 /* kernels launched by this call */
    double bvh_nodes_per_ray;                /* mean wide nodes fetched per ray (counters build only, else 0) */
    double bvh_tris_per_ray;                 /* mean triangles fetched per ray (counters build only, else 0) */
    double shadow_nodes_per_ray;             /* same for the shadow (any-hit) rays */
    double shadow_tris_per_ray;
    double extend_seconds;                   /* pt: summed device time of the extend launches (CUDA events) */
    double shade_seconds;                    /* pt: ... of the shade launches */
    double shadow_seconds;                   /* pt: ... of the shadow launches */
    int64_t extend_launches;                 /* pt: number of extend launches (= wavefront iterations) */
    int64_t extend_rays;                     /* pt: closest-hit queries */
    int64_t shadow_rays;                     /* pt: shadow queries */
} LmcuRenderStats;

typedef struct LmcuBuildOptions {
    int32_t max_leaf_tris;                   /* 1..3 triangles per leaf slot (0 = default 2) */
    int32_t two_level;                       /* instancing-aware layout (the analogue of accel::embreeinstanced,
                                                plugin/accel_embree/accel_embree_instanced.cpp:96-165): one tree per mesh over
                                                its object-space triangles + one tree over the primitives, instead of one tree
                                                over the flattened scene.  Same hits (the leaf test transforms the vertices
                                                with the flattened build's operations); memory and build time follow the
                                                UNIQUE triangle count (C4: 4.95 MB and 10 ms against 4.75 GB and 40 ms); the
                                                walk is slower (C4: 0.78 x the flattened tree's paths per second: an
                                                object-space ray per instance entered, a vertex transform per leaf test).
                                                1 = on (when every transform is an invertible affine map), -1 = off,
                                                0 = automatic: flattened while it fits — two-level only when the flattened
                                                build of an instanced scene would need more than half of the free device
                                                memory.  A two-level accel is rebuilt, not serialised: lmcu_accel_save()
                                                answers LMCU_ERR_UNSUPPORTED. */
    int32_t reserved_[6];
} LmcuBuildOptions;

typedef struct LmcuBuildStats {
    double elapsed;                          /* seconds, device time of flatten + build */
    int64_t num_triangles;
    int64_t num_wide_nodes;
    int64_t node_bytes;
    int64_t triangle_bytes;
    int32_t triangles_fp32;                  /* 1 if every world-space vertex is float-representable (48 B triangles) */
    int32_t two_level;                       /* 1 if the two-level layout was built: num_triangles counts the flattened triangles,
                                                num_wide_nodes / node_bytes / triangle_bytes what is resident (all trees, the
                                                object-space triangles and the instance table) */
    double sah_cost;                         /* surface-area cost of the wide tree: expected wide-node visits + triangle tests
                                                of a random line through the scene box (sum of box area / root box area
                                                over nodes, plus the same over leaf slots times their triangle count) */
} LmcuBuildStats;

typedef struct LmcuScene LmcuScene;

/* -- library ---------------------------------------------------------------------------------- */
const char* lmcu_last_error(void);
int lmcu_device_count(void);
const char* lmcu_version(void);

/* Device blocks (path pools, films, the builder's temporaries) are recycled per device inside the library: a freed block is
 * parked and handed out again on the next request of the same size, at most LMCU_CACHE_GB (default 24) GiB per device — over
 * the cap the blocks parked longest ago go back to the driver.  lmcu_cache_trim() hands every parked block of `device` back
 * (after a device synchronise); lmcu_cache_stats() reports what is parked. */
int lmcu_cache_trim(int device);
int lmcu_cache_stats(int device, int64_t* parked_bytes, int64_t* parked_blocks);
/* Measured read bandwidth of the L2 (GB/s) over a `bytes`-sized buffer that every CTA streams repeatedly: the roofline
 * denominator for traversal of an L2-resident BVH (bench.py). */
int lmcu_probe_l2_bandwidth(int device, int64_t bytes, double* gb_per_s);

/* -- scene + accel: replaces lm::Accel::build / intersect (include/lm/accel.h:38,70) ---------- */
int lmcu_scene_create(int device, LmcuScene** out);
void lmcu_scene_destroy(LmcuScene* scene);
/* Run all later work of this scene on `stream` (a cudaStream_t of the caller, e.g. torch's current
 * stream; NULL restores the scene's own stream). */
int lmcu_scene_set_stream(LmcuScene* scene, void* stream);
/* Upload + flatten (accel_sahbvh.cpp:132-151: world p = M * (p,1) in FP64, glm operation order). */
int lmcu_scene_upload(LmcuScene* scene, const LmcuSceneDesc* desc);
/* Build the wide BVH (replaces accel_sahbvh.cpp:156-266). */
int lmcu_accel_build(LmcuScene* scene, const LmcuBuildOptions* opts, LmcuBuildStats* stats);
/* Closest hit for a batch of rays in HOST memory (replaces accel_sahbvh.cpp:269-299 per ray). */
int lmcu_accel_intersect(LmcuScene* scene, const LmcuRay* rays, int64_t n, LmcuHit* hits);
/* ONE ray: the scalar lm::Accel::intersect contract (include/lm/accel.h:70, served by accel_sahbvh.cpp:269-299; called concurrently from every OpenMP
 * thread by Scene::intersect, src/scene.cpp:320).  Thread-safe and lock-free: each calling thread owns a pinned,
 * device-mapped (ray, hit) slot and a stream; a call is one kernel launch and one stream synchronise, no allocation
 * and no memcpy.  Same hit, bit for bit, as lmcu_accel_intersect. */
int lmcu_accel_intersect1(LmcuScene* scene, const LmcuRay* ray, LmcuHit* hit);
/* Same with DEVICE buffers on `stream` (a cudaStream_t, may be NULL); asynchronous.  Batches in flight on different
 * streams are independent (each launch has its own work cursor).  A traversal-stack overflow (deeper than 48 wide-node
 * groups: never seen) is reported by the next lmcu_accel_intersect / lmcu_render_* call on the scene. */
int lmcu_accel_intersect_device(LmcuScene* scene, const void* d_rays, int64_t n, void* d_hits, void* stream);
/* Mean wide nodes / triangles fetched per ray by the last counted batch (counters variant of the kernel). */
int lmcu_accel_intersect_counted(LmcuScene* scene, const LmcuRay* rays, int64_t n, LmcuHit* hits,
                                 double* nodes_per_ray, double* tris_per_ray);

/* Serialisation of the built accel: replaces Component::save / load of an Accel (include/lm/component.h:210-266;
 * accel::sahbvh archives nodes_, trs_, indices_, src/accel/accel_sahbvh.cpp:122-124).  The blob holds the wide nodes, the
 * triangles in leaf order, the shading records and the primitive table, with a checksum; lmcu_accel_load() restores them
 * without running the builder.  On a scene WITH an uploaded description the blob must match it (triangle and primitive
 * counts), and the renderers work as after lmcu_accel_build(); on a fresh scene the result serves lmcu_accel_intersect*
 * only.  A blob is valid for the library version that wrote it. */
int lmcu_accel_save_size(LmcuScene* scene, int64_t* bytes);
int lmcu_accel_save(LmcuScene* scene, void* buffer, int64_t bytes);
int lmcu_accel_load(LmcuScene* scene, const void* buffer, int64_t bytes);

/* -- renderers: replace lm::Renderer::render (include/lm/renderer.h:25) ------------------------ */
/* film_rgb: HOST buffer of width*height*3 floats, or NULL to keep the result only in the device film. */
int lmcu_render_raycast(LmcuScene* scene, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats);
int lmcu_render_ao(LmcuScene* scene, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats);
int lmcu_render_pt(LmcuScene* scene, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats);
/* renderer::lt (src/renderer/renderer_lt.cpp:56-167): light tracing with next-event estimation toward the pinhole — paths start
 * on a light (Light::sample_ray of a uniformly chosen light), every non-specular surface vertex is connected to the camera and
 * the contribution splatted at the raster position of the connection (arbitrary-pixel splats).  Keys used: max_verts, seed,
 * scheduler (sample: num_samples light paths; time: render_time, samples_per_iter), sample_offset, rescale (film *=
 * w*h/processed, :162); the scheduler is always the image-space one (:44-47), primary_ray_sampling_mode and spp are ignored. */
int lmcu_render_lt(LmcuScene* scene, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats);
/* renderer::bdpt (src/renderer/renderer_bdpt.cpp:304-407, include/lm/bidir.h): per sample one eye and one light subpath of up
 * to max_verts vertices (no Russian roulette), every connection (s, t) with min_verts <= s + t <= max_verts, power-heuristic
 * weights over all strategies of the connected path.  Keys used: max_verts (2..20), min_verts (0 = 2), seed, scheduler (sample:
 * num_samples; time: render_time, samples_per_iter), sample_offset, rescale (film *= w*h/processed, :390-391).  Image-space
 * scheduler only (renderer_bdpt.cpp:54-56).  bdpt_strategies selects renderer::pt_naive_path / pt_nee_path / lt_nee_path: the
 * same subpaths and connections restricted to one strategy family, without weights; the subpath no strategy extends is
 * not walked past its endpoint. */
int lmcu_render_bdpt(LmcuScene* scene, const LmcuRenderParams* p, float* film_rgb, LmcuRenderStats* stats);
/* Device film of the last render (width*height*3 floats) for the multi-GPU film reduce
 * (sample split: every rank renders with rescale=0, NCCL-reduces this buffer, then scales). */
int lmcu_film_device_ptr(LmcuScene* scene, void** d_film, int64_t* num_floats);
/* film *= s on the device (film::bitmap::rescale, src/film/film_bitmap.cpp:244-248). */
int lmcu_film_rescale(LmcuScene* scene, double s);
/* Copy the device film to a host buffer. */
int lmcu_film_download(LmcuScene* scene, float* film_rgb);

/* -- multi-GPU film combine over NVLink peer memory (sample split) -----------------------
 * Every GPU renders its own samples with rescale=0 into its own device film.  The root then runs ONE
 * kernel that reads the peers' films directly over NVLink (P2P loads), adds them to its own film and
 * applies film::rescale (src/film/film_bitmap.cpp:244-248, called once by renderer_pt.cpp:307-309) as
 * the epilogue: film_root = (film_root + sum_p film_p) * scale.  Peers in the SAME process pass their
 * lmcu_film_device_ptr() after lmcu_enable_peer_access(); peers in OTHER processes (one rank per GPU)
 * export a CUDA IPC handle and the root opens it.  The caller orders the kernels (stream/event/barrier):
 * the peers' renders must have completed before the call and their films must stay untouched until
 * it returns (it synchronises the root's stream). */
#define LMCU_IPC_HANDLE_BYTES 64
#define LMCU_MAX_PEERS 15
int lmcu_film_ipc_export(LmcuScene* scene, void* handle /* LMCU_IPC_HANDLE_BYTES */);
int lmcu_ipc_open(int device, const void* handle, void** d_ptr);
int lmcu_ipc_close(int device, void* d_ptr);
int lmcu_enable_peer_access(int device, int peer_device);
int lmcu_film_reduce_peers(LmcuScene* root, void* const* peer_films, int num_peers, double scale);
/* Slice-parallel form of the same combine (one call per GPU, all at once) — the multi-GPU counterpart of Film::accum +
 * Film::rescale (src/film/film_bitmap.cpp:220-234,244-248): out[i] = scale * sum_k films[k][i] for
 * i in [begin, end).  `films` = this GPU's own film and the peers' films (peer-mapped), `out` = the root's film
 * (peer-mapped on every GPU but the root; it may also be one of `films`).  With the film cut into G slices no GPU
 * receives more than (G-1)/G of a film and all NVLink ports work at once, where lmcu_film_reduce_peers pulls G-1 whole
 * films into the root.  `begin` must be a multiple of 4 floats.  Runs on `scene`'s device and stream and synchronises
 * it; the caller barriers before (all renders complete) and after (all slices stored). */
int lmcu_film_reduce_slice(LmcuScene* scene, void* const* films, int num_films, void* out, int64_t begin, int64_t end, double scale);

#ifdef __cplusplus
}
#endif
#endif /* LMCU_H */
