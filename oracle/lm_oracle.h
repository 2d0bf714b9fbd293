/* lm_oracle.h — C ABI of the CPU restatement (oracle/lm_oracle.c).  TEST INFRASTRUCTURE ONLY. */
#ifndef LM_ORACLE_H
#define LM_ORACLE_H
#include "../include/lmcu.h"

typedef struct OrcScene OrcScene;

/* Flatten + build the reference's binary SAH BVH (accel_sahbvh.cpp:127-266). Returns NULL on error. */
OrcScene* orc_scene_create(const LmcuSceneDesc* desc);
void orc_scene_destroy(OrcScene* s);
/* Accel_SAHBVH::intersect for a batch (accel_sahbvh.cpp:269-299); OpenMP over rays. */
void orc_intersect(const OrcScene* s, const LmcuRay* rays, int64_t n, LmcuHit* hits);
/* renderer::raycast / renderer::ao / renderer::pt into a double film [h][w][3]; num_threads <= 0: all cores.
 * With num_threads == 1 the RNG stream is the reference's (mt19937 seeded seed+0, samples in index order). */
void orc_render_raycast(const OrcScene* s, const LmcuRenderParams* p, double* film);
void orc_render_ao(const OrcScene* s, const LmcuRenderParams* p, double* film, int num_threads);
void orc_render_pt(const OrcScene* s, const LmcuRenderParams* p, double* film, int num_threads);
/* std::mt19937 + std::uniform_real_distribution<double> as libstdc++ implements them (math.h:204-220) */
void orc_rng_fill(uint32_t seed, double* out, int n);

#endif
