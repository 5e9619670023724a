/*
 * pt_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the wavefront loop of nkkk98/MyGPURaytracer
 * (apps/src/pathtrace.cu:248-655, apps/src/intersections.h,
 * apps/src/interactions.h) with the glm 0.9.6.3 expression trees and the
 * thrust::minstd_rand / uniform_real_distribution arithmetic written out.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (mygpuraytracer_b200/)
 * never links, imports or calls anything in this directory.
 *
 * Parity pin: with trig_mode = B2PT_TRIG_NATIVE (libm) this oracle is
 * bit-identical to a host build of the reference's own headers
 * (oracle/_ref/ref_cpu, built by oracle/Makefile from /root/reference) on
 * every stage of every shipped scene -- tests/test_oracle_vs_ref.py -- and
 * it reproduces the golden vectors of SURVEY.md section 8c
 * (tests/test_oracle_golden.py, tests/golden/).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
 * No FMA contraction, IEEE divide and sqrt: any implementation with the same
 * operation order produces the same bits.
 */
#ifndef PT_ORACLE_H_
#define PT_ORACLE_H_

#include "../include/b2pt.h" /* POD types only */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG (apps/src/intersections.h:12-20, apps/src/pathtrace.cu:66-70,
 *      thrust/random/detail/linear_congruential_engine.inl,
 *      thrust/random/detail/uniform_real_distribution.inl) ----------------- */
uint32_t oracle_utilhash(uint32_t a);
uint32_t oracle_seed(int32_t iter, int32_t index, int32_t depth); /* engine state after seeding */
uint32_t oracle_minstd_next(uint32_t* state);
float oracle_uniform(uint32_t* state, float a, float b);

/* ---- geometry primitives (apps/src/intersections.h) ---------------------- */
/* Each returns t (-1 on a miss) and fills normal[3]. */
float oracle_box_test(const B2ptGeom* g, const float o[3], const float d[3], float normal[3]);
float oracle_sphere_test(const B2ptGeom* g, const float o[3], const float d[3], float normal[3]);
/* glm::intersectRayTriangle; returns 1 on a hit and fills bary (u, v, t). */
int oracle_ray_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
                        const float v2[3], float bary[3]);
/* meshIntersectionTest: brute force over the geom's faces. */
float oracle_mesh_test(const B2ptScene* s, const B2ptGeom* g, const float o[3], const float d[3],
                       float normal[3], float uv[2], int32_t* face);

/* calculateRandomDirectionInHemisphere, apps/src/interactions.h:12-44. */
void oracle_hemisphere(const float n[3], uint32_t* rng, int32_t trig_mode, float out[3]);

/* The portable sin/cos shared bit-for-bit with the CUDA kernels. */
void oracle_sincos_portable(float x, float* s, float* c);

/* ---- stages (all arrays SoA, caller-allocated) ------------------------------ */

/* generateRayFromCamera, apps/src/pathtrace.cu:248-297.  P = W*H entries. */
int oracle_generate(const B2ptCamera* cam, const B2ptOptions* opt, int32_t iter, int32_t trace_depth,
                    float* origin, float* dir, float* color, int32_t* pixel, int32_t* bounces);

/* computeIntersections, apps/src/pathtrace.cu:303-386.  On a miss: t=-1,
 * normal=0, uv=0, material=0 (the memset of :595), geom=-1, face=-1. */
int oracle_intersect(const B2ptScene* s, int32_t n, const float* origin, const float* dir, float* t,
                     float* normal, float* uv, int32_t* geom, int32_t* face, int32_t* material);

/* thrust::sort_by_key with sortByMaterial, apps/src/pathtrace.cu:512-516,612:
 * perm[k] = pre-sort index of sorted slot k (stable, descending material). */
int oracle_sort_perm(int32_t n, const int32_t* material, int32_t* perm);

/* shadeFakeMaterial + scatterRay, apps/src/pathtrace.cu:397-498,
 * apps/src/interactions.h:112-258.  Arrays are in SORTED slot order; the
 * path arrays are updated in place.  depth is the value after depth++ (:614).
 * albedo may be NULL. */
int oracle_shade(const B2ptScene* s, const B2ptOptions* opt, int32_t iter, int32_t depth, int32_t n,
                 const float* hit_t, const float* hit_normal, const float* hit_uv,
                 const int32_t* hit_geom, const int32_t* hit_material, float* origin, float* dir,
                 float* color, const int32_t* pixel, int32_t* bounces, float* albedo);

/* thrust::stable_partition with isTerminate, apps/src/pathtrace.cu:518-522,649:
 * perm[k] = source slot of output slot k; returns the number of live paths. */
int oracle_partition_perm(int32_t n, const int32_t* bounces, int32_t* perm);

/* finalGather, apps/src/pathtrace.cu:501-510. */
int oracle_gather(int32_t n, float* image, const float* color, const int32_t* pixel);

/* Whole iterations: iter_first, iter_first+stride, ... (count of them), the
 * loop of apps/src/pathtrace.cu:527-655.  image/albedo are W*H*3 floats and
 * are accumulated into (albedo written at iter==1 only).  n_live, if not
 * NULL, receives trace_depth+1 live counts of the LAST iteration;
 * segments, if not NULL, receives the total number of path segments traced
 * (sum over iterations and depths of the live count). */
int oracle_render(const B2ptScene* s, const B2ptOptions* opt, int32_t iter_first, int32_t count,
                  int32_t stride, float* image, float* albedo, int32_t* n_live, int64_t* segments);

/* Order in which the two lens draws of glm::vec2(uDOF(rng), uDOF(rng))
 * (apps/src/pathtrace.cu:285) are taken: 0 = left to right (nvcc device
 * code), 1 = right to left (g++ host build of the same line). */
void oracle_set_dof_arg_order(int right_to_left);

/* Number of OpenMP threads the stage loops will use. */
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
