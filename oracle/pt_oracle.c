/*
 * pt_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See pt_oracle.h for the contract and the parity pin.
 *
 * Every function cites the reference lines it restates.  glm 0.9.6.3
 * expression trees (apps/external/include/glm):
 *   dot(vec3)      (x*x' + y*y') + z*z'                detail/func_geometric.inl:65-71
 *   cross          (ay*bz - by*az, az*bx - bz*ax, ax*by - bx*ay)          :134-142
 *   normalize      v * (1 / sqrt(dot(v,v)))            :154-159, func_exponential.inl:150-153
 *   length         sqrt(dot(v,v))                      :95-100
 *   reflect        I - (N*dot(N,I))*2                  :176-179
 *   refract        (eta*I - (eta*d + sqrt(k))*N) * float(k>=0)             :193-200
 *   mat4*vec4      (m0*v0 + m1*v1) + (m2*v2 + m3*v3)   detail/type_mat4x4.inl:617-628
 *   mat3*vec3      (m0*v0 + m1*v1) + m2*v2             detail/type_mat3x3.inl:487-493
 *   min/max/abs    x<y?x:y / x>y?x:y / x>=0?x:-x       detail/func_common.inl:56,413,434
 */
#include "pt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small vector helpers, written so that the operation order is explicit      */
/* ------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }
static inline void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 cross3(v3 a, v3 b) {
  return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 normalize3(v3 a) { return muls(a, 1.0f / sqrtf(dot3(a, a))); }
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
static inline float glm_min(float x, float y) { return x < y ? x : y; }
static inline float glm_max(float x, float y) { return x > y ? x : y; }
static inline float glm_abs(float x) { return x >= 0.0f ? x : -x; }
static inline v3 reflect3(v3 I, v3 N) { return sub(I, muls(muls(N, dot3(N, I)), 2.0f)); }

/* vec3(m * vec4(v, w)); m column-major. */
static inline v3 mat4_mul(const float* m, v3 v, float w) {
  v3 r;
  r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * w);
  r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * w);
  r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
  return r;
}

/* ------------------------------------------------------------------------- */
/* RNG                                                                        */
/* ------------------------------------------------------------------------- */

/* utilhash, apps/src/intersections.h:12-20. */
uint32_t oracle_utilhash(uint32_t a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

/* makeSeededRandomEngine, apps/src/pathtrace.cu:66-70, then
 * linear_congruential_engine::seed: x = s mod m, 0 -> 1 (m = 2^31-1). */
uint32_t oracle_seed(int32_t iter, int32_t index, int32_t depth) {
  uint32_t h = oracle_utilhash(0x80000000u | ((uint32_t)depth << 22) | (uint32_t)iter) ^
               oracle_utilhash((uint32_t)index);
  uint32_t x = h % 2147483647u;
  return x == 0u ? 1u : x;
}

/* minstd_rand: x <- 48271 x mod (2^31 - 1). */
uint32_t oracle_minstd_next(uint32_t* state) {
  *state = (uint32_t)(((uint64_t)(*state) * 48271u) % 2147483647u);
  return *state;
}

/* uniform_real_distribution<float>(a, b): float(x - min) / (1 + float(max - min))
 * with min = 1, max = 2^31 - 2; the denominator rounds to 2^31. */
float oracle_uniform(uint32_t* state, float a, float b) {
  float r = (float)(oracle_minstd_next(state) - 1u);
  r /= (1.0f + (float)(2147483646u - 1u));
  return (r * (b - a)) + a;
}

/* ------------------------------------------------------------------------- */
/* trig                                                                       */
/* ------------------------------------------------------------------------- */

/* Portable sin/cos: three-term Cody-Waite reduction by pi/2 and degree-7/8
 * minimax polynomials (Cephes sinf/cosf coefficients).  Uses +,-,* and one
 * rintf only, so that the CUDA kernels (built with -fmad=false) reproduce it
 * bit for bit.  Accurate to ~1 ulp for |x| < 100. */
void oracle_sincos_portable(float x, float* s, float* c) {
  float kf = rintf(x * 0.636619772f);
  int k = (int)kf;
  float r = x - kf * 1.5703125f;
  r = r - kf * 4.837512969970703125e-4f;
  r = r - kf * 7.54978995489188216e-8f;
  float z = r * r;
  float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
  float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z -
             0.5f * z + 1.0f;
  switch (k & 3) {
    case 0: *s = sp; *c = cp; break;
    case 1: *s = cp; *c = -sp; break;
    case 2: *s = -sp; *c = -cp; break;
    default: *s = -cp; *c = sp; break;
  }
}

static inline void sincos_mode(float x, int32_t mode, float* s, float* c) {
  if (mode == B2PT_TRIG_PORTABLE) {
    oracle_sincos_portable(x, s, c);
  } else {
    *s = sinf(x);
    *c = cosf(x);
  }
}

/* pow((1.0 - cosTheta), 5) in double, apps/src/interactions.h:153,192. */
static inline double pow5_mode(double a, int32_t mode) {
  if (mode == B2PT_TRIG_PORTABLE) {
    double a2 = a * a;
    return a2 * a2 * a;
  }
  return pow(a, 5.0);
}

/* glm::pow(x, exponent) -> powf, apps/src/interactions.h:128,204. */
static inline float powf_mode(float x, float e, int32_t mode) {
  if (e == 0.0f) return 1.0f; /* powf(anything, 0) == 1, also for NaN */
  (void)mode;
  return powf(x, e);
}

/* ------------------------------------------------------------------------- */
/* intersections.h                                                            */
/* ------------------------------------------------------------------------- */

/* getPointOnRay, apps/src/intersections.h:27-29. */
static inline v3 point_on_ray(v3 o, v3 d, float t) {
  return add(o, muls(normalize3(d), t - .0001f));
}

/* boxIntersectionTest, apps/src/intersections.h:48-90. */
float oracle_box_test(const B2ptGeom* g, const float o[3], const float d[3], float normal[3]) {
  v3 ro = ld3(o), rd = ld3(d);
  v3 qo = mat4_mul(g->inverse_transform, ro, 1.0f);
  v3 qd = normalize3(mat4_mul(g->inverse_transform, rd, 0.0f));
  float qo_a[3] = {qo.x, qo.y, qo.z};
  float qd_a[3] = {qd.x, qd.y, qd.z};
  float tmin = -1e38f, tmax = 1e38f;
  float tmin_n[3] = {0, 0, 0}, tmax_n[3] = {0, 0, 0};
  for (int xyz = 0; xyz < 3; ++xyz) {
    float qdxyz = qd_a[xyz];
    float t1 = (-0.5f - qo_a[xyz]) / qdxyz;
    float t2 = (+0.5f - qo_a[xyz]) / qdxyz;
    float ta = glm_min(t1, t2);
    float tb = glm_max(t1, t2);
    float n[3] = {0, 0, 0};
    n[xyz] = t2 < t1 ? +1.0f : -1.0f;
    if (ta > 0 && ta > tmin) {
      tmin = ta;
      memcpy(tmin_n, n, sizeof n);
    }
    if (tb < tmax) {
      tmax = tb;
      memcpy(tmax_n, n, sizeof n);
    }
  }
  if (tmax >= tmin && tmax > 0) {
    if (tmin <= 0) {
      tmin = tmax;
      memcpy(tmin_n, tmax_n, sizeof tmin_n);
    }
    v3 ip = mat4_mul(g->transform, point_on_ray(qo, qd, tmin), 1.0f);
    v3 nn = normalize3(mat4_mul(g->inv_transpose, ld3(tmin_n), 0.0f));
    st3(normal, nn);
    return length3(sub(ro, ip));
  }
  return -1.0f;
}

/* sphereIntersectionTest, apps/src/intersections.h:102-144. */
float oracle_sphere_test(const B2ptGeom* g, const float o[3], const float d[3], float normal[3]) {
  v3 r_o = ld3(o), r_d = ld3(d);
  v3 ro = mat4_mul(g->inverse_transform, r_o, 1.0f);
  v3 rd = normalize3(mat4_mul(g->inverse_transform, r_d, 0.0f));
  float vDotDirection = dot3(ro, rd);
  float radicand = vDotDirection * vDotDirection - (dot3(ro, ro) - 0.25f /* powf(.5, 2) */);
  if (radicand < 0) return -1.0f;
  float squareRoot = sqrtf(radicand);
  float firstTerm = -vDotDirection;
  float t1 = firstTerm + squareRoot;
  float t2 = firstTerm - squareRoot;
  float t;
  int outside;
  if (t1 < 0 && t2 < 0) {
    return -1.0f;
  } else if (t1 > 0 && t2 > 0) {
    t = fminf(t1, t2); /* unqualified min() -> CUDA's float overload */
    outside = 1;
  } else {
    t = fmaxf(t1, t2);
    outside = 0;
  }
  v3 obj = point_on_ray(ro, rd, t);
  v3 ip = mat4_mul(g->transform, obj, 1.0f);
  v3 nn = normalize3(mat4_mul(g->inv_transpose, obj, 0.0f));
  if (!outside) nn = neg(nn);
  st3(normal, nn);
  return length3(sub(r_o, ip));
}

/* glm::intersectRayTriangle, apps/external/include/glm/gtx/intersect.inl:36-74
 * (back faces culled: a < FLT_EPSILON misses). */
int oracle_ray_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
                        const float v2[3], float bary[3]) {
  v3 orig = ld3(o), dir = ld3(d), a0 = ld3(v0);
  v3 e1 = sub(ld3(v1), a0);
  v3 e2 = sub(ld3(v2), a0);
  v3 p = cross3(dir, e2);
  float a = dot3(e1, p);
  if (a < FLT_EPSILON) return 0;
  float f = 1.0f / a;
  v3 s = sub(orig, a0);
  bary[0] = f * dot3(s, p);
  if (bary[0] < 0.0f) return 0;
  if (bary[0] > 1.0f) return 0;
  v3 q = cross3(s, e1);
  bary[1] = f * dot3(dir, q);
  if (bary[1] < 0.0f) return 0;
  if (bary[1] + bary[0] > 1.0f) return 0;
  bary[2] = f * dot3(e2, q);
  return bary[2] >= 0.0f;
}

/* Texture fetch as written at all six sites (e.g. intersections.h:270-276):
 * px = (int)(v*h)*w + (int)(u*w); three bytes at px*channels; /255.f.  The
 * reference does not bound the index (Q12); the oracle clamps it into the
 * image only to stay memory-safe, which changes nothing for uv in [0,1). */
static inline v3 fetch_texel(const B2ptTexture* tx, float u, float v) {
  int coordU = (int)(u * (float)tx->width);
  int coordV = (int)(v * (float)tx->height);
  long long pixelID = (long long)coordV * tx->width + coordU;
  long long last = (long long)tx->width * tx->height - 1;
  if (pixelID < 0) pixelID = 0;
  if (pixelID > last) pixelID = last;
  /* Three consecutive bytes, whatever the channel count: with 1 or 2 channels the reference reads into the next
   * texel(s), and past the end of the image for the last ones (undefined there; read as the last byte here). */
  const long long end = ((long long)tx->width * tx->height) * tx->channels - 1;
  const long long at = pixelID * tx->channels;
  unsigned int colR = tx->texels[at], colG = tx->texels[at + 1 > end ? end : at + 1], colB = tx->texels[at + 2 > end ? end : at + 2];
  return V((float)colR / 255.f, (float)colG / 255.f, (float)colB / 255.f);
}

static inline const B2ptTexture* geom_tex(const B2ptScene* s, int32_t idx) {
  if (idx < 0 || idx >= s->n_textures) return NULL;
  const B2ptTexture* t = &s->textures[idx];
  return t->channels ? t : NULL;
}

/* The four maps of an OBJ hit: the reference takes them from the geom (one material per OBJ,
 * apps/src/scene.cpp:134-231).  With B2ptScene::material_textures (per-face materials, an extension the
 * reference does not have: it discards tinyobj's material_ids, scene.cpp:121-122) they come from the
 * material of the face that was hit.  which: 0 kd, 1 ks, 2 bump, 3 ke. */
static inline const B2ptTexture* obj_tex(const B2ptScene* s, const B2ptGeom* g, int32_t material, int which) {
  if (s->material_textures && material >= 0 && material < s->n_materials)
    return geom_tex(s, s->material_textures[4 * (size_t)material + which]);
  const int32_t idx = which == 0 ? g->tex_kd : which == 1 ? g->tex_ks : which == 2 ? g->tex_bump : g->tex_ke;
  return geom_tex(s, idx);
}
static inline int32_t face_material_of(const B2ptScene* s, const B2ptGeom* g, int32_t face) {
  return (s->face_material && face >= 0) ? s->face_material[(size_t)g->face_begin + face] : g->material_id;
}

/* meshIntersectionTest, apps/src/intersections.h:207-282. */
float oracle_mesh_test(const B2ptScene* s, const B2ptGeom* g, const float o[3], const float d[3],
                       float normal[3], float uv[2], int32_t* face) {
  v3 r_o = ld3(o), r_d = ld3(d);
  v3 qo = mat4_mul(g->inverse_transform, r_o, 1.0f);
  v3 qd = normalize3(mat4_mul(g->inverse_transform, r_d, 0.0f));
  float qo_a[3] = {qo.x, qo.y, qo.z}, qd_a[3] = {qd.x, qd.y, qd.z};
  float tmin = FLT_MAX;
  int nearest = -1;
  float tu = 0, tv = 0;
  const float* pos = s->face_pos + (size_t)g->face_begin * 9;
  const float* fuv = s->face_uv + (size_t)g->face_begin * 6;
  for (int j = 0; j < g->face_count; j++) {
    const float* tri = pos + (size_t)j * 9;
    float bary[3];
    if (oracle_ray_triangle(qo_a, qd_a, tri, tri + 3, tri + 6, bary)) {
      float w = 1 - bary[0] - bary[1];
      v3 p = add(add(muls(ld3(tri), w), muls(ld3(tri + 3), bary[0])), muls(ld3(tri + 6), bary[1]));
      float t = length3(sub(qo, p)); /* glm::distance(p, q.origin) = length(q.origin - p) */
      if (t < tmin) {
        tmin = t;
        nearest = j;
        const float* tuv = fuv + (size_t)j * 6;
        tu = (w * tuv[0] + bary[0] * tuv[2]) + bary[1] * tuv[4];
        tv = (w * tuv[1] + bary[0] * tuv[3]) + bary[1] * tuv[5];
      }
    }
  }
  if (nearest == -1) return -1.0f;
  const float* tri = pos + (size_t)nearest * 9;
  const float* tuv = fuv + (size_t)nearest * 6;
  v3 e1 = sub(ld3(tri + 3), ld3(tri));
  v3 e2 = sub(ld3(tri + 6), ld3(tri));
  v3 objn = normalize3(cross3(e1, e2));
  v3 nn = normalize3(mat4_mul(g->inv_transpose, objn, 0.0f));
  const B2ptTexture* bump = obj_tex(s, g, face_material_of(s, g, nearest), 2);
  if (g->type == B2PT_OBJ && bump) {
    float d1x = tuv[2] - tuv[0], d1y = tuv[3] - tuv[1];
    float d2x = tuv[4] - tuv[0], d2y = tuv[5] - tuv[1];
    float f = 1.0f / (d1x * d2y - d2x * d1y);
    v3 tangent = V(f * (d2y * e1.x - d1y * e2.x), f * (d2y * e1.y - d1y * e2.y),
                   f * (d2y * e1.z - d1y * e2.z));
    tangent = normalize3(tangent);
    v3 bitangent = V(f * (-d2x * e1.x + d1x * e2.x), f * (-d2x * e1.y + d1x * e2.y),
                     f * (-d2x * e1.z + d1x * e2.z));
    bitangent = normalize3(bitangent);
    v3 T = normalize3(mat4_mul(g->transform, tangent, 0.0f));
    v3 B = normalize3(mat4_mul(g->transform, bitangent, 0.0f));
    v3 N = nn;
    v3 tsn = normalize3(fetch_texel(bump, tu, tv));
    tsn = normalize3(V(tsn.x * 2.0f - 1.0f, tsn.y * 2.0f - 1.0f, tsn.z * 2.0f - 1.0f));
    v3 r = V((T.x * tsn.x + B.x * tsn.y) + N.x * tsn.z, (T.y * tsn.x + B.y * tsn.y) + N.y * tsn.z,
             (T.z * tsn.x + B.z * tsn.y) + N.z * tsn.z);
    nn = normalize3(r);
  }
  st3(normal, nn);
  uv[0] = tu;
  uv[1] = tv;
  *face = nearest;
  return tmin;
}

/* ------------------------------------------------------------------------- */
/* interactions.h                                                             */
/* ------------------------------------------------------------------------- */

/* calculateRandomDirectionInHemisphere, apps/src/interactions.h:12-44. */
static v3 hemisphere(v3 normal, uint32_t* rng, int32_t trig_mode) {
  float up = sqrtf(oracle_uniform(rng, 0.0f, 1.0f));
  float over = sqrtf(1 - up * up);
  float around = oracle_uniform(rng, 0.0f, 1.0f) * 6.2831853071795864769252867665590057683943f;
  v3 dnn;
  const float SQRT_OF_ONE_THIRD = 0.5773502691896257645091487805019574556476f;
  if (fabsf(normal.x) < SQRT_OF_ONE_THIRD) {
    dnn = V(1, 0, 0);
  } else if (fabsf(normal.y) < SQRT_OF_ONE_THIRD) {
    dnn = V(0, 1, 0);
  } else {
    dnn = V(0, 0, 1);
  }
  v3 p1 = normalize3(cross3(normal, dnn));
  v3 p2 = normalize3(cross3(normal, p1));
  float sn, cs;
  sincos_mode(around, trig_mode, &sn, &cs);
  return add(add(muls(normal, up), muls(p1, cs * over)), muls(p2, sn * over));
}

void oracle_hemisphere(const float n[3], uint32_t* rng, int32_t trig_mode, float out[3]) {
  st3(out, hemisphere(ld3(n), rng, trig_mode));
}

/* ------------------------------------------------------------------------- */
/* stages                                                                     */
/* ------------------------------------------------------------------------- */

static int g_dof_args_right_to_left = 0;
void oracle_set_dof_arg_order(int right_to_left) { g_dof_args_right_to_left = right_to_left; }

/* ConcentricSampleDisk, apps/src/pathtrace.cu:225-239. */
static void concentric_disk(float px, float py, int32_t trig_mode, float* ox, float* oy) {
  float ux = 2.f * px - 1.0f, uy = 2.f * py - 1.0f;
  if (ux == 0 && uy == 0) {
    *ox = 0;
    *oy = 0;
    return;
  }
  float theta, r;
  if (fabsf(ux) > fabsf(uy)) {
    r = ux;
    theta = 0.785398f * (uy / ux);
  } else {
    r = uy;
    theta = 1.570796f - 0.785398f * (ux / uy);
  }
  float sn, cs;
  sincos_mode(theta, trig_mode, &sn, &cs);
  *ox = r * cs;
  *oy = r * sn;
}

/* generateRayFromCamera, apps/src/pathtrace.cu:248-297. */
int oracle_generate(const B2ptCamera* cam, const B2ptOptions* opt, int32_t iter, int32_t trace_depth,
                    float* origin, float* dir, float* color, int32_t* pixel, int32_t* bounces) {
  const int W = cam->resolution[0], H = cam->resolution[1];
  const v3 view = ld3(cam->view), right = ld3(cam->right), up = ld3(cam->up);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      int index = x + (y * W);
      uint32_t rng = oracle_seed(iter, index, trace_depth);
      v3 o = ld3(cam->position);
      float antia_x = (float)x, antia_y = (float)y;
      if (opt->antialiasing) {
        uint32_t rngAA = oracle_seed(iter, index, trace_depth);
        antia_x += oracle_uniform(&rngAA, -0.5f, 0.5f);
        antia_y += oracle_uniform(&rngAA, -0.5f, 0.5f);
      }
      v3 d = normalize3(sub(sub(view, muls(muls(right, cam->pixel_length[0]), antia_x - (float)W * 0.5f)),
                            muls(muls(up, cam->pixel_length[1]), antia_y - (float)H * 0.5f)));
      if (opt->depth_of_field && opt->lens_radius > 0) {
        /* glm::vec2(uDOF(rng), uDOF(rng)): the evaluation order of the two
         * arguments is unspecified in C++.  g++ (the reference's host pass,
         * oracle/_ref/ref_cpu) draws the SECOND argument first; nvcc's device
         * front end draws left to right (oracle/_ref/ref_gpu_dof).  The
         * oracle follows the device unless told otherwise. */
        float u0, u1;
        if (g_dof_args_right_to_left) {
          u1 = oracle_uniform(&rng, 0.0f, 1.0f);
          u0 = oracle_uniform(&rng, 0.0f, 1.0f);
        } else {
          u0 = oracle_uniform(&rng, 0.0f, 1.0f);
          u1 = oracle_uniform(&rng, 0.0f, 1.0f);
        }
        float lx, ly;
        concentric_disk(u0, u1, opt->trig_mode, &lx, &ly);
        lx = opt->lens_radius * lx;
        ly = opt->lens_radius * ly;
        float ft = glm_abs(opt->focal_distance / d.z);
        v3 pFocus = add(o, muls(d, ft));
        o = add(o, V(lx, ly, 0.0f));
        d = normalize3(sub(pFocus, o));
      }
      st3(origin + 3 * (size_t)index, o);
      st3(dir + 3 * (size_t)index, d);
      st3(color + 3 * (size_t)index, V(1.0f, 1.0f, 1.0f));
      pixel[index] = index;
      bounces[index] = trace_depth;
    }
  }
  return 0;
}

/* computeIntersections, apps/src/pathtrace.cu:303-386. */
int oracle_intersect(const B2ptScene* s, int32_t n, const float* origin, const float* dir, float* t_out,
                     float* normal, float* uv, int32_t* geom, int32_t* face, int32_t* material) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < n; ++i) {
    const float* o = origin + 3 * (size_t)i;
    const float* d = dir + 3 * (size_t)i;
    float t_min = FLT_MAX;
    int hit = -1, hit_face = -1;
    float nrm[3] = {0, 0, 0}, huv[2] = {0, 0};
    for (int gi = 0; gi < s->n_geoms; ++gi) {
      const B2ptGeom* g = &s->geoms[gi];
      float t = -1.0f, tn[3] = {0, 0, 0}, tuv[2] = {0, 0};
      int tf = -1;
      if (g->type == B2PT_CUBE) {
        t = oracle_box_test(g, o, d, tn);
      } else if (g->type == B2PT_SPHERE) {
        t = oracle_sphere_test(g, o, d, tn);
      } else if (g->type == B2PT_OBJ) {
        t = oracle_mesh_test(s, g, o, d, tn, tuv, &tf);
      }
      if (t > 0.0f && t_min > t) {
        t_min = t;
        hit = gi;
        hit_face = tf;
        memcpy(nrm, tn, sizeof nrm);
        memcpy(huv, tuv, sizeof huv);
      }
    }
    if (hit == -1) {
      t_out[i] = -1.0f;
      normal[3 * (size_t)i] = normal[3 * (size_t)i + 1] = normal[3 * (size_t)i + 2] = 0.0f;
      uv[2 * (size_t)i] = uv[2 * (size_t)i + 1] = 0.0f;
      geom[i] = -1;
      face[i] = -1;
      material[i] = 0; /* cudaMemset of pathtrace.cu:595 */
    } else {
      t_out[i] = t_min;
      memcpy(normal + 3 * (size_t)i, nrm, sizeof nrm);
      memcpy(uv + 2 * (size_t)i, huv, sizeof huv);
      geom[i] = hit;
      face[i] = hit_face;
      material[i] = face_material_of(s, &s->geoms[hit], hit_face);
    }
  }
  return 0;
}

/* sort_by_key(sortByMaterial): the unique stable order by descending
 * materialId, computed as a counting sort. */
int oracle_sort_perm(int32_t n, const int32_t* material, int32_t* perm) {
  int32_t maxm = 0;
  for (int i = 0; i < n; ++i) {
    if (material[i] < 0) return B2PT_ERR_INVALID;
    if (material[i] > maxm) maxm = material[i];
  }
  int64_t* start = (int64_t*)calloc((size_t)maxm + 2, sizeof(int64_t));
  if (!start) return B2PT_ERR_NOMEM;
  for (int i = 0; i < n; ++i) start[maxm - material[i] + 1]++;
  for (int k = 0; k <= maxm; ++k) start[k + 1] += start[k];
  for (int i = 0; i < n; ++i) perm[start[maxm - material[i]]++] = i;
  free(start);
  return 0;
}

/* scatterRay, apps/src/interactions.h:112-258.  Returns with the segment
 * updated; *bounces may be set to 1 by the emissive-texture branch (:184). */
static void scatter(const B2ptScene* s, const B2ptOptions* opt, v3* o, v3* d, v3* color, int32_t* bounces,
                    v3 intersect, v3 n, float u, float v, int32_t geom_id, const B2ptMaterial* m,
                    uint32_t* rng) {
  const int32_t tm = opt->trig_mode;
  v3 spec_color = ld3(m->specular_color);
  if (m->has_reflective > 0) {
    v3 reflectDir = reflect3(*d, n);
    float spec = powf_mode(glm_max(dot3(neg(*d), reflectDir), 0.0f), m->specular_exponent, tm);
    *color = mulv(*color, muls(spec_color, m->has_reflective * spec));
    *o = add(intersect, muls(n, 0.01f));
    *d = reflectDir;
  } else if (m->has_refractive > 0) {
    float IoR1 = 1.0f, IoR2 = m->index_of_refraction;
    float cosTheta = dot3(neg(*d), n);
    if (cosTheta < 0) {
      n = muls(n, -1.0f); /* surfaceNormal *= -1 (int -> float) */
      IoR1 = IoR2;
      IoR2 = 1.0f;
      cosTheta = fabsf(cosTheta);
    }
    float sinTheta = (float)sqrt(1.0 - (double)(cosTheta * cosTheta));
    if (IoR1 / IoR2 * sinTheta > 1.0f) {
      *d = reflect3(*d, n);
    } else {
      float r0 = ((IoR1 - IoR2) / (IoR1 + IoR2)) * ((IoR1 - IoR2) / (IoR1 + IoR2));
      float coeff = (float)((double)r0 + (double)(1.0f - r0) * pow5_mode(1.0 - (double)cosTheta, tm));
      float random = oracle_uniform(rng, 0.0f, 1.0f);
      if (random < coeff) {
        *d = reflect3(*d, n);
      } else {
        float eta = IoR1 / IoR2;
        float dv = dot3(n, *d);
        float k = 1.0f - eta * eta * (1.0f - dv * dv);
        v3 r = sub(muls(*d, eta), muls(n, eta * dv + sqrtf(k)));
        *d = muls(r, (float)(k >= 0.0f));
      }
    }
    *color = mulv(*color, spec_color);
    *o = add(intersect, muls(*d, 0.01f));
  } else if (s->geoms[geom_id].type == B2PT_OBJ) {
    const B2ptGeom* g = &s->geoms[geom_id];
    const int32_t mat_id = (int32_t)(m - s->materials);
    const B2ptTexture* ke = obj_tex(s, g, mat_id, 3);
    const B2ptTexture* ks = obj_tex(s, g, mat_id, 1);
    const B2ptTexture* kd = obj_tex(s, g, mat_id, 0);
    v3 emission = V(0, 0, 0);
    if (ke) emission = fetch_texel(ke, u, v);
    if (emission.x > FLT_EPSILON || emission.y > FLT_EPSILON || emission.z > FLT_EPSILON) {
      *color = mulv(*color, muls(emission, 5.0f));
      *bounces = 1;
      return;
    }
    float IoR1 = 1.0f, IoR2 = m->index_of_refraction;
    float cosTheta = dot3(neg(*d), n);
    float r0 = ((IoR1 - IoR2) / (IoR1 + IoR2)) * ((IoR1 - IoR2) / (IoR1 + IoR2));
    float coeff = (float)((double)r0 + (double)(1.0f - r0) * pow5_mode(1.0 - (double)cosTheta, tm));
    float random = oracle_uniform(rng, 0.0f, 1.0f);
    if (random < coeff) {
      v3 reflectDir = reflect3(*d, n);
      float spec = powf_mode(glm_max(dot3(neg(*d), reflectDir), 0.0f), 0.0f, tm);
      v3 specColor = ks ? fetch_texel(ks, u, v) : spec_color;
      specColor = muls(specColor, spec);
      *color = mulv(*color, specColor);
      *o = add(intersect, muls(n, 0.01f));
      *d = reflectDir;
    } else {
      v3 diffuseColor = kd ? fetch_texel(kd, u, v) : ld3(m->color);
      *color = mulv(*color, diffuseColor);
      *d = hemisphere(n, rng, tm);
      *o = add(intersect, muls(*d, 0.01f));
    }
  } else {
    *d = hemisphere(n, rng, tm);
    *o = add(intersect, muls(*d, 0.01f));
    *color = mulv(*color, ld3(m->color));
  }
}

/* shadeFakeMaterial, apps/src/pathtrace.cu:397-498. */
int oracle_shade(const B2ptScene* s, const B2ptOptions* opt, int32_t iter, int32_t depth, int32_t n,
                 const float* hit_t, const float* hit_normal, const float* hit_uv,
                 const int32_t* hit_geom, const int32_t* hit_material, float* origin, float* dir,
                 float* color, const int32_t* pixel, int32_t* bounces, float* albedo) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int idx = 0; idx < n; ++idx) {
    const float t = hit_t[idx];
    const size_t i3 = 3 * (size_t)idx;
    /* albedo AOV, :412-462 */
    if (albedo && iter == 1 && depth == 1) {
      float* a = albedo + 3 * (size_t)pixel[idx];
      if (t > 0.0f) {
        const B2ptMaterial* m = &s->materials[hit_material[idx]];
        st3(a, ld3(m->color));
        const B2ptGeom* g = &s->geoms[hit_geom[idx]];
        if (g->type == B2PT_OBJ) {
          const B2ptTexture* ke = obj_tex(s, g, hit_material[idx], 3);
          const B2ptTexture* kd = obj_tex(s, g, hit_material[idx], 0);
          v3 emission = V(0, 0, 0);
          if (ke) emission = fetch_texel(ke, hit_uv[2 * (size_t)idx], hit_uv[2 * (size_t)idx + 1]);
          if (emission.x > FLT_EPSILON || emission.y > FLT_EPSILON || emission.z > FLT_EPSILON) {
            st3(a, muls(emission, 5.0f));
          } else if (kd) {
            st3(a, fetch_texel(kd, hit_uv[2 * (size_t)idx], hit_uv[2 * (size_t)idx + 1]));
          }
        } else if (m->emittance > 0.0f) {
          st3(a, muls(ld3(m->color), m->emittance));
        } else if (m->has_refractive > 0.0f) {
          st3(a, ld3(m->specular_color));
        }
      } else {
        st3(a, V(0, 0, 0));
      }
    }
    if (t > 0.0f) {
      /* apps/src/pathtrace.cu:467 seeds with the array slot (after sort + compaction): B2PT_RNG_SLOT.
       * B2PT_RNG_PIXEL is the second mode of north_star: a counter keyed on (pixel, iteration, depth), which no
       * longer depends on where the path sits in the arrays (no reference behaviour; this restatement defines it). */
      uint32_t rng = opt->rng_mode == B2PT_RNG_PIXEL ? oracle_seed(iter, pixel[idx], depth) : oracle_seed(iter, idx, 0);
      const B2ptMaterial* m = &s->materials[hit_material[idx]];
      if (m->emittance > 0.0f) {
        st3(color + i3, mulv(ld3(color + i3), muls(ld3(m->color), m->emittance)));
        bounces[idx] = 0;
      } else if (bounces[idx] == 1) {
        st3(color + i3, V(0, 0, 0));
        bounces[idx] = 0;
      } else {
        v3 o = ld3(origin + i3), d = ld3(dir + i3), c = ld3(color + i3);
        int32_t b = bounces[idx];
        v3 intersect = add(o, muls(d, t));
        scatter(s, opt, &o, &d, &c, &b, intersect, ld3(hit_normal + i3), hit_uv[2 * (size_t)idx],
                hit_uv[2 * (size_t)idx + 1], hit_geom[idx], m, &rng);
        st3(origin + i3, o);
        st3(dir + i3, d);
        st3(color + i3, c);
        bounces[idx] = b - 1;
      }
    } else {
      st3(color + i3, V(0, 0, 0));
      bounces[idx] = 0;
    }
  }
  return 0;
}

/* stable_partition(isTerminate): remainingBounces > 0 first, both halves stable. */
int oracle_partition_perm(int32_t n, const int32_t* bounces, int32_t* perm) {
  int32_t live = 0;
  for (int i = 0; i < n; ++i) live += bounces[i] > 0;
  int32_t a = 0, b = live;
  for (int i = 0; i < n; ++i) {
    if (bounces[i] > 0) perm[a++] = i; else perm[b++] = i;
  }
  return live;
}

/* finalGather, apps/src/pathtrace.cu:501-510 (PI of :44). */
int oracle_gather(int32_t n, float* image, const float* color, const int32_t* pixel) {
  const float PI_GATHER = 3.14159265358f;
  for (int i = 0; i < n; ++i) {
    float* px = image + 3 * (size_t)pixel[i];
    px[0] += color[3 * (size_t)i] * PI_GATHER;
    px[1] += color[3 * (size_t)i + 1] * PI_GATHER;
    px[2] += color[3 * (size_t)i + 2] * PI_GATHER;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* whole iterations                                                           */
/* ------------------------------------------------------------------------- */
static void permute_f(int32_t n, int w, const int32_t* perm, float* a, float* tmp) {
  for (int k = 0; k < n; ++k) memcpy(tmp + (size_t)k * w, a + (size_t)perm[k] * w, sizeof(float) * w);
  memcpy(a, tmp, sizeof(float) * (size_t)n * w);
}
static void permute_i(int32_t n, const int32_t* perm, int32_t* a, int32_t* tmp) {
  for (int k = 0; k < n; ++k) tmp[k] = a[perm[k]];
  memcpy(a, tmp, sizeof(int32_t) * (size_t)n);
}

int oracle_render(const B2ptScene* s, const B2ptOptions* opt, int32_t iter_first, int32_t count,
                  int32_t stride, float* image, float* albedo, int32_t* n_live, int64_t* segments) {
  const int P = s->camera.resolution[0] * s->camera.resolution[1];
  const int D = s->trace_depth;
  if (P <= 0 || D < 0 || stride <= 0) return B2PT_ERR_INVALID;
  size_t p = (size_t)P;
  float* origin = malloc(p * 12), *dir = malloc(p * 12), *color = malloc(p * 12);
  int32_t* pixel = malloc(p * 4), *bounces = malloc(p * 4);
  float* ht = malloc(p * 4), *hn = malloc(p * 12), *huv = malloc(p * 8);
  int32_t* hg = malloc(p * 4), *hf = malloc(p * 4), *hm = malloc(p * 4);
  int32_t* perm = malloc(p * 4);
  float* tmpf = malloc(p * 12);
  int32_t* tmpi = malloc(p * 4);
  if (!origin || !dir || !color || !pixel || !bounces || !ht || !hn || !huv || !hg || !hf || !hm || !perm ||
      !tmpf || !tmpi)
    return B2PT_ERR_NOMEM;
  int64_t seg = 0;
  for (int it = 0; it < count; ++it) {
    int iter = iter_first + it * stride;
    oracle_generate(&s->camera, opt, iter, D, origin, dir, color, pixel, bounces);
    int n = P, depth = 0;
    if (n_live) {
      for (int k = 0; k <= D; ++k) n_live[k] = 0;
    }
    /* while (!iterationComplete), apps/src/pathtrace.cu:584-652.  The
     * reference only stops when the live count reaches zero. */
    while (n > 0) {
      if (n_live && depth <= D) n_live[depth] = n;
      seg += n;
      oracle_intersect(s, n, origin, dir, ht, hn, huv, hg, hf, hm);
      if (opt->sort_by_material) {
        oracle_sort_perm(n, hm, perm);
        permute_f(n, 1, perm, ht, tmpf);
        permute_f(n, 3, perm, hn, tmpf);
        permute_f(n, 2, perm, huv, tmpf);
        permute_i(n, perm, hg, tmpi);
        permute_i(n, perm, hm, tmpi);
        permute_f(n, 3, perm, origin, tmpf);
        permute_f(n, 3, perm, dir, tmpf);
        permute_f(n, 3, perm, color, tmpf);
        permute_i(n, perm, pixel, tmpi);
        permute_i(n, perm, bounces, tmpi);
      }
      depth++;
      oracle_shade(s, opt, iter, depth, n, ht, hn, huv, hg, hm, origin, dir, color, pixel, bounces, albedo);
      int live = oracle_partition_perm(n, bounces, perm);
      permute_f(n, 3, perm, origin, tmpf);
      permute_f(n, 3, perm, dir, tmpf);
      permute_f(n, 3, perm, color, tmpf);
      permute_i(n, perm, pixel, tmpi);
      permute_i(n, perm, bounces, tmpi);
      n = live;
    }
    oracle_gather(P, image, color, pixel);
  }
  if (segments) *segments = seg;
  free(origin); free(dir); free(color); free(pixel); free(bounces);
  free(ht); free(hn); free(huv); free(hg); free(hf); free(hm); free(perm); free(tmpf); free(tmpi);
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
