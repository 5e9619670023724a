// pt_math.cuh -- device arithmetic of the path-tracing kernels (sm_100a).
//
// The whole library is compiled with -fmad=false: nvcc never contracts a*b+c
// into an FMA, so every expression below rounds exactly as written.  The
// expression trees are the ones glm 0.9.6.3 evaluates in the reference
// (apps/external/include/glm):
//   dot(vec3)   (x*x' + y*y') + z*z'                       detail/func_geometric.inl:65-71
//   cross       (ay*bz - by*az, az*bx - bz*ax, ax*by - bx*ay)           :134-142
//   normalize   v * (1 / sqrt(dot(v,v)))    :154-159, detail/func_exponential.inl:150-153
//   reflect     I - (N*dot(N,I))*2                                      :176-179
//   mat4*vec4   (m0*v0 + m1*v1) + (m2*v2 + m3*v3)          detail/type_mat4x4.inl:617-628
// With IEEE divide and square root (nvcc defaults) this makes hit distances,
// normals and scattered rays bit-identical to a -fmad=false build of the
// reference and to the CPU oracle.  Where speed matters and exactness does
// not (BVH slab tests only prune), fused operations are requested explicitly
// with fmaf().
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace b2pt {

struct V3 {
  float x, y, z;
};

__device__ __forceinline__ V3 mk(float x, float y, float z) {
  V3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 mulv(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 reflect(V3 I, V3 N) { return I - (N * dot(N, I)) * 2.0f; }
__device__ __forceinline__ float glm_min(float x, float y) { return x < y ? x : y; }
__device__ __forceinline__ float glm_max(float x, float y) { return x > y ? x : y; }
__device__ __forceinline__ float glm_abs(float x) { return x >= 0.0f ? x : -x; }

// Rows 0..2 of a column-major mat4, stored as three float4 rows
// (row r = m[r], m[4+r], m[8+r], m[12+r]); vec3(M * vec4(v, w)).
struct Mat34 {
  float4 r0, r1, r2;
};
__device__ __forceinline__ V3 xform(const Mat34& m, V3 v, float w) {
  return mk((m.r0.x * v.x + m.r0.y * v.y) + (m.r0.z * v.z + m.r0.w * w),
            (m.r1.x * v.x + m.r1.y * v.y) + (m.r1.z * v.z + m.r1.w * w),
            (m.r2.x * v.x + m.r2.y * v.y) + (m.r2.z * v.z + m.r2.w * w));
}

// ---- RNG: thrust::minstd_rand + uniform_real_distribution<float> -----------------
// utilhash, apps/src/intersections.h:12-20.
__device__ __forceinline__ uint32_t utilhash(uint32_t a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}
// makeSeededRandomEngine (apps/src/pathtrace.cu:66-70) followed by
// linear_congruential_engine::seed: x = h mod (2^31-1), 0 -> 1.
__device__ __forceinline__ uint32_t rng_seed(int iter, int index, int depth) {
  uint32_t h = utilhash(0x80000000u | ((uint32_t)depth << 22) | (uint32_t)iter) ^ utilhash((uint32_t)index);
  uint32_t x = h % 2147483647u;
  return x == 0u ? 1u : x;
}
// x <- 48271 x mod (2^31-1); u = float(x-1) / 2^31 * (b-a) + a.
// 2^31 == 1 (mod 2^31-1), so the 47-bit product folds with one add and one
// conditional subtract instead of a 64-bit division.
__device__ __forceinline__ float rng_uniform(uint32_t& s, float a, float b) {
  uint64_t p = (uint64_t)s * 48271ull;
  uint32_t f = (uint32_t)(p & 0x7fffffffull) + (uint32_t)(p >> 31);
  s = f >= 2147483647u ? f - 2147483647u : f;
  float r = (float)(s - 1u);
  r = r / 2147483648.0f;
  return (r * (b - a)) + a;
}

// ---- trig ---------------------------------------------------------------------------
// PORTABLE: the +,-,* polynomial shared bit-for-bit with oracle/pt_oracle.c.
__device__ __forceinline__ void sincos_portable(float x, float* s, float* c) {
  float kf = rintf(x * 0.636619772f);
  int k = (int)kf;
  float r = x - kf * 1.5703125f;
  r = r - kf * 4.837512969970703125e-4f;
  r = r - kf * 7.54978995489188216e-8f;
  float z = r * r;
  float sp = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
  float cp = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
  switch (k & 3) {
    case 0: *s = sp; *c = cp; break;
    case 1: *s = cp; *c = -sp; break;
    case 2: *s = -sp; *c = -cp; break;
    default: *s = -cp; *c = sp; break;
  }
}
// NATIVE: libdevice sinf/cosf, what `cos(around)` / `sin(around)` compile to
// in the reference's device code (apps/src/interactions.h:42-43).
template <int TRIG>
__device__ __forceinline__ void sincos_mode(float x, float* s, float* c) {
  if (TRIG == 1) {
    sincos_portable(x, s, c);
  } else {
    *s = sinf(x);
    *c = cosf(x);
  }
}
// pow((1.0 - cosTheta), 5) in double (apps/src/interactions.h:153,192).
template <int TRIG>
__device__ __forceinline__ double pow5_mode(double a) {
  if (TRIG == 1) {
    double a2 = a * a;
    return a2 * a2 * a;
  }
  return pow(a, 5);
}
__device__ __forceinline__ float powf_exponent(float x, float e) {
  if (e == 0.0f) return 1.0f;  // powf(x, 0) == 1 for every x, NaN included
  return powf(x, e);
}

}  // namespace b2pt
