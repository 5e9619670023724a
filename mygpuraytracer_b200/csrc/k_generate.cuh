// k_generate.cuh -- camera rays (generateRayFromCamera, apps/src/pathtrace.cu:248-297).
#pragma once

#include "pt_device.cuh"

namespace b2pt {

struct GenParams {
  DevCamera cam;
  int trace_depth;
  int antialiasing;
  int depth_of_field;
  float lens_radius;
  float focal_distance;
};

// Start-of-iteration bookkeeping: advance the iteration number kept in device
// memory (so a captured CUDA graph replays without parameter updates) and zero
// the per-iteration counters.
__global__ void k_iter_begin(Counters* c, int* iter_state, int n_paths, int depth_slots) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  if (tid == 0) {
    iter_state[0] = iter_state[1];       // current iteration
    iter_state[1] += iter_state[2];      // next = current + stride
    c->serial += 1;
    c->n_live[0] = n_paths;
  }
  for (int i = tid; i < depth_slots; i += nth) {
    if (i > 0) c->n_live[i] = 0;
    c->ray_ticket[i] = 0;
    c->mesh_count[i] = 0;
    c->long_count[i] = 0;
    c->long_ticket[i] = 0;
    c->sort_ticket[i] = 0;
  }
  if (tid == 0) c->n_live[depth_slots] = 0;
  unsigned int* h = &c->hist[0][0];
  unsigned int* hl = &c->hist_live[0][0];
  for (int i = tid; i < depth_slots * kMaxMaterials; i += nth) {
    h[i] = 0;
    hl[i] = 0;
  }
}

// ConcentricSampleDisk, apps/src/pathtrace.cu:225-239.
template <int TRIG>
__device__ __forceinline__ void concentric_disk(float px, float py, float* ox, float* oy) {
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
  sincos_mode<TRIG>(theta, &sn, &cs);
  *ox = r * cs;
  *oy = r * sn;
}

// One thread per pixel; a warp covers 32 consecutive pixels of a row, so the
// three 128-bit stores per path are fully coalesced (the reference writes a
// 44-byte AoS element from an 8x8 block).
// The primary ray of pixel `index` (generateRayFromCamera, pathtrace.cu:248-297).
template <int TRIG>
__device__ __forceinline__ void camera_ray(const GenParams& gp, int iter, int index, V3* origin, V3* direction) {
  {
    const int x = index % gp.cam.res_x;
    const int y = index / gp.cam.res_x;
    uint32_t rng = rng_seed(iter, index, gp.trace_depth);
    V3 o = gp.cam.position;
    float ax = (float)x, ay = (float)y;
    if (gp.antialiasing) {
      uint32_t rng_aa = rng_seed(iter, index, gp.trace_depth);
      ax += rng_uniform(rng_aa, -0.5f, 0.5f);
      ay += rng_uniform(rng_aa, -0.5f, 0.5f);
    }
    V3 d = normalize(gp.cam.view - (gp.cam.right * gp.cam.plx) * (ax - (float)gp.cam.res_x * 0.5f) -
                     (gp.cam.up * gp.cam.ply) * (ay - (float)gp.cam.res_y * 0.5f));
    if (gp.depth_of_field && gp.lens_radius > 0) {
      // glm::vec2(uDOF(rng), uDOF(rng)), pathtrace.cu:285: device code draws
      // the arguments left to right.
      float u0 = rng_uniform(rng, 0.0f, 1.0f);
      float u1 = rng_uniform(rng, 0.0f, 1.0f);
      float lx, ly;
      concentric_disk<TRIG>(u0, u1, &lx, &ly);
      lx = gp.lens_radius * lx;
      ly = gp.lens_radius * ly;
      float ft = glm_abs(gp.focal_distance / d.z);
      V3 focus = o + d * ft;
      o = o + mk(lx, ly, 0.0f);
      d = normalize(focus - o);
    }
    *origin = o;
    *direction = d;
  }
}

template <int TRIG>
__global__ void __launch_bounds__(256) k_generate(GenParams gp, const int* __restrict__ iter_state, PathBuf out) {
  const int P = gp.cam.res_x * gp.cam.res_y;
  const int iter = iter_state[0];
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < P; index += gridDim.x * blockDim.x) {
    V3 o, d;
    camera_ray<TRIG>(gp, iter, index, &o, &d);
    out.s0[index] = make_float4(o.x, o.y, o.z, __int_as_float(index));
    out.s1[index] = make_float4(d.x, d.y, d.z, __int_as_float(gp.trace_depth));
    out.s2[index] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
  }
}

}  // namespace b2pt
