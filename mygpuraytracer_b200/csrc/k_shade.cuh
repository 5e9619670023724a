// k_shade.cuh -- shade + scatter + stream compaction + gather, one kernel.
//
// Reference stages fused here:
//   shadeFakeMaterial      apps/src/pathtrace.cu:397-498
//   scatterRay             apps/src/interactions.h:112-258
//   stable_partition       apps/src/pathtrace.cu:518-522,649   (isTerminate)
//   finalGather            apps/src/pathtrace.cu:501-510
//
// A thread shades sorted slot j: it gathers the hit record and the path state
// through the sort permutation (the only pass that touches the payload of the
// sort), seeds the reference's RNG with (iter, j, 0) and scatters.  Survivors
// are written to the other path buffer at their rank among the survivors
// (computed by the kernel that ranked the survival flags: k_sort_material, or
// k_rank_live when the material sort is off), which is exactly the live prefix
// thrust::stable_partition produces.  A path that dies adds color*PI to its
// pixel at once: every pixel owns exactly one path per iteration and dies
// exactly once, so the accumulation needs no atomics and adds the same values
// in the same order as the reference's end-of-iteration finalGather.
#pragma once

#include <float.h>

#include "k_intersect.cuh"  // fetch_texel
#include "pt_device.cuh"

namespace b2pt {

constexpr int kShadeThreads = 256;
constexpr int kShadeWarps = kShadeThreads / 32;

struct ShadeParams {
  DevScene scene;
  PathBuf in;
  PathBuf out;
  HitBuf hits;
  const int* perm;  // NULL: identity (SORT_BY_MATERIAL 0)
  const int* apos;  // survivors in front of each sorted slot (k_sort_material / k_rank_live)
  Counters* ctr;
  float* image;
  float* albedo;
  const int* iter_state;
  int depth;      // 0-based loop index; the reference's `depth` argument is depth+1
  int rng_pixel;  // 0: seed with the slot (reference); 1: seed with (pixel, depth)
  // stage recording (parity tests), all NULL otherwise
  float4* rec_s0;
  float4* rec_s1;
  float4* rec_s2;
  int* rec_dead;  // pixelIndex of dead paths in stable order
  int* rec_live;  // pixelIndex of live paths in stable order
  const uint8_t* live;  // the prediction of k_intersect (checked in record mode)
};

// calculateRandomDirectionInHemisphere, apps/src/interactions.h:12-44.
template <int TRIG>
__device__ __forceinline__ V3 hemisphere(V3 normal, uint32_t& rng) {
  const float up = sqrtf(rng_uniform(rng, 0.0f, 1.0f));
  const float over = sqrtf(1 - up * up);
  const float around = rng_uniform(rng, 0.0f, 1.0f) * 6.2831853071795864769252867665590057683943f;
  const float kSqrtOneThird = 0.5773502691896257645091487805019574556476f;
  V3 dnn;
  if (fabsf(normal.x) < kSqrtOneThird) {
    dnn = mk(1, 0, 0);
  } else if (fabsf(normal.y) < kSqrtOneThird) {
    dnn = mk(0, 1, 0);
  } else {
    dnn = mk(0, 0, 1);
  }
  const V3 p1 = normalize(cross(normal, dnn));
  const V3 p2 = normalize(cross(normal, p1));
  float sn, cs;
  sincos_mode<TRIG>(around, &sn, &cs);
  return (normal * up + p1 * (cs * over)) + p2 * (sn * over);
}

// Survivors go to slot `pos` of the next depth's path buffer; a path that dies
// is gathered: image[pixelIndex] += color * PI (finalGather, pathtrace.cu:501-510,
// PI of :44).  Adding an exact zero leaves the non-negative accumulator unchanged.
template <bool RECORD>
__device__ __forceinline__ void write_result(const ShadeParams& p, int j, unsigned int pos, bool alive, V3 o, V3 d, V3 col,
                                             int pixel, int bounces) {
  if (alive) {
    p.out.s0[pos] = make_float4(o.x, o.y, o.z, __int_as_float(pixel));
    p.out.s1[pos] = make_float4(d.x, d.y, d.z, __int_as_float(bounces));
    p.out.s2[pos] = make_float4(col.x, col.y, col.z, 0.0f);
    if (RECORD) p.rec_live[pos] = pixel;
  } else {
    if (col.x != 0.0f || col.y != 0.0f || col.z != 0.0f) {
      float* px = p.image + 3 * (size_t)pixel;
      px[0] += col.x * 3.14159265358f;
      px[1] += col.y * 3.14159265358f;
      px[2] += col.z * 3.14159265358f;
    }
    if (RECORD) p.rec_dead[j - (int)pos] = pixel;
  }
}

// Shade sorted slot j (shadeFakeMaterial + scatterRay, pathtrace.cu:397-498, interactions.h:112-258): gathers
// the hit record and the path state through the sort permutation, returns the scattered path.
template <int TRIG>
__device__ __forceinline__ void shade_slot(const ShadeParams& p, int j, int iter, int ref_depth, V3& o, V3& d, V3& col,
                                           int& pixel, int& bounces) {
  const int i = p.perm ? p.perm[j] : j;
  const float4 h0 = p.hits.h0[i];
  const float4 h1 = p.hits.h1[i];
  const float4 s0 = p.in.s0[i], s1 = p.in.s1[i], s2 = p.in.s2[i];
  o = mk(s0.x, s0.y, s0.z);
  d = mk(s1.x, s1.y, s1.z);
  col = mk(s2.x, s2.y, s2.z);
  pixel = __float_as_int(s0.w);
  bounces = __float_as_int(s1.w);
  const float t = h0.x;
  const int gm = __float_as_int(h1.z);
  const int geom_id = gm & 0xffff;
  const int mat_id = (gm >> 16) & 0xffff;
  const float tu = h1.x, tv = h1.y;

  // albedo AOV, pathtrace.cu:412-462 (iteration 1, first shade only)
  if (p.albedo != nullptr && iter == 1 && ref_depth == 1) {
    V3 a = mk(0, 0, 0);
    if (t > 0.0f) {
      const DevMaterial& m = p.scene.materials[mat_id];
      a = mk(m.color[0], m.color[1], m.color[2]);
      const DevGeom& G = p.scene.geoms[geom_id];
      if (G.type == 3) {
        const DevMesh& M = p.scene.meshes[G.mesh];
        const DevTexture& ke = obj_tex(p.scene, M, mat_id, 3);
        const DevTexture& kd = obj_tex(p.scene, M, mat_id, 0);
        V3 emission = mk(0, 0, 0);
        if (ke.channels) emission = fetch_texel(ke, tu, tv);
        if (emission.x > FLT_EPSILON || emission.y > FLT_EPSILON || emission.z > FLT_EPSILON) {
          a = emission * 5.0f;
        } else if (kd.channels) {
          a = fetch_texel(kd, tu, tv);
        }
      } else if (m.emittance > 0.0f) {
        a = a * m.emittance;
      } else if (m.has_refractive > 0.0f) {
        a = mk(m.specular_color[0], m.specular_color[1], m.specular_color[2]);
      }
    }
    float* ap = p.albedo + 3 * (size_t)pixel;
    ap[0] = a.x;
    ap[1] = a.y;
    ap[2] = a.z;
  }

  if (t > 0.0f) {
    uint32_t rng = p.rng_pixel ? rng_seed(iter, pixel, ref_depth) : rng_seed(iter, j, 0);
    const DevMaterial m = p.scene.materials[mat_id];
    const V3 mcol = mk(m.color[0], m.color[1], m.color[2]);
    const V3 scol = mk(m.specular_color[0], m.specular_color[1], m.specular_color[2]);
    if (m.emittance > 0.0f) {
      col = mulv(col, mcol * m.emittance);
      bounces = 0;
    } else if (bounces == 1) {
      col = mk(0, 0, 0);
      bounces = 0;
    } else {
      // scatterRay, interactions.h:112-258
      const V3 x = o + d * t;
      V3 nrm = mk(h0.y, h0.z, h0.w);
      bool done = false;
      if (m.has_reflective > 0) {
        const V3 rdir = reflect(d, nrm);
        const float spec = powf_exponent(glm_max(dot(-d, rdir), 0.0f), m.specular_exponent);
        col = mulv(col, scol * (m.has_reflective * spec));
        o = x + nrm * 0.01f;
        d = rdir;
      } else if (m.has_refractive > 0) {
        float ior1 = 1.0f, ior2 = m.ior;
        float cos_t = dot(-d, nrm);
        if (cos_t < 0) {
          nrm = nrm * -1.0f;
          ior1 = ior2;
          ior2 = 1.0f;
          cos_t = fabsf(cos_t);
        }
        const float sin_t = (float)sqrt(1.0 - (double)(cos_t * cos_t));
        if (ior1 / ior2 * sin_t > 1.0f) {
          d = reflect(d, nrm);
        } else {
          const float r0 = ((ior1 - ior2) / (ior1 + ior2)) * ((ior1 - ior2) / (ior1 + ior2));
          const float coeff = (float)((double)r0 + (double)(1.0f - r0) * pow5_mode<TRIG>(1.0 - (double)cos_t));
          const float rnd = rng_uniform(rng, 0.0f, 1.0f);
          if (rnd < coeff) {
            d = reflect(d, nrm);
          } else {
            // glm::refract, detail/func_geometric.inl:193-200
            const float eta = ior1 / ior2;
            const float dv = dot(nrm, d);
            const float k = 1.0f - eta * eta * (1.0f - dv * dv);
            d = (d * eta - nrm * (eta * dv + sqrtf(k))) * (k >= 0.0f ? 1.0f : 0.0f);
          }
        }
        col = mulv(col, scol);
        o = x + d * 0.01f;
      } else {
        const DevGeom& G = p.scene.geoms[geom_id];
        if (G.type == 3) {
          const DevMesh& M = p.scene.meshes[G.mesh];
          const DevTexture& ke = obj_tex(p.scene, M, mat_id, 3);
          V3 emission = mk(0, 0, 0);
          if (ke.channels) emission = fetch_texel(ke, tu, tv);
          if (emission.x > FLT_EPSILON || emission.y > FLT_EPSILON || emission.z > FLT_EPSILON) {
            col = mulv(col, emission * 5.0f);
            bounces = 1;  // interactions.h:184; decremented to 0 below
            done = true;
          }
          if (!done) {
            const float ior1 = 1.0f, ior2 = m.ior;
            const float cos_t = dot(-d, nrm);
            const float r0 = ((ior1 - ior2) / (ior1 + ior2)) * ((ior1 - ior2) / (ior1 + ior2));
            const float coeff = (float)((double)r0 + (double)(1.0f - r0) * pow5_mode<TRIG>(1.0 - (double)cos_t));
            const float rnd = rng_uniform(rng, 0.0f, 1.0f);
            if (rnd < coeff) {
              const V3 rdir = reflect(d, nrm);
              const DevTexture& ks = obj_tex(p.scene, M, mat_id, 1);
              V3 sc = ks.channels ? fetch_texel(ks, tu, tv) : scol;
              sc = sc * 1.0f;  // spec = pow(x, 0.0f) == 1, interactions.h:204,214
              col = mulv(col, sc);
              o = x + nrm * 0.01f;
              d = rdir;
            } else {
              const DevTexture& kd = obj_tex(p.scene, M, mat_id, 0);
              const V3 dc = kd.channels ? fetch_texel(kd, tu, tv) : mcol;
              col = mulv(col, dc);
              d = hemisphere<TRIG>(nrm, rng);
              o = x + d * 0.01f;
            }
          }
        } else {
          d = hemisphere<TRIG>(nrm, rng);
          o = x + d * 0.01f;
          col = mulv(col, mcol);
        }
      }
      bounces -= 1;
    }
  } else {
    col = mk(0, 0, 0);
    bounces = 0;
  }
}

// The compaction ranks come from k_sort_material or k_rank_live (apos), so the kernel is embarrassingly parallel:
// no shared memory, no barrier, no look-back; the CTAs stride over 256-slot tiles.
template <int TRIG, bool RECORD>
__global__ void __launch_bounds__(kShadeThreads) k_shade_compact(ShadeParams p) {
  const int tid = threadIdx.x;
  const int n = p.ctr->n_live[p.depth];
  const int iter = p.iter_state[0];
  const int ref_depth = p.depth + 1;
  for (unsigned int tile = blockIdx.x; (long long)tile * kShadeThreads < (long long)n; tile += gridDim.x) {
    const int j = (int)tile * kShadeThreads + tid;
    if (j >= n) continue;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 0), col = mk(0, 0, 0);
    int pixel = 0, bounces = 0;
    shade_slot<TRIG>(p, j, iter, ref_depth, o, d, col, pixel, bounces);
    const bool alive = bounces > 0;
    if (RECORD) {
      p.rec_s0[j] = make_float4(o.x, o.y, o.z, __int_as_float(pixel));
      p.rec_s1[j] = make_float4(d.x, d.y, d.z, __int_as_float(bounces));
      p.rec_s2[j] = make_float4(col.x, col.y, col.z, 0.0f);
    }
    // stable compaction: the rank among the survivors was computed when the flags were
    write_result<RECORD>(p, j, (unsigned int)p.apos[j], alive, o, d, col, pixel, bounces);
    if (RECORD && alive != (p.live[p.perm ? p.perm[j] : j] != 0)) atomicAdd(&p.ctr->pred_mismatch, 1u);
  }
}

}  // namespace b2pt
