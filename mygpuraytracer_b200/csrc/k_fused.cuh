// k_fused.cuh -- the analytic intersection of depth d+1 fused into the kernel that produces its rays.
//
// A wavefront iteration is a chain of short dependent kernels; k_intersect_analytic is bound by instruction
// issue, k_shade_compact by the latency of its gathers.  Here the thread that scatters a path (or generates a
// camera ray) also finds its closest cube / sphere at once: the ray never goes back to HBM to be read again,
// and eight launches per iteration disappear.  Measured: one context 1.435 -> 1.39 ms per iteration (the two
// phases of a tile serialise inside a CTA, so the intersection arithmetic only partly fills the issue slots the
// shade phase leaves idle); with several contexts sharing the GPU the unfused kernels are 2 % better and stay
// in use (b2pt.cu).
//
//  k_generate_trace   generateRayFromCamera + computeIntersections of depth 0 (analytic geoms)
//  k_shade_trace      shadeFakeMaterial + scatterRay + stable_partition + finalGather of depth d, then
//                     computeIntersections of depth d+1 (analytic geoms) for the survivors.  The survivors
//                     of a 256-slot tile own consecutive output slots (their compaction ranks come from
//                     k_sort_material), so they are staged in shared memory and re-distributed: thread k
//                     takes survivor k, which keeps the second phase at full lane utilisation and makes the
//                     stores of the path state coalesced.
// Both write exactly what the unfused kernels write (same device functions, same order of operations per
// path), so results are bit-identical; the unfused kernels stay in use where the records are read back
// between the stages (record_stages), without the material sort, and to fill the first-bounce cache.
#pragma once

#include "k_generate.cuh"
#include "k_intersect.cuh"
#include "k_shade.cuh"

namespace b2pt {

template <int TRIG>
__global__ void __launch_bounds__(256) k_generate_trace(GenParams gp, const int* __restrict__ iter_state, PathBuf out,
                                                        IsectParams next) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  const int tid = threadIdx.x, lane = tid & 31;
  analytic_stage(next.scene, sgeom, shist, slive);
  __syncthreads();
  const int P = gp.cam.res_x * gp.cam.res_y;
  const int iter = iter_state[0];
  for (int base = (blockIdx.x * blockDim.x + tid) & ~31; base < P; base += gridDim.x * blockDim.x) {
    const int index = base + lane;
    const bool valid = index < P;
    AnalyticHit r;
    r.mat = 0;
    r.survives = r.want_mesh = false;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    if (valid) {
      camera_ray<TRIG>(gp, iter, index, &o, &d);
      out.s0[index] = make_float4(o.x, o.y, o.z, __int_as_float(index));
      out.s1[index] = make_float4(d.x, d.y, d.z, __int_as_float(gp.trace_depth));
      out.s2[index] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
      analytic_trace(sgeom, next.scene.n_geoms, next.scene.materials, o, d, gp.trace_depth, &r);
    }
    analytic_commit(next, shist, slive, index, valid, r, o, d);
  }
  __syncthreads();
  analytic_flush(next, shist, slive);
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&next.ctr->segments, (unsigned long long)P);
}

// Needs the compaction ranks of k_sort_material (p.apos): a tile's survivors then own the consecutive output
// slots apos[first slot of the tile] ... + count - 1.
constexpr int kFusedThreads = 256;  // slots per tile of k_shade_trace (128 measured the same)

template <int TRIG>
__global__ void __launch_bounds__(kFusedThreads) k_shade_trace(ShadeParams p, IsectParams next) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  __shared__ float4 st0[kFusedThreads], st1[kFusedThreads], st2[kFusedThreads];  // survivors of the tile, by rank
  const int tid = threadIdx.x;
  analytic_stage(next.scene, sgeom, shist, slive);
  __syncthreads();
  const int n = p.ctr->n_live[p.depth];
  const int iter = p.iter_state[0];
  const int ref_depth = p.depth + 1;
  for (unsigned int tile = blockIdx.x; (long long)tile * kFusedThreads < (long long)n; tile += gridDim.x) {
    const int j0 = (int)tile * kFusedThreads;
    const int j = j0 + tid;
    const bool valid = j < n;
    const unsigned int pos0 = (unsigned int)p.apos[j0];  // survivors in front of the tile
    bool alive = false;
    if (valid) {
      V3 o = mk(0, 0, 0), d = mk(0, 0, 0), col = mk(0, 0, 0);
      int pixel = 0, bounces = 0;
      shade_slot<TRIG>(p, j, iter, ref_depth, o, d, col, pixel, bounces);
      alive = bounces > 0;
      if (alive) {
        const unsigned int k = (unsigned int)p.apos[j] - pos0;
        st0[k] = make_float4(o.x, o.y, o.z, __int_as_float(pixel));
        st1[k] = make_float4(d.x, d.y, d.z, __int_as_float(bounces));
        st2[k] = make_float4(col.x, col.y, col.z, 0.0f);
      } else if (col.x != 0.0f || col.y != 0.0f || col.z != 0.0f) {
        // finalGather (pathtrace.cu:501-510): the path dies here, and only here, in this iteration
        float* px = p.image + 3 * (size_t)pixel;
        px[0] += col.x * 3.14159265358f;
        px[1] += col.y * 3.14159265358f;
        px[2] += col.z * 3.14159265358f;
      }
    }
    const int count = __syncthreads_count(alive);
    // ---- second phase: thread k owns survivor k of the tile ----
    const bool have = tid < count;
    AnalyticHit r;
    r.mat = 0;
    r.survives = r.want_mesh = false;
    const int slot = (int)pos0 + tid;
    V3 no = mk(0, 0, 0), nd = mk(0, 0, 1);
    if (have) {
      const float4 a = st0[tid], b = st1[tid], c = st2[tid];
      p.out.s0[slot] = a;
      p.out.s1[slot] = b;
      p.out.s2[slot] = c;
      no = mk(a.x, a.y, a.z);
      nd = mk(b.x, b.y, b.z);
      analytic_trace(sgeom, next.scene.n_geoms, next.scene.materials, no, nd, __float_as_int(b.w), &r);
    }
    analytic_commit(next, shist, slive, slot, have, r, no, nd);
    __syncthreads();  // the staging arrays are reused by the next tile
  }
  __syncthreads();
  analytic_flush(next, shist, slive);
  // every survivor of this depth is a segment of the next one
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&next.ctr->segments, (unsigned long long)p.ctr->n_live[p.depth + 1]);
}

}  // namespace b2pt
