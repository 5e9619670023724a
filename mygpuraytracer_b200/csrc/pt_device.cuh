// pt_device.cuh -- device-side data layout of the wavefront loop.
//
// HBM layout (P = W*H paths, all arrays 16-byte aligned, read and written with
// 128-bit accesses):
//
//   path state, SoA, double buffered (A/B)          48 B/path
//     s0[i] = (origin.xyz, pixelIndex)      PathSegment::ray.origin / pixelIndex
//     s1[i] = (direction.xyz, remainingBounces)
//     s2[i] = (color.rgb, unused)           PathSegment::color (throughput)
//   hit records, SoA                                 32 B/path
//     h0[i] = (t, normal.xyz)               ShadeableIntersection::t / surfaceNormal
//     h1[i] = (u, v, geom | material<<16, face)
//   sort keys          uint8 material id              1 B/path
//   sort permutation   int32                          4 B/path
//
// The reference's AoS structs are apps/src/sceneStructs.h:105-121 (44-byte
// PathSegment, 32-byte ShadeableIntersection used as the sort key).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_math.cuh"

namespace b2pt {

constexpr int kMaxMaterials = 256;  // sort key is one byte
constexpr int kMaxGeoms = 64;       // analytic geoms + meshes, staged in shared memory
constexpr int kMaxDepth = 62;

struct PathBuf {
  float4* s0;
  float4* s1;
  float4* s2;
};

struct HitBuf {
  float4* h0;
  float4* h1;
};

struct DevTexture {
  const uint8_t* texels;
  int w, h, channels;
};

// One geom as the kernels read it (Geom, apps/src/sceneStructs.h:50-70).
struct DevGeom {
  Mat34 inv;   // inverseTransform
  Mat34 fwd;   // transform
  Mat34 invT;  // invTranspose
  int type;
  int material;
  int mesh;   // index into DevScene::meshes, -1 for analytic geoms
  int rigid;  // 1 if object-space and world-space distances agree (scale == 1)
  float4 wmin;  // padded world-space bounding box; .w = distance slack of the exact test
  float4 wmax;
};

struct DevMesh {
  const float4* nodes;    // 8 x float4 (one 128-byte line) per 4-wide BVH node, child-major (see k_lbvh.cuh)
  const float4* tris;     // 3 x float4 per triangle in BVH leaf (Morton) order: (v0,face) (v1,-) (v2,-)
  const float* face_pos;  // 9 floats per face, original order (the LBVH build and the brute-force validation kernel)
  const float* face_uv;   // 6 floats per face, original order
  const float4* face_rec; // the same 15 floats per face in ONE aligned 64-byte record (v0 v1 v2 | uv0 uv1 uv2 | pad): what
                          // the winner of a walk gathers -- two sectors and four 128-bit loads instead of 15 scalar loads
                          // over up to five sectors
  int n_faces;
  int root;  // 0: the root node; < 0: a leaf code (a mesh of at most kLeafTris triangles has no nodes)
  int geom;  // the geom this mesh belongs to (meshes are numbered in geom order)
  int pad_;
  const int* face_mat;  // scene material of every face (B2ptScene::face_material), or NULL: the geom's material
  DevTexture kd, ks, bump, ke;  // the geom's four maps, in this order (obj_tex indexes them)
};

struct DevMaterial {  // Material, apps/src/sceneStructs.h:72-82
  float color[3];
  float specular_exponent;
  float specular_color[3];
  float has_reflective;
  float has_refractive;
  float ior;
  float emittance;
};

struct DevCamera {  // Camera, apps/src/sceneStructs.h:84-93
  int res_x, res_y;
  V3 position, view, up, right;
  float plx, ply;
};

struct DevScene {
  const DevGeom* geoms;
  const DevMesh* meshes;
  const DevMaterial* materials;
  const DevTexture* mat_tex;  // [4 * n_materials] (kd, ks, bump, ke) per material (B2ptScene::material_textures), or NULL
  int n_geoms, n_meshes, n_materials;
};

// Per-iteration device counters; zeroed by k_iter_begin.  Nothing in the
// depth loop is read back by the host: every kernel takes its element count
// from here (the reference syncs three times per depth).
struct Counters {
  unsigned int serial;                    // iterations started (lookback epochs)
  int n_live[kMaxDepth + 2];              // paths entering depth d
  unsigned int ray_ticket[kMaxDepth + 1];    // head of the mesh-ray queue (k_intersect_mesh)
  unsigned int mesh_count[kMaxDepth + 1];    // tail of the mesh-ray queue (k_intersect_analytic)
  unsigned int long_count[kMaxDepth + 1];    // tail of the long-walk queue (k_mesh_walk)
  unsigned int long_ticket[kMaxDepth + 1];   // head of the long-walk queue (k_mesh_walk_long)
  unsigned int sort_ticket[kMaxDepth + 1];   // tile order of the sort
  unsigned int hist[kMaxDepth + 1][kMaxMaterials];       // material histogram per depth
  unsigned int hist_live[kMaxDepth + 1][kMaxMaterials];  // ... of the paths that will survive the shade
  unsigned int pred_mismatch;                            // record mode: shade disagreed with the prediction
  unsigned long long segments;            // total path segments since creation
};

// Decoupled look-back status word: epoch in the high half so the arrays are
// never cleared; flag 1 = tile aggregate, 2 = inclusive prefix.
__device__ __forceinline__ unsigned long long lb_pack(unsigned int epoch, unsigned int flag, unsigned int value) {
  return ((unsigned long long)epoch << 32) | ((unsigned long long)flag << 30) | (unsigned long long)value;
}
__device__ __forceinline__ unsigned int lb_epoch(unsigned long long w) { return (unsigned int)(w >> 32); }
__device__ __forceinline__ unsigned int lb_flag(unsigned long long w) { return (unsigned int)(w >> 30) & 3u; }
__device__ __forceinline__ unsigned int lb_value(unsigned long long w) { return (unsigned int)w & 0x3fffffffu; }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Decoupled look-back by a FULL WARP: publishes `total` as the aggregate of
// `tile`, inspects 32 predecessors per step (each lane spins on its own status
// word), sums the aggregates down to the nearest inclusive prefix, publishes
// the tile's inclusive prefix and returns its exclusive prefix (to all lanes).
// `status` points at this scan's array; consecutive tiles are `stride` words apart.
__device__ __forceinline__ unsigned int lookback_warp(unsigned long long* status, size_t stride, unsigned int tile,
                                                      unsigned int epoch, unsigned int total) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) st_volatile_u64(status + (size_t)tile * stride, lb_pack(epoch, tile == 0 ? 2u : 1u, total));
  unsigned int excl = 0;
  for (int base = (int)tile - 1; base >= 0; base -= 32) {
    const int t = base - lane;
    unsigned int flag = 2u, value = 0u;  // lanes before tile 0 act as an empty inclusive prefix
    if (t >= 0) {
      unsigned long long w;
      do {
        w = ld_volatile_u64(status + (size_t)t * stride);
      } while (lb_epoch(w) != epoch || lb_flag(w) == 0u);
      flag = lb_flag(w);
      value = lb_value(w);
    }
    const unsigned int incl = __ballot_sync(0xffffffffu, flag == 2u);
    const int first = incl ? __ffs(incl) - 1 : 32;  // nearest inclusive prefix among these 32
    unsigned int v = lane <= first ? value : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if (incl) break;
  }
  if (tile > 0 && lane == 0) st_volatile_u64(status + (size_t)tile * stride, lb_pack(epoch, 2u, excl + total));
  return excl;
}

}  // namespace b2pt
