// k_lbvh.cuh -- on-GPU LBVH build for OBJ geoms.
//
// The reference has no acceleration structure: meshIntersectionTest walks every
// face for every ray (apps/src/intersections.h:216-230, "//TODO BVH" at
// apps/src/pathtrace.cu:331).  This builds, per mesh and entirely on the
// device:
//   1. centroid bounds (block reduction + ordered-int atomics)
//   2. 30-bit Morton codes of the triangle centroids
//   3. LSD radix sort of (code, face) pairs -- four onesweep passes (k_sort.cuh)
//   4. the Karras 2012 hierarchy (one thread per internal node, duplicates
//      broken by index)
//   5. bottom-up refit with one atomic counter per internal node
//   6. 4-wide traversal nodes with fat leaves (k_emit_wide4): every maximal subtree of at most
//      kLeafTris triangles becomes ONE leaf (a contiguous range of the Morton order: Karras'
//      nodes cover index ranges); a wide node takes a binary node's two children and opens the
//      inner one with the largest surface area twice (surface-area greedy).  One node is one
//      128-byte line with the boxes stored per axis.  Only the nodes reachable from the root are
//      kept (k_wide_levels marks them level by level, a scan numbers them, k_compact_nodes moves
//      them), so the tree the walk touches is contiguous: 250 500 triangles -> 71 k nodes = 9 MB
//      with kLeafTris = 2, instead of one node per binary node = 32 MB.  Triangles stay in leaf
//      order as 3 x float4 (v0|face id, v1, v2); the depth of the wide tree bounds the traversal
//      stacks.
// Boxes are inflated by `pad` (a few 1e-6 of the mesh extent) so that the
// conservative slab test can never cull a triangle the reference's exact test
// would accept: the BVH prunes, it never decides.
#pragma once

#include <float.h>

#include "pt_device.cuh"

namespace b2pt {

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned int o) {
  const unsigned int b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f;
  memcpy(&f, &b, 4);
  return f;
#endif
}

struct TriBounds {
  unsigned int lo[3];  // ordered-int min of centroids
  unsigned int hi[3];
  int max_depth;
  int wide_depth;  // inner 4-wide nodes on the longest root-to-leaf path (k_wide_levels)
};

__global__ void k_bounds_init(TriBounds* b) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int k = 0; k < 3; ++k) {
      b->lo[k] = 0xffffffffu;
      b->hi[k] = 0u;
    }
    b->max_depth = 0;
    b->wide_depth = 0;
  }
}

__device__ __forceinline__ V3 tri_centroid(const float* fp) {
  const float minx = fminf(fminf(fp[0], fp[3]), fp[6]), maxx = fmaxf(fmaxf(fp[0], fp[3]), fp[6]);
  const float miny = fminf(fminf(fp[1], fp[4]), fp[7]), maxy = fmaxf(fmaxf(fp[1], fp[4]), fp[7]);
  const float minz = fminf(fminf(fp[2], fp[5]), fp[8]), maxz = fmaxf(fmaxf(fp[2], fp[5]), fp[8]);
  return mk(0.5f * (minx + maxx), 0.5f * (miny + maxy), 0.5f * (minz + maxz));
}

__global__ void __launch_bounds__(256) k_centroid_bounds(const float* __restrict__ face_pos, int n, TriBounds* out) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float fp[9];
    for (int k = 0; k < 9; ++k) fp[k] = face_pos[9 * (size_t)i + k];
    const V3 c = tri_centroid(fp);
    lo[0] = fminf(lo[0], c.x); lo[1] = fminf(lo[1], c.y); lo[2] = fminf(lo[2], c.z);
    hi[0] = fmaxf(hi[0], c.x); hi[1] = fmaxf(hi[1], c.y); hi[2] = fmaxf(hi[2], c.z);
  }
  for (int k = 0; k < 3; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 3; ++k) {
      atomicMin(&out->lo[k], f2ord(lo[k]));
      atomicMax(&out->hi[k], f2ord(hi[k]));
    }
  }
}

#ifndef B2PT_MORTON_UNIFORM
#define B2PT_MORTON_UNIFORM 1
#endif
__device__ __forceinline__ unsigned int expand10(unsigned int v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

__global__ void __launch_bounds__(256) k_morton(const float* __restrict__ face_pos, int n, const TriBounds* b,
                                                uint32_t* code, uint32_t* val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float fp[9];
  for (int k = 0; k < 9; ++k) fp[k] = face_pos[9 * (size_t)i + k];
  const V3 c = tri_centroid(fp);
  const float lx = ord2f(b->lo[0]), ly = ord2f(b->lo[1]), lz = ord2f(b->lo[2]);
  float ex = fmaxf(ord2f(b->hi[0]) - lx, 1e-30f), ey = fmaxf(ord2f(b->hi[1]) - ly, 1e-30f);
  float ez = fmaxf(ord2f(b->hi[2]) - lz, 1e-30f);
#if B2PT_MORTON_UNIFORM
  // cubic cells: the grid spans the longest axis on all three, so the radix splits cut boxes towards cubes
  // instead of keeping the aspect ratio of the mesh (tools/exp_tree_quality.c: -4 % node visits, -6 % long
  // walks on the stand-in hull, whose extent is 4.6 x 1.9 x 4.8)
  ex = ey = ez = fmaxf(ex, fmaxf(ey, ez));
#endif
  const unsigned int qx = (unsigned int)fminf(fmaxf((c.x - lx) / ex * 1024.0f, 0.0f), 1023.0f);
  const unsigned int qy = (unsigned int)fminf(fmaxf((c.y - ly) / ey * 1024.0f, 0.0f), 1023.0f);
  const unsigned int qz = (unsigned int)fminf(fmaxf((c.z - lz) / ez * 1024.0f, 0.0f), 1023.0f);
  code[i] = (expand10(qx) << 2) | (expand10(qy) << 1) | expand10(qz);
  val[i] = (uint32_t)i;
}

// Leaves in Morton order: triangle triples and padded boxes.
__global__ void __launch_bounds__(256) k_leaves(const float* __restrict__ face_pos, const uint32_t* __restrict__ sorted_face,
                                                int n, float pad, float4* tris, float4* leaf_box) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int f = (int)sorted_face[s];
  float fp[9];
  for (int k = 0; k < 9; ++k) fp[k] = face_pos[9 * (size_t)f + k];
  tris[3 * (size_t)s + 0] = make_float4(fp[0], fp[1], fp[2], __int_as_float(f));
  tris[3 * (size_t)s + 1] = make_float4(fp[3], fp[4], fp[5], 0.0f);
  tris[3 * (size_t)s + 2] = make_float4(fp[6], fp[7], fp[8], 0.0f);
  leaf_box[2 * (size_t)s + 0] = make_float4(fminf(fminf(fp[0], fp[3]), fp[6]) - pad, fminf(fminf(fp[1], fp[4]), fp[7]) - pad,
                                            fminf(fminf(fp[2], fp[5]), fp[8]) - pad, 0.0f);
  leaf_box[2 * (size_t)s + 1] = make_float4(fmaxf(fmaxf(fp[0], fp[3]), fp[6]) + pad, fmaxf(fmaxf(fp[1], fp[4]), fp[7]) + pad,
                                            fmaxf(fmaxf(fp[2], fp[5]), fp[8]) + pad, 0.0f);
}

// delta(i, j): common prefix of codes i and j, ties broken by index (Karras 2012, sec. 4).
__device__ __forceinline__ int lbvh_delta(const uint32_t* __restrict__ code, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint32_t a = code[i], b = code[j];
  if (a == b) return 32 + __clz((unsigned int)i ^ (unsigned int)j);
  return __clz(a ^ b);
}

// One thread per internal node i in [0, n-2]: children and parent links.
// child encoding: >= 0 internal node, < 0 ~leaf slot.  parent[] has n-1 internal
// entries followed by n leaf entries.
__global__ void __launch_bounds__(256) k_karras(const uint32_t* __restrict__ code, int n, int2* children, int* parent,
                                                int2* range) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (lbvh_delta(code, n, i, i + 1) - lbvh_delta(code, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(code, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(code, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (lbvh_delta(code, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvh_delta(code, n, i, j);
  int s = 0;
  for (int div = 2, t = (l + div - 1) / div; ; div <<= 1, t = (l + div - 1) / div) {
    if (lbvh_delta(code, n, i, i + (s + t) * d) > dnode) s += t;
    if (t <= 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int left = (lo == gamma) ? ~gamma : gamma;
  const int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  children[i] = make_int2(left, right);
  range[i] = make_int2(lo, hi);  // the leaves (Morton slots) under this node, inclusive
  if (left >= 0) parent[left] = i; else parent[(n - 1) + gamma] = i;
  if (right >= 0) parent[right] = i; else parent[(n - 1) + gamma + 1] = i;
  if (i == 0) parent[0] = -1;
}

// Bottom-up refit: the second thread to reach a node computes its box.
__global__ void __launch_bounds__(256) k_refit(int n, const int2* __restrict__ children, const int* __restrict__ parent,
                                               const float4* __restrict__ leaf_box, float4* node_box, int* visit,
                                               TriBounds* info) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int node = parent[(n - 1) + s];
  int depth = 1;
  while (node >= 0) {
    if (atomicAdd(&visit[node], 1) == 0) return;  // first arrival: the sibling subtree is not done yet
    __threadfence();
    const int2 c = children[node];
    const volatile float4* lb = c.x >= 0 ? node_box + 2 * (size_t)c.x : leaf_box + 2 * (size_t)(~c.x);
    const volatile float4* rb = c.y >= 0 ? node_box + 2 * (size_t)c.y : leaf_box + 2 * (size_t)(~c.y);
    const float lminx = lb[0].x, lminy = lb[0].y, lminz = lb[0].z, lmaxx = lb[1].x, lmaxy = lb[1].y, lmaxz = lb[1].z;
    const float rminx = rb[0].x, rminy = rb[0].y, rminz = rb[0].z, rmaxx = rb[1].x, rmaxy = rb[1].y, rmaxz = rb[1].z;
    node_box[2 * (size_t)node + 0] = make_float4(fminf(lminx, rminx), fminf(lminy, rminy), fminf(lminz, rminz), 0.0f);
    node_box[2 * (size_t)node + 1] = make_float4(fmaxf(lmaxx, rmaxx), fmaxf(lmaxy, rmaxy), fmaxf(lmaxz, rmaxz), 0.0f);
    __threadfence();
    node = parent[node];
    ++depth;
  }
  (void)depth;
}

// Depth of every leaf (tree statistics only).
__global__ void __launch_bounds__(256) k_tree_depth(int n, const int* __restrict__ parent, TriBounds* info) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int node = parent[(n - 1) + s];
  int depth = 1;
  while (node >= 0) {
    node = parent[node];
    ++depth;
  }
  atomicMax(&info->max_depth, depth);
}

// The 15 floats of a face (positions, texture coordinates) as one aligned 64-byte record (DevMesh::face_rec).
__global__ void __launch_bounds__(256) k_face_records(const float* __restrict__ face_pos, const float* __restrict__ face_uv, int n,
                                                      float4* rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = face_pos + 9 * (size_t)i;
  const float* u = face_uv + 6 * (size_t)i;
  rec[4 * (size_t)i + 0] = make_float4(p[0], p[1], p[2], p[3]);
  rec[4 * (size_t)i + 1] = make_float4(p[4], p[5], p[6], p[7]);
  rec[4 * (size_t)i + 2] = make_float4(p[8], u[0], u[1], u[2]);
  rec[4 * (size_t)i + 3] = make_float4(u[3], u[4], u[5], 0.0f);
}

// ---- 4-wide traversal nodes with fat leaves -----------------------------------------------------------
// child ids:  >= 0 (and < kEmptyChild)  inner node index
//             kEmptyChild               unused slot
//             < 0                       leaf: ~((first Morton slot << 3) | (triangles - 1))
// node layout (8 x float4 = one 128-byte line), boxes per axis for the four slots:
//   f0 = min.x[0..3]  f1 = max.x   f2 = min.y  f3 = max.y   f4 = min.z  f5 = max.z
//   f6 = child ids as int bits   f7 = pad
// (both planes of an axis sit in one aligned 32-byte half-sector: the walk picks the plane a ray enters
// through with a 16-byte address offset).  An unused slot has min = +FLT_MAX, max = -FLT_MAX.
constexpr int kEmptyChild = 0x40000000;
#ifndef B2PT_WIDE
#define B2PT_WIDE 4
#endif
// Children per traversal node: 4, or 8 as TWO of the 128-byte lines above side by side (slots 0..3 in the first line,
// 4..7 in the second: the walk runs the same four-slot box test on each).  kNodeF4 float4s per node.
constexpr int kWide = B2PT_WIDE;
static_assert(kWide == 4 || kWide == 8, "4- or 8-wide nodes");
constexpr int kNodeF4 = 2 * kWide;
#ifndef B2PT_LEAF_TRIS
#define B2PT_LEAF_TRIS 2
#endif
constexpr int kLeafTris = B2PT_LEAF_TRIS;  // triangles per leaf, at most
static_assert(kLeafTris >= 1 && kLeafTris <= 8, "leaf codes keep the count in three bits");

__host__ __device__ __forceinline__ int leaf_code(int first, int count) { return ~((first << 3) | (count - 1)); }
__host__ __device__ __forceinline__ int leaf_first(int code) { return (~code) >> 3; }
__host__ __device__ __forceinline__ int leaf_count(int code) { return ((~code) & 7) + 1; }

__device__ __forceinline__ float box_area(const float4* b) {
  const float dx = b[1].x - b[0].x, dy = b[1].y - b[0].y, dz = b[1].z - b[0].z;
  return dx * dy + dy * dz + dz * dx;
}

// One thread per binary node: its wide node (only those reachable from the root survive the compaction).
// Child ids are still BINARY node indices here; k_compact_nodes renumbers them.
__global__ void __launch_bounds__(256) k_emit_wide4(int n, const int2* __restrict__ children, const int2* __restrict__ range,
                                                    const float4* __restrict__ leaf_box, const float4* __restrict__ node_box,
                                                    float4* nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  // a child can be opened if it is an inner node with more triangles than a leaf may hold
  auto openable = [&](int c) { return c >= 0 && range[c].y - range[c].x + 1 > kLeafTris; };
  int id[kWide];
  for (int q = 0; q < kWide; ++q) id[q] = kEmptyChild;
  int k = 2;
  {
    const int2 c = children[i];
    id[0] = c.x;
    id[1] = c.y;
  }
  for (int round = 0; round < kWide - 2; ++round) {
    int pick = -1;
    float best_area = -1.0f;
    for (int q = 0; q < k; ++q) {
      if (!openable(id[q])) continue;
      const float a = box_area(node_box + 2 * (size_t)id[q]);
      if (a > best_area) {
        best_area = a;
        pick = q;
      }
    }
    if (pick < 0) break;
    const int2 g = children[id[pick]];
    id[pick] = g.x;
    id[k++] = g.y;
  }
  float lo[3][kWide], hi[3][kWide];
  int code[kWide];
  for (int q = 0; q < kWide; ++q) {
    float4 b0 = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.0f), b1 = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.0f);
    code[q] = kEmptyChild;
    if (id[q] != kEmptyChild) {
      const float4* b = id[q] >= 0 ? node_box + 2 * (size_t)id[q] : leaf_box + 2 * (size_t)(~id[q]);
      b0 = b[0];
      b1 = b[1];
      if (id[q] < 0) code[q] = leaf_code(~id[q], 1);
      else if (!openable(id[q])) code[q] = leaf_code(range[id[q]].x, range[id[q]].y - range[id[q]].x + 1);
      else code[q] = id[q];
    }
    lo[0][q] = b0.x; lo[1][q] = b0.y; lo[2][q] = b0.z;
    hi[0][q] = b1.x; hi[1][q] = b1.y; hi[2][q] = b1.z;
  }
  for (int h = 0; h < kWide / 4; ++h) {  // one 128-byte line per four slots
    float4* o = nodes + kNodeF4 * (size_t)i + 8 * h;
    const int q = 4 * h;
    for (int a = 0; a < 3; ++a) {
      o[2 * a] = make_float4(lo[a][q], lo[a][q + 1], lo[a][q + 2], lo[a][q + 3]);
      o[2 * a + 1] = make_float4(hi[a][q], hi[a][q + 1], hi[a][q + 2], hi[a][q + 3]);
    }
    o[6] = make_float4(__int_as_float(code[q]), __int_as_float(code[q + 1]), __int_as_float(code[q + 2]), __int_as_float(code[q + 3]));
    o[7] = make_float4(__int_as_float(k), 0.0f, 0.0f, 0.0f);  // children of the whole node (not read by the walk)
  }
}

// Reachability and depth of the wide tree, one level per launch: the nodes reached at `level` mark their
// inner children with level + 1 (wdepth[0] = 1 for the root before the first launch, 0 = not reachable).  A
// walk's stack holds at most three entries per inner wide node on its path plus the four of the node being
// opened, which is what the traversal stacks are sized against.
__global__ void __launch_bounds__(256) k_wide_levels(int n, const float4* __restrict__ nodes, int* wdepth, int level,
                                                     TriBounds* info) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1 || wdepth[i] != level) return;
  for (int h = 0; h < kWide / 4; ++h) {
    const float4 cf = nodes[kNodeF4 * (size_t)i + 8 * h + 6];
    const int ch[4] = {__float_as_int(cf.x), __float_as_int(cf.y), __float_as_int(cf.z), __float_as_int(cf.w)};
    for (int k = 0; k < 4; ++k)
      if (ch[k] >= 0 && ch[k] != kEmptyChild) wdepth[ch[k]] = level + 1;
  }
  atomicMax(&info->wide_depth, level);
}

__global__ void __launch_bounds__(256) k_reach_flags(int n, const int* __restrict__ wdepth, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n - 1) flag[i] = wdepth[i] != 0;
}

// Move the reachable nodes to their compact slots (slot[] = exclusive scan of the flags) and renumber the
// inner child ids.  The root (binary node 0) stays node 0.
__global__ void __launch_bounds__(256) k_compact_nodes(int n, const float4* __restrict__ src, const int* __restrict__ wdepth,
                                                       const int* __restrict__ slot, float4* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1 || wdepth[i] == 0) return;
  for (int h = 0; h < kWide / 4; ++h) {
    const float4* s = src + kNodeF4 * (size_t)i + 8 * h;
    float4* d = dst + kNodeF4 * (size_t)slot[i] + 8 * h;
    for (int k = 0; k < 6; ++k) d[k] = s[k];
    const float4 cf = s[6];
    int ch[4] = {__float_as_int(cf.x), __float_as_int(cf.y), __float_as_int(cf.z), __float_as_int(cf.w)};
    for (int k = 0; k < 4; ++k)
      if (ch[k] >= 0 && ch[k] != kEmptyChild) ch[k] = slot[ch[k]];
    d[6] = make_float4(__int_as_float(ch[0]), __int_as_float(ch[1]), __int_as_float(ch[2]), __int_as_float(ch[3]));
    d[7] = s[7];
  }
}

}  // namespace b2pt
