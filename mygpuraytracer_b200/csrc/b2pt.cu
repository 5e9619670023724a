// b2pt.cu -- context, wavefront loop and the C ABI of include/b2pt.h.
//
// Host side of the hot path: what pathtraceInit / pathtrace / pathtraceFree do
// in the reference (apps/src/pathtrace.cu:130-223, 527-671), re-designed so
// that one iteration is a fixed sequence of kernels with no host round trip:
//
//   k_iter_begin                       (iteration number and counters live on the device)
//   k_generate                         generateRayFromCamera
//   for d in 0 .. depth-1:
//     k_intersect_analytic             memset + computeIntersections (cubes, spheres) + sort key + histograms
//     k_mesh_walk, k_mesh_walk_long,   computeIntersections (OBJ geoms): LBVH walk, long walks, record completion
//     k_mesh_finish
//     k_sort_material_few / k_sort_material   thrust::sort_by_key  -> 4-byte permutation + compaction ranks
//                                        (<= 8 materials: the packed-counter kernel; else the 256-bin one)
//     k_shade_compact                  shadeFakeMaterial + stable_partition + finalGather
//
// Every kernel reads its element count from device memory, so the sequence is
// identical for every iteration and is captured once into a CUDA graph.
// Compiled with -fmad=false (see pt_math.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b2pt.h"
#include "k_generate.cuh"
#include "k_intersect.cuh"
#include "k_lbvh.cuh"
#include "k_prims.cuh"
#include "k_shade.cuh"
#include "k_sort.cuh"
#include "k_walk.cuh"
#include "k_fused.cuh"
#include "host/hdr_writer.h"
#include "host/png_writer.h"
#include "pt_device.cuh"

using namespace b2pt;

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      char b_[512];                                                                                      \
      snprintf(b_, sizeof b_, "CUDA error (%s:%d): %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return fail(B2PT_ERR_CUDA, b_);                                                                    \
    }                                                                                                    \
  } while (0)

extern "C" const char* b2pt_last_error(void) { return g_last_error.c_str(); }
// used by host/scene_loader.cpp (same shared object) to report through the same channel
extern "C" void b2pt_set_last_error_(const char* msg) { g_last_error = msg ? msg : ""; }
extern "C" int b2pt_abi_version(void) { return B2PT_ABI_VERSION; }
extern "C" int b2pt_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(B2PT_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  return n;
}

extern "C" void b2pt_default_options(B2ptOptions* o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->struct_size = sizeof(B2ptOptions);
  o->device = 0;
  o->antialiasing = 1;       // ANTIALIASING 1, pathtrace.cu:39
  o->depth_of_field = 0;     // DEPTH_OF_FIELD 0, pathtrace.cu:36
  o->lens_radius = 0.8f;     // pathtrace.cu:279
  o->focal_distance = 11.0f; // pathtrace.cu:280
  o->sort_by_material = 1;   // SORT_BY_MATERIAL 1, pathtrace.cu:38
  o->cache_first_bounce = 0;
  o->trig_mode = B2PT_TRIG_NATIVE;
  o->rng_mode = B2PT_RNG_SLOT;
  o->use_bvh = 1;
  o->record_stages = 0;
  o->use_graph = 1;
  o->concurrent_contexts = 1;
  o->persistent_host_albedo = 0;  // copy the albedo AOV every call, like pathtrace.cu:666-668
}

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
struct MeshBuild {
  float4* nodes = nullptr;
  float4* tris = nullptr;
  int root = 0;        // DevMesh::root
  int wide_depth = 0;  // inner wide nodes on the longest root-to-leaf path
  float* face_pos = nullptr;
  float* face_uv = nullptr;
  float4* face_rec = nullptr;
  size_t bvh_bytes = 0;
  B2ptBvhInfo info{};
};

struct StageRecord {
  int n = 0;
  std::vector<float4> in_s0, in_s1, in_s2, h0, h1, sh_s0, sh_s1, sh_s2;
  std::vector<int> perm, live, dead;
  int n_live_out = 0;
};

struct SceneStore {
  int device = 0;
  std::vector<void*> allocs;
  ~SceneStore() {
    cudaSetDevice(device);
    for (void* p : allocs) cudaFree(p);
  }
};

struct B2ptCtx {
  int device = 0;
  cudaStream_t stream = nullptr;      // where work is queued
  cudaStream_t own_stream = nullptr;  // created by the context
  B2ptOptions opt{};
  int W = 0, H = 0, P = 0, depth = 0, loop_depth = 0;
  int n_geoms = 0, n_materials = 0;
  int sm_count = 0;
  float origin_world = 0.0f;  // rays start inside [-origin_world, origin_world]^3 (BVH padding, set_camera check)

  std::vector<void*> allocs;  // every cudaMalloc of this context, freed in b2pt_destroy
  // geoms, materials, BVHs, triangles and textures: owned by the first context of a scene and shared (read
  // only) with the contexts created from it by b2pt_create_shared; freed with the last of them
  std::shared_ptr<SceneStore> scene_store;
  bool alloc_to_scene = false;
  std::vector<MeshBuild> meshes;
  std::vector<int> geom_mesh;
  DevScene dscene{};
  GenParams gen{};

  PathBuf buf[2]{};
  HitBuf hits[2]{};        // hit records, sort keys and survival flags of depth d live in buffer d & 1:
  uint8_t* key[2]{};       // the fused shade kernel writes those of depth d+1 while it reads those of depth d
  uint8_t* live[2]{};
  bool fuse = true;        // analytic intersection fused into generate / shade (k_fused.cuh)
  int* perm = nullptr;
  int* apos = nullptr;
  unsigned long long* sort_status_live = nullptr;
  Counters* ctr = nullptr;
  int* iter_state = nullptr;
  unsigned long long* sort_status = nullptr;
  float* image = nullptr;
  float* albedo = nullptr;
  float* image_target = nullptr;  // where the gather accumulates (own image or caller's buffer)

  // stage recording
  float4 *rec_s0 = nullptr, *rec_s1 = nullptr, *rec_s2 = nullptr;
  int *rec_live = nullptr, *rec_dead = nullptr;
  std::vector<StageRecord> records;

  unsigned long long* trav_stats = nullptr;  // [depth][32], B2PT_TRAVERSAL_STATS=1 only
  void* l2_window_ptr = nullptr;
  size_t l2_window_bytes = 0;
  // first-bounce cache (pathtrace.cu:586-610, intended semantics): with AA and DOF off the camera rays are the
  // same every iteration, so depth 0 keeps its own hit/key/live buffers, filled once and reused afterwards
  // the albedo AOV only changes on iteration 1 (pathtrace.cu:412) and on a reset: b2pt_pathtrace skips the
  // D2H copy into a host buffer that already holds the current version
  uint64_t albedo_version = 1, albedo_host_version = 0;
  const float* albedo_host_last = nullptr;
  bool fb_enabled = false, fb_valid = false;
  HitBuf fb_hits{};
  uint8_t *fb_key = nullptr, *fb_live = nullptr;
  unsigned int* fb_hist = nullptr;  // [2][kMaxMaterials]: hist[0], hist_live[0] of the cached depth
  int walk_grid = 0, finish_grid = 0, long_grid = 0, long_walk = 24;
  int long_lanes = 32;  // lanes per long walk: 32 alone, 16 when contexts share the SMs (B2PT_LONG_LANES)
  int2* long_queue = nullptr;
  float4* long_best = nullptr;
  int2* long_stack = nullptr;
  int* long_n = nullptr;
  int long_cap = 0, long_carry = kLongCarry;
  int shade_stride_grid = 0, gen_trace_grid = 0;
  int isect_grid = 0, analytic_grid = 0, sort_grid = 0, shade_grid = 0, gen_grid = 0;
  bool sort_general = false;  // B2PT_SORT_GENERAL=1: the 256-bin material sort even when the few-materials kernel applies (tests)
  float4* mesh_queue = nullptr;  // [3][P]: the rays that walk a mesh, see IsectParams::queue
  cudaEvent_t ev_loop_a = nullptr, ev_loop_b = nullptr;
  bool loop_timed = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_kernels = 0;
  int64_t launches = 0;

  template <typename T>
  int dalloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(B2PT_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    (alloc_to_scene && scene_store ? scene_store->allocs : allocs).push_back(q);
    *p = (T*)q;
    return 0;
  }
};

// Keep the largest mesh's BVH (nodes + triangles) resident in L2: the walk is a
// chain of dependent loads, so its speed is set by the latency of each step
// (L2 hit ~250 cycles, HBM ~600+), and the streaming path-state / texture
// traffic of the other kernels would otherwise keep evicting it.
static void apply_l2_policy(B2ptCtx* c, cudaStream_t s) {
  if (!c->l2_window_ptr || !c->l2_window_bytes) return;
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof v);
  v.accessPolicyWindow.base_ptr = c->l2_window_ptr;
  v.accessPolicyWindow.num_bytes = c->l2_window_bytes;
  v.accessPolicyWindow.hitRatio = 1.0f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

static Mat34 rows_of(const float* m) {
  Mat34 r;
  r.r0 = make_float4(m[0], m[4], m[8], m[12]);
  r.r1 = make_float4(m[1], m[5], m[9], m[13]);
  r.r2 = make_float4(m[2], m[6], m[10], m[14]);
  return r;
}

// 1 if the upper 3x3 of the transform is orthonormal (object-space distances
// equal world-space distances).
static int is_rigid(const float* m) {
  for (int a = 0; a < 3; ++a)
    for (int b = a; b < 3; ++b) {
      double d = 0;
      for (int k = 0; k < 3; ++k) d += (double)m[a * 4 + k] * (double)m[b * 4 + k];
      if (std::fabs(d - (a == b ? 1.0 : 0.0)) > 1e-5) return 0;
    }
  return 1;
}

// Padded world-space box of a geom whose object-space bounds are [lo, hi], and
// the distance slack of its exact test (see may_beat in k_intersect.cuh).
static void world_box(const float* m, const float lo[3], const float hi[3], DevGeom* D) {
  float wl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, wh[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int c = 0; c < 8; ++c) {
    const float p[3] = {(c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]};
    for (int r = 0; r < 3; ++r) {
      const float w = m[r] * p[0] + m[4 + r] * p[1] + m[8 + r] * p[2] + m[12 + r];
      wl[r] = std::min(wl[r], w);
      wh[r] = std::max(wh[r], w);
    }
  }
  float ext = 0.0f, scale = 0.0f;
  for (int r = 0; r < 3; ++r) ext = std::max(ext, std::max(std::fabs(wl[r]), std::fabs(wh[r])));
  for (int c = 0; c < 3; ++c)
    scale = std::max(scale, std::sqrt(m[c * 4] * m[c * 4] + m[c * 4 + 1] * m[c * 4 + 1] + m[c * 4 + 2] * m[c * 4 + 2]));
  const float pad = 1e-3f + 1e-4f * ext;
  D->wmin = make_float4(wl[0] - pad, wl[1] - pad, wl[2] - pad, 2e-3f + 1.5e-4f * scale);
  D->wmax = make_float4(wh[0] + pad, wh[1] + pad, wh[2] + pad, 0.0f);
}

static DevCamera to_dev_camera(const B2ptCamera& c) {
  DevCamera d;
  d.res_x = c.resolution[0];
  d.res_y = c.resolution[1];
  d.position = {c.position[0], c.position[1], c.position[2]};
  d.view = {c.view[0], c.view[1], c.view[2]};
  d.up = {c.up[0], c.up[1], c.up[2]};
  d.right = {c.right[0], c.right[1], c.right[2]};
  d.plx = c.pixel_length[0];
  d.ply = c.pixel_length[1];
  return d;
}

// ---------------------------------------------------------------------------------
// radix sort of pairs (device arrays), used by the LBVH build and exported
// ---------------------------------------------------------------------------------
struct Scratch {  // frees everything on scope exit
  std::vector<void*> p;
  ~Scratch() { for (void* q : p) cudaFree(q); }
  template <typename T>
  cudaError_t get(T** out, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) p.push_back(q);
    *out = (T*)q;
    return e;
  }
};

struct RadixTemps {
  unsigned int* hist = nullptr;     // [4][256]
  unsigned int* tickets = nullptr;  // [4]
  unsigned long long* status = nullptr;
  uint32_t* key_alt = nullptr;
  uint32_t* val_alt = nullptr;
};

static std::atomic<unsigned int> g_radix_epoch{1};  // contexts may be created from several host threads

static int radix_temps_alloc(RadixTemps* t, int n) {
  const int tiles = (n + kSortTile - 1) / kSortTile;
  CK(cudaMalloc(&t->hist, 4 * 256 * sizeof(unsigned int)));
  CK(cudaMalloc(&t->tickets, 4 * sizeof(unsigned int)));
  CK(cudaMalloc(&t->status, (size_t)tiles * 256 * sizeof(unsigned long long)));
  CK(cudaMalloc(&t->key_alt, (size_t)std::max(n, 1) * 4));
  CK(cudaMalloc(&t->val_alt, (size_t)std::max(n, 1) * 4));
  return 0;
}
static void radix_temps_free(RadixTemps* t) {
  cudaFree(t->hist);
  cudaFree(t->tickets);
  cudaFree(t->status);
  cudaFree(t->key_alt);
  cudaFree(t->val_alt);
  *t = RadixTemps();
}
struct RadixTempsGuard {  // frees on every return path
  RadixTemps t;
  ~RadixTempsGuard() { radix_temps_free(&t); }
};
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

// Sorts n pairs in place (result ends in key/val after four passes).  Asynchronous on `s`; the temporaries
// must stay alive until the stream has run the passes.
static int radix_sort_pairs_dev(uint32_t* key, uint32_t* val, int n, cudaStream_t s, int64_t* launches, const RadixTemps& t) {
  if (n <= 1) return 0;
  const int tiles = (n + kSortTile - 1) / kSortTile;
  CK(cudaMemsetAsync(t.hist, 0, 4 * 256 * sizeof(unsigned int), s));
  CK(cudaMemsetAsync(t.tickets, 0, 4 * sizeof(unsigned int), s));
  CK(cudaMemsetAsync(t.status, 0, (size_t)tiles * 256 * sizeof(unsigned long long), s));
  k_radix_hist<<<std::min(tiles * 4, 1184), 256, 0, s>>>(key, n, t.hist);
  uint32_t *ki = key, *vi = val, *ko = t.key_alt, *vo = t.val_alt;
  for (int pass = 0; pass < 4; ++pass) {
    RadixPassPolicy pol;
    pol.key_in = ki;
    pol.val_in = vi;
    pol.key_out = ko;
    pol.val_out = vo;
    pol.hist = t.hist + pass * 256;
    pol.ticket_ = t.tickets + pass;
    pol.status_ = t.status;
    pol.epoch_ = g_radix_epoch.fetch_add(1u);
    pol.n_ = n;
    pol.shift = pass * 8;
    k_onesweep_pass<RadixPassPolicy><<<tiles, kSortThreads, 0, s>>>(pol);
    std::swap(ki, ko);
    std::swap(vi, vo);
  }
  if (launches) *launches += 5;
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// LBVH build for one mesh
// ---------------------------------------------------------------------------------
static int build_mesh(B2ptCtx* c, const float* pos_host, const float* uv_host, int n, float origin_radius, MeshBuild* out) {
  cudaStream_t s = c->stream;
  int rc;
  if ((rc = c->dalloc(&out->face_pos, (size_t)n * 9))) return rc;
  if ((rc = c->dalloc(&out->face_uv, (size_t)n * 6))) return rc;
  CK(cudaMemcpyAsync(out->face_pos, pos_host, (size_t)n * 36, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(out->face_uv, uv_host, (size_t)n * 24, cudaMemcpyHostToDevice, s));
  if ((rc = c->dalloc(&out->face_rec, (size_t)n * 4))) return rc;
  k_face_records<<<(n + 255) / 256, 256, 0, s>>>(out->face_pos, out->face_uv, n, out->face_rec);
  c->launches += 1;
  out->info.n_faces = n;

  // pad: a few 1e-6 of the mesh extent (host pass over the positions)
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = 0; i < (size_t)n * 3; ++i)
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], pos_host[i * 3 + k]);
      hi[k] = std::max(hi[k], pos_host[i * 3 + k]);
    }
  float ext = 0.0f;
  for (int k = 0; k < 3; ++k) ext = std::max(ext, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
  // second term: the FMA slab test of k_walk.cuh displaces a plane by up to ~2^-23 * (largest coordinate of a
  // ray origin in this mesh's object space); origin_radius bounds that coordinate (pad = 2^-19 * it)
  const float pad = 4e-6f * std::max(ext, 1.0f) + 1.9073486328125e-6f * origin_radius;

  // temporaries, events and the radix buffers are released on every return path
  Scratch tmp;
  TriBounds* bounds = nullptr;
  uint32_t *code = nullptr, *val = nullptr;
  float4 *leaf_box = nullptr, *node_box = nullptr, *tris_tmp = nullptr, *nodes_tmp = nullptr;
  int2 *children = nullptr, *range = nullptr;
  int *parent = nullptr, *wdepth = nullptr, *flag = nullptr, *slot = nullptr;
  unsigned int* scan_ticket = nullptr;
  unsigned long long* scan_status = nullptr;
  const int inner = std::max(n - 1, 1);
  const int scan_tiles = (inner + kScanTile - 1) / kScanTile;
  CK(tmp.get(&bounds, 1));
  CK(tmp.get(&code, (size_t)n));
  CK(tmp.get(&val, (size_t)n));
  CK(tmp.get(&leaf_box, (size_t)n * 2));
  CK(tmp.get(&node_box, (size_t)inner * 2));
  CK(tmp.get(&tris_tmp, (size_t)n * 3));
  CK(tmp.get(&nodes_tmp, (size_t)inner * kNodeF4));
  CK(tmp.get(&children, (size_t)inner));
  CK(tmp.get(&range, (size_t)inner));
  CK(tmp.get(&parent, (size_t)(2 * n)));
  CK(tmp.get(&wdepth, (size_t)inner));
  CK(tmp.get(&flag, (size_t)inner));
  CK(tmp.get(&slot, (size_t)inner));
  CK(tmp.get(&scan_ticket, 1));
  CK(tmp.get(&scan_status, (size_t)scan_tiles));
  RadixTempsGuard rguard;  // allocated up front: build_ms below is device time, not cudaMalloc latency
  RadixTemps& rtemps = rguard.t;
  if ((rc = radix_temps_alloc(&rtemps, n))) return rc;

  // build_ms is DEVICE time: the build has two small read-backs (tree depth, node count) and one allocation of
  // the final size in the middle, so it is timed as three segments between them and the segments are added
  EventPair ev, ev2, ev3;
  CK(cudaEventCreate(&ev.a));
  CK(cudaEventCreate(&ev.b));
  CK(cudaEventCreate(&ev2.a));
  CK(cudaEventCreate(&ev2.b));
  CK(cudaEventCreate(&ev3.a));
  CK(cudaEventCreate(&ev3.b));
  CK(cudaEventRecord(ev.a, s));
  const int blocks = (n + 255) / 256;
  k_bounds_init<<<1, 32, 0, s>>>(bounds);
  k_centroid_bounds<<<std::min(blocks, 1184), 256, 0, s>>>(out->face_pos, n, bounds);
  k_morton<<<blocks, 256, 0, s>>>(out->face_pos, n, bounds, code, val);
  c->launches += 3;
  if ((rc = radix_sort_pairs_dev(code, val, n, s, &c->launches, rtemps))) return rc;
  k_leaves<<<blocks, 256, 0, s>>>(out->face_pos, val, n, pad, tris_tmp, leaf_box);
  c->launches += 1;
  TriBounds hb;
  memset(&hb, 0, sizeof hb);
  int n_nodes = 0;  // wide nodes reachable from the root
  const bool tree = n > kLeafTris;  // a mesh of at most kLeafTris triangles is a single leaf
  if (tree) {
    CK(cudaMemsetAsync(wdepth, 0, (size_t)(n - 1) * sizeof(int), s));
    const int iblocks = (n - 1 + 255) / 256;
    k_karras<<<iblocks, 256, 0, s>>>(code, n, children, parent, range);
    k_refit<<<blocks, 256, 0, s>>>(n, children, parent, leaf_box, node_box, wdepth, bounds);
    k_tree_depth<<<blocks, 256, 0, s>>>(n, parent, bounds);
    k_emit_wide4<<<iblocks, 256, 0, s>>>(n, children, range, leaf_box, node_box, nodes_tmp);
    c->launches += 4;
    // which wide nodes the root reaches, and how deep the wide tree is: one level per launch (wdepth[] served
    // the refit as its visit counters and is free again); a binary tree of depth D gives at most D wide levels,
    // and D is known by now (one small read-back)
    CK(cudaMemsetAsync(wdepth, 0, (size_t)(n - 1) * sizeof(int), s));
    const int one = 1;
    CK(cudaMemcpyAsync(wdepth, &one, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ev.b, s));
    CK(cudaMemcpyAsync(&hb, bounds, sizeof hb, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(ev2.a, s));
    const int levels = std::max(1, std::min(hb.max_depth, 64));
    for (int level = 1; level <= levels; ++level) k_wide_levels<<<iblocks, 256, 0, s>>>(n, nodes_tmp, wdepth, level, bounds);
    c->launches += levels;
    // number the reachable nodes (exclusive scan of the flags) ...
    k_reach_flags<<<iblocks, 256, 0, s>>>(n, wdepth, flag);
    CK(cudaMemsetAsync(scan_ticket, 0, sizeof(unsigned int), s));
    CK(cudaMemsetAsync(scan_status, 0, (size_t)scan_tiles * sizeof(unsigned long long), s));
    k_scan_family<0><<<scan_tiles, kScanThreads, 0, s>>>(flag, nullptr, n - 1, slot, nullptr, scan_ticket, scan_status, 1u, nullptr);
    c->launches += 2;
    CK(cudaEventRecord(ev2.b, s));
    int last[2] = {0, 0};
    CK(cudaMemcpyAsync(&last[0], slot + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&last[1], flag + (n - 2), sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    n_nodes = last[0] + last[1];
  }
  // ... and keep only those: nodes and triangles share one allocation so that a single L2 access-policy window
  // can keep the whole acceleration structure resident
  const size_t node_f4 = (size_t)std::max(n_nodes, 1) * kNodeF4, tri_f4 = (size_t)n * 3;
  if ((rc = c->dalloc(&out->nodes, node_f4 + tri_f4))) return rc;
  out->tris = out->nodes + node_f4;
  out->bvh_bytes = (node_f4 + tri_f4) * sizeof(float4);
  out->info.n_nodes = n_nodes;
  if (!tree) CK(cudaEventRecord(ev.b, s));
  CK(cudaEventRecord(ev3.a, s));
  if (tree) {
    k_compact_nodes<<<(n - 1 + 255) / 256, 256, 0, s>>>(n, nodes_tmp, wdepth, slot, out->nodes);
    c->launches += 1;
  } else {
    CK(cudaMemsetAsync(out->nodes, 0, node_f4 * sizeof(float4), s));
  }
  CK(cudaMemcpyAsync(out->tris, tris_tmp, tri_f4 * sizeof(float4), cudaMemcpyDeviceToDevice, s));
  CK(cudaEventRecord(ev3.b, s));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  {
    float m1 = 0.0f, m2 = 0.0f, m3 = 0.0f;
    CK(cudaEventElapsedTime(&m1, ev.a, ev.b));
    if (tree) CK(cudaEventElapsedTime(&m2, ev2.a, ev2.b));
    CK(cudaEventElapsedTime(&m3, ev3.a, ev3.b));
    out->info.build_ms = m1 + m2 + m3;
  }
  CK(cudaMemcpy(&hb, bounds, sizeof hb, cudaMemcpyDeviceToHost));
  out->info.max_depth = tree ? hb.max_depth : 1;
  out->wide_depth = tree ? hb.wide_depth : 0;
  out->root = tree ? 0 : leaf_code(0, n);
  if (getenv("B2PT_TRAVERSAL_STATS"))
    fprintf(stderr, "[b2pt bvh] %d triangles: binary depth %d, %d wide nodes (%.1f MB), wide depth %d, leaves of <= %d triangles\n", n,
            out->info.max_depth, n_nodes, n_nodes * 16.0 * kNodeF4 / 1e6, out->wide_depth, kLeafTris);
  // a walk's stack holds at most three entries per inner wide node on its path (nearest child first, the other
  // hits pushed); k_mesh_walk_long's depth-first fallback relies on the same bound
  if (out->info.max_depth > 64 || (kWide - 1) * out->wide_depth > kWalkShort + kWalkSpill)
    return fail(B2PT_ERR_RANGE, "LBVH deeper than the traversal stack");
  return 0;
}

// ---------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------
static int upload_texture(B2ptCtx* c, const B2ptScene* sc, int idx, DevTexture* out, std::vector<DevTexture>* cache) {
  out->texels = nullptr;
  out->w = out->h = out->channels = 0;
  if (idx < 0) return 0;
  if (idx >= sc->n_textures) return fail(B2PT_ERR_INVALID, "texture index out of range");
  if (cache && (*cache)[(size_t)idx].texels) {  // several materials / geoms may name the same map
    *out = (*cache)[(size_t)idx];
    return 0;
  }
  const B2ptTexture& t = sc->textures[idx];
  if (t.channels == 0 || t.texels == nullptr) return 0;
  if (t.channels < 1 || t.width <= 0 || t.height <= 0)
    return fail(B2PT_ERR_INVALID, "textures need >= 1 channel and positive dimensions");
  // The fetch reads three consecutive bytes per texel whatever the channel count, like the reference
  // (apps/src/interactions.h:199-212): with 1 or 2 channels that runs into the next texels and, for the last ones,
  // past the end of the image (undefined in the reference).  Two padding bytes that repeat the last byte make that
  // read defined (the oracle reads the same) without a bounds check per texel.
  uint8_t* d = nullptr;
  const size_t bytes = (size_t)t.width * t.height * t.channels;
  int rc = c->dalloc(&d, bytes + 2);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d, t.texels, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(d + bytes, t.texels[bytes - 1], 2, c->stream));
  out->texels = d;
  out->w = t.width;
  out->h = t.height;
  out->channels = t.channels;
  if (cache) (*cache)[(size_t)idx] = *out;
  return 0;
}

static int create_impl(const B2ptScene* sc, const B2ptOptions* opt_in, B2ptCtx* c, B2ptCtx* parent = nullptr) {
  B2ptOptions opt;
  b2pt_default_options(&opt);
  if (opt_in) {
    if (opt_in->struct_size == 0 || opt_in->struct_size > sizeof(B2ptOptions))
      return fail(B2PT_ERR_INVALID, "B2ptOptions.struct_size is not set");
    memcpy(&opt, opt_in, opt_in->struct_size);
    opt.struct_size = sizeof(B2ptOptions);
  }
  c->opt = opt;
  const int W = sc->camera.resolution[0], H = sc->camera.resolution[1];
  if (W <= 0 || H <= 0 || (long long)W * H >= (1ll << 30)) return fail(B2PT_ERR_INVALID, "bad resolution");
  if (sc->n_geoms < 0 || sc->n_geoms > kMaxGeoms) return fail(B2PT_ERR_RANGE, "at most 64 geoms are supported");
  if (sc->n_materials <= 0 || sc->n_materials > kMaxMaterials)
    return fail(B2PT_ERR_RANGE, "between 1 and 256 materials are supported");
  if (sc->trace_depth < 0 || sc->trace_depth > kMaxDepth) return fail(B2PT_ERR_RANGE, "trace depth must be in [0, 62]");
  if (sc->n_geoms > 0 && !sc->geoms) return fail(B2PT_ERR_INVALID, "geoms is NULL");
  if (!sc->materials) return fail(B2PT_ERR_INVALID, "materials is NULL");
  c->W = W;
  c->H = H;
  c->P = W * H;
  c->depth = sc->trace_depth;
  c->loop_depth = std::max(sc->trace_depth, 1);  // the reference's while loop always runs once
  c->n_geoms = sc->n_geoms;
  c->n_materials = sc->n_materials;

  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (opt.device < 0 || opt.device >= ndev) return fail(B2PT_ERR_INVALID, "no such CUDA device");
  c->device = opt.device;
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  c->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaEventCreate(&c->ev_loop_a));
  CK(cudaEventCreate(&c->ev_loop_b));

  int rc;
  if (parent) {
    // the scene already lives on this device: share it (read only) instead of uploading and building again
    if (parent->device != c->device) return fail(B2PT_ERR_INVALID, "a shared context must be on its parent's device");
    if (parent->n_geoms != sc->n_geoms || parent->n_materials != sc->n_materials)
      return fail(B2PT_ERR_INVALID, "b2pt_create_shared: `scene` is not the scene the parent was created from");
    c->scene_store = parent->scene_store;
    c->origin_world = parent->origin_world;
    c->geom_mesh = parent->geom_mesh;
    c->meshes = parent->meshes;
    c->dscene = parent->dscene;
    c->l2_window_ptr = parent->l2_window_ptr;
    c->l2_window_bytes = parent->l2_window_bytes;
    if (c->l2_window_ptr) apply_l2_policy(c, c->stream);
  } else {
  c->scene_store = std::make_shared<SceneStore>();
  c->scene_store->device = c->device;
  c->alloc_to_scene = true;
  // ---- where rays can start: every geom's world box and the camera, with a 4x margin for camera moves ----
  float origin_world = 0.0f;
  for (int k = 0; k < 3; ++k) origin_world = std::max(origin_world, std::fabs(sc->camera.position[k]));
  for (int g = 0; g < sc->n_geoms; ++g) {
    const B2ptGeom& G = sc->geoms[g];
    float lo[3] = {-0.5f, -0.5f, -0.5f}, hi[3] = {0.5f, 0.5f, 0.5f};
    if (G.type == B2PT_OBJ && G.face_count > 0 && sc->face_pos && G.face_begin >= 0 &&
        (long long)G.face_begin + G.face_count <= sc->n_faces) {
      for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; }
      const float* fp = sc->face_pos + (size_t)G.face_begin * 9;
      for (size_t v = 0; v < (size_t)G.face_count * 3; ++v)
        for (int k = 0; k < 3; ++k) {
          lo[k] = std::min(lo[k], fp[v * 3 + k]);
          hi[k] = std::max(hi[k], fp[v * 3 + k]);
        }
    }
    DevGeom tmp;
    world_box(G.transform, lo, hi, &tmp);
    origin_world = std::max(origin_world, std::max({std::fabs(tmp.wmin.x), std::fabs(tmp.wmin.y), std::fabs(tmp.wmin.z),
                                                     std::fabs(tmp.wmax.x), std::fabs(tmp.wmax.y), std::fabs(tmp.wmax.z)}));
  }
  origin_world *= 4.0f;
  c->origin_world = origin_world;
  // ---- geoms, meshes, textures -----------------------------------------------------
  std::vector<DevGeom> hg(sc->n_geoms);
  std::vector<DevMesh> hm;
  std::vector<DevTexture> tex_cache((size_t)std::max(sc->n_textures, 0));
  for (DevTexture& t : tex_cache) t.texels = nullptr;
  // per-face materials (B2ptScene::face_material): one device copy of the ids, every mesh points at its faces
  int* face_mat_dev = nullptr;
  if (sc->face_material && sc->n_faces > 0) {
    for (int f = 0; f < sc->n_faces; ++f)
      if (sc->face_material[f] < 0 || sc->face_material[f] >= sc->n_materials)
        return fail(B2PT_ERR_INVALID, "face_material entry out of range");
    if ((rc = c->dalloc(&face_mat_dev, (size_t)sc->n_faces))) return rc;
    CK(cudaMemcpyAsync(face_mat_dev, sc->face_material, (size_t)sc->n_faces * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  c->geom_mesh.assign(sc->n_geoms, -1);
  for (int g = 0; g < sc->n_geoms; ++g) {
    const B2ptGeom& G = sc->geoms[g];
    DevGeom& D = hg[g];
    D.inv = rows_of(G.inverse_transform);
    D.fwd = rows_of(G.transform);
    D.invT = rows_of(G.inv_transpose);
    D.type = G.type;
    D.material = G.material_id;
    D.mesh = -1;
    D.rigid = is_rigid(G.transform);
    {
      const float lo[3] = {-0.5f, -0.5f, -0.5f}, hi[3] = {0.5f, 0.5f, 0.5f};  // unit cube / r=0.5 sphere
      world_box(G.transform, lo, hi, &D);
    }
    if (G.material_id < 0 || G.material_id >= sc->n_materials)
      return fail(B2PT_ERR_INVALID, "geom material id out of range");
    if (G.type == B2PT_OBJ && G.face_count > 0) {
      if (G.face_begin < 0 || (long long)G.face_begin + G.face_count > sc->n_faces || !sc->face_pos || !sc->face_uv)
        return fail(B2PT_ERR_INVALID, "geom face range out of bounds");
      MeshBuild mb;
      float origin_radius = 0.0f;  // the cube [-origin_world, origin_world]^3 seen from object space
      for (int cn = 0; cn < 8; ++cn) {
        const float p[3] = {(cn & 1) ? origin_world : -origin_world, (cn & 2) ? origin_world : -origin_world,
                            (cn & 4) ? origin_world : -origin_world};
        const float* m = G.inverse_transform;
        for (int r = 0; r < 3; ++r)
          origin_radius = std::max(origin_radius, std::fabs(m[r] * p[0] + m[4 + r] * p[1] + m[8 + r] * p[2] + m[12 + r]));
      }
      if ((rc = build_mesh(c, sc->face_pos + (size_t)G.face_begin * 9, sc->face_uv + (size_t)G.face_begin * 6,
                           G.face_count, origin_radius, &mb)))
        return rc;
      {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        const float* fp = sc->face_pos + (size_t)G.face_begin * 9;
        for (size_t v = 0; v < (size_t)G.face_count * 3; ++v)
          for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], fp[v * 3 + k]);
            hi[k] = std::max(hi[k], fp[v * 3 + k]);
          }
        world_box(G.transform, lo, hi, &D);
      }
      DevMesh M{};
      M.nodes = mb.nodes;
      M.tris = mb.tris;
      M.face_pos = mb.face_pos;
      M.face_uv = mb.face_uv;
      M.face_rec = mb.face_rec;
      M.n_faces = G.face_count;
      M.root = mb.root;
      M.geom = g;
      M.face_mat = face_mat_dev ? face_mat_dev + G.face_begin : nullptr;
      if ((rc = upload_texture(c, sc, G.tex_kd, &M.kd, &tex_cache))) return rc;
      if ((rc = upload_texture(c, sc, G.tex_ks, &M.ks, &tex_cache))) return rc;
      if ((rc = upload_texture(c, sc, G.tex_bump, &M.bump, &tex_cache))) return rc;
      if ((rc = upload_texture(c, sc, G.tex_ke, &M.ke, &tex_cache))) return rc;
      D.mesh = (int)hm.size();
      c->geom_mesh[g] = D.mesh;
      hm.push_back(M);
      c->meshes.push_back(mb);
    }
  }
  if (!c->meshes.empty() && !getenv("B2PT_NO_L2_PERSIST")) {
    size_t best = 0;
    for (size_t k = 1; k < c->meshes.size(); ++k)
      if (c->meshes[k].bvh_bytes > c->meshes[best].bvh_bytes) best = k;
    const size_t want = c->meshes[best].bvh_bytes;
    const size_t cap = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, (size_t)prop.accessPolicyMaxWindowSize);
    if (cap > 0) {
      c->l2_window_ptr = c->meshes[best].nodes;
      c->l2_window_bytes = std::min(want, cap);
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_window_bytes) != cudaSuccess) {
        cudaGetLastError();
        c->l2_window_ptr = nullptr;
      }
      apply_l2_policy(c, c->stream);
    }
  }
  DevGeom* dg = nullptr;
  DevMesh* dm = nullptr;
  DevMaterial* dmat = nullptr;
  if ((rc = c->dalloc(&dg, hg.size()))) return rc;
  if ((rc = c->dalloc(&dm, hm.size()))) return rc;
  if ((rc = c->dalloc(&dmat, (size_t)sc->n_materials))) return rc;
  if (!hg.empty()) CK(cudaMemcpyAsync(dg, hg.data(), hg.size() * sizeof(DevGeom), cudaMemcpyHostToDevice, c->stream));
  if (!hm.empty()) CK(cudaMemcpyAsync(dm, hm.data(), hm.size() * sizeof(DevMesh), cudaMemcpyHostToDevice, c->stream));
  static_assert(sizeof(DevMaterial) == sizeof(B2ptMaterial), "material layout");
  CK(cudaMemcpyAsync(dmat, sc->materials, (size_t)sc->n_materials * sizeof(DevMaterial), cudaMemcpyHostToDevice, c->stream));
  DevTexture* dmt = nullptr;
  if (sc->material_textures) {  // the four maps of every material (kd, ks, bump, ke)
    std::vector<DevTexture> hmt(4 * (size_t)sc->n_materials);
    for (size_t k = 0; k < hmt.size(); ++k)
      if ((rc = upload_texture(c, sc, sc->material_textures[k], &hmt[k], &tex_cache))) return rc;
    if ((rc = c->dalloc(&dmt, hmt.size()))) return rc;
    CK(cudaMemcpyAsync(dmt, hmt.data(), hmt.size() * sizeof(DevTexture), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // hmt is a local
  }
  c->dscene.geoms = dg;
  c->dscene.meshes = dm;
  c->dscene.materials = dmat;
  c->dscene.mat_tex = dmt;
  c->dscene.n_geoms = sc->n_geoms;
  c->dscene.n_meshes = (int)hm.size();
  c->dscene.n_materials = sc->n_materials;
  c->alloc_to_scene = false;
  }  // !parent

  // ---- wavefront buffers -------------------------------------------------------------
  const size_t P = (size_t)c->P;
  for (int b = 0; b < 2; ++b) {
    if ((rc = c->dalloc(&c->buf[b].s0, P))) return rc;
    if ((rc = c->dalloc(&c->buf[b].s1, P))) return rc;
    if ((rc = c->dalloc(&c->buf[b].s2, P))) return rc;
  }
  for (int b = 0; b < 2; ++b) {
    if ((rc = c->dalloc(&c->hits[b].h0, P))) return rc;
    if ((rc = c->dalloc(&c->hits[b].h1, P))) return rc;
    if ((rc = c->dalloc(&c->key[b], P + kSortTile))) return rc;   // k_sort_material_few reads whole 16-byte groups
    if ((rc = c->dalloc(&c->live[b], P + kSortTile))) return rc;  // k_rank_live reads whole 16-byte groups
  }
  if ((rc = c->dalloc(&c->perm, P))) return rc;
  if ((rc = c->dalloc(&c->apos, P + kSortTile))) return rc;  // ... and writes whole int4s
  if ((rc = c->dalloc(&c->mesh_queue, 3 * P))) return rc;
  c->long_cap = (int)std::max<size_t>(P / 4, 4096);
  if ((rc = c->dalloc(&c->long_queue, (size_t)c->long_cap))) return rc;
  if ((rc = c->dalloc(&c->long_best, (size_t)c->long_cap))) return rc;
  if ((rc = c->dalloc(&c->long_stack, (size_t)c->long_cap * kLongCarry))) return rc;
  if ((rc = c->dalloc(&c->long_n, (size_t)c->long_cap))) return rc;
  if ((rc = c->dalloc(&c->ctr, 1))) return rc;
  if ((rc = c->dalloc(&c->iter_state, 4))) return rc;
  c->sort_grid = (int)((P + kSortTile - 1) / kSortTile);
  c->shade_grid = (int)((P + kShadeThreads - 1) / kShadeThreads);
  c->gen_grid = (int)std::min<size_t>((P + 255) / 256, (size_t)c->sm_count * 8);
  if ((rc = c->dalloc(&c->sort_status, (size_t)c->sort_grid * 256))) return rc;
  if ((rc = c->dalloc(&c->sort_status_live, (size_t)c->sort_grid * 256))) return rc;
  if ((rc = c->dalloc(&c->image, P * 3))) return rc;
  if ((rc = c->dalloc(&c->albedo, P * 3))) return rc;
  c->image_target = c->image;
  CK(cudaMemsetAsync(c->ctr, 0, sizeof(Counters), c->stream));
  CK(cudaMemsetAsync(c->iter_state, 0, 4 * sizeof(int), c->stream));
  CK(cudaMemsetAsync(c->sort_status, 0, (size_t)c->sort_grid * 256 * 8, c->stream));
  CK(cudaMemsetAsync(c->sort_status_live, 0, (size_t)c->sort_grid * 256 * 8, c->stream));
  CK(cudaMemsetAsync(c->image, 0, P * 12, c->stream));
  CK(cudaMemsetAsync(c->albedo, 0, P * 12, c->stream));
  c->fb_enabled = opt.cache_first_bounce && !opt.antialiasing && !opt.depth_of_field;
  if (c->fb_enabled) {
    if ((rc = c->dalloc(&c->fb_hits.h0, P))) return rc;
    if ((rc = c->dalloc(&c->fb_hits.h1, P))) return rc;
    if ((rc = c->dalloc(&c->fb_key, P))) return rc;
    if ((rc = c->dalloc(&c->fb_live, P + kSortTile))) return rc;
    if ((rc = c->dalloc(&c->fb_hist, (size_t)2 * kMaxMaterials))) return rc;
  }
  if (opt.record_stages) {
    if ((rc = c->dalloc(&c->rec_s0, P))) return rc;
    if ((rc = c->dalloc(&c->rec_s1, P))) return rc;
    if ((rc = c->dalloc(&c->rec_s2, P))) return rc;
    if ((rc = c->dalloc(&c->rec_live, P))) return rc;
    if ((rc = c->dalloc(&c->rec_dead, P))) return rc;
  }

  if (getenv("B2PT_TRAVERSAL_STATS")) {
    if ((rc = c->dalloc(&c->trav_stats, (size_t)32 * (kMaxDepth + 1)))) return rc;
    CK(cudaMemsetAsync(c->trav_stats, 0, sizeof(unsigned long long) * 32 * (kMaxDepth + 1), c->stream));
  }
  // persistent grid of the intersect kernel: SMs x resident CTAs
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_intersect_mesh_brute, kIsectThreads, 0));
  c->isect_grid = c->sm_count * std::max(occ, 1);
  if (c->trav_stats)
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mesh_walk<true>, kWalkThreads, 0));
  else
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mesh_walk<false>, kWalkThreads, 0));
  c->walk_grid = c->sm_count * std::max(occ, 1);
  const int walk_occ = std::max(occ, 1);  // B2PT_WALK_CTAS is applied below, after the shared-SM sizing
  c->finish_grid = c->sm_count * 8;
  c->shade_stride_grid = c->sm_count * 12;
  c->gen_trace_grid = (int)std::min<size_t>((P + 255) / 256, (size_t)c->sm_count * 8);
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mesh_walk_long<32>, kCoopThreads, 0));
  c->long_grid = c->sm_count * std::max(occ, 1);
  // the fused kernels shorten the chain of one context (1.435 -> 1.39 ms) but cost 2 % of aggregate throughput
  // when several contexts share the GPU (their two phases serialise inside a CTA)
  c->fuse = opt.concurrent_contexts <= 1;
  if (const char* e = getenv("B2PT_FUSE")) c->fuse = atoi(e) != 0;
  if (const char* e = getenv("B2PT_SORT_GENERAL")) c->sort_general = atoi(e) != 0;
  if (const char* e = getenv("B2PT_LONG_CARRY")) c->long_carry = std::max(0, std::min(atoi(e), kLongCarry));  // tests: 0 = always restart at the root
  if (const char* e = getenv("B2PT_LONG_CAP")) c->long_cap = std::max(1, std::min(atoi(e), c->long_cap));      // tests: a full hand-off queue
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_intersect_analytic, 256, 0));
  c->analytic_grid = (int)std::min<size_t>((P + 255) / 256, (size_t)c->sm_count * std::max(occ, 1) * 2);
  if (opt.concurrent_contexts > 1) {
    // several contexts render on this GPU at once: a share of the SMs each, so that their kernels co-reside
    // and fill each other's tails (measured: 4 contexts, 1.18 -> 1.03 ms per iteration in aggregate)
    c->analytic_grid = std::min(c->analytic_grid, c->sm_count * 2);
    c->walk_grid = c->sm_count;
    c->long_lanes = 16;
    c->long_walk = 32;  // with the SMs shared, throughput counts, not the length of the launch: 24 -> 32 steps is -1.4 % per iteration
    c->long_grid = c->sm_count * 2;
    c->finish_grid = c->sm_count * 2;
    c->shade_stride_grid = c->sm_count * 4;
  }
  // experiment knobs: resident CTAs per SM of the persistent / grid-stride kernels
  if (const char* e = getenv("B2PT_LONG_WALK")) c->long_walk = std::max(atoi(e), 1);
  if (const char* e = getenv("B2PT_LONG_LANES")) c->long_lanes = atoi(e) == 16 ? 16 : 32;
  if (const char* e = getenv("B2PT_WALK_CTAS")) c->walk_grid = c->sm_count * std::max(1, std::min(atoi(e), walk_occ));
  if (const char* e = getenv("B2PT_LONG_CTAS")) c->long_grid = c->sm_count * std::max(1, atoi(e));
  if (const char* e = getenv("B2PT_ANALYTIC_CTAS")) c->analytic_grid = c->sm_count * std::max(1, atoi(e));
  if (const char* e = getenv("B2PT_FINISH_CTAS")) c->finish_grid = c->sm_count * std::max(1, atoi(e));
  if (const char* e = getenv("B2PT_SHADE_CTAS")) c->shade_stride_grid = c->sm_count * std::max(1, atoi(e));

  c->gen.cam = to_dev_camera(sc->camera);
  c->gen.trace_depth = sc->trace_depth;
  c->gen.antialiasing = opt.antialiasing;
  c->gen.depth_of_field = opt.depth_of_field;
  c->gen.lens_radius = opt.lens_radius;
  c->gen.focal_distance = opt.focal_distance;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" void b2pt_destroy(B2ptCtx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->trav_stats) {
    std::vector<unsigned long long> h((size_t)32 * (kMaxDepth + 1));
    cudaMemcpy(h.data(), c->trav_stats, h.size() * 8, cudaMemcpyDeviceToHost);
    for (int d = 0; d < c->loop_depth; ++d) {
      const unsigned long long* s = &h[(size_t)32 * d];
      if (!s[0]) continue;
      fprintf(stderr, "[b2pt traversal] depth %d: %llu walks, node steps/walk %.2f, leaf steps/walk %.2f; per warp: %llu warps, "
                      "loop iterations avg %.1f max %llu, refill passes avg %.1f, lanes per node step %.1f\n",
              d, s[0], (double)s[1] / s[0], (double)s[2] / s[0], s[21], s[21] ? (double)s[22] / s[21] : 0.0, s[23],
              s[21] ? (double)s[26] / s[21] : 0.0, s[30] ? (double)s[29] / s[30] : 0.0);
    }
  }
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  if (c->graph) cudaGraphDestroy(c->graph);
  for (void* p : c->allocs) cudaFree(p);
  if (c->ev_loop_a) cudaEventDestroy(c->ev_loop_a);
  if (c->ev_loop_b) cudaEventDestroy(c->ev_loop_b);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

static int create_ctx(const B2ptScene* scene, const B2ptOptions* opt, B2ptCtx* parent, B2ptCtx** out);

extern "C" int b2pt_create(const B2ptScene* scene, const B2ptOptions* opt, B2ptCtx** out) {
  return create_ctx(scene, opt, nullptr, out);
}

extern "C" int b2pt_create_shared(B2ptCtx* parent, const B2ptScene* scene, const B2ptOptions* opt, B2ptCtx** out) {
  if (!parent) return fail(B2PT_ERR_INVALID, "parent is NULL");
  return create_ctx(scene, opt, parent, out);
}

static int create_ctx(const B2ptScene* scene, const B2ptOptions* opt, B2ptCtx* parent, B2ptCtx** out) {
  if (!scene || !out) return fail(B2PT_ERR_INVALID, "scene and out must not be NULL");
  *out = nullptr;
  B2ptCtx* c = new (std::nothrow) B2ptCtx();
  if (!c) return fail(B2PT_ERR_NOMEM, "out of host memory");
  int rc = create_impl(scene, opt, c, parent);
  if (rc != 0) {
    std::string keep = g_last_error;
    b2pt_destroy(c);
    g_last_error = keep;
    return rc;
  }
  *out = c;
  return 0;
}

// ---------------------------------------------------------------------------------
// one iteration
// ---------------------------------------------------------------------------------
__global__ void k_set_iter(int* iter_state, int first, int stride) {
  iter_state[1] = first;
  iter_state[2] = stride;
}

// First-bounce cache: save / restore the depth-0 material histograms (the hit records themselves stay
// in the depth-0 buffers and are simply not recomputed).
__global__ void k_first_bounce_hist(Counters* ctr, unsigned int* saved, int n_paths, int restore) {
  for (int i = threadIdx.x; i < kMaxMaterials; i += blockDim.x) {
    if (restore) {
      ctr->hist[0][i] = saved[i];
      ctr->hist_live[0][i] = saved[kMaxMaterials + i];
    } else {
      saved[i] = ctr->hist[0][i];
      saved[kMaxMaterials + i] = ctr->hist_live[0][i];
    }
  }
  if (restore && threadIdx.x == 0) atomicAdd(&ctr->segments, (unsigned long long)n_paths);
}

template <int TRIG>
static void launch_generate(B2ptCtx* c, const IsectParams* next) {
  if (next)
    k_generate_trace<TRIG><<<c->gen_trace_grid, 256, 0, c->stream>>>(c->gen, c->iter_state, c->buf[0], *next);
  else
    k_generate<TRIG><<<c->gen_grid, 256, 0, c->stream>>>(c->gen, c->iter_state, c->buf[0]);
}

template <int TRIG, bool RECORD>
static void launch_shade(B2ptCtx* c, const ShadeParams& sp) {
  k_shade_compact<TRIG, RECORD><<<std::min(c->shade_grid, c->shade_stride_grid), kShadeThreads, 0, c->stream>>>(sp);
}

static void unpack3(const std::vector<float4>& v, int n, float* dst) {
  for (int i = 0; i < n; ++i) {
    dst[3 * i] = v[i].x;
    dst[3 * i + 1] = v[i].y;
    dst[3 * i + 2] = v[i].z;
  }
}

// Queue the kernels of one iteration (also the body that gets graph-captured).
struct KernelTimes {  // optional per-launch event timing (b2pt_profile_iteration)
  std::vector<cudaEvent_t> ev;
  std::vector<int> kind;  // kind of the launch that FOLLOWS event i
  cudaStream_t s = nullptr;
  void mark(int k) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
    kind.push_back(k);
  }
};

static int enqueue_iteration(B2ptCtx* c, bool record, bool time_loop, bool capturing = false, KernelTimes* kt = nullptr) {
  cudaStream_t s = c->stream;
  const int slots = c->loop_depth + 1;
  k_iter_begin<<<8, 256, 0, s>>>(c->ctr, c->iter_state, c->P, slots);
  if (kt) kt->mark(0);
  // Fused form (k_fused.cuh): the kernel that produces the rays of depth d+1 also intersects them with the
  // analytic geoms.  Not when the records are read back between the stages, without the compaction ranks of
  // the material sort, or for the depth whose hits the first-bounce cache keeps.
  const bool fused = c->fuse && !record;
  auto next_params = [&](int depth) {
    IsectParams np;
    memset(&np, 0, sizeof np);
    np.scene = c->dscene;
    np.out = c->hits[depth & 1];
    np.key = c->key[depth & 1];
    np.live = c->live[depth & 1];
    np.ctr = c->ctr;
    np.depth = depth;
    np.queue = c->mesh_queue;
    np.queue_cap = c->P;
    return np;
  };
  const bool fused_gen = fused && !c->fb_enabled;
  {
    const IsectParams np = next_params(0);
    if (c->opt.trig_mode == B2PT_TRIG_PORTABLE) launch_generate<1>(c, fused_gen ? &np : nullptr);
    else launch_generate<0>(c, fused_gen ? &np : nullptr);
  }
  c->launches += 2;
  if (time_loop) CK(cudaEventRecordWithFlags(c->ev_loop_a, s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  if (record) c->records.assign(c->loop_depth, StageRecord());
  for (int d = 0; d < c->loop_depth; ++d) {
    PathBuf in = c->buf[d & 1], out = c->buf[(d + 1) & 1];
    StageRecord* R = record ? &c->records[d] : nullptr;
    int n = 0;
    if (record) {
      CK(cudaStreamSynchronize(s));
      CK(cudaMemcpy(&n, &c->ctr->n_live[d], sizeof(int), cudaMemcpyDeviceToHost));
      R->n = n;
      R->in_s0.resize(n); R->in_s1.resize(n); R->in_s2.resize(n);
      CK(cudaMemcpy(R->in_s0.data(), in.s0, (size_t)n * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->in_s1.data(), in.s1, (size_t)n * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->in_s2.data(), in.s2, (size_t)n * 16, cudaMemcpyDeviceToHost));
    }
    const bool fb = c->fb_enabled && d == 0;
    const HitBuf hits_d = fb ? c->fb_hits : c->hits[d & 1];
    uint8_t* const key_d = fb ? c->fb_key : c->key[d & 1];
    uint8_t* const live_d = fb ? c->fb_live : c->live[d & 1];
    const bool analytic_done = d == 0 ? fused_gen : fused;  // by the kernel that produced this depth's rays
    IsectParams ip;
    ip.scene = c->dscene;
    ip.in = in;
    ip.out = hits_d;
    ip.key = key_d;
    ip.live = live_d;
    ip.ctr = c->ctr;
    ip.depth = d;
    ip.stats = c->trav_stats ? c->trav_stats + 32 * d : nullptr;
    if (kt) kt->mark(1);
    ip.queue = c->mesh_queue;
    ip.queue_cap = c->P;
    ip.long_queue = c->long_queue;
    // (several meshes are safe: a ray that is handed off is finished by k_mesh_walk_long, remaining meshes included)
    ip.long_walk = c->long_walk;
    ip.long_cap = c->long_cap;
    ip.long_carry = c->long_carry;
    ip.long_best = c->long_best;
    ip.long_stack = c->long_stack;
    ip.long_n = c->long_n;
    if (fb && c->fb_valid) {
      // the depth-0 records are still in the first-bounce buffers: only the histograms come back
      k_first_bounce_hist<<<1, 256, 0, s>>>(c->ctr, c->fb_hist, c->P, 1);
      c->launches += 1;
    } else {
      if (!analytic_done) {
        k_intersect_analytic<<<c->analytic_grid, 256, 0, s>>>(ip);
        c->launches += 1;
      }
      if (c->dscene.n_meshes > 0) {
        if (c->opt.use_bvh) {
          if (kt) kt->mark(4);
          if (c->trav_stats)
            k_mesh_walk<true><<<c->walk_grid, kWalkThreads, 0, s>>>(ip);
          else
            k_mesh_walk<false><<<c->walk_grid, kWalkThreads, 0, s>>>(ip);
          if (kt) kt->mark(5);
          if (c->long_lanes == 32)
            k_mesh_walk_long<32><<<c->long_grid, kCoopThreads, 0, s>>>(ip);
          else
            k_mesh_walk_long<16><<<c->long_grid, kCoopThreads, 0, s>>>(ip);
          if (kt) kt->mark(6);
          k_mesh_finish<<<c->finish_grid, 256, 0, s>>>(ip);
          c->launches += 3;
        } else {
          k_intersect_mesh_brute<<<c->isect_grid, kIsectThreads, 0, s>>>(ip);
          c->launches += 1;
        }
      }
      if (fb) {
        k_first_bounce_hist<<<1, 256, 0, s>>>(c->ctr, c->fb_hist, c->P, 0);
        c->launches += 1;
      }
    }
    {
      MatSortParams mp;
      mp.key = key_d;
      mp.live = live_d;
      mp.perm = c->perm;
      mp.apos = c->apos;
      mp.ctr = c->ctr;
      mp.status = c->sort_status;
      mp.status_live = c->sort_status_live;
      mp.depth = d;
      if (kt) kt->mark(2);
      if (c->opt.sort_by_material && c->n_materials <= kFewMaterials && !c->sort_general)
        k_sort_material_few<<<c->sort_grid, kSortThreads, 0, s>>>(mp, c->n_materials);
      else if (c->opt.sort_by_material)
        k_sort_material<<<c->sort_grid, kSortThreads, 0, s>>>(mp);
      else  // SORT_BY_MATERIAL 0: slot order is kept, only the compaction ranks are needed
        k_rank_live<<<c->sort_grid, kSortThreads, 0, s>>>(mp);
      c->launches += 1;
    }
    if (record) {
      CK(cudaStreamSynchronize(s));
      R->h0.resize(n); R->h1.resize(n); R->perm.resize(n);
      CK(cudaMemcpy(R->h0.data(), hits_d.h0, (size_t)n * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->h1.data(), hits_d.h1, (size_t)n * 16, cudaMemcpyDeviceToHost));
      if (c->opt.sort_by_material) {
        CK(cudaMemcpy(R->perm.data(), c->perm, (size_t)n * 4, cudaMemcpyDeviceToHost));
      } else {
        for (int i = 0; i < n; ++i) R->perm[i] = i;
      }
    }
    ShadeParams sp;
    sp.scene = c->dscene;
    sp.in = in;
    sp.out = out;
    sp.hits = hits_d;
    sp.perm = c->opt.sort_by_material ? c->perm : nullptr;  // NULL: identity
    sp.apos = c->apos;
    sp.live = live_d;
    sp.ctr = c->ctr;
    sp.image = c->image_target;
    sp.albedo = c->albedo;
    sp.iter_state = c->iter_state;
    sp.depth = d;
    sp.rng_pixel = c->opt.rng_mode == B2PT_RNG_PIXEL;
    sp.rec_s0 = c->rec_s0; sp.rec_s1 = c->rec_s1; sp.rec_s2 = c->rec_s2;
    sp.rec_dead = c->rec_dead; sp.rec_live = c->rec_live;
    const bool portable = c->opt.trig_mode == B2PT_TRIG_PORTABLE;
    if (kt) kt->mark(3);
    if (fused) {
      const IsectParams np = next_params(d + 1);
      const int tiles = (c->P + kFusedThreads - 1) / kFusedThreads;
      const int grid = std::min(tiles, c->shade_stride_grid * (kShadeThreads / kFusedThreads));
      if (portable) k_shade_trace<1><<<grid, kFusedThreads, 0, s>>>(sp, np);
      else k_shade_trace<0><<<grid, kFusedThreads, 0, s>>>(sp, np);
    } else if (record) {
      if (portable) launch_shade<1, true>(c, sp); else launch_shade<0, true>(c, sp);
    } else {
      if (portable) launch_shade<1, false>(c, sp); else launch_shade<0, false>(c, sp);
    }
    c->launches += 1;
    if (record) {
      CK(cudaStreamSynchronize(s));
      int live = 0;
      CK(cudaMemcpy(&live, &c->ctr->n_live[d + 1], sizeof(int), cudaMemcpyDeviceToHost));
      R->n_live_out = live;
      R->sh_s0.resize(n); R->sh_s1.resize(n); R->sh_s2.resize(n);
      R->live.resize(live); R->dead.resize(n - live);
      CK(cudaMemcpy(R->sh_s0.data(), c->rec_s0, (size_t)n * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->sh_s1.data(), c->rec_s1, (size_t)n * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->sh_s2.data(), c->rec_s2, (size_t)n * 16, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->live.data(), c->rec_live, (size_t)live * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(R->dead.data(), c->rec_dead, (size_t)(n - live) * 4, cudaMemcpyDeviceToHost));
    }
  }
  if (time_loop) CK(cudaEventRecordWithFlags(c->ev_loop_b, s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  if (kt) kt->mark(-1);
  CK(cudaGetLastError());
  if (c->fb_enabled && !capturing) c->fb_valid = true;
  if (record) {
    // the survival flags k_intersect predicted must be what the shade then did
    unsigned int mism = 0;
    CK(cudaStreamSynchronize(s));
    CK(cudaMemcpy(&mism, &c->ctr->pred_mismatch, sizeof mism, cudaMemcpyDeviceToHost));
    if (mism) return fail(B2PT_ERR_STATE, "internal: survival prediction disagreed with the shade kernel");
  }
  return 0;
}

static int profile_impl(B2ptCtx* c, int32_t iter, float* kinds, int n_kinds, float* total) {
  CK(cudaSetDevice(c->device));
  if (iter <= 1) c->albedo_version += 1;
  k_set_iter<<<1, 1, 0, c->stream>>>(c->iter_state, iter, 1);
  c->launches += 1;
  KernelTimes kt;
  kt.s = c->stream;
  cudaEvent_t e0;
  CK(cudaEventCreate(&e0));
  CK(cudaEventRecord(e0, c->stream));
  int rc = enqueue_iteration(c, false, true, false, &kt);
  if (rc == 0) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail(B2PT_ERR_CUDA, cudaGetErrorString(e));
  }
  for (int k = 0; k < n_kinds; ++k) kinds[k] = 0.0f;
  *total = 0.0f;
  if (rc == 0) {
    for (size_t i = 0; i + 1 < kt.ev.size(); ++i) {
      float t = 0.0f;
      cudaEventElapsedTime(&t, kt.ev[i], kt.ev[i + 1]);
      if (kt.kind[i] >= 0 && kt.kind[i] < n_kinds) kinds[kt.kind[i]] += t;
    }
    cudaEventElapsedTime(total, e0, kt.ev.back());
  }
  for (cudaEvent_t e : kt.ev) cudaEventDestroy(e);
  cudaEventDestroy(e0);
  c->loop_timed = true;
  return rc;
}

extern "C" int b2pt_profile_iteration(B2ptCtx* c, int32_t iter, float ms[5]) {
  if (!c || !ms) return fail(B2PT_ERR_INVALID, "ctx and ms must not be NULL");
  float k[7], total;
  int rc = profile_impl(c, iter, k, 7, &total);
  ms[0] = k[0];
  ms[1] = k[1] + k[4] + k[5] + k[6];
  ms[2] = k[2];
  ms[3] = k[3];
  ms[4] = total;
  return rc;
}

extern "C" int b2pt_profile_kernels(B2ptCtx* c, int32_t iter, float ms[B2PT_PROF_COUNT]) {
  if (!c || !ms) return fail(B2PT_ERR_INVALID, "ctx and ms must not be NULL");
  float k[7], total;
  int rc = profile_impl(c, iter, k, 7, &total);
  ms[B2PT_PROF_GENERATE] = k[0];
  ms[B2PT_PROF_ANALYTIC] = k[1];
  ms[B2PT_PROF_WALK] = k[4];
  ms[B2PT_PROF_WALK_LONG] = k[5];
  ms[B2PT_PROF_FINISH] = k[6];
  ms[B2PT_PROF_SORT] = k[2];
  ms[B2PT_PROF_SHADE] = k[3];
  ms[B2PT_PROF_ITERATION] = total;
  return rc;
}

extern "C" int b2pt_render(B2ptCtx* c, int32_t iter_first, int32_t iter_count, int32_t iter_stride) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  if (iter_count < 0 || iter_stride <= 0) return fail(B2PT_ERR_INVALID, "iter_count >= 0 and iter_stride > 0 required");
  if (iter_count == 0) return 0;
  CK(cudaSetDevice(c->device));
  if (iter_first <= 1) c->albedo_version += 1;  // iteration 1 writes the albedo AOV
  k_set_iter<<<1, 1, 0, c->stream>>>(c->iter_state, iter_first, iter_stride);
  c->launches += 1;
  const bool record = c->opt.record_stages != 0;
  if (record || !c->opt.use_graph) {
    for (int i = 0; i < iter_count; ++i) {
      int rc = enqueue_iteration(c, record && i == iter_count - 1, i == iter_count - 1);
      if (rc) return rc;
    }
    c->loop_timed = true;
    return 0;
  }
  if (c->fb_enabled && !c->fb_valid) {
    // the iteration that fills the first-bounce cache runs outside the graph; the graph replays the cached form
    int rc = enqueue_iteration(c, false, iter_count == 1);
    if (rc) return rc;
    if (--iter_count == 0) {
      c->loop_timed = true;
      return 0;
    }
  }
  if (!c->graph_exec) {
    const int64_t before = c->launches;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_iteration(c, false, true, true);
    cudaError_t e = cudaStreamEndCapture(c->stream, &c->graph);
    if (rc || e != cudaSuccess) {  // do not keep a half-captured graph
      if (c->graph) cudaGraphDestroy(c->graph);
      c->graph = nullptr;
      c->launches = before;
      if (rc) return rc;
      return fail(B2PT_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    }
    c->graph_kernels = (int)(c->launches - before);
    c->launches = before;
    CK(cudaGraphInstantiate(&c->graph_exec, c->graph, 0));
  }
  for (int i = 0; i < iter_count; ++i) {
    CK(cudaGraphLaunch(c->graph_exec, c->stream));
    c->launches += c->graph_kernels;
  }
  c->loop_timed = true;
  return 0;
}

extern "C" int b2pt_sync(B2ptCtx* c) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int b2pt_read_accum(B2ptCtx* c, float* image_host, float* albedo_host) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  const size_t bytes = (size_t)c->P * 12;
  if (image_host) CK(cudaMemcpyAsync(image_host, c->image_target, bytes, cudaMemcpyDeviceToHost, c->stream));
  if (albedo_host) CK(cudaMemcpyAsync(albedo_host, c->albedo, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int b2pt_pathtrace(B2ptCtx* c, int32_t iter, float* image_host, float* albedo_host) {
  int rc = b2pt_render(c, iter, 1, 1);
  if (rc) return rc;
  // only a caller that vouches for its buffer (persistent_host_albedo) may skip the copy of an unchanged AOV
  const bool albedo_current = c->opt.persistent_host_albedo && albedo_host && albedo_host == c->albedo_host_last &&
                              c->albedo_host_version == c->albedo_version;
  rc = b2pt_read_accum(c, image_host, albedo_current ? nullptr : albedo_host);
  if (rc == 0 && albedo_host) {
    c->albedo_host_last = albedo_host;
    c->albedo_host_version = c->albedo_version;
  }
  return rc;
}

extern "C" int b2pt_reset_accum(B2ptCtx* c) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(c->image_target, 0, (size_t)c->P * 12, c->stream));
  CK(cudaMemsetAsync(c->albedo, 0, (size_t)c->P * 12, c->stream));
  c->albedo_version += 1;
  return 0;
}

extern "C" int b2pt_set_camera(B2ptCtx* c, const B2ptCamera* cam) {
  if (!c || !cam) return fail(B2PT_ERR_INVALID, "ctx and cam must not be NULL");
  if (cam->resolution[0] != c->W || cam->resolution[1] != c->H)
    return fail(B2PT_ERR_INVALID, "the resolution of a context cannot change");
  for (int k = 0; k < 3; ++k)
    if (!(std::fabs(cam->position[k]) <= c->origin_world))
      return fail(B2PT_ERR_RANGE, "camera position is outside the region the BVH boxes were padded for; create a new context");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  c->gen.cam = to_dev_camera(*cam);
  c->fb_valid = false;
  if (c->graph_exec) {  // the camera is baked into the captured kernel arguments
    cudaGraphExecDestroy(c->graph_exec);
    cudaGraphDestroy(c->graph);
    c->graph_exec = nullptr;
    c->graph = nullptr;
  }
  return b2pt_reset_accum(c);
}

extern "C" int b2pt_set_stream(B2ptCtx* c, void* cuda_stream) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  apply_l2_policy(c, c->stream);
  if (c->graph_exec) {  // a captured graph is tied to the stream it was captured on
    cudaGraphExecDestroy(c->graph_exec);
    cudaGraphDestroy(c->graph);
    c->graph_exec = nullptr;
    c->graph = nullptr;
  }
  return 0;
}

extern "C" float* b2pt_device_image(B2ptCtx* c) { return c ? c->image_target : nullptr; }
extern "C" float* b2pt_device_albedo(B2ptCtx* c) { return c ? c->albedo : nullptr; }
extern "C" void* b2pt_stream(B2ptCtx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int64_t b2pt_launch_count(B2ptCtx* c) { return c ? c->launches : 0; }

extern "C" int b2pt_set_device_image(B2ptCtx* c, float* image_dev) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  c->image_target = image_dev ? image_dev : c->image;
  if (c->graph_exec) {
    cudaGraphExecDestroy(c->graph_exec);
    cudaGraphDestroy(c->graph);
    c->graph_exec = nullptr;
    c->graph = nullptr;
  }
  return 0;
}

extern "C" float b2pt_last_loop_ms(B2ptCtx* c) {
  if (!c || !c->loop_timed) return -1.0f;
  cudaSetDevice(c->device);
  if (cudaEventSynchronize(c->ev_loop_b) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, c->ev_loop_a, c->ev_loop_b) != cudaSuccess) return -1.0f;
  return ms;
}

extern "C" int b2pt_live_counts(B2ptCtx* c, int32_t* n_live, int32_t cap) {
  if (!c || !n_live) return fail(B2PT_ERR_INVALID, "ctx and n_live must not be NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  Counters* h = (Counters*)malloc(sizeof(Counters));
  if (!h) return fail(B2PT_ERR_NOMEM, "out of host memory");
  cudaError_t e = cudaMemcpy(h, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost);
  int k = 0;
  if (e == cudaSuccess)
    for (; k < cap && k <= c->loop_depth; ++k) n_live[k] = h->n_live[k];
  free(h);
  if (e != cudaSuccess) return fail(B2PT_ERR_CUDA, cudaGetErrorString(e));
  return k;
}

extern "C" int b2pt_walk_counts(B2ptCtx* c, int32_t* walks, int32_t* long_walks, int32_t cap) {
  if (!c || !walks) return fail(B2PT_ERR_INVALID, "ctx and walks must not be NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  Counters* h = (Counters*)malloc(sizeof(Counters));
  if (!h) return fail(B2PT_ERR_NOMEM, "out of host memory");
  cudaError_t e = cudaMemcpy(h, c->ctr, sizeof(Counters), cudaMemcpyDeviceToHost);
  int k = 0;
  if (e == cudaSuccess)
    for (; k < cap && k < c->loop_depth; ++k) {
      walks[k] = (int32_t)h->mesh_count[k];
      if (long_walks) long_walks[k] = (int32_t)h->long_count[k];
    }
  free(h);
  if (e != cudaSuccess) return fail(B2PT_ERR_CUDA, cudaGetErrorString(e));
  return k;
}

// ---------------------------------------------------------------------------------
// tone map (sendImageToPBO / sendDenosiedImageToPBO, pathtrace.cu:73-116)
// ---------------------------------------------------------------------------------
__global__ void k_tonemap(const float* __restrict__ src, int n, int iter, uchar4* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float r = src[3 * (size_t)i], g = src[3 * (size_t)i + 1], b = src[3 * (size_t)i + 2];
  int cr, cg, cb;
  if (iter > 0) {
    cr = (int)(r / iter * 255.0);
    cg = (int)(g / iter * 255.0);
    cb = (int)(b / iter * 255.0);
  } else {
    cr = (int)(r * 255.0);
    cg = (int)(g * 255.0);
    cb = (int)(b * 255.0);
  }
  uchar4 o;
  o.x = (unsigned char)min(max(cr, 0), 255);
  o.y = (unsigned char)min(max(cg, 0), 255);
  o.z = (unsigned char)min(max(cb, 0), 255);
  o.w = 0;
  dst[i] = o;
}

// used by pipe.cu: the same kernel on a stream of the caller's choice
extern "C" int b2pt_tonemap_on_stream_(B2ptCtx* c, const float* src_dev, int32_t iter, uint8_t* rgba8_dev, void* stream) {
  if (!c || !rgba8_dev || !src_dev) return fail(B2PT_ERR_INVALID, "ctx, src_dev and rgba8_dev must not be NULL");
  CK(cudaSetDevice(c->device));
  k_tonemap<<<(c->P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src_dev, c->P, iter, (uchar4*)rgba8_dev);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int b2pt_tonemap_rgba8(B2ptCtx* c, const float* src_dev, int32_t iter, uint8_t* rgba8_dev) {
  if (!c || !rgba8_dev) return fail(B2PT_ERR_INVALID, "ctx and rgba8_dev must not be NULL");
  CK(cudaSetDevice(c->device));
  k_tonemap<<<(c->P + 255) / 256, 256, 0, c->stream>>>(src_dev ? src_dev : c->image_target, c->P, iter, (uchar4*)rgba8_dev);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------
// output hand-off: saveImage / savePNG (main.cpp:115-165, image.cpp:22-39) and the
// denoiser inputs (main.cpp:189-203)
// ---------------------------------------------------------------------------------
// image::savePNG quantisation: (unsigned char)(glm::clamp(v, 0, 1) * 255.f) per channel,
// v = image / samples (saveImage, main.cpp:126) or the albedo as it is (:129);
// saveImage mirrors x (setPixel(width - 1 - x, y, ...)).
__device__ __forceinline__ unsigned char png_quant(float v) {
  float c = (v < 0.0f) ? 0.0f : v;  // glm::max(x, 0): (x < 0) ? 0 : x
  c = (1.0f < c) ? 1.0f : c;        // glm::min(x, 1): (1 < x) ? 1 : x
  c = c * 255.f;
  return c == c ? (unsigned char)c : (unsigned char)0;  // NaN: the reference's cast is undefined, 0 here
}

__global__ void k_resolve_rgb8(const float* __restrict__ src, int w, int h, float samples, int divide, int mirror_x,
                               unsigned char* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int x = i % w, y = i / w;
  float r = src[3 * (size_t)i], g = src[3 * (size_t)i + 1], b = src[3 * (size_t)i + 2];
  if (divide) {
    r = r / samples;
    g = g / samples;
    b = b / samples;
  }
  const size_t o = 3 * ((size_t)y * w + (mirror_x ? w - 1 - x : x));
  dst[o] = png_quant(r);
  dst[o + 1] = png_quant(g);
  dst[o + 2] = png_quant(b);
}

// inputColor[index] = image[index] / (float)iteration (main.cpp:194-199), Float3 as OIDN's setImage expects.
__global__ void k_resolve_color(const float* __restrict__ src, size_t n3, float iter, float* __restrict__ dst) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i] / iter;
}

static int resolve_rgb8_host(B2ptCtx* c, int32_t aov, int32_t samples, int32_t mirror_x, std::vector<uint8_t>* out) {
  if (aov != B2PT_AOV_IMAGE && aov != B2PT_AOV_ALBEDO) return fail(B2PT_ERR_INVALID, "aov must be B2PT_AOV_IMAGE or B2PT_AOV_ALBEDO");
  if (aov == B2PT_AOV_IMAGE && samples <= 0) return fail(B2PT_ERR_INVALID, "samples must be > 0");
  CK(cudaSetDevice(c->device));
  unsigned char* d = nullptr;
  CK(cudaMalloc(&d, (size_t)c->P * 3));
  const float* src = aov == B2PT_AOV_IMAGE ? c->image_target : c->albedo;
  k_resolve_rgb8<<<(c->P + 255) / 256, 256, 0, c->stream>>>(src, c->W, c->H, (float)samples, aov == B2PT_AOV_IMAGE, mirror_x, d);
  c->launches += 1;
  out->resize((size_t)c->P * 3);
  cudaError_t e = cudaMemcpyAsync(out->data(), d, out->size(), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(B2PT_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

extern "C" int b2pt_resolve_rgb8(B2ptCtx* c, int32_t aov, int32_t samples, int32_t mirror_x, uint8_t* rgb8_host) {
  if (!c || !rgb8_host) return fail(B2PT_ERR_INVALID, "ctx and rgb8_host must not be NULL");
  std::vector<uint8_t> v;
  int rc = resolve_rgb8_host(c, aov, samples, mirror_x, &v);
  if (rc) return rc;
  memcpy(rgb8_host, v.data(), v.size());
  return 0;
}

extern "C" int b2pt_save_png(B2ptCtx* c, int32_t aov, int32_t samples, const char* path) {
  if (!c || !path) return fail(B2PT_ERR_INVALID, "ctx and path must not be NULL");
  std::vector<uint8_t> v;
  int rc = resolve_rgb8_host(c, aov, samples, 1, &v);
  if (rc) return rc;
  const std::string err = b2pt_host::write_png_rgb8(path, c->W, c->H, v.data());
  if (!err.empty()) return fail(B2PT_ERR_IO, err);
  return 0;
}

extern "C" int b2pt_resolve_color(B2ptCtx* c, int32_t iter, float* color_dev, float* color_host) {
  if (!c) return fail(B2PT_ERR_INVALID, "ctx is NULL");
  if (iter <= 0) return fail(B2PT_ERR_INVALID, "iter must be > 0");
  if (!color_dev && !color_host) return fail(B2PT_ERR_INVALID, "color_dev and color_host are both NULL");
  CK(cudaSetDevice(c->device));
  float* d = color_dev;
  if (!d) CK(cudaMalloc(&d, (size_t)c->P * 12));
  k_resolve_color<<<c->sm_count * 4, 256, 0, c->stream>>>(c->image_target, (size_t)c->P * 3, (float)iter, d);
  c->launches += 1;
  cudaError_t e = cudaSuccess;
  if (color_host) {
    e = cudaMemcpyAsync(color_host, d, (size_t)c->P * 12, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  }
  if (!color_dev) cudaFree(d);
  if (e != cudaSuccess) return fail(B2PT_ERR_CUDA, cudaGetErrorString(e));
  return 0;
}

// saveImage + image::saveHDR (main.cpp:115-135,163, image.cpp:41-45): the mirrored image / samples (or the
// albedo as it is) as a Radiance RGBE file.  The division runs on the device (k_resolve_color, IEEE like
// glm's vec3 / float), mirror and encoding on the host (csrc/host/hdr_writer.h).
extern "C" int b2pt_save_hdr(B2ptCtx* c, int32_t aov, int32_t samples, const char* path) {
  if (!c || !path) return fail(B2PT_ERR_INVALID, "ctx and path must not be NULL");
  if (aov != B2PT_AOV_IMAGE && aov != B2PT_AOV_ALBEDO) return fail(B2PT_ERR_INVALID, "aov must be B2PT_AOV_IMAGE or B2PT_AOV_ALBEDO");
  std::vector<float> px((size_t)c->P * 3);
  int rc;
  if (aov == B2PT_AOV_IMAGE) {
    if (samples <= 0) return fail(B2PT_ERR_INVALID, "samples must be > 0");
    rc = b2pt_resolve_color(c, samples, nullptr, px.data());
  } else {
    rc = b2pt_read_accum(c, nullptr, px.data());
  }
  if (rc) return rc;
  std::vector<float> mirrored(px.size());
  for (int y = 0; y < c->H; ++y)
    for (int x = 0; x < c->W; ++x)
      memcpy(&mirrored[((size_t)y * c->W + (c->W - 1 - x)) * 3], &px[((size_t)y * c->W + x) * 3], 12);
  const std::string err = b2pt_host::write_hdr_rgb(path, c->W, c->H, mirrored.data());
  if (!err.empty()) return fail(B2PT_ERR_IO, err);
  return 0;
}

// ---------------------------------------------------------------------------------
// stage dumps
// ---------------------------------------------------------------------------------
extern "C" int64_t b2pt_stage_read(B2ptCtx* c, int32_t depth, int32_t stage, void* dst, int64_t bytes) {
  if (!c || !dst) return fail(B2PT_ERR_INVALID, "ctx and dst must not be NULL");
  if (depth < 0 || depth >= (int)c->records.size()) return fail(B2PT_ERR_RANGE, "no record for that depth");
  const StageRecord& R = c->records[depth];
  const int n = R.n;
  int cols = 1;
  switch (stage) {
    case B2PT_STAGE_RAY_ORIGIN: case B2PT_STAGE_RAY_DIR: case B2PT_STAGE_HIT_NORMAL: case B2PT_STAGE_SHADED_COLOR:
    case B2PT_STAGE_SHADED_ORIGIN: case B2PT_STAGE_SHADED_DIR: cols = 3; break;
    case B2PT_STAGE_HIT_UV: cols = 2; break;
    default: cols = 1;
  }
  const int64_t need = (int64_t)n * cols * 4;
  if (bytes < need) return fail(B2PT_ERR_RANGE, "destination too small");
  float* f = (float*)dst;
  int32_t* q = (int32_t*)dst;
  switch (stage) {
    case B2PT_STAGE_RAY_ORIGIN: unpack3(R.in_s0, n, f); break;
    case B2PT_STAGE_RAY_DIR: unpack3(R.in_s1, n, f); break;
    case B2PT_STAGE_RAY_PIXEL: for (int i = 0; i < n; ++i) memcpy(&q[i], &R.in_s0[i].w, 4); break;
    case B2PT_STAGE_HIT_T: for (int i = 0; i < n; ++i) f[i] = R.h0[i].x; break;
    case B2PT_STAGE_HIT_NORMAL:
      for (int i = 0; i < n; ++i) { f[3 * i] = R.h0[i].y; f[3 * i + 1] = R.h0[i].z; f[3 * i + 2] = R.h0[i].w; }
      break;
    case B2PT_STAGE_HIT_UV: for (int i = 0; i < n; ++i) { f[2 * i] = R.h1[i].x; f[2 * i + 1] = R.h1[i].y; } break;
    case B2PT_STAGE_HIT_GEOM:
      for (int i = 0; i < n; ++i) { int gm; memcpy(&gm, &R.h1[i].z, 4); q[i] = (int16_t)(gm & 0xffff); }
      break;
    case B2PT_STAGE_HIT_FACE: for (int i = 0; i < n; ++i) memcpy(&q[i], &R.h1[i].w, 4); break;
    case B2PT_STAGE_HIT_MATERIAL:
      for (int i = 0; i < n; ++i) { int gm; memcpy(&gm, &R.h1[i].z, 4); q[i] = (gm >> 16) & 0xffff; }
      break;
    case B2PT_STAGE_SORT_PERM: memcpy(q, R.perm.data(), (size_t)n * 4); break;
    case B2PT_STAGE_SHADED_COLOR: unpack3(R.sh_s2, n, f); break;
    case B2PT_STAGE_SHADED_BOUNCES: for (int i = 0; i < n; ++i) memcpy(&q[i], &R.sh_s1[i].w, 4); break;
    case B2PT_STAGE_SHADED_ORIGIN: unpack3(R.sh_s0, n, f); break;
    case B2PT_STAGE_SHADED_DIR: unpack3(R.sh_s1, n, f); break;
    case B2PT_STAGE_PARTITION_PIXEL:
      memcpy(q, R.live.data(), R.live.size() * 4);
      memcpy(q + R.live.size(), R.dead.data(), R.dead.size() * 4);
      break;
    default: return fail(B2PT_ERR_INVALID, "unknown stage id");
  }
  return need;
}

extern "C" int b2pt_bvh_info(B2ptCtx* c, int32_t geom, B2ptBvhInfo* info) {
  if (!c || !info) return fail(B2PT_ERR_INVALID, "ctx and info must not be NULL");
  if (geom < 0 || geom >= c->n_geoms || c->geom_mesh[geom] < 0) return fail(B2PT_ERR_RANGE, "geom has no BVH");
  *info = c->meshes[c->geom_mesh[geom]].info;
  return 0;
}

// ---------------------------------------------------------------------------------
// standalone primitives on host arrays
// ---------------------------------------------------------------------------------

template <int MODE>
static int scan_family_host(int n, const int* in_host, const uint8_t* flags_host, int* out_host, int* count_out) {
  if (n < 0) return fail(B2PT_ERR_INVALID, "n < 0");
  if (n == 0) { if (count_out) *count_out = 0; return 0; }
  Scratch S;
  int *din = nullptr, *dout = nullptr, *ddead = nullptr, *dcount = nullptr;
  uint8_t* dflags = nullptr;
  unsigned int* ticket = nullptr;
  unsigned long long* status = nullptr;
  const int tiles = (n + kScanTile - 1) / kScanTile;
  if (MODE == 2) { CK(S.get(&dflags, (size_t)n)); CK(cudaMemcpy(dflags, flags_host, (size_t)n, cudaMemcpyHostToDevice)); }
  else { CK(S.get(&din, (size_t)n)); CK(cudaMemcpy(din, in_host, (size_t)n * 4, cudaMemcpyHostToDevice)); }
  CK(S.get(&dout, (size_t)n));
  CK(S.get(&ddead, (size_t)n));
  CK(S.get(&dcount, 1));
  CK(S.get(&ticket, 1));
  CK(S.get(&status, (size_t)tiles));
  CK(cudaMemset(ticket, 0, 4));
  CK(cudaMemset(dcount, 0, 4));
  CK(cudaMemset(status, 0, (size_t)tiles * 8));
  k_scan_family<MODE><<<tiles, kScanThreads>>>(din, dflags, n, dout, ddead, ticket, status, 1u, dcount);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  int count = 0;
  CK(cudaMemcpy(&count, dcount, 4, cudaMemcpyDeviceToHost));
  if (MODE == 0) {
    CK(cudaMemcpy(out_host, dout, (size_t)n * 4, cudaMemcpyDeviceToHost));
  } else if (MODE == 1) {
    CK(cudaMemcpy(out_host, dout, (size_t)count * 4, cudaMemcpyDeviceToHost));
  } else {
    CK(cudaMemcpy(out_host, dout, (size_t)count * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_host + count, ddead, (size_t)(n - count) * 4, cudaMemcpyDeviceToHost));
  }
  if (count_out) *count_out = count;
  return 0;
}

extern "C" int b2pt_scan_exclusive_i32(int32_t n, int32_t* out_host, const int32_t* in_host) {
  if (n > 0 && (!out_host || !in_host)) return fail(B2PT_ERR_INVALID, "NULL array");
  return scan_family_host<0>(n, in_host, nullptr, out_host, nullptr);
}

extern "C" int b2pt_compact_nonzero_i32(int32_t n, int32_t* out_host, const int32_t* in_host) {
  if (n > 0 && (!out_host || !in_host)) return fail(B2PT_ERR_INVALID, "NULL array");
  int count = 0;
  int rc = scan_family_host<1>(n, in_host, nullptr, out_host, &count);
  return rc ? rc : count;
}

extern "C" int b2pt_partition_perm(int32_t n, const uint8_t* flags_host, int32_t* perm_host) {
  if (n > 0 && (!flags_host || !perm_host)) return fail(B2PT_ERR_INVALID, "NULL array");
  int count = 0;
  int rc = scan_family_host<2>(n, nullptr, flags_host, perm_host, &count);
  return rc ? rc : count;
}

extern "C" int b2pt_sort_desc_perm(int32_t n, const int32_t* keys_host, int32_t* perm_host) {
  if (n < 0) return fail(B2PT_ERR_INVALID, "n < 0");
  if (n == 0) return 0;
  if (!keys_host || !perm_host) return fail(B2PT_ERR_INVALID, "NULL array");
  std::vector<uint8_t> k8((size_t)n);
  for (int i = 0; i < n; ++i) {
    if (keys_host[i] < 0 || keys_host[i] > 255) return fail(B2PT_ERR_RANGE, "keys must be in [0, 255]");
    k8[i] = (uint8_t)keys_host[i];
  }
  Scratch S;
  uint8_t* dkey = nullptr;
  int* dperm = nullptr;
  Counters* ctr = nullptr;
  unsigned long long* status = nullptr;
  const int tiles = (n + kSortTile - 1) / kSortTile;
  CK(S.get(&dkey, (size_t)n));
  CK(S.get(&dperm, (size_t)n));
  CK(S.get(&ctr, 1));
  CK(S.get(&status, (size_t)tiles * 256));
  CK(cudaMemcpy(dkey, k8.data(), (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemset(ctr, 0, sizeof(Counters)));
  CK(cudaMemset(status, 0, (size_t)tiles * 256 * 8));
  CK(cudaMemcpy(&ctr->n_live[0], &n, 4, cudaMemcpyHostToDevice));
  const unsigned int one = 1;
  CK(cudaMemcpy(&ctr->serial, &one, 4, cudaMemcpyHostToDevice));
  k_key_hist_u8<<<std::min((n + 255) / 256, 1184), 256>>>(dkey, n, &ctr->hist[0][0]);
  MaterialSortPolicy mp;
  mp.key = dkey;
  mp.perm = dperm;
  mp.ctr = ctr;
  mp.status_ = status;
  mp.depth = 0;
  k_onesweep_pass<MaterialSortPolicy><<<tiles, kSortThreads>>>(mp);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(perm_host, dperm, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// hist_live for the stand-alone entry below: counts key[i] where live[i] != 0.
__global__ void k_key_hist_live_u8(const uint8_t* __restrict__ key, const uint8_t* __restrict__ live, int n, unsigned int* hist) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (live[i]) atomicAdd(&hist[key[i]], 1u);
}

extern "C" int b2pt_sort_material_ranks(int32_t n, const uint8_t* material_host, const uint8_t* live_host, int32_t n_materials,
                                        int32_t general, int32_t* perm_host, int32_t* rank_host) {
  if (n < 0 || n_materials <= 0 || n_materials > kMaxMaterials) return fail(B2PT_ERR_INVALID, "bad n or n_materials");
  if (n == 0) return 0;
  if (!material_host || !live_host || !perm_host || !rank_host) return fail(B2PT_ERR_INVALID, "NULL array");
  for (int i = 0; i < n; ++i)
    if (material_host[i] >= n_materials) return fail(B2PT_ERR_RANGE, "material id >= n_materials");
  Scratch S;
  uint8_t *dkey = nullptr, *dlive = nullptr;
  int *dperm = nullptr, *dapos = nullptr;
  Counters* ctr = nullptr;
  unsigned long long *status = nullptr, *status_live = nullptr;
  const int tiles = (n + kSortTile - 1) / kSortTile;
  const size_t padded = (size_t)tiles * kSortTile;  // the kernels read whole 16-byte groups
  CK(S.get(&dkey, padded));
  CK(S.get(&dlive, padded));
  CK(S.get(&dperm, padded));
  CK(S.get(&dapos, padded));
  CK(S.get(&ctr, 1));
  CK(S.get(&status, (size_t)tiles * 256));
  CK(S.get(&status_live, (size_t)tiles * 256));
  CK(cudaMemset(dkey, 0, padded));
  CK(cudaMemset(dlive, 0, padded));
  CK(cudaMemcpy(dkey, material_host, (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dlive, live_host, (size_t)n, cudaMemcpyHostToDevice));
  CK(cudaMemset(ctr, 0, sizeof(Counters)));
  CK(cudaMemset(status, 0, (size_t)tiles * 256 * 8));
  CK(cudaMemset(status_live, 0, (size_t)tiles * 256 * 8));
  CK(cudaMemcpy(&ctr->n_live[0], &n, 4, cudaMemcpyHostToDevice));
  const unsigned int one = 1;
  CK(cudaMemcpy(&ctr->serial, &one, 4, cudaMemcpyHostToDevice));
  const int hg = std::min((n + 255) / 256, 1184);
  k_key_hist_u8<<<hg, 256>>>(dkey, n, &ctr->hist[0][0]);
  k_key_hist_live_u8<<<hg, 256>>>(dkey, dlive, n, &ctr->hist_live[0][0]);
  MatSortParams mp;
  mp.key = dkey;
  mp.live = dlive;
  mp.perm = dperm;
  mp.apos = dapos;
  mp.ctr = ctr;
  mp.status = status;
  mp.status_live = status_live;
  mp.depth = 0;
  if (n_materials <= kFewMaterials && !general)
    k_sort_material_few<<<tiles, kSortThreads>>>(mp, n_materials);
  else
    k_sort_material<<<tiles, kSortThreads>>>(mp);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(perm_host, dperm, (size_t)n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rank_host, dapos, (size_t)n * 4, cudaMemcpyDeviceToHost));
  int next = 0;
  CK(cudaMemcpy(&next, &ctr->n_live[1], 4, cudaMemcpyDeviceToHost));
  return next;
}

extern "C" int b2pt_radix_sort_pairs_u32(int32_t n, uint32_t* keys_host, uint32_t* vals_host) {
  if (n < 0) return fail(B2PT_ERR_INVALID, "n < 0");
  if (n == 0) return 0;
  if (!keys_host || !vals_host) return fail(B2PT_ERR_INVALID, "NULL array");
  Scratch S;
  uint32_t *dk = nullptr, *dv = nullptr;
  CK(S.get(&dk, (size_t)n));
  CK(S.get(&dv, (size_t)n));
  CK(cudaMemcpy(dk, keys_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, vals_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  RadixTemps rt;
  int rc = radix_temps_alloc(&rt, n);
  if (rc == 0) rc = radix_sort_pairs_dev(dk, dv, n, 0, nullptr, rt);
  if (rc == 0 && cudaStreamSynchronize(0) != cudaSuccess) rc = fail(B2PT_ERR_CUDA, "radix sort failed");
  radix_temps_free(&rt);
  if (rc) return rc;
  CK(cudaMemcpy(keys_host, dk, (size_t)n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vals_host, dv, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}
