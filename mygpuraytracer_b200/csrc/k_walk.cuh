// k_walk.cuh -- the mesh half of computeIntersections (apps/src/pathtrace.cu:303-386,
// apps/src/intersections.h:207-282) as a persistent, warp-synchronous BVH walk.
//
//  k_mesh_walk    Persistent warps drain the queue k_intersect_analytic filled.  Every lane owns one ray and one
//                 walk; the warp advances in warp-uniform steps -- a NODE step (all lanes holding an inner node
//                 test its four child boxes) or a LEAF step (all lanes holding a leaf run the exact test on its
//                 up to kLeafTris triangles), whichever has more lanes waiting -- so the two code paths never
//                 serialise inside one step.  The traversal stack holds (node, entry distance) pairs: the first
//                 kWalkShort entries per lane live in shared memory, the rest spill to local memory; entries that
//                 fell behind the closest hit are dropped at pop time without touching memory.
//                 Lanes whose walk ended are refilled from the queue as soon as kRefillMin of them are idle (first
//                 32 rays per warp static, then one ticket atomic per refill); a queue entry carries the ray itself
//                 (origin, direction, closest analytic hit, survival flag: 48 bytes written by the kernel that
//                 produced the ray), so a refill is the ticket followed by ONE coalesced round trip.
//                 The winner's (t, barycentrics, face, geom) goes straight into the ray's hit record; key and
//                 material histogram are patched.  A walk of more than `long_walk` steps is handed off with its
//                 state (closest hit, stack) so that it cannot hold its warp.
//  k_mesh_walk_long  The handed-off walks, 16 lanes per ray: the lanes of a group take the top entries of a
//                 shared stack each round (a lane tests the four boxes of its node, or the triangles of its
//                 leaf), reduce the closest (t, face) and append the surviving children by a scan.  A ray that
//                 was handed off is FINISHED here: after the mesh it came with, the group walks the ray's
//                 remaining meshes itself, so a ray never has two walks in flight (several OBJ geoms are safe).
//  k_mesh_finish  One thread per queued ray, full warps: rays the mesh won get their uv, geometric normal,
//                 bump-mapped normal and survival flag (intersections.h:226,235-279, interactions.h:171-186).
//
// Parity rules are those of k_intersect.cuh: the BVH only prunes, the leaf test is glm::intersectRayTriangle
// exactly, the closest hit is the lexicographic minimum of (t, face id) within a mesh and of (t, geom id) across
// geoms -- what the reference's strict `<` loops return -- whatever the order triangles are visited in.
//
// Measured and dropped in round 2 (profiles/r02_notes.md, tools/attic/): four lanes per ray (one child box / one
// triangle per lane, per-quad refill): bit-exact, but 118.7 M warp instructions per depth-1 launch against 46.7 M
// here -- the per-ray set-up and fold, which 32 rays share in this kernel, are paid by 8; and a static assignment
// of queue entries to lanes with the next ray prefetched by cp.async into a shared-memory mailbox: refills cost
// nothing any more, but lanes that drew short walks run dry while the queue still holds rays (9.5 instead of 13
// lanes per node step, walk 0.53 -> 0.68 ms).
#pragma once

#include "k_intersect.cuh"

namespace b2pt {

constexpr int kWalkThreads = 256;
#ifndef B2PT_WALK_MINBLOCKS
#define B2PT_WALK_MINBLOCKS 3
#endif
constexpr int kWalkMinBlocks = B2PT_WALK_MINBLOCKS;  // resident CTAs per SM the register budget is set for
#ifndef B2PT_WALK_SORT
#define B2PT_WALK_SORT 5  // comparators of the child sort in a node step (5: full order; 3 / 4: experiments)
#endif
#ifndef B2PT_WALK_SHORT
#define B2PT_WALK_SHORT (B2PT_WIDE == 8 ? 20 : 16)
#endif
constexpr int kWalkShort = B2PT_WALK_SHORT;   // (node, tn) entries per lane in shared memory: deep enough that the
                                              // spill below is the exception (its code stays off the fast path)
constexpr int kWalkStackTotal = kWide == 8 ? 160 : 96;  // (kWide - 1) entries per level of the wide tree: 32 / 22 levels (checked at build time)
constexpr int kWalkSpill = kWalkStackTotal - kWalkShort;   // spill entries per lane in local memory
#ifndef B2PT_REFILL_MIN
#define B2PT_REFILL_MIN 12
#endif
constexpr int kRefillMin = B2PT_REFILL_MIN;  // waiting lanes that trigger a refill pass
constexpr int kWalkMeshes = 4;     // meshes whose geom is staged in shared memory
constexpr int kLongCarry = 32;     // stack entries a long walk carries over to k_mesh_walk_long
constexpr int kWalkDone = 0x7fffffff;
constexpr int kNoGeom = 0x7fffffff;

// Box test of the four children of a wide node with one FMA per plane.
// The ray keeps id = 1/d and noid = -(o * id).  A node stores, per axis, the four
// children's min planes and max planes in one aligned 32-byte pair (k_lbvh.cuh).
// Per axis the plane the ray enters through ("near") is the min plane when
// d >= 0 and the max plane otherwise, so the walk fetches near and far planes
// through three per-ray byte offsets (offx/offy/offz: near; far = off ^ 16) and
// needs no min/max between the two planes of an axis:
//     tn = max(near.x*id.x + noid.x, near.y*.., near.z*.., 0)
//     tf = min(far.x*id.x + noid.x,  far.y*..,  far.z*..,  lim)
// Rounding: id is 1/d to ~1 ulp (MUFU.RCP), o*id is rounded once and the FMA once;
// together that displaces each plane by at most ~2^-21 * max(|o|, |plane|) in
// position space whatever the size of id.  The build pads every box by 2^-19 *
// (largest coordinate a ray origin can have in object space) on top of the
// extent term, so the test stays conservative.
// NaNs (0 * inf, inf - inf for rays parallel to an axis) are dropped by
// fminf / fmaxf: that axis then does not constrain the box.  (The near / far plane is chosen by the SIGN of
// the direction, never by comparing the two products, which is what keeps that true.)
struct WideHit {
  float tn[4];
  bool ok[4];
};
__device__ __forceinline__ void wide_slab_fma(const float4* node, int offx, int offy, int offz, V3 id, V3 noid, float lim,
                                              WideHit* h, float4* cf) {
  const char* nb = reinterpret_cast<const char*>(node);
  const float4 nx = __ldg(reinterpret_cast<const float4*>(nb + offx));
  const float4 fx = __ldg(reinterpret_cast<const float4*>(nb + (offx ^ 16)));
  const float4 ny = __ldg(reinterpret_cast<const float4*>(nb + offy));
  const float4 fy = __ldg(reinterpret_cast<const float4*>(nb + (offy ^ 16)));
  const float4 nz = __ldg(reinterpret_cast<const float4*>(nb + offz));
  const float4 fz = __ldg(reinterpret_cast<const float4*>(nb + (offz ^ 16)));
  *cf = __ldg(node + 6);
#define B2PT_ONE(k, c)                                                                                         \
  {                                                                                                            \
    const float a = fmaxf(fmaxf(fmaf(nx.c, id.x, noid.x), fmaf(ny.c, id.y, noid.y)),                           \
                          fmaxf(fmaf(nz.c, id.z, noid.z), 0.0f));                                              \
    const float b = fminf(fminf(fmaf(fx.c, id.x, noid.x), fmaf(fy.c, id.y, noid.y)),                           \
                          fminf(fmaf(fz.c, id.z, noid.z), lim));                                               \
    h->tn[k] = a;                                                                                              \
    h->ok[k] = a <= b;                                                                                         \
  }
  B2PT_ONE(0, x) B2PT_ONE(1, y) B2PT_ONE(2, z) B2PT_ONE(3, w)
#undef B2PT_ONE
}
__device__ __forceinline__ void slab_offsets(V3 id, int* offx, int* offy, int* offz) {
  *offx = id.x >= 0.0f ? 0 : 16;
  *offy = id.y >= 0.0f ? 32 : 48;
  *offz = id.z >= 0.0f ? 64 : 80;
}

__device__ __forceinline__ bool walk_is_inner(int node) { return (unsigned int)node < 0x40000000u; }

// The ray of a walk in the object space of geom G, and where its pruning starts.
struct WalkRay {
  V3 qo, qd, id, noid;
  float tbest, lim;
};
__device__ __forceinline__ void walk_ray_setup(const DevGeom& G, V3 o, V3 d, float t_min, WalkRay* r) {
  r->qo = xform(G.inv, o, 1.0f);
  r->qd = normalize(xform(G.inv, d, 0.0f));
  r->id = rcp_fast(r->qd);
  r->noid = mk(-(r->qo.x * r->id.x), -(r->qo.y * r->id.y), -(r->qo.z * r->id.z));
  const float t_limit = (G.rigid && t_min < FLT_MAX) ? t_min * 1.0001f + 1e-5f : FLT_MAX;
  r->tbest = t_limit;
  r->lim = t_limit >= FLT_MAX ? FLT_MAX : t_limit * 1.00001f + 1e-6f;
}

// The fold of a finished walk: the mesh is the closest geom so far if (t, geom) beats the record.
__device__ __forceinline__ bool mesh_wins(float tbest, int best, int g, float t_min, int hit) {
  return best >= 0 && tbest > 0.0f && (tbest < t_min || (tbest == t_min && g < hit));
}

// The exact test on the triangles of one leaf: closest (t, face) by lexicographic order.
__device__ __forceinline__ float leaf_exact(const float4* __restrict__ tris, int code, V3 qo, V3 qd, float* u, float* v, int* fid) {
  const int first = leaf_first(code), cnt = leaf_count(code);
  float t = -1.0f;
  *fid = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < kLeafTris; ++k) {
    if (k < cnt) {
      const float4* tp = tris + 3 * (size_t)(first + k);
      const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
      float uu, vv;
      const float tt = tri_exact(qo, qd, mk(ta.x, ta.y, ta.z), mk(tb.x, tb.y, tb.z), mk(tc.x, tc.y, tc.z), &uu, &vv);
      const int ff = __float_as_int(ta.w);
      if (tt >= 0.0f && (t < 0.0f || tt < t || (tt == t && ff < *fid))) {
        t = tt;
        *u = uu;
        *v = vv;
        *fid = ff;
      }
    }
  }
  return t;
}

struct WalkMesh {  // what the walk keeps of a mesh in shared memory
  const float4* nodes;
  const float4* tris;
  const int* face_mat;
  int root, geom;
};

template <bool STATS>
__global__ void __launch_bounds__(kWalkThreads, kWalkMinBlocks) k_mesh_walk(IsectParams p) {
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  __shared__ int smat[kMaxGeoms];            // material of every geom
  __shared__ DevGeom sgeom[kWalkMeshes];     // the geoms of the first kWalkMeshes meshes (the others are read from global memory)
  __shared__ WalkMesh smesh[kWalkMeshes];
  __shared__ int2 sstack[kWalkShort * kWalkThreads];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const unsigned int total = p.ctr->mesh_count[p.depth];
  if (total == 0) return;
  const int n_meshes = p.scene.n_meshes;
  for (int i = tid; i < kMaxMaterials; i += kWalkThreads) {
    shist[i] = 0;
    slive[i] = 0;
  }
  for (int i = tid; i < p.scene.n_geoms; i += kWalkThreads) smat[i] = __ldg(&p.scene.geoms[i].material);
  {
    const int staged = min(n_meshes, kWalkMeshes);
    constexpr int kWords = (int)(sizeof(DevGeom) / 16);
    for (int i = tid; i < staged * kWords; i += kWalkThreads) {
      const int mm = i / kWords, w = i - mm * kWords;
      const float4* src = reinterpret_cast<const float4*>(p.scene.geoms + __ldg(&p.scene.meshes[mm].geom));
      reinterpret_cast<float4*>(sgeom + mm)[w] = __ldg(src + w);
    }
    if (tid < staged) {
      const DevMesh& M = p.scene.meshes[tid];
      smesh[tid].nodes = M.nodes;
      smesh[tid].tris = M.tris;
      smesh[tid].face_mat = M.face_mat;
      smesh[tid].root = M.root;
      smesh[tid].geom = M.geom;
    }
  }
  __syncthreads();
  int2* const sst = sstack + tid;
  int2 spill[kWalkSpill];
  const int long_walk = p.long_walk;

  // ---- lane state ----
  int ray = -1;           // path slot this lane is walking for, -1: none
  int m = 0;              // mesh being walked
  int g = 0;              // its geom
  int node = kWalkDone;   // >= 0 inner node, < 0 leaf code, kWalkDone: walk ended
  int sp = 0;
  const float4* nodes = nullptr;
  const float4* tris = nullptr;
  WalkRay R;
  R.qo = R.qd = R.id = R.noid = mk(0, 0, 1);
  R.tbest = R.lim = 0.0f;
  int offx = 0, offy = 32, offz = 64;
  float bu = 0.0f, bv = 0.0f;
  int best = -1;
  float t_min = FLT_MAX;  // closest hit so far over all geoms (analytic + meshes already walked)
  int hit = kNoGeom;
  int old_material = 0;   // ... and its material (the one the histograms counted this ray under)
  int steps = 0;
  bool handed = false, was_live = false;
  V3 wo = mk(0, 0, 0), wd = mk(0, 0, 1);  // the ray in world space (read again only when there are several meshes)
  unsigned int s_walks = 0, s_nodes = 0, s_leaves = 0, s_iters = 0, s_refills = 0, s_node_iters = 0, s_node_lanes = 0;  // STATS

  unsigned int* head = &p.ctr->ray_ticket[p.depth];
  bool exhausted = false, first_fetch = true;
  const unsigned int static_part = gridDim.x * (kWalkThreads / 32) * 32u;
  const float4* const queue_o = p.queue;                          // (origin, path slot)
  const float4* const queue_d = p.queue + (size_t)p.queue_cap;    // (direction, closest analytic t)
  const float4* const queue_x = p.queue + 2 * (size_t)p.queue_cap;  // (geom | material << 16, survival flag, -, -)

#define B2PT_WALK_PUSH(c, t)                                                            \
  {                                                                                     \
    const int2 e = make_int2((c), __float_as_int(t));                                   \
    if (sp < kWalkShort) sst[sp * kWalkThreads] = e;                                    \
    else if (sp < kWalkShort + kWalkSpill) spill[sp - kWalkShort] = e;                  \
    ++sp;                                                                               \
  }
// Next thing to do: entries that fell behind the closest hit are dropped without touching memory.  The common
// case -- the stack lies in shared memory -- is one LDS and one compare per entry; the spill is a rare side path.
#define B2PT_WALK_POP_NEXT()                                                            \
  {                                                                                     \
    node = kWalkDone;                                                                   \
    while (sp > 0) {                                                                    \
      --sp;                                                                             \
      int2 e;                                                                           \
      if (sp < kWalkShort) e = sst[sp * kWalkThreads];                                  \
      else e = spill[min(sp - kWalkShort, kWalkSpill - 1)];                             \
      if (__int_as_float(e.y) <= R.lim) {                                               \
        node = e.x;                                                                     \
        break;                                                                          \
      }                                                                                 \
    }                                                                                   \
  }

  // Next mesh >= m0 whose box this ray crosses in front of t_min; sets the walk up in its object space.
  auto setup = [&](int m0) -> bool {
    const V3 idw = rcp_fast(wd);
    const V3 noidw = mk(-(wo.x * idw.x), -(wo.y * idw.y), -(wo.z * idw.z));
    for (int mm = m0; mm < n_meshes; ++mm) {
      const bool in_smem = mm < kWalkMeshes;
      const int gg = in_smem ? smesh[mm].geom : p.scene.meshes[mm].geom;
      const DevGeom& G = in_smem ? sgeom[mm] : p.scene.geoms[gg];
      if (!may_beat(G, idw, noidw, t_min, G.rigid != 0)) continue;
      nodes = in_smem ? smesh[mm].nodes : p.scene.meshes[mm].nodes;
      tris = in_smem ? smesh[mm].tris : p.scene.meshes[mm].tris;
      walk_ray_setup(G, wo, wd, t_min, &R);
      slab_offsets(R.id, &offx, &offy, &offz);
      best = -1;
      sp = 0;
      steps = 0;
      node = in_smem ? smesh[mm].root : p.scene.meshes[mm].root;  // < 0: the whole mesh is one leaf
      m = mm;
      g = gg;
      return true;
    }
    return false;
  };

  while (true) {
    const bool can_node = walk_is_inner(node);
    const bool has_leaf = node < 0;
    const unsigned int nm = __ballot_sync(0xffffffffu, can_node);
    const unsigned int lm = __ballot_sync(0xffffffffu, has_leaf);
    const unsigned int busy = nm | lm;
    const bool idle = !can_node && !has_leaf;
    if (STATS) {
      ++s_iters;
      if (nm) {
        ++s_node_iters;
        s_node_lanes += __popc(nm);
      }
    }
    const bool want_refill = busy == 0u || (!exhausted && __popc(~busy) >= kRefillMin);
    if (want_refill) {
      // ---- refill: idle lanes fold their result, move on to the ray's next mesh or start their next ray ----
      if (STATS) ++s_refills;
      if (idle && ray >= 0) {
        if (STATS) ++s_walks;
        if (!handed && mesh_wins(R.tbest, best, g, t_min, hit)) {
          // the mesh is the closest geom so far: (t, barycentrics, face) into the record, k_mesh_finish does the rest
          // (the record this fold replaces is an analytic geom's or an earlier mesh's: its material is in the record)
          const int old_mat = hit == kNoGeom ? 0 : old_material;
          const int mat = mesh_face_material(m < kWalkMeshes ? smesh[m].face_mat : p.scene.meshes[m].face_mat, best, smat[g]);
          if (mat != old_mat) {
            atomicAdd(&shist[mat], 1);
            atomicSub(&shist[old_mat], 1);
          }
          if (was_live) {
            atomicSub(&slive[old_mat], 1);
            p.live[ray] = 0;
            was_live = false;
          }
          p.key[ray] = (uint8_t)mat;
          reinterpret_cast<float*>(p.out.h0 + ray)[0] = R.tbest;
          p.out.h1[ray] = make_float4(bu, bv, __int_as_float((g & 0xffff) | (mat << 16)), __int_as_float(best));
          t_min = R.tbest;
          hit = g;
          old_material = mat;
        }
        // a handed-off ray is finished by k_mesh_walk_long, meshes included
        if (handed || m + 1 >= n_meshes || !setup(m + 1)) ray = -1;
      }
      {
        const bool need = ray < 0;
        const unsigned int need_m = __ballot_sync(0xffffffffu, need);
        if (need_m != 0u && !exhausted) {
          const int cnt = __popc(need_m);
          unsigned int qb = 0;
          if (first_fetch) {
            // every warp's first 32 rays are assigned statically: thousands of warps hitting one ticket word at kernel
            // start is a queue of same-address atomics; the ticket hands out what lies behind the static part
            qb = (unsigned int)(blockIdx.x * (kWalkThreads / 32) + (tid >> 5)) * 32u;
            first_fetch = false;
          } else {
            if (lane == 0) qb = static_part + atomicAdd(head, (unsigned int)cnt);
            qb = __shfl_sync(0xffffffffu, qb, 0);
          }
          if (qb + (unsigned int)cnt >= total) exhausted = true;
          const unsigned int q = qb + (unsigned int)__popc(need_m & ((1u << lane) - 1u));
          if (need && q < total) {
            const float4 a = __ldg(queue_o + q), b = __ldg(queue_d + q), c = __ldg(queue_x + q);
            ray = __float_as_int(a.w);
            wo = mk(a.x, a.y, a.z);
            wd = mk(b.x, b.y, b.z);
            t_min = b.w > 0.0f ? b.w : FLT_MAX;
            hit = b.w > 0.0f ? (__float_as_int(c.x) & 0xffff) : kNoGeom;
            old_material = b.w > 0.0f ? ((__float_as_int(c.x) >> 16) & 0xffff) : 0;
            was_live = __float_as_int(c.y) != 0;
            handed = false;
            if (!setup(0)) ray = -1;
          }
        }
        if (__ballot_sync(0xffffffffu, node != kWalkDone) == 0u && exhausted) break;
      }
      continue;
    }
    // One kind of step per iteration, whichever has more lanes waiting: a NODE step (the four child boxes of
    // every lane's inner node) or a LEAF step (the exact triangle tests), so the two code paths never
    // serialise inside one iteration.
    const bool node_step = __popc(nm) >= __popc(lm);
    if (node_step) {
      if (can_node) {
        if (STATS) ++s_nodes;
        float tn[kWide];
        int ch[kWide];
#pragma unroll
        for (int h = 0; h < kWide / 4; ++h) {  // the same four-slot box test on each 128-byte line of the node
          WideHit wh;
          float4 cf;
          wide_slab_fma(nodes + kNodeF4 * (size_t)node + 8 * h, offx, offy, offz, R.id, R.noid, R.lim, &wh, &cf);
          ch[4 * h + 0] = wh.ok[0] ? __float_as_int(cf.x) : kEmptyChild;
          ch[4 * h + 1] = wh.ok[1] ? __float_as_int(cf.y) : kEmptyChild;
          ch[4 * h + 2] = wh.ok[2] ? __float_as_int(cf.z) : kEmptyChild;
          ch[4 * h + 3] = wh.ok[3] ? __float_as_int(cf.w) : kEmptyChild;
#pragma unroll
          for (int k = 0; k < 4; ++k) tn[4 * h + k] = wh.ok[k] ? wh.tn[k] : FLT_MAX;
        }
#define B2PT_CSWAP(a, b)                                   \
  if (tn[b] < tn[a]) {                                     \
    const float tt = tn[a]; tn[a] = tn[b]; tn[b] = tt;     \
    const int cc = ch[a]; ch[a] = ch[b]; ch[b] = cc;       \
  }
        if (kWide == 4) {
#if B2PT_WALK_SORT == 3   // nearest child only: the other hits go on the stack as they come
          B2PT_CSWAP(0, 1) B2PT_CSWAP(2, 3) B2PT_CSWAP(0, 2)
#elif B2PT_WALK_SORT == 4  // the two nearest in order
          B2PT_CSWAP(0, 1) B2PT_CSWAP(2, 3) B2PT_CSWAP(0, 2) B2PT_CSWAP(1, 2)
#else
          B2PT_CSWAP(0, 1) B2PT_CSWAP(2, 3) B2PT_CSWAP(0, 2) B2PT_CSWAP(1, 3) B2PT_CSWAP(1, 2)
#endif
        } else {
          // eight children: the nearest comes to slot 0 (seven comparators), the others keep their slot order -- a
          // full sort is 19 comparators, and the model of the walk on the renderer's rays prices the missing
          // order at +5 % steps (tools/exp_wide_leaf.c, EXP_UNSORTED)
          B2PT_CSWAP(0, 1) B2PT_CSWAP(2, 3) B2PT_CSWAP(4 % kWide, 5 % kWide) B2PT_CSWAP(6 % kWide, 7 % kWide)
          B2PT_CSWAP(0, 2) B2PT_CSWAP(4 % kWide, 6 % kWide) B2PT_CSWAP(0, 4 % kWide)
        }
#undef B2PT_CSWAP
        if (ch[0] != kEmptyChild) {
          // nearest first; the others go on the stack (farthest first when they are sorted).  With room for all of
          // them in shared memory (the rule) the pushes are predicated stores; the spill path keeps the general form.
          if (sp + (kWide - 1) <= kWalkShort) {
            int2* top = sst + sp * kWalkThreads;
            int pushed = 0;
#pragma unroll
            for (int k = kWide - 1; k >= 1; --k) {
              const bool pk = ch[k] != kEmptyChild;
              if (pk) top[0] = make_int2(ch[k], __float_as_int(tn[k]));
              top += pk ? kWalkThreads : 0;
              pushed += pk ? 1 : 0;
            }
            sp += pushed;
          } else {
#pragma unroll
            for (int k = kWide - 1; k >= 1; --k)
              if (ch[k] != kEmptyChild) B2PT_WALK_PUSH(ch[k], tn[k])
          }
          node = ch[0];
        } else {
          B2PT_WALK_POP_NEXT()
        }
        ++steps;
      }
    } else {
      if (has_leaf) {
        if (STATS) ++s_leaves;
        float u = 0.0f, v = 0.0f;
        int fid;
        const float t = leaf_exact(tris, node, R.qo, R.qd, &u, &v, &fid);
        if (t >= 0.0f && (t < R.tbest || (t == R.tbest && fid < best))) {
          R.tbest = t;
          best = fid;
          bu = u;
          bv = v;
          R.lim = t * 1.00001f + 1e-6f;
        }
        B2PT_WALK_POP_NEXT()
        ++steps;
      }
    }
    // A walk that outgrew its lane is handed to k_mesh_walk_long together with its state: the closest hit so
    // far and the traversal stack plus what the lane was about to do (up to kLongCarry entries; more restarts
    // at the root, still pruned by the carried hit).  One atomic per warp; if the hand-off queue is full the
    // lane simply keeps walking.
    const bool hand = steps > long_walk && node != kWalkDone;
    const unsigned int hm = __ballot_sync(0xffffffffu, hand);
    if (hm != 0u) {
      unsigned int lb = 0;
      if (lane == __ffs(hm) - 1) lb = atomicAdd(&p.ctr->long_count[p.depth], (unsigned int)__popc(hm));
      lb = __shfl_sync(0xffffffffu, lb, __ffs(hm) - 1);
      if (hand) {
        const unsigned int lq = lb + (unsigned int)__popc(hm & ((1u << lane) - 1u));
        if (lq < (unsigned int)p.long_cap) {
          p.long_queue[lq] = make_int2(ray, m);
          p.long_best[lq] = make_float4(R.tbest, bu, bv, __int_as_float(best));
          int2* ent = p.long_stack + (size_t)lq * kLongCarry;
          int ne = -1;
          if (sp + 1 <= p.long_carry) {
            for (int k = 0; k < sp; ++k) ent[k] = k < kWalkShort ? sst[k * kWalkThreads] : spill[k - kWalkShort];
            ent[sp] = make_int2(node, __float_as_int(0.0f));
            ne = sp + 1;
          }
          p.long_n[lq] = ne;
          node = kWalkDone;
          handed = true;
        } else {
          steps = -0x40000000;
        }
      }
    }
  }
#undef B2PT_WALK_PUSH
#undef B2PT_WALK_POP_NEXT
  if (STATS) {
    atomicAdd(&p.stats[0], (unsigned long long)s_walks);
    atomicAdd(&p.stats[1], (unsigned long long)s_nodes);
    atomicAdd(&p.stats[2], (unsigned long long)s_leaves);
    if (lane == 0) {
      atomicAdd(&p.stats[21], 1ull);
      atomicAdd(&p.stats[22], (unsigned long long)s_iters);
      atomicMax(&p.stats[23], (unsigned long long)s_iters);
      atomicAdd(&p.stats[26], (unsigned long long)s_refills);
      atomicAdd(&p.stats[29], (unsigned long long)s_node_lanes);
      atomicAdd(&p.stats[30], (unsigned long long)s_node_iters);
    }
  }
  __syncthreads();
  for (int i = tid; i < kMaxMaterials; i += kWalkThreads) {
    const int c = shist[i];
    if (c) atomicAdd(&p.ctr->hist[p.depth][i], (unsigned int)c);
    const int cl = slive[i];
    if (cl) atomicAdd(&p.ctr->hist_live[p.depth][i], (unsigned int)cl);
  }
}

// Long walks, kCoopGroup lanes per ray (two rays per warp).  A group shares one stack of (node, entry distance)
// pairs in shared memory, seeded with the state k_mesh_walk handed over; each round its lanes take the top entries:
// a lane with an inner node tests its four child boxes, a lane with a leaf runs the exact test on its triangles,
// the closest (t, face) of the round is reduced across the group and the surviving children are appended with a
// segmented scan.  The visiting order is no longer nearest-first, which is harmless: the winner is the
// lexicographic minimum of (t, face id) whatever the order, and entries behind it are dropped.  Close to the
// capacity of the stack a group falls back to one entry per round (depth-first, growth <= 3 per level).  The loop is
// warp-uniform: the groups of a warp run their rounds in lockstep and refill independently.  When the mesh a ray
// came with is done, the group folds the result and walks the ray's remaining meshes from their roots.
constexpr int kCoopThreads = 128;
// Lanes per long walk: 16 when several contexts share the SMs (two rays per warp), 32 when a context has the GPU to
// itself -- there a launch ends with its longest walk, and a whole warp per ray shortens it (long walks 0.277 ->
// 0.250 ms per iteration full-width; with four contexts 32 lanes cost 1 % of the aggregate).  kCoopGroup is the
// template argument of the kernel; the constants below depend on it.
template <int kCoopGroup>
__global__ void __launch_bounds__(kCoopThreads) k_mesh_walk_long(IsectParams p) {
  constexpr int kCoopCap = 32 * kCoopGroup;                                           // stack entries per group (32 KB per CTA in all)
  constexpr int kCoopDfs = kCoopCap - kWalkStackTotal - (kWide - 1) * kCoopGroup;     // above this only one entry per round is taken
  constexpr unsigned int kCoopMask = kCoopGroup == 32 ? 0xffffffffu : ((1u << (kCoopGroup & 31)) - 1u);
  __shared__ int2 cstack[(kCoopThreads / kCoopGroup) * kCoopCap];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (kCoopGroup - 1);           // lane within the group
  const int gfirst = lane & ~(kCoopGroup - 1);      // first lane of the group
  int2* const st = cstack + (threadIdx.x / kCoopGroup) * kCoopCap;
  const unsigned int total = min(p.ctr->long_count[p.depth], (unsigned int)p.long_cap);
  unsigned int* head = &p.ctr->long_ticket[p.depth];
  const int n_meshes = p.scene.n_meshes;

  // ---- group state (replicated in the lanes of the group) ----
  int ray = -1, m = 0, g = 0, hit = kNoGeom, old_material = 0, sp = 0, best = -1;
  float t_min = FLT_MAX, bu = 0.0f, bv = 0.0f;
  WalkRay R;
  R.qo = R.qd = R.id = R.noid = mk(0, 0, 1);
  R.tbest = R.lim = 0.0f;
  V3 wo = mk(0, 0, 0), wd = mk(0, 0, 1);  // the ray in world space
  int offx = 0, offy = 32, offz = 64;
  const float4* nodes = nullptr;
  const float4* tris = nullptr;
  bool exhausted = false;

  while (true) {
    if (sp == 0) {  // group-uniform: fold the finished walk, go on with the ray's next mesh or fetch the next job
      if (ray >= 0) {
        if (mesh_wins(R.tbest, best, g, t_min, hit)) {
          const int old_mat = hit == kNoGeom ? 0 : old_material;
          const int mat = mesh_face_material(p.scene.meshes[m].face_mat, best, p.scene.geoms[g].material);
          if (gl == 0) {
            if (mat != old_mat) {
              atomicAdd(&p.ctr->hist[p.depth][mat], 1u);
              atomicSub(&p.ctr->hist[p.depth][old_mat], 1u);
            }
            if (p.live[ray]) {
              atomicSub(&p.ctr->hist_live[p.depth][old_mat], 1u);
              p.live[ray] = 0;
            }
            p.key[ray] = (uint8_t)mat;
            reinterpret_cast<float*>(p.out.h0 + ray)[0] = R.tbest;
            p.out.h1[ray] = make_float4(bu, bv, __int_as_float((g & 0xffff) | (mat << 16)), __int_as_float(best));
          }
          t_min = R.tbest;
          hit = g;
          old_material = mat;
        }
        // the ray's remaining meshes (k_mesh_walk dropped the ray when it handed the walk off)
        bool more = false;
        const V3 idw = rcp_fast(wd);
        const V3 noidw = mk(-(wo.x * idw.x), -(wo.y * idw.y), -(wo.z * idw.z));
        for (int mm = m + 1; mm < n_meshes && !more; ++mm) {
          const DevMesh& M = p.scene.meshes[mm];
          const DevGeom& G = p.scene.geoms[M.geom];
          if (!may_beat(G, idw, noidw, t_min, G.rigid != 0)) continue;
          walk_ray_setup(G, wo, wd, t_min, &R);
          slab_offsets(R.id, &offx, &offy, &offz);
          nodes = M.nodes;
          tris = M.tris;
          best = -1;
          m = mm;
          g = M.geom;
          if (gl == 0) st[0] = make_int2(M.root, __float_as_int(0.0f));
          sp = 1;
          more = true;
        }
        if (!more) ray = -1;
      }
      if (ray < 0 && !exhausted) {
        unsigned int q = 0;
        if (gl == 0) q = atomicAdd(head, 1u);
        q = __shfl_sync(kCoopMask << gfirst, q, gfirst);
        if (q >= total) {
          exhausted = true;
        } else {
          const int2 job = p.long_queue[q];
          ray = job.x;
          m = job.y;
          const DevMesh& M = p.scene.meshes[m];
          g = M.geom;
          const DevGeom& G = p.scene.geoms[g];
          const float4 a = p.in.s0[ray];
          const float4 b = p.in.s1[ray];
          const float t0 = reinterpret_cast<const float*>(p.out.h0 + ray)[0];
          const int gm0 = __float_as_int(p.out.h1[ray].z);
          t_min = t0 > 0.0f ? t0 : FLT_MAX;
          hit = t0 > 0.0f ? (gm0 & 0xffff) : kNoGeom;
          old_material = t0 > 0.0f ? ((gm0 >> 16) & 0xffff) : 0;
          wo = mk(a.x, a.y, a.z);
          wd = mk(b.x, b.y, b.z);
          walk_ray_setup(G, wo, wd, t_min, &R);
          slab_offsets(R.id, &offx, &offy, &offz);
          // the state k_mesh_walk handed over: closest triangle so far and (if it fitted) the traversal stack
          const float4 carried = p.long_best[q];
          const int n_carried = p.long_n[q];
          best = __float_as_int(carried.w);
          if (best >= 0) {
            R.tbest = carried.x;
            R.lim = R.tbest * 1.00001f + 1e-6f;
          }
          bu = carried.y;
          bv = carried.z;
          nodes = M.nodes;
          tris = M.tris;
          if (n_carried > 0) {
            for (int k = gl; k < n_carried; k += kCoopGroup) st[k] = p.long_stack[(size_t)q * kLongCarry + k];
            sp = n_carried;
          } else {
            if (gl == 0) st[0] = make_int2(M.root, __float_as_int(0.0f));
            sp = 1;
          }
        }
      }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, sp > 0)) break;

    // ---- one round ----
    const int take = sp > kCoopDfs ? 1 : min(sp, kCoopGroup);
    int node = kWalkDone;
    if (gl < take) {
      const int2 e = st[sp - 1 - gl];
      if (__int_as_float(e.y) <= R.lim) node = e.x;
    }
    sp -= take;
    __syncwarp();
    int ch[kWide];
    float tn[kWide];
#pragma unroll
    for (int k = 0; k < kWide; ++k) {
      ch[k] = kEmptyChild;
      tn[k] = FLT_MAX;
    }
    float t = -1.0f, u = 0.0f, v = 0.0f;
    int fid = 0x7fffffff;
    if (walk_is_inner(node)) {
#pragma unroll
      for (int h = 0; h < kWide / 4; ++h) {
        WideHit wh;
        float4 cf;
        wide_slab_fma(nodes + kNodeF4 * (size_t)node + 8 * h, offx, offy, offz, R.id, R.noid, R.lim, &wh, &cf);
        tn[4 * h + 0] = wh.tn[0]; tn[4 * h + 1] = wh.tn[1]; tn[4 * h + 2] = wh.tn[2]; tn[4 * h + 3] = wh.tn[3];
        ch[4 * h + 0] = wh.ok[0] ? __float_as_int(cf.x) : kEmptyChild;
        ch[4 * h + 1] = wh.ok[1] ? __float_as_int(cf.y) : kEmptyChild;
        ch[4 * h + 2] = wh.ok[2] ? __float_as_int(cf.z) : kEmptyChild;
        ch[4 * h + 3] = wh.ok[3] ? __float_as_int(cf.w) : kEmptyChild;
      }
    } else if (node < 0) {
      t = leaf_exact(tris, node, R.qo, R.qd, &u, &v, &fid);
    }
    // closest (t, face) of this round within the group; t >= 0, so the bit patterns order like the values
    if (__any_sync(0xffffffffu, t >= 0.0f)) {
      unsigned long long key = t >= 0.0f ? ((unsigned long long)__float_as_uint(t) << 32) | (unsigned int)fid : ~0ull;
      unsigned long long mn = key;
#pragma unroll
      for (int o = kCoopGroup / 2; o > 0; o >>= 1) {
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, mn, o, kCoopGroup);
        mn = y < mn ? y : mn;
      }
      const unsigned int winners = __ballot_sync(0xffffffffu, key == mn && mn != ~0ull) & (kCoopMask << gfirst);
      // the shuffles run on the whole warp (this block is warp-uniform); groups without a hit read themselves
      const int src = winners ? __ffs(winners) - 1 : lane;
      const float su = __shfl_sync(0xffffffffu, u, src), sv = __shfl_sync(0xffffffffu, v, src);
      if (winners) {  // group-uniform
        const float tr = __uint_as_float((unsigned int)(mn >> 32));
        const int fr = (int)(unsigned int)mn;
        if (tr < R.tbest || (tr == R.tbest && fr < best)) {
          R.tbest = tr;
          best = fr;
          bu = su;
          bv = sv;
          R.lim = tr * 1.00001f + 1e-6f;
        }
      }
    }
    // append the children that are still in front of the closest hit (segmented scan over the group)
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kWide; ++k) {
      if (ch[k] != kEmptyChild && !(tn[k] <= R.lim)) ch[k] = kEmptyChild;
      cnt += ch[k] != kEmptyChild;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < kCoopGroup; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o, kCoopGroup);
      if (gl >= o) incl += y;
    }
    const int all = __shfl_sync(0xffffffffu, incl, kCoopGroup - 1, kCoopGroup);
    int w = sp + incl - cnt;
#pragma unroll
    for (int k = 0; k < kWide; ++k)
      if (ch[k] != kEmptyChild) st[w++] = make_int2(ch[k], __float_as_int(tn[k]));
    sp += all;
    __syncwarp();
  }
}

// Finish the records of the rays a mesh won: the walks left (t | bu, bv, geom|mat, face).
__global__ void __launch_bounds__(256) k_mesh_finish(IsectParams p) {
  __shared__ int slive[kMaxMaterials];
  const int tid = threadIdx.x;
  const unsigned int total = p.ctr->mesh_count[p.depth];
  if (total == 0) return;
  for (int i = tid; i < kMaxMaterials; i += blockDim.x) slive[i] = 0;
  __syncthreads();
  for (unsigned int q = blockIdx.x * blockDim.x + tid; q < total; q += gridDim.x * blockDim.x) {
    const int i = __float_as_int(p.queue[q].w);
    const float4 h1 = p.out.h1[i];
    const int face = __float_as_int(h1.w);
    if (face < 0) continue;  // an analytic geom (or nothing) is closer
    const int gm = __float_as_int(h1.z);
    const int g = gm & 0xffff, mat = (gm >> 16) & 0xffff;
    const DevGeom& G = p.scene.geoms[g];
    const DevMesh& M = p.scene.meshes[G.mesh];
    const float t = reinterpret_cast<const float*>(p.out.h0 + i)[0];
    V3 nrm;
    float tu, tv;
    mesh_record(G, M, obj_tex(p.scene, M, mat, 2), face, h1.x, h1.y, &nrm, &tu, &tv);
    p.out.h0[i] = make_float4(t, nrm.x, nrm.y, nrm.z);
    p.out.h1[i] = make_float4(tu, tv, h1.z, h1.w);
    // survival: an emissive texel turns the hit into a light (interactions.h:171-186);
    // scatterRay only looks at the emission map in its OBJ branch (not reflective, not refractive)
    bool emissive = false;
    const DevMaterial& mm = p.scene.materials[mat];
    const DevTexture& ke = obj_tex(p.scene, M, mat, 3);
    if (ke.channels && !(__ldg(&mm.has_reflective) > 0) && !(__ldg(&mm.has_refractive) > 0)) {
      const V3 e = fetch_texel(ke, tu, tv);
      emissive = e.x > FLT_EPSILON || e.y > FLT_EPSILON || e.z > FLT_EPSILON;
    }
    const int bounces = __float_as_int(p.in.s1[i].w);
    if (will_survive(p.scene.materials, mat, bounces, emissive)) {
      p.live[i] = 1;
      atomicAdd(&slive[mat], 1);
    }
  }
  __syncthreads();
  for (int i = tid; i < kMaxMaterials; i += blockDim.x) {
    const int cl = slive[i];
    if (cl) atomicAdd(&p.ctr->hist_live[p.depth][i], (unsigned int)cl);
  }
}

}  // namespace b2pt
