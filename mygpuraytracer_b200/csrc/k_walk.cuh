// k_walk.cuh -- the mesh half of computeIntersections (apps/src/pathtrace.cu:303-386,
// apps/src/intersections.h:207-282) as a persistent, warp-synchronous BVH walk.
//
//  k_mesh_walk    Persistent warps drain the queue k_intersect_analytic filled.
//                 Every lane owns one ray and one walk; the warp advances in
//                 warp-uniform steps -- a NODE step (all lanes holding an inner
//                 node test its four child boxes) or a LEAF step (all lanes
//                 holding a triangle run the exact test), whichever has more
//                 lanes waiting -- so the two code paths never serialise inside
//                 one step.  Lanes whose walk ended are refilled from the queue as
//                 soon as kRefillMin of them are idle (one atomic per refill), so
//                 a long walk no longer holds 31 idle lanes hostage (the batch
//                 version ran with 7 of 32 lanes active).  The traversal stack
//                 holds (node, entry distance) pairs: the first kWalkShort entries
//                 per lane live in shared memory, the rest spill to local memory;
//                 entries that fell behind the closest hit are dropped at pop
//                 time without touching memory.
//                 The winner's (t, barycentrics, face, geom) goes straight into
//                 the ray's hit record; key and material histogram are patched.
//                 A walk of more than `long_walk` steps is handed off with its
//                 state (closest hit, stack) so that it cannot hold its CTA.
//  k_mesh_walk_long  The handed-off walks, 16 lanes per ray: the lanes of a group
//                 take the top entries of a shared stack each round, reduce the
//                 closest (t, face) and append the surviving children by a scan.
//  k_mesh_finish  One thread per queued ray, full warps: rays the mesh won get
//                 their uv, geometric normal, bump-mapped normal and survival
//                 flag (intersections.h:226,235-279, interactions.h:171-186).
//
// Parity rules are those of k_intersect.cuh: the BVH only prunes, the leaf test
// is glm::intersectRayTriangle exactly, ties go to the lowest face / geom id.
#pragma once

#include "k_intersect.cuh"

namespace b2pt {

constexpr int kWalkThreads = 256;
constexpr int kWalkMinBlocks = 3;  // resident CTAs per SM the register budget is set for
constexpr int kWalkShort = 8;      // (node, tn) entries per lane in shared memory
constexpr int kWalkSpill = 88;     // spill entries per lane: 96 in all = 3 per level of a 32-level wide tree
constexpr int kRefillMin = 8;      // idle lanes that trigger a refill
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
// fminf / fmaxf: that axis then does not constrain the box.
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

template <bool STATS>
__global__ void __launch_bounds__(kWalkThreads, kWalkMinBlocks) k_mesh_walk(IsectParams p) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  __shared__ int2 sstack[kWalkShort * kWalkThreads];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n_geoms = p.scene.n_geoms;
  const unsigned int total = p.ctr->mesh_count[p.depth];
  if (total == 0) return;
  const long long clk_start = STATS ? clock64() : 0;
  long long clk_loop = 0, clk_refill = 0, clk_node = 0, clk_leaf = 0, clk_handoff = 0;
  unsigned int w_leaf_iters = 0;
  unsigned int w_iters = 0, w_refills = 0, w_node_iters = 0, w_node_lanes = 0;
  unsigned int a_walks = 0, a_nodes = 0, a_tris = 0, a_maxn = 0, a_maxt = 0;
  __shared__ unsigned long long sstat[32];  // per-CTA copy of p.stats, flushed once
  if (STATS && tid < 32) sstat[tid] = 0ull;
  {
    const float4* src = reinterpret_cast<const float4*>(p.scene.geoms);
    float4* dst = reinterpret_cast<float4*>(sgeom);
    const int words = n_geoms * (int)(sizeof(DevGeom) / 16);
    for (int i = tid; i < words; i += kWalkThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < kMaxMaterials; i += kWalkThreads) {
      shist[i] = 0;
      slive[i] = 0;
    }
  }
  __syncthreads();
  unsigned int* head = &p.ctr->ray_ticket[p.depth];
  int2* const sst = sstack + tid;
  int2 spill[kWalkSpill];

  // ---- lane state ----
  int ray = -1;           // path slot this lane is walking for, -1: none
  int g = 0;              // geom being walked
  int node = kWalkDone;   // >= 0 inner node, < 0 ~leaf slot, kWalkDone: walk ended
  int sp = 0;
  const float4* nodes = nullptr;
  const float4* tris = nullptr;
  V3 qo = mk(0, 0, 0), qd = mk(0, 0, 1), id = mk(0, 0, 1);
  V3 noid = mk(0, 0, 0);
  int offx = 0, offy = 32, offz = 64;
  float tbest = 0.0f, lim = 0.0f, bu = 0.0f, bv = 0.0f;
  int best = -1;
  float t_min = FLT_MAX;  // closest hit so far over all geoms (analytic + meshes already walked)
  int hit = kNoGeom;
  unsigned int n_nodes = 0, n_tris = 0;
  int steps = 0;
  bool exhausted = false;
  bool first_fetch = true;
  const unsigned int static_part = gridDim.x * (kWalkThreads / 32) * 32u;
  const int long_walk = p.long_walk;

#define B2PT_WALK_PUSH(c, t)                                                            \
  {                                                                                     \
    const int2 e = make_int2((c), __float_as_int(t));                                   \
    if (sp < kWalkShort) sst[sp * kWalkThreads] = e;                                    \
    else if (sp < kWalkShort + kWalkSpill) spill[sp - kWalkShort] = e;                  \
    ++sp;                                                                               \
  }
// Next thing to do: entries that fell behind the closest hit are dropped without touching memory.
#define B2PT_WALK_POP_NEXT()                                                            \
  {                                                                                     \
    node = kWalkDone;                                                                   \
    while (sp > 0) {                                                                    \
      --sp;                                                                             \
      const int2 e = sp < kWalkShort ? sst[sp * kWalkThreads] : spill[min(sp - kWalkShort, kWalkSpill - 1)]; \
      if (__int_as_float(e.y) <= lim) {                                                 \
        node = e.x;                                                                     \
        break;                                                                          \
      }                                                                                 \
    }                                                                                   \
  }

  // Find the next mesh geom >= g0 whose box this ray crosses in front of t_min and
  // set the walk up in its object space.
  auto setup = [&](V3 o, V3 d, int g0) -> bool {
    const V3 idw = rcp_fast(d);
    const V3 noidw = mk(-(o.x * idw.x), -(o.y * idw.y), -(o.z * idw.z));
    for (int gg = g0; gg < n_geoms; ++gg) {
      const DevGeom& G = sgeom[gg];
      if (G.type != 3 || G.mesh < 0) continue;
      if (!may_beat(G, idw, noidw, t_min, G.rigid != 0)) continue;
      const DevMesh& M = p.scene.meshes[G.mesh];
      nodes = M.nodes;
      tris = M.tris;
      qo = xform(G.inv, o, 1.0f);
      qd = normalize(xform(G.inv, d, 0.0f));
      id = rcp_fast(qd);
      noid = mk(-(qo.x * id.x), -(qo.y * id.y), -(qo.z * id.z));
      slab_offsets(id, &offx, &offy, &offz);
      const float t_limit = (G.rigid && t_min < FLT_MAX) ? t_min * 1.0001f + 1e-5f : FLT_MAX;
      tbest = t_limit;
      lim = t_limit >= FLT_MAX ? FLT_MAX : t_limit * 1.00001f + 1e-6f;
      best = -1;
      sp = 0;
      steps = 0;
      node = M.root;  // < 0: a mesh of one triangle
      g = gg;
      return true;
    }
    return false;
  };

  if (STATS) clk_loop = clock64();
  while (true) {
    const bool can_node = walk_is_inner(node);
    const bool has_leaf = node < 0;
    const unsigned int nm = __ballot_sync(0xffffffffu, can_node);
    const unsigned int lm = __ballot_sync(0xffffffffu, has_leaf);
    const unsigned int busy = nm | lm;
    if (STATS) {
      ++w_iters;
      if (nm) {
        ++w_node_iters;
        w_node_lanes += __popc(nm);
      }
    }
    if (busy == 0u || (!exhausted && __popc(~busy) >= kRefillMin)) {
      const long long clk_r = STATS ? clock64() : 0;
      // ---- refill: idle lanes fold their result, move on to the ray's next mesh or fetch a new ray ----
      if (!can_node && !has_leaf && ray >= 0) {
        if (STATS) {  // lane-local, flushed once at the end
          ++a_walks;
          a_nodes += n_nodes;
          a_tris += n_tris;
          a_maxn = max(a_maxn, n_nodes);
          a_maxt = max(a_maxt, n_tris);
          atomicAdd(&sstat[5 + min(15u, 31u - __clz((n_nodes + n_tris) | 1u))], 1ull);  // log2 histogram of steps
        }
        if (best >= 0 && tbest > 0.0f && (tbest < t_min || (tbest == t_min && g < hit))) {
          // the mesh is the closest geom so far: (t, barycentrics, face) into the record, k_mesh_finish does the rest
          const int old_mat = hit == kNoGeom ? 0 : sgeom[hit].material;
          const int mat = sgeom[g].material;
          if (mat != old_mat) {
            atomicAdd(&shist[mat], 1);
            atomicSub(&shist[old_mat], 1);
          }
          if (p.live[ray]) {
            atomicSub(&slive[old_mat], 1);
            p.live[ray] = 0;
          }
          p.key[ray] = (uint8_t)mat;
          reinterpret_cast<float*>(p.out.h0 + ray)[0] = tbest;
          p.out.h1[ray] = make_float4(bu, bv, __int_as_float((g & 0xffff) | (mat << 16)), __int_as_float(best));
          t_min = tbest;
          hit = g;
        }
        bool more = false;
        if (p.scene.n_meshes > 1) {
          const float4 a = p.in.s0[ray];
          const float4 b = p.in.s1[ray];
          more = setup(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), g + 1);
        }
        if (!more) ray = -1;
      }
      const bool need = ray < 0;
      const unsigned int need_m = __ballot_sync(0xffffffffu, need);
      if (need_m != 0u && !exhausted) {
        const int cnt = __popc(need_m);
        unsigned int qb = 0;
        if (first_fetch) {
          // every warp's first 32 rays are assigned statically: 3552 warps hitting one ticket word at kernel
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
          ray = p.queue[q];
          const float4 a = p.in.s0[ray];
          const float4 b = p.in.s1[ray];
          const float t0 = reinterpret_cast<const float*>(p.out.h0 + ray)[0];
          const int gm = __float_as_int(p.out.h1[ray].z);
          t_min = t0 > 0.0f ? t0 : FLT_MAX;
          hit = t0 > 0.0f ? (gm & 0xffff) : kNoGeom;
          if (STATS) n_nodes = n_tris = 0;
          if (!setup(mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), 0)) ray = -1;
        }
      }
      if (STATS) {
        ++w_refills;
        clk_refill += clock64() - clk_r;
      }
      if (__ballot_sync(0xffffffffu, node != kWalkDone) == 0u && exhausted) break;
      continue;
    }
    // One kind of step per iteration, whichever has more lanes waiting: a NODE step (the four child boxes of
    // every lane's inner node) or a LEAF step (the exact triangle test), so the two code paths never
    // serialise inside one iteration.
    const long long clk_step = STATS ? clock64() : 0;
    const bool node_step = __popc(nm) >= __popc(lm);
    if (node_step) {
      if (can_node) {
        if (STATS) ++n_nodes;
        WideHit wh;
        float4 cf;
        wide_slab_fma(nodes + 8 * (size_t)node, offx, offy, offz, id, noid, lim, &wh, &cf);
        float tn[4] = {wh.tn[0], wh.tn[1], wh.tn[2], wh.tn[3]};
        int ch[4] = {wh.ok[0] ? __float_as_int(cf.x) : kEmptyChild, wh.ok[1] ? __float_as_int(cf.y) : kEmptyChild,
                     wh.ok[2] ? __float_as_int(cf.z) : kEmptyChild, wh.ok[3] ? __float_as_int(cf.w) : kEmptyChild};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ch[k] == kEmptyChild) tn[k] = FLT_MAX;
#define B2PT_CSWAP(a, b)                                   \
  if (tn[b] < tn[a]) {                                     \
    const float tt = tn[a]; tn[a] = tn[b]; tn[b] = tt;     \
    const int cc = ch[a]; ch[a] = ch[b]; ch[b] = cc;       \
  }
        B2PT_CSWAP(0, 1) B2PT_CSWAP(2, 3) B2PT_CSWAP(0, 2) B2PT_CSWAP(1, 3) B2PT_CSWAP(1, 2)
#undef B2PT_CSWAP
        if (ch[0] != kEmptyChild) {
          // nearest first; the others go on the stack farthest first
          if (ch[3] != kEmptyChild) B2PT_WALK_PUSH(ch[3], tn[3])
          if (ch[2] != kEmptyChild) B2PT_WALK_PUSH(ch[2], tn[2])
          if (ch[1] != kEmptyChild) B2PT_WALK_PUSH(ch[1], tn[1])
          node = ch[0];
        } else {
          B2PT_WALK_POP_NEXT()
        }
        ++steps;
      }
    } else {
      if (has_leaf) {
        if (STATS) ++n_tris;
        const float4* tp = tris + 3 * (size_t)(~node);
        const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
        float u, v;
        const float t = tri_exact(qo, qd, mk(ta.x, ta.y, ta.z), mk(tb.x, tb.y, tb.z), mk(tc.x, tc.y, tc.z), &u, &v);
        if (t >= 0.0f) {
          const int fid = __float_as_int(ta.w);
          if (t < tbest || (t == tbest && fid < best)) {
            tbest = t;
            best = fid;
            bu = u;
            bv = v;
            lim = t * 1.00001f + 1e-6f;
          }
        }
        B2PT_WALK_POP_NEXT()
        ++steps;
      }
    }
    if (STATS) {
      __syncwarp();
      const long long dt = clock64() - clk_step;
      if (node_step) clk_node += dt; else { clk_leaf += dt; ++w_leaf_iters; }
    }
    const long long clk_hand = STATS ? clock64() : 0;
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
          p.long_queue[lq] = make_int2(ray, g);
          p.long_best[lq] = make_float4(tbest, bu, bv, __int_as_float(best));
          int2* ent = p.long_stack + (size_t)lq * kLongCarry;
          int ne = -1;
          if (sp + 1 <= p.long_carry) {
            for (int k = 0; k < sp; ++k) ent[k] = k < kWalkShort ? sst[k * kWalkThreads] : spill[k - kWalkShort];
            ent[sp] = make_int2(node, __float_as_int(0.0f));
            ne = sp + 1;
          }
          p.long_n[lq] = ne;
          node = kWalkDone;
          best = -1;
        } else {
          steps = -0x40000000;
        }
      }
    }
    if (STATS) clk_handoff += clock64() - clk_hand;
  }
#undef B2PT_WALK_PUSH
#undef B2PT_WALK_POP_NEXT
  if (STATS) {
    atomicAdd(&sstat[0], (unsigned long long)a_walks);
    atomicAdd(&sstat[1], (unsigned long long)a_nodes);
    atomicAdd(&sstat[2], (unsigned long long)a_tris);
    atomicMax(&sstat[3], (unsigned long long)a_maxn);
    atomicMax(&sstat[4], (unsigned long long)a_maxt);
    if (lane == 0) {
      const long long now = clock64();
      atomicAdd(&sstat[21], 1ull);
      atomicAdd(&sstat[22], (unsigned long long)w_iters);
      atomicMax(&sstat[23], (unsigned long long)w_iters);
      atomicAdd(&sstat[24], (unsigned long long)(now - clk_start));
      atomicMax(&sstat[25], (unsigned long long)(now - clk_start));
      atomicAdd(&sstat[26], (unsigned long long)w_refills);
      atomicAdd(&sstat[27], (unsigned long long)clk_refill);
      atomicAdd(&sstat[28], (unsigned long long)(clk_loop - clk_start));
      atomicAdd(&sstat[29], (unsigned long long)w_node_lanes);
      atomicAdd(&sstat[30], (unsigned long long)w_node_iters);
      atomicAdd(&sstat[31], (unsigned long long)clk_node);
      atomicAdd(&sstat[20], (unsigned long long)clk_leaf);
      atomicAdd(&sstat[19], (unsigned long long)w_leaf_iters);
      atomicAdd(&sstat[18], (unsigned long long)clk_handoff);
    }
  }
  __syncthreads();
  if (STATS && tid < 32 && sstat[tid]) {
    if (tid == 3 || tid == 4 || tid == 23 || tid == 25) atomicMax(&p.stats[tid], sstat[tid]); else atomicAdd(&p.stats[tid], sstat[tid]);
  }
  for (int i = tid; i < kMaxMaterials; i += kWalkThreads) {
    const int c = shist[i];
    if (c) atomicAdd(&p.ctr->hist[p.depth][i], (unsigned int)c);
    const int cl = slive[i];
    if (cl) atomicAdd(&p.ctr->hist_live[p.depth][i], (unsigned int)cl);
  }
}

// Long walks, kCoopGroup lanes per ray (two rays per warp).  A group shares one
// stack of (node, entry distance) pairs in shared memory, seeded with the state
// k_mesh_walk handed over; each round its lanes take the top entries, inner nodes
// test their four child boxes, leaves run the exact triangle test, the closest
// (t, face) of the round is reduced across the group and the surviving children
// are appended with a segmented scan.  The visiting order is no longer
// nearest-first, which is harmless: the winner is the lexicographic minimum of
// (t, face id) whatever the order, and entries behind it are dropped.  Close to
// the capacity of the stack a group falls back to one entry per round
// (depth-first, growth <= 3 per level).  The loop is warp-uniform: the groups
// of a warp run their rounds in lockstep and refill independently.  (8 lanes per
// ray are better for aggregate throughput, 32 for one context alone; 16 is the
// compromise, profiles/r01_notes.md.)
constexpr int kCoopThreads = 128;
constexpr int kCoopGroup = 16;
constexpr int kCoopCap = 32 * kCoopGroup;                        // stack entries per group (32 KB per CTA in all)
constexpr int kCoopDfs = kCoopCap - 96 - 3 * kCoopGroup;         // above this only one entry per round is taken
constexpr unsigned int kCoopMask = kCoopGroup == 32 ? 0xffffffffu : ((1u << (kCoopGroup & 31)) - 1u);

__global__ void __launch_bounds__(kCoopThreads) k_mesh_walk_long(IsectParams p) {
  __shared__ int2 cstack[(kCoopThreads / kCoopGroup) * kCoopCap];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (kCoopGroup - 1);           // lane within the group
  const int gfirst = lane & ~(kCoopGroup - 1);      // first lane of the group
  int2* const st = cstack + (threadIdx.x / kCoopGroup) * kCoopCap;
  const unsigned int total = min(p.ctr->long_count[p.depth], (unsigned int)p.long_cap);
  unsigned int* head = &p.ctr->long_ticket[p.depth];

  // ---- group state (replicated in the lanes of the group) ----
  int ray = -1, g = 0, hit = kNoGeom, sp = 0, best = -1;
  float t_min = FLT_MAX, tbest = 0.0f, lim = 0.0f, bu = 0.0f, bv = 0.0f;
  V3 qo = mk(0, 0, 0), qd = mk(0, 0, 1), id = mk(0, 0, 1), noid = mk(0, 0, 0);
  int offx = 0, offy = 32, offz = 64;
  const float4* nodes = nullptr;
  const float4* tris = nullptr;
  bool exhausted = false;

  while (true) {
    if (sp == 0) {  // group-uniform: fold the finished walk, fetch the next one
      if (ray >= 0) {
        if (gl == 0 && best >= 0 && tbest > 0.0f && (tbest < t_min || (tbest == t_min && g < hit))) {
          const int old_mat = hit == kNoGeom ? 0 : p.scene.geoms[hit].material;
          const int mat = p.scene.geoms[g].material;
          if (mat != old_mat) {
            atomicAdd(&p.ctr->hist[p.depth][mat], 1u);
            atomicSub(&p.ctr->hist[p.depth][old_mat], 1u);
          }
          if (p.live[ray]) {
            atomicSub(&p.ctr->hist_live[p.depth][old_mat], 1u);
            p.live[ray] = 0;
          }
          p.key[ray] = (uint8_t)mat;
          reinterpret_cast<float*>(p.out.h0 + ray)[0] = tbest;
          p.out.h1[ray] = make_float4(bu, bv, __int_as_float((g & 0xffff) | (mat << 16)), __int_as_float(best));
        }
        ray = -1;
      }
      if (!exhausted) {
        unsigned int q = 0;
        if (gl == 0) q = atomicAdd(head, 1u);
        q = __shfl_sync(kCoopMask << gfirst, q, gfirst);
        if (q >= total) {
          exhausted = true;
        } else {
          const int2 job = p.long_queue[q];
          ray = job.x;
          g = job.y;
          const DevGeom& G = p.scene.geoms[g];
          const DevMesh& M = p.scene.meshes[G.mesh];
          const float4 a = p.in.s0[ray];
          const float4 b = p.in.s1[ray];
          const float t0 = reinterpret_cast<const float*>(p.out.h0 + ray)[0];
          const int gm0 = __float_as_int(p.out.h1[ray].z);
          t_min = t0 > 0.0f ? t0 : FLT_MAX;
          hit = t0 > 0.0f ? (gm0 & 0xffff) : kNoGeom;
          qo = xform(G.inv, mk(a.x, a.y, a.z), 1.0f);
          qd = normalize(xform(G.inv, mk(b.x, b.y, b.z), 0.0f));
          id = rcp_fast(qd);
          noid = mk(-(qo.x * id.x), -(qo.y * id.y), -(qo.z * id.z));
          slab_offsets(id, &offx, &offy, &offz);
          const float t_limit = (G.rigid && t_min < FLT_MAX) ? t_min * 1.0001f + 1e-5f : FLT_MAX;
          // the state k_mesh_walk handed over: closest triangle so far and (if it fitted) the traversal stack
          const float4 carried = p.long_best[q];
          const int n_carried = p.long_n[q];
          tbest = carried.x;
          best = __float_as_int(carried.w);
          bu = carried.y;
          bv = carried.z;
          lim = best >= 0 ? tbest * 1.00001f + 1e-6f : (t_limit >= FLT_MAX ? FLT_MAX : t_limit * 1.00001f + 1e-6f);
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
      if (__int_as_float(e.y) <= lim) node = e.x;
    }
    sp -= take;
    __syncwarp();
    int ch[4] = {kEmptyChild, kEmptyChild, kEmptyChild, kEmptyChild};
    float tn[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
    float t = -1.0f, u = 0.0f, v = 0.0f;
    int fid = 0x7fffffff;
    if (walk_is_inner(node)) {
      WideHit wh;
      float4 cf;
      wide_slab_fma(nodes + 8 * (size_t)node, offx, offy, offz, id, noid, lim, &wh, &cf);
      tn[0] = wh.tn[0]; tn[1] = wh.tn[1]; tn[2] = wh.tn[2]; tn[3] = wh.tn[3];
      ch[0] = wh.ok[0] ? __float_as_int(cf.x) : kEmptyChild;
      ch[1] = wh.ok[1] ? __float_as_int(cf.y) : kEmptyChild;
      ch[2] = wh.ok[2] ? __float_as_int(cf.z) : kEmptyChild;
      ch[3] = wh.ok[3] ? __float_as_int(cf.w) : kEmptyChild;
    } else if (node < 0) {
      const float4* tp = tris + 3 * (size_t)(~node);
      const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
      t = tri_exact(qo, qd, mk(ta.x, ta.y, ta.z), mk(tb.x, tb.y, tb.z), mk(tc.x, tc.y, tc.z), &u, &v);
      fid = __float_as_int(ta.w);
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
        if (tr < tbest || (tr == tbest && fr < best)) {
          tbest = tr;
          best = fr;
          bu = su;
          bv = sv;
          lim = tr * 1.00001f + 1e-6f;
        }
      }
    }
    // append the children that are still in front of the closest hit (segmented scan over the group)
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (ch[k] != kEmptyChild && !(tn[k] <= lim)) ch[k] = kEmptyChild;
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
    for (int k = 0; k < 4; ++k)
      if (ch[k] != kEmptyChild) st[w++] = make_int2(ch[k], __float_as_int(tn[k]));
    sp += all;
    __syncwarp();
  }
}

// Finish the records of the rays a mesh won: k_mesh_walk left (t | bu, bv, geom|mat, face).
__global__ void __launch_bounds__(256) k_mesh_finish(IsectParams p) {
  __shared__ int slive[kMaxMaterials];
  const int tid = threadIdx.x;
  const unsigned int total = p.ctr->mesh_count[p.depth];
  if (total == 0) return;
  for (int i = tid; i < kMaxMaterials; i += blockDim.x) slive[i] = 0;
  __syncthreads();
  for (unsigned int q = blockIdx.x * blockDim.x + tid; q < total; q += gridDim.x * blockDim.x) {
    const int i = p.queue[q];
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
    mesh_record(G, M, face, h1.x, h1.y, &nrm, &tu, &tv);
    p.out.h0[i] = make_float4(t, nrm.x, nrm.y, nrm.z);
    p.out.h1[i] = make_float4(tu, tv, h1.z, h1.w);
    // survival: an emissive texel turns the hit into a light (interactions.h:171-186);
    // scatterRay only looks at the emission map in its OBJ branch (not reflective, not refractive)
    bool emissive = false;
    const DevMaterial& mm = p.scene.materials[mat];
    if (M.ke.channels && !(__ldg(&mm.has_reflective) > 0) && !(__ldg(&mm.has_refractive) > 0)) {
      const V3 e = fetch_texel(M.ke, tu, tv);
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
