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
//  k_mesh_finish  One thread per queued ray, full warps: rays the mesh won get
//                 their uv, geometric normal, bump-mapped normal and survival
//                 flag (intersections.h:226,235-279, interactions.h:171-186).
//
// Parity rules are those of k_intersect.cuh: the BVH only prunes, the leaf test
// is glm::intersectRayTriangle exactly, ties go to the lowest face / geom id.
#pragma once

#include "k_intersect.cuh"

namespace b2pt {

#ifndef B2PT_WALK_THREADS
#define B2PT_WALK_THREADS 256
#endif
#ifndef B2PT_WALK_MINBLOCKS
#define B2PT_WALK_MINBLOCKS 3
#endif
#ifndef B2PT_REFILL_MIN
#define B2PT_REFILL_MIN 8
#endif
#ifndef B2PT_WALK_PREFETCH
#define B2PT_WALK_PREFETCH 0
#endif
constexpr int kWalkThreads = B2PT_WALK_THREADS;
constexpr int kWalkShort = 8;    // (node, tn) entries per lane in shared memory
constexpr int kWalkSpill = 88;   // spill entries per lane: 96 in all = 3 per level of a 32-level wide tree
constexpr int kRefillMin = B2PT_REFILL_MIN;    // idle lanes that trigger a refill
constexpr int kWalkDone = 0x7fffffff;
constexpr int kNoGeom = 0x7fffffff;

__device__ __forceinline__ bool walk_is_inner(int node) { return (unsigned int)node < 0x40000000u; }

template <bool STATS>
__global__ void __launch_bounds__(kWalkThreads, B2PT_WALK_MINBLOCKS) k_mesh_walk(IsectParams p) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  __shared__ int2 sstack[kWalkShort * kWalkThreads];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n_geoms = p.scene.n_geoms;
  const unsigned int total = p.ctr->mesh_count[p.depth];
  if (total == 0) return;
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
  float tbest = 0.0f, lim = 0.0f, bu = 0.0f, bv = 0.0f;
  int best = -1;
  float t_min = FLT_MAX;  // closest hit so far over all geoms (analytic + meshes already walked)
  int hit = kNoGeom;
  unsigned int n_nodes = 0, n_tris = 0;
  int steps = 0;
  bool exhausted = false;
  const int long_walk = p.long_walk;

#define B2PT_WALK_PUSH(c, t)                                                            \
  {                                                                                     \
    const int2 e = make_int2((c), __float_as_int(t));                                   \
    if (sp < kWalkShort) sst[sp * kWalkThreads] = e;                                    \
    else if (sp < kWalkShort + kWalkSpill) spill[sp - kWalkShort] = e;                  \
    ++sp;                                                                               \
  }
// A walk that outgrew its lane is handed, whole, to k_mesh_walk_long (one warp per ray).
#define B2PT_WALK_STEP_DONE()                                                           \
  if (++steps > long_walk && node != kWalkDone) {                                       \
    p.long_queue[atomicAdd(&p.ctr->long_count[p.depth], 1u)] = make_int2(ray, g);       \
    node = kWalkDone;                                                                   \
    best = -1;                                                                          \
  }
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
    const V3 idw = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    for (int gg = g0; gg < n_geoms; ++gg) {
      const DevGeom& G = sgeom[gg];
      if (G.type != 3 || G.mesh < 0) continue;
      if (!may_beat(G, o, idw, t_min, G.rigid != 0)) continue;
      const DevMesh& M = p.scene.meshes[G.mesh];
      nodes = M.nodes;
      tris = M.tris;
      qo = xform(G.inv, o, 1.0f);
      qd = normalize(xform(G.inv, d, 0.0f));
      id = mk(1.0f / qd.x, 1.0f / qd.y, 1.0f / qd.z);
      const float t_limit = (G.rigid && t_min < FLT_MAX) ? t_min * 1.0001f + 1e-5f : FLT_MAX;
      tbest = t_limit;
      lim = t_limit >= FLT_MAX ? FLT_MAX : t_limit * 1.00001f + 1e-6f;
      best = -1;
      sp = 0;
      steps = 0;
      node = M.root;
      g = gg;
      return true;
    }
    return false;
  };

  while (true) {
    const bool is_node = walk_is_inner(node);
    const bool is_leaf = node < 0;
    const unsigned int nm = __ballot_sync(0xffffffffu, is_node);
    const unsigned int lm = __ballot_sync(0xffffffffu, is_leaf);
    const unsigned int busy = nm | lm;
    if (busy == 0u || (!exhausted && __popc(~busy) >= kRefillMin)) {
      // ---- refill: idle lanes fold their result, move on to the ray's next mesh or fetch a new ray ----
      if (!is_node && !is_leaf && ray >= 0) {
        if (STATS) {
          atomicAdd(&p.stats[0], 1ull);
          atomicAdd(&p.stats[1], (unsigned long long)n_nodes);
          atomicAdd(&p.stats[2], (unsigned long long)n_tris);
          atomicMax(&p.stats[3], (unsigned long long)n_nodes);
          atomicMax(&p.stats[4], (unsigned long long)n_tris);
          atomicAdd(&p.stats[5 + min(15u, 31u - __clz((n_nodes + n_tris) | 1u))], 1ull);  // log2 histogram of steps
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
        if (lane == 0) qb = atomicAdd(head, (unsigned int)cnt);
        qb = __shfl_sync(0xffffffffu, qb, 0);
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
      if (__ballot_sync(0xffffffffu, node != kWalkDone) == 0u && exhausted) break;
      continue;
    }
    if (__popc(nm) >= __popc(lm)) {
      // ---- node step ----
      if (is_node) {
        if (STATS) ++n_nodes;
        const float4* n = nodes + 8 * (size_t)node;
        const float4 lx = __ldg(n), ly = __ldg(n + 1), lz = __ldg(n + 2);
        const float4 hx = __ldg(n + 3), hy = __ldg(n + 4), hz = __ldg(n + 5);
        const float4 cf = __ldg(n + 6);
        float tn[4];
        int ch[4] = {__float_as_int(cf.x), __float_as_int(cf.y), __float_as_int(cf.z), __float_as_int(cf.w)};
        {
          float tf;
          slab(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, qo, id, &tn[0], &tf);
          if (!(tn[0] <= tf && tf >= 0.0f && tn[0] <= lim)) ch[0] = kEmptyChild;
          slab(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, qo, id, &tn[1], &tf);
          if (!(tn[1] <= tf && tf >= 0.0f && tn[1] <= lim)) ch[1] = kEmptyChild;
          slab(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, qo, id, &tn[2], &tf);
          if (!(tn[2] <= tf && tf >= 0.0f && tn[2] <= lim)) ch[2] = kEmptyChild;
          slab(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, qo, id, &tn[3], &tf);
          if (!(tn[3] <= tf && tf >= 0.0f && tn[3] <= lim)) ch[3] = kEmptyChild;
        }
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
          if (ch[1] != kEmptyChild) {
            B2PT_WALK_PUSH(ch[1], tn[1])
#if B2PT_WALK_PREFETCH
            // the entry on top of the stack is the likeliest next visit: have it in L1 by then
            const void* pf = ch[1] >= 0 ? (const void*)(nodes + 8 * (size_t)ch[1]) : (const void*)(tris + 3 * (size_t)(~ch[1]));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
#endif
          }
          node = ch[0];
        } else {
          B2PT_WALK_POP_NEXT()
        }
        B2PT_WALK_STEP_DONE()
      }
    } else {
      // ---- leaf step ----
      if (is_leaf) {
        if (STATS) ++n_tris;
        const float4* tp = tris + 3 * (size_t)(~node);
        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        float u, v;
        const float t = tri_exact(qo, qd, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), &u, &v);
        if (t >= 0.0f) {
          const int fid = __float_as_int(a.w);
          if (t < tbest || (t == tbest && fid < best)) {
            tbest = t;
            best = fid;
            bu = u;
            bv = v;
            lim = t * 1.00001f + 1e-6f;
          }
        }
        B2PT_WALK_POP_NEXT()
        B2PT_WALK_STEP_DONE()
      }
    }
  }
#undef B2PT_WALK_PUSH
#undef B2PT_WALK_STEP_DONE
#undef B2PT_WALK_POP_NEXT
  __syncthreads();
  for (int i = tid; i < kMaxMaterials; i += kWalkThreads) {
    const int c = shist[i];
    if (c) atomicAdd(&p.ctr->hist[p.depth][i], (unsigned int)c);
    const int cl = slive[i];
    if (cl) atomicAdd(&p.ctr->hist_live[p.depth][i], (unsigned int)cl);
  }
}

// One warp per long walk.  The warp shares one stack of (node, entry distance)
// pairs in shared memory; each round the 32 lanes take the top 32 entries, inner
// nodes test their four child boxes, leaves run the exact triangle test, the
// closest (t, face) of the round is reduced across the warp and the surviving
// children are appended with a warp scan.  The visiting order is no longer
// nearest-first, which is harmless: the winner is the lexicographic minimum of
// (t, face id) whatever the order, and entries behind it are dropped.  Close to
// the capacity of the stack the warp falls back to one entry per round
// (depth-first, growth <= 3 per level).
constexpr int kCoopThreads = 128;
constexpr int kCoopCap = 1024;

__global__ void __launch_bounds__(kCoopThreads) k_mesh_walk_long(IsectParams p) {
  __shared__ int2 cstack[(kCoopThreads / 32) * kCoopCap];
  const int lane = threadIdx.x & 31;
  int2* const st = cstack + (threadIdx.x >> 5) * kCoopCap;
  const unsigned int total = p.ctr->long_count[p.depth];
  unsigned int* head = &p.ctr->long_ticket[p.depth];
  while (true) {
    unsigned int q = 0;
    if (lane == 0) q = atomicAdd(head, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= total) break;
    const int2 job = p.long_queue[q];
    const int ray = job.x, g = job.y;
    const DevGeom& G = p.scene.geoms[g];
    const DevMesh& M = p.scene.meshes[G.mesh];
    const float4 a = p.in.s0[ray];
    const float4 b = p.in.s1[ray];
    const float t0 = reinterpret_cast<const float*>(p.out.h0 + ray)[0];
    const int gm0 = __float_as_int(p.out.h1[ray].z);
    const float t_min = t0 > 0.0f ? t0 : FLT_MAX;
    const int hit = t0 > 0.0f ? (gm0 & 0xffff) : kNoGeom;
    const V3 qo = xform(G.inv, mk(a.x, a.y, a.z), 1.0f);
    const V3 qd = normalize(xform(G.inv, mk(b.x, b.y, b.z), 0.0f));
    const V3 id = mk(1.0f / qd.x, 1.0f / qd.y, 1.0f / qd.z);
    const float t_limit = (G.rigid && t_min < FLT_MAX) ? t_min * 1.0001f + 1e-5f : FLT_MAX;
    float tbest = t_limit;
    float lim = t_limit >= FLT_MAX ? FLT_MAX : t_limit * 1.00001f + 1e-6f;
    int best = -1;
    float bu = 0.0f, bv = 0.0f;
    const float4* nodes = M.nodes;
    const float4* tris = M.tris;
    __syncwarp();
    if (lane == 0) st[0] = make_int2(M.root, __float_as_int(0.0f));
    int sp = 1;
    __syncwarp();
    while (sp > 0) {
      const int take = sp > kCoopCap - 256 ? 1 : min(sp, 32);
      int node = kWalkDone;
      if (lane < take) {
        const int2 e = st[sp - 1 - lane];
        if (__int_as_float(e.y) <= lim) node = e.x;
      }
      sp -= take;
      __syncwarp();
      int ch[4];
      float tn[4];
      float t = -1.0f, u = 0.0f, v = 0.0f;
      int fid = 0x7fffffff;
      if (walk_is_inner(node)) {
        const float4* n = nodes + 8 * (size_t)node;
        const float4 lx = __ldg(n), ly = __ldg(n + 1), lz = __ldg(n + 2);
        const float4 hx = __ldg(n + 3), hy = __ldg(n + 4), hz = __ldg(n + 5);
        const float4 cf = __ldg(n + 6);
        ch[0] = __float_as_int(cf.x); ch[1] = __float_as_int(cf.y); ch[2] = __float_as_int(cf.z); ch[3] = __float_as_int(cf.w);
        float tf;
        slab(lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, qo, id, &tn[0], &tf);
        if (!(tn[0] <= tf && tf >= 0.0f)) ch[0] = kEmptyChild;
        slab(lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, qo, id, &tn[1], &tf);
        if (!(tn[1] <= tf && tf >= 0.0f)) ch[1] = kEmptyChild;
        slab(lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, qo, id, &tn[2], &tf);
        if (!(tn[2] <= tf && tf >= 0.0f)) ch[2] = kEmptyChild;
        slab(lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, qo, id, &tn[3], &tf);
        if (!(tn[3] <= tf && tf >= 0.0f)) ch[3] = kEmptyChild;
      } else {
        ch[0] = ch[1] = ch[2] = ch[3] = kEmptyChild;
        tn[0] = tn[1] = tn[2] = tn[3] = FLT_MAX;
        if (node < 0) {
          const float4* tp = tris + 3 * (size_t)(~node);
          const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
          t = tri_exact(qo, qd, mk(ta.x, ta.y, ta.z), mk(tb.x, tb.y, tb.z), mk(tc.x, tc.y, tc.z), &u, &v);
          fid = __float_as_int(ta.w);
        }
      }
      // closest (t, face) of this round; t >= 0, so the bit patterns order like the values
      const unsigned int tb_ = t >= 0.0f ? __float_as_uint(t) : 0xffffffffu;
      const unsigned int mn = __reduce_min_sync(0xffffffffu, tb_);
      if (mn != 0xffffffffu) {
        const unsigned int fm = __reduce_min_sync(0xffffffffu, tb_ == mn ? (unsigned int)fid : 0xffffffffu);
        const float tr = __uint_as_float(mn);
        const int src = __ffs(__ballot_sync(0xffffffffu, tb_ == mn && (unsigned int)fid == fm)) - 1;
        const float su = __shfl_sync(0xffffffffu, u, src), sv = __shfl_sync(0xffffffffu, v, src);
        if (tr < tbest || (tr == tbest && (int)fm < best)) {
          tbest = tr;
          best = (int)fm;
          bu = su;
          bv = sv;
          lim = tr * 1.00001f + 1e-6f;
        }
      }
      // append the children that are still in front of the closest hit
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ch[k] != kEmptyChild && !(tn[k] <= lim)) ch[k] = kEmptyChild;
        cnt += ch[k] != kEmptyChild;
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const int all = __shfl_sync(0xffffffffu, incl, 31);
      int w = sp + incl - cnt;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ch[k] != kEmptyChild) st[w++] = make_int2(ch[k], __float_as_int(tn[k]));
      sp += all;
      __syncwarp();
    }
    if (lane == 0 && best >= 0 && tbest > 0.0f && (tbest < t_min || (tbest == t_min && g < hit))) {
      const int old_mat = hit == kNoGeom ? 0 : p.scene.geoms[hit].material;
      const int mat = G.material;
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
