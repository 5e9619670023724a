// k_intersect.cuh -- closest hit (computeIntersections, apps/src/pathtrace.cu:303-386): the analytic geoms.
//
//  k_intersect_analytic    one thread per ray against the cubes and spheres (staged in shared memory once per
//                          CTA): a cheap padded world-box pre-test, then the reference's exact tests; writes the
//                          hit record, the sort key, the survival flag and the material histograms, and queues
//                          the rays that may hit a mesh for k_walk.cuh.
//  k_intersect_mesh_brute  validation only (use_bvh = 0): the reference's loop over every face.
// Shared device functions: the exact cube / sphere / triangle tests, the texel fetch, the mesh part of a hit
// record (uv, normals, bump map).
//
// Parity rules:
//  * pre-tests only prune: boxes are padded and the slab tests are conservative; what decides is the
//    reference's arithmetic, evaluated in glm's operation order without FMA contraction (pt_math.cuh);
//  * the mesh leaf test is glm::intersectRayTriangle exactly (gtx/intersect.inl:44-73) followed by
//    t = distance(p, origin) in object space (apps/src/intersections.h:219-223);
//  * ties are resolved as the reference's strict `<` loops do: lowest face id within a mesh
//    (intersections.h:223), lowest geom id across geoms (pathtrace.cu:360);
//  * per-geom work that only the winner needs (normals, uv, the bump texel) is deferred until the closest
//    geom is known; the values are identical.
#pragma once

#include <float.h>

#include "k_lbvh.cuh"
#include "pt_device.cuh"

namespace b2pt {

constexpr int kIsectThreads = 256;

struct IsectParams {
  DevScene scene;
  PathBuf in;
  HitBuf out;
  uint8_t* key;
  uint8_t* live;  // 1 if the path will survive the coming shade (decided here, see will_survive)
  Counters* ctr;
  int depth;
  float4* queue;              // rays that must walk a mesh (filled by k_intersect_analytic), three planes of queue_cap
                              // entries: (origin, path slot) (direction, closest analytic t) (geom|material<<16, survival flag)
  int queue_cap;
  int2* long_queue;           // (ray, geom) walks that outgrew one lane (filled by k_mesh_walk)
  int long_walk;              // steps after which k_mesh_walk hands a walk to k_mesh_walk_long
  int long_cap;               // capacity of the hand-off queue
  int long_carry;             // stack entries a hand-off may carry (<= kLongCarry; fewer only in tests)
  float4* long_best;          // [long_cap] (t, bu, bv, face) of the closest triangle found so far
  int2* long_stack;           // [long_cap][kLongCarry] carried traversal stack
  int* long_n;                // [long_cap] carried entries, -1: restart at the root
  unsigned long long* stats;  // optional traversal statistics (B2PT_TRAVERSAL_STATS=1), else NULL
};

// Conservative ray/box slab test; NaNs (0 * inf) are dropped by fminf/fmaxf.
__device__ __forceinline__ void slab(float bx0, float by0, float bz0, float bx1, float by1, float bz1, V3 o, V3 id,
                                     float* tn, float* tf) {
  float x0 = (bx0 - o.x) * id.x, x1 = (bx1 - o.x) * id.x;
  float y0 = (by0 - o.y) * id.y, y1 = (by1 - o.y) * id.y;
  float z0 = (bz0 - o.z) * id.z, z1 = (bz1 - o.z) * id.z;
  *tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
  *tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * 1.0000004f;
}

// glm::intersectRayTriangle + t = distance(p, q.origin); returns t or -1.
__device__ __forceinline__ float tri_exact(V3 qo, V3 qd, V3 v0, V3 v1, V3 v2, float* bu, float* bv) {
  V3 e1 = v1 - v0;
  V3 e2 = v2 - v0;
  V3 p = cross(qd, e2);
  float a = dot(e1, p);
  if (a < FLT_EPSILON) return -1.0f;
  float f = 1.0f / a;
  V3 s = qo - v0;
  float u = f * dot(s, p);
  if (u < 0.0f) return -1.0f;
  if (u > 1.0f) return -1.0f;
  V3 q = cross(s, e1);
  float v = f * dot(qd, q);
  if (v < 0.0f) return -1.0f;
  if (v + u > 1.0f) return -1.0f;
  float tt = f * dot(e2, q);
  if (!(tt >= 0.0f)) return -1.0f;
  float w = 1 - u - v;
  V3 pt = (v0 * w + v1 * u) + v2 * v;
  *bu = u;
  *bv = v;
  return length(qo - pt);
}

// The reference's loop over every face, in face order (intersections.h:216-230).
__device__ __forceinline__ float mesh_brute(const DevMesh& m, V3 qo, V3 qd, int* face, float* bu, float* bv) {
  float tmin = FLT_MAX;
  int nearest = -1;
  for (int j = 0; j < m.n_faces; ++j) {
    const float* fp = m.face_pos + 9 * (size_t)j;
    float u, v;
    const float t = tri_exact(qo, qd, mk(__ldg(fp), __ldg(fp + 1), __ldg(fp + 2)),
                              mk(__ldg(fp + 3), __ldg(fp + 4), __ldg(fp + 5)),
                              mk(__ldg(fp + 6), __ldg(fp + 7), __ldg(fp + 8)), &u, &v);
    if (t >= 0.0f && t < tmin) {
      tmin = t;
      nearest = j;
      *bu = u;
      *bv = v;
    }
  }
  *face = nearest;
  return nearest >= 0 ? tmin : -1.0f;
}

// Texel fetch as written at all six sites of the reference
// (e.g. intersections.h:270-276): px = (int)(v*h)*w + (int)(u*w), three bytes
// at px*channels, /255.f.  The index is clamped into the image for memory
// safety only (the reference does not bound it).
__device__ __forceinline__ V3 fetch_texel(const DevTexture& tx, float u, float v) {
  const int cu = (int)(u * (float)tx.w);
  const int cv = (int)(v * (float)tx.h);
  long long id = (long long)cv * tx.w + cu;
  const long long last = (long long)tx.w * tx.h - 1;
  id = id < 0 ? 0 : (id > last ? last : id);
  const uint8_t* px = tx.texels + id * tx.channels;
  const unsigned int r = __ldg(px), g = __ldg(px + 1), b = __ldg(px + 2);
  return mk((float)r / 255.f, (float)g / 255.f, (float)b / 255.f);
}

// The material of a mesh face and the four maps of an OBJ hit.  The reference has one material per OBJ and
// keeps the maps in the geom (apps/src/scene.cpp:134-231); with per-face materials (B2ptScene::face_material /
// material_textures, an extension: the reference discards tinyobj's ids, scene.cpp:121-122) both come from the
// face that was hit.  which: 0 kd, 1 ks, 2 bump, 3 ke.
__device__ __forceinline__ int mesh_face_material(const int* __restrict__ face_mat, int face, int geom_material) {
  return face_mat ? __ldg(face_mat + face) : geom_material;
}
__device__ __forceinline__ const DevTexture& obj_tex(const DevScene& sc, const DevMesh& M, int material, int which) {
  return sc.mat_tex ? sc.mat_tex[4 * material + which] : (&M.kd)[which];
}

// 1/x to ~1 ulp (MUFU.RCP) for tests that only prune: the exact tests never see it.
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ V3 rcp_fast(V3 d) { return mk(rcp_fast(d.x), rcp_fast(d.y), rcp_fast(d.z)); }

// Ray against a geom's padded world-space box, one FMA per plane: id ~ 1/d, noid = -(o * id).
// The boxes are padded by >= 1e-3 (world_box), far more than the rounding of the FMA form; NaNs (rays
// parallel to an axis) are dropped by fminf/fmaxf, i.e. that axis does not constrain the box.
__device__ __forceinline__ void world_slab(const DevGeom& G, V3 id, V3 noid, float* tn, float* tf) {
  const float x0 = fmaf(G.wmin.x, id.x, noid.x), x1 = fmaf(G.wmax.x, id.x, noid.x);
  const float y0 = fmaf(G.wmin.y, id.y, noid.y), y1 = fmaf(G.wmax.y, id.y, noid.y);
  const float z0 = fmaf(G.wmin.z, id.z, noid.z), z1 = fmaf(G.wmax.z, id.z, noid.z);
  *tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
  *tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
}

// Conservative world-space pre-test against a geom's padded bounding box.
// Returns false only if the exact test is certain to miss, or certain to
// return a distance that cannot beat t_best (strictly closer hits win).
__device__ __forceinline__ bool may_beat(const DevGeom& G, V3 id, V3 noid, float t_best, bool world_metric) {
  float tn, tf;
  world_slab(G, id, noid, &tn, &tf);
  if (!(tn <= tf) || tf < 0.0f) return false;
  // G.wmin.w = slack: the exact tests pull the hit point back by 1e-4 in OBJECT
  // space (getPointOnRay), i.e. by up to 1e-4 * scale in world space
  if (world_metric && t_best < FLT_MAX && tn > t_best * 1.0001f + G.wmin.w) return false;
  return true;
}

// The mesh part of the winner's record: uv, geometric normal, bump map
// (apps/src/intersections.h:226,235-279).
__device__ __forceinline__ void mesh_record(const DevGeom& G, const DevMesh& M, const DevTexture& bump, int face, float bu,
                                            float bv, V3* nrm_out, float* tu_out, float* tv_out) {
  const float4* fr = M.face_rec + 4 * (size_t)face;
  const float4 r0 = __ldg(fr), r1 = __ldg(fr + 1), r2 = __ldg(fr + 2), r3 = __ldg(fr + 3);
  const V3 v0 = mk(r0.x, r0.y, r0.z);
  const V3 v1 = mk(r0.w, r1.x, r1.y);
  const V3 v2 = mk(r1.z, r1.w, r2.x);
  const float u0x = r2.y, u0y = r2.z, u1x = r2.w, u1y = r3.x;
  const float u2x = r3.y, u2y = r3.z;
  const float w = 1 - bu - bv;
  const float tu = (w * u0x + bu * u1x) + bv * u2x;
  const float tv = (w * u0y + bu * u1y) + bv * u2y;
  const V3 e1 = v1 - v0, e2 = v2 - v0;
  V3 nrm = normalize(xform(G.invT, normalize(cross(e1, e2)), 0.0f));
  if (bump.channels) {
    const float d1x = u1x - u0x, d1y = u1y - u0y, d2x = u2x - u0x, d2y = u2y - u0y;
    const float f = 1.0f / (d1x * d2y - d2x * d1y);
    V3 tang = mk(f * (d2y * e1.x - d1y * e2.x), f * (d2y * e1.y - d1y * e2.y), f * (d2y * e1.z - d1y * e2.z));
    tang = normalize(tang);
    V3 bit = mk(f * (-d2x * e1.x + d1x * e2.x), f * (-d2x * e1.y + d1x * e2.y), f * (-d2x * e1.z + d1x * e2.z));
    bit = normalize(bit);
    const V3 T = normalize(xform(G.fwd, tang, 0.0f));
    const V3 B = normalize(xform(G.fwd, bit, 0.0f));
    V3 tsn = normalize(fetch_texel(bump, tu, tv));
    tsn = normalize(mk(tsn.x * 2.0f - 1.0f, tsn.y * 2.0f - 1.0f, tsn.z * 2.0f - 1.0f));
    nrm = normalize(mk((T.x * tsn.x + B.x * tsn.y) + nrm.x * tsn.z, (T.y * tsn.x + B.y * tsn.y) + nrm.y * tsn.z,
                       (T.z * tsn.x + B.z * tsn.y) + nrm.z * tsn.z));
  }
  *nrm_out = nrm;
  *tu_out = tu;
  *tv_out = tv;
}

// Whether a path survives the coming shade is already decided by its hit record
// (shadeFakeMaterial, apps/src/pathtrace.cu:463-496): a miss, an emissive
// material, the last bounce and an emissive texel (interactions.h:182-186) kill
// it, everything else scatters.  Computing the flag here lets the material sort
// rank the survivors as well, so the shade kernel needs no scan of its own.
__device__ __forceinline__ bool will_survive(const DevMaterial* __restrict__ mats, int mat, int bounces, bool emissive_texel) {
  return !(__ldg(&mats[mat].emittance) > 0.0f) && bounces > 1 && !emissive_texel;
}

// One exact test of an analytic geom: boxIntersectionTest (apps/src/intersections.h:48-90) or
// sphereIntersectionTest (:102-144) with the arithmetic the two share written ONCE -- the ray into object space in
// front, getPointOnRay + the point back into world space + the distance behind -- so that a warp whose lanes hold
// cubes and spheres runs those parts together and only the middle (the three slabs / the quadratic) per type.
// Every expression is evaluated as the reference writes it, in glm's order (pt_math.cuh): the bits are the same.
// Returns the world-space distance or -1; *aux: the box's object-space axis normal / the sphere's object-space hit
// point; *kind: 1 box, 2 sphere from outside, 3 sphere from inside.
__device__ __forceinline__ float analytic_exact(const DevGeom& G, V3 o, V3 d, V3* aux, int* kind) {
  const V3 qo = xform(G.inv, o, 1.0f);
  const V3 qd = normalize(xform(G.inv, d, 0.0f));
  float ts = 0.0f;
  bool hit = false;
  int k = 1;
  V3 a = mk(0, 0, 0);
  if (G.type == 1) {
    float tmin = -1e38f, tmax = 1e38f;
    V3 nmin = mk(0, 0, 0), nmax = mk(0, 0, 0);
    {
      const float t1 = (-0.5f - qo.x) / qd.x, t2 = (+0.5f - qo.x) / qd.x;
      const float ta = glm_min(t1, t2), tb = glm_max(t1, t2);
      const V3 nn = mk(t2 < t1 ? +1.0f : -1.0f, 0, 0);
      if (ta > 0 && ta > tmin) { tmin = ta; nmin = nn; }
      if (tb < tmax) { tmax = tb; nmax = nn; }
    }
    {
      const float t1 = (-0.5f - qo.y) / qd.y, t2 = (+0.5f - qo.y) / qd.y;
      const float ta = glm_min(t1, t2), tb = glm_max(t1, t2);
      const V3 nn = mk(0, t2 < t1 ? +1.0f : -1.0f, 0);
      if (ta > 0 && ta > tmin) { tmin = ta; nmin = nn; }
      if (tb < tmax) { tmax = tb; nmax = nn; }
    }
    {
      const float t1 = (-0.5f - qo.z) / qd.z, t2 = (+0.5f - qo.z) / qd.z;
      const float ta = glm_min(t1, t2), tb = glm_max(t1, t2);
      const V3 nn = mk(0, 0, t2 < t1 ? +1.0f : -1.0f);
      if (ta > 0 && ta > tmin) { tmin = ta; nmin = nn; }
      if (tb < tmax) { tmax = tb; nmax = nn; }
    }
    if (tmax >= tmin && tmax > 0) {
      if (tmin <= 0) { tmin = tmax; nmin = nmax; }
      hit = true;
      ts = tmin;
      a = nmin;
    }
  } else {
    const float vdd = dot(qo, qd);
    const float radicand = vdd * vdd - (dot(qo, qo) - 0.25f);
    if (!(radicand < 0)) {
      const float sq = sqrtf(radicand);
      const float first = -vdd;
      const float t1 = first + sq, t2 = first - sq;
      if (!(t1 < 0 && t2 < 0)) {
        hit = true;
        if (t1 > 0 && t2 > 0) { ts = fminf(t1, t2); k = 2; } else { ts = fmaxf(t1, t2); k = 3; }
      }
    }
  }
  if (!hit) return -1.0f;
  // getPointOnRay (:27-29) normalises the direction again
  const V3 p = qo + normalize(qd) * (ts - .0001f);
  if (k != 1) a = p;
  *aux = a;
  *kind = k;
  return length(o - xform(G.fwd, p, 1.0f));
}

// Closest hit in two kernels per depth.
//
//  k_intersect_analytic  one thread per ray against the analytic geoms (cubes,
//     spheres): a cheap padded world-box pre-test skips geoms that cannot win,
//     the survivors run the exact reference tests.  The record of the best
//     analytic hit (or the miss), the sort key and the material histogram are
//     written at once.  Rays whose path crosses a mesh's box in front of that
//     hit are appended to a device queue (one atomic per warp).
//  k_mesh_walk (k_walk.cuh)  persistent warps drain the queue.  Rays the mesh
//     wins get their record, key and histogram entry rewritten.
// The closest hit is the lexicographic minimum of (t, geom id) over all geoms,
// which is what the reference's strict `<` loop in geom order returns.
// The analytic half of one ray: closest cube / sphere, its record, and whether a mesh has to be walked.
struct AnalyticHit {
  float4 h0, h1;
  int mat;
  bool survives, want_mesh;
};
__device__ __forceinline__ void analytic_trace(const DevGeom* sgeom, int n_geoms, const DevMaterial* __restrict__ mats, V3 o,
                                               V3 d, int bounces, AnalyticHit* r) {
  const V3 id = rcp_fast(d);
  const V3 noid = mk(-(o.x * id.x), -(o.y * id.y), -(o.z * id.z));
  float t_min = FLT_MAX;
  int hit = -1, kind = 0;
  V3 aux = mk(0, 0, 0);  // box: axis normal; sphere: object-space point
  // Pass 1 (warp-coherent, cheap): which geoms does this ray's path cross, and which of them does it enter
  // first?  For meshes: the closest box entry (minus the distance slack) of the rigid ones, and whether a
  // non-rigid one is crossed at all.
  unsigned long long cand = 0ull;
  float mesh_tn = FLT_MAX, near_tn = FLT_MAX;
  int g = -1;
  bool mesh_any = false;
  for (int k = 0; k < n_geoms; ++k) {
    const DevGeom& G = sgeom[k];
    float tn, tf;
    world_slab(G, id, noid, &tn, &tf);
    if (!(tn <= tf) || tf < 0.0f) continue;  // most geoms end here: the early exit is cheaper than predicating the rest
    if (G.type == 1 || G.type == 0) {
      cand |= 1ull << k;
      if (tn < near_tn) {
        near_tn = tn;
        g = k;
      }
    } else if (G.type == 3 && G.mesh >= 0) {
      if (G.rigid) mesh_tn = fminf(mesh_tn, tn - G.wmin.w); else mesh_any = true;
    }
  }
  // Pass 2: the exact tests, in ROUNDS the lanes of a warp run together.  A lane starts with the candidate it
  // enters first -- usually the one it hits, after which may_beat drops the others -- and then skips ahead to its
  // next candidate that can still win BEFORE it joins the next round, so a round is entered with every lane that
  // has work (0.8 - 1 exact tests per ray, 1.8 - 1.9 rounds per warp of scattered rays on the Cornell scenes:
  // tools/exp_analytic_model.py).  The result does not depend on the order: the winner is the lexicographic
  // minimum of (t, geom id), which is what the reference's strict `<` loop in geom order keeps.
  if (g >= 0) cand &= ~(1ull << g);
  while (g >= 0) {
    float t;
    V3 taux = mk(0, 0, 0);
    int tkind = 1;
    t = analytic_exact(sgeom[g], o, d, &taux, &tkind);
    if (t > 0.0f && (t < t_min || (t == t_min && g < hit))) {
      t_min = t;
      hit = g;
      kind = tkind;
      aux = taux;
    }
    g = -1;
    while (cand) {
      const int k = __ffsll((long long)cand) - 1;
      cand &= cand - 1ull;
      if (may_beat(sgeom[k], id, noid, t_min, true)) {
        g = k;
        break;
      }
    }
  }
  r->mat = 0;
  r->survives = false;
  if (hit < 0) {
    r->h0 = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
    r->h1 = make_float4(0.0f, 0.0f, __int_as_float(0xffff), __int_as_float(-1));
  } else {
    const DevGeom& G = sgeom[hit];
    V3 nrm = normalize(xform(G.invT, aux, 0.0f));
    if (kind == 3) nrm = -nrm;
    r->mat = G.material;
    r->survives = will_survive(mats, r->mat, bounces, false);
    r->h0 = make_float4(t_min, nrm.x, nrm.y, nrm.z);
    r->h1 = make_float4(0.0f, 0.0f, __int_as_float((hit & 0xffff) | (r->mat << 16)), __int_as_float(-1));
  }
  // a mesh has to be walked if its box is entered in front of the closest analytic hit (a rigid mesh
  // reports world-space distances, so t_min bounds it; any other mesh is always walked)
  r->want_mesh = mesh_any || (mesh_tn < FLT_MAX && (t_min >= FLT_MAX || mesh_tn <= t_min * 1.0001f));
}

// Store the record of slot i, count it in the CTA's histograms and queue it for the mesh walk.  Called by
// WHOLE WARPS (lanes without a ray pass valid = false): the histogram and queue updates are warp-aggregated.
__device__ __forceinline__ void analytic_commit(const IsectParams& p, int* shist, int* slive, int i, bool valid,
                                                const AnalyticHit& r, V3 o, V3 d) {
  const int lane = threadIdx.x & 31;
  if (valid) {
    p.out.h0[i] = r.h0;
    p.out.h1[i] = r.h1;
    p.key[i] = (uint8_t)r.mat;
    p.live[i] = r.survives ? 1 : 0;
  }
  const unsigned int active = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned int peers = __match_any_sync(active, r.mat);
    if (lane == __ffs(peers) - 1) atomicAdd(&shist[r.mat], __popc(peers));
    const unsigned int lpeers = peers & __ballot_sync(active, r.survives);
    if (lpeers && lane == __ffs(lpeers) - 1) atomicAdd(&slive[r.mat], __popc(lpeers));
  }
  const bool want = valid && r.want_mesh;
  const unsigned int mm = __ballot_sync(0xffffffffu, want);
  if (mm) {
    unsigned int qbase = 0;
    if (lane == 0) qbase = atomicAdd(&p.ctr->mesh_count[p.depth], (unsigned int)__popc(mm));
    qbase = __shfl_sync(0xffffffffu, qbase, 0);
    if (want) {
      // the queue entry carries the ray: the walk's refill is then one coalesced round trip instead of a chain of gathers
      const unsigned int q = qbase + __popc(mm & ((1u << lane) - 1u));
      p.queue[q] = make_float4(o.x, o.y, o.z, __int_as_float(i));
      p.queue[(size_t)p.queue_cap + q] = make_float4(d.x, d.y, d.z, r.h0.x);
      p.queue[2 * (size_t)p.queue_cap + q] = make_float4(r.h1.z, __int_as_float(r.survives ? 1 : 0), 0.0f, 0.0f);
    }
  }
}

// Shared-memory set-up / flush of the two helpers above.
__device__ __forceinline__ void analytic_stage(const DevScene& scene, DevGeom* sgeom, int* shist, int* slive) {
  const float4* src = reinterpret_cast<const float4*>(scene.geoms);
  float4* dst = reinterpret_cast<float4*>(sgeom);
  const int words = scene.n_geoms * (int)(sizeof(DevGeom) / 16);
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
  for (int i = threadIdx.x; i < kMaxMaterials; i += blockDim.x) {
    shist[i] = 0;
    slive[i] = 0;
  }
}
__device__ __forceinline__ void analytic_flush(const IsectParams& p, const int* shist, const int* slive) {
  for (int i = threadIdx.x; i < kMaxMaterials; i += blockDim.x) {
    const int c = shist[i];
    if (c) atomicAdd(&p.ctr->hist[p.depth][i], (unsigned int)c);
    const int cl = slive[i];
    if (cl) atomicAdd(&p.ctr->hist_live[p.depth][i], (unsigned int)cl);
  }
}

__global__ void __launch_bounds__(256) k_intersect_analytic(IsectParams p) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  const int tid = threadIdx.x, lane = tid & 31;
  analytic_stage(p.scene, sgeom, shist, slive);
  __syncthreads();
  const int n = p.ctr->n_live[p.depth];
  // whole warps stride over the rays so that the ballots of analytic_commit are warp-wide
  for (int base = (blockIdx.x * blockDim.x + tid) & ~31; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + lane;
    const bool valid = i < n;
    AnalyticHit r;
    r.mat = 0;
    r.survives = r.want_mesh = false;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    if (valid) {
      const float4 a = p.in.s0[i];
      const float4 b = p.in.s1[i];
      o = mk(a.x, a.y, a.z);
      d = mk(b.x, b.y, b.z);
      analytic_trace(sgeom, p.scene.n_geoms, p.scene.materials, o, d, __float_as_int(b.w), &r);
    }
    analytic_commit(p, shist, slive, i, valid, r, o, d);
  }
  __syncthreads();
  analytic_flush(p, shist, slive);
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&p.ctr->segments, (unsigned long long)n);
}

// Validation path (use_bvh = 0): the reference's brute-force loop over every face
// (intersections.h:216-230) for the queued rays, 32 at a time per persistent warp.
// The product path is k_walk.cuh.
__global__ void __launch_bounds__(kIsectThreads, 3) k_intersect_mesh_brute(IsectParams p) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ int shist[kMaxMaterials];
  __shared__ int slive[kMaxMaterials];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n_geoms = p.scene.n_geoms;
  const unsigned int total = p.ctr->mesh_count[p.depth];
  if (total == 0) return;
  {
    const float4* src = reinterpret_cast<const float4*>(p.scene.geoms);
    float4* dst = reinterpret_cast<float4*>(sgeom);
    const int words = n_geoms * (int)(sizeof(DevGeom) / 16);
    for (int i = tid; i < words; i += kIsectThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < kMaxMaterials; i += kIsectThreads) {
      shist[i] = 0;
      slive[i] = 0;
    }
  }
  __syncthreads();
  unsigned int* head = &p.ctr->ray_ticket[p.depth];
  while (true) {
    unsigned int qb = 0;
    if (lane == 0) qb = atomicAdd(head, 32u);
    qb = __shfl_sync(0xffffffffu, qb, 0);
    if (qb >= total) break;
    if (qb + lane < total) {
      const int i = __float_as_int(p.queue[qb + lane].w);
      const float4 a = p.in.s0[i];
      const float4 b = p.in.s1[i];
      const float4 h0 = p.out.h0[i];
      const int gm = __float_as_int(p.out.h1[i].z);
      const V3 o = mk(a.x, a.y, a.z), d = mk(b.x, b.y, b.z);
      const V3 id = rcp_fast(d);
      const V3 noid = mk(-(o.x * id.x), -(o.y * id.y), -(o.z * id.z));
      float t_min = h0.x > 0.0f ? h0.x : FLT_MAX;
      int hit = h0.x > 0.0f ? (gm & 0xffff) : 0x7fffffff;
      const int old_mat = (gm >> 16) & 0xffff;
      int mhit = -1, mface = -1;
      float mbu = 0, mbv = 0;
      for (int g = 0; g < n_geoms; ++g) {
        const DevGeom& G = sgeom[g];
        if (G.type != 3 || G.mesh < 0) continue;
        if (!may_beat(G, id, noid, t_min, G.rigid != 0)) continue;
        const DevMesh& M = p.scene.meshes[G.mesh];
        const V3 qo = xform(G.inv, o, 1.0f);
        const V3 qd = normalize(xform(G.inv, d, 0.0f));
        float bu = 0, bv = 0, t;
        int tface = -1;
        t = mesh_brute(M, qo, qd, &tface, &bu, &bv);
        if (t > 0.0f && (t < t_min || (t == t_min && g < hit))) {
          t_min = t; hit = g; mhit = g; mface = tface; mbu = bu; mbv = bv;
        }
      }
      if (mhit >= 0) {
        const DevGeom& G = sgeom[mhit];
        V3 nrm;
        float tu, tv;
        const int mat = mesh_face_material(p.scene.meshes[G.mesh].face_mat, mface, G.material);
        mesh_record(G, p.scene.meshes[G.mesh], obj_tex(p.scene, p.scene.meshes[G.mesh], mat, 2), mface, mbu, mbv, &nrm, &tu, &tv);
        p.out.h0[i] = make_float4(t_min, nrm.x, nrm.y, nrm.z);
        p.out.h1[i] = make_float4(tu, tv, __int_as_float((mhit & 0xffff) | (mat << 16)), __int_as_float(mface));
        p.key[i] = (uint8_t)mat;
        // survival: an emissive texel turns the hit into a light (interactions.h:171-186)
        const DevMesh& MM = p.scene.meshes[G.mesh];
        bool emissive = false;
        const DevMaterial& mm = p.scene.materials[mat];
        // scatterRay only looks at the emission map in its OBJ branch (not reflective, not refractive)
        const DevTexture& ke = obj_tex(p.scene, MM, mat, 3);
        if (ke.channels && !(__ldg(&mm.has_reflective) > 0) && !(__ldg(&mm.has_refractive) > 0)) {
          const V3 e = fetch_texel(ke, tu, tv);
          emissive = e.x > FLT_EPSILON || e.y > FLT_EPSILON || e.z > FLT_EPSILON;
        }
        const bool survives = will_survive(p.scene.materials, mat, __float_as_int(b.w), emissive);
        const bool old_live = p.live[i] != 0;
        p.live[i] = survives ? 1 : 0;
        if (mat != old_mat) {
          atomicAdd(&shist[mat], 1);
          atomicSub(&shist[old_mat], 1);
        }
        if (old_live) atomicSub(&slive[old_mat], 1);
        if (survives) atomicAdd(&slive[mat], 1);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = tid; i < kMaxMaterials; i += kIsectThreads) {
    const int c = shist[i];
    if (c) atomicAdd(&p.ctr->hist[p.depth][i], (unsigned int)c);
    const int cl = slive[i];
    if (cl) atomicAdd(&p.ctr->hist_live[p.depth][i], (unsigned int)cl);
  }
}

}  // namespace b2pt
