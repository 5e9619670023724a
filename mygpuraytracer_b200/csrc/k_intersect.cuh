// k_intersect.cuh -- closest hit (computeIntersections, apps/src/pathtrace.cu:303-386).
//
// Persistent kernel: the grid is sized to the machine (SMs x resident CTAs),
// each warp claims 32 rays at a time from a device ticket and keeps going
// until the live range [0, n_live[depth]) is exhausted, so the cost of a few
// long BVH walks is spread over the whole chip.  Analytic geoms (cubes,
// spheres) are staged in shared memory once per CTA; OBJ geoms are walked
// through their LBVH (k_lbvh.cuh) with a per-thread short stack in shared
// memory that spills to local memory.
//
// Parity rules:
//  * the BVH only prunes.  Boxes are inflated at build time and the slab test
//    is conservative; the leaf test is glm::intersectRayTriangle exactly
//    (gtx/intersect.inl:44-73) followed by t = distance(p, origin) in object
//    space (apps/src/intersections.h:219-223);
//  * ties are resolved as the reference's strict `<` loops do: lowest face id
//    within a mesh (intersections.h:223), lowest geom id across geoms
//    (pathtrace.cu:360);
//  * per-geom work that only the winner needs (normals, uv, the bump texel)
//    is deferred until the closest geom is known; the values are identical.
#pragma once

#include <float.h>

#include "pt_device.cuh"

namespace b2pt {

constexpr int kIsectThreads = 256;
constexpr int kShortStack = 12;   // entries per thread in shared memory
constexpr int kLocalStack = 52;   // spill entries per thread (LBVH depth <= 64)

struct IsectParams {
  DevScene scene;
  PathBuf in;
  HitBuf out;
  uint8_t* key;
  Counters* ctr;
  int depth;
};

struct StackRef {
  int* sm;     // shared-memory base for this thread (stride = blockDim)
  int* local;  // local-memory spill
  int sp;
  __device__ __forceinline__ void push(int v) {
    if (sp < kShortStack) sm[sp * kIsectThreads] = v; else local[sp - kShortStack] = v;
    ++sp;
  }
  __device__ __forceinline__ int pop() {
    --sp;
    return sp < kShortStack ? sm[sp * kIsectThreads] : local[sp - kShortStack];
  }
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

// Walk the LBVH of one mesh.  t_limit bounds the search (FLT_MAX, or the
// closest analytic hit so far when the geom is rigid).  Returns the object
// space distance of the closest triangle, or -1.
__device__ __forceinline__ float mesh_traverse(const DevMesh& m, V3 qo, V3 qd, float t_limit, int* face, float* bu,
                                               float* bv, StackRef st) {
  const V3 id = mk(1.0f / qd.x, 1.0f / qd.y, 1.0f / qd.z);
  float tbest = t_limit;
  float lim = t_limit >= FLT_MAX ? FLT_MAX : t_limit * 1.00001f + 1e-6f;
  int best = -1;
  int node = m.root;
  st.sp = 0;
  while (true) {
    if (node >= 0) {
      const float4* n = m.nodes + 4 * (size_t)node;
      const float4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3);
      float tnl, tfl, tnr, tfr;
      slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, qo, id, &tnl, &tfl);
      slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, qo, id, &tnr, &tfr);
      const bool hl = tnl <= tfl && tfl >= 0.0f && tnl <= lim;
      const bool hr = tnr <= tfr && tfr >= 0.0f && tnr <= lim;
      const int cl = __float_as_int(n3.x), cr = __float_as_int(n3.y);
      if (hl && hr) {
        const bool left_first = tnl <= tnr;
        st.push(left_first ? cr : cl);
        node = left_first ? cl : cr;
        continue;
      } else if (hl) {
        node = cl;
        continue;
      } else if (hr) {
        node = cr;
        continue;
      }
    } else {
      const int slot = ~node;
      const float4* tp = m.tris + 3 * (size_t)slot;
      const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
      float u, v;
      const float t = tri_exact(qo, qd, mk(a.x, a.y, a.z), mk(b.x, b.y, b.z), mk(c.x, c.y, c.z), &u, &v);
      if (t >= 0.0f) {
        const int fid = __float_as_int(a.w);
        if (t < tbest || (t == tbest && fid < best)) {
          tbest = t;
          best = fid;
          *bu = u;
          *bv = v;
          lim = t * 1.00001f + 1e-6f;
        }
      }
    }
    if (st.sp == 0) break;
    node = st.pop();
  }
  *face = best;
  return best >= 0 ? tbest : -1.0f;
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

template <bool USE_BVH>
__global__ void __launch_bounds__(kIsectThreads, 2) k_intersect(IsectParams p) {
  __shared__ DevGeom sgeom[kMaxGeoms];
  __shared__ unsigned int shist[kMaxMaterials];
  __shared__ int sstack[USE_BVH ? kShortStack * kIsectThreads : 1];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int n_geoms = p.scene.n_geoms;
  {
    // stage the geoms: 160-byte structs copied as 16-byte words
    const float4* src = reinterpret_cast<const float4*>(p.scene.geoms);
    float4* dst = reinterpret_cast<float4*>(sgeom);
    const int words = n_geoms * (int)(sizeof(DevGeom) / 16);
    for (int i = tid; i < words; i += kIsectThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < kMaxMaterials; i += kIsectThreads) shist[i] = 0;
  }
  __syncthreads();

  const int n = p.ctr->n_live[p.depth];
  unsigned int* ticket = &p.ctr->ray_ticket[p.depth];
  int local_stack[USE_BVH ? kLocalStack : 1];

  while (true) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(ticket, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= (unsigned int)n) break;
    const int i = (int)base + lane;
    const bool valid = i < n;
    int mat = 0;
    if (valid) {
      const float4 a = p.in.s0[i];
      const float4 b = p.in.s1[i];
      const V3 o = mk(a.x, a.y, a.z), d = mk(b.x, b.y, b.z);

      float t_min = FLT_MAX;
      int hit = -1;
      int kind = 0;         // 1 box, 2 sphere (outside), 3 sphere (inside), 4 mesh
      V3 aux = mk(0, 0, 0); // box: axis normal; sphere: object-space point; mesh: (u, v, -)
      int face = -1;

      for (int g = 0; g < n_geoms; ++g) {
        const DevGeom& G = sgeom[g];
        float t = -1.0f;
        V3 taux = mk(0, 0, 0);
        int tkind = 0, tface = -1;
        if (G.type == 1 /* CUBE */) {
          // boxIntersectionTest, apps/src/intersections.h:48-90
          const V3 qo = xform(G.inv, o, 1.0f);
          const V3 qd = normalize(xform(G.inv, d, 0.0f));
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
            // getPointOnRay (:27-29) normalises the direction again
            const V3 ip = xform(G.fwd, qo + normalize(qd) * (tmin - .0001f), 1.0f);
            t = length(o - ip);
            taux = nmin;
            tkind = 1;
          }
        } else if (G.type == 0 /* SPHERE */) {
          // sphereIntersectionTest, apps/src/intersections.h:102-144
          const V3 ro = xform(G.inv, o, 1.0f);
          const V3 rd = normalize(xform(G.inv, d, 0.0f));
          const float vdd = dot(ro, rd);
          const float radicand = vdd * vdd - (dot(ro, ro) - 0.25f);
          if (!(radicand < 0)) {
            const float sq = sqrtf(radicand);
            const float first = -vdd;
            const float t1 = first + sq, t2 = first - sq;
            if (!(t1 < 0 && t2 < 0)) {
              float ts;
              if (t1 > 0 && t2 > 0) { ts = fminf(t1, t2); tkind = 2; } else { ts = fmaxf(t1, t2); tkind = 3; }
              const V3 obj = ro + normalize(rd) * (ts - .0001f);
              const V3 ip = xform(G.fwd, obj, 1.0f);
              t = length(o - ip);
              taux = obj;
            }
          }
        } else if (G.type == 3 /* OBJ */ && G.mesh >= 0) {
          // meshIntersectionTest, apps/src/intersections.h:207-282
          const DevMesh& M = p.scene.meshes[G.mesh];
          const V3 qo = xform(G.inv, o, 1.0f);
          const V3 qd = normalize(xform(G.inv, d, 0.0f));
          float bu = 0, bv = 0;
          if (USE_BVH) {
            StackRef st;
            st.sm = sstack + tid;
            st.local = local_stack;
            st.sp = 0;
            const float lim = (G.rigid && t_min < FLT_MAX) ? t_min * 1.0001f + 1e-5f : FLT_MAX;
            t = mesh_traverse(M, qo, qd, lim, &tface, &bu, &bv, st);
          } else {
            t = mesh_brute(M, qo, qd, &tface, &bu, &bv);
          }
          taux = mk(bu, bv, 0);
          tkind = 4;
        }
        if (t > 0.0f && t_min > t) {
          t_min = t;
          hit = g;
          kind = tkind;
          aux = taux;
          face = tface;
        }
      }

      float4 h0, h1;
      if (hit < 0) {
        h0 = make_float4(-1.0f, 0.0f, 0.0f, 0.0f);
        h1 = make_float4(0.0f, 0.0f, __int_as_float(0xffff), __int_as_float(-1));
      } else {
        const DevGeom& G = sgeom[hit];
        V3 nrm;
        float tu = 0.0f, tv = 0.0f;
        if (kind == 1) {
          nrm = normalize(xform(G.invT, aux, 0.0f));
        } else if (kind == 2 || kind == 3) {
          nrm = normalize(xform(G.invT, aux, 0.0f));
          if (kind == 3) nrm = -nrm;
        } else {
          const DevMesh& M = p.scene.meshes[G.mesh];
          const float* fp = M.face_pos + 9 * (size_t)face;
          const float* fu = M.face_uv + 6 * (size_t)face;
          const V3 v0 = mk(__ldg(fp), __ldg(fp + 1), __ldg(fp + 2));
          const V3 v1 = mk(__ldg(fp + 3), __ldg(fp + 4), __ldg(fp + 5));
          const V3 v2 = mk(__ldg(fp + 6), __ldg(fp + 7), __ldg(fp + 8));
          const float u0x = __ldg(fu), u0y = __ldg(fu + 1), u1x = __ldg(fu + 2), u1y = __ldg(fu + 3);
          const float u2x = __ldg(fu + 4), u2y = __ldg(fu + 5);
          const float bu = aux.x, bv = aux.y;
          const float w = 1 - bu - bv;
          tu = (w * u0x + bu * u1x) + bv * u2x;
          tv = (w * u0y + bu * u1y) + bv * u2y;
          const V3 e1 = v1 - v0, e2 = v2 - v0;
          nrm = normalize(xform(G.invT, normalize(cross(e1, e2)), 0.0f));
          if (M.bump.channels) {
            // normal map, intersections.h:245-279
            const float d1x = u1x - u0x, d1y = u1y - u0y, d2x = u2x - u0x, d2y = u2y - u0y;
            const float f = 1.0f / (d1x * d2y - d2x * d1y);
            V3 tang = mk(f * (d2y * e1.x - d1y * e2.x), f * (d2y * e1.y - d1y * e2.y), f * (d2y * e1.z - d1y * e2.z));
            tang = normalize(tang);
            V3 bit = mk(f * (-d2x * e1.x + d1x * e2.x), f * (-d2x * e1.y + d1x * e2.y),
                        f * (-d2x * e1.z + d1x * e2.z));
            bit = normalize(bit);
            const V3 T = normalize(xform(G.fwd, tang, 0.0f));
            const V3 B = normalize(xform(G.fwd, bit, 0.0f));
            V3 tsn = normalize(fetch_texel(M.bump, tu, tv));
            tsn = normalize(mk(tsn.x * 2.0f - 1.0f, tsn.y * 2.0f - 1.0f, tsn.z * 2.0f - 1.0f));
            nrm = normalize(mk((T.x * tsn.x + B.x * tsn.y) + nrm.x * tsn.z, (T.y * tsn.x + B.y * tsn.y) + nrm.y * tsn.z,
                               (T.z * tsn.x + B.z * tsn.y) + nrm.z * tsn.z));
          }
        }
        mat = G.material;
        h0 = make_float4(t_min, nrm.x, nrm.y, nrm.z);
        h1 = make_float4(tu, tv, __int_as_float((hit & 0xffff) | (mat << 16)), __int_as_float(face));
      }
      p.out.h0[i] = h0;
      p.out.h1[i] = h1;
      p.key[i] = (uint8_t)mat;
    }
    // material histogram for the one-pass sort: one shared atomic per distinct
    // material per warp
    const unsigned int active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned int peers = __match_any_sync(active, mat);
      if (lane == __ffs(peers) - 1) atomicAdd(&shist[mat], (unsigned int)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = tid; i < kMaxMaterials; i += kIsectThreads) {
    const unsigned int c = shist[i];
    if (c) atomicAdd(&p.ctr->hist[p.depth][i], c);
  }
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&p.ctr->segments, (unsigned long long)n);
}

}  // namespace b2pt
