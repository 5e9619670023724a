// exp_tree_quality.c -- CPU experiment, not part of the product: how much better than the Karras LBVH could
// the hierarchy over the same Morton order be?  Builds (A) the Karras-equivalent radix tree and (B) a PLOC tree
// (Meister & Bittner 2018: mutual nearest neighbours inside a window of the Morton order) over the triangles in
// tris.bin (n x 9 float32), collapses both to 4-wide nodes the way k_emit_wide4 does and walks sample rays
// (closest hit, nearest child first, entries behind the hit culled at pop) counting node and triangle steps.
//   gcc -O2 -o exp_tree_quality tools/exp_tree_quality.c -lm && ./exp_tree_quality tris.bin [radius] [uniform]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; } Box;
static int n;
static float* tri;
static Box* leafbox;
static uint64_t* key;  // code << 32 | face
static int* order;

static Box bunion(Box a, Box b) {
  Box r;
  for (int k = 0; k < 3; ++k) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
  return r;
}
static float barea(Box b) {
  float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return dx * dy + dy * dz + dz * dx;
}
static uint32_t expand10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static int cmp64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

// binary tree: inner nodes 0..n-2, child >= 0 inner, < 0 ~leaf slot (slot = position in Morton order)
typedef struct { int l, r; } Kids;
static Kids* kids;
static Box* nbox;
static int nnodes, root;

static uint64_t skey(int i) { return ((key[i] >> 32) << 32) | (uint32_t)i; }  // code, ties by index
static int build_radix(int lo, int hi) {  // leaves lo..hi inclusive
  if (lo == hi) return ~lo;
  uint64_t a = skey(lo), b = skey(hi);
  int bit = 63 - __builtin_clzll(a ^ b);
  int L = lo, R = hi;  // last index with bit == 0
  while (L < R) { int m = (L + R + 1) >> 1; if ((skey(m) >> bit) & 1) R = m - 1; else L = m; }
  int me = nnodes++;
  int l = build_radix(lo, L), r = build_radix(L + 1, hi);
  kids[me].l = l; kids[me].r = r;
  return me;
}
static Box getbox(int c) { return c < 0 ? leafbox[~c] : nbox[c]; }
static Box refit(int c) {
  if (c < 0) return leafbox[~c];
  Box b = bunion(refit(kids[c].l), refit(kids[c].r));
  nbox[c] = b;
  return b;
}
static void build_ploc(int radius) {
  int m = n;
  int* C = malloc(sizeof(int) * n), *C2 = malloc(sizeof(int) * n), *nn = malloc(sizeof(int) * n);
  for (int i = 0; i < n; ++i) C[i] = ~i;
  nnodes = 0;
  int iters = 0;
  while (m > 1) {
    for (int i = 0; i < m; ++i) {
      Box bi = getbox(C[i]);
      float best = INFINITY; int bj = -1;
      int j0 = i - radius < 0 ? 0 : i - radius, j1 = i + radius >= m ? m - 1 : i + radius;
      for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        float a = barea(bunion(bi, getbox(C[j])));
        if (a < best) { best = a; bj = j; }
      }
      nn[i] = bj;
    }
    int o = 0;
    for (int i = 0; i < m; ++i) {
      int j = nn[i];
      if (nn[j] == i) {
        if (i < j) {
          int me = nnodes++;
          kids[me].l = C[i]; kids[me].r = C[j];
          nbox[me] = bunion(getbox(C[i]), getbox(C[j]));
          C2[o++] = me;
        }
      } else C2[o++] = C[i];
    }
    int* t = C; C = C2; C2 = t;
    m = o; ++iters;
  }
  root = C[0];
  fprintf(stderr, "ploc: %d iterations, %d nodes\n", iters, nnodes);
  free(C); free(C2); free(nn);
}
static double sah_cost(void) {
  double ar = barea(nbox[root]), c = 0;
  for (int i = 0; i < nnodes; ++i) c += barea(nbox[i]) / ar;
  double cl = 0;
  for (int i = 0; i < n; ++i) cl += barea(leafbox[i]) / ar;
  return c + cl;
}
// ---- 4-wide collapse as in k_emit_wide4 (surface-area greedy) ----
#define EMPTY 0x40000000
typedef struct { int id[4]; Box b[4]; } Wide;
static Wide* wide;
static void emit_wide(void) {
  for (int i = 0; i < nnodes; ++i) {
    int id[4] = {kids[i].l, kids[i].r, EMPTY, EMPTY}, k = 2;
    for (int round = 0; round < 2; ++round) {
      int pick = -1; float ba = -1;
      for (int q = 0; q < k; ++q) { if (id[q] < 0) continue; float a = barea(nbox[id[q]]); if (a > ba) { ba = a; pick = q; } }
      if (pick < 0) break;
      Kids g = kids[id[pick]]; id[pick] = g.l; id[k++] = g.r;
    }
    for (int q = 0; q < 4; ++q) { wide[i].id[q] = id[q]; if (id[q] != EMPTY) wide[i].b[q] = getbox(id[q]); }
  }
}
static int tri_hit(const float* o, const float* d, const float* T, float* tout) {
  float e1[3], e2[3], p[3], s[3], q[3];
  for (int k = 0; k < 3; ++k) { e1[k] = T[3 + k] - T[k]; e2[k] = T[6 + k] - T[k]; }
  p[0] = d[1] * e2[2] - d[2] * e2[1]; p[1] = d[2] * e2[0] - d[0] * e2[2]; p[2] = d[0] * e2[1] - d[1] * e2[0];
  float a = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
  if (a < 1.1920929e-7f) return 0;
  float f = 1.0f / a;
  for (int k = 0; k < 3; ++k) s[k] = o[k] - T[k];
  float u = f * (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]);
  if (u < 0) return 0;
  q[0] = s[1] * e1[2] - s[2] * e1[1]; q[1] = s[2] * e1[0] - s[0] * e1[2]; q[2] = s[0] * e1[1] - s[1] * e1[0];
  float v = f * (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]);
  if (v < 0 || u + v > 1) return 0;
  float t = f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
  if (t <= 0) return 0;
  *tout = t;
  return 1;
}
static uint64_t rs = 88172645463325252ull;
static double rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (rs >> 11) * (1.0 / 9007199254740992.0); }
static void walk(const float* o, const float* d, int* nodes_out, int* tris_out) {
  static int st[4096]; static float stt[4096];
  int sp = 0, nn_ = 0, nt = 0;
  float tb = INFINITY;
  int node = root;
  float id[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
  while (1) {
    if (node >= 0) {
      ++nn_;
      Wide* w = &wide[node];
      int ch[4]; float tn[4];
      for (int q = 0; q < 4; ++q) {
        ch[q] = EMPTY; tn[q] = INFINITY;
        if (w->id[q] == EMPTY) continue;
        float a = 0, b = tb;
        for (int k = 0; k < 3; ++k) {
          float t0 = (w->b[q].lo[k] - o[k]) * id[k], t1 = (w->b[q].hi[k] - o[k]) * id[k];
          a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1));
        }
        if (a <= b) { ch[q] = w->id[q]; tn[q] = a; }
      }
#define CS(a, b) if (tn[b] < tn[a]) { float tt = tn[a]; tn[a] = tn[b]; tn[b] = tt; int cc = ch[a]; ch[a] = ch[b]; ch[b] = cc; }
      CS(0, 1) CS(2, 3) CS(0, 2) CS(1, 3) CS(1, 2)
      if (ch[0] != EMPTY) {
        for (int q = 3; q >= 1; --q) if (ch[q] != EMPTY) { st[sp] = ch[q]; stt[sp++] = tn[q]; }
        node = ch[0];
        continue;
      }
    } else {
      ++nt;
      float t;
      if (tri_hit(o, d, tri + 9 * (size_t)order[~node], &t) && t < tb) tb = t;
    }
    node = EMPTY;
    while (sp > 0) { --sp; if (stt[sp] <= tb) { node = st[sp]; break; } }
    if (node == EMPTY) break;
  }
  *nodes_out = nn_; *tris_out = nt;
}
static void sample(const char* name) {
  emit_wide();
  Box rb = nbox[root];
  float c[3], e[3];
  for (int k = 0; k < 3; ++k) { c[k] = 0.5f * (rb.lo[k] + rb.hi[k]); e[k] = 0.5f * (rb.hi[k] - rb.lo[k]); }
  rs = 88172645463325252ull;
  const int R = 400000;
  double sn = 0, st_ = 0; long over24 = 0, over24steps = 0; int mx = 0;
  for (int r = 0; r < R; ++r) {
    float o[3], d[3];
    if (r & 1) {  // from outside towards the box
      double z = 2 * rnd() - 1, ph = 6.283185307 * rnd(), s = sqrt(1 - z * z);
      o[0] = c[0] + 8 * s * cos(ph); o[1] = c[1] + 8 * s * sin(ph); o[2] = c[2] + 8 * z;
      for (int k = 0; k < 3; ++k) d[k] = c[k] + 1.3f * e[k] * (float)(2 * rnd() - 1) - o[k];
    } else {  // leaving the surface: cosine-ish hemisphere about the geometric normal
      int f = (int)(rnd() * n); if (f >= n) f = n - 1;
      const float* T = tri + 9 * (size_t)f;
      double u = rnd(), v = rnd(); if (u + v > 1) { u = 1 - u; v = 1 - v; }
      float e1[3], e2[3], nr[3];
      for (int k = 0; k < 3; ++k) { e1[k] = T[3 + k] - T[k]; e2[k] = T[6 + k] - T[k]; }
      nr[0] = e1[1] * e2[2] - e1[2] * e2[1]; nr[1] = e1[2] * e2[0] - e1[0] * e2[2]; nr[2] = e1[0] * e2[1] - e1[1] * e2[0];
      float ln = sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
      for (int k = 0; k < 3; ++k) nr[k] /= ln;
      double z = 2 * rnd() - 1, ph = 6.283185307 * rnd(), s = sqrt(1 - z * z);
      d[0] = nr[0] + (float)(s * cos(ph)); d[1] = nr[1] + (float)(s * sin(ph)); d[2] = nr[2] + (float)z;
      for (int k = 0; k < 3; ++k) o[k] = T[k] + (float)u * e1[k] + (float)v * e2[k];
      float ld0 = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      if (ld0 < 1e-3f) { d[0] = nr[0]; d[1] = nr[1]; d[2] = nr[2]; ld0 = 1; }
      for (int k = 0; k < 3; ++k) o[k] += 0.01f * d[k] / ld0;
    }
    float ld = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int k = 0; k < 3; ++k) d[k] /= ld;
    int a, b;
    walk(o, d, &a, &b);
    sn += a; st_ += b;
    if (a + b > 24) { ++over24; over24steps += a + b - 24; }
    if (a + b > mx) mx = a + b;
  }
  printf("%-28s SAH %.2f  nodes/walk %.2f  tris/walk %.2f  walks over 24 steps %.2f %%  (steps beyond 24: %.2f per walk)  max %d\n",
         name, sah_cost(), sn / R, st_ / R, 100.0 * over24 / R, (double)over24steps / R, mx);
}
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb");
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  n = (int)(sz / 36);
  tri = malloc(sz);
  if (fread(tri, 1, sz, f) != (size_t)sz) return 1;
  fclose(f);
  int radius = argc > 2 ? atoi(argv[2]) : 16;
  leafbox = malloc(sizeof(Box) * n); key = malloc(8 * n); order = malloc(4 * n);
  kids = malloc(sizeof(Kids) * n); nbox = malloc(sizeof(Box) * n); wide = malloc(sizeof(Wide) * n);
  for (int uniform = 0; uniform < 2; ++uniform) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float* cen = malloc(12 * (size_t)n);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        float a = fminf(fminf(tri[9 * i + k], tri[9 * i + 3 + k]), tri[9 * i + 6 + k]);
        float b = fmaxf(fmaxf(tri[9 * i + k], tri[9 * i + 3 + k]), tri[9 * i + 6 + k]);
        cen[3 * i + k] = 0.5f * (a + b);
        lo[k] = fminf(lo[k], cen[3 * i + k]); hi[k] = fmaxf(hi[k], cen[3 * i + k]);
      }
    float ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    if (uniform) { float m = fmaxf(ext[0], fmaxf(ext[1], ext[2])); ext[0] = ext[1] = ext[2] = m; }
    for (int i = 0; i < n; ++i) {
      uint32_t q[3];
      for (int k = 0; k < 3; ++k) q[k] = (uint32_t)fminf(fmaxf((cen[3 * i + k] - lo[k]) / ext[k] * 1024.0f, 0.0f), 1023.0f);
      uint32_t code = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
      key[i] = ((uint64_t)code << 32) | (uint32_t)i;
    }
    free(cen);
    qsort(key, n, 8, cmp64);
    for (int i = 0; i < n; ++i) {
      int fidx = (int)(uint32_t)key[i];
      order[i] = fidx;
      for (int k = 0; k < 3; ++k) {
        leafbox[i].lo[k] = fminf(fminf(tri[9 * fidx + k], tri[9 * fidx + 3 + k]), tri[9 * fidx + 6 + k]) - 1.2e-4f;
        leafbox[i].hi[k] = fmaxf(fmaxf(tri[9 * fidx + k], tri[9 * fidx + 3 + k]), tri[9 * fidx + 6 + k]) + 1.2e-4f;
      }
    }
    nnodes = 0;
    root = build_radix(0, n - 1);
    refit(root);
    sample(uniform ? "karras, uniform cells" : "karras, per-axis cells");
    build_ploc(radius);
    char nm[64];
    snprintf(nm, sizeof nm, "ploc r=%d, %s", radius, uniform ? "uniform" : "per-axis");
    sample(nm);
  }
  return 0;
}
