// exp_wide_leaf.c -- CPU experiment, not part of the product: how many dependent steps does a mesh walk take with
// W-wide nodes and leaves of up to L triangles, on the rays the renderer really traces?
//
//   gcc -O2 -o /tmp/exp_wide_leaf tools/exp_wide_leaf.c -lm
//   python tools/dump_rays.py DIR            (tris.bin + rays.bin from the CPU oracle's stage dumps)
//   /tmp/exp_wide_leaf DIR/tris.bin DIR/rays.bin
//
// Builds the Karras-equivalent radix tree over 30-bit Morton codes (cubic cells, as k_lbvh.cuh), turns every
// maximal subtree of at most L triangles into one leaf (a contiguous range of the Morton order), collapses the rest
// into W-wide nodes by opening the child with the largest surface area (as k_emit_wide4 does for W = 4), and walks
// the dumped rays exactly as k_mesh_walk does: nearest child first, the other hits pushed farthest first with
// their entry distance, entries behind the closest hit dropped at pop.  Two levels of numbers:
//   per walk     node steps, leaf steps, exact triangle tests, walks longer than the hand-off threshold
//   per warp     the warp-synchronous schedule of k_mesh_walk (one kind of step per iteration, whichever has more
//                lanes waiting; 32 rays per warp, no refill = the one-wave regime of a full-width grid, or R rays
//                per lane with refill at 8 idle lanes = the shared-SM regime): iterations by kind and active lanes
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; } Box;
static int n;
static float* tri;
static Box* leafbox;
static uint64_t* key;
static int* order;
typedef struct { int l, r; } Kids;
static Kids* kids;
static Box* nbox;
static int* ncount;  // triangles under an inner node
static int* nfirst;  // first Morton slot under an inner node
static int nnodes, root;

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
static uint64_t skey(int i) { return ((key[i] >> 32) << 32) | (uint32_t)i; }
static int build_radix(int lo, int hi) {
  if (lo == hi) return ~lo;
  uint64_t a = skey(lo), b = skey(hi);
  int bit = 63 - __builtin_clzll(a ^ b);
  int L = lo, R = hi;
  while (L < R) { int m = (L + R + 1) >> 1; if ((skey(m) >> bit) & 1) R = m - 1; else L = m; }
  int me = nnodes++;
  int l = build_radix(lo, L), r = build_radix(L + 1, hi);
  kids[me].l = l; kids[me].r = r;
  ncount[me] = hi - lo + 1;
  nfirst[me] = lo;
  return me;
}
static Box refit(int c) {
  if (c < 0) return leafbox[~c];
  Box b = bunion(refit(kids[c].l), refit(kids[c].r));
  nbox[c] = b;
  return b;
}

// ---- wide nodes with fat leaves ----
#define EMPTY 0x40000000
#define MAXW 16
typedef struct { int id[MAXW]; Box b[MAXW]; int k; } Wide;  // id >= 0 inner node, < 0: ~(first << 4 | count - 1)
static Wide* wide;
static int W = 4, L = 1;
static int is_leaf_sub(int c) { return c < 0 || ncount[c] <= L; }
static int leaf_code(int c) { return c < 0 ? ~((~c) << 4) : ~((nfirst[c] << 4) | (ncount[c] - 1)); }
static Box cbox(int c) { return c < 0 ? leafbox[~c] : nbox[c]; }
static char* reach;
static long reach_nodes;
static void emit_wide(void) {
  for (int i = 0; i < nnodes; ++i) {
    int id[MAXW], k = 2;
    id[0] = kids[i].l; id[1] = kids[i].r;
    while (k < W) {
      int pick = -1; float ba = -1;
      for (int q = 0; q < k; ++q) { if (is_leaf_sub(id[q])) continue; float a = barea(nbox[id[q]]); if (a > ba) { ba = a; pick = q; } }
      if (pick < 0) break;
      Kids g = kids[id[pick]]; id[pick] = g.l; id[k++] = g.r;
    }
    wide[i].k = k;
    for (int q = 0; q < k; ++q) { wide[i].b[q] = cbox(id[q]); wide[i].id[q] = is_leaf_sub(id[q]) ? leaf_code(id[q]) : id[q]; }
  }
  memset(reach, 0, (size_t)n);
  reach_nodes = 0;
  if (!is_leaf_sub(root)) {
    int* st = malloc(sizeof(int) * (size_t)n); int sp = 0; st[sp++] = root;
    while (sp) { int c = st[--sp]; reach[c] = 1; ++reach_nodes; for (int q = 0; q < wide[c].k; ++q) if (wide[c].id[q] >= 0) st[sp++] = wide[c].id[q]; }
    free(st);
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
  if (u < 0 || u > 1) return 0;
  q[0] = s[1] * e1[2] - s[2] * e1[1]; q[1] = s[2] * e1[0] - s[0] * e1[2]; q[2] = s[0] * e1[1] - s[1] * e1[0];
  float v = f * (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]);
  if (v < 0 || u + v > 1) return 0;
  float t = f * (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]);
  if (t < 0) return 0;
  *tout = t;
  return 1;
}

// ---- one walk as a resumable state machine (so that the warp schedule can interleave walks) ----
typedef struct {
  float o[3], d[3], id[3], tb;
  int node;  // >= 0 inner, < 0 leaf code, EMPTY: done
  int sp, nn, nl, nt;
  int st[256]; float stt[256];
} Walk;
static int sorted_order = 1;  // 1: push farthest first (full sort); 0: nearest first, the rest unsorted
static void walk_begin(Walk* w, const float* ray) {
  for (int k = 0; k < 3; ++k) { w->o[k] = ray[k]; w->d[k] = ray[3 + k]; w->id[k] = 1.0f / ray[3 + k]; }
  w->tb = ray[6];
  w->sp = w->nn = w->nl = w->nt = 0;
  w->node = is_leaf_sub(root) ? leaf_code(root) : root;
}
static void walk_pop(Walk* w) {
  w->node = EMPTY;
  while (w->sp > 0) { --w->sp; if (w->stt[w->sp] <= w->tb) { w->node = w->st[w->sp]; break; } }
}
static void walk_node_step(Walk* w) {
  ++w->nn;
  Wide* nd = &wide[w->node];
  int ch[MAXW]; float tn[MAXW]; int k = 0;
  for (int q = 0; q < nd->k; ++q) {
    float a = 0, b = w->tb;
    for (int x = 0; x < 3; ++x) {
      float t0 = (nd->b[q].lo[x] - w->o[x]) * w->id[x], t1 = (nd->b[q].hi[x] - w->o[x]) * w->id[x];
      a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1));
    }
    if (a <= b) { ch[k] = nd->id[q]; tn[k] = a; ++k; }
  }
  if (k == 0) { walk_pop(w); return; }
  if (sorted_order) {
    for (int i = 1; i < k; ++i) { float t = tn[i]; int c = ch[i]; int j = i - 1; while (j >= 0 && tn[j] > t) { tn[j + 1] = tn[j]; ch[j + 1] = ch[j]; --j; } tn[j + 1] = t; ch[j + 1] = c; }
  } else {
    int m = 0; for (int i = 1; i < k; ++i) if (tn[i] < tn[m]) m = i;
    float t = tn[0]; tn[0] = tn[m]; tn[m] = t; int c = ch[0]; ch[0] = ch[m]; ch[m] = c;
  }
  for (int q = k - 1; q >= 1; --q) { w->st[w->sp] = ch[q]; w->stt[w->sp++] = tn[q]; }
  w->node = ch[0];
}
static void walk_leaf_step(Walk* w) {
  ++w->nl;
  const int code = ~w->node, first = code >> 4, cnt = (code & 15) + 1;
  for (int j = 0; j < cnt; ++j) {
    float t;
    ++w->nt;
    if (tri_hit(w->o, w->d, tri + 9 * (size_t)order[first + j], &t) && t < w->tb) w->tb = t;
  }
  walk_pop(w);
}

static float* rays; static int nrays;
static int handoff = 24;

static void per_walk(const char* name) {
  double sn = 0, sl = 0, st = 0; long over = 0, beyond = 0; int mx = 0;
  long hist[64] = {0};
  for (int r = 0; r < nrays; ++r) {
    Walk w; walk_begin(&w, rays + 8 * (size_t)r);
    while (w.node != EMPTY) { if (w.node >= 0) walk_node_step(&w); else walk_leaf_step(&w); }
    sn += w.nn; sl += w.nl; st += w.nt;
    int steps = w.nn + w.nl;
    if (steps > handoff) { ++over; beyond += steps - handoff; }
    if (steps > mx) mx = steps;
    ++hist[steps < 63 ? steps : 63];
  }
  printf("%-14s nodes %5.2f leaves %5.2f tris %5.2f steps %5.2f | > %d steps: %5.2f %% (+%.2f/walk) max %d | wide nodes %ld (%.1f MB at %d B)\n",
         name, sn / nrays, sl / nrays, st / nrays, (sn + sl) / nrays, handoff, 100.0 * over / nrays, (double)beyond / nrays, mx,
         reach_nodes, reach_nodes * (W <= 4 ? 128.0 : 256.0) / 1e6, W <= 4 ? 128 : 256);
}

static int refill_min = 8;  // idle lanes that trigger a refill (EXP_REFILL_MIN)
// The warp-synchronous schedule of k_mesh_walk.  rays_per_lane = 1: every warp gets 32 rays and never refills.
static void per_warp(const char* name, int rays_per_lane, int fused) {
  const int nwarps = (nrays + 32 * rays_per_lane - 1) / (32 * rays_per_lane);
  long it_node = 0, it_leaf = 0, it_both = 0, lanes_node = 0, lanes_leaf = 0, refills = 0, handed = 0; long maxit = 0;
  int next = 0;
  // static first batch per warp (as the kernel), then a shared queue
  Walk* w = malloc(sizeof(Walk) * 32);
  int steps[32];
  int qhead = nwarps * 32 < nrays ? nwarps * 32 : nrays;
  for (int wp = 0; wp < nwarps; ++wp) {
    int active[32];
    for (int l = 0; l < 32; ++l) {
      int r = wp * 32 + l;
      active[l] = r < nrays && r < nwarps * 32;
      if (active[l]) { walk_begin(&w[l], rays + 8 * (size_t)r); steps[l] = 0; if (w[l].node == EMPTY) active[l] = 0; }
    }
    long its = 0;
    (void)next;
    while (1) {
      int nn_ = 0, nl_ = 0, idle = 0;
      for (int l = 0; l < 32; ++l) { if (!active[l]) { ++idle; continue; } if (w[l].node >= 0) ++nn_; else ++nl_; }
      if (nn_ + nl_ == 0 || (idle >= refill_min && qhead < nrays && rays_per_lane > 1)) {
        int got = 0;
        for (int l = 0; l < 32 && qhead < nrays && rays_per_lane > 1; ++l)
          if (!active[l]) { walk_begin(&w[l], rays + 8 * (size_t)qhead++); steps[l] = 0; active[l] = w[l].node != EMPTY; ++got; }
        if (got) { ++refills; continue; }
        if (nn_ + nl_ == 0) break;
      }
      ++its;
      if (fused) {  // one iteration serves both kinds (costs both code paths)
        ++it_both; lanes_node += nn_; lanes_leaf += nl_;
        for (int l = 0; l < 32; ++l) if (active[l]) { if (w[l].node >= 0) walk_node_step(&w[l]); else walk_leaf_step(&w[l]); ++steps[l]; }
      } else if (nn_ >= nl_) {
        ++it_node; lanes_node += nn_;
        for (int l = 0; l < 32; ++l) if (active[l] && w[l].node >= 0) { walk_node_step(&w[l]); ++steps[l]; }
      } else {
        ++it_leaf; lanes_leaf += nl_;
        for (int l = 0; l < 32; ++l) if (active[l] && w[l].node < 0) { walk_leaf_step(&w[l]); ++steps[l]; }
      }
      for (int l = 0; l < 32; ++l) if (active[l]) { if (w[l].node == EMPTY) active[l] = 0; else if (steps[l] > handoff) { active[l] = 0; ++handed; } }
    }
    if (its > maxit) maxit = its;
  }
  free(w);
  const long its = it_node + it_leaf + it_both;
  printf("  %-22s warps %5d  iterations/warp %6.1f (node %5.1f leaf %5.1f) max %ld  lanes/node-it %4.1f lanes/leaf-it %4.1f  refills/warp %.1f  handed off %.2f %%\n",
         name, nwarps, (double)its / nwarps, (double)(it_node + it_both) / nwarps, (double)(it_leaf) / nwarps, maxit,
         (it_node + it_both) ? (double)lanes_node / (it_node + it_both) : 0.0, (it_leaf + it_both) ? (double)lanes_leaf / (it_leaf + it_both) : 0.0,
         (double)refills / nwarps, 100.0 * handed / nrays);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: exp_wide_leaf tris.bin rays.bin [handoff]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  n = (int)(sz / 36); tri = malloc(sz);
  if (fread(tri, 1, sz, f) != (size_t)sz) return 1;
  fclose(f);
  f = fopen(argv[2], "rb");
  fseek(f, 0, SEEK_END); sz = ftell(f); fseek(f, 0, SEEK_SET);
  nrays = (int)(sz / 32); rays = malloc(sz);
  if (fread(rays, 1, sz, f) != (size_t)sz) return 1;
  fclose(f);
  if (argc > 3) handoff = atoi(argv[3]);
  if (getenv("EXP_REFILL_MIN")) refill_min = atoi(getenv("EXP_REFILL_MIN"));
  leafbox = malloc(sizeof(Box) * n); key = malloc(8 * (size_t)n); order = malloc(4 * (size_t)n);
  kids = malloc(sizeof(Kids) * n); nbox = malloc(sizeof(Box) * n); wide = malloc(sizeof(Wide) * (size_t)n);
  ncount = malloc(4 * (size_t)n); nfirst = malloc(4 * (size_t)n); reach = malloc(n);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  float* cen = malloc(12 * (size_t)n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      float a = fminf(fminf(tri[9 * i + k], tri[9 * i + 3 + k]), tri[9 * i + 6 + k]);
      float b = fmaxf(fmaxf(tri[9 * i + k], tri[9 * i + 3 + k]), tri[9 * i + 6 + k]);
      cen[3 * i + k] = 0.5f * (a + b);
      lo[k] = fminf(lo[k], cen[3 * i + k]); hi[k] = fmaxf(hi[k], cen[3 * i + k]);
    }
  float m = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
  for (int i = 0; i < n; ++i) {
    uint32_t q[3];
    for (int k = 0; k < 3; ++k) q[k] = (uint32_t)fminf(fmaxf((cen[3 * i + k] - lo[k]) / m * 1024.0f, 0.0f), 1023.0f);
    key[i] = ((uint64_t)((expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2])) << 32) | (uint32_t)i;
  }
  free(cen);
  qsort(key, n, 8, cmp64);
  for (int i = 0; i < n; ++i) {
    int fi = (int)(uint32_t)key[i];
    order[i] = fi;
    for (int k = 0; k < 3; ++k) {
      leafbox[i].lo[k] = fminf(fminf(tri[9 * fi + k], tri[9 * fi + 3 + k]), tri[9 * fi + 6 + k]) - 1.2e-4f;
      leafbox[i].hi[k] = fmaxf(fmaxf(tri[9 * fi + k], tri[9 * fi + 3 + k]), tri[9 * fi + 6 + k]) + 1.2e-4f;
    }
  }
  nnodes = 0;
  root = build_radix(0, n - 1);
  refit(root);
  printf("%d triangles, %d rays, hand-off after %d steps\n", n, nrays, handoff);
  const int Ws[] = {4, 6, 8}, Ls[] = {1, 2, 4, 8};
  const int only = getenv("EXP_ONLY") != NULL;  // EXP_ONLY=1: the shipped shape (W = 4, L = 2) only
  for (int a = 0; a < (only ? 1 : 3); ++a)
    for (int b = (only ? 1 : 0); b < (only ? 2 : 4); ++b) {
      W = Ws[a]; L = Ls[b];
      emit_wide();
      char nm[64];
      snprintf(nm, sizeof nm, "W=%d L=%d", W, L);
      sorted_order = getenv("EXP_UNSORTED") ? 0 : 1;  // EXP_UNSORTED=1: nearest child first, the other hits pushed in slot order
      per_walk(nm);
      per_warp("one wave, alternating", 1, 0);
      per_warp("one wave, fused step", 1, 1);
      per_warp("3 rays/lane, alternating", 3, 0);
      per_warp("3 rays/lane, fused", 3, 1);
      per_warp("12 rays/lane, alternating", 12, 0);
    }
  return 0;
}
