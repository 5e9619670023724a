// k_sort.cuh -- single-pass radix partitioning (onesweep) and its two users.
//
// (1) Material sort.  Replaces thrust::sort_by_key(dev_intersections, dev_paths,
//     sortByMaterial()) (apps/src/pathtrace.cu:512-516,612): a STABLE sort by
//     DESCENDING materialId whose 32-byte keys and 44-byte values the reference
//     moves through every merge pass.  Here the key is one byte, written by
//     k_intersect together with a per-depth material histogram, and the output
//     is only the permutation perm[sorted slot] = pre-sort slot (4 B/path); the
//     payload is gathered once, by the shade kernel.  The stable order by
//     (255 - material) is the unique stable descending order, so the
//     permutation is identical to thrust's.
// (2) LSD radix sort of (Morton code, face) pairs for the LBVH build: four
//     8-bit passes of the same kernel.
//
// One pass = one kernel: a tile ranks its 4096 keys locally (warp match_any),
// publishes its per-bin counts, and resolves its per-bin exclusive prefix by
// decoupled look-back over the tiles before it.  Tiles are handed out by a
// device ticket so a tile's predecessors are always resident or finished.
// Bin bases come from a histogram computed beforehand (by k_intersect for the
// material sort, by k_radix_hist for the LSD sort).
//
// Kernels: k_onesweep_pass<Policy> (the LSD passes, and the stand-alone b2pt_sort_desc_perm), k_sort_material (the
// renderer's material sort with compaction ranks, any number of materials), k_sort_material_few (the same result for
// scenes with at most 8 materials: packed counters instead of match_any rounds, what every shipped scene uses),
// k_rank_live (SORT_BY_MATERIAL 0 and the pixel-keyed RNG: compaction ranks only).
#pragma once

#include "k_prims.cuh"
#include "pt_device.cuh"

namespace b2pt {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRows = 16;                           // 32-key rows per warp
constexpr int kSortTile = kSortWarps * kSortRows * 32;  // 4096 keys
static_assert(kSortThreads == kMaxMaterials, "one thread per bin");

// Policy for the material sort of depth `depth`.
struct MaterialSortPolicy {
  const uint8_t* key;
  int* perm;
  Counters* ctr;
  unsigned long long* status_;
  int depth;
  __device__ __forceinline__ int n() const { return ctr->n_live[depth]; }
  __device__ __forceinline__ unsigned int* ticket() const { return &ctr->sort_ticket[depth]; }
  __device__ __forceinline__ unsigned long long* status() const { return status_; }
  __device__ __forceinline__ unsigned int epoch() const {
    return ctr->serial * (unsigned int)(kMaxDepth + 1) + (unsigned int)depth + 1u;
  }
  // ascending digit == descending material
  __device__ __forceinline__ unsigned int digit(int idx) const { return 255u - (unsigned int)__ldg(key + idx); }
  __device__ __forceinline__ unsigned int bin_total(int bin) const { return ctr->hist[depth][255 - bin]; }
  __device__ __forceinline__ void scatter(int idx, unsigned int pos) const { perm[pos] = idx; }
  static constexpr bool kFewBins = true;  // a handful of materials: one warp looks back per bin
};

// Policy for one 8-bit pass of the LSD pair sort.
struct RadixPassPolicy {
  const uint32_t* key_in;
  const uint32_t* val_in;
  uint32_t* key_out;
  uint32_t* val_out;
  const unsigned int* hist;  // 256 bin totals of this digit
  unsigned int* ticket_;
  unsigned long long* status_;
  unsigned int epoch_;
  int n_;
  int shift;
  __device__ __forceinline__ int n() const { return n_; }
  __device__ __forceinline__ unsigned int* ticket() const { return ticket_; }
  __device__ __forceinline__ unsigned long long* status() const { return status_; }
  __device__ __forceinline__ unsigned int epoch() const { return epoch_; }
  __device__ __forceinline__ unsigned int digit(int idx) const { return (__ldg(key_in + idx) >> shift) & 255u; }
  __device__ __forceinline__ unsigned int bin_total(int bin) const { return hist[bin]; }
  __device__ __forceinline__ void scatter(int idx, unsigned int pos) const {
    key_out[pos] = key_in[idx];
    val_out[pos] = val_in[idx];
  }
  static constexpr bool kFewBins = false;  // all 256 bins are populated: one thread per bin
};

template <typename Policy>
__global__ void __launch_bounds__(kSortThreads) k_onesweep_pass(Policy p) {
  __shared__ unsigned int wcnt[kSortWarps][256];
  __shared__ unsigned int bin_base[256];
  __shared__ unsigned int tile_excl[256];
  __shared__ unsigned int warp_tot[kSortWarps];
  __shared__ unsigned int s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n();
  if (tid == 0) s_tile = atomicAdd(p.ticket(), 1u);
  for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const unsigned int tile = s_tile;
  if ((long long)tile * kSortTile >= (long long)n) return;
  const unsigned int epoch = p.epoch();

  // ---- local ranking: warp w owns 512 consecutive keys, 32 at a time ----------
  const int wbase = (int)tile * kSortTile + warp * (kSortRows * 32);
  unsigned short lrank[kSortRows];
  unsigned short ldig[kSortRows];
  // every key of the tile is loaded before the ranking chain starts (see k_sort_material)
  unsigned short draw[kSortRows];
#pragma unroll
  for (int r = 0; r < kSortRows; ++r) {
    const int idx = wbase + r * 32 + lane;
    draw[r] = idx < n ? (unsigned short)p.digit(idx) : (unsigned short)0;
  }
#pragma unroll
  for (int r = 0; r < kSortRows; ++r) {
    const int idx = wbase + r * 32 + lane;
    const bool valid = idx < n;
    const unsigned int dig = valid ? (unsigned int)draw[r] : 256u + (unsigned int)lane;  // invalid lanes match nobody
    const unsigned int peers = __match_any_sync(0xffffffffu, dig);
    unsigned int before = 0;
    if (valid) before = wcnt[warp][dig];
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) wcnt[warp][dig] = before + __popc(peers);
    __syncwarp();
    lrank[r] = (unsigned short)(before + __popc(peers & ((1u << lane) - 1u)));
    ldig[r] = (unsigned short)dig;
  }
  __syncthreads();

  // ---- per bin: offsets of the warps inside the tile, then look-back -----------
  __shared__ unsigned int tile_cnt[256];
  __shared__ unsigned int bin_tot[256];
  {
    const int b = tid;  // one thread per bin
    unsigned int run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const unsigned int c = wcnt[w][b];
      wcnt[w][b] = run;
      run += c;
    }
    const unsigned int total = p.bin_total(b);
    tile_cnt[b] = run;
    bin_tot[b] = total;
    unsigned int excl = 0;
    if (!Policy::kFewBins && total != 0) {  // bins nobody holds are never read
      unsigned long long* mine = p.status() + (size_t)tile * 256 + b;
      st_volatile_u64(mine, lb_pack(epoch, tile == 0 ? 2u : 1u, run));
      if (tile > 0) {
        for (int t = (int)tile - 1; t >= 0; --t) {
          const unsigned long long* theirs = p.status() + (size_t)t * 256 + b;
          unsigned long long w;
          do {
            w = ld_volatile_u64(theirs);
          } while (lb_epoch(w) != epoch || lb_flag(w) == 0u);
          excl += lb_value(w);
          if (lb_flag(w) == 2u) break;
        }
        st_volatile_u64(mine, lb_pack(epoch, 2u, excl + run));
      }
    }
    tile_excl[b] = excl;
    // exclusive scan of the bin totals -> bin bases
    unsigned int v = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    unsigned int add = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) add += (w < warp) ? warp_tot[w] : 0u;
    bin_base[b] = v - total + add;
  }
  if (Policy::kFewBins) {
    // warp w resolves bins w, w+8, ...; only bins that exist anywhere are looked back
    for (int b = warp; b < 256; b += kSortWarps) {
      if (bin_tot[b] == 0) continue;  // warp-uniform
      const unsigned int e = lookback_warp(p.status() + b, 256, tile, epoch, tile_cnt[b]);
      if (lane == 0) tile_excl[b] = e;
    }
  }
  __syncthreads();

  // ---- scatter ----------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < kSortRows; ++r) {
    const int idx = wbase + r * 32 + lane;
    if (idx < n) {
      const unsigned int dig = ldig[r];
      p.scatter(idx, bin_base[dig] + tile_excl[dig] + wcnt[warp][dig] + lrank[r]);
    }
  }
}

// All four 8-bit digit histograms of a key array in one read.
__global__ void __launch_bounds__(256) k_radix_hist(const uint32_t* __restrict__ key, int n, unsigned int* hist /*[4][256]*/) {
  __shared__ unsigned int sh[4][256];
  for (int i = threadIdx.x; i < 1024; i += 256) (&sh[0][0])[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint32_t k = key[i];
    atomicAdd(&sh[0][k & 255u], 1u);
    atomicAdd(&sh[1][(k >> 8) & 255u], 1u);
    atomicAdd(&sh[2][(k >> 16) & 255u], 1u);
    atomicAdd(&sh[3][(k >> 24) & 255u], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += 256) {
    const unsigned int c = (&sh[0][0])[i];
    if (c) atomicAdd(&hist[i], c);
  }
}

// Material sort + compaction ranks in one pass.
//
// For every path the kernel produces its sorted slot j (stable, descending
// material: the permutation of thrust::sort_by_key at pathtrace.cu:612) AND the
// number A of SURVIVING paths in front of it in that sorted order.  k_intersect
// already decided which paths survive the shade (will_survive), so
//   survivor at sorted slot j  -> slot A of the next depth   (the live prefix of
//   dead path at sorted slot j -> place j - A of the dead tail    stable_partition, :649)
// and the shade kernel needs no scan, no barrier and no look-back of its own.
// Ranks inside a tile come from warp match_any on the material digit, with a
// ballot of the survivors folded in; tile prefixes come from two decoupled
// look-backs per material that exists (all paths / survivors).
struct MatSortParams {
  const uint8_t* key;
  const uint8_t* live;
  int* perm;   // perm[j] = pre-sort slot
  int* apos;   // apos[j] = survivors in front of sorted slot j
  Counters* ctr;
  unsigned long long* status;       // [tiles][256] all paths
  unsigned long long* status_live;  // [tiles][256] survivors
  int depth;
};

__global__ void __launch_bounds__(kSortThreads) k_sort_material(MatSortParams p) {
  __shared__ unsigned int wcnt[kSortWarps][256];
  __shared__ unsigned int wlive[kSortWarps][256];
  __shared__ unsigned int bin_base[256], live_base[256], tile_excl[256], tile_lexcl[256];
  __shared__ unsigned int tile_cnt[256], tile_lcnt[256], bin_tot[256], bin_ltot[256];
  __shared__ unsigned int warp_tot[kSortWarps], warp_ltot[kSortWarps];
  __shared__ unsigned int s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.ctr->n_live[p.depth];
  if (tid == 0) s_tile = atomicAdd(&p.ctr->sort_ticket[p.depth], 1u);
  for (int i = tid; i < kSortWarps * 256; i += kSortThreads) {
    (&wcnt[0][0])[i] = 0;
    (&wlive[0][0])[i] = 0;
  }
  __syncthreads();
  const unsigned int tile = s_tile;
  if ((long long)tile * kSortTile >= (long long)n) return;
  const unsigned int epoch = p.ctr->serial * (unsigned int)(kMaxDepth + 1) + (unsigned int)p.depth + 1u;

  const int wbase = (int)tile * kSortTile + warp * (kSortRows * 32);
  unsigned short lrank[kSortRows], larank[kSortRows], ldig[kSortRows];
  // all 2 x 16 loads of the tile are issued before the first rank is computed: the ranking loop below is a chain
  // of warp barriers, and a load inside it would add one memory round trip per row
  unsigned char kraw[kSortRows], lraw[kSortRows];
#pragma unroll
  for (int r = 0; r < kSortRows; ++r) {
    const int idx = wbase + r * 32 + lane;
    kraw[r] = idx < n ? __ldg(p.key + idx) : (unsigned char)0;
    lraw[r] = idx < n ? __ldg(p.live + idx) : (unsigned char)0;
  }
#pragma unroll
  for (int r = 0; r < kSortRows; ++r) {
    const int idx = wbase + r * 32 + lane;
    const bool valid = idx < n;
    const unsigned int dig = valid ? 255u - (unsigned int)kraw[r] : 256u + (unsigned int)lane;
    const bool alive = valid && lraw[r] != 0;
    const unsigned int peers = __match_any_sync(0xffffffffu, dig);
    const unsigned int lpeers = peers & __ballot_sync(0xffffffffu, alive);
    unsigned int before = 0, lbefore = 0;
    if (valid) {
      before = wcnt[warp][dig];
      lbefore = wlive[warp][dig];
    }
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) {
      wcnt[warp][dig] = before + __popc(peers);
      wlive[warp][dig] = lbefore + __popc(lpeers);
    }
    __syncwarp();
    const unsigned int lt = (1u << lane) - 1u;
    lrank[r] = (unsigned short)(before + __popc(peers & lt));
    larank[r] = (unsigned short)(lbefore + __popc(lpeers & lt));
    ldig[r] = (unsigned short)dig;
  }
  __syncthreads();
  {
    const int b = tid;  // one thread per bin
    unsigned int run = 0, lrun = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const unsigned int c = wcnt[w][b], cl = wlive[w][b];
      wcnt[w][b] = run;
      wlive[w][b] = lrun;
      run += c;
      lrun += cl;
    }
    const unsigned int total = p.ctr->hist[p.depth][255 - b];
    const unsigned int ltotal = p.ctr->hist_live[p.depth][255 - b];
    tile_cnt[b] = run;
    tile_lcnt[b] = lrun;
    bin_tot[b] = total;
    bin_ltot[b] = ltotal;
    tile_excl[b] = 0;
    tile_lexcl[b] = 0;
    // exclusive scans of the two histograms over the bins -> bases
    unsigned int v = total, vl = ltotal;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, v, o);
      const unsigned int tl = __shfl_up_sync(0xffffffffu, vl, o);
      if (lane >= o) {
        v += t;
        vl += tl;
      }
    }
    if (lane == 31) {
      warp_tot[warp] = v;
      warp_ltot[warp] = vl;
    }
    __syncthreads();
    unsigned int add = 0, ladd = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      add += (w < warp) ? warp_tot[w] : 0u;
      ladd += (w < warp) ? warp_ltot[w] : 0u;
    }
    bin_base[b] = v - total + add;
    live_base[b] = vl - ltotal + ladd;
    // the number of survivors is the live count of the next depth
    if (tile == 0 && b == 255) p.ctr->n_live[p.depth + 1] = (int)(vl + ladd);
  }
  // warp w resolves bins w, w+8, ...; only bins that exist anywhere are looked back
  for (int b = warp; b < 256; b += kSortWarps) {
    if (bin_tot[b] == 0) continue;  // warp-uniform
    const unsigned int e = lookback_warp(p.status + b, 256, tile, epoch, tile_cnt[b]);
    unsigned int el = 0;
    if (bin_ltot[b] != 0) el = lookback_warp(p.status_live + b, 256, tile, epoch, tile_lcnt[b]);
    if (lane == 0) {
      tile_excl[b] = e;
      tile_lexcl[b] = el;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRows; ++r) {
    const int idx = wbase + r * 32 + lane;
    if (idx < n) {
      const unsigned int dig = ldig[r];
      const unsigned int j = bin_base[dig] + tile_excl[dig] + wcnt[warp][dig] + lrank[r];
      p.perm[j] = idx;
      p.apos[j] = (int)(live_base[dig] + tile_lexcl[dig] + wlive[warp][dig] + larank[r]);
    }
  }
}

// The same result for scenes with at most kFewMaterials materials (every scene the reference ships), without the
// chain of 16 match_any rounds and without 256-bin bookkeeping: a thread takes 16 CONSECUTIVE slots (one 16-byte
// load of keys, one of survival flags), ranks them against its own earlier slots with 4-bit counters packed in a
// register (two halves of 8 slots, 8 bins x 4 bits each), the per-thread totals are scanned across the CTA as
// 16-bit pairs (8 words: 4 for all paths, 4 for the survivors), one warp per bin resolves the two tile prefixes
// in ONE look-back loop, and every slot's position is base[bin] + its rank.  Thread order is slot order, so the
// permutation is the stable one.  ~1.1 warp instructions per key instead of 4.2, 40 registers instead of 118.
constexpr int kFewMaterials = 8;

// Two decoupled look-backs (all paths / survivors of one bin) advanced together by a full warp.
__device__ __forceinline__ void lookback_warp2(unsigned long long* st_a, unsigned long long* st_b, size_t stride, unsigned int tile,
                                               unsigned int epoch, unsigned int total_a, unsigned int total_b,
                                               unsigned int* excl_a, unsigned int* excl_b) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) {
    st_volatile_u64(st_a + (size_t)tile * stride, lb_pack(epoch, tile == 0 ? 2u : 1u, total_a));
    st_volatile_u64(st_b + (size_t)tile * stride, lb_pack(epoch, tile == 0 ? 2u : 1u, total_b));
  }
  unsigned int ea = 0, eb = 0;
  bool done_a = false, done_b = false;  // warp-uniform
  for (int base = (int)tile - 1; base >= 0 && !(done_a && done_b); base -= 32) {
    const int t = base - lane;
    unsigned int fa = 2u, va = 0u, fb = 2u, vb = 0u;  // lanes before tile 0 act as an empty inclusive prefix
    if (t >= 0) {
      unsigned long long wa, wb;
      if (!done_a) {
        do {
          wa = ld_volatile_u64(st_a + (size_t)t * stride);
        } while (lb_epoch(wa) != epoch || lb_flag(wa) == 0u);
        fa = lb_flag(wa);
        va = lb_value(wa);
      }
      if (!done_b) {
        do {
          wb = ld_volatile_u64(st_b + (size_t)t * stride);
        } while (lb_epoch(wb) != epoch || lb_flag(wb) == 0u);
        fb = lb_flag(wb);
        vb = lb_value(wb);
      }
    }
    const unsigned int ia = __ballot_sync(0xffffffffu, fa == 2u), ib = __ballot_sync(0xffffffffu, fb == 2u);
    const int first_a = ia ? __ffs(ia) - 1 : 32, first_b = ib ? __ffs(ib) - 1 : 32;
    unsigned int xa = (!done_a && lane <= first_a) ? va : 0u, xb = (!done_b && lane <= first_b) ? vb : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      xa += __shfl_xor_sync(0xffffffffu, xa, o);
      xb += __shfl_xor_sync(0xffffffffu, xb, o);
    }
    ea += xa;
    eb += xb;
    done_a = done_a || ia != 0u;
    done_b = done_b || ib != 0u;
  }
  if (tile > 0 && lane == 0) {
    st_volatile_u64(st_a + (size_t)tile * stride, lb_pack(epoch, 2u, ea + total_a));
    st_volatile_u64(st_b + (size_t)tile * stride, lb_pack(epoch, 2u, eb + total_b));
  }
  *excl_a = ea;
  *excl_b = eb;
}

__global__ void __launch_bounds__(kSortThreads, 4) k_sort_material_few(MatSortParams p, int n_mat) {
  __shared__ unsigned int s_warp[kSortWarps][8];          // warp totals, 16-bit pairs: [0..3] all paths, [4..7] survivors
  __shared__ unsigned int s_wexcl[kSortWarps][8];         // ... exclusive over the warps of the CTA
  __shared__ unsigned int s_tot[2][kFewMaterials];        // tile totals per bin: all / survivors
  __shared__ unsigned int s_base[2][kFewMaterials];       // bin base + tile prefix: all / survivors
  __shared__ unsigned int s_ex[kFewMaterials][kSortThreads];  // per thread and bin: slots of the tile in front (all | survivors << 16)
  __shared__ __align__(16) unsigned int s_slot[kSortTile];  // per slot of the tile: bin << 28 | survivors in front << 14 | paths in front
  __shared__ unsigned int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.ctr->n_live[p.depth];
  // the grid is sized for a full frame: CTAs beyond the live paths leave before they queue up for a ticket
  // (exactly ceil(n / tile) CTAs stay, so the tickets they draw are still 0 .. tiles-1)
  if ((long long)blockIdx.x * kSortTile >= (long long)n) return;
  if (tid == 0) s_tile = atomicAdd(&p.ctr->sort_ticket[p.depth], 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const unsigned int epoch = p.ctr->serial * (unsigned int)(kMaxDepth + 1) + (unsigned int)p.depth + 1u;
  const int tile_base = (int)tile * kSortTile;
  const int base = tile_base + tid * 16;
  unsigned int kw[4] = {0u, 0u, 0u, 0u}, lw[4] = {0u, 0u, 0u, 0u};  // (both arrays are allocated in whole tiles)
  if (base < n) {
    const uint4 a = *reinterpret_cast<const uint4*>(p.key + base);
    const uint4 b = *reinterpret_cast<const uint4*>(p.live + base);
    kw[0] = a.x; kw[1] = a.y; kw[2] = a.z; kw[3] = a.w;
    lw[0] = b.x; lw[1] = b.y; lw[2] = b.z; lw[3] = b.w;
  }
  // the histograms of the depth (bin totals), wanted after the scan: issue the loads now
  unsigned int tot = 0, ltot = 0;
  if (lane < n_mat) {
    tot = p.ctr->hist[p.depth][n_mat - 1 - lane];
    ltot = p.ctr->hist_live[p.depth][n_mat - 1 - lane];
  }
  // ---- ranks inside the thread: 4-bit counters, bin c = n_mat - 1 - material (ascending bin = descending material) ----
  unsigned int cnt[2] = {0u, 0u}, lcnt[2] = {0u, 0u};  // per half: 8 bins x 4 bits (a half holds 8 slots: counts <= 8)
  unsigned int rk[2] = {0u, 0u}, lrk[2] = {0u, 0u};    // per half: 8 slots x 4 bits, rank within the THREAD (<= 15)
  unsigned int bin_of[2] = {0u, 0u};                   // per half: 8 slots x 4 bits
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int h = k >> 3, q = k & 7;
    const bool valid = base + k < n;
    const unsigned int mat = (kw[k >> 2] >> (8 * (k & 3))) & 0xffu;
    const bool alive = valid && ((lw[k >> 2] >> (8 * (k & 3))) & 0xffu) != 0u;
    const unsigned int c = ((unsigned int)(n_mat - 1) - mat) & 7u;  // (an invalid slot reads material 0: a bin that exists)
    const unsigned int sh = 4u * c;
    unsigned int r = (cnt[h] >> sh) & 15u, lr = (lcnt[h] >> sh) & 15u;
    if (h == 1) {
      r += (cnt[0] >> sh) & 15u;
      lr += (lcnt[0] >> sh) & 15u;
    }
    rk[h] |= r << (4 * q);
    lrk[h] |= lr << (4 * q);
    bin_of[h] |= c << (4 * q);
    cnt[h] += (valid ? 1u : 0u) << sh;
    lcnt[h] += (alive ? 1u : 0u) << sh;
  }
  // ---- thread totals as 16-bit pairs (bins 2j, 2j+1), inclusive scan over the warp ----
  unsigned int v[8];
  {
    const unsigned int lo = (cnt[0] & 0x0f0f0f0fu) + (cnt[1] & 0x0f0f0f0fu);                // bins 0 2 4 6, one byte each
    const unsigned int hi = ((cnt[0] >> 4) & 0x0f0f0f0fu) + ((cnt[1] >> 4) & 0x0f0f0f0fu);  // bins 1 3 5 7
    const unsigned int llo = (lcnt[0] & 0x0f0f0f0fu) + (lcnt[1] & 0x0f0f0f0fu);
    const unsigned int lhi = ((lcnt[0] >> 4) & 0x0f0f0f0fu) + ((lcnt[1] >> 4) & 0x0f0f0f0fu);
    v[0] = __byte_perm(lo, hi, 0x0400) & 0x00ff00ffu;
    v[1] = __byte_perm(lo, hi, 0x0501) & 0x00ff00ffu;
    v[2] = __byte_perm(lo, hi, 0x0602) & 0x00ff00ffu;
    v[3] = __byte_perm(lo, hi, 0x0703) & 0x00ff00ffu;
    v[4] = __byte_perm(llo, lhi, 0x0400) & 0x00ff00ffu;
    v[5] = __byte_perm(llo, lhi, 0x0501) & 0x00ff00ffu;
    v[6] = __byte_perm(llo, lhi, 0x0602) & 0x00ff00ffu;
    v[7] = __byte_perm(llo, lhi, 0x0703) & 0x00ff00ffu;
  }
  unsigned int own[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) own[i] = v[i];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned int y = __shfl_up_sync(0xffffffffu, v[i], o);
      if (lane >= o) v[i] += y;  // a warp holds 512 slots: the 16-bit fields cannot carry
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_warp[warp][i] = v[i];
  }
  __syncthreads();
  if (tid < kSortWarps * 8) {  // (warp w, word i): sum of the warps in front; the last warp's inclusive sum is the tile total
    const int w = tid >> 3, i = tid & 7;
    unsigned int run = 0;
    for (int ww = 0; ww < w; ++ww) run += s_warp[ww][i];  // (a tile holds 4096 slots: the 16-bit fields cannot carry)
    s_wexcl[w][i] = run;
    if (w == kSortWarps - 1) {
      const unsigned int last = s_warp[w][i];
      s_tot[i >> 2][2 * (i & 3)] = (run & 0xffffu) + (last & 0xffffu);
      s_tot[i >> 2][2 * (i & 3) + 1] = (run >> 16) + (last >> 16);
    }
  }
  __syncthreads();
  // ---- publish the tile's counts at once (the tiles behind wait for them), resolve the prefixes later ----
  unsigned int bin_total = 0, bin_ltotal = 0, pre = 0, lpre = 0, lall = 0;
  if (warp < n_mat) {
    const int b = warp;
    pre = lane < b ? tot : 0u;
    lpre = lane < b ? ltot : 0u;
    lall = ltot;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {  // n_mat <= 8: lanes 0..7 hold the bins
      pre += __shfl_xor_sync(0xffffffffu, pre, o);
      lpre += __shfl_xor_sync(0xffffffffu, lpre, o);
      lall += __shfl_xor_sync(0xffffffffu, lall, o);
    }
    bin_total = __shfl_sync(0xffffffffu, tot, b);
    bin_ltotal = __shfl_sync(0xffffffffu, ltot, b);
    if (lane == 0 && tile > 0) {  // (tile 0 publishes its inclusive prefix inside the look-back)
      if (bin_total != 0u) st_volatile_u64(p.status + (size_t)tile * 256 + b, lb_pack(epoch, 1u, s_tot[0][b]));
      if (bin_ltotal != 0u) st_volatile_u64(p.status_live + (size_t)tile * 256 + b, lb_pack(epoch, 1u, s_tot[1][b]));
    }
  }
  // ---- every slot's place inside the tile, written in SLOT order to shared memory ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned int ea = v[i] - own[i] + s_wexcl[warp][i];          // paths of the tile in front of this thread, bins 2i, 2i+1
    const unsigned int el = v[4 + i] - own[4 + i] + s_wexcl[warp][4 + i];  // survivors
    s_ex[2 * i][tid] = (ea & 0xffffu) | (el << 16);
    s_ex[2 * i + 1][tid] = (ea >> 16) | (el & 0xffff0000u);
  }
  // (a thread reads back only what it wrote itself: no barrier)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned int pk[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const unsigned int c = (bin_of[h] >> (4 * q)) & 15u;
      const unsigned int e = s_ex[c][tid];
      const unsigned int ra = (e & 0xffffu) + ((rk[h] >> (4 * q)) & 15u);   // < 4096
      const unsigned int rl = (e >> 16) + ((lrk[h] >> (4 * q)) & 15u);      // < 4096
      pk[q] = (c << 28) | (rl << 14) | ra;
    }
    uint4* dst = reinterpret_cast<uint4*>(s_slot + tid * 16 + 8 * h);
    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  // ---- per bin (one warp each): base of the bin from the histograms, tile prefixes by look-back ----
  if (warp < n_mat) {
    const int b = warp;
    unsigned int e = 0, el = 0;
    if (bin_total != 0u) {  // warp-uniform; bins nobody holds are never read
      if (bin_ltotal != 0u) lookback_warp2(p.status + b, p.status_live + b, 256, tile, epoch, s_tot[0][b], s_tot[1][b], &e, &el);
      else e = lookback_warp(p.status + b, 256, tile, epoch, s_tot[0][b]);
    }
    if (lane == 0) {
      s_base[0][b] = pre + e;
      s_base[1][b] = lpre + el;
      // the number of survivors is the live count of the next depth
      if (tile == 0 && b == 0) p.ctr->n_live[p.depth + 1] = (int)lall;
    }
  }
  __syncthreads();
  // ---- scatter, consecutive threads on consecutive slots: runs of one material become coalesced stores ----
#pragma unroll 4
  for (int r = 0; r < 16; ++r) {
    const int s = r * kSortThreads + tid;
    const int idx = tile_base + s;
    if (idx < n) {
      const unsigned int w = s_slot[s];
      const unsigned int c = w >> 28;
      const unsigned int j = s_base[0][c] + (w & 0x3fffu);
      p.perm[j] = idx;
      p.apos[j] = (int)(s_base[1][c] + ((w >> 14) & 0x3fffu));
    }
  }
}

// Compaction ranks WITHOUT a sort (SORT_BY_MATERIAL 0, apps/src/pathtrace.cu:38,611-613): paths are shaded in slot
// order, so all the shade kernel needs is apos[j] = survivors in front of slot j (the live prefix of
// thrust::stable_partition, :649) and the live count of the next depth.  One pass over the 1-byte survival flags:
// 16 flags per thread, block scan, decoupled look-back across the 4096-slot tiles.  With the pixel-keyed RNG the
// image does not depend on the order paths sit in the arrays, so this replaces the material sort altogether there.
__global__ void __launch_bounds__(kSortThreads) k_rank_live(MatSortParams p) {
  __shared__ unsigned int smem[kScanThreads / 32 + 1];
  __shared__ unsigned int s_tile, s_excl;
  static_assert(kScanThreads == kSortThreads && kSortTile == kSortThreads * 16, "16 flags per thread");
  const int tid = threadIdx.x;
  const int n = p.ctr->n_live[p.depth];
  if (tid == 0) s_tile = atomicAdd(&p.ctr->sort_ticket[p.depth], 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  if ((long long)tile * kSortTile >= (long long)n) return;
  const unsigned int epoch = p.ctr->serial * (unsigned int)(kMaxDepth + 1) + (unsigned int)p.depth + 1u;
  const int base = (int)tile * kSortTile + tid * 16;
  unsigned int f[4] = {0u, 0u, 0u, 0u};  // 16 flags, one byte each (the arrays are allocated in whole tiles' worth of bytes)
  if (base < n) {
    const uint4 v = *reinterpret_cast<const uint4*>(p.live + base);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  unsigned int alive = 0u;  // bit k: slot base + k survives
#pragma unroll
  for (int k = 0; k < 16; ++k)
    if (base + k < n && ((f[k >> 2] >> (8 * (k & 3))) & 0xffu) != 0u) alive |= 1u << k;
  unsigned int total;
  const unsigned int excl = block_exclusive_scan((unsigned int)__popc(alive), &total, smem);
  if (tid < 32) {
    const unsigned int e = lookback_warp(p.status_live, 256, tile, epoch, total);
    if (tid == 0) {
      s_excl = e;
      if (((long long)tile + 1) * kSortTile >= (long long)n) p.ctr->n_live[p.depth + 1] = (int)(e + total);
    }
  }
  __syncthreads();
  const unsigned int first = s_excl + excl;
  if (base < n) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = base + 4 * q;
      if (j >= n) break;
      int4 o;
      o.x = (int)(first + __popc(alive & ((1u << (4 * q)) - 1u)));
      o.y = (int)(first + __popc(alive & ((1u << (4 * q + 1)) - 1u)));
      o.z = (int)(first + __popc(alive & ((1u << (4 * q + 2)) - 1u)));
      o.w = (int)(first + __popc(alive & ((1u << (4 * q + 3)) - 1u)));
      *reinterpret_cast<int4*>(p.apos + j) = o;  // (apos holds whole tiles: the last int4 may pass n)
    }
  }
}

}  // namespace b2pt
