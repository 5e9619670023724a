// k_prims.cuh -- standalone scan / compaction kernels.
//
// The same warp-ballot + decoupled look-back scheme the shade kernel uses for
// its stable compaction, exposed on plain int arrays.  They back the C-ABI
// entry points that serve the surface of the reference's side library
// (StreamCompaction::Efficient::scan / compact, apps/stream_compaction/
// efficient.h:8-11; kernMapToBoolean / kernScatter, common.cu:25-49; the
// Blelloch up/down sweeps of efficient.cu:14-31 take 2*log2(N) launches, this
// takes one) and make the look-back machinery testable in isolation.
#pragma once

#include "pt_device.cuh"

namespace b2pt {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // consecutive elements per thread
constexpr int kScanTile = kScanThreads * kScanItems;

// Block-wide exclusive scan of one value per thread; returns the exclusive
// prefix and the block total.
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* total, unsigned int* smem /*[8+1]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = lane < kScanThreads / 32 ? smem[lane] : 0u;
    unsigned int wi = w;
#pragma unroll
    for (int o = 1; o < kScanThreads / 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < kScanThreads / 32) smem[lane] = wi - w;
    if (lane == kScanThreads / 32 - 1) smem[kScanThreads / 32] = wi;
  }
  __syncthreads();
  *total = smem[kScanThreads / 32];
  return incl - v + smem[warp];
}

// MODE 0: out[i] = exclusive prefix sum of in.
// MODE 1: stable compaction of the non-zero elements of in; *count = kept.
// MODE 2: in is a byte predicate; perm gets kept indices at [rank] and the
//         dropped indices at dead[rank_dead]; *count = kept.
template <int MODE>
__global__ void __launch_bounds__(kScanThreads) k_scan_family(const int* __restrict__ in, const uint8_t* __restrict__ flags,
                                                              int n, int* out, int* dead, unsigned int* ticket,
                                                              unsigned long long* status, unsigned int epoch, int* count) {
  __shared__ unsigned int smem[kScanThreads / 32 + 1];
  __shared__ unsigned int s_tile, s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long base = (long long)tile * kScanTile + (long long)threadIdx.x * kScanItems;
  if ((long long)tile * kScanTile >= (long long)n) return;
  int v[kScanItems];
  unsigned int w[kScanItems];
  unsigned int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const long long i = base + k;
    v[k] = 0;
    if (i < n) v[k] = (MODE == 2) ? (int)flags[i] : in[i];
    w[k] = (MODE == 0) ? (unsigned int)v[k] : (v[k] != 0 ? 1u : 0u);
    sum += w[k];
  }
  unsigned int total;
  unsigned int excl = block_exclusive_scan(sum, &total, smem);
  if (threadIdx.x < 32) {
    const unsigned int e = lookback_warp(status, 1, tile, epoch, total);
    if (threadIdx.x == 0) {
      s_excl = e;
      if (MODE != 0 && ((long long)tile + 1) * kScanTile >= (long long)n) *count = (int)(e + total);
    }
  }
  __syncthreads();
  excl += s_excl;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const long long i = base + k;
    if (i < n) {
      if (MODE == 0) {
        out[i] = (int)excl;
      } else if (MODE == 1) {
        if (w[k]) out[excl] = v[k];
      } else {
        if (w[k]) out[excl] = (int)i; else dead[i - (long long)excl] = (int)i;
      }
    }
    excl += w[k];
  }
}

__global__ void k_key_hist_u8(const uint8_t* __restrict__ key, int n, unsigned int* hist) {
  __shared__ unsigned int sh[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sh[key[i]], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

}  // namespace b2pt
