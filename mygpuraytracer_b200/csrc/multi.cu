// multi.cu -- pathtrace() over several GPUs: one FRAME = one iteration per GPU, combined on the device.
//
// The reference accumulates image[pixel] += color * PI once per iteration (apps/src/pathtrace.cu:508) and hands
// the running sum to the host after every call (:662-668).  Iterations are independent, so G GPUs render G
// consecutive iterations at once: member g of a frame that starts at iteration i renders i + g into its own
// zeroed image (its "contribution"), with `lanes` contexts per GPU rendering the next frames ahead exactly as
// a B2ptPipe does on one GPU (csrc/pipe.cu).  What is new here is how a frame is combined:
//
//   k_frame_reduce   ONE kernel per member that is reduce-scatter + accumulate + clear + forward at once.
//                    Member g owns pixel slice g of the running sum.  It reads slice g of EVERY member's
//                    contribution through NVLink peer memory, adds them to its slice IN ITERATION ORDER
//                    (s = ((s + c_0) + c_1) + ...), stores zero back over every contribution it consumed and
//                    writes the new sum both to its own slice and into member 0's whole image.
//
// A contribution is 0 or exactly color*PI (a pixel's one path dies once per iteration), so the additions are the
// reference's sequential `image[pixel] += color*PI` in the same order: the image is BIT-IDENTICAL to the one a
// single GPU accumulates, which an NCCL reduce (tree / ring order) cannot give.  Per frame a member pulls
// (G-1)/G of one image over NVLink, spread evenly over all links (all-to-all), and only member 0 touches PCIe.
//
// Two hosts are served by the same objects:
//   * one process per GPU (torch.distributed / MPI style): a B2ptShard per rank; device pointers are exchanged
//     once as CUDA IPC handles (b2pt_shard_export / b2pt_shard_connect), and the two rendezvous of a frame
//     ("every contribution is complete", "every slice has been consumed") are a barrier the HOST queues on
//     b2pt_shard_stream() between the three phases -- with NCCL that is a 4-byte all-reduce, so no kernel here
//     ever spins on a flag.  b2pt_shard_frame_image() + b2pt_shard_frame_merge() is the plain form for hosts
//     that would rather call ncclReduce on the contribution themselves.
//   * one process, several GPUs (the reference's own host, apps/src/main.cpp): b2pt_multi_* owns one shard per
//     device, enables peer access, and replaces the barriers by cross-device event waits.  Every member then
//     copies ITS slice to the host image over its own PCIe link.
//
// Built on the public C ABI of include/b2pt.h only.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/b2pt.h"

extern "C" void b2pt_set_last_error_(const char* msg);

namespace {

int mfail(int code, const std::string& msg) {
  b2pt_set_last_error_(msg.c_str());
  return code;
}
#define MCK(expr)                                                                                          \
  do {                                                                                                     \
    cudaError_t e_ = (expr);                                                                               \
    if (e_ != cudaSuccess) return mfail(B2PT_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_));  \
  } while (0)

constexpr int kMaxMembers = B2PT_MAX_MEMBERS;

struct ReduceArgs {
  float* sum;                // this member's slice of the running sum (len floats)
  float* forward;            // member 0's whole image + offset of the slice, or NULL
  float* contrib[kMaxMembers];  // every member's contribution + offset of the slice, in iteration order
  int members;
  unsigned int len;          // floats in the slice
};

__device__ __forceinline__ bool nz(const float4& a) { return a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f; }

// Grid-stride over the float4s of the slice; the loads of all members are issued before the first add so that
// the NVLink round trips overlap.  M = members when it is a small known number (full unroll), 0 = generic.
template <int M>
__global__ void __launch_bounds__(256) k_frame_reduce(ReduceArgs a) {
  const int members = M ? M : a.members;
  const unsigned int n4 = a.len / 4;
  float4* s4 = reinterpret_cast<float4*>(a.sum);
  float4* f4 = reinterpret_cast<float4*>(a.forward);
  const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 c[M ? M : kMaxMembers];
#pragma unroll
    for (int m = 0; m < (M ? M : kMaxMembers); ++m)
      if (m < members) c[m] = __ldcg(reinterpret_cast<const float4*>(a.contrib[m]) + i);
    float4 s = s4[i];
#pragma unroll
    for (int m = 0; m < (M ? M : kMaxMembers); ++m)
      if (m < members && nz(c[m])) {  // x + 0 == x for the non-negative sums: skipping keeps the bits and the stores
        s.x += c[m].x;
        s.y += c[m].y;
        s.z += c[m].z;
        s.w += c[m].w;
        __stcg(reinterpret_cast<float4*>(a.contrib[m]) + i, zero);
      }
    s4[i] = s;
    if (f4) f4[i] = s;
  }
  // the last slice of an image whose float count is not a multiple of four
  if (blockIdx.x == 0 && threadIdx.x < (a.len & 3u)) {
    const unsigned int i = (a.len & ~3u) + threadIdx.x;
    float s = a.sum[i];
    for (int m = 0; m < members; ++m) {
      const float c = a.contrib[m][i];
      if (c != 0.0f) {
        s += c;
        a.contrib[m][i] = 0.0f;
      }
    }
    a.sum[i] = s;
    if (a.forward) a.forward[i] = s;
  }
}

void launch_reduce(const ReduceArgs& a, int sm_count, cudaStream_t st) {
  const unsigned int n4 = a.len / 4;
  const int grid = (int)std::max(1u, std::min((n4 + 255u) / 256u, (unsigned int)sm_count * 8u));
  switch (a.members) {
    case 1: k_frame_reduce<1><<<grid, 256, 0, st>>>(a); break;
    case 2: k_frame_reduce<2><<<grid, 256, 0, st>>>(a); break;
    case 4: k_frame_reduce<4><<<grid, 256, 0, st>>>(a); break;
    case 8: k_frame_reduce<8><<<grid, 256, 0, st>>>(a); break;
    default: k_frame_reduce<0><<<grid, 256, 0, st>>>(a); break;
  }
}

struct SLane {
  B2ptCtx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  float* image = nullptr;          // the contribution (cudaMalloc: exportable)
  cudaEvent_t rendered = nullptr;  // the lane's render of `iter` is complete
  cudaEvent_t released = nullptr;  // every member has consumed (and cleared) the contribution
  int iter = 0;
  bool busy = false;
  bool released_valid = false;
  // peers' pointers for this lane index, by member (own entry = local pointer)
  float* peer_image[kMaxMembers] = {};
  float* peer_albedo[kMaxMembers] = {};
};

struct Slice {
  size_t off = 0, len = 0;  // floats
};

}  // namespace

struct B2ptShard {
  int rank = 0, world = 1, device = 0, sm_count = 148;
  std::vector<SLane> lanes;
  int head = 0;
  int next_first = 0;  // first iteration of the frame the head lane holds
  bool have_seq = false;
  size_t floats = 0;
  Slice slice;          // what this member reduces
  float* slice_sum = nullptr;  // [slice.len]
  float* full_sum = nullptr;   // member 0: whole image (others: peer pointer after connect, or NULL)
  bool owns_full = false;
  cudaStream_t merge = nullptr;
  cudaEvent_t ready = nullptr;    // in-process barrier 1: this member's contribution of the frame is complete
  cudaEvent_t reduced = nullptr;  // in-process barrier 2: this member's slice has been reduced
  bool connected = false;
  std::vector<void*> opened;      // cudaIpcOpenMemHandle results
  int cur = -1;                   // lane of the frame in flight (between begin and end)
  int cur_first = 0;
  int phase = 0;                  // 0 idle, 1 after begin, 2 after reduce / merge
  int albedo_member = 0, albedo_lane = 0;
  bool albedo_valid = false;
  bool albedo_skip_unchanged = false;
  uint64_t albedo_version = 1, albedo_host_version = 0;
  const float* albedo_host_last = nullptr;
  int64_t reduces = 0, misses = 0;
};

namespace {

Slice slice_of(size_t floats, int world, int r) {
  Slice s;
  const size_t per = ((floats + (size_t)world * 4 - 1) / ((size_t)world * 4)) * 4;  // multiple of 4 floats
  s.off = std::min(per * (size_t)r, floats);
  s.len = std::min(per, floats - s.off);
  return s;
}

int shard_enqueue(B2ptShard* s, int lane_index, int iter, bool after_release) {
  SLane& L = s->lanes[(size_t)lane_index];
  if (after_release && L.released_valid) MCK(cudaStreamWaitEvent(L.stream, L.released, 0));
  int rc = b2pt_render(L.ctx, iter, 1, 1);
  if (rc) return rc;
  MCK(cudaEventRecord(L.rendered, L.stream));
  L.iter = iter;
  L.busy = true;
  return 0;
}

// Drop whatever was speculated; lanes k = 0.. render first + rank + k*world.
int shard_restart(B2ptShard* s, int first) {
  const size_t bytes = s->floats * sizeof(float);
  for (SLane& L : s->lanes) {
    MCK(cudaStreamSynchronize(L.stream));
    if (L.busy) MCK(cudaMemsetAsync(L.image, 0, bytes, L.stream));
    L.busy = false;
  }
  const int n = (int)s->lanes.size();
  for (int k = 0; k < n; ++k) {
    const long long it = (long long)first + s->rank + (long long)k * s->world;
    if (it > 0x7fffffffLL) break;
    int rc = shard_enqueue(s, k, (int)it, false);
    if (rc) return rc;
  }
  s->head = 0;
  s->next_first = first;
  s->have_seq = true;
  return 0;
}

}  // namespace

extern "C" void b2pt_shard_destroy(B2ptShard* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (SLane& L : s->lanes) {
    if (L.stream) cudaStreamSynchronize(L.stream);
  }
  if (s->merge) cudaStreamSynchronize(s->merge);
  for (SLane& L : s->lanes) {
    if (L.rendered) cudaEventDestroy(L.rendered);
    if (L.released) cudaEventDestroy(L.released);
    b2pt_destroy(L.ctx);
    if (L.image) cudaFree(L.image);
  }
  for (void* p : s->opened) cudaIpcCloseMemHandle(p);
  if (s->ready) cudaEventDestroy(s->ready);
  if (s->reduced) cudaEventDestroy(s->reduced);
  if (s->merge) cudaStreamDestroy(s->merge);
  if (s->slice_sum && s->slice_sum != s->full_sum) cudaFree(s->slice_sum);
  if (s->owns_full && s->full_sum) cudaFree(s->full_sum);
  delete s;
}

extern "C" int b2pt_shard_create(const B2ptScene* scene, const B2ptOptions* opt, int32_t rank, int32_t world, int32_t lanes,
                                 B2ptShard** out) {
  if (!scene || !out) return mfail(B2PT_ERR_INVALID, "scene and out must not be NULL");
  *out = nullptr;
  if (world < 1 || world > kMaxMembers) return mfail(B2PT_ERR_RANGE, "1 <= world <= B2PT_MAX_MEMBERS required");
  if (rank < 0 || rank >= world) return mfail(B2PT_ERR_RANGE, "0 <= rank < world required");
  if (lanes < 1 || lanes > 16) return mfail(B2PT_ERR_RANGE, "1 <= lanes <= 16 required");
  B2ptOptions o;
  b2pt_default_options(&o);
  if (opt) {
    if (opt->struct_size != sizeof(B2ptOptions)) return mfail(B2PT_ERR_INVALID, "B2ptOptions.struct_size mismatch");
    o = *opt;
  }
  if (o.record_stages) return mfail(B2PT_ERR_INVALID, "stage recording needs a plain context (b2pt_create)");
  o.concurrent_contexts = lanes;
  B2ptShard* s = new (std::nothrow) B2ptShard();
  if (!s) return mfail(B2PT_ERR_NOMEM, "out of host memory");
  s->rank = rank;
  s->world = world;
  s->device = o.device;
  s->albedo_skip_unchanged = o.persistent_host_albedo != 0;
  s->floats = (size_t)scene->camera.resolution[0] * (size_t)scene->camera.resolution[1] * 3;
  s->slice = slice_of(s->floats, world, rank);
  s->lanes.resize((size_t)lanes);
  const size_t bytes = s->floats * sizeof(float);
  int rc = 0;
  cudaError_t e = cudaSetDevice(s->device);
  if (e == cudaSuccess) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, s->device) == cudaSuccess) s->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&s->merge, cudaStreamNonBlocking);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->reduced, cudaEventDisableTiming);
  if (e == cudaSuccess && rank == 0) {
    e = cudaMalloc(&s->full_sum, bytes);
    s->owns_full = e == cudaSuccess;
    if (e == cudaSuccess) e = cudaMemsetAsync(s->full_sum, 0, bytes, s->merge);
  }
  if (world == 1) {
    s->slice_sum = s->full_sum;  // the one slice IS the image
  } else {
    if (e == cudaSuccess) e = cudaMalloc(&s->slice_sum, std::max<size_t>(s->slice.len, 4) * sizeof(float));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->slice_sum, 0, std::max<size_t>(s->slice.len, 4) * sizeof(float), s->merge);
  }
  if (e != cudaSuccess) rc = mfail(B2PT_ERR_CUDA, std::string("b2pt_shard_create: ") + cudaGetErrorString(e));
  for (int k = 0; k < lanes && rc == 0; ++k) {
    SLane& L = s->lanes[(size_t)k];
    rc = k > 0 ? b2pt_create_shared(s->lanes[0].ctx, scene, &o, &L.ctx) : b2pt_create(scene, &o, &L.ctx);
    if (rc) break;
    L.stream = (cudaStream_t)b2pt_stream(L.ctx);
    if (cudaMalloc(&L.image, bytes) != cudaSuccess || cudaMemsetAsync(L.image, 0, bytes, L.stream) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.rendered, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.released, cudaEventDisableTiming) != cudaSuccess) {
      rc = mfail(B2PT_ERR_CUDA, "b2pt_shard_create: lane allocation failed");
      break;
    }
    rc = b2pt_set_device_image(L.ctx, L.image);  // the lane gathers straight into the buffer that is reduced
    L.peer_image[rank] = L.image;
    L.peer_albedo[rank] = b2pt_device_albedo(L.ctx);
  }
  if (rc == 0 && cudaStreamSynchronize(s->merge) != cudaSuccess) rc = mfail(B2PT_ERR_CUDA, "b2pt_shard_create: sync failed");
  if (rc) {
    std::string keep = b2pt_last_error();
    b2pt_shard_destroy(s);
    b2pt_set_last_error_(keep.c_str());
    return rc;
  }
  s->connected = world == 1;
  *out = s;
  return 0;
}

// ---- pointer exchange ------------------------------------------------------------------------------------------
// Exported blob of one member: int32 lanes, int32 has_full, then cudaIpcMemHandle_t (64 B) x (2*lanes + 1):
// lane images, lane albedos, member 0's whole image.
static size_t export_bytes(int lanes) { return 8 + sizeof(cudaIpcMemHandle_t) * (size_t)(2 * lanes + 1); }

extern "C" int64_t b2pt_shard_export_size(B2ptShard* s) { return s ? (int64_t)export_bytes((int)s->lanes.size()) : 0; }

extern "C" int b2pt_shard_export(B2ptShard* s, void* blob, int64_t bytes) {
  if (!s || !blob) return mfail(B2PT_ERR_INVALID, "shard and blob must not be NULL");
  const int lanes = (int)s->lanes.size();
  if (bytes < (int64_t)export_bytes(lanes)) return mfail(B2PT_ERR_RANGE, "blob too small (b2pt_shard_export_size)");
  MCK(cudaSetDevice(s->device));
  char* p = (char*)blob;
  memset(p, 0, export_bytes(lanes));
  int32_t head[2] = {lanes, s->owns_full ? 1 : 0};
  memcpy(p, head, 8);
  cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(p + 8);
  for (int k = 0; k < lanes; ++k) {
    MCK(cudaIpcGetMemHandle(&h[k], s->lanes[(size_t)k].image));
    MCK(cudaIpcGetMemHandle(&h[lanes + k], b2pt_device_albedo(s->lanes[(size_t)k].ctx)));
  }
  if (s->owns_full) MCK(cudaIpcGetMemHandle(&h[2 * lanes], s->full_sum));
  return 0;
}

// The albedo AOV is a sub-range of one of the context's allocations only if the context allocated it on its
// own: b2pt.cu does (one cudaMalloc per buffer), so the handle's base address is the buffer itself.
extern "C" int b2pt_shard_connect(B2ptShard* s, const void* blobs, int64_t bytes_per_member) {
  if (!s || !blobs) return mfail(B2PT_ERR_INVALID, "shard and blobs must not be NULL");
  if (s->connected) return mfail(B2PT_ERR_STATE, "the shard is already connected");
  const int lanes = (int)s->lanes.size();
  if (bytes_per_member < (int64_t)export_bytes(lanes)) return mfail(B2PT_ERR_RANGE, "blobs too small");
  MCK(cudaSetDevice(s->device));
  for (int m = 0; m < s->world; ++m) {
    if (m == s->rank) continue;
    const char* p = (const char*)blobs + (size_t)m * (size_t)bytes_per_member;
    int32_t head[2];
    memcpy(head, p, 8);
    if (head[0] != lanes) return mfail(B2PT_ERR_INVALID, "every member of a job needs the same number of lanes");
    const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(p + 8);
    for (int k = 0; k < lanes; ++k) {
      void* q = nullptr;
      MCK(cudaIpcOpenMemHandle(&q, h[k], cudaIpcMemLazyEnablePeerAccess));
      s->opened.push_back(q);
      s->lanes[(size_t)k].peer_image[m] = (float*)q;
      MCK(cudaIpcOpenMemHandle(&q, h[lanes + k], cudaIpcMemLazyEnablePeerAccess));
      s->opened.push_back(q);
      s->lanes[(size_t)k].peer_albedo[m] = (float*)q;
    }
    if (m == 0) {
      if (!head[1]) return mfail(B2PT_ERR_INVALID, "member 0 exported no image");
      void* q = nullptr;
      MCK(cudaIpcOpenMemHandle(&q, h[2 * lanes], cudaIpcMemLazyEnablePeerAccess));
      s->opened.push_back(q);
      s->full_sum = (float*)q;
    }
  }
  s->connected = true;
  return 0;
}

// Same process: take the peers' pointers directly (b2pt_multi_*; peer access is the caller's business).
static int shard_connect_local(B2ptShard* s, B2ptShard* const* all) {
  for (int m = 0; m < s->world; ++m) {
    if (m == s->rank) continue;
    if (all[m]->lanes.size() != s->lanes.size()) return mfail(B2PT_ERR_INVALID, "lane counts differ");
    for (size_t k = 0; k < s->lanes.size(); ++k) {
      s->lanes[k].peer_image[m] = all[m]->lanes[k].image;
      s->lanes[k].peer_albedo[m] = b2pt_device_albedo(all[m]->lanes[k].ctx);
    }
  }
  if (s->rank != 0) s->full_sum = all[0]->full_sum;
  s->connected = true;
  return 0;
}

// ---- one frame, three phases -----------------------------------------------------------------------------------
extern "C" int b2pt_shard_frame_begin(B2ptShard* s, int32_t first_iter) {
  if (!s) return mfail(B2PT_ERR_INVALID, "shard is NULL");
  if (s->phase != 0) return mfail(B2PT_ERR_STATE, "b2pt_shard_frame_begin: the previous frame has not ended");
  if (first_iter < 1) return mfail(B2PT_ERR_INVALID, "iterations are numbered from 1");
  MCK(cudaSetDevice(s->device));
  const int mine = first_iter + s->rank;
  if (!(s->have_seq && s->next_first == first_iter && s->lanes[(size_t)s->head].busy && s->lanes[(size_t)s->head].iter == mine)) {
    if (s->have_seq) s->misses += 1;
    int rc = shard_restart(s, first_iter);
    if (rc) return rc;
  }
  s->cur = s->head;
  s->cur_first = first_iter;
  SLane& L = s->lanes[(size_t)s->cur];
  // bound the run-ahead of a rank that never waits for pixels: the frame that used this lane before must be done
  if (L.released_valid) MCK(cudaEventSynchronize(L.released));
  MCK(cudaStreamWaitEvent(s->merge, L.rendered, 0));
  MCK(cudaEventRecord(s->ready, s->merge));
  if (first_iter <= 1 && 1 - first_iter < s->world) {  // iteration 1 is in this frame: its renderer holds the albedo AOV
    s->albedo_member = 1 - first_iter;
    s->albedo_lane = s->cur;
    s->albedo_valid = true;
    s->albedo_version += 1;
  }
  s->phase = 1;
  return 0;
}

extern "C" void* b2pt_shard_stream(B2ptShard* s) { return s ? (void*)s->merge : nullptr; }

extern "C" int b2pt_shard_frame_reduce(B2ptShard* s) {
  if (!s) return mfail(B2PT_ERR_INVALID, "shard is NULL");
  if (s->phase != 1) return mfail(B2PT_ERR_STATE, "b2pt_shard_frame_reduce: call b2pt_shard_frame_begin first");
  if (!s->connected) return mfail(B2PT_ERR_STATE, "the shard is not connected to its peers (b2pt_shard_connect)");
  MCK(cudaSetDevice(s->device));
  SLane& L = s->lanes[(size_t)s->cur];
  if (s->slice.len > 0) {
    ReduceArgs a;
    memset(&a, 0, sizeof a);
    a.sum = s->slice_sum;
    a.forward = (s->full_sum && s->full_sum != s->slice_sum) ? s->full_sum + s->slice.off : nullptr;
    a.members = s->world;
    a.len = (unsigned int)s->slice.len;
    for (int m = 0; m < s->world; ++m) a.contrib[m] = L.peer_image[m] + s->slice.off;
    launch_reduce(a, s->sm_count, s->merge);
    MCK(cudaGetLastError());
    s->reduces += 1;
  }
  MCK(cudaEventRecord(s->reduced, s->merge));
  s->phase = 2;
  return 0;
}

// The plain form: the host reduces the contribution itself (ncclReduce to member 0 on b2pt_shard_stream()) ...
extern "C" float* b2pt_shard_frame_image(B2ptShard* s) {
  if (!s || s->phase != 1) return nullptr;
  return s->lanes[(size_t)s->cur].image;
}
// ... and then member 0 folds the reduced contribution into the running sum; every member clears its own.
extern "C" int b2pt_shard_frame_merge(B2ptShard* s) {
  if (!s) return mfail(B2PT_ERR_INVALID, "shard is NULL");
  if (s->phase != 1) return mfail(B2PT_ERR_STATE, "b2pt_shard_frame_merge: call b2pt_shard_frame_begin first");
  MCK(cudaSetDevice(s->device));
  SLane& L = s->lanes[(size_t)s->cur];
  if (s->rank == 0) {
    ReduceArgs a;
    memset(&a, 0, sizeof a);
    a.sum = s->full_sum;
    a.forward = nullptr;
    a.members = 1;
    a.len = (unsigned int)s->floats;
    a.contrib[0] = L.image;
    launch_reduce(a, s->sm_count, s->merge);
    MCK(cudaGetLastError());
    s->reduces += 1;
  } else {
    MCK(cudaMemsetAsync(L.image, 0, s->floats * sizeof(float), s->merge));
  }
  MCK(cudaEventRecord(s->reduced, s->merge));
  s->phase = 3;
  return 0;
}

// slice_to_host: this member copies ITS slice into image_host (same process, every member has the pointer);
// otherwise member 0 copies the whole image and the others copy nothing.
static int shard_frame_end(B2ptShard* s, float* image_host, float* albedo_host, bool slice_to_host, bool wait) {
  if (!s) return mfail(B2PT_ERR_INVALID, "shard is NULL");
  if (s->phase < 2) return mfail(B2PT_ERR_STATE, "b2pt_shard_frame_end: reduce or merge the frame first");
  MCK(cudaSetDevice(s->device));
  SLane& L = s->lanes[(size_t)s->cur];
  // everything queued on the merge stream so far (the host's second barrier included) precedes the release
  MCK(cudaEventRecord(L.released, s->merge));
  L.released_valid = true;
  if (image_host) {
    if (s->phase == 2 && slice_to_host) {
      if (s->slice.len)
        MCK(cudaMemcpyAsync(image_host + s->slice.off, s->slice_sum, s->slice.len * sizeof(float), cudaMemcpyDeviceToHost, s->merge));
    } else if (s->rank == 0) {
      MCK(cudaMemcpyAsync(image_host, s->full_sum, s->floats * sizeof(float), cudaMemcpyDeviceToHost, s->merge));
    }
  }
  if (albedo_host && s->rank == 0 && s->albedo_valid &&
      !(s->albedo_skip_unchanged && albedo_host == s->albedo_host_last && s->albedo_host_version == s->albedo_version)) {
    // the renderer of iteration 1 holds the AOV; its render preceded the first barrier of that frame
    const float* src = s->lanes[(size_t)s->albedo_lane].peer_albedo[s->albedo_member];
    if (src) {
      MCK(cudaMemcpyAsync(albedo_host, src, s->floats * sizeof(float), cudaMemcpyDefault, s->merge));
      s->albedo_host_last = albedo_host;
      s->albedo_host_version = s->albedo_version;
    }
  }
  // the lane moves on to the frame `lanes` frames ahead as soon as every member has consumed its contribution
  L.busy = false;
  const int n = (int)s->lanes.size();
  const long long next = (long long)s->cur_first + s->rank + (long long)n * s->world;
  if (next <= 0x7fffffffLL) {
    int rc = shard_enqueue(s, s->cur, (int)next, true);
    if (rc) return rc;
  }
  s->head = (s->head + 1) % n;
  s->next_first = s->cur_first + s->world;
  s->phase = 0;
  s->cur = -1;
  if (wait) MCK(cudaStreamSynchronize(s->merge));
  return 0;
}

extern "C" int b2pt_shard_frame_end(B2ptShard* s, float* image_host, float* albedo_host) {
  // a member that takes pixels waits for them; the others return at once and are throttled by frame_begin
  return shard_frame_end(s, image_host, albedo_host, false, image_host != nullptr || albedo_host != nullptr);
}

extern "C" int b2pt_shard_reset(B2ptShard* s, const B2ptCamera* cam) {
  if (!s) return mfail(B2PT_ERR_INVALID, "shard is NULL");
  if (s->phase != 0) return mfail(B2PT_ERR_STATE, "b2pt_shard_reset: a frame is in flight");
  MCK(cudaSetDevice(s->device));
  for (SLane& L : s->lanes) {
    int rc = cam ? b2pt_set_camera(L.ctx, cam) : b2pt_reset_accum(L.ctx);  // both zero the lane's image and albedo
    if (rc) return rc;
    rc = b2pt_sync(L.ctx);
    if (rc) return rc;
    L.busy = false;
  }
  if (s->slice_sum != s->full_sum) MCK(cudaMemsetAsync(s->slice_sum, 0, std::max<size_t>(s->slice.len, 4) * sizeof(float), s->merge));
  if (s->owns_full) MCK(cudaMemsetAsync(s->full_sum, 0, s->floats * sizeof(float), s->merge));
  MCK(cudaStreamSynchronize(s->merge));
  s->albedo_version += 1;
  s->albedo_valid = false;
  s->have_seq = false;
  s->head = 0;
  return 0;
}

extern "C" float* b2pt_shard_device_image(B2ptShard* s) { return (s && s->rank == 0) ? s->full_sum : nullptr; }
extern "C" float* b2pt_shard_device_slice(B2ptShard* s, int64_t* offset_floats, int64_t* len_floats) {
  if (!s) return nullptr;
  if (offset_floats) *offset_floats = (int64_t)s->slice.off;
  if (len_floats) *len_floats = (int64_t)s->slice.len;
  return s->slice_sum;
}
extern "C" int32_t b2pt_shard_lanes(B2ptShard* s) { return s ? (int32_t)s->lanes.size() : 0; }
extern "C" B2ptCtx* b2pt_shard_lane(B2ptShard* s, int32_t k) {
  return (s && k >= 0 && k < (int32_t)s->lanes.size()) ? s->lanes[(size_t)k].ctx : nullptr;
}
extern "C" int64_t b2pt_shard_launch_count(B2ptShard* s) {
  if (!s) return 0;
  int64_t n = s->reduces;
  for (SLane& L : s->lanes) n += b2pt_launch_count(L.ctx);
  return n;
}
extern "C" int64_t b2pt_shard_misses(B2ptShard* s) { return s ? s->misses : 0; }
extern "C" int b2pt_shard_sync(B2ptShard* s) {
  if (!s) return mfail(B2PT_ERR_INVALID, "shard is NULL");
  MCK(cudaSetDevice(s->device));
  MCK(cudaStreamSynchronize(s->merge));
  return 0;
}

// =====================================================================================================================
// one process, several GPUs
// =====================================================================================================================
struct B2ptMulti {
  std::vector<B2ptShard*> members;
  std::vector<int> devices;
  bool host_registered = false;
  float* registered_image = nullptr;
};

extern "C" void b2pt_multi_destroy(B2ptMulti* m) {
  if (!m) return;
  for (B2ptShard* s : m->members) b2pt_shard_destroy(s);
  delete m;
}

extern "C" int b2pt_multi_create(const B2ptScene* scene, const B2ptOptions* opt, int32_t n_devices, const int32_t* devices,
                                 int32_t lanes, B2ptMulti** out) {
  if (!scene || !out) return mfail(B2PT_ERR_INVALID, "scene and out must not be NULL");
  *out = nullptr;
  if (n_devices < 1 || n_devices > kMaxMembers) return mfail(B2PT_ERR_RANGE, "1 <= n_devices <= B2PT_MAX_MEMBERS required");
  int visible = 0;
  MCK(cudaGetDeviceCount(&visible));
  B2ptOptions o;
  b2pt_default_options(&o);
  if (opt) {
    if (opt->struct_size != sizeof(B2ptOptions)) return mfail(B2PT_ERR_INVALID, "B2ptOptions.struct_size mismatch");
    o = *opt;
  }
  B2ptMulti* m = new (std::nothrow) B2ptMulti();
  if (!m) return mfail(B2PT_ERR_NOMEM, "out of host memory");
  int rc = 0;
  for (int g = 0; g < n_devices && rc == 0; ++g) {
    const int dev = devices ? devices[g] : g;  // the same device may appear twice (members then share it)
    if (dev < 0 || dev >= visible) {
      rc = mfail(B2PT_ERR_INVALID, "b2pt_multi_create: no such CUDA device");
      break;
    }
    m->devices.push_back(dev);
  }
  // peer access in both directions between every pair of distinct devices
  for (size_t a = 0; a < m->devices.size() && rc == 0; ++a)
    for (size_t b = 0; b < m->devices.size() && rc == 0; ++b) {
      if (m->devices[a] == m->devices[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->devices[a], m->devices[b]);
      if (!can) {
        rc = mfail(B2PT_ERR_STATE, "b2pt_multi_create: the devices cannot access each other's memory (no NVLink / P2P)");
        break;
      }
      cudaSetDevice(m->devices[a]);
      const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        rc = mfail(B2PT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
      cudaGetLastError();
    }
  for (int g = 0; g < n_devices && rc == 0; ++g) {
    B2ptOptions og = o;
    og.device = m->devices[(size_t)g];
    B2ptShard* s = nullptr;
    rc = b2pt_shard_create(scene, &og, g, n_devices, lanes, &s);
    if (rc == 0) m->members.push_back(s);
  }
  for (size_t g = 0; g < m->members.size() && rc == 0; ++g)
    if (n_devices > 1) rc = shard_connect_local(m->members[g], m->members.data());
  if (rc) {
    std::string keep = b2pt_last_error();
    b2pt_multi_destroy(m);
    b2pt_set_last_error_(keep.c_str());
    return rc;
  }
  *out = m;
  return 0;
}

// One frame: iterations first_iter .. first_iter + members - 1, one per member; on return image_host holds the
// running sum including all of them (bit-identical to a single GPU accumulating them one after the other).
extern "C" int b2pt_multi_pathtrace(B2ptMulti* m, int32_t first_iter, float* image_host, float* albedo_host) {
  if (!m) return mfail(B2PT_ERR_INVALID, "multi is NULL");
  const size_t G = m->members.size();
  int rc;
  for (B2ptShard* s : m->members)
    if ((rc = b2pt_shard_frame_begin(s, first_iter))) return rc;
  // rendezvous 1: every contribution of the frame is complete before anybody reads it
  for (size_t a = 0; a < G; ++a) {
    MCK(cudaSetDevice(m->members[a]->device));
    for (size_t b = 0; b < G; ++b)
      if (a != b) MCK(cudaStreamWaitEvent(m->members[a]->merge, m->members[b]->ready, 0));
  }
  for (B2ptShard* s : m->members)
    if ((rc = b2pt_shard_frame_reduce(s))) return rc;
  // rendezvous 2: every slice has been consumed before a contribution is rendered into again (and before member 0
  // hands out the whole image)
  for (size_t a = 0; a < G; ++a) {
    MCK(cudaSetDevice(m->members[a]->device));
    for (size_t b = 0; b < G; ++b)
      if (a != b) MCK(cudaStreamWaitEvent(m->members[a]->merge, m->members[b]->reduced, 0));
  }
  // every member copies its slice over its own PCIe link; member 0 also delivers the albedo AOV
  for (B2ptShard* s : m->members)
    if ((rc = shard_frame_end(s, image_host, s->rank == 0 ? albedo_host : nullptr, true, false))) return rc;
  if (image_host || albedo_host)
    for (B2ptShard* s : m->members)
      if ((rc = b2pt_shard_sync(s))) return rc;
  return 0;
}

extern "C" int b2pt_multi_reset(B2ptMulti* m, const B2ptCamera* cam) {
  if (!m) return mfail(B2PT_ERR_INVALID, "multi is NULL");
  for (B2ptShard* s : m->members) {
    int rc = b2pt_shard_reset(s, cam);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int32_t b2pt_multi_members(B2ptMulti* m) { return m ? (int32_t)m->members.size() : 0; }
extern "C" B2ptShard* b2pt_multi_member(B2ptMulti* m, int32_t g) {
  return (m && g >= 0 && g < (int32_t)m->members.size()) ? m->members[(size_t)g] : nullptr;
}
extern "C" float* b2pt_multi_device_image(B2ptMulti* m) { return (m && !m->members.empty()) ? m->members[0]->full_sum : nullptr; }
extern "C" int64_t b2pt_multi_launch_count(B2ptMulti* m) {
  int64_t n = 0;
  if (m)
    for (B2ptShard* s : m->members) n += b2pt_shard_launch_count(s);
  return n;
}
