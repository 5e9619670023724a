// pipe.cu -- pathtrace() for a host that asks for one iteration at a time, pipelined.
//
// The reference's host loop calls pathtrace(pbo, 0, ++iteration) once per frame (apps/src/main.cpp:255)
// and gets the running sum back in scene->state.image every time (apps/src/pathtrace.cu:663-668).  Served
// one call at a time that is 1.36 ms of kernels, which leave most of the GPU idle in their tails, followed
// by 0.47 ms of PCIe during which the GPU does nothing at all.  Iteration numbers are predictable, so a
// B2ptPipe keeps `lanes` contexts busy with the NEXT iterations while the host consumes the current one:
//
//   lane k renders iteration i + k*stride into its own zeroed image (its "contribution")
//   b2pt_pipe_pathtrace(i):  wait for the lane that holds i
//                            sum += contribution, contribution = 0        (k_pipe_merge, copy stream)
//                            lane renders i + lanes*stride                (as soon as the merge is done)
//                            D2H of sum -> host                           (overlaps the other lanes' kernels)
//
// Every iteration is still rendered exactly once and merged in call order.  A pixel receives at most one
// addition per iteration (its one path dies once, apps/src/pathtrace.cu:501-510), so a lane's contribution
// is 0 or 0 + color*PI = color*PI exactly, and sum + contribution has the bits of the reference's
// image[pixel] += color*PI: results are bit-identical to b2pt_pathtrace on one context
// (tests/test_gpu_pipe.py).  A call that does not continue the sequence (another stride, a restart) drops
// the speculated iterations and starts over from the requested one.
//
// Built on the public C ABI of include/b2pt.h only.
#include <cuda_runtime.h>

#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/b2pt.h"

extern "C" void b2pt_set_last_error_(const char* msg);
extern "C" int b2pt_tonemap_on_stream_(B2ptCtx* c, const float* src_dev, int32_t iter, uint8_t* rgba8_dev, void* stream);

namespace {

int pipe_fail(int code, const std::string& msg) {
  b2pt_set_last_error_(msg.c_str());
  return code;
}
#define PCK(expr)                                                                                         \
  do {                                                                                                    \
    cudaError_t e_ = (expr);                                                                              \
    if (e_ != cudaSuccess) return pipe_fail(B2PT_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_)); \
  } while (0)

// sum += c; c = 0.  n4 float4s followed by `tail` floats.
__global__ void __launch_bounds__(256) k_pipe_merge(float* __restrict__ sum, float* __restrict__ c, size_t n4, int tail) {
  float4* s4 = reinterpret_cast<float4*>(sum);
  float4* c4 = reinterpret_cast<float4*>(c);
  const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = c4[i];
    if (a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f) {
      float4 s = s4[i];
      s.x += a.x;
      s.y += a.y;
      s.z += a.z;
      s.w += a.w;
      s4[i] = s;
      c4[i] = zero;
    }
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
    const size_t i = 4 * n4 + threadIdx.x;
    sum[i] += c[i];
    c[i] = 0.0f;
  }
}

struct Lane {
  B2ptCtx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t rendered = nullptr;  // the lane's render of `iter` is complete
  cudaEvent_t merged = nullptr;    // its contribution has been folded into the sum
  int iter = 0;
  bool busy = false;
  bool rendered_valid = false;     // `rendered` has been recorded at least once
};

}  // namespace

struct B2ptPipe {
  int device = 0;
  std::vector<Lane> lanes;
  int head = 0;
  int stride = 1;
  int last_iter = 0;
  bool have_last = false;
  size_t floats = 0;  // W*H*3
  float* sum = nullptr;
  cudaStream_t copy = nullptr;
  int sm_count = 148;
  int albedo_lane = 0;
  bool albedo_skip_unchanged = false;  // B2ptOptions.persistent_host_albedo
  uint64_t albedo_version = 1, albedo_host_version = 0;
  const float* albedo_host_last = nullptr;
  int64_t merges = 0;
  int64_t misses = 0;
  int last_lane = 0;
  bool track_loop_ms = false;  // switched on by the first b2pt_pipe_last_loop_ms call
  float last_loop_ms = -1.0f;
};

static int pipe_enqueue(B2ptPipe* p, Lane& L, int lane_index, int iter, bool after_merge) {
  if (after_merge) PCK(cudaStreamWaitEvent(L.stream, L.merged, 0));
  int rc = b2pt_render(L.ctx, iter, 1, 1);
  if (rc) return rc;
  PCK(cudaEventRecord(L.rendered, L.stream));
  L.rendered_valid = true;
  L.iter = iter;
  L.busy = true;
  if (iter == 1) {  // iteration 1, and only it, writes the albedo AOV (apps/src/pathtrace.cu:412)
    p->albedo_lane = lane_index;
    p->albedo_version += 1;
  }
  return 0;
}

// Drop whatever was speculated and start the sequence iter, iter + stride, ...
static int pipe_restart(B2ptPipe* p, int iter) {
  const size_t bytes = p->floats * sizeof(float);
  for (Lane& L : p->lanes) {
    if (!L.busy) continue;
    PCK(cudaStreamSynchronize(L.stream));
    PCK(cudaMemsetAsync(b2pt_device_image(L.ctx), 0, bytes, L.stream));
    L.busy = false;
  }
  const int n = (int)p->lanes.size();
  for (int k = 0; k < n; ++k) {
    const long long it = (long long)iter + (long long)k * p->stride;
    if (it > 0x7fffffffLL) break;
    int rc = pipe_enqueue(p, p->lanes[k], k, (int)it, false);
    if (rc) return rc;
  }
  p->head = 0;
  return 0;
}

extern "C" void b2pt_pipe_destroy(B2ptPipe* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (Lane& L : p->lanes) {
    if (L.stream) cudaStreamSynchronize(L.stream);
    if (L.rendered) cudaEventDestroy(L.rendered);
    if (L.merged) cudaEventDestroy(L.merged);
    b2pt_destroy(L.ctx);
  }
  if (p->copy) {
    cudaStreamSynchronize(p->copy);
    cudaStreamDestroy(p->copy);
  }
  if (p->sum) cudaFree(p->sum);
  delete p;
}

extern "C" int b2pt_pipe_create(const B2ptScene* scene, const B2ptOptions* opt, int32_t lanes, B2ptPipe** out) {
  if (!scene || !out) return pipe_fail(B2PT_ERR_INVALID, "scene and out must not be NULL");
  *out = nullptr;
  if (lanes < 1 || lanes > 16) return pipe_fail(B2PT_ERR_RANGE, "1 <= lanes <= 16 required");
  B2ptOptions o;
  b2pt_default_options(&o);
  if (opt) {
    if (opt->struct_size != sizeof(B2ptOptions)) return pipe_fail(B2PT_ERR_INVALID, "B2ptOptions.struct_size mismatch");
    o = *opt;
  }
  if (o.record_stages) return pipe_fail(B2PT_ERR_INVALID, "stage recording needs a plain context (b2pt_create)");
  o.concurrent_contexts = lanes;
  B2ptPipe* p = new (std::nothrow) B2ptPipe();
  if (!p) return pipe_fail(B2PT_ERR_NOMEM, "out of host memory");
  p->albedo_skip_unchanged = o.persistent_host_albedo != 0;
  p->device = o.device;
  p->floats = (size_t)scene->camera.resolution[0] * (size_t)scene->camera.resolution[1] * 3;
  p->lanes.resize((size_t)lanes);
  int rc = 0;
  cudaError_t e = cudaSetDevice(p->device);
  if (e == cudaSuccess) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) == cudaSuccess) p->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&p->copy, cudaStreamNonBlocking);
  }
  if (e == cudaSuccess) e = cudaMalloc(&p->sum, p->floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMemsetAsync(p->sum, 0, p->floats * sizeof(float), p->copy);
  if (e != cudaSuccess) rc = pipe_fail(B2PT_ERR_CUDA, std::string("b2pt_pipe_create: ") + cudaGetErrorString(e));
  for (int k = 0; k < lanes && rc == 0; ++k) {
    Lane& L = p->lanes[(size_t)k];
    // one copy of the scene on the device: every lane walks the same BVH, which then stays resident in L2
    // (B2PT_PIPE_PRIVATE_SCENES=1: a copy per lane, for A/B measurements)
    const bool share = k > 0 && !getenv("B2PT_PIPE_PRIVATE_SCENES");
    rc = share ? b2pt_create_shared(p->lanes[0].ctx, scene, &o, &L.ctx) : b2pt_create(scene, &o, &L.ctx);
    if (rc) break;
    L.stream = (cudaStream_t)b2pt_stream(L.ctx);
    if (cudaEventCreateWithFlags(&L.rendered, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&L.merged, cudaEventDisableTiming) != cudaSuccess)
      rc = pipe_fail(B2PT_ERR_CUDA, "b2pt_pipe_create: cudaEventCreate failed");
  }
  if (rc == 0 && cudaStreamSynchronize(p->copy) != cudaSuccess) rc = pipe_fail(B2PT_ERR_CUDA, "b2pt_pipe_create: sync failed");
  if (rc) {
    std::string keep = b2pt_last_error();
    b2pt_pipe_destroy(p);
    b2pt_set_last_error_(keep.c_str());
    return rc;
  }
  *out = p;
  return 0;
}

extern "C" int b2pt_pipe_pathtrace(B2ptPipe* p, int32_t iter, float* image_host, float* albedo_host) {
  if (!p) return pipe_fail(B2PT_ERR_INVALID, "pipe is NULL");
  PCK(cudaSetDevice(p->device));
  const int n = (int)p->lanes.size();
  if (!(p->lanes[(size_t)p->head].busy && p->lanes[(size_t)p->head].iter == iter)) {
    if (p->have_last) {
      p->misses += 1;
      p->stride = iter > p->last_iter ? iter - p->last_iter : 1;  // the sequence the host is really on
    }
    int rc = pipe_restart(p, iter);
    if (rc) return rc;
  }
  const int li = p->head;
  Lane& L = p->lanes[(size_t)li];
  PCK(cudaStreamWaitEvent(p->copy, L.rendered, 0));
  k_pipe_merge<<<p->sm_count * 4, 256, 0, p->copy>>>(p->sum, b2pt_device_image(L.ctx), p->floats / 4, (int)(p->floats % 4));
  PCK(cudaGetLastError());
  PCK(cudaEventRecord(L.merged, p->copy));
  p->merges += 1;
  const size_t bytes = p->floats * sizeof(float);
  if (image_host) PCK(cudaMemcpyAsync(image_host, p->sum, bytes, cudaMemcpyDeviceToHost, p->copy));
  if (albedo_host && !(p->albedo_skip_unchanged && albedo_host == p->albedo_host_last &&
                       p->albedo_host_version == p->albedo_version)) {
    // the albedo AOV lives in the context that rendered (or is rendering) iteration 1: wait for that render,
    // which need not be the one this call consumes (a host that starts at iteration 0 finds iteration 1 one
    // lane ahead)
    Lane& A = p->lanes[(size_t)p->albedo_lane];
    if (A.rendered_valid) PCK(cudaStreamWaitEvent(p->copy, A.rendered, 0));
    PCK(cudaMemcpyAsync(albedo_host, b2pt_device_albedo(p->lanes[(size_t)p->albedo_lane].ctx), bytes, cudaMemcpyDeviceToHost,
                        p->copy));
    p->albedo_host_last = albedo_host;
    p->albedo_host_version = p->albedo_version;
  }
  if (p->track_loop_ms) {
    // the lane's loop events are recorded again by its next render: read them before it is re-armed
    PCK(cudaEventSynchronize(L.rendered));
    p->last_loop_ms = b2pt_last_loop_ms(L.ctx);
  }
  // the lane moves on to the iteration `lanes` steps ahead as soon as its contribution has been consumed
  L.busy = false;
  const long long next = (long long)iter + (long long)n * p->stride;
  if (next <= 0x7fffffffLL) {
    int rc = pipe_enqueue(p, L, li, (int)next, true);
    if (rc) return rc;
  }
  p->head = (p->head + 1) % n;
  p->last_iter = iter;
  p->have_last = true;
  p->last_lane = li;
  PCK(cudaStreamSynchronize(p->copy));
  return 0;
}

extern "C" int b2pt_pipe_reset(B2ptPipe* p, const B2ptCamera* cam) {
  if (!p) return pipe_fail(B2PT_ERR_INVALID, "pipe is NULL");
  PCK(cudaSetDevice(p->device));
  for (Lane& L : p->lanes) {
    int rc = cam ? b2pt_set_camera(L.ctx, cam) : b2pt_reset_accum(L.ctx);  // both zero the lane's image and albedo
    if (rc) return rc;
    rc = b2pt_sync(L.ctx);
    if (rc) return rc;
    L.busy = false;
  }
  PCK(cudaMemsetAsync(p->sum, 0, p->floats * sizeof(float), p->copy));
  PCK(cudaStreamSynchronize(p->copy));
  p->albedo_version += 1;
  p->have_last = false;
  p->stride = 1;
  p->head = 0;
  return 0;
}

extern "C" int b2pt_pipe_tonemap_rgba8(B2ptPipe* p, const float* src_dev, int32_t iter, uint8_t* rgba8_dev) {
  if (!p || !rgba8_dev) return pipe_fail(B2PT_ERR_INVALID, "pipe and rgba8_dev must not be NULL");
  PCK(cudaSetDevice(p->device));
  // the copy stream holds only merges and copies, never the iterations the lanes render ahead
  int rc = b2pt_tonemap_on_stream_(p->lanes[0].ctx, src_dev ? src_dev : p->sum, iter, rgba8_dev, (void*)p->copy);
  if (rc) return rc;
  PCK(cudaStreamSynchronize(p->copy));
  return 0;
}

extern "C" float* b2pt_pipe_device_image(B2ptPipe* p) { return p ? p->sum : nullptr; }
extern "C" float* b2pt_pipe_device_albedo(B2ptPipe* p) {
  return p ? b2pt_device_albedo(p->lanes[(size_t)p->albedo_lane].ctx) : nullptr;
}
extern "C" int32_t b2pt_pipe_lanes(B2ptPipe* p) { return p ? (int32_t)p->lanes.size() : 0; }
extern "C" B2ptCtx* b2pt_pipe_lane(B2ptPipe* p, int32_t k) {
  return (p && k >= 0 && k < (int32_t)p->lanes.size()) ? p->lanes[(size_t)k].ctx : nullptr;
}
extern "C" int64_t b2pt_pipe_launch_count(B2ptPipe* p) {
  if (!p) return 0;
  int64_t n = p->merges;
  for (Lane& L : p->lanes) n += b2pt_launch_count(L.ctx);
  return n;
}
extern "C" int64_t b2pt_pipe_misses(B2ptPipe* p) { return p ? p->misses : 0; }
extern "C" float b2pt_pipe_last_loop_ms(B2ptPipe* p) {
  if (!p) return -1.0f;
  if (!p->track_loop_ms) {  // costs one host wait per call, so it is only paid by hosts that ask
    p->track_loop_ms = true;
    return -1.0f;
  }
  return p->last_loop_ms;
}
