// ref_gpu_main.cu -- TEST INFRASTRUCTURE (oracle/_ref/ref_gpu*).
//
// Headless driver around the UNMODIFIED reference CUDA path tracer.  The
// reference translation unit is included textually (REF_PATHTRACE_CU points at
// /root/reference/apps/src/pathtrace.cu, or at a sed-generated variant under
// oracle/_ref/gen/ that only flips one of its compile-time switches), so its
// kernels, functors, static device pointers and pathtrace() are the
// reference's own.  This file adds
//   --time  : calls the reference's pathtrace() and reports its own timer
//             (the "loop" window, apps/src/pathtrace.cu:583-653) plus the
//             wall time of the whole call;
//   --dump  : replays the host loop of pathtrace() (apps/src/pathtrace.cu:
//             572-655) launching the reference's kernels and thrust calls
//             verbatim, copying every stage to .npy files in between.
// It is the authority for the bit-exact gates (hit ids, sort and partition
// permutations) and the GPU comparator of BASELINE.md (B-GPU).
#ifndef REF_PATHTRACE_CU
#error "define REF_PATHTRACE_CU to the reference pathtrace.cu"
#endif
#include REF_PATHTRACE_CU

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "ref_common.h"

static void die_on_cuda(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "ref_gpu: CUDA error at %s: %s\n", what, cudaGetErrorString(e));
    exit(1);
  }
}

// One iteration of apps/src/pathtrace.cu:572-655 with stage dumps.
static void ref_iteration_with_dumps(int iter, const std::string& outdir, bool dump, std::vector<int32_t>& nlive,
                                     long long& segments) {
  const int traceDepth = hst_scene->state.traceDepth;
  const Camera& cam = hst_scene->state.camera;
  const int pixelcount = cam.resolution.x * cam.resolution.y;
  const dim3 blockSize2d(8, 8);
  const dim3 blocksPerGrid2d((cam.resolution.x + blockSize2d.x - 1) / blockSize2d.x,
                             (cam.resolution.y + blockSize2d.y - 1) / blockSize2d.y);
  const int blockSize1d = 128;
  RefStageWriter w(outdir);
  std::vector<PathSegment> hp(pixelcount);
  std::vector<ShadeableIntersection> hi(pixelcount);

  generateRayFromCamera<<<blocksPerGrid2d, blockSize2d>>>(cam, iter, traceDepth, dev_paths);
  die_on_cuda("generate");
  int depth = 0;
  int num_paths = pixelcount;
  std::fill(nlive.begin(), nlive.end(), 0);
  while (num_paths > 0) {
    if (depth <= traceDepth) nlive[depth] = num_paths;
    segments += num_paths;
    dim3 nb = (num_paths + blockSize1d - 1) / blockSize1d;
    if (dump) {
      cudaMemcpy(hp.data(), dev_paths, sizeof(PathSegment) * num_paths, cudaMemcpyDeviceToHost);
      w.paths(depth, "in", hp.data(), num_paths);
    }
    cudaMemset(dev_intersections, 0, pixelcount * sizeof(ShadeableIntersection));
    computeIntersections<<<nb, blockSize1d>>>(depth, num_paths, dev_paths, dev_geoms, hst_scene->geoms.size(),
                                             dev_intersections);
    die_on_cuda("intersect");
    if (dump) {
      cudaMemcpy(hi.data(), dev_intersections, sizeof(ShadeableIntersection) * num_paths, cudaMemcpyDeviceToHost);
      w.hits(depth, hi.data(), num_paths);
    }
#if SORT_BY_MATERIAL
    thrust::sort_by_key(thrust::device, dev_intersections, dev_intersections + num_paths, dev_paths,
                        sortByMaterial());
    if (dump) {
      cudaMemcpy(hp.data(), dev_paths, sizeof(PathSegment) * num_paths, cudaMemcpyDeviceToHost);
      std::vector<int32_t> px(num_paths);
      for (int i = 0; i < num_paths; ++i) px[i] = hp[i].pixelIndex;
      ref_write_npy(w.name(depth, "sorted_pixel"), "<i4", px.data(), 4, num_paths, 1);
    }
#endif
    depth++;
    shadeFakeMaterial<<<nb, blockSize1d>>>(iter, num_paths, dev_intersections, dev_paths, dev_materials, dev_geoms,
                                          depth, dev_albedo);
    die_on_cuda("shade");
    if (dump) {
      cudaMemcpy(hp.data(), dev_paths, sizeof(PathSegment) * num_paths, cudaMemcpyDeviceToHost);
      w.paths(depth - 1, "shaded", hp.data(), num_paths);
    }
    PathSegment* dev_path_end = thrust::stable_partition(thrust::device, dev_paths, dev_paths + num_paths, isTerminate());
    if (dump) {
      cudaMemcpy(hp.data(), dev_paths, sizeof(PathSegment) * num_paths, cudaMemcpyDeviceToHost);
      std::vector<int32_t> px(num_paths);
      for (int i = 0; i < num_paths; ++i) px[i] = hp[i].pixelIndex;
      ref_write_npy(w.name(depth - 1, "part_pixel"), "<i4", px.data(), 4, num_paths, 1);
    }
    num_paths = dev_path_end - dev_paths;
  }
  dim3 numBlocksPixels = (pixelcount + blockSize1d - 1) / blockSize1d;
  finalGather<<<numBlocksPixels, blockSize1d>>>(pixelcount, dev_image, dev_paths);
  die_on_cuda("gather");
  if (dump) ref_write_npy(outdir + "/nlive.npy", "<i4", nlive.data(), 4, traceDepth + 1, 1);
}

int main(int argc, char** argv) {
  std::string scene_path, outdir, b2s;
  int iter_first = 1, iters = 1, dump_iter = -1, warmup = 0;
  bool time_mode = false;
  for (int i = 1; i < argc; ++i) {
    std::string s = argv[i];
    auto next = [&]() { return std::string(argv[++i]); };
    if (s == "--scene") scene_path = next();
    else if (s == "--out") outdir = next();
    else if (s == "--b2s") b2s = next();
    else if (s == "--iter-first") iter_first = atoi(next().c_str());
    else if (s == "--iters") iters = atoi(next().c_str());
    else if (s == "--warmup") warmup = atoi(next().c_str());
    else if (s == "--dump-iter") dump_iter = atoi(next().c_str());
    else if (s == "--time") time_mode = true;
    else { fprintf(stderr, "unknown arg %s\n", s.c_str()); return 2; }
  }
  if (scene_path.empty()) {
    fprintf(stderr, "usage: ref_gpu --scene S.txt [--out DIR] [--b2s FILE] [--iter-first N] [--iters K] "
                    "[--warmup W] [--dump-iter N] [--time]\n");
    return 2;
  }
  Scene* scene = new Scene(scene_path);
  ref_apply_orbit_camera(scene);
  if (!b2s.empty()) ref_write_b2s(*scene, b2s.c_str());
  const int P = scene->state.camera.resolution.x * scene->state.camera.resolution.y;
  const int D = scene->state.traceDepth;

  pathtraceFree();  // apps/src/main.cpp:245-248 order: Free then Init
  pathtraceInit(scene);
  die_on_cuda("init");

  if (time_mode) {
    // The reference's own entry point, untouched.
    for (int i = 0; i < warmup; ++i) pathtrace(NULL, 0, iter_first + i);
    pathtraceFree();
    pathtraceInit(scene);
    double loop_ms = 0.0;
    std::vector<double> each;  // wall time of every call (pathtrace() ends with a device sync, pathtrace.cu:670)
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) {
      auto c0 = std::chrono::steady_clock::now();
      pathtrace(NULL, 0, iter_first + i);
      cudaDeviceSynchronize();
      each.push_back(1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count());
      loop_ms += timer().getGpuElapsedTimeForPreviousOperation();
    }
    cudaDeviceSynchronize();
    double call_ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<double> sorted = each;
    std::sort(sorted.begin(), sorted.end());
    const double med = sorted.empty() ? 0.0 : (sorted.size() & 1 ? sorted[sorted.size() / 2]
                                                                : 0.5 * (sorted[sorted.size() / 2 - 1] + sorted[sorted.size() / 2]));
    printf("REF_GPU_RESULT {\"iters\": %d, \"warmup\": %d, \"loop_ms_per_iter\": %.4f, \"call_ms_per_iter\": %.4f, "
           "\"call_ms_min\": %.4f, \"call_ms_median\": %.4f, \"call_ms_max\": %.4f, "
           "\"mpaths_per_s_loop\": %.4f, \"mpaths_per_s_call\": %.4f, \"width\": %d, \"height\": %d, \"depth\": %d}\n",
           iters, warmup, loop_ms / iters, call_ms / iters, sorted.empty() ? 0.0 : sorted.front(), med,
           sorted.empty() ? 0.0 : sorted.back(), (double)P * iters / (loop_ms * 1e-3) / 1e6,
           (double)P * iters / (call_ms * 1e-3) / 1e6, scene->state.camera.resolution.x,
           scene->state.camera.resolution.y, D);
  } else {
    std::vector<int32_t> nlive(D + 2, 0);
    long long segments = 0;
    for (int i = 0; i < iters; ++i) {
      int iter = iter_first + i;
      ref_iteration_with_dumps(iter, outdir, !outdir.empty() && iter == dump_iter, nlive, segments);
    }
    printf("REF_GPU_RESULT {\"iters\": %d, \"segments\": %lld}\n", iters, segments);
  }
  if (!outdir.empty()) {
    std::vector<glm::vec3> img(P), alb(P);
    cudaMemcpy(img.data(), dev_image, sizeof(glm::vec3) * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(alb.data(), dev_albedo, sizeof(glm::vec3) * P, cudaMemcpyDeviceToHost);
    ref_write_npy(outdir + "/image.npy", "<f4", img.data(), 4, P, 3);
    ref_write_npy(outdir + "/albedo.npy", "<f4", alb.data(), 4, P, 3);
  }
  pathtraceFree();
  return 0;
}
