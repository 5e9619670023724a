// ref_adapter_main.cpp -- TEST INFRASTRUCTURE (oracle/_ref/ref_adapter).
//
// A headless stand-in for the reference's host loop (apps/src/main.cpp:221-281, runCuda) built from the
// reference's OWN Scene / scene.cpp / pathtrace.h, linked against integration/pathtrace_b2pt.cpp instead of
// apps/src/pathtrace.cu.  It proves the drop-in: same five symbols, same call order (Free, Init, then
// pathtrace(pbo, 0, ++iteration) once per frame), scene->state.image / .albedo as the output contract.
//   ref_adapter --scene S.txt [--iters K] [--iter-first N] [--out DIR]
#include <chrono>
#include <string>
#include <vector>

#include "pathtrace.h"
#include "ref_common.h"

int main(int argc, char** argv) {
  std::string scene_path, outdir;
  int iter_first = 1, iters = 1;
  for (int i = 1; i < argc; ++i) {
    std::string s = argv[i];
    auto next = [&]() { return std::string(argv[++i]); };
    if (s == "--scene") scene_path = next();
    else if (s == "--out") outdir = next();
    else if (s == "--iter-first") iter_first = atoi(next().c_str());
    else if (s == "--iters") iters = atoi(next().c_str());
    else { fprintf(stderr, "unknown arg %s\n", s.c_str()); return 2; }
  }
  if (scene_path.empty()) {
    fprintf(stderr, "usage: ref_adapter --scene S.txt [--iters K] [--iter-first N] [--out DIR]\n");
    return 2;
  }
  Scene* scene = new Scene(scene_path);
  ref_apply_orbit_camera(scene);  // main.cpp:67-81,222-240
  const int P = scene->state.camera.resolution.x * scene->state.camera.resolution.y;
  pathtraceFree();  // main.cpp:245-248: Free, then Init
  pathtraceInit(scene);
  double loop_ms = 0.0;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; ++i) {
    pathtrace(NULL, 0, iter_first + i);  // main.cpp:261
    loop_ms += timer().getGpuElapsedTimeForPreviousOperation();  // main.cpp:263
  }
  const double call_ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("REF_ADAPTER_RESULT {\"iters\": %d, \"timer_ms_per_iter\": %.4f, \"call_ms_per_iter\": %.4f, "
         "\"mpaths_per_s_call\": %.4f, \"width\": %d, \"height\": %d}\n",
         iters, loop_ms / iters, call_ms / iters, (double)P * iters / (call_ms * 1e-3) / 1e6,
         scene->state.camera.resolution.x, scene->state.camera.resolution.y);
  if (!outdir.empty()) {
    ref_write_npy(outdir + "/image.npy", "<f4", scene->state.image.data(), 4, P, 3);
    ref_write_npy(outdir + "/albedo.npy", "<f4", scene->state.albedo.data(), 4, P, 3);
  }
  pathtraceFree();
  return 0;
}
