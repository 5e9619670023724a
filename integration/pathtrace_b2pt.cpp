// pathtrace_b2pt.cpp -- the adapter a MyGPURaytracer maintainer adds.
//
// It REPLACES apps/src/pathtrace.cu in the reference's build and keeps the five
// symbols of apps/src/pathtrace.h:6-10, forwarding them to the C ABI of
// include/b2pt.h.  main.cpp, preview.cpp, scene.cpp and every header of the
// reference stay untouched; this file is compiled against the reference's own
// headers ("pathtrace.h", "scene.h", glm) and linked with libb2pt.so.
//
// `make -C oracle _ref/ref_adapter` builds it together with the reference's
// scene loader and a headless host (oracle/ref_driver/ref_adapter_main.cpp);
// tests/test_gpu_vs_reference.py then checks that the reference's own host code
// driving this adapter produces the image the reference's own pathtrace.cu does.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pathtrace.h"  // the reference's header: timer, pathtraceInit, pathtraceFree, pathtrace, sendToGPU

#include "b2pt.h"  // include/b2pt.h

static Scene* hst_scene = NULL;
// main.cpp asks for one iteration per call with consecutive numbers (main.cpp:255): a B2ptPipe renders the
// next ones on the same GPU while this one is copied to scene->state.image (B2PT_PIPE_LANES=1: no look-ahead)
static B2ptPipe* pipe = NULL;
static float* dev_denoised = NULL;
static bool albedo_on_host = false;

static void b2pt_check(int rc, const char* what) {
  if (rc < 0) {  // the reference's checkCUDAError prints and exits (apps/src/pathtrace.cu:46-64)
    fprintf(stderr, "b2pt error (%s): %s\n", what, b2pt_last_error());
    exit(EXIT_FAILURE);
  }
}

PerformanceTimer& timer() {
  static PerformanceTimer t;
  return t;
}

void pathtraceInit(Scene* scene) {
  hst_scene = scene;
  std::vector<B2ptGeom> geoms;
  std::vector<B2ptTexture> tex;
  std::vector<float> pos, uv;
  // pathtraceInit indexes the four texture vectors by geom id (apps/src/pathtrace.cu:146-168)
  auto add_tex = [&](const std::vector<Texture>& v, size_t i) -> int {
    if (i >= v.size() || !v[i].channels || !v[i].image) return -1;
    B2ptTexture t;
    memset(&t, 0, sizeof t);
    t.width = v[i].width;
    t.height = v[i].height;
    t.channels = v[i].channels;
    t.texels = v[i].image;
    tex.push_back(t);
    return (int)tex.size() - 1;
  };
  for (size_t i = 0; i < scene->geoms.size(); ++i) {
    const Geom& g = scene->geoms[i];
    B2ptGeom o;
    memset(&o, 0, sizeof o);
    o.type = (int)g.type;
    o.material_id = g.materialid;
    memcpy(o.transform, &g.transform[0][0], 64);  // glm::mat4 is 16 column-major floats
    memcpy(o.inverse_transform, &g.inverseTransform[0][0], 64);
    memcpy(o.inv_transpose, &g.invTranspose[0][0], 64);
    o.face_begin = (int)(pos.size() / 9);
    o.face_count = g.faceSize;
    for (int k = 0; k < g.faceSize; ++k) {
      const Face& f = scene->allFaces[i][k];
      const Vertex* vs[3] = {&f.v0, &f.v1, &f.v2};
      for (int j = 0; j < 3; ++j) {
        pos.push_back(vs[j]->position.x);
        pos.push_back(vs[j]->position.y);
        pos.push_back(vs[j]->position.z);
        uv.push_back(vs[j]->texcoord.x);
        uv.push_back(vs[j]->texcoord.y);
      }
    }
    o.tex_kd = add_tex(scene->kdTextures, i);
    o.tex_ks = add_tex(scene->ksTextures, i);
    o.tex_bump = add_tex(scene->bumpTextures, i);
    o.tex_ke = add_tex(scene->keTextures, i);
    geoms.push_back(o);
  }
  B2ptScene s;
  memset(&s, 0, sizeof s);
  s.n_geoms = (int)geoms.size();
  s.geoms = geoms.data();
  s.n_materials = (int)scene->materials.size();
  static_assert(sizeof(Material) == sizeof(B2ptMaterial), "identical 44-byte layout");
  s.materials = reinterpret_cast<const B2ptMaterial*>(scene->materials.data());
  s.n_textures = (int)tex.size();
  s.textures = tex.data();
  s.n_faces = (int)(pos.size() / 9);
  s.face_pos = pos.data();
  s.face_uv = uv.data();
  static_assert(sizeof(Camera) == sizeof(B2ptCamera), "identical 84-byte layout");
  memcpy(&s.camera, &scene->state.camera, sizeof(B2ptCamera));
  s.trace_depth = scene->state.traceDepth;
  s.iterations = (int)scene->state.iterations;
  B2ptOptions opt;
  b2pt_default_options(&opt);  // = the macros of apps/src/pathtrace.cu:36-42
  // scene->state.albedo is one vector sized by the loader and written by nobody but pathtrace(): the library
  // may skip the copy of an albedo AOV that has not changed since the previous call (it changes on iteration 1)
  opt.persistent_host_albedo = 1;
  int lanes = 4;
  if (const char* e = getenv("B2PT_PIPE_LANES")) lanes = atoi(e) > 0 ? atoi(e) : 1;
  b2pt_check(b2pt_pipe_create(&s, &opt, lanes, &pipe), "pathtraceInit");
  cudaMalloc(&dev_denoised, sizeof(glm::vec3) * scene->state.image.size());
  albedo_on_host = false;
  // state.image / state.albedo are sized once by the loader (scene.cpp:379-381): page-lock them so that the
  // per-iteration device-to-host copy of pathtrace() runs at PCIe speed instead of through a staging buffer
  cudaHostRegister(scene->state.image.data(), sizeof(glm::vec3) * scene->state.image.size(), cudaHostRegisterDefault);
  cudaHostRegister(scene->state.albedo.data(), sizeof(glm::vec3) * scene->state.albedo.size(), cudaHostRegisterDefault);
  cudaGetLastError();  // registration is an optimisation: a failure only means pageable copies
}

void pathtraceFree() {  // tolerates the Free-before-Init of apps/src/main.cpp:245-248
  if (pipe && hst_scene) {
    cudaHostUnregister(hst_scene->state.image.data());
    cudaHostUnregister(hst_scene->state.albedo.data());
    cudaGetLastError();
  }
  b2pt_pipe_destroy(pipe);
  pipe = NULL;
  cudaFree(dev_denoised);
  dev_denoised = NULL;
}

void pathtrace(uchar4* /*pbo*/, int /*frame*/, int iter) {  // pbo and frame are unused upstream too (AI_DENOISE 1)
  // apps/src/pathtrace.cu:663-668: running sum and albedo AOV to scene->state every call; the albedo only
  // changes on iteration 1 (pathtrace.cu:412), the pipe skips the copies that would rewrite the same bytes
  timer().startGpuTimer();  // main.cpp:263 reads timer().getGpuElapsedTimeForPreviousOperation()
  b2pt_check(b2pt_pipe_pathtrace(pipe, iter, reinterpret_cast<float*>(hst_scene->state.image.data()),
                                 reinterpret_cast<float*>(hst_scene->state.albedo.data())),
             "pathtrace");
  timer().endGpuTimer();
  albedo_on_host = true;
}

void sendToGPU(uchar4* pbo, int /*iter*/) {  // apps/src/pathtrace.cu:673-685
  cudaMemcpy(dev_denoised, hst_scene->state.output.data(), sizeof(glm::vec3) * hst_scene->state.output.size(),
             cudaMemcpyHostToDevice);
  // synchronous for the caller: main.cpp unmaps the PBO right after this call (main.cpp:270-271) and the next
  // frame overwrites dev_denoised
  b2pt_check(b2pt_pipe_tonemap_rgba8(pipe, dev_denoised, 0, reinterpret_cast<uint8_t*>(pbo)), "sendToGPU");
}
