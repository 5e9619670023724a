// ref_cpu_main.cu -- TEST INFRASTRUCTURE (oracle/_ref/ref_cpu).
//
// A HOST build of the reference's own math: it includes the unmodified
// apps/src/intersections.h and apps/src/interactions.h (both
// __host__ __device__) and the unmodified loader apps/src/scene.cpp, and runs
// them on the CPU with OpenMP.  Only the __global__ kernel bodies and the host
// loop of apps/src/pathtrace.cu cannot be reused on a CPU; they are restated
// below, each block citing the lines it follows.  Compiled with nvcc (host
// pass) so that unqualified min/max/cos/sin/sqrt/abs/pow resolve to the same
// float overloads the reference gets (route A).
//
// Purpose: (1) pin oracle/pt_oracle.c bit-for-bit, (2) write the .b2s POD
// scene and golden stage dumps, (3) serve as the "reference" CPU baseline.
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <thrust/random.h>

#include "sceneStructs.h"
#include "scene.h"
#include "glm/glm.hpp"
#include "utilities.h"
#include "intersections.h"
#include "interactions.h"

#include "ref_common.h"

// apps/src/pathtrace.cu:44 (overrides utilities.h:12 for finalGather)
#define REF_GATHER_PI 3.14159265358f

// makeSeededRandomEngine, apps/src/pathtrace.cu:66-70 (defined in the .cu, so
// not reachable from a host build; utilhash is the reference's own).
static thrust::default_random_engine ref_make_rng(int iter, int index, int depth) {
  int h = utilhash((1u << 31) | (depth << 22) | iter) ^ utilhash(index);
  return thrust::default_random_engine(h);
}

// ConcentricSampleDisk, apps/src/pathtrace.cu:225-239.
static glm::vec2 ref_concentric(const glm::vec2& point) {
  glm::vec2 uOffset = 2.f * point - glm::vec2(1, 1);
  if (uOffset.x == 0 && uOffset.y == 0) return glm::vec2(0, 0);
  float theta, r;
  if (std::abs(uOffset.x) > std::abs(uOffset.y)) {
    r = uOffset.x;
    theta = 0.785398f * (uOffset.y / uOffset.x);
  } else {
    r = uOffset.y;
    theta = 1.570796f - 0.785398f * (uOffset.x / uOffset.y);
  }
  return r * glm::vec2(std::cos(theta), std::sin(theta));
}

struct Args {
  std::string scene, outdir;
  int iter_first = 1, iters = 1, dump_iter = -1, threads = 0;
  bool aa = true, dof = false, sort = true, time_only = false;
  std::string b2s;
};

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    std::string s = argv[i];
    auto next = [&]() { return std::string(argv[++i]); };
    if (s == "--scene") a.scene = next();
    else if (s == "--out") a.outdir = next();
    else if (s == "--b2s") a.b2s = next();
    else if (s == "--iter-first") a.iter_first = atoi(next().c_str());
    else if (s == "--iters") a.iters = atoi(next().c_str());
    else if (s == "--dump-iter") a.dump_iter = atoi(next().c_str());
    else if (s == "--threads") a.threads = atoi(next().c_str());
    else if (s == "--no-aa") a.aa = false;
    else if (s == "--dof") a.dof = true;
    else if (s == "--no-sort") a.sort = false;
    else { fprintf(stderr, "unknown arg %s\n", s.c_str()); return 2; }
  }
  if (a.scene.empty()) {
    fprintf(stderr, "usage: ref_cpu --scene S.txt [--out DIR] [--b2s FILE] [--iter-first N] [--iters K] "
                    "[--dump-iter N] [--threads T] [--no-aa] [--dof] [--no-sort]\n");
    return 2;
  }
#ifdef _OPENMP
  if (a.threads > 0) omp_set_num_threads(a.threads);
#endif
  Scene* scene = new Scene(a.scene);
  ref_apply_orbit_camera(scene);
  if (!a.b2s.empty() && !ref_write_b2s(*scene, a.b2s.c_str())) {
    fprintf(stderr, "cannot write %s\n", a.b2s.c_str());
    return 1;
  }
  // pathtraceInit, apps/src/pathtrace.cu:140-169: geoms carry pointers to
  // their faces and textures (host memory here).
  for (size_t i = 0; i < scene->geoms.size(); ++i) {
    Geom& g = scene->geoms[i];
    g.dev_faces = scene->allFaces[i].data();
    // the reference indexes these vectors unguarded (UB for OBJ files whose
    // MTL names no maps); a missing entry means "no texture"
    g.kd = i < scene->kdTextures.size() ? scene->kdTextures[i] : Texture();
    g.ks = i < scene->ksTextures.size() ? scene->ksTextures[i] : Texture();
    g.ke = i < scene->keTextures.size() ? scene->keTextures[i] : Texture();
    g.bump = i < scene->bumpTextures.size() ? scene->bumpTextures[i] : Texture();
  }
  if (a.iters <= 0) return 0;

  const Camera cam = scene->state.camera;
  const int traceDepth = scene->state.traceDepth;
  const int P = cam.resolution.x * cam.resolution.y;
  Geom* geoms = scene->geoms.data();
  const int geoms_size = (int)scene->geoms.size();
  Material* materials = scene->materials.data();

  std::vector<PathSegment> paths(P), ptmp(P);
  std::vector<ShadeableIntersection> isects(P), itmp(P);
  std::vector<glm::vec3> image(P, glm::vec3(0.0f)), albedo(P, glm::vec3(0.0f));
  std::vector<int> order(P);
  std::vector<int32_t> nlive(traceDepth + 2, 0);
  long long segments = 0;
  auto t0 = std::chrono::steady_clock::now();

  for (int it = 0; it < a.iters; ++it) {
    const int iter = a.iter_first + it;
    const bool dump = (!a.outdir.empty() && iter == a.dump_iter);
    RefStageWriter w(a.outdir);

    // generateRayFromCamera, apps/src/pathtrace.cu:248-297
#pragma omp parallel for schedule(static)
    for (int y = 0; y < cam.resolution.y; ++y)
      for (int x = 0; x < cam.resolution.x; ++x) {
        int index = x + (y * cam.resolution.x);
        PathSegment& segment = paths[index];
        thrust::default_random_engine rng = ref_make_rng(iter, index, traceDepth);
        segment.ray.origin = cam.position;
        segment.color = glm::vec3(1.0f, 1.0f, 1.0f);
        float antia_x = x;
        float antia_y = y;
        if (a.aa) {
          thrust::default_random_engine rngANTIA = ref_make_rng(iter, index, traceDepth);
          thrust::uniform_real_distribution<float> uANTIA(-0.5, 0.5);
          antia_x += uANTIA(rngANTIA);
          antia_y += uANTIA(rngANTIA);
        }
        segment.ray.direction = glm::normalize(cam.view
            - cam.right * cam.pixelLength.x * ((float)antia_x - (float)cam.resolution.x * 0.5f)
            - cam.up * cam.pixelLength.y * ((float)antia_y - (float)cam.resolution.y * 0.5f));
        if (a.dof) {
          float lensRadius = 0.8f;
          float focalDistance = 11.0f;
          thrust::uniform_real_distribution<float> uDOF(0, 1);
          if (lensRadius > 0) {
            glm::vec2 pLens = lensRadius * ref_concentric(glm::vec2(uDOF(rng), uDOF(rng)));
            float ft = glm::abs(focalDistance / segment.ray.direction.z);
            glm::vec3 pFocus = segment.ray.origin + segment.ray.direction * ft;
            segment.ray.origin += glm::vec3(pLens.x, pLens.y, 0);
            segment.ray.direction = normalize(pFocus - segment.ray.origin);
          }
        }
        segment.pixelIndex = index;
        segment.remainingBounces = traceDepth;
      }

    int depth = 0;
    int num_paths = P;
    std::fill(nlive.begin(), nlive.end(), 0);
    // while (!iterationComplete), apps/src/pathtrace.cu:584-652
    while (num_paths > 0) {
      if (depth <= traceDepth) nlive[depth] = num_paths;
      segments += num_paths;
      if (dump) w.paths(depth, "in", paths.data(), num_paths);
      // cudaMemset(dev_intersections, 0, ...), :595
      memset(isects.data(), 0, sizeof(ShadeableIntersection) * (size_t)P);
      // computeIntersections, :303-386
#pragma omp parallel for schedule(dynamic, 256)
      for (int path_index = 0; path_index < num_paths; ++path_index) {
        PathSegment pathSegment = paths[path_index];
        float t = -1.0f;
        glm::vec3 normal;
        float t_min = FLT_MAX;
        int hit_geom_index = -1;
        bool outside = true;
        glm::vec2 uv = glm::vec2(0.0f, 0.0f);
        glm::vec3 tmp_intersect, tmp_normal;
        glm::vec2 tmp_uv;
        for (int i = 0; i < geoms_size; i++) {
          Geom& geom = geoms[i];
          if (geom.type == CUBE) {
            t = boxIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, outside);
          } else if (geom.type == SPHERE) {
            t = sphereIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, outside);
          } else if (geom.type == OBJ) {
            t = meshIntersectionTest(geom, pathSegment.ray, tmp_intersect, tmp_normal, tmp_uv, outside);
          }
          if (t > 0.0f && t_min > t) {
            t_min = t;
            hit_geom_index = i;
            normal = tmp_normal;
            uv = tmp_uv;
          }
        }
        if (hit_geom_index == -1) {
          isects[path_index].t = -1.0f;
        } else {
          isects[path_index].t = t_min;
          isects[path_index].materialId = geoms[hit_geom_index].materialid;
          isects[path_index].surfaceNormal = normal;
          isects[path_index].geomId = hit_geom_index;
          isects[path_index].texcoord = uv;
        }
      }
      if (dump) w.hits(depth, isects.data(), num_paths);
      // thrust::sort_by_key(..., sortByMaterial()), :512-516, :612
      if (a.sort) {
        for (int i = 0; i < num_paths; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.begin() + num_paths,
                         [&](int x, int y) { return isects[x].materialId > isects[y].materialId; });
        for (int k = 0; k < num_paths; ++k) {
          itmp[k] = isects[order[k]];
          ptmp[k] = paths[order[k]];
        }
        std::copy(itmp.begin(), itmp.begin() + num_paths, isects.begin());
        std::copy(ptmp.begin(), ptmp.begin() + num_paths, paths.begin());
        if (dump) ref_write_npy(w.name(depth, "sort_perm"), "<i4", order.data(), 4, num_paths, 1);
      }
      if (dump) {
        std::vector<int32_t> px(num_paths);
        for (int i = 0; i < num_paths; ++i) px[i] = paths[i].pixelIndex;
        ref_write_npy(w.name(depth, "sorted_pixel"), "<i4", px.data(), 4, num_paths, 1);
      }
      depth++;
      // shadeFakeMaterial, :397-498
#pragma omp parallel for schedule(dynamic, 256)
      for (int idx = 0; idx < num_paths; ++idx) {
        ShadeableIntersection intersection = isects[idx];
        if (iter == 1 && depth == 1) {
          if (intersection.t > 0.0f) {
            Material material = materials[intersection.materialId];
            glm::vec3 materialColor = material.color;
            albedo[paths[idx].pixelIndex] = materialColor;
            Geom geom = geoms[intersection.geomId];
            if (geom.type == OBJ) {
              glm::vec3 emission(0.0f);
              if (geom.ke.channels) {
                int coordU = (int)(intersection.texcoord.x * geom.ke.width);
                int coordV = (int)(intersection.texcoord.y * geom.ke.height);
                int pixelID = coordV * geom.ke.width + coordU;
                unsigned int colR = (unsigned int)geom.ke.image[pixelID * geom.ke.channels];
                unsigned int colG = (unsigned int)geom.ke.image[pixelID * geom.ke.channels + 1];
                unsigned int colB = (unsigned int)geom.ke.image[pixelID * geom.ke.channels + 2];
                emission = glm::vec3(colR / 255.f, colG / 255.f, colB / 255.f);
              }
              if (emission.x > FLT_EPSILON || emission.y > FLT_EPSILON || emission.z > FLT_EPSILON) {
                albedo[paths[idx].pixelIndex] = (emission * 5.0f);
              } else if (geom.kd.channels) {
                int coordU = (int)(intersection.texcoord.x * geom.kd.width);
                int coordV = (int)(intersection.texcoord.y * geom.kd.height);
                int pixelID = coordV * geom.kd.width + coordU;
                unsigned int colR = (unsigned int)geom.kd.image[pixelID * geom.kd.channels];
                unsigned int colG = (unsigned int)geom.kd.image[pixelID * geom.kd.channels + 1];
                unsigned int colB = (unsigned int)geom.kd.image[pixelID * geom.kd.channels + 2];
                albedo[paths[idx].pixelIndex] = glm::vec3(colR / 255.f, colG / 255.f, colB / 255.f);
              }
            } else if (material.emittance > 0.0f) {
              albedo[paths[idx].pixelIndex] = materialColor * material.emittance;
            } else if (material.hasRefractive > 0.0f) {
              albedo[paths[idx].pixelIndex] = material.specular.color;
            }
          } else {
            albedo[paths[idx].pixelIndex] = glm::vec3(0.0f);
          }
        }
        if (intersection.t > 0.0f) {
          thrust::default_random_engine rng = ref_make_rng(iter, idx, 0);
          Material material = materials[intersection.materialId];
          glm::vec3 materialColor = material.color;
          if (material.emittance > 0.0f) {
            paths[idx].color *= (materialColor * material.emittance);
            paths[idx].remainingBounces = 0;
          } else if (paths[idx].remainingBounces == 1) {
            paths[idx].color = glm::vec3(0.0);
            paths[idx].remainingBounces = 0;
          } else {
            scatterRay(paths[idx], paths[idx].ray.origin + intersection.t * paths[idx].ray.direction,
                       intersection, material, rng, geoms, iter, depth);
            paths[idx].remainingBounces -= 1;
          }
        } else {
          paths[idx].color = glm::vec3(0.0f);
          paths[idx].remainingBounces = 0;
        }
      }
      if (dump) w.paths(depth - 1, "shaded", paths.data(), num_paths);
      // thrust::stable_partition(..., isTerminate()), :518-522, :649
      auto mid = std::stable_partition(paths.begin(), paths.begin() + num_paths,
                                       [](const PathSegment& p) { return p.remainingBounces > 0; });
      if (dump) {
        std::vector<int32_t> px(num_paths);
        for (int i = 0; i < num_paths; ++i) px[i] = paths[i].pixelIndex;
        ref_write_npy(w.name(depth - 1, "part_pixel"), "<i4", px.data(), 4, num_paths, 1);
      }
      num_paths = (int)(mid - paths.begin());
    }
    // finalGather, :501-510
    for (int index = 0; index < P; ++index) {
      PathSegment iterationPath = paths[index];
      image[iterationPath.pixelIndex] += iterationPath.color * REF_GATHER_PI;
    }
    if (dump) ref_write_npy(a.outdir + "/nlive.npy", "<i4", nlive.data(), 4, traceDepth + 1, 1);
  }
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (!a.outdir.empty()) {
    ref_write_npy(a.outdir + "/image.npy", "<f4", image.data(), 4, P, 3);
    ref_write_npy(a.outdir + "/albedo.npy", "<f4", albedo.data(), 4, P, 3);
  }
  int threads = 1;
#ifdef _OPENMP
  threads = omp_get_max_threads();
#endif
  printf("REF_CPU_RESULT {\"iters\": %d, \"seconds\": %.6f, \"ms_per_iter\": %.3f, \"mpaths_per_s\": %.6f, "
         "\"segments\": %lld, \"threads\": %d, \"width\": %d, \"height\": %d, \"depth\": %d}\n",
         a.iters, sec, 1e3 * sec / a.iters, (double)P * a.iters / sec / 1e6, segments, threads,
         cam.resolution.x, cam.resolution.y, traceDepth);
  return 0;
}
