// ref_common.h -- TEST INFRASTRUCTURE (oracle/_ref builds only).
//
// Glue compiled TOGETHER WITH the unmodified reference sources under
// /root/reference/apps/src (never copied into this repo): a headless stand-in
// for the camera set-up of apps/src/main.cpp, a writer for the POD scene file
// (.b2s) that tests feed to both the oracle and the CUDA path ("same-POD
// inputs"), and a minimal .npy writer for stage dumps.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "scene.h"         // the reference's own Scene (apps/src/scene.h)
#include "sceneStructs.h"  // the reference's own PODs

#include "../../include/b2pt.h"

// main() of apps/src/main.cpp:67-81 followed by the camchanged block of
// runCuda(), apps/src/main.cpp:222-240.  main.cpp itself cannot be built
// here (GLFW/GLEW/OIDN), so its camera arithmetic is restated.
static inline void ref_apply_orbit_camera(Scene* scene) {
  Camera& cam = scene->state.camera;
  glm::vec3 view = cam.view;
  glm::vec3 viewXZ = glm::vec3(view.x, 0.0f, view.z);
  glm::vec3 viewZY = glm::vec3(0.0f, view.y, view.z);
  float phi = glm::acos(glm::dot(glm::normalize(viewXZ), glm::vec3(0, 0, -1)));
  float theta = glm::acos(glm::dot(glm::normalize(viewZY), glm::vec3(0, 1, 0)));
  float zoom = glm::length(cam.position - cam.lookAt);

  glm::vec3 cameraPosition;
  cameraPosition.x = zoom * sinf(phi) * sinf(theta);
  cameraPosition.y = zoom * cosf(theta);
  cameraPosition.z = zoom * cosf(phi) * sinf(theta);
  cam.view = -glm::normalize(cameraPosition);
  glm::vec3 v = cam.view;
  glm::vec3 u = glm::vec3(0, 1, 0);
  glm::vec3 r = glm::cross(v, u);
  cam.up = glm::cross(r, v);
  cam.right = r;
  cameraPosition += cam.lookAt;
  cam.position = cameraPosition;
}

// ---- .b2s : POD scene file --------------------------------------------------
//   char magic[8] = "B2SCENE1"
//   int32 n_geoms, n_materials, n_textures, n_faces, trace_depth, iterations
//   B2ptCamera camera
//   B2ptGeom geoms[n_geoms]            (texture fields are indices)
//   B2ptMaterial materials[n_materials]
//   per texture: int32 w, h, channels, reserved; then w*h*channels bytes
//   float face_pos[n_faces*9]; float face_uv[n_faces*6]
static inline bool ref_write_b2s(const Scene& sc, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  std::vector<B2ptGeom> geoms;
  std::vector<B2ptTexture> texs;
  std::vector<const unsigned char*> texdata;
  std::vector<float> pos, uv;
  auto add_tex = [&](const std::vector<Texture>& v, size_t i) -> int32_t {
    if (i >= v.size() || v[i].channels == 0 || v[i].image == NULL) return -1;
    B2ptTexture t;
    memset(&t, 0, sizeof t);
    t.width = v[i].width;
    t.height = v[i].height;
    t.channels = v[i].channels;
    texs.push_back(t);
    texdata.push_back(v[i].image);
    return (int32_t)texs.size() - 1;
  };
  // pathtraceInit indexes the four texture vectors by geom id
  // (apps/src/pathtrace.cu:146-168); do the same.
  for (size_t i = 0; i < sc.geoms.size(); ++i) {
    const Geom& g = sc.geoms[i];
    B2ptGeom o;
    memset(&o, 0, sizeof o);
    o.type = (int32_t)g.type;
    o.material_id = g.materialid;
    memcpy(o.transform, &g.transform[0][0], 64);
    memcpy(o.inverse_transform, &g.inverseTransform[0][0], 64);
    memcpy(o.inv_transpose, &g.invTranspose[0][0], 64);
    o.face_begin = (int32_t)(pos.size() / 9);
    o.face_count = g.faceSize;
    for (int k = 0; k < g.faceSize; ++k) {
      const Face& fc = sc.allFaces[i][k];
      const Vertex* vs[3] = {&fc.v0, &fc.v1, &fc.v2};
      for (int j = 0; j < 3; ++j) {
        pos.push_back(vs[j]->position.x);
        pos.push_back(vs[j]->position.y);
        pos.push_back(vs[j]->position.z);
        uv.push_back(vs[j]->texcoord.x);
        uv.push_back(vs[j]->texcoord.y);
      }
    }
    o.tex_kd = add_tex(sc.kdTextures, i);
    o.tex_ks = add_tex(sc.ksTextures, i);
    o.tex_bump = add_tex(sc.bumpTextures, i);
    o.tex_ke = add_tex(sc.keTextures, i);
    geoms.push_back(o);
  }
  int32_t hdr[6] = {(int32_t)geoms.size(), (int32_t)sc.materials.size(), (int32_t)texs.size(),
                    (int32_t)(pos.size() / 9), sc.state.traceDepth, (int32_t)sc.state.iterations};
  fwrite("B2SCENE1", 1, 8, f);
  fwrite(hdr, 4, 6, f);
  static_assert(sizeof(Camera) == sizeof(B2ptCamera), "camera layout");
  static_assert(sizeof(Material) == sizeof(B2ptMaterial), "material layout");
  fwrite(&sc.state.camera, sizeof(Camera), 1, f);
  fwrite(geoms.data(), sizeof(B2ptGeom), geoms.size(), f);
  fwrite(sc.materials.data(), sizeof(Material), sc.materials.size(), f);
  for (size_t i = 0; i < texs.size(); ++i) {
    int32_t th[4] = {texs[i].width, texs[i].height, texs[i].channels, 0};
    fwrite(th, 4, 4, f);
    fwrite(texdata[i], 1, (size_t)texs[i].width * texs[i].height * texs[i].channels, f);
  }
  fwrite(pos.data(), 4, pos.size(), f);
  fwrite(uv.data(), 4, uv.size(), f);
  fclose(f);
  return true;
}

// ---- .npy writer (format 1.0, C order, little endian) -------------------------
static inline bool ref_write_npy(const std::string& path, const char* descr, const void* data,
                                 size_t elem_size, size_t rows, size_t cols) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  char dict[256];
  if (cols > 1)
    snprintf(dict, sizeof dict, "{'descr': '%s', 'fortran_order': False, 'shape': (%zu, %zu), }", descr, rows, cols);
  else
    snprintf(dict, sizeof dict, "{'descr': '%s', 'fortran_order': False, 'shape': (%zu,), }", descr, rows);
  std::string h(dict);
  size_t total = 10 + h.size() + 1;
  size_t pad = (64 - total % 64) % 64;
  h.append(pad, ' ');
  h.push_back('\n');
  unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(h.size() & 255),
                           (unsigned char)(h.size() >> 8)};
  fwrite(pre, 1, 10, f);
  fwrite(h.data(), 1, h.size(), f);
  fwrite(data, elem_size, rows * cols, f);
  fclose(f);
  return true;
}

// Split reference AoS buffers into the SoA stage arrays the tests compare.
struct RefStageWriter {
  std::string dir;
  explicit RefStageWriter(const std::string& d) : dir(d) {}
  std::string name(int depth, const char* what) const {
    char b[64];
    snprintf(b, sizeof b, "/d%02d_%s.npy", depth, what);
    return dir + b;
  }
  void paths(int depth, const char* prefix, const PathSegment* p, size_t n) const {
    std::vector<float> o(n * 3), d(n * 3), c(n * 3);
    std::vector<int32_t> px(n), rb(n);
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) {
        o[i * 3 + k] = p[i].ray.origin[k];
        d[i * 3 + k] = p[i].ray.direction[k];
        c[i * 3 + k] = p[i].color[k];
      }
      px[i] = p[i].pixelIndex;
      rb[i] = p[i].remainingBounces;
    }
    std::string pre(prefix);
    ref_write_npy(name(depth, (pre + "_origin").c_str()), "<f4", o.data(), 4, n, 3);
    ref_write_npy(name(depth, (pre + "_dir").c_str()), "<f4", d.data(), 4, n, 3);
    ref_write_npy(name(depth, (pre + "_color").c_str()), "<f4", c.data(), 4, n, 3);
    ref_write_npy(name(depth, (pre + "_pixel").c_str()), "<i4", px.data(), 4, n, 1);
    ref_write_npy(name(depth, (pre + "_bounces").c_str()), "<i4", rb.data(), 4, n, 1);
  }
  void hits(int depth, const ShadeableIntersection* h, size_t n) const {
    std::vector<float> t(n), nr(n * 3), uv(n * 2);
    std::vector<int32_t> g(n), m(n);
    for (size_t i = 0; i < n; ++i) {
      t[i] = h[i].t;
      for (int k = 0; k < 3; ++k) nr[i * 3 + k] = h[i].surfaceNormal[k];
      uv[i * 2] = h[i].texcoord.x;
      uv[i * 2 + 1] = h[i].texcoord.y;
      g[i] = h[i].geomId;
      m[i] = h[i].materialId;
    }
    ref_write_npy(name(depth, "hit_t"), "<f4", t.data(), 4, n, 1);
    ref_write_npy(name(depth, "hit_normal"), "<f4", nr.data(), 4, n, 3);
    ref_write_npy(name(depth, "hit_uv"), "<f4", uv.data(), 4, n, 2);
    ref_write_npy(name(depth, "hit_geom"), "<i4", g.data(), 4, n, 1);
    ref_write_npy(name(depth, "hit_material"), "<i4", m.data(), 4, n, 1);
  }
};
