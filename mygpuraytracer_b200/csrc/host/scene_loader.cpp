// scene_loader.cpp -- the scenes/<name>.txt format, OBJ/MTL meshes and textures.
//
// Drop-in surface of apps/src/scene.cpp (Scene::Scene :10-36, loadMaterial
// :388-423, loadCamera :324-386, loadGeom :236-322, loadObj :38-234) and of the
// camera set-up main.cpp performs before the first pathtrace() call
// (apps/src/main.cpp:67-81 and 222-240), producing the POD scene of
// include/b2pt.h.  Differences from the reference, all deliberate:
//   * errors are returned (B2PT_ERR_IO / B2PT_ERR_INVALID), never exit()/throw;
//   * every geom gets four texture slots, so OBJ files whose MTL names no maps
//     are safe (the reference indexes short vectors);
//   * "..\\textures\\x.jpg" in an MTL is normalised to forward slashes and
//     looked up relative to the working directory, the scene file, the OBJ and
//     the MTL, so the shipped spaceship MTL loads on Linux;
//   * RES / ITERATIONS / DEPTH can be overridden (every shipped scene says
//     800x800, 5000, 8).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../../include/b2pt.h"
#include "hostmath.h"
#include "image_decode.h"

using namespace b2host;

namespace {

thread_local std::string g_loader_error;

struct LoadedTexture {
  int w = 0, h = 0, c = 0;
  std::vector<uint8_t> texels;
};

}  // namespace

struct B2ptLoadedScene {
  std::vector<B2ptGeom> geoms;
  std::vector<B2ptMaterial> materials;
  std::vector<LoadedTexture> tex_store;
  std::vector<B2ptTexture> textures;
  std::vector<float> face_pos, face_uv;
  bool per_face_materials = false;           // B2ptLoadOverrides::per_face_materials
  std::vector<int32_t> face_material;        // [n_faces] when per_face_materials
  std::vector<int32_t> material_textures;    // [4 * n_materials] (kd, ks, bump, ke) when per_face_materials
  B2ptScene view{};
  std::string image_name;
  std::string scene_dir;
  std::string warnings;  // one line per thing the loader tolerated the way the reference does (b2pt_scene_warnings)
};

extern "C" const char* b2pt_last_error(void);
// defined in b2pt.cu: lets the loader report through the same channel
extern "C" void b2pt_set_last_error_(const char* msg);

namespace {

int fail(int code, const std::string& msg) {
  b2pt_set_last_error_(msg.c_str());
  return code;
}

// utilityCore::safeGetline, apps/src/utilities.cpp:82-112: \n, \r\n and \r end a line.
bool get_line(std::istream& is, std::string& t) {
  t.clear();
  bool any = false;
  for (;;) {
    int c = is.get();
    if (c == EOF) return any || !t.empty();
    any = true;
    if (c == '\n') return true;
    if (c == '\r') {
      if (is.peek() == '\n') is.get();
      return true;
    }
    t += (char)c;
  }
}

std::vector<std::string> tokens_of(const std::string& s) {
  std::istringstream ss(s);
  std::vector<std::string> r;
  std::string w;
  while (ss >> w) r.push_back(w);
  return r;
}

std::string dir_of(const std::string& p) {
  size_t k = p.find_last_of('/');
  return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

bool file_exists(const std::string& p) {
  std::ifstream f(p.c_str(), std::ios::binary);
  return f.good();
}

std::string slashes(std::string s) {
  std::string r;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '\\') {
      if (r.empty() || r.back() != '/') r += '/';
    } else {
      r += s[i];
    }
  }
  return r;
}

std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

std::string find_file(const std::string& name, const std::vector<std::string>& dirs) {
  if (file_exists(name)) return name;
  for (const std::string& d : dirs) {
    std::string p = d + "/" + name;
    if (file_exists(p)) return p;
  }
  return std::string();
}

struct MtlInfo {
  float kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0}, ke[3] = {0, 0, 0};  // tinyobjloader 2.0.0 starts every material at zero
  float ior = 1.0f;
  std::string map_kd, map_ks, map_ke, map_bump;
  std::string name;
  bool found = false;
};

// The file name of a map_* / bump statement.  tinyobjloader 2.0.0 accepts the MTL texture options in front
// of it and lets each swallow a fixed number of blank-separated arguments, numeric or not
// (apps/src/tiny_obj_loader.h:1235-1325); whatever follows the last option, blanks included, is the name.
static std::string texture_name_of(const std::string& args) {
  static const struct {
    const char* name;
    int nargs;
  } opts[] = {{"-blendu", 1}, {"-blendv", 1}, {"-clamp", 1}, {"-boost", 1}, {"-bm", 1},      {"-o", 3},         {"-s", 3},
              {"-t", 3},      {"-type", 1},   {"-texres", 1}, {"-imfchan", 1}, {"-mm", 2},    {"-colorspace", 1}};
  auto blank = [](char c) { return c == ' ' || c == '\t'; };
  size_t i = 0;
  for (;;) {
    while (i < args.size() && blank(args[i])) ++i;
    if (i >= args.size()) return std::string();
    int nargs = -1;
    for (const auto& o : opts) {
      const size_t n = strlen(o.name);
      if (args.compare(i, n, o.name) == 0 && i + n < args.size() && blank(args[i + n])) {
        nargs = o.nargs;
        i += n;
        break;
      }
    }
    if (nargs < 0) return args.substr(i);
    for (int k = 0; k < nargs; ++k) {
      while (i < args.size() && blank(args[i])) ++i;
      while (i < args.size() && !blank(args[i])) ++i;
    }
  }
}

// The materials of an MTL file in file order (the reference uses objMaterials[0] only, scene.cpp:68,134; the others
// matter with B2ptLoadOverrides::per_face_materials).
std::vector<MtlInfo> parse_mtl(const std::string& path) {
  std::vector<MtlInfo> all;
  std::ifstream f(path.c_str());
  std::string line;
  int count = 0;
  while (get_line(f, line)) {
    std::string t = trim(line);
    if (t.empty() || t[0] == '#') continue;
    std::vector<std::string> tk = tokens_of(t);
    const std::string& k = tk[0];
    if (k == "newmtl") {
      ++count;
      all.emplace_back();
      all.back().found = true;
      all.back().name = trim(t.substr(k.size()));
      continue;
    }
    if (count < 1) continue;
    MtlInfo& m = all.back();
    auto rest = [&]() { return slashes(trim(texture_name_of(t.substr(k.size())))); };
    auto f3 = [&](float* o) {
      for (int i = 0; i < 3 && i + 1 < (int)tk.size(); ++i) o[i] = (float)atof(tk[i + 1].c_str());
    };
    if (k == "Kd") f3(m.kd);
    else if (k == "Ks") f3(m.ks);
    else if (k == "Ke") f3(m.ke);
    else if (k == "Ni" && tk.size() > 1) m.ior = (float)atof(tk[1].c_str());
    else if (k == "map_Kd") m.map_kd = rest();
    else if (k == "map_Ks") m.map_ks = rest();
    else if (k == "map_Ke") m.map_ke = rest();
    else if (k == "map_Bump" || k == "map_bump" || k == "bump") m.map_bump = rest();
  }
  return all;
}

struct ObjIndex {
  int v = -1, vt = -1, vn = -1;
};

bool parse_index(const std::string& tok, int nv, int nvt, int nvn, ObjIndex* out) {
  int vals[3] = {0, 0, 0};
  bool has[3] = {false, false, false};
  int field = 0;
  size_t i = 0;
  while (i <= tok.size() && field < 3) {
    size_t j = tok.find('/', i);
    if (j == std::string::npos) j = tok.size();
    if (j > i) {
      vals[field] = atoi(tok.substr(i, j - i).c_str());
      has[field] = true;
    }
    ++field;
    i = j + 1;
  }
  if (!has[0]) return false;
  auto fix = [](int idx, int n) { return idx > 0 ? idx - 1 : (idx < 0 ? n + idx : -1); };
  out->v = fix(vals[0], nv);
  out->vt = has[1] ? fix(vals[1], nvt) : -1;
  out->vn = has[2] ? fix(vals[2], nvn) : -1;
  return out->v >= 0 && out->v < nv;
}

// Crossing-number test of (tx, ty) against the triangle (px, py) (the pnpoly test tinyobjloader uses,
// apps/src/tiny_obj_loader.h:1414-1426), in single precision and in the same order of operations.
static bool crossing_inside3(const float* px, const float* py, float tx, float ty) {
  bool in = false;
  for (int i = 0, j = 2; i < 3; j = i++) {
    if (((py[i] > ty) != (py[j] > ty)) && (tx < (px[j] - px[i]) * (ty - py[i]) / (py[j] - py[i]) + px[i])) in = !in;
  }
  return in;
}

// loadObj, apps/src/scene.cpp:38-234.  Faces are emitted in file order; quads
// are split along the shorter diagonal as tinyobjloader 2.0.0 does
// (apps/src/tiny_obj_loader.h:1509-1553); polygons with more corners go through
// its ear clipping, restated below (tiny_obj_loader.h:1569-1845), so that concave
// polygons come out as the same triangles in the same order (tests/golden/hardobj.obj).
int load_obj(B2ptLoadedScene* S, const std::string& obj_path, const std::vector<std::string>& search, B2ptGeom* g) {
  std::string path = find_file(slashes(obj_path), search);
  if (path.empty()) return fail(B2PT_ERR_IO, "cannot open OBJ file " + obj_path);
  std::ifstream f(path.c_str());
  std::vector<float> v, vt;
  int nvn = 0;
  std::string line;
  // materials accumulate over the mtllib statements, a usemtl is resolved against what has been loaded when it is
  // read, an unknown name means "no material" (-1): tinyobjloader 2.0.0, apps/src/tiny_obj_loader.h:2729-2810
  std::vector<MtlInfo> mtls;
  std::map<std::string, int> mtl_index;
  std::vector<std::string> tex_dirs = search;
  int cur_mat = -1;
  std::vector<int> face_mids;  // tinyobj's mesh.material_ids, one per emitted triangle
  g->face_begin = (int)(S->face_pos.size() / 9);
  auto emit = [&](const ObjIndex& a, const ObjIndex& b, const ObjIndex& c) {
    face_mids.push_back(cur_mat);
    const ObjIndex* t[3] = {&a, &b, &c};
    for (int k = 0; k < 3; ++k) {
      for (int j = 0; j < 3; ++j) S->face_pos.push_back(v[3 * (size_t)t[k]->v + j]);
      if (t[k]->vt >= 0 && 2 * (size_t)t[k]->vt + 1 < vt.size()) {
        S->face_uv.push_back(vt[2 * (size_t)t[k]->vt]);
        S->face_uv.push_back(vt[2 * (size_t)t[k]->vt + 1]);
      } else {
        S->face_uv.push_back(0.0f);  // glm::vec2() of an unset Vertex::texcoord
        S->face_uv.push_back(0.0f);
      }
    }
  };
  while (get_line(f, line)) {
    const char* p = line.c_str();
    while (*p == ' ' || *p == '\t') ++p;
    if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
      char* e;
      for (int k = 0; k < 3; ++k) {
        v.push_back((float)strtod(p + (k == 0 ? 1 : 0), &e));
        p = e;
      }
    } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
      char* e;
      p += 2;
      for (int k = 0; k < 2; ++k) {
        vt.push_back((float)strtod(p, &e));
        p = e;
      }
    } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
      ++nvn;
    } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
      std::vector<std::string> tk = tokens_of(p + 1);
      std::vector<ObjIndex> idx;
      bool ok = true;
      for (const std::string& t : tk) {
        ObjIndex oi;
        if (!parse_index(t, (int)(v.size() / 3), (int)(vt.size() / 2), nvn, &oi)) {
          ok = false;
          break;
        }
        idx.push_back(oi);
      }
      if (!ok || idx.size() < 3) continue;
      if (idx.size() == 3) {
        emit(idx[0], idx[1], idx[2]);
      } else if (idx.size() == 4) {
        const float* a = &v[3 * (size_t)idx[0].v];
        const float* b = &v[3 * (size_t)idx[1].v];
        const float* c = &v[3 * (size_t)idx[2].v];
        const float* d = &v[3 * (size_t)idx[3].v];
        const float e02[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        const float e13[3] = {d[0] - b[0], d[1] - b[1], d[2] - b[2]};
        const float s02 = e02[0] * e02[0] + e02[1] * e02[1] + e02[2] * e02[2];
        const float s13 = e13[0] * e13[0] + e13[1] * e13[1] + e13[2] * e13[2];
        if (s02 < s13) {
          emit(idx[0], idx[1], idx[2]);
          emit(idx[0], idx[2], idx[3]);
        } else {
          emit(idx[0], idx[1], idx[3]);
          emit(idx[1], idx[2], idx[3]);
        }
      } else {
        // Projection plane: the first corner that is not degenerate decides which coordinate is dropped
        // (the one along which its normal is strictly largest; x is kept on ties).
        int ax0 = 1, ax1 = 2;
        const size_t n0 = idx.size();
        for (size_t k = 0; k < n0; ++k) {
          const float* p0 = &v[3 * (size_t)idx[k].v];
          const float* p1 = &v[3 * (size_t)idx[(k + 1) % n0].v];
          const float* p2 = &v[3 * (size_t)idx[(k + 2) % n0].v];
          const float e0x = p1[0] - p0[0], e0y = p1[1] - p0[1], e0z = p1[2] - p0[2];
          const float e1x = p2[0] - p1[0], e1y = p2[1] - p1[1], e1z = p2[2] - p1[2];
          const float nx = std::fabs(e0y * e1z - e0z * e1y);
          const float ny = std::fabs(e0z * e1x - e0x * e1z);
          const float nz = std::fabs(e0x * e1y - e0y * e1x);
          const float eps = std::numeric_limits<float>::epsilon();
          if (nx > eps || ny > eps || nz > eps) {
            if (!(nx > ny && nx > nz)) {
              ax0 = 0;
              if (nz > nx && nz > ny) ax1 = 1;
            }
            break;
          }
        }
        // Ear clipping: walk a candidate corner around the remaining polygon; a corner is cut off when it turns
        // the way the sign test says and no other remaining vertex lies inside its triangle.  The walk gives up
        // after a full round without progress (what is left is then dropped, as the reference's loader does).
        std::vector<ObjIndex> rest = idx;
        size_t corner = 0, budget = rest.size(), seen = rest.size();
        while (rest.size() > 3 && budget > 0) {
          const size_t n = rest.size();
          if (corner >= n) corner -= n;
          if (seen != n) {
            seen = n;
            budget = n;
          } else {
            --budget;
          }
          ObjIndex t[3];
          float px[3], py[3];
          for (int k = 0; k < 3; ++k) {
            t[k] = rest[(corner + k) % n];
            px[k] = v[3 * (size_t)t[k].v + ax0];
            py[k] = v[3 * (size_t)t[k].v + ax1];
          }
          const float e0x = px[1] - px[0], e0y = py[1] - py[0];
          const float e1x = px[2] - px[1], e1y = py[2] - py[1];
          const float turn = e0x * e1y - e0y * e1x;
          const float sign_ref = (px[0] * py[1] - py[0] * px[1]) * 0.5f;
          if (turn * sign_ref < 0.0f) {
            ++corner;
            continue;
          }
          bool blocked = false;
          for (size_t o = 3; o < n && !blocked; ++o) {
            const ObjIndex& q = rest[(corner + o) % n];
            blocked = crossing_inside3(px, py, v[3 * (size_t)q.v + ax0], v[3 * (size_t)q.v + ax1]);
          }
          if (blocked) {
            ++corner;
            continue;
          }
          emit(t[0], t[1], t[2]);
          rest.erase(rest.begin() + (long)((corner + 1) % n));
        }
        if (rest.size() == 3) emit(rest[0], rest[1], rest[2]);
      }
    } else if (strncmp(p, "usemtl", 6) == 0) {
      const std::vector<std::string> tk = tokens_of(p + 6);
      const auto it = tk.empty() ? mtl_index.end() : mtl_index.find(tk[0]);
      cur_mat = it == mtl_index.end() ? -1 : it->second;
    } else if (strncmp(p, "mtllib", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
      const std::string mtllib = trim(std::string(p + 6));
      if (mtllib.empty()) continue;
      std::vector<std::string> mdirs = {dir_of(path) + "/materials", dir_of(path), "../models/materials"};
      for (const std::string& d : search) mdirs.push_back(d);
      // `mtllib a.mtl b.mtl`: tinyobjloader splits the statement at blanks and takes the first file that loads
      // (apps/src/tiny_obj_loader.h:2761-2806); a name that exists as written (blanks and all) is tried first
      std::vector<std::string> names = {mtllib};
      for (const std::string& n : tokens_of(mtllib)) names.push_back(n);
      for (const std::string& n : names) {
        const std::string mp = find_file(slashes(n), mdirs);
        if (mp.empty()) continue;
        for (MtlInfo& mi : parse_mtl(mp)) {
          mtl_index.insert(std::make_pair(mi.name, (int)mtls.size()));  // the first definition of a name wins
          mtls.push_back(std::move(mi));
        }
        tex_dirs.push_back(dir_of(mp));
        break;
      }
    }
  }
  g->face_count = (int)(S->face_pos.size() / 9) - g->face_begin;

  // material + textures of the first MTL material (objMaterials[0], scene.cpp:68,134)
  const MtlInfo mtl = mtls.empty() ? MtlInfo() : mtls[0];
  tex_dirs.push_back(dir_of(path));
  std::map<std::string, int> tex_cache;  // several materials may name the same file
  auto load_tex = [&](const std::string& name) -> int {
    if (name.empty()) return -1;
    const auto hit = tex_cache.find(name);
    if (hit != tex_cache.end()) return hit->second;
    std::string tp = find_file(name, tex_dirs);
    LoadedTexture t;
    if (tp.empty() || !decode_image_file(tp, /*flip_vertically=*/true, &t.w, &t.h, &t.c, &t.texels)) {
      // "Failed to load ... texture file": the reference pushes an empty Texture (scene.cpp:150-154) and prints
      // the name; here the slot stays empty and the name goes to b2pt_scene_warnings()
      S->warnings += "texture '" + name + "' " + (tp.empty() ? "not found" : "could not be decoded (" + tp + ")") +
                     ": the map is left empty, as in the reference\n";
      tex_cache[name] = -1;
      return -1;
    }
    S->tex_store.push_back(std::move(t));
    tex_cache[name] = (int)S->tex_store.size() - 1;
    return (int)S->tex_store.size() - 1;
  };
  g->tex_kd = load_tex(mtl.map_kd);
  g->tex_ks = load_tex(mtl.map_ks);
  g->tex_bump = load_tex(mtl.map_bump);
  g->tex_ke = load_tex(mtl.map_ke);

  // "New material for this object", scene.cpp:220-231
  auto scene_material = [](const MtlInfo& mi) {
    B2ptMaterial m;
    memset(&m, 0, sizeof m);
    for (int k = 0; k < 3; ++k) {
      m.specular_color[k] = mi.ks[k];
      m.color[k] = mi.kd[k];
    }
    m.specular_exponent = 0.0f;
    m.index_of_refraction = mi.ior;
    m.emittance = mi.ke[0];
    m.has_reflective = 0.0f;
    m.has_refractive = 0.0f;
    return m;
  };
  const int base = (int)S->materials.size();
  S->materials.push_back(scene_material(mtl));
  g->material_id = base;
  if (S->per_face_materials) {
    // Every MTL material becomes a scene material with its own four maps, converted the way the reference converts
    // material 0; a face without a (known) usemtl takes material 0, the one the reference shades everything with.
    S->material_textures.resize(4 * (size_t)base, -1);
    const int tex0[4] = {g->tex_kd, g->tex_ks, g->tex_bump, g->tex_ke};
    S->material_textures.insert(S->material_textures.end(), tex0, tex0 + 4);
    for (size_t k = 1; k < mtls.size(); ++k) {
      S->materials.push_back(scene_material(mtls[k]));
      const int tx[4] = {load_tex(mtls[k].map_kd), load_tex(mtls[k].map_ks), load_tex(mtls[k].map_bump), load_tex(mtls[k].map_ke)};
      S->material_textures.insert(S->material_textures.end(), tx, tx + 4);
    }
    S->face_material.resize((size_t)g->face_begin, 0);
    for (int mid : face_mids) S->face_material.push_back(base + ((mid > 0 && mid < (int)mtls.size()) ? mid : 0));
  }
  return 0;
}

}  // namespace

extern "C" int b2pt_scene_load(const char* path, const B2ptLoadOverrides* ov, B2ptLoadedScene** out) {
  if (!path || !out) return fail(B2PT_ERR_INVALID, "path and out must not be NULL");
  *out = nullptr;
  std::ifstream in(path);
  if (!in.is_open()) return fail(B2PT_ERR_IO, std::string("cannot open scene file ") + path);
  B2ptLoadedScene* S = new B2ptLoadedScene();
  S->per_face_materials = ov && ov->per_face_materials != 0;
  S->scene_dir = dir_of(path);
  const std::vector<std::string> search = {S->scene_dir, S->scene_dir + "/../bin", "."};
  B2ptCamera cam;
  memset(&cam, 0, sizeof cam);
  int trace_depth = 0, iterations = 0;
  bool have_camera = false;
  int rc = 0;
  std::string line;
  // Scene::Scene, scene.cpp:19-35
  while (rc == 0 && get_line(in, line)) {
    if (line.empty()) continue;
    std::vector<std::string> tk = tokens_of(line);
    if (tk.empty()) continue;
    if (tk[0] == "MATERIAL" && tk.size() > 1) {
      // loadMaterial, scene.cpp:388-423: exactly seven property lines
      if (atoi(tk[1].c_str()) != (int)S->materials.size()) {
        rc = fail(B2PT_ERR_INVALID, "MATERIAL id does not match expected number of materials");
        break;
      }
      B2ptMaterial m;
      memset(&m, 0, sizeof m);
      for (int i = 0; i < 7; ++i) {
        if (!get_line(in, line)) break;
        std::vector<std::string> t = tokens_of(line);
        if (t.empty()) continue;
        auto f = [&](size_t k) { return k < t.size() ? (float)atof(t[k].c_str()) : 0.0f; };
        if (t[0] == "RGB") { m.color[0] = f(1); m.color[1] = f(2); m.color[2] = f(3); }
        else if (t[0] == "SPECEX") m.specular_exponent = f(1);
        else if (t[0] == "SPECRGB") { m.specular_color[0] = f(1); m.specular_color[1] = f(2); m.specular_color[2] = f(3); }
        else if (t[0] == "REFL") m.has_reflective = f(1);
        else if (t[0] == "REFR") m.has_refractive = f(1);
        else if (t[0] == "REFRIOR") m.index_of_refraction = f(1);
        else if (t[0] == "EMITTANCE") m.emittance = f(1);
      }
      S->materials.push_back(m);
    } else if (tk[0] == "CAMERA") {
      // loadCamera, scene.cpp:324-386
      float fovy = 0.0f;
      for (int i = 0; i < 5; ++i) {
        if (!get_line(in, line)) break;
        std::vector<std::string> t = tokens_of(line);
        if (t.empty()) continue;
        if (t[0] == "RES" && t.size() > 2) { cam.resolution[0] = atoi(t[1].c_str()); cam.resolution[1] = atoi(t[2].c_str()); }
        else if (t[0] == "FOVY" && t.size() > 1) fovy = (float)atof(t[1].c_str());
        else if (t[0] == "ITERATIONS" && t.size() > 1) iterations = atoi(t[1].c_str());
        else if (t[0] == "DEPTH" && t.size() > 1) trace_depth = atoi(t[1].c_str());
        else if (t[0] == "FILE" && t.size() > 1) S->image_name = t[1];
      }
      float up_in[3] = {0, 0, 0};
      while (get_line(in, line) && !line.empty()) {
        std::vector<std::string> t = tokens_of(line);
        if (t.size() < 4) continue;
        float* dst = t[0] == "EYE" ? cam.position : t[0] == "LOOKAT" ? cam.look_at : t[0] == "UP" ? up_in : nullptr;
        if (dst) for (int k = 0; k < 3; ++k) dst[k] = (float)atof(t[k + 1].c_str());
      }
      if (ov) {
        if (ov->width > 0 && ov->height > 0) { cam.resolution[0] = ov->width; cam.resolution[1] = ov->height; }
        if (ov->iterations > 0) iterations = ov->iterations;
        if (ov->depth > 0) trace_depth = ov->depth;
      }
      // the reference allocates resolution.x * resolution.y pixels without looking (scene.cpp:379-381); a
      // library refuses what b2pt_create would refuse, here, with the file name in the message
      if (cam.resolution[0] <= 0 || cam.resolution[1] <= 0 || (long long)cam.resolution[0] * cam.resolution[1] >= (1ll << 30)) {
        rc = fail(B2PT_ERR_INVALID, "CAMERA: RES must be positive and below 2^30 pixels");
        break;
      }
      if (trace_depth < 0 || trace_depth > 62) {
        rc = fail(B2PT_ERR_RANGE, "CAMERA: DEPTH must be in [0, 62]");
        break;
      }
      const float kPi = 3.1415926535897932384626422832795028841971f;
      // yscaled = tan(fovy * (PI / 180)): the FULL angle
      const float yscaled = std::tan(fovy * (kPi / 180));
      const float xscaled = (yscaled * cam.resolution[0]) / cam.resolution[1];
      const float fovx = (std::atan(xscaled) * 180) / kPi;
      cam.fov[0] = fovx;
      cam.fov[1] = fovy;
      cam.pixel_length[0] = 2 * xscaled / (float)cam.resolution[0];
      cam.pixel_length[1] = 2 * yscaled / (float)cam.resolution[1];
      float d[3] = {cam.look_at[0] - cam.position[0], cam.look_at[1] - cam.position[1], cam.look_at[2] - cam.position[2]};
      normalize3(d, cam.view);
      // main(), main.cpp:67-81: orbit angles from the loaded view
      const float vxz[3] = {cam.view[0], 0.0f, cam.view[2]}, vzy[3] = {0.0f, cam.view[1], cam.view[2]};
      float nxz[3], nzy[3];
      normalize3(vxz, nxz);
      normalize3(vzy, nzy);
      const float fwd[3] = {0, 0, -1}, upv[3] = {0, 1, 0};
      const float phi = std::acos(dot3(nxz, fwd));
      const float theta = std::acos(dot3(nzy, upv));
      const float pl[3] = {cam.position[0] - cam.look_at[0], cam.position[1] - cam.look_at[1], cam.position[2] - cam.look_at[2]};
      const float zoom = std::sqrt(dot3(pl, pl));
      // runCuda(), main.cpp:222-240: the camera the kernels actually see (Q1)
      float cp[3] = {zoom * std::sin(phi) * std::sin(theta), zoom * std::cos(theta), zoom * std::cos(phi) * std::sin(theta)};
      float nv[3];
      normalize3(cp, nv);
      for (int k = 0; k < 3; ++k) cam.view[k] = -nv[k];
      cross3(cam.view, upv, cam.right);   // NOT normalised
      cross3(cam.right, cam.view, cam.up);
      for (int k = 0; k < 3; ++k) cam.position[k] = cp[k] + cam.look_at[k];
      have_camera = true;
    } else if (tk[0] == "OBJECT" && tk.size() > 1) {
      // loadGeom, scene.cpp:236-322
      if (atoi(tk[1].c_str()) != (int)S->geoms.size()) {
        rc = fail(B2PT_ERR_INVALID, "OBJECT id does not match expected number of geoms");
        break;
      }
      B2ptGeom g;
      memset(&g, 0, sizeof g);
      g.tex_kd = g.tex_ks = g.tex_bump = g.tex_ke = -1;
      g.type = -1;
      std::string obj_file;
      if (get_line(in, line) && !line.empty()) {
        std::string t = trim(line);
        if (t == "sphere") g.type = B2PT_SPHERE;
        else if (t == "cube") g.type = B2PT_CUBE;
        else if (t == "triangle") g.type = B2PT_TRIANGLE;
        else if (t == "obj") {
          g.type = B2PT_OBJ;
          if (get_line(in, line) && !line.empty()) obj_file = trim(line);
        }
      }
      if (g.type < 0) {
        rc = fail(B2PT_ERR_INVALID, "unknown OBJECT type in scene file");
        break;
      }
      if (g.type != B2PT_OBJ) {
        if (get_line(in, line) && !line.empty()) {
          std::vector<std::string> t = tokens_of(line);
          if (t.size() > 1) g.material_id = atoi(t[1].c_str());
        }
      } else {
        g.material_id = -1;
      }
      float tr[3] = {0, 0, 0}, ro[3] = {0, 0, 0}, sc[3] = {0, 0, 0};
      while (get_line(in, line) && !line.empty()) {
        std::vector<std::string> t = tokens_of(line);
        if (t.size() < 4) continue;
        float* dst = t[0] == "TRANS" ? tr : t[0] == "ROTAT" ? ro : t[0] == "SCALE" ? sc : nullptr;
        if (dst) for (int k = 0; k < 3; ++k) dst[k] = (float)atof(t[k + 1].c_str());
      }
      const M4 T = build_transform(tr, ro, sc);
      const M4 I = inverse(T), IT = inverse_transpose(T);
      memcpy(g.transform, T.m, 64);
      memcpy(g.inverse_transform, I.m, 64);
      memcpy(g.inv_transpose, IT.m, 64);
      if (g.type == B2PT_OBJ) {
        rc = load_obj(S, obj_file, search, &g);
        if (rc) break;
      }
      S->geoms.push_back(g);
    }
  }
  if (rc == 0 && !have_camera) rc = fail(B2PT_ERR_INVALID, "scene file has no CAMERA block");
  if (rc == 0) {
    for (const B2ptGeom& g : S->geoms)
      if (g.material_id < 0 || g.material_id >= (int)S->materials.size())
        rc = fail(B2PT_ERR_INVALID, "geom refers to a material that does not exist");
  }
  if (rc) {
    delete S;
    return rc;
  }
  for (const LoadedTexture& t : S->tex_store) {
    B2ptTexture x;
    memset(&x, 0, sizeof x);
    x.width = t.w;
    x.height = t.h;
    x.channels = t.c;
    x.texels = t.texels.data();
    S->textures.push_back(x);
  }
  B2ptScene& V = S->view;
  V.n_geoms = (int)S->geoms.size();
  V.n_materials = (int)S->materials.size();
  V.n_textures = (int)S->textures.size();
  V.n_faces = (int)(S->face_pos.size() / 9);
  V.geoms = S->geoms.data();
  V.materials = S->materials.data();
  V.textures = S->textures.data();
  V.face_pos = S->face_pos.data();
  V.face_uv = S->face_uv.data();
  V.camera = cam;
  V.trace_depth = trace_depth;
  V.iterations = iterations;
  V.face_material = nullptr;
  V.material_textures = nullptr;
  if (S->per_face_materials) {
    S->face_material.resize((size_t)V.n_faces, 0);
    S->material_textures.resize(4 * (size_t)V.n_materials, -1);
    V.face_material = S->face_material.data();
    V.material_textures = S->material_textures.data();
  }
  *out = S;
  return 0;
}

extern "C" const B2ptScene* b2pt_scene_view(const B2ptLoadedScene* s) { return s ? &s->view : nullptr; }
extern "C" const char* b2pt_scene_image_name(const B2ptLoadedScene* s) { return s ? s->image_name.c_str() : ""; }
extern "C" const char* b2pt_scene_warnings(const B2ptLoadedScene* s) { return s ? s->warnings.c_str() : ""; }
extern "C" void b2pt_scene_free(B2ptLoadedScene* s) { delete s; }
