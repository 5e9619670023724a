// ref_objmat_main.cpp -- TEST INFRASTRUCTURE (oracle/_ref/ref_objmat).
//
// What tinyobjloader 2.0.0 -- the reference's vendored copy, apps/src/tiny_obj_loader.h, called exactly as
// apps/src/scene.cpp:40-58 calls it -- reports for an OBJ file: the per-face material ids the reference reads and
// discards (scene.cpp:121-122) and the materials of the MTL file(s).  Prints one JSON object; the per-face
// material golden of tests/golden/ is made from it (tests/golden/make_multimat_golden.py).
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"

#include <cstdio>
#include <iostream>

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: ref_objmat file.obj mtl_search_path\n");
    return 2;
  }
  tinyobj::ObjReaderConfig reader_config;
  reader_config.mtl_search_path = argv[2];  // "../models/materials" in scene.cpp:41
  tinyobj::ObjReader reader;
  if (!reader.ParseFromFile(argv[1], reader_config)) {
    fprintf(stderr, "TinyObjReader: %s\n", reader.Error().c_str());
    return 1;
  }
  auto& shapes = reader.GetShapes();
  auto& mats = reader.GetMaterials();
  printf("{\"material_ids\": [");
  bool first = true;
  for (size_t s = 0; s < shapes.size(); s++)      // the loop order of scene.cpp:73-76
    for (size_t f = 0; f < shapes[s].mesh.num_face_vertices.size(); f++) {
      printf("%s%d", first ? "" : ", ", shapes[s].mesh.material_ids[f]);
      first = false;
    }
  printf("], \"materials\": [");
  for (size_t m = 0; m < mats.size(); ++m) {
    const tinyobj::material_t& t = mats[m];
    printf("%s{\"name\": \"%s\", \"kd\": [%.9g, %.9g, %.9g], \"ks\": [%.9g, %.9g, %.9g], \"ke\": [%.9g, %.9g, %.9g], \"ior\": %.9g, "
           "\"map_kd\": \"%s\", \"map_ks\": \"%s\", \"map_ke\": \"%s\", \"map_bump\": \"%s\"}",
           m ? ", " : "", t.name.c_str(), t.diffuse[0], t.diffuse[1], t.diffuse[2], t.specular[0], t.specular[1], t.specular[2],
           t.emission[0], t.emission[1], t.emission[2], t.ior, t.diffuse_texname.c_str(), t.specular_texname.c_str(),
           t.emissive_texname.c_str(), t.bump_texname.c_str());
  }
  printf("]}\n");
  return 0;
}
