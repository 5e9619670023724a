// ref_hdr_main.cpp -- TEST INFRASTRUCTURE.  Drives the reference's own image
// class (apps/src/image.cpp, compiled where it lies) the way saveImage() would
// with its commented-out last line (apps/src/main.cpp:115-135,163): reads a raw
// accumulation buffer, mirrors x, divides by the sample count and lets
// image::saveHDR (image.cpp:41-45, stbi_write_hdr) encode.
//   ref_hdr <in.raw> <w> <h> <samples> <divide 0|1> <out base name>
// in.raw = w*h*3 little-endian floats in the renderer's pixel order.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "image.h"

int main(int argc, char** argv) {
  if (argc != 7) {
    fprintf(stderr, "usage: ref_hdr in.raw w h samples divide out_base\n");
    return 2;
  }
  const int width = atoi(argv[2]), height = atoi(argv[3]);
  const float samples = (float)atoi(argv[4]);
  const int divide = atoi(argv[5]);
  std::vector<glm::vec3> buf((size_t)width * height);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(buf.data(), sizeof(glm::vec3), buf.size(), f) != buf.size()) {
    fprintf(stderr, "cannot read %s\n", argv[1]);
    return 1;
  }
  fclose(f);
  image img(width, height);
  for (int x = 0; x < width; x++) {
    for (int y = 0; y < height; y++) {
      int index = x + (y * width);
      glm::vec3 pix = buf[index];
      if (divide) img.setPixel(width - 1 - x, y, glm::vec3(pix) / samples);
      else img.setPixel(width - 1 - x, y, glm::vec3(pix));
    }
  }
  img.saveHDR(argv[6]);
  return 0;
}
