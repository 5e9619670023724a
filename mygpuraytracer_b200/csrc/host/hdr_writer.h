// hdr_writer.h -- Radiance .hdr (RGBE) writer for the saveImage hand-off: image::saveHDR
// (apps/src/image.cpp:41-45, stbi_write_hdr; the call sits commented out at apps/src/main.cpp:163).
// The PIXELS are what the reference writes -- Ward's shared-exponent encoding exactly as
// stb_image_write states it, pinned against the reference's own image.cpp by
// tests/golden/hdr_golden.npz -- while the container stores flat scanlines instead of stb's
// run-length encoded ones (both are valid "32-bit_rle_rgbe" files; a flat scanline cannot be
// mistaken for an RLE one below 32768 pixels per row, because a pixel whose red and green
// mantissas are 2 has a blue mantissa >= 128 in the place of the row length's high byte).
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

namespace b2pt_host {

// One pixel: m = largest channel; below 1e-32 -> 0 0 0 0; else frexp(m) = (f, e),
// scale = float(f) * 256.0f / m in single precision, mantissas truncated, exponent byte e + 128.
inline void rgbe_from_rgb(const float* rgb, uint8_t* out) {
  const float m = rgb[0] > (rgb[1] > rgb[2] ? rgb[1] : rgb[2]) ? rgb[0] : (rgb[1] > rgb[2] ? rgb[1] : rgb[2]);
  if ((double)m < 1e-32) {
    out[0] = out[1] = out[2] = out[3] = 0;
    return;
  }
  int e = 0;
  const float f = (float)frexp((double)m, &e);
  const float scale = f * 256.0f / m;
  out[0] = (uint8_t)(int)(rgb[0] * scale);
  out[1] = (uint8_t)(int)(rgb[1] * scale);
  out[2] = (uint8_t)(int)(rgb[2] * scale);
  out[3] = (uint8_t)(e + 128);
}

// `rgb`: width*height*3 floats, row 0 first, exactly the pixels handed to image::saveHDR.
// Returns an empty string or the error text.
inline std::string write_hdr_rgb(const char* path, int width, int height, const float* rgb) {
  if (width <= 0 || height <= 0) return "write_hdr_rgb: empty image";
  FILE* f = fopen(path, "wb");
  if (!f) return std::string("cannot open ") + path + " for writing";
  fprintf(f, "#?RADIANCE\n# Written by b2pt (pixels of stb_image_write's stbi_write_hdr)\nFORMAT=32-bit_rle_rgbe\n");
  fprintf(f, "EXPOSURE=          1.0000000000000\n\n-Y %d +X %d\n", height, width);
  std::vector<uint8_t> row((size_t)width * 4);
  bool ok = true;
  for (int y = 0; y < height && ok; ++y) {
    for (int x = 0; x < width; ++x) rgbe_from_rgb(rgb + ((size_t)y * width + x) * 3, &row[(size_t)x * 4]);
    ok = fwrite(row.data(), 1, row.size(), f) == row.size();
  }
  if (fclose(f) != 0) ok = false;
  return ok ? std::string() : std::string("short write to ") + path;
}

}  // namespace b2pt_host
