// png_writer.h -- minimal PNG (8-bit RGB) writer for the saveImage hand-off
// (apps/src/main.cpp:115-165 -> image::savePNG, apps/src/image.cpp:22-39, which
// calls stb_image_write).  The PIXELS are what the reference writes; the
// container uses stored (uncompressed) deflate blocks and filter type 0, so
// the file bytes differ from stb's while every decoder yields the same image.
#pragma once

#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

namespace b2pt_host {

inline uint32_t png_crc32(const uint8_t* p, size_t n, uint32_t crc = 0) {
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    ready = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
  return ~crc;
}

inline void png_put32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24));
  v.push_back((uint8_t)(x >> 16));
  v.push_back((uint8_t)(x >> 8));
  v.push_back((uint8_t)x);
}

inline void png_chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
  png_put32(out, (uint32_t)data.size());
  const size_t at = out.size();
  out.insert(out.end(), type, type + 4);
  out.insert(out.end(), data.begin(), data.end());
  png_put32(out, png_crc32(out.data() + at, out.size() - at));
}

// rgb: h rows of w*3 bytes, top row first.  Returns an empty string or an error text.
inline std::string write_png_rgb8(const char* path, int w, int h, const uint8_t* rgb) {
  if (w <= 0 || h <= 0) return "bad image size";
  std::vector<uint8_t> raw;
  raw.reserve((size_t)h * ((size_t)w * 3 + 1));
  for (int y = 0; y < h; ++y) {
    raw.push_back(0);  // filter type 0 (None)
    raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
  }
  std::vector<uint8_t> z;
  z.push_back(0x78);
  z.push_back(0x01);  // zlib header, no preset dictionary
  uint32_t a = 1, b = 0;
  size_t pos = 0;
  while (pos < raw.size()) {
    const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
    z.push_back(pos + n == raw.size() ? 1 : 0);  // BFINAL, BTYPE=00 (stored)
    z.push_back((uint8_t)(n & 0xff));
    z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xff));
    z.push_back((uint8_t)((~n >> 8) & 0xff));
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    for (size_t i = 0; i < n; ++i) {  // Adler-32
      a += raw[pos + i];
      if (a >= 65521u) a -= 65521u;
      b += a;
      if (b >= 65521u) b -= 65521u;
    }
    pos += n;
  }
  png_put32(z, (b << 16) | a);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  std::vector<uint8_t> ihdr;
  png_put32(ihdr, (uint32_t)w);
  png_put32(ihdr, (uint32_t)h);
  ihdr.push_back(8);  // bit depth
  ihdr.push_back(2);  // colour type: RGB
  ihdr.push_back(0);
  ihdr.push_back(0);
  ihdr.push_back(0);
  png_chunk(out, "IHDR", ihdr);
  png_chunk(out, "IDAT", z);
  png_chunk(out, "IEND", std::vector<uint8_t>());
  FILE* f = fopen(path, "wb");
  if (!f) return std::string("cannot open ") + path + " for writing";
  const size_t wrote = fwrite(out.data(), 1, out.size(), f);
  const int rc = fclose(f);
  if (wrote != out.size() || rc != 0) return std::string("short write to ") + path;
  return std::string();
}

}  // namespace b2pt_host
