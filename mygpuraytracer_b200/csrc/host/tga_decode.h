// tga_decode.h -- Truevision TGA reader for textures, producing what stbi_load(path, &w, &h, &c, 0) produces
// (the reference loads every texture map with stb_image, apps/src/scene.cpp:126-214; TGA is the format stb tries
// last because it has no signature, apps/src/stb_image.h:5700-5940):
//   * image types 1 / 2 / 3 (colour-mapped, true colour, grey) and their run-length encoded forms 9 / 10 / 11;
//   * channels = bits / 8 of the pixel (of the palette entry for colour-mapped files); 15- and 16-bit pixels
//     are RGB555 expanded with (v * 255) / 31 to three channels, 16-bit grey is grey + alpha;
//   * rows are stored bottom-up unless bit 5 of the descriptor says top-down; BGR(A) becomes RGB(A).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace b2host {
namespace tga {

inline int channels_of(int bits, bool grey, bool* rgb16) {
  *rgb16 = false;
  switch (bits) {
    case 8: return 1;
    case 16:
      if (grey) return 2;
      *rgb16 = true;
      return 3;
    case 15: *rgb16 = true; return 3;
    case 24: return 3;
    case 32: return 4;
    default: return 0;
  }
}

// The header test stb applies before it accepts a file as TGA.
inline bool looks_like_tga(const uint8_t* d, size_t n) {
  if (n < 18) return false;
  const int cmap = d[1], type = d[2], bpp = d[16];
  if (cmap > 1) return false;
  if (cmap == 1) {
    if (type != 1 && type != 9) return false;
    const int pb = d[7];
    if (pb != 8 && pb != 15 && pb != 16 && pb != 24 && pb != 32) return false;
  } else if (type != 2 && type != 3 && type != 10 && type != 11) {
    return false;
  }
  if ((d[12] | (d[13] << 8)) < 1 || (d[14] | (d[15] << 8)) < 1) return false;
  if (cmap == 1 && bpp != 8 && bpp != 16) return false;
  return bpp == 8 || bpp == 15 || bpp == 16 || bpp == 24 || bpp == 32;
}

inline bool decode(const uint8_t* d, size_t n, int* w_out, int* h_out, int* c_out, std::vector<uint8_t>* out, std::string* err) {
  if (!looks_like_tga(d, n)) { *err = "not a TGA file"; return false; }
  const int id_len = d[0], indexed = d[1];
  int type = d[2];
  const bool rle = type >= 8;
  if (rle) type -= 8;
  const int pal_start = d[3] | (d[4] << 8), pal_len = d[5] | (d[6] << 8), pal_bits = d[7];
  const int w = d[12] | (d[13] << 8), h = d[14] | (d[15] << 8), bpp = d[16];
  const bool bottom_up = ((d[17] >> 5) & 1) == 0;
  bool rgb16 = false;
  const int ch = indexed ? channels_of(pal_bits, false, &rgb16) : channels_of(bpp, type == 3, &rgb16);
  if (!ch) { *err = "TGA: unsupported pixel format"; return false; }
  size_t pos = 18 + (size_t)id_len;
  auto need = [&](size_t k) { return pos + k <= n; };
  auto read_rgb555 = [&](uint8_t* o) {
    const int px = d[pos] | (d[pos + 1] << 8);
    pos += 2;
    o[0] = (uint8_t)((((px >> 10) & 31) * 255) / 31);
    o[1] = (uint8_t)((((px >> 5) & 31) * 255) / 31);
    o[2] = (uint8_t)(((px & 31) * 255) / 31);
  };
  std::vector<uint8_t> palette;
  if (indexed) {
    if (pal_len == 0) { *err = "TGA: empty palette"; return false; }
    pos += (size_t)pal_start;
    palette.assign((size_t)pal_len * ch, 0);
    if (rgb16) {
      if (!need((size_t)pal_len * 2)) { *err = "TGA: truncated palette"; return false; }
      for (int i = 0; i < pal_len; ++i) read_rgb555(&palette[(size_t)i * ch]);
    } else {
      if (!need(palette.size())) { *err = "TGA: truncated palette"; return false; }
      memcpy(palette.data(), d + pos, palette.size());
      pos += palette.size();
    }
  }
  {
    // the pixel data must be able to cover the image before anything of that size is allocated: a raw pixel takes
    // bpp / 8 bytes, a run-length packet at least 2 bytes per 128 pixels
    const uint64_t pixels = (uint64_t)w * (uint64_t)h, left = pos <= n ? n - pos : 0;
    const uint64_t src_bytes = (uint64_t)((bpp + 7) / 8);
    if (pixels > (1ull << 28) || (!rle && pixels * src_bytes > left) || (rle && pixels > left * 128)) {
      *err = "TGA: truncated pixel data";
      return false;
    }
  }
  out->assign((size_t)w * h * ch, 0);
  uint8_t px[4] = {0, 0, 0, 0};
  int run = 0;
  bool repeating = false, fetch = true;
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    if (rle) {
      if (run == 0) {
        if (!need(1)) { *err = "TGA: truncated pixel data"; return false; }
        const int cmd = d[pos++];
        run = 1 + (cmd & 127);
        repeating = (cmd >> 7) != 0;
        fetch = true;
      } else if (!repeating) {
        fetch = true;
      }
    } else {
      fetch = true;
    }
    if (fetch) {
      if (indexed) {
        const size_t k = bpp == 8 ? 1 : 2;
        if (!need(k)) { *err = "TGA: truncated pixel data"; return false; }
        int idx = k == 1 ? d[pos] : (d[pos] | (d[pos + 1] << 8));
        pos += k;
        if (idx >= pal_len) idx = 0;
        for (int j = 0; j < ch; ++j) px[j] = palette[(size_t)idx * ch + j];
      } else if (rgb16) {
        if (!need(2)) { *err = "TGA: truncated pixel data"; return false; }
        read_rgb555(px);
      } else {
        if (!need((size_t)ch)) { *err = "TGA: truncated pixel data"; return false; }
        for (int j = 0; j < ch; ++j) px[j] = d[pos + j];
        pos += (size_t)ch;
      }
      fetch = false;
    }
    for (int j = 0; j < ch; ++j) (*out)[i * ch + j] = px[j];
    --run;
  }
  if (bottom_up) {
    const size_t row = (size_t)w * ch;
    for (int y = 0; y * 2 < h; ++y)
      for (size_t k = 0; k < row; ++k) std::swap((*out)[(size_t)y * row + k], (*out)[(size_t)(h - 1 - y) * row + k]);
  }
  if (ch >= 3 && !rgb16)
    for (size_t i = 0; i < (size_t)w * h; ++i) std::swap((*out)[i * ch], (*out)[i * ch + 2]);
  *w_out = w;
  *h_out = h;
  *c_out = ch;
  return true;
}

}  // namespace tga
}  // namespace b2host
