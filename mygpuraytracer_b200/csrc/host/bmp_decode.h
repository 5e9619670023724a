// bmp_decode.h -- Windows BMP reader for textures, producing what stbi_load(path, &w, &h, &c, 0) produces
// (the reference loads every texture map with stb_image, apps/src/scene.cpp:126-214; BMP rules at
// apps/src/stb_image.h:5300-5640):
//   * headers of 12 (OS/2), 40, 56, 108 and 124 bytes; 1-, 4- and 8-bit palette images, 16-, 24- and 32-bit
//     true colour, BI_RGB and BI_BITFIELDS (RLE and embedded JPEG / PNG are rejected, as stb does);
//   * three channels, or four when the format carries an alpha mask (plain 32-bit files count: their fourth
//     byte is the alpha, and if it is zero everywhere the image is made opaque);
//   * mask fields narrower than eight bits are widened by bit replication; rows are stored bottom-up unless
//     the height is negative.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace b2host {
namespace bmp {

inline int high_bit(uint32_t v) {
  int n = -1;
  while (v) { ++n; v >>= 1; }
  return n;
}
inline int bit_count(uint32_t v) {
  int n = 0;
  while (v) { n += (int)(v & 1u); v >>= 1; }
  return n;
}
// A masked field moved to the top of a byte and replicated downwards to fill it.
inline int widen(uint32_t v, int shift, int bits) {
  static const unsigned mul[9] = {0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01};
  static const unsigned shr[9] = {0, 0, 0, 1, 0, 2, 4, 6, 0};
  if (shift < 0) v <<= -shift; else v >>= shift;
  v >>= (8 - bits);
  return (int)((unsigned)v * mul[bits]) >> shr[bits];
}

inline bool decode(const uint8_t* d, size_t n, int* w_out, int* h_out, int* c_out, std::vector<uint8_t>* out, std::string* err) {
  size_t pos = 0;
  bool short_read = false;
  auto u8 = [&]() -> uint32_t { if (pos < n) return d[pos++]; short_read = true; return 0; };
  auto u16 = [&]() -> uint32_t { uint32_t a = u8(); return a | (u8() << 8); };
  auto u32 = [&]() -> uint32_t { uint32_t a = u16(); return a | (u16() << 16); };
  if (n < 26 || d[0] != 'B' || d[1] != 'M') { *err = "not a BMP file"; return false; }
  pos = 10;
  const int offset = (int)u32();
  const int hsz = (int)u32();
  if (offset < 0) { *err = "BMP: bad data offset"; return false; }
  if (hsz != 12 && hsz != 40 && hsz != 56 && hsz != 108 && hsz != 124) { *err = "BMP: unknown header size"; return false; }
  int w, hh;
  if (hsz == 12) { w = (int)u16(); hh = (int)u16(); } else { w = (int)u32(); hh = (int)u32(); }
  if (u16() != 1) { *err = "BMP: bad plane count"; return false; }
  const int bpp = (int)u16();
  uint32_t mr = 0, mg = 0, mb = 0, ma = 0;
  unsigned all_a = 255;
  int extra_read = 14;
  auto default_masks = [&]() {
    if (bpp == 16) { mr = 31u << 10; mg = 31u << 5; mb = 31u; }
    else if (bpp == 32) { mr = 0xffu << 16; mg = 0xffu << 8; mb = 0xffu; ma = 0xffu << 24; all_a = 0; }
    else mr = mg = mb = ma = 0;
  };
  if (hsz != 12) {
    const int compress = (int)u32();
    if (compress == 1 || compress == 2) { *err = "BMP: run-length encoded files are not supported"; return false; }
    if (compress >= 4 || compress < 0) { *err = "BMP: unsupported compression"; return false; }
    if (compress == 3 && bpp != 16 && bpp != 32) { *err = "BMP: bit fields need 16 or 32 bits per pixel"; return false; }
    pos += 20;  // image size, resolution, colours used / important
    if (hsz == 40 || hsz == 56) {
      if (hsz == 56) pos += 16;
      if (bpp == 16 || bpp == 32) {
        if (compress == 0) default_masks();
        else {
          mr = u32(); mg = u32(); mb = u32();
          extra_read += 12;
          if (mr == mg && mg == mb) { *err = "BMP: bad masks"; return false; }
        }
      }
    } else {
      mr = u32(); mg = u32(); mb = u32(); ma = u32();
      if (compress != 3) default_masks();
      pos += 52;
      if (hsz == 124) pos += 16;
    }
  }
  if (short_read || pos > n) { *err = "BMP: truncated header"; return false; }
  const bool bottom_up = hh > 0;
  const int h = std::abs(hh);
  if (w <= 0 || h <= 0 || (uint64_t)w * (uint64_t)h > (1ull << 28)) { *err = "BMP: bad dimensions"; return false; }
  int psize = 0;
  if (hsz == 12) { if (bpp < 24) psize = (offset - extra_read - 24) / 3; }
  else if (bpp < 16) psize = (offset - extra_read - hsz) >> 2;
  if (psize == 0 && (size_t)offset != pos) { *err = "BMP: bad data offset"; return false; }
  const int ch = (bpp == 24 && ma == 0xff000000u) ? 3 : (ma ? 4 : 3);
  // the pixel data must be present before anything of the image's size is allocated
  const uint64_t row_src = bpp < 16 ? (((uint64_t)w * bpp + 7) / 8 + 3) / 4 * 4 : ((uint64_t)w * (bpp / 8) + 3) / 4 * 4;
  if (bpp != 1 && bpp != 4 && bpp != 8 && bpp != 16 && bpp != 24 && bpp != 32) { *err = "BMP: bad bit depth"; return false; }
  if ((uint64_t)offset > n || row_src * (uint64_t)h > n - (uint64_t)offset + 3) { *err = "BMP: truncated pixel data"; return false; }
  out->assign((size_t)w * h * ch, 0);
  size_t z = 0;
  if (bpp < 16) {
    if (psize <= 0 || psize > 256) { *err = "BMP: bad palette"; return false; }
    uint8_t pal[256][3];
    memset(pal, 0, sizeof pal);
    for (int i = 0; i < psize; ++i) {
      pal[i][2] = (uint8_t)u8();
      pal[i][1] = (uint8_t)u8();
      pal[i][0] = (uint8_t)u8();
      if (hsz != 12) u8();
    }
    const long skip = (long)offset - extra_read - hsz - (long)psize * (hsz == 12 ? 3 : 4);
    if (skip < 0) { *err = "BMP: bad data offset"; return false; }
    pos += (size_t)skip;
    const int width = bpp == 1 ? (w + 7) >> 3 : bpp == 4 ? (w + 1) >> 1 : w;
    const int pad = (-width) & 3;
    for (int j = 0; j < h; ++j) {
      if (bpp == 1) {
        int bit = 7;
        uint32_t v = u8();
        for (int i = 0; i < w; ++i) {
          const int c = (int)(v >> bit) & 1;
          (*out)[z++] = pal[c][0]; (*out)[z++] = pal[c][1]; (*out)[z++] = pal[c][2];
          if (ch == 4) (*out)[z++] = 255;
          if (i + 1 == w) break;
          if (--bit < 0) { bit = 7; v = u8(); }
        }
      } else {
        for (int i = 0; i < w; i += 2) {
          uint32_t v = u8(), v2 = 0;
          if (bpp == 4) { v2 = v & 15; v >>= 4; }
          (*out)[z++] = pal[v][0]; (*out)[z++] = pal[v][1]; (*out)[z++] = pal[v][2];
          if (ch == 4) (*out)[z++] = 255;
          if (i + 1 == w) break;
          v = bpp == 8 ? u8() : v2;
          (*out)[z++] = pal[v][0]; (*out)[z++] = pal[v][1]; (*out)[z++] = pal[v][2];
          if (ch == 4) (*out)[z++] = 255;
        }
      }
      pos += (size_t)pad;
    }
  } else {
    const long skip = (long)offset - extra_read - hsz;
    if (skip < 0) { *err = "BMP: bad data offset"; return false; }
    pos += (size_t)skip;
    const int width = bpp == 24 ? 3 * w : bpp == 16 ? 2 * w : 0;
    const int pad = (-width) & 3;
    int easy = 0;
    if (bpp == 24) easy = 1;
    else if (bpp == 32 && mb == 0xffu && mg == 0xff00u && mr == 0x00ff0000u && ma == 0xff000000u) easy = 2;
    int rs = 0, gs = 0, bs = 0, as = 0, rc = 0, gc = 0, bc = 0, ac = 0;
    if (!easy) {
      if (!mr || !mg || !mb) { *err = "BMP: bad masks"; return false; }
      rs = high_bit(mr) - 7; rc = bit_count(mr);
      gs = high_bit(mg) - 7; gc = bit_count(mg);
      bs = high_bit(mb) - 7; bc = bit_count(mb);
      as = high_bit(ma) - 7; ac = bit_count(ma);
      if (rc > 8 || gc > 8 || bc > 8 || ac > 8) { *err = "BMP: bad masks"; return false; }
    }
    for (int j = 0; j < h; ++j) {
      for (int i = 0; i < w; ++i) {
        if (easy) {
          (*out)[z + 2] = (uint8_t)u8();
          (*out)[z + 1] = (uint8_t)u8();
          (*out)[z + 0] = (uint8_t)u8();
          z += 3;
          const unsigned a = easy == 2 ? u8() : 255u;
          all_a |= a;
          if (ch == 4) (*out)[z++] = (uint8_t)a;
        } else {
          const uint32_t v = bpp == 16 ? u16() : u32();
          (*out)[z++] = (uint8_t)widen(v & mr, rs, rc);
          (*out)[z++] = (uint8_t)widen(v & mg, gs, gc);
          (*out)[z++] = (uint8_t)widen(v & mb, bs, bc);
          const unsigned a = ma ? (unsigned)widen(v & ma, as, ac) : 255u;
          all_a |= a;
          if (ch == 4) (*out)[z++] = (uint8_t)a;
        }
      }
      pos += (size_t)pad;
    }
  }
  if (ch == 4 && all_a == 0)
    for (size_t i = 3; i < out->size(); i += 4) (*out)[i] = 255;
  if (bottom_up) {
    const size_t row = (size_t)w * ch;
    for (int y = 0; y < h / 2; ++y)
      for (size_t k = 0; k < row; ++k) std::swap((*out)[(size_t)y * row + k], (*out)[(size_t)(h - 1 - y) * row + k]);
  }
  *w_out = w;
  *h_out = h;
  *c_out = ch;
  return true;
}

}  // namespace bmp
}  // namespace b2host
