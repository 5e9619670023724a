// png_decode.h -- PNG reader for textures, producing what stbi_load(path, &w, &h, &c, 0) produces
// (the reference loads every texture map with stb_image, apps/src/scene.cpp:126-214):
//   * channels: grey 1, grey+alpha 2, RGB 3, RGBA 4, palette 3 (4 with a tRNS chunk); a tRNS chunk on a
//     grey / RGB image adds an alpha channel that is 0 for the transparent colour and 255 elsewhere;
//   * 16-bit samples are reduced to their high byte, 1/2/4-bit grey samples are scaled by 255/85/17,
//     palette indices are looked up unscaled;
//   * chunk CRCs and the Adler-32 are not verified (stb does not verify them either).
// Non-interlaced files only (Adam7 is rejected with an error).  Written from the PNG / zlib / DEFLATE
// specifications (RFC 2083, 1950, 1951); pinned against the reference's loader by tests/golden/png.
#pragma once

#include <cstdint>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace b2host {
namespace png {

// ---- DEFLATE (RFC 1951) ------------------------------------------------------------------------
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  bool fail = false;
  uint32_t get(int k) {  // k <= 16, least significant bit first
    while (n < k) {
      if (p >= end) {
        fail = true;
        return 0;
      }
      acc |= (uint32_t)*p++ << n;
      n += 8;
    }
    const uint32_t v = acc & ((1u << k) - 1u);
    acc >>= k;
    n -= k;
    return v;
  }
  uint32_t peek(int k) {  // up to k <= 16 bits without consuming them; fewer are valid at the very end of the data
    while (n < k && p < end) {
      acc |= (uint32_t)*p++ << n;
      n += 8;
    }
    return acc & ((1u << k) - 1u);
  }
  void drop(int k) {
    if (k > n) {
      fail = true;
      acc = 0;
      n = 0;
    } else {
      acc >>= k;
      n -= k;
    }
  }
  void to_byte_boundary() {  // drop the rest of the current byte, give whole buffered bytes back
    p -= n / 8;
    acc = 0;
    n = 0;
  }
};

constexpr int kFastBits = 9;
struct Huffman {  // canonical code: how many codes of each length, symbols ordered by (length, value)
  uint16_t count[16];
  uint16_t symbol[288];
  uint16_t fast[1 << kFastBits];  // codes of up to kFastBits bits, indexed by the next bits of the stream: (length << 9) | symbol
  bool build(const uint8_t* len, int n) {
    memset(count, 0, sizeof count);
    memset(fast, 0, sizeof fast);
    for (int i = 0; i < n; ++i) ++count[len[i]];
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < n; ++i)
      if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    int left = 1;  // over-subscribed codes are invalid; incomplete ones are tolerated (single-code trees)
    for (int l = 1; l <= 15; ++l) {
      left = (left << 1) - count[l];
      if (left < 0) return false;
    }
    count[0] = 0;
    // the short codes once more, as a table over the next kFastBits bits of the stream (which delivers a code's
    // bits most significant first, so the index is the code reversed, with every combination of the bits behind it)
    int next_code[16];
    next_code[0] = 0;
    for (int l = 1, code = 0; l <= 15; ++l) {
      code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
      next_code[l] = code;
    }
    for (int i = 0; i < n; ++i) {
      const int l = len[i];
      if (!l) continue;
      const int code = next_code[l]++;
      if (l > kFastBits) continue;
      int rev = 0;
      for (int k = 0; k < l; ++k) rev |= ((code >> k) & 1) << (l - 1 - k);
      for (int hi = 0; hi < (1 << (kFastBits - l)); ++hi) fast[rev | (hi << l)] = (uint16_t)((l << 9) | i);
    }
    return true;
  }
  int decode(BitReader& b) const {
    const uint16_t e = fast[b.peek(kFastBits)];
    if (e) {
      b.drop(e >> 9);
      return b.fail ? -1 : (e & 511);
    }
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
      code |= (int)b.get(1);
      if (b.fail) return -1;
      const int c = count[l];
      if (code - c < first) return symbol[index + (code - first)];
      index += c;
      first = (first + c) << 1;
      code <<= 1;
    }
    return -1;
  }
};

// `cap`: the number of bytes the caller can use (the scanlines of the image); the stream is decoded no further, so
// a small file cannot make the decoder allocate without bound.
inline bool inflate(const uint8_t* src, size_t n, size_t cap, std::vector<uint8_t>* out, std::string* err) {
  static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  static const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  if (n < 2) { *err = "PNG: zlib stream too short"; return false; }
  if ((src[0] & 15) != 8 || (src[1] & 0x20)) { *err = "PNG: unsupported zlib header"; return false; }
  BitReader b{src + 2, src + n};
  Huffman lit, dist;
  bool last = false;
  while (!last && out->size() < cap) {
    last = b.get(1) != 0;
    const uint32_t type = b.get(2);
    if (b.fail) { *err = "PNG: truncated deflate stream"; return false; }
    if (type == 0) {
      b.to_byte_boundary();
      if (b.end - b.p < 4) { *err = "PNG: truncated stored block"; return false; }
      const uint32_t len = b.p[0] | (b.p[1] << 8), nlen = b.p[2] | (b.p[3] << 8);
      b.p += 4;
      if ((len ^ 0xffffu) != nlen || (size_t)(b.end - b.p) < len) { *err = "PNG: corrupt stored block"; return false; }
      out->insert(out->end(), b.p, b.p + std::min<size_t>(len, cap - out->size()));
      b.p += len;
      continue;
    }
    if (type == 3) { *err = "PNG: invalid deflate block type"; return false; }
    uint8_t lens[320];
    if (type == 1) {
      for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      lit.build(lens, 288);
      for (int i = 0; i < 30; ++i) lens[i] = 5;
      dist.build(lens, 30);
    } else {
      const int hlit = (int)b.get(5) + 257, hdist = (int)b.get(5) + 1, hclen = (int)b.get(4) + 4;
      if (hlit > 286 || hdist > 30) { *err = "PNG: bad deflate code counts"; return false; }
      uint8_t cl[19];
      memset(cl, 0, sizeof cl);
      for (int i = 0; i < hclen; ++i) cl[cl_order[i]] = (uint8_t)b.get(3);
      Huffman clh;
      if (b.fail || !clh.build(cl, 19)) { *err = "PNG: bad code-length code"; return false; }
      int i = 0;
      while (i < hlit + hdist) {
        const int s = clh.decode(b);
        if (s < 0) { *err = "PNG: bad code lengths"; return false; }
        if (s < 16) {
          lens[i++] = (uint8_t)s;
          continue;
        }
        int rep, val = 0;
        if (s == 16) {
          if (i == 0) { *err = "PNG: repeat without a previous length"; return false; }
          val = lens[i - 1];
          rep = 3 + (int)b.get(2);
        } else if (s == 17) {
          rep = 3 + (int)b.get(3);
        } else {
          rep = 11 + (int)b.get(7);
        }
        if (i + rep > hlit + hdist) { *err = "PNG: code lengths overflow"; return false; }
        while (rep--) lens[i++] = (uint8_t)val;
      }
      if (b.fail || !lit.build(lens, hlit) || !dist.build(lens + hlit, hdist)) { *err = "PNG: bad Huffman tables"; return false; }
    }
    for (;;) {
      const int s = lit.decode(b);
      if (s < 0) { *err = "PNG: bad literal/length code"; return false; }
      if (out->size() >= cap) return true;  // everything the image needs has been decoded
      if (s < 256) {
        out->push_back((uint8_t)s);
        continue;
      }
      if (s == 256) break;
      if (s > 285) { *err = "PNG: bad length symbol"; return false; }
      const int len = len_base[s - 257] + (int)b.get(len_extra[s - 257]);
      const int ds = dist.decode(b);
      if (ds < 0 || ds > 29) { *err = "PNG: bad distance code"; return false; }
      const size_t d = (size_t)dist_base[ds] + b.get(dist_extra[ds]);
      if (b.fail || d > out->size()) { *err = "PNG: distance reaches before the start"; return false; }
      size_t from = out->size() - d;
      for (int k = 0; k < len && out->size() < cap; ++k) out->push_back((*out)[from++]);
    }
  }
  return true;
}

// ---- PNG (RFC 2083) -------------------------------------------------------------------------------
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline bool decode(const uint8_t* data, size_t size, int* w_out, int* h_out, int* c_out, std::vector<uint8_t>* out, std::string* err) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  if (size < 8 || memcmp(data, sig, 8) != 0) { *err = "not a PNG file"; return false; }
  size_t pos = 8;
  int w = 0, h = 0, depth = 0, ctype = 0;
  bool have_header = false, have_trns = false;
  uint8_t palette[256 * 4];
  int pal_len = 0;
  uint16_t trns_colour[3] = {0, 0, 0};
  memset(palette, 0, sizeof palette);
  for (int i = 0; i < 256; ++i) palette[4 * i + 3] = 255;
  std::vector<uint8_t> z;
  while (pos + 12 <= size) {
    const uint32_t len = be32(data + pos);
    const uint8_t* type = data + pos + 4;
    const uint8_t* body = data + pos + 8;
    if (len > size - pos - 12) { *err = "PNG: truncated chunk"; return false; }
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) { *err = "PNG: bad IHDR"; return false; }
      w = (int)be32(body);
      h = (int)be32(body + 4);
      depth = body[8];
      ctype = body[9];
      if (body[12] != 0) { *err = "PNG: interlaced files are not supported"; return false; }
      if (w <= 0 || h <= 0 || body[10] != 0 || body[11] != 0) { *err = "PNG: bad IHDR"; return false; }
      if (w > (1 << 24) || h > (1 << 24) || (uint64_t)w * (uint64_t)h > (1ull << 28)) { *err = "PNG: image too large"; return false; }
      const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                            (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                            ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
      if (!depth_ok) { *err = "PNG: unsupported colour type / bit depth"; return false; }
      have_header = true;
    } else if (!memcmp(type, "PLTE", 4)) {
      pal_len = (int)(len / 3);
      if (pal_len > 256 || pal_len * 3 != (int)len) { *err = "PNG: bad PLTE"; return false; }
      for (int i = 0; i < pal_len; ++i) memcpy(palette + 4 * i, body + 3 * i, 3);
    } else if (!memcmp(type, "tRNS", 4)) {
      if (!have_header) { *err = "PNG: tRNS before IHDR"; return false; }
      if (ctype == 3) {
        if ((int)len > pal_len) { *err = "PNG: bad tRNS"; return false; }
        for (uint32_t i = 0; i < len; ++i) palette[4 * i + 3] = body[i];
      } else if (ctype == 0 || ctype == 2) {
        const int nc = ctype == 0 ? 1 : 3;
        if ((int)len != 2 * nc) { *err = "PNG: bad tRNS"; return false; }
        for (int k = 0; k < nc; ++k) trns_colour[k] = (uint16_t)((body[2 * k] << 8) | body[2 * k + 1]);
      } else {
        *err = "PNG: tRNS with an alpha colour type";
        return false;
      }
      have_trns = true;
    } else if (!memcmp(type, "IDAT", 4)) {
      z.insert(z.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (!have_header || z.empty()) { *err = "PNG: no image data"; return false; }
  if (ctype == 3 && pal_len == 0) { *err = "PNG: palette image without PLTE"; return false; }
  const int file_ch = ctype == 3 ? 1 : ((ctype & 2) ? 3 : 1) + ((ctype & 4) ? 1 : 0);
  const size_t row_bytes = ((size_t)w * file_ch * depth + 7) / 8;
  const size_t bpp = (size_t)(file_ch * depth + 7) / 8;  // filter distance, at least one byte
  std::vector<uint8_t> raw;
  if (!inflate(z.data(), z.size(), (row_bytes + 1) * (size_t)h, &raw, err)) return false;
  if (raw.size() < (row_bytes + 1) * (size_t)h) { *err = "PNG: not enough image data"; return false; }
  // ---- undo the scanline filters in place ----
  std::vector<uint8_t> zero(row_bytes, 0);
  for (int y = 0; y < h; ++y) {
    uint8_t* cur = raw.data() + (size_t)y * (row_bytes + 1) + 1;
    const uint8_t* up = y ? cur - (row_bytes + 1) : zero.data();
    const int ft = cur[-1];
    if (ft > 4) { *err = "PNG: bad filter type"; return false; }
    for (size_t i = 0; i < row_bytes; ++i) {
      const int a = i >= bpp ? cur[i - bpp] : 0, bb = up[i], c = i >= bpp ? up[i - bpp] : 0;
      int pred = 0;
      if (ft == 1) pred = a;
      else if (ft == 2) pred = bb;
      else if (ft == 3) pred = (a + bb) >> 1;
      else if (ft == 4) {
        const int p = a + bb - c, pa = p > a ? p - a : a - p, pb = p > bb ? p - bb : bb - p, pc = p > c ? p - c : c - p;
        pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c);
      }
      cur[i] = (uint8_t)(cur[i] + pred);
    }
  }
  // ---- samples -> 8-bit output with stb_image's channel rules ----
  const bool add_alpha = have_trns && (ctype == 0 || ctype == 2);
  const int out_ch = ctype == 3 ? (have_trns ? 4 : 3) : file_ch + (add_alpha ? 1 : 0);
  out->assign((size_t)w * h * out_ch, 0);
  static const int grey_scale[9] = {0, 255, 85, 0, 17, 0, 0, 0, 1};
  for (int y = 0; y < h; ++y) {
    const uint8_t* cur = raw.data() + (size_t)y * (row_bytes + 1) + 1;
    uint8_t* o = out->data() + (size_t)y * w * out_ch;
    for (int x = 0; x < w; ++x) {
      if (ctype == 3) {
        int idx;
        if (depth == 8) idx = cur[x];
        else {
          const int per = 8 / depth, shift = (per - 1 - x % per) * depth;
          idx = (cur[x / per] >> shift) & ((1 << depth) - 1);
        }
        const uint8_t* pe = palette + 4 * idx;
        o[0] = pe[0];
        o[1] = pe[1];
        o[2] = pe[2];
        if (out_ch == 4) o[3] = pe[3];
        o += out_ch;
        continue;
      }
      bool transparent = add_alpha;
      for (int k = 0; k < file_ch; ++k) {
        int v8;
        uint16_t raw_v;
        if (depth == 16) {
          raw_v = (uint16_t)((cur[2 * (x * file_ch + k)] << 8) | cur[2 * (x * file_ch + k) + 1]);
          v8 = raw_v >> 8;
        } else if (depth == 8) {
          raw_v = cur[x * file_ch + k];
          v8 = raw_v;
        } else {  // 1, 2 or 4 bit grey
          const int per = 8 / depth, shift = (per - 1 - x % per) * depth;
          raw_v = (uint16_t)((cur[x / per] >> shift) & ((1 << depth) - 1));
          v8 = raw_v * grey_scale[depth];
        }
        if (add_alpha && k < 3 && raw_v != (depth == 16 ? trns_colour[k] : (uint16_t)(trns_colour[k] & 255))) transparent = false;
        o[k] = (uint8_t)v8;
      }
      if (add_alpha) o[file_ch] = transparent ? 0 : 255;
      o += out_ch;
    }
  }
  *w_out = w;
  *h_out = h;
  *c_out = out_ch;
  return true;
}

}  // namespace png
}  // namespace b2host
