// image_decode.h -- 8-bit texture decoding for the scene loader.
//
// The reference decodes textures with stb_image (apps/src/scene.cpp:133-218,
// stbi_load with the native channel count and a vertical flip).  stb_image is
// third-party code vendored in the reference tree; this is an independent
// baseline-JPEG (ITU T.81 sequential DCT, Huffman, 8 bit) and binary PPM reader
// that reproduces the two numeric choices which decide the texel bytes:
//   * the integer inverse DCT "derived from jidctint -- DCT_ISLOW" with 12-bit
//     constants, +512 / >>10 column pass and +65536+(128<<17) / >>17 row pass
//     (apps/src/stb_image.h:2392-2489);
//   * the 20-bit fixed-point YCbCr->RGB conversion with the 0xffff0000 mask on
//     the Cb term of green (apps/src/stb_image.h:3598-3623).
//   * the "JFIF-centred" chroma upsampling (tent filters for 2x1, 1x2 and 2x2, replication otherwise,
//     apps/src/stb_image.h:3388-3468, 3836-3880).
// The output is byte-identical to stbi_load for baseline files -- the seven shipped 4096x4096
// textures (4:4:4) and the 4:2:0 / 4:2:2 / grey / restart-interval files of
// tests/golden/jpeg (tests/test_loader.py checks both against texels dumped by the reference
// loader).  Progressive files (spectral selection and successive approximation, T.81 Annex G) are decoded
// into a coefficient store first and transformed when the last scan is in, like stb_image does.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bmp_decode.h"
#include "png_decode.h"
#include "tga_decode.h"

namespace b2host {

namespace jpg {

struct Huff {
  // canonical code tables: for each length 1..16, first code, first symbol index
  int mincode[17], maxcode[18], valptr[17];
  uint8_t vals[256];
  uint16_t fast[512];  // 9-bit lookahead: (len << 8) | symbol, 0 = slow path
  bool ok = false;
  bool build(const uint8_t counts[16], const uint8_t* symbols, int n) {
    ok = false;
    if (n < 0 || n > 256) return false;
    memcpy(vals, symbols, (size_t)n);
    int code = 0, k = 0;
    memset(fast, 0, sizeof fast);
    for (int len = 1; len <= 16; ++len) {
      valptr[len] = k;
      mincode[len] = code;
      if (code + counts[len - 1] > (1 << len)) return false;  // more codes of this length than the prefix code has room for
      for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
        if (len <= 9) {
          const int base = code << (9 - len);
          for (int j = 0; j < (1 << (9 - len)); ++j) fast[base + j] = (uint16_t)((len << 8) | vals[k]);
        }
      }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    ok = true;
    return true;
  }
};

struct Component {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int dc_pred = 0;
  int w_blocks = 0, h_blocks = 0;  // allocated size in 8x8 blocks
  std::vector<uint8_t> plane;      // w_blocks*8 x h_blocks*8 samples
  std::vector<short> coef;         // progressive files: 64 coefficients per block, w_blocks x h_blocks blocks
};

struct Decoder {
  const uint8_t* p = nullptr;
  const uint8_t* end = nullptr;
  uint32_t bitbuf = 0;
  int bitcnt = 0;
  bool hit_marker = false;
  int width = 0, height = 0, ncomp = 0;
  int hmax = 1, vmax = 1;
  uint16_t quant[4][64];
  Huff dc[4], ac[4];
  Component comp[4];
  int restart_interval = 0;
  bool progressive = false;
  int ss = 0, se = 63, ah = 0, al = 0, eobrun = 0;  // spectral selection / successive approximation of the scan

  int byte() { return p < end ? *p++ : 0; }
  int word() { int a = byte(); return (a << 8) | byte(); }

  void fill() {
    while (bitcnt <= 24) {
      int b = 0;
      if (!hit_marker && p < end) {
        b = *p++;
        if (b == 0xff) {
          int c = p < end ? *p : 0;
          if (c == 0) {
            ++p;
          } else {  // a marker: stop feeding, pad with zeros
            --p;
            hit_marker = true;
            b = 0;
          }
        }
      }
      bitbuf |= (uint32_t)b << (24 - bitcnt);
      bitcnt += 8;
    }
  }
  int bits(int n) {
    if (n == 0) return 0;
    if (bitcnt < n) fill();
    const int v = (int)(bitbuf >> (32 - n));
    bitbuf <<= n;
    bitcnt -= n;
    return v;
  }
  int decode(const Huff& h) {
    if (bitcnt < 16) fill();
    const uint16_t f = h.fast[bitbuf >> 23];
    if (f) {
      const int len = f >> 8;
      bitbuf <<= len;
      bitcnt -= len;
      return f & 255;
    }
    int code = (int)(bitbuf >> 23);
    int len = 10;
    code = (int)(bitbuf >> (32 - len));
    while (len <= 16 && code > h.maxcode[len]) {
      ++len;
      code = (int)(bitbuf >> (32 - len));
    }
    if (len > 16) return -1;
    bitbuf <<= len;
    bitcnt -= len;
    return h.vals[h.valptr[len] + code - h.mincode[len]];
  }
  // T.81 F.2.2.1 EXTEND
  int receive_extend(int s) {
    if (s == 0) return 0;
    const int v = bits(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
  }
};

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline uint8_t clamp_u8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

// One 8-point pass of the islow inverse DCT with 12-bit constants.  in[]: the
// eight inputs; even part -> e0..e3, odd part -> o0..o3, all scaled by 4096.
struct Idct1D {
  int e0, e1, e2, e3, o0, o1, o2, o3;
};
inline Idct1D idct_1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
  // fixed-point constants: (int)(c * 4096 + 0.5)
  const int C0_541 = 2217, Cm1_847 = -7567, C0_765 = 3135, C1_175 = 4816, C0_298 = 1223, C2_053 = 8410;
  const int C3_072 = 12586, C1_501 = 6149, Cm0_899 = -3685, Cm2_562 = -10497, Cm1_961 = -8034, Cm0_390 = -1597;
  Idct1D r;
  int z = (s2 + s6) * C0_541;
  const int a2 = z + s6 * Cm1_847;
  const int a3 = z + s2 * C0_765;
  const int a0 = (s0 + s4) * 4096;
  const int a1 = (s0 - s4) * 4096;
  r.e0 = a0 + a3;
  r.e3 = a0 - a3;
  r.e1 = a1 + a2;
  r.e2 = a1 - a2;
  int b0 = s7, b1 = s5, b2 = s3, b3 = s1;
  int q3 = b0 + b2, q4 = b1 + b3, q1 = b0 + b3, q2 = b1 + b2;
  const int q5 = (q3 + q4) * C1_175;
  b0 = b0 * C0_298;
  b1 = b1 * C2_053;
  b2 = b2 * C3_072;
  b3 = b3 * C1_501;
  q1 = q5 + q1 * Cm0_899;
  q2 = q5 + q2 * Cm2_562;
  q3 = q3 * Cm1_961;
  q4 = q4 * Cm0_390;
  r.o3 = b3 + (q1 + q4);
  r.o2 = b2 + (q2 + q3);
  r.o1 = b1 + (q2 + q4);
  r.o0 = b0 + (q1 + q3);
  return r;
}

inline void idct_block(uint8_t* out, int stride, const short d[64]) {
  int tmp[64];
  for (int c = 0; c < 8; ++c) {
    if (d[8 + c] == 0 && d[16 + c] == 0 && d[24 + c] == 0 && d[32 + c] == 0 && d[40 + c] == 0 && d[48 + c] == 0 &&
        d[56 + c] == 0) {
      const int dcterm = d[c] * 4;
      for (int r = 0; r < 8; ++r) tmp[r * 8 + c] = dcterm;
      continue;
    }
    Idct1D k = idct_1d(d[c], d[8 + c], d[16 + c], d[24 + c], d[32 + c], d[40 + c], d[48 + c], d[56 + c]);
    k.e0 += 512; k.e1 += 512; k.e2 += 512; k.e3 += 512;
    tmp[0 * 8 + c] = (k.e0 + k.o3) >> 10;
    tmp[7 * 8 + c] = (k.e0 - k.o3) >> 10;
    tmp[1 * 8 + c] = (k.e1 + k.o2) >> 10;
    tmp[6 * 8 + c] = (k.e1 - k.o2) >> 10;
    tmp[2 * 8 + c] = (k.e2 + k.o1) >> 10;
    tmp[5 * 8 + c] = (k.e2 - k.o1) >> 10;
    tmp[3 * 8 + c] = (k.e3 + k.o0) >> 10;
    tmp[4 * 8 + c] = (k.e3 - k.o0) >> 10;
  }
  for (int r = 0; r < 8; ++r) {
    const int* v = tmp + r * 8;
    uint8_t* o = out + (size_t)r * stride;
    Idct1D k = idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    const int bias = 65536 + (128 << 17);
    k.e0 += bias; k.e1 += bias; k.e2 += bias; k.e3 += bias;
    o[0] = clamp_u8((k.e0 + k.o3) >> 17);
    o[7] = clamp_u8((k.e0 - k.o3) >> 17);
    o[1] = clamp_u8((k.e1 + k.o2) >> 17);
    o[6] = clamp_u8((k.e1 - k.o2) >> 17);
    o[2] = clamp_u8((k.e2 + k.o1) >> 17);
    o[5] = clamp_u8((k.e2 - k.o1) >> 17);
    o[3] = clamp_u8((k.e3 + k.o0) >> 17);
    o[4] = clamp_u8((k.e3 - k.o0) >> 17);
  }
}

inline bool decode_block(Decoder& D, Component& c, short blk[64]) {
  memset(blk, 0, 64 * sizeof(short));
  const int t = D.decode(D.dc[c.td]);
  if (t < 0 || t > 15) return false;
  const int diff = t ? D.receive_extend(t) : 0;
  c.dc_pred += diff;
  blk[0] = (short)(c.dc_pred * D.quant[c.tq][0]);
  for (int k = 1; k < 64;) {
    const int rs = D.decode(D.ac[c.ta]);
    if (rs < 0) return false;
    const int s = rs & 15, r = rs >> 4;
    if (s == 0) {
      if (rs != 0xf0) break;  // end of block
      k += 16;
    } else {
      k += r;
      if (k > 63) return false;
      const int zig = kZigzag[k];
      blk[zig] = (short)(D.receive_extend(s) * D.quant[c.tq][zig]);
      ++k;
    }
  }
  return true;
}

// Progressive scans (T.81 Annex G): a DC scan sets or refines coefficient 0 of every block of its components,
// an AC scan sets or refines the band ss..se of one component; end-of-band runs span blocks.
inline bool prog_block_dc(Decoder& D, Component& c, short* blk) {
  if (D.ah == 0) {
    const int t = D.decode(D.dc[c.td]);
    if (t < 0 || t > 15) return false;
    const int diff = t ? D.receive_extend(t) : 0;
    c.dc_pred += diff;
    blk[0] = (short)(c.dc_pred * (1 << D.al));
  } else if (D.bits(1)) {
    blk[0] = (short)(blk[0] + (1 << D.al));
  }
  return true;
}

inline bool prog_block_ac(Decoder& D, Component& c, short* blk) {
  const Huff& h = D.ac[c.ta];
  if (D.ah == 0) {
    if (D.eobrun) {
      --D.eobrun;
      return true;
    }
    for (int k = D.ss; k <= D.se;) {
      const int rs = D.decode(h);
      if (rs < 0) return false;
      const int s = rs & 15, r = rs >> 4;
      if (s == 0) {
        if (r < 15) {
          D.eobrun = 1 << r;
          if (r) D.eobrun += D.bits(r);
          --D.eobrun;
          break;
        }
        k += 16;
      } else {
        k += r;
        if (k > 63) return false;
        blk[kZigzag[k++]] = (short)(D.receive_extend(s) * (1 << D.al));
      }
    }
    return true;
  }
  const short bit = (short)(1 << D.al);
  auto refine = [&](short* p) {
    if (D.bits(1) && (*p & bit) == 0) *p = (short)(*p > 0 ? *p + bit : *p - bit);
  };
  if (D.eobrun) {
    --D.eobrun;
    for (int k = D.ss; k <= D.se; ++k) {
      short* p = &blk[kZigzag[k]];
      if (*p != 0) refine(p);
    }
    return true;
  }
  int k = D.ss;
  do {
    const int rs = D.decode(h);
    if (rs < 0) return false;
    int s = rs & 15, r = rs >> 4;
    if (s == 0) {
      if (r < 15) {
        D.eobrun = (1 << r) - 1;
        if (r) D.eobrun += D.bits(r);
        r = 64;  // the rest of the band only gets refinement bits
      }
      // r == 15: sixteen zero coefficients, handled by the run below
    } else {
      if (s != 1) return false;
      s = D.bits(1) ? bit : -bit;
    }
    while (k <= D.se) {
      short* p = &blk[kZigzag[k++]];
      if (*p != 0) {
        refine(p);
      } else {
        if (r == 0) {
          *p = (short)s;
          break;
        }
        --r;
      }
    }
  } while (k <= D.se);
  return true;
}

inline bool decode(const uint8_t* data, size_t size, int* w, int* h, int* channels, std::vector<uint8_t>* out, std::string* err) {
  Decoder D;
  D.p = data;
  D.end = data + size;
  memset(D.quant, 0, sizeof D.quant);
  if (size < 4 || D.byte() != 0xff || D.byte() != 0xd8) { *err = "not a JPEG file"; return false; }
  bool have_frame = false, done = false;
  int scan_comps[4], n_scan = 0;
  while (!done && D.p < D.end) {
    int m = D.byte();
    if (m != 0xff) continue;
    while (m == 0xff) m = D.byte();
    if (m == 0xd9) break;
    if (m == 0x00) continue;  // a stuffed 0xff00 left over from entropy-coded data
    if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
    int len = D.word();
    if (len < 2) { *err = "bad JPEG segment length"; return false; }
    const uint8_t* seg_end = D.p + len - 2;
    if (seg_end > D.end) { *err = "truncated JPEG segment"; return false; }
    if (m == 0xdb) {  // DQT
      while (D.p < seg_end) {
        const int pq = D.byte();
        const int prec = pq >> 4, id = pq & 15;
        if (id > 3) { *err = "bad DQT"; return false; }
        for (int i = 0; i < 64; ++i) D.quant[id][kZigzag[i]] = (uint16_t)(prec ? D.word() : D.byte());
      }
    } else if (m == 0xc4) {  // DHT
      while (D.p < seg_end) {
        const int tc = D.byte();
        uint8_t counts[16];
        int n = 0;
        for (int i = 0; i < 16; ++i) { counts[i] = (uint8_t)D.byte(); n += counts[i]; }
        if (n > 256 || (tc & 15) > 3) { *err = "bad DHT"; return false; }
        uint8_t syms[256];
        for (int i = 0; i < n; ++i) syms[i] = (uint8_t)D.byte();
        if ((tc >> 4) > 1 || !((tc >> 4) ? D.ac : D.dc)[tc & 15].build(counts, syms, n)) { *err = "bad DHT"; return false; }
      }
    } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {  // SOF0/1/2
      D.progressive = m == 0xc2;
      if (D.byte() != 8) { *err = "only 8-bit JPEG is supported"; return false; }
      D.height = D.word();
      D.width = D.word();
      D.ncomp = D.byte();
      if ((D.ncomp != 1 && D.ncomp != 3) || D.width <= 0 || D.height <= 0) { *err = "unsupported JPEG frame"; return false; }
      if (have_frame) { *err = "more than one JPEG frame"; return false; }
      if ((uint64_t)D.width * (uint64_t)D.height > (1ull << 28)) { *err = "JPEG image too large"; return false; }
      D.hmax = D.vmax = 1;
      for (int i = 0; i < D.ncomp; ++i) {
        D.comp[i].id = D.byte();
        const int hv = D.byte();
        D.comp[i].h = hv >> 4;
        D.comp[i].v = hv & 15;
        D.comp[i].tq = D.byte();
        if (D.comp[i].h < 1 || D.comp[i].h > 4 || D.comp[i].v < 1 || D.comp[i].v > 4 || D.comp[i].tq > 3) { *err = "bad SOF"; return false; }
        D.hmax = D.comp[i].h > D.hmax ? D.comp[i].h : D.hmax;
        D.vmax = D.comp[i].v > D.vmax ? D.comp[i].v : D.vmax;
      }
      for (int i = 0; i < D.ncomp; ++i)
        if (D.hmax % D.comp[i].h != 0 || D.vmax % D.comp[i].v != 0) { *err = "bad JPEG sampling factors"; return false; }
      const int mcux = (D.width + 8 * D.hmax - 1) / (8 * D.hmax), mcuy = (D.height + 8 * D.vmax - 1) / (8 * D.vmax);
      for (int i = 0; i < D.ncomp; ++i) {
        D.comp[i].w_blocks = mcux * D.comp[i].h;
        D.comp[i].h_blocks = mcuy * D.comp[i].v;
        D.comp[i].plane.assign((size_t)D.comp[i].w_blocks * 8 * D.comp[i].h_blocks * 8, 0);
        if (D.progressive) D.comp[i].coef.assign((size_t)D.comp[i].w_blocks * D.comp[i].h_blocks * 64, 0);
      }
      have_frame = true;
    } else if (m == 0xdd) {  // DRI
      D.restart_interval = D.word();
    } else if (m == 0xda) {  // SOS
      if (!have_frame) { *err = "SOS before SOF"; return false; }
      n_scan = D.byte();
      if (n_scan < 1 || n_scan > D.ncomp) { *err = "bad SOS"; return false; }
      for (int i = 0; i < n_scan; ++i) {
        const int id = D.byte(), tt = D.byte();
        int which = -1;
        for (int k = 0; k < D.ncomp; ++k) if (D.comp[k].id == id) which = k;
        if (which < 0) { *err = "bad SOS component"; return false; }
        D.comp[which].td = tt >> 4;
        D.comp[which].ta = tt & 15;
        if (D.comp[which].td > 3 || D.comp[which].ta > 3) { *err = "bad SOS table"; return false; }
        scan_comps[i] = which;
      }
      D.ss = D.byte();
      D.se = D.byte();
      {
        const int aa = D.byte();
        D.ah = aa >> 4;
        D.al = aa & 15;
      }
      if (D.progressive) {
        if (D.ss > 63 || D.se > 63 || D.ss > D.se || D.ah > 13 || D.al > 13) { *err = "bad SOS"; return false; }
        if (D.ss == 0 && D.se != 0) { *err = "JPEG scan mixes DC and AC coefficients"; return false; }
        if (D.ss != 0 && n_scan != 1) { *err = "interleaved AC scan"; return false; }
      } else {
        D.ss = 0;
        D.se = 63;
        D.ah = D.al = 0;
      }
      for (int i = 0; i < n_scan; ++i) {
        const Component& sc = D.comp[scan_comps[i]];
        const bool need_dc = !D.progressive || (D.ss == 0 && D.ah == 0), need_ac = !D.progressive || D.ss != 0;
        if ((need_dc && !D.dc[sc.td].ok) || (need_ac && !D.ac[sc.ta].ok)) {
          *err = "JPEG scan uses a Huffman table that was never defined";
          return false;
        }
      }
      D.eobrun = 0;
      D.p = seg_end;
      // entropy-coded segment
      D.bitbuf = 0; D.bitcnt = 0; D.hit_marker = false;
      for (int i = 0; i < D.ncomp; ++i) D.comp[i].dc_pred = 0;
      short blk[64];
      int todo = D.restart_interval ? D.restart_interval : 0x7fffffff;
      auto restart_if_needed = [&]() -> bool {
        if (--todo > 0) return true;
        // expect RSTn
        D.bitbuf = 0; D.bitcnt = 0;
        if (!D.hit_marker) {  // consume padding up to the marker
          while (D.p + 1 < D.end && !(D.p[0] == 0xff && D.p[1] >= 0xd0 && D.p[1] <= 0xd7)) {
            if (D.p[0] == 0xff && D.p[1] != 0) return true;  // some other marker: let the outer loop see it
            ++D.p;
          }
        }
        if (D.p + 1 < D.end && D.p[0] == 0xff && D.p[1] >= 0xd0 && D.p[1] <= 0xd7) {
          D.p += 2;
          D.hit_marker = false;
          for (int i = 0; i < D.ncomp; ++i) D.comp[i].dc_pred = 0;
          D.eobrun = 0;
          todo = D.restart_interval;
        }
        return true;
      };
      if (n_scan == 1) {
        Component& c = D.comp[scan_comps[0]];
        const int bw = (((D.width * c.h + D.hmax - 1) / D.hmax) + 7) / 8, bh = (((D.height * c.v + D.vmax - 1) / D.vmax) + 7) / 8;
        for (int by = 0; by < bh; ++by)
          for (int bx = 0; bx < bw; ++bx) {
            if (D.progressive) {
              short* cb = &c.coef[((size_t)by * c.w_blocks + bx) * 64];
              if (!(D.ss == 0 ? prog_block_dc(D, c, cb) : prog_block_ac(D, c, cb))) { *err = "corrupt JPEG data"; return false; }
            } else {
              if (!decode_block(D, c, blk)) { *err = "corrupt JPEG data"; return false; }
              idct_block(&c.plane[((size_t)by * 8) * c.w_blocks * 8 + (size_t)bx * 8], c.w_blocks * 8, blk);
            }
            restart_if_needed();
          }
      } else {
        const int mcux = (D.width + 8 * D.hmax - 1) / (8 * D.hmax), mcuy = (D.height + 8 * D.vmax - 1) / (8 * D.vmax);
        for (int my = 0; my < mcuy; ++my)
          for (int mx = 0; mx < mcux; ++mx) {
            for (int i = 0; i < n_scan; ++i) {
              Component& c = D.comp[scan_comps[i]];
              for (int y = 0; y < c.v; ++y)
                for (int x = 0; x < c.h; ++x) {
                  if (D.progressive) {  // only DC scans are interleaved (checked above)
                    short* cb = &c.coef[(((size_t)my * c.v + y) * c.w_blocks + ((size_t)mx * c.h + x)) * 64];
                    if (!prog_block_dc(D, c, cb)) { *err = "corrupt JPEG data"; return false; }
                    continue;
                  }
                  if (!decode_block(D, c, blk)) { *err = "corrupt JPEG data"; return false; }
                  const size_t px = ((size_t)mx * c.h + x) * 8, py = ((size_t)my * c.v + y) * 8;
                  idct_block(&c.plane[py * c.w_blocks * 8 + px], c.w_blocks * 8, blk);
                }
            }
            restart_if_needed();
          }
      }
      // continue scanning for markers after the entropy-coded data
      if (D.hit_marker) D.hit_marker = false;
      continue;
    }
    D.p = seg_end;
  }
  if (!have_frame) { *err = "JPEG has no frame"; return false; }
  if (D.progressive) {
    // all scans are in: dequantise (16-bit, as stb_image does) and transform every block
    for (int k = 0; k < D.ncomp; ++k) {
      Component& c = D.comp[k];
      for (int by = 0; by < c.h_blocks; ++by)
        for (int bx = 0; bx < c.w_blocks; ++bx) {
          short* cb = &c.coef[((size_t)by * c.w_blocks + bx) * 64];
          for (int i = 0; i < 64; ++i) cb[i] = (short)(cb[i] * D.quant[c.tq][i]);
          idct_block(&c.plane[((size_t)by * 8) * c.w_blocks * 8 + (size_t)bx * 8], c.w_blocks * 8, cb);
        }
    }
  }
  *w = D.width;
  *h = D.height;
  *channels = D.ncomp;
  out->assign((size_t)D.width * D.height * D.ncomp, 0);
  // Chroma upsampling as stb_image does it (apps/src/stb_image.h:3388-3468, 3836-3880): "JFIF-centred" tent
  // filters for the factors 2x1, 1x2 and 2x2, plain replication for everything else, and per component a
  // little state machine that picks the nearer and the farther source row of every output row.
  struct Upsampler {
    int hs = 1, vs = 1, w_lores = 0, rows = 0, ystep = 0, row0 = 0, row1 = 0, ypos = 0;
    std::vector<uint8_t> line;
  };
  Upsampler up[3];
  for (int k = 0; k < D.ncomp; ++k) {
    const Component& c = D.comp[k];
    Upsampler& u = up[k];
    u.hs = D.hmax / c.h;
    u.vs = D.vmax / c.v;
    u.ystep = u.vs >> 1;
    u.w_lores = (D.width + u.hs - 1) / u.hs;
    u.rows = (D.height * c.v + D.vmax - 1) / D.vmax;
    u.line.assign((size_t)u.w_lores * u.hs + 8, 0);
  }
  for (int y = 0; y < D.height; ++y) {
    const uint8_t* row[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < D.ncomp; ++k) {
      const Component& c = D.comp[k];
      Upsampler& u = up[k];
      const size_t stride = (size_t)c.w_blocks * 8;
      const bool lower = u.ystep >= (u.vs >> 1);
      const uint8_t* near = c.plane.data() + stride * (size_t)(lower ? u.row1 : u.row0);
      const uint8_t* far = c.plane.data() + stride * (size_t)(lower ? u.row0 : u.row1);
      uint8_t* o = u.line.data();
      const int w = u.w_lores;
      if (u.hs == 1 && u.vs == 1) {
        row[k] = near;
      } else if (u.hs == 1 && u.vs == 2) {
        for (int i = 0; i < w; ++i) o[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        row[k] = o;
      } else if (u.hs == 2 && u.vs == 1) {
        if (w == 1) {
          o[0] = o[1] = near[0];
        } else {
          o[0] = near[0];
          o[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
          int i = 1;
          for (; i < w - 1; ++i) {
            const int n = 3 * near[i] + 2;
            o[2 * i] = (uint8_t)((n + near[i - 1]) >> 2);
            o[2 * i + 1] = (uint8_t)((n + near[i + 1]) >> 2);
          }
          o[2 * i] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
          o[2 * i + 1] = near[w - 1];
        }
        row[k] = o;
      } else if (u.hs == 2 && u.vs == 2) {
        if (w == 1) {
          o[0] = o[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
        } else {
          int t1 = 3 * near[0] + far[0];
          o[0] = (uint8_t)((t1 + 2) >> 2);
          for (int i = 1; i < w; ++i) {
            const int t0 = t1;
            t1 = 3 * near[i] + far[i];
            o[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            o[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
          }
          o[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
        }
        row[k] = o;
      } else {
        for (int i = 0; i < w; ++i)
          for (int j = 0; j < u.hs; ++j) o[i * u.hs + j] = near[i];
        row[k] = o;
      }
      if (++u.ystep >= u.vs) {
        u.ystep = 0;
        u.row0 = u.row1;
        if (++u.ypos < u.rows) u.row1 += 1;
      }
    }
    uint8_t* o = out->data() + (size_t)y * D.width * D.ncomp;
    if (D.ncomp == 1) {
      for (int x = 0; x < D.width; ++x) o[x] = row[0][x];
    } else {
      for (int x = 0; x < D.width; ++x) {
        // 20-bit fixed point: (int)(c * 4096.0f + 0.5f) << 8
        const int yf = ((int)row[0][x] << 20) + (1 << 19);
        const int cb = (int)row[1][x] - 128, cr = (int)row[2][x] - 128;
        int r = yf + cr * (5743 << 8);
        int g = yf + (cr * -(2925 << 8)) + ((cb * -(1410 << 8)) & 0xffff0000);
        int b = yf + cb * (7258 << 8);
        r >>= 20;
        g >>= 20;
        b >>= 20;
        o[3 * x] = clamp_u8(r);
        o[3 * x + 1] = clamp_u8(g);
        o[3 * x + 2] = clamp_u8(b);
      }
    }
  }
  return true;
}

}  // namespace jpg

inline bool read_file(const std::string& path, std::vector<uint8_t>* buf) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf->resize(n > 0 ? (size_t)n : 0);
  size_t got = n > 0 ? fread(buf->data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == buf->size();
}

// Binary PPM (P6, maxval 255) and PGM (P5).
inline bool decode_pnm(const std::vector<uint8_t>& buf, int* w, int* h, int* c, std::vector<uint8_t>* out) {
  if (buf.size() < 3 || buf[0] != 'P' || (buf[1] != '6' && buf[1] != '5')) return false;
  size_t i = 2;
  int vals[3], k = 0;
  while (k < 3 && i < buf.size()) {
    while (i < buf.size() && (buf[i] == ' ' || buf[i] == '\n' || buf[i] == '\r' || buf[i] == '\t')) ++i;
    if (i < buf.size() && buf[i] == '#') {
      while (i < buf.size() && buf[i] != '\n') ++i;
      continue;
    }
    int v = 0;
    bool any = false;
    while (i < buf.size() && buf[i] >= '0' && buf[i] <= '9') { v = v * 10 + (buf[i] - '0'); ++i; any = true; }
    if (!any) return false;
    vals[k++] = v;
  }
  if (k < 3 || vals[2] != 255) return false;
  ++i;  // single whitespace after maxval
  *w = vals[0];
  *h = vals[1];
  *c = buf[1] == '6' ? 3 : 1;
  const size_t need = (size_t)*w * *h * *c;
  if (buf.size() < i + need) return false;
  out->assign(buf.begin() + (long)i, buf.begin() + (long)(i + need));
  return true;
}

// stbi_load(path, &w, &h, &c, 0) with stbi_set_flip_vertically_on_load(flip).
inline bool decode_image_file(const std::string& path, bool flip_vertically, int* w, int* h, int* c, std::vector<uint8_t>* texels) {
  std::vector<uint8_t> buf;
  if (!read_file(path, &buf)) return false;
  std::string err;
  bool ok = false;
  if (buf.size() > 2 && buf[0] == 0xff && buf[1] == 0xd8) ok = jpg::decode(buf.data(), buf.size(), w, h, c, texels, &err);
  else if (buf.size() > 8 && buf[0] == 0x89 && buf[1] == 'P' && buf[2] == 'N' && buf[3] == 'G') ok = png::decode(buf.data(), buf.size(), w, h, c, texels, &err);
  else if (buf.size() > 2 && buf[0] == 'B' && buf[1] == 'M') ok = bmp::decode(buf.data(), buf.size(), w, h, c, texels, &err);
  else ok = decode_pnm(buf, w, h, c, texels);
  if (!ok && tga::looks_like_tga(buf.data(), buf.size())) ok = tga::decode(buf.data(), buf.size(), w, h, c, texels, &err);  // no signature: last
  if (!ok) return false;
  if (flip_vertically) {
    const size_t row = (size_t)*w * *c;
    std::vector<uint8_t> tmp(row);
    for (int y = 0; y < *h / 2; ++y) {
      uint8_t* a = texels->data() + (size_t)y * row;
      uint8_t* b = texels->data() + (size_t)(*h - 1 - y) * row;
      memcpy(tmp.data(), a, row);
      memcpy(a, b, row);
      memcpy(b, tmp.data(), row);
    }
  }
  return true;
}

}  // namespace b2host
