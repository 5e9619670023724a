// fuzz_image_decoders.cpp -- mutation fuzzer for the texture decoders of the scene loader (csrc/host/image_decode.h,
// png_decode.h, tga_decode.h, bmp_decode.h): a library must reject a damaged texture file, never crash or hang on it.
//   g++ -std=c++17 -O1 -g -fwrapv -fsanitize=address,undefined -fno-sanitize=signed-integer-overflow \
//       -I mygpuraytracer_b200/csrc/host tools/fuzz_image_decoders.cpp -o fuzz && ./fuzz file.jpg|png|tga|bmp iterations seed
// Bit flips, truncations, 4-byte overwrites and injected 0xff bytes; tests/test_image_fuzz.py runs it on every fixture.
#include "image_decode.h"
#include <random>
#include <cstdio>
int main(int argc, char** argv) {
  std::vector<uint8_t> data;
  if (!b2host::read_file(argv[1], &data)) return 2;
  const int n = atoi(argv[2]);
  std::mt19937 rng((unsigned)atoi(argv[3]));
  int ok = 0, bad = 0;
  for (int i = 0; i < n; ++i) {
    std::vector<uint8_t> b = data;
    const int mode = i % 4;
    if (mode == 0) { int k = 1 + rng() % 6; while (k--) b[2 + rng() % (b.size() - 2)] ^= (uint8_t)(1 + rng() % 255); }
    else if (mode == 1) b.resize(4 + rng() % (b.size() - 4));
    else if (mode == 2) { size_t k = 2 + rng() % (b.size() - 10); for (int j = 0; j < 4; ++j) b[k + j] = (uint8_t)rng(); }
    else { size_t k = 2 + rng() % (b.size() - 2); b[k] = 0xff; }
    int w, h, c; std::vector<uint8_t> out; std::string err;
    bool r;
    if (data[0] == 0xff) r = b2host::jpg::decode(b.data(), b.size(), &w, &h, &c, &out, &err);
    else if (data[0] == 0x89) r = b2host::png::decode(b.data(), b.size(), &w, &h, &c, &out, &err);
    else if (data[0] == 'B' && data[1] == 'M') r = b2host::bmp::decode(b.data(), b.size(), &w, &h, &c, &out, &err);
    else r = b2host::tga::decode(b.data(), b.size(), &w, &h, &c, &out, &err);
    if (r) ++ok; else ++bad;
    if (argc > 4) fprintf(stderr, "%d %d\n", i, (int)r);
  }
  printf("decoded %d rejected %d\n", ok, bad);
  return 0;
}
