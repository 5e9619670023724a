// hostmath.h -- load-time matrix and camera arithmetic of the scene loader.
//
// The reference builds every Geom's transform, inverse and inverse-transpose
// with glm 0.9.6.3 on the host (apps/src/utilities.cpp:65-72,
// apps/src/scene.cpp:301-304).  Hit distances are compared across geoms with a
// strict `<`, so a 1-ulp difference in a matrix could flip a closest-hit id;
// the functions below therefore evaluate the same expression trees in the same
// order (glm/gtc/matrix_transform.inl:40-134, glm/detail/type_mat4x4.inl:37-92
// and :686-703, glm/gtc/matrix_inverse.inl:95-146).  tests/test_loader.py pins
// the result bit-for-bit against the reference loader's own output
// (tests/golden/*.b2s).  Matrices are float[16], column-major: m[c*4+r].
#pragma once

#include <cmath>
#include <cstring>

namespace b2host {

struct M4 {
  float m[16];
  float& at(int c, int r) { return m[c * 4 + r]; }
  float at(int c, int r) const { return m[c * 4 + r]; }
};

struct F4 {
  float v[4];
};

inline F4 col(const M4& a, int c) { return F4{{a.m[c * 4], a.m[c * 4 + 1], a.m[c * 4 + 2], a.m[c * 4 + 3]}}; }
inline void set_col(M4& a, int c, F4 x) { std::memcpy(a.m + c * 4, x.v, 16); }
inline F4 operator*(F4 a, float s) { return F4{{a.v[0] * s, a.v[1] * s, a.v[2] * s, a.v[3] * s}}; }
inline F4 operator+(F4 a, F4 b) { return F4{{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]}}; }

inline M4 identity() {
  M4 r;
  std::memset(r.m, 0, sizeof r.m);
  r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
  return r;
}

// glm::translate(m, v): Result[3] = m[0]*v0 + m[1]*v1 + m[2]*v2 + m[3]
inline M4 translate(const M4& a, const float v[3]) {
  M4 r = a;
  set_col(r, 3, ((col(a, 0) * v[0] + col(a, 1) * v[1]) + col(a, 2) * v[2]) + col(a, 3));
  return r;
}

// glm::rotate(m, angle, axis) with an already unit axis along x, y or z.
inline M4 rotate(const M4& a, float angle, const float axis_in[3]) {
  const float c = std::cos(angle), s = std::sin(angle);
  // normalize(v) = v * (1 / sqrt(dot(v, v)))
  const float inv = 1.0f / std::sqrt((axis_in[0] * axis_in[0] + axis_in[1] * axis_in[1]) + axis_in[2] * axis_in[2]);
  const float ax[3] = {axis_in[0] * inv, axis_in[1] * inv, axis_in[2] * inv};
  const float t[3] = {(1.0f - c) * ax[0], (1.0f - c) * ax[1], (1.0f - c) * ax[2]};
  float R[3][3];
  R[0][0] = c + t[0] * ax[0];
  R[0][1] = 0 + t[0] * ax[1] + s * ax[2];
  R[0][2] = 0 + t[0] * ax[2] - s * ax[1];
  R[1][0] = 0 + t[1] * ax[0] - s * ax[2];
  R[1][1] = c + t[1] * ax[1];
  R[1][2] = 0 + t[1] * ax[2] + s * ax[0];
  R[2][0] = 0 + t[2] * ax[0] + s * ax[1];
  R[2][1] = 0 + t[2] * ax[1] - s * ax[0];
  R[2][2] = c + t[2] * ax[2];
  M4 r;
  for (int k = 0; k < 3; ++k) set_col(r, k, (col(a, 0) * R[k][0] + col(a, 1) * R[k][1]) + col(a, 2) * R[k][2]);
  set_col(r, 3, col(a, 3));
  return r;
}

// glm::scale(m, v): Result[i] = m[i] * v[i]
inline M4 scale(const M4& a, const float v[3]) {
  M4 r;
  for (int k = 0; k < 3; ++k) set_col(r, k, col(a, k) * v[k]);
  set_col(r, 3, col(a, 3));
  return r;
}

// operator*(mat4, mat4): Result[j] = A0*B[j][0] + A1*B[j][1] + A2*B[j][2] + A3*B[j][3]
inline M4 mul(const M4& a, const M4& b) {
  M4 r;
  for (int j = 0; j < 4; ++j)
    set_col(r, j, ((col(a, 0) * b.at(j, 0) + col(a, 1) * b.at(j, 1)) + col(a, 2) * b.at(j, 2)) + col(a, 3) * b.at(j, 3));
  return r;
}

// buildTransformationMatrix, apps/src/utilities.cpp:65-72: T * (Rx*Ry*Rz) * S, degrees.
inline M4 build_transform(const float t[3], const float rdeg[3], const float s[3]) {
  const float kPi = 3.1415926535897932384626422832795028841971f;
  const float X[3] = {1, 0, 0}, Y[3] = {0, 1, 0}, Z[3] = {0, 0, 1};
  const M4 T = translate(identity(), t);
  M4 R = rotate(identity(), rdeg[0] * kPi / 180, X);
  R = mul(R, rotate(identity(), rdeg[1] * kPi / 180, Y));
  R = mul(R, rotate(identity(), rdeg[2] * kPi / 180, Z));
  const M4 S = scale(identity(), s);
  return mul(mul(T, R), S);
}

// glm::inverse(mat4): cofactors, then one multiply by 1/det.
inline M4 inverse(const M4& m) {
  auto M = [&](int c, int r) { return m.at(c, r); };
  const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
  const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
  const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
  const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
  M4 inv;
  for (int k = 0; k < 4; ++k) {
    inv.at(0, k) = ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]) * sa[k];
    inv.at(1, k) = ((v0[k] * f0[k] - v2[k] * f3[k]) + v3[k] * f4[k]) * sb[k];
    inv.at(2, k) = ((v0[k] * f1[k] - v1[k] * f3[k]) + v3[k] * f5[k]) * sa[k];
    inv.at(3, k) = ((v0[k] * f2[k] - v1[k] * f4[k]) + v2[k] * f5[k]) * sb[k];
  }
  const float d0 = M(0, 0) * inv.at(0, 0), d1 = M(0, 1) * inv.at(1, 0), d2 = M(0, 2) * inv.at(2, 0), d3 = M(0, 3) * inv.at(3, 0);
  const float one_over_det = 1.0f / ((d0 + d1) + (d2 + d3));
  M4 r;
  for (int k = 0; k < 16; ++k) r.m[k] = inv.m[k] * one_over_det;
  return r;
}

// glm::inverseTranspose(mat4): cofactor matrix divided by the determinant.
inline M4 inverse_transpose(const M4& m) {
  auto M = [&](int c, int r) { return m.at(c, r); };
  const float s00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), s01 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  const float s02 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), s03 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  const float s04 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), s05 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  const float s06 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), s07 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
  const float s08 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), s09 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
  const float s10 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), s11 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
  const float s12 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), s13 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
  const float s14 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3), s15 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  const float s16 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3), s17 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
  const float s18 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  M4 inv;
  inv.at(0, 0) = +((M(1, 1) * s00 - M(1, 2) * s01) + M(1, 3) * s02);
  inv.at(0, 1) = -((M(1, 0) * s00 - M(1, 2) * s03) + M(1, 3) * s04);
  inv.at(0, 2) = +((M(1, 0) * s01 - M(1, 1) * s03) + M(1, 3) * s05);
  inv.at(0, 3) = -((M(1, 0) * s02 - M(1, 1) * s04) + M(1, 2) * s05);
  inv.at(1, 0) = -((M(0, 1) * s00 - M(0, 2) * s01) + M(0, 3) * s02);
  inv.at(1, 1) = +((M(0, 0) * s00 - M(0, 2) * s03) + M(0, 3) * s04);
  inv.at(1, 2) = -((M(0, 0) * s01 - M(0, 1) * s03) + M(0, 3) * s05);
  inv.at(1, 3) = +((M(0, 0) * s02 - M(0, 1) * s04) + M(0, 2) * s05);
  inv.at(2, 0) = +((M(0, 1) * s06 - M(0, 2) * s07) + M(0, 3) * s08);
  inv.at(2, 1) = -((M(0, 0) * s06 - M(0, 2) * s09) + M(0, 3) * s10);
  inv.at(2, 2) = +((M(0, 0) * s11 - M(0, 1) * s09) + M(0, 3) * s12);
  inv.at(2, 3) = -((M(0, 0) * s08 - M(0, 1) * s10) + M(0, 2) * s12);
  inv.at(3, 0) = -((M(0, 1) * s13 - M(0, 2) * s14) + M(0, 3) * s15);
  inv.at(3, 1) = +((M(0, 0) * s13 - M(0, 2) * s16) + M(0, 3) * s17);
  inv.at(3, 2) = -((M(0, 0) * s14 - M(0, 1) * s16) + M(0, 3) * s18);
  inv.at(3, 3) = +((M(0, 0) * s15 - M(0, 1) * s17) + M(0, 2) * s18);
  const float det = ((M(0, 0) * inv.at(0, 0) + M(0, 1) * inv.at(0, 1)) + M(0, 2) * inv.at(0, 2)) + M(0, 3) * inv.at(0, 3);
  for (int k = 0; k < 16; ++k) inv.m[k] = inv.m[k] / det;
  return inv;
}

// vec3 helpers in glm's evaluation order
inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline void cross3(const float a[3], const float b[3], float o[3]) {
  const float x = a[1] * b[2] - b[1] * a[2], y = a[2] * b[0] - b[2] * a[0], z = a[0] * b[1] - b[0] * a[1];
  o[0] = x;
  o[1] = y;
  o[2] = z;
}
inline void normalize3(const float a[3], float o[3]) {
  const float inv = 1.0f / std::sqrt(dot3(a, a));
  o[0] = a[0] * inv;
  o[1] = a[1] * inv;
  o[2] = a[2] * inv;
}

}  // namespace b2host
