// eig3.h -- closed-form eigen-decomposition of a symmetric 3x3 (covariance) matrix, host only.
//
// The reference obtains split directions and surface normals from
// Eigen::SelfAdjointEigenSolver<Matrix3d>::computeDirect (tools/mad_tree.cpp:59-61).  Eigen is a
// third-party dependency that is not vendored in the reference tree (FetchContent, pinned 3.4.0), so
// this is a restatement of that routine's published algorithm: shift by trace/3, scale by the largest
// |coefficient|, trigonometric roots of the characteristic cubic (ascending), eigenvector of the
// best-separated eigenvalue as a normalised cross product of two columns of (A - lambda I), second one
// the same way (or orthogonalised when nearly degenerate), middle one by a cross product.
// Column 0 of the result is the normal (smallest eigenvalue), column 2 the split direction.
#pragma once
#include <cmath>
#include <limits>

#include "arith.h"

namespace madicp {

struct Sym3 {  // lower triangle of a symmetric matrix
  double xx, yx, zx, yy, zy, zz;
};

namespace eig3_detail {

inline void cubic_roots(const Sym3& a, double r[3]) {
  const double third = 1.0 / 3.0;
  const double sqrt3 = std::sqrt(3.0);
  const double c0 = a.xx * a.yy * a.zz + 2.0 * a.yx * a.zx * a.zy - a.xx * a.zy * a.zy - a.yy * a.zx * a.zx -
                    a.zz * a.yx * a.yx;
  const double c1 = a.xx * a.yy - a.yx * a.yx + a.xx * a.zz - a.zx * a.zx + a.yy * a.zz - a.zy * a.zy;
  const double c2 = a.xx + a.yy + a.zz;
  const double c2_3 = c2 * third;
  double a_3 = (c2 * c2_3 - c1) * third;
  if (a_3 < 0.0) a_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_3 * (2.0 * c2_3 * c2_3 - c1));
  double q = a_3 * a_3 * a_3 - half_b * half_b;
  if (q < 0.0) q = 0.0;
  const double rho = std::sqrt(a_3);
  const double theta = std::atan2(std::sqrt(q), half_b) * third;
  const double ct = std::cos(theta), st = std::sin(theta);
  r[0] = c2_3 - rho * (ct + sqrt3 * st);
  r[1] = c2_3 - rho * (ct - sqrt3 * st);
  r[2] = c2_3 + 2.0 * rho * ct;
}

// Null vector of the (rank-2) symmetric matrix m (full 3x3, column-major m[c*3+r]); also returns the
// column used as pivot ("representative").
inline void null_vector(const double m[9], double out[3], double rep[3]) {
  int p = 0;
  double best = std::fabs(m[0]);
  if (std::fabs(m[4]) > best) {
    best = std::fabs(m[4]);
    p = 1;
  }
  if (std::fabs(m[8]) > best) p = 2;
  for (int i = 0; i < 3; ++i) rep[i] = m[p * 3 + i];
  const double* u = m + ((p + 1) % 3) * 3;
  const double* w = m + ((p + 2) % 3) * 3;
  const double a[3] = {rep[1] * u[2] - rep[2] * u[1], rep[2] * u[0] - rep[0] * u[2], rep[0] * u[1] - rep[1] * u[0]};
  const double b[3] = {rep[1] * w[2] - rep[2] * w[1], rep[2] * w[0] - rep[0] * w[2], rep[0] * w[1] - rep[1] * w[0]};
  const double na = dot3(a[0], a[1], a[2], a[0], a[1], a[2]);
  const double nb = dot3(b[0], b[1], b[2], b[0], b[1], b[2]);
  if (na > nb) {
    const double s = std::sqrt(na);
    for (int i = 0; i < 3; ++i) out[i] = a[i] / s;
  } else {
    const double s = std::sqrt(nb);
    for (int i = 0; i < 3; ++i) out[i] = b[i] / s;
  }
}
}  // namespace eig3_detail

// cov: symmetric input (lower triangle used).  V: eigenvectors, column-major (V[c*3+r]), ascending.
inline void eig3_symmetric(const Sym3& cov, double V[9]) {
  using namespace eig3_detail;
  const double shift = (cov.xx + cov.yy + cov.zz) / 3.0;
  Sym3 s = cov;
  s.xx -= shift;
  s.yy -= shift;
  s.zz -= shift;
  // max |coeff| over the full 3x3 in column-major visiting order (first maximum wins; symmetric, so
  // visiting the six unique entries in that order is equivalent)
  const double full[9] = {s.xx, s.yx, s.zx, s.yx, s.yy, s.zy, s.zx, s.zy, s.zz};
  double scale = std::fabs(full[0]);
  for (int i = 1; i < 9; ++i) {
    const double v = std::fabs(full[i]);
    if (v > scale) scale = v;
  }
  if (scale > 0.0) {
    s.xx /= scale; s.yx /= scale; s.zx /= scale;
    s.yy /= scale; s.zy /= scale; s.zz /= scale;
  }
  double lam[3];
  cubic_roots(s, lam);
  const double eps = std::numeric_limits<double>::epsilon();
  if ((lam[2] - lam[0]) <= eps) {
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double gap_hi = lam[2] - lam[1];
  const double gap_lo = lam[1] - lam[0];
  int k = 0, l = 2;
  double d0 = gap_hi, d1 = gap_lo;
  if (d0 > d1) {
    k = 2;
    l = 0;
    d0 = d1;
  }
  (void) gap_hi;
  double m[9] = {s.xx, s.yx, s.zx, s.yx, s.yy, s.zy, s.zx, s.zy, s.zz};
  m[0] -= lam[k];
  m[4] -= lam[k];
  m[8] -= lam[k];
  double vk[3], vl[3];
  null_vector(m, vk, vl);
  // NOTE: when d0 was swapped, d1 keeps the ORIGINAL lower gap (as the restated routine does)
  if (d0 <= 2.0 * eps * d1) {
    const double d = dot3(vk[0], vk[1], vk[2], vl[0], vl[1], vl[2]);
    for (int i = 0; i < 3; ++i) vl[i] -= d * vl[i];
    const double n = norm3(vl[0], vl[1], vl[2]);
    for (int i = 0; i < 3; ++i) vl[i] /= n;
  } else {
    double m2[9] = {s.xx, s.yx, s.zx, s.yx, s.yy, s.zy, s.zx, s.zy, s.zz};
    m2[0] -= lam[l];
    m2[4] -= lam[l];
    m2[8] -= lam[l];
    double dummy[3];
    null_vector(m2, vl, dummy);
  }
  for (int i = 0; i < 3; ++i) {
    V[k * 3 + i] = vk[i];
    V[l * 3 + i] = vl[i];
  }
  const double* c2 = V + 6;
  const double* c0 = V;
  double mid[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
  const double n = norm3(mid[0], mid[1], mid[2]);
  for (int i = 0; i < 3; ++i) V[3 + i] = mid[i] / n;
}

}  // namespace madicp
