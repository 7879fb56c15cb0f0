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
#include <math.h>

#include "arith.h"

namespace madicp {

struct Sym3 {  // lower triangle of a symmetric matrix
  double xx, yx, zx, yy, zy, zz;
};

// The decomposition is split around its three libm calls (atan2, cos, sin -- the only operations of the whole
// MAD-tree build that are not IEEE-exact basic arithmetic) so that host and device run the SAME arithmetic
// before and after them, and the device-side build (gpu_tree.cu) can have the host's glibc evaluate them:
//   eig3_prepare(cov)        -> shifted/scaled matrix, cubic coefficients, the two arguments of atan2
//   [theta = atan2(sq, half_b) / 3, ct = cos(theta), st = sin(theta)]   (host libm)
//   eig3_finish(mid, ct, st) -> eigenvectors
struct Eig3Mid {
  Sym3 s;          // (cov - shift*I) / scale
  double c2_3, rho;
  double sq, half_b;  // theta = atan2(sq, half_b) / 3
};

namespace eig3_detail {

// Null vector of the (rank-2) symmetric matrix m (full 3x3, column-major m[c*3+r]); also returns the
// column used as pivot ("representative").
MADICP_HD void null_vector(const double m[9], double out[3], double rep[3]) {
  int p = 0;
  double best = fabs(m[0]);
  if (fabs(m[4]) > best) {
    best = fabs(m[4]);
    p = 1;
  }
  if (fabs(m[8]) > best) p = 2;
  for (int i = 0; i < 3; ++i) rep[i] = m[p * 3 + i];
  const double* u = m + ((p + 1) % 3) * 3;
  const double* w = m + ((p + 2) % 3) * 3;
  const double a[3] = {sub_(mul_(rep[1], u[2]), mul_(rep[2], u[1])), sub_(mul_(rep[2], u[0]), mul_(rep[0], u[2])),
                       sub_(mul_(rep[0], u[1]), mul_(rep[1], u[0]))};
  const double b[3] = {sub_(mul_(rep[1], w[2]), mul_(rep[2], w[1])), sub_(mul_(rep[2], w[0]), mul_(rep[0], w[2])),
                       sub_(mul_(rep[0], w[1]), mul_(rep[1], w[0]))};
  const double na = dot3(a[0], a[1], a[2], a[0], a[1], a[2]);
  const double nb = dot3(b[0], b[1], b[2], b[0], b[1], b[2]);
  if (na > nb) {
    const double s = sqrt(na);
    for (int i = 0; i < 3; ++i) out[i] = a[i] / s;
  } else {
    const double s = sqrt(nb);
    for (int i = 0; i < 3; ++i) out[i] = b[i] / s;
  }
}
}  // namespace eig3_detail

// cov: symmetric input (lower triangle used).
MADICP_HD void eig3_prepare(const Sym3& cov, Eig3Mid& mid) {
  const double shift = add_(add_(cov.xx, cov.yy), cov.zz) / 3.0;
  Sym3 s = cov;
  s.xx = sub_(s.xx, shift);
  s.yy = sub_(s.yy, shift);
  s.zz = sub_(s.zz, shift);
  // max |coeff| over the full 3x3 in column-major visiting order (first maximum wins; symmetric, so
  // visiting the six unique entries in that order is equivalent)
  const double full[9] = {s.xx, s.yx, s.zx, s.yx, s.yy, s.zy, s.zx, s.zy, s.zz};
  double scale = fabs(full[0]);
  for (int i = 1; i < 9; ++i) {
    const double v = fabs(full[i]);
    if (v > scale) scale = v;
  }
  if (scale > 0.0) {
    s.xx /= scale; s.yx /= scale; s.zx /= scale;
    s.yy /= scale; s.zy /= scale; s.zz /= scale;
  }
  mid.s = s;
  const Sym3& a = s;
  const double third = 1.0 / 3.0;
  // c0 = xx*yy*zz + 2*yx*zx*zy - xx*zy*zy - yy*zx*zx - zz*yx*yx   (left to right)
  double c0 = mul_(mul_(a.xx, a.yy), a.zz);
  c0 = add_(c0, mul_(mul_(mul_(2.0, a.yx), a.zx), a.zy));
  c0 = sub_(c0, mul_(mul_(a.xx, a.zy), a.zy));
  c0 = sub_(c0, mul_(mul_(a.yy, a.zx), a.zx));
  c0 = sub_(c0, mul_(mul_(a.zz, a.yx), a.yx));
  // c1 = xx*yy - yx*yx + xx*zz - zx*zx + yy*zz - zy*zy
  double c1 = sub_(mul_(a.xx, a.yy), mul_(a.yx, a.yx));
  c1 = add_(c1, mul_(a.xx, a.zz));
  c1 = sub_(c1, mul_(a.zx, a.zx));
  c1 = add_(c1, mul_(a.yy, a.zz));
  c1 = sub_(c1, mul_(a.zy, a.zy));
  const double c2 = add_(add_(a.xx, a.yy), a.zz);
  const double c2_3 = mul_(c2, third);
  double a_3 = mul_(sub_(mul_(c2, c2_3), c1), third);
  if (a_3 < 0.0) a_3 = 0.0;
  const double half_b = mul_(0.5, add_(c0, mul_(c2_3, sub_(mul_(mul_(2.0, c2_3), c2_3), c1))));
  double q = sub_(mul_(mul_(a_3, a_3), a_3), mul_(half_b, half_b));
  if (q < 0.0) q = 0.0;
  mid.c2_3 = c2_3;
  mid.rho = sqrt(a_3);
  mid.sq = sqrt(q);
  mid.half_b = half_b;
}

// ct = cos(theta), st = sin(theta), theta = atan2(mid.sq, mid.half_b) * (1.0 / 3.0).
// V: eigenvectors, column-major (V[c*3+r]), ascending eigenvalues.
MADICP_HD void eig3_finish(const Eig3Mid& mid, double ct, double st, double V[9]) {
  using namespace eig3_detail;
  const Sym3& s = mid.s;
  const double sqrt3 = sqrt(3.0);
  double lam[3];
  lam[0] = sub_(mid.c2_3, mul_(mid.rho, add_(ct, mul_(sqrt3, st))));
  lam[1] = sub_(mid.c2_3, mul_(mid.rho, sub_(ct, mul_(sqrt3, st))));
  lam[2] = add_(mid.c2_3, mul_(mul_(2.0, mid.rho), ct));
  const double eps = 2.220446049250313e-16;  // std::numeric_limits<double>::epsilon()
  if (sub_(lam[2], lam[0]) <= eps) {
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double gap_hi = sub_(lam[2], lam[1]);
  const double gap_lo = sub_(lam[1], lam[0]);
  int k = 0, l = 2;
  double d0 = gap_hi, d1 = gap_lo;
  if (d0 > d1) {
    k = 2;
    l = 0;
    d0 = d1;
  }
  double m[9] = {s.xx, s.yx, s.zx, s.yx, s.yy, s.zy, s.zx, s.zy, s.zz};
  m[0] = sub_(m[0], lam[k]);
  m[4] = sub_(m[4], lam[k]);
  m[8] = sub_(m[8], lam[k]);
  double vk[3], vl[3];
  null_vector(m, vk, vl);
  // NOTE: when d0 was swapped, d1 keeps the ORIGINAL lower gap (as the restated routine does)
  if (d0 <= mul_(mul_(2.0, eps), d1)) {
    const double d = dot3(vk[0], vk[1], vk[2], vl[0], vl[1], vl[2]);
    for (int i = 0; i < 3; ++i) vl[i] = sub_(vl[i], mul_(d, vl[i]));
    const double n = norm3(vl[0], vl[1], vl[2]);
    for (int i = 0; i < 3; ++i) vl[i] /= n;
  } else {
    double m2[9] = {s.xx, s.yx, s.zx, s.yx, s.yy, s.zy, s.zx, s.zy, s.zz};
    m2[0] = sub_(m2[0], lam[l]);
    m2[4] = sub_(m2[4], lam[l]);
    m2[8] = sub_(m2[8], lam[l]);
    double dummy[3];
    null_vector(m2, vl, dummy);
  }
  // runtime k/l are 0 or 2: write through selects (keeps V in registers on the device)
  for (int i = 0; i < 3; ++i) {
    V[i] = (k == 0) ? vk[i] : vl[i];
    V[6 + i] = (k == 0) ? vl[i] : vk[i];
  }
  const double* c2 = V + 6;
  const double* c0 = V;
  double mid_v[3] = {sub_(mul_(c2[1], c0[2]), mul_(c2[2], c0[1])), sub_(mul_(c2[2], c0[0]), mul_(c2[0], c0[2])),
                     sub_(mul_(c2[0], c0[1]), mul_(c2[1], c0[0]))};
  const double n = norm3(mid_v[0], mid_v[1], mid_v[2]);
  for (int i = 0; i < 3; ++i) V[3 + i] = mid_v[i] / n;
}

#if !defined(__CUDA_ARCH__)
// cov -> eigenvectors on the host (libm's atan2 / cos / sin)
inline void eig3_symmetric(const Sym3& cov, double V[9]) {
  Eig3Mid mid;
  eig3_prepare(cov, mid);
  const double theta = atan2(mid.sq, mid.half_b) * (1.0 / 3.0);
  eig3_finish(mid, cos(theta), sin(theta), V);
}
#endif

}  // namespace madicp
