// solve6.h -- Gauss-Newton state update shared by host and device:
//   dx = H.ldlt().solve(-b);  dX = (expMapSO3(dx[3:6]), dx[0:3]);  X = X * dX
// (reference: odometry/mad_icp.cpp:105-117, tools/lie_algebra.h:33-52).
// The factorisation is the pivoted (largest |diagonal|) lower LDL^T that Eigen 3.4's
// LDLT<Matrix6d> performs, with the pseudo-inverse of D in the solve, so a rank-deficient
// or all-zero H yields dx = 0 on the null space instead of NaN.
#pragma once
#include <float.h>

#include "arith.h"

namespace madicp {

// H: 6x6 symmetric, only the lower triangle (r >= c, index r*6+c or c*6+r alike) is read.
MADICP_HD void ldlt6_solve_neg(const double* H, const double* b, double* x) {
  double A[6][6];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c <= r; ++c) A[r][c] = H[r * 6 + c];
  int perm[6];
  bool all_zero = false;
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < 6; ++i) {
      const double a = fabs(A[i][i]);
      if (a > best) {
        best = a;
        p = i;
      }
    }
    perm[k] = p;
    if (p != k) {  // symmetric row/column interchange on the lower triangle
      for (int j = 0; j < k; ++j) {
        const double t = A[k][j];
        A[k][j] = A[p][j];
        A[p][j] = t;
      }
      for (int i = p + 1; i < 6; ++i) {
        const double t = A[i][k];
        A[i][k] = A[i][p];
        A[i][p] = t;
      }
      {
        const double t = A[k][k];
        A[k][k] = A[p][p];
        A[p][p] = t;
      }
      for (int i = k + 1; i < p; ++i) {
        const double t = A[i][k];
        A[i][k] = A[p][i];
        A[p][i] = t;
      }
    }
    if (k > 0) {
      double w[6];
      for (int j = 0; j < k; ++j) w[j] = A[j][j] * A[k][j];
      double s = 0.0;
      for (int j = 0; j < k; ++j) s += A[k][j] * w[j];
      A[k][k] -= s;
      for (int i = k + 1; i < 6; ++i) {
        double s2 = 0.0;
        for (int j = 0; j < k; ++j) s2 += A[i][j] * w[j];
        A[i][k] -= s2;
      }
    }
    const double d = A[k][k];
    const bool ok = fabs(d) > 0.0;
    if (k == 0 && !ok) {
      all_zero = true;
      break;
    }
    if (ok)
      for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
  }
  if (all_zero)
    for (int j = 0; j < 6; ++j) perm[j] = j;
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = -b[i];
  for (int k = 0; k < 6; ++k)
    if (perm[k] != k) {
      const double t = y[k];
      y[k] = y[perm[k]];
      y[perm[k]] = t;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
  for (int i = 0; i < 6; ++i) y[i] = (fabs(A[i][i]) > DBL_MIN) ? y[i] / A[i][i] : 0.0;
  for (int i = 5; i >= 0; --i)
    for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
  for (int k = 5; k >= 0; --k)
    if (perm[k] != k) {
      const double t = y[k];
      y[k] = y[perm[k]];
      y[perm[k]] = t;
    }
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// Rodrigues with the reference's small-angle branch (theta^2 < 1e-8 -> I + [w]x). R row-major 3x3.
MADICP_HD void expmap_so3(double wx, double wy, double wz, double* R) {
  const double th2 = dot3(wx, wy, wz, wx, wy, wz);
  const double W[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
  if (th2 < 1e-8) {
    for (int i = 0; i < 9; ++i) R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + W[i];
    return;
  }
  const double th = sqrt(th2);
  double K[9], oK[9];
  for (int i = 0; i < 9; ++i) K[i] = W[i] / th;
  const double hs = sin(th / 2.0);
  const double omc = 2.0 * hs * hs;
  const double s = sin(th);
  for (int i = 0; i < 9; ++i) oK[i] = omc * K[i];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double kk = dot3(oK[r * 3], oK[r * 3 + 1], oK[r * 3 + 2], K[c], K[3 + c], K[6 + c]);
      R[r * 3 + c] = (((r == c) ? 1.0 : 0.0) + s * K[r * 3 + c]) + kk;
    }
}

// X (row-major 3x4) <- X * [exp(w) | t] with dx = [t, w] solved from (H, b).
MADICP_HD void gn_update_pose(const double* H, const double* b, double* X, double* dx_out) {
  double dx[6];
  ldlt6_solve_neg(H, b, dx);
  double dR[9], dX[12], Xn[12];
  expmap_so3(dx[3], dx[4], dx[5], dR);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) dX[r * 4 + c] = dR[r * 3 + c];
    dX[r * 4 + 3] = dx[r];
  }
  iso_mul(X, dX, Xn);
  for (int i = 0; i < 12; ++i) X[i] = Xn[i];
  if (dx_out)
    for (int i = 0; i < 6; ++i) dx_out[i] = dx[i];
}

}  // namespace madicp
