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

// H: 6x6, only the lower triangle (H[r*6+c], r >= c) is read, as Eigen's LDLT<.., Lower> does.
// Written with compile-time loop bounds and select-based interchanges only (no run-time array
// index), so on the device the whole factorisation lives in registers: the update is on the
// critical path between two Gauss-Newton rounds and a generic pointer-indexed version costs tens of
// microseconds of single-thread latency per round.
MADICP_HD void ldlt6_solve_neg(const double* H, const double* b, double* x) {
  double A[6][6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) A[r][c] = (c <= r) ? H[r * 6 + c] : 0.0;
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] = -b[i];
  int perm[6];
  bool all_zero = false;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    // pivot: largest |diagonal| of the trailing block, first maximum wins
    int p = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const double a = fabs(A[i][i]);
      if (a > best) {
        best = a;
        p = i;
      }
    }
    perm[k] = p;
    // symmetric interchange k <-> p on the lower triangle, and on the right-hand side (P b is
    // applied on the fly: later transpositions only touch positions >= k)
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const bool sw = (p == i);
#pragma unroll
      for (int j = 0; j < k; ++j) {
        const double u = A[k][j], w = A[i][j];
        A[k][j] = sw ? w : u;
        A[i][j] = sw ? u : w;
      }
#pragma unroll
      for (int r = i + 1; r < 6; ++r) {
        const double u = A[r][k], w = A[r][i];
        A[r][k] = sw ? w : u;
        A[r][i] = sw ? u : w;
      }
      {
        const double u = A[k][k], w = A[i][i];
        A[k][k] = sw ? w : u;
        A[i][i] = sw ? u : w;
      }
#pragma unroll
      for (int r = k + 1; r < i; ++r) {
        const double u = A[r][k], w = A[i][r];
        A[r][k] = sw ? w : u;
        A[i][r] = sw ? u : w;
      }
      {
        const double u = y[k], w = y[i];
        y[k] = sw ? w : u;
        y[i] = sw ? u : w;
      }
    }
    if (k > 0) {
      double w[6];
#pragma unroll
      for (int j = 0; j < k; ++j) w[j] = A[j][j] * A[k][j];
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < k; ++j) s += A[k][j] * w[j];
      A[k][k] -= s;
#pragma unroll
      for (int i = k + 1; i < 6; ++i) {
        double s2 = 0.0;
#pragma unroll
        for (int j = 0; j < k; ++j) s2 += A[i][j] * w[j];
        A[i][k] -= s2;
      }
    }
    const double d = A[k][k];
    const bool ok = fabs(d) > 0.0;
    if (k == 0 && !ok) all_zero = true;  // whole diagonal is zero: nothing to factor
    if (ok && !all_zero) {
#pragma unroll
      for (int i = k + 1; i < 6; ++i) A[i][k] /= d;
    }
  }
  (void) perm;
  if (all_zero) {
    // Eigen stops factoring when the first pivot is exactly zero; D = 0 makes the solve return 0
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = 0.0;
    return;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] = (fabs(A[i][i]) > DBL_MIN) ? y[i] / A[i][i] : 0.0;
#pragma unroll
  for (int i = 5; i >= 0; --i)
#pragma unroll
    for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
  // x = P^T y: undo the transpositions in reverse order
#pragma unroll
  for (int k = 5; k >= 0; --k) {
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      const bool sw = (perm[k] == i);
      const double u = y[k], w = y[i];
      y[k] = sw ? w : u;
      y[i] = sw ? u : w;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// Rodrigues with the reference's small-angle branch (theta^2 < 1e-8 -> I + [w]x). R row-major 3x3.
MADICP_HD void expmap_so3(double wx, double wy, double wz, double* R) {
  const double th2 = dot3(wx, wy, wz, wx, wy, wz);
  const double W[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
  if (th2 < 1e-8) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + W[i];
    return;
  }
  const double th = sqrt(th2);
  double K[9], oK[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) K[i] = W[i] / th;
  const double hs = sin(th / 2.0);
  const double omc = 2.0 * hs * hs;
  const double s = sin(th);
#pragma unroll
  for (int i = 0; i < 9; ++i) oK[i] = omc * K[i];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
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
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) dX[r * 4 + c] = dR[r * 3 + c];
    dX[r * 4 + 3] = dx[r];
  }
  iso_mul(X, dX, Xn);
#pragma unroll
  for (int i = 0; i < 12; ++i) X[i] = Xn[i];
  if (dx_out) {
#pragma unroll
    for (int i = 0; i < 6; ++i) dx_out[i] = dx[i];
  }
}

}  // namespace madicp
