// pose_math.h -- host-side SE(3) helpers shared by the C++ facade (csrc/facade/pipeline.hpp) and the
// library's ingest code (csrc/ingest.cpp).  Plain doubles, the reference's operand order, no FMA
// (compile with -ffp-contract=off): poses 3x4 row-major [R|t].
#pragma once
#include <cmath>

namespace madicp_pose {

struct Pose {  // 3x4 row-major [R|t]
  double m[12];
};
inline Pose poseIdentity() { return Pose{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}}; }
inline Pose poseMul(const Pose& A, const Pose& B) {
  Pose C;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      C.m[r * 4 + c] = (A.m[r * 4] * B.m[c] + A.m[r * 4 + 1] * B.m[4 + c]) + A.m[r * 4 + 2] * B.m[8 + c];
    C.m[r * 4 + 3] = ((A.m[r * 4] * B.m[3] + A.m[r * 4 + 1] * B.m[7]) + A.m[r * 4 + 2] * B.m[11]) + A.m[r * 4 + 3];
  }
  return C;
}
inline Pose poseInverse(const Pose& T) {
  Pose I;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) I.m[r * 4 + c] = T.m[c * 4 + r];
  for (int r = 0; r < 3; ++r)
    I.m[r * 4 + 3] = -((I.m[r * 4] * T.m[3] + I.m[r * 4 + 1] * T.m[7]) + I.m[r * 4 + 2] * T.m[11]);
  return I;
}
inline void poseApply(const Pose& T, const double* p, double* o) {  // o may alias p
  const double x = p[0], y = p[1], z = p[2];
  for (int r = 0; r < 3; ++r) o[r] = ((T.m[r * 4] * x + T.m[r * 4 + 1] * y) + T.m[r * 4 + 2] * z) + T.m[r * 4 + 3];
}
// tools/lie_algebra.h:39-52 (small-angle branch theta^2 < 1e-8 -> I + [w]x)
inline Pose poseFromTwist(const double t[3], const double w[3]) {
  Pose P = poseIdentity();
  const double th2 = (w[0] * w[0] + w[1] * w[1]) + w[2] * w[2];
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double R[9];
  if (th2 < 1e-8) {
    for (int i = 0; i < 9; ++i) R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + W[i];
  } else {
    const double th = std::sqrt(th2);
    double K[9], oK[9];
    for (int i = 0; i < 9; ++i) K[i] = W[i] / th;
    const double omc = 2.0 * std::sin(th / 2.0) * std::sin(th / 2.0), s = std::sin(th);
    for (int i = 0; i < 9; ++i) oK[i] = omc * K[i];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        R[r * 3 + c] = (((r == c) ? 1.0 : 0.0) + s * K[r * 3 + c]) +
                       ((oK[r * 3] * K[c] + oK[r * 3 + 1] * K[3 + c]) + oK[r * 3 + 2] * K[6 + c]);
  }
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) P.m[r * 4 + c] = R[r * 3 + c];
    P.m[r * 4 + 3] = t[r];
  }
  return P;
}
// tools/lie_algebra.h:54-89
inline void logSO3(const Pose& T, double w[3]) {
  const double R11 = T.m[0], R12 = T.m[1], R13 = T.m[2], R21 = T.m[4], R22 = T.m[5], R23 = T.m[6], R31 = T.m[8],
               R32 = T.m[9], R33 = T.m[10];
  const double tr = R11 + R22 + R33;
  if (tr + 1.0 < 1e-10) {
    if (std::fabs(R33 + 1.0) > 1e-5) {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R33);
      w[0] = f * R13; w[1] = f * R23; w[2] = f * (1.0 + R33);
    } else if (std::fabs(R22 + 1.0) > 1e-5) {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R22);
      w[0] = f * R12; w[1] = f * (1.0 + R22); w[2] = f * R32;
    } else {
      const double f = M_PI / std::sqrt(2.0 + 2.0 * R11);
      w[0] = f * (1.0 + R11); w[1] = f * R21; w[2] = f * R31;
    }
    return;
  }
  double mag;
  const double tr_3 = tr - 3.0;
  if (tr_3 < -1e-7) {
    const double theta = std::acos((tr - 1.0) / 2.0);
    mag = theta / (2.0 * std::sin(theta));
  } else {
    mag = 0.5 - tr_3 * tr_3 / 12.0;
  }
  w[0] = mag * (R32 - R23); w[1] = mag * (R13 - R31); w[2] = mag * (R21 - R12);
}
}  // namespace madicp_pose
