// ingest.cpp -- host-side preparation of an incoming scan before its tree is built.
//
// madicp_deskew is the reference's Pipeline::deskew (odometry/pipeline.cpp:79-123): the points are sorted
// by azimuth, the sweep is cut into 1024 chunks, every chunk gets the pose interpolated from the last
// relative motion, and the cloud is rewritten IN SORTED ORDER with those poses applied.  The order matters
// downstream (the tree's sums run in array order), so the result must be the reference's permutation
// exactly -- and equal azimuths are the rule in lidar data (the beams of one firing column), so the
// order std::sort happens to leave among them is part of the result.  That order depends only on the
// sequence of comparison outcomes, not on what is being moved: the same std::sort over 16-byte
// (azimuth, index) records takes the same decisions as the reference's over its 32-byte (azimuth, point)
// pairs and ends with the same permutation, at half the traffic.  Everything else (atan2 per point, the
// chunk poses, the transform) is a pure per-point function and is spread over the worker pool.  The
// reference spends ~10 ms per 131k-point scan here on one core.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/madicp_b200.h"
#include "host_pool.hpp"
#include "pose_math.h"

namespace madicp {
void set_error(const std::string& msg);
}

namespace {
using madicp_host::default_init_allocator;
using madicp_host::for_chunks;
using madicp_pose::Pose;

constexpr int kChunks = 1024;  // tools/constants.h:31

struct Item {
  double key;
  int32_t idx;
  int32_t pad;
};
struct P3 {
  double v[3];
};

template <class T>
using RawVec = std::vector<T, default_init_allocator<T>>;

// sorted keys -> chunk number of every position (the reference advances by at most one chunk per point)
void sweep(const Item* it, int64_t n, double resolution, RawVec<int32_t>& cid, int32_t& last) {
  double angle = M_PI - resolution;
  int32_t c = 0;
  for (int64_t i = n - 1; i >= 0; --i) {
    if (it[i].key < angle) {
      angle -= resolution;
      ++c;
    }
    cid[size_t(i)] = c;
  }
  last = c;
}

}  // namespace

extern "C" int madicp_deskew(double* points_xyz, int64_t n, const double T_prev[12], const double T_now[12], double sensor_hz,
                             int num_threads) {
  if (!points_xyz || !T_prev || !T_now || n <= 0 || n > (int64_t(1) << 30) || !(sensor_hz > 0.0)) {
    madicp::set_error("madicp_deskew: bad arguments");
    return MADICP_ERR_INVALID;
  }
  int threads = num_threads < 1 ? 1 : (num_threads > 64 ? 64 : num_threads);
  if (n < 20000) threads = 1;
  const size_t un = size_t(n);
  // relative motion of the last two poses as a constant twist (pipeline.cpp:80-86)
  const double ts = 1. / sensor_hz;
  Pose a, b;
  std::memcpy(a.m, T_prev, sizeof(a.m));
  std::memcpy(b.m, T_now, sizeof(b.m));
  const Pose rel = madicp_pose::poseMul(madicp_pose::poseInverse(a), b);
  double w[3];
  madicp_pose::logSO3(rel, w);
  const double vel[6] = {rel.m[3] / ts, rel.m[7] / ts, rel.m[11] / ts, w[0] / ts, w[1] / ts, w[2] / ts};
  const double resolution = 2 * M_PI / double(kChunks), delta = ts / double(kChunks - 1);

  const bool timing = std::getenv("MADTREE_TIMING") != nullptr;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) {
    return std::chrono::duration<double, std::milli>(y - x).count();
  };
  const auto t0 = now();
  // azimuth of every point (pipeline.cpp:91-95)
  // (the same pass keeps a copy of the points: the result is written over the input in another order)
  RawVec<Item> items(un);
  RawVec<P3> in(un);
  for_chunks(threads, un, 8192, [&](size_t c0, size_t c1) {
    for (size_t i = c0; i < c1; ++i) {
      items[i] = Item{std::atan2(points_xyz[3 * i + 1], points_xyz[3 * i]), int32_t(i), 0};
      std::memcpy(in[i].v, points_xyz + 3 * i, sizeof(double) * 3);
    }
  });
  const auto t1 = now();

  // the reference's sort (pipeline.cpp:97-99), on records that carry an index instead of the point
  std::sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.key < y.key; });

  const auto t2 = now();
  // which chunk each sorted position falls in, then the chunk poses (t accumulates as in the reference)
  RawVec<int32_t> cid(un);
  int32_t last = 0;
  sweep(items.data(), n, resolution, cid, last);
  std::vector<Pose> meas(size_t(last) + 1);
  {
    double t = -ts;
    for (int32_t c = 0; c <= last; ++c) {
      const double tr[3] = {vel[0] * t, vel[1] * t, vel[2] * t}, ro[3] = {vel[3] * t, vel[4] * t, vel[5] * t};
      meas[size_t(c)] = madicp_pose::poseFromTwist(tr, ro);
      t += delta;
    }
  }
  // (*curr_cloud)[i] = meas_pose_to_robot * sorted[i].second  (pipeline.cpp:121)
  for_chunks(threads, un, 8192, [&](size_t c0, size_t c1) {
    for (size_t i = c0; i < c1; ++i) madicp_pose::poseApply(meas[size_t(cid[i])], in[size_t(items[i].idx)].v, points_xyz + 3 * i);
  });
  if (timing)
    std::fprintf(stderr, "madicp_deskew: n=%lld threads=%d azimuths %.2f ms, sort %.2f ms%s, sweep+poses+apply %.2f ms\n", (long long) n,
                 threads, ms(t0, t1), ms(t1, t2), "", ms(t2, now()));
  return MADICP_OK;
}
