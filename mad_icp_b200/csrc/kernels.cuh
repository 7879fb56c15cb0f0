// kernels.cuh -- sm_100a device code of the registration hot path.
//
//   K1  k_search        MADtree::bestMatchingLeafFast for every (moving leaf, keyframe)
//                       (reference: tools/mad_tree.cpp:144-152 called from odometry/mad_icp.cpp:78-79)
//   K2  k_linearize     gate + errorAndJacobian + Huber + weight + H/b accumulation
//                       (reference: odometry/mad_icp.cpp:59-72, 81-101), deterministic reduction
//   K3  k_solve         updateState (reference: odometry/mad_icp.cpp:105-117)
//   GN  k_gn_loop       all of the above for `iters` rounds in one persistent cooperative kernel
//                       (reference loop: odometry/pipeline.cpp:166-193), optional in-kernel
//                       all-reduce of H/b across GPUs through peer mailboxes (NVLink stores)
//
// Memory/branch bound FP64 work: no tensor cores.  A node visit is one 64-byte record = two
// 256-bit read-only loads (LDG.E.256); moving leaves are laid out in getLeafs (DFS) order so the
// lanes of a warp walk nearly the same path and their loads coalesce / hit L1 at the top levels.
// Compiled with -fmad=false; the predicate chain uses __d*_rn intrinsics (arith.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/madicp_b200.h"
#include "arith.h"
#include "solve6.h"

namespace madicp {

constexpr int kMaxSlots = 64;      // keyframe slots addressable by one launch
constexpr int kBlock = 256;        // threads per CTA for every kernel here
constexpr int kWarps = kBlock / 32;
constexpr int kAcc = 27;           // 21 lower-triangle entries of H + 6 entries of b
constexpr int kMaxPeers = 16;
constexpr int kMailboxSlots = 2;   // double-buffered by round parity

struct ModelView {  // passed by value (constant bank): the active keyframes of this device
  const madtree_rec_t* recs[kMaxSlots];
  int K;
};

struct IcpParams {
  double min_ball, rho_ker_sqrt, b_ratio;
};

// Control block + results of one registration, in device global memory.
struct GnState {
  int ticket;     // monotonically increasing arrival counter (reset by the host before a launch)
  int round;      // number of completed rounds (release/acquire flag)
  int n_matched;  // matched moving leaves in the last round
  int pad;
  double H[36];   // last round, full symmetric
  double b[6];
  double X_trace[(MADICP_MAX_ITERS + 1) * 12];  // pose before round i; [iters] = final pose
};

// LL-style mailbox cell: a double split into two 32-bit halves, each paired with a 32-bit epoch
// flag, written with ONE 16-byte store so data and flags arrive together (no fence on the wire).
struct __align__(16) LLCell {
  uint32_t lo, flag_lo, hi, flag_hi;
};
struct Mailbox {  // lives on every rank; cell [slot][src_rank][i] is written by src_rank
  LLCell cell[kMailboxSlots][kMaxPeers][32];
};

struct PeerView {
  Mailbox* box[kMaxPeers];  // box[r] = rank r's mailbox mapped into this process (box[rank] local)
  int rank, world;
  uint32_t epoch_base;      // first epoch value of this launch (monotonic across launches)
};

// ---------------------------------------------------------------------------------------------
struct Rec {  // a node record in registers
  double mx, my, mz, dx, dy, dz, bbox0;
  int link;
};

__device__ __forceinline__ Rec load_rec(const madtree_rec_t* p) {
  double a0, a1, a2, a3, b0, b1, b2, b3;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3) : "l"(p));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3)
               : "l"(reinterpret_cast<const char*>(p) + 32));
  Rec r;
  r.mx = a0; r.my = a1; r.mz = a2; r.dx = a3; r.dy = b0; r.dz = b1; r.bbox0 = b2;
  r.link = __double2loint(b3);
  return r;
}

// Greedy single-path descent; returns the record index of the leaf and the leaf record.
__device__ __forceinline__ int descend(const madtree_rec_t* __restrict__ recs, double qx, double qy, double qz,
                                       Rec& leaf) {
  int node = 0;
  Rec r = load_rec(recs);
  while (r.link >= 0) {
    const double s = plane_side(qx, qy, qz, r.mx, r.my, r.mz, r.dx, r.dy, r.dz);
    node = r.link + ((s < 0.0) ? 0 : 1);
    r = load_rec(recs + node);
  }
  leaf = r;
  return node;
}

// Contribution of one correspondence to the 27 accumulators.  Returns false when gated out.
// acc layout: H lower triangle row-major (r>=c): idx = r*(r+1)/2 + c  (21 values), then b[0..5].
__device__ __forceinline__ bool linearize_one(const double* __restrict__ X, const IcpParams& P, double px, double py,
                                              double pz, double mlx, double mly, double mlz, const Rec& f,
                                              double* acc) {
  const double src_ball = P.min_ball + P.b_ratio * norm3(px, py, pz);
  const double ex = mlx - f.mx, ey = mly - f.my, ez = mlz - f.mz;
  if (norm3(ex, ey, ez) > src_ball) return false;
  const double e = dot3(ex, ey, ez, f.dx, f.dy, f.dz);
  double J[6];
  J[0] = dot3(f.dx, f.dy, f.dz, X[0], X[4], X[8]);
  J[1] = dot3(f.dx, f.dy, f.dz, X[1], X[5], X[9]);
  J[2] = dot3(f.dx, f.dy, f.dz, X[2], X[6], X[10]);
  const double n0 = -J[0], n1 = -J[1], n2 = -J[2];
  J[3] = n1 * pz + n2 * (-py);
  J[4] = n0 * (-pz) + n2 * px;
  J[5] = n0 * py + n1 * (-px);
  double scale = 1.0;
  const double chi = fabs(e);
  if (chi > P.rho_ker_sqrt) scale = P.rho_ker_sqrt / chi;
  const double w = 1.0 - f.bbox0 / P.min_ball;
  scale *= w * w;
  double sJ[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) sJ[i] = scale * J[i];
  int k = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) acc[k++] += sJ[r] * J[c];
#pragma unroll
  for (int r = 0; r < 6; ++r) acc[21 + r] += sJ[r] * e;
  return true;
}

__device__ __forceinline__ double shfl_down_f64(double v, int off) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_down_sync(0xffffffffu, lo, off);
  hi = __shfl_down_sync(0xffffffffu, hi, off);
  return __hiloint2double(hi, lo);
}

// Deterministic CTA reduction of acc[27] -> out[27] (global).  Fixed shuffle tree inside a warp,
// warps combined in warp order.
__device__ __forceinline__ void block_reduce_store(double* acc, double (*s_warp)[kAcc], double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    double v = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += shfl_down_f64(v, off);
    if (lane == 0) s_warp[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double s = s_warp[0][threadIdx.x];
#pragma unroll
    for (int w2 = 1; w2 < kWarps; ++w2) s += s_warp[w2][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// Sum `nblk` per-CTA partials (global, written by other SMs -> read with ld.cg) into s_tot[27].
__device__ __forceinline__ void final_reduce(const double* partial, int nblk, double (*s_warp)[kAcc], double* s_tot) {
  const int j = threadIdx.x & 31, g = threadIdx.x >> 5;
  __syncthreads();
  if (j < kAcc) {
    double s = 0.0;
    for (int blk = g; blk < nblk; blk += kWarps) s += __ldcg(partial + size_t(blk) * kAcc + j);
    s_warp[g][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double s = s_warp[0][threadIdx.x];
#pragma unroll
    for (int w2 = 1; w2 < kWarps; ++w2) s += s_warp[w2][threadIdx.x];
    s_tot[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void unpack_Hb(const double* tot, double* H, double* b) {
  int k = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c <= r; ++c) {
      H[r * 6 + c] = tot[k];
      H[c * 6 + r] = tot[k];
      ++k;
    }
  for (int r = 0; r < 6; ++r) b[r] = tot[21 + r];
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace madicp
