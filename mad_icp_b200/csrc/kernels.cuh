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
// Memory/branch bound FP64 work (no tcgen05: there is no dense contraction; the only tensor-pipe
// use is the register-saving FP64 DMMA fold of the per-correspondence outer products, see
// warp_accumulate).  A node visit is one 64-byte record = two 256-bit read-only loads
// (LDG.E.256); moving leaves are laid out in getLeafs (DFS) order so the lanes of a warp walk
// nearly the same path and their loads coalesce / hit L1 at the top levels.
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
constexpr int kAcc = 48;           // 6 rows x 8 cols of the accumulator tile: H(r,c) at r*8+c, b(r) at r*8+6
constexpr int kStage = 13;         // doubles staged per correspondence: sJ[6], J[6], e
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
  double H[36];   // last round; H[r*6+c] = sum (scale*J_r)*J_c, both triangles accumulated independently
  double b[6];
  double X_trace[(MADICP_MAX_ITERS + 1) * 12];  // pose before round i; [iters] = final pose
};

// LL-style mailbox cell: a double split into two 32-bit halves, each paired with a 32-bit epoch
// flag, written with ONE 16-byte store so data and flags arrive together (no fence on the wire).
struct __align__(16) LLCell {
  uint32_t lo, flag_lo, hi, flag_hi;
};
struct Mailbox {  // lives on every rank; cell [slot][src_rank][i] is written by src_rank
  LLCell cell[kMailboxSlots][kMaxPeers][kAcc];
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

// One correspondence (reference: odometry/mad_icp.cpp:81-101): gate, error, Jacobian, Huber scale,
// planarity weight.  Fills v = {sJ[0..5] = scale*J, J[0..5], e}; returns false (v untouched) when
// the gate rejects the pair.  FP64, no FMA, operand order as arith.h.
__device__ __forceinline__ bool linearize_one(const double* __restrict__ X, const IcpParams& P, double px, double py,
                                              double pz, double mlx, double mly, double mlz, const Rec& f, double* v) {
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
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    v[i] = scale * J[i];
    v[6 + i] = J[i];
  }
  v[12] = e;
  return true;
}

// H += sJ^T J, b += sJ^T e for the 32 correspondences a warp holds, on the FP64 tensor pipe:
// D(8x8) += A(8x4) * B(4x8) with A[r][k] = sJ_r(item k), B[k][c] = J_c(item k) (c<6), e(item k)
// (c==6), zero padding elsewhere -- 8 DMMA.8x8x4 per 32 items.  The point is not FLOPs (there are
// few) but registers: the running sums are the 2-double C fragment instead of 42 scalars per
// thread, which is what lets the descent run at 4 CTAs/SM.  H(r,c) = sum (scale*J_r)*J_c is formed
// for both triangles independently, like the reference's `scale * J.transpose() * J`.
// stage: this warp's [32][kStage] doubles in shared memory.  v: this lane's 13 values (zeros if
// the lane has no correspondence).
__device__ __forceinline__ void warp_accumulate(double* stage, const double* v, double& c0, double& c1) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < kStage; ++i) stage[lane * kStage + i] = v[i];
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const double* it = stage + (4 * s + t) * kStage;
    const double a = (g < 6) ? it[g] : 0.0;                      // rows 6,7 of A are padding
    const double b = (g < 7) ? it[6 + (g < 7 ? g : 6)] : 0.0;    // col 6 of B = e, col 7 padding
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
  }
  __syncwarp();
}

// Deterministic CTA reduction of the warps' C fragments -> out[kAcc] (global): warps are combined
// in warp order.  s_red: [kWarps][64] doubles.
__device__ __forceinline__ void block_reduce_store(double c0, double c1, double (*s_red)[64], double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  s_red[warp][g * 8 + 2 * t] = c0;
  s_red[warp][g * 8 + 2 * t + 1] = c1;
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double s = s_red[0][threadIdx.x];
#pragma unroll
    for (int w2 = 1; w2 < kWarps; ++w2) s += s_red[w2][threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// Sum `nblk` per-CTA partials (global, written by other SMs -> read with ld.cg) into s_tot[kAcc]:
// four interleaved strands over the CTA index, combined in strand order.
__device__ __forceinline__ void final_reduce(const double* partial, int nblk, double (*s_red)[64], double* s_tot) {
  const int j = threadIdx.x & 63, g = threadIdx.x >> 6;
  __syncthreads();
  if (j < kAcc) {
    double s = 0.0;
    for (int blk = g; blk < nblk; blk += kBlock / 64) s += __ldcg(partial + size_t(blk) * kAcc + j);
    s_red[g][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double s = s_red[0][threadIdx.x];
#pragma unroll
    for (int w2 = 1; w2 < kBlock / 64; ++w2) s += s_red[w2][threadIdx.x];
    s_tot[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void unpack_Hb(const double* tot, double* H, double* b) {
  for (int r = 0; r < 6; ++r) {
    for (int c = 0; c < 6; ++c) H[r * 6 + c] = tot[r * 8 + c];
    b[r] = tot[r * 8 + 6];
  }
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
