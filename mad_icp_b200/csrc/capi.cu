// capi.cu -- kernels' launch side and the madicp_* C ABI (include/madicp_b200.h).
// No CPU fallback: every compute entry point launches the sm_100a kernels of kernels.cuh.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace madicp {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

// =============================================================================================
// Kernels
// =============================================================================================

// K1: one thread per (keyframe k, moving leaf q), item w = k*L + q so a warp holds 32 consecutive
// leaves (DFS order => spatially coherent) of one keyframe.
__global__ void __launch_bounds__(kBlock)
k_search(const __grid_constant__ ModelView model, const double* __restrict__ moving, int L,
         const double* __restrict__ Xp, int* __restrict__ hit, int* __restrict__ ordinals) {
  double X[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) X[i] = Xp[i];
  const int64_t total = int64_t(model.K) * L;
  for (int64_t w = int64_t(blockIdx.x) * kBlock + threadIdx.x; w < total; w += int64_t(gridDim.x) * kBlock) {
    const int k = int(w / L), q = int(w - int64_t(k) * L);
    const double px = moving[3 * q], py = moving[3 * q + 1], pz = moving[3 * q + 2];
    double mx, my, mz;
    iso_apply(X, px, py, pz, mx, my, mz);
    Rec leaf;
    const int node = descend(model.recs[k], mx, my, mz, leaf);
    if (hit) hit[w] = node;
    if (ordinals) ordinals[w] = -1 - leaf.link;
  }
}

// K2: reads K1's leaf record index, folds the 6x7 H/b tile per warp (FP64 DMMA), per-CTA partials, and
// the last CTA to arrive folds the partials in CTA order into st->H / st->b.
__global__ void __launch_bounds__(kBlock)
k_linearize(const __grid_constant__ ModelView model, const double* __restrict__ moving, int L,
            const double* __restrict__ Xp, const __grid_constant__ IcpParams P, const int* __restrict__ hit,
            unsigned char* __restrict__ matched, double* __restrict__ partial, GnState* st) {
  __shared__ double s_stage[kWarps][32 * kStage];
  __shared__ double s_red[kWarps][64];
  __shared__ double s_tot[kAcc];
  __shared__ int s_last;
  double X[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) X[i] = Xp[i];
  double c0 = 0.0, c1 = 0.0;
  const int64_t total = int64_t(model.K) * L;
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count: every lane takes part in the DMMA fold, lanes past the end stage zeros
  for (int64_t w0 = int64_t(blockIdx.x) * kBlock + (threadIdx.x - lane); w0 < total; w0 += int64_t(gridDim.x) * kBlock) {
    const int64_t w = w0 + lane;
    double v[kStage];
#pragma unroll
    for (int i = 0; i < kStage; ++i) v[i] = 0.0;
    if (w < total) {
      const int k = int(w / L), q = int(w - int64_t(k) * L);
      const double px = moving[3 * q], py = moving[3 * q + 1], pz = moving[3 * q + 2];
      double mx, my, mz;
      iso_apply(X, px, py, pz, mx, my, mz);
      const Rec f = load_rec(model.recs[k] + hit[w]);
      if (linearize_one(X, P, px, py, pz, mx, my, mz, f, v) && matched) matched[q] = 1;
    }
    warp_accumulate(s_stage[threadIdx.x >> 5], v, c0, c1);
  }
  block_reduce_store(c0, c1, s_red, partial + size_t(blockIdx.x) * kAcc);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1) == int(gridDim.x) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  final_reduce(partial, gridDim.x, s_red, s_tot);
  if (threadIdx.x == 0) {
    unpack_Hb(s_tot, st->H, st->b);
    st->ticket = 0;
  }
}

// K3: updateState for H,b already on the device (single thread; ~1 us).
__global__ void k_solve(const double* __restrict__ H, const double* __restrict__ b, double* X) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double Hl[36], bl[6], Xl[12];
    for (int i = 0; i < 36; ++i) Hl[i] = H[i];
    for (int i = 0; i < 6; ++i) bl[i] = b[i];
    for (int i = 0; i < 12; ++i) Xl[i] = X[i];
    gn_update_pose(Hl, bl, Xl, nullptr);
    for (int i = 0; i < 12; ++i) X[i] = Xl[i];
  }
}

// In-kernel all-reduce of the 48-value accumulator tile across GPUs (called by ONE CTA per rank).
// Every rank stores its partial into every rank's mailbox (own included) with 16-byte LL cells,
// then spins on its own mailbox until all `world` partials of this epoch are present and sums them
// in rank order -> identical bits on every rank.
__device__ __forceinline__ void peer_allreduce(const PeerView& pv, uint32_t epoch, double* s_tot,
                                               double (*s_stage)[kAcc]) {
  const int slot = int(epoch & 1u);
  for (int idx = threadIdx.x; idx < pv.world * kAcc; idx += kBlock) {
    const int r = idx / kAcc, i = idx - r * kAcc;
    const double v = s_tot[i];
    const uint32_t lo = uint32_t(__double2loint(v)), hi = uint32_t(__double2hiint(v));
    LLCell* dst = &pv.box[r]->cell[slot][pv.rank][i];
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(lo), "r"(epoch), "r"(hi), "r"(epoch)
                 : "memory");
  }
  for (int idx = threadIdx.x; idx < pv.world * kAcc; idx += kBlock) {
    const int r = idx / kAcc, i = idx - r * kAcc;
    const LLCell* src = &pv.box[pv.rank]->cell[slot][r][i];
    uint32_t lo, f0, hi, f1;
    do {
      asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src)
                   : "memory");
    } while (f0 != epoch || f1 != epoch);
    s_stage[r][i] = __hiloint2double(int(hi), int(lo));
  }
  __syncthreads();
  if (threadIdx.x < kAcc) {
    double s = s_stage[0][threadIdx.x];
    for (int r = 1; r < pv.world; ++r) s += s_stage[r][threadIdx.x];
    s_tot[threadIdx.x] = s;
  }
  __syncthreads();
}

struct GnArgs {
  ModelView model;
  IcpParams P;
  PeerView peers;  // world <= 1 => single GPU
  const double* moving;
  int L;
  int iters;
  unsigned char* matched;               // local matched flags (L bytes), zeroed by the host
  unsigned char* peer_matched[kMaxPeers];  // world > 1: every rank's matched array (peer mapped)
  double* partial;                      // gridDim.x * kAcc
  GnState* st;
  long long* dbg;                       // nullable: per-round SM-clock stamps (madicp_debug_timing)
};

// GN: the whole ICP loop.  Persistent cooperative grid (all CTAs co-resident); one software grid
// barrier per round: CTAs publish partials, take a ticket, the last one reduces / exchanges /
// solves and releases st->round, the others spin on it with ld.acquire.gpu.
__global__ void __launch_bounds__(kBlock, 4)
k_gn_loop(const __grid_constant__ GnArgs A) {
  // shared memory: the per-warp staging tiles are only live inside the item loop, so the reduction
  // scratch and the peer staging alias them
  __shared__ __align__(16) double s_buf[kWarps * 32 * kStage];
  __shared__ double s_tot[kAcc];
  __shared__ double s_X[12];
  __shared__ int s_last;
  __shared__ int s_count[kWarps];
  double(*s_red)[64] = reinterpret_cast<double(*)[64]>(s_buf);               // [kWarps][64]
  double(*s_peer)[kAcc] = reinterpret_cast<double(*)[kAcc]>(s_buf + kWarps * 64);  // [kMaxPeers][kAcc]
  static_assert(kWarps * 64 + kMaxPeers * kAcc <= kWarps * 32 * kStage, "scratch must fit in the staging tiles");
  GnState* st = A.st;
  const int64_t total = int64_t(A.model.K) * A.L;
  const bool multi = A.peers.world > 1;
  const int lane = threadIdx.x & 31;
  double* stage = s_buf + (threadIdx.x >> 5) * (32 * kStage);
  for (int it = 0; it < A.iters; ++it) {
    if (threadIdx.x == 0 && it > 0)
      while (ld_acquire_gpu(&st->round) < it) {}
    __syncthreads();  // also: everyone is done with s_buf/s_X of the previous round
    if (threadIdx.x < 12) s_X[threadIdx.x] = __ldcg(&st->X_trace[it * 12 + threadIdx.x]);
    __syncthreads();
    const bool last_round = (it == A.iters - 1);
    double c0 = 0.0, c1 = 0.0;
    long long t_begin = 0;
    if (A.dbg && threadIdx.x == 0) t_begin = clock64();
    for (int64_t w0 = int64_t(blockIdx.x) * kBlock + (threadIdx.x - lane); w0 < total;
         w0 += int64_t(gridDim.x) * kBlock) {
      const int64_t w = w0 + lane;
      double v[kStage];
#pragma unroll
      for (int i = 0; i < kStage; ++i) v[i] = 0.0;
      if (w < total) {
        const int k = int(w / A.L), q = int(w - int64_t(k) * A.L);
        const double px = A.moving[3 * q], py = A.moving[3 * q + 1], pz = A.moving[3 * q + 2];
        double mx, my, mz;
        iso_apply(s_X, px, py, pz, mx, my, mz);
        Rec f;
        descend(A.model.recs[k], mx, my, mz, f);
        if (linearize_one(s_X, A.P, px, py, pz, mx, my, mz, f, v) && last_round) {
          if (multi) {
            for (int r = 0; r < A.peers.world; ++r) A.peer_matched[r][q] = 1;
          } else {
            A.matched[q] = 1;
          }
        }
      }
      warp_accumulate(stage, v, c0, c1);
    }
    __syncthreads();  // staging tiles are dead; s_red aliases them
    if (A.dbg && threadIdx.x == 0 && blockIdx.x == 0) A.dbg[it * 8 + 0] = clock64() - t_begin;  // item phase, CTA 0
    block_reduce_store(c0, c1, s_red, A.partial + size_t(blockIdx.x) * kAcc);
    if (multi && last_round)
      __threadfence_system();  // matched flags stored to peers become visible before our LL cells
    else
      __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1) == (it + 1) * int(gridDim.x) - 1);
    __syncthreads();
    if (s_last) {
      long long t0 = 0, t1 = 0, t2 = 0;
      if (A.dbg && threadIdx.x == 0) {
        t0 = clock64();
        A.dbg[it * 8 + 1] = t0 - t_begin;  // round start -> last CTA arrived (that CTA's clock)
      }
      __threadfence();
      final_reduce(A.partial, gridDim.x, s_red, s_tot);
      if (A.dbg && threadIdx.x == 0) t1 = clock64();
      if (multi) {
        if (last_round) __threadfence_system();
        peer_allreduce(A.peers, A.peers.epoch_base + uint32_t(it) + 1u, s_tot, s_peer);
        if (last_round) __threadfence_system();
      }
      if (last_round) {  // count matched moving leaves (all writers are done: they took tickets)
        int c = 0;
        for (int q = threadIdx.x; q < A.L; q += kBlock) c += (__ldcv(A.matched + q) != 0);
        for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
        if (lane == 0) s_count[threadIdx.x >> 5] = c;
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        if (A.dbg) t2 = clock64();
        double H[36], b[6], Xn[12];
        unpack_Hb(s_tot, H, b);
        for (int i = 0; i < 12; ++i) Xn[i] = s_X[i];
        gn_update_pose(H, b, Xn, nullptr);
        for (int i = 0; i < 12; ++i) st->X_trace[(it + 1) * 12 + i] = Xn[i];
        if (last_round) {
          for (int i = 0; i < 36; ++i) st->H[i] = H[i];
          for (int i = 0; i < 6; ++i) st->b[i] = b[i];
          int c = 0;
          for (int w2 = 0; w2 < kWarps; ++w2) c += s_count[w2];
          st->n_matched = c;
        }
        __threadfence();
        st_release_gpu(&st->round, it + 1);
        if (A.dbg) {
          A.dbg[it * 8 + 2] = t1 - t0;         // fold of the per-CTA partials
          A.dbg[it * 8 + 3] = t2 - t1;         // peer exchange + matched count
          A.dbg[it * 8 + 4] = clock64() - t2;  // solve + pose update + publish
        }
      }
    }
  }
}

// MADtreeWrapper::searchCloud / searchCloudDist: arbitrary query points against one slot.
__global__ void __launch_bounds__(kBlock)
k_search_cloud(const madtree_rec_t* __restrict__ recs, const double* __restrict__ q, int64_t n,
               int* __restrict__ ordinals, double* __restrict__ points, double* __restrict__ normals,
               double* __restrict__ dists) {
  for (int64_t i = int64_t(blockIdx.x) * kBlock + threadIdx.x; i < n; i += int64_t(gridDim.x) * kBlock) {
    const double qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    Rec f;
    descend(recs, qx, qy, qz, f);
    if (ordinals) ordinals[i] = -1 - f.link;
    if (points) {
      points[3 * i] = f.mx; points[3 * i + 1] = f.my; points[3 * i + 2] = f.mz;
    }
    if (normals) {
      normals[3 * i] = f.dx; normals[3 * i + 1] = f.dy; normals[3 * i + 2] = f.dz;
    }
    if (dists) dists[i] = norm3(qx - f.mx, qy - f.my, qz - f.mz);
  }
}

}  // namespace madicp

// =============================================================================================
// Context
// =============================================================================================
using namespace madicp;

namespace {
struct Slot {
  madtree_rec_t* d_recs = nullptr;
  int n_nodes = 0, n_leaves = 0;
  size_t cap_nodes = 0;
};
constexpr size_t kMatchedCap = size_t(1) << 20;  // bytes reserved for matched flags (max moving leaves)

// One cudaMalloc, exported over CUDA IPC: mailbox + matched flags.  The flags are double-buffered by
// registration-call parity: peers store into buffer (call & 1) during their last round while the
// owner zeroes buffer ((call + 1) & 1) ahead of the NEXT call, so a zeroing can never race a peer.
struct CommBlock {
  Mailbox box;
  unsigned char matched[2][kMatchedCap];
};
}  // namespace

struct madicp_ctx {
  int device = 0;
  int max_keyframes = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int sm_count = 0;
  std::vector<Slot> slots;
  IcpParams P{0.2, 0.31622776601683794, 0.02};
  double* d_moving = nullptr;
  unsigned char* d_step_matched = nullptr;  // matched flags of the step API (madicp_linearize)
  int L = 0;
  size_t cap_moving = 0;
  uint32_t call_seq = 0;  // registrations enqueued so far (selects the matched buffer)
  int* d_hit = nullptr;
  int* d_ord = nullptr;
  size_t cap_items = 0;
  double* d_partial = nullptr;
  size_t cap_partial = 0;
  GnState* d_state = nullptr;
  double* d_X = nullptr;  // 12 (step API pose) + 36 + 6 scratch
  CommBlock* d_comm = nullptr;
  double* h_pinned = nullptr;  // 12 + 36 + 6 + ... staging
  GnState* h_state = nullptr;  // pinned mirror
  unsigned char* h_matched = nullptr;
  int gn_grid = 0;
  int last_iters = 0;
  long long* d_dbg = nullptr;  // MADICP_MAX_ITERS x 8 clock stamps when debug timing is on
  int64_t launches = 0;
  // peers
  int rank = 0, world = 1;
  CommBlock* peer_comm[kMaxPeers] = {};
  uint32_t epoch = 0;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                               \
      return MADICP_ERR_CUDA;                                                                      \
    }                                                                                              \
  } while (0)

static ModelView make_view(const madicp_ctx* c) {
  ModelView v;
  v.K = 0;
  for (int s = 0; s < c->max_keyframes; ++s)
    if (c->slots[s].n_nodes > 0) v.recs[v.K++] = c->slots[s].d_recs;
  for (int i = v.K; i < kMaxSlots; ++i) v.recs[i] = nullptr;
  return v;
}

static int ensure_items(madicp_ctx* c, size_t items) {
  if (items <= c->cap_items) return MADICP_OK;
  if (c->d_hit) cudaFree(c->d_hit);
  if (c->d_ord) cudaFree(c->d_ord);
  c->d_hit = c->d_ord = nullptr;
  c->cap_items = 0;
  CK(cudaMalloc(&c->d_hit, items * sizeof(int)));
  CK(cudaMalloc(&c->d_ord, items * sizeof(int)));
  c->cap_items = items;
  return MADICP_OK;
}

static int grid_for(const madicp_ctx* c, int64_t items) {
  int64_t g = (items + kBlock - 1) / kBlock;
  const int64_t cap = int64_t(c->sm_count) * 8;  // 8 CTAs of 256 threads fill an SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

extern "C" {

const char* madicp_last_error(void) { return g_error.c_str(); }
int madicp_abi_version(void) { return 1; }

int madicp_create(madicp_ctx_t** out, int device, int max_keyframes) {
  if (!out || max_keyframes < 1 || max_keyframes > kMaxSlots) {
    set_error("madicp_create: bad arguments (1 <= max_keyframes <= 64)");
    return MADICP_ERR_INVALID;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) {
    set_error(std::string("madicp_create: no usable CUDA device (") +
              (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
              "); this library has no CPU fallback");
    return MADICP_ERR_CUDA;
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("madicp_create: device is not sm_100-class; kernels are built for sm_100a only");
    return MADICP_ERR_CUDA;
  }
  madicp_ctx* c = new (std::nothrow) madicp_ctx;
  if (!c) return MADICP_ERR_NOMEM;
  c->device = device;
  c->max_keyframes = max_keyframes;
  c->sm_count = prop.multiProcessorCount;
  c->slots.resize(max_keyframes);
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaMalloc(&c->d_state, sizeof(GnState)));
  CK(cudaMemset(c->d_state, 0, sizeof(GnState)));
  CK(cudaMalloc(&c->d_X, sizeof(double) * 64));
  CK(cudaMalloc(&c->d_comm, sizeof(CommBlock)));
  CK(cudaMemset(c->d_comm, 0, sizeof(CommBlock)));
  CK(cudaMallocHost(&c->h_pinned, sizeof(double) * 64));
  CK(cudaMallocHost(&c->h_state, sizeof(GnState)));
  CK(cudaMallocHost(&c->h_matched, kMatchedCap));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gn_loop, kBlock, 0));
  if (per_sm < 1) {
    set_error("madicp_create: k_gn_loop does not fit on an SM");
    return MADICP_ERR_CUDA;
  }
  c->gn_grid = per_sm * c->sm_count;
  c->cap_partial = size_t(std::max(c->gn_grid, c->sm_count * 8)) * kAcc;
  CK(cudaMalloc(&c->d_partial, c->cap_partial * sizeof(double)));
  c->peer_comm[0] = c->d_comm;
  *out = c;
  return MADICP_OK;
}

void madicp_destroy(madicp_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int r = 0; r < c->world; ++r)
    if (c->world > 1 && r != c->rank && c->peer_comm[r]) cudaIpcCloseMemHandle(c->peer_comm[r]);
  for (Slot& s : c->slots)
    if (s.d_recs) cudaFree(s.d_recs);
  cudaFree(c->d_moving);
  cudaFree(c->d_step_matched);
  cudaFree(c->d_hit);
  cudaFree(c->d_ord);
  cudaFree(c->d_partial);
  cudaFree(c->d_state);
  cudaFree(c->d_X);
  cudaFree(c->d_comm);
  cudaFree(c->d_dbg);
  cudaFreeHost(c->h_pinned);
  cudaFreeHost(c->h_state);
  cudaFreeHost(c->h_matched);
  cudaStreamDestroy(c->own_stream);
  delete c;
}

int madicp_set_params(madicp_ctx_t* c, double min_ball, double rho_ker, double b_ratio) {
  if (!c || !(min_ball > 0) || rho_ker < 0) {
    set_error("madicp_set_params: bad arguments");
    return MADICP_ERR_INVALID;
  }
  c->P.min_ball = min_ball;
  c->P.rho_ker_sqrt = sqrt(rho_ker);
  c->P.b_ratio = b_ratio;
  return MADICP_OK;
}

int madicp_set_stream(madicp_ctx_t* c, void* s) {
  if (!c) return MADICP_ERR_INVALID;
  c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
  return MADICP_OK;
}
void* madicp_get_stream(const madicp_ctx_t* c) { return c ? static_cast<void*>(c->stream) : nullptr; }

int madicp_put_keyframe_records(madicp_ctx_t* c, int slot, const madtree_rec_t* recs, int n_nodes, int n_leaves) {
  if (!c || !recs || slot < 0 || slot >= c->max_keyframes || n_nodes < 1 || n_leaves < 1) {
    set_error("madicp_put_keyframe: bad arguments");
    return MADICP_ERR_INVALID;
  }
  CK(cudaSetDevice(c->device));
  Slot& s = c->slots[slot];
  if (size_t(n_nodes) > s.cap_nodes) {
    CK(cudaStreamSynchronize(c->stream));
    if (s.d_recs) cudaFree(s.d_recs);
    s.d_recs = nullptr;
    s.cap_nodes = 0;
    CK(cudaMalloc(&s.d_recs, size_t(n_nodes) * sizeof(madtree_rec_t)));
    s.cap_nodes = n_nodes;
  }
  CK(cudaMemcpyAsync(s.d_recs, recs, size_t(n_nodes) * sizeof(madtree_rec_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // caller may free/modify the host tree right after
  s.n_nodes = n_nodes;
  s.n_leaves = n_leaves;
  return MADICP_OK;
}

int madicp_put_keyframe(madicp_ctx_t* c, int slot, const madtree_t* tree) {
  if (!tree) {
    set_error("madicp_put_keyframe: null tree");
    return MADICP_ERR_INVALID;
  }
  return madicp_put_keyframe_records(c, slot, madtree_records(tree), madtree_num_nodes(tree), madtree_num_leaves(tree));
}

int madicp_drop_keyframe(madicp_ctx_t* c, int slot) {
  if (!c || slot < 0 || slot >= c->max_keyframes) return MADICP_ERR_INVALID;
  c->slots[slot].n_nodes = 0;  // memory is kept for reuse by the next keyframe in this slot
  c->slots[slot].n_leaves = 0;
  return MADICP_OK;
}

int madicp_num_keyframes(const madicp_ctx_t* c) {
  if (!c) return MADICP_ERR_INVALID;
  int k = 0;
  for (const Slot& s : c->slots) k += (s.n_nodes > 0);
  return k;
}
int madicp_active_slots(const madicp_ctx_t* c, int* out, int cap) {
  if (!c) return MADICP_ERR_INVALID;
  int k = 0;
  for (int s = 0; s < c->max_keyframes; ++s)
    if (c->slots[s].n_nodes > 0) {
      if (out && k < cap) out[k] = s;
      ++k;
    }
  return k;
}
int madicp_keyframe_leaves(const madicp_ctx_t* c, int slot) {
  if (!c || slot < 0 || slot >= c->max_keyframes || c->slots[slot].n_nodes == 0) return MADICP_ERR_INVALID;
  return c->slots[slot].n_leaves;
}
int64_t madicp_model_nodes(const madicp_ctx_t* c) {
  if (!c) return MADICP_ERR_INVALID;
  int64_t n = 0;
  for (const Slot& s : c->slots) n += s.n_nodes;
  return n;
}
int64_t madicp_kernel_launches(const madicp_ctx_t* c) { return c ? c->launches : 0; }

int madicp_set_moving(madicp_ctx_t* c, const double* means, int L) {
  if (!c || !means || L < 1 || size_t(L) > kMatchedCap) {
    set_error("madicp_set_moving: bad arguments (1 <= L <= 1048576)");
    return MADICP_ERR_INVALID;
  }
  CK(cudaSetDevice(c->device));
  if (size_t(L) > c->cap_moving) {
    CK(cudaStreamSynchronize(c->stream));
    if (c->d_moving) cudaFree(c->d_moving);
    if (c->d_step_matched) cudaFree(c->d_step_matched);
    c->d_moving = nullptr;
    c->d_step_matched = nullptr;
    c->cap_moving = 0;
    const size_t cap = size_t(L) + size_t(L) / 4 + 1024;
    CK(cudaMalloc(&c->d_moving, cap * 3 * sizeof(double)));
    CK(cudaMalloc(&c->d_step_matched, cap));
    c->cap_moving = cap;
  }
  CK(cudaMemcpyAsync(c->d_moving, means, size_t(L) * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->L = L;
  return MADICP_OK;
}

static int check_ready(madicp_ctx* c, const char* who) {
  if (!c) return MADICP_ERR_INVALID;
  if (c->L < 1 || !c->d_moving) {
    set_error(std::string(who) + ": no moving leaves (call madicp_set_moving first)");
    return MADICP_ERR_STATE;
  }
  if (madicp_num_keyframes(c) < 1 && c->world <= 1) {
    set_error(std::string(who) + ": no keyframe uploaded");
    return MADICP_ERR_STATE;
  }
  return MADICP_OK;
}

static int launch_search(madicp_ctx* c, const ModelView& mv, const double* d_X, bool want_ord) {
  const int64_t items = int64_t(mv.K) * c->L;
  int rc = ensure_items(c, size_t(items));
  if (rc) return rc;
  k_search<<<grid_for(c, items), kBlock, 0, c->stream>>>(mv, c->d_moving, c->L, d_X, c->d_hit,
                                                        want_ord ? c->d_ord : nullptr);
  c->launches++;
  CK(cudaGetLastError());
  return MADICP_OK;
}

int madicp_search(madicp_ctx_t* c, const double X[12], int32_t* out) {
  int rc = check_ready(c, "madicp_search");
  if (rc) return rc;
  if (!X || !out) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  const ModelView mv = make_view(c);
  if (mv.K < 1) {
    set_error("madicp_search: no keyframe on this device");
    return MADICP_ERR_STATE;
  }
  memcpy(c->h_pinned, X, 12 * sizeof(double));
  CK(cudaMemcpyAsync(c->d_X, c->h_pinned, 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  rc = launch_search(c, mv, c->d_X, true);
  if (rc) return rc;
  CK(cudaMemcpyAsync(out, c->d_ord, size_t(mv.K) * c->L * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return MADICP_OK;
}

int madicp_linearize(madicp_ctx_t* c, const double X[12], double H[36], double b[6], uint8_t* matched) {
  int rc = check_ready(c, "madicp_linearize");
  if (rc) return rc;
  if (!X || !H || !b) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  const ModelView mv = make_view(c);
  if (mv.K < 1) {
    set_error("madicp_linearize: no keyframe on this device");
    return MADICP_ERR_STATE;
  }
  memcpy(c->h_pinned, X, 12 * sizeof(double));
  CK(cudaMemcpyAsync(c->d_X, c->h_pinned, 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  rc = launch_search(c, mv, c->d_X, false);
  if (rc) return rc;
  CK(cudaMemsetAsync(c->d_state, 0, 16, c->stream));
  CK(cudaMemsetAsync(c->d_step_matched, 0, size_t(c->L), c->stream));
  const int64_t items = int64_t(mv.K) * c->L;
  const int grid = grid_for(c, items);
  k_linearize<<<grid, kBlock, 0, c->stream>>>(mv, c->d_moving, c->L, c->d_X, c->P, c->d_hit, c->d_step_matched,
                                             c->d_partial, c->d_state);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c->h_state, c->d_state, offsetof(GnState, X_trace), cudaMemcpyDeviceToHost, c->stream));
  if (matched) CK(cudaMemcpyAsync(c->h_matched, c->d_step_matched, size_t(c->L), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(H, c->h_state->H, sizeof(double) * 36);
  memcpy(b, c->h_state->b, sizeof(double) * 6);
  if (matched) memcpy(matched, c->h_matched, size_t(c->L));
  return MADICP_OK;
}

int madicp_solve_update(madicp_ctx_t* c, const double H[36], const double b[6], double X[12]) {
  if (!c || !H || !b || !X) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  memcpy(c->h_pinned, X, 12 * sizeof(double));
  memcpy(c->h_pinned + 12, H, 36 * sizeof(double));
  memcpy(c->h_pinned + 48, b, 6 * sizeof(double));
  CK(cudaMemcpyAsync(c->d_X, c->h_pinned, 54 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_solve<<<1, 32, 0, c->stream>>>(c->d_X + 12, c->d_X + 48, c->d_X);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c->h_pinned, c->d_X, 12 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(X, c->h_pinned, 12 * sizeof(double));
  return MADICP_OK;
}

int madicp_register_async(madicp_ctx_t* c, int iters, const double X0[12]) {
  int rc = check_ready(c, "madicp_register");
  if (rc) return rc;
  if (!X0 || iters < 1 || iters > MADICP_MAX_ITERS) {
    set_error("madicp_register: bad arguments (1 <= iters <= 64)");
    return MADICP_ERR_INVALID;
  }
  CK(cudaSetDevice(c->device));
  GnArgs A;
  A.model = make_view(c);
  A.P = c->P;
  A.peers.rank = c->rank;
  A.peers.world = c->world;
  A.peers.epoch_base = c->epoch;
  const int mb = int(c->call_seq & 1u);
  for (int r = 0; r < kMaxPeers; ++r) {
    A.peers.box[r] = (r < c->world) ? &c->peer_comm[r]->box : nullptr;
    A.peer_matched[r] = (r < c->world) ? c->peer_comm[r]->matched[mb] : nullptr;
  }
  A.moving = c->d_moving;
  A.L = c->L;
  A.iters = iters;
  A.matched = c->d_comm->matched[mb];
  A.partial = c->d_partial;
  A.st = c->d_state;
  A.dbg = c->d_dbg;
  c->epoch += uint32_t(iters);
  // control words + initial pose in one small pinned H2D copy
  GnState* hs = c->h_state;
  hs->ticket = 0;
  hs->round = 0;
  hs->n_matched = 0;
  hs->pad = 0;
  CK(cudaMemcpyAsync(c->d_state, hs, 16, cudaMemcpyHostToDevice, c->stream));
  memcpy(c->h_pinned, X0, 12 * sizeof(double));
  CK(cudaMemcpyAsync(c->d_state->X_trace, c->h_pinned, 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  // zero the flags buffer of the NEXT call (nobody can be writing it yet; see CommBlock)
  CK(cudaMemsetAsync(c->d_comm->matched[mb ^ 1], 0, std::min(kMatchedCap, c->cap_moving), c->stream));
  void* args[] = {&A};
  CK(cudaLaunchCooperativeKernel((void*) k_gn_loop, dim3(c->gn_grid), dim3(kBlock), args, 0, c->stream));
  c->launches++;
  c->last_iters = iters;
  c->call_seq++;
  return MADICP_OK;
}

int madicp_register_fetch(madicp_ctx_t* c, double X[12], double H[36], double b[6], uint8_t* matched, int* n_matched) {
  if (!c || c->last_iters < 1) {
    set_error("madicp_register_fetch: nothing was enqueued");
    return MADICP_ERR_STATE;
  }
  CK(cudaSetDevice(c->device));
  const int it = c->last_iters;
  CK(cudaMemcpyAsync(c->h_state, c->d_state, offsetof(GnState, X_trace), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(c->h_state->X_trace + it * 12, c->d_state->X_trace + it * 12, 12 * sizeof(double),
                     cudaMemcpyDeviceToHost, c->stream));
  if (matched)
    CK(cudaMemcpyAsync(c->h_matched, c->d_comm->matched[(c->call_seq - 1u) & 1u], size_t(c->L), cudaMemcpyDeviceToHost,
                       c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (X) memcpy(X, c->h_state->X_trace + it * 12, 12 * sizeof(double));
  if (H) memcpy(H, c->h_state->H, 36 * sizeof(double));
  if (b) memcpy(b, c->h_state->b, 6 * sizeof(double));
  if (matched) memcpy(matched, c->h_matched, size_t(c->L));
  if (n_matched) *n_matched = c->h_state->n_matched;
  return MADICP_OK;
}

int madicp_register(madicp_ctx_t* c, int iters, double X[12], double H[36], double b[6], uint8_t* matched,
                    int* n_matched) {
  int rc = madicp_register_async(c, iters, X);
  if (rc) return rc;
  return madicp_register_fetch(c, X, H, b, matched, n_matched);
}

int madicp_register_trace(madicp_ctx_t* c, double* X_trace, int max_rounds) {
  if (!c || !X_trace || c->last_iters < 1) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  int rows = c->last_iters + 1;
  if (rows > max_rounds) rows = max_rounds;
  CK(cudaMemcpyAsync(c->h_state->X_trace, c->d_state->X_trace, size_t(rows) * 12 * sizeof(double),
                     cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(X_trace, c->h_state->X_trace, size_t(rows) * 12 * sizeof(double));
  return rows;
}

int madicp_search_cloud(madicp_ctx_t* c, int slot, const double* q, int64_t n, int32_t* ordinals, double* points,
                        double* normals, double* dists) {
  if (!c || !q || n < 1 || slot < 0 || slot >= c->max_keyframes || c->slots[slot].n_nodes == 0) {
    set_error("madicp_search_cloud: bad arguments or empty slot");
    return MADICP_ERR_INVALID;
  }
  CK(cudaSetDevice(c->device));
  double *d_q = nullptr, *d_out = nullptr;
  int* d_o = nullptr;
  CK(cudaMalloc(&d_q, size_t(n) * 3 * sizeof(double)));
  CK(cudaMalloc(&d_out, size_t(n) * 7 * sizeof(double)));
  CK(cudaMalloc(&d_o, size_t(n) * sizeof(int)));
  CK(cudaMemcpyAsync(d_q, q, size_t(n) * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double* d_p = d_out;
  double* d_n = d_out + size_t(n) * 3;
  double* d_d = d_out + size_t(n) * 6;
  k_search_cloud<<<grid_for(c, n), kBlock, 0, c->stream>>>(c->slots[slot].d_recs, d_q, n, d_o, points ? d_p : nullptr,
                                                          normals ? d_n : nullptr, dists ? d_d : nullptr);
  c->launches++;
  CK(cudaGetLastError());
  if (ordinals) CK(cudaMemcpyAsync(ordinals, d_o, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (points) CK(cudaMemcpyAsync(points, d_p, size_t(n) * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (normals) CK(cudaMemcpyAsync(normals, d_n, size_t(n) * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (dists) CK(cudaMemcpyAsync(dists, d_d, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_q);
  cudaFree(d_out);
  cudaFree(d_o);
  return MADICP_OK;
}

// ------------------------------------------------------------------------------ multi-GPU
int madicp_comm_export(madicp_ctx_t* c, void* handle_out) {
  if (!c || !handle_out) return MADICP_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == MADICP_IPC_HANDLE_BYTES, "IPC handle size");
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->d_comm);
  if (e != cudaSuccess) {
    set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    return MADICP_ERR_COMM;
  }
  memcpy(handle_out, &h, sizeof(h));
  return MADICP_OK;
}

int madicp_comm_connect(madicp_ctx_t* c, int rank, int world, const void* all_handles) {
  if (!c || !all_handles || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) {
    set_error("madicp_comm_connect: bad arguments (world <= 16)");
    return MADICP_ERR_INVALID;
  }
  CK(cudaSetDevice(c->device));
  const char* hs = static_cast<const char*>(all_handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->peer_comm[r] = c->d_comm;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + size_t(r) * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error(std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
      return MADICP_ERR_COMM;
    }
    c->peer_comm[r] = static_cast<CommBlock*>(p);
  }
  c->rank = rank;
  c->world = world;
  c->epoch = 0;
  return MADICP_OK;
}

int madicp_comm_world(const madicp_ctx_t* c) { return c ? c->world : MADICP_ERR_INVALID; }

// ------------------------------------------------------------------------------ debug
int madicp_debug_timing(madicp_ctx_t* c, int enable, int64_t* out, int max_rounds) {
  if (!c) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  int rows = 0;
  if (out && c->d_dbg) {
    rows = std::min(max_rounds, c->last_iters);
    CK(cudaMemcpy(out, c->d_dbg, size_t(rows) * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  if (enable && !c->d_dbg) {
    CK(cudaMalloc(&c->d_dbg, MADICP_MAX_ITERS * 8 * sizeof(long long)));
    CK(cudaMemset(c->d_dbg, 0, MADICP_MAX_ITERS * 8 * sizeof(long long)));
  } else if (!enable && c->d_dbg) {
    cudaFree(c->d_dbg);
    c->d_dbg = nullptr;
  }
  return rows;
}

int madicp_set_gn_grid(madicp_ctx_t* c, int ctas_per_sm) {
  if (!c || ctas_per_sm < 1) return MADICP_ERR_INVALID;
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gn_loop, kBlock, 0));
  if (ctas_per_sm > per_sm) ctas_per_sm = per_sm;
  c->gn_grid = ctas_per_sm * c->sm_count;
  return ctas_per_sm;
}

}  // extern "C"
