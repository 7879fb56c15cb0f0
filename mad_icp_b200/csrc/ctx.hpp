// ctx.hpp -- the context and device-tree objects behind the opaque handles of include/madicp_b200.h, shared by
// the translation units of libmadicp_b200.so (capi.cu: registration + keyframe lifecycle; gpu_tree.cu: the
// device-side MAD-tree build and ingest).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace madicp {
void set_error(const std::string& msg);

constexpr int kMaxLevels = 4096;  // deepest tree a keyframe slot / device tree can hold (level table entries)
struct Slot {  // slot s owns pool indices [s*pool_cap, (s+1)*pool_cap) and quad records [s*quad_cap, ...)
  int n_nodes = 0, n_leaves = 0, n_levels = 0;
};
constexpr size_t kMatchedCap = size_t(1) << 20;  // bytes reserved for matched flags (max moving leaves)

// One cudaMalloc, exported over CUDA IPC: mailbox + matched flags.  The flags are double-buffered by
// registration-call parity: peers store into buffer (call & 1) during their last round while the
// owner zeroes buffer ((call + 1) & 1) ahead of the NEXT call, so a zeroing can never race a peer.
struct CommBlock {
  Mailbox box;
  unsigned char matched[2][kMatchedCap];
};
}  // namespace madicp

struct madicp_ctx;

// A MAD-tree resident in device memory (sensor frame): the 64-byte breadth-first records, the level table
// and the getLeafs table.  Built on the device (gpu_tree.cu) or uploaded from a host-built tree.
struct madtree_gpu {
  madicp_ctx* ctx = nullptr;
  madtree_rec_t* recs = nullptr;  // n_nodes
  int* lvl = nullptr;             // n_levels + 1
  int* leaf_of = nullptr;         // n_leaves: getLeafs ordinal -> breadth-first index
  size_t cap_nodes = 0;           // capacity of the arrays (one allocation, carved)
  void* block = nullptr;
  int n_nodes = 0, n_leaves = 0, n_levels = 0;
  std::vector<int> h_lvl;         // host copy of the level table
  // full per-node data kept by the device build for audits (madtree_gpu_export): null for uploaded trees
  double* full = nullptr;         // n_nodes x 16: mean 3, eigenvectors 9 (column-major), bbox 3, num_points
  int64_t n_points = 0;
  uint64_t build_seq = 0;
};

struct madicp_ctx {
  int device = 0;
  int max_keyframes = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  int sm_count = 0;
  std::vector<madicp::Slot> slots;
  // keyframe pool: parallel arrays, pool_cap nodes per slot (kernels.cuh: ModelView)
  size_t pool_cap = 0;
  madtree_rec_t* d_pool_recs = nullptr;
  int* d_pool_child0 = nullptr;    // per node: first quad record of the grandchildren (even-depth internal nodes)
  int* d_pool_rec_of = nullptr;    // per node: its quad record (even-depth nodes)
  int* d_pool_lvl = nullptr;       // per slot: level table, kMaxLevels + 1 entries
  size_t quad_cap = 0;             // 4-ary records per slot (= 2 * pool_cap)
  madicp::QuadRec* d_quad = nullptr;
  double* d_pool_ww = nullptr;     // per node: planarity weight of a leaf (what the quad leaf codes also carry)
  // pinned staging rings for the small stream-ordered uploads of a promotion (pose, level table)
  static constexpr int kXformRing = 64;
  double* d_xform = nullptr;
  double* h_xform = nullptr;
  int* h_lvl = nullptr;
  cudaEvent_t xform_done[kXformRing] = {};
  uint32_t xform_seq = 0;
  std::vector<madtree_gpu*> tree_cache;  // freed device trees keep their memory for the next scan
  std::vector<void*> tree_slabs;         // the allocations the trees are carved from
  cudaEvent_t tree_free_ev = nullptr;    // recorded on the context's stream at every madtree_gpu_free
  std::mutex tree_mu;                    // ... builders on other host threads allocate from it too
  void* build_state = nullptr;           // gpu_tree.cu: working memory of the device build (lazily created)
  long long* d_dbg_cta = nullptr;  // MADICP_MAX_ITERS x grid item-phase cycles when debug timing is on
  madicp::IcpParams P{0.2, 0.31622776601683794, 0.02};
  double* d_moving = nullptr;               // raw L x 3 means as uploaded / gathered
  madicp::Moving4* d_mov4 = nullptr;        // prepared (mean, gate radius) records the kernels read
  bool mov4_stale = true;                   // params changed / new means since the last preparation
  unsigned char* d_step_matched = nullptr;  // matched flags of the step API (madicp_linearize)
  int L = 0;
  size_t cap_moving = 0;
  uint32_t call_seq = 0;  // registrations enqueued so far (selects the matched buffer)
  int* d_hit = nullptr;
  int* d_ord = nullptr;
  size_t cap_items = 0;
  double* d_cloud_q = nullptr;  // madicp_search_cloud scratch: queries (3n) + outputs (7n), ordinals
  int* d_cloud_o = nullptr;
  size_t cap_cloud = 0;
  double* d_partial = nullptr;      // per-CTA tiles of the step kernel (k_linearize)
  madicp::LLCell* d_tiles = nullptr;  // per-CTA tiles of the persistent kernel, epoch-tagged (cap_partial cells)
  size_t cap_partial = 0;
  int* d_memo_leaf = nullptr;      // path memo of the persistent kernel (GnArgs), grid x item_stride each
  float* d_memo_margin = nullptr;
  size_t cap_memo = 0;
  bool use_memo = true;            // MADICP_NO_MEMO=1 / madicp_debug_set_memo(0): walk every item in every round
  madicp::GnState* d_state = nullptr;
  double* d_X = nullptr;  // 12 (step API pose) + 36 + 6 scratch
  madicp::CommBlock* d_comm = nullptr;
  double* h_pinned = nullptr;       // 12 + 36 + 6 + ... staging
  madicp::GnState* h_state = nullptr;  // pinned mirror (results)
  unsigned char* h_matched = nullptr;
  int gn_grid = 0;
  bool gn_auto = true;  // pick the shape per launch from the item count (pick_shape)
  int gn_threads = 1024;
  const void* gn_kernel = nullptr;
  size_t gn_smem = 0;
  // one-CTA-per-SM shapes the automatic choice considers, with the cost of one full pass of each (any unit):
  // a prior until madicp_calibrate measures them on the resident workload
  static constexpr int kNumAutoShapes = 6;
  static constexpr int kAutoShapes[kNumAutoShapes] = {768, 1024, 896, 704, 640, 512};
  double pass_cost[kNumAutoShapes] = {9500.0, 11500.0, 10400.0, 9500.0, 8400.0, 7600.0};
  bool calibrated = false;
  int last_iters = 0;
  long long* d_dbg = nullptr;  // MADICP_MAX_ITERS x 8 clock stamps when debug timing is on
  std::atomic<int64_t> launches{0};
  // peers
  int rank = 0, world = 1;
  madicp::CommBlock* peer_comm[madicp::kMaxPeers] = {};
  uint32_t epoch = 0;
  uint32_t pose_epoch = 1;  // GnState::X_ll epochs (never reset: the cells are zeroed once)
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      madicp::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                       \
      return MADICP_ERR_CUDA;                                                                      \
    }                                                                                              \
  } while (0)

// No C++ exception may cross the C ABI: every entry point that can allocate host memory runs inside these.
#define MADICP_TRY try {
#define MADICP_CATCH(who)                                          \
  }                                                                \
  catch (const std::bad_alloc&) {                                  \
    madicp::set_error(std::string(who) + ": out of host memory");  \
    return MADICP_ERR_NOMEM;                                       \
  }                                                                \
  catch (const std::exception& e_) {                               \
    madicp::set_error(std::string(who) + ": " + e_.what());        \
    return MADICP_ERR_INVALID;                                     \
  }

// capi.cu
int madicp_tree_alloc(madicp_ctx* c, size_t cap_nodes, madtree_gpu** out);
// gpu_tree.cu
void madicp_gpu_build_release(madicp_ctx* c);
