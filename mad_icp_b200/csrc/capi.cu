// capi.cu -- kernels' launch side and the madicp_* C ABI (include/madicp_b200.h).
// No CPU fallback: every compute entry point launches the sm_100a kernels of kernels.cuh.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "device_kernels.cuh"

namespace madicp {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

}  // namespace madicp

// =============================================================================================
// Context
// =============================================================================================
using namespace madicp;

namespace {
struct Slot {  // slot s owns pool indices [s*pool_cap, (s+1)*pool_cap) and heap positions [s*heap_cap, ...)
  int n_nodes = 0, n_leaves = 0;
  std::vector<int> heap_pos;    // node -> position in the implicit heap (kept to re-home the slot on growth)
  std::vector<int> quad_pos;    // node -> 4-ary record * 4 + slot
  std::vector<int> quad_child;  // even-depth node -> first record of its grandchildren
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
  // keyframe pool: three parallel arrays, pool_cap nodes per slot (kernels.cuh: ModelView)
  size_t pool_cap = 0;
  madtree_rec_t* d_pool_recs = nullptr;
  int* d_pool_links = nullptr;
  size_t heap_cap = 0;  // heap positions per slot: 2^(depth+1) + skew
  FastRec* d_heap = nullptr;
  int* d_bfs_of = nullptr;
  FastRec* d_pool_fast = nullptr;  // breadth-first copy of the shadows (walk_mode 0)
  size_t quad_cap = 0;             // 4-ary records per slot (= 2 * pool_cap)
  QuadRec* d_quad = nullptr;
  int walk_mode = 4;
  bool heap_ok = true;  // false once a keyframe deeper than the implicit-heap limit has been seen
  long long* d_dbg_cta = nullptr;  // MADICP_MAX_ITERS x grid item-phase cycles when debug timing is on
  int* d_heap_pos = nullptr;  // upload scratch, pool_cap ints
  IcpParams P{0.2, 0.31622776601683794, 0.02};
  double* d_moving = nullptr;               // raw L x 3 means as uploaded
  Moving4* d_mov4 = nullptr;                // prepared (mean, gate radius) records the kernels read
  bool mov4_stale = true;                   // params changed / new means since the last preparation
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
  GnState* h_state = nullptr;  // pinned mirror (results)
  // pinned ring of launch headers (control words + initial pose): a header may only be rewritten once the
  // copy that reads it has executed, so back-to-back asynchronous registrations stay correct
  static constexpr int kInRing = 16;
  unsigned char* h_in = nullptr;
  cudaEvent_t in_done[kInRing] = {};
  unsigned char* h_matched = nullptr;
  int gn_grid = 0;
  bool gn_auto = true;  // pick the shape per launch from the item count (see pick_shape)
  int gn_threads = 1024;
  const void* gn_kernel = nullptr;
  size_t gn_smem = 0;
  int last_iters = 0;
  long long* d_dbg = nullptr;  // MADICP_MAX_ITERS x 8 clock stamps when debug timing is on
  int64_t launches = 0;
  // peers
  int rank = 0, world = 1;
  CommBlock* peer_comm[kMaxPeers] = {};
  uint32_t epoch = 0;
  uint32_t pose_epoch = 1;  // GnState::X_ll epochs (never reset: the cells are zeroed once)
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
  v.recs = c->d_pool_recs;
  v.links = c->d_pool_links;
  v.heap = c->d_heap;
  v.bfs_of = c->d_bfs_of;
  v.fast = c->d_pool_fast;
  v.quad = c->d_quad;
  v.walk_mode = c->walk_mode;
  v.K = 0;
  for (int s = 0; s < c->max_keyframes; ++s)
    if (c->slots[s].n_nodes > 0) {
      v.root[v.K] = int(size_t(s) * c->heap_cap);
      v.broot[v.K] = int(size_t(s) * c->pool_cap);
      v.qroot[v.K] = int(size_t(s) * c->quad_cap);
      ++v.K;
    }
  for (int i = v.K; i < kMaxSlots; ++i) v.root[i] = v.broot[i] = v.qroot[i] = 0;
  return v;
}

// (Re)builds the shadows (heap order), the heap->record map and the absolute links of slot `s` from
// its exact records in the pool.
static int prepare_slot(madicp_ctx* c, int s) {
  const int n = c->slots[s].n_nodes;
  const size_t off = size_t(s) * c->pool_cap, hoff = size_t(s) * c->heap_cap;
  CK(cudaMemcpyAsync(c->d_heap_pos, c->slots[s].heap_pos.data(), size_t(n) * sizeof(int), cudaMemcpyHostToDevice,
                     c->stream));
  CK(cudaMemcpyAsync(c->d_heap_pos + c->pool_cap, c->slots[s].quad_pos.data(), size_t(n) * sizeof(int),
                     cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_heap_pos + 2 * c->pool_cap, c->slots[s].quad_child.data(), size_t(n) * sizeof(int),
                     cudaMemcpyHostToDevice, c->stream));
  k_prepare_slot<<<(n + kStepBlock - 1) / kStepBlock, kStepBlock, 0, c->stream>>>(
      c->d_pool_recs + off, c->d_heap_pos, c->d_heap_pos + c->pool_cap, c->d_heap_pos + 2 * c->pool_cap, n, int(off),
      int(hoff), c->P.min_ball, c->d_pool_links + off, c->d_heap,
      c->d_bfs_of, c->d_pool_fast + off, c->d_quad, int(size_t(s) * c->quad_cap));
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));  // heap_pos scratch is reused by the next slot
  return MADICP_OK;
}

// Makes every slot at least `need` nodes large and the heap at least `need_heap` positions per slot.
// Growing re-homes the resident keyframes (device to device) and rebuilds their shadows; it only
// happens when a larger or deeper tree than any before shows up.
static int ensure_pool(madicp_ctx* c, size_t need, size_t need_heap) {
  const bool grow_pool = need > c->pool_cap, grow_heap = need_heap > c->heap_cap;
  if (!grow_pool && !grow_heap) return MADICP_OK;
  CK(cudaStreamSynchronize(c->stream));
  if (grow_pool) {
    // slot stride = 2^n + 40 nodes: a power-of-two stride would put the roots and upper levels of all
    // keyframes (the hottest lines of every walk) on the same cache sets
    size_t cap = size_t(1) << 16;
    while (cap + 40 < need || cap + 40 <= c->pool_cap) cap <<= 1;
    cap += 40;
    if (cap * size_t(c->max_keyframes) > size_t(0x7fffffff)) {
      set_error("keyframe pool would exceed 2^31 nodes");
      return MADICP_ERR_NOMEM;
    }
    madtree_rec_t* recs = nullptr;
    int* links = nullptr;
    int* hp = nullptr;
    FastRec* fast = nullptr;
    const size_t total = cap * size_t(c->max_keyframes);
    CK(cudaMalloc(&recs, total * sizeof(madtree_rec_t)));
    CK(cudaMalloc(&links, total * sizeof(int)));
    CK(cudaMalloc(&fast, total * sizeof(FastRec)));
    CK(cudaMalloc(&hp, 3 * cap * sizeof(int)));
    for (int s = 0; s < c->max_keyframes; ++s)
      if (c->slots[s].n_nodes > 0)
        CK(cudaMemcpyAsync(recs + size_t(s) * cap, c->d_pool_recs + size_t(s) * c->pool_cap,
                           size_t(c->slots[s].n_nodes) * sizeof(madtree_rec_t), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_pool_recs);
    cudaFree(c->d_pool_links);
    cudaFree(c->d_heap_pos);
    cudaFree(c->d_pool_fast);
    cudaFree(c->d_quad);
    c->d_quad = nullptr;
    CK(cudaMalloc(&c->d_quad, 2 * cap * size_t(c->max_keyframes) * sizeof(QuadRec)));
    c->quad_cap = 2 * cap;
    c->d_pool_fast = fast;
    c->d_pool_recs = recs;
    c->d_pool_links = links;
    c->d_heap_pos = hp;
    c->pool_cap = cap;
  }
  if (grow_heap) {
    size_t cap = size_t(1) << 19;  // depth 18
    while (cap < need_heap) cap <<= 1;
    cap += 40;
    if (cap * size_t(c->max_keyframes) > size_t(0x7fffffff)) {
      set_error("keyframe tree too deep for the implicit-heap layout (depth limit reached)");
      return MADICP_ERR_NOMEM;
    }
    cudaFree(c->d_heap);
    cudaFree(c->d_bfs_of);
    c->d_heap = nullptr;
    c->d_bfs_of = nullptr;
    const size_t total = cap * size_t(c->max_keyframes);
    CK(cudaMalloc(&c->d_heap, total * sizeof(FastRec)));
    CK(cudaMalloc(&c->d_bfs_of, total * sizeof(int)));
    CK(cudaMemsetAsync(c->d_heap, 0, total * sizeof(FastRec), c->stream));  // never-visited positions are prefetched only
    c->heap_cap = cap;
  }
  for (int s = 0; s < c->max_keyframes; ++s)
    if (c->slots[s].n_nodes > 0) {
      int rc = prepare_slot(c, s);
      if (rc) return rc;
    }
  return MADICP_OK;
}

// index of `slot` among the active slots (the k of ModelView::root[k])
static int slot_rank(const madicp_ctx* c, int slot) {
  int k = 0;
  for (int s = 0; s < slot; ++s) k += (c->slots[s].n_nodes > 0);
  return k;
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

// Persistent-kernel shapes: (threads per CTA, CTAs per SM) -> an instantiation; the pair fixes the
// register budget (64K registers / (THREADS*CTAS)).  Selected at create time (default 1024x1, or env
// MADICP_GN_SHAPE="threads,ctas") and through madicp_set_gn_grid.
struct GnShape {
  int threads, ctas;
  const void* fn;
  size_t smem;
};
template <int THREADS, int CTAS>
static GnShape gn_shape() {
  return GnShape{THREADS, CTAS, reinterpret_cast<const void*>(k_gn_loop<THREADS, CTAS>), gn_dynamic_smem<THREADS>()};
}
static const GnShape* gn_shapes(int* n) {
  static const GnShape table[] = {
      gn_shape<1024, 1>(), gn_shape<896, 1>(), gn_shape<768, 1>(), gn_shape<704, 1>(), gn_shape<640, 1>(),
      gn_shape<512, 1>(),  gn_shape<512, 2>(),
      gn_shape<256, 2>(),  gn_shape<256, 3>(), gn_shape<256, 4>(),
  };
  *n = int(sizeof(table) / sizeof(table[0]));
  return table;
}
static int configure_gn(madicp_ctx* c, int threads, int ctas) {
  int n = 0;
  const GnShape* t = gn_shapes(&n);
  for (int i = 0; i < n; ++i)
    if (t[i].threads == threads && t[i].ctas == ctas) {
      CK(cudaFuncSetAttribute(t[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(t[i].smem + kGnMapMaxBytes)));
      // ask for the smallest shared-memory carve-out that fits: the rest of the 228 KB is L1 for the tree
      CK(cudaFuncSetAttribute(t[i].fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                              int((t[i].smem * size_t(ctas) + 2048) * 100 / (228 * 1024)) + 1));
      int per_sm = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, t[i].fn, threads, t[i].smem));
      if (per_sm < ctas) {
        set_error("persistent kernel shape does not fit on an SM");
        return MADICP_ERR_CUDA;
      }
      c->gn_threads = threads;
      c->gn_grid = ctas * c->sm_count;
      c->gn_kernel = t[i].fn;
      c->gn_smem = t[i].smem;
      return MADICP_OK;
    }
  set_error("unsupported persistent-kernel shape (threads per CTA, CTAs per SM)");
  return MADICP_ERR_INVALID;
}

// The item phase of a round costs (passes) x (time of one pass); a pass walks one warp-item per
// resident warp and its time grows with the number of resident warps (L1 contention, and fewer
// registers per thread).  Measured on B200 at cfg3 (profiles/r01zg_probe_shapes.txt), cycles per
// full pass: 512 threads 7.6k, 640: 8.4k, 704: 9.5k, 768: 9.5k, 896: 10.4k, 1024: 11.5k.  With W warps
// per SM and n warp-items per SM the passes are ceil(n / W): pick the one-CTA-per-SM shape that minimises
// the product (ties go to the earlier entry).
static int pick_shape(madicp_ctx* c, int64_t items) {
  if (!c->gn_auto) return MADICP_OK;
  const double per_sm = double((items + 31) / 32) / double(c->sm_count);
  int best = 1024;
  double best_cost = 1e300;
  static const struct { int threads; double pass_cycles; } kShapes[] = {
      {768, 9500.0}, {1024, 11500.0}, {896, 10400.0}, {704, 9500.0}, {640, 8400.0}, {512, 7600.0}};
  for (const auto& sh : kShapes) {
    const double passes = ceil(per_sm / double(sh.threads / 32));
    const double cost = passes * sh.pass_cycles;
    if (cost < best_cost) {
      best_cost = cost;
      best = sh.threads;
    }
  }
  if (best == c->gn_threads && c->gn_grid == c->sm_count) return MADICP_OK;
  return configure_gn(c, best, 1);
}

static int grid_for(const madicp_ctx* c, int64_t items) {
  int64_t g = (items + kStepBlock - 1) / kStepBlock;
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
  CK(cudaMallocHost(&c->h_in, size_t(madicp_ctx::kInRing) * 128));
  for (int i = 0; i < madicp_ctx::kInRing; ++i) CK(cudaEventCreateWithFlags(&c->in_done[i], cudaEventDisableTiming));
  CK(cudaMallocHost(&c->h_matched, kMatchedCap));
  int threads = 1024, ctas = 1;
  if (const char* e = getenv("MADICP_GN_SHAPE"))
    if (sscanf(e, "%d,%d", &threads, &ctas) == 2) c->gn_auto = false;
  int rc = configure_gn(c, threads, ctas);
  if (rc) return rc;
  c->cap_partial = size_t(c->sm_count) * 8 * kAcc;
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
  cudaFree(c->d_pool_recs);
  cudaFree(c->d_pool_links);
  cudaFree(c->d_heap);
  cudaFree(c->d_bfs_of);
  cudaFree(c->d_heap_pos);
  cudaFree(c->d_pool_fast);
  cudaFree(c->d_quad);
  cudaFree(c->d_dbg_cta);
  cudaFree(c->d_moving);
  cudaFree(c->d_mov4);
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
  cudaFreeHost(c->h_in);
  for (int i = 0; i < madicp_ctx::kInRing; ++i)
    if (c->in_done[i]) cudaEventDestroy(c->in_done[i]);
  cudaFreeHost(c->h_matched);
  cudaStreamDestroy(c->own_stream);
  delete c;
}

int madicp_set_params(madicp_ctx_t* c, double min_ball, double rho_ker, double b_ratio) {
  if (!c || !(min_ball > 0) || rho_ker < 0) {
    set_error("madicp_set_params: bad arguments");
    return MADICP_ERR_INVALID;
  }
  const bool reweigh = (min_ball != c->P.min_ball);
  c->P.min_ball = min_ball;
  c->P.rho_ker_sqrt = sqrt(rho_ker);
  c->P.b_ratio = b_ratio;
  c->mov4_stale = true;  // the gate radius depends on min_ball and b_ratio
  if (reweigh) {         // leaf planarity weights (1 - bbox0/min_ball)^2 live in the leaf shadows
    CK(cudaSetDevice(c->device));
    for (int s = 0; s < c->max_keyframes; ++s)
      if (c->slots[s].n_nodes > 0) {
        int rc = prepare_slot(c, s);
        if (rc) return rc;
      }
  }
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
  // heap position of every node (host, O(n)): children of the node at position h sit at 2h+1, 2h+2
  std::vector<int> heap_pos(size_t(n_nodes), 0);
  int64_t max_pos = 0;
  for (int i = 0; i < n_nodes; ++i) {
    const int link = recs[i].link;
    if (link < 0) continue;
    if (link < 1 || link + 1 >= n_nodes || link <= i) {
      set_error("madicp_put_keyframe: records are not a breadth-first tree with adjacent siblings");
      return MADICP_ERR_INVALID;
    }
    // the implicit binary heap (walk modes 1-3, kept for measurements) is limited to 20 levels; deeper nodes
    // get no heap position and those modes are switched off -- the default 4-ary records have no depth limit
    const int64_t h = (heap_pos[size_t(i)] < 0) ? -1 : 2 * int64_t(heap_pos[size_t(i)]) + 1;
    if (h < 0 || h + 1 >= (int64_t(1) << 21)) {
      heap_pos[size_t(link)] = heap_pos[size_t(link) + 1] = -1;
      c->heap_ok = false;
      if (c->walk_mode >= 1 && c->walk_mode <= 3) c->walk_mode = 4;
      continue;
    }
    heap_pos[size_t(link)] = int(h);
    heap_pos[size_t(link) + 1] = int(h + 1);
    if (h + 1 > max_pos) max_pos = h + 1;
  }
  // dense 4-ary records: breadth-first over the even-depth nodes; the (up to four) grandchildren of a node
  // get four contiguous records.  depth parity from the heap position (depth = floor(log2(pos + 1))).
  std::vector<int> quad_pos(size_t(n_nodes), 0), quad_child(size_t(n_nodes), 0);
  {
    int next_rec = 1;  // record 0 = the root
    std::vector<int> order{0};  // even-depth nodes in breadth-first order; their record = quad_pos >> 2
    for (size_t h = 0; h < order.size(); ++h) {
      const int i = order[h];
      const int rec = quad_pos[size_t(i)] >> 2;
      const int l0 = recs[i].link;
      if (l0 < 0) continue;  // a leaf at an even depth: p0 holds the leaf code
      bool any = false;
      for (int s0 = 0; s0 < 2; ++s0) {
        const int ch = l0 + s0;
        quad_pos[size_t(ch)] = rec * 4 + 1 + s0;
        const int l1 = recs[ch].link;
        if (l1 < 0) continue;
        for (int s1 = 0; s1 < 2; ++s1) {
          quad_pos[size_t(l1 + s1)] = (next_rec + 2 * s0 + s1) * 4;
          order.push_back(l1 + s1);
          any = true;
        }
      }
      quad_child[size_t(i)] = next_rec;
      if (any) next_rec += 4;
    }
    if (size_t(next_rec) > 2 * (size_t(n_nodes) + 64)) {
      set_error("madicp_put_keyframe: internal error (4-ary record count)");
      return MADICP_ERR_INVALID;
    }
  }
  int rc = ensure_pool(c, size_t(n_nodes), size_t(8 * max_pos + 16));  // room for the 3-level look-ahead prefetch
  if (rc) return rc;
  Slot& s = c->slots[slot];
  CK(cudaMemcpyAsync(c->d_pool_recs + size_t(slot) * c->pool_cap, recs, size_t(n_nodes) * sizeof(madtree_rec_t),
                     cudaMemcpyHostToDevice, c->stream));
  s.n_nodes = n_nodes;
  s.n_leaves = n_leaves;
  s.heap_pos.swap(heap_pos);
  s.quad_pos.swap(quad_pos);
  s.quad_child.swap(quad_child);
  rc = prepare_slot(c, slot);
  if (rc) return rc;
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

static int prepare_moving(madicp_ctx* c) {
  if (!c->mov4_stale || c->L < 1) return MADICP_OK;
  k_prepare_moving<<<(c->L + kStepBlock - 1) / kStepBlock, kStepBlock, 0, c->stream>>>(c->d_moving, c->L, c->P, c->d_mov4);
  c->launches++;
  CK(cudaGetLastError());
  c->mov4_stale = false;
  return MADICP_OK;
}

int madicp_set_moving(madicp_ctx_t* c, const double* means, int L) {
  if (!c || !means || L < 1 || size_t(L) > kMatchedCap) {
    set_error("madicp_set_moving: bad arguments (1 <= L <= 1048576)");
    return MADICP_ERR_INVALID;
  }
  CK(cudaSetDevice(c->device));
  if (size_t(L) > c->cap_moving) {
    CK(cudaStreamSynchronize(c->stream));
    if (c->d_moving) cudaFree(c->d_moving);
    if (c->d_mov4) cudaFree(c->d_mov4);
    if (c->d_step_matched) cudaFree(c->d_step_matched);
    c->d_moving = nullptr;
    c->d_mov4 = nullptr;
    c->d_step_matched = nullptr;
    c->cap_moving = 0;
    const size_t cap = size_t(L) + size_t(L) / 4 + 1024;
    CK(cudaMalloc(&c->d_moving, cap * 3 * sizeof(double)));
    CK(cudaMalloc(&c->d_mov4, cap * sizeof(Moving4)));
    CK(cudaMalloc(&c->d_step_matched, cap));
    c->cap_moving = cap;
  }
  CK(cudaMemcpyAsync(c->d_moving, means, size_t(L) * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->L = L;
  c->mov4_stale = true;
  return prepare_moving(c);
}

static int check_ready(madicp_ctx* c, const char* who) {
  if (!c) return MADICP_ERR_INVALID;
  if (c->L < 1 || !c->d_moving) {
    set_error(std::string(who) + ": no moving leaves (call madicp_set_moving first)");
    return MADICP_ERR_STATE;
  }
  if (int64_t(madicp_num_keyframes(c)) * int64_t(c->L) >= (int64_t(1) << 31)) {
    set_error(std::string(who) + ": keyframes x moving leaves must stay below 2^31 (32-bit item index)");
    return MADICP_ERR_INVALID;
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
  rc = prepare_moving(c);
  if (rc) return rc;
  k_search<<<grid_for(c, items), kStepBlock, 0, c->stream>>>(mv, c->d_mov4, c->L, d_X, c->d_hit,
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
  k_linearize<<<grid, kStepBlock, 0, c->stream>>>(mv, c->d_mov4, c->L, c->d_X, c->P, c->d_hit, c->d_step_matched,
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
  rc = prepare_moving(c);
  if (rc) return rc;
  rc = pick_shape(c, int64_t(madicp_num_keyframes(c)) * c->L);
  if (rc) return rc;
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
  A.moving = c->d_mov4;
  A.L = c->L;
  A.iters = iters;
  A.matched = c->d_comm->matched[mb];
  A.partial = c->d_partial;
  // Item map in shared memory: 4 bytes per CTA-local item, taken only if it neither exceeds the reserve
  // nor pushes the CTA into the next shared-memory carve-out (that would shrink L1, which holds the tree).
  size_t map_bytes = 0;
  {
    const size_t per_cta = size_t(madicp_num_keyframes(c)) * (size_t(c->L) / size_t(c->gn_grid) + 8) * 4 + 128;
    auto bucket = [](size_t bytes) {
      const size_t kb[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
      for (size_t b : kb)
        if (bytes + 1024 <= b * 1024) return b;  // 1 KB of static shared memory + system reserve
      return size_t(1 << 20);
    };
    const size_t ctas = size_t(c->gn_grid / c->sm_count);
    if (per_cta <= kGnMapMaxBytes && c->L < (1 << 26) && bucket((c->gn_smem + per_cta) * ctas) == bucket(c->gn_smem * ctas))
      map_bytes = per_cta;
  }
  A.map_in_smem = map_bytes ? 1 : 0;
  A.st = c->d_state;
  A.dbg = c->d_dbg;
  A.dbg_cta = c->d_dbg ? c->d_dbg_cta : nullptr;
  c->epoch += uint32_t(iters);
  A.pose_epoch = c->pose_epoch;
  c->pose_epoch += uint32_t(iters);
  // control words + initial pose: one small pinned H2D copy from the next header of the ring
  static_assert(offsetof(GnState, X_out) <= 128, "launch header must fit a ring entry");
  const int ring = int(c->call_seq % madicp_ctx::kInRing);
  if (c->call_seq >= madicp_ctx::kInRing) CK(cudaEventSynchronize(c->in_done[ring]));
  GnState* hs = reinterpret_cast<GnState*>(c->h_in + size_t(ring) * 128);
  hs->ticket = 0;
  hs->round = 0;
  hs->n_matched = 0;
  hs->pad = 0;
  memcpy(hs->X_in, X0, 12 * sizeof(double));
  CK(cudaMemcpyAsync(c->d_state, hs, offsetof(GnState, X_out), cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(c->in_done[ring], c->stream));
  // zero the flags buffer of the NEXT call (nobody can be writing it yet; see CommBlock)
  CK(cudaMemsetAsync(c->d_comm->matched[mb ^ 1], 0, std::min(kMatchedCap, c->cap_moving), c->stream));
  void* args[] = {&A};
  CK(cudaLaunchCooperativeKernel(c->gn_kernel, dim3(c->gn_grid), dim3(c->gn_threads), args, c->gn_smem + map_bytes, c->stream));
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
  CK(cudaMemcpyAsync(c->h_state, c->d_state, offsetof(GnState, X_trace), cudaMemcpyDeviceToHost, c->stream));
  if (matched)
    CK(cudaMemcpyAsync(c->h_matched, c->d_comm->matched[(c->call_seq - 1u) & 1u], size_t(c->L), cudaMemcpyDeviceToHost,
                       c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (X) memcpy(X, c->h_state->X_out, 12 * sizeof(double));
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
  k_search_cloud<<<grid_for(c, n), kStepBlock, 0, c->stream>>>(make_view(c), slot_rank(c, slot), d_q, n, d_o,
                                                              points ? d_p : nullptr, normals ? d_n : nullptr,
                                                              dists ? d_d : nullptr);
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
    CK(cudaMalloc(&c->d_dbg_cta, size_t(MADICP_MAX_ITERS) * c->sm_count * 8 * sizeof(long long)));
  } else if (!enable && c->d_dbg) {
    cudaFree(c->d_dbg);
    cudaFree(c->d_dbg_cta);
    c->d_dbg = nullptr;
    c->d_dbg_cta = nullptr;
  }
  return rows;
}

int madicp_debug_cta_cycles(madicp_ctx_t* c, int64_t* out, int cap) {
  if (!c || !out || !c->d_dbg_cta) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  const int n = std::min(cap, c->last_iters * c->gn_grid);
  CK(cudaMemcpy(out, c->d_dbg_cta, size_t(n) * sizeof(long long), cudaMemcpyDeviceToHost));
  return c->gn_grid;
}

int madicp_set_walk_mode(madicp_ctx_t* c, int mode) {
  if (!c || mode < 0 || mode > 4) return MADICP_ERR_INVALID;
  if (mode >= 1 && mode <= 3 && !c->heap_ok) {
    set_error("madicp_set_walk_mode: a resident keyframe is deeper than the implicit-heap limit (20 levels)");
    return MADICP_ERR_STATE;
  }
  c->walk_mode = mode;
  return MADICP_OK;
}

int madicp_set_gn_grid(madicp_ctx_t* c, int threads_per_cta, int ctas_per_sm) {
  if (!c || ctas_per_sm < 1) return MADICP_ERR_INVALID;
  CK(cudaSetDevice(c->device));
  if (threads_per_cta == 0) {  // back to automatic selection
    c->gn_auto = true;
    return 0;
  }
  int rc = configure_gn(c, threads_per_cta, ctas_per_sm);
  if (rc) return rc;
  c->gn_auto = false;
  return c->gn_grid / c->sm_count;
}

}  // extern "C"
