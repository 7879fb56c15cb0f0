// flat_tree.cpp -- host-side MAD-tree builder producing the breadth-first 64-byte record layout the
// sm_100a kernels walk (include/madicp_b200.h: madtree_*).
//
// What it computes is the reference's MADtree (tools/mad_tree.cpp:47-130 build, :154-163 leaf order,
// :165-172 applyTransform; helpers tools/utils.h:38-97); how it is organised is not: instead of one
// 152-byte heap object per node linked by pointers, nodes live in index-linked arrays created by an
// explicit-stack depth-first expansion (so node id == DFS pre-order position and leaves come out in
// getLeafs order for free), followed by one breadth-first renumbering pass that makes siblings
// adjacent and emits the device records.  Subtrees below the top levels are independent index
// ranges of the point array, so they are expanded by separate threads and spliced back in pre-order.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/madicp_b200.h"
#include "arith.h"
#include "eig3.h"

namespace madicp {
void set_error(const std::string& msg);
}

namespace {
using madicp::dot3;
using madicp::norm3;

struct Node {
  double mean[3];
  double ev[9];  // column-major eigenvectors: col0 normal, col2 split direction
  double bbox[3];
  int32_t npts;
  int32_t left, right;   // node ids (pre-order), -1 when absent; <= -2 while building: frontier reference
  int32_t leaf_ordinal;  // getLeafs position, -1 for internal nodes
};

// Pending range [begin,end) of the point array plus what the reference reaches through pointers
// (plane_predecessor, the parent chain) carried by value, so subtrees can be expanded independently:
//   pp_col0  : eigenvectors.col(0) of the plane predecessor, if one was set above (mad_tree.cpp:65-67,90-93)
//   anc_col0 : col(0) of the nearest ancestor with >= 3 points, or of the root (the walk of mad_tree.cpp:68-73)
struct Job {
  int64_t begin, end;
  int32_t parent;  // node id of the parent within the same arena, -1 for the arena's root
  bool is_right;
  bool has_pp, is_root;
  double pp_col0[3], anc_col0[3];
  int depth;
};

struct Builder {
  double* pts;  // n x 3, reordered in place
  double b_max, b_min;

  // statistics of one range -> fills mean/ev/bbox/npts of `nd`
  void stats(Node& nd, int64_t begin, int64_t end) const {
    double sx = 0, sy = 0, sz = 0;
    double cxx = 0, cyx = 0, czx = 0, cyy = 0, czy = 0, czz = 0;
    int k = 0;
    for (int64_t i = begin; i != end; ++i) {
      const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
      sx += x; sy += y; sz += z;
      cxx += x * x; cyx += y * x; czx += z * x;
      cyy += y * y; czy += z * y; czz += z * z;
      ++k;
    }
    const double inv = 1. / k;
    sx *= inv; sy *= inv; sz *= inv;
    cxx *= inv; cyx *= inv; czx *= inv; cyy *= inv; czy *= inv; czz *= inv;
    cxx -= sx * sx; cyx -= sy * sx; czx -= sz * sx;
    cyy -= sy * sy; czy -= sz * sy; czz -= sz * sz;
    const double f = double(k) / double(k - 1);
    madicp::Sym3 c{cxx * f, cyx * f, czx * f, cyy * f, czy * f, czz * f};
    nd.mean[0] = sx; nd.mean[1] = sy; nd.mean[2] = sz;
    madicp::eig3_symmetric(c, nd.ev);
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int64_t i = begin; i != end; ++i) {
      const double dx = pts[3 * i] - sx, dy = pts[3 * i + 1] - sy, dz = pts[3 * i + 2] - sz;
      for (int a = 0; a < 3; ++a) {
        const double v = dot3(nd.ev[3 * a], nd.ev[3 * a + 1], nd.ev[3 * a + 2], dx, dy, dz);
        lo[a] = (v < lo[a]) ? v : lo[a];  // NaN never replaces (one-point ranges give NaN axes)
        hi[a] = (hi[a] < v) ? v : hi[a];
      }
    }
    for (int a = 0; a < 3; ++a) nd.bbox[a] = hi[a] - lo[a];
    nd.npts = k;
  }

  // Hoare-style unstable partition of the reference (tools/utils.h:38-52): order of the two halves
  // matters because child sums are accumulated in array order.
  int64_t partition(int64_t begin, int64_t end, const Node& nd) const {
    int64_t lo = begin, hi = end;
    const double nx = nd.ev[6], ny = nd.ev[7], nz = nd.ev[8];
    while (lo != hi) {
      double* p = pts + 3 * lo;
      if (madicp::plane_side(p[0], p[1], p[2], nd.mean[0], nd.mean[1], nd.mean[2], nx, ny, nz) < 0.0) {
        ++lo;
      } else {
        double* q = pts + 3 * (hi - 1);
        std::swap(p[0], q[0]);
        std::swap(p[1], q[1]);
        std::swap(p[2], q[2]);
        --hi;
      }
    }
    return hi;
  }

  // Leaf finalisation (tools/mad_tree.cpp:64-88).
  void make_leaf(Node& nd, const Job& j) const {
    if (j.has_pp) {
      nd.ev[0] = j.pp_col0[0]; nd.ev[1] = j.pp_col0[1]; nd.ev[2] = j.pp_col0[2];
    } else if (nd.npts < 3 && !j.is_root) {
      nd.ev[0] = j.anc_col0[0]; nd.ev[1] = j.anc_col0[1]; nd.ev[2] = j.anc_col0[2];
    }
    // nearest cloud point to the centroid; first minimum wins.  The reference writes each new
    // minimum through a reference to *begin, i.e. into the first slot of the range.
    double best = std::numeric_limits<double>::max();
    double* first = pts + 3 * j.begin;
    for (int64_t i = j.begin; i != j.end; ++i) {
      const double vx = pts[3 * i], vy = pts[3 * i + 1], vz = pts[3 * i + 2];
      const double d = norm3(vx - nd.mean[0], vy - nd.mean[1], vz - nd.mean[2]);
      if (d < best) {
        first[0] = vx; first[1] = vy; first[2] = vz;
        best = d;
      }
    }
    nd.mean[0] = first[0]; nd.mean[1] = first[1]; nd.mean[2] = first[2];
  }

  // One node: statistics, leaf test, leaf finalisation or split.  Returns true for an internal node and
  // then fills the context its children inherit (`child`, ranges not set) and the split position.
  bool process(Node& nd, const Job& j, Job& child, int64_t& mid) const {
    nd.left = nd.right = -1;
    nd.leaf_ordinal = -1;
    stats(nd, j.begin, j.end);
    if (nd.bbox[2] < b_max) {
      make_leaf(nd, j);
      return false;
    }
    child = j;
    child.is_root = false;
    child.depth = j.depth + 1;
    if (!j.has_pp && nd.bbox[0] < b_min) {  // this node becomes the plane predecessor of its subtree
      child.has_pp = true;
      child.pp_col0[0] = nd.ev[0]; child.pp_col0[1] = nd.ev[1]; child.pp_col0[2] = nd.ev[2];
    }
    if (nd.npts >= 3 || j.is_root) {  // where the "fewer than 3 points" walk of a descendant leaf stops
      child.anc_col0[0] = nd.ev[0]; child.anc_col0[1] = nd.ev[1]; child.anc_col0[2] = nd.ev[2];
    }
    mid = partition(j.begin, j.end, nd);
    return true;
  }

  // Expand `root` depth first, appending nodes to `nodes` in pre-order (ids local to `nodes`).
  void expand(std::vector<Node>& nodes, const Job& root) const {
    std::vector<Job> stack;
    stack.push_back(root);
    while (!stack.empty()) {
      const Job j = stack.back();
      stack.pop_back();
      const int32_t id = int32_t(nodes.size());
      nodes.emplace_back();
      if (j.parent >= 0) (j.is_right ? nodes[size_t(j.parent)].right : nodes[size_t(j.parent)].left) = id;
      Job c;
      int64_t mid = 0;
      if (!process(nodes.back(), j, c, mid)) continue;
      c.parent = id;
      // right first so the left child is popped (and numbered) next: pre-order, left before right
      Job r = c, l = c;
      r.begin = mid; r.end = j.end; r.is_right = true;
      l.begin = j.begin; l.end = mid; l.is_right = false;
      stack.push_back(r);
      stack.push_back(l);
    }
  }
};

// Runs fn(i) for i in [0,n) on up to `threads` threads (the calling thread included).
template <class F>
void parallel_for(size_t n, int threads, F&& fn) {
  if (n == 0) return;
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    for (size_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
  };
  std::vector<std::thread> pool;
  const size_t extra = std::min(size_t(threads > 1 ? threads - 1 : 0), n - 1);
  for (size_t i = 0; i < extra; ++i) pool.emplace_back(worker);
  worker();
  for (std::thread& th : pool) th.join();
}

// Copies arena `src` behind `dst` (pre-order is preserved inside an arena); returns the offset.
int32_t splice(std::vector<Node>& dst, const std::vector<Node>& src) {
  const int32_t off = int32_t(dst.size());
  for (const Node& n : src) {
    dst.push_back(n);
    Node& m = dst.back();
    if (m.left >= 0) m.left += off;
    if (m.right >= 0) m.right += off;
  }
  return off;
}

// Emits the top skeleton in pre-order, replacing frontier references by the spliced subtrees.
int32_t emit(std::vector<Node>& out, const std::vector<Node>& top, int32_t id, const std::vector<std::vector<Node>>& sub) {
  const int32_t me = int32_t(out.size());
  out.push_back(top[size_t(id)]);
  const int32_t l = top[size_t(id)].left, r = top[size_t(id)].right;
  if (l == -1) return me;  // leaf
  const int32_t nl = (l <= -2) ? splice(out, sub[size_t(-2 - l)]) : emit(out, top, l, sub);
  const int32_t nr = (r <= -2) ? splice(out, sub[size_t(-2 - r)]) : emit(out, top, r, sub);
  out[size_t(me)].left = nl;
  out[size_t(me)].right = nr;
  return me;
}

}  // namespace

// Node ids are positions in the tree's final pre-order node array.
struct madtree {
  std::vector<double> pts;
  std::vector<Node> nodes;          // DFS pre-order
  std::vector<int32_t> leaf_nodes;  // getLeafs order -> node id
  std::vector<int32_t> bfs_index;   // node id -> breadth-first position
  std::vector<madtree_rec_t> recs;  // breadth-first records
  double b_max = 0, b_min = 0;

  void refresh_records() {
    recs.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
      const Node& n = nodes[i];
      madtree_rec_t& r = recs[bfs_index[i]];
      for (int a = 0; a < 3; ++a) r.mean[a] = n.mean[a];
      if (n.left < 0) {
        for (int a = 0; a < 3; ++a) r.dir[a] = n.ev[a];
        r.bbox0 = n.bbox[0];
        r.link = -1 - n.leaf_ordinal;
      } else {
        for (int a = 0; a < 3; ++a) r.dir[a] = n.ev[6 + a];
        r.bbox0 = n.bbox[0];
        r.link = bfs_index[n.left];
      }
      r.num_points = n.npts;
    }
  }
};

extern "C" {

int madtree_build(const double* points_xyz, int64_t n, double b_max, double b_min, int num_threads, madtree_t** out) {
  if (!points_xyz || !out || n <= 0) {
    madicp::set_error("madtree_build: null pointer or empty cloud (the reference dereferences *begin on an empty range)");
    return MADICP_ERR_INVALID;
  }
  const bool timing = getenv("MADTREE_TIMING") != nullptr;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = now();
  madtree* t = new (std::nothrow) madtree;
  if (!t) return MADICP_ERR_NOMEM;
  t->pts.assign(points_xyz, points_xyz + 3 * n);
  t->b_max = b_max;
  t->b_min = b_min;
  Builder B{t->pts.data(), b_max, b_min};
  Job root{};
  root.begin = 0;
  root.end = n;
  root.parent = -1;
  root.is_root = true;
  int threads = num_threads;
  if (threads > 64) threads = 64;
  if (threads <= 1 || n < 20000) {
    t->nodes.reserve(size_t(n / 2 + 16));
    B.expand(t->nodes, root);
  } else {
    // The sums of a node are accumulated in array order (that order defines the result), so a single
    // node cannot be split across threads -- but disjoint ranges are independent (the reference uses
    // std::async on the top log2(num_threads) levels, mad_tree.cpp:99-129).  The top levels are done
    // level by level (all nodes of a level in parallel) down to ~4 ranges per worker, the subtrees
    // below on the same pool, then one splice in pre-order.
    int depth = 0;
    while ((1 << depth) < 4 * threads) ++depth;
    std::vector<Node> top(1);
    std::vector<Job> level{root};    // jobs of the current level; job.parent = id of ITS node in `top`
    level[0].parent = 0;
    std::vector<Job> frontier;
    std::vector<int32_t> frontier_parent;  // (parent id << 1) | is_right
    for (int d = 0; d < depth && !level.empty(); ++d) {
      std::vector<Job> ctx(level.size());
      std::vector<int64_t> mids(level.size(), 0);
      std::vector<char> internal(level.size(), 0);
      parallel_for(level.size(), threads, [&](size_t i) {
        internal[i] = B.process(top[size_t(level[i].parent)], level[i], ctx[i], mids[i]) ? 1 : 0;
      });
      std::vector<Job> next;
      for (size_t i = 0; i < level.size(); ++i) {
        if (!internal[i]) continue;
        const int32_t me = level[i].parent;
        for (int side = 0; side < 2; ++side) {
          Job c = ctx[i];
          c.begin = side ? mids[i] : level[i].begin;
          c.end = side ? level[i].end : mids[i];
          c.is_right = side != 0;
          if (d + 1 == depth) {  // stop here: this range becomes an independent subtree job
            (side ? top[size_t(me)].right : top[size_t(me)].left) = -2 - int32_t(frontier.size());
            c.parent = -1;
            frontier.push_back(c);
          } else {
            const int32_t id = int32_t(top.size());
            top.emplace_back();
            (side ? top[size_t(me)].right : top[size_t(me)].left) = id;
            c.parent = id;
            next.push_back(c);
          }
        }
      }
      level.swap(next);
    }
    const auto ta = now();
    std::vector<std::vector<Node>> sub(frontier.size());
    parallel_for(frontier.size(), threads, [&](size_t k) {
      sub[k].reserve(size_t((frontier[k].end - frontier[k].begin) / 2 + 16));
      B.expand(sub[k], frontier[k]);
    });
    const auto tb = now();
    if (timing) std::fprintf(stderr, "  top %.2f ms (%zu ranges), pool %.2f ms\n", ms(t0, ta), frontier.size(), ms(ta, tb));
    size_t total = top.size();
    for (const auto& v : sub) total += v.size();
    t->nodes.reserve(total);
    emit(t->nodes, top, 0, sub);
  }
  const auto t1 = now();
  // leaves in pre-order == getLeafs order (left subtree fully before right subtree)
  for (size_t i = 0; i < t->nodes.size(); ++i)
    if (t->nodes[i].left < 0) {
      t->nodes[i].leaf_ordinal = int32_t(t->leaf_nodes.size());
      t->leaf_nodes.push_back(int32_t(i));
    }
  // breadth-first numbering with adjacent siblings
  t->bfs_index.assign(t->nodes.size(), -1);
  std::vector<int32_t> order;
  order.reserve(t->nodes.size());
  order.push_back(0);
  t->bfs_index[0] = 0;
  for (size_t h = 0; h < order.size(); ++h) {
    const Node& nd = t->nodes[order[h]];
    if (nd.left >= 0) {
      t->bfs_index[nd.left] = int32_t(order.size());
      order.push_back(nd.left);
      t->bfs_index[nd.right] = int32_t(order.size());
      order.push_back(nd.right);
    }
  }
  t->refresh_records();
  if (timing)
    std::fprintf(stderr, "madtree_build: n=%lld threads=%d expand %.2f ms, order+records %.2f ms\n", (long long) n, threads,
                 ms(t0, t1), ms(t1, now()));
  *out = t;
  return MADICP_OK;
}

void madtree_free(madtree_t* t) { delete t; }
int madtree_num_nodes(const madtree_t* t) { return t ? int(t->nodes.size()) : MADICP_ERR_INVALID; }
int madtree_num_leaves(const madtree_t* t) { return t ? int(t->leaf_nodes.size()) : MADICP_ERR_INVALID; }

int madtree_apply_transform(madtree_t* t, const double X[12]) {
  if (!t || !X) return MADICP_ERR_INVALID;
  for (Node& n : t->nodes) {
    double o[3];
    madicp::iso_apply(X, n.mean[0], n.mean[1], n.mean[2], o[0], o[1], o[2]);
    n.mean[0] = o[0]; n.mean[1] = o[1]; n.mean[2] = o[2];
    double e[9];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r)
        e[c * 3 + r] = dot3(X[r * 4], X[r * 4 + 1], X[r * 4 + 2], n.ev[c * 3], n.ev[c * 3 + 1], n.ev[c * 3 + 2]);
    std::memcpy(n.ev, e, sizeof(e));
  }
  t->refresh_records();
  return MADICP_OK;
}

int madtree_leaves(const madtree_t* t, double* means, double* normals, double* bbox0, int32_t* num_points) {
  if (!t) return MADICP_ERR_INVALID;
  for (size_t i = 0; i < t->leaf_nodes.size(); ++i) {
    const Node& n = t->nodes[t->leaf_nodes[i]];
    for (int a = 0; a < 3; ++a) {
      if (means) means[3 * i + a] = n.mean[a];
      if (normals) normals[3 * i + a] = n.ev[a];
    }
    if (bbox0) bbox0[i] = n.bbox[0];
    if (num_points) num_points[i] = n.npts;
  }
  return MADICP_OK;
}

const madtree_rec_t* madtree_records(const madtree_t* t) { return t ? t->recs.data() : nullptr; }

int madtree_export(const madtree_t* t, double* mean, double* eigenvectors, double* bbox, int32_t* num_points,
                   int32_t* left, int32_t* right, int32_t* leaf_ordinal) {
  if (!t) return MADICP_ERR_INVALID;
  for (size_t i = 0; i < t->nodes.size(); ++i) {
    const Node& n = t->nodes[i];
    if (mean) std::memcpy(mean + 3 * i, n.mean, 24);
    if (eigenvectors) std::memcpy(eigenvectors + 9 * i, n.ev, 72);
    if (bbox) std::memcpy(bbox + 3 * i, n.bbox, 24);
    if (num_points) num_points[i] = n.npts;
    if (left) left[i] = n.left;
    if (right) right[i] = n.right;
    if (leaf_ordinal) leaf_ordinal[i] = n.leaf_ordinal;
  }
  return MADICP_OK;
}

}  // extern "C"
