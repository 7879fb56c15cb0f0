/* madicp_b200.h -- C ABI of libmadicp_b200.so: the B200 (sm_100a) implementation of MAD-ICP's
 * per-scan registration hot path.  Plain pointers and sizes only; no C++/torch types cross this
 * boundary.  The reference has no FFI of its own -- its boundary is the C++ class API
 * (MADtree / MADicp) that Pipeline and the pybind wrappers call -- so each entry point below cites
 * the reference member it stands in for (paths relative to mad_icp/src/ in rvp-group/mad-icp
 * v0.0.10).  The C++ facade (mad_icp_b200/csrc/facade/) and the pybind modules sit on top of this
 * header and keep the reference's names; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns MADICP_OK (0) or a negative MADICP_ERR_*; nothing throws across the ABI;
 *     madicp_last_error() returns a message for the last failure on the calling thread.
 *   - host buffers are caller-owned; device memory, streams and peer mappings are library-owned.
 *   - poses are 3x4 row-major [R|t] doubles (X[r*4+c]); H is 6x6 (symmetric, both triangles filled),
 *     b is 6, ordered [t_x t_y t_z w_x w_y w_z] as in the reference (odometry/mad_icp.cpp:112-115).
 *   - a context is driven by one host thread at a time.
 *   - there is NO CPU fallback: every madicp_* compute call runs CUDA kernels and fails with
 *     MADICP_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef MADICP_B200_H
#define MADICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MADICP_OK 0
#define MADICP_ERR_INVALID (-1) /* bad argument (null pointer, size, slot out of range, empty cloud) */
#define MADICP_ERR_CUDA (-2)    /* CUDA runtime / launch failure, or no usable GPU */
#define MADICP_ERR_STATE (-3)   /* call order violated (e.g. register before set_moving) */
#define MADICP_ERR_NOMEM (-4)
#define MADICP_ERR_COMM (-5) /* peer (multi-GPU) set-up failure */

#define MADICP_MAX_ITERS 64

/* ----------------------------------------------------------------------------------------------
 * Flat MAD-tree node record: 64 bytes, 64-byte aligned, breadth-first order, the two children of
 * a node adjacent (right = link + 1).  One 64-byte record = two 256-bit loads on sm_100a.
 *   internal node : mean = centroid, dir = eigenvectors.col(2) (split direction), link = index of
 *                   the left child (>= 1)
 *   leaf          : mean = cloud point nearest the centroid, dir = eigenvectors.col(0) (surface
 *                   normal, possibly inherited), bbox0 = bbox(0), link = -1 - leaf_ordinal where
 *                   leaf_ordinal is the position in MADtree::getLeafs order (DFS, left first)
 * replaces: struct MADtree fields mean_/eigenvectors_/bbox_/left_/right_ (tools/mad_tree.h:91-98)
 * -------------------------------------------------------------------------------------------- */
typedef struct madtree_rec {
  double mean[3];
  double dir[3];
  double bbox0;
  int32_t link;
  int32_t num_points;
} madtree_rec_t;

typedef struct madtree madtree_t;       /* host-resident flat MAD-tree (built on the CPU) */
typedef struct madicp_ctx madicp_ctx_t; /* one GPU: keyframe slots, moving leaves, stream, GN state */

const char* madicp_last_error(void);
/* Version of this ABI (bumped on any signature change). */
int madicp_abi_version(void);

/* ================================ host side: MAD-tree ======================================== */

/* MADtree::MADtree(vec, begin, end, b_max, b_min, 0, max_parallel_level, nullptr, nullptr)
 * (tools/mad_tree.cpp:33-130).  points_xyz: n x 3 doubles (std::vector<Eigen::Vector3d> layout).
 * Like the reference the build reorders a private copy of the cloud.  n == 0 is rejected
 * (the reference dereferences *begin, UB).  num_threads <= 1 builds serially. */
int madtree_build(const double* points_xyz, int64_t n, double b_max, double b_min, int num_threads, madtree_t** out);
void madtree_free(madtree_t* t);
int madtree_num_nodes(const madtree_t* t);
int madtree_num_leaves(const madtree_t* t);
/* MADtree::applyTransform(r, t) (tools/mad_tree.cpp:165-172) on every node; X = [R|t] row-major. */
int madtree_apply_transform(madtree_t* t, const double X[12]);
/* MADtree::getLeafs order (tools/mad_tree.cpp:154-163).  Any output pointer may be NULL.
 * means/normals: L x 3, bbox0: L, num_points: L. */
int madtree_leaves(const madtree_t* t, double* means, double* normals, double* bbox0, int32_t* num_points);
/* Breadth-first 64-byte records (see madtree_rec_t); valid until the next apply_transform/free. */
const madtree_rec_t* madtree_records(const madtree_t* t);
/* Depth of the tree + 1, and the level table: out[d] = breadth-first index of the first node of depth d,
 * out[levels] = node count (cap >= levels + 1).  Returns the number of levels. */
int madtree_num_levels(const madtree_t* t);
int madtree_level_offsets(const madtree_t* t, int32_t* out, int cap);
/* getLeafs order -> breadth-first record index, L entries.  Returns L. */
int madtree_leaf_records(const madtree_t* t, int32_t* out);
/* Full per-node dump in DFS pre-order for audits/tests: mean n x 3, eigenvectors n x 9 (column-major),
 * bbox n x 3, num_points n, left/right pre-order index (-1 on leaves), leaf_ordinal (-1 on internal). */
int madtree_export(const madtree_t* t, double* mean, double* eigenvectors, double* bbox, int32_t* num_points,
                   int32_t* left, int32_t* right, int32_t* leaf_ordinal);

/* ================================ device side: registration ================================== */

/* Creates a context on CUDA device `device` with `max_keyframes` model slots.
 * replaces: MADicp::MADicp (odometry/mad_icp.cpp:31-39) + the keyframe deque of Pipeline
 * (odometry/pipeline.h:85). */
int madicp_create(madicp_ctx_t** out, int device, int max_keyframes);
void madicp_destroy(madicp_ctx_t* ctx);
/* MADicp ctor parameters: min_ball (= b_max), rho_ker (the kernel stores sqrt(rho_ker) as the
 * reference does, mad_icp.cpp:32), b_ratio. */
int madicp_set_params(madicp_ctx_t* ctx, double min_ball, double rho_ker, double b_ratio);
/* Launch all work on this CUDA stream (a cudaStream_t / torch stream handle) instead of the
 * context's own non-blocking stream.  NULL restores the internal stream. */
int madicp_set_stream(madicp_ctx_t* ctx, void* cuda_stream);
void* madicp_get_stream(const madicp_ctx_t* ctx);

/* Upload a (map-frame, i.e. already applyTransform-ed) keyframe tree into model slot `slot`
 * (replaces pushing a Frame onto keyframes_, odometry/pipeline.cpp:250-257). Overwrites the slot. */
int madicp_put_keyframe(madicp_ctx_t* ctx, int slot, const madtree_t* tree);
/* The same for a SENSOR-frame tree plus its pose: MADtree::applyTransform(R, t) (tools/mad_tree.cpp:165-172,
 * called at odometry/pipeline.cpp:224) runs on the device, fused into the upload, with the reference's operand
 * order and no FMA -- bit-identical to madtree_apply_transform + madicp_put_keyframe.  X = NULL: no transform.
 * Asynchronous on the context's stream; the host tree may be freed as soon as the call returns. */
int madicp_put_keyframe_transformed(madicp_ctx_t* ctx, int slot, const madtree_t* tree, const double X[12]);
/* Caller-supplied records (validated: one breadth-first tree, siblings adjacent, ordinals a permutation). */
int madicp_put_keyframe_records(madicp_ctx_t* ctx, int slot, const madtree_rec_t* recs, int n_nodes, int n_leaves);
int madicp_drop_keyframe(madicp_ctx_t* ctx, int slot); /* keyframes_.pop_front(), pipeline.cpp:253-256 */
int madicp_num_keyframes(const madicp_ctx_t* ctx);     /* active slots on THIS device */
/* Active slots in ascending slot order; returns count. */
int madicp_active_slots(const madicp_ctx_t* ctx, int* slots_out, int cap);
int madicp_keyframe_leaves(const madicp_ctx_t* ctx, int slot); /* leaves in slot, <0 if empty */

/* MADicp::setMoving (mad_icp.cpp:51-53): sensor-frame means of the current scan's leaves, in
 * getLeafs order, L x 3 doubles on the HOST; copied to the device. */
int madicp_set_moving(madicp_ctx_t* ctx, const double* means_xyz, int L);
/* Wait for everything enqueued on the context's stream. */
int madicp_synchronize(madicp_ctx_t* ctx);

/* ---------------- device-resident MAD-trees ----------------------------------
 * A madtree_gpu_t is a sensor-frame MAD-tree living in device memory of ONE context: breadth-first records,
 * level table, getLeafs table.  The scan's tree never has to exist on the host: build (or upload) -> the
 * scan's leaves become the moving leaves -> on promotion the tree is transformed and laid out in a keyframe
 * slot, all on the device, all asynchronous on the context's stream. */
typedef struct madtree_gpu madtree_gpu_t;
/* MADtree::MADtree(...) (tools/mad_tree.cpp:33-130) ON THE DEVICE: points_xyz n x 3 doubles on the host (one
 * H2D copy), same tree bit for bit as madtree_build / the reference (split order, NaN nodes, normal inheritance).
 * The three libm calls per node of Eigen's computeDirect (atan2, cos, sin: glibc is not correctly rounded, so no
 * device implementation can reproduce its bits) are served by the host between two kernels of a level. */
int madtree_gpu_build(madicp_ctx_t* ctx, const double* points_xyz, int64_t n, double b_max, double b_min,
                      madtree_gpu_t** out);
/* The same from a cloud already on the device (madicp_ingest). */
int madtree_gpu_build_resident(madicp_ctx_t* ctx, double b_max, double b_min, madtree_gpu_t** out);
/* A batch of scans at once (look-ahead: the tree of a scan depends on the pose estimates only when it is deskewed, so
 * the trees of the next scans can be built before their turn): `count` clouds (all float32 or all float64, host
 * memory) -> `count` trees, built as ONE forest.  The build's latency is that of its dependent-add chains, which a
 * batch runs side by side, so sixteen trees cost little more than one.  out[count]. */
int madtree_gpu_build_batch(madicp_ctx_t* ctx, const void* const* clouds, const int64_t* n_points, int is_f32, int count,
                            double b_max, double b_min, madtree_gpu_t** out);
/* Early upload for the NEXT madtree_gpu_build_batch: starts copying `cloud` (host memory, n_points x 3 float32 or
 * float64; it must stay valid and unchanged until that batch call returns) to the device now, on a copy stream of its
 * own, so that the copy runs under the registrations of earlier scans.  Staged clouds lie back to back in the build
 * lane's working buffer: the batch call uses the longest prefix of its clouds[] that was staged in this order (same
 * pointers and sizes) and copies the rest itself; any other build / ingest call on the context discards what was
 * staged.  reserve_points: total points the batch will hold (sizes the lane on first use; 0 = this cloud only).
 * Returns MADICP_OK whether or not the cloud could be staged. */
int madicp_stage_cloud(madicp_ctx_t* ctx, const void* cloud, int64_t n_points, int is_f32, int64_t reserve_points);
/* Gives up whatever madicp_stage_cloud staged and returns once nothing reads the staged host buffers any more (their
 * uploads and the background sums of their roots): call it before releasing a staged cloud that was never built. */
int madicp_stage_discard(madicp_ctx_t* ctx);
/* Upload of a host-built tree (records + tables), asynchronous. */
int madtree_gpu_upload(madicp_ctx_t* ctx, const madtree_t* tree, madtree_gpu_t** out);
void madtree_gpu_free(madtree_gpu_t* t);
int madtree_gpu_num_nodes(const madtree_gpu_t* t);
int madtree_gpu_num_leaves(const madtree_gpu_t* t);
int madtree_gpu_num_levels(const madtree_gpu_t* t);
/* Device -> host (synchronises): the breadth-first records and/or the getLeafs table.  Either may be NULL. */
int madtree_gpu_download(const madtree_gpu_t* t, madtree_rec_t* recs_out, int32_t* leaf_records_out);
/* Audit dump of a DEVICE-BUILT tree in breadth-first order: mean n x 3, eigenvectors n x 9 (column-major), bbox
 * n x 3, num_points n (any may be NULL).  Valid for the most recently built tree of the context.  Synchronises. */
int madtree_gpu_export(const madtree_gpu_t* t, double* mean, double* eigenvectors, double* bbox, int32_t* num_points);
/* MADicp::setMoving with the leaves of a device tree (no host copy of the means). */
int madicp_set_moving_tree(madicp_ctx_t* ctx, const madtree_gpu_t* t);
/* The current moving-leaf means back on the host (L x 3), e.g. for Pipeline::currentLeaves.  Returns L. */
int madicp_get_moving(madicp_ctx_t* ctx, double* means_out, int cap_leaves);
/* Keyframe promotion from a device tree: D2D copy + applyTransform(X) (NULL: none) + layout, no host sync. */
int madicp_put_keyframe_tree(madicp_ctx_t* ctx, int slot, const madtree_gpu_t* t, const double X[12]);

/* Ingest of a raw scan on the device (odometry/pipeline.cpp:79-123 and the float32 -> float64
 * conversion of the readers / pybind/eigen_stl_bindings.h:44-60).  xyz: n x 3 float32 (is_f32 != 0) or float64 on
 * the host, copied as is.  deskew == 0: conversion only.  deskew != 0: Pipeline::deskew -- the azimuth sort (the
 * reference's std::sort permutation, ties included) and the <= 1024 chunk poses come from the host
 * (threaded; atan2/sin/cos are glibc's), the gather by that permutation, the conversion and the per-chunk rigid
 * transform (reference operand order, no FMA) run on the device.  The result is the device-resident cloud that
 * madtree_gpu_build_resident consumes; points_out (nullable, n x 3 doubles) receives a copy. */
int madicp_ingest(madicp_ctx_t* ctx, const void* xyz, int64_t n, int is_f32, int deskew, const double T_prev[12],
                  const double T_now[12], double sensor_hz, int num_threads, double* points_out);

/* K1 only -- MADtree::bestMatchingLeafFast (tools/mad_tree.cpp:144-152) of X*mean for every moving
 * leaf against every active keyframe.  out_ordinals: K_active x L int32 on the host (row k = k-th
 * active slot in ascending slot order); values are getLeafs ordinals of the matched leaf. */
int madicp_search(madicp_ctx_t* ctx, const double X[12], int32_t* out_ordinals);

/* One linearisation at pose X: resetAdders() + update(tree) for every active keyframe
 * (mad_icp.cpp:41-49,74-103) WITHOUT updateState.  H (36), b (6) summed over this device's
 * keyframes; matched (nullable, L bytes) receives 1 where any keyframe passed the gate, else 0. */
int madicp_linearize(madicp_ctx_t* ctx, const double X[12], double H[36], double b[6], uint8_t* matched);

/* MADicp::updateState (mad_icp.cpp:105-117) on the device for caller-supplied H, b. */
int madicp_solve_update(madicp_ctx_t* ctx, const double H[36], const double b[6], double X_inout[12]);

/* The whole ICP loop of Pipeline::compute / MADicpWrapper::compute (odometry/pipeline.cpp:166-193,
 * pybind/tools/mad_icp_wrapper.h:72-81) in ONE persistent kernel: `iters` rounds of
 * {clear matched on the last round; resetAdders; update over all keyframes; updateState}.
 * With peers connected (madicp_comm_connect) every round all-reduces H/b across the GPUs inside the
 * kernel.  Outputs (nullable): final pose, H/b of the last round (Pipeline reads H_adder_,
 * pipeline.cpp:223), matched flags of the last round (L bytes), their count. */
int madicp_register(madicp_ctx_t* ctx, int iters, double X_inout[12], double H_last[36], double b_last[6],
                    uint8_t* matched_last, int* n_matched);
/* Same, split for pipelining / device-resident timing: enqueue on the stream with the resident
 * moving leaves, then fetch (synchronises). */
int madicp_register_async(madicp_ctx_t* ctx, int iters, const double X0[12]);
int madicp_register_fetch(madicp_ctx_t* ctx, double X[12], double H_last[36], double b_last[6], uint8_t* matched_last,
                          int* n_matched);
/* The same plus Frame::weight_ = det(H^-1) of the last round's H (odometry/pipeline.cpp:223), computed by the
 * solve thread on the device. */
int madicp_register_fetch_weight(madicp_ctx_t* ctx, double X[12], double H_last[36], double b_last[6],
                                 uint8_t* matched_last, int* n_matched, double* weight);
/* A loop the `realtime` budget cut short (odometry/pipeline.cpp:167-169): `iters` < MAX_ICP_ITS rounds ran, so the
 * clear of the matched flags that belongs to round MAX_ICP_ITS-1 (pipeline.cpp:172-176) never happened and the
 * flags are the union over all rounds. */
int madicp_register_partial_async(madicp_ctx_t* ctx, int iters, const double X0[12]);
/* Per-round poses of the last madicp_register* call: (iters+1) x 12 doubles (X before round i; the
 * last row is the final pose).  Debug/parity aid. */
int madicp_register_trace(madicp_ctx_t* ctx, double* X_trace, int max_rounds);

/* Per round of the last madicp_register* call: how many (moving leaf, keyframe) pairs the kernel actually walked; the
 * others provably kept the leaf of their last walk (path memo, kernels.cuh).  Returns the number of rounds written. */
int madicp_register_walked(madicp_ctx_t* ctx, int32_t* walked, int max_rounds);

/* Pipeline::deskew (odometry/pipeline.cpp:79-123), host side: sorts the n points by azimuth, cuts the
 * sweep into 1024 chunks, applies to every chunk the pose interpolated from the relative motion of the
 * last two estimates (T_prev, T_now: 3x4 row-major), and rewrites points_xyz in sorted order -- the
 * reference's permutation exactly, ties of its unstable sort included.  Runs on num_threads host
 * threads; no device work. */
int madicp_deskew(double* points_xyz, int64_t n, const double T_prev[12], const double T_now[12], double sensor_hz,
                  int num_threads);

/* MADtreeWrapper::searchCloud / searchCloudDist (pybind/tools/mad_tree_wrapper.h:48-67): nearest-leaf
 * search of n host query points in slot `slot`.  Any output may be NULL: ordinals n, points n x 3
 * (leaf mean), normals n x 3, dists n. */
int madicp_search_cloud(madicp_ctx_t* ctx, int slot, const double* queries_xyz, int64_t n, int32_t* ordinals,
                        double* points, double* normals, double* dists);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t madicp_kernel_launches(const madicp_ctx_t* ctx);
/* Sum over active keyframes of nodes / leaves (sizing for the algorithmic-bytes formula). */
int64_t madicp_model_nodes(const madicp_ctx_t* ctx);

/* ================================ multi-GPU (one process per GPU) ============================= */
/* Keyframes shard across ranks (slot s lives on rank s % world, done by the caller); the only
 * exchange is the 27-value H/b sum per round and the matched flags at the end.  Each rank exports a
 * 64-byte CUDA IPC handle of its mailbox; the host side all-gathers them (torch.distributed) and
 * hands every rank the full table.  After connect, madicp_register* runs the all-reduce inside the
 * persistent kernel with peer stores over NVLink; every rank sums the partials in rank order, so all
 * ranks hold bit-identical H, b and X. */
#define MADICP_IPC_HANDLE_BYTES 64
int madicp_comm_export(madicp_ctx_t* ctx, void* handle_out /* 64 bytes */);
int madicp_comm_connect(madicp_ctx_t* ctx, int rank, int world, const void* all_handles /* world x 64 bytes */);
int madicp_comm_world(const madicp_ctx_t* ctx);

/* Measures, on the resident model and moving leaves, the per-pass cost of each persistent-kernel shape the
 * automatic choice considers (a few one-round registrations each) and uses it from then on.  Optional: without it
 * a prior measured on B200 is used.  Returns the number of shapes measured. */
int madicp_calibrate(madicp_ctx_t* ctx, const double X0[12]);

/* Tuning and debug entry points (clock stamps, kernel shape override) are declared in madicp_b200_debug.h. */

#ifdef __cplusplus
}
#endif
#endif /* MADICP_B200_H */
