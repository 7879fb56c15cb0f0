/* madicp_b200_debug.h -- tuning and diagnostic entry points of libmadicp_b200.so.  NOT part of the drop-in
 * surface (include/madicp_b200.h): nothing a user of the reference's API needs; used by scripts/ and tests. */
#ifndef MADICP_B200_DEBUG_H
#define MADICP_B200_DEBUG_H

#include "madicp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-round SM-clock stamps of the persistent kernel: enable != 0 switches recording on for the
 * following launches; out (nullable) receives rounds x 8 int64 of the LAST launch:
 * [0] item phase of CTA 0, [1] round start -> last CTA arrived, [2] fold of the per-CTA partials,
 * [3] peer exchange + matched count, [4] solve + publish (cycles).  Returns rows written. */
int madicp_debug_timing(madicp_ctx_t* ctx, int enable, int64_t* out, int max_rounds);
/* Item-phase cycles of every CTA for the rounds of the last launch (rounds x grid int64, debug timing
 * must be on).  Returns the grid size. */
int madicp_debug_cta_cycles(madicp_ctx_t* ctx, int64_t* out, int cap);
/* Plane 1..4 of the per-CTA stamps of the last launch (rounds x grid int64, debug timing on): %globaltimer in ns at the
 * start of the round's items, at their end, and after the CTA's tile went out (plane 0 = madicp_debug_cta_cycles).
 * Rows 6 and 7 of madicp_debug_timing are on the same clock: all tiles folded, next pose handed out.  Plane 4, first
 * 16 entries of a round: CTA 0's fold trace in SM cycles ([0] entry, [1..12] end of thread 0's sweeps, [13] done,
 * [14] number of sweeps). */
int madicp_debug_cta_stamps(madicp_ctx_t* ctx, int plane, int64_t* out, int cap);
/* Shape of the persistent kernel: threads per CTA and resident CTAs per SM; supported pairs are
 * (1024,1) default, (768,1), (512,1), (512,2), (256,2), (256,3), (256,4); env MADICP_GN_SHAPE="t,c"
 * selects one at create time.  By default the library picks among the one-CTA-per-SM shapes per
 * launch from the item count; threads_per_cta = 0 restores that.  Returns the CTAs per SM in effect. */
int madicp_set_gn_grid(madicp_ctx_t* ctx, int threads_per_cta, int ctas_per_sm);


/* Path memo of the persistent kernel (kernels.cuh, descend_t): enable = 0 walks every (leaf, keyframe) pair in every
 * round, as round 1 did.  Results are identical either way (the memo only skips walks it has proved unchanged);
 * this switch exists for A/B measurements and for the test that checks exactly that. */
int madicp_debug_set_memo(madicp_ctx_t* ctx, int enable);

/* Diagnostic for madicp_deskew's sort: n pseudo-random keys over `distinct` values, sorted by std::sort
 * and by the threaded restatement of it; returns how many positions of the two permutations differ (0). */
int64_t madicp_debug_sort_check(int64_t n, uint32_t seed, int64_t distinct, int num_threads);

#ifdef __cplusplus
}
#endif
#endif /* MADICP_B200_DEBUG_H */
