/* oracle/restate -- CPU restatement of the Ansel develop hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so; the product (libb200iop.so and ansel_b200/) never links or calls it.
 *
 * Numerical contract of every function in this directory: plain C11 evaluated with
 * FLT_EVAL_METHOD == 0, no floating-point contraction (-ffp-contract=off), no fast-math
 * re-association -- i.e. the source-level semantics of the reference ("ref-strict" in
 * SURVEY.md section 8c).  Each function is pinned against the reference's own source compiled
 * with those flags (oracle/_ref/libref_strict.so) by tests/test_oracle_pin.py.
 */
#ifndef B200_ORACLE_COMMON_H
#define B200_ORACLE_COMMON_H
#include <stddef.h>
#include <stdint.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CFA colour of a site: develop/imageop_math.h:190-193 */
static inline int orc_fc(size_t row, size_t col, uint32_t filters)
{
  const unsigned sh = (unsigned)((((row << 1) & 14) + (col & 1)) << 1);
  return (int)((filters >> sh) & 3u);
}

/* flush-to-zero / denormals-are-zero for the calling thread: system/fp_mode.h:45-62 */
void orc_fp_fast_mode(void);
void orc_fp_fast_mode_all(void);

/* dt_dev_get_roi_filters (develop/imageop.c:139-142): the CFA word seen from a ROI origin (pipe_ends_oracle.c) */
uint32_t orc_roi_filters(uint32_t filters, int x, int y);

#ifdef __cplusplus
}
#endif
#endif
