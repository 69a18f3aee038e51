/* oracle/_ref wrapper: the LMMSE demosaicer (Zhang & Wu; RawTherapee's, tiled for darktable).  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_lmmse.c:
 *     iop/demosaic/lmmse.c :51-576   tile constants, limf, median3f, median9f, calc_gamma, lmmse_demosaic
 *     iop/demosaic.c       :1207-1213  the two gamma tables (the loop body that fills them)
 * develop/imageop_math.h :175-219 (FC) comes through gen_imageop_math.c.
 *
 * This translation unit is compiled WITHOUT OpenMP: the reference zeroes its six tile planes once per thread and carries them from tile
 * to tile (lmmse.c:166-175), and rows / columns a short last tile does not rewrite keep what the thread's previous tile left there --
 * the pixels within 4 of the frame's bottom and right edge then depend on which thread ran which tiles.  Pinned is the serial walk of
 * the tiles (ref_lmmse), and next to it a walk with the planes zeroed in front of every tile (ref_lmmse_fresh: the same lines, one
 * tile per call), which is what the oracle's fresh mode and the CUDA kernel reproduce.
 */
#undef _OPENMP
#include "ref_piece.h"
#include <stdio.h>
#include "gen_imageop_math.c"
#define INLINE inline
#define dt_control_log(...) do { } while(0)
#define _(s) s
static inline float *dt_pixelpipe_cache_alloc_align_float_cache(size_t n, int id) { (void)id; return aligned_alloc(64, ((n * sizeof(float) + 63) / 64) * 64); }
#define dt_pixelpipe_cache_free_align(p) free((void *)(p))
#include "system/mem_alloc.h"
#include "gen_lmmse.c"

/* iop/demosaic.c:1207-1213 */
void ref_lmmse_gamma_tables(float *gamma_in, float *gamma_out)
{
  struct { float *lmmse_gamma_in, *lmmse_gamma_out; } g = { gamma_in, gamma_out }, *gd = &g;
#include "gen_lmmse_tables.c"
}

void ref_lmmse(float *out, const float *in, int width, int height, uint32_t filters, int mode, const float processed_maximum[3])
{
  static float *gin = NULL, *gout = NULL;
  if(!gin)
  {
    gin = malloc(65536 * sizeof(float));
    gout = malloc(65536 * sizeof(float));
    ref_lmmse_gamma_tables(gin, gout);
  }
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  for(int k = 0; k < 3; k++) piece.dsc_in.processed_maximum[k] = processed_maximum[k];
  dt_iop_roi_t roi_in = { 0, 0, width, height, 1.0 }, roi_out = roi_in;
  lmmse_demosaic(&piece, out, in, &roi_out, &roi_in, filters, (uint32_t)mode, gin, gout);
}
