/* oracle/_ref wrapper: highlights, "guided laplacians" mode (iop/highlights/laplacian.c).  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_hl_laplacian*.c:
 *     iop/highlights/common.h  :431-476 (mode enum, params == data), :615-630 (MAX_NUM_SCALES, DS_FACTOR, scale / variant enums)
 *     iop/highlights/gather.h  :66-79   dt_hl_cfa_t, _hl_cfa_strategy
 *     iop/highlights/gather.c  :66-541  _interpolate_and_mask (+ X-Trans, passthrough), _compute_laplacian_normalization,
 *                                       _remosaic_and_replace (+ X-Trans, passthrough)
 *     pixel/box_filters.c      :40-49 181-195 228-254 290-305 350-405 508-574 645-702 766-825 890-924 949-971: the four-channel box mean
 *     pixel/fast_guided_filter.h :98-155 interpolate_bilinear
 *     iop/highlights/laplacian.c :76-575 scale_type, guide_laplacians, heat_PDE_diffusion, wavelets_process, process_laplacian
 * pixel/bspline.h (decompose_2D_Bspline) and iop/noise_generator.h are included unmodified.
 *
 * _compute_laplacian_normalization() is an OpenMP float reduction: its value depends on the number of threads and on the order they
 * finish in (at 45 MP a single thread's sum stops growing once the addends fall under half an ulp of it).  The wrapper records the
 * vector the run used (ref_hl_laplacian's `normalization` out-argument), so that a test can hand the same vector to the code under
 * test; with `force` set the recorded vector is imposed instead.
 */
#include "ref_piece.h"
#include "gen_imageop_math.c"
#include "system/mem_alloc.h"
#include "math/openmp_maths.h"
#include "pixel/dwt.h"
#include "iop/noise_generator.h"
#include "caches/pixelpipe_cache_alloc.h"
#include "pixel/bspline.h"
#include <stdio.h>

static inline float dt_dev_get_module_scale(const dt_dev_pixelpipe_t *pipe, const dt_iop_roi_t *roi) { return pipe->iscale / roi->scale; }
/* develop/imageop.c:139-142 -> rawspeed's ColorFilterArray::shiftDcrawFilter (third party; restated in ref_highlights.c) */
uint32_t ref_roi_filters(uint32_t filters, int x, int y);
static inline uint32_t dt_dev_get_roi_filters(const dt_dev_pixelpipe_iop_t *piece, const dt_iop_roi_t *roi)
{
  return ref_roi_filters(piece->dsc_in.filters, roi->x, roi->y);
}

#include "gen_hl_laplacian_types.c"
static int dt_box_mean_4ch(float *const buf, const int height, const int width, const int radius, const unsigned iterations);
static inline int dt_box_mean(float *const buf, const size_t height, const size_t width, const int ch, const int radius,
                              const unsigned iterations)
{ /* pixel/box_filters.c:1050-1073: ch == 4 -> dt_box_mean_4ch */
  if(ch != 4) abort();
  return dt_box_mean_4ch(buf, (int)height, (int)width, radius, iterations);
}
#include "gen_hl_laplacian_box.c"
#include "gen_hl_laplacian_gather.c"

static float hl_norm_used[4];
static int hl_norm_force = 0;
static void hl_normalization_hook(const float *input, const dt_iop_roi_t *roi_in, uint32_t filters, const uint8_t (*xtrans)[6],
                                  dt_aligned_pixel_t normalization)
{
  if(hl_norm_force)
    for(int c = 0; c < 4; c++) normalization[c] = hl_norm_used[c];
  else
  {
    _compute_laplacian_normalization(input, roi_in, filters, xtrans, normalization);
    for(int c = 0; c < 4; c++) hl_norm_used[c] = normalization[c];
  }
}
#define _compute_laplacian_normalization hl_normalization_hook
#include "gen_hl_laplacian.c"
#undef _compute_laplacian_normalization

/* clips as process() builds them (iop/highlights.c:764-766); normalization: out (force == 0) or in (force != 0) */
int ref_hl_laplacian(const float *in, float *out, int x, int y, int width, int height, uint32_t filters, const uint8_t xtrans[36],
                     const float clips[4], int iterations, int scales, float noise_level, float solid_color, float iscale, float roi_scale,
                     float normalization[4], int force)
{
  dt_iop_highlights_data_t d;
  memset(&d, 0, sizeof(d));
  d.mode = DT_IOP_HIGHLIGHTS_LAPLACIAN;
  d.iterations = iterations;
  d.scales = scales;
  d.noise_level = noise_level;
  d.solid_color = solid_color;
  dt_dev_pixelpipe_t pipe = { 1, 0, iscale, 0 };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = &d;
  piece.roi_in = piece.roi_out = (dt_iop_roi_t){ x, y, width, height, roi_scale };
  piece.dsc_in.filters = filters;
  piece.dsc_in.channels = filters ? 1 : 4;
  if(xtrans) memcpy(piece.dsc_in.xtrans, xtrans, 36);
  hl_norm_force = force;
  if(force)
    for(int c = 0; c < 4; c++) hl_norm_used[c] = normalization[c];
  dt_aligned_pixel_t c4 = { clips[0], clips[1], clips[2], clips[3] };
  const int err = process_laplacian(NULL, &pipe, &piece, in, out, &piece.roi_in, &piece.roi_out, c4);
  for(int c = 0; c < 4; c++) normalization[c] = hl_norm_used[c];
  hl_norm_force = 0;
  return err;
}
