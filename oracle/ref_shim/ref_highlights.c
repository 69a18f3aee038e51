/* oracle/_ref wrapper: highlights (clip mode and the "nothing to reconstruct" bypass).  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile cuts verbatim into oracle/_ref/gen_highlights.c:
 *     iop/highlights/common.h :218 (DT_HL_MIN_CLIPPED_PIXELS), :431-476 (mode enum, params == data)
 *     iop/highlights/clip.c   :60-85  process_clip
 *     iop/highlights/common.h :618-619 SQRT3, SQRT12 (long double);  iop/highlights/lch.c :315-411 process_lch_bayer
 *     iop/highlights/lch.c    :65-204 interp_pix_xtrans, interpolate_color_xtrans;  :412-537 process_lch_xtrans
 *     iop/highlights/lch.c    :206-303 interpolate_color;  iop/highlights/inpaint.c :63-82 process_inpaint_bayer
 *     iop/highlights.c        :232-302 _hl_count_thresholds, _hl_count_clipped, _hl_copy_input;  :679-789 process()
 * The other reconstruction modes ( guided Laplacians, harmonic transposition) are separate translation
 * units of 18 k lines that are not built here: their entry points abort (the tests reach them only through the bypass).
 */
#include "ref_piece.h"
#include "gen_imageop_math.c"
#include <stdio.h>
static inline void dt_iop_image_copy_by_size(float *const out, const float *const in, const size_t width, const size_t height, const size_t ch)
{ /* common/imagebuf.h:91-95 -> dt_iop_image_copy: a copy */
  memcpy(out, in, sizeof(float) * width * height * ch);
}
typedef struct dt_iop_highlights_gui_data_t { int show_visualize; } dt_iop_highlights_gui_data_t;
#define dt_iop_gui_data(self) NULL
#define DT_DEV_PIXELPIPE_DISPLAY_PASSTHRU (1 << 12)
/* develop/imageop.c:139-142 -> imageio/imageio_rawspeed.cc:146-151 -> rawspeed's ColorFilterArray::shiftDcrawFilter
 * (third party, not under /root/reference/src; ColorFilterArray.cpp:143-170 restated): an odd x swaps the two colours of
 * every row pair of the 8x2 pattern word, y rotates it by four bits per row */
uint32_t ref_roi_filters(uint32_t filters, int x, int y)
{
  if(!filters || filters == 9u) return filters;
  if(abs(x) & 1)
    for(int n = 0; n < 8; n++)
    {
      const int i = n * 4, j = i + 2;
      const uint32_t t = ((filters >> i) ^ (filters >> j)) & 3u;
      filters ^= (t << i) | (t << j);
    }
  if(y == 0) return filters;
  y *= 4;
  y = y >= 0 ? y % 32 : 32 - ((-y) % 32);
  if(y != 0 && y != 32) filters = (filters >> y) | (filters << (32 - y));
  return filters;
}
static inline uint32_t dt_dev_get_roi_filters(const dt_dev_pixelpipe_iop_t *piece, const dt_iop_roi_t *roi)
{
  return ref_roi_filters(piece->dsc_in.filters, roi->x, roi->y);
}
#define NOT_BUILT(name) do { fprintf(stderr, "oracle/_ref: highlights %s is not built\n", name); abort(); } while(0)
#define process_visualize(...) NOT_BUILT("process_visualize")
static inline int process_laplacian_stub(void) { NOT_BUILT("process_laplacian"); return 1; }
#define process_laplacian(...) process_laplacian_stub()
#define process_harmonic(...) process_laplacian_stub()
#define process highlights_process
#include "gen_highlights.c"
#undef process

static uint8_t ref_highlights_xtrans[6][6];
/* the sensor's X-Trans table for the following ref_highlights() calls with filters == 9 */
void ref_highlights_set_xtrans(const uint8_t xtrans[36]) { memcpy(ref_highlights_xtrans, xtrans, 36); }
int ref_highlights(const float *in, float *out, int x, int y, int width, int height, uint32_t filters, int channels, int mode, float clip,
                   const float processed_maximum[4], int mask_display)
{
  dt_iop_highlights_data_t d;
  memset(&d, 0, sizeof(d));
  d.mode = mode;
  d.clip = clip;
  dt_dev_pixelpipe_t pipe = { 1, mask_display, 1.0f, 0 };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = &d;
  piece.roi_in = piece.roi_out = (dt_iop_roi_t){ x, y, width, height, 1.0 };
  piece.dsc_in.filters = filters;
  piece.dsc_in.channels = channels;
  for(int k = 0; k < 4; k++) piece.dsc_in.processed_maximum[k] = processed_maximum[k];
  memcpy(piece.dsc_in.xtrans, ref_highlights_xtrans, 36);
  return highlights_process(NULL, &pipe, &piece, in, out);
}
size_t ref_highlights_sizeof_data(void) { return sizeof(dt_iop_highlights_data_t); }
