/* oracle/_ref wrapper: rawprepare (black/white normalisation of the sensor data, the first module of the pipe).
 * TEST INFRASTRUCTURE ONLY.
 *
 * iop/rawprepare.c carries GUI and database code; oracle/Makefile cuts verbatim into oracle/_ref/gen_rawprepare.c:
 *     :94-111   dt_iop_rawprepare_data_t       :206-210  compute_proper_crop
 *     :413-418  BL                             :466-633  process()
 * common/dng_opcode.h's dt_dng_gain_map_t (:37-55) is restated below because that header pulls common/image.h.
 */
#include "ref_piece.h"
typedef struct dt_dng_gain_map_t
{
  uint32_t top, left, bottom, right, plane, planes, row_pitch, col_pitch, map_points_v, map_points_h;
  double map_spacing_v, map_spacing_h, map_origin_v, map_origin_h;
  uint32_t map_planes;
  float map_gain[];
} dt_dng_gain_map_t;
#define process rawprepare_process
#include "gen_rawprepare.c"
#undef process

/* sub/div: the four per-CFA-site black levels and (white - black) ranges commit_params() leaves (:722-750);
 * gain: NULL or 4 maps of map_w*map_h floats with the given geometry (origin and spacing relative to the full image) */
int ref_rawprepare(const void *in, void *out, int in_width, int in_height, int out_x, int out_y, int out_width, int out_height,
                   double roi_scale, int crop_x, int crop_y, const float sub[4], const float div[4], uint32_t filters, int channels,
                   int datatype, int buf_w, int buf_h, const float *gain, int map_w, int map_h, double spacing_h, double spacing_v,
                   double origin_h, double origin_v)
{
  dt_iop_rawprepare_data_t d;
  memset(&d, 0, sizeof(d));
  d.x = crop_x;
  d.y = crop_y;
  for(int k = 0; k < 4; k++)
  {
    d.sub[k] = sub[k];
    d.div[k] = div[k];
  }
  dt_dng_gain_map_t *maps[4] = { 0 };
  if(gain)
  {
    d.apply_gainmaps = 1;
    for(int f = 0; f < 4; f++)
    {
      maps[f] = calloc(1, sizeof(dt_dng_gain_map_t) + sizeof(float) * map_w * map_h);
      maps[f]->map_points_h = map_w;
      maps[f]->map_points_v = map_h;
      maps[f]->map_spacing_h = spacing_h;
      maps[f]->map_spacing_v = spacing_v;
      maps[f]->map_origin_h = origin_h;
      maps[f]->map_origin_v = origin_v;
      memcpy(maps[f]->map_gain, gain + (size_t)f * map_w * map_h, sizeof(float) * map_w * map_h);
      d.gainmaps[f] = maps[f];
    }
  }
  dt_dev_pixelpipe_t pipe = { 1, 0, 1.0f, 0 };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = &d;
  piece.buf_in = (dt_iop_roi_t){ 0, 0, buf_w, buf_h, 1.0 };
  piece.roi_in = (dt_iop_roi_t){ 0, 0, in_width, in_height, roi_scale };
  piece.roi_out = (dt_iop_roi_t){ out_x, out_y, out_width, out_height, roi_scale };
  piece.dsc_in.filters = filters;
  piece.dsc_in.channels = channels;
  piece.dsc_in.datatype = datatype;
  const int rc = rawprepare_process(NULL, &pipe, &piece, in, out);
  for(int f = 0; f < 4; f++) free(maps[f]);
  return rc;
}
size_t ref_rawprepare_sizeof_data(void) { return sizeof(dt_iop_rawprepare_data_t); }
