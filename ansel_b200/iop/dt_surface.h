/* dt_surface.h -- the slice of the reference's operator surface the hot-path modules touch.
 *
 * In the reference tree these names come from src/develop/imageop.h, src/develop/pixelpipe_hb.h,
 * src/pixel/format.h and src/develop/tiling.h; a maintainer dropping the adapters of this
 * directory into src/iop/ includes those headers instead and deletes this file.
 * Only members the adapters read are declared; names and meaning follow the reference so the
 * adapter bodies compile unchanged against the real headers.
 */
#ifndef B200_DT_SURFACE_H
#define B200_DT_SURFACE_H
#include <stddef.h>
#include <stdint.h>
#include "b200iop.h"

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
typedef int gboolean;

typedef b200_roi_t dt_iop_roi_t;              /* src/pixel/format.h:48-52 */
typedef b200_tiling_t dt_develop_tiling_t;    /* src/develop/tiling.h:39-58 */
typedef void *cl_mem;                         /* the process_cl slot carries a device pointer */

/* src/pixel/format.h:80-119 */
typedef struct dt_iop_buffer_dsc_t
{
  unsigned int channels;
  int datatype;
  size_t bpp;
  uint32_t filters;
  uint8_t xtrans[6][6];
  struct { uint16_t raw_black_level, raw_white_point; } rawprepare;
  struct { int enabled; float coeffs[4] __attribute__((aligned(16))); } temperature;
  float processed_maximum[4] __attribute__((aligned(16)));
  int cst;
} dt_iop_buffer_dsc_t;

/* src/common/image.h: the two members the demosaic/denoise bodies read */
typedef struct dt_image_t
{
  float exif_iso;
  uint32_t flags;
} dt_image_t;

typedef struct dt_develop_t
{
  dt_image_t image_storage;
  int gui_attached;
} dt_develop_t;

/* src/colorprofiles/iop_profile.h:122-146: the two matrices filmic's gamut mapping reads */
typedef struct dt_iop_order_iccprofile_info_t
{
  float matrix_in[3][4] __attribute__((aligned(16)));  /* RGB -> XYZ(D50); NaN = not a matrix profile */
  float matrix_out[3][4] __attribute__((aligned(16))); /* XYZ(D50) -> RGB */
} dt_iop_order_iccprofile_info_t;

/* src/develop/pixelpipe_hb.h: dt_dev_pixelpipe_t */
typedef struct dt_dev_pixelpipe_t
{
  dt_develop_t *dev;
  int type;          /* dt_dev_pixelpipe_type_t */
  int mask_display;
  int devid;         /* device reserved for this pipe (dt_opencl_reserve_device_for_pipe analogue) */
  float iscale;
  void *stream;      /* per-pipe CUDA stream, NULL = default stream */
  dt_iop_order_iccprofile_info_t *work_profile_info;   /* pixelpipe_hb.h: set by colorin's commit */
  dt_iop_order_iccprofile_info_t *output_profile_info; /* set by colorout's commit */
} dt_dev_pixelpipe_t;
/* src/colorprofiles/iop_profile.c: the two getters filmic's process() calls (filmicrgb.c:2714-2715) */
static inline const dt_iop_order_iccprofile_info_t *dt_ioppr_get_pipe_work_profile_info(const dt_dev_pixelpipe_t *pipe)
{
  return pipe->work_profile_info;
}
static inline const dt_iop_order_iccprofile_info_t *dt_ioppr_get_pipe_output_profile_info(const dt_dev_pixelpipe_t *pipe)
{
  return pipe->output_profile_info;
}

struct dt_iop_module_t;
/* src/develop/pixelpipe_hb.h:101-166 */
typedef struct dt_dev_pixelpipe_iop_t
{
  struct dt_iop_module_t *module;
  void *data;
  size_t data_size;
  int enabled;
  dt_iop_roi_t buf_in, buf_out;
  dt_iop_roi_t roi_in, roi_out;
  int process_cl_ready;
  int process_tiling_ready;
  dt_iop_buffer_dsc_t dsc_in, dsc_out;
} dt_dev_pixelpipe_iop_t;

/* src/develop/imageop.h:226-377 */
typedef struct dt_iop_module_t
{
  char op[20];
  dt_develop_t *dev;
  void *global_data;
} dt_iop_module_t;

/* build the C-ABI view of a piece: exactly the fields SURVEY.md appendix D lists */
static inline void b200_piece_from_dt(b200_piece_t *p, const dt_iop_module_t *self, const dt_dev_pixelpipe_t *pipe,
                                      const dt_dev_pixelpipe_iop_t *piece)
{
  p->roi_in = piece->roi_in;
  p->roi_out = piece->roi_out;
  p->filters = piece->dsc_in.filters;
  for(int i = 0; i < 6; i++)
    for(int j = 0; j < 6; j++) p->xtrans[i][j] = piece->dsc_in.xtrans[i][j];
  p->channels = piece->dsc_in.channels;
  p->datatype = piece->dsc_in.datatype;
  for(int k = 0; k < 4; k++)
  {
    p->processed_maximum[k] = piece->dsc_in.processed_maximum[k];
    p->wb_coeffs[k] = piece->dsc_in.temperature.coeffs[k];
  }
  p->buf_in_width = piece->buf_in.width;
  p->buf_in_height = piece->buf_in.height;
  p->pipe_type = pipe->type;
  p->mask_display = pipe->mask_display;
  p->iscale = pipe->iscale;
  const dt_develop_t *dev = self && self->dev ? self->dev : pipe->dev;
  p->exif_iso = dev ? dev->image_storage.exif_iso : 0.0f;
  p->image_flags = dev ? dev->image_storage.flags : 0u;
  p->devid = pipe->devid;
  p->data = piece->data;
  p->data_size = piece->data_size;
}
#endif
