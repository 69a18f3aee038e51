/* b200iop.h -- C ABI of libb200iop.so: the Ansel develop-pixelpipe hot path on B200 (sm_100a).
 *
 * Drop-in boundary.  A reference module keeps its dt_iop_module_t
 * surface (src/iop/iop_api.h:80-351) in plain C and forwards the body of
 *     process()            iop_api.h:265-266   -> b200_<op>_process_host()   host pointers, H2D + kernels + D2H
 *     process_cl()         iop_api.h:292-293   -> b200_<op>_process_dev()    device pointers + stream, no copies
 *     tiling_callback()    iop_api.h:121-122   -> b200_<op>_tiling()
 * to this library.  ansel_b200/iop/ holds those C adapters; INTEGRATION.md shows where they slot
 * into the reference tree.  No C++/CUDA/torch type crosses this header: plain pointers, sizes,
 * and PODs whose layout mirrors the reference structs they stand for (each cited below).
 *
 * Return convention of every int-returning entry point: 0 = success, non-zero = error
 * (b200_last_error() describes it).  That is process()'s convention (develop/pixelpipe_cpu.c:117-133);
 * a process_cl() adapter returns `rc == 0` because that slot means TRUE = success
 * (develop/pixelpipe_gpu.c:358).
 *
 * Threading: entry points are re-entrant; up to four pipes may run the same module concurrently
 * (doc/reorganisation.md:77-81).  Device scratch is per calling thread and per device.
 * There is no CPU fallback anywhere in this library: without a CUDA device calls fail.
 */
#ifndef B200IOP_H
#define B200IOP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------------ */
enum
{
  B200_OK = 0,
  B200_ERR_CUDA = 1,        /* a CUDA runtime/driver call failed */
  B200_ERR_ARG = 2,         /* NULL / inconsistent arguments */
  B200_ERR_UNSUPPORTED = 3, /* valid in the reference, not built here (say so, never guess) */
  B200_ERR_NODEVICE = 4,    /* no usable sm_100 device */
  B200_ERR_NOMEM = 5
};

/* ---- mirrors of the reference's operator-surface types ----------------------------------- */

/* dt_iop_roi_t, src/pixel/format.h:48-52 (identical layout) */
typedef struct b200_roi_t
{
  int x, y, width, height;
  double scale;
} b200_roi_t;

/* dt_develop_tiling_t, src/develop/tiling.h:39-58 (identical layout) */
typedef struct b200_tiling_t
{
  float factor, factor_cl, maxbuf, maxbuf_cl;
  unsigned overhead, overlap, xalign, yalign;
} b200_tiling_t;

/* dt_dev_pixelpipe_type_t, src/develop/pixelpipe.h:41-45 (same values) */
enum
{
  B200_PIPE_NONE = 0,
  B200_PIPE_EXPORT = 1,
  B200_PIPE_FULL = 2,
  B200_PIPE_PREVIEW = 3,
  B200_PIPE_THUMBNAIL = 4
};

/* Exactly the fields the hot-path process() bodies read from `piece`, `pipe` and `self`
 * (SURVEY.md appendix D): dt_dev_pixelpipe_iop_t src/develop/pixelpipe_hb.h:101-166,
 * dt_iop_buffer_dsc_t src/pixel/format.h:80-119. */
typedef struct b200_piece_t
{
  b200_roi_t roi_in, roi_out;  /* piece->roi_in / roi_out */
  uint32_t filters;            /* piece->dsc_in.filters (sensor phase only; the ROI shift of
                                  develop/imageop.c:139-142 is applied inside the library) */
  uint8_t xtrans[6][6];        /* piece->dsc_in.xtrans */
  uint32_t channels;           /* piece->dsc_in.channels */
  float processed_maximum[4];  /* piece->dsc_in.processed_maximum */
  float wb_coeffs[4];          /* piece->dsc_in.temperature.coeffs */
  int buf_in_width, buf_in_height; /* piece->buf_in */
  int pipe_type;               /* pipe->type */
  int mask_display;            /* pipe->mask_display */
  double iscale;               /* pipe->iscale */
  float exif_iso;              /* self->dev->image_storage.exif_iso */
  uint32_t image_flags;        /* self->dev->image_storage.flags */
  int devid;                   /* pipe->devid: CUDA device ordinal, < 0 = current device */
  const void *data;            /* piece->data: the module's b200_<op>_data_t */
  size_t data_size;            /* piece->data_size */
} b200_piece_t;

/* ---- library lifetime -------------------------------------------------------------------- */
int b200_abi_version(void);
/* Bind up to ndev devices (0 = all visible).  Fails with B200_ERR_NODEVICE when there is no
 * CUDA device -- there is no CPU path behind this ABI.  Stands where dt_opencl_init() does
 * (src/common/opencl.c). */
int b200_init(int ndev);
void b200_shutdown(void);
int b200_device_count(void);
/* thread-local, never NULL */
const char *b200_last_error(void);

/* integer CFA phase: dt_dev_get_roi_filters() develop/imageop.c:139-142 ->
 * dt_rawspeed_crop_dcraw_filters() imageio/imageio_rawspeed.cc:146-151 ->
 * ColorFilterArray::shiftDcrawFilter() external/rawspeed/.../ColorFilterArray.cpp:143-170 */
uint32_t b200_roi_filters(uint32_t filters, int roi_x, int roi_y);
/* FC(), develop/imageop_math.h:190-193 */
int b200_fc(int row, int col, uint32_t filters);

/* ---- demosaic (src/iop/demosaic.c) -------------------------------------------------------- */
/* dt_iop_demosaic_method_t values used here, iop/demosaic.c:109-135 */
enum
{
  B200_DEMOSAIC_PPG = 0,
  B200_DEMOSAIC_AMAZE = 1,
  B200_DEMOSAIC_VNG4 = 2,
  B200_DEMOSAIC_RCD = 5,
  B200_DEMOSAIC_LMMSE = 6
};
/* dt_iop_demosaic_data_t, iop/demosaic.c:238-247 (identical layout) */
typedef struct b200_demosaic_data_t
{
  uint32_t green_eq;
  uint32_t color_smoothing;
  uint32_t demosaicing_method;
  uint32_t lmmse_refine;
  float median_thrs;
  double CAM_to_RGB[3][4];
  float dual_thrs;
} b200_demosaic_data_t;

/* process(), iop/demosaic.c:1043-1253: in = 1-channel float mosaic roi_in, out = RGBA float roi_out */
int b200_demosaic_process_host(const b200_piece_t *piece, const void *in, void *out);
/* process_cl() slot, iop/demosaic/rcd.c:568-850: device pointers, `stream` is a cudaStream_t (NULL = default) */
int b200_demosaic_process_dev(const b200_piece_t *piece, const void *d_in, void *d_out, void *stream);
/* tiling_callback(), iop/demosaic.c:1916-2013 */
void b200_demosaic_tiling(const b200_piece_t *piece, b200_tiling_t *tiling);

#ifdef __cplusplus
}
#endif
#endif /* B200IOP_H */
