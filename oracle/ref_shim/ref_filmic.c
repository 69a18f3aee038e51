/* oracle/_ref wrapper: filmic rgb (AgX colour science and its helpers).  TEST INFRASTRUCTURE ONLY.
 *
 * iop/filmicrgb.c carries its GTK GUI in the same translation unit and cannot be compiled whole
 * here.  oracle/Makefile cuts its pixel and parameter code out verbatim (oracle/ref_shim/slice.py)
 * into oracle/_ref/gen_filmicrgb.c:
 *     :96-110     constants                        :146-274   enums, spline struct, params struct
 *     :360-400    dt_iop_filmicrgb_data_t          :466-684   spline geometry helpers
 *     :948-2649   every pixel function (norms, log/spline tone mapping, Ych, gamut mapping, v1..v5, AgX)
 *     :3665-4113  filmic_sigmoid_scale, dt_iop_filmic_rgb_compute_spline, commit_params
 * This file declares the handful of names that cut expects and exposes plain-C entry points.
 */
#include <glib.h>
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#include <assert.h>
#include "system/macros.h"
#include "system/mem_alloc.h"
#include "system/openmp.h"
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#else
#include "system/target_clones.h"
#endif
#include "system/simd.h"
#include "math/math.h"
#include "math/matrices.h"
#include "pixel/format.h"
#include "common/colorspaces_inline_conversions.h"
#include "pixel/chromatic_adaptation.h"
#include "colorprofiles/iop_profile.h"
#include "iop/noise_generator.h"
#include "math/gaussian_elimination.h"

/* --- names from develop/, caches/, common/ the cut refers to ---------------------------------- */
typedef void dt_iop_params_t;
typedef struct dt_iop_module_t { int dummy; } dt_iop_module_t;
typedef enum { DT_DEV_PIXELPIPE_NONE = 0, DT_DEV_PIXELPIPE_EXPORT = 1, DT_DEV_PIXELPIPE_FULL = 2 } dt_dev_pixelpipe_type_t;
typedef struct dt_dev_pixelpipe_t { dt_dev_pixelpipe_type_t type; float iscale; } dt_dev_pixelpipe_t;
typedef struct dt_dev_pixelpipe_iop_t { void *data; dt_iop_roi_t buf_in, buf_out; } dt_dev_pixelpipe_iop_t;
typedef struct dt_colorprofiles_settings_t
{
  dt_colorspaces_color_mode_t mode;
  dt_colorspaces_color_profile_type_t softproof_type;
  char softproof_filename[512];
  dt_iop_color_intent_t softproof_intent;
} dt_colorprofiles_settings_t;
static void dt_colorprofiles_get_settings(dt_colorprofiles_settings_t *s) { memset(s, 0, sizeof(*s)); s->mode = DT_PROFILE_NORMAL; }
#define g_strlcpy(d, s, n) strncpy((d), (s), (n))
#define DT_PIXEL_APPLY_DPI(x) (x)
#define dt_control_log(...) ((void)0)
#define dt_print(...) ((void)0)
#define _(s) (s)
static inline float dt_dev_get_module_scale(const dt_dev_pixelpipe_t *pipe, const dt_iop_roi_t *roi) { return pipe->iscale / roi->scale; }
/* caches/pixelpipe_cache_alloc.h is self-contained; its three extern helpers are defined in ref_nlm.c */
#include "caches/pixelpipe_cache_alloc.h"
#include "pixel/bspline.h"

#include "gen_filmicrgb.c"

/* ---- plain-C entry points ------------------------------------------------------------------- */
size_t ref_filmic_sizeof_params(void) { return sizeof(dt_iop_filmicrgb_params_t); }
size_t ref_filmic_sizeof_data(void) { return sizeof(dt_iop_filmicrgb_data_t); }
size_t ref_filmic_offsetof_spline(void) { return offsetof(dt_iop_filmicrgb_data_t, spline); }
size_t ref_filmic_offsetof_noise_distribution(void) { return offsetof(dt_iop_filmicrgb_data_t, noise_distribution); }
size_t ref_filmic_sizeof_spline(void) { return sizeof(dt_iop_filmic_rgb_spline_t); }

/* the $DEFAULT values of dt_iop_filmicrgb_params_t, filmicrgb.c:244-274 */
void ref_filmic_default_params(void *out)
{
  dt_iop_filmicrgb_params_t p;
  memset(&p, 0, sizeof(p));
  p.grey_point_source = 18.45f;
  p.black_point_source = -8.0f;
  p.white_point_source = 4.0f;
  p.reconstruct_threshold = 16.0f;
  p.reconstruct_feather = 3.0f;
  p.reconstruct_bloom_vs_details = 100.0f;
  p.reconstruct_grey_vs_color = 100.0f;
  p.reconstruct_structure_vs_texture = 100.0f;
  p.security_factor = 0.0f;
  p.grey_point_target = 18.45f;
  p.black_point_target = 0.01517634f;
  p.white_point_target = 100.0f;
  p.output_power = 4.0f;
  p.latitude = 10.0f;
  p.contrast = 1.18f;
  p.saturation = 0.0f;
  p.balance = 0.0f;
  p.noise_level = 0.05f;
  p.preserve_color = DT_FILMIC_METHOD_MAX_RGB;
  p.version = DT_FILMIC_COLORSCIENCE_V8;
  p.auto_hardness = TRUE;
  p.custom_grey = FALSE;
  p.high_quality_reconstruction = 1;
  p.noise_distribution = DT_NOISE_POISSONIAN;
  p.shadows = DT_FILMIC_CURVE_SIGMOID;
  p.highlights = DT_FILMIC_CURVE_SIGMOID;
  p.compensate_icc_black = FALSE;
  p.spline_version = DT_FILMIC_SPLINE_VERSION_V3;
  memcpy(out, &p, sizeof(p));
}

/* commit_params(), filmicrgb.c:4005-4113: params -> data, through the reference's own code */
void ref_filmic_commit(const void *params, void *data_out)
{
  dt_iop_filmicrgb_params_t p;
  memcpy(&p, params, sizeof(p));
  dt_iop_filmicrgb_data_t *d = aligned_alloc(64, ((sizeof(dt_iop_filmicrgb_data_t) + 63) / 64) * 64);
  memset(d, 0, sizeof(*d));
  dt_iop_module_t self = { 0 };
  dt_dev_pixelpipe_t pipe = { DT_DEV_PIXELPIPE_EXPORT, 1.0f };
  dt_dev_pixelpipe_iop_t piece;
  memset(&piece, 0, sizeof(piece));
  piece.data = d;
  commit_params(&self, &p, &pipe, &piece);
  memcpy(data_out, d, sizeof(*d));
  free(d);
}

static void fill_profile(dt_iop_order_iccprofile_info_t *pr, const float min9[9], const float mout9[9])
{
  memset(pr, 0, sizeof(*pr));
  for(int i = 0; i < 3; i++)
    for(int j = 0; j < 3; j++)
    {
      pr->matrix_in[i][j] = min9[3 * i + j];
      pr->matrix_out[i][j] = mout9[3 * i + j];
    }
}

/* the AgX branch of process(), filmicrgb.c:2846-2856 -> filmic_agx() :2495-2587.
 * work_*: RGB->XYZ(D50) and back of the pipe's work profile; export_*: same for the output profile
 * or NULL when it is not a matrix profile (use_output_profile = 0). */
void ref_filmic_agx(const float *in, float *out, size_t width, size_t height, const void *data,
                    const float work_in[9], const float work_out[9], const float *export_in, const float *export_out)
{
  dt_iop_filmicrgb_data_t *d = aligned_alloc(64, ((sizeof(dt_iop_filmicrgb_data_t) + 63) / 64) * 64);
  memcpy(d, data, sizeof(*d));
  dt_iop_order_iccprofile_info_t work, expo;
  fill_profile(&work, work_in, work_out);
  if(export_in) fill_profile(&expo, export_in, export_out);
  const float white_display = powf(d->spline.y[4], d->output_power);
  const float black_display = powf(d->spline.y[0], d->output_power);
  filmic_agx(in, out, &work, export_in ? &expo : NULL, d, d->spline, width, height, 4, black_display, white_display);
  free(d);
}

/* the per-call matrix set-up of filmic_agx(): filmic_v4_prepare_matrices :2033-2063 and
 * filmic_agx_prepare_bracket :2390-2459.  out: input, output, export_input, export_output, inset, outset,
 * three rows of four floats each. */
void ref_filmic_prepare(int version, const float work_in[9], const float work_out[9], const float *export_in,
                        const float *export_out, float out[72])
{
  dt_iop_order_iccprofile_info_t work, expo;
  fill_profile(&work, work_in, work_out);
  if(export_in) fill_profile(&expo, export_in, export_out);
  dt_colormatrix_t M[6];
  memset(M, 0, sizeof(M));
  filmic_v4_prepare_matrices(M[0], M[1], M[2], M[3], &work, export_in ? &expo : NULL);
  filmic_agx_prepare_bracket(&work, (dt_iop_filmicrgb_colorscience_type_t)version, M[4], M[5]);
  for(int k = 0; k < 6; k++)
    for(int i = 0; i < 3; i++)
      for(int j = 0; j < 4; j++) out[12 * k + 4 * i + j] = M[k][i][j];
}

/* the non-AgX branches of process(), filmicrgb.c:2857-2887: filmic_v5 and the v1..v4 colour sciences with and
 * without chroma preservation.  Same profile conventions as ref_filmic_agx(). */
int ref_filmic_legacy(const float *in, float *out, size_t width, size_t height, const void *data, const float work_in[9],
                      const float work_out[9], const float *export_in, const float *export_out)
{
  dt_iop_filmicrgb_data_t *data_ = aligned_alloc(64, ((sizeof(dt_iop_filmicrgb_data_t) + 63) / 64) * 64);
  memcpy(data_, data, sizeof(*data_));
  const dt_iop_filmicrgb_data_t *const d = data_;
  dt_iop_order_iccprofile_info_t work_, expo_;
  fill_profile(&work_, work_in, work_out);
  if(export_in) fill_profile(&expo_, export_in, export_out);
  const dt_iop_order_iccprofile_info_t *const work_profile = &work_, *const export_profile = export_in ? &expo_ : NULL;
  const size_t ch = 4;
  const float white_display = powf(d->spline.y[4], d->output_power);
  const float black_display = powf(d->spline.y[0], d->output_power);
  int rc = 0;
  if(d->version == DT_FILMIC_COLORSCIENCE_V5)
    filmic_v5(in, out, work_profile, export_profile, d, d->spline, width, height, ch, black_display, white_display);
  else if(d->preserve_color == DT_FILMIC_METHOD_NONE)
  {
    if(d->version == DT_FILMIC_COLORSCIENCE_V1)
      filmic_split_v1(in, out, work_profile, d, d->spline, width, height);
    else if(d->version == DT_FILMIC_COLORSCIENCE_V2 || d->version == DT_FILMIC_COLORSCIENCE_V3)
      filmic_split_v2_v3(in, out, work_profile, d, d->spline, width, height);
    else if(d->version == DT_FILMIC_COLORSCIENCE_V4)
      filmic_split_v4(in, out, work_profile, export_profile, d, d->spline, d->preserve_color, width, height, ch, d->version, black_display,
                      white_display);
    else
      rc = 1;
  }
  else
  {
    if(d->version == DT_FILMIC_COLORSCIENCE_V1)
      filmic_chroma_v1(in, out, work_profile, d, d->spline, d->preserve_color, width, height);
    else if(d->version == DT_FILMIC_COLORSCIENCE_V2 || d->version == DT_FILMIC_COLORSCIENCE_V3)
      filmic_chroma_v2_v3(in, out, work_profile, d, d->spline, d->preserve_color, width, height, ch, d->version);
    else if(d->version == DT_FILMIC_COLORSCIENCE_V4)
      filmic_chroma_v4(in, out, work_profile, export_profile, d, d->spline, d->preserve_color, width, height, ch, d->version, black_display,
                       white_display);
    else
      rc = 1;
  }
  free(data_);
  return rc;
}

/* The highlight reconstruction in front of the tone mapping, process() filmicrgb.c:2729-2838, replayed on the cut
 * functions (mask_clipped_pixels :1201, inpaint_noise :1230, reconstruct_highlights :1430, compute_ratios :2604,
 * restore_ratios :2622).  Returns 1 with the reconstructed frame in `out`, 0 when fewer than 10 pixels are clipped
 * (out = in: tone mapping reads the input), -1 on allocation failure.  mask_out (optional): the clipping mask. */
int ref_filmic_reconstruct(const float *in, float *out, float *mask_out, size_t width, size_t height, const void *data, float iscale,
                           double roi_scale, int buf_w, int buf_h)
{
  dt_iop_filmicrgb_data_t *data_ = aligned_alloc(64, ((sizeof(dt_iop_filmicrgb_data_t) + 63) / 64) * 64);
  memcpy(data_, data, sizeof(*data_));
  const dt_iop_filmicrgb_data_t *const d = data_;
  dt_dev_pixelpipe_t pipe_ = { DT_DEV_PIXELPIPE_EXPORT, iscale };
  const dt_dev_pixelpipe_t *const pipe = &pipe_;
  dt_dev_pixelpipe_iop_t piece_;
  memset(&piece_, 0, sizeof(piece_));
  piece_.data = data_;
  piece_.buf_in = (dt_iop_roi_t){ 0, 0, buf_w, buf_h, 1.0 };
  piece_.buf_out = piece_.buf_in;
  const dt_dev_pixelpipe_iop_t *const piece = &piece_;
  dt_iop_roi_t roi = { 0, 0, (int)width, (int)height, roi_scale };
  const dt_iop_roi_t *const roi_in = &roi, *const roi_out = &roi;
  const size_t ch = 4, npx = width * height;
  int rc = 0;
  float *mask = aligned_alloc(64, ((npx * sizeof(float) + 63) / 64) * 64);
  float *inpainted = aligned_alloc(64, npx * 4 * sizeof(float));
  float *norms = aligned_alloc(64, ((npx * sizeof(float) + 63) / 64) * 64);
  float *ratios = aligned_alloc(64, npx * 4 * sizeof(float));
  float *reconstructed = out;
  const float scale = fmaxf(dt_dev_get_module_scale(pipe, roi_in), 1.f);
  const int recover_highlights = mask_clipped_pixels(in, mask, d->normalize, d->reconstruct_feather, roi_out->width, roi_out->height, 4);
  if(mask_out) memcpy(mask_out, mask, npx * sizeof(float));
  if(recover_highlights)
  {
    rc = 1;
    inpaint_noise(in, mask, inpainted, d->noise_level / scale, d->reconstruct_threshold, d->noise_distribution, roi_out->width, roi_out->height);
    if(reconstruct_highlights(pipe, inpainted, mask, reconstructed, DT_FILMIC_RECONSTRUCT_RGB, ch, d, piece, roi_in, roi_out)) rc = -1;
    if(rc == 1 && d->high_quality_reconstruction > 0)
      for(int i = 0; i < d->high_quality_reconstruction; i++)
      {
        compute_ratios(reconstructed, norms, ratios, NULL, DT_FILMIC_METHOD_EUCLIDEAN_NORM_V1, roi_out->width, roi_out->height);
        if(reconstruct_highlights(pipe, ratios, mask, reconstructed, DT_FILMIC_RECONSTRUCT_RATIOS, ch, d, piece, roi_in, roi_out))
        {
          rc = -1;
          break;
        }
        restore_ratios(reconstructed, norms, roi_out->width, roi_out->height);
      }
  }
  else
    memcpy(out, in, npx * 4 * sizeof(float));
  free(mask);
  free(inpainted);
  free(norms);
  free(ratios);
  free(data_);
  return rc;
}
