/* oracle/_ref link stubs.  TEST INFRASTRUCTURE ONLY.
 * The reference translation units compiled by oracle/_ref (ref_color.c -> colorprofiles/conversion.c)
 * reference lcms2 and other parts of lib_ansel from code the harness never enters.  They are
 * resolved here, in a translation unit that sees none of the reference's prototypes, to stubs
 * that abort loudly if ever reached. */
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
typedef void *cmsHTRANSFORM;
typedef void *cmsHPROFILE;

#define REF_ABORT(name)                                                     \
  do                                                                        \
  {                                                                         \
    fprintf(stderr, "oracle/_ref: %s reached -- not part of the pinned path\n", name); \
    abort();                                                                \
  } while(0)

/* --- names conversion.c needs from the rest of lib_ansel / lcms2, never called here --------- */
cmsHTRANSFORM cmsCreateProofingTransform() { REF_ABORT("cmsCreateProofingTransform"); }
cmsHTRANSFORM cmsCreateTransform() { REF_ABORT("cmsCreateTransform"); }
void cmsDeleteTransform() { REF_ABORT("cmsDeleteTransform"); }
int cmsGetColorSpace() { REF_ABORT("cmsGetColorSpace"); }
cmsHPROFILE cmsOpenProfileFromMem() { REF_ABORT("cmsOpenProfileFromMem"); }
int cmsSaveProfileToMem() { REF_ABORT("cmsSaveProfileToMem"); }
void *dt_colorprofiles_get_settings() { REF_ABORT("dt_colorprofiles_get_settings"); }
void dt_colorspaces_cleanup_profile() { REF_ABORT("dt_colorspaces_cleanup_profile"); }
int dt_colorspaces_get_matrix_from_input_profile() { REF_ABORT("dt_colorspaces_get_matrix_from_input_profile"); }
int dt_colorspaces_get_matrix_from_output_profile() { REF_ABORT("dt_colorspaces_get_matrix_from_output_profile"); }
void *dt_colorspaces_get_profile() { REF_ABORT("dt_colorspaces_get_profile"); }
void dt_colorspaces_lock_profile() { REF_ABORT("dt_colorspaces_lock_profile"); }
void dt_colorspaces_unlock_profile() { REF_ABORT("dt_colorspaces_unlock_profile"); }
void dt_colorspaces_transform_rgba_float_row() { REF_ABORT("dt_colorspaces_transform_rgba_float_row"); }
void dt_ioppr_init_unbounded_coeffs() { REF_ABORT("dt_ioppr_init_unbounded_coeffs"); }
void dt_print(unsigned int flags, const char *fmt, ...) { (void)flags; (void)fmt; }
void *dt_alloc_align(size_t size) { return aligned_alloc(64, (size + 63) / 64 * 64); }

