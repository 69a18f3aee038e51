/* Stand-in for <lcms2.h> (LittleCMS 2 is not installed in this image).  TEST INFRASTRUCTURE ONLY.
 * oracle/_ref compiles the reference's colorprofiles/conversion.c to reach its matrix/TRC apply
 * loops (_apply_matrix, _apply_target_curves); the profile-building half of that file calls
 * lcms2, which is only declared here and resolved to aborting stubs in ref_color.c -- the
 * harness never enters it (it fills dt_colorspaces_conversion_t directly). */
#ifndef B200_ORACLE_LCMS2_STUB_H
#define B200_ORACLE_LCMS2_STUB_H
#include <stdint.h>
#include <stddef.h>
typedef void *cmsHPROFILE;
typedef void *cmsHTRANSFORM;
typedef void *cmsContext;
typedef void cmsToneCurve;
typedef uint32_t cmsUInt32Number;
typedef int32_t cmsInt32Number;
typedef uint16_t cmsUInt16Number;
typedef uint8_t cmsUInt8Number;
typedef double cmsFloat64Number;
typedef float cmsFloat32Number;
typedef int cmsBool;
typedef uint32_t cmsTagSignature;
typedef uint32_t cmsColorSpaceSignature;
typedef uint32_t cmsProfileClassSignature;
typedef uint32_t cmsInfoType;
typedef struct { double X, Y, Z; } cmsCIEXYZ;
typedef struct { double x, y, Y; } cmsCIExyY;
typedef struct { cmsCIExyY Red, Green, Blue; } cmsCIExyYTRIPLE;
typedef struct { cmsCIEXYZ Red, Green, Blue; } cmsCIEXYZTRIPLE;
typedef struct { double L, a, b; } cmsCIELab;
typedef void cmsMLU;
#define TYPE_RGBA_FLT 1
#define TYPE_RGB_FLT 2
#define TYPE_LabA_FLT 3
#define TYPE_Lab_FLT 4
#define TYPE_XYZA_FLT 5
#define TYPE_XYZ_FLT 6
#define TYPE_RGB_DBL 7
#define TYPE_XYZ_DBL 8
#define TYPE_Lab_DBL 9
#define TYPE_RGBA_8 10
#define TYPE_RGB_8 11
#define TYPE_BGRA_8 12
#define TYPE_RGB_16 13
#define TYPE_RGBA_16 14
#define INTENT_PERCEPTUAL 0
#define INTENT_RELATIVE_COLORIMETRIC 1
#define INTENT_SATURATION 2
#define INTENT_ABSOLUTE_COLORIMETRIC 3
#define cmsFLAGS_NOCACHE 0x0040
#define cmsFLAGS_NOOPTIMIZE 0x0100
#define cmsFLAGS_SOFTPROOFING 0x4000
#define cmsFLAGS_GAMUTCHECK 0x1000
#define cmsFLAGS_BLACKPOINTCOMPENSATION 0x2000
#define cmsFLAGS_COPY_ALPHA 0x04000000
#define cmsSigRedColorantTag 0x7258595A
#define cmsSigGreenColorantTag 0x6758595A
#define cmsSigBlueColorantTag 0x6258595A
#define cmsSigRedTRCTag 0x72545243
#define cmsSigGreenTRCTag 0x67545243
#define cmsSigBlueTRCTag 0x62545243
#define cmsSigMediaWhitePointTag 0x77747074
#define cmsSigRgbData 0x52474220
#define cmsSigLabData 0x4C616220
#define cmsSigXYZData 0x58595A20
#define cmsSigGrayData 0x47524159
#define cmsSigDisplayClass 0x6D6E7472
#define cmsSigInputClass 0x73636E72
#define cmsSigOutputClass 0x70727472
#define cmsInfoDescription 0
#define LCMS_USED_AS_INPUT 0
#define LCMS_USED_AS_OUTPUT 1
#define LCMS_USED_AS_PROOF 2
#endif
