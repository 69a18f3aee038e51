/* Minimal stand-in for <glib.h>, used ONLY by the oracle/_ref build (test infrastructure).
 * The reference's low-level headers include glib for a handful of typedefs and macros; the
 * real glib is not installed in this image.  Nothing from glib's implementation is needed by
 * the pixel code that oracle/_ref compiles.  Not part of the product. */
#ifndef B200_ORACLE_GLIB_STUB_H
#define B200_ORACLE_GLIB_STUB_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

typedef int gboolean;
typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef size_t gsize;
typedef float gfloat;
typedef double gdouble;
typedef unsigned char guint8;
typedef uint16_t guint16;
typedef uint32_t guint32;
typedef uint64_t guint64;
typedef int32_t gint32;
typedef int64_t gint64;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef CLAMP
#define CLAMP(x, lo, hi) (((x) > (hi)) ? (hi) : (((x) < (lo)) ? (lo) : (x)))
#endif
#define G_GSIZE_FORMAT "zu"
#define G_LIKELY(x) __builtin_expect(!!(x), 1)
#define G_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define G_BEGIN_DECLS
#define G_END_DECLS
typedef struct _GList { void *data; struct _GList *next, *prev; } GList;
typedef struct _GHashTable GHashTable;
typedef struct _GSList { void *data; struct _GSList *next; } GSList;
#define g_malloc malloc
#define g_malloc0(n) calloc(1, (n))
#define g_free free
#endif
