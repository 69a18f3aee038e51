/* oracle/_ref wrapper: the reference's edge-aware a-trous wavelets.  TEST INFRASTRUCTURE ONLY.
 * Textually includes the unmodified /root/reference/src/pixel/eaw.c, which exports
 *   eaw_dn_decompose() :242-326  and  eaw_synthesize() :157-175
 * as plain C symbols; the harness calls them directly. */
#include <glib.h>
#ifdef REF_STRICT
#define __DT_CLONE_TARGETS__
#endif
#include "pixel/eaw.c"
