/* Shadows src/iop/iop_api.h for the oracle/_ref build only.  TEST INFRASTRUCTURE ONLY.
 * The real header is an X-macro list of module entry points that includes cairo and GTK; the pixel
 * translation units compiled by oracle/_ref (pixel/nlmeans_core.c) include it without using it. */
