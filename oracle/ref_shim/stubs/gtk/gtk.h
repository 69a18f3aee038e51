/* Stand-in for <gtk/gtk.h>: opaque GUI types only.  TEST INFRASTRUCTURE ONLY (oracle/_ref).
 * The reference mixes GUI declarations into headers its pixel code includes; no GTK function is called
 * by anything oracle/_ref executes. */
#ifndef B200_ORACLE_GTK_STUB_H
#define B200_ORACLE_GTK_STUB_H
#include <glib.h>
typedef struct _GtkWidget GtkWidget;
typedef struct _GtkLabel GtkLabel;
typedef struct _GtkBox GtkBox;
typedef struct _GtkNotebook GtkNotebook;
typedef struct _GtkDrawingArea GtkDrawingArea;
typedef struct _GtkWindow GtkWindow;
typedef struct _GtkAccelGroup GtkAccelGroup;
typedef struct _GtkTreeModel GtkTreeModel;
typedef struct _GtkTreeIter { int stamp; void *a, *b, *c; } GtkTreeIter;
typedef struct _GtkStyleContext GtkStyleContext;
typedef struct _GtkAllocation { int x, y, width, height; } GtkAllocation;
typedef struct _GdkRGBA { double red, green, blue, alpha; } GdkRGBA;
typedef struct _GdkEvent GdkEvent;
typedef struct _GdkEventButton GdkEventButton;
typedef struct _GdkEventMotion GdkEventMotion;
typedef struct _GdkEventScroll GdkEventScroll;
typedef struct _GdkEventKey GdkEventKey;
typedef struct _GdkEventCrossing GdkEventCrossing;
typedef struct _GdkPixbuf GdkPixbuf;
typedef struct _GdkDevice GdkDevice;
typedef struct _cairo cairo_t;
typedef struct _cairo_surface cairo_surface_t;
typedef unsigned int GdkModifierType;
typedef int GtkOrientation;
typedef int GtkAlign;
typedef int GtkStateFlags;
typedef int GtkPositionType;
#endif
