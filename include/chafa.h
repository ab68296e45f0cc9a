/* chafa.h -- public C API of the B200-native canvas renderer.
 *
 * Drop-in for the part of libchafa's API that drives the canvas-rendering path
 * (reference: hpjansson/chafa 1.19.0, chafa/chafa.h and the headers it pulls in).
 * Same type names, enumerator values, function names, argument meaning and
 * error behaviour (bad arguments are ignored, there is no error channel on
 * draw/print).  Behind it everything that touches pixels runs as sm_100a CUDA
 * kernels; there is no CPU fallback.
 *
 * Each declaration cites the reference interface it replaces.  GLib itself is
 * not required: when <glib.h> has not been included first, the handful of GLib
 * typedefs the API uses are declared here with identical layout.
 *
 * Extra, non-reference entry points (device-resident buffers, batching, bands
 * for multi-GPU) live in chafa_b200.h.
 */
#ifndef CHAFA_B200_CHAFA_H
#define CHAFA_B200_CHAFA_H

#include <stddef.h>
#include <stdint.h>
#include <stdarg.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ */
/* GLib-compatible basic types (only if GLib's own are not in scope)          */

#if !defined(__G_LIB_H__) && !defined(GLIBSHIM_GLIB_H)
typedef char gchar;
typedef int gint;
typedef gint gboolean;
typedef unsigned int guint;
typedef unsigned char guchar;
typedef uint8_t guint8;
typedef uint16_t guint16;
typedef uint32_t guint32;
typedef uint64_t guint64;
typedef int16_t gint16;
typedef int32_t gint32;
typedef int64_t gint64;
typedef float gfloat;
typedef double gdouble;
typedef size_t gsize;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef guint32 gunichar;
typedef guint32 GQuark;

/* Layout-compatible with GLib's GString / GError */
typedef struct _GString
{
    gchar *str;
    gsize len;
    gsize allocated_len;
}
GString;

typedef struct _GError
{
    GQuark domain;
    gint code;
    gchar *message;
}
GError;
#define CHAFA_B200_OWN_GLIB_TYPES 1

/* The handful of GLib calls a caller of this path makes on what the library returns
 * (examples/example.c:59 frees the printed GString with g_string_free).  Everything the library
 * hands out is malloc()ed, as GLib's own allocations are since 2.46. */
#include <stdlib.h>
#ifndef TRUE
# define TRUE 1
#endif
#ifndef FALSE
# define FALSE 0
#endif
static inline void
g_free (gpointer mem)
{
    free (mem);
}
static inline gchar *
g_string_free (GString *string, gboolean free_segment)
{
    gchar *segment;
    if (!string)
        return NULL;
    segment = string->str;
    if (free_segment)
    {
        free (segment);
        segment = NULL;
    }
    free (string);
    return segment;
}
static inline void
g_strfreev (gchar **str_array)
{
    size_t i;
    if (!str_array)
        return;
    for (i = 0; str_array [i]; i++)
        free (str_array [i]);
    free (str_array);
}
#endif

#ifndef CHAFA_API
# define CHAFA_API __attribute__ ((visibility ("default")))
#endif

#define CHAFA_MAJOR_VERSION 1
#define CHAFA_MINOR_VERSION 19
#define CHAFA_MICRO_VERSION 0

/* ------------------------------------------------------------------------ */
/* Enumerations (reference: chafa/chafa-common.h:48-230)                      */

typedef enum
{
    CHAFA_PIXEL_RGBA8_PREMULTIPLIED,
    CHAFA_PIXEL_BGRA8_PREMULTIPLIED,
    CHAFA_PIXEL_ARGB8_PREMULTIPLIED,
    CHAFA_PIXEL_ABGR8_PREMULTIPLIED,
    CHAFA_PIXEL_RGBA8_UNASSOCIATED,
    CHAFA_PIXEL_BGRA8_UNASSOCIATED,
    CHAFA_PIXEL_ARGB8_UNASSOCIATED,
    CHAFA_PIXEL_ABGR8_UNASSOCIATED,
    CHAFA_PIXEL_RGB8,
    CHAFA_PIXEL_BGR8,
    CHAFA_PIXEL_MAX
}
ChafaPixelType;

typedef enum
{
    CHAFA_ALIGN_START,
    CHAFA_ALIGN_END,
    CHAFA_ALIGN_CENTER,
    CHAFA_ALIGN_MAX
}
ChafaAlign;

typedef enum
{
    CHAFA_TUCK_STRETCH,
    CHAFA_TUCK_FIT,
    CHAFA_TUCK_SHRINK_TO_FIT,
    CHAFA_TUCK_MAX
}
ChafaTuck;

typedef enum
{
    CHAFA_COLOR_EXTRACTOR_AVERAGE,
    CHAFA_COLOR_EXTRACTOR_MEDIAN,
    CHAFA_COLOR_EXTRACTOR_MAX
}
ChafaColorExtractor;

typedef enum
{
    CHAFA_COLOR_SPACE_RGB,
    CHAFA_COLOR_SPACE_DIN99D,
    CHAFA_COLOR_SPACE_MAX
}
ChafaColorSpace;

typedef enum
{
    CHAFA_DITHER_MODE_NONE,
    CHAFA_DITHER_MODE_ORDERED,
    CHAFA_DITHER_MODE_DIFFUSION,
    CHAFA_DITHER_MODE_NOISE,
    CHAFA_DITHER_MODE_MAX
}
ChafaDitherMode;

typedef enum
{
    CHAFA_OPTIMIZATION_REUSE_ATTRIBUTES = (1 << 0),
    CHAFA_OPTIMIZATION_SKIP_CELLS = (1 << 1),
    CHAFA_OPTIMIZATION_REPEAT_CELLS = (1 << 2),
    CHAFA_OPTIMIZATION_NONE = 0,
    CHAFA_OPTIMIZATION_ALL = 0x7fffffff
}
ChafaOptimizations;

typedef enum
{
    CHAFA_CANVAS_MODE_TRUECOLOR,
    CHAFA_CANVAS_MODE_INDEXED_256,
    CHAFA_CANVAS_MODE_INDEXED_240,
    CHAFA_CANVAS_MODE_INDEXED_16,
    CHAFA_CANVAS_MODE_FGBG_BGFG,
    CHAFA_CANVAS_MODE_FGBG,
    CHAFA_CANVAS_MODE_INDEXED_8,
    CHAFA_CANVAS_MODE_INDEXED_16_8,
    CHAFA_CANVAS_MODE_MAX
}
ChafaCanvasMode;

typedef enum
{
    CHAFA_PIXEL_MODE_SYMBOLS,
    CHAFA_PIXEL_MODE_SIXELS,
    CHAFA_PIXEL_MODE_KITTY,
    CHAFA_PIXEL_MODE_ITERM2,
    CHAFA_PIXEL_MODE_MAX
}
ChafaPixelMode;

typedef enum
{
    CHAFA_PASSTHROUGH_NONE,
    CHAFA_PASSTHROUGH_SCREEN,
    CHAFA_PASSTHROUGH_TMUX,
    CHAFA_PASSTHROUGH_MAX
}
ChafaPassthrough;

/* ------------------------------------------------------------------------ */
/* Features / threads (reference: chafa/chafa-features.h:29-55)               */

typedef enum
{
    CHAFA_FEATURE_MMX = (1 << 0),
    CHAFA_FEATURE_SSE41 = (1 << 1),
    CHAFA_FEATURE_POPCNT = (1 << 2),
    CHAFA_FEATURE_AVX2 = (1 << 3)
}
ChafaFeatures;

CHAFA_API ChafaFeatures chafa_get_builtin_features (void);
CHAFA_API ChafaFeatures chafa_get_supported_features (void);
CHAFA_API gchar *chafa_describe_features (ChafaFeatures features);
/* The reference's worker threads are replaced by GPU work; the thread count is
 * kept as a stored setting only (chafa/chafa-features.c:226-282). */
CHAFA_API gint chafa_get_n_threads (void);
CHAFA_API void chafa_set_n_threads (gint n);
CHAFA_API gint chafa_get_n_actual_threads (void);

/* ------------------------------------------------------------------------ */
/* Symbol map (reference: chafa/chafa-symbol-map.h:33-126)                    */

#define CHAFA_SYMBOL_WIDTH_PIXELS 8
#define CHAFA_SYMBOL_HEIGHT_PIXELS 8

typedef enum
{
    CHAFA_SYMBOL_TAG_NONE        = 0,
    CHAFA_SYMBOL_TAG_SPACE       = (1 <<  0),
    CHAFA_SYMBOL_TAG_SOLID       = (1 <<  1),
    CHAFA_SYMBOL_TAG_STIPPLE     = (1 <<  2),
    CHAFA_SYMBOL_TAG_BLOCK       = (1 <<  3),
    CHAFA_SYMBOL_TAG_BORDER      = (1 <<  4),
    CHAFA_SYMBOL_TAG_DIAGONAL    = (1 <<  5),
    CHAFA_SYMBOL_TAG_DOT         = (1 <<  6),
    CHAFA_SYMBOL_TAG_QUAD        = (1 <<  7),
    CHAFA_SYMBOL_TAG_HHALF       = (1 <<  8),
    CHAFA_SYMBOL_TAG_VHALF       = (1 <<  9),
    CHAFA_SYMBOL_TAG_HALF        = ((1 << 8) | (1 << 9)),
    CHAFA_SYMBOL_TAG_INVERTED    = (1 << 10),
    CHAFA_SYMBOL_TAG_BRAILLE     = (1 << 11),
    CHAFA_SYMBOL_TAG_TECHNICAL   = (1 << 12),
    CHAFA_SYMBOL_TAG_GEOMETRIC   = (1 << 13),
    CHAFA_SYMBOL_TAG_ASCII       = (1 << 14),
    CHAFA_SYMBOL_TAG_ALPHA       = (1 << 15),
    CHAFA_SYMBOL_TAG_DIGIT       = (1 << 16),
    CHAFA_SYMBOL_TAG_ALNUM       = ((1 << 15) | (1 << 16)),
    CHAFA_SYMBOL_TAG_NARROW      = (1 << 17),
    CHAFA_SYMBOL_TAG_WIDE        = (1 << 18),
    CHAFA_SYMBOL_TAG_AMBIGUOUS   = (1 << 19),
    CHAFA_SYMBOL_TAG_UGLY        = (1 << 20),
    CHAFA_SYMBOL_TAG_LEGACY      = (1 << 21),
    CHAFA_SYMBOL_TAG_SEXTANT     = (1 << 22),
    CHAFA_SYMBOL_TAG_WEDGE       = (1 << 23),
    CHAFA_SYMBOL_TAG_LATIN       = (1 << 24),
    CHAFA_SYMBOL_TAG_IMPORTED    = (1 << 25),
    CHAFA_SYMBOL_TAG_OCTANT      = (1 << 26),
    CHAFA_SYMBOL_TAG_EXTRA       = (1 << 30),
    CHAFA_SYMBOL_TAG_BAD         = ((1 << 19) | (1 << 20)),
    CHAFA_SYMBOL_TAG_ALL         = ~((1 << 30) | (1 << 19) | (1 << 20))
}
ChafaSymbolTags;

typedef struct ChafaSymbolMap ChafaSymbolMap;

CHAFA_API ChafaSymbolMap *chafa_symbol_map_new (void);
CHAFA_API ChafaSymbolMap *chafa_symbol_map_copy (const ChafaSymbolMap *symbol_map);
CHAFA_API void chafa_symbol_map_ref (ChafaSymbolMap *symbol_map);
CHAFA_API void chafa_symbol_map_unref (ChafaSymbolMap *symbol_map);
CHAFA_API void chafa_symbol_map_add_by_tags (ChafaSymbolMap *symbol_map, ChafaSymbolTags tags);
CHAFA_API void chafa_symbol_map_remove_by_tags (ChafaSymbolMap *symbol_map, ChafaSymbolTags tags);
CHAFA_API void chafa_symbol_map_add_by_range (ChafaSymbolMap *symbol_map, gunichar first, gunichar last);
CHAFA_API void chafa_symbol_map_remove_by_range (ChafaSymbolMap *symbol_map, gunichar first, gunichar last);
/* Selector grammar as in the reference (chafa/chafa-symbol-map.c:924-1043).
 * On a syntax error returns FALSE and sets *error (domain "g-option-context-error-quark",
 * code 1 = BAD_VALUE); the map is left unchanged. */
CHAFA_API gboolean chafa_symbol_map_apply_selectors (ChafaSymbolMap *symbol_map,
                                                     const gchar *selectors, GError **error);
CHAFA_API gboolean chafa_symbol_map_get_allow_builtin_glyphs (ChafaSymbolMap *symbol_map);
CHAFA_API void chafa_symbol_map_set_allow_builtin_glyphs (ChafaSymbolMap *symbol_map,
                                                          gboolean use_builtin_glyphs);
CHAFA_API void chafa_symbol_map_add_glyph (ChafaSymbolMap *symbol_map, gunichar code_point,
                                           ChafaPixelType pixel_format, gpointer pixels,
                                           gint width, gint height, gint rowstride);
CHAFA_API gboolean chafa_symbol_map_get_glyph (ChafaSymbolMap *symbol_map, gunichar code_point,
                                               ChafaPixelType pixel_format, gpointer *pixels_out,
                                               gint *width_out, gint *height_out, gint *rowstride_out);

/* ------------------------------------------------------------------------ */
/* Canvas config (reference: chafa/chafa-canvas-config.h:33-141)              */

typedef struct ChafaCanvasConfig ChafaCanvasConfig;

CHAFA_API ChafaCanvasConfig *chafa_canvas_config_new (void);
CHAFA_API ChafaCanvasConfig *chafa_canvas_config_copy (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_ref (ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_unref (ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_get_geometry (const ChafaCanvasConfig *config, gint *width_out, gint *height_out);
CHAFA_API void chafa_canvas_config_set_geometry (ChafaCanvasConfig *config, gint width, gint height);
CHAFA_API void chafa_canvas_config_get_cell_geometry (const ChafaCanvasConfig *config, gint *cell_width_out, gint *cell_height_out);
CHAFA_API void chafa_canvas_config_set_cell_geometry (ChafaCanvasConfig *config, gint cell_width, gint cell_height);
CHAFA_API ChafaCanvasMode chafa_canvas_config_get_canvas_mode (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_canvas_mode (ChafaCanvasConfig *config, ChafaCanvasMode mode);
CHAFA_API ChafaColorExtractor chafa_canvas_config_get_color_extractor (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_color_extractor (ChafaCanvasConfig *config, ChafaColorExtractor color_extractor);
CHAFA_API ChafaColorSpace chafa_canvas_config_get_color_space (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_color_space (ChafaCanvasConfig *config, ChafaColorSpace color_space);
CHAFA_API const ChafaSymbolMap *chafa_canvas_config_peek_symbol_map (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_symbol_map (ChafaCanvasConfig *config, const ChafaSymbolMap *symbol_map);
CHAFA_API const ChafaSymbolMap *chafa_canvas_config_peek_fill_symbol_map (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_fill_symbol_map (ChafaCanvasConfig *config, const ChafaSymbolMap *fill_symbol_map);
CHAFA_API gfloat chafa_canvas_config_get_transparency_threshold (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_transparency_threshold (ChafaCanvasConfig *config, gfloat alpha_threshold);
CHAFA_API guint32 chafa_canvas_config_get_fg_color (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_fg_color (ChafaCanvasConfig *config, guint32 fg_color_packed_rgb);
CHAFA_API guint32 chafa_canvas_config_get_bg_color (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_bg_color (ChafaCanvasConfig *config, guint32 bg_color_packed_rgb);
CHAFA_API gfloat chafa_canvas_config_get_work_factor (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_work_factor (ChafaCanvasConfig *config, gfloat work_factor);
CHAFA_API gboolean chafa_canvas_config_get_preprocessing_enabled (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_preprocessing_enabled (ChafaCanvasConfig *config, gboolean preprocessing_enabled);
CHAFA_API ChafaDitherMode chafa_canvas_config_get_dither_mode (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_dither_mode (ChafaCanvasConfig *config, ChafaDitherMode dither_mode);
CHAFA_API void chafa_canvas_config_get_dither_grain_size (const ChafaCanvasConfig *config, gint *width_out, gint *height_out);
CHAFA_API void chafa_canvas_config_set_dither_grain_size (ChafaCanvasConfig *config, gint width, gint height);
CHAFA_API gfloat chafa_canvas_config_get_dither_intensity (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_dither_intensity (ChafaCanvasConfig *config, gfloat intensity);
CHAFA_API ChafaPixelMode chafa_canvas_config_get_pixel_mode (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_pixel_mode (ChafaCanvasConfig *config, ChafaPixelMode pixel_mode);
CHAFA_API ChafaOptimizations chafa_canvas_config_get_optimizations (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_optimizations (ChafaCanvasConfig *config, ChafaOptimizations optimizations);
CHAFA_API gboolean chafa_canvas_config_get_fg_only_enabled (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_fg_only_enabled (ChafaCanvasConfig *config, gboolean fg_only_enabled);
CHAFA_API ChafaPassthrough chafa_canvas_config_get_passthrough (const ChafaCanvasConfig *config);
CHAFA_API void chafa_canvas_config_set_passthrough (ChafaCanvasConfig *config, ChafaPassthrough passthrough);

/* ------------------------------------------------------------------------ */
/* Terminal info: the emit half only (reference: chafa/chafa-term-info.h)      */

#define CHAFA_TERM_SEQ_LENGTH_MAX 96
#define CHAFA_TERM_SEQ_ARGS_MAX 24

typedef enum
{
#include "chafa-term-seq-enum.inc"
    CHAFA_TERM_SEQ_MAX
}
ChafaTermSeq;

typedef enum
{
    CHAFA_TERM_INFO_ERROR_SEQ_TOO_LONG,
    CHAFA_TERM_INFO_ERROR_BAD_ESCAPE,
    CHAFA_TERM_INFO_ERROR_BAD_ARGUMENTS
}
ChafaTermInfoError;

typedef struct ChafaTermInfo ChafaTermInfo;
typedef struct ChafaTermDb ChafaTermDb;

CHAFA_API GQuark chafa_term_info_error_quark (void);
CHAFA_API ChafaTermInfo *chafa_term_info_new (void);
CHAFA_API ChafaTermInfo *chafa_term_info_copy (const ChafaTermInfo *term_info);
CHAFA_API void chafa_term_info_ref (ChafaTermInfo *term_info);
CHAFA_API void chafa_term_info_unref (ChafaTermInfo *term_info);
CHAFA_API const gchar *chafa_term_info_get_seq (ChafaTermInfo *term_info, ChafaTermSeq seq);
CHAFA_API gint chafa_term_info_set_seq (ChafaTermInfo *term_info, ChafaTermSeq seq,
                                        const gchar *str, GError **error);
CHAFA_API gboolean chafa_term_info_have_seq (const ChafaTermInfo *term_info, ChafaTermSeq seq);
CHAFA_API void chafa_term_info_supplement (ChafaTermInfo *term_info, ChafaTermInfo *source);
/* Generic emitter: formats @seq with unsigned arguments into a new malloc'd string
 * (chafa/chafa-term-info.c:1080-1160). Returns NULL if the sequence is not defined. */
CHAFA_API gchar *chafa_term_info_emit_seq (ChafaTermInfo *term_info, ChafaTermSeq seq, ...);
#include "chafa-term-seq-emit.inc"

/* Only the part of the terminal database the print path needs: the built-in
 * fallback description (chafa/chafa-term-db.c:569-582, 1316-1327).  Environment
 * probing is out of scope. */
CHAFA_API ChafaTermDb *chafa_term_db_get_default (void);
CHAFA_API ChafaTermInfo *chafa_term_db_get_fallback_info (ChafaTermDb *term_db);

/* ------------------------------------------------------------------------ */
/* Frame / image / placement wrappers (reference: chafa/chafa-frame.h,          */
/* chafa-image.h, chafa-placement.h)                                           */

typedef struct ChafaFrame ChafaFrame;
typedef struct ChafaImage ChafaImage;
typedef struct ChafaPlacement ChafaPlacement;

CHAFA_API ChafaFrame *chafa_frame_new (gconstpointer data, ChafaPixelType pixel_type,
                                       gint width, gint height, gint rowstride);
CHAFA_API ChafaFrame *chafa_frame_new_steal (gpointer data, ChafaPixelType pixel_type,
                                             gint width, gint height, gint rowstride);
CHAFA_API ChafaFrame *chafa_frame_new_borrow (gpointer data, ChafaPixelType pixel_type,
                                              gint width, gint height, gint rowstride);
CHAFA_API void chafa_frame_ref (ChafaFrame *frame);
CHAFA_API void chafa_frame_unref (ChafaFrame *frame);

CHAFA_API ChafaImage *chafa_image_new (void);
CHAFA_API void chafa_image_ref (ChafaImage *image);
CHAFA_API void chafa_image_unref (ChafaImage *image);
CHAFA_API void chafa_image_set_frame (ChafaImage *image, ChafaFrame *frame);

CHAFA_API ChafaPlacement *chafa_placement_new (ChafaImage *image, gint id);
CHAFA_API void chafa_placement_ref (ChafaPlacement *placement);
CHAFA_API void chafa_placement_unref (ChafaPlacement *placement);
CHAFA_API ChafaTuck chafa_placement_get_tuck (ChafaPlacement *placement);
CHAFA_API void chafa_placement_set_tuck (ChafaPlacement *placement, ChafaTuck tuck);
CHAFA_API ChafaAlign chafa_placement_get_halign (ChafaPlacement *placement);
CHAFA_API void chafa_placement_set_halign (ChafaPlacement *placement, ChafaAlign align);
CHAFA_API ChafaAlign chafa_placement_get_valign (ChafaPlacement *placement);
CHAFA_API void chafa_placement_set_valign (ChafaPlacement *placement, ChafaAlign align);

/* ------------------------------------------------------------------------ */
/* Canvas (reference: chafa/chafa-canvas.h:33-83)                             */

typedef struct ChafaCanvas ChafaCanvas;

/* chafa/chafa-canvas.c:513-631.  NULL config = defaults.  Returns NULL when the
 * config has a non-positive size -- or when no CUDA device / kernel image is
 * available (there is no CPU path to fall back to; CHAFA_B200_VERBOSE=1 prints why). */
CHAFA_API ChafaCanvas *chafa_canvas_new (const ChafaCanvasConfig *config);
CHAFA_API ChafaCanvas *chafa_canvas_new_similar (ChafaCanvas *orig);
CHAFA_API void chafa_canvas_ref (ChafaCanvas *canvas);
CHAFA_API void chafa_canvas_unref (ChafaCanvas *canvas);
CHAFA_API const ChafaCanvasConfig *chafa_canvas_peek_config (ChafaCanvas *canvas);
CHAFA_API void chafa_canvas_set_placement (ChafaCanvas *canvas, ChafaPlacement *placement);
/* chafa/chafa-canvas.c:786-805: the hot path.  @src_pixels is HOST memory, borrowed
 * for the duration of the call. */
CHAFA_API void chafa_canvas_draw_all_pixels (ChafaCanvas *canvas, ChafaPixelType src_pixel_type,
                                             const guint8 *src_pixels,
                                             gint src_width, gint src_height, gint src_rowstride);
/* chafa/chafa-canvas.c:866-922.  @term_info may be NULL (fallback sequences).  The
 * returned GString is owned by the caller (g_string_free / chafa_b200_string_free). */
CHAFA_API GString *chafa_canvas_print (ChafaCanvas *canvas, ChafaTermInfo *term_info);
CHAFA_API void chafa_canvas_print_rows (ChafaCanvas *canvas, ChafaTermInfo *term_info,
                                        GString ***array_out, gint *array_len_out);
CHAFA_API gchar **chafa_canvas_print_rows_strv (ChafaCanvas *canvas, ChafaTermInfo *term_info);
CHAFA_API gunichar chafa_canvas_get_char_at (ChafaCanvas *canvas, gint x, gint y);
CHAFA_API gint chafa_canvas_set_char_at (ChafaCanvas *canvas, gint x, gint y, gunichar c);
CHAFA_API void chafa_canvas_get_colors_at (ChafaCanvas *canvas, gint x, gint y, gint *fg_out, gint *bg_out);
CHAFA_API void chafa_canvas_set_colors_at (ChafaCanvas *canvas, gint x, gint y, gint fg, gint bg);
CHAFA_API void chafa_canvas_get_raw_colors_at (ChafaCanvas *canvas, gint x, gint y, gint *fg_out, gint *bg_out);
CHAFA_API void chafa_canvas_set_raw_colors_at (ChafaCanvas *canvas, gint x, gint y, gint fg, gint bg);
CHAFA_API void chafa_canvas_set_contents_rgba8 (ChafaCanvas *canvas, const guint8 *src_pixels,
                                                gint src_width, gint src_height, gint src_rowstride);
CHAFA_API GString *chafa_canvas_build_ansi (ChafaCanvas *canvas);

/* ------------------------------------------------------------------------ */
/* Utilities (reference: chafa/chafa-util.h)                                   */

CHAFA_API void chafa_calc_canvas_geometry (gint src_width, gint src_height,
                                           gint *dest_width_inout, gint *dest_height_inout,
                                           gfloat font_ratio, gboolean zoom, gboolean stretch);
CHAFA_API void chafa_free_gstring_array (GString **gsa);

#ifdef __cplusplus
}
#endif

#endif /* CHAFA_B200_CHAFA_H */
