/* ref_probe.c -- OUR code, compiled INTO oracle/_ref/libchafa_ref.so next to the
 * reference's own files.  TEST INFRASTRUCTURE ONLY.
 *
 * Read-only accessors into the reference's internal state, so that tests and the
 * table generator (tools/gen_tables.py) can see what the reference actually built:
 * the built-in glyph set, the compiled symbol-table order, per-cell results before
 * printing, the fixed palette in both colour spaces, the scaler LUTs, the fallback
 * terminal sequences.  Nothing here computes anything on its own. */

#include "config.h"
#include "chafa.h"
#include "internal/chafa-private.h"
#include "internal/chafa-canvas-internal.h"
#include "internal/chafa-pixops.h"
#include "internal/smolscale/smolscale.h"

int
ref_probe_version (void)
{
    return 2;
}

/* ---- built-in glyph set (chafa/internal/chafa-symbols.c:572-666) ---- */

int
ref_probe_builtin_symbol (int wide, int index, guint32 *c_out, guint32 *tags_out,
                          guint64 *bitmap0_out, guint64 *bitmap1_out)
{
    chafa_init ();

    if (!wide)
    {
        const ChafaSymbol *sym = &chafa_symbols [index];
        if (index < 0 || index >= CHAFA_N_SYMBOLS_MAX || sym->c == 0)
            return 0;
        *c_out = sym->c;
        *tags_out = sym->sc;
        *bitmap0_out = sym->bitmap;
        *bitmap1_out = 0;
        return 1;
    }
    else
    {
        const ChafaSymbol2 *sym = &chafa_symbols2 [index];
        if (index < 0 || index >= CHAFA_N_SYMBOLS_MAX || sym->sym [0].c == 0)
            return 0;
        *c_out = sym->sym [0].c;
        *tags_out = sym->sym [0].sc;
        *bitmap0_out = sym->sym [0].bitmap;
        *bitmap1_out = sym->sym [1].bitmap;
        return 1;
    }
}

guint32
ref_probe_tags_for_char (guint32 c)
{
    chafa_init ();
    return chafa_get_tags_for_char (c);
}

/* ---- compiled symbol table of a map (chafa/chafa-symbol-map.c:384-470) ---- */

int
ref_probe_symbol_map_compiled (ChafaSymbolMap *map, int wide, int max_out,
                               guint32 *c_out, guint64 *bitmap0_out, guint64 *bitmap1_out)
{
    int i, n;

    chafa_init ();
    /* Prepare a private copy so the caller's map is not mutated */
    map = chafa_symbol_map_copy (map);
    chafa_symbol_map_prepare (map);

    n = wide ? map->n_symbols2 : map->n_symbols;
    for (i = 0; i < n && i < max_out; i++)
    {
        if (!wide)
        {
            c_out [i] = map->symbols [i].c;
            bitmap0_out [i] = map->symbols [i].bitmap;
            bitmap1_out [i] = 0;
        }
        else
        {
            c_out [i] = map->symbols2 [i].sym [0].c;
            bitmap0_out [i] = map->symbols2 [i].sym [0].bitmap;
            bitmap1_out [i] = map->symbols2 [i].sym [1].bitmap;
        }
    }

    chafa_symbol_map_unref (map);
    return n;
}

/* ---- canvas internals (chafa/internal/chafa-canvas-internal.h:30-88) ---- */

const void *
ref_probe_canvas_cells (ChafaCanvas *canvas)
{
    return canvas->cells;
}

void
ref_probe_canvas_info (ChafaCanvas *canvas, gint *out)
{
    out [0] = canvas->width_pixels;
    out [1] = canvas->height_pixels;
    out [2] = canvas->work_factor_int;
    out [3] = canvas->blank_char;
    out [4] = canvas->solid_char;
    out [5] = canvas->consider_inverted;
    out [6] = canvas->extract_colors;
    out [7] = canvas->use_quantized_error;
    out [8] = canvas->config.symbol_map.n_symbols;
    out [9] = canvas->config.symbol_map.n_symbols2;
    out [10] = canvas->config.fill_symbol_map.n_symbols;
    out [11] = canvas->config.fill_symbol_map.n_symbols2;
}

/* Run only the pixel-preparation half of draw_all_pixels and hand back the
 * prepared pixmap (chafa/internal/chafa-symbol-renderer.c:959-976,
 * chafa/internal/chafa-pixops.c:677-791).  Caller frees with ref_probe_free. */
void *
ref_probe_prepare_pixels (ChafaCanvas *canvas, ChafaPixelType src_pixel_type,
                          const guint8 *src_pixels, gint src_width, gint src_height,
                          gint src_rowstride, gint halign, gint valign, gint tuck)
{
    ChafaPixel *pixels;

    pixels = g_try_new (ChafaPixel, (gsize) canvas->width_pixels * canvas->height_pixels);
    if (!pixels)
        return NULL;

    chafa_prepare_pixel_data_for_symbols (&canvas->fg_palette, &canvas->dither,
                                          canvas->config.color_space,
                                          canvas->config.preprocessing_enabled,
                                          canvas->work_factor_int,
                                          src_pixel_type, src_pixels,
                                          src_width, src_height, src_rowstride,
                                          pixels,
                                          canvas->width_pixels, canvas->height_pixels,
                                          canvas->config.cell_width, canvas->config.cell_height,
                                          (ChafaAlign) halign, (ChafaAlign) valign, (ChafaTuck) tuck);
    return pixels;
}

void
ref_probe_free (void *p)
{
    g_free (p);
}

/* ---- plain scaler call with Chafa's arguments (chafa-pixops.c:760-785) ---- */

int
ref_probe_scale (const void *src, int src_type, int sw, int sh, int sstride,
                 const guint8 *bg_rgba, void *dest, int dw, int dh,
                 int px, int py, int pw, int ph, int nearest)
{
    SmolScaleCtx *ctx;
    SmolFlags flags = SMOL_CLEAR_DEST;

    if (nearest)
        flags |= SMOL_INTERP_NEAREST;

    ctx = smol_scale_new_full (src, (SmolPixelType) src_type, sw, sh, sstride,
                               bg_rgba, SMOL_PIXEL_RGBA8_UNASSOCIATED,
                               NULL, SMOL_PIXEL_RGBA8_UNASSOCIATED, dw, dh, dw * 4,
                               px * SMOL_SUBPIXEL_MUL, py * SMOL_SUBPIXEL_MUL,
                               pw * SMOL_SUBPIXEL_MUL, ph * SMOL_SUBPIXEL_MUL,
                               SMOL_COMPOSITE_SRC_OVER_COLOR_SRC_ALPHA, SMOL_OPACITY_MAX,
                               flags, NULL, NULL);
    if (!ctx)
        return 0;
    smol_scale_batch_full (ctx, dest, 0, dh);
    smol_scale_destroy (ctx);
    return 1;
}

/* ---- fixed palette (chafa/internal/chafa-palette.c:131-238) ---- */

guint32
ref_probe_fixed_palette_color (int index, int color_space)
{
    ChafaPalette pal;
    const ChafaColor *col;

    chafa_palette_init (&pal, CHAFA_PALETTE_TYPE_FIXED_256);
    col = chafa_palette_get_color (&pal, (ChafaColorSpace) color_space, index);
    return (guint32) col->ch [0] | ((guint32) col->ch [1] << 8)
        | ((guint32) col->ch [2] << 16) | ((guint32) col->ch [3] << 24);
}

guint32
ref_probe_rgb_to_din99d (guint32 rgba)
{
    ChafaColor in, out;

    in.ch [0] = rgba & 0xff;
    in.ch [1] = (rgba >> 8) & 0xff;
    in.ch [2] = (rgba >> 16) & 0xff;
    in.ch [3] = (rgba >> 24) & 0xff;
    chafa_color_rgb_to_din99d (&in, &out);
    return (guint32) out.ch [0] | ((guint32) out.ch [1] << 8)
        | ((guint32) out.ch [2] << 16) | ((guint32) out.ch [3] << 24);
}

/* All 2^24 colours through chafa_color_rgb_to_din99d () (chafa-color.c:130-168); alpha byte
 * of the result cleared.  Entry index = b << 16 | g << 8 | r. */
void
ref_probe_din99d_table (guint32 *out)
{
    guint32 i;

    for (i = 0; i < (1u << 24); i++)
        out [i] = ref_probe_rgb_to_din99d (i | 0xff000000u) & 0x00ffffffu;
}

int
ref_probe_palette_lookup (ChafaCanvas *canvas, int which, guint32 rgba, int color_space)
{
    ChafaColor in;

    in.ch [0] = rgba & 0xff;
    in.ch [1] = (rgba >> 8) & 0xff;
    in.ch [2] = (rgba >> 16) & 0xff;
    in.ch [3] = (rgba >> 24) & 0xff;
    return chafa_palette_lookup_nearest (which ? &canvas->bg_palette : &canvas->fg_palette,
                                         (ChafaColorSpace) color_space, &in, NULL);
}

/* ---- dither textures (chafa/internal/chafa-dither.c:55-84) ---- */

int
ref_probe_dither_texture (ChafaCanvas *canvas, gint *out, int max_out)
{
    int n = 0, i;

    if (canvas->dither.mode == CHAFA_DITHER_MODE_ORDERED)
        n = 256;
    else if (canvas->dither.mode == CHAFA_DITHER_MODE_NOISE)
        n = 64 * 64 * 3;
    for (i = 0; i < n && i < max_out; i++)
        out [i] = canvas->dither.texture_data [i];
    return n;
}

/* ---- sixel internals: indexed image after quantisation ---- */

int
ref_probe_sixel_image (ChafaCanvas *canvas, gint *w_out, gint *h_out,
                       const guint8 **pixels_out, gint *n_colors_out, guint32 *pal_out)
{
    ChafaSixelRenderer *r = canvas->pixel_renderer;
    int i;

    if (canvas->config.pixel_mode != CHAFA_PIXEL_MODE_SIXELS || !r || !r->image)
        return 0;
    *w_out = r->image->width;
    *h_out = r->image->height;
    *pixels_out = r->image->pixels;
    *n_colors_out = chafa_palette_get_n_colors (&r->image->palette);
    for (i = 0; i < 256; i++)
    {
        const ChafaColor *col = chafa_palette_get_color (&r->image->palette, CHAFA_COLOR_SPACE_RGB, i);
        pal_out [i] = (guint32) col->ch [0] | ((guint32) col->ch [1] << 8)
            | ((guint32) col->ch [2] << 16) | ((guint32) col->ch [3] << 24);
    }
    return 1;
}
