/* host_config.cpp -- ChafaCanvasConfig: option container, host only.
 *
 * Replaces chafa/chafa-canvas-config.c:45-935 of the reference: same defaults,
 * same argument checks (an out-of-range argument leaves the config unchanged). */

#include "internal.h"

ChafaCanvasConfig::ChafaCanvasConfig ()
{
    /* Defaults: block + border + space, no wide symbols; empty fill map
     * (chafa/chafa-canvas-config.c:45-76). */
    Selector s;
    s.is_range = false;
    s.first = s.last = 0;
    s.additive = true;
    s.tags = CHAFA_SYMBOL_TAG_BLOCK;
    symbol_map.selectors.push_back (s);
    s.tags = CHAFA_SYMBOL_TAG_BORDER;
    symbol_map.selectors.push_back (s);
    s.tags = CHAFA_SYMBOL_TAG_SPACE;
    symbol_map.selectors.push_back (s);
    s.additive = false;
    s.tags = CHAFA_SYMBOL_TAG_WIDE;
    symbol_map.selectors.push_back (s);
}

ChafaCanvasConfig::ChafaCanvasConfig (const ChafaCanvasConfig &o)
    : refs (1), width (o.width), height (o.height), cell_width (o.cell_width), cell_height (o.cell_height),
      canvas_mode (o.canvas_mode), color_space (o.color_space), dither_mode (o.dither_mode),
      color_extractor (o.color_extractor), pixel_mode (o.pixel_mode),
      dither_grain_width (o.dither_grain_width), dither_grain_height (o.dither_grain_height),
      dither_intensity (o.dither_intensity), fg_color_packed_rgb (o.fg_color_packed_rgb),
      bg_color_packed_rgb (o.bg_color_packed_rgb), alpha_threshold (o.alpha_threshold),
      work_factor (o.work_factor), symbol_map (o.symbol_map), fill_symbol_map (o.fill_symbol_map),
      preprocessing_enabled (o.preprocessing_enabled), fg_only_enabled (o.fg_only_enabled),
      optimizations (o.optimizations), passthrough (o.passthrough)
{
}

#define CHECK(c) do { RETURN_IF_FAIL ((c) != NULL); RETURN_IF_FAIL ((c)->refs.load () > 0); } while (0)
#define CHECK_VAL(c, v) do { RETURN_VAL_IF_FAIL ((c) != NULL, (v)); RETURN_VAL_IF_FAIL ((c)->refs.load () > 0, (v)); } while (0)

extern "C" {

ChafaCanvasConfig *
chafa_canvas_config_new (void)
{
    return new ChafaCanvasConfig ();
}

ChafaCanvasConfig *
chafa_canvas_config_copy (const ChafaCanvasConfig *config)
{
    RETURN_VAL_IF_FAIL (config != NULL, NULL);
    return new ChafaCanvasConfig (*config);
}

void
chafa_canvas_config_ref (ChafaCanvasConfig *config)
{
    CHECK (config);
    config->refs++;
}

void
chafa_canvas_config_unref (ChafaCanvasConfig *config)
{
    CHECK (config);
    if (--config->refs == 0)
        delete config;
}

void
chafa_canvas_config_get_geometry (const ChafaCanvasConfig *config, gint *width_out, gint *height_out)
{
    CHECK (config);
    if (width_out) *width_out = config->width;
    if (height_out) *height_out = config->height;
}

void
chafa_canvas_config_set_geometry (ChafaCanvasConfig *config, gint width, gint height)
{
    CHECK (config);
    RETURN_IF_FAIL (width > 0);
    RETURN_IF_FAIL (height > 0);
    config->width = width;
    config->height = height;
}

void
chafa_canvas_config_get_cell_geometry (const ChafaCanvasConfig *config, gint *cell_width_out, gint *cell_height_out)
{
    CHECK (config);
    if (cell_width_out) *cell_width_out = config->cell_width;
    if (cell_height_out) *cell_height_out = config->cell_height;
}

void
chafa_canvas_config_set_cell_geometry (ChafaCanvasConfig *config, gint cell_width, gint cell_height)
{
    CHECK (config);
    RETURN_IF_FAIL (cell_width > 0);
    RETURN_IF_FAIL (cell_height > 0);
    config->cell_width = cell_width;
    config->cell_height = cell_height;
}

ChafaCanvasMode
chafa_canvas_config_get_canvas_mode (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_CANVAS_MODE_TRUECOLOR);
    return config->canvas_mode;
}

void
chafa_canvas_config_set_canvas_mode (ChafaCanvasConfig *config, ChafaCanvasMode mode)
{
    CHECK (config);
    RETURN_IF_FAIL ((unsigned) mode < CHAFA_CANVAS_MODE_MAX);
    config->canvas_mode = mode;
}

ChafaColorExtractor
chafa_canvas_config_get_color_extractor (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_COLOR_EXTRACTOR_AVERAGE);
    return config->color_extractor;
}

void
chafa_canvas_config_set_color_extractor (ChafaCanvasConfig *config, ChafaColorExtractor color_extractor)
{
    CHECK (config);
    RETURN_IF_FAIL ((unsigned) color_extractor < CHAFA_COLOR_EXTRACTOR_MAX);
    config->color_extractor = color_extractor;
}

ChafaColorSpace
chafa_canvas_config_get_color_space (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_COLOR_SPACE_RGB);
    return config->color_space;
}

void
chafa_canvas_config_set_color_space (ChafaCanvasConfig *config, ChafaColorSpace color_space)
{
    CHECK (config);
    RETURN_IF_FAIL ((unsigned) color_space < CHAFA_COLOR_SPACE_MAX);
    config->color_space = color_space;
}

const ChafaSymbolMap *
chafa_canvas_config_peek_symbol_map (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, NULL);
    return &config->symbol_map;
}

void
chafa_canvas_config_set_symbol_map (ChafaCanvasConfig *config, const ChafaSymbolMap *symbol_map)
{
    CHECK (config);
    RETURN_IF_FAIL (symbol_map != NULL);
    config->symbol_map = *symbol_map;
}

const ChafaSymbolMap *
chafa_canvas_config_peek_fill_symbol_map (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, NULL);
    return &config->fill_symbol_map;
}

void
chafa_canvas_config_set_fill_symbol_map (ChafaCanvasConfig *config, const ChafaSymbolMap *fill_symbol_map)
{
    CHECK (config);
    RETURN_IF_FAIL (fill_symbol_map != NULL);
    config->fill_symbol_map = *fill_symbol_map;
}

gfloat
chafa_canvas_config_get_transparency_threshold (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 0.0f);
    return (gfloat) (1.0 - (config->alpha_threshold / 256.0));
}

void
chafa_canvas_config_set_transparency_threshold (ChafaCanvasConfig *config, gfloat alpha_threshold)
{
    CHECK (config);
    RETURN_IF_FAIL (alpha_threshold >= 0.0f);
    RETURN_IF_FAIL (alpha_threshold <= 1.0f);
    /* Stored inverted, as an opacity threshold, truncated to int */
    config->alpha_threshold = (int) (256.0f * (1.0f - alpha_threshold));
}

guint32
chafa_canvas_config_get_fg_color (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 0);
    return config->fg_color_packed_rgb;
}

void
chafa_canvas_config_set_fg_color (ChafaCanvasConfig *config, guint32 fg_color_packed_rgb)
{
    CHECK (config);
    config->fg_color_packed_rgb = fg_color_packed_rgb;
}

guint32
chafa_canvas_config_get_bg_color (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 0);
    return config->bg_color_packed_rgb;
}

void
chafa_canvas_config_set_bg_color (ChafaCanvasConfig *config, guint32 bg_color_packed_rgb)
{
    CHECK (config);
    config->bg_color_packed_rgb = bg_color_packed_rgb;
}

gfloat
chafa_canvas_config_get_work_factor (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 1.0f);
    return config->work_factor;
}

void
chafa_canvas_config_set_work_factor (ChafaCanvasConfig *config, gfloat work_factor)
{
    CHECK (config);
    RETURN_IF_FAIL (work_factor >= 0.0f && work_factor <= 1.0f);
    config->work_factor = work_factor;
}

gboolean
chafa_canvas_config_get_preprocessing_enabled (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 0);
    return config->preprocessing_enabled ? 1 : 0;
}

void
chafa_canvas_config_set_preprocessing_enabled (ChafaCanvasConfig *config, gboolean preprocessing_enabled)
{
    CHECK (config);
    config->preprocessing_enabled = !!preprocessing_enabled;
}

ChafaDitherMode
chafa_canvas_config_get_dither_mode (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_DITHER_MODE_NONE);
    return config->dither_mode;
}

void
chafa_canvas_config_set_dither_mode (ChafaCanvasConfig *config, ChafaDitherMode dither_mode)
{
    CHECK (config);
    RETURN_IF_FAIL ((unsigned) dither_mode < CHAFA_DITHER_MODE_MAX);
    config->dither_mode = dither_mode;
}

void
chafa_canvas_config_get_dither_grain_size (const ChafaCanvasConfig *config, gint *width_out, gint *height_out)
{
    CHECK (config);
    if (width_out) *width_out = config->dither_grain_width;
    if (height_out) *height_out = config->dither_grain_height;
}

void
chafa_canvas_config_set_dither_grain_size (ChafaCanvasConfig *config, gint width, gint height)
{
    CHECK (config);
    RETURN_IF_FAIL (width == 1 || width == 2 || width == 4 || width == 8);
    RETURN_IF_FAIL (height == 1 || height == 2 || height == 4 || height == 8);
    config->dither_grain_width = width;
    config->dither_grain_height = height;
}

gfloat
chafa_canvas_config_get_dither_intensity (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 1.0f);
    return config->dither_intensity;
}

void
chafa_canvas_config_set_dither_intensity (ChafaCanvasConfig *config, gfloat intensity)
{
    CHECK (config);
    RETURN_IF_FAIL (intensity >= 0.0f);
    config->dither_intensity = intensity;
}

ChafaPixelMode
chafa_canvas_config_get_pixel_mode (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_PIXEL_MODE_SYMBOLS);
    return config->pixel_mode;
}

void
chafa_canvas_config_set_pixel_mode (ChafaCanvasConfig *config, ChafaPixelMode pixel_mode)
{
    CHECK (config);
    RETURN_IF_FAIL ((unsigned) pixel_mode < CHAFA_PIXEL_MODE_MAX);
    config->pixel_mode = pixel_mode;
}

ChafaOptimizations
chafa_canvas_config_get_optimizations (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_OPTIMIZATION_NONE);
    return (ChafaOptimizations) config->optimizations;
}

void
chafa_canvas_config_set_optimizations (ChafaCanvasConfig *config, ChafaOptimizations optimizations)
{
    CHECK (config);
    config->optimizations = (int) optimizations;
}

gboolean
chafa_canvas_config_get_fg_only_enabled (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, 0);
    return config->fg_only_enabled ? 1 : 0;
}

void
chafa_canvas_config_set_fg_only_enabled (ChafaCanvasConfig *config, gboolean fg_only_enabled)
{
    CHECK (config);
    config->fg_only_enabled = !!fg_only_enabled;
}

ChafaPassthrough
chafa_canvas_config_get_passthrough (const ChafaCanvasConfig *config)
{
    CHECK_VAL (config, CHAFA_PASSTHROUGH_NONE);
    return config->passthrough;
}

void
chafa_canvas_config_set_passthrough (ChafaCanvasConfig *config, ChafaPassthrough passthrough)
{
    CHECK (config);
    config->passthrough = passthrough;
}

} /* extern "C" */
