/* host_glyph.cpp -- user glyph import (placeholder: the scaler route for
 * premultiplied output without compositing is not built yet). */

#include "internal.h"

bool
glyph_import_device (ChafaPixelType pixel_format, const void *pixels, int width, int height, int rowstride,
                     bool wide, uint64_t *bitmaps_out)
{
    (void) pixel_format; (void) pixels; (void) width; (void) height; (void) rowstride; (void) wide;
    (void) bitmaps_out;
    b200_set_error ("chafa_symbol_map_add_glyph: glyph import is not built yet");
    fprintf (stderr, "chafa-b200: chafa_symbol_map_add_glyph: glyph import is not built yet\n");
    return false;
}
