/* host_sixel.cpp -- sixel pixel mode (placeholder until the sixel kernels land). */

#include "internal.h"
#include "device.h"

struct ChafaCanvas;

void
sixel_state_free (void *state)
{
    (void) state;
}

bool
sixel_draw (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int w, int h, int rowstride)
{
    (void) c; (void) d_src; (void) type; (void) w; (void) h; (void) rowstride;
    b200_set_error ("sixel mode not built yet");
    fprintf (stderr, "chafa-b200: sixel mode not built yet\n");
    return false;
}

GString *
sixel_print (ChafaCanvas *c, ChafaTermInfo *ti)
{
    (void) c; (void) ti;
    return nullptr;
}

extern "C" int
chafa_b200_canvas_get_sixel_image (ChafaCanvas *canvas, uint8_t *pixels_out, int *width_out, int *height_out,
                                   int *n_colors_out, uint32_t *palette_out)
{
    (void) canvas; (void) pixels_out; (void) width_out; (void) height_out; (void) n_colors_out; (void) palette_out;
    return 0;
}
