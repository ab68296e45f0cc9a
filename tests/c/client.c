/* A plain-C caller of the drop-in: configure, draw a small gradient, print it whole, then by rows
 * (chafa_canvas_print_rows_strv, chafa-canvas.c:1000-1032) and emit a few typed control sequences
 * (chafa-term-seq-def.h).  tests/test_c_client.py links it against include/chafa.h +
 * libchafa_b200.so on CPU and runs it on the GPU box; the same calls go to the compiled reference
 * through ctypes and the outputs are compared byte for byte.
 *
 * Output format: sections "== name <n bytes>\n" followed by the bytes. */

#include <chafa.h>
#include <stdio.h>
#include <string.h>

#define W 64
#define H 48

static void
section (const char *name, const char *data, size_t len)
{
    printf ("== %s %zu\n", name, len);
    fwrite (data, 1, len, stdout);
    fputc ('\n', stdout);
}

int
main (void)
{
    static guint8 pixels [W * H * 4];
    ChafaSymbolMap *symbol_map;
    ChafaCanvasConfig *config;
    ChafaCanvas *canvas;
    ChafaTermInfo *term_info;
    GString *gs;
    gchar **rows;
    gchar buf [CHAFA_TERM_SEQ_LENGTH_MAX + 1], *end;
    size_t total = 0;
    int x, y, i;

    for (y = 0; y < H; y++)
        for (x = 0; x < W; x++)
        {
            guint8 *p = pixels + (y * W + x) * 4;
            p [0] = (guint8) (x * 4);
            p [1] = (guint8) (y * 5);
            p [2] = (guint8) ((x ^ y) * 3);
            p [3] = (x + y) % 11 == 0 ? 0 : 255;
        }

    symbol_map = chafa_symbol_map_new ();
    chafa_symbol_map_add_by_tags (symbol_map, CHAFA_SYMBOL_TAG_ALL);

    config = chafa_canvas_config_new ();
    chafa_canvas_config_set_geometry (config, 16, 6);
    chafa_canvas_config_set_symbol_map (config, symbol_map);
    chafa_canvas_config_set_canvas_mode (config, CHAFA_CANVAS_MODE_INDEXED_256);

    canvas = chafa_canvas_new (config);
    if (!canvas)
    {
        fprintf (stderr, "chafa_canvas_new failed\n");
        return 2;
    }
    chafa_canvas_draw_all_pixels (canvas, CHAFA_PIXEL_RGBA8_UNASSOCIATED, pixels, W, H, W * 4);

    gs = chafa_canvas_print (canvas, NULL);
    section ("print", gs->str, gs->len);
    g_string_free (gs, TRUE);

    rows = chafa_canvas_print_rows_strv (canvas, NULL);
    for (i = 0; rows [i]; i++)
    {
        char name [32];
        snprintf (name, sizeof (name), "row%d", i);
        section (name, rows [i], strlen (rows [i]));
        total += strlen (rows [i]);
    }
    printf ("== n_rows %d total %zu\n", i, total);
    g_strfreev (rows);

    term_info = chafa_term_db_get_fallback_info (chafa_term_db_get_default ());
    end = chafa_term_info_emit_cursor_to_pos (term_info, buf, 12, 34);
    section ("cursor_to_pos", buf, (size_t) (end - buf));
    end = chafa_term_info_emit_set_color_fgbg_direct (term_info, buf, 1, 2, 3, 250, 251, 252);
    section ("set_color_fgbg_direct", buf, (size_t) (end - buf));
    end = chafa_term_info_emit_reset_attributes (term_info, buf);
    section ("reset_attributes", buf, (size_t) (end - buf));
    end = chafa_term_info_emit_begin_sixels (term_info, buf, 0, 1, 0);
    section ("begin_sixels", buf, (size_t) (end - buf));
    chafa_term_info_unref (term_info);

    chafa_canvas_unref (canvas);
    chafa_canvas_config_unref (config);
    chafa_symbol_map_unref (symbol_map);
    return 0;
}
