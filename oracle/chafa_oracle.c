/* chafa_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the symbols-mode
 * hot path of hpjansson/chafa 1.19: cell matching on a prepared pixmap, row resolve and
 * ANSI printing.  Nothing in the product links or calls this; only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline leg may load it.
 *
 * PINNED: tests/test_oracle.py checks this file against (a) the golden vectors in
 * the .npz files in tests/golden/, which were produced by the reference's own C files compiled as
 * oracle/_ref/libchafa_ref.so (generator: tests/golden/gen_golden.py), and (b) live against
 * that library whenever it is present.
 *
 * Coverage: RGB colour space; canvas modes TRUECOLOR, INDEXED_256, INDEXED_240,
 * INDEXED_16; average colour extractor; narrow and wide symbols; work factors below
 * (Hamming-prefilter path) and at or above 0.8 (exhaustive path); no fill symbols, no
 * foreground-only mode.  The scaler, dithering and sixel mode are checked against
 * oracle/_ref directly and through goldens, not restated here.
 *
 * Arithmetic follows the reference's AVX2 build, which is what oracle/_ref is on x86-64
 * hosts: mean colours divide by a rounded reciprocal multiply,
 * the cell error sums all four channels. */

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MODE_TRUECOLOR 0
#define MODE_INDEXED_256 1
#define MODE_INDEXED_240 2
#define MODE_INDEXED_16 3

#define ERROR_MAX (INT32_MAX / 8)
#define PEN_TRANSPARENT 256
#define PEN_FG 257
#define OPT_REUSE_ATTRIBUTES 1

typedef struct
{
    uint32_t c, fg, bg;             /* chafa/internal/chafa-canvas-internal.h:30-37 */
}
OracleCell;

typedef struct
{
    int32_t width, height;          /* cells */
    int32_t mode;
    int32_t work_factor_int;        /* work_factor * 10 + 0.5 (chafa-canvas.c:556) */
    int32_t alpha_threshold;
    int32_t optimizations;
    uint32_t blank_char, solid_char;
    int32_t n_syms;
    const uint64_t *bitmaps;        /* compiled table order (chafa-symbol-map.c:384-425) */
    const uint32_t *chars;
    int32_t n_syms2;
    const uint64_t *bitmaps2;       /* left, right per wide symbol */
    const uint32_t *chars2;
}
OracleCanvas;

typedef struct { uint8_t ch [4]; } Color;

/* ---- fixed palette (chafa/internal/chafa-palette.c:131-238) --------------------- */

static Color fixed_palette [259];
static uint8_t cube_index [256];
static int palette_ready;

static void
init_palette (void)
{
    static const uint32_t first16 [16] =
    {
        0x000000, 0x800000, 0x007000, 0x707000, 0x000070, 0x700070, 0x007070, 0xc0c0c0,
        0x404040, 0xff0000, 0x00ff00, 0xffff00, 0x0000ff, 0xff00ff, 0x00ffff, 0xffffff
    };
    static const int levels [6] = { 0x00, 0x5f, 0x87, 0xaf, 0xd7, 0xff };
    int i, v, level;

    if (palette_ready)
        return;
    for (i = 0; i < 16; i++)
    {
        fixed_palette [i].ch [0] = first16 [i] >> 16;
        fixed_palette [i].ch [1] = (first16 [i] >> 8) & 0xff;
        fixed_palette [i].ch [2] = first16 [i] & 0xff;
    }
    for (i = 0; i < 216; i++)
    {
        fixed_palette [16 + i].ch [0] = levels [i / 36];
        fixed_palette [16 + i].ch [1] = levels [(i / 6) % 6];
        fixed_palette [16 + i].ch [2] = levels [i % 6];
    }
    for (i = 0; i < 24; i++)
        fixed_palette [232 + i].ch [0] = fixed_palette [232 + i].ch [1] = fixed_palette [232 + i].ch [2] = 8 + 10 * i;
    memset (fixed_palette [256].ch, 0x80, 3);   /* transparent */
    memset (fixed_palette [257].ch, 0xff, 3);   /* default foreground */
    memset (fixed_palette [258].ch, 0x00, 3);   /* default background */
    for (i = 0; i < 259; i++)
        fixed_palette [i].ch [3] = 0xff;
    fixed_palette [256].ch [3] = 0;

    for (v = 0, level = 0; v < 256; v++)
    {
        while (level < 5 && v >= (levels [level] + levels [level + 1]) / 2)
            level++;
        cube_index [v] = level;
    }
    palette_ready = 1;
}

/* chafa/internal/chafa-color.h:145-148: three channels */
static int
diff3 (const Color *a, const Color *b)
{
    int d0 = (int) b->ch [0] - a->ch [0], d1 = (int) b->ch [1] - a->ch [1], d2 = (int) b->ch [2] - a->ch [2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

typedef struct { int index [2]; int error [2]; } Candidates2;

/* chafa-palette.c:96-125 */
static void
cand2_update (Candidates2 *c, int index, int error)
{
    if (error < c->error [0])
    {
        c->index [1] = c->index [0];
        c->index [0] = index;
        c->error [1] = c->error [0];
        c->error [0] = error;
    }
    else if (error < c->error [1])
    {
        c->index [1] = index;
        c->error [1] = error;
    }
}

static int
cand2_try (Candidates2 *c, const Color *color, int index)
{
    int error = diff3 (color, &fixed_palette [index]);
    cand2_update (c, index, error);
    return error;
}

/* chafa-palette.c:253-305: cube entry, then the gray-ramp walk with its quirks (pen 244
 * recorded with pen 245's error; one entry read past either end of the ramp) */
static void
pick_cube_and_grays (Candidates2 *c, const Color *color)
{
    int i, step, error, last_error;

    cand2_try (c, color, 16 + cube_index [color->ch [0]] * 36 + cube_index [color->ch [1]] * 6
               + cube_index [color->ch [2]]);

    i = 232 + 12;
    last_error = cand2_try (c, color, i);
    error = diff3 (color, &fixed_palette [i + 1]);
    if (error < last_error)
    {
        cand2_update (c, i, error);
        last_error = error;
        step = 1;
        i++;
    }
    else
        step = -1;

    do
    {
        i += step;
        error = diff3 (color, &fixed_palette [i]);
        if (error > last_error)
            break;
        cand2_update (c, i, error);
        last_error = error;
    }
    while (i >= 232 && i <= 255);
}

/* chafa-palette.c:955-1037, fixed palettes, RGB */
static int
lookup_nearest (int mode, int alpha_threshold, const Color *color)
{
    Candidates2 c = { { -1, -1 }, { INT32_MAX, INT32_MAX } };
    int i;

    if (color->ch [3] < alpha_threshold)
        return PEN_TRANSPARENT;

    if (mode == MODE_INDEXED_256 || mode == MODE_INDEXED_240)
    {
        pick_cube_and_grays (&c, color);
        if (mode == MODE_INDEXED_256)
            for (i = 0; i < 16; i++)
                cand2_try (&c, color, i);
    }
    else
    {
        for (i = 0; i < 16; i++)
            cand2_try (&c, color, i);
    }
    return c.index [0];
}

/* ---- 8x8 work cell (chafa/internal/chafa-work-cell.c) ---------------------------- */

typedef struct { Color px [64]; } WorkCell;

static void
fetch_cell (WorkCell *wc, const uint8_t *pixmap, int width_px, int cx, int cy)
{
    int y;
    for (y = 0; y < 8; y++)
        memcpy (&wc->px [y * 8], pixmap + ((size_t) (cy * 8 + y) * width_px + cx * 8) * 4, 32);
}

/* chafa-work-cell.c:293-338: widest channel; first minimum, last maximum */
static void
contrasting_pair (const WorkCell *wc, Color *bg_out, Color *fg_out)
{
    int min_i [4] = { 0, 0, 0, 0 }, max_i [4] = { 0, 0, 0, 0 };
    int ch, i, best_ch = 0, best_range;

    for (ch = 0; ch < 4; ch++)
        for (i = 1; i < 64; i++)
        {
            if (wc->px [i].ch [ch] < wc->px [min_i [ch]].ch [ch])
                min_i [ch] = i;
            if (wc->px [i].ch [ch] >= wc->px [max_i [ch]].ch [ch])
                max_i [ch] = i;
        }

    best_range = wc->px [max_i [0]].ch [0] - wc->px [min_i [0]].ch [0];
    for (ch = 1; ch < 4; ch++)
    {
        int range = wc->px [max_i [ch]].ch [ch] - wc->px [min_i [ch]].ch [ch];
        if (range > best_range)
        {
            best_range = range;
            best_ch = ch;
        }
    }
    *bg_out = wc->px [min_i [best_ch]];
    *fg_out = wc->px [max_i [best_ch]];
}

/* chafa-work-cell.c:121-143: MSB is the top-left pixel */
static uint64_t
cell_bitmap (const WorkCell *wc, const Color *bg, const Color *fg)
{
    uint64_t bm = 0;
    int i;
    for (i = 0; i < 64; i++)
    {
        bm <<= 1;
        if (diff3 (&wc->px [i], bg) > diff3 (&wc->px [i], fg))
            bm |= 1;
    }
    return bm;
}

/* chafa-avx2.c:125-165: sum * floor (32768 / n), rounded, >> 15; skipped for n <= 1 */
static uint8_t
div_rounded (int sum, int n)
{
    if (n <= 1)
        return (uint8_t) sum;
    return (uint8_t) ((sum * (32768 / n) + 16384) >> 15);
}

/* chafa-work-cell.c:76-102 */
static void
mean_colors (const WorkCell *wc, uint64_t bitmap, Color *fg_out, Color *bg_out)
{
    int sum [2] [4] = { { 0 } }, n [2] = { 0, 0 };
    int i, ch;

    for (i = 0; i < 64; i++)
    {
        int which = (bitmap >> (63 - i)) & 1;
        n [which]++;
        for (ch = 0; ch < 4; ch++)
            sum [which] [ch] += wc->px [i].ch [ch];
    }
    for (ch = 0; ch < 4; ch++)
    {
        bg_out->ch [ch] = div_rounded (sum [0] [ch], n [0]);
        fg_out->ch [ch] = div_rounded (sum [1] [ch], n [1]);
    }
}

/* chafa-avx2.c:39-81: four channels */
static int
cell_error (const WorkCell *wc, uint64_t bitmap, const Color *fg, const Color *bg)
{
    int error = 0, i, ch;
    for (i = 0; i < 64; i++)
    {
        const Color *c = ((bitmap >> (63 - i)) & 1) ? fg : bg;
        for (ch = 0; ch < 4; ch++)
        {
            int d = (int) c->ch [ch] - wc->px [i].ch [ch];
            error += d * d;
        }
    }
    return error;
}

/* chafa-color.h:96-106 */
static Color
average_2 (Color a, Color b)
{
    Color o;
    int ch;
    for (ch = 0; ch < 4; ch++)
        o.ch [ch] = (a.ch [ch] >> 1) + (b.ch [ch] >> 1);
    return o;
}

/* ---- candidate search (chafa/chafa-symbol-map.c:1153-1332) ----------------------- */

typedef struct { int sym, hd; } Cand;

static int
popcount64 (uint64_t v)
{
    int n = 0;
    while (v)
    {
        v &= v - 1;
        n++;
    }
    return n;
}

static void
cand_insert (Cand *cands, Cand nc)
{
    int i = 8 - 1;
    while (i)
    {
        i--;
        if (nc.hd >= cands [i].hd)
        {
            memmove (cands + i + 2, cands + i + 1, (8 - 2 - i) * sizeof (Cand));
            cands [i + 1] = nc;
            return;
        }
    }
    memmove (cands + 1, cands, 7 * sizeof (Cand));
    cands [0] = nc;
}

/* wide: two bitmaps per symbol, distances up to 128 */
static int
find_candidates (const uint64_t *bitmaps, int n_syms, int wide, uint64_t bm_a, uint64_t bm_b, int n_wanted,
                 Cand *out)
{
    const int max_dist = wide ? 128 : 64;
    Cand cands [8];
    int i, n;

    for (i = 0; i < 8; i++)
    {
        cands [i].sym = 0;
        cands [i].hd = max_dist + 1;
    }
    for (i = 0; i < n_syms; i++)
    {
        Cand c;
        int hd = wide ? popcount64 (bitmaps [2 * i] ^ bm_a) + popcount64 (bitmaps [2 * i + 1] ^ bm_b)
                      : popcount64 (bitmaps [i] ^ bm_a);
        c.sym = i;
        c.hd = hd;
        if (c.hd < cands [7].hd)
            cand_insert (cands, c);
        c.hd = max_dist - hd;
        if (c.hd < cands [7].hd)
            cand_insert (cands, c);
    }
    for (n = 0; n < 8 && cands [n].hd <= max_dist; n++)
        ;
    if (n > n_wanted)
        n = n_wanted;
    memcpy (out, cands, n * sizeof (Cand));
    return n;
}

/* ---- per-cell selection (chafa/internal/chafa-symbol-renderer.c:187-471) ---------- */

typedef struct { int sym; Color fg, bg; int error; } Eval;
typedef struct { int sym; Color fg, bg; int error [2]; } Eval2;

static void
eval_narrow (const OracleCanvas *cv, const WorkCell *wc, int sym, Eval *best)
{
    Color fg, bg;
    int error;

    mean_colors (wc, cv->bitmaps [sym], &fg, &bg);
    error = cell_error (wc, cv->bitmaps [sym], &fg, &bg);
    if (error < best->error)
    {
        best->sym = sym;
        best->fg = fg;
        best->bg = bg;
        best->error = error;
    }
}

static void
eval_wide (const OracleCanvas *cv, const WorkCell *wa, const WorkCell *wb, int sym, Eval2 *best)
{
    Color afg, abg, bfg, bbg, fg, bg;
    int ea, eb;

    mean_colors (wa, cv->bitmaps2 [2 * sym], &afg, &abg);
    mean_colors (wb, cv->bitmaps2 [2 * sym + 1], &bfg, &bbg);
    fg = average_2 (afg, bfg);
    bg = average_2 (abg, bbg);
    ea = cell_error (wa, cv->bitmaps2 [2 * sym], &fg, &bg);
    eb = cell_error (wb, cv->bitmaps2 [2 * sym + 1], &fg, &bg);
    if (ea + eb < best->error [0] + best->error [1])
    {
        best->sym = sym;
        best->fg = fg;
        best->bg = bg;
        best->error [0] = ea;
        best->error [1] = eb;
    }
}

static int
n_candidates_wanted (const OracleCanvas *cv)
{
    int n = cv->work_factor_int;
    return n < 1 ? 1 : n > 8 ? 8 : n;
}

static void
pick_narrow (const OracleCanvas *cv, const WorkCell *wc, Eval *best)
{
    int i;

    best->sym = -1;
    best->error = ERROR_MAX;
    if (cv->work_factor_int >= 8)
    {
        for (i = 0; i < cv->n_syms; i++)
            eval_narrow (cv, wc, i, best);
    }
    else
    {
        Cand cands [8];
        Color bg, fg;
        int n;

        contrasting_pair (wc, &bg, &fg);
        n = find_candidates (cv->bitmaps, cv->n_syms, 0, cell_bitmap (wc, &bg, &fg), 0,
                             n_candidates_wanted (cv), cands);
        for (i = 0; i < n; i++)
            eval_narrow (cv, wc, cands [i].sym, best);
    }
}

static void
pick_wide (const OracleCanvas *cv, const WorkCell *wa, const WorkCell *wb, Eval2 *best)
{
    int i;

    best->sym = -1;
    best->error [0] = best->error [1] = ERROR_MAX;
    if (cv->work_factor_int >= 8)
    {
        for (i = 0; i < cv->n_syms2; i++)
            eval_wide (cv, wa, wb, i, best);
    }
    else
    {
        Cand cands [8];
        Color abg, afg, bbg, bfg, bg, fg;
        int n;

        contrasting_pair (wa, &abg, &afg);
        contrasting_pair (wb, &bbg, &bfg);
        bg = average_2 (abg, bbg);
        fg = average_2 (afg, bfg);
        n = find_candidates (cv->bitmaps2, cv->n_syms2, 1, cell_bitmap (wa, &bg, &fg), cell_bitmap (wb, &bg, &fg),
                             n_candidates_wanted (cv), cands);
        for (i = 0; i < n; i++)
            eval_wide (cv, wa, wb, cands [i].sym, best);
    }
}

/* chafa-symbol-renderer.c:699-729; packed colour is A << 24 | R << 16 | G << 8 | B */
static void
cell_colors (const OracleCanvas *cv, const Color *fg, const Color *bg, OracleCell *cell)
{
    if (cv->mode == MODE_TRUECOLOR)
    {
        cell->fg = ((uint32_t) fg->ch [3] << 24) | ((uint32_t) fg->ch [0] << 16) | ((uint32_t) fg->ch [1] << 8) | fg->ch [2];
        cell->bg = ((uint32_t) bg->ch [3] << 24) | ((uint32_t) bg->ch [0] << 16) | ((uint32_t) bg->ch [1] << 8) | bg->ch [2];
    }
    else
    {
        cell->fg = lookup_nearest (cv->mode, cv->alpha_threshold, fg);
        cell->bg = lookup_nearest (cv->mode, cv->alpha_threshold, bg);
    }
}

/* chafa-symbol-renderer.c:797-895 without fill symbols */
void
oracle_update_cells (const OracleCanvas *cv, const uint8_t *pixmap, OracleCell *cells)
{
    int cx, cy;

    init_palette ();

    for (cy = 0; cy < cv->height; cy++)
    {
        OracleCell *row = cells + (size_t) cy * cv->width;
        WorkCell prev_wc, wc;
        int prev_error = 0;

        for (cx = 0; cx < cv->width; cx++)
        {
            Eval best;
            int error = ERROR_MAX;

            row [cx].c = ' ';
            row [cx].fg = row [cx].bg = 0;
            fetch_cell (&wc, pixmap, cv->width * 8, cx, cy);

            if (cv->n_syms > 0)
            {
                pick_narrow (cv, &wc, &best);
                row [cx].c = cv->chars [best.sym];
                cell_colors (cv, &best.fg, &best.bg, &row [cx]);
                error = best.error;
            }

            if (cx >= 1 && row [cx - 1].c != 0 && cv->n_syms2 > 0)
            {
                Eval2 wbest;

                pick_wide (cv, &prev_wc, &wc, &wbest);
                if (wbest.sym >= 0 && wbest.error [0] + wbest.error [1] < prev_error + error)
                {
                    OracleCell left;
                    left.c = cv->chars2 [wbest.sym];
                    cell_colors (cv, &wbest.fg, &wbest.bg, &left);
                    row [cx - 1] = left;
                    row [cx] = left;
                    row [cx].c = left.c == cv->solid_char ? left.c : 0;
                    error = wbest.error [1];
                }
            }

            if (row [cx].c != 0 && (row [cx].c == ' ' || row [cx].fg == row [cx].bg))
            {
                row [cx].c = cv->blank_char;
                if (cv->blank_char == ' ' && cx > 0)
                {
                    row [cx].fg = row [cx - 1].fg;
                    if (cv->mode == MODE_TRUECOLOR)
                        row [cx].fg |= 0xff000000u;
                    else if (row [cx].fg == PEN_TRANSPARENT)
                        row [cx].fg = PEN_FG;
                }
            }

            prev_wc = wc;
            prev_error = error;
        }
    }
}

/* ---- printing (chafa/internal/chafa-canvas-printer.c:56-763, fallback sequences of
 * chafa/chafa-term-db.c:569-582: no repeat-character sequence) ----------------------- */

typedef struct
{
    char *p;
    uint32_t fg, bg;    /* truecolor: 0xAARRGGBB after thresholding; indexed: pen */
    int inverted;
}
Printer;

static void
put_str (Printer *pr, const char *s)
{
    size_t n = strlen (s);
    memcpy (pr->p, s, n);
    pr->p += n;
}

static void
put_u8 (Printer *pr, unsigned v)
{
    v &= 0xff;
    if (v >= 100) *pr->p++ = '0' + v / 100;
    if (v >= 10) *pr->p++ = '0' + (v / 10) % 10;
    *pr->p++ = '0' + v % 10;
}

static void
put_utf8 (Printer *pr, uint32_t c)
{
    if (c < 0x80)
        *pr->p++ = c;
    else if (c < 0x800)
    {
        *pr->p++ = 0xc0 | (c >> 6);
        *pr->p++ = 0x80 | (c & 0x3f);
    }
    else if (c < 0x10000)
    {
        *pr->p++ = 0xe0 | (c >> 12);
        *pr->p++ = 0x80 | ((c >> 6) & 0x3f);
        *pr->p++ = 0x80 | (c & 0x3f);
    }
    else
    {
        *pr->p++ = 0xf0 | (c >> 18);
        *pr->p++ = 0x80 | ((c >> 12) & 0x3f);
        *pr->p++ = 0x80 | ((c >> 6) & 0x3f);
        *pr->p++ = 0x80 | (c & 0x3f);
    }
}

static void
reset_attributes (Printer *pr, int mode)
{
    put_str (pr, "\033[0m");
    pr->inverted = 0;
    if (mode == MODE_TRUECOLOR)
    {
        pr->fg &= 0x00ffffffu;
        pr->bg &= 0x00ffffffu;
    }
    else
        pr->fg = pr->bg = PEN_TRANSPARENT;
}

static void
put_rgb (Printer *pr, uint32_t c)
{
    put_u8 (pr, c >> 16);
    *pr->p++ = ';';
    put_u8 (pr, c >> 8);
    *pr->p++ = ';';
    put_u8 (pr, c);
}

/* chafa-canvas-printer.c:139-213 */
static void
attributes_truecolor (Printer *pr, int optimizations, uint32_t fg, uint32_t bg, int inverted)
{
    int fg_on = (fg >> 24) != 0, bg_on = (bg >> 24) != 0;

    if (optimizations & OPT_REUSE_ATTRIBUTES)
    {
        if ((pr->inverted && !inverted) || ((pr->fg >> 24) != 0 && !fg_on) || ((pr->bg >> 24) != 0 && !bg_on))
            reset_attributes (pr, MODE_TRUECOLOR);
        if (!pr->inverted && inverted)
            put_str (pr, "\033[7m");

        if (fg != pr->fg)
        {
            if (bg != pr->bg && bg_on)
            {
                put_str (pr, "\033[38;2;");
                put_rgb (pr, fg);
                put_str (pr, ";48;2;");
                put_rgb (pr, bg);
                *pr->p++ = 'm';
            }
            else if (fg_on)
            {
                put_str (pr, "\033[38;2;");
                put_rgb (pr, fg);
                *pr->p++ = 'm';
            }
        }
        else if (bg != pr->bg && bg_on)
        {
            put_str (pr, "\033[48;2;");
            put_rgb (pr, bg);
            *pr->p++ = 'm';
        }
    }
    else
    {
        reset_attributes (pr, MODE_TRUECOLOR);
        if (inverted)
            put_str (pr, "\033[7m");
        if (fg_on)
        {
            put_str (pr, "\033[38;2;");
            put_rgb (pr, fg);
            if (bg_on)
            {
                put_str (pr, ";48;2;");
                put_rgb (pr, bg);
            }
            *pr->p++ = 'm';
        }
        else if (bg_on)
        {
            put_str (pr, "\033[48;2;");
            put_rgb (pr, bg);
            *pr->p++ = 'm';
        }
    }
    pr->fg = fg;
    pr->bg = bg;
    pr->inverted = inverted;
}

static void
put_pen_fg (Printer *pr, int mode, uint32_t pen)
{
    if (mode == MODE_INDEXED_16)
        put_u8 (pr, (pen & 0xff) + ((pen & 0xff) < 8 ? 30 : 82));
    else
    {
        put_str (pr, "38;5;");
        put_u8 (pr, pen);
    }
}

static void
put_pen_bg (Printer *pr, int mode, uint32_t pen)
{
    if (mode == MODE_INDEXED_16)
        put_u8 (pr, (pen & 0xff) + ((pen & 0xff) < 8 ? 40 : 92));
    else
    {
        put_str (pr, "48;5;");
        put_u8 (pr, pen);
    }
}

/* chafa-canvas-printer.c:254-341, 380-433 */
static void
attributes_indexed (Printer *pr, int mode, int optimizations, uint32_t fg, uint32_t bg, int inverted)
{
    if (optimizations & OPT_REUSE_ATTRIBUTES)
    {
        if ((pr->inverted && !inverted) || (pr->fg != PEN_TRANSPARENT && fg == PEN_TRANSPARENT)
            || (pr->bg != PEN_TRANSPARENT && bg == PEN_TRANSPARENT))
            reset_attributes (pr, mode);
        if (!pr->inverted && inverted)
            put_str (pr, "\033[7m");

        if (fg != pr->fg)
        {
            if (bg != pr->bg && bg != PEN_TRANSPARENT)
            {
                put_str (pr, "\033[");
                put_pen_fg (pr, mode, fg);
                *pr->p++ = ';';
                put_pen_bg (pr, mode, bg);
                *pr->p++ = 'm';
            }
            else if (fg != PEN_TRANSPARENT)
            {
                put_str (pr, "\033[");
                put_pen_fg (pr, mode, fg);
                *pr->p++ = 'm';
            }
        }
        else if (bg != pr->bg && bg != PEN_TRANSPARENT)
        {
            put_str (pr, "\033[");
            put_pen_bg (pr, mode, bg);
            *pr->p++ = 'm';
        }
    }
    else
    {
        reset_attributes (pr, mode);
        if (inverted)
            put_str (pr, "\033[7m");
        if (fg != PEN_TRANSPARENT)
        {
            put_str (pr, "\033[");
            put_pen_fg (pr, mode, fg);
            if (bg != PEN_TRANSPARENT)
            {
                *pr->p++ = ';';
                put_pen_bg (pr, mode, bg);
            }
            *pr->p++ = 'm';
        }
        else if (bg != PEN_TRANSPARENT)
        {
            put_str (pr, "\033[");
            put_pen_bg (pr, mode, bg);
            *pr->p++ = 'm';
        }
    }
    pr->fg = fg;
    pr->bg = bg;
    pr->inverted = inverted;
}

static uint32_t
threshold_argb (uint32_t argb, int alpha_threshold)
{
    return (int) (argb >> 24) < alpha_threshold ? (argb & 0x00ffffffu) : (argb | 0xff000000u);
}

/* Returns the number of bytes written; @out must hold (cells + rows) * 64 bytes */
size_t
oracle_print (const OracleCanvas *cv, const OracleCell *cells, char *out)
{
    Printer pr;
    int x, y;

    pr.p = out;
    pr.fg = pr.bg = cv->mode == MODE_TRUECOLOR ? 0 : PEN_TRANSPARENT;
    pr.inverted = 0;

    for (y = 0; y < cv->height; y++)
    {
        const OracleCell *row = cells + (size_t) y * cv->width;

        if (y == 0)
            reset_attributes (&pr, cv->mode);

        for (x = 0; x < cv->width; x++)
        {
            const OracleCell *cell = &row [x];
            int both_transparent;

            if (cell->c == 0)
                continue;

            if (cv->mode == MODE_TRUECOLOR)
            {
                uint32_t fg = threshold_argb (cell->fg, cv->alpha_threshold);
                uint32_t bg = threshold_argb (cell->bg, cv->alpha_threshold);
                int fg_on = (fg >> 24) != 0, bg_on = (bg >> 24) != 0;

                if (!fg_on && bg_on)
                    attributes_truecolor (&pr, cv->optimizations, bg, fg, 1);
                else
                    attributes_truecolor (&pr, cv->optimizations, fg, bg, 0);
                both_transparent = !fg_on && !bg_on;
            }
            else
            {
                uint32_t fg = cell->fg, bg = cell->bg;

                if (fg == PEN_TRANSPARENT && bg != PEN_TRANSPARENT)
                    attributes_indexed (&pr, cv->mode, cv->optimizations, bg, fg, 1);
                else
                    attributes_indexed (&pr, cv->mode, cv->optimizations, fg, bg, 0);
                both_transparent = fg == PEN_TRANSPARENT && bg == PEN_TRANSPARENT;
            }

            if (both_transparent)
            {
                *pr.p++ = ' ';
                if (x < cv->width - 1 && row [x + 1].c == 0)
                    *pr.p++ = ' ';
            }
            else
                put_utf8 (&pr, cell->c);
        }

        reset_attributes (&pr, cv->mode);
        if (y < cv->height - 1)
            *pr.p++ = '\n';
    }
    return (size_t) (pr.p - out);
}
