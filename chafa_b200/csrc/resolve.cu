/* resolve.cu -- kernel K4: turn per-cell and per-pair match results into the final row
 * of canvas cells.
 *
 * Restates the serial part of update_cells_row (chafa/internal/chafa-symbol-renderer.c:
 * 797-895): greedy replacement of two narrow cells by one wide symbol when that lowers
 * the summed error, fill symbols for featureless cells (apply_fill 544-654,
 * apply_fill_fg_only 485-542), blank-character substitution and the foreground-colour
 * carry from the previous cell.  Rows are independent; within a row the walk is serial
 * in cx, so one thread owns one row.  All heavy work (matching) was done by K3 / K3w. */

#include <cuda_runtime.h>

#include "device.h"
#include "palette.cuh"

#define MODE_TRUECOLOR 0
#define MODE_INDEXED_256 1
#define MODE_INDEXED_240 2
#define MODE_INDEXED_16 3
#define MODE_FGBG_BGFG 4
#define MODE_FGBG 5
#define MODE_INDEXED_8 6
#define MODE_INDEXED_16_8 7

__device__ __forceinline__ uint32_t
transparent_cell_color_r (int mode)
{
    return mode == MODE_TRUECOLOR ? 0x00808080u : (uint32_t) PAL_INDEX_TRANSPARENT;
}

/* chafa/chafa-symbol-map.c:1335-1426 on the fill map (sorted by popcount) */
__device__ static int
find_closest_popcount (const DevSymbols &syms, int popcount)
{
    int n = syms.n_fill;
    int i = 0, j = n - 1;

    while (i < j)
    {
        int k = (i + j + 1) / 2;
        if (popcount < (int) syms.fill_popcounts [k])
            j = k - 1;
        else
            i = k;
    }
    if (i < n - 1
        && abs (popcount - (int) syms.fill_popcounts [i + 1]) < abs (popcount - (int) syms.fill_popcounts [i]))
        i++;
    return i;
}

__device__ static void
find_fill_candidate (const DevSymbols &syms, int popcount, bool do_inverse, int &sym_out, bool &inverted_out)
{
    int sym = find_closest_popcount (syms, popcount);
    int dist = abs (popcount - (int) syms.fill_popcounts [sym]);
    bool inverted = false;

    if (do_inverse && dist != 0)
    {
        int sym2 = find_closest_popcount (syms, 64 - popcount);
        int d2 = abs (64 - popcount - (int) syms.fill_popcounts [sym2]);
        if (d2 < dist)
        {
            sym = sym2;
            inverted = true;
        }
    }
    sym_out = sym;
    inverted_out = inverted;
}

__device__ static void
apply_fill_fg_only (const DevCanvas &cv, uint32_t mean, DevCell &cell)
{
    if (cv.syms.n_fill == 0)
        return;

    if (cv.canvas_mode == MODE_TRUECOLOR)
    {
        cell.fg = pack_argb (mean);
    }
    else
    {
        ColorCandidates cc;
        palette_lookup_nearest (cv.fg_palette, cv.color_space, mean, &cc);
        cell.fg = (uint32_t) cc.index [0];
    }
    cell.bg = transparent_cell_color_r (cv.canvas_mode);

    int fg_value = (CH (cv.default_fg, 0) + CH (cv.default_fg, 1) + CH (cv.default_fg, 2)) / 3;
    int bg_value = (CH (cv.default_bg, 0) + CH (cv.default_bg, 1) + CH (cv.default_bg, 2)) / 3;
    int mean_value = (CH (mean, 0) + CH (mean, 1) + CH (mean, 2)) / 3;
    int n_bits = ((mean_value * 64) + 128) / 255;
    if (fg_value < bg_value)
        n_bits = 64 - n_bits;

    int sym;
    bool inv;
    find_fill_candidate (cv.syms, n_bits, false, sym, inv);
    cell.c = cv.syms.fill_chars [sym];
}

__device__ static void
apply_fill (const DevCanvas &cv, uint32_t mean, DevCell &cell)
{
    if (cv.syms.n_fill == 0)
        return;

    int mode = cv.canvas_mode;
    int sym;
    bool inverted;
    ColorCandidates cc;

    if (mode == MODE_TRUECOLOR)
    {
        cell.bg = cell.fg = pack_argb (mean);
        find_fill_candidate (cv.syms, 0, false, sym, inverted);
        cell.c = cv.syms.fill_chars [sym];
        return;
    }
    else if (mode == MODE_INDEXED_256 || mode == MODE_INDEXED_240 || mode == MODE_INDEXED_16 || mode == MODE_INDEXED_8)
    {
        palette_lookup_nearest (cv.fg_palette, cv.color_space, mean, &cc);
    }
    else if (mode == MODE_INDEXED_16_8)
    {
        ColorCandidates cb;
        palette_lookup_nearest (cv.fg_palette, cv.color_space, mean, &cc);
        palette_lookup_nearest (cv.bg_palette, cv.color_space, mean, &cb);
        if (cc.index [0] != cb.index [0])
        {
            if (cc.index [1] == cb.index [0])
                cc.index [1] = cb.index [1];
            cc.index [0] = cb.index [0];
        }
    }
    else
    {
        cc.index [0] = PAL_INDEX_FG;
        cc.index [1] = PAL_INDEX_BG;
    }

    uint32_t col0 = cv.fg_palette->colors [cc.index [0]];
    uint32_t col1 = cv.fg_palette->colors [cc.index [1]];

    if (mode == MODE_FGBG_BGFG || mode == MODE_FGBG)
        col1 |= 0xff000000u;

    int best_i = 0, best_error = 0x7fffffff;
    for (int i = 0; i <= 64; i++)
    {
        /* The interpolated alpha is computed by the reference too but does not enter
         * the 3-channel difference */
        int c0 = (CH (col0, 0) * (64 - i) + CH (col1, 0) * i) / 64;
        int c1 = (CH (col0, 1) * (64 - i) + CH (col1, 1) * i) / 64;
        int c2 = (CH (col0, 2) * (64 - i) + CH (col1, 2) * i) / 64;
        uint32_t col2 = (uint32_t) (c0 & 0xff) | ((uint32_t) (c1 & 0xff) << 8) | ((uint32_t) (c2 & 0xff) << 16);
        int error = color_diff3 (mean, col2);
        if (error < best_error)
        {
            best_i = (mode == MODE_FGBG) ? 64 - i : i;
            best_error = error;
        }
    }

    find_fill_candidate (cv.syms, best_i, cv.consider_inverted && mode != MODE_INDEXED_16_8, sym, inverted);

    if (mode != MODE_FGBG && mode != MODE_INDEXED_16_8)
    {
        if (best_i == 0)
            cc.index [1] = cc.index [0];
        else if (best_i == 64)
            cc.index [0] = cc.index [1];
    }

    if (inverted)
    {
        cell.fg = (uint32_t) cc.index [0];
        cell.bg = (uint32_t) cc.index [1];
    }
    else
    {
        cell.fg = (uint32_t) cc.index [1];
        cell.bg = (uint32_t) cc.index [0];
    }
    cell.c = cv.syms.fill_chars [sym];
}

__global__ void
resolve_rows_kernel (DevCanvas cv, const DevCellEval *__restrict__ evals,
                     const DevWideEval *__restrict__ wide_evals, DevCell *__restrict__ cells)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cv.n_rows)
        return;

    const DevCellEval *ev = evals + (size_t) r * cv.width;
    const DevWideEval *wv = wide_evals ? wide_evals + (size_t) r * cv.width : nullptr;
    DevCell *row = cells + (size_t) (cv.first_row + r) * cv.width;
    const bool have_wide = wv != nullptr && cv.syms.n2 > 0;
    const int mode = cv.canvas_mode;

    DevCell prev;           /* cells [cx - 1] as it currently stands */
    int prev_error = 0;
    prev.c = 0;
    prev.fg = prev.bg = 0;

    for (int cx = 0; cx < cv.width; cx++)
    {
        DevCellEval e = ev [cx];
        DevCell cur;
        int cur_error = e.error;

        cur.c = e.c;
        cur.fg = e.fg;
        cur.bg = e.bg;

        if (have_wide && cx >= 1 && prev.c != 0)
        {
            DevWideEval w = wv [cx];

            /* error sums may exceed SYMBOL_ERROR_MAX * 2 only when no symbols exist; int is enough */
            if (w.error_a + w.error_b < prev_error + cur_error)
            {
                prev.c = w.c;
                prev.fg = w.fg;
                prev.bg = w.bg;
                cur.c = 0;
                cur.fg = w.fg;
                cur.bg = w.bg;
                if (w.c == cv.solid_char)
                    cur.c = w.c;
                cur_error = w.error_b;
                row [cx - 1] = prev;
            }
        }

        if (cur.c != 0 && (cur.c == ' ' || cur.c == 0x2588 || cur.fg == cur.bg))
        {
            if (cv.fg_only)
            {
                apply_fill_fg_only (cv, e.mean, cur);
                cur.bg = transparent_cell_color_r (mode);
            }
            else
            {
                apply_fill (cv, e.mean, cur);
            }
        }

        if (cur.c != 0 && (cur.c == ' ' || cur.fg == cur.bg))
        {
            cur.c = cv.blank_char;

            if (cv.blank_char == ' ' && cx > 0)
            {
                cur.fg = prev.fg;
                if (mode == MODE_TRUECOLOR)
                    cur.fg |= 0xff000000u;
                else if (cur.fg == PAL_INDEX_TRANSPARENT)
                    cur.fg = PAL_INDEX_FG;
            }
        }

        row [cx] = cur;
        prev = cur;
        prev_error = cur_error;
    }
}

int
dev_launch_resolve (const DevCanvas *cv, const DevCellEval *evals, const DevWideEval *wide_evals,
                    DevCell *cells, void *stream)
{
    if (cv->n_rows <= 0)
        return 0;
    int threads = 32;
    int blocks = (cv->n_rows + threads - 1) / threads;
    resolve_rows_kernel<<<blocks, threads, 0, (cudaStream_t) stream>>> (*cv, evals, wide_evals, cells);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

/* Host-side call of the same picker code, for the cell colour setters of the public
 * API (chafa/chafa-canvas.c:118-131); not on the rendering path. */
int
host_palette_lookup_nearest (const DevPalette *pal, int color_space, uint32_t color)
{
    return palette_lookup_nearest (pal, color_space, color, nullptr);
}
