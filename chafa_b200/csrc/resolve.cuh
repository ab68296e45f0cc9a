/* resolve.cuh -- kernel K4: turn per-cell and per-pair match results into the final row
 * of canvas cells.
 *
 * Restates the serial part of update_cells_row (chafa/internal/chafa-symbol-renderer.c:
 * 797-895): greedy replacement of two narrow cells by one wide symbol when that lowers
 * the summed error, fill symbols for featureless cells (apply_fill 544-654,
 * apply_fill_fg_only 485-542), blank-character substitution and the foreground-colour
 * carry from the previous cell.  Rows are independent; within a row the walk is serial
 * in cx, so one thread owns one row.  All heavy work (matching) was done by K3 / K3w. */

#ifndef CHAFA_B200_RESOLVE_CUH
#define CHAFA_B200_RESOLVE_CUH

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

/* ------------------------------------------------------------------------ */
/* Row walk as scans.
 *
 * The reference's loop carries three things from cell cx - 1 to cell cx:
 *
 *  (1) whether cx - 1 became the right half of a wide symbol (then no wide symbol is
 *      tried at cx), and which error it ended up with.  A cell is in one of three
 *      states -- N: narrow result kept; M: right half of an adopted wide symbol;
 *      S: like M but the pair was reverted to the (narrow) solid character by the
 *      16/8 quantiser, so it still counts as a left neighbour.  The state of cx is a
 *      function of the state of cx - 1 and of per-cell data, i.e. a 3 -> 3 map per
 *      cell; maps compose associatively, so an inclusive scan yields every state.
 *  (2) the left cell is overwritten when a wide symbol is adopted: a plain store
 *      after the cell's own result has been written.
 *  (3) featureless cells copy the foreground of their left neighbour as it stands
 *      at that moment, so runs of them copy from the nearest cell to the left that
 *      did not copy: a max-scan over "index of the last non-copying cell".  The
 *      adjustment applied on copying is idempotent, so applying it once is enough.
 *
 * One thread block owns one row; RESOLVE_THREADS cells per sweep with block-level
 * carries between sweeps. */

#define RESOLVE_THREADS 512
#define RESOLVE_WARPS (RESOLVE_THREADS / 32)

__device__ __forceinline__ uint32_t
map_apply (uint32_t m, uint32_t x)
{
    return (m >> (2 * x)) & 3u;
}

/* (first, then second) */
__device__ __forceinline__ uint32_t
map_compose (uint32_t first, uint32_t second)
{
    return map_apply (second, map_apply (first, 0)) | (map_apply (second, map_apply (first, 1)) << 2)
        | (map_apply (second, map_apply (first, 2)) << 4);
}

#define MAP_IDENTITY (0u | (1u << 2) | (2u << 4))

/* Row @r of the band: called by every thread of a block of RESOLVE_THREADS (it synchronises the block) */
__device__ __forceinline__ void
resolve_row (const DevCanvas &cv, const DevCellEval *__restrict__ evals,
             const DevWideEval *__restrict__ wide_evals, DevCell *__restrict__ cells, const int r)
{
    __shared__ uint32_t s_warp_map [RESOLVE_WARPS];
    __shared__ int s_warp_src [RESOLVE_WARPS];
    __shared__ uint32_t s_fg [RESOLVE_THREADS];
    __shared__ uint32_t s_carry_state, s_carry_fg;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevCellEval *ev = evals + (size_t) r * cv.width;
    const DevWideEval *wv = wide_evals ? wide_evals + (size_t) r * cv.width : nullptr;
    DevCell *row = cells + (size_t) (cv.first_row + r) * cv.width;
    const bool have_wide = wv != nullptr && cv.syms.n2 > 0;
    const int mode = cv.canvas_mode;

    if (tid == 0)
    {
        s_carry_state = 0;
        s_carry_fg = 0;
    }
    __syncthreads ();

    for (int base = 0; base < cv.width; base += RESOLVE_THREADS)
    {
        const int cx = base + tid;
        const bool valid = cx < cv.width;
        DevCellEval e;
        DevWideEval w;

        e.c = ' ';
        e.fg = e.bg = e.mean = 0;
        e.error = 0;
        w.c = 0;
        w.fg = w.bg = 0;
        w.error_a = w.error_b = 0;
        if (valid)
        {
            e = ev [cx];
            if (have_wide)
                w = wv [cx];
        }

        /* errors of the left neighbour */
        int left_err = __shfl_up_sync (0xffffffffu, e.error, 1);
        int left_werr = __shfl_up_sync (0xffffffffu, w.error_b, 1);
        if (lane == 0)
        {
            /* from the previous warp / sweep: read directly (L1/L2 hit) */
            if (cx >= 1 && valid)
            {
                left_err = ev [cx - 1].error;
                left_werr = have_wide ? wv [cx - 1].error_b : 0;
            }
        }

        /* (1) state maps */
        uint32_t map = MAP_IDENTITY;
        if (valid)
        {
            uint32_t to_n = 0, to_m = 0;
            bool try_wide = have_wide && cx >= 1;
            int wsum = w.error_a + w.error_b;
            bool cond_n = try_wide && wsum < left_err + e.error;
            bool cond_s = try_wide && wsum < left_werr + e.error;
            to_m = (w.c == cv.solid_char) ? 2u : 1u;
            map = (cond_n ? to_m : to_n) | (0u << 2) | ((cond_s ? to_m : to_n) << 4);
        }

        uint32_t incl = map;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            uint32_t other = __shfl_up_sync (0xffffffffu, incl, d);
            if (lane >= d)
                incl = map_compose (other, incl);
        }
        if (lane == 31)
            s_warp_map [warp] = incl;
        __syncthreads ();
        uint32_t prefix = MAP_IDENTITY;
        for (int k = 0; k < warp; k++)
            prefix = map_compose (prefix, s_warp_map [k]);
        uint32_t total = prefix;
        for (int k = warp; k < RESOLVE_WARPS; k++)
            total = map_compose (total, s_warp_map [k]);
        const uint32_t state_in = s_carry_state;
        const uint32_t state = map_apply (map_compose (prefix, incl), state_in);
        const uint32_t state_out = map_apply (total, state_in);
        const bool merged = valid && state != 0;

        /* the cell as it stands after its own step */
        DevCell cur;
        if (merged)
        {
            cur.c = state == 2 ? w.c : 0;
            cur.fg = w.fg;
            cur.bg = w.bg;
        }
        else
        {
            cur.c = e.c;
            cur.fg = e.fg;
            cur.bg = e.bg;
        }

        bool copies = false;
        if (valid)
        {
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
                copies = cv.blank_char == ' ' && cx > 0;
            }
        }

        /* (3) foreground copy.  A merged cell copies from the left half written by its
         * own step, which is known here. */
        if (copies && merged)
        {
            cur.fg = w.fg;
            if (mode == MODE_TRUECOLOR)
                cur.fg |= 0xff000000u;
            else if (cur.fg == PAL_INDEX_TRANSPARENT)
                cur.fg = PAL_INDEX_FG;
            copies = false;
        }

        s_fg [tid] = cur.fg;
        int src = copies ? -1 : tid;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int other = __shfl_up_sync (0xffffffffu, src, d);
            if (lane >= d)
                src = max (src, other);
        }
        if (lane == 31)
            s_warp_src [warp] = src;
        __syncthreads ();
        for (int k = 0; k < warp; k++)
            src = max (src, s_warp_src [k]);
        if (copies)
        {
            uint32_t fg = src >= 0 ? s_fg [src] : s_carry_fg;
            if (mode == MODE_TRUECOLOR)
                fg |= 0xff000000u;
            else if (fg == PAL_INDEX_TRANSPARENT)
                fg = PAL_INDEX_FG;
            cur.fg = fg;
        }

        /* (2) stores: own cell first, then the left half of adopted wide symbols */
        if (valid)
            row [cx] = cur;
        __syncthreads ();
        if (merged)
        {
            DevCell left;
            left.c = w.c;
            left.fg = w.fg;
            left.bg = w.bg;
            row [cx - 1] = left;
        }

        /* carries for the next sweep */
        if (tid == RESOLVE_THREADS - 1)
        {
            s_carry_state = state_out;
            s_carry_fg = cur.fg;
        }
        __syncthreads ();
    }
}

#endif
