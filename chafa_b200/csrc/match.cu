/* match.cu -- kernel K3 / K3w: per-cell symbol and colour selection.
 *
 * One warp owns one 8x8 cell (K3) or one pair of adjacent cells (K3w).  The symbol
 * bitmaps sit in shared memory for the lifetime of a persistent block.  Per cell:
 *
 *   contrasting colour pair -> 64-bit outline bitmap -> Hamming distance to every
 *   symbol (xor + popc) -> stable top-K by (distance, evaluation order) -> for each
 *   candidate: masked fg/bg mean colours, full 4-channel squared error -> argmin
 *   (first wins) -> cell colours (packed ARGB or palette pens).
 *
 * Work factors >= 0.8 skip the prefilter and evaluate every symbol.
 *
 * Arithmetic follows the reference's AVX2 build:
 *   chafa/internal/chafa-work-cell.c:76-102, 121-143, 293-338
 *   chafa/internal/chafa-avx2.c:39-81 (4-channel error), 83-165 (mean = mulhrs by 32768/n)
 *   chafa/chafa-symbol-map.c:1153-1332 (candidate order)
 *   chafa/internal/chafa-symbol-renderer.c:187-471, 656-753 (evaluation, cell colours)
 *
 * Lane l of the warp holds pixel l (rows 0-3) and pixel l + 32 (rows 4-7) of the
 * cell; the bitmap's MSB is pixel 0. */

#include <cuda_runtime.h>

#include "device.h"
#include "palette.cuh"

#define WARPS_PER_BLOCK 8
#define CAND_LIST_MAX 128
#define CAND_LIST_STRIDE (CAND_LIST_MAX + 1)   /* + the list length */
/* candidate key: distance (<= 128) above the sequence number 2 * symbol + inverted; 23 bits of
 * sequence number = four million symbols, beyond anything a symbol map holds */
#define CAND_KEY_SHIFT 23
#define CAND_KEY_SEQ_MASK 0x7fffffu
#define FULL 0xffffffffu

#define MODE_TRUECOLOR 0
#define MODE_INDEXED_256 1
#define MODE_INDEXED_240 2
#define MODE_INDEXED_16 3
#define MODE_FGBG_BGFG 4
#define MODE_FGBG 5
#define MODE_INDEXED_8 6
#define MODE_INDEXED_16_8 7

#define SYMBOL_ERROR_MAX (0x7fffffff / 8)

/* floor (32768 / n) for n = 0..64 (n = 0 -> 0) */
__constant__ uint16_t c_invdiv16 [65] =
{
    0, 32768, 16384, 10922, 8192, 6553, 5461, 4681, 4096, 3640, 3276, 2978, 2730, 2520, 2340, 2184,
    2048, 1927, 1820, 1724, 1638, 1560, 1489, 1424, 1365, 1310, 1260, 1213, 1170, 1129, 1092, 1057,
    1024, 992, 963, 936, 910, 885, 862, 840, 819, 799, 780, 762, 744, 728, 712, 697,
    682, 668, 655, 642, 630, 618, 606, 595, 585, 574, 564, 555, 546, 537, 528, 520, 512
};

struct CellPix
{
    uint32_t p0, p1;              /* this lane's two pixels */
    uint32_t rb0, ga0, rb1, ga1;  /* the same, split into two 16-bit fields per word */
    uint32_t tot_rb, tot_ga;      /* whole-cell channel sums, two 16-bit fields per word */
};

__device__ __forceinline__ void
load_cell (CellPix &cp, const uint32_t *pixels, int width_px, int cx, int cy, int lane)
{
    const uint32_t *base = pixels + (size_t) (cy * 8 + (lane >> 3)) * width_px + cx * 8 + (lane & 7);
    cp.p0 = __ldg (base);
    cp.p1 = __ldg (base + (size_t) 4 * width_px);

    cp.rb0 = cp.p0 & 0x00ff00ffu;
    cp.ga0 = (cp.p0 >> 8) & 0x00ff00ffu;
    cp.rb1 = cp.p1 & 0x00ff00ffu;
    cp.ga1 = (cp.p1 >> 8) & 0x00ff00ffu;
    cp.tot_rb = __reduce_add_sync (FULL, cp.rb0 + cp.rb1);
    cp.tot_ga = __reduce_add_sync (FULL, cp.ga0 + cp.ga1);
}

/* Channel sums -> colour: mulhrs (sum, floor (32768 / n)) per channel
 * (chafa/internal/chafa-avx2.c:125-165).  The reference skips the division when
 * n <= 1; the formula gives the same result there (n = 1: (sum * 32768 + 16384) >> 15
 * == sum; n = 0: sum == 0).  The quotient never exceeds 255. */
__device__ __forceinline__ uint32_t
sums_to_color (uint32_t rb, uint32_t ga, int n)
{
    uint32_t inv = c_invdiv16 [n];
    uint32_t r = ((rb & 0xffffu) * inv + 16384u) >> 15;
    uint32_t b = ((rb >> 16) * inv + 16384u) >> 15;
    uint32_t g = ((ga & 0xffffu) * inv + 16384u) >> 15;
    uint32_t a = ((ga >> 16) * inv + 16384u) >> 15;
    return r | (g << 8) | (b << 16) | (a << 24);
}

/* Mean colour of all 64 pixels (chafa-work-cell.c:104-119) */
__device__ __forceinline__ uint32_t
cell_mean_color (const CellPix &cp)
{
    return sums_to_color (cp.tot_rb, cp.tot_ga, 64);
}

/* Does the symbol cover this lane's pixel in the upper / lower half of the cell?
 * The bitmap's MSB is pixel 0, so lane l owns bit 31 - l of each half. */
__device__ __forceinline__ void
lane_bits (uint64_t bm, int lane, bool &b0, bool &b1)
{
    b0 = (int) ((uint32_t) (bm >> 32) << lane) < 0;
    b1 = (int) ((uint32_t) bm << lane) < 0;
}

/* Masked mean colours for one symbol (chafa-work-cell.c:76-102) */
__device__ __forceinline__ void
mean_colors_for_symbol (const CellPix &cp, uint64_t bm, int lane, uint32_t &fg_out, uint32_t &bg_out)
{
    bool b0, b1;
    lane_bits (bm, lane, b0, b1);

    uint32_t rb = (b0 ? cp.rb0 : 0u) + (b1 ? cp.rb1 : 0u);
    uint32_t ga = (b0 ? cp.ga0 : 0u) + (b1 ? cp.ga1 : 0u);
    uint32_t fg_rb = __reduce_add_sync (FULL, rb);
    uint32_t fg_ga = __reduce_add_sync (FULL, ga);
    int n_fg = __popcll (bm);

    fg_out = sums_to_color (fg_rb, fg_ga, n_fg);
    bg_out = sums_to_color (cp.tot_rb - fg_rb, cp.tot_ga - fg_ga, 64 - n_fg);
}

/* ---- median extractor (chafa-work-cell.c:147-160, 206-291, 340-389) -------------
 * The reference orders the 64 pixels by one channel with a stable counting sort, i.e. by
 * (value, pixel index), and takes the pixel in the middle of those covered (fg) or not
 * covered (bg) by the symbol -- along the channel in which that pixel group has the
 * widest range.  Here: per-group channel extrema by warp reductions, the rank-n key by
 * a bitwise search with two ballots per bit. */

__device__ __forceinline__ uint32_t
channel_of (uint32_t px, int ch)
{
    return (px >> (8 * ch)) & 0xffu;
}

/* Channel with the widest range among the pixels selected by (s0, s1); first wins ties */
__device__ __forceinline__ int
widest_channel (const CellPix &cp, bool s0, bool s1)
{
    int best_ch = 0, best_range = -0x7fffffff;
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
    {
        uint32_t v0 = channel_of (cp.p0, ch), v1 = channel_of (cp.p1, ch);
        uint32_t mn = __reduce_min_sync (FULL, min (s0 ? v0 : 0xffffu, s1 ? v1 : 0xffffu));
        uint32_t mx = __reduce_max_sync (FULL, max (s0 ? v0 + 1 : 0u, s1 ? v1 + 1 : 0u));
        int range = (int) mx - 1 - (int) mn;
        if (range > best_range)
        {
            best_range = range;
            best_ch = ch;
        }
    }
    return best_ch;
}

/* Colour of the pixel of rank @n (0-based) among the selected ones, ordered by
 * (value of channel @ch, pixel index) */
__device__ __forceinline__ uint32_t
nth_pixel_by_channel (const CellPix &cp, bool s0, bool s1, int ch, int n, int lane)
{
    uint32_t k0 = (channel_of (cp.p0, ch) << 6) | (uint32_t) lane;
    uint32_t k1 = (channel_of (cp.p1, ch) << 6) | (uint32_t) (lane + 32);
    uint32_t result = 0;

#pragma unroll
    for (int b = 13; b >= 0; b--)
    {
        uint32_t trial = result | (1u << b);
        int below = __popc (__ballot_sync (FULL, s0 && k0 < trial)) + __popc (__ballot_sync (FULL, s1 && k1 < trial));
        if (below <= n)
            result = trial;
    }

    uint32_t idx = result & 63u;
    uint32_t a = __shfl_sync (FULL, cp.p0, idx & 31), b = __shfl_sync (FULL, cp.p1, idx & 31);
    return (idx & 32u) ? b : a;
}

__device__ void
median_colors_for_symbol (const CellPix &cp, uint64_t bm, int lane, uint32_t &fg_out, uint32_t &bg_out)
{
    bool b0, b1;
    lane_bits (bm, lane, b0, b1);
    int n_fg = __popcll (bm);

    if (n_fg == 0)
    {
        int ch = widest_channel (cp, true, true);
        fg_out = bg_out = nth_pixel_by_channel (cp, true, true, ch, 64 / 2, lane);
    }
    else if (n_fg == 64)
    {
        int ch = widest_channel (cp, true, true);
        fg_out = bg_out = nth_pixel_by_channel (cp, true, true, ch, 64 / 2, lane);
    }
    else
    {
        int fg_ch = widest_channel (cp, b0, b1);
        int bg_ch = widest_channel (cp, !b0, !b1);
        fg_out = nth_pixel_by_channel (cp, b0, b1, fg_ch, n_fg / 2, lane);
        bg_out = nth_pixel_by_channel (cp, !b0, !b1, bg_ch, (64 - n_fg) / 2, lane);
    }
}

/* chafa-symbol-renderer.c:71-78.  The extractor is a template parameter of the kernels so
 * that the (rare, register-hungry) median code does not weigh on the default path. */
template <bool MEDIAN>
__device__ __forceinline__ void
symbol_colors (const CellPix &cp, uint64_t bm, int lane, uint32_t &fg_out, uint32_t &bg_out)
{
    if (MEDIAN)
        median_colors_for_symbol (cp, bm, lane, fg_out, bg_out);
    else
        mean_colors_for_symbol (cp, bm, lane, fg_out, bg_out);
}

/* Sum over the cell of the 4-channel squared distance to the fg or bg colour
 * (chafa-avx2.c:39-81) */
__device__ __forceinline__ int
cell_error (const CellPix &cp, uint64_t bm, int lane, uint32_t fg, uint32_t bg)
{
    bool b0, b1;
    lane_bits (bm, lane, b0, b1);

    uint32_t d0 = __vabsdiffu4 (b0 ? fg : bg, cp.p0);
    uint32_t d1 = __vabsdiffu4 (b1 ? fg : bg, cp.p1);
    unsigned e = __dp4a (d0, d0, 0u);
    e = __dp4a (d1, d1, e);
    return (int) __reduce_add_sync (FULL, e);
}

/* Darkest / brightest pixel along the channel with the widest range: first minimum,
 * last maximum (chafa-work-cell.c:293-338).  colors [0] = bg, [1] = fg. */
__device__ __forceinline__ void
contrasting_pair (const CellPix &cp, int lane, uint32_t &bg_out, uint32_t &fg_out)
{
    uint32_t i0 = (uint32_t) lane, i1 = (uint32_t) lane + 32u;
    uint32_t kmin [4], kmax [4];

    /* key = value << 6 | pixel index: the smallest key is the first minimum, the
     * largest key the last maximum.  One warp reduction per channel and extreme. */
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
    {
        uint32_t a = (((cp.p0 >> (8 * ch)) & 0xffu) << 6) | i0;
        uint32_t b = (((cp.p1 >> (8 * ch)) & 0xffu) << 6) | i1;
        kmin [ch] = __reduce_min_sync (FULL, min (a, b));
        kmax [ch] = __reduce_max_sync (FULL, max (a, b));
    }

    int best_ch = 0;
    int best_range = (int) (kmax [0] >> 6) - (int) (kmin [0] >> 6);
#pragma unroll
    for (int ch = 1; ch < 4; ch++)
    {
        int range = (int) (kmax [ch] >> 6) - (int) (kmin [ch] >> 6);
        if (range > best_range)
        {
            best_range = range;
            best_ch = ch;
        }
    }

    uint32_t imin = kmin [0] & 63u, imax = kmax [0] & 63u;
    if (best_ch == 1) { imin = kmin [1] & 63u; imax = kmax [1] & 63u; }
    else if (best_ch == 2) { imin = kmin [2] & 63u; imax = kmax [2] & 63u; }
    else if (best_ch == 3) { imin = kmin [3] & 63u; imax = kmax [3] & 63u; }

    uint32_t a0 = __shfl_sync (FULL, cp.p0, imin & 31), a1 = __shfl_sync (FULL, cp.p1, imin & 31);
    uint32_t b0 = __shfl_sync (FULL, cp.p0, imax & 31), b1 = __shfl_sync (FULL, cp.p1, imax & 31);
    bg_out = (imin & 32u) ? a1 : a0;
    fg_out = (imax & 32u) ? b1 : b0;
}

/* bit = pixel is closer to colors [1] (fg) than to colors [0] (bg); MSB = pixel 0
 * (chafa-work-cell.c:121-143) */
__device__ __forceinline__ uint64_t
cell_to_bitmap (const CellPix &cp, uint32_t bg, uint32_t fg)
{
    bool s0 = color_diff3 (cp.p0, bg) > color_diff3 (cp.p0, fg);
    bool s1 = color_diff3 (cp.p1, bg) > color_diff3 (cp.p1, fg);
    uint32_t hi = __brev (__ballot_sync (FULL, s0));
    uint32_t lo = __brev (__ballot_sync (FULL, s1));
    return ((uint64_t) hi << 32) | lo;
}

/* chafa/internal/chafa-color.h:96-106 */
__device__ __forceinline__ uint32_t
color_average_2 (uint32_t a, uint32_t b)
{
    return ((a >> 1) & 0x7f7f7f7fu) + ((b >> 1) & 0x7f7f7f7fu);
}

/* ------------------------------------------------------------------------ */
/* Candidate search: stable top-K of the sequence                             */
/*   (sym 0, plain), (sym 0, inverted), (sym 1, plain), ...                   */
/* ordered by distance.  key = distance << 23 | sequence number is unique, so */
/* the K smallest keys are the answer.  WIDE: distance over both halves.      */

template <bool WIDE>
__device__ __forceinline__ int
hamming (const uint64_t *s_bitmaps, int s, uint64_t bm_a, uint64_t bm_b)
{
    if (WIDE)
        return __popcll (s_bitmaps [2 * s] ^ bm_a) + __popcll (s_bitmaps [2 * s + 1] ^ bm_b);
    return __popcll (s_bitmaps [s] ^ bm_a);
}

/* One pass over the symbols computes every distance once.  The distances stay in
 * shared memory as bytes (4 per lane and group of 128 symbols; 0xff past the end of the
 * table), and each lane keeps the smallest plain-or-inverted distance of each of its 4
 * byte positions: 128 partitions of the symbol table.  The k-th smallest partition
 * minimum bounds the k-th smallest distance overall from above, and is nearly always
 * equal to it because the best k symbols rarely share a partition.  A second sweep over
 * the packed bytes collects every (symbol, orientation) at or below the bound -- a
 * handful -- and the short list is ranked by (distance, evaluation order). */
template <bool WIDE>
__device__ int
find_candidates (const uint64_t *s_bitmaps, int n_syms, uint64_t bm_a, uint64_t bm_b,
                 bool do_inverse, int k_max, uint32_t *s_list, uint32_t *s_hd, int lane)
{
    const uint32_t max_dist = WIDE ? 128 : 64;
    const int n_entries = do_inverse ? n_syms * 2 : n_syms;
    const int k = min (k_max, n_entries);
    const int n_groups = (n_syms + 127) >> 7;
    const int n_full = n_syms >> 7;
    uint32_t mn [4] = { 0xffu, 0xffu, 0xffu, 0xffu }, mx [4] = { 0, 0, 0, 0 };

    for (int g = 0; g < n_full; g++)
    {
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            uint32_t hd = (uint32_t) hamming<WIDE> (s_bitmaps, g * 128 + 32 * j + lane, bm_a, bm_b);
            mn [j] = min (mn [j], hd);
            mx [j] = max (mx [j], hd);
            packed |= hd << (8 * j);
        }
        s_hd [g * 32 + lane] = packed;
    }
    if (n_full < n_groups)
    {
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            int s = n_full * 128 + 32 * j + lane;
            uint32_t hd = 0xffu;
            if (s < n_syms)
            {
                hd = (uint32_t) hamming<WIDE> (s_bitmaps, s, bm_a, bm_b);
                mn [j] = min (mn [j], hd);
                mx [j] = max (mx [j], hd);
            }
            packed |= hd << (8 * j);
        }
        s_hd [n_full * 32 + lane] = packed;
    }

    /* Partition minima; a partition without symbols keeps 0xff */
    if (do_inverse)
    {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (lane + 32 * j < n_syms)
                mn [j] = min (mn [j], max_dist - mx [j]);
    }

    uint32_t threshold = 0xffu;
    for (int r = 0; r < k; r++)
    {
        uint32_t lane_min = min (min (mn [0], mn [1]), min (mn [2], mn [3]));
        threshold = __reduce_min_sync (FULL, lane_min);
        uint32_t holders = __ballot_sync (FULL, lane_min == threshold);
        /* retire one partition per round: the first matching one of the lowest holder */
        bool retire = lane == __ffs (holders) - 1;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            bool is = retire && mn [j] == threshold;
            mn [j] = is ? 0xffu : mn [j];
            retire = retire && !is;
        }
    }
    if (threshold > max_dist)
        threshold = max_dist;   /* fewer than k partitions hold symbols: take everything */

    /* Collection.  Lanes append their own hits; the order does not matter.  Bytes are
     * compared two at a time in 16-bit fields: bit 15 of (0x8000 | a) - b is a >= b. */
    uint32_t *s_count = s_list + CAND_LIST_MAX;
    if (lane == 0)
        *s_count = 0;
    __syncwarp ();

    const uint32_t t2 = (0x8000u | threshold) * 0x00010001u;
    const uint32_t u2 = (max_dist - threshold) * 0x00010001u;

    for (int g = 0; g < n_groups; g++)
    {
        uint32_t p = s_hd [g * 32 + lane];
        uint32_t lo = p & 0x00ff00ffu, hi = (p >> 8) & 0x00ff00ffu;
        uint32_t hit_lo = t2 - lo, hit_hi = t2 - hi;
        if (do_inverse)
        {
            hit_lo |= (lo | 0x80008000u) - u2;
            hit_hi |= (hi | 0x80008000u) - u2;
        }
        /* bit 8 j <-> byte j */
        uint32_t hit = ((hit_lo & 0x80008000u) >> 15) | ((hit_hi & 0x80008000u) >> 7);

        while (hit)
        {
            int j = (__ffs (hit) - 1) >> 3;
            uint32_t hd = (p >> (8 * j)) & 0xffu;
            uint32_t s = (uint32_t) (g * 128 + 32 * j + lane);
            hit &= hit - 1;

            if (hd == 0xffu)
                continue;                           /* past the end of the table */
            if (hd <= threshold)
            {
                uint32_t slot = atomicAdd (s_count, 1u);
                if (slot < CAND_LIST_MAX) s_list [slot] = (hd << CAND_KEY_SHIFT) | (2 * s);
            }
            if (do_inverse && max_dist - hd <= threshold)
            {
                uint32_t slot = atomicAdd (s_count, 1u);
                if (slot < CAND_LIST_MAX) s_list [slot] = ((max_dist - hd) << CAND_KEY_SHIFT) | (2 * s + 1);
            }
        }
    }
    __syncwarp ();
    const int n_list = (int) *s_count;

    if (n_list <= 32)
    {
        /* Rank the short list; keys are distinct */
        uint32_t key = lane < n_list ? s_list [lane] : 0xffffffffu;
        int rank = 0;
        for (int i = 0; i < n_list; i++)
            rank += __shfl_sync (FULL, key, i) < key ? 1 : 0;
        __syncwarp ();
        if (lane < n_list && rank < k)
            s_list [rank] = key;
        __syncwarp ();
        return min (k, n_list);
    }

    if (n_list <= CAND_LIST_MAX)
    {
        /* Ties at the bound: same ranking, several keys per lane */
        uint32_t keys [CAND_LIST_MAX / 32];
        int rank [CAND_LIST_MAX / 32];
#pragma unroll
        for (int j = 0; j < CAND_LIST_MAX / 32; j++)
        {
            int idx = lane + 32 * j;
            keys [j] = idx < n_list ? s_list [idx] : 0xffffffffu;
            rank [j] = 0;
        }
        for (int i = 0; i < n_list; i++)
        {
            uint32_t other = s_list [i];
#pragma unroll
            for (int j = 0; j < CAND_LIST_MAX / 32; j++)
                rank [j] += other < keys [j] ? 1 : 0;
        }
        __syncwarp ();
#pragma unroll
        for (int j = 0; j < CAND_LIST_MAX / 32; j++)
            if (lane + 32 * j < n_list && rank [j] < k)
                s_list [rank [j]] = keys [j];
        __syncwarp ();
        return k;
    }

    /* Massive ties (e.g. flat cells against a table full of equal popcounts): exact
     * selection from the packed bytes, one sweep per candidate */
    uint32_t prev = 0;
    bool have_prev = false;
    for (int r = 0; r < k; r++)
    {
        uint32_t best = 0xffffffffu;
        for (int g = 0; g < n_groups; g++)
        {
            uint32_t p = s_hd [g * 32 + lane];
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                uint32_t hd = (p >> (8 * j)) & 0xffu;
                uint32_t s = (uint32_t) (g * 128 + 32 * j + lane);
                if (hd == 0xffu)
                    continue;
                uint32_t ka = (hd << CAND_KEY_SHIFT) | (2 * s);
                if (!have_prev || ka > prev) best = min (best, ka);
                if (do_inverse)
                {
                    uint32_t kb = ((max_dist - hd) << CAND_KEY_SHIFT) | (2 * s + 1);
                    if (!have_prev || kb > prev) best = min (best, kb);
                }
            }
        }
        best = __reduce_min_sync (FULL, best);
        __syncwarp ();
        if (lane == 0) s_list [r] = best;
        prev = best;
        have_prev = true;
    }
    __syncwarp ();
    return k;
}

/* ------------------------------------------------------------------------ */
/* Cell colours (chafa-symbol-renderer.c:656-729)                             */

__device__ __forceinline__ uint32_t
transparent_cell_color (int mode)
{
    return mode == MODE_TRUECOLOR ? 0x00808080u : (uint32_t) PAL_INDEX_TRANSPARENT;
}

__device__ inline void
update_cell_colors (const DevCanvas &cv, uint32_t &c_inout, uint32_t col_fg, uint32_t col_bg,
                    uint32_t &fg_out, uint32_t &bg_out, int lane)
{
    int mode = cv.canvas_mode;
    int cs = cv.color_space;

    if (mode == MODE_TRUECOLOR)
    {
        fg_out = pack_argb (col_fg);
        bg_out = pack_argb (col_bg);
    }
    else if (mode == MODE_INDEXED_16_8)
    {
        uint32_t fg = (uint32_t) palette_lookup_nearest_warp (cv.fg_palette, cs, col_fg, lane);
        uint32_t bg = (uint32_t) palette_lookup_nearest_warp (cv.fg_palette, cs, col_bg, lane);

        if (fg == bg && fg >= 8 && fg <= 15)
        {
            if (cv.solid_char)
            {
                c_inout = cv.solid_char;
                bg = (uint32_t) palette_lookup_nearest_warp (cv.bg_palette, cs, col_fg, lane);
            }
            else
            {
                fg = bg = (uint32_t) palette_lookup_nearest_warp (cv.bg_palette, cs, col_fg, lane);
            }
        }
        else
        {
            bg = (uint32_t) palette_lookup_nearest_warp (cv.bg_palette, cs, col_bg, lane);
        }
        fg_out = fg;
        bg_out = bg;
    }
    else
    {
        fg_out = (uint32_t) palette_lookup_nearest_warp (cv.fg_palette, cs, col_fg, lane);
        bg_out = (uint32_t) palette_lookup_nearest_warp (cv.bg_palette, cs, col_bg, lane);
    }

    if (cv.fg_only)
        bg_out = transparent_cell_color (mode);
}

/* Colour pair used to score a symbol: the extracted means, or the nearest pens of
 * both palettes when scoring against quantised colours (16/8 mode)
 * (chafa-symbol-renderer.c:115-147) */
__device__ __forceinline__ void
error_colors (const DevCanvas &cv, uint32_t &fg, uint32_t &bg, int lane)
{
    if (cv.use_quantized_error)
    {
        int fi = palette_lookup_nearest_warp (cv.fg_palette, cv.color_space, fg, lane);
        int bi = palette_lookup_nearest_warp (cv.bg_palette, cv.color_space, bg, lane);
        fg = cv.fg_palette->colors [fi];
        bg = cv.bg_palette->colors [bi];
    }
}

struct SubsetTable
{
    uint32_t rb [16] [16];   /* [group of 4 pixels] [subset] */
    uint32_t ga [16] [16];
    uint32_t q [16] [16];
};

struct MaskedSums
{
    uint32_t rb, ga, q;
};

__device__ __forceinline__ MaskedSums
masked_sums (const SubsetTable &t, uint64_t bm)
{
    MaskedSums m = { 0, 0, 0 };
    uint32_t hi = (uint32_t) (bm >> 32), lo = (uint32_t) bm;
#pragma unroll
    for (int g = 0; g < 8; g++)
    {
        uint32_t nib = (hi >> (28 - 4 * g)) & 15u;
        m.rb += t.rb [g] [nib];
        m.ga += t.ga [g] [nib];
        m.q += t.q [g] [nib];
    }
#pragma unroll
    for (int g = 0; g < 8; g++)
    {
        uint32_t nib = (lo >> (28 - 4 * g)) & 15u;
        m.rb += t.rb [8 + g] [nib];
        m.ga += t.ga [8 + g] [nib];
        m.q += t.q [8 + g] [nib];
    }
    return m;
}

/* sum over the pixels behind (rb, ga, q, n) of the squared 4-channel distance to @c */
__device__ __forceinline__ int
closed_form_error (uint32_t c, uint32_t rb, uint32_t ga, uint32_t q, int n)
{
    int c0 = (int) (c & 0xff), c1 = (int) ((c >> 8) & 0xff), c2 = (int) ((c >> 16) & 0xff), c3 = (int) (c >> 24);
    int dot = c0 * (int) (rb & 0xffffu) + c2 * (int) (rb >> 16) + c1 * (int) (ga & 0xffffu) + c3 * (int) (ga >> 16);
    int norm = c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3;
    return (int) q - 2 * dot + n * norm;
}

/* Warp-level build of one cell's subset tables: the cell's pixels go through a 256-byte
 * staging area so that lane (g, half) can fetch the four pixels of group g with one
 * 128-bit load; it then fills the eight subsets whose first pixel is in (half = 1) or
 * out (half = 0), each from the previous one with one addition per field. */
__device__ __forceinline__ void
build_subset_table_warp (SubsetTable &t, const CellPix &cp, uint32_t *s_px, int lane)
{
    s_px [lane] = cp.p0;
    s_px [lane + 32] = cp.p1;
    __syncwarp ();

    const int g = lane >> 1, half = lane & 1;
    const uint4 px = *(const uint4 *) (s_px + 4 * g);
    const uint32_t v [4] = { px.x, px.y, px.z, px.w };
    uint32_t rb [4], ga [4], q [4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        rb [k] = v [k] & 0x00ff00ffu;
        ga [k] = (v [k] >> 8) & 0x00ff00ffu;
        q [k] = __dp4a (v [k], v [k], 0u);
    }

    uint32_t e_rb [8], e_ga [8], e_q [8];
    e_rb [0] = half ? rb [0] : 0u;
    e_ga [0] = half ? ga [0] : 0u;
    e_q [0] = half ? q [0] : 0u;
    /* subset bit 2 <-> pixel 1, bit 1 <-> pixel 2, bit 0 <-> pixel 3 */
#pragma unroll
    for (int sub = 1; sub < 8; sub++)
    {
        int low = sub & -sub;                   /* lowest set bit */
        int px_i = low == 1 ? 3 : low == 2 ? 2 : 1;
        e_rb [sub] = e_rb [sub ^ low] + rb [px_i];
        e_ga [sub] = e_ga [sub ^ low] + ga [px_i];
        e_q [sub] = e_q [sub ^ low] + q [px_i];
    }
#pragma unroll
    for (int sub = 0; sub < 8; sub++)
    {
        t.rb [g] [half * 8 + sub] = e_rb [sub];
        t.ga [g] [half * 8 + sub] = e_ga [sub];
        t.q [g] [half * 8 + sub] = e_q [sub];
    }
    __syncwarp ();
}

/* ------------------------------------------------------------------------ */
/* K3: narrow symbols                                                         */

/* TABLE: candidates are scored side by side, one lane each, from the cell's subset tables
 * with the closed-form error (see K3x below) instead of one warp-wide pass per candidate.
 * Applies to the default configuration (average extractor, colours from the image, error
 * against the extracted colours); the other variants keep the per-candidate loop. */
template <bool MEDIAN, bool TABLE>
__global__ void __launch_bounds__ (WARPS_PER_BLOCK * 32, MEDIAN ? 1 : 4)
match_cells_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevCellEval *__restrict__ evals,
                    int bitmaps_in_global)
{
    extern __shared__ uint64_t s_mem [];
    /* symbol tables too large for shared memory (whole fonts imported as glyphs) are read in place */
    const uint64_t *s_bitmaps = bitmaps_in_global ? cv.syms.bitmaps : s_mem;
    uint32_t *s_lists = (uint32_t *) (s_mem + (bitmaps_in_global ? 0 : cv.syms.n));
    uint32_t *s_hds = s_lists + WARPS_PER_BLOCK * CAND_LIST_STRIDE;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_syms = cv.syms.n;
    uint32_t *s_hd = s_hds + warp * (((n_syms + 127) >> 7) * 32);
    /* TABLE: per warp one SubsetTable + 64 staged pixels, 16-byte aligned, after the distance bytes */
    uint8_t *s_tab_base = (uint8_t *) (((uintptr_t) (s_hds + WARPS_PER_BLOCK * (((n_syms + 127) >> 7) * 32)) + 15)
                                      & ~(uintptr_t) 15);
    SubsetTable *s_tab = (SubsetTable *) (s_tab_base + (size_t) warp * (sizeof (SubsetTable) + 256));
    uint32_t *s_px = (uint32_t *) ((uint8_t *) s_tab + sizeof (SubsetTable));

    if (!bitmaps_in_global)
        for (int i = threadIdx.x; i < n_syms; i += blockDim.x)
            s_mem [i] = cv.syms.bitmaps [i];
    __syncthreads ();

    uint32_t *s_list = s_lists + warp * CAND_LIST_STRIDE;
    const int n_cells = cv.width * cv.n_rows;
    const bool exhaustive = cv.work_factor_int >= 8;

    for (int cell = blockIdx.x * WARPS_PER_BLOCK + warp; cell < n_cells; cell += gridDim.x * WARPS_PER_BLOCK)
    {
        int cy = cv.first_row + cell / cv.width;
        int cx = cell % cv.width;
        CellPix cp;

        load_cell (cp, pixels, cv.width_px, cx, cy, lane);

        DevCellEval out;
        out.c = ' ';
        out.fg = out.bg = 0;
        out.error = SYMBOL_ERROR_MAX;
        out.mean = cell_mean_color (cp);

        if (n_syms > 0)
        {
            int n_cand;

            if (exhaustive)
            {
                n_cand = n_syms;
            }
            else
            {
                uint32_t pair_bg, pair_fg;

                if (cv.extract_colors && !cv.fg_only)
                    contrasting_pair (cp, lane, pair_bg, pair_fg);
                else
                {
                    pair_bg = cv.default_bg;
                    pair_fg = cv.default_fg;
                }

                uint64_t bitmap = cell_to_bitmap (cp, pair_bg, pair_fg);
                n_cand = find_candidates<false> (s_bitmaps, n_syms, bitmap, 0, cv.consider_inverted != 0,
                                                 cv.n_candidates, s_list, s_hd, lane);
            }

            int best_sym = -1, best_error = SYMBOL_ERROR_MAX;
            uint32_t best_fg = 0, best_bg = 0;

            if (TABLE && !exhaustive)
            {
                /* lane i scores candidate i; the first smallest error wins */
                build_subset_table_warp (*s_tab, cp, s_px, lane);

                uint32_t key = 0xffffffffu, fg = 0, bg = 0;
                int sym = 0;
                if (lane < n_cand)
                {
                    sym = (int) ((s_list [lane] & CAND_KEY_SEQ_MASK) >> 1);
                    uint64_t bm = s_bitmaps [sym];
                    MaskedSums m = masked_sums (*s_tab, bm);
                    int n_fg = __popcll (bm);
                    fg = sums_to_color (m.rb, m.ga, n_fg);
                    bg = sums_to_color (cp.tot_rb - m.rb, cp.tot_ga - m.ga, 64 - n_fg);
                    uint32_t tot_q = 0;
#pragma unroll
                    for (int g = 0; g < 16; g++)
                        tot_q += s_tab->q [g] [15];
                    int error = closed_form_error (fg, m.rb, m.ga, m.q, n_fg)
                        + closed_form_error (bg, cp.tot_rb - m.rb, cp.tot_ga - m.ga, tot_q - m.q, 64 - n_fg);
                    if (error < SYMBOL_ERROR_MAX)
                        key = ((uint32_t) error << 3) | (uint32_t) lane;
                }
                uint32_t win = __reduce_min_sync (FULL, key);
                if (win != 0xffffffffu)
                {
                    int src = (int) (win & 7u);
                    best_error = (int) (win >> 3);
                    best_sym = __shfl_sync (FULL, sym, src);
                    best_fg = __shfl_sync (FULL, fg, src);
                    best_bg = __shfl_sync (FULL, bg, src);
                }
                n_cand = 0;
            }

            for (int i = 0; i < n_cand; i++)
            {
                int sym = exhaustive ? i : (int) ((s_list [i] & CAND_KEY_SEQ_MASK) >> 1);
                uint64_t bm = s_bitmaps [sym];
                uint32_t fg, bg;

                if (cv.fg_only)
                {
                    fg = cv.default_fg;
                    bg = cv.default_bg;
                }
                else
                {
                    symbol_colors<MEDIAN> (cp, bm, lane, fg, bg);
                }

                uint32_t efg = fg, ebg = bg;
                error_colors (cv, efg, ebg, lane);
                int error = cell_error (cp, bm, lane, efg, ebg);

                if (error < best_error)
                {
                    best_error = error;
                    best_sym = sym;
                    best_fg = fg;
                    best_bg = bg;
                }
            }
            __syncwarp ();

            if (best_sym >= 0)
            {
                if (cv.extract_colors && cv.fg_only)
                    symbol_colors<MEDIAN> (cp, s_bitmaps [best_sym], lane, best_fg, best_bg);

                out.c = cv.syms.chars [best_sym];
                out.error = best_error;
                update_cell_colors (cv, out.c, best_fg, best_bg, out.fg, out.bg, lane);
            }
        }

        if (lane == 0)
            evals [(size_t) (cy - cv.first_row) * cv.width + cx] = out;
    }
}

/* ------------------------------------------------------------------------ */
/* K3w: wide symbols over the pair (cx - 1, cx); result stored at cx          */

template <bool MEDIAN, bool TABLE>
__global__ void __launch_bounds__ (WARPS_PER_BLOCK * 32, MEDIAN ? 1 : (TABLE ? 3 : 4))
match_wide_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevWideEval *__restrict__ wide_evals,
                   int bitmaps_in_global)
{
    extern __shared__ uint64_t s_mem [];
    const uint64_t *s_bitmaps = bitmaps_in_global ? cv.syms.bitmaps2 : s_mem;
    uint32_t *s_lists = (uint32_t *) (s_mem + (bitmaps_in_global ? 0 : 2 * cv.syms.n2));
    uint32_t *s_hds = s_lists + WARPS_PER_BLOCK * CAND_LIST_STRIDE;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_syms = cv.syms.n2;
    uint32_t *s_hd = s_hds + warp * (((n_syms + 127) >> 7) * 32);
    uint8_t *s_tab_base = (uint8_t *) (((uintptr_t) (s_hds + WARPS_PER_BLOCK * (((n_syms + 127) >> 7) * 32)) + 15)
                                      & ~(uintptr_t) 15);
    SubsetTable *s_tab = (SubsetTable *) (s_tab_base + (size_t) warp * (2 * sizeof (SubsetTable) + 256));
    uint32_t *s_px = (uint32_t *) ((uint8_t *) s_tab + 2 * sizeof (SubsetTable));

    if (!bitmaps_in_global)
        for (int i = threadIdx.x; i < 2 * n_syms; i += blockDim.x)
            s_mem [i] = cv.syms.bitmaps2 [i];
    __syncthreads ();

    uint32_t *s_list = s_lists + warp * CAND_LIST_STRIDE;
    const int pairs_per_row = cv.width - 1;
    const int n_pairs = pairs_per_row * cv.n_rows;
    const bool exhaustive = cv.work_factor_int >= 8;

    for (int pair = blockIdx.x * WARPS_PER_BLOCK + warp; pair < n_pairs; pair += gridDim.x * WARPS_PER_BLOCK)
    {
        int cy = cv.first_row + pair / pairs_per_row;
        int cx = pair % pairs_per_row + 1;
        CellPix ca, cb;

        load_cell (ca, pixels, cv.width_px, cx - 1, cy, lane);
        load_cell (cb, pixels, cv.width_px, cx, cy, lane);

        int n_cand;

        if (exhaustive)
        {
            n_cand = n_syms;
        }
        else
        {
            uint32_t pair_bg, pair_fg;

            if (cv.canvas_mode == MODE_FGBG || cv.canvas_mode == MODE_FGBG_BGFG)
            {
                pair_bg = cv.default_bg;
                pair_fg = cv.default_fg;
            }
            else
            {
                uint32_t abg, afg, bbg, bfg;
                contrasting_pair (ca, lane, abg, afg);
                contrasting_pair (cb, lane, bbg, bfg);
                pair_bg = color_average_2 (abg, bbg);
                pair_fg = color_average_2 (afg, bfg);
            }

            uint64_t bm_a = cell_to_bitmap (ca, pair_bg, pair_fg);
            uint64_t bm_b = cell_to_bitmap (cb, pair_bg, pair_fg);
            n_cand = find_candidates<true> (s_bitmaps, n_syms, bm_a, bm_b, cv.consider_inverted != 0,
                                            cv.n_candidates, s_list, s_hd, lane);
        }

        int best_sym = -1, best_ea = SYMBOL_ERROR_MAX, best_eb = SYMBOL_ERROR_MAX;
        uint32_t best_fg = 0, best_bg = 0;

        if (TABLE && !exhaustive)
        {
            build_subset_table_warp (s_tab [0], ca, s_px, lane);
            build_subset_table_warp (s_tab [1], cb, s_px, lane);

            uint32_t key = 0xffffffffu, fg = 0, bg = 0;
            int sym = 0, ea = 0, eb = 0;
            if (lane < n_cand)
            {
                sym = (int) ((s_list [lane] & CAND_KEY_SEQ_MASK) >> 1);
                uint64_t bm_a = s_bitmaps [2 * sym], bm_b = s_bitmaps [2 * sym + 1];
                MaskedSums ma = masked_sums (s_tab [0], bm_a), mb = masked_sums (s_tab [1], bm_b);
                int na = __popcll (bm_a), nb = __popcll (bm_b);
                uint32_t qa = 0, qb = 0;
#pragma unroll
                for (int g = 0; g < 16; g++)
                {
                    qa += s_tab [0].q [g] [15];
                    qb += s_tab [1].q [g] [15];
                }
                fg = color_average_2 (sums_to_color (ma.rb, ma.ga, na), sums_to_color (mb.rb, mb.ga, nb));
                bg = color_average_2 (sums_to_color (ca.tot_rb - ma.rb, ca.tot_ga - ma.ga, 64 - na),
                                      sums_to_color (cb.tot_rb - mb.rb, cb.tot_ga - mb.ga, 64 - nb));
                ea = closed_form_error (fg, ma.rb, ma.ga, ma.q, na)
                    + closed_form_error (bg, ca.tot_rb - ma.rb, ca.tot_ga - ma.ga, qa - ma.q, 64 - na);
                eb = closed_form_error (fg, mb.rb, mb.ga, mb.q, nb)
                    + closed_form_error (bg, cb.tot_rb - mb.rb, cb.tot_ga - mb.ga, qb - mb.q, 64 - nb);
                if (ea + eb < 2 * SYMBOL_ERROR_MAX)
                    key = ((uint32_t) (ea + eb) << 3) | (uint32_t) lane;
            }
            uint32_t win = __reduce_min_sync (FULL, key);
            if (win != 0xffffffffu)
            {
                int src = (int) (win & 7u);
                best_sym = __shfl_sync (FULL, sym, src);
                best_fg = __shfl_sync (FULL, fg, src);
                best_bg = __shfl_sync (FULL, bg, src);
                best_ea = __shfl_sync (FULL, ea, src);
                best_eb = __shfl_sync (FULL, eb, src);
            }
            n_cand = 0;
        }

        for (int i = 0; i < n_cand; i++)
        {
            int sym = exhaustive ? i : (int) ((s_list [i] & CAND_KEY_SEQ_MASK) >> 1);
            uint64_t bm_a = s_bitmaps [2 * sym], bm_b = s_bitmaps [2 * sym + 1];
            uint32_t fg, bg;

            if (cv.fg_only)
            {
                fg = cv.default_fg;
                bg = cv.default_bg;
            }
            else
            {
                uint32_t afg, abg, bfg, bbg;
                symbol_colors<MEDIAN> (ca, bm_a, lane, afg, abg);
                symbol_colors<MEDIAN> (cb, bm_b, lane, bfg, bbg);
                fg = color_average_2 (afg, bfg);
                bg = color_average_2 (abg, bbg);
            }

            uint32_t efg = fg, ebg = bg;
            error_colors (cv, efg, ebg, lane);
            int ea = cell_error (ca, bm_a, lane, efg, ebg);
            int eb = cell_error (cb, bm_b, lane, efg, ebg);

            if (ea + eb < best_ea + best_eb)
            {
                best_ea = ea;
                best_eb = eb;
                best_sym = sym;
                best_fg = fg;
                best_bg = bg;
            }
        }
        __syncwarp ();

        DevWideEval out;
        out.c = 0;
        out.fg = out.bg = 0;
        out.error_a = out.error_b = SYMBOL_ERROR_MAX;

        if (best_sym >= 0)
        {
            if (cv.extract_colors && cv.fg_only)
            {
                uint32_t afg, abg, bfg, bbg;
                symbol_colors<MEDIAN> (ca, s_bitmaps [2 * best_sym], lane, afg, abg);
                symbol_colors<MEDIAN> (cb, s_bitmaps [2 * best_sym + 1], lane, bfg, bbg);
                best_fg = color_average_2 (afg, bfg);
                best_bg = color_average_2 (abg, bbg);
            }

            out.c = cv.syms.chars2 [best_sym];
            out.error_a = best_ea;
            out.error_b = best_eb;
            update_cell_colors (cv, out.c, best_fg, best_bg, out.fg, out.bg, lane);
        }

        if (lane == 0)
            wide_evals [(size_t) (cy - cv.first_row) * cv.width + cx] = out;
    }
}

/* ------------------------------------------------------------------------ */
/* K3x: exhaustive matching (work factor >= 0.8): every symbol is scored for every     */
/* cell (chafa-symbol-renderer.c:270-339).                                              */
/*                                                                                      */
/* One THREAD per (cell, symbol) instead of one warp: the per-symbol quantities are     */
/* masked sums over the cell's 64 pixels, and a masked sum over 64 pixels is the sum    */
/* of 16 table entries if the cell first tabulates, for each group of 4 pixels, the     */
/* sums of all 16 subsets (3 KB of shared memory per cell, built by 256 threads in one  */
/* step).  Tabulated per subset: the channel sums (two 16-bit fields per word) and the  */
/* sum of squared pixel norms.  With S = masked channel sums, Q = masked sum of squares */
/* and c = the (integer) mean colour the reference computes from S, the reference's     */
/* error  sum_i sum_ch (c_ch - px_ch)^2  is exactly  Q - 2 c.S + n |c|^2  in integers,  */
/* for the covered and the uncovered pixels separately -- no second pass over pixels.   */
/* A block walks a run of cells of one row, keeps the tables of the previous cell, and  */
/* scores the wide symbols on each adjacent pair from the two tables.                   */

#define XH_THREADS 256
#define XH_SEGMENT 32

__global__ void __launch_bounds__ (XH_THREADS)
match_exhaustive_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevCellEval *__restrict__ evals,
                         DevWideEval *__restrict__ wide_evals, int bitmaps_in_global)
{
    extern __shared__ uint64_t s_mem [];
    const int n_wide_tab = wide_evals ? cv.syms.n2 : 0;
    const uint64_t *s_bitmaps = bitmaps_in_global ? cv.syms.bitmaps : s_mem;                    /* n */
    const uint64_t *s_bitmaps2 = bitmaps_in_global ? cv.syms.bitmaps2 : s_mem + cv.syms.n;      /* 2 * n2 */
    SubsetTable *s_tab = (SubsetTable *) (s_mem + (bitmaps_in_global ? 0 : cv.syms.n + 2 * n_wide_tab));   /* 2 tables */
    __shared__ unsigned long long s_best, s_best2;
    __shared__ int s_err_a [XH_THREADS / 32], s_err_b [XH_THREADS / 32];
    __shared__ unsigned long long s_key2 [XH_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_syms = cv.syms.n, n_syms2 = wide_evals ? cv.syms.n2 : 0;
    const int segs_per_row = (cv.width + XH_SEGMENT - 1) / XH_SEGMENT;

    if (!bitmaps_in_global)
    {
        for (int i = tid; i < n_syms; i += XH_THREADS)
            s_mem [i] = cv.syms.bitmaps [i];
        for (int i = tid; i < 2 * n_syms2; i += XH_THREADS)
            s_mem [n_syms + i] = cv.syms.bitmaps2 [i];
    }

    for (int seg = blockIdx.x; seg < cv.n_rows * segs_per_row; seg += gridDim.x)
    {
        const int r = seg / segs_per_row;
        const int cy = cv.first_row + r;
        const int cx_begin = (seg % segs_per_row) * XH_SEGMENT;
        const int cx_end = min (cx_begin + XH_SEGMENT, cv.width);
        /* one cell of lead-in so the first pair of the segment has its left tables */
        const int cx_first = (n_syms2 > 0 && cx_begin > 0) ? cx_begin - 1 : cx_begin;

        for (int cx = cx_first; cx < cx_end; cx++)
        {
            SubsetTable &cur = s_tab [cx & 1];
            const SubsetTable &prev = s_tab [(cx & 1) ^ 1];

            __syncthreads ();
            {
                /* entry (group, subset) of this cell's tables; the subset's MSB is the
                 * group's first pixel, like the bitmap's */
                int g = tid >> 4, sub = tid & 15;
                const uint32_t *px = pixels + (size_t) (cy * 8 + (g >> 1)) * cv.width_px + cx * 8 + (g & 1) * 4;
                uint32_t rb = 0, ga = 0, q = 0;
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    uint32_t p = __ldg (px + k);
                    if ((sub >> (3 - k)) & 1)
                    {
                        rb += p & 0x00ff00ffu;
                        ga += (p >> 8) & 0x00ff00ffu;
                        q = __dp4a (p, p, q);
                    }
                }
                cur.rb [g] [sub] = rb;
                cur.ga [g] [sub] = ga;
                cur.q [g] [sub] = q;
                if (tid == 0)
                {
                    s_best = ~0ull;
                    s_best2 = ~0ull;
                }
            }
            __syncthreads ();

            /* whole-cell sums: the full subset of every group */
            uint32_t tot_rb = 0, tot_ga = 0, tot_q = 0;
#pragma unroll
            for (int g = 0; g < 16; g++)
            {
                tot_rb += cur.rb [g] [15];
                tot_ga += cur.ga [g] [15];
                tot_q += cur.q [g] [15];
            }

            if (cx >= cx_begin)
            {
                unsigned long long best = ~0ull;
                for (int s = tid; s < n_syms; s += XH_THREADS)
                {
                    uint64_t bm = s_bitmaps [s];
                    MaskedSums m = masked_sums (cur, bm);
                    int n_fg = __popcll (bm);
                    uint32_t fg = sums_to_color (m.rb, m.ga, n_fg);
                    uint32_t bg = sums_to_color (tot_rb - m.rb, tot_ga - m.ga, 64 - n_fg);
                    int error = closed_form_error (fg, m.rb, m.ga, m.q, n_fg)
                        + closed_form_error (bg, tot_rb - m.rb, tot_ga - m.ga, tot_q - m.q, 64 - n_fg);
                    if (error < SYMBOL_ERROR_MAX)
                        best = min (best, ((unsigned long long) (unsigned) error << 32) | (unsigned) s);
                }
                /* first symbol with the smallest error: min over (error, index) */
                unsigned hi = __reduce_min_sync (FULL, (unsigned) (best >> 32));
                unsigned lo = __reduce_min_sync (FULL, (unsigned) (best >> 32) == hi ? (unsigned) best : 0xffffffffu);
                if (lane == 0 && hi != 0xffffffffu)
                    atomicMin (&s_best, ((unsigned long long) hi << 32) | lo);
            }

            if (n_syms2 > 0 && cx >= 1 && cx >= cx_begin)
            {
                /* left-cell totals */
                uint32_t ptot_rb = 0, ptot_ga = 0, ptot_q = 0;
#pragma unroll
                for (int g = 0; g < 16; g++)
                {
                    ptot_rb += prev.rb [g] [15];
                    ptot_ga += prev.ga [g] [15];
                    ptot_q += prev.q [g] [15];
                }

                unsigned long long best = ~0ull;
                int best_ea = 0, best_eb = 0;
                for (int s = tid; s < n_syms2; s += XH_THREADS)
                {
                    uint64_t bm_a = s_bitmaps2 [2 * s], bm_b = s_bitmaps2 [2 * s + 1];
                    MaskedSums ma = masked_sums (prev, bm_a), mb = masked_sums (cur, bm_b);
                    int na = __popcll (bm_a), nb = __popcll (bm_b);
                    uint32_t fg = color_average_2 (sums_to_color (ma.rb, ma.ga, na), sums_to_color (mb.rb, mb.ga, nb));
                    uint32_t bg = color_average_2 (sums_to_color (ptot_rb - ma.rb, ptot_ga - ma.ga, 64 - na),
                                                   sums_to_color (tot_rb - mb.rb, tot_ga - mb.ga, 64 - nb));
                    int ea = closed_form_error (fg, ma.rb, ma.ga, ma.q, na)
                        + closed_form_error (bg, ptot_rb - ma.rb, ptot_ga - ma.ga, ptot_q - ma.q, 64 - na);
                    int eb = closed_form_error (fg, mb.rb, mb.ga, mb.q, nb)
                        + closed_form_error (bg, tot_rb - mb.rb, tot_ga - mb.ga, tot_q - mb.q, 64 - nb);
                    unsigned long long key = ((unsigned long long) (unsigned) (ea + eb) << 32) | (unsigned) s;
                    if (ea + eb < 2 * SYMBOL_ERROR_MAX && key < best)
                    {
                        best = key;
                        best_ea = ea;
                        best_eb = eb;
                    }
                }
                unsigned hi = __reduce_min_sync (FULL, (unsigned) (best >> 32));
                unsigned lo = __reduce_min_sync (FULL, (unsigned) (best >> 32) == hi ? (unsigned) best : 0xffffffffu);
                unsigned long long wkey = ((unsigned long long) hi << 32) | lo;
                if (best == wkey && hi != 0xffffffffu)
                {
                    /* this lane holds the warp's winner: publish its halves */
                    s_err_a [warp] = best_ea;
                    s_err_b [warp] = best_eb;
                }
                if (lane == 0)
                {
                    s_key2 [warp] = hi != 0xffffffffu ? wkey : ~0ull;
                    if (hi != 0xffffffffu)
                        atomicMin (&s_best2, wkey);
                }
            }
            __syncthreads ();

            if (cx < cx_begin)
                continue;

            if (warp == 0)
            {
                DevCellEval out;
                out.c = ' ';
                out.fg = out.bg = 0;
                out.error = SYMBOL_ERROR_MAX;
                out.mean = sums_to_color (tot_rb, tot_ga, 64);
                unsigned long long key = s_best;
                if (key != ~0ull)
                {
                    int sym = (int) (unsigned) key;
                    uint64_t bm = s_bitmaps [sym];
                    MaskedSums m = masked_sums (cur, bm);
                    int n_fg = __popcll (bm);
                    uint32_t fg = sums_to_color (m.rb, m.ga, n_fg);
                    uint32_t bg = sums_to_color (tot_rb - m.rb, tot_ga - m.ga, 64 - n_fg);
                    out.c = cv.syms.chars [sym];
                    out.error = (int) (key >> 32);
                    update_cell_colors (cv, out.c, fg, bg, out.fg, out.bg, lane);
                }
                if (lane == 0)
                    evals [(size_t) r * cv.width + cx] = out;
            }
            else if (warp == 1 && n_syms2 > 0 && cx >= 1)
            {
                DevWideEval out;
                out.c = 0;
                out.fg = out.bg = 0;
                out.error_a = out.error_b = SYMBOL_ERROR_MAX;
                unsigned long long key = s_best2;
                if (key != ~0ull)
                {
                    uint32_t ptot_rb = 0, ptot_ga = 0;
#pragma unroll
                    for (int g = 0; g < 16; g++)
                    {
                        ptot_rb += prev.rb [g] [15];
                        ptot_ga += prev.ga [g] [15];
                    }
                    int sym = (int) (unsigned) key;
                    uint64_t bm_a = s_bitmaps2 [2 * sym], bm_b = s_bitmaps2 [2 * sym + 1];
                    MaskedSums ma = masked_sums (prev, bm_a), mb = masked_sums (cur, bm_b);
                    int na = __popcll (bm_a), nb = __popcll (bm_b);
                    uint32_t fg = color_average_2 (sums_to_color (ma.rb, ma.ga, na), sums_to_color (mb.rb, mb.ga, nb));
                    uint32_t bg = color_average_2 (sums_to_color (ptot_rb - ma.rb, ptot_ga - ma.ga, 64 - na),
                                                   sums_to_color (tot_rb - mb.rb, tot_ga - mb.ga, 64 - nb));
                    for (int w = 0; w < XH_THREADS / 32; w++)
                        if (s_key2 [w] == key)
                        {
                            out.error_a = s_err_a [w];
                            out.error_b = s_err_b [w];
                        }
                    out.c = cv.syms.chars2 [sym];
                    update_cell_colors (cv, out.c, fg, bg, out.fg, out.bg, lane);
                }
                if (lane == 0)
                    wide_evals [(size_t) r * cv.width + cx] = out;
            }
        }
    }
}

/* ------------------------------------------------------------------------ */

static int g_num_sms = 0;

static int
num_sms ()
{
    if (!g_num_sms)
    {
        int dev = 0;
        cudaGetDevice (&dev);
        cudaDeviceGetAttribute (&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0)
            g_num_sms = 148;
    }
    return g_num_sms;
}


/* Shared-memory budget of the scalar matchers.  The symbol bitmaps live in shared memory while
 * they leave room for a few blocks per SM (whole fonts imported through
 * chafa_symbol_map_add_glyph do not: those tables are read in place, from L1/L2).  Adds the bitmap
 * bytes to *smem when they are to be staged, opts the kernel in to more than 48 KB on the current
 * device (the attribute is per device and per kernel; setting it is cheap), and returns whether the
 * kernel reads the bitmaps from global memory; -1 if even the per-warp scratch does not fit. */
#define MATCH_SMEM_STAGE_MAX (96 * 1024)
#define MATCH_SMEM_BLOCK_MAX (227 * 1024)

template <typename K>
static int
place_bitmaps (K kernel, size_t table_bytes, size_t *smem)
{
    int in_global = *smem + table_bytes > MATCH_SMEM_STAGE_MAX;
    if (!in_global)
        *smem += table_bytes;
    if (*smem > MATCH_SMEM_BLOCK_MAX)
        return -1;
    if (*smem > 48 * 1024
        && cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) *smem) != cudaSuccess)
    {
        cudaGetLastError ();
        return -1;
    }
    return in_global;
}

/* Side-by-side candidate scoring from subset tables: the default configuration */
static bool
match_table_applies (const DevCanvas *cv)
{
    return cv->work_factor_int < 8 && !cv->fg_only && !cv->use_quantized_error && cv->color_extractor == 0
        && cv->extract_colors;
}

/* c_invdiv16 is initialised statically, so it is present on every device the module
 * gets loaded on */
static int
upload_tables ()
{
    return 0;
}

/* The exhaustive kernel covers the common case; the quantised-error (16/8), foreground-only
 * and median variants stay on the warp-per-cell kernels. */
bool
dev_match_exhaustive_applies (const DevCanvas *cv)
{
    return cv->work_factor_int >= 8 && !cv->fg_only && !cv->use_quantized_error && cv->color_extractor == 0
        && cv->syms.n > 0;
}

int
dev_launch_match_exhaustive (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals,
                             DevWideEval *wide_evals, void *stream)
{
    int err = upload_tables ();
    int n_segs = cv->n_rows * ((cv->width + XH_SEGMENT - 1) / XH_SEGMENT);

    if (err || n_segs <= 0)
        return err;

    int n2 = wide_evals ? cv->syms.n2 : 0;
    size_t tables = (size_t) cv->syms.n * 8 + (size_t) n2 * 16;
    size_t smem = 2 * sizeof (SubsetTable);
    int in_global = place_bitmaps (match_exhaustive_kernel, tables, &smem);
    if (in_global < 0)
        return (int) cudaErrorInvalidValue;
    int grid = std::min (n_segs, num_sms () * 6);
    match_exhaustive_kernel<<<grid, XH_THREADS, smem, (cudaStream_t) stream>>> (*cv, pixels, evals, wide_evals, in_global);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_match (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, void *stream)
{
    int n_cells = cv->width * cv->n_rows;
    int err = upload_tables ();

    if (err || n_cells <= 0)
        return err;

    bool table = match_table_applies (cv);
    size_t tables = (size_t) cv->syms.n * 8;
    size_t smem = WARPS_PER_BLOCK * CAND_LIST_STRIDE * 4
        + (size_t) WARPS_PER_BLOCK * ((cv->syms.n + 127) / 128) * 128
        + 16 + (table ? (size_t) WARPS_PER_BLOCK * (sizeof (SubsetTable) + 256) : 0);
    int blocks_needed = (n_cells + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int grid = std::min (blocks_needed, num_sms () * 8);
    auto kernel = table ? match_cells_kernel<false, true>
                        : cv->color_extractor == 0 ? match_cells_kernel<false, false> : match_cells_kernel<true, false>;
    int in_global = place_bitmaps (kernel, tables, &smem);
    if (in_global < 0)
        return (int) cudaErrorInvalidValue;
    kernel<<<grid, WARPS_PER_BLOCK * 32, smem, (cudaStream_t) stream>>> (*cv, pixels, evals, in_global);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_match_wide (const DevCanvas *cv, const uint32_t *pixels, DevWideEval *wide_evals, void *stream)
{
    int n_pairs = (cv->width - 1) * cv->n_rows;
    int err = upload_tables ();

    if (err || cv->syms.n2 <= 0 || n_pairs <= 0 || !wide_evals)
        return err;

    bool table = match_table_applies (cv);
    size_t tables = (size_t) cv->syms.n2 * 16;
    size_t smem = WARPS_PER_BLOCK * CAND_LIST_STRIDE * 4
        + (size_t) WARPS_PER_BLOCK * ((cv->syms.n2 + 127) / 128) * 128
        + 16 + (table ? (size_t) WARPS_PER_BLOCK * (2 * sizeof (SubsetTable) + 256) : 0);
    int blocks_needed = (n_pairs + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int grid = std::min (blocks_needed, num_sms () * 8);
    auto kernel = table ? match_wide_kernel<false, true>
                        : cv->color_extractor == 0 ? match_wide_kernel<false, false> : match_wide_kernel<true, false>;
    int in_global = place_bitmaps (kernel, tables, &smem);
    if (in_global < 0)
        return (int) cudaErrorInvalidValue;
    kernel<<<grid, WARPS_PER_BLOCK * 32, smem, (cudaStream_t) stream>>> (*cv, pixels, wide_evals, in_global);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
