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
#define CAND_LIST_MAX 96
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
__constant__ uint16_t c_invdiv16 [65];

static const uint16_t h_invdiv16 [65] =
{
    0, 32768, 16384, 10922, 8192, 6553, 5461, 4681, 4096, 3640, 3276, 2978, 2730, 2520, 2340, 2184,
    2048, 1927, 1820, 1724, 1638, 1560, 1489, 1424, 1365, 1310, 1260, 1213, 1170, 1129, 1092, 1057,
    1024, 992, 963, 936, 910, 885, 862, 840, 819, 799, 780, 762, 744, 728, 712, 697,
    682, 668, 655, 642, 630, 618, 606, 595, 585, 574, 564, 555, 546, 537, 528, 520, 512
};

struct CellPix
{
    uint32_t p0, p1;          /* this lane's two pixels */
    uint32_t tot_rb, tot_ga;  /* whole-cell channel sums, two 16-bit fields per word */
};

__device__ __forceinline__ void
load_cell (CellPix &cp, const uint32_t *pixels, int width_px, int cx, int cy, int lane)
{
    const uint32_t *base = pixels + (size_t) (cy * 8 + (lane >> 3)) * width_px + cx * 8 + (lane & 7);
    cp.p0 = __ldg (base);
    cp.p1 = __ldg (base + (size_t) 4 * width_px);

    uint32_t rb = (cp.p0 & 0x00ff00ffu) + (cp.p1 & 0x00ff00ffu);
    uint32_t ga = ((cp.p0 >> 8) & 0x00ff00ffu) + ((cp.p1 >> 8) & 0x00ff00ffu);
    cp.tot_rb = __reduce_add_sync (FULL, rb);
    cp.tot_ga = __reduce_add_sync (FULL, ga);
}

/* mulhrs (sum, floor (32768 / n)) -- chafa/internal/chafa-avx2.c:151-165 */
__device__ __forceinline__ uint32_t
div_field (uint32_t sum, int n)
{
    if (n <= 1)
        return sum;
    return (sum * (uint32_t) c_invdiv16 [n] + 16384u) >> 15;
}

__device__ __forceinline__ uint32_t
sums_to_color (uint32_t rb, uint32_t ga, int n)
{
    uint32_t r = div_field (rb & 0xffffu, n) & 0xffu;
    uint32_t b = div_field (rb >> 16, n) & 0xffu;
    uint32_t g = div_field (ga & 0xffffu, n) & 0xffu;
    uint32_t a = div_field (ga >> 16, n) & 0xffu;
    return r | (g << 8) | (b << 16) | (a << 24);
}

/* Mean colour of all 64 pixels (chafa-work-cell.c:104-119) */
__device__ __forceinline__ uint32_t
cell_mean_color (const CellPix &cp)
{
    return sums_to_color (cp.tot_rb, cp.tot_ga, 64);
}

__device__ __forceinline__ void
lane_bits (uint64_t bm, int lane, uint32_t &b0, uint32_t &b1)
{
    b0 = ((uint32_t) (bm >> 32) >> (31 - lane)) & 1u;
    b1 = ((uint32_t) bm >> (31 - lane)) & 1u;
}

/* Masked mean colours for one symbol (chafa-work-cell.c:76-102) */
__device__ __forceinline__ void
mean_colors_for_symbol (const CellPix &cp, uint64_t bm, int lane, uint32_t &fg_out, uint32_t &bg_out)
{
    uint32_t b0, b1;
    lane_bits (bm, lane, b0, b1);

    uint32_t m0 = 0u - b0, m1 = 0u - b1;
    uint32_t rb = ((cp.p0 & m0) & 0x00ff00ffu) + ((cp.p1 & m1) & 0x00ff00ffu);
    uint32_t ga = (((cp.p0 & m0) >> 8) & 0x00ff00ffu) + (((cp.p1 & m1) >> 8) & 0x00ff00ffu);
    uint32_t fg_rb = __reduce_add_sync (FULL, rb);
    uint32_t fg_ga = __reduce_add_sync (FULL, ga);
    int n_fg = __popcll (bm);

    fg_out = sums_to_color (fg_rb, fg_ga, n_fg);
    bg_out = sums_to_color (cp.tot_rb - fg_rb, cp.tot_ga - fg_ga, 64 - n_fg);
}

/* Sum over the cell of the 4-channel squared distance to the fg or bg colour
 * (chafa-avx2.c:39-81) */
__device__ __forceinline__ int
cell_error (const CellPix &cp, uint64_t bm, int lane, uint32_t fg, uint32_t bg)
{
    uint32_t b0, b1;
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
    uint32_t k0 [2], k1 [2];

    /* key = value << 6 | pixel index; two channels per word */
    k0 [0] = (((cp.p0 & 0xffu) << 6 | i0) << 16) | (((cp.p0 >> 8) & 0xffu) << 6 | i0);
    k0 [1] = ((((cp.p0 >> 16) & 0xffu) << 6 | i0) << 16) | (((cp.p0 >> 24) & 0xffu) << 6 | i0);
    k1 [0] = (((cp.p1 & 0xffu) << 6 | i1) << 16) | (((cp.p1 >> 8) & 0xffu) << 6 | i1);
    k1 [1] = ((((cp.p1 >> 16) & 0xffu) << 6 | i1) << 16) | (((cp.p1 >> 24) & 0xffu) << 6 | i1);

    uint32_t mn [2], mx [2];
    mn [0] = __vminu2 (k0 [0], k1 [0]);
    mn [1] = __vminu2 (k0 [1], k1 [1]);
    mx [0] = __vmaxu2 (k0 [0], k1 [0]);
    mx [1] = __vmaxu2 (k0 [1], k1 [1]);

#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
    {
        mn [0] = __vminu2 (mn [0], __shfl_xor_sync (FULL, mn [0], s));
        mn [1] = __vminu2 (mn [1], __shfl_xor_sync (FULL, mn [1], s));
        mx [0] = __vmaxu2 (mx [0], __shfl_xor_sync (FULL, mx [0], s));
        mx [1] = __vmaxu2 (mx [1], __shfl_xor_sync (FULL, mx [1], s));
    }

    /* channel order 0..3 = hi(word0), lo(word0), hi(word1), lo(word1) */
    uint32_t kmin [4] = { mn [0] >> 16, mn [0] & 0xffffu, mn [1] >> 16, mn [1] & 0xffffu };
    uint32_t kmax [4] = { mx [0] >> 16, mx [0] & 0xffffu, mx [1] >> 16, mx [1] & 0xffffu };

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
/* ordered by distance.  key = distance << 13 | sequence number is unique, so */
/* the K smallest keys are the answer.  WIDE: distance over both halves.      */

template <bool WIDE>
__device__ __forceinline__ int
hamming (const uint64_t *s_bitmaps, int s, uint64_t bm_a, uint64_t bm_b)
{
    if (WIDE)
        return __popcll (s_bitmaps [2 * s] ^ bm_a) + __popcll (s_bitmaps [2 * s + 1] ^ bm_b);
    return __popcll (s_bitmaps [s] ^ bm_a);
}

template <bool WIDE>
__device__ int
find_candidates (const uint64_t *s_bitmaps, int n_syms, uint64_t bm_a, uint64_t bm_b,
                 bool do_inverse, int k_max, uint32_t *s_list, int lane)
{
    const int max_dist = WIDE ? 128 : 64;
    int n_entries = do_inverse ? n_syms * 2 : n_syms;
    int k = min (k_max, n_entries);

    /* Pass 1: each lane's smallest key.  The k-th smallest of those bounds the
     * k-th smallest key overall from above. */
    uint32_t lane_min = 0xffffffffu;
    for (int s = lane; s < n_syms; s += 32)
    {
        int hd = hamming<WIDE> (s_bitmaps, s, bm_a, bm_b);
        lane_min = min (lane_min, ((uint32_t) hd << 13) | (uint32_t) (2 * s));
        if (do_inverse)
            lane_min = min (lane_min, ((uint32_t) (max_dist - hd) << 13) | (uint32_t) (2 * s + 1));
    }

    uint32_t threshold = 0xffffffffu;
    if (n_syms >= 32 || n_syms >= k)
    {
        uint32_t mine = lane_min;
        for (int r = 0; r < k; r++)
        {
            threshold = __reduce_min_sync (FULL, mine);
            if (mine == threshold)
                mine = 0xffffffffu;
        }
    }

    /* Pass 2: collect every key <= threshold */
    int n_list = 0;
    for (int base = 0; base < n_syms; base += 32)
    {
        int s = base + lane;
        uint32_t ka = 0xffffffffu, kb = 0xffffffffu;
        if (s < n_syms)
        {
            int hd = hamming<WIDE> (s_bitmaps, s, bm_a, bm_b);
            ka = ((uint32_t) hd << 13) | (uint32_t) (2 * s);
            if (do_inverse)
                kb = ((uint32_t) (max_dist - hd) << 13) | (uint32_t) (2 * s + 1);
        }
        bool pa = s < n_syms && ka <= threshold;
        bool pb = s < n_syms && do_inverse && kb <= threshold;
        uint32_t ma = __ballot_sync (FULL, pa), mb = __ballot_sync (FULL, pb);
        uint32_t lt = (1u << lane) - 1u;
        if (pa)
        {
            int slot = n_list + __popc (ma & lt);
            if (slot < CAND_LIST_MAX) s_list [slot] = ka;
        }
        n_list += __popc (ma);
        if (pb)
        {
            int slot = n_list + __popc (mb & lt);
            if (slot < CAND_LIST_MAX) s_list [slot] = kb;
        }
        n_list += __popc (mb);
    }
    __syncwarp ();

    if (n_list > CAND_LIST_MAX)
    {
        /* Rare (massive ties): exact selection, one full scan per candidate */
        uint32_t prev = 0;
        bool have_prev = false;
        for (int r = 0; r < k; r++)
        {
            uint32_t best = 0xffffffffu;
            for (int s = lane; s < n_syms; s += 32)
            {
                int hd = hamming<WIDE> (s_bitmaps, s, bm_a, bm_b);
                uint32_t ka = ((uint32_t) hd << 13) | (uint32_t) (2 * s);
                if (!have_prev || ka > prev) best = min (best, ka);
                if (do_inverse)
                {
                    uint32_t kb = ((uint32_t) (max_dist - hd) << 13) | (uint32_t) (2 * s + 1);
                    if (!have_prev || kb > prev) best = min (best, kb);
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

    /* Sort the short list by rank; keys are distinct */
    uint32_t mine [CAND_LIST_MAX / 32];
    int rank [CAND_LIST_MAX / 32];
#pragma unroll
    for (int j = 0; j < CAND_LIST_MAX / 32; j++)
    {
        int idx = lane + 32 * j;
        mine [j] = idx < n_list ? s_list [idx] : 0xffffffffu;
        rank [j] = 0;
    }
    for (int i = 0; i < n_list; i++)
    {
        uint32_t other = s_list [i];
#pragma unroll
        for (int j = 0; j < CAND_LIST_MAX / 32; j++)
            rank [j] += other < mine [j] ? 1 : 0;
    }
    __syncwarp ();
#pragma unroll
    for (int j = 0; j < CAND_LIST_MAX / 32; j++)
    {
        int idx = lane + 32 * j;
        if (idx < n_list && rank [j] < k)
            s_list [rank [j]] = mine [j];
    }
    __syncwarp ();
    return min (k, n_list);
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
        uint32_t fg = (uint32_t) palette_lookup_nearest (cv.fg_palette, cs, col_fg, nullptr);
        uint32_t bg = (uint32_t) palette_lookup_nearest (cv.fg_palette, cs, col_bg, nullptr);

        if (fg == bg && fg >= 8 && fg <= 15)
        {
            if (cv.solid_char)
            {
                c_inout = cv.solid_char;
                bg = (uint32_t) palette_lookup_nearest (cv.bg_palette, cs, col_fg, nullptr);
            }
            else
            {
                fg = bg = (uint32_t) palette_lookup_nearest (cv.bg_palette, cs, col_fg, nullptr);
            }
        }
        else
        {
            bg = (uint32_t) palette_lookup_nearest (cv.bg_palette, cs, col_bg, nullptr);
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
error_colors (const DevCanvas &cv, uint32_t &fg, uint32_t &bg)
{
    if (cv.use_quantized_error)
    {
        int fi = palette_lookup_nearest (cv.fg_palette, cv.color_space, fg, nullptr);
        int bi = palette_lookup_nearest (cv.bg_palette, cv.color_space, bg, nullptr);
        fg = cv.fg_palette->colors [fi];
        bg = cv.bg_palette->colors [bi];
    }
}

/* ------------------------------------------------------------------------ */
/* K3: narrow symbols                                                         */

__global__ void __launch_bounds__ (WARPS_PER_BLOCK * 32)
match_cells_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevCellEval *__restrict__ evals)
{
    extern __shared__ uint64_t s_mem [];
    uint64_t *s_bitmaps = s_mem;
    uint32_t *s_lists = (uint32_t *) (s_bitmaps + cv.syms.n);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_syms = cv.syms.n;

    for (int i = threadIdx.x; i < n_syms; i += blockDim.x)
        s_bitmaps [i] = cv.syms.bitmaps [i];
    __syncthreads ();

    uint32_t *s_list = s_lists + warp * CAND_LIST_MAX;
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
                                                 cv.n_candidates, s_list, lane);
            }

            int best_sym = -1, best_error = SYMBOL_ERROR_MAX;
            uint32_t best_fg = 0, best_bg = 0;

            for (int i = 0; i < n_cand; i++)
            {
                int sym = exhaustive ? i : (int) ((s_list [i] & 0x1fffu) >> 1);
                uint64_t bm = s_bitmaps [sym];
                uint32_t fg, bg;

                if (cv.fg_only)
                {
                    fg = cv.default_fg;
                    bg = cv.default_bg;
                }
                else
                {
                    mean_colors_for_symbol (cp, bm, lane, fg, bg);
                }

                uint32_t efg = fg, ebg = bg;
                error_colors (cv, efg, ebg);
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
                    mean_colors_for_symbol (cp, s_bitmaps [best_sym], lane, best_fg, best_bg);

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

__global__ void __launch_bounds__ (WARPS_PER_BLOCK * 32)
match_wide_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevWideEval *__restrict__ wide_evals)
{
    extern __shared__ uint64_t s_mem [];
    uint64_t *s_bitmaps = s_mem;
    uint32_t *s_lists = (uint32_t *) (s_bitmaps + 2 * cv.syms.n2);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_syms = cv.syms.n2;

    for (int i = threadIdx.x; i < 2 * n_syms; i += blockDim.x)
        s_bitmaps [i] = cv.syms.bitmaps2 [i];
    __syncthreads ();

    uint32_t *s_list = s_lists + warp * CAND_LIST_MAX;
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
                                            cv.n_candidates, s_list, lane);
        }

        int best_sym = -1, best_ea = SYMBOL_ERROR_MAX, best_eb = SYMBOL_ERROR_MAX;
        uint32_t best_fg = 0, best_bg = 0;

        for (int i = 0; i < n_cand; i++)
        {
            int sym = exhaustive ? i : (int) ((s_list [i] & 0x1fffu) >> 1);
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
                mean_colors_for_symbol (ca, bm_a, lane, afg, abg);
                mean_colors_for_symbol (cb, bm_b, lane, bfg, bbg);
                fg = color_average_2 (afg, bfg);
                bg = color_average_2 (abg, bbg);
            }

            uint32_t efg = fg, ebg = bg;
            error_colors (cv, efg, ebg);
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
                mean_colors_for_symbol (ca, s_bitmaps [2 * best_sym], lane, afg, abg);
                mean_colors_for_symbol (cb, s_bitmaps [2 * best_sym + 1], lane, bfg, bbg);
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

static int
upload_tables ()
{
    static bool tables_ready = false;

    if (!tables_ready)
    {
        cudaError_t err = cudaMemcpyToSymbol (c_invdiv16, h_invdiv16, sizeof (h_invdiv16));
        if (err != cudaSuccess)
            return (int) err;
        tables_ready = true;
    }
    return 0;
}

int
dev_launch_match (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, void *stream)
{
    int n_cells = cv->width * cv->n_rows;
    int err = upload_tables ();

    if (err || n_cells <= 0)
        return err;

    size_t smem = (size_t) cv->syms.n * 8 + WARPS_PER_BLOCK * CAND_LIST_MAX * 4;
    int blocks_needed = (n_cells + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int grid = std::min (blocks_needed, num_sms () * 8);
    match_cells_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, (cudaStream_t) stream>>> (*cv, pixels, evals);
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

    size_t smem = (size_t) cv->syms.n2 * 16 + WARPS_PER_BLOCK * CAND_LIST_MAX * 4;
    int blocks_needed = (n_pairs + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    int grid = std::min (blocks_needed, num_sms () * 8);
    match_wide_kernel<<<grid, WARPS_PER_BLOCK * 32, smem, (cudaStream_t) stream>>> (*cv, pixels, wide_evals);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
