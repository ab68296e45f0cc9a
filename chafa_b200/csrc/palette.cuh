/* palette.cuh -- device-side colour helpers and fixed-palette pickers.
 *
 * Colours are packed little-endian as the reference's ChafaColor lies in memory:
 * ch0 | ch1 << 8 | ch2 << 16 | ch3 << 24 (ch3 = alpha).
 *
 * Restates chafa/internal/chafa-palette.c:96-125 (candidate tracking), 253-386
 * (fixed pickers, including the order-dependent tie rules and the gray-ramp walk
 * that records pen 244 with pen 245's error and reads one entry past either end
 * of the ramp) and 955-1037 (lookup). */
#ifndef CHAFA_B200_PALETTE_CUH
#define CHAFA_B200_PALETTE_CUH

#include "device.h"

#define HD __host__ __device__

#define CH(col, i) (int) (((col) >> ((i) * 8)) & 0xffu)

/* 3-channel sum of squared differences (chafa/internal/chafa-color.h:145-148) */
HD __forceinline__ int
color_diff3 (uint32_t a, uint32_t b)
{
    int d0 = CH (b, 0) - CH (a, 0);
    int d1 = CH (b, 1) - CH (a, 1);
    int d2 = CH (b, 2) - CH (a, 2);
    return d0 * d0 + d1 * d1 + d2 * d2;
}

/* Internal order -> the reference's packed ARGB cell colour (chafa-color.c:28-36) */
__device__ __forceinline__ uint32_t
pack_argb (uint32_t col)
{
    return __byte_perm (col, 0, 0x3012);
}

__device__ __forceinline__ uint32_t
unpack_argb (uint32_t packed)
{
    return __byte_perm (packed, 0, 0x3012);
}

struct ColorCandidates
{
    int index [2];
    int error [2];
};

HD __forceinline__ void
cand_init (ColorCandidates &c)
{
    c.index [0] = c.index [1] = -1;
    c.error [0] = c.error [1] = 0x7fffffff;
}

HD __forceinline__ void
cand_update (ColorCandidates &c, int index, int error)
{
    if (error < c.error [0])
    {
        c.index [1] = c.index [0];
        c.index [0] = index;
        c.error [1] = c.error [0];
        c.error [0] = error;
    }
    else if (error < c.error [1])
    {
        c.index [1] = index;
        c.error [1] = error;
    }
}

HD __forceinline__ int
cand_update_index (ColorCandidates &c, const DevPalette *pal, uint32_t color, int index)
{
    int error = color_diff3 (color, pal->fixed [index]);
    cand_update (c, index, error);
    return error;
}

HD inline void
pick_fixed_216_cube (ColorCandidates &c, const DevPalette *pal, uint32_t color)
{
    int i = 16 + pal->cube_index [CH (color, 0)] * 36 + pal->cube_index [CH (color, 1)] * 6
        + pal->cube_index [CH (color, 2)];
    cand_update_index (c, pal, color, i);
}

HD inline void
pick_fixed_24_grays (ColorCandidates &c, const DevPalette *pal, uint32_t color)
{
    int i = 232 + 12, step, error, last_error;

    last_error = cand_update_index (c, pal, color, i);
    error = color_diff3 (color, pal->fixed [i + 1]);
    if (error < last_error)
    {
        cand_update (c, i, error);
        last_error = error;
        step = 1;
        i++;
    }
    else
    {
        step = -1;
    }

    do
    {
        i += step;
        error = color_diff3 (color, pal->fixed [i]);
        if (error > last_error)
            break;
        cand_update (c, i, error);
        last_error = error;
    }
    while (i >= 232 && i <= 255);
}

HD inline void
pick_fixed_range (ColorCandidates &c, const DevPalette *pal, uint32_t color, int first, int end)
{
    for (int i = first; i < end; i++)
        cand_update_index (c, pal, color, i);
}

/* Sequential lookup, executed redundantly by whoever calls it.  @color_space 0 = RGB. */
HD inline int
palette_lookup_nearest (const DevPalette *pal, int color_space, uint32_t color, ColorCandidates *cand_out)
{
    ColorCandidates c;
    cand_init (c);

    if (CH (color, 3) < pal->alpha_threshold)
    {
        c.index [0] = c.index [1] = pal->transparent_index;
        c.error [0] = c.error [1] = 0;
    }
    else if (pal->type == PAL_FIXED_256)
    {
        if (color_space == 0)
        {
            pick_fixed_216_cube (c, pal, color);
            pick_fixed_24_grays (c, pal, color);
            pick_fixed_range (c, pal, color, 0, 16);
        }
        else
        {
            pick_fixed_range (c, pal, color, 0, 256);
        }
    }
    else if (pal->type == PAL_FIXED_240)
    {
        if (color_space == 0)
        {
            pick_fixed_216_cube (c, pal, color);
            pick_fixed_24_grays (c, pal, color);
        }
        else
        {
            pick_fixed_range (c, pal, color, 16, 256);
        }
    }
    else if (pal->type == PAL_FIXED_16)
    {
        pick_fixed_range (c, pal, color, 0, 16);
    }
    else if (pal->type == PAL_FIXED_8)
    {
        pick_fixed_range (c, pal, color, 0, 8);
    }
    else
    {
        cand_update (c, PAL_INDEX_FG, color_diff3 (color, pal->colors [PAL_INDEX_FG]));
        cand_update (c, PAL_INDEX_BG, color_diff3 (color, pal->colors [PAL_INDEX_BG]));
    }

    if (pal->transparent_index < 256)
    {
        if (c.index [0] == pal->transparent_index)
        {
            c.index [0] = c.index [1];
            c.error [0] = c.error [1];
        }
        else
        {
            if (c.index [0] == PAL_INDEX_TRANSPARENT) c.index [0] = pal->transparent_index;
            if (c.index [1] == PAL_INDEX_TRANSPARENT) c.index [1] = pal->transparent_index;
        }
    }

    if (cand_out)
        *cand_out = c;
    return c.index [0];
}

#ifdef __CUDACC__

/* Warp-cooperative lookup: every lane calls it with the same arguments and gets the
 * best pen (index [0] of the sequential version) back.
 *
 * The sequential pickers evaluate a fixed sequence of (pen, error) pairs and keep the
 * first one with the strictly smallest error.  Here every pair of the sequence is
 * scored by its own lane and the winner is min over (error << 6 | position in the
 * sequence).  The only data-dependent part of the sequence is the gray-ramp walk
 * (chafa-palette.c:276-305): it starts at pen 244, turns towards 245 or 243, and keeps
 * going while the error does not increase -- including one entry past either end of
 * the ramp (pens 231 and 256), and recording pen 244 a second time with pen 245's
 * error when it turns upwards.  Which ramp entries it visits follows from the errors of
 * the 26 pens 231..256, one per lane, and two ballots. */
__device__ inline int
palette_lookup_nearest_warp (const DevPalette *pal, int color_space, uint32_t color, int lane)
{
    const unsigned FULLMASK = 0xffffffffu;
    const int type = pal->type;

    if (CH (color, 3) < pal->alpha_threshold || pal->transparent_index < 256)
        return palette_lookup_nearest (pal, color_space, color, nullptr);

    if (color_space != 0 && (type == PAL_FIXED_256 || type == PAL_FIXED_240))
    {
        /* linear scan over the pens in the canvas's colour space; lowest pen wins ties */
        int first = type == PAL_FIXED_240 ? 16 : 0;
        unsigned best = 0xffffffffu;

        for (int i = first + lane; i < 256; i += 32)
        {
            unsigned key = ((unsigned) color_diff3 (color, pal->fixed [i]) << 9) | (unsigned) i;
            best = min (best, key);
        }
        best = __reduce_min_sync (FULLMASK, best);
        return (int) (best & 0x1ff);
    }

    if (type == PAL_FIXED_16 || type == PAL_FIXED_8)
    {
        int n = type == PAL_FIXED_16 ? 16 : 8;
        unsigned key = 0xffffffffu;
        if (lane < n)
            key = ((unsigned) color_diff3 (color, pal->fixed [lane]) << 5) | (unsigned) lane;
        return (int) (__reduce_min_sync (FULLMASK, key) & 31u);
    }

    if (type != PAL_FIXED_256 && type != PAL_FIXED_240)
        return palette_lookup_nearest (pal, color_space, color, nullptr);

    /* RGB, 256 / 240 pens: cube entry, gray-ramp walk, then (256 only) pens 0..15 */
    int cube = 16 + pal->cube_index [CH (color, 0)] * 36 + pal->cube_index [CH (color, 1)] * 6
        + pal->cube_index [CH (color, 2)];
    unsigned key = ((unsigned) color_diff3 (color, pal->fixed [cube]) << 6) | 0u;

    /* lane L < 26 scores pen 231 + L */
    int e = lane < 26 ? color_diff3 (color, pal->fixed [231 + lane]) : 0x7fffffff;
    int e_prev = __shfl_up_sync (FULLMASK, e, 1);
    int e_next = __shfl_down_sync (FULLMASK, e, 1);
    int e244 = __shfl_sync (FULLMASK, e, 13), e245 = __shfl_sync (FULLMASK, e, 14);
    bool up = e245 < e244;
    unsigned brk_up = __ballot_sync (FULLMASK, lane >= 15 && lane < 26 && e > e_prev) & 0x03ff8000u;
    unsigned brk_dn = __ballot_sync (FULLMASK, lane <= 12 && e > e_next) & 0x00001fffu;

    if (lane == 13)
    {
        key = min (key, ((unsigned) e << 6) | 1u);
    }
    else if (up)
    {
        int stop = brk_up ? __ffs (brk_up) - 1 : 26;
        if (lane == 14)
            key = min (key, ((unsigned) e << 6) | 2u);          /* pen 244 again, with 245's error */
        else if (lane >= 15 && lane < stop)
            key = min (key, ((unsigned) e << 6) | (unsigned) (lane - 12));   /* pen 246 -> position 3 */
    }
    else
    {
        int stop = brk_dn ? 31 - __clz (brk_dn) : -1;
        if (lane <= 12 && lane > stop)
            key = min (key, ((unsigned) e << 6) | (unsigned) (14 - lane));   /* pen 243 -> position 2 */
    }

    if (type == PAL_FIXED_256 && lane < 16)
        key = min (key, ((unsigned) color_diff3 (color, pal->fixed [lane]) << 6) | (unsigned) (32 + lane));

    unsigned best = __reduce_min_sync (FULLMASK, key);
    int pos = (int) (best & 63u);

    if (pos == 0)
        return cube;
    if (pos >= 32)
        return pos - 32;
    if (pos <= 2 && (pos == 1 || up))
        return 244;
    return up ? 243 + pos : 245 - pos;
}

#endif /* __CUDACC__ */

#endif /* CHAFA_B200_PALETTE_CUH */
