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

/* Warp-cooperative variant for the long linear scans (DIN99d 256/240-colour):
 * every lane must call it with the same arguments; returns the best index only.
 * Ties go to the lowest index, as the sequential strict-< scan does. */
__device__ inline int
palette_lookup_nearest_warp (const DevPalette *pal, int color_space, uint32_t color, int lane)
{
    bool linear_scan = color_space != 0 && (pal->type == PAL_FIXED_256 || pal->type == PAL_FIXED_240);

    if (!linear_scan || CH (color, 3) < pal->alpha_threshold || pal->transparent_index < 256)
        return palette_lookup_nearest (pal, color_space, color, nullptr);

    int first = pal->type == PAL_FIXED_240 ? 16 : 0;
    unsigned best = 0xffffffffu;

    for (int i = first + lane; i < 256; i += 32)
    {
        unsigned key = ((unsigned) color_diff3 (color, pal->fixed [i]) << 9) | (unsigned) i;
        best = min (best, key);
    }
    best = __reduce_min_sync (0xffffffffu, best);
    return (int) (best & 0x1ff);
}

#endif /* CHAFA_B200_PALETTE_CUH */
