/* scale.cu -- kernel K1: scale the source image to the canvas pixmap, composite it over
 * the background colour and repack to RGBA8, in one pass.
 *
 * One thread produces one destination pixel.  It gathers the source samples the
 * reference's separable filters would touch for that pixel and applies the same
 * integer arithmetic in the same order, so the result is bit-identical:
 *
 *   unpack  8-bit sRGB -> 11-bit linear via LUT, premultiplied by (alpha + 1)
 *           (smolscale-generic.c:924-947 for unassociated sources, 742-800 for
 *            premultiplied / 24-bit ones)
 *   H pass  copy | one | nearest | bilinear with 0-3 halvings | box
 *           (smolscale-generic.c:1991-2315), edge opacity (1971-1989)
 *   V pass  same filter family over H-filtered rows (2350-3328)
 *   composite over the background colour keeping source alpha (3475-3607)
 *   pack    un-premultiply by reciprocal LUT, linear -> sRGB LUT (1644-1716)
 *
 * The reference works on two uint64 "parts" per pixel holding four 32-bit lanes; the
 * lanes never interact for valid input, so four uint32 lanes are used here.  Lane 3
 * is alpha.  The filter plan (which filter, sample offsets and weights) is computed on
 * the host (host_scale.cpp) and read from global memory.
 *
 * HBM-bound: 4 B read + 4 B written per pixel at 1:1; source rows are read through L1
 * / L2 when the vertical filter touches several rows per output pixel. */

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "device.h"
#include "prep_pixel.cuh"

enum
{
    F_COPY = 0, F_ONE, F_BILINEAR_0H, F_BILINEAR_1H, F_BILINEAR_2H, F_BILINEAR_3H, F_BOX, F_NEAREST
};

#define BOXES_MULTIPLIER 16777216ull

struct Lanes
{
    uint32_t l [4];
};

struct ScaleShared
{
    uint16_t from_srgb [256];
    uint8_t to_srgb [2048];
    uint32_t inv_div [4] [256];   /* p8, p8l, p16, p16l */
};

/* The same tables where they lie in global memory (read through L1): for kernels whose
 * common path does not need them */
struct ScaleGlobal
{
    const uint16_t *from_srgb;
    const uint8_t *to_srgb;
    const uint32_t *inv_div [4];
};

/* Pixel formats, chafa.h order: 0-3 premultiplied RGBA/BGRA/ARGB/ABGR, 4-7 the same
 * unassociated, 8 RGB8, 9 BGR8.  Returns r, g, b, a bytes. */
__device__ __forceinline__ void
fetch_rgba (const DevScale &p, int x, int y, uint32_t &r, uint32_t &g, uint32_t &b, uint32_t &a)
{
    const uint8_t *row = p.src + (size_t) y * (size_t) p.src_rowstride;
    int t = p.src_pixel_type;

    if (t >= 8)
    {
        const uint8_t *q = row + (size_t) x * 3;
        uint32_t c0 = q [0], c1 = q [1], c2 = q [2];
        a = 0xff;
        g = c1;
        if (t == 8) { r = c0; b = c2; } else { r = c2; b = c0; }
        return;
    }

    uint32_t v;
    if ((((uintptr_t) row) & 3) == 0)
        v = __ldg ((const uint32_t *) row + x);
    else
    {
        const uint8_t *q = row + (size_t) x * 4;
        v = q [0] | ((uint32_t) q [1] << 8) | ((uint32_t) q [2] << 16) | ((uint32_t) q [3] << 24);
    }

    uint32_t c0 = v & 0xff, c1 = (v >> 8) & 0xff, c2 = (v >> 16) & 0xff, c3 = v >> 24;
    switch (t & 3)
    {
        case 0: r = c0; g = c1; b = c2; a = c3; break;   /* RGBA */
        case 1: r = c2; g = c1; b = c0; a = c3; break;   /* BGRA */
        case 2: r = c1; g = c2; b = c3; a = c0; break;   /* ARGB */
        default: r = c3; g = c2; b = c1; a = c0; break;  /* ABGR */
    }
}

__device__ __forceinline__ bool
src_is_unassociated (int t)
{
    return t >= 4 && t <= 7;
}

/* Source pixel -> internal lanes.  x == src width yields the zero pad pixel the
 * reference keeps at the end of every unpacked row. */
template <class Tables>
__device__ __forceinline__ Lanes
unpack (const DevScale &p, const Tables &s, int x, int y)
{
    Lanes o;

    if (x >= p.h.src_size_px)
    {
        o.l [0] = o.l [1] = o.l [2] = o.l [3] = 0;
        return o;
    }

    uint32_t c [3], a;
    fetch_rgba (p, x, y, c [0], c [1], c [2], a);

    if (src_is_unassociated (p.src_pixel_type) && p.plain)
    {
        /* unassociated -> premultiplied output: 11-bit linear, premultiplied on 8 bits
         * (smolscale-generic.c:881-892, 441-448) */
#pragma unroll
        for (int i = 0; i < 3; i++)
            o.l [i] = (((p.linear ? (uint32_t) s.from_srgb [c [i]] : c [i]) * (a + 1)) >> 8) & 0x7ff;
        o.l [3] = (a << 8) | 0xff;
    }
    else if (src_is_unassociated (p.src_pixel_type))
    {
        /* premul16: value * (alpha + 1) */
#pragma unroll
        for (int i = 0; i < 3; i++)
            o.l [i] = (p.linear ? (uint32_t) s.from_srgb [c [i]] : c [i]) * (a + 1);
        o.l [3] = (a << 8) | 0xff;
    }
    else if (p.linear)
    {
        /* premul8 -> unassociated -> linear -> premul8 on 11 bits */
        uint32_t inv = s.inv_div [0] [a];
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t u = (uint32_t) (((uint64_t) c [i] * inv) >> 13) & 0xff;
            o.l [i] = (((uint32_t) s.from_srgb [u] * (a + 1)) >> 8) & 0x7ff;
        }
        o.l [3] = (a << 8) | 0xff;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 3; i++)
            o.l [i] = c [i];
        o.l [3] = a;
    }
    return o;
}

__device__ __forceinline__ uint32_t
lerp_lane (uint32_t pv, uint32_t qv, uint32_t f)
{
    /* ((p - q) * F >> 8) + q on a 24-bit lane == (p * F + q * (256 - F)) >> 8 */
    return ((pv * f + qv * (256u - f)) >> 8) & 0x00ffffffu;
}

__device__ __forceinline__ Lanes
lerp (const Lanes &a, const Lanes &b, uint32_t f)
{
    Lanes o;
#pragma unroll
    for (int i = 0; i < 4; i++)
        o.l [i] = lerp_lane (a.l [i], b.l [i], f);
    return o;
}

__device__ __forceinline__ Lanes
weight (const Lanes &a, uint32_t w)
{
    Lanes o;
#pragma unroll
    for (int i = 0; i < 4; i++)
        o.l [i] = (uint32_t) (((uint64_t) a.l [i] * w) >> 8) & 0x00ffffffu;
    return o;
}

__device__ __forceinline__ void
add_to (Lanes &acc, const Lanes &a)
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        acc.l [i] += a.l [i];
}

__device__ __forceinline__ Lanes
box_finalize (const Lanes &acc, uint32_t span_mul)
{
    Lanes o;
#pragma unroll
    for (int i = 0; i < 4; i++)
        o.l [i] = (uint32_t) (((uint64_t) acc.l [i] * span_mul + BOXES_MULTIPLIER / 2) / BOXES_MULTIPLIER);
    return o;
}

__device__ __forceinline__ void
box_span (uint32_t precalc, uint32_t step, uint32_t &ofs0, uint32_t &f0, uint32_t &f1, uint32_t &n)
{
    uint32_t o0 = precalc, o1 = precalc + step;
    f0 = 256 - (o0 % 256);
    f1 = o1 % 256;
    o0 /= 256;
    o1 /= 256;
    ofs0 = o0;
    n = o1 - o0 - 1;
}

/* H-filtered lanes of visible column @xr on source row @sy */
template <class Tables>
__device__ __forceinline__ Lanes
hsample (const DevScale &p, const Tables &s, int xr, int sy)
{
    const DevScaleDim &h = p.h;
    Lanes o;

    switch (h.filter)
    {
        case F_COPY:
            o = unpack (p, s, xr + h.clip_before_px, sy);
            break;
        case F_ONE:
            o = unpack (p, s, 0, sy);
            break;
        case F_NEAREST:
            o = unpack (p, s, (int) (h.precalc [xr] & 0xffff), sy);
            break;
        case F_BOX:
        {
            uint32_t ofs0, f0, f1, n;
            box_span (h.precalc [xr], h.span_step, ofs0, f0, f1, n);
            Lanes acc = weight (unpack (p, s, (int) ofs0, sy), f0);
            for (uint32_t i = 0; i < n; i++)
                add_to (acc, unpack (p, s, (int) (ofs0 + 1 + i), sy));
            add_to (acc, weight (unpack (p, s, (int) (ofs0 + 1 + n), sy), f1));
            o = box_finalize (acc, h.span_mul);
            break;
        }
        default:
        {
            int nh = h.n_halvings;
            Lanes acc;
            acc.l [0] = acc.l [1] = acc.l [2] = acc.l [3] = 0;
            for (int i = 0; i < (1 << nh); i++)
            {
                uint32_t pc = h.precalc [(xr << nh) + i];
                int ofs = (int) (pc & 0xffff);
                Lanes a = unpack (p, s, ofs, sy);
                Lanes b = unpack (p, s, ofs + 1, sy);
                add_to (acc, lerp (a, b, pc >> 16));
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
                o.l [i] = (acc.l [i] >> nh) & 0x00ffffffu;
            break;
        }
    }

    /* Horizontal edge opacity is applied to every H-filtered row */
    if (xr == 0)
        o = weight (o, (uint32_t) h.first_opacity);
    if (xr == h.placement_size_px - 1)
        o = weight (o, (uint32_t) h.last_opacity);
    return o;
}

__device__ __forceinline__ Lanes
apply_v_opacity (const DevScale &p, const Lanes &in, int yr)
{
    if (yr == 0 && p.v.first_opacity < 256)
        return weight (in, (uint32_t) p.v.first_opacity);
    if (yr == p.v.placement_size_px - 1 && p.v.last_opacity < 256)
        return weight (in, (uint32_t) p.v.last_opacity);
    return in;
}

template <class Tables>
__device__ __forceinline__ Lanes
vsample (const DevScale &p, const Tables &s, int xr, int yr)
{
    const DevScaleDim &v = p.v;

    switch (v.filter)
    {
        case F_COPY:
            return hsample (p, s, xr, yr + v.clip_before_px);
        case F_ONE:
            return apply_v_opacity (p, hsample (p, s, xr, 0), yr);
        case F_NEAREST:
            return hsample (p, s, xr, (int) (v.precalc [yr] & 0xffff));
        case F_BOX:
        {
            uint32_t ofs0, w1, w2, n;
            box_span (v.precalc [yr], v.span_step, ofs0, w1, w2, n);
            Lanes acc = weight (hsample (p, s, xr, (int) ofs0), w1);
            for (uint32_t i = 0; i < n; i++)
                add_to (acc, hsample (p, s, xr, (int) (ofs0 + 1 + i)));
            if (w2 > 0 && ofs0 + 1 + n < (uint32_t) v.src_size_px)
                add_to (acc, weight (hsample (p, s, xr, (int) (ofs0 + 1 + n)), w2));
            return apply_v_opacity (p, box_finalize (acc, v.span_mul), yr);
        }
        default:
        {
            int nh = v.n_halvings;
            Lanes acc;
            acc.l [0] = acc.l [1] = acc.l [2] = acc.l [3] = 0;
            for (int i = 0; i < (1 << nh); i++)
            {
                uint32_t pc = v.precalc [(yr << nh) + i];
                int ofs = (int) (pc & 0xffff);
                Lanes a = hsample (p, s, xr, ofs);
                Lanes b = hsample (p, s, xr, ofs + 1);
                add_to (acc, lerp (a, b, pc >> 16));
            }
            Lanes o;
#pragma unroll
            for (int i = 0; i < 4; i++)
                o.l [i] = (acc.l [i] >> nh) & 0x00ffffffu;
            return apply_v_opacity (p, o, yr);
        }
    }
}

/* Composite over the opaque colour (p.composite == 1) or not at all (2), all four lanes alike
 * (smolscale-generic.c:3331-3436), then pack to unassociated RGBA8 (1600-1716).  The colour is
 * unpacked like a source pixel of alpha 255: channel * 256 / linear channel * 256 for the 16-bit
 * premultiplied formats, the (linear) channel itself for the 8-bit ones; alpha lane all ones. */
template <class Tables>
__device__ __forceinline__ uint32_t
composite_over_and_pack (const DevScale &p, const Tables &s, const Lanes &in)
{
    const uint32_t bg [3] = { p.bg_rgb & 0xff, (p.bg_rgb >> 8) & 0xff, (p.bg_rgb >> 16) & 0xff };
    const bool unassoc = src_is_unassociated (p.src_pixel_type);
    const bool over = p.composite == 1;
    uint32_t t [4] = { in.l [0], in.l [1], in.l [2], in.l [3] };
    uint32_t out [3], a;

    if (unassoc || p.linear)
    {
        if (over)
        {
            uint32_t a_src = (in.l [3] >> 8) & 0xff;
            uint32_t nz = (a_src + 0xff) >> 8;
            uint32_t w = 0x100 - a_src - nz;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                uint32_t colv;
                if (i == 3)
                    colv = 0xffffu;
                else if (unassoc)
                    colv = (p.linear ? (uint32_t) s.from_srgb [bg [i]] : bg [i]) * 256u;
                else
                    colv = (uint32_t) s.from_srgb [bg [i]] & 0x7ff;
                t [i] = in.l [i] * nz + ((uint32_t) (((uint64_t) colv * w + 0x80) >> 8) & 0x00ffffffu);
            }
        }
        a = (t [3] >> 8) & 0xff;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            if (unassoc)
            {
                uint32_t inv = s.inv_div [p.linear ? 3 : 2] [a];
                if (p.linear)
                    out [i] = s.to_srgb [(uint32_t) (((uint64_t) t [i] * inv) >> 19) & 0x7ff];
                else
                    out [i] = (uint32_t) (((uint64_t) t [i] * inv) >> 16) & 0xff;
            }
            else
            {
                uint32_t inv = s.inv_div [1] [a];
                out [i] = s.to_srgb [(uint32_t) (((uint64_t) t [i] * inv) >> 11) & 0x7ff];
            }
        }
    }
    else
    {
        if (over)
        {
            uint32_t a_src = in.l [3] & 0xff;
            uint32_t nz = (a_src + 0xff) >> 8;
            uint32_t w = 0xff - a_src;
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                uint32_t tt = (i == 3 ? 0xffu : bg [i]) * w + 0x80;
                t [i] = in.l [i] * nz + (((tt + ((tt >> 8) & 0x00ffffffu)) >> 8) & 0x00ffffffu);
            }
        }
        a = t [3] & 0xff;
        uint32_t inv = s.inv_div [0] [a];
#pragma unroll
        for (int i = 0; i < 3; i++)
            out [i] = (uint32_t) (((uint64_t) t [i] * inv) >> 13) & 0xff;
    }

    return out [0] | (out [1] << 8) | (out [2] << 16) | (a << 24);
}

/* Composite over the (opaque) colour, keep the source alpha, repack to RGBA8 */
template <class Tables>
__device__ __forceinline__ uint32_t
composite_and_pack (const DevScale &p, const Tables &s, const Lanes &in)
{
    if (p.composite != 0)
        return composite_over_and_pack (p, s, in);

    uint32_t bg [3] = { p.bg_rgb & 0xff, (p.bg_rgb >> 8) & 0xff, (p.bg_rgb >> 16) & 0xff };
    uint32_t out [3], a;
    bool unassoc = src_is_unassociated (p.src_pixel_type);

    if (unassoc || p.linear)
    {
        a = (in.l [3] >> 8) & 0xff;
        uint32_t nz = (a + 0xff) >> 8;
        uint32_t w = 0x100 - a - nz;

#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t colv;
            if (unassoc)
                colv = (p.linear ? (uint32_t) s.from_srgb [bg [i]] : bg [i]) * 256u;   /* premul16, alpha 255 */
            else
                colv = (((uint32_t) s.from_srgb [bg [i]] * 256u) >> 8) & 0x7ff;           /* premul8 linear */

            uint32_t t = in.l [i] * nz + ((uint32_t) (((uint64_t) colv * w + 0x80) >> 8) & 0x00ffffffu);

            if (unassoc)
            {
                t = ((t >> 8) & 0xffffu) * (a + 1);
                uint32_t inv = s.inv_div [p.linear ? 3 : 2] [a];
                if (p.linear)
                    out [i] = s.to_srgb [(uint32_t) (((uint64_t) t * inv) >> 19) & 0x7ff];
                else
                    out [i] = (uint32_t) (((uint64_t) t * inv) >> 16) & 0xff;
            }
            else
            {
                t = (((t & 0xfffu) * (a + 1)) >> 8) & 0xfffu;
                uint32_t inv = s.inv_div [1] [a];
                out [i] = s.to_srgb [(uint32_t) (((uint64_t) t * inv) >> 11) & 0x7ff];
            }
        }
    }
    else
    {
        /* premul8, compressed gamma (only when the source is > 8191x the output) */
        a = in.l [3] & 0xff;
        uint32_t nz = (a + 0xff) >> 8;
        uint32_t w = 0xff - a;
        uint32_t inv = s.inv_div [0] [a];

#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t t = bg [i] * w + 0x80;
            t = in.l [i] * nz + (((t + ((t >> 8) & 0x00ffffffu)) >> 8) & 0x00ffffffu);
            t = (((t + 1) * (a + 1) - 1) >> 8) & 0xff;
            out [i] = (uint32_t) (((uint64_t) t * inv) >> 13) & 0xff;
        }
    }

    return out [0] | (out [1] << 8) | (out [2] << 16) | (a << 24);
}

/* Plain route: un-premultiply, linear -> sRGB, premultiply again, RGBA8 premultiplied out
 * (smolscale-generic.c:1651-1662, 360-377) */
template <class Tables>
__device__ __forceinline__ uint32_t
pack_premultiplied (const DevScale &p, const Tables &s, const Lanes &in)
{
    uint32_t out [3], a;

    if (p.linear)
    {
        a = (in.l [3] >> 8) & 0xff;
        uint32_t inv = s.inv_div [1] [a];
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            uint32_t u = s.to_srgb [(uint32_t) (((uint64_t) in.l [i] * inv) >> 11) & 0x7ff];
            out [i] = (((u + 1) * (a + 1) - 1) >> 8) & 0xff;
        }
    }
    else
    {
        a = in.l [3] & 0xff;
#pragma unroll
        for (int i = 0; i < 3; i++)
            out [i] = in.l [i] & 0xff;
    }
    return out [0] | (out [1] << 8) | (out [2] << 16) | (a << 24);
}

/* HF, VF >= 0: the kernel is specialised for that pair of filters (copy or bilinear with 0-3
 * halvings) on unassociated 32-bit sources, composited over the background colour: it works on a
 * copy of the parameter block in which those fields are compile-time constants, so the filter and
 * format switches of the general code fold away (700-2000 instructions instead of 24000).
 * HF = VF = -1: everything at run time. */
template <int HF, int VF>
__global__ void __launch_bounds__ (256)
scale_kernel (const __grid_constant__ DevScale p_in)
{
    __shared__ ScaleShared s;

    DevScale p = p_in;
    if (HF >= 0)
    {
        p.h.filter = HF;
        p.h.n_halvings = HF >= F_BILINEAR_0H ? HF - F_BILINEAR_0H : 0;
        p.v.filter = VF;
        p.v.n_halvings = VF >= F_BILINEAR_0H ? VF - F_BILINEAR_0H : 0;
        p.src_pixel_type = 4 | (p_in.src_pixel_type & 3);
        p.linear = 1;
        p.plain = 0;
        p.composite = 0;
    }

    for (int i = threadIdx.x; i < 256; i += blockDim.x)
    {
        s.from_srgb [i] = p.from_srgb [i];
        s.inv_div [0] [i] = p.inv_div [i];
        s.inv_div [1] [i] = p.inv_div [256 + i];
        s.inv_div [2] [i] = p.inv_div [512 + i];
        s.inv_div [3] [i] = p.inv_div [768 + i];
    }
    for (int i = threadIdx.x; i < 2048; i += blockDim.x)
        s.to_srgb [i] = p.to_srgb [i];
    __syncthreads ();

    /* Pixel written where the placement does not reach: the colour at zero coverage */
    Lanes zero;
    zero.l [0] = zero.l [1] = zero.l [2] = zero.l [3] = 0;
    const uint32_t clear_pixel = p.plain ? 0u : composite_and_pack (p, s, zero);

    const bool copy_1to1 = p.h.filter == F_COPY && p.v.filter == F_COPY && p.linear && !p.plain;

    /* blockIdx.x: column strip of 256 pixels; blockIdx.y strides over the rows */
    const int x = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= p.dest_w)
        return;

    for (int row = (int) blockIdx.y; row < p.n_rows; row += (int) gridDim.y)
    {
        int y = p.first_row + row;
        int xr = x - p.h.placement_ofs_px;
        int yr = y - p.v.placement_ofs_px;
        uint32_t out = clear_pixel;

        if (xr >= 0 && xr < p.h.placement_size_px && yr >= 0 && yr < p.v.placement_size_px)
        {
            bool done = false;

            if (copy_1to1)
            {
                /* Pixel-aligned 1:1 placement: an opaque source pixel comes out unchanged.
                 * linear -> premultiply by 256 -> composite with zero background weight ->
                 * un-premultiply -> sRGB round-trips exactly for every byte value at alpha
                 * 255 (the LUTs are built for that; checked for all 256 values in
                 * tests/test_cpu.py).  Only the channel order is applied. */
                uint32_t r, g, b, a;
                fetch_rgba (p, xr + p.h.clip_before_px, yr + p.v.clip_before_px, r, g, b, a);
                if (a == 0xff)
                {
                    out = r | (g << 8) | (b << 16) | 0xff000000u;
                    done = true;
                }
            }
            if (!done && ((p.plain && p.src_pixel_type == 0) || (p.composite == 2 && p.src_pixel_type == 4))
                && p.h.filter == F_COPY && p.v.filter == F_COPY)
            {
                /* same format, 1:1: the reference copies the bytes (smolscale.c:1272-1283) */
                uint32_t r, g, b, a;
                fetch_rgba (p, xr + p.h.clip_before_px, yr + p.v.clip_before_px, r, g, b, a);
                out = r | (g << 8) | (b << 16) | (a << 24);
                done = true;
            }
            if (!done)
                out = p.plain ? pack_premultiplied (p, s, vsample (p, s, xr, yr))
                              : composite_and_pack (p, s, vsample (p, s, xr, yr));
        }

        if (p.post_on)
            out = prep_dither_convert (out, x, y, p.post);
        p.dest [(size_t) y * p.dest_w + x] = out;
    }
}

/* One pixel of the 1:1 route that is not opaque: the general arithmetic, tables in global memory.
 * Kept out of line so that the streaming loop stays small. */
__device__ __forceinline__ uint32_t
scale_copy_translucent_inline (const DevScale &p, uint32_t rgba)
{
    ScaleGlobal g;
    g.from_srgb = p.from_srgb;
    g.to_srgb = p.to_srgb;
    g.inv_div [0] = p.inv_div;
    g.inv_div [1] = p.inv_div + 256;
    g.inv_div [2] = p.inv_div + 512;
    g.inv_div [3] = p.inv_div + 768;
    /* unpack () of an unassociated source pixel, not plain: premul16 = (linear) value * (alpha + 1) */
    Lanes in;
    uint32_t a = rgba >> 24;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        uint32_t c = (rgba >> (8 * i)) & 0xffu;
        in.l [i] = (p.linear ? (uint32_t) g.from_srgb [c] : c) * (a + 1);
    }
    in.l [3] = (a << 8) | 0xff;
    return composite_and_pack (p, g, in);
}

__device__ __noinline__ uint32_t
scale_copy_translucent (const DevScale *pp, uint32_t rgba)
{
    return scale_copy_translucent_inline (*pp, rgba);
}

/* 1:1 pixel-aligned placement of a 32-bit source: four pixels per thread, 128-bit loads and
 * stores, four rows in flight per thread.  Opaque pixels pass through (see scale_kernel); the
 * rest take the general route.  Needs 16-byte aligned rows on both sides.  The parameter block
 * is read from a copy in global memory by the out-of-line route only. */
template <int T, int DM>    /* channel order (type & 3) and dither mode of the post step as constants, or -1 */
__global__ void __launch_bounds__ (256)
scale_copy4_kernel (const __grid_constant__ DevScale p_in)
{
    DevScale p = p_in;
    if (T >= 0)
    {
        p.src_pixel_type = 4 | T;
        p.post_on = DM != 0;
        p.post.dither_mode = DM;
        p.post.to_din99d = 0;
        p.composite = 0;
        p.linear = 1;
        p.plain = 0;
    }
    const int x4 = (int) (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= p.dest_w)
        return;
    const int t = p.src_pixel_type & 3;
    const int sx = x4 - p.h.placement_ofs_px + p.h.clip_before_px;
    const int sy0 = p.first_row - p.v.placement_ofs_px + p.v.clip_before_px;
    const uint8_t *src = p.src + (size_t) sx * 4;
    const bool raw_translucent = p.composite == 2 && t == 0;     /* same format, no colour: bytes as they are */

    for (int row0 = (int) blockIdx.y * 4; row0 < p.n_rows; row0 += (int) gridDim.y * 4)
    {
        uint4 in [4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (row0 + r < p.n_rows)
                in [r] = __ldg ((const uint4 *) (src + (size_t) (sy0 + row0 + r) * (size_t) p.src_rowstride));

#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            if (row0 + r >= p.n_rows)
                break;
            const int y = p.first_row + row0 + r;
            uint32_t o [4] = { in [r].x, in [r].y, in [r].z, in [r].w };

            if (t != 0)
            {
                const uint32_t sel = t == 1 ? 0x3012u : t == 2 ? 0x0321u : 0x0123u;     /* BGRA, ARGB, ABGR */
#pragma unroll
                for (int k = 0; k < 4; k++)
                    o [k] = __byte_perm (o [k], 0, sel);
            }

            if ((o [0] & o [1] & o [2] & o [3]) < 0xff000000u && !raw_translucent)
            {
#pragma unroll 1
                for (int k = 0; k < 4; k++)
                    if (o [k] < 0xff000000u)
                        o [k] = T >= 0 ? scale_copy_translucent_inline (p, o [k])      /* constants folded */
                                       : scale_copy_translucent (&p_in, o [k]);
            }
            if (p.post_on)
            {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    o [k] = prep_dither_convert (o [k], x4 + k, y, p.post);
            }
            *(uint4 *) (p.dest + (size_t) y * p.dest_w + x4) = make_uint4 (o [0], o [1], o [2], o [3]);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* Row-staged scaler: the reducing (and enlarging) filters from TMA-staged source tiles.
 *
 * A block owns a tile of the destination -- blockDim.x columns by tiles.rows_per_block rows -- and
 * the rectangle of source pixels its filters read.  One thread issues the TMA loads
 * (cp.async.bulk.tensor.2d) that bring that rectangle into shared memory as boxes of eight source
 * rows, each completing on its own mbarrier, and the block starts on the first rows while the
 * rest are in flight; other resident blocks cover the latency.  A thread owns one destination
 * column and walks down it: every source row is unpacked and filtered horizontally exactly once
 * per column (the reference's row driver does the same per row, smolscale.c:582-702), the last two
 * H-filtered rows stay in registers because consecutive vertical samples share a row, and the
 * vertical samples are accumulated as they complete.  The arithmetic is the reference's, in its
 * order (smolscale-generic.c:925-947, 1991-2315, 2466-2541).
 *
 * Applies to unassociated 32-bit sources composited over the canvas colour (the `spec` condition
 * of dev_launch_scale) with 16-byte aligned rows, filters copy / bilinear with up to 2 (H) or 3 (V)
 * halvings.  Reading past the image is defined for TMA (zero fill); the pad pixel of the reference
 * (all lanes zero) is substituted explicitly. */

__device__ __forceinline__ uint32_t
st_smem_u32 (const void *ptr)
{
    return (uint32_t) __cvta_generic_to_shared (ptr);
}

__device__ __forceinline__ void
st_mbar_init (uint32_t bar, uint32_t count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void
st_mbar_expect_tx (uint32_t bar, uint32_t bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
st_mbar_wait (uint32_t bar, uint32_t parity)
{
    asm volatile ("{\n"
                  ".reg .pred p;\n"
                  "WAIT_%=:\n"
                  "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                  "@p bra DONE_%=;\n"
                  "bra WAIT_%=;\n"
                  "DONE_%=:\n"
                  "}" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void
st_tma_load_2d (uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile ("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                  :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ __forceinline__ int
st_src_first (const DevScaleDim &d, int r)
{
    if (d.filter == F_COPY)
        return r + d.clip_before_px;
    return (int) (d.precalc [r << d.n_halvings] & 0xffff);
}

__device__ __forceinline__ int
st_src_last (const DevScaleDim &d, int r)
{
    if (d.filter == F_COPY)
        return r + d.clip_before_px;
    return (int) (d.precalc [(r << d.n_halvings) + (1 << d.n_halvings) - 1] & 0xffff) + 1;
}

/* The tables the row-staged kernel reads */
struct __align__ (16) StagedTables
{
    uint16_t from_srgb [256];
    uint8_t to_srgb [2048];
    uint32_t inv16l [256];          /* inv_div [3]: reciprocals for 16-bit premultiplied linear values */
};
#define ST_VSAMPLES_MAX (64 << 3)   /* vertical samples of one block: rows_per_block <= 64, eight each at most */

/* premul16 lanes of an unassociated source pixel: (linear) channel * (alpha + 1)
 * (smolscale-generic.c:925-947); @sel: byte selector that brings the pixel to R, G, B, A order */
__device__ __forceinline__ Lanes
st_unpack (const StagedTables &s, uint32_t v, uint32_t sel)
{
    v = __byte_perm (v, 0, sel);
    const uint32_t a1 = (v >> 24) + 1;
    Lanes o;
    o.l [0] = (uint32_t) s.from_srgb [v & 0xff] * a1;
    o.l [1] = (uint32_t) s.from_srgb [(v >> 8) & 0xff] * a1;
    o.l [2] = (uint32_t) s.from_srgb [(v >> 16) & 0xff] * a1;
    o.l [3] = (a1 << 8) - 1;                                    /* (a << 8) | 0xff */
    return o;
}

/* lerp () without the 24-bit masks: the lanes of this route stay below 2^19 (11-bit linear value
 * times at most 256), so the masks of the general code never clear a bit */
__device__ __forceinline__ void
st_lerp_add (Lanes &acc, const Lanes &a, const Lanes &b, uint32_t f)
{
    const uint32_t g = 256u - f;
#pragma unroll
    for (int i = 0; i < 4; i++)
        acc.l [i] += (a.l [i] * f + b.l [i] * g) >> 8;
}

/* composite_and_pack () for this route (unassociated source, linear light, colour under the image,
 * source alpha kept).  Every product whose bits the general code takes from a 64-bit result has
 * them in the low word here: colv * w < 2^27, and bits 19-29 of t * inv lie in the low word. */
__device__ __forceinline__ uint32_t
st_composite_pack (const StagedTables &s, const uint32_t colv [3], const Lanes &in)
{
    const uint32_t a = (in.l [3] >> 8) & 0xff;
    const uint32_t nz = (a + 0xff) >> 8;
    const uint32_t w = 0x100 - a - nz;
    const uint32_t inv = s.inv16l [a];
    uint32_t out = a << 24;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        uint32_t t = in.l [i] * nz + (((colv [i] * w + 0x80) >> 8) & 0x00ffffffu);
        t = ((t >> 8) & 0xffffu) * (a + 1);
        out |= (uint32_t) s.to_srgb [((t * inv) >> 19) & 0x7ff] << (8 * i);
    }
    return out;
}

template <int HF, int VF>
__global__ void __launch_bounds__ (128)
scale_tma_kernel (const __grid_constant__ DevScale p_in, const __grid_constant__ CUtensorMap src_map)
{
    extern __shared__ __align__ (128) uint8_t s_tile_bytes [];
    __shared__ StagedTables s;
    __shared__ float s_linear_f [(HF == F_COPY && VF != F_COPY) ? 256 : 1];     /* from_srgb as floats (float walk) */
    __shared__ uint32_t s_vsamples [VF == F_COPY ? 1 : ST_VSAMPLES_MAX];
    __shared__ __align__ (8) uint64_t s_bars [SCALE_TILE_MAX_CHUNKS];

    DevScale p = p_in;
    p.h.filter = HF;
    p.h.n_halvings = HF >= F_BILINEAR_0H ? HF - F_BILINEAR_0H : 0;
    p.v.filter = VF;
    p.v.n_halvings = VF >= F_BILINEAR_0H ? VF - F_BILINEAR_0H : 0;
    constexpr int NH = HF >= F_BILINEAR_0H ? HF - F_BILINEAR_0H : 0;
    constexpr int NV = VF >= F_BILINEAR_0H ? VF - F_BILINEAR_0H : 0;

    const int tid = (int) threadIdx.x, tw = (int) blockDim.x;
    const int box_w = p.tiles.box_w;
    const uint32_t *tile = (const uint32_t *) s_tile_bytes;
    const int order = p_in.src_pixel_type & 3;
    const uint32_t sel = order == 0 ? 0x3210u : order == 1 ? 0x3012u : order == 2 ? 0x0321u : 0x0123u;

    /* this block's destination rows and columns, and the source rectangle under them */
    const int row0 = p.first_row + (int) blockIdx.y * p.tiles.rows_per_block;
    const int row1 = min (row0 + p.tiles.rows_per_block, p.first_row + p.n_rows);
    const int yr_a = max (row0 - p.v.placement_ofs_px, 0);
    const int yr_b = min (row1 - p.v.placement_ofs_px, p.v.placement_size_px);
    const int x0 = (int) blockIdx.x * tw;
    const int xr_a = max (x0 - p.h.placement_ofs_px, 0);
    const int xr_b = min (x0 + tw - p.h.placement_ofs_px, p.h.placement_size_px);
    int src_x0 = 0, src_y0 = 0, n_chunks = 0, n_src_rows = 0;
    if (yr_a < yr_b && xr_a < xr_b)
    {
        src_x0 = st_src_first (p.h, xr_a) & ~3;        /* a box starts on a 16-byte boundary of its row */
        src_y0 = st_src_first (p.v, yr_a);
        n_src_rows = st_src_last (p.v, yr_b - 1) - src_y0 + 1;
        n_chunks = (n_src_rows + SCALE_TILE_CHUNK_ROWS - 1) / SCALE_TILE_CHUNK_ROWS;
    }

    if (tid == 0 && n_chunks > 0)
    {
        for (int c = 0; c < n_chunks; c++)
            st_mbar_init (st_smem_u32 (&s_bars [c]), 1);
        asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t box_bytes = (uint32_t) box_w * SCALE_TILE_CHUNK_ROWS * 4u;
        for (int c = 0; c < n_chunks; c++)
        {
            const uint32_t bar = st_smem_u32 (&s_bars [c]);
            st_mbar_expect_tx (bar, box_bytes);
            st_tma_load_2d (st_smem_u32 (s_tile_bytes) + (uint32_t) c * box_bytes, &src_map, src_x0,
                            src_y0 + c * SCALE_TILE_CHUNK_ROWS, bar);
        }
    }

    /* the tables, while the tiles are on their way: 224 16-byte pieces (each table is an
     * allocation of its own, hence aligned) */
    /* (loops strided by the block size, a run-time value here: kept rolled, or the compiler derives
     * their trip counts with an integer division of some forty instructions each) */
#pragma unroll 1
    for (int i = tid; i < 224; i += tw)
    {
        const uint4 *src = i < 32 ? (const uint4 *) p.from_srgb + i
                         : i < 160 ? (const uint4 *) p.to_srgb + (i - 32)
                                   : (const uint4 *) (p.inv_div + 768) + (i - 160);
        ((uint4 *) &s) [i] = __ldg (src);
    }
    if (HF == F_COPY && VF != F_COPY)
    {
#pragma unroll 1
        for (int i = tid; i < 256; i += tw)
            s_linear_f [i] = (float) p.from_srgb [i];
    }
    /* this block's vertical samples: tile row of the upper source row | weight << 16 */
    if (VF != F_COPY)
    {
#pragma unroll 1
        for (int i = tid; i < ((yr_b - yr_a) << NV); i += tw)
        {
            const uint32_t pc = p.v.precalc [(yr_a << NV) + i];
            s_vsamples [i] = ((pc & 0xffffu) - (uint32_t) src_y0) | (pc & 0xffff0000u);
        }
    }
    __syncthreads ();       /* tables, samples, and the barriers' initialisation */

    /* Is every source pixel under this tile opaque?  Then the alpha lane is 0xffff throughout, the
     * premultiplication is a shift, and compositing leaves the colour alone: the tile takes a walk
     * with three lanes and without those steps (same values, fewer instructions).  Tiles at the
     * right and bottom edge of the image (pad pixel, rows past the end) and placements with
     * fractional edge coverage take the general walk. */
    bool tile_opaque = n_chunks > 0 && p.h.first_opacity == 256 && p.h.last_opacity == 256
        && p.v.first_opacity == 256 && p.v.last_opacity == 256
        && src_x0 + box_w <= p.h.src_size_px && src_y0 + n_src_rows <= p.v.src_size_px;
    if (__syncthreads_or (tile_opaque))        /* (block-uniform: one answer for everybody) */
    {
        for (int c = 0; c < n_chunks; c++)
            st_mbar_wait (st_smem_u32 (&s_bars [c]), 0);
        /* the rows of the tile lie back to back (box_w pixels each, a multiple of four) */
        uint32_t all_bits = 0xffffffffu;
        const uint4 *tile4 = (const uint4 *) tile;
        const int n4 = (n_src_rows * box_w) >> 2;
#pragma unroll 1
        for (int i = tid; i < n4; i += tw)
        {
            const uint4 v = tile4 [i];
            all_bits &= (v.x & v.y) & (v.z & v.w);
        }
        const uint32_t alpha_mask = order < 2 ? 0xff000000u : 0x000000ffu;
        tile_opaque = __syncthreads_and ((all_bits & alpha_mask) == alpha_mask) != 0;
    }

    const int x = x0 + tid;
    if (x >= p.dest_w)
        return;

    /* the canvas colour as the compositing step wants it, and the pixel where nothing is placed */
    const uint32_t colv [3] = { (uint32_t) s.from_srgb [p.bg_rgb & 0xff] * 256u,
                                (uint32_t) s.from_srgb [(p.bg_rgb >> 8) & 0xff] * 256u,
                                (uint32_t) s.from_srgb [(p.bg_rgb >> 16) & 0xff] * 256u };
    Lanes zero;
    zero.l [0] = zero.l [1] = zero.l [2] = zero.l [3] = 0;

    const int xr = x - p.h.placement_ofs_px;
    const bool col_in = xr >= 0 && xr < p.h.placement_size_px && yr_a < yr_b;
    /* (only blocks that reach outside the placement need it) */
    const bool any_clear = !col_in || row0 < yr_a + p.v.placement_ofs_px || row1 > yr_b + p.v.placement_ofs_px;
    const uint32_t clear_pixel = any_clear ? st_composite_pack (s, colv, zero) : 0u;
    uint32_t *dst = p.dest + (size_t) row0 * p.dest_w + x;
    int y = row0;

    auto put = [&] (uint32_t out)
    {
        if (p.post_on)
            out = prep_dither_convert (out, x, y, p.post);
        *dst = out;
        dst += p.dest_w;
        y++;
    };

    if (!col_in)
    {
        while (y < row1)
            put (clear_pixel);
        return;
    }
    while (y < yr_a + p.v.placement_ofs_px)     /* rows above the placement */
        put (clear_pixel);

    /* this column's horizontal taps: offsets into a tile row, weights, and whether a tap is the
     * reference's pad pixel past the end of the row (all lanes zero; TMA delivers zero bytes there,
     * which unpack to an alpha lane of 0xff) */
    int tap_ofs [1 << NH];
    uint32_t tap_f [1 << NH];
    bool tap_pad [1 << NH];
    if (HF == F_COPY)
    {
        tap_ofs [0] = xr + p.h.clip_before_px - src_x0;
        tap_f [0] = 0;
        tap_pad [0] = false;
    }
    else
    {
#pragma unroll
        for (int i = 0; i < (1 << NH); i++)
        {
            uint32_t pc = p.h.precalc [(xr << NH) + i];
            tap_ofs [i] = (int) (pc & 0xffff) - src_x0;
            tap_f [i] = pc >> 16;
            tap_pad [i] = (int) (pc & 0xffff) + 1 >= p.h.src_size_px;
        }
    }
    const bool edge_col = (xr == 0 && p.h.first_opacity < 256)
        || (xr == p.h.placement_size_px - 1 && p.h.last_opacity < 256);

    /* rows are visited top to bottom: a box is waited for when its first row comes up */
    auto box_wait = [&] (int lr)
    {
        if ((lr & (SCALE_TILE_CHUNK_ROWS - 1)) == 0)
            st_mbar_wait (st_smem_u32 (&s_bars [lr / SCALE_TILE_CHUNK_ROWS]), 0);
    };

    /* OPAQUE: three lanes.  A pixel's lanes are the 11-bit linear values themselves (the general
     * walk carries them times 256 = alpha + 1); the first blend of such values, p f + q (256 - f),
     * is what the general walk gets from ((256 p) f + (256 q) (256 - f)) >> 8, so from there on the
     * two walks hold the same numbers.  RAW_H: the H-filtered row is still unscaled (H = copy). */
    auto walk = [&] (auto opaque_tag)
    {
        constexpr bool OPAQUE = decltype (opaque_tag)::value;
        constexpr int NL = OPAQUE ? 3 : 4;
        constexpr bool RAW_H = OPAQUE && HF == F_COPY;

        auto unpack = [&] (uint32_t v) -> Lanes
        {
            if (!OPAQUE)
                return st_unpack (s, v, sel);
            v = __byte_perm (v, 0, sel);
            Lanes o;
            o.l [0] = s.from_srgb [v & 0xff];
            o.l [1] = s.from_srgb [(v >> 8) & 0xff];
            o.l [2] = s.from_srgb [(v >> 16) & 0xff];
            o.l [3] = 0;
            return o;
        };
        /* acc += blend of a and b; @first: the operands are unscaled pixel values */
        auto blend_add = [&] (Lanes &acc, const Lanes &a, const Lanes &b, uint32_t f, bool first)
        {
            const uint32_t g = 256u - f;
#pragma unroll
            for (int i = 0; i < NL; i++)
            {
                const uint32_t v = a.l [i] * f + b.l [i] * g;
                acc.l [i] += first ? v : v >> 8;
            }
        };
        auto pack = [&] (const Lanes &o, bool raw) -> uint32_t
        {
            if (!OPAQUE)
                return st_composite_pack (s, colv, o);
            /* alpha 255: no colour shows through; ((l >> 8) & 0xffff) * 256 * inv >> 19 */
            const uint32_t inv = s.inv16l [255];
            uint32_t out = 0xff000000u;
#pragma unroll
            for (int i = 0; i < 3; i++)
            {
                const uint32_t t = raw ? o.l [i] << 8 : o.l [i] & 0x00ffff00u;
                out |= (uint32_t) s.to_srgb [((t * inv) >> 19) & 0x7ff] << (8 * i);
            }
            return out;
        };

        /* H-filtered lanes of this column on tile row @lr */
        auto hrow = [&] (int lr) -> Lanes
        {
            const uint32_t *row = tile + lr * box_w;
            Lanes o;
            if (HF == F_COPY)
            {
                o = unpack (row [tap_ofs [0]]);
            }
            else
            {
                o = zero;
#pragma unroll
                for (int i = 0; i < (1 << NH); i++)
                {
                    Lanes a = unpack (row [tap_ofs [i]]);
                    Lanes b = unpack (row [tap_ofs [i] + 1]);
                    if (!OPAQUE && tap_pad [i])
                        b.l [3] = 0;
                    blend_add (o, a, b, tap_f [i], OPAQUE);
                }
#pragma unroll
                for (int i = 0; i < NL; i++)
                    o.l [i] >>= NH;
            }
            if (!OPAQUE && edge_col)
            {
                if (xr == 0)
                    o = weight (o, (uint32_t) p.h.first_opacity);
                if (xr == p.h.placement_size_px - 1)
                    o = weight (o, (uint32_t) p.h.last_opacity);
            }
            return o;
        };

        if (VF == F_COPY)
        {
            for (int lr = 0; lr < yr_b - yr_a; lr++)
            {
                box_wait (lr);
                put (pack (hrow (lr), RAW_H));
            }
            return;
        }

        /* The vertical samples of a column, in order: sample k of row yr blends tile rows
         * (ofs, ofs + 1) with weight f; it is complete when row ofs + 1 has been filtered.
         * Offsets never decrease and never skip a row.  Two rows per trip, so that the upper and
         * the lower row swap roles without a copy. */
        const int n_samples = (yr_b - yr_a) << NV;
        const bool v_edge = (yr_a == 0 && p.v.first_opacity < 256)
            || (yr_b == p.v.placement_size_px && p.v.last_opacity < 256);
        int yr = yr_a, k = 0;
        uint32_t pc = s_vsamples [0];
        Lanes acc = zero, ra = zero, rb = zero;

        auto consume = [&] (const Lanes &upper, const Lanes &lower, int lr) -> bool
        {
            while ((int) (pc & 0xffffu) + 1 == lr)
            {
                blend_add (acc, upper, lower, pc >> 16, RAW_H);
                k++;
                if ((k & ((1 << NV) - 1)) == 0)
                {
                    Lanes o;
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        o.l [i] = acc.l [i] >> NV;
                    if (!OPAQUE && v_edge)
                        o = apply_v_opacity (p, o, yr);
                    put (pack (o, false));
                    acc = zero;
                    yr++;
                    if (k == n_samples)
                        return true;
                }
                pc = s_vsamples [k];
            }
            return false;
        };

        for (int lr = 0;; lr += 2)
        {
            box_wait (lr);
            rb = hrow (lr);
            if (consume (ra, rb, lr))
                break;
            ra = hrow (lr + 1);
            if (consume (rb, ra, lr + 1))
                break;
        }
    };

    if (tile_opaque && HF == F_COPY && VF != F_COPY && s.inv16l [255] == 2048u)
    {
        /* Opaque tile, columns copied, rows blended: the walk in FP32.  (The reciprocal of alpha 255
         * is 2048 in the reference's table, which makes the un-premultiplication ((v << 8) * inv) >> 19
         * the identity on 11-bit values; checked rather than assumed.)  Every value on the way -- an
         * 11-bit linear channel, p f + q (256 - f) <= 2047 * 256, the sum of at most eight of those --
         * is an integer below 2^24, so FFMA computes it exactly, at twice the rate of the integer
         * multiply-add; the linearisation table is read as floats and the only conversions are the
         * three table indices of a finished pixel (floor (acc / 2^(NV + 8)): a power-of-two scaling,
         * exact, then truncation). */
        const float to_index = 1.0f / (float) (1 << (NV + 8));
        const int n_samples = (yr_b - yr_a) << NV;
        int k = 0;
        uint32_t pc = s_vsamples [0];
        float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;

        auto hrow_f = [&] (int lr, float &c0, float &c1, float &c2)
        {
            const uint32_t v = __byte_perm (tile [lr * box_w + tap_ofs [0]], 0, sel);
            c0 = s_linear_f [v & 0xff];
            c1 = s_linear_f [(v >> 8) & 0xff];
            c2 = s_linear_f [(v >> 16) & 0xff];
        };
        auto consume_f = [&] (float u0, float u1, float u2, float l0, float l1, float l2, int lr) -> bool
        {
            while ((int) (pc & 0xffffu) + 1 == lr)
            {
                const float f = (float) (pc >> 16), g = 256.0f - f;
                acc0 = fmaf (u0, f, fmaf (l0, g, acc0));
                acc1 = fmaf (u1, f, fmaf (l1, g, acc1));
                acc2 = fmaf (u2, f, fmaf (l2, g, acc2));
                k++;
                if ((k & ((1 << NV) - 1)) == 0)
                {
                    const uint32_t i0 = (uint32_t) (acc0 * to_index), i1 = (uint32_t) (acc1 * to_index),
                                   i2 = (uint32_t) (acc2 * to_index);
                    put (0xff000000u | (uint32_t) s.to_srgb [i0] | ((uint32_t) s.to_srgb [i1] << 8)
                         | ((uint32_t) s.to_srgb [i2] << 16));
                    acc0 = acc1 = acc2 = 0.0f;
                    if (k == n_samples)
                        return true;
                }
                pc = s_vsamples [k];
            }
            return false;
        };

        for (int lr = 0;; lr += 2)
        {
            box_wait (lr);
            hrow_f (lr, b0, b1, b2);
            if (consume_f (a0, a1, a2, b0, b1, b2, lr))
                break;
            hrow_f (lr + 1, a0, a1, a2);
            if (consume_f (b0, b1, b2, a0, a1, a2, lr + 1))
                break;
        }
    }
    else if (tile_opaque)
        walk (std::true_type ());
    else
        walk (std::false_type ());

    while (y < row1)                            /* rows below the placement */
        put (clear_pixel);
}

static std::atomic<unsigned long long> g_staged_launches { 0 };

typedef CUresult (*EncodeTiledFn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn
encode_tiled_fn ()
{
    static EncodeTiledFn fn = [] () -> EncodeTiledFn
    {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess
            || q != cudaDriverEntryPointSuccess)
        {
            cudaGetLastError ();
            return nullptr;
        }
        return (EncodeTiledFn) sym;
    } ();
    return fn;
}

/* Returns 0 when the launch does not apply (the caller falls back to the gathering kernel),
 * 1 when launched, -1 on a CUDA error (left for cudaGetLastError). */
static int
launch_scale_tma (const DevScale *p, cudaStream_t st)
{
    const ScaleTiles &tl = p->tiles;
    if (tl.tile_w <= 0 || tl.n_chunks <= 0 || tl.n_chunks > SCALE_TILE_MAX_CHUNKS)
        return 0;
    if ((((uintptr_t) p->src) & 15) != 0 || (p->src_rowstride & 15) != 0 || p->post.to_din99d)
        return 0;
    if (p->h.filter != F_COPY && p->h.n_halvings > 2)
        return 0;
    EncodeTiledFn encode = encode_tiled_fn ();
    if (!encode)
        return 0;

    CUtensorMap map;
    const cuuint64_t dims [2] = { (cuuint64_t) p->h.src_size_px, (cuuint64_t) p->v.src_size_px };
    const cuuint64_t strides [1] = { (cuuint64_t) p->src_rowstride };
    const cuuint32_t box [2] = { (cuuint32_t) tl.box_w, SCALE_TILE_CHUNK_ROWS };
    const cuuint32_t elem [2] = { 1, 1 };
    if (encode (&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *) p->src, dims, strides, box, elem,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;

    const size_t smem = (size_t) tl.n_chunks * SCALE_TILE_CHUNK_ROWS * tl.box_w * 4;
    dim3 grid ((unsigned) ((p->dest_w + tl.tile_w - 1) / tl.tile_w),
               (unsigned) ((p->n_rows + tl.rows_per_block - 1) / tl.rows_per_block));
    bool launched = false;
    cudaError_t err = cudaSuccess;

#define TMA_CASE(HF, VF) \
    if (!launched && p->h.filter == HF && p->v.filter == VF) \
    { \
        if (smem > 48 * 1024) \
            err = (cudaError_t) dev_opt_in_smem ((const void *) scale_tma_kernel<HF, VF>, smem); \
        if (err == cudaSuccess) \
            scale_tma_kernel<HF, VF><<<grid, tl.tile_w, smem, st>>> (*p, map); \
        launched = true; \
    }
#define TMA_ROW(HF) \
    TMA_CASE (HF, F_COPY) TMA_CASE (HF, F_BILINEAR_0H) TMA_CASE (HF, F_BILINEAR_1H) \
    TMA_CASE (HF, F_BILINEAR_2H) TMA_CASE (HF, F_BILINEAR_3H)
    TMA_ROW (F_COPY) TMA_ROW (F_BILINEAR_0H) TMA_ROW (F_BILINEAR_1H) TMA_ROW (F_BILINEAR_2H)
#undef TMA_ROW
#undef TMA_CASE
    if (!launched)
        return 0;
    if (err == cudaSuccess)
        g_staged_launches.fetch_add (1, std::memory_order_relaxed);
    return err == cudaSuccess ? 1 : -1;
}

unsigned long long
dev_scale_staged_launches ()
{
    return g_staged_launches.load ();
}

int
dev_launch_scale (const DevScale *p, void *stream)
{
    size_t n_px = (size_t) p->dest_w * (size_t) p->n_rows;
    if (n_px == 0)
        return 0;

    const int sms = dev_sm_count ();

    /* whole-canvas 1:1 placement of an unassociated 32-bit source with 16-byte aligned rows */
    bool copy4 = p->h.filter == F_COPY && p->v.filter == F_COPY && p->linear && !p->plain
        && p->src_pixel_type >= 4 && p->src_pixel_type <= 7
        && p->h.placement_ofs_px == 0 && p->h.placement_size_px == p->dest_w && (p->dest_w & 3) == 0
        && p->v.placement_ofs_px <= p->first_row
        && p->v.placement_ofs_px + p->v.placement_size_px >= p->first_row + p->n_rows
        && (p->src_rowstride & 15) == 0 && (((uintptr_t) p->src + (size_t) p->h.clip_before_px * 4) & 15) == 0
        && (((uintptr_t) p->dest) & 15) == 0;
    if (copy4)
    {
        unsigned strips4 = (unsigned) ((p->dest_w / 4 + 255) / 256);
        int row_groups = (p->n_rows + 3) / 4;
        unsigned rows4 = (unsigned) std::max (1, std::min (row_groups, (sms * 8 + (int) strips4 - 1) / (int) strips4));
        dim3 grid4 (strips4, rows4);
        cudaStream_t st4 = (cudaStream_t) stream;
        int t4 = p->src_pixel_type & 3;
        int dm4 = p->post_on ? p->post.dither_mode : 0;
        bool plain4 = p->composite == 0 && !(p->post_on && p->post.to_din99d) && (!p->post_on || dm4 == 1);
        if (plain4 && t4 == 0 && dm4 == 0) scale_copy4_kernel<0, 0><<<grid4, 256, 0, st4>>> (*p);
        else if (plain4 && t4 == 0 && dm4 == 1) scale_copy4_kernel<0, 1><<<grid4, 256, 0, st4>>> (*p);
        else if (plain4 && t4 == 1 && dm4 == 0) scale_copy4_kernel<1, 0><<<grid4, 256, 0, st4>>> (*p);
        else if (plain4 && t4 == 1 && dm4 == 1) scale_copy4_kernel<1, 1><<<grid4, 256, 0, st4>>> (*p);
        else scale_copy4_kernel<-1, -1><<<grid4, 256, 0, st4>>> (*p);
        dev_count_launch (1);
        return (int) cudaGetLastError ();
    }

    unsigned strips = (unsigned) ((p->dest_w + 255) / 256);
    unsigned row_blocks = (unsigned) std::max (1, std::min (p->n_rows, (sms * 8 + (int) strips - 1) / (int) strips));
    auto simple = [] (int f) { return f == F_COPY || (f >= F_BILINEAR_0H && f <= F_BILINEAR_3H); };
    bool spec = simple (p->h.filter) && simple (p->v.filter) && p->src_pixel_type >= 4 && p->src_pixel_type <= 7
        && p->linear == 1 && p->plain == 0 && p->composite == 0
        && p->h.n_halvings == (p->h.filter >= F_BILINEAR_0H ? p->h.filter - F_BILINEAR_0H : 0)
        && p->v.n_halvings == (p->v.filter >= F_BILINEAR_0H ? p->v.filter - F_BILINEAR_0H : 0);
    dim3 grid (strips, row_blocks);
    cudaStream_t st = (cudaStream_t) stream;
    bool launched = false;
    static const bool no_tma = getenv ("CHAFA_B200_SCALE_GATHER") != nullptr;     /* A/B measurements */
    if (spec && !no_tma)
    {
        int r = launch_scale_tma (p, st);
        if (r != 0)
        {
            dev_count_launch (1);
            return r < 0 ? (int) cudaErrorUnknown : (int) cudaGetLastError ();
        }
    }
    if (spec)
    {
#define SCALE_CASE(HF, VF) \
        if (!launched && p->h.filter == HF && p->v.filter == VF) \
        { \
            scale_kernel<HF, VF><<<grid, 256, 0, st>>> (*p); \
            launched = true; \
        }
#define SCALE_ROW(HF) \
        SCALE_CASE (HF, F_COPY) SCALE_CASE (HF, F_BILINEAR_0H) SCALE_CASE (HF, F_BILINEAR_1H) \
        SCALE_CASE (HF, F_BILINEAR_2H) SCALE_CASE (HF, F_BILINEAR_3H)
        SCALE_ROW (F_COPY) SCALE_ROW (F_BILINEAR_0H) SCALE_ROW (F_BILINEAR_1H) SCALE_ROW (F_BILINEAR_2H)
        SCALE_ROW (F_BILINEAR_3H)
#undef SCALE_ROW
#undef SCALE_CASE
    }
    if (!launched)
        scale_kernel<-1, -1><<<grid, 256, 0, st>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
