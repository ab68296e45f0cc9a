/* prep_pixel.cuh -- the elementwise half of the preprocessing pass (ordered / noise dither,
 * RGB -> DIN99d), shared by prep.cu (pass 2, in place) and scale.cu (applied to every pixel the
 * scaler writes when no whole-image statistic is needed, which saves one pass over the pixmap).
 *
 * Reference: chafa/internal/chafa-pixops.c:298-313, chafa-dither.c:106-144, chafa-color.c:83-168. */
#ifndef CHAFA_B200_PREP_PIXEL_CUH
#define CHAFA_B200_PREP_PIXEL_CUH

#include "device.h"

/* chafa-color.c:83-168 */
__device__ __noinline__ static uint32_t
rgb_to_din99d (uint32_t px)
{
    double rgbf [3], xyz [3], f3 [3], lab [3];

#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        double v = (double) ((px >> (8 * i)) & 0xff) / 255.0;
        rgbf [i] = v <= 0.04045 ? (v / 12.92) : pow ((v + 0.055) / 1.044, 2.4);
    }

    xyz [0] = 0.4124564 * rgbf [0] + 0.3575761 * rgbf [1] + 0.1804375 * rgbf [2];
    xyz [1] = 0.2126729 * rgbf [0] + 0.7151522 * rgbf [1] + 0.0721750 * rgbf [2];
    xyz [2] = 0.0193339 * rgbf [0] + 0.1191920 * rgbf [1] + 0.9503041 * rgbf [2];

    xyz [0] = 1.12 * xyz [0] - 0.12 * xyz [2];

    const double wp [3] = { 0.95047, 1.0, 1.08883 };
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        double v = xyz [i] / wp [i];
        f3 [i] = v > (216.0 / 24389.0) ? cbrt (v) : ((24389.0 / 27.0) * v + 16.0) / 116.0;
    }
    lab [0] = 116.0 * f3 [1] - 16.0;
    lab [1] = 500.0 * (f3 [0] - f3 [1]);
    lab [2] = 200.0 * (f3 [1] - f3 [2]);

    double adj_L = 325.22 * log (1.0 + 0.0036 * lab [0]);
    double ee = 0.6427876096865393 * lab [1] + 0.766044443118978 * lab [2];
    double f = 1.14 * (0.6427876096865393 * lab [2] - 0.766044443118978 * lab [1]);
    double G = sqrt (ee * ee + f * f);
    double C = 22.5 * log (1.0 + 0.06 * G);
    double h = atan2 (f, ee) + 0.8726646;
    while (h < 0.0) h += 6.283185;
    while (h > 6.283185) h -= 6.283185;

    /* double -> guint8 stores: truncation, as the compiled reference does for in-range values */
    uint32_t c0 = (uint32_t) (int) (adj_L * 2.5) & 0xff;
    uint32_t c1 = (uint32_t) (int) (C * cos (h) * 2.5 + 128.0) & 0xff;
    uint32_t c2 = (uint32_t) (int) (C * sin (h) * 2.5 + 128.0) & 0xff;
    return c0 | (c1 << 8) | (c2 << 16) | (px & 0xff000000u);
}

/* dither @px (already normalised) at pixmap position (x, y), then convert the colour space */
__device__ __forceinline__ uint32_t
prep_dither_convert (uint32_t px, int x, int y, const DevPrepPixel &q)
{
    if (q.dither_mode == 1 || q.dither_mode == 3)
    {
        int ti = (((y >> q.grain_h_shift) & q.tex_size_mask) << q.tex_size_shift)
            + ((x >> q.grain_w_shift) & q.tex_size_mask);
        int m0, m1, m2;

        if (q.dither_mode == 1)
        {
            m0 = m1 = m2 = (int) (short) __ldg (q.texture + ti);
        }
        else
        {
            m0 = (int) (short) __ldg (q.texture + ti * 3);
            m1 = (int) (short) __ldg (q.texture + ti * 3 + 1);
            m2 = (int) (short) __ldg (q.texture + ti * 3 + 2);
        }
        /* c = clamp (c + m, 0, 255) per colour channel, two channels per register in 16-bit lanes
         * (ch0 | ch2 << 16 and ch1 | alpha << 16; alpha gets + 0) */
        uint32_t rb = px & 0x00ff00ffu, ga = (px >> 8) & 0x00ff00ffu;
        rb = __vadd2 (rb, ((uint32_t) m0 & 0xffffu) | ((uint32_t) m2 << 16));
        ga = __vadd2 (ga, (uint32_t) m1 & 0xffffu);
        rb = __vmins2 (__vmaxs2 (rb, 0u), 0x00ff00ffu);
        ga = __vmins2 (__vmaxs2 (ga, 0u), 0x00ff00ffu);
        px = rb | (ga << 8);
    }

    if (q.to_din99d)
    {
        /* a pure function of the 24-bit colour: read it from the per-device table when there is one */
        if (q.din99d_lut)
            px = __ldg (q.din99d_lut + (px & 0x00ffffffu)) | (px & 0xff000000u);
        else
            px = rgb_to_din99d (px);
    }
    return px;
}

#endif /* CHAFA_B200_PREP_PIXEL_CUH */
