/* prep.cu -- kernel K2: in-place preprocessing of the canvas pixmap.
 *
 * Pass 1 (chafa/internal/chafa-pixops.c:477-530): optional saturation boost, 2048-bin
 * intensity histogram (3R + 4G + B) of the pixels with alpha > 127.
 * Bounds (pixops.c:110-140): crop percentiles of the histogram, one thread.
 * Pass 2 (pixops.c:568-625): intensity normalisation (142-190), ordered / noise dither
 * (298-313; chafa-dither.c:106-144), RGB -> DIN99d (chafa-color.c:83-168, in double as
 * the reference).  Floyd-Steinberg diffusion is a serial chain and lives in fsdither.cu.
 *
 * All three are streaming, HBM-bound passes over Wpx * Hpx * 4 bytes. */

#include <cuda_runtime.h>

#include "device.h"

#define INTENSITY_MAX 2048
#define FIXED_MULT 4096

__device__ __forceinline__ uint32_t
boost_saturation (uint32_t px)
{
    float r = (float) (px & 0xff), g = (float) ((px >> 8) & 0xff), b = (float) ((px >> 16) & 0xff);
    float P = __fsqrt_rn (__fadd_rn (__fadd_rn (__fmul_rn (__fmul_rn (r, r), .299f),
                                                __fmul_rn (__fmul_rn (g, g), .587f)),
                                     __fmul_rn (__fmul_rn (b, b), .144f)));
    int c0 = (int) __fadd_rn (P, __fmul_rn (__fsub_rn (r, P), 2.0f));
    int c1 = (int) __fadd_rn (P, __fmul_rn (__fsub_rn (g, P), 2.0f));
    int c2 = (int) __fadd_rn (P, __fmul_rn (__fsub_rn (b, P), 2.0f));
    c0 = min (max (c0, 0), 255);
    c1 = min (max (c1, 0), 255);
    c2 = min (max (c2, 0), 255);
    return (px & 0xff000000u) | (uint32_t) c0 | ((uint32_t) c1 << 8) | ((uint32_t) c2 << 16);
}

__global__ void __launch_bounds__ (256)
prep_pass1_kernel (DevPrep p)
{
    __shared__ int s_hist [INTENSITY_MAX];
    __shared__ int s_count;

    if (p.build_histogram)
    {
        for (int i = threadIdx.x; i < INTENSITY_MAX; i += blockDim.x)
            s_hist [i] = 0;
        if (threadIdx.x == 0)
            s_count = 0;
        __syncthreads ();
    }

    size_t n = (size_t) p.width * p.n_rows;
    uint32_t *base = p.pixels + (size_t) p.first_row * p.width;
    int local_count = 0;

    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    {
        uint32_t px = base [i];

        if (p.boost_saturation)
        {
            px = boost_saturation (px);
            base [i] = px;
        }
        if (p.build_histogram && (px >> 24) > 127)
        {
            int v = (int) (px & 0xff) * 3 + (int) ((px >> 8) & 0xff) * 4 + (int) ((px >> 16) & 0xff);
            atomicAdd (&s_hist [v], 1);
            local_count++;
        }
    }

    if (p.build_histogram)
    {
        atomicAdd (&s_count, local_count);
        __syncthreads ();
        for (int i = threadIdx.x; i < INTENSITY_MAX; i += blockDim.x)
            if (s_hist [i])
                atomicAdd (&p.hist [i], s_hist [i]);
        if (threadIdx.x == 0 && s_count)
            atomicAdd (&p.hist [INTENSITY_MAX], s_count);
    }
}

__global__ void
prep_bounds_kernel (DevPrep p)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;

    long long n_samples = p.hist [INTENSITY_MAX];
    long long pixels_crop = (n_samples * (((long long) p.crop_pct * 1024) / 100)) / 1024;
    int i, t;

    for (i = 0, t = (int) pixels_crop; i < INTENSITY_MAX; i++)
    {
        t -= p.hist [i];
        if (t <= 0)
            break;
    }
    p.bounds [0] = i;

    for (i = INTENSITY_MAX - 1, t = (int) pixels_crop; i >= 0; i--)
    {
        t -= p.hist [i];
        if (t <= 0)
            break;
    }
    p.bounds [1] = i;
}

__device__ __forceinline__ int
normalize_ch (int v, int vmin, int factor)
{
    int vt = (v - vmin) * factor;
    vt /= FIXED_MULT;
    return min (max (vt, 0), 255);
}

/* chafa-color.c:83-168 */
__device__ static uint32_t
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

__global__ void __launch_bounds__ (256)
prep_pass2_kernel (DevPrep p)
{
    size_t n = (size_t) p.width * p.n_rows;
    uint32_t *base = p.pixels + (size_t) p.first_row * p.width;
    int hmin = 0, hmax = 0, factor = 0;
    bool do_norm = false;

    if (p.normalize)
    {
        hmin = p.bounds [0];
        hmax = p.bounds [1];
        if (hmin != hmax)
        {
            do_norm = true;
            factor = ((INTENSITY_MAX - 1) * FIXED_MULT) / (hmax - hmin);
        }
    }

    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    {
        uint32_t px = base [i];
        int c0 = (int) (px & 0xff), c1 = (int) ((px >> 8) & 0xff), c2 = (int) ((px >> 16) & 0xff);

        if (do_norm)
        {
            c0 = normalize_ch (c0, hmin / 8, factor);
            c1 = normalize_ch (c1, hmin / 8, factor);
            c2 = normalize_ch (c2, hmin / 8, factor);
        }

        if (p.dither_mode == 1 || p.dither_mode == 3)
        {
            int y = p.first_row + (int) (i / (size_t) p.width);
            int x = (int) (i % (size_t) p.width);
            int ti = (((y >> p.grain_h_shift) & p.tex_size_mask) << p.tex_size_shift)
                + ((x >> p.grain_w_shift) & p.tex_size_mask);

            if (p.dither_mode == 1)
            {
                int m = (int) (short) p.texture [ti];
                c0 = min (max (c0 + m, 0), 255);
                c1 = min (max (c1 + m, 0), 255);
                c2 = min (max (c2 + m, 0), 255);
            }
            else
            {
                c0 = min (max (c0 + (int) (short) p.texture [ti * 3], 0), 255);
                c1 = min (max (c1 + (int) (short) p.texture [ti * 3 + 1], 0), 255);
                c2 = min (max (c2 + (int) (short) p.texture [ti * 3 + 2], 0), 255);
            }
        }

        px = (px & 0xff000000u) | (uint32_t) c0 | ((uint32_t) c1 << 8) | ((uint32_t) c2 << 16);

        if (p.to_din99d)
            px = rgb_to_din99d (px);

        base [i] = px;
    }
}

static unsigned
grid_for (size_t n)
{
    int dev = 0, sms = 148;
    cudaGetDevice (&dev);
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t) sms * 8)
        blocks = (size_t) sms * 8;
    return (unsigned) (blocks ? blocks : 1);
}

int
dev_launch_prep_pass1 (const DevPrep *p, void *stream)
{
    size_t n = (size_t) p->width * p->n_rows;
    if (!n || (!p->boost_saturation && !p->build_histogram))
        return 0;
    prep_pass1_kernel<<<grid_for (n), 256, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_prep_bounds (const DevPrep *p, void *stream)
{
    prep_bounds_kernel<<<1, 32, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_prep_pass2 (const DevPrep *p, void *stream)
{
    size_t n = (size_t) p->width * p->n_rows;
    if (!n)
        return 0;
    prep_pass2_kernel<<<grid_for (n), 256, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
