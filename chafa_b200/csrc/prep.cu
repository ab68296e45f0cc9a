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

#include <map>
#include <mutex>

#include "device.h"
#include "prep_pixel.cuh"

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

__global__ void __launch_bounds__ (256)
prep_pass2_kernel (DevPrep p)
{
    size_t n = (size_t) p.width * p.n_rows;
    uint32_t *base = p.pixels + (size_t) p.first_row * p.width;
    int hmin = 0, hmax = 0, factor = 0;
    bool do_norm = false;
    DevPrepPixel q;
    q.dither_mode = p.dither_mode;
    q.grain_w_shift = p.grain_w_shift;
    q.grain_h_shift = p.grain_h_shift;
    q.tex_size_shift = p.tex_size_shift;
    q.tex_size_mask = p.tex_size_mask;
    q.to_din99d = p.to_din99d;
    q.texture = p.texture;
    q.din99d_lut = p.din99d_lut;

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

        px = (px & 0xff000000u) | (uint32_t) c0 | ((uint32_t) c1 << 8) | ((uint32_t) c2 << 16);
        if (q.dither_mode == 1 || q.dither_mode == 3 || q.to_din99d)
            px = prep_dither_convert (px, (int) (i % (size_t) p.width), p.first_row + (int) (i / (size_t) p.width), q);

        base [i] = px;
    }
}

__global__ void __launch_bounds__ (256)
din99d_lut_kernel (uint32_t *lut)
{
    uint32_t rgb = blockIdx.x * blockDim.x + threadIdx.x;
    lut [rgb] = rgb_to_din99d (rgb) & 0x00ffffffu;
}

static std::mutex g_din_mutex;
static std::map<int, const uint32_t *> g_din_luts;

const uint32_t *
dev_din99d_lut (int device, void *stream)
{
    std::lock_guard<std::mutex> lock (g_din_mutex);
    auto it = g_din_luts.find (device);
    if (it != g_din_luts.end ())
        return it->second;

    uint32_t *lut = nullptr;
    if (cudaMalloc (&lut, (size_t) 4 << 24) != cudaSuccess)
    {
        cudaGetLastError ();
        g_din_luts [device] = nullptr;
        return nullptr;
    }
    din99d_lut_kernel<<<(1u << 24) / 256, 256, 0, (cudaStream_t) stream>>> (lut);
    dev_count_launch (1);
    if (cudaGetLastError () != cudaSuccess || cudaStreamSynchronize ((cudaStream_t) stream) != cudaSuccess)
    {
        cudaFree (lut);
        lut = nullptr;
    }
    g_din_luts [device] = lut;
    return lut;
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
