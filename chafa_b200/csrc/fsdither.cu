/* fsdither.cu -- Floyd-Steinberg error diffusion over dither grains, symbols mode.
 *
 * Restates fs_dither / fs_dither_grain (chafa/internal/chafa-pixops.c:210-282, 315-447).
 * The reference walks the grains of the pixmap serpentine -- even grain rows left to
 * right, odd rows right to left -- so the first grain of a row needs the error left by
 * the last grains of the previous one: the whole image is ONE dependency chain and
 * there is no row wavefront to exploit under byte identity (SURVEY.md finding 6).
 *
 * So this kernel is latency-bound by construction.  One warp walks the chain; within a
 * grain the lanes own the pixels (up to 8x8 = two per lane), the sums are warp
 * reductions, and the pen search runs on every lane redundantly (fixed palettes only:
 * truecolor symbols never dither).  The two error rows live in global scratch that
 * stays in L1.
 *
 * Arithmetic notes: the reference keeps errors in 16-bit integers, divides with
 * truncation towards zero, and scales the quantisation error by the float dither
 * intensity before truncating to 16 bits; all reproduced here (no FMA contraction). */

#include <cuda_runtime.h>

#include "device.h"
#include "palette.cuh"

#define FULL 0xffffffffu

struct Err
{
    short ch [4];
};

__device__ __forceinline__ void
err_add (Err *e, int c0, int c1, int c2)
{
    e->ch [0] = (short) (e->ch [0] + c0);
    e->ch [1] = (short) (e->ch [1] + c1);
    e->ch [2] = (short) (e->ch [2] + c2);
}

/* One grain.  err_in is read before any output is written: the outputs may alias
 * each other but never the input slot. */
__device__ __forceinline__ void
fs_grain (const DevFsDither &p, uint32_t *pixel, int lane, const Err *err_in,
          Err *out0, Err *out1, Err *out2, Err *out3)
{
    const int gw = 1 << p.grain_w_shift, gh = 1 << p.grain_h_shift;
    const int gshift = p.grain_w_shift + p.grain_h_shift;
    const int n_px = gw * gh;
    Err ein = *err_in;
    int acc [3] = { 0, 0, 0 }, over [3] = { 0, 0, 0 };

    for (int k = lane; k < n_px; k += 32)
    {
        uint32_t *q = pixel + (size_t) (k >> p.grain_w_shift) * p.width + (k & (gw - 1));
        uint32_t px = *q;
        uint32_t outpx = px & 0xff000000u;

#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            int v = (int) ((px >> (8 * i)) & 0xff) + ein.ch [i];
            if (v < 0)
            {
                over [i] += v;
                v = 0;
            }
            else if (v > 255)
            {
                over [i] += v - 255;
                v = 255;
            }
            acc [i] += v;
            outpx |= (uint32_t) v << (8 * i);
        }
        *q = outpx;
    }

    uint32_t acol = 0xff000000u;
    int mean [3];
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        acc [i] = __reduce_add_sync (FULL, acc [i]);
        over [i] = __reduce_add_sync (FULL, over [i]);
        mean [i] = (int) (short) acc [i] >> gshift;
        acol |= (uint32_t) (mean [i] & 0xff) << (8 * i);
    }

    int index = palette_lookup_nearest_warp (p.palette, p.color_space, acol, lane);
    uint32_t col = p.palette->colors [index];

    if (lane == 0)
    {
        int ne [3];
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            float q = __fmul_rn ((float) (mean [i] - (int) ((col >> (8 * i)) & 0xff)), p.intensity);
            float s = __fadd_rn ((float) ((int) (short) over [i] >> gshift), q);
            ne [i] = (int) (short) (int) s;
        }
        err_add (out0, ne [0] * 7 / 16, ne [1] * 7 / 16, ne [2] * 7 / 16);
        err_add (out1, ne [0] * 1 / 16, ne [1] * 1 / 16, ne [2] * 1 / 16);
        err_add (out2, ne [0] * 5 / 16, ne [1] * 5 / 16, ne [2] * 5 / 16);
        err_add (out3, ne [0] * 3 / 16, ne [1] * 3 / 16, ne [2] * 3 / 16);
    }
    __syncwarp ();
}

__global__ void __launch_bounds__ (32)
fs_dither_symbols_kernel (DevFsDither p)
{
    const int lane = threadIdx.x;
    const int gw = 1 << p.grain_w_shift;
    const int wg = p.width >> p.grain_w_shift;
    const int rows = p.height >> p.grain_h_shift;
    Err *row0 = (Err *) p.scratch + 1;
    Err *row1 = (Err *) p.scratch + (wg + 2) + 1;

    for (int i = lane; i < (wg + 2) * 2; i += 32)
    {
        Err z = { { 0, 0, 0, 0 } };
        ((Err *) p.scratch) [i] = z;
    }
    __syncwarp ();

    for (int y = 0; y < rows; y++)
    {
        for (int i = lane; i < wg + 2; i += 32)
        {
            Err z = { { 0, 0, 0, 0 } };
            (row1 - 1) [i] = z;
        }
        __syncwarp ();

        uint32_t *line = p.pixels + (size_t) (y << p.grain_h_shift) * p.width;

        if (!(y & 1))
        {
            fs_grain (p, line, lane, row0, row0 + 1, row1 + 1, row1, row1 + 1);
            if (wg >= 2)
            {
                int x;
                for (x = 1; ((x + 1) << p.grain_w_shift) < p.width; x++)
                    fs_grain (p, line + x * gw, lane, row0 + x, row0 + x + 1, row1 + x + 1, row1 + x, row1 + x - 1);
                fs_grain (p, line + x * gw, lane, row0 + x, row1 + x, row1 + x, row1 + x - 1, row1 + x - 1);
            }
        }
        else
        {
            if (wg >= 2)
            {
                fs_grain (p, line + (wg - 1) * gw, lane, row0 + wg - 1, row0 + wg - 2, row1 + wg - 2,
                          row1 + wg - 1, row1 + wg - 2);
                for (int x = wg - 2; x > 0; x--)
                    fs_grain (p, line + x * gw, lane, row0 + x, row0 + x - 1, row1 + x - 1, row1 + x, row1 + x + 1);
            }
            fs_grain (p, line, lane, row0, row1, row1, row1 + 1, row1 + 1);
        }

        Err *t = row0;
        row0 = row1;
        row1 = t;
    }
}

size_t
dev_fs_dither_symbols_scratch_bytes (int width, int grain_w_shift)
{
    return ((size_t) (width >> grain_w_shift) + 2) * 2 * sizeof (Err);
}

int
dev_launch_fs_dither_symbols (const DevFsDither *p, void *stream)
{
    if (p->width <= 0 || p->height <= 0)
        return 0;
    fs_dither_symbols_kernel<<<1, 32, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
