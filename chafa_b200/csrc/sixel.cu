/* sixel.cu -- kernels of the sixel pixel mode.
 *
 *   K6a  the generated palette is built in palette_build.cu
 *   K6b  per-pixel pen lookup, no / ordered / noise dither
 *        (chafa/internal/chafa-indexed-image.c:52-170; chafa-color-table.c:315-390)
 *   K6c  Floyd-Steinberg diffusion, per pixel, serpentine starting right-to-left on
 *        row 0 (chafa-indexed-image.c:172-304; chafa-palette.c:1039-1076)
 *   K6d  band encoder: six rows of pens -> per-pen run-length coded sixel text
 *        (chafa/internal/chafa-sixel-renderer.c:128-373)
 *
 * The dynamic palette's pen search is NOT a nearest-colour search: it is a binary search
 * on the first principal component followed by a two-way scan that stops as soon as the
 * squared distance along that component exceeds the best full distance, with ties going
 * to the entry scanned last, and with one read past the end of the table when the
 * colour projects beyond the last entry.  The walk is replayed exactly here, entry by
 * entry, from a copy of the table in shared memory.
 *
 * K6c is one dependency chain per image: a single thread walks it.
 * Throughput comes from many images in flight on separate streams, not from one. */

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "device.h"
#include "palette.cuh"

#define FULL 0xffffffffu

static unsigned grid_for (size_t n, int per_block);

/* ------------------------------------------------------------------------ */
/* Pen search                                                                 */

struct TableView
{
    const DevColorTableEntry *entries;   /* n_entries + 1 (a zero record follows the last) */
    const uint32_t *pens;
    int n_entries;
    int ev [2] [3];
    int avg [3];
    int mul [2];
};

__device__ __forceinline__ int
table_color_diff (uint32_t a, uint32_t b)
{
    int n = ((int) (b & 0xff) - (int) (a & 0xff)) * 32;
    int diff = n * n;
    n = ((int) ((b >> 8) & 0xff) - (int) ((a >> 8) & 0xff)) * 32;
    diff += n * n;
    n = ((int) ((b >> 16) & 0xff) - (int) ((a >> 16) & 0xff)) * 32;
    diff += n * n;
    return diff;
}

__device__ __forceinline__ int
table_project (const int *v, const int *axis, int mul)
{
    long long d = (long long) v [0] * axis [0] + (long long) v [1] * axis [1] + (long long) v [2] * axis [2];
    return (int) ((d * mul) / ((1 << 14) / 32));
}

__device__ __forceinline__ bool
table_refine (const TableView &t, uint32_t want, const int *v, int j, int &best_pen, long long &best_diff)
{
    const DevColorTableEntry e = t.entries [j];
    long long a = (long long) e.v [0] - v [0];
    a *= a;
    if (a > best_diff)
        return false;

    long long b = (long long) e.v [1] - v [1];
    b *= b;
    if (b <= best_diff)
    {
        long long d = table_color_diff (t.pens [e.pen & 255], want);
        if (d <= best_diff)
        {
            best_pen = j;
            best_diff = d;
        }
    }
    return true;
}

__device__ int
table_find_nearest_pen (const TableView &t, uint32_t want)
{
    int p [3] = { (int) (want & 0xff) * 32 - t.avg [0], (int) ((want >> 8) & 0xff) * 32 - t.avg [1],
                  (int) ((want >> 16) & 0xff) * 32 - t.avg [2] };
    int v [2] = { table_project (p, t.ev [0], t.mul [0]), table_project (p, t.ev [1], t.mul [1]) };
    int i = 0, j = t.n_entries;

    while (i != j)
    {
        int n = i + (j - i) / 2;
        if (v [0] > t.entries [n].v [0])
            i = n + 1;
        else
            j = n;
    }

    const int m = j;
    long long best_diff = 0x7fffffffffffffffll;
    int best_pen = 0;

    for (j = m; j >= 0; j--)
        if (!table_refine (t, want, v, j, best_pen, best_diff))
            break;
    for (j = m + 1; j < t.n_entries; j++)
        if (!table_refine (t, want, v, j, best_pen, best_diff))
            break;

    return t.entries [best_pen].pen;
}

struct SharedTable
{
    DevColorTableEntry entries [257];
    uint32_t pens [256];
};

__device__ __forceinline__ void
load_table (SharedTable &s, TableView &t, const DevColorTable *g)
{
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
    {
        s.entries [i] = g->entries [i];
        s.pens [i] = g->pens [i];
    }
    if (threadIdx.x == 0)
    {
        s.entries [256].v [0] = s.entries [256].v [1] = 0;
        s.entries [256].pen = 0;
    }
    t.entries = s.entries;
    t.pens = s.pens;
    t.n_entries = g->n_entries;
#pragma unroll
    for (int k = 0; k < 2; k++)
    {
        t.mul [k] = g->eigen_mul [k];
#pragma unroll
        for (int c = 0; c < 3; c++)
            t.ev [k] [c] = g->eigenvectors [k] [c];
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
        t.avg [c] = g->average [c];
    __syncthreads ();
}

/* chafa-color.c:83-168, as in prep.cu */
__device__ uint32_t sixel_rgb_to_din99d (uint32_t px);

/* Pen of one colour: transparency, then the dynamic table or the fixed pickers.
 * Returns the palette index before the first-colour offset is removed. */
__device__ __forceinline__ int
lookup_pen (const DevSixel &p, const TableView &t, uint32_t color)
{
    if (p.dynamic)
    {
        if ((int) (color >> 24) < p.alpha_threshold)
            return p.transparent_index;
        return table_find_nearest_pen (t, color & 0xffffffu);
    }
    return palette_lookup_nearest (p.fixed_palette, p.color_space, color, nullptr);
}

/* Table index of a colour.  Neighbouring pixels of an image have neighbouring colours, so
 * the table is laid out in bricks of 8 x 4 x 4 colours (one 128-byte line each): the
 * chain's loads then mostly hit the L1 of the SM it runs on instead of going to L2. */
__device__ __forceinline__ uint32_t
lut_index (uint32_t color)
{
    uint32_t c0 = color & 0xff, c1 = (color >> 8) & 0xff, c2 = (color >> 16) & 0xff;
    uint32_t brick = (c0 >> 3) | ((c1 >> 2) << 5) | ((c2 >> 2) << 11);
    return (brick << 7) | (c0 & 7u) | ((c1 & 3u) << 3) | ((c2 & 3u) << 5);
}

/* Pen of every 24-bit colour, for the diffusion chain: the pen search is a pure function
 * of the colour once the palette exists, so all 2^24 answers are computed in parallel
 * (a few ms) and the serial chain pays one dependent load per pixel instead of a table
 * walk.  Indexed by ch0 | ch1 << 8 | ch2 << 16 of an opaque colour in the canvas's
 * colour space. */
__global__ void __launch_bounds__ (256)
sixel_lut_kernel (DevSixel p)
{
    __shared__ SharedTable s_table;
    TableView t;

    if (p.dynamic)
        load_table (s_table, t, p.table);

    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < (1u << 24); c += gridDim.x * blockDim.x)
        p.lut [lut_index (c)] = (uint8_t) lookup_pen (p, t, c | 0xff000000u);
}

/* ------------------------------------------------------------------------ */
/* K6b                                                                        */

__global__ void __launch_bounds__ (256)
sixel_quantise_kernel (DevSixel p)
{
    __shared__ SharedTable s_table;
    TableView t;

    if (p.dynamic)
        load_table (s_table, t, p.table);

    const size_t n_img = (size_t) p.width * p.image_height;
    const size_t n_px = (size_t) p.width * p.height;

    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_img; i += (size_t) gridDim.x * blockDim.x)
    {
        if (i >= n_px)
        {
            p.indexed [i] = (uint8_t) p.transparent_index;      /* rows added to reach a multiple of six */
            continue;
        }

        uint32_t px = p.pixels [i];
        int index;

        if (p.dither_mode == 1 || p.dither_mode == 3)
        {
            int y = (int) (i / (size_t) p.width), x = (int) (i % (size_t) p.width);
            int ti = (((y >> p.grain_h_shift) & p.tex_size_mask) << p.tex_size_shift)
                + ((x >> p.grain_w_shift) & p.tex_size_mask);
            int c [3];
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                int m = p.dither_mode == 1 ? p.texture [ti] : p.texture [ti * 3 + k];
                c [k] = min (max ((int) ((px >> (8 * k)) & 0xff) + (int) (short) m, 0), 255);
            }
            px = (px & 0xff000000u) | (uint32_t) c [0] | ((uint32_t) c [1] << 8) | ((uint32_t) c [2] << 16);
        }

        if ((int) (px >> 24) < p.alpha_threshold)
        {
            index = p.transparent_index;
        }
        else
        {
            if (p.color_space != 0)
                px = sixel_rgb_to_din99d (px);
            index = lookup_pen (p, t, px) - p.first_color;
        }
        p.indexed [i] = (uint8_t) index;
    }
}

/* ------------------------------------------------------------------------ */
/* K6b': the reference's lookup cache, replayed exactly.
 *
 * Without diffusion the reference answers a pixel from a 16384-slot direct-mapped cache
 * keyed on the colour with the low bit of every channel dropped, and on a miss looks up
 * the pixel's full colour and stores the pen under the reduced key
 * (chafa-indexed-image.c:52-82, chafa-color-hash.h:35-62).  A pixel therefore gets the
 * pen of the FIRST pixel of the current run of equal reduced keys among the pixels that
 * map to its slot, in scan order, within its thread batch (one cache per batch,
 * chafa-indexed-image.c:306-334).  Transparent pixels bypass the cache.
 *
 * Replayed in parallel: stable radix sort of the opaque pixels by (batch, slot) keeps scan
 * order inside each group; a pixel heads a run if its group or reduced key differs from
 * its predecessor's; a max-scan carries the head's position forward; every pixel then
 * copies the exact pen of its run's head. */

struct MaxU32
{
    __host__ __device__ uint32_t operator() (uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__device__ __forceinline__ uint32_t
dither_pixel (const DevSixel &p, uint32_t px, int x, int y)
{
    if (p.dither_mode != 1 && p.dither_mode != 3)
        return px;
    int ti = (((y >> p.grain_h_shift) & p.tex_size_mask) << p.tex_size_shift)
        + ((x >> p.grain_w_shift) & p.tex_size_mask);
    int c [3];
#pragma unroll
    for (int k = 0; k < 3; k++)
    {
        int m = p.dither_mode == 1 ? p.texture [ti] : p.texture [ti * 3 + k];
        c [k] = min (max ((int) ((px >> (8 * k)) & 0xff) + (int) (short) m, 0), 255);
    }
    return (px & 0xff000000u) | (uint32_t) c [0] | ((uint32_t) c [1] << 8) | ((uint32_t) c [2] << 16);
}

__global__ void __launch_bounds__ (256)
sixel_cache_keys_kernel (DevSixel p, const uint16_t *__restrict__ row_batch, uint32_t *__restrict__ sort_keys,
                         uint32_t *__restrict__ idx, uint32_t *__restrict__ color_keys)
{
    const size_t n_px = (size_t) p.width * p.height;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += (size_t) gridDim.x * blockDim.x)
    {
        int y = (int) (i / (size_t) p.width), x = (int) (i % (size_t) p.width);
        uint32_t px = dither_pixel (p, p.pixels [i], x, y);
        uint32_t key = px & 0x00fefefeu;
        uint32_t slot = (key ^ (key >> 7) ^ (key >> 14)) & 16383u;
        bool opaque = (int) (px >> 24) >= p.alpha_threshold;
        sort_keys [i] = opaque ? (((uint32_t) row_batch [y] << 14) | slot) : 0xffffffffu;
        idx [i] = (uint32_t) i;
        color_keys [i] = key;
    }
}

__global__ void __launch_bounds__ (256)
sixel_cache_heads_kernel (size_t n, const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ sorted_idx,
                          const uint32_t *__restrict__ color_keys, uint32_t *__restrict__ head_pos)
{
    for (size_t j = (size_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t) gridDim.x * blockDim.x)
    {
        bool head = j == 0 || sorted_keys [j] != sorted_keys [j - 1]
            || color_keys [sorted_idx [j]] != color_keys [sorted_idx [j - 1]];
        head_pos [j] = head ? (uint32_t) j : 0u;
    }
}

__global__ void __launch_bounds__ (256)
sixel_cache_gather_kernel (size_t n, const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ sorted_idx,
                           const uint32_t *__restrict__ head_pos, const uint8_t *__restrict__ exact,
                           uint8_t *__restrict__ out)
{
    for (size_t j = (size_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t) gridDim.x * blockDim.x)
    {
        uint32_t i = sorted_idx [j];
        /* transparent pixels keep their own (transparent) pen */
        out [i] = sorted_keys [j] == 0xffffffffu ? exact [i] : exact [sorted_idx [head_pos [j]]];
    }
}

size_t
dev_sixel_cache_scratch_bytes (size_t n_px)
{
    size_t sort_tmp = 0, scan_tmp = 0;
    cub::DeviceRadixSort::SortPairs (nullptr, sort_tmp, (const uint32_t *) nullptr, (uint32_t *) nullptr,
                                     (const uint32_t *) nullptr, (uint32_t *) nullptr, (int) n_px);
    cub::DeviceScan::InclusiveScan (nullptr, scan_tmp, (const uint32_t *) nullptr, (uint32_t *) nullptr, MaxU32 (),
                                    (int) n_px);
    size_t tmp = (sort_tmp > scan_tmp ? sort_tmp : scan_tmp) + 256;
    return n_px * 4 * 6 + tmp + 1024;
}

/* @exact: pens from the exact per-pixel lookup (K6b); @out: pens as the reference's cached
 * lookup produces them.  @scratch: dev_sixel_cache_scratch_bytes (). */
int
dev_launch_sixel_cache_replay (const DevSixel *p, const uint16_t *d_row_batch, const uint8_t *exact, uint8_t *out,
                               void *scratch, size_t scratch_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    size_t n = (size_t) p->width * p->height;
    if (!n)
        return 0;

    uint32_t *keys_in = (uint32_t *) scratch, *keys_out = keys_in + n, *idx_in = keys_out + n, *idx_out = idx_in + n,
             *color_keys = idx_out + n, *head_pos = color_keys + n;
    void *tmp = (void *) (((uintptr_t) (head_pos + n) + 255) & ~(uintptr_t) 255);
    size_t tmp_bytes = scratch_bytes - (size_t) ((uint8_t *) tmp - (uint8_t *) scratch);

    sixel_cache_keys_kernel<<<grid_for (n, 256), 256, 0, st>>> (*p, d_row_batch, keys_in, idx_in, color_keys);
    size_t need = tmp_bytes;
    cudaError_t err = cub::DeviceRadixSort::SortPairs (tmp, need, keys_in, keys_out, idx_in, idx_out, (int) n, 0, 32, st);
    if (err != cudaSuccess)
        return (int) err;
    sixel_cache_heads_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, keys_out, idx_out, color_keys, head_pos);
    need = tmp_bytes;
    err = cub::DeviceScan::InclusiveScan (tmp, need, head_pos, head_pos, MaxU32 (), (int) n, st);
    if (err != cudaSuccess)
        return (int) err;
    sixel_cache_gather_kernel<<<grid_for (n, 256), 256, 0, st>>> (n, keys_out, idx_out, head_pos, exact, out);
    /* rows added to reach a multiple of six */
    size_t n_img = (size_t) p->width * p->image_height;
    if (n_img > n)
        cudaMemsetAsync (out + n, p->transparent_index, n_img - n, st);
    dev_count_launch (5);
    return (int) cudaGetLastError ();
}

/* ------------------------------------------------------------------------ */
/* K6c                                                                        */

struct Err16
{
    short ch [4];
};

__device__ __forceinline__ void
fs_spread (Err16 *out, const short *err, int factor, float intensity)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        float v = __fadd_rn ((float) out->ch [i], __fmul_rn ((float) ((int) err [i] * factor), intensity));
        out->ch [i] = (short) (int) v;
    }
}

__device__ __forceinline__ uint8_t
fs_pixel (const DevSixel &p, const TableView &t, uint32_t px, Err16 error_in,
          Err16 *out0, Err16 *out1, Err16 *out2, Err16 *out3)
{
    int index;
    short err [3] = { 0, 0, 0 };

    if ((int) (px >> 24) < p.alpha_threshold)
    {
        index = p.transparent_index;
    }
    else
    {
        if (p.color_space != 0 && !p.preconverted)
            px = sixel_rgb_to_din99d (px);

        short comp [3];
        uint32_t color = px & 0xff000000u;
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
            float e = __fdiv_rn (__fmul_rn ((float) error_in.ch [i], 0.9f), 16.0f);
            comp [i] = (short) (int) __fadd_rn ((float) (int) ((px >> (8 * i)) & 0xff), e);
            color |= (uint32_t) min (max ((int) comp [i], 0), 255) << (8 * i);
        }

        index = p.lut ? (int) p.lut [lut_index (color)] : lookup_pen (p, t, color);
        if (index != p.transparent_index)
        {
            uint32_t found = p.pal_colors [index];
#pragma unroll
            for (int i = 0; i < 3; i++)
                err [i] = (short) ((int) comp [i] - (int) ((found >> (8 * i)) & 0xff));
        }
        index -= p.first_color;
    }

    /* the four targets may alias at the row ends: strictly in this order */
    fs_spread (out0, err, 7, p.intensity);
    fs_spread (out1, err, 1, p.intensity);
    fs_spread (out2, err, 5, p.intensity);
    fs_spread (out3, err, 3, p.intensity);
    return (uint8_t) index;
}

/* The same chain with everything off the critical path taken off it: pens come from the
 * 2^24-entry table, the source pixels are already in the canvas's colour space, the three
 * next-row accumulators that a pixel touches live in registers (each next-row element is
 * written to memory once, when its last contribution has arrived), and the previous
 * row's error for the next pixel is fetched before it is needed.  What remains per pixel
 * is: compensate -> clamp -> one table load -> palette colour -> error -> carry.
 *
 * In processing order (k = 0 .. w-1, x = k on odd rows, w-1-k on even rows) a pixel
 * sends 7/16 to the next pixel of its row and 1/16, 5/16, 3/16 to the next row ahead of,
 * at, and behind its own position; the first pixel of a row has nothing behind it and
 * sends its 3/16 ahead, the last one has nothing ahead and sends its 7/16 and 1/16
 * straight down and its 5/16 and 3/16 behind (chafa-indexed-image.c:207-266).  Every
 * addition is float and truncates to 16 bits, so the order of contributions to one
 * element is kept: it is the processing order. */
struct Acc3
{
    short ch [3];
};

/* a += err * factor * intensity in float, truncated to 16 bits.  At intensity 1 (the
 * default) every term is a small integer, the float sum is exact and the truncation a
 * no-op: plain integer arithmetic gives the same bits without four conversions on the
 * critical path. */
template <bool UNIT>
__device__ __forceinline__ void
acc_add (Acc3 &a, const int *err, int factor, float intensity)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        if (UNIT)
            a.ch [i] = (short) ((int) a.ch [i] + err [i] * factor);
        else
            a.ch [i] = (short) (int) __fadd_rn ((float) a.ch [i], __fmul_rn ((float) (err [i] * factor), intensity));
    }
}

template <bool UNIT>
__global__ void __launch_bounds__ (32)
sixel_fs_fast_kernel (DevSixel p)
{
    __shared__ uint32_t s_colors [256];

    for (int i = threadIdx.x; i < 256; i += 32)
        s_colors [i] = p.pal_colors [i];

    const int w = p.width;
    Err16 *row0 = (Err16 *) p.scratch, *row1 = row0 + w;
    const Err16 zero = { { 0, 0, 0, 0 } };

    for (int i = threadIdx.x; i < w; i += 32)
        row0 [i] = zero;
    for (size_t i = (size_t) w * p.height + threadIdx.x; i < (size_t) w * p.image_height; i += 32)
        p.indexed [i] = (uint8_t) p.transparent_index;
    __syncwarp ();

    if (threadIdx.x != 0)
        return;

    const float intensity = p.intensity;

    for (int y = 0; y < p.height; y++)
    {
        const int dir = (y & 1) ? 1 : -1;
        const int x0 = (y & 1) ? 0 : w - 1;
        const uint32_t *in = p.pixels + (size_t) y * w;
        uint8_t *out = p.indexed + (size_t) y * w;

        Acc3 a_prev = { { 0, 0, 0 } }, a_cur = { { 0, 0, 0 } }, a_next = { { 0, 0, 0 } };
        Err16 e_in = row0 [x0];
        Err16 e_ahead = row0 [x0 + dir];
        uint32_t px = in [x0];
        uint32_t px_ahead = in [x0 + dir];

        for (int k = 0; k < w; k++)
        {
            const int x = x0 + k * dir;
            const bool first = k == 0, last = k == w - 1;

            /* prefetch for pixel k + 2 (independent of the chain) */
            Err16 e_ahead2 = zero;
            uint32_t px_ahead2 = 0;
            if (k + 2 < w)
            {
                e_ahead2 = row0 [x + 2 * dir];
                px_ahead2 = in [x + 2 * dir];
            }

            int err [3] = { 0, 0, 0 };
            int index;

            if ((int) (px >> 24) < p.alpha_threshold)
            {
                index = p.transparent_index;
            }
            else
            {
                int comp [3];
                uint32_t color = 0;
#pragma unroll
                for (int i = 0; i < 3; i++)
                {
                    float e = __fmul_rn (__fmul_rn ((float) e_in.ch [i], 0.9f), 0.0625f);
                    comp [i] = (int) (short) (int) __fadd_rn ((float) (int) ((px >> (8 * i)) & 0xff), e);
                    color |= (uint32_t) min (max (comp [i], 0), 255) << (8 * i);
                }
                index = (int) p.lut [lut_index (color)];
                if (index != p.transparent_index)
                {
                    uint32_t found = s_colors [index & 255];
#pragma unroll
                    for (int i = 0; i < 3; i++)
                        err [i] = (int) (short) (comp [i] - (int) ((found >> (8 * i)) & 0xff));
                }
                index -= p.first_color;
            }
            out [x] = (uint8_t) index;

            if (!last)
            {
                /* 7/16 to the next pixel of this row: its incoming error, on top of what the
                 * previous row left there */
                Acc3 carry = { { e_ahead.ch [0], e_ahead.ch [1], e_ahead.ch [2] } };
                acc_add<UNIT> (carry, err, 7, intensity);
                acc_add<UNIT> (a_next, err, 1, intensity);
                acc_add<UNIT> (a_cur, err, 5, intensity);
                if (first)
                    acc_add<UNIT> (a_next, err, 3, intensity);
                else
                    acc_add<UNIT> (a_prev, err, 3, intensity);

                if (!first)
                {
                    Err16 done = { { a_prev.ch [0], a_prev.ch [1], a_prev.ch [2], 0 } };
                    row1 [x - dir] = done;
                }
                a_prev = a_cur;
                a_cur = a_next;
                a_next.ch [0] = a_next.ch [1] = a_next.ch [2] = 0;

                e_in.ch [0] = carry.ch [0];
                e_in.ch [1] = carry.ch [1];
                e_in.ch [2] = carry.ch [2];
                e_ahead = e_ahead2;
                px = px_ahead;
                px_ahead = px_ahead2;
            }
            else
            {
                acc_add<UNIT> (a_cur, err, 7, intensity);
                acc_add<UNIT> (a_cur, err, 1, intensity);
                acc_add<UNIT> (a_prev, err, 5, intensity);
                acc_add<UNIT> (a_prev, err, 3, intensity);
                Err16 d0 = { { a_prev.ch [0], a_prev.ch [1], a_prev.ch [2], 0 } };
                Err16 d1 = { { a_cur.ch [0], a_cur.ch [1], a_cur.ch [2], 0 } };
                row1 [x - dir] = d0;
                row1 [x] = d1;
            }
        }

        Err16 *tmp = row0;
        row0 = row1;
        row1 = tmp;
    }
}

__global__ void __launch_bounds__ (32)
sixel_fs_kernel (DevSixel p)
{
    __shared__ SharedTable s_table;
    TableView t;

    if (p.dynamic && !p.lut)
        load_table (s_table, t, p.table);

    const int w = p.width;
    Err16 *row0 = (Err16 *) p.scratch, *row1 = row0 + w;
    const Err16 zero = { { 0, 0, 0, 0 } };

    for (int i = threadIdx.x; i < w; i += 32)
        row0 [i] = zero;
    /* rows added to reach a multiple of six */
    for (size_t i = (size_t) w * p.height + threadIdx.x; i < (size_t) w * p.image_height; i += 32)
        p.indexed [i] = (uint8_t) p.transparent_index;
    __syncwarp ();

    for (int y = 0; y < p.height; y++)
    {
        for (int i = threadIdx.x; i < w; i += 32)
            row1 [i] = zero;
        __syncwarp ();

        if (threadIdx.x == 0)
        {
            const uint32_t *in = p.pixels + (size_t) y * w;
            uint8_t *out = p.indexed + (size_t) y * w;

            if (w == 1)
            {
                /* both passes degenerate to one pixel; the forward pass writes x = 0 twice */
                if (y & 1)
                {
                    out [0] = fs_pixel (p, t, in [0], row0 [0], row0 + 1, row1 + 1, row1, row1 + 1);
                    out [0] = fs_pixel (p, t, in [0], row0 [0], row1, row1, row1 - 1, row1 - 1);
                }
                else
                {
                    out [0] = fs_pixel (p, t, in [0], row0 [0], row0 - 1, row1 - 1, row1, row1 - 1);
                    out [0] = fs_pixel (p, t, in [0], row0 [0], row1, row1, row1 + 1, row1 + 1);
                }
            }
            else if (y & 1)
            {
                int x;
                out [0] = fs_pixel (p, t, in [0], row0 [0], row0 + 1, row1 + 1, row1, row1 + 1);
                for (x = 1; x < w - 1; x++)
                    out [x] = fs_pixel (p, t, in [x], row0 [x], row0 + x + 1, row1 + x + 1, row1 + x, row1 + x - 1);
                out [x] = fs_pixel (p, t, in [x], row0 [x], row1 + x, row1 + x, row1 + x - 1, row1 + x - 1);
            }
            else
            {
                int x = w - 1;
                out [x] = fs_pixel (p, t, in [x], row0 [x], row0 + x - 1, row1 + x - 1, row1 + x, row1 + x - 1);
                for (x--; x >= 1; x--)
                    out [x] = fs_pixel (p, t, in [x], row0 [x], row0 + x - 1, row1 + x - 1, row1 + x, row1 + x + 1);
                out [0] = fs_pixel (p, t, in [0], row0 [0], row1, row1, row1 + 1, row1 + 1);
            }
        }
        __syncwarp ();

        Err16 *tmp = row0;
        row0 = row1;
        row1 = tmp;
    }
}

/* ------------------------------------------------------------------------ */
/* K6d                                                                        */

/* Text of one run of @n equal sixel characters (chafa-sixel-renderer.c:201-236) */
template <bool WRITE>
__device__ __forceinline__ int
put_run (uint8_t *out, uint8_t c, int n)
{
    int len = 0;

    for (;;)
    {
        if (n < 4)
        {
            do
            {
                if (WRITE) out [len] = c;
                len++;
            }
            while (--n);
            return len;
        }

        int m = n < 255 ? n : 255;
        if (WRITE)
        {
            out [len] = '!';
            int k = len + 1;
            if (m >= 100) out [k++] = (uint8_t) ('0' + m / 100);
            if (m >= 10) out [k++] = (uint8_t) ('0' + (m / 10) % 10);
            out [k++] = (uint8_t) ('0' + m % 10);
            out [k] = c;
        }
        len += 1 + (m >= 100 ? 3 : m >= 10 ? 2 : 1) + 1;
        if (n < 255)
            return len;
        n -= 255;
        if (n == 0)
            return len;
    }
}

/* Which pens occur in each band?  Block per band; the flags go to pen_ofs, which the layout kernel
 * overwrites after the measuring pass has used them to skip the pens a band does not contain. */
__global__ void __launch_bounds__ (256)
sixel_band_pens_kernel (DevSixelEncode p)
{
    __shared__ uint32_t s_present [256];
    const int band = blockIdx.x;
    const uint8_t *rows = p.indexed + (size_t) band * 6 * p.width;

    s_present [threadIdx.x] = 0;
    __syncthreads ();
    for (int i = threadIdx.x; i < 6 * p.width; i += 256)
        s_present [rows [i]] = 1;
    __syncthreads ();
    if ((int) threadIdx.x < p.n_colors)
        p.pen_ofs [(size_t) band * p.n_colors + threadIdx.x] = s_present [threadIdx.x];
}

/* One warp per (band, pen): run-length code the pen's row of sixel characters.
 * MEASURE: body length into body_len.  Otherwise: write at the final offset. */
template <bool WRITE>
__global__ void __launch_bounds__ (256)
sixel_encode_kernel (DevSixelEncode p)
{
    const int lane = threadIdx.x & 31;
    const int wid = (int) ((size_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    if (wid >= p.n_bands * p.n_colors)
        return;

    const int band = wid / p.n_colors, pen = wid % p.n_colors;
    if (pen == p.transparent_index)
    {
        if (!WRITE && lane == 0)
            p.body_len [wid] = 0;
        return;
    }

    const uint8_t *rows = p.indexed + (size_t) band * 6 * p.width;
    const bool force = (band == 0 || (band + 1) * 6 >= p.height) && pen == p.first_pen;
    uint8_t *out = nullptr;

    if (!WRITE && !force && p.n_colors <= 256 && p.pen_ofs [wid] == 0)
    {
        /* the pen does not occur in this band: its row is one run of '?', which is dropped */
        if (lane == 0)
            p.body_len [wid] = 0;
        return;
    }
    if (WRITE)
    {
        if (p.body_len [wid] == 0)
            return;
        out = p.out + p.band_ofs [band] + p.pen_ofs [wid] + (p.prefix_len [wid] & 0xffu);
    }

    /* carry between chunks of 32 columns: the open run */
    int run_char = -1, run_len = 0;
    int total = 0;

    for (int base = 0; base < p.width; base += 32)
    {
        int x = base + lane;
        int c = -1;
        if (x < p.width)
        {
            int bits = 0;
#pragma unroll
            for (int r = 0; r < 6; r++)
                bits |= (rows [(size_t) r * p.width + x] == pen ? 1 : 0) << r;
            c = '?' + bits;
        }

        /* a run ends at x if the next column differs (or x is the last column) */
        int next = __shfl_down_sync (FULL, c, 1);
        if (lane == 31)
            next = -2;                       /* decided at the next chunk */
        bool is_last = x == p.width - 1;
        bool ends = x < p.width && (is_last || (lane < 31 && next != c));
        /* lane 31's run end is unknown unless it is the last column: defer by treating the
         * chunk boundary as "continues", and close it when the next chunk starts */
        uint32_t end_mask = __ballot_sync (FULL, ends);

        /* first: does the carried run end exactly at the chunk boundary? */
        int first_c = __shfl_sync (FULL, c, 0);
        int carried_len = 0, carried_char = run_char;
        if (run_char >= 0 && first_c != run_char)
        {
            carried_len = run_len;           /* flushed by lane 0 below, before its own run */
            run_char = -1;
            run_len = 0;
        }

        /* start of the run this lane ends: after the previous run end in this chunk */
        uint32_t below = end_mask & ((1u << lane) - 1u);
        int start_lane = below ? 32 - __clz (below) : 0;
        int len_here = lane - start_lane + 1;
        bool continues_carry = ends && below == 0 && run_char >= 0;     /* run began in an earlier chunk */
        int my_len = ends ? len_here + (continues_carry ? run_len : 0) : 0;

        /* the trailing '?' run of the row is dropped unless forced */
        bool emit = ends && !(is_last && c == '?' && !force);
        int my_text = 0;
        if (emit)
            my_text = put_run<false> (nullptr, (uint8_t) c, my_len);
        int carried_text = 0;
        if (lane == 0 && carried_len > 0)
            carried_text = put_run<false> (nullptr, (uint8_t) carried_char, carried_len);

        /* exclusive prefix of text lengths over lanes; lane 0's carried run goes first */
        int mine = my_text + carried_text;
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            int o = __shfl_up_sync (FULL, incl, d);
            if (lane >= d)
                incl += o;
        }
        int excl = incl - mine;
        int chunk_total = __shfl_sync (FULL, incl, 31);

        if (WRITE)
        {
            uint8_t *dst = out + total + excl;
            if (lane == 0 && carried_len > 0)
                dst += put_run<true> (dst, (uint8_t) carried_char, carried_len);
            if (emit)
                put_run<true> (dst, (uint8_t) c, my_len);
        }
        total += chunk_total;

        /* new carry: columns after the last run end of this chunk */
        int n_valid = min (32, p.width - base);
        int last_end = end_mask ? 32 - __clz (end_mask) : 0;       /* lanes [last_end, n_valid) are open */
        if (last_end < n_valid)
        {
            int open_c = __shfl_sync (FULL, c, last_end);
            int open_len = n_valid - last_end;
            if (end_mask == 0 && run_char >= 0)
                run_len += open_len;                                /* whole chunk continues the carry */
            else
            {
                run_char = open_c;
                run_len = open_len;
            }
        }
        else
        {
            run_char = -1;
            run_len = 0;
        }
    }

    if (!WRITE && lane == 0)
        p.body_len [wid] = (uint32_t) total;
}

/* Per band: prefix lengths ('$' when an earlier pen of the band printed, '#' + pen),
 * pen offsets inside the band and the band's length (+ '-' unless it is the last) */
__global__ void
sixel_band_layout_kernel (DevSixelEncode p)
{
    int band = blockIdx.x * blockDim.x + threadIdx.x;
    if (band >= p.n_bands)
        return;

    uint32_t ofs = 0;
    bool need_cr = false;
    for (int pen = 0; pen < p.n_colors; pen++)
    {
        int wid = band * p.n_colors + pen;
        uint32_t body = p.body_len [wid];
        uint32_t prefix = 0, cr_flag = 0;
        if (body)
        {
            prefix = (need_cr ? 1u : 0u) + 1u + (pen >= 100 ? 3u : pen >= 10 ? 2u : 1u);
            cr_flag = need_cr ? 0x80000000u : 0u;
            need_cr = true;
        }
        p.prefix_len [wid] = prefix | cr_flag;      /* low byte: length; bit 31: starts with '$' */
        p.pen_ofs [wid] = ofs;
        ofs += prefix + body;
    }
    bool is_last = (band + 1) * 6 >= p.height;
    p.band_len [band] = ofs + (is_last ? 0u : 1u);
}

__global__ void
sixel_band_offsets_kernel (DevSixelEncode p)
{
    if (blockIdx.x || threadIdx.x)
        return;
    unsigned long long acc = 0;
    for (int b = 0; b < p.n_bands; b++)
    {
        p.band_ofs [b] = acc;
        acc += p.band_len [b];
    }
    p.band_ofs [p.n_bands] = acc;
}

/* Prefixes and band separators; the bodies are written by sixel_encode_kernel<true> */
__global__ void __launch_bounds__ (256)
sixel_prefix_kernel (DevSixelEncode p)
{
    int wid = blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= p.n_bands * p.n_colors)
        return;
    int band = wid / p.n_colors, pen = wid % p.n_colors;

    if (pen == 0)
    {
        bool is_last = (band + 1) * 6 >= p.height;
        if (!is_last)
            p.out [p.band_ofs [band + 1] - 1] = '-';
    }

    uint32_t pl = p.prefix_len [wid];
    if (!(pl & 0xffu))
        return;
    uint8_t *dst = p.out + p.band_ofs [band] + p.pen_ofs [wid];
    if (pl & 0x80000000u)
        *dst++ = '$';
    *dst++ = '#';
    if (pen >= 100) *dst++ = (uint8_t) ('0' + pen / 100);
    if (pen >= 10) *dst++ = (uint8_t) ('0' + (pen / 10) % 10);
    *dst++ = (uint8_t) ('0' + pen % 10);
}

/* ------------------------------------------------------------------------ */

/* chafa-color.c:83-168 */
__device__ uint32_t
sixel_rgb_to_din99d (uint32_t px)
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

    uint32_t c0 = (uint32_t) (int) (adj_L * 2.5) & 0xff;
    uint32_t c1 = (uint32_t) (int) (C * cos (h) * 2.5 + 128.0) & 0xff;
    uint32_t c2 = (uint32_t) (int) (C * sin (h) * 2.5 + 128.0) & 0xff;
    return c0 | (c1 << 8) | (c2 << 16) | (px & 0xff000000u);
}

/* ------------------------------------------------------------------------ */
/* Launchers                                                                  */

static unsigned
grid_for (size_t n, int per_block)
{
    int dev = 0, sms = 148;
    cudaGetDevice (&dev);
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t blocks = (n + per_block - 1) / per_block;
    if (blocks > (size_t) sms * 8)
        blocks = (size_t) sms * 8;
    return (unsigned) (blocks ? blocks : 1);
}

int
dev_launch_sixel_quantise (const DevSixel *p, void *stream)
{
    size_t n = (size_t) p->width * p->image_height;
    if (!n)
        return 0;
    if (p->dither_mode == 2)
    {
        if (p->lut)
        {
            sixel_lut_kernel<<<grid_for ((size_t) 1 << 24, 256), 256, 0, (cudaStream_t) stream>>> (*p);
            dev_count_launch (1);
        }
        if (p->lut && p->width >= 3 && (p->color_space == 0 || p->preconverted))
        {
            if (p->intensity == 1.0f)
                sixel_fs_fast_kernel<true><<<1, 32, 0, (cudaStream_t) stream>>> (*p);
            else
                sixel_fs_fast_kernel<false><<<1, 32, 0, (cudaStream_t) stream>>> (*p);
        }
        else
            sixel_fs_kernel<<<1, 32, 0, (cudaStream_t) stream>>> (*p);
    }
    else
        sixel_quantise_kernel<<<grid_for (n, 256), 256, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_sixel_encode_measure (const DevSixelEncode *p, void *stream)
{
    int n_warps = p->n_bands * p->n_colors;
    if (n_warps <= 0)
        return 0;
    cudaStream_t st = (cudaStream_t) stream;
    if (p->n_colors <= 256)
    {
        sixel_band_pens_kernel<<<p->n_bands, 256, 0, st>>> (*p);
        dev_count_launch (1);
    }
    sixel_encode_kernel<false><<<(n_warps + 7) / 8, 256, 0, st>>> (*p);
    sixel_band_layout_kernel<<<(p->n_bands + 63) / 64, 64, 0, st>>> (*p);
    sixel_band_offsets_kernel<<<1, 32, 0, st>>> (*p);
    dev_count_launch (3);
    return (int) cudaGetLastError ();
}

int
dev_launch_sixel_encode_emit (const DevSixelEncode *p, void *stream)
{
    int n_warps = p->n_bands * p->n_colors;
    if (n_warps <= 0)
        return 0;
    cudaStream_t st = (cudaStream_t) stream;
    sixel_prefix_kernel<<<(n_warps + 255) / 256, 256, 0, st>>> (*p);
    sixel_encode_kernel<true><<<(n_warps + 7) / 8, 256, 0, st>>> (*p);
    dev_count_launch (2);
    return (int) cudaGetLastError ();
}
