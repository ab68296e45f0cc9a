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

#include <cstdlib>

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
    bool narrow;                /* every squared distance of the walk fits a positive int: see load_table () */
    /* the same table as arrays (shared memory), for the 32-bit walk: projections of entry j on the two
     * axes and the colour of its pen, one load each */
    const int *v0, *v1;
    const uint32_t *entry_color;
};

/* sum over the three channels of (32 (b - a))^2 = 1024 sum (b - a)^2: per-byte absolute differences
 * (VABSDIFF4) and their dot product with themselves (IDP.4A) */
__device__ __forceinline__ int
table_color_diff (uint32_t a, uint32_t b)
{
    const uint32_t d = __vabsdiffu4 (a & 0x00ffffffu, b & 0x00ffffffu);
    return (int) (__dp4a (d, d, 0u) << 10);
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

/* table_refine () in 32-bit arithmetic from the array form of the table: the same comparisons on
 * the same values whenever TableView::narrow holds (all of a, b, d below 2^31, and so is the running
 * best after the first candidate, which is always taken) */
__device__ __forceinline__ bool
table_refine_narrow (const TableView &t, uint32_t want, int w0, int w1, int j, int &best_pen, int &best_diff)
{
    int a = t.v0 [j] - w0;
    a *= a;
    if (a > best_diff)
        return false;

    int b = t.v1 [j] - w1;
    b *= b;
    if (b <= best_diff)
    {
        int d = table_color_diff (t.entry_color [j], want);
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

    if (t.narrow)
    {
        while (i != j)
        {
            int n = i + (j - i) / 2;
            if (v [0] > t.v0 [n])
                i = n + 1;
            else
                j = n;
        }
        const int m = j;
        int best_diff = 0x7fffffff, best_pen = 0, from = m;
        if (m == t.n_entries)
        {
            /* past the last entry: the walk starts on the zero record that follows it, whose
             * projections are outside the range the 32-bit form covers; with nothing found yet it
             * passes both tests whatever they are and becomes the first candidate */
            best_pen = m;
            best_diff = table_color_diff (t.entry_color [m], want);
            from = m - 1;
        }
        for (j = from; j >= 0; j--)
            if (!table_refine_narrow (t, want, v [0], v [1], j, best_pen, best_diff))
                break;
        for (j = m + 1; j < t.n_entries; j++)
            if (!table_refine_narrow (t, want, v [0], v [1], j, best_pen, best_diff))
                break;
        return t.entries [best_pen].pen;
    }

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
    int v0 [257], v1 [257];
    uint32_t entry_color [257];
    int narrow;
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

    /* Can the walk run in 32-bit arithmetic?  Its squared terms are (entry projection - colour
     * projection)^2 per axis and a colour distance below 2^28.  The projections of a colour are
     * largest, in magnitude, at a corner of the colour cube (a linear form, truncated towards zero),
     * so eight corners and the table's own entries bound every difference; below 46 340 its square
     * is a positive int.  Tables with unit-length axes have differences of a few ten thousand at
     * most; degenerate axes make the scale factor large, and those tables keep the 64-bit walk. */
    if (threadIdx.x == 0)
    {
        /* (the projections carry a large common offset -- the table's average is kept with five more
         * fractional bits than the colours it is subtracted from, as in the reference -- so it is the
         * range of the differences that is bounded, not the magnitudes) */
        bool ok = true;
        for (int k = 0; k < 2; k++)
        {
            long long want_min = 0, want_max = 0, entry_min = 0, entry_max = 0;
            for (int corner = 0; corner < 8; corner++)
            {
                int q [3] = { ((corner & 1) ? 255 : 0) * 32 - t.avg [0], ((corner & 2) ? 255 : 0) * 32 - t.avg [1],
                              ((corner & 4) ? 255 : 0) * 32 - t.avg [2] };
                long long d = (long long) q [0] * t.ev [k] [0] + (long long) q [1] * t.ev [k] [1] + (long long) q [2] * t.ev [k] [2];
                long long pr = (d * t.mul [k]) / ((1 << 14) / 32);
                if (pr != (long long) (int) pr)
                    ok = false;                 /* table_project () would have truncated it */
                want_min = corner == 0 || pr < want_min ? pr : want_min;
                want_max = corner == 0 || pr > want_max ? pr : want_max;
            }
            for (int i = 0; i < t.n_entries && i < 257; i++)
            {
                long long a = s.entries [i].v [k];
                entry_min = i == 0 || a < entry_min ? a : entry_min;
                entry_max = i == 0 || a > entry_max ? a : entry_max;
            }
            /* one unit of slack for the truncation inside the cube */
            if (want_max - entry_min + 1 >= 46340 || entry_max - want_min + 1 >= 46340)
                ok = false;
        }
        s.narrow = ok && t.n_entries > 0;
    }
    __syncthreads ();
    for (int i = threadIdx.x; i < 257; i += blockDim.x)
    {
        s.v0 [i] = s.entries [i].v [0];
        s.v1 [i] = s.entries [i].v [1];
        s.entry_color [i] = s.pens [s.entries [i].pen & 255];
    }
    __syncthreads ();
    t.narrow = s.narrow != 0;
    t.v0 = s.v0;
    t.v1 = s.v1;
    t.entry_color = s.entry_color;
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

/* The diffusion chain's table: the pen of every 24-bit colour, one byte each (16 MB).
 * Neighbouring pixels of an image have neighbouring colours, and error diffusion spreads the
 * colours it looks up a few dozen levels around them: the table is laid out in bricks of
 * 8 x 4 x 4 colours (one 128-byte line each) so that this cloud stays resident in the first-level
 * cache of the SM that runs the chain (measured: one byte per colour keeps 98 % of the chain's
 * loads in L1; four bytes per colour -- pen and pen colour in one load -- drops that to 57 % and
 * loses more than the saved shared-memory load gains).
 *
 * Byte index of colour (c0, c1, c2): c0 & 7 | (c1 & 3) << 3 | (c2 & 3) << 5 | (c0 >> 3) << 7
 * | (c1 >> 2) << 12 | (c2 >> 2) << 18.  The table is aligned to its size, so the address is the
 * base with the index OR-ed into its low 24 bits (no carry: the chain computes a 32-bit value). */
#define FS_LUT_BYTES ((size_t) 1 << 24)

__device__ __forceinline__ uint32_t
lut_index (uint32_t c0, uint32_t c1, uint32_t c2)
{
    return (c0 & 7u) | ((c1 & 3u) << 3) | ((c2 & 3u) << 5) | ((c0 >> 3) << 7) | ((c1 >> 2) << 12) | ((c2 >> 2) << 18);
}

/* Low address word of the entry of (c0, c1, c2), each in 0..255: two shifted copies per channel
 * merged bit-field by bit-field (select under a mask; stray bits of one merge are overwritten by
 * the next).  Four dependent operations. */
__device__ __forceinline__ uint32_t
lut_address_lo (uint32_t base_lo, uint32_t c0, uint32_t c1, uint32_t c2)
{
    const uint32_t a = (c0 & ~0xf80u) | ((c0 << 4) & 0xf80u);                       /* bits 0-2, 7-11; 3-6 stray */
    const uint32_t b = ((c1 << 3) & ~0x3f000u) | ((c1 << 10) & 0x3f000u);          /* bits 3-4, 12-17; 5-10 stray */
    const uint32_t c = ((c2 << 5) & ~0xfffc0000u) | ((c2 * 65536u + base_lo) & 0xfffc0000u);   /* 5-6, 18-31 */
    const uint32_t ab = (a & ~0x3f018u) | (b & 0x3f018u);                          /* 5-6 stray */
    return (ab & ~0xfffc0060u) | (c & 0xfffc0060u);
}

/* Pen of every 24-bit colour: the pen search is a pure function of the colour once the palette
 * exists, so all 2^24 answers are computed in parallel (a few ms) and the serial chain pays one
 * load per pixel instead of a table walk.  Indexed by an opaque colour in the canvas's colour
 * space. */
__global__ void __launch_bounds__ (256)
sixel_lut_kernel (DevSixel p)
{
    __shared__ SharedTable s_table;
    TableView t;

    if (p.dynamic)
        load_table (s_table, t, p.table);

    /* The table is filled in the order of its bytes: a warp then owns 32 consecutive bytes, i.e. a
     * slab of 8 x 4 x 1 colours (the inverse of lut_index ()) -- stores that coalesce, and 32 colours
     * closer to one another than 32 in a line, whose walks over the pen table are of similar length
     * (a warp walks as far as its farthest lane). */
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (1u << 24); i += gridDim.x * blockDim.x)
    {
        const uint32_t c0 = (i & 7u) | (((i >> 7) & 31u) << 3);
        const uint32_t c1 = ((i >> 3) & 3u) | (((i >> 12) & 63u) << 2);
        const uint32_t c2 = ((i >> 5) & 3u) | (((i >> 18) & 63u) << 2);
        p.lut [i] = (uint8_t) lookup_pen (p, t, c0 | (c1 << 8) | (c2 << 16) | 0xff000000u);
    }
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
            int y = (int) (i / (size_t) p.width) + p.y0, x = (int) (i % (size_t) p.width);
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
        int y = (int) (i / (size_t) p.width) + p.y0, x = (int) (i % (size_t) p.width);
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

        index = p.lut ? (int) p.lut [lut_index (color & 0xff, (color >> 8) & 0xff, (color >> 16) & 0xff)]
                      : lookup_pen (p, t, color);
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

/* Summary of the full table for the chain: one byte per brick of 8 x 8 x 8 colours (32 KB,
 * resident in the shared memory of the SM that runs the chain) holding the pen if all of the
 * brick's 512 colours share one, else FS_MIXED.  Where the palette's pens are far apart relative to
 * a brick (images whose colours fill the cube), most lookups end here at shared-memory latency
 * whatever the access pattern; where they are dense (smooth images: the palette follows the
 * image's colours) most bricks are mixed and the chain is better off going straight to the full
 * table through the first-level cache.  The chain times both ways on the image itself.  One warp
 * per brick: its 512 entries are four 128-byte lines of the full table, 16 bytes per lane. */
#define FS_MIXED 0xffu      /* (a brick that is all pen 255 -- fixed palettes only -- is looked up in the full table) */

__global__ void __launch_bounds__ (256)
sixel_brick_table_kernel (DevSixel p)
{
    const int lane = threadIdx.x & 31;
    const int brick = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      /* b0 | b1 << 5 | b2 << 10 */
    if (brick >= 32768)
        return;

    const uint32_t b0 = brick & 31, b1 = (brick >> 5) & 31, b2 = brick >> 10;
    const uint32_t q = lane >> 3;
    const uint32_t line = b0 | ((2 * b1 + (q & 1)) << 5) | ((2 * b2 + (q >> 1)) << 11);
    const uint4 v = *(const uint4 *) (p.lut + ((size_t) line << 7) + (lane & 7) * 16);

    const uint32_t lead = __shfl_sync (FULL, v.x & 0xff, 0);
    const bool same = v.x == lead * 0x01010101u && v.y == v.x && v.z == v.x && v.w == v.x;
    const bool whole = __all_sync (FULL, same);
    if (lane == 0)
        p.brick_pens [brick] = (uint8_t) (whole ? lead : FS_MIXED);
}

/* One Floyd-Steinberg chain over the image, row by row.  A block of FS_THREADS threads alternates
 * between two phases per row:
 *
 *   A  thread 0 walks the row's pixels in serpentine order.  Per pixel (intensity 1, the default:
 *      all integer): incoming error e (16-bit, x16 scale) -> compensated channel
 *      trunc ((160 px + 9 e) / 160) -- what the reference's float expression px + e * 0.9f / 16
 *      truncates to for every 16-bit e (tests/test_cpu.py checks all of them) -> clamp -> pen
 *      from the brick table or the full table -> pen colour -> error -> 7/16 of it onto the next
 *      pixel's incoming error.  The per-pixel errors and pens go to shared-memory rows; nothing
 *      else is on the chain.
 *   B  all threads: the next row's incoming errors from this row's errors (each element is the
 *      ordered sum 1/16 of the pixel before it, 5/16 of the one above it, 3/16 of the one after
 *      it in processing order, 16-bit truncation after every addition as in the reference; the
 *      first and last pixel of a row redirect their shares, chafa-indexed-image.c:207-266), the
 *      row's pens out to global memory, the next row's pixels in.
 *
 * Serpentine order: row 0 runs right to left (odd rows left to right, chafa-indexed-image.c:214). */
#define FS_THREADS 256

template <bool UNIT>
__device__ __forceinline__ int
fs_add (int acc, int err, int factor, float intensity)
{
    if (UNIT)
        return (int) (short) (acc + err * factor);
    return (int) (short) (int) __fadd_rn ((float) acc, __fmul_rn ((float) (err * factor), intensity));
}

/* next-row element @j (processing order, row of @w pixels): the shares that land on it, added in
 * the order the reference adds them */
template <bool UNIT>
__device__ __forceinline__ void
fs_next_row_element (int j, int w, const uint2 *packed, int x_of_j, int dir, float intensity, int *acc)
{
    struct Unpacked
    {
        short ch [3];
    };
    auto errs_at = [packed] (int x)
    {
        const uint2 v = packed [x];
        Unpacked e = { { (short) (v.x & 0xffff), (short) (v.x >> 16), (short) (v.y & 0xffff) } };
        return e;
    };
    const bool j_last = j == w - 1;
#pragma unroll
    for (int c = 0; c < 3; c++)
        acc [c] = 0;
    if (j >= 1)
    {
        /* the pixel before: 1/16 ahead; the row's first pixel also sends its 3/16 ahead */
        const Unpacked e = errs_at (x_of_j - dir);
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 1, intensity);
            if (j == 1)
                acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 3, intensity);
        }
    }
    {
        const Unpacked e = errs_at (x_of_j);
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            if (!j_last)
                acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 5, intensity);
            else
            {
                /* the row's last pixel: 7/16 and 1/16 go straight down */
                acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 7, intensity);
                acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 1, intensity);
            }
        }
    }
    if (j + 1 <= w - 1)
    {
        const Unpacked e = errs_at (x_of_j + dir);
        const bool next_last = j + 1 == w - 1;
#pragma unroll
        for (int c = 0; c < 3; c++)
        {
            if (next_last)
                acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 5, intensity);   /* ... and 5/16 and 3/16 behind */
            acc [c] = fs_add<UNIT> (acc [c], e.ch [c], 3, intensity);
        }
    }
}

/* Shared memory of the chain's block: palette colours, brick table, then per pixel of a row the
 * incoming error (two words: channel 0 | channel 1 << 16, channel 2), the produced error with the
 * pen (two words: error 0 | error 1 << 16, error 2 | pen << 16) and the pixel.  Rows carry two
 * pad elements at either end so that the chain's look-ahead needs no bounds check.  Kept small:
 * what shared memory does not take is first-level cache for the full table. */
#define FS_FIXED_BYTES (1024 + 32768)
#define FS_PAD 2

__host__ __device__ static inline size_t
fs_smem_bytes (int w)
{
    return FS_FIXED_BYTES + (((size_t) (w + 2 * FS_PAD) * (8 + 8 + 4) + 15) & ~(size_t) 15);
}

/* (a & ~m) | (b & m) */
__device__ __forceinline__ uint32_t
bit_select (uint32_t a, uint32_t b, uint32_t m)
{
    uint32_t r;
    asm ("lop3.b32 %0, %1, %2, %3, 0xd8;" : "=r" (r) : "r" (a), "r" (b), "r" (m));
    return r;
}

__device__ __forceinline__ uint32_t
lds_u32 (uint32_t shared_address)
{
    uint32_t v;
    asm volatile ("ld.shared.u32 %0, [%1];" : "=r" (v) : "r" (shared_address));
    return v;
}

__device__ __forceinline__ uint32_t
lds_u8 (uint32_t shared_address)
{
    uint32_t v;
    asm volatile ("ld.shared.u8 %0, [%1];" : "=r" (v) : "r" (shared_address));
    return v;
}

/* One row of the chain.  @BRICKS: consult the brick table first.  @DYNAMIC: generated palette (the
 * table never answers the transparent pen for an opaque colour and pens start at 0).  @SKIP: the
 * row has transparent pixels, which are stepped over without the arithmetic (their error is
 * zero and their pen fixed); rows without them run a body free of branches. */
template <bool UNIT, bool DYNAMIC, bool BRICKS, bool SKIP>
__device__ __forceinline__ void
fs_chain_row (const DevSixel &p, int y, const uint32_t *s_colors, const uint8_t *s_bricks, const int2 *s_in,
              uint2 *s_out, const uint32_t *s_px)
{
    const uint32_t colors_at = (uint32_t) __cvta_generic_to_shared (s_colors);
    const uint32_t bricks_at = (uint32_t) __cvta_generic_to_shared (s_bricks);
    const int w = p.width;
    const int dir = (y & 1) ? 1 : -1;
    const int x0 = (y & 1) ? 0 : w - 1;
    const float intensity = p.intensity;
    const int transparent = p.transparent_index, first_color = DYNAMIC ? 0 : p.first_color;
    const int alpha_threshold = p.alpha_threshold;
    const uint32_t lut_lo = (uint32_t) (uintptr_t) p.lut;
    const uint64_t lut_hi = (uint64_t) (uintptr_t) p.lut & 0xffffffff00000000ull;

    int2 in = s_in [x0];
    int e0 = (int) (short) in.x, e1 = in.x >> 16, e2 = in.y;
    uint32_t px = s_px [x0];
    uint32_t px_next = s_px [x0 + dir];
    int2 in_next = s_in [x0 + dir];
    /* 160 x channel of the current pixel, prepared one pixel ahead so that the chain only adds 9 e */
    int t0 = (int) __byte_perm (px, 0, 0x4440) * 160, t1 = (int) __byte_perm (px, 0, 0x4441) * 160,
        t2 = (int) __byte_perm (px, 0, 0x4442) * 160;

#pragma unroll 2
    for (int k = 0; k < w; k++)
    {
        const int x = x0 + k * dir;

        /* two pixels ahead (the rows are padded): off the chain */
        const uint32_t px_next2 = s_px [x + 2 * dir];
        const int2 in_next2 = s_in [x + 2 * dir];
        const int t0_next = (int) __byte_perm (px_next, 0, 0x4440) * 160, t1_next = (int) __byte_perm (px_next, 0, 0x4441) * 160,
                  t2_next = (int) __byte_perm (px_next, 0, 0x4442) * 160;
        const int n0 = (int) (short) in_next.x, n1 = in_next.x >> 16, n2 = in_next.y;

        /* No branches in here: the two pixels of the unrolled body are one block of code, so the
         * next pixel's preparation is scheduled into this pixel's waits.  A transparent pixel runs
         * through the same arithmetic and its results are discarded. */
        const bool opaque = (int) (px >> 24) >= alpha_threshold;
        if (SKIP && !opaque)
        {
            e0 = n0;
            e1 = n1;
            e2 = n2;
            s_out [x] = make_uint2 (0u, (uint32_t) transparent << 16);
            px = px_next;
            px_next = px_next2;
            in_next = in_next2;
            t0 = t0_next;
            t1 = t1_next;
            t2 = t2_next;
            continue;
        }
        int c0, c1, c2;
        if (UNIT)
        {
            c0 = (e0 * 9 + t0) / 160;
            c1 = (e1 * 9 + t1) / 160;
            c2 = (e2 * 9 + t2) / 160;
        }
        else
        {
            c0 = (int) (short) (int) __fadd_rn ((float) (int) (px & 0xff), __fmul_rn (__fmul_rn ((float) e0, 0.9f), 0.0625f));
            c1 = (int) (short) (int) __fadd_rn ((float) (int) ((px >> 8) & 0xff), __fmul_rn (__fmul_rn ((float) e1, 0.9f), 0.0625f));
            c2 = (int) (short) (int) __fadd_rn ((float) (int) ((px >> 16) & 0xff), __fmul_rn (__fmul_rn ((float) e2, 0.9f), 0.0625f));
        }
        const uint32_t q0 = (uint32_t) min (max (c0, 0), 255), q1 = (uint32_t) min (max (c1, 0), 255),
                       q2 = (uint32_t) min (max (c2, 0), 255);

        uint32_t found_pen = FS_MIXED;
        if (BRICKS)
            found_pen = lds_u8 (bricks_at + ((q0 >> 3) | ((q1 >> 3) << 5) | ((q2 >> 3) << 10)));
        {
            /* address of the table entry: the base with the entry's index in its low 24 bits,
             * merged field by field (stray bits of one merge are overwritten by the next) */
            const uint32_t fa = bit_select (q0, q0 << 4, 0xf80u);                       /* bits 0-2, 7-11; 3-6 stray */
            const uint32_t fb = bit_select (q1 << 3, q1 << 10, 0x3f000u);               /* bits 3-4, 12-17; 5-10 stray */
            const uint32_t fc = bit_select (q2 << 5, q2 * 65536u + lut_lo, 0xfffc0000u);   /* bits 5-6, 18-31 */
            const uint32_t lo = bit_select (bit_select (fa, fb, 0x3f018u), fc, 0xfffc0060u);
            if (BRICKS)
                asm volatile ("{ .reg .pred p; setp.eq.u32 p, %0, 255; @p ld.global.u8 %0, [%1]; }"
                              : "+r" (found_pen) : "l" (lut_hi | lo));
            else
                asm volatile ("ld.global.u8 %0, [%1];" : "=r" (found_pen) : "l" (lut_hi | lo));
        }

        const uint32_t found = lds_u32 (colors_at + found_pen * 4);
        const bool keep = opaque && (DYNAMIC || (int) found_pen != transparent);
        const int r0 = keep ? (int) (short) (c0 - (int) __byte_perm (found, 0, 0x4440)) : 0;
        const int r1 = keep ? (int) (short) (c1 - (int) __byte_perm (found, 0, 0x4441)) : 0;
        const int r2 = keep ? (int) (short) (c2 - (int) __byte_perm (found, 0, 0x4442)) : 0;
        const int pen = opaque ? (int) found_pen - first_color : transparent;

        /* 7/16 onto the next pixel's incoming error (what the previous row left there) */
        e0 = fs_add<UNIT> (n0, r0, 7, intensity);
        e1 = fs_add<UNIT> (n1, r1, 7, intensity);
        e2 = fs_add<UNIT> (n2, r2, 7, intensity);

        s_out [x] = make_uint2 (__byte_perm ((uint32_t) r0, (uint32_t) r1, 0x5410),
                                __byte_perm ((uint32_t) r2, (uint32_t) pen, 0x5410));

        px = px_next;
        px_next = px_next2;
        in_next = in_next2;
        t0 = t0_next;
        t1 = t1_next;
        t2 = t2_next;
    }
}

template <bool UNIT, bool DYNAMIC>
__device__ __forceinline__ void
fs_rows_block (const DevSixel &p)
{
    extern __shared__ __align__ (16) unsigned char fs_smem [];
    __shared__ int s_row_has_transparent [2];       /* of the row being walked / being loaded */
    const int tid = threadIdx.x;
    const int w = p.width;

    uint32_t *s_colors = (uint32_t *) fs_smem;
    uint8_t *s_bricks = fs_smem + 1024;
    if (tid < 2)
        s_row_has_transparent [tid] = 0;
    __syncthreads ();
    int2 *s_in = (int2 *) (fs_smem + FS_FIXED_BYTES) + FS_PAD;
    uint2 *s_out = (uint2 *) (s_in + w + FS_PAD) + FS_PAD;
    uint32_t *s_px = (uint32_t *) (s_out + w + FS_PAD) + FS_PAD;

    s_colors [tid] = p.pal_colors [tid] & 0x00ffffffu;
    for (int i = tid; i < 32768 / 16; i += FS_THREADS)
        ((uint4 *) s_bricks) [i] = ((const uint4 *) p.brick_pens) [i];
    for (int i = tid - FS_PAD; i < w + FS_PAD; i += FS_THREADS)
    {
        const uint32_t px = i >= 0 && i < w ? p.pixels [i] : 0xff000000u;
        s_in [i] = make_int2 (0, 0);
        s_px [i] = px;
        if ((int) (px >> 24) < p.alpha_threshold)
            s_row_has_transparent [0] = 1;
    }
    /* rows added to reach a multiple of six */
    for (size_t i = (size_t) w * p.height + tid; i < (size_t) w * p.image_height; i += FS_THREADS)
        p.indexed [i] = (uint8_t) p.transparent_index;
    __syncthreads ();

    /* Brick table or not: every 64 rows the chain times two rows each way and keeps the faster for
     * the rest (the output does not depend on it) */
    long long t_with = 0, t_without = 0;
    bool use_bricks = true;

    for (int y = 0; y < p.height; y++)
    {
        if (tid == 0)
        {
            const int phase = y & 63;
            bool bricks = phase < 2 ? true : phase < 4 ? false : use_bricks;
            if (p.brick_policy)
                bricks = p.brick_policy == 1;
            const long long t0 = clock64 ();
            const bool skip = s_row_has_transparent [y & 1] != 0;
            if (bricks && skip)
                fs_chain_row<UNIT, DYNAMIC, true, true> (p, y, s_colors, s_bricks, s_in, s_out, s_px);
            else if (bricks)
                fs_chain_row<UNIT, DYNAMIC, true, false> (p, y, s_colors, s_bricks, s_in, s_out, s_px);
            else if (skip)
                fs_chain_row<UNIT, DYNAMIC, false, true> (p, y, s_colors, s_bricks, s_in, s_out, s_px);
            else
                fs_chain_row<UNIT, DYNAMIC, false, false> (p, y, s_colors, s_bricks, s_in, s_out, s_px);
            const long long dt = clock64 () - t0;
            if (phase == 0)
                t_with = dt;
            else if (phase == 1)
                t_with += dt;
            else if (phase == 2)
                t_without = dt;
            else if (phase == 3)
                use_bricks = t_with < t_without + dt;
        }
        __syncthreads ();

        const int dir = (y & 1) ? 1 : -1;
        if (tid == 0)
            s_row_has_transparent [(y + 1) & 1] = 0;
        __syncthreads ();
        for (int x = tid; x < w; x += FS_THREADS)
        {
            const int j = dir > 0 ? x : w - 1 - x;
            int a [3];
            fs_next_row_element<UNIT> (j, w, s_out, x, dir, p.intensity, a);
            s_in [x] = make_int2 ((int) __byte_perm ((uint32_t) a [0], (uint32_t) a [1], 0x5410), a [2]);
            p.indexed [(size_t) y * w + x] = (uint8_t) (s_out [x].y >> 16);
            if (y + 1 < p.height)
            {
                const uint32_t px = p.pixels [(size_t) (y + 1) * w + x];
                s_px [x] = px;
                if ((int) (px >> 24) < p.alpha_threshold)
                    s_row_has_transparent [(y + 1) & 1] = 1;
            }
        }
        __syncthreads ();
    }
}

template <bool UNIT, bool DYNAMIC>
__global__ void __launch_bounds__ (FS_THREADS)
sixel_fs_rows_kernel (DevSixel p)
{
    fs_rows_block<UNIT, DYNAMIC> (p);
}

/* Many images, one launch: block b walks the chain of image b.  A process has 32 hardware queues,
 * so chains launched one per stream stop overlapping beyond 32 images; a grid has no such limit
 * (one block per SM: 148 chains side by side). */
template <bool UNIT, bool DYNAMIC>
__global__ void __launch_bounds__ (FS_THREADS)
sixel_fs_rows_batch_kernel (const DevSixel *__restrict__ images)
{
    const DevSixel p = images [blockIdx.x];
    fs_rows_block<UNIT, DYNAMIC> (p);
}

template <bool UNIT, bool DYNAMIC>
static cudaError_t
launch_fs_rows (const DevSixel *p, cudaStream_t st)
{
    const size_t smem = fs_smem_bytes (p->width);
    cudaError_t err = (cudaError_t) dev_opt_in_smem ((const void *) sixel_fs_rows_kernel<UNIT, DYNAMIC>, smem);
    if (err != cudaSuccess)
        return err;
    sixel_fs_rows_kernel<UNIT, DYNAMIC><<<1, FS_THREADS, smem, st>>> (*p);
    return cudaSuccess;
}

template <bool UNIT, bool DYNAMIC>
static cudaError_t
launch_fs_rows_batch (const DevSixel *d_images, int n, int max_width, cudaStream_t st)
{
    const size_t smem = fs_smem_bytes (max_width);
    cudaError_t err = (cudaError_t) dev_opt_in_smem ((const void *) sixel_fs_rows_batch_kernel<UNIT, DYNAMIC>, smem);
    if (err != cudaSuccess)
        return err;
    sixel_fs_rows_batch_kernel<UNIT, DYNAMIC><<<n, FS_THREADS, smem, st>>> (d_images);
    return cudaSuccess;
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
    const int image_band = band + p.band0;
    const bool force = (image_band == 0 || (image_band + 1) * 6 >= p.height) && pen == p.first_pen;
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
    bool is_last = (band + p.band0 + 1) * 6 >= p.height;
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
        bool is_last = (band + p.band0 + 1) * 6 >= p.height;
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

/* Diffusion: does the block-per-image chain kernel apply (else the one-warp fallback walks the table)? */
bool
dev_sixel_chain_applies (const DevSixel *p)
{
    return p->dither_mode == 2 && p->lut && p->brick_pens && p->width >= 3 && (p->color_space == 0 || p->preconverted)
        && fs_smem_bytes (p->width) <= 227 * 1024;
}

/* The tables of one image's chain (pen of every colour, brick summary), without the chain */
int
dev_launch_sixel_chain_tables (const DevSixel *p, void *stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    sixel_lut_kernel<<<grid_for ((size_t) 1 << 24, 256), 256, 0, st>>> (*p);
    sixel_brick_table_kernel<<<32768 / 8, 256, 0, st>>> (*p);
    dev_count_launch (2);
    return (int) cudaGetLastError ();
}

/* The chains of @n images (device array of their parameters) in one launch.  All of them must be
 * of the same kind: @unit_intensity (dither intensity exactly 1), @dynamic (generated palette). */
int
dev_launch_sixel_chain_batch (const DevSixel *d_images, int n, int max_width, bool unit_intensity, bool dynamic,
                              void *stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    if (n <= 0)
        return 0;
    cudaError_t err = unit_intensity
        ? (dynamic ? launch_fs_rows_batch<true, true> (d_images, n, max_width, st)
                   : launch_fs_rows_batch<true, false> (d_images, n, max_width, st))
        : (dynamic ? launch_fs_rows_batch<false, true> (d_images, n, max_width, st)
                   : launch_fs_rows_batch<false, false> (d_images, n, max_width, st));
    if (err != cudaSuccess)
        return (int) err;
    dev_count_launch (1);
    return (int) cudaGetLastError ();
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
        if (dev_sixel_chain_applies (p))
        {
            cudaStream_t st = (cudaStream_t) stream;
            sixel_brick_table_kernel<<<32768 / 8, 256, 0, st>>> (*p);
            cudaError_t err = p->intensity == 1.0f
                ? (p->dynamic ? launch_fs_rows<true, true> (p, st) : launch_fs_rows<true, false> (p, st))
                : (p->dynamic ? launch_fs_rows<false, true> (p, st) : launch_fs_rows<false, false> (p, st));
            if (err != cudaSuccess)
                return (int) err;
            dev_count_launch (1);
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
