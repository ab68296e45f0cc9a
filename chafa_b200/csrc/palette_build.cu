/* palette_build.cu -- the generated ("dynamic") 256-colour palette of the sixel mode, built on
 * the device from the scaled image: no host arithmetic and no host round trip inside a draw.
 *
 * What the reference computes (chafa/internal/chafa-palette.c:388-953, chafa-pca.c:29-162,
 * chafa-color-table.c:92-191, 288-313) is a chain of float results whose last bits decide integer
 * outputs, so every float operation below is an explicitly rounded one (__fmul_rn, __fadd_rn,
 * __fdiv_rn, __fsqrt_rn: never contracted to FMA) applied in the operand order of the pinned
 * reference build (oracle/_ref, gcc -O2 -ffast-math: e.g. the common factor of a squared
 * distance is applied after the three weighted squares have been summed).  The structure is
 * this file's own:
 *
 *   KP1  sampling          every n-th pixel into colour-cube bins, exact integer sums (atomics);
 *                          a dense second pass into a second bin set runs only if the first saw
 *                          fewer than 256 opaque samples (decided on the device)
 *   KP2  float bins        bins whose float sum would leave the exact range (count * 255 >=
 *                          2^24) are re-accumulated in float in sample order, one thread each
 *   KP3  compaction        occupied bins packed to the front in cube order (block scan), sums
 *                          turned into mean colours and quantised counts
 *   KP4  first neighbours  for every bin the cheapest merge partner among the bins after it:
 *                          one warp per bin; the reference's scan is order dependent (a
 *                          candidate is skipped when its count term alone reaches the running
 *                          best), which is replayed exactly with one ballot per accepted
 *                          candidate instead of one step per candidate
 *   KP5  merging           one block: thread 0 owns the lazy-update heap; whenever the top bin's
 *                          partner is stale the whole block re-derives it (all candidates
 *                          evaluated in parallel, then a few ordered selection rounds)
 *   KP6  finish            surviving bins -> colours, duplicate removal on the sixel 0..100
 *                          scale, transparent pen, optional DIN99d copy, two principal axes by
 *                          power iteration (the sum over pens is a serial float chain, one
 *                          thread per channel; everything around it is parallel), projections,
 *                          stable sort of the pen table by the first projection
 */

#include <cuda_runtime.h>

#include <cfloat>

#include "device.h"
#include "prep_pixel.cuh"

#define FULL 0xffffffffu
#define NO_BIN 0xffffffffu
#define GONE 0xffffu            /* merge stamp of a bin that has been absorbed */
#define MERGE_THREADS 1024
#define SMALL_BINS 4096         /* up to here the merge keeps all of its state in shared memory */
#define N_COLS 255              /* colours asked of the merge (chafa-palette.c:938) */
#define RISKY_MAX 1024

enum { CTR_SAMPLES_A = 0, CTR_SAMPLES_B, CTR_RISKY, CTR_BINS, CTR_FLAT_WEIGHTS, CTR_FIND_CALLS, CTR_MAX = 16 };

/* Views into the workspace (dev_palette_build_work_bytes ()) */
struct Work
{
    uint32_t *bins_a, *bins_b;      /* max_bins x { sum0, sum1, sum2, count } */
    uint32_t *ctr;
    uint32_t *risky_ids;
    float *risky_sums;
    float *acc0, *acc1, *acc2, *cnt, *err;
    float *cand_bound, *cand_err;
    uint16_t *nearest, *tm, *mtm, *heap;
    float *heap_err;
};

__host__ __device__ static inline size_t
align16 (size_t n)
{
    return (n + 15) & ~(size_t) 15;
}

/* The head of the workspace -- both bin sets, the counters and the float sums of the bins beyond
 * float's exact range -- is what band owners of a sharded still exchange (bins and counters are
 * summed across the owners, the float sums travel from owner to owner in row order), so it may live
 * in a buffer of the caller's (@head); otherwise it is the start of @base. */
__host__ __device__ static inline size_t
head_bytes_for (int max_bins)
{
    return (size_t) max_bins * 32 + CTR_MAX * 4 + RISKY_MAX * 16;
}

__host__ __device__ static Work
carve (void *base, void *head, int max_bins)
{
    Work w;
    size_t m = (size_t) max_bins;
    uint8_t *p = (uint8_t *) (head ? head : base);
    w.bins_a = (uint32_t *) p; p += m * 16;
    w.bins_b = (uint32_t *) p; p += m * 16;
    w.ctr = (uint32_t *) p; p += CTR_MAX * 4;
    w.risky_sums = (float *) p; p += RISKY_MAX * 16;
    if (head)
        p = (uint8_t *) base;
    w.risky_ids = (uint32_t *) p; p += RISKY_MAX * 4;
    w.acc0 = (float *) p; p += m * 4;
    w.acc1 = (float *) p; p += m * 4;
    w.acc2 = (float *) p; p += m * 4;
    w.cnt = (float *) p; p += m * 4;
    w.err = (float *) p; p += m * 4;
    w.cand_bound = (float *) p; p += m * 4;
    w.cand_err = (float *) p; p += m * 4;
    w.nearest = (uint16_t *) p; p += align16 (m * 2);
    w.tm = (uint16_t *) p; p += align16 (m * 2);
    w.mtm = (uint16_t *) p; p += align16 (m * 2);
    w.heap = (uint16_t *) p; p += align16 ((m + 1) * 2);
    w.heap_err = (float *) p; p += align16 ((m + 1) * 4);
    return w;
}

size_t
dev_palette_build_work_bytes (int bits_per_ch)
{
    size_t m = (size_t) 1 << (bits_per_ch * 3);
    return m * 32 + CTR_MAX * 4 + RISKY_MAX * 20 + m * 28 + 3 * align16 (m * 2) + align16 ((m + 1) * 2)
        + align16 ((m + 1) * 4) + 64;
}

/* bytes of the head, all of which must be zero before KP1; the first dev_palette_build_reduce_words ()
 * 32-bit words of it are integer sums */
size_t
dev_palette_build_head_bytes (int bits_per_ch)
{
    return head_bytes_for (1 << (bits_per_ch * 3));
}

size_t
dev_palette_build_reduce_words (int bits_per_ch)
{
    return ((size_t) 1 << (bits_per_ch * 3)) * 8 + CTR_MAX;
}

/* ------------------------------------------------------------------------ */
/* KP1: sampling (chafa-palette.c:513-536)                                    */

__device__ __forceinline__ uint32_t
cube_index (uint32_t px, int bits_per_ch)
{
    uint32_t mask = (0xffu << (8 - bits_per_ch)) & 0xffu;
    /* Shift counts as the reference computes them; the middle one is negative for 3 bits per
     * channel, where the compiled reference shifts by the count modulo 32 */
    unsigned sh0 = (unsigned) (bits_per_ch * 2 - (8 - bits_per_ch)) & 31u;
    unsigned sh1 = (unsigned) (bits_per_ch - (8 - bits_per_ch)) & 31u;
    unsigned sh2 = (unsigned) (8 - bits_per_ch);
    uint32_t c0 = px & 0xff, c1 = (px >> 8) & 0xff, c2 = (px >> 16) & 0xff;
    return (((c0 & mask) << sh0) | ((c1 & mask) << sh1) | ((c2 & mask) >> sh2)) & ((1u << (bits_per_ch * 3)) - 1u);
}

/* @gate: NULL, or the sample count of the sparse pass -- the dense pass only runs below 256 */
__global__ void __launch_bounds__ (256)
palette_sample_kernel (const uint32_t *__restrict__ pixels, size_t first_pixel, size_t end_pixel, size_t step,
                       int bits_per_ch, int alpha_threshold, uint32_t *__restrict__ bins,
                       uint32_t *__restrict__ n_samples, const uint32_t *__restrict__ gate)
{
    if (gate && *gate >= 256u)
        return;

    /* the samples are pixels 0, step, 2 step ... of the whole image; this launch takes those in
     * [first_pixel, end_pixel) */
    const size_t k_first = (first_pixel + step - 1) / step, k_end = (end_pixel + step - 1) / step;
    unsigned local = 0;

    for (size_t k = k_first + (size_t) blockIdx.x * blockDim.x + threadIdx.x; k < k_end;
         k += (size_t) gridDim.x * blockDim.x)
    {
        uint32_t px = pixels [k * step];
        if ((int) (px >> 24) < alpha_threshold)
            continue;
        uint32_t *b = bins + (size_t) cube_index (px, bits_per_ch) * 4;
        atomicAdd (&b [0], px & 0xff);
        atomicAdd (&b [1], (px >> 8) & 0xff);
        atomicAdd (&b [2], (px >> 16) & 0xff);
        atomicAdd (&b [3], 1u);
        local++;
    }

    local = __reduce_add_sync (FULL, local);
    if ((threadIdx.x & 31) == 0 && local)
        atomicAdd (n_samples, local);
}

__device__ __forceinline__ bool
dense_pass_used (const uint32_t *ctr)
{
    return ctr [CTR_SAMPLES_A] < 256u;
}

/* ------------------------------------------------------------------------ */
/* KP2: bins beyond the exact range of float                                  */

/* Listed in cube order by one block, so that every band owner of a sharded still numbers them
 * alike (the float sums are handed from owner to owner by list position) */
__device__ __forceinline__ uint32_t block_exclusive_scan (uint32_t v, uint32_t *total, uint32_t *s_warp);

__global__ void __launch_bounds__ (1024)
palette_risky_kernel (Work w, int max_bins)
{
    __shared__ uint32_t s_warp [33];
    const uint32_t *bins = dense_pass_used (w.ctr) ? w.bins_b : w.bins_a;
    const int per = max_bins / 1024 > 0 ? max_bins / 1024 : 1;
    const int first = threadIdx.x * per;

    uint32_t mine = 0;
    for (int b = first; b < first + per && b < max_bins; b++)
        mine += (unsigned long long) bins [(size_t) b * 4 + 3] * 255u >= (1u << 24) ? 1u : 0u;

    uint32_t total;
    uint32_t k = block_exclusive_scan (mine, &total, s_warp);
    for (int b = first; b < first + per && b < max_bins; b++)
        if ((unsigned long long) bins [(size_t) b * 4 + 3] * 255u >= (1u << 24))
        {
            if (k < RISKY_MAX)
                w.risky_ids [k] = (uint32_t) b;
            k++;
        }
    if (threadIdx.x == 0)
        w.ctr [CTR_RISKY] = total;
}

/* The reference adds sample after sample into float sums; that equals the integer sum until a
 * sum passes 2^24.  Where it can, the float accumulation is redone exactly as the reference does
 * it: one thread per such bin walks the samples in order. */
__global__ void __launch_bounds__ (32)
palette_float_bins_kernel (Work w, const uint32_t *__restrict__ pixels, size_t first_pixel, size_t end_pixel,
                           size_t step, int bits_per_ch, int alpha_threshold)
{
    uint32_t n_risky = min (w.ctr [CTR_RISKY], (uint32_t) RISKY_MAX);
    if (dense_pass_used (w.ctr))
        step = 1;

    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_risky; k += gridDim.x * blockDim.x)
    {
        uint32_t mine = w.risky_ids [k];
        /* the sums so far: zero, or what the owners of the rows above have accumulated */
        float a0 = w.risky_sums [k * 4 + 0], a1 = w.risky_sums [k * 4 + 1], a2 = w.risky_sums [k * 4 + 2],
              count = w.risky_sums [k * 4 + 3];

        for (size_t i = (first_pixel + step - 1) / step * step; i < end_pixel; i += step)
        {
            uint32_t px = pixels [i];
            if ((int) (px >> 24) < alpha_threshold || cube_index (px, bits_per_ch) != mine)
                continue;
            a0 = __fadd_rn (a0, (float) (px & 0xff));
            a1 = __fadd_rn (a1, (float) ((px >> 8) & 0xff));
            a2 = __fadd_rn (a2, (float) ((px >> 16) & 0xff));
            count = __fadd_rn (count, 1.0f);
        }
        w.risky_sums [k * 4 + 0] = a0;
        w.risky_sums [k * 4 + 1] = a1;
        w.risky_sums [k * 4 + 2] = a2;
        w.risky_sums [k * 4 + 3] = count;
    }
}

/* ------------------------------------------------------------------------ */
/* KP3: occupied bins to the front, means and quantised counts                */

/* exclusive prefix sum of one value per thread over a block of 1024; returns the total too */
__device__ __forceinline__ uint32_t
block_exclusive_scan (uint32_t v, uint32_t *total, uint32_t *s_warp /* 33 words */)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        uint32_t t = __shfl_up_sync (FULL, inc, d);
        if (lane >= d)
            inc += t;
    }
    if (lane == 31)
        s_warp [warp] = inc;
    __syncthreads ();
    if (warp == 0)
    {
        uint32_t wv = lane < (int) (blockDim.x >> 5) ? s_warp [lane] : 0u;
        uint32_t winc = wv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            uint32_t t = __shfl_up_sync (FULL, winc, d);
            if (lane >= d)
                winc += t;
        }
        s_warp [lane] = winc - wv;
        if (lane == 31)
            s_warp [32] = winc;
    }
    __syncthreads ();
    *total = s_warp [32];
    uint32_t r = s_warp [warp] + inc - v;
    __syncthreads ();
    return r;
}

__global__ void __launch_bounds__ (1024)
palette_compact_kernel (Work w, int max_bins)
{
    __shared__ uint32_t s_warp [33];
    const uint32_t *bins = dense_pass_used (w.ctr) ? w.bins_b : w.bins_a;
    const uint32_t n_risky = min (w.ctr [CTR_RISKY], (uint32_t) RISKY_MAX);
    const int per = max_bins / 1024 > 0 ? max_bins / 1024 : 1;      /* bins per thread, consecutive */
    const int first = threadIdx.x * per;

    uint32_t mine = 0;
    for (int b = first; b < first + per && b < max_bins; b++)
        mine += bins [(size_t) b * 4 + 3] != 0u;

    uint32_t n_bins;
    uint32_t at = block_exclusive_scan (mine, &n_bins, s_warp);

    /* Weights and the count transform (chafa-palette.c:586-603 with 255 colours asked for):
     * luma weights and square-rooted counts, unless there are more than 33 bins per colour */
    float weight = fminf (0.9f, __fdiv_rn ((float) N_COLS * 1.0f, (float) n_bins));
    bool flat = n_bins > 0 && weight < .03f;

    for (int b = first; b < first + per && b < max_bins; b++)
    {
        const uint32_t *s = bins + (size_t) b * 4;
        if (s [3] == 0u)
            continue;
        float f0 = (float) s [0], f1 = (float) s [1], f2 = (float) s [2], fc = (float) s [3];
        for (uint32_t k = 0; k < n_risky; k++)
            if (w.risky_ids [k] == (uint32_t) b)
            {
                f0 = w.risky_sums [k * 4];
                f1 = w.risky_sums [k * 4 + 1];
                f2 = w.risky_sums [k * 4 + 2];
                fc = w.risky_sums [k * 4 + 3];
            }
        float inv = __fdiv_rn (1.0f, fc);
        w.acc0 [at] = __fmul_rn (f0, inv);
        w.acc1 [at] = __fmul_rn (f1, inv);
        w.acc2 [at] = __fmul_rn (f2, inv);
        w.cnt [at] = flat ? fc : __fsqrt_rn (fc);
        w.tm [at] = 0;
        w.mtm [at] = 0;
        at++;
    }

    if (threadIdx.x == 0)
    {
        w.ctr [CTR_BINS] = n_bins;
        w.ctr [CTR_FLAT_WEIGHTS] = flat ? 1u : 0u;
    }
}

/* ------------------------------------------------------------------------ */
/* Cost of merging two bins                                                   */

struct BinView
{
    float a0, a1, a2, n;
};

/* @bound: the count term n1 n2 / (n1 + n2); a candidate is not even looked at when it reaches
 * the running best.  @cost: the full merge cost.  Both non-negative. */
__device__ __forceinline__ void
merge_cost (const BinView &me, const BinView &other, float w0, float w1, float w2, float *bound, float *cost)
{
    float count_term = __fdiv_rn (__fmul_rn (other.n, me.n), __fadd_rn (other.n, me.n));
    float half = __fmul_rn (count_term, .5f);
    float d0 = __fsub_rn (other.a0, me.a0), d1 = __fsub_rn (other.a1, me.a1), d2 = __fsub_rn (other.a2, me.a2);

    float t = __fadd_rn (__fadd_rn (__fmul_rn (__fmul_rn (d0, d0), w0), __fmul_rn (__fmul_rn (d1, d1), w1)),
                         __fmul_rn (__fmul_rn (d2, d2), w2));
    float c = __fmul_rn (t, half);

    /* the same difference seen along luma and the two chroma axes */
    const float axes [3] [3] = { { .299f, .587f, .114f }, { -0.14713f, -0.28886f, 0.436f }, { 0.615f, -0.51499f, -0.10001f } };
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        float q0 = __fmul_rn (axes [j] [0], d0), q1 = __fmul_rn (axes [j] [1], d1), q2 = __fmul_rn (axes [j] [2], d2);
        float q = __fadd_rn (__fadd_rn (__fmul_rn (q0, q0), __fmul_rn (q1, q1)), __fmul_rn (q2, q2));
        c = __fadd_rn (c, __fmul_rn (q, half));
    }
    *bound = count_term;
    *cost = c;
}

/* The reference walks the candidates in order with a running best E (initially FLT_MAX): a
 * candidate replaces it iff its count term and its cost are both below E (the partial sums the
 * reference also tests are monotone, so they decide nothing more).  Hence one number per
 * candidate: accept iff max (bound, cost) < E, and E becomes the cost. */
__device__ __forceinline__ float
accept_level (float bound, float cost)
{
    return fmaxf (bound, cost);
}

__device__ __forceinline__ void
weights_of (const Work &w, float *w0, float *w1, float *w2)
{
    bool flat = w.ctr [CTR_FLAT_WEIGHTS] != 0u;
    *w0 = flat ? 1.0f : .299f;
    *w1 = flat ? 1.0f : .587f;
    *w2 = flat ? 1.0f : .114f;
}

/* ------------------------------------------------------------------------ */
/* KP4: first partner of every bin (all bins alive)                           */

__global__ void __launch_bounds__ (256)
palette_first_partner_kernel (Work w)
{
    const int n = (int) w.ctr [CTR_BINS];
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    float w0, w1, w2;
    weights_of (w, &w0, &w1, &w2);

    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += n_warps)
    {
        BinView me = { w.acc0 [i], w.acc1 [i], w.acc2 [i], w.cnt [i] };
        float best = FLT_MAX;
        int partner = 0;

        for (int base = i + 1; base < n; base += 32)
        {
            int j = base + lane;
            float level = FLT_MAX, cost = 0.0f;
            if (j < n)
            {
                BinView other = { w.acc0 [j], w.acc1 [j], w.acc2 [j], w.cnt [j] };
                float bound;
                merge_cost (me, other, w0, w1, w2, &bound, &cost);
                level = accept_level (bound, cost);
            }
            /* the first lane below the running best is accepted by the sequential walk too
             * (nothing before it was); lanes after it are re-tested against the new best */
            unsigned pending = __ballot_sync (FULL, level < best);
            while (pending)
            {
                int l = __ffs (pending) - 1;
                best = __shfl_sync (FULL, cost, l);
                partner = base + l;
                pending = __ballot_sync (FULL, level < best) & ~((2u << l) - 1u);
            }
        }
        if (lane == 0)
        {
            w.err [i] = best;
            w.nearest [i] = (uint16_t) partner;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* KP5: merging                                                               */

struct MergeState
{
    float *acc0, *acc1, *acc2, *cnt, *err, *cand_level, *cand_cost;
    uint16_t *nearest, *tm, *mtm, *heap;
    float *heap_err;            /* heap_err [l] = err [heap [l]]: a bin's cost only changes while it sits at the top */
};

/* restore the heap below position 1 for bin @b (chafa-palette.c:671-680; 16-bit positions as
 * in the reference).  The costs of the bins in the heap are kept beside their positions, so a level
 * costs one round of shared-memory loads instead of two dependent ones (position, then cost): the
 * heap is walked by one thread while the block waits. */
__device__ __forceinline__ void
heap_sink (const MergeState &s, uint16_t b)
{
    const float e = s.err [b];
    const uint16_t count = s.heap [0];
    uint16_t l, l2;
    for (l = 1, l2 = 2; l2 <= count; l = l2, l2 = (uint16_t) (l * 2))
    {
        float e2 = s.heap_err [l2];
        uint16_t h = s.heap [l2];
        if (l2 < count)
        {
            const float e3 = s.heap_err [l2 + 1];
            const uint16_t h3 = s.heap [l2 + 1];
            if (e2 > e3)
            {
                l2++;
                e2 = e3;
                h = h3;
            }
        }
        if (e <= e2)
            break;
        s.heap [l] = h;
        s.heap_err [l] = e2;
    }
    s.heap [l] = b;
    s.heap_err [l] = e;
}

/* The whole block re-derives the partner of bin @b among the live bins after it.
 * Returns (uniformly) the cost and the partner. */
__device__ __forceinline__ void
block_find_partner (const MergeState &s, int n, int b, float w0, float w1, float w2, uint32_t *s_slot,
                    float *best_out, int *partner_out)
{
    const int tid = threadIdx.x;
    BinView me = { s.acc0 [b], s.acc1 [b], s.acc2 [b], s.cnt [b] };

    for (int j = b + 1 + tid; j < n; j += MERGE_THREADS)
    {
        float level = FLT_MAX, cost = 0.0f;
        if (s.mtm [j] != GONE)
        {
            BinView other = { s.acc0 [j], s.acc1 [j], s.acc2 [j], s.cnt [j] };
            float bound;
            merge_cost (me, other, w0, w1, w2, &bound, &cost);
            level = accept_level (bound, cost);
        }
        s.cand_level [j] = level;
        s.cand_cost [j] = cost;
    }

    float best = FLT_MAX;
    int partner = 0, pos = b;

    /* Selection rounds: the first candidate after @pos below the running best is what the
     * sequential walk accepts next.  Three slots in rotation, one barrier per round. */
    for (int round = 0; ; round++)
    {
        uint32_t local = NO_BIN;
        int j = b + 1 + tid;
        if (pos >= j)
            j += ((pos - j) / MERGE_THREADS + 1) * MERGE_THREADS;
        for (; j < n; j += MERGE_THREADS)
            if (s.cand_level [j] < best)
            {
                local = (uint32_t) j;
                break;
            }
        uint32_t *slot = &s_slot [round % 3];
        local = __reduce_min_sync (FULL, local);
        if ((tid & 31) == 0 && local != NO_BIN)
            atomicMin (slot, local);
        __syncthreads ();
        uint32_t next = *slot;
        if (tid == 0)
            s_slot [(round + 2) % 3] = NO_BIN;      /* last read before this round's barrier */
        if (next == NO_BIN)
            break;
        best = s.cand_cost [next];
        partner = (int) next;
        pos = (int) next;
    }
    *best_out = best;
    *partner_out = partner;
}

__global__ void __launch_bounds__ (MERGE_THREADS)
palette_merge_kernel (Work w)
{
    extern __shared__ __align__ (16) unsigned char smem [];
    __shared__ uint32_t s_slot [3];
    __shared__ int s_cmd, s_bin;

    const int tid = threadIdx.x;
    const int n = (int) w.ctr [CTR_BINS];
    const bool in_smem = n <= SMALL_BINS;
    float w0, w1, w2;
    weights_of (w, &w0, &w1, &w2);

    MergeState s;
    if (in_smem)
    {
        float *f = (float *) smem;
        s.acc0 = f; s.acc1 = f + SMALL_BINS; s.acc2 = f + 2 * SMALL_BINS; s.cnt = f + 3 * SMALL_BINS;
        s.err = f + 4 * SMALL_BINS; s.cand_level = f + 5 * SMALL_BINS; s.cand_cost = f + 6 * SMALL_BINS;
        uint16_t *h = (uint16_t *) (f + 7 * SMALL_BINS);
        s.nearest = h; s.tm = h + SMALL_BINS; s.mtm = h + 2 * SMALL_BINS; s.heap = h + 3 * SMALL_BINS;
        s.heap_err = (float *) (((uintptr_t) (s.heap + SMALL_BINS + 8) + 15) & ~(uintptr_t) 15);
        for (int i = tid; i < n; i += MERGE_THREADS)
        {
            s.acc0 [i] = w.acc0 [i]; s.acc1 [i] = w.acc1 [i]; s.acc2 [i] = w.acc2 [i]; s.cnt [i] = w.cnt [i];
            s.err [i] = w.err [i]; s.nearest [i] = w.nearest [i]; s.tm [i] = 0; s.mtm [i] = 0;
        }
    }
    else
    {
        s.acc0 = w.acc0; s.acc1 = w.acc1; s.acc2 = w.acc2; s.cnt = w.cnt; s.err = w.err;
        s.cand_level = w.cand_bound; s.cand_cost = w.cand_err;
        s.nearest = w.nearest; s.tm = w.tm; s.mtm = w.mtm; s.heap = w.heap;
        s.heap_err = w.heap_err;
    }
    if (tid < 3)
        s_slot [tid] = NO_BIN;
    __syncthreads ();

    const int surplus = n - N_COLS;
    int merged = 0;             /* thread 0's */
    uint32_t find_calls = 0;

    if (tid == 0)
    {
        /* heap keyed by each bin's cost, built by inserting the bins in order (chafa-palette.c:609-626):
         * the layout decides which of several equally cheap bins is merged first */
        s.heap [0] = 0;
        for (int i = 0; i < n; i++)
        {
            float e = s.err [i];
            uint16_t l = ++s.heap [0], l2;
            for (; l > 1; l = l2)
            {
                l2 = l >> 1;
                const float e2 = s.heap_err [l2];
                if (e2 <= e)
                    break;
                s.heap [l] = s.heap [l2];
                s.heap_err [l] = e2;
            }
            s.heap [l] = (uint16_t) i;
            s.heap_err [l] = e;
        }
    }

    while (surplus > 0)
    {
        if (tid == 0)
        {
            int cmd = 0;
            while (merged < surplus)
            {
                uint16_t b = s.heap [1];
                uint16_t seen = s.tm [b], stamp = s.mtm [b];

                if (seen >= stamp && s.mtm [s.nearest [b]] <= seen)
                {
                    /* both up to date: absorb the partner (chafa-palette.c:683-703) */
                    uint16_t p = s.nearest [b];
                    float n1 = s.cnt [b], n2 = s.cnt [p];
                    float total = __fadd_rn (n1, n2);
                    float inv = __fdiv_rn (1.0f, total);
                    s.acc0 [b] = __fmul_rn (rintf (__fadd_rn (__fmul_rn (s.acc0 [p], n2), __fmul_rn (s.acc0 [b], n1))), inv);
                    s.acc1 [b] = __fmul_rn (rintf (__fadd_rn (__fmul_rn (s.acc1 [p], n2), __fmul_rn (s.acc1 [b], n1))), inv);
                    s.acc2 [b] = __fmul_rn (rintf (__fadd_rn (__fmul_rn (s.acc2 [p], n2), __fmul_rn (s.acc2 [b], n1))), inv);
                    s.cnt [b] = total;
                    s.mtm [b] = (uint16_t) ++merged;
                    s.mtm [p] = GONE;
                    continue;
                }
                if (stamp == GONE)
                {
                    /* absorbed since it was queued: drop it */
                    b = s.heap [1] = s.heap [s.heap [0]];
                    s.heap [0]--;
                    heap_sink (s, b);           /* (reads err [b], which equals the cost stored with it) */
                    continue;
                }
                cmd = 1;
                s_bin = b;
                break;
            }
            s_cmd = cmd;
        }
        __syncthreads ();
        if (!s_cmd)
            break;

        int b = s_bin, partner;
        float best;
        block_find_partner (s, n, b, w0, w1, w2, s_slot, &best, &partner);
        if (tid == 0)
        {
            s.err [b] = best;
            s.nearest [b] = (uint16_t) partner;
            s.tm [b] = (uint16_t) merged;
            heap_sink (s, (uint16_t) b);
            find_calls++;
        }
    }
    __syncthreads ();

    if (in_smem)
        for (int i = tid; i < n; i += MERGE_THREADS)
        {
            w.acc0 [i] = s.acc0 [i]; w.acc1 [i] = s.acc1 [i]; w.acc2 [i] = s.acc2 [i];
            w.mtm [i] = s.mtm [i];
        }
    if (tid == 0)
        w.ctr [CTR_FIND_CALLS] = find_calls;
}

/* ------------------------------------------------------------------------ */
/* KP6: colours, clean-up, pen table                                          */

#define FINISH_THREADS 256

__device__ __forceinline__ float
dot3_rn (float a0, float a1, float a2, float b0, float b1, float b2)
{
    return __fadd_rn (__fadd_rn (__fmul_rn (a0, b0), __fmul_rn (a1, b1)), __fmul_rn (a2, b2));
}

/* One principal axis of the @n centred pen colours in s_v by power iteration
 * (chafa-pca.c:29-74): start vector normalize (.11, .23, .51), at most 1000 rounds, stop when
 * the squared residual |r lambda - s|^2 drops below 0.0001^2.  s = sum_i v_i (v_i . r) is a
 * float sum in pen order -- one thread per channel adds it up serially; the products are
 * computed by all threads.  Every thread returns the same axis. */
__device__ void
principal_axis (int n, const float (*s_v) [256], float (*s_p) [256], float *s_sum, float *axis)
{
    const int tid = threadIdx.x;
    float r0, r1, r2;
    {
        const float c0 = .11f, c1 = .23f, c2 = .51f;
        float inv = __fdiv_rn (1.0f, __fsqrt_rn (dot3_rn (c0, c1, c2, c0, c1, c2)));
        r0 = __fmul_rn (c0, inv);
        r1 = __fmul_rn (c1, inv);
        r2 = __fmul_rn (c2, inv);
    }
    const float stop = __fmul_rn (0.0001f, 0.0001f);

    if (n <= 0)
    {
        axis [0] = axis [1] = axis [2] = 0.0f;
        return;
    }

    /* The reference always runs into its iteration limit here: its stopping test compares an
     * absolute residual of 1e-4 with sums of the order of 1e9, which float cannot resolve.  The
     * iteration is a deterministic map of r, though, so once r repeats -- the same three floats as
     * one or two iterations earlier -- the remaining iterations are known without running them: a
     * fixed point stays, a cycle of two alternates and the limit's parity picks the member.  (Both
     * members' residuals have been tested by then and did not stop the loop.)  Typically some
     * tens of iterations instead of 1000. */
    float p0 = r0, p1 = r1, p2 = r2;            /* r after the previous iteration */
    float q0 = r0, q1 = r1, q2 = r2;            /* ... and after the one before */
    const int limit = 1000;

    for (int it = 0; it < limit; it++)
    {
        if (tid < n)
        {
            float v0 = s_v [0] [tid], v1 = s_v [1] [tid], v2 = s_v [2] [tid];
            float u = dot3_rn (v0, v1, v2, r0, r1, r2);
            s_p [0] [tid] = __fmul_rn (v0, u);
            s_p [1] [tid] = __fmul_rn (v1, u);
            s_p [2] [tid] = __fmul_rn (v2, u);
        }
        __syncthreads ();
        if (tid < 3)
        {
            float acc = 0.0f;
            for (int i = 0; i < n; i++)
                acc = __fadd_rn (acc, s_p [tid] [i]);
            s_sum [tid] = acc;
        }
        __syncthreads ();
        float s0 = s_sum [0], s1 = s_sum [1], s2 = s_sum [2];
        float lambda = dot3_rn (s0, s1, s2, r0, r1, r2);
        float t0 = __fsub_rn (__fmul_rn (r0, lambda), s0), t1 = __fsub_rn (__fmul_rn (r1, lambda), s1),
              t2 = __fsub_rn (__fmul_rn (r2, lambda), s2);
        float residual = __fadd_rn (__fadd_rn (__fmul_rn (t1, t1), __fmul_rn (t0, t0)), __fmul_rn (t2, t2));
        float mag2 = dot3_rn (s0, s1, s2, s0, s1, s2);
        q0 = p0; q1 = p1; q2 = p2;
        p0 = r0; p1 = r1; p2 = r2;
        if (mag2 == 0.0f)
        {
            r0 = r1 = r2 = 0.0f;
        }
        else
        {
            float inv = __fdiv_rn (1.0f, __fsqrt_rn (mag2));
            r0 = __fmul_rn (s0, inv);
            r1 = __fmul_rn (s1, inv);
            r2 = __fmul_rn (s2, inv);
        }
        if (stop > residual)
            break;
        const bool same1 = __float_as_uint (r0) == __float_as_uint (p0) && __float_as_uint (r1) == __float_as_uint (p1)
            && __float_as_uint (r2) == __float_as_uint (p2);
        const bool same2 = __float_as_uint (r0) == __float_as_uint (q0) && __float_as_uint (r1) == __float_as_uint (q1)
            && __float_as_uint (r2) == __float_as_uint (q2);
        if (same1)
            break;                              /* fixed point */
        if (same2 && it >= 2)
        {
            /* r_it = r_(it - 2): after the last iteration (limit - 1) it is r_it or r_(it - 1) */
            if (((limit - 1 - it) & 1) != 0)
            {
                r0 = p0;
                r1 = p1;
                r2 = p2;
            }
            break;
        }
    }
    axis [0] = r0;
    axis [1] = r1;
    axis [2] = r2;
}

__device__ __forceinline__ int
project_fixed (const int *v, const int *axis, int mul)
{
    long long d = (long long) v [0] * axis [0] + (long long) v [1] * axis [1] + (long long) v [2] * axis [2];
    return (int) ((d * mul) / ((1 << 14) / 32));
}

/* sixel colour components run 0..100; two pens that agree there are one pen (chafa-palette.c:762-805) */
__device__ __forceinline__ int
sixel_scale_distance (uint32_t a, uint32_t b)
{
    int diff = 0;
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        int t = (int) (((a >> (8 * c)) & 0xff) * 100) / 256 - (int) (((b >> (8 * c)) & 0xff) * 100) / 256;
        diff += t * t;
    }
    return diff;
}

__global__ void __launch_bounds__ (FINISH_THREADS)
palette_finish_kernel (Work w, DevPaletteBuild p)
{
    __shared__ uint32_t s_rgb [PAL_INDEX_MAX], s_cs [PAL_INDEX_MAX];
    __shared__ uint32_t s_scan [FINISH_THREADS + 1];
    __shared__ int s_n_colors;
    __shared__ float s_v [3] [256], s_prod [3] [256], s_sum [3];
    __shared__ int s_key [256];

    const int tid = threadIdx.x;
    const int n = (int) w.ctr [CTR_BINS];

    /* start from the caller's table (the fixed palette with the canvas's FG / BG pens) */
    for (int i = tid; i < PAL_INDEX_MAX; i += FINISH_THREADS)
        s_rgb [i] = p.base_colors [i];

    /* surviving bins in cube order -> colours, truncated (chafa-palette.c:705-721) */
    const int per = (n + FINISH_THREADS - 1) / FINISH_THREADS;
    const int first = tid * per;
    uint32_t alive = 0;
    for (int i = first; i < first + per && i < n; i++)
        alive += w.mtm [i] != GONE;
    s_scan [tid + 1] = alive;
    __syncthreads ();
    if (tid == 0)
    {
        s_scan [0] = 0;
        for (int i = 1; i <= FINISH_THREADS; i++)
            s_scan [i] += s_scan [i - 1];
    }
    __syncthreads ();
    {
        uint32_t at = s_scan [tid];
        for (int i = first; i < first + per && i < n; i++)
            if (w.mtm [i] != GONE)
            {
                if (at < 256u)
                    s_rgb [at] = ((uint32_t) (int) w.acc0 [i] & 0xff) | (((uint32_t) (int) w.acc1 [i] & 0xff) << 8)
                        | (((uint32_t) (int) w.acc2 [i] & 0xff) << 16) | 0xff000000u;
                at++;
            }
    }
    __syncthreads ();

    if (tid == 0)
    {
        int n_colors = (int) min (s_scan [FINISH_THREADS], 256u);
        int best_diff = 0x7fffffff, best_pair = 1, j = 1;

        for (int i = 1; i < n_colors; i++)
        {
            int diff = sixel_scale_distance (s_rgb [j - 1], s_rgb [i]);
            if (diff == 0)
                continue;
            if (diff < best_diff)
            {
                best_pair = j - 1;
                best_diff = diff;
            }
            s_rgb [j++] = s_rgb [i];
        }
        n_colors = j;

        /* the transparent pen's slot: appended, or in place of one of the closest pair */
        if (p.transparent_index < 256)
        {
            if (n_colors < 256)
                s_rgb [n_colors++] = s_rgb [p.transparent_index];
            else
                s_rgb [best_pair] = s_rgb [p.transparent_index];
        }
        s_n_colors = n_colors;
    }
    __syncthreads ();
    const int n_colors = s_n_colors;

    for (int i = tid; i < PAL_INDEX_MAX; i += FINISH_THREADS)
    {
        uint32_t c = s_rgb [i];
        if (p.color_space == 1 && i < n_colors)
            c = rgb_to_din99d (c);
        s_cs [i] = c;
    }
    __syncthreads ();

    /* pen table: pens [0, n_colors) without the transparent one, centred, projected on two axes */
    const bool is_pen = tid < n_colors && tid < 256 && tid != p.transparent_index;
    uint32_t n_before, n_pens;
    {
        /* pens are a prefix of the indices minus at most one: position = index - (index > transparent) */
        n_pens = (uint32_t) min (n_colors, 256);
        if (p.transparent_index < (int) n_pens)
            n_pens--;
        n_before = (uint32_t) tid - (tid > p.transparent_index ? 1u : 0u);
    }
    const uint32_t my_color = s_cs [tid < PAL_INDEX_MAX ? tid : 0] & 0x00ffffffu;

    if (is_pen)
    {
        s_v [0] [n_before] = __fmul_rn ((float) (my_color & 0xff), 32.0f);
        s_v [1] [n_before] = __fmul_rn ((float) ((my_color >> 8) & 0xff), 32.0f);
        s_v [2] [n_before] = __fmul_rn ((float) ((my_color >> 16) & 0xff), 32.0f);
    }
    __syncthreads ();
    if (tid < 3)
    {
        float acc = 0.0f;
        for (uint32_t i = 0; i < n_pens; i++)
            acc = __fadd_rn (acc, s_v [tid] [i]);
        s_sum [tid] = n_pens ? __fmul_rn (acc, __fdiv_rn (1.0f, (float) n_pens)) : 0.0f;
    }
    __syncthreads ();
    const float mean [3] = { s_sum [0], s_sum [1], s_sum [2] };
    __syncthreads ();
    if (is_pen)
        for (int c = 0; c < 3; c++)
            s_v [c] [n_before] = __fsub_rn (s_v [c] [n_before], mean [c]);
    __syncthreads ();

    float axis [2] [3];
    principal_axis ((int) n_pens, s_v, s_prod, s_sum, axis [0]);
    if (is_pen)
    {
        /* remove the first component from every pen (chafa-pca.c:76-96) */
        float v0 = s_v [0] [n_before], v1 = s_v [1] [n_before], v2 = s_v [2] [n_before];
        float score = dot3_rn (v0, v1, v2, axis [0] [0], axis [0] [1], axis [0] [2]);
        s_v [0] [n_before] = __fsub_rn (v0, __fmul_rn (score, axis [0] [0]));
        s_v [1] [n_before] = __fsub_rn (v1, __fmul_rn (score, axis [0] [1]));
        s_v [2] [n_before] = __fsub_rn (v2, __fmul_rn (score, axis [0] [2]));
    }
    __syncthreads ();
    principal_axis ((int) n_pens, s_v, s_prod, s_sum, axis [1]);

    /* fixed point, 5 fractional bits, rounded to nearest even like cvtss2si */
    int ev [2] [3], avg [3], mul [2];
    for (int k = 0; k < 2; k++)
    {
        for (int c = 0; c < 3; c++)
            ev [k] [c] = __float2int_rn (__fmul_rn (axis [k] [c], 32.0f));
        int m = ev [k] [0] * ev [k] [0] + ev [k] [1] * ev [k] [1] + ev [k] [2] * ev [k] [2];
        mul [k] = (1 << 14) / max (m, 1);
    }
    for (int c = 0; c < 3; c++)
        avg [c] = __float2int_rn (__fmul_rn (mean [c], 32.0f));

    int key = 0, key2 = 0;
    if (is_pen)
    {
        int v [3] = { (int) (my_color & 0xff) * 32 - avg [0], (int) ((my_color >> 8) & 0xff) * 32 - avg [1],
                      (int) ((my_color >> 16) & 0xff) * 32 - avg [2] };
        key = project_fixed (v, ev [0], mul [0]);
        key2 = project_fixed (v, ev [1], mul [1]);
        s_key [n_before] = key;
    }
    __syncthreads ();

    /* table sorted by the first projection; equal keys keep pen order (the reference's qsort is
     * glibc's merge sort) */
    DevColorTable *t = p.table_out;
    t->entries [tid].v [0] = t->entries [tid].v [1] = t->entries [tid].pen = 0;
    t->pens [tid] = is_pen ? my_color : 0xffffffffu;
    __syncthreads ();
    if (is_pen)
    {
        uint32_t rank = 0;
        for (uint32_t i = 0; i < n_pens; i++)
            rank += s_key [i] < key || (s_key [i] == key && i < n_before);
        t->entries [rank].v [0] = key;
        t->entries [rank].v [1] = key2;
        t->entries [rank].pen = tid;
    }
    if (tid == 0)
    {
        t->n_entries = (int) n_pens;
        for (int k = 0; k < 2; k++)
        {
            t->eigen_mul [k] = mul [k];
            for (int c = 0; c < 3; c++)
                t->eigenvectors [k] [c] = ev [k] [c];
        }
        for (int c = 0; c < 3; c++)
            t->average [c] = avg [c];
        p.result_out->n_colors = n_colors;
        p.result_out->n_bins = n;
        p.result_out->n_partner_searches = (int) w.ctr [CTR_FIND_CALLS];
    }
    for (int i = tid; i < PAL_INDEX_MAX; i += FINISH_THREADS)
    {
        p.pal_colors_out [i] = s_cs [i];
        p.result_out->colors_rgb [i] = s_rgb [i];
    }
}

/* ------------------------------------------------------------------------ */
/* Launcher                                                                   */

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

/* Queues the whole build on @stream; nothing is waited for.  Returns 0 or a cudaError_t. */
static Work
work_of (const DevPaletteBuild *p, int *max_bins_out)
{
    const int max_bins = 1 << (p->bits_per_ch * 3);
    *max_bins_out = max_bins;
    return carve (p->work, p->head, max_bins);
}

/* KP1 over the pixels [sample_first, sample_end) of the image */
int
dev_launch_palette_sample (const DevPaletteBuild *p, void *stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    int max_bins;
    Work w = work_of (p, &max_bins);
    cudaError_t err;

    if ((err = cudaMemsetAsync (w.bins_a, 0, head_bytes_for (max_bins), st)) != cudaSuccess)
        return (int) err;

    size_t n_mine = p->sample_end - p->sample_first;
    size_t n_sparse = (n_mine + p->step - 1) / p->step;
    palette_sample_kernel<<<grid_for (n_sparse, 256), 256, 0, st>>> (p->pixels, p->sample_first, p->sample_end, p->step,
                                                                    p->bits_per_ch, p->alpha_threshold, w.bins_a,
                                                                    w.ctr + CTR_SAMPLES_A, nullptr);
    /* too many transparent pixels: sample densely (chafa-palette.c:561-569).  Gated by this
     * launch's own count: a band owner that saw 256 samples knows the whole image has, and if the
     * whole image has not, no owner has and all of them ran the dense pass. */
    palette_sample_kernel<<<grid_for (n_mine, 256), 256, 0, st>>> (p->pixels, p->sample_first, p->sample_end, 1,
                                                                  p->bits_per_ch, p->alpha_threshold, w.bins_b,
                                                                  w.ctr + CTR_SAMPLES_B, w.ctr + CTR_SAMPLES_A);
    dev_count_launch (2);
    return (int) cudaGetLastError ();
}

/* KP2 over the same pixels, continuing from the float sums in the head */
int
dev_launch_palette_float_bins (const DevPaletteBuild *p, void *stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    int max_bins;
    Work w = work_of (p, &max_bins);

    palette_risky_kernel<<<1, 1024, 0, st>>> (w, max_bins);
    palette_float_bins_kernel<<<8, 32, 0, st>>> (w, p->pixels, p->sample_first, p->sample_end, p->step, p->bits_per_ch,
                                                 p->alpha_threshold);
    dev_count_launch (2);
    return (int) cudaGetLastError ();
}

/* KP3-KP6 from the (complete) head */
int
dev_launch_palette_finish (const DevPaletteBuild *p, void *stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    int max_bins;
    Work w = work_of (p, &max_bins);
    cudaError_t err;

    palette_compact_kernel<<<1, 1024, 0, st>>> (w, max_bins);
    palette_first_partner_kernel<<<max (max_bins / 8, 1), 256, 0, st>>> (w);

    /* room for SMALL_BINS bins in shared memory, used whenever the image occupies no more (decided
     * in the kernel from the bin count, whatever the cube resolution) */
    size_t smem = (size_t) SMALL_BINS * (7 * 4 + 3 * 2) + ((size_t) SMALL_BINS + 8) * 2 + ((size_t) SMALL_BINS + 8) * 4 + 16;
    err = cudaFuncSetAttribute (palette_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess)
        return (int) err;
    palette_merge_kernel<<<1, MERGE_THREADS, smem, st>>> (w);
    palette_finish_kernel<<<1, FINISH_THREADS, 0, st>>> (w, *p);
    dev_count_launch (4);
    return (int) cudaGetLastError ();
}

int
dev_launch_palette_build (const DevPaletteBuild *p, void *stream)
{
    int err = dev_launch_palette_sample (p, stream);
    if (!err)
        err = dev_launch_palette_float_bins (p, stream);
    if (!err)
        err = dev_launch_palette_finish (p, stream);
    return err;
}
