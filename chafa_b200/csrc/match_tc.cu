/* match_tc.cu -- kernel K3t: per-cell symbol and colour selection, one THREAD per cell,
 * Hamming distances on the 5th-generation tensor cores.
 *
 * Covers the default configuration of the fast path (work factor < 0.8, average colour
 * extractor, colours from the image, error against the extracted colours; narrow AND wide
 * symbols in one launch).  Everything else stays on the kernels in match.cu.
 *
 * Reference behaviour restated (same files as match.cu):
 *   chafa/internal/chafa-work-cell.c:76-102, 121-143, 293-338
 *   chafa/internal/chafa-avx2.c:39-81, 83-165
 *   chafa/chafa-symbol-map.c:1153-1332 (candidate order)
 *   chafa/internal/chafa-symbol-renderer.c:187-471, 656-753
 *
 * Formulation.  With pixels coded -1 (covered) / +1 (not covered), the dot product of a
 * cell's outline bitmap and a symbol's bitmap over the 64 pixels is D = 64 - 2 hd, and the
 * inverted symbol scores -D.  So the distance table "every cell x every symbol" is one int8
 * matrix product  A (cells x 64) . B^T (symbols x 64), which tcgen05.mma (kind::i8) computes
 * 128 cells x 128 symbols at a time into tensor memory.  A tensor-memory lane is one row
 * of the product, i.e. one cell, and tcgen05.ld hands each thread the row of its lane:
 * the thread that owns a cell gets that cell's distances to 128 consecutive symbols in 64
 * registers (two 16-bit values each) without any shuffling.
 *
 * The reference's candidate list is the k smallest of the 2N keys (distance, sequence
 * number); since a symbol's better orientation has distance (64 - |D|) / 2, that is the k
 * largest |D| with ties in sequence order.  Two sweeps over the product (the MMAs are
 * simply issued twice, they cost a few percent of the ALU work):
 *   sweep 1: running 16-bit SIMD max / min per register position = 16 partitions of the
 *            symbol table; the k-th largest partition maximum T is a lower bound of the
 *            k-th largest |D|;
 *   sweep 2: per chunk of 128 symbols, every value with |D| >= T sets a bit in a per-thread hit
 *            mask (one SIMD add, one SIMD compare with two predicate outputs, two predicated
 *            ORs per register).  The flagged symbols -- a handful -- are ranked exactly by
 *            (distance, sequence number) right after the chunk, and once k candidates are
 *            known T rises to "strictly better than the k-th", because a later symbol loses
 *            every tie.  Massive ties (flat cells) therefore cost one chunk, not the table.
 * A degenerate bound (fewer than k partitions hold symbols) falls back to a sequential scan.
 *
 * Candidates are scored with the closed-form error (see match.cu, K3x): masked channel sums
 * by IDP.4A of planar pixel registers against the symbol's -1/+1 bytes, which are the rows
 * of the B operand already in shared memory.
 *
 * Block = 4 groups of 128 threads; a group owns 128 tensor-memory columns and walks tiles of
 * 127 cells (+ 1 cell of overlap, because a wide symbol spans the pair (c, c + 1)). */

#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <cstdlib>
#include <type_traits>

#include "device.h"
#include "palette.cuh"

#define TC_GROUPS 4
#define TC_GROUP_THREADS 128
#define TC_THREADS (TC_GROUPS * TC_GROUP_THREADS)
#define TC_TILE_STRIDE 127
#define TC_CHUNK 128
#define TC_TMEM_COLS 512
#define TC_A_BYTES 16384
#define TC_X_BYTES 2048
#define TC_GROUP_BYTES (TC_A_BYTES + TC_X_BYTES)

#define MODE_TRUECOLOR 0
#define MODE_INDEXED_16_8 7
#define SYMBOL_ERROR_MAX (0x7fffffff / 8)

/* floor (32768 / n) for n = 0..64 (n = 0 -> 0) */
__constant__ uint16_t c_tc_invdiv16 [65] =
{
    0, 32768, 16384, 10922, 8192, 6553, 5461, 4681, 4096, 3640, 3276, 2978, 2730, 2520, 2340, 2184,
    2048, 1927, 1820, 1724, 1638, 1560, 1489, 1424, 1365, 1310, 1260, 1213, 1170, 1129, 1092, 1057,
    1024, 992, 963, 936, 910, 885, 862, 840, 819, 799, 780, 762, 744, 728, 712, 697,
    682, 668, 655, 642, 630, 618, 606, 595, 585, 574, 564, 555, 546, 537, 528, 520, 512
};

/* ---- PTX wrappers ------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t
smem_u32 (const void *p)
{
    return (uint32_t) __cvta_generic_to_shared (p);
}

__device__ __forceinline__ void
mbar_init (uint32_t bar, uint32_t count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

/* Bounded wait: a protocol error must end in a trap, not in a hung device */
__device__ __forceinline__ void
mbar_wait (uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    for (;;)
    {
        uint32_t ok;
        asm volatile ("{\n"
                      ".reg .pred p;\n"
                      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
                      "selp.u32 %0, 1, 0, p;\n"
                      "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(2000u) : "memory");
        if (ok)
            return;
        if (++spins > 2000000u)
            __trap ();
    }
}

__device__ __forceinline__ void
group_barrier (int group)
{
    asm volatile ("bar.sync %0, %1;" :: "r"(group + 1), "n"(TC_GROUP_THREADS) : "memory");
}

__device__ __forceinline__ void tc_fence_before () { asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after () { asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld () { asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem () { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }

/* D (tensor memory) = or += A (shared) . B^T (shared), signed 8-bit operands, 32-bit accumulators */
__device__ __forceinline__ void
tc_mma_i8 (uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile ("{\n"
                  ".reg .pred p;\n"
                  "setp.ne.b32 p, %4, 0;\n"
                  "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
                  "}" :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void
tc_commit (uint32_t bar)
{
    asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

#include "tmem_ld.inc"

__device__ __forceinline__ void
cp_async_16 (void *smem, const void *gmem)
{
    asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32 (smem)), "l"(gmem) : "memory");
}

/* 8 bytes, or 8 zero bytes if !@valid (the source address must still be a legal one) */
__device__ __forceinline__ void
cp_async_8_zfill (void *smem, const void *gmem, bool valid)
{
    asm volatile ("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32 (smem)), "l"(gmem), "r"(valid ? 8 : 0)
                  : "memory");
}

/* Shared-memory operand descriptor, K-major, no swizzle: 8 rows x 16 bytes core matrices stored
 * as 128 contiguous bytes; SBO = distance between 8-row groups, LBO = distance between the two
 * 16-byte K slices of one MMA (K = 32 bytes). */
__device__ __forceinline__ uint64_t
smem_desc (uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t) ((addr >> 4) & 0x3fffu)
        | ((uint64_t) ((lbo_bytes >> 4) & 0x3fffu) << 16)
        | ((uint64_t) ((sbo_bytes >> 4) & 0x3fffu) << 32)
        | (1ull << 46);
}

/* kind::i8 instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, M = 128, N = 128 */
#define TC_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ uint32_t
dp4a_us (uint32_t pixels, uint32_t signs, uint32_t acc)
{
    uint32_t d;
    asm ("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pixels), "r"(signs), "r"(acc));
    return d;
}

/* ---- per-group state ------------------------------------------------------------------ */

struct TcGroup
{
    int group, t;                 /* group in block, thread in group */
    uint32_t a_addr;              /* shared: A operand / exchange area */
    uint8_t *a_ptr;
    uint32_t *xpair;              /* [2] [128] contrast colours, [2] [128] right-half bitmap */
    uint32_t bar;                 /* shared address of the group's mbarrier */
    uint32_t *arrivals;           /* shared: warps of the group that have emptied tensor memory (running count) */
    uint32_t parity;
    uint32_t tmem;                /* this group's first column, lane 0 */
    uint32_t tmem_ld;             /* the same for this warp's lanes */
};

/* ---- small helpers -------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t
tc_color_average_2 (uint32_t a, uint32_t b)
{
    return ((a >> 1) & 0x7f7f7f7fu) + ((b >> 1) & 0x7f7f7f7fu);
}

/* chafa-avx2.c:125-165: mean = mulhrs (sum, floor (32768 / n)) */
__device__ __forceinline__ uint32_t
tc_sums_to_color (const uint32_t (&s) [4], int n, const uint16_t *s_invdiv)
{
    uint32_t inv = s_invdiv [n];
    uint32_t c0 = (s [0] * inv + 16384u) >> 15;
    uint32_t c1 = (s [1] * inv + 16384u) >> 15;
    uint32_t c2 = (s [2] * inv + 16384u) >> 15;
    uint32_t c3 = (s [3] * inv + 16384u) >> 15;
    return c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
}

/* n |c|^2 - 2 c.S : the part of sum_i |c - px_i|^2 that depends on the colour */
__device__ __forceinline__ int
tc_color_term (uint32_t c, const uint32_t (&s) [4], int n)
{
    int c0 = (int) (c & 0xff), c1 = (int) ((c >> 8) & 0xff), c2 = (int) ((c >> 16) & 0xff), c3 = (int) (c >> 24);
    int dot = c0 * (int) s [0] + c1 * (int) s [1] + c2 * (int) s [2] + c3 * (int) s [3];
    int norm = c0 * c0 + c1 * c1 + c2 * c2 + c3 * c3;
    return n * norm - 2 * dot;
}

template <int KMAX>
__device__ __forceinline__ void
tc_insert (uint32_t (&top) [KMAX], uint32_t key)
{
#pragma unroll
    for (int i = 0; i < KMAX; i++)
    {
        uint32_t lo = min (top [i], key);
        key = max (top [i], key);
        top [i] = lo;
    }
}

/* 16 pixels (4 words of -1 / +1 bytes) of an outline bitmap; @rev has bit p = pixel p */
__device__ __forceinline__ uint4
tc_expand16 (uint32_t rev, int shift)
{
    uint32_t w [4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        uint32_t nib = (rev >> (shift + 4 * i)) & 15u;
        uint32_t s = (nib * 0x00204081u) & 0x01010101u;     /* bit b -> byte b */
        w [i] = s * 0xfeu + 0x01010101u;                    /* 1 -> 0xff (-1), 0 -> 0x01 */
    }
    return make_uint4 (w [0], w [1], w [2], w [3]);
}

/* Row @row of a 128-row operand tile: K slice c (16 bytes) lives at c * 2048 + row * 16 */
__device__ __forceinline__ void
tc_write_a_half (uint8_t *a, int row, int first_slice, uint64_t bm)
{
    uint32_t rev_hi = __brev ((uint32_t) (bm >> 32)), rev_lo = __brev ((uint32_t) bm);
    *(uint4 *) (a + (first_slice + 0) * 2048 + row * 16) = tc_expand16 (rev_hi, 0);
    *(uint4 *) (a + (first_slice + 1) * 2048 + row * 16) = tc_expand16 (rev_hi, 16);
    *(uint4 *) (a + (first_slice + 2) * 2048 + row * 16) = tc_expand16 (rev_lo, 0);
    *(uint4 *) (a + (first_slice + 3) * 2048 + row * 16) = tc_expand16 (rev_lo, 16);
}

/* ---- cell colours (chafa-symbol-renderer.c:656-729), one thread ---------------------------- */

/* Nearest pen of a fixed palette.  The pickers are pure functions of the 24-bit colour (for a
 * palette type and colour space), so they are tabulated once per device: 2^24 pens, 16 MB, built
 * by running the sequential picker (palette.cuh) on every colour.  A lookup is then one byte
 * load -- and exact by construction, quirks of the gray-ramp walk included. */
__global__ void
tc_build_pen_lut_kernel (const DevPalette *__restrict__ pal_src, int color_space, uint32_t *__restrict__ lut)
{
    __shared__ DevPalette pal;
    const uint32_t *src = (const uint32_t *) pal_src;
    uint32_t *dst = (uint32_t *) &pal;
    for (int i = threadIdx.x; i < (int) (sizeof (DevPalette) / 4); i += blockDim.x)
        dst [i] = src [i];
    __syncthreads ();
    if (threadIdx.x == 0)
    {
        pal.alpha_threshold = 0;
        pal.transparent_index = PAL_INDEX_TRANSPARENT;
    }
    __syncthreads ();

    uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    uint32_t packed = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < 4; i++)
        packed |= (uint32_t) palette_lookup_nearest (&pal, color_space, (first + i) | 0xff000000u, nullptr) << (8 * i);
    lut [first >> 2] = packed;
}

__device__ __noinline__ uint32_t
tc_palette_lookup_seq (const DevPalette *pal, int cs, uint32_t color)
{
    return (uint32_t) palette_lookup_nearest (pal, cs, color, nullptr);
}

struct TcColorCtx
{
    int mode, cs, alpha_threshold, fg_only;
    const DevPalette *fg_palette, *bg_palette;
    const uint8_t *lut;             /* NULL: sequential pickers */
    uint32_t solid_char;
};

__device__ __forceinline__ uint32_t
tc_pen (const TcColorCtx &x, const DevPalette *pal, uint32_t color)
{
    if (!x.lut)
        return tc_palette_lookup_seq (pal, x.cs, color);
    if ((int) (color >> 24) < x.alpha_threshold)
        return PAL_INDEX_TRANSPARENT;
    return __ldg (x.lut + (color & 0x00ffffffu));
}

__device__ __forceinline__ uint32_t
tc_transparent_cell_color (int mode)
{
    return mode == MODE_TRUECOLOR ? 0x00808080u : (uint32_t) PAL_INDEX_TRANSPARENT;
}

__device__ __forceinline__ void
tc_update_cell_colors (const TcColorCtx &x, uint32_t &c_inout, uint32_t col_fg, uint32_t col_bg,
                       uint32_t &fg_out, uint32_t &bg_out)
{
    if (x.mode == MODE_TRUECOLOR)
    {
        fg_out = pack_argb (col_fg);
        bg_out = pack_argb (col_bg);
    }
    else if (x.mode == MODE_INDEXED_16_8)
    {
        uint32_t fg = tc_pen (x, x.fg_palette, col_fg);
        uint32_t bg = tc_pen (x, x.fg_palette, col_bg);

        if (fg == bg && fg >= 8 && fg <= 15)
        {
            if (x.solid_char)
            {
                c_inout = x.solid_char;
                bg = tc_pen (x, x.bg_palette, col_fg);
            }
            else
            {
                fg = bg = tc_pen (x, x.bg_palette, col_fg);
            }
        }
        else
        {
            bg = tc_pen (x, x.bg_palette, col_bg);
        }
        fg_out = fg;
        bg_out = bg;
    }
    else
    {
        fg_out = tc_pen (x, x.fg_palette, col_fg);
        bg_out = tc_pen (x, x.bg_palette, col_bg);
    }

    if (x.fg_only)
        bg_out = tc_transparent_cell_color (x.mode);
}

/* ---- candidate search ------------------------------------------------------------------- */

/* hit |= bit_lo if the low half of @u >= the low half of @thr (unsigned 16-bit), |= bit_hi likewise for
 * the high halves: one VIMNMX.U16x2 with two predicate outputs and two predicated ORs */
__device__ __forceinline__ void
tc_flag_pair (uint32_t &hit, uint32_t u, uint32_t thr, uint32_t bit_lo, uint32_t bit_hi)
{
    asm ("{\n"
         ".reg .pred ph, pl;\n"
         ".reg .b32 mx;\n"
         ".reg .u16 a0, a1, b0, b1;\n"
         "max.u16x2 mx, %1, %2;\n"
         "mov.b32 {a0, a1}, mx;\n"
         "mov.b32 {b0, b1}, %1;\n"
         "setp.eq.u16 pl, a0, b0;\n"
         "setp.eq.u16 ph, a1, b1;\n"
         "@pl or.b32 %0, %0, %3;\n"
         "@ph or.b32 %0, %0, %4;\n"
         "}" : "+r"(hit) : "r"(u), "r"(thr), "r"(bit_lo), "r"(bit_hi));
}

/* One thread issues the MMAs of chunk @chunk (128 symbols): KSTEPS slices of K = 32 */
template <int KSTEPS>
__device__ __forceinline__ void
tc_issue_chunk (const TcGroup &g, uint32_t b_addr, uint32_t b_lbo, int chunk)
{
    uint64_t a_desc = smem_desc (g.a_addr, 2048, 128);
    uint64_t b_desc = smem_desc (b_addr + (uint32_t) chunk * (TC_CHUNK * 16), b_lbo, 128);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ks++)
    {
        tc_mma_i8 (g.tmem, a_desc, b_desc, TC_IDESC, ks > 0);
        a_desc += (2 * 2048) >> 4;
        b_desc += (2 * b_lbo) >> 4;
    }
    tc_commit (g.bar);
}

/* Distance of symbol @s to the thread's bitmap(s), as the key of its better orientation:
 * distance << 13 | sequence number (2 s plain, 2 s + 1 inverted; plain wins a tie) */
template <bool WIDE>
__device__ __forceinline__ uint32_t
tc_symbol_key (const uint64_t *s_bitmaps, int s, uint64_t bm_a, uint64_t bm_b)
{
    const int max_dist = WIDE ? 128 : 64;
    int hd;
    if (WIDE)
        hd = __popcll (s_bitmaps [2 * s] ^ bm_a) + __popcll (s_bitmaps [2 * s + 1] ^ bm_b);
    else
        hd = __popcll (s_bitmaps [s] ^ bm_a);
    int inv = max_dist - hd;
    return hd <= inv ? ((uint32_t) hd << 13) | (uint32_t) (2 * s) : ((uint32_t) inv << 13) | (uint32_t) (2 * s + 1);
}

/* The k smallest keys (distance << 13 | sequence number) of this thread's cell or cell pair,
 * ascending in top [0..k).  The A operand must be in place (written, fenced, group-synchronised). */
template <bool WIDE, int KMAX>
__device__ __forceinline__ void
tc_select (TcGroup &g, const uint64_t *s_bitmaps, int n_syms, uint32_t b_addr, uint32_t b_lbo,
           uint64_t bm_a, uint64_t bm_b, int k, uint32_t (&top) [KMAX])
{
    constexpr int KSTEPS = WIDE ? 4 : 2;
    const int max_dist = WIDE ? 128 : 64;
    const int n_chunks = (n_syms + TC_CHUNK - 1) / TC_CHUNK;
    const bool leader = g.t == 0;

    uint32_t mx [8], mn [8];
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        mx [i] = 0x80008000u;
        mn [i] = 0x7fff7fffu;
    }
#pragma unroll
    for (int i = 0; i < KMAX; i++)
        top [i] = 0xffffffffu;

    int bound = 0;                  /* T: candidates have |D| >= T */

    if (leader)
    {
        tc_fence_after ();
        tc_issue_chunk<KSTEPS> (g, b_addr, b_lbo, 0);
    }

    for (int q = 0; q < 2 * n_chunks; q++)
    {
        /* The 128 columns are read as two halves: the second load is in flight while the first half's
         * registers are processed, and only then does this warp report the buffer drained. */
        uint32_t ra [32], rb [32];

        mbar_wait (g.bar, g.parity);
        g.parity ^= 1u;
        tc_fence_after ();
        tmem_ld_64cols_pack16 (g.tmem_ld, ra);
        tmem_wait_ld_32 (ra);
        tmem_ld_64cols_pack16 (g.tmem_ld + 64, rb);

        const bool sweep1 = q < n_chunks;
        uint32_t hit [4] = { 0, 0, 0, 0 };      /* sweep 2: bit b of word w <-> symbol 32 w + b of the chunk */
        /* |D| >= T  <=>  (uint16) (D + T - 1) >= 2 T - 1 */
        const uint32_t add2 = (uint32_t) ((bound - 1) & 0xffff) * 0x00010001u;
        const uint32_t thr2 = (uint32_t) ((2 * bound - 1) & 0xffff) * 0x00010001u;

        auto half = [&] (uint32_t (&r) [32], int first_word)
        {
            if (sweep1)
            {
                /* 16 running maxima and minima (register position x half) */
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int i = 0; i < 8; i++)
                    {
                        mx [i] = __vimax3_s16x2 (mx [i], r [16 * h + i], r [16 * h + 8 + i]);
                        mn [i] = __vimin3_s16x2 (mn [i], r [16 * h + i], r [16 * h + 8 + i]);
                    }
            }
            else if (bound > 0)
            {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    tc_flag_pair (hit [first_word + (j >> 4)], __vadd2 (r [j], add2), thr2, 1u << (2 * (j & 15)),
                                  2u << (2 * (j & 15)));
            }
        };

        half (ra, 0);
        tmem_wait_ld_32 (rb);
        tc_fence_before ();
        /* The warp that empties tensor memory last issues the next chunk: nobody waits for anybody,
         * and the MMA runs under the processing of this chunk's registers. */
        if (q + 1 < 2 * n_chunks && (g.t & 31) == 0)
        {
            __threadfence_block ();
            uint32_t before = atomicAdd (g.arrivals, 1u);
            if ((before & 3u) == 3u)
            {
                __threadfence_block ();
                tc_fence_after ();
                tc_issue_chunk<KSTEPS> (g, b_addr, b_lbo, q + 1 < n_chunks ? q + 1 : q + 1 - n_chunks);
            }
        }
        half (rb, 2);

        if (sweep1)
        {
            if (q == n_chunks - 1)
            {
                /* bound = k-th largest partition maximum of |D| */
                int v [16];
#pragma unroll
                for (int i = 0; i < 8; i++)
                {
                    uint32_t a = __vmaxs2 (mx [i], __vnegss2 (mn [i]));
                    v [2 * i] = max ((int) (short) (a & 0xffffu), 0);
                    v [2 * i + 1] = max ((int) (short) (a >> 16), 0);
                }
                /* Batcher's odd-even merge sort for 16 inputs (63 compare-exchanges, depth 10; checked on all
                 * 2^16 0/1 inputs by tools/gen_sort_network.py): ascending, so the k-th largest is v [16 - k] */
                {
                    constexpr unsigned char net [63] [2] = {
                        {0,1}, {2,3}, {4,5}, {6,7}, {8,9}, {10,11}, {12,13}, {14,15}, {0,2}, {1,3}, {4,6}, {5,7},
                        {8,10}, {9,11}, {12,14}, {13,15}, {1,2}, {5,6}, {9,10}, {13,14}, {0,4}, {1,5}, {2,6}, {3,7},
                        {8,12}, {9,13}, {10,14}, {11,15}, {2,4}, {3,5}, {10,12}, {11,13}, {1,2}, {3,4}, {5,6}, {9,10},
                        {11,12}, {13,14}, {0,8}, {1,9}, {2,10}, {3,11}, {4,12}, {5,13}, {6,14}, {7,15}, {4,8}, {5,9},
                        {6,10}, {7,11}, {2,4}, {3,5}, {6,8}, {7,9}, {10,12}, {11,13}, {1,2}, {3,4}, {5,6}, {7,8},
                        {9,10}, {11,12}, {13,14} };
#pragma unroll
                    for (int c = 0; c < 63; c++)
                    {
                        int lo = min (v [net [c] [0]], v [net [c] [1]]), hi = max (v [net [c] [0]], v [net [c] [1]]);
                        v [net [c] [0]] = lo;
                        v [net [c] [1]] = hi;
                    }
                    bound = v [8];
#pragma unroll
                    for (int i = 9; i < 16; i++)
                        if (16 - i >= k)
                            bound = v [i];      /* ends at i = 16 - k */
                }
            }
        }
        else if (bound > 0)
        {
            /* rank the flagged symbols */
            const int sym_base = (q - n_chunks) * TC_CHUNK;
#pragma unroll
            for (int w = 0; w < 4; w++)
            {
                uint32_t m = hit [w];
                while (m)
                {
                    int b = __ffs (m) - 1;
                    m &= m - 1;
                    tc_insert<KMAX> (top, tc_symbol_key<WIDE> (s_bitmaps, sym_base + 32 * w + b, bm_a, bm_b));
                }
            }

            /* k candidates known: a later symbol has to be strictly better than the k-th */
            uint32_t kth = top [0];
#pragma unroll
            for (int i = 1; i < KMAX; i++)
                if (i < k)
                    kth = top [i];
            if (kth != 0xffffffffu)
                bound = max (bound, max_dist - 2 * (int) (kth >> 13) + 2);
        }
    }

    if (bound <= 0)
    {
        /* degenerate bound (tiny symbol tables): exact sequential scan */
        for (int s = 0; s < n_syms; s++)
        {
            int hd;
            if (WIDE)
                hd = __popcll (s_bitmaps [2 * s] ^ bm_a) + __popcll (s_bitmaps [2 * s + 1] ^ bm_b);
            else
                hd = __popcll (s_bitmaps [s] ^ bm_a);
            tc_insert<KMAX> (top, ((uint32_t) hd << 13) | (uint32_t) (2 * s));
            tc_insert<KMAX> (top, ((uint32_t) (max_dist - hd) << 13) | (uint32_t) (2 * s + 1));
        }
    }
}

/* ---- masked sums ---------------------------------------------------------------------- */

/* sum over the covered pixels of each channel: (total - sum_i px_i sign_i) / 2 with sign = -1 on
 * covered pixels.  @row = shared address of the symbol's 16 bytes in K slice @first_slice; the
 * next slices follow at @slice_stride bytes. */
__device__ __forceinline__ void
tc_masked_sums (const uint32_t (&px) [4] [16], const uint32_t (&tot) [4], const uint8_t *row, uint32_t slice_stride,
                uint32_t (&s_fg) [4])
{
    uint32_t d [4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        uint4 m = *(const uint4 *) (row + c * slice_stride);
        uint32_t w [4] = { m.x, m.y, m.z, m.w };
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
                d [ch] = dp4a_us (px [ch] [4 * c + i], w [i], d [ch]);
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
        s_fg [ch] = (tot [ch] - d [ch]) >> 1;
}

/* ---- the kernel ----------------------------------------------------------------------- */

/* KEXACT: the number of candidates is KMAX for narrow and wide symbols alike (the usual case: both
 * tables hold more than KMAX / 2 symbols), which makes the candidate loops compile-time bounded */
template <int KMAX, bool KEXACT>
__global__ void __launch_bounds__ (TC_THREADS, 1)
match_tc_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevCellEval *__restrict__ evals,
                 DevWideEval *__restrict__ wide_evals, int n_tiles)
{
    extern __shared__ __align__ (128) uint8_t s_mem [];
    __shared__ __align__ (8) unsigned long long s_bars [TC_GROUPS];
    __shared__ uint32_t s_tmem_base;
    __shared__ uint32_t s_arrivals [TC_GROUPS];
    __shared__ uint16_t s_invdiv [65];

    const int n_syms = cv.syms.n;
    const int n_syms2 = wide_evals ? cv.syms.n2 : 0;
    const int npad = cv.syms.npad, n2pad = wide_evals ? cv.syms.n2pad : 0;
    uint8_t *s_bn = s_mem;
    uint8_t *s_bw = s_bn + (size_t) npad * 64;
    uint8_t *s_groups = s_bw + (size_t) n2pad * 128;
    uint64_t *s_bitmaps = (uint64_t *) (s_groups + (size_t) TC_GROUPS * TC_GROUP_BYTES);      /* npad */
    uint64_t *s_bitmaps2 = s_bitmaps + npad;                                                  /* 2 * n2pad */

    const int tid = threadIdx.x, warp = tid >> 5;

    TcGroup g;
    g.group = tid >> 7;
    g.t = tid & 127;
    g.a_ptr = s_groups + (size_t) g.group * TC_GROUP_BYTES;
    g.a_addr = smem_u32 (g.a_ptr);
    g.xpair = (uint32_t *) (g.a_ptr + TC_A_BYTES);
    g.bar = smem_u32 (&s_bars [g.group]);
    g.arrivals = &s_arrivals [g.group];
    g.parity = 0;

    /* B operands (the symbol tables as -1 / +1 bytes, already in operand layout in global memory) and
     * the bitmap tables: asynchronous copies that land while the first tile's bitmaps are computed */
    {
        const uint8_t *src = cv.syms.bimage;
        for (int i = tid; i < npad * 4; i += TC_THREADS)
            cp_async_16 (s_bn + (size_t) i * 16, src + (size_t) i * 16);
        src = cv.syms.bimage2;
        if (n_syms2 > 0)
            for (int i = tid; i < n2pad * 8; i += TC_THREADS)
                cp_async_16 (s_bw + (size_t) i * 16, src + (size_t) i * 16);
        for (int i = tid; i < npad; i += TC_THREADS)
            cp_async_8_zfill (s_bitmaps + i, cv.syms.bitmaps + min (i, n_syms - 1), i < n_syms);
        for (int i = tid; i < 2 * n2pad; i += TC_THREADS)
            cp_async_8_zfill (s_bitmaps2 + i, cv.syms.bitmaps2 + min (i, max (2 * n_syms2 - 1, 0)), i < 2 * n_syms2);
        asm volatile ("cp.async.commit_group;" ::: "memory");
        if (tid < 65)
            s_invdiv [tid] = c_tc_invdiv16 [tid];
    }
    bool tables_ready = false;
    if (tid < TC_GROUPS)
    {
        mbar_init (smem_u32 (&s_bars [tid]), 1);
        s_arrivals [tid] = 0;
    }
    if (warp == 0)
    {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                      :: "r"(smem_u32 (&s_tmem_base)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem ();
    tc_fence_before ();
    __syncthreads ();
    tc_fence_after ();

    g.tmem = s_tmem_base + (uint32_t) g.group * TC_CHUNK;
    g.tmem_ld = g.tmem + ((uint32_t) ((warp & 3) * 32) << 16);

    TcColorCtx cx_colors;
    cx_colors.mode = cv.canvas_mode;
    cx_colors.cs = cv.color_space;
    cx_colors.alpha_threshold = cv.alpha_threshold;
    cx_colors.fg_only = cv.fg_only;
    cx_colors.fg_palette = cv.fg_palette;
    cx_colors.bg_palette = cv.bg_palette;
    cx_colors.lut = cv.pen_lut;
    cx_colors.solid_char = cv.solid_char;

    const int n_cells = cv.width * cv.n_rows;
    const int k_narrow = KEXACT ? KMAX : min (cv.n_candidates, 2 * n_syms);
    const int k_wide = KEXACT ? KMAX : min (cv.n_candidates, 2 * n_syms2);

    for (int tile = blockIdx.x + gridDim.x * g.group; tile < n_tiles; tile += gridDim.x * TC_GROUPS)
    {
        const int cell_raw = tile * TC_TILE_STRIDE + g.t;
        const bool cell_valid = cell_raw < n_cells;
        const int cell = cell_valid ? cell_raw : n_cells - 1;
        const int cx = cell % cv.width, cy = cv.first_row + cell / cv.width;
        const uint4 *cell_px = (const uint4 *) (pixels + (size_t) cy * 8 * cv.width_px + cx * 8);
        const int row_q = cv.width_px >> 2;         /* row pitch in uint4 (width_px is a multiple of 8) */

        /* does this thread own the pair (cell, cell + 1)? */
        const bool pair_valid = n_syms2 > 0 && cell_valid && g.t < TC_TILE_STRIDE && cell_raw + 1 < n_cells
            && cx != cv.width - 1;

        uint64_t bm_n, bm_l, bm_r;      /* outline bitmaps: own colours, pair with the right / left neighbour */

        /* ---- A: contrasting colours and outline bitmaps ---- */
        {
            uint32_t px [64];
#pragma unroll
            for (int y = 0; y < 8; y++)
            {
                uint4 a = __ldg (cell_px + (size_t) y * row_q), b = __ldg (cell_px + (size_t) y * row_q + 1);
                px [8 * y + 0] = a.x; px [8 * y + 1] = a.y; px [8 * y + 2] = a.z; px [8 * y + 3] = a.w;
                px [8 * y + 4] = b.x; px [8 * y + 5] = b.y; px [8 * y + 6] = b.z; px [8 * y + 7] = b.w;
            }

            /* chafa-work-cell.c:293-338: first minimum / last maximum along the widest channel */
            /* key = value << 6 | pixel index (14 bits); pixels i and i + 32 side by side in one
             * register, 16-bit SIMD min / max */
            uint32_t kmin [4], kmax [4];
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
            {
                uint32_t mn2 = 0xffffffffu, mx2 = 0;
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                {
                    uint32_t ka = (__byte_perm (px [i], px [i + 32], 0x4440 + ch + (ch << 8)) & 0x00ff00ffu) * 64u
                        + ((uint32_t) i | ((uint32_t) (i + 32) << 16));
                    uint32_t kb = (__byte_perm (px [i + 1], px [i + 33], 0x4440 + ch + (ch << 8)) & 0x00ff00ffu) * 64u
                        + ((uint32_t) (i + 1) | ((uint32_t) (i + 33) << 16));
                    mn2 = __vimin3_u16x2 (mn2, ka, kb);
                    mx2 = __vimax3_u16x2 (mx2, ka, kb);
                }
                kmin [ch] = min (mn2 & 0xffffu, mn2 >> 16);
                kmax [ch] = max (mx2 & 0xffffu, mx2 >> 16);
            }
            int best_ch = 0;
            int best_range = (int) (kmax [0] >> 6) - (int) (kmin [0] >> 6);
#pragma unroll
            for (int ch = 1; ch < 4; ch++)
            {
                int range = (int) (kmax [ch] >> 6) - (int) (kmin [ch] >> 6);
                if (range > best_range)
                {
                    best_range = range;
                    best_ch = ch;
                }
            }
            uint32_t imin = kmin [0] & 63u, imax = kmax [0] & 63u;
            if (best_ch == 1) { imin = kmin [1] & 63u; imax = kmax [1] & 63u; }
            else if (best_ch == 2) { imin = kmin [2] & 63u; imax = kmax [2] & 63u; }
            else if (best_ch == 3) { imin = kmin [3] & 63u; imax = kmax [3] & 63u; }

            const uint32_t *cell_px32 = (const uint32_t *) cell_px;
            uint32_t own_bg = __ldg (cell_px32 + (size_t) (imin >> 3) * cv.width_px + (imin & 7u));
            uint32_t own_fg = __ldg (cell_px32 + (size_t) (imax >> 3) * cv.width_px + (imax & 7u));

            g.xpair [g.t] = own_bg;
            g.xpair [128 + g.t] = own_fg;
            group_barrier (g.group);
            uint32_t lbg = own_bg, lfg = own_fg, rbg = own_bg, rfg = own_fg;
            if (g.t > 0) { lbg = g.xpair [g.t - 1]; lfg = g.xpair [128 + g.t - 1]; }
            if (g.t < 127) { rbg = g.xpair [g.t + 1]; rfg = g.xpair [128 + g.t + 1]; }

            /* colour pairs: own; pair (this, right) ; pair (left, this) */
            uint32_t bgc [3], fgc [3];
            int negc [3];
            bgc [0] = own_bg; fgc [0] = own_fg;
            bgc [1] = tc_color_average_2 (own_bg, rbg); fgc [1] = tc_color_average_2 (own_fg, rfg);
            bgc [2] = tc_color_average_2 (lbg, own_bg); fgc [2] = tc_color_average_2 (lfg, own_fg);
#pragma unroll
            for (int v = 0; v < 3; v++)
            {
                bgc [v] &= 0x00ffffffu;
                fgc [v] &= 0x00ffffffu;
                /* diff3 (px, bg) > diff3 (px, fg)  <=>  2 px.(fg - bg) + |bg|^2 - |fg|^2 > 0 */
                negc [v] = (int) __dp4a (fgc [v], fgc [v], 0u) - (int) __dp4a (bgc [v], bgc [v], 0u);
            }

            uint32_t bits [3] [2];
#pragma unroll
            for (int v = 0; v < 3; v++)
#pragma unroll
                for (int half = 0; half < 2; half++)
                {
                    uint32_t acc = 0;
#pragma unroll
                    for (int i = 0; i < 32; i++)
                    {
                        uint32_t p = px [32 * half + i];
                        int d = (int) __dp4a (p, bgc [v], 0u) - (int) __dp4a (p, fgc [v], 0u);
                        int w = 2 * d + negc [v];            /* negative <=> bit set */
                        acc = __funnelshift_l ((uint32_t) w, acc, 1);
                    }
                    bits [v] [half] = acc;
                }
            bm_n = ((uint64_t) bits [0] [0] << 32) | bits [0] [1];
            bm_l = ((uint64_t) bits [1] [0] << 32) | bits [1] [1];
            bm_r = ((uint64_t) bits [2] [0] << 32) | bits [2] [1];
        }

        uint32_t top_n [KMAX], top_w [KMAX];
#pragma unroll
        for (int i = 0; i < KMAX; i++)
            top_n [i] = top_w [i] = 0xffffffffu;

        if (!tables_ready)
        {
            asm volatile ("cp.async.wait_all;" ::: "memory");
            fence_async_smem ();
            __syncthreads ();
            tables_ready = true;
        }

        /* ---- B: narrow candidates ---- */
        if (n_syms > 0)
        {
            tc_write_a_half (g.a_ptr, g.t, 0, bm_n);
            fence_async_smem ();
            group_barrier (g.group);
            tc_select<false, KMAX> (g, s_bitmaps, n_syms, smem_u32 (s_bn), (uint32_t) npad * 16, bm_n, 0, k_narrow, top_n);
        }

        /* ---- C: wide candidates for the pair (this cell, right neighbour) ---- */
        uint64_t bm_pair_b = 0;
        if (n_syms2 > 0)
        {
            /* publish the right-half bitmap to the left neighbour (row t - 1, K slices 4..7) */
            uint64_t *xbm = (uint64_t *) (g.xpair + 256);
            xbm [g.t] = bm_r;
            tc_write_a_half (g.a_ptr, g.t, 0, bm_l);
            if (g.t > 0)
                tc_write_a_half (g.a_ptr, g.t - 1, 4, bm_r);
            else
                tc_write_a_half (g.a_ptr, 127, 4, 0);
            fence_async_smem ();
            group_barrier (g.group);
            bm_pair_b = g.t < 127 ? xbm [g.t + 1] : 0;
            tc_select<true, KMAX> (g, s_bitmaps2, n_syms2, smem_u32 (s_bw), (uint32_t) n2pad * 16, bm_l, bm_pair_b, k_wide,
                             top_w);
        }

        /* ---- D: scoring.  The A area is free now: exchange area for the pair's right half ---- */
        uint16_t *xcand = (uint16_t *) g.a_ptr;                       /* [8] [128] */
        uint2 *xsum = (uint2 *) (g.a_ptr + 2048);                     /* [8] [128] */
        uint32_t *xtot = (uint32_t *) (g.a_ptr + 2048 + 8192);        /* [5] [128] */

        if (n_syms2 > 0)
        {
            group_barrier (g.group);                                  /* everyone is done with the operand */
#pragma unroll
            for (int i = 0; i < KMAX; i++)
                xcand [i * 128 + g.t] = (uint16_t) ((top_w [i] & 0x1fffu) >> 1);
        }

        uint32_t pl [4] [16];
        uint32_t tot [4] = { 0, 0, 0, 0 };
        uint32_t tot_q = 0;
#pragma unroll
        for (int y = 0; y < 8; y++)
#pragma unroll
            for (int half = 0; half < 2; half++)
            {
                uint4 a = __ldg (cell_px + (size_t) y * row_q + half);
                uint32_t t01 = __byte_perm (a.x, a.y, 0x5140), u01 = __byte_perm (a.x, a.y, 0x7362);
                uint32_t t23 = __byte_perm (a.z, a.w, 0x5140), u23 = __byte_perm (a.z, a.w, 0x7362);
                int gi = 2 * y + half;
                pl [0] [gi] = __byte_perm (t01, t23, 0x5410);
                pl [1] [gi] = __byte_perm (t01, t23, 0x7632);
                pl [2] [gi] = __byte_perm (u01, u23, 0x5410);
                pl [3] [gi] = __byte_perm (u01, u23, 0x7632);
#pragma unroll
                for (int ch = 0; ch < 4; ch++)
                {
                    tot [ch] = __dp4a (pl [ch] [gi], 0x01010101u, tot [ch]);
                    tot_q = __dp4a (pl [ch] [gi], pl [ch] [gi], tot_q);
                }
            }

        if (n_syms2 > 0)
        {
            xtot [0 * 128 + g.t] = tot [0];
            xtot [1 * 128 + g.t] = tot [1];
            xtot [2 * 128 + g.t] = tot [2];
            xtot [3 * 128 + g.t] = tot [3];
            xtot [4 * 128 + g.t] = tot_q;
            group_barrier (g.group);

            /* right half of the left neighbour's candidates, on this thread's pixels */
            if (g.t > 0)
            {
#pragma unroll 1
                for (int i = 0; i < k_wide; i++)
                {
                    int sym = xcand [i * 128 + g.t - 1];
                    uint32_t s_fg [4] = { 0, 0, 0, 0 };
                    if (sym < n_syms2)
                        tc_masked_sums (pl, tot, s_bw + ((size_t) 4 * n2pad + sym) * 16, (uint32_t) n2pad * 16, s_fg);
                    xsum [i * 128 + g.t - 1] = make_uint2 (s_fg [0] | (s_fg [1] << 16), s_fg [2] | (s_fg [3] << 16));
                }
            }
            group_barrier (g.group);
        }

        /* narrow: first smallest error wins (chafa-symbol-renderer.c:341-397) */
        {
            DevCellEval out;
            out.c = ' ';
            out.fg = out.bg = 0;
            out.error = SYMBOL_ERROR_MAX;
            out.mean = tc_sums_to_color (tot, 64, s_invdiv);

            uint32_t best_key = 0xffffffffu, best_fg = 0, best_bg = 0;
            int best_sym = -1;
#pragma unroll
            for (int i = 0; i < KMAX; i++)
            {
                if (i < k_narrow && top_n [i] != 0xffffffffu)
                {
                    int sym = (int) ((top_n [i] & 0x1fffu) >> 1);
                    int n_fg = __popcll (s_bitmaps [sym]);
                    uint32_t s_fg [4], s_bg [4];
                    tc_masked_sums (pl, tot, s_bn + (size_t) sym * 16, (uint32_t) npad * 16, s_fg);
#pragma unroll
                    for (int ch = 0; ch < 4; ch++)
                        s_bg [ch] = tot [ch] - s_fg [ch];
                    uint32_t fg = tc_sums_to_color (s_fg, n_fg, s_invdiv);
                    uint32_t bg = tc_sums_to_color (s_bg, 64 - n_fg, s_invdiv);
                    int error = (int) tot_q + tc_color_term (fg, s_fg, n_fg) + tc_color_term (bg, s_bg, 64 - n_fg);
                    uint32_t key = ((uint32_t) error << 3) | (uint32_t) i;
                    if (error < SYMBOL_ERROR_MAX && key < best_key)
                    {
                        best_key = key;
                        best_sym = sym;
                        best_fg = fg;
                        best_bg = bg;
                    }
                }
            }
            if (best_sym >= 0)
            {
                out.error = (int) (best_key >> 3);
                out.c = __ldg (cv.syms.chars + best_sym);
                tc_update_cell_colors (cx_colors, out.c, best_fg, best_bg, out.fg, out.bg);
            }
            if (cell_valid && g.t < TC_TILE_STRIDE)
                evals [cell] = out;
        }

        /* wide: the pair (this cell, right neighbour); stored at the right cell */
        if (n_syms2 > 0)
        {
            DevWideEval out;
            out.c = 0;
            out.fg = out.bg = 0;
            out.error_a = out.error_b = SYMBOL_ERROR_MAX;

            if (g.t < 127)
            {
                uint32_t rtot [4], rtot_q;
                rtot [0] = xtot [0 * 128 + g.t + 1];
                rtot [1] = xtot [1 * 128 + g.t + 1];
                rtot [2] = xtot [2 * 128 + g.t + 1];
                rtot [3] = xtot [3 * 128 + g.t + 1];
                rtot_q = xtot [4 * 128 + g.t + 1];

                uint32_t best_key = 0xffffffffu, best_fg = 0, best_bg = 0;
                int best_sym = -1, best_ea = SYMBOL_ERROR_MAX, best_eb = SYMBOL_ERROR_MAX;
#pragma unroll
                for (int i = 0; i < KMAX; i++)
                {
                    if (i < k_wide && top_w [i] != 0xffffffffu)
                    {
                        int sym = (int) ((top_w [i] & 0x1fffu) >> 1);
                        int na = __popcll (s_bitmaps2 [2 * sym]);
                        int nb = __popcll (s_bitmaps2 [2 * sym + 1]);
                        uint32_t a_fg [4], a_bg [4], b_fg [4], b_bg [4];
                        tc_masked_sums (pl, tot, s_bw + (size_t) sym * 16, (uint32_t) n2pad * 16, a_fg);
                        uint2 rs = xsum [i * 128 + g.t];
                        b_fg [0] = rs.x & 0xffffu; b_fg [1] = rs.x >> 16; b_fg [2] = rs.y & 0xffffu; b_fg [3] = rs.y >> 16;
#pragma unroll
                        for (int ch = 0; ch < 4; ch++)
                        {
                            a_bg [ch] = tot [ch] - a_fg [ch];
                            b_bg [ch] = rtot [ch] - b_fg [ch];
                        }
                        uint32_t fg = tc_color_average_2 (tc_sums_to_color (a_fg, na, s_invdiv),
                                                          tc_sums_to_color (b_fg, nb, s_invdiv));
                        uint32_t bg = tc_color_average_2 (tc_sums_to_color (a_bg, 64 - na, s_invdiv),
                                                          tc_sums_to_color (b_bg, 64 - nb, s_invdiv));
                        int ea = (int) tot_q + tc_color_term (fg, a_fg, na) + tc_color_term (bg, a_bg, 64 - na);
                        int eb = (int) rtot_q + tc_color_term (fg, b_fg, nb) + tc_color_term (bg, b_bg, 64 - nb);
                        uint32_t key = ((uint32_t) (ea + eb) << 3) | (uint32_t) i;
                        if (ea + eb < 2 * SYMBOL_ERROR_MAX && key < best_key)
                        {
                            best_key = key;
                            best_sym = sym;
                            best_fg = fg;
                            best_bg = bg;
                            best_ea = ea;
                            best_eb = eb;
                        }
                    }
                }
                if (best_sym >= 0)
                {
                    out.error_a = best_ea;
                    out.error_b = best_eb;
                    out.c = __ldg (cv.syms.chars2 + best_sym);
                    tc_update_cell_colors (cx_colors, out.c, best_fg, best_bg, out.fg, out.bg);
                }
            }
            if (pair_valid)
                wide_evals [cell + 1] = out;
        }

        /* the exchange area becomes the next tile's operand */
        group_barrier (g.group);
    }

    if (!tables_ready)
    {
        asm volatile ("cp.async.wait_all;" ::: "memory");
        __syncthreads ();
    }
    tc_fence_before ();
    __syncthreads ();
    if (warp == 0)
    {
        tc_fence_after ();
        asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "n"(TC_TMEM_COLS) : "memory");
    }
}

/* ---- K3xt: exhaustive matching (work factor >= 0.8) on the tensor cores ------------------------
 *
 * Every symbol is scored for every cell (chafa-symbol-renderer.c:270-339).  What a score needs from
 * the pixels is the masked channel sums S_fg [ch] = sum over the covered pixels of channel ch, and
 * that is a matrix product too: (cells x 64 pixel values of one channel, u8) . (symbols x 64
 * coverage bytes, 0/1)^T.  Four MMAs per step (one per channel plane of the tile) fill 4 x 32
 * tensor-memory columns with the sums of 32 symbols; the thread that owns the cell reads them as
 * packed 16-bit pairs and evaluates the closed-form error (see K3x in match.cu) -- about 55
 * instructions per (cell, symbol) where the table-driven kernel spends about 250.
 * Wide symbols: the pair's right cell is the next row of the same operand, i.e. the same
 * shared-memory image read 16 bytes further on (rows are 16 bytes apart in every K slice), so the
 * left-half and right-half sums of a pair land in the lane of its left cell. */

#define XT_GROUPS 3
#define XT_THREADS (XT_GROUPS * TC_GROUP_THREADS)
#define XT_PLANE_BYTES 8192
#define XT_A_BYTES (4 * XT_PLANE_BYTES + 128)
#define XT_X_BYTES 2560
#define XT_GROUP_BYTES (XT_A_BYTES + XT_X_BYTES)
#define XT_NCHUNK 32
#define XT_WCHUNK 16
/* kind::i8, unsigned x unsigned -> s32, K-major, M = 128 */
#define XT_IDESC(n) ((2u << 4) | (((uint32_t) (n) >> 3) << 17) | ((128u >> 4) << 24))

/* (sum * inv + 16384) >> 15 as one wide multiply-add: @sum2 = 2 sum, @inv16 = inv << 16 */
__device__ __forceinline__ uint32_t
xt_mean (uint32_t sum2, uint32_t inv16)
{
    return (uint32_t) (((unsigned long long) sum2 * inv16 + 0x80000000ull) >> 32);
}

/* n c^2 - 2 c S for one channel; @sum2 = 2 S */
__device__ __forceinline__ int
xt_term (int c, int sum2, int n)
{
    return c * (n * c - sum2);
}

template <int DUMMY>
__global__ void __launch_bounds__ (XT_THREADS, 1)
match_xtc_kernel (DevCanvas cv, const uint32_t *__restrict__ pixels, DevCellEval *__restrict__ evals,
                  DevWideEval *__restrict__ wide_evals, int n_tiles, int npad, int n2pad)
{
    extern __shared__ __align__ (128) uint8_t s_mem [];
    __shared__ __align__ (8) unsigned long long s_bars [XT_GROUPS];
    __shared__ uint32_t s_tmem_base;
    __shared__ uint32_t s_arrivals [XT_GROUPS];
    __shared__ uint16_t s_invdiv [65];
    __shared__ int s_alpha_const, s_alpha_uniform;

    const int n_syms = cv.syms.n;
    const int n_syms2 = wide_evals ? cv.syms.n2 : 0;
    if (!wide_evals)
        n2pad = 0;
    uint8_t *s_bn = s_mem;                                            /* npad x 64, bytes 0 / 1 */
    uint8_t *s_bw = s_bn + (size_t) npad * 64;                        /* n2pad x 128 */
    uint2 *s_symc = (uint2 *) (s_bw + (size_t) n2pad * 128);          /* npad: inv (n) << 16 | n, inv (64 - n) << 16 */
    uint4 *s_symc2 = (uint4 *) (s_symc + npad);                       /* n2pad: the same for both halves */
    uint8_t *s_groups = (uint8_t *) (s_symc2 + n2pad);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int group = tid >> 7, t = tid & 127;
    uint8_t *a_ptr = s_groups + (size_t) group * XT_GROUP_BYTES;
    const uint32_t a_addr = smem_u32 (a_ptr);
    uint32_t *xtot = (uint32_t *) (a_ptr + XT_A_BYTES);               /* [5] [128] */
    const uint32_t bar = smem_u32 (&s_bars [group]);
    uint32_t parity = 0;

    if (tid < 65)
        s_invdiv [tid] = c_tc_invdiv16 [tid];
    __syncthreads ();

    /* A fully opaque cell's alpha channel contributes the same two terms to every symbol's error:
     * both means are 255 whatever the coverage n (checked here for every n with the arithmetic of the
     * sweep), so warps whose 32 cells are opaque sweep three channels and add the constant. */
    if (tid == 0)
    {
        int first = 0, uniform = 1;
        for (int n = 0; n <= 64; n++)
        {
            const uint32_t sf2 = 510u * (uint32_t) n, sb2 = 510u * (uint32_t) (64 - n);
            const int v = xt_term ((int) xt_mean (sf2, (uint32_t) s_invdiv [n] << 16), (int) sf2, n)
                + xt_term ((int) xt_mean (sb2, (uint32_t) s_invdiv [64 - n] << 16), (int) sb2, 64 - n);
            if (n == 0)
                first = v;
            else if (v != first)
                uniform = 0;
        }
        s_alpha_const = first;
        s_alpha_uniform = uniform;
    }

    /* coverage bytes from the -1 / +1 operand images: 0xff -> 1, 0x01 -> 0; the images are laid out for
     * cv.syms.npad / n2pad rows per K slice, here for npad / n2pad */
    for (int i = tid; i < npad * 4; i += XT_THREADS)
    {
        int c = i / npad, r = i % npad;
        uint4 v = make_uint4 (0, 0, 0, 0);
        if (r < n_syms)
        {
            v = __ldg ((const uint4 *) cv.syms.bimage + (size_t) c * cv.syms.npad + r);
            v.x = (v.x >> 7) & 0x01010101u; v.y = (v.y >> 7) & 0x01010101u;
            v.z = (v.z >> 7) & 0x01010101u; v.w = (v.w >> 7) & 0x01010101u;
        }
        ((uint4 *) s_bn) [i] = v;
    }
    for (int i = tid; i < n2pad * 8; i += XT_THREADS)
    {
        int c = i / n2pad, r = i % n2pad;
        uint4 v = make_uint4 (0, 0, 0, 0);
        if (r < n_syms2)
        {
            v = __ldg ((const uint4 *) cv.syms.bimage2 + (size_t) c * cv.syms.n2pad + r);
            v.x = (v.x >> 7) & 0x01010101u; v.y = (v.y >> 7) & 0x01010101u;
            v.z = (v.z >> 7) & 0x01010101u; v.w = (v.w >> 7) & 0x01010101u;
        }
        ((uint4 *) s_bw) [i] = v;
    }
    for (int i = tid; i < npad; i += XT_THREADS)
    {
        /* Padding rows score exactly the cell's sum of squares (zero sums, zero reciprocals, so both colours
         * come out 0).  No real narrow symbol exceeds that: with c = the rounded mean of S over n, the term
         * c (n c - 2 S) is <= 0 (n c <= S + n / 2 <= 2 S when S >= n / 2, and c = 0 otherwise).  Padding comes
         * last and ties lose, so it can never win and the narrow sweep needs no bounds check. */
        int n = i < n_syms ? __popcll (__ldg (cv.syms.bitmaps + i)) : 0;
        s_symc [i] = i < n_syms
            ? make_uint2 ((uint32_t) n | ((uint32_t) s_invdiv [n] << 16), (uint32_t) s_invdiv [64 - n] << 16)
            : make_uint2 (0u, 0u);
    }
    for (int i = tid; i < n2pad; i += XT_THREADS)
    {
        int na = i < n_syms2 ? __popcll (__ldg (cv.syms.bitmaps2 + 2 * i)) : 0;
        int nb = i < n_syms2 ? __popcll (__ldg (cv.syms.bitmaps2 + 2 * i + 1)) : 0;
        s_symc2 [i] = i < n_syms2
            ? make_uint4 ((uint32_t) na | ((uint32_t) s_invdiv [na] << 16), (uint32_t) s_invdiv [64 - na] << 16,
                          (uint32_t) nb | ((uint32_t) s_invdiv [nb] << 16), (uint32_t) s_invdiv [64 - nb] << 16)
            : make_uint4 (0u, 0u, 0u, 0u);
    }
    if (tid < XT_GROUPS)
    {
        mbar_init (smem_u32 (&s_bars [tid]), 1);
        s_arrivals [tid] = 0;
    }
    if (warp == 0)
    {
        asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                      :: "r"(smem_u32 (&s_tmem_base)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem ();
    tc_fence_before ();
    __syncthreads ();
    tc_fence_after ();

    const uint32_t tmem = s_tmem_base + (uint32_t) group * 128;
    const uint32_t tmem_ld = tmem + ((uint32_t) ((warp & 3) * 32) << 16);
    const uint32_t bn_addr = smem_u32 (s_bn), bw_addr = smem_u32 (s_bw);

    TcColorCtx cx_colors;
    cx_colors.mode = cv.canvas_mode;
    cx_colors.cs = cv.color_space;
    cx_colors.alpha_threshold = cv.alpha_threshold;
    cx_colors.fg_only = cv.fg_only;
    cx_colors.fg_palette = cv.fg_palette;
    cx_colors.bg_palette = cv.bg_palette;
    cx_colors.lut = cv.pen_lut;
    cx_colors.solid_char = cv.solid_char;

    const int n_cells = cv.width * cv.n_rows;
    const int n_chunks = npad / XT_NCHUNK, n_chunks2 = n2pad / XT_WCHUNK;

    /* one thread issues the MMAs of a step; @step < n_chunks: narrow, else wide */
    auto issue = [&] (int step)
    {
        if (step < n_chunks)
        {
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
            {
                uint64_t a_desc = smem_desc (a_addr + ch * XT_PLANE_BYTES, 2048, 128);
                uint64_t b_desc = smem_desc (bn_addr + (uint32_t) step * (XT_NCHUNK * 16), (uint32_t) npad * 16, 128);
#pragma unroll
                for (int ks = 0; ks < 2; ks++)
                {
                    tc_mma_i8 (tmem + ch * XT_NCHUNK, a_desc, b_desc, XT_IDESC (XT_NCHUNK), ks > 0);
                    a_desc += (2 * 2048) >> 4;
                    b_desc += (2 * (uint32_t) npad * 16) >> 4;
                }
            }
        }
        else
        {
            int w = step - n_chunks;
#pragma unroll
            for (int side = 0; side < 2; side++)
#pragma unroll
                for (int ch = 0; ch < 4; ch++)
                {
                    /* side 1: the right cell of the pair = the next row of the plane */
                    uint64_t a_desc = smem_desc (a_addr + ch * XT_PLANE_BYTES + side * 16, 2048, 128);
                    uint64_t b_desc = smem_desc (bw_addr + (uint32_t) (side * 4 * n2pad + w * XT_WCHUNK) * 16,
                                                 (uint32_t) n2pad * 16, 128);
#pragma unroll
                    for (int ks = 0; ks < 2; ks++)
                    {
                        tc_mma_i8 (tmem + (side * 4 + ch) * XT_WCHUNK, a_desc, b_desc, XT_IDESC (XT_WCHUNK), ks > 0);
                        a_desc += (2 * 2048) >> 4;
                        b_desc += (2 * (uint32_t) n2pad * 16) >> 4;
                    }
                }
        }
        tc_commit (bar);
    };

    for (int tile = blockIdx.x + gridDim.x * group; tile < n_tiles; tile += gridDim.x * XT_GROUPS)
    {
        const int cell_raw = tile * TC_TILE_STRIDE + t;
        const bool cell_valid = cell_raw < n_cells;
        const int cell = cell_valid ? cell_raw : n_cells - 1;
        const int cx = cell % cv.width, cy = cv.first_row + cell / cv.width;
        const uint4 *cell_px = (const uint4 *) (pixels + (size_t) cy * 8 * cv.width_px + cx * 8);
        const int row_q = cv.width_px >> 2;
        const bool pair_valid = n_syms2 > 0 && cell_valid && t < TC_TILE_STRIDE && cell_raw + 1 < n_cells
            && cx != cv.width - 1;

        /* ---- operand: the tile's four channel planes; whole-cell sums ---- */
        uint32_t tot [4] = { 0, 0, 0, 0 }, tot_q = 0;
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            uint32_t w [4] [4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                int gi = 4 * c + i;        /* pixels 4 gi .. 4 gi + 3: row gi / 2, half gi % 2 */
                uint4 a = __ldg (cell_px + (size_t) (gi >> 1) * row_q + (gi & 1));
                uint32_t t01 = __byte_perm (a.x, a.y, 0x5140), u01 = __byte_perm (a.x, a.y, 0x7362);
                uint32_t t23 = __byte_perm (a.z, a.w, 0x5140), u23 = __byte_perm (a.z, a.w, 0x7362);
                w [0] [i] = __byte_perm (t01, t23, 0x5410);
                w [1] [i] = __byte_perm (t01, t23, 0x7632);
                w [2] [i] = __byte_perm (u01, u23, 0x5410);
                w [3] [i] = __byte_perm (u01, u23, 0x7632);
#pragma unroll
                for (int ch = 0; ch < 4; ch++)
                {
                    tot [ch] = __dp4a (w [ch] [i], 0x01010101u, tot [ch]);
                    tot_q = __dp4a (w [ch] [i], w [ch] [i], tot_q);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
                *(uint4 *) (a_ptr + ch * XT_PLANE_BYTES + c * 2048 + t * 16)
                    = make_uint4 (w [ch] [0], w [ch] [1], w [ch] [2], w [ch] [3]);
        }
        xtot [0 * 128 + t] = tot [0];
        xtot [1 * 128 + t] = tot [1];
        xtot [2 * 128 + t] = tot [2];
        xtot [3 * 128 + t] = tot [3];
        xtot [4 * 128 + t] = tot_q;
        fence_async_smem ();
        group_barrier (group);

        const int n_steps = n_chunks + n_chunks2;
        if (t == 0 && n_steps > 0)
        {
            tc_fence_after ();
            issue (0);
        }

        int best_e = SYMBOL_ERROR_MAX, best_s = -1;
        int best2_sum = 2 * SYMBOL_ERROR_MAX, best2_s = -1, best2_ea = SYMBOL_ERROR_MAX, best2_eb = SYMBOL_ERROR_MAX;
        uint32_t best2_fg = 0, best2_bg = 0;
        uint32_t rtot2 [4] = { 0, 0, 0, 0 }, rtot_q = 0;      /* right neighbour's sums, doubled */
        if (n_syms2 > 0 && t < 127)
        {
            rtot2 [0] = 2 * xtot [0 * 128 + t + 1];
            rtot2 [1] = 2 * xtot [1 * 128 + t + 1];
            rtot2 [2] = 2 * xtot [2 * 128 + t + 1];
            rtot2 [3] = 2 * xtot [3 * 128 + t + 1];
            rtot_q = xtot [4 * 128 + t + 1];
        }
        const uint32_t tot2 [4] = { 2 * tot [0], 2 * tot [1], 2 * tot [2], 2 * tot [3] };
        const int alpha_const = s_alpha_const;
        const bool opaque_warp = s_alpha_uniform && __all_sync (0xffffffffu, tot [3] == 64u * 255u);

        for (int step = 0; step < n_steps; step++)
        {
            uint32_t r [64];

            mbar_wait (bar, parity);
            parity ^= 1u;
            tc_fence_after ();
            if (step < n_chunks)
            {
#pragma unroll
                for (int ch = 0; ch < 4; ch++)
                    tmem_ld_32cols_pack16 (tmem_ld + ch * XT_NCHUNK, *(uint32_t (*) [16]) &r [16 * ch]);
            }
            else
            {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    tmem_ld_16cols_pack16 (tmem_ld + q * XT_WCHUNK, *(uint32_t (*) [8]) &r [8 * q]);
            }
            tc_wait_ld ();
            tc_fence_before ();
            if (step + 1 < n_steps && (t & 31) == 0)
            {
                __threadfence_block ();
                uint32_t before = atomicAdd (&s_arrivals [group], 1u);
                if ((before & 3u) == 3u)
                {
                    __threadfence_block ();
                    tc_fence_after ();
                    issue (step + 1);
                }
            }

            if (step < n_chunks)
            {
                /* r [16 ch + j]: channel ch of symbols 2 j (low half) and 2 j + 1.  All sums are carried
                 * doubled (2 S): that is what the error terms need, and it lets the mean be one wide
                 * multiply-add. */
                const int s0 = step * XT_NCHUNK;
                auto sweep = [&] (auto n_channels, int e0)
                {
                    constexpr int NCH = decltype (n_channels)::value;
#pragma unroll
                    for (int j = 0; j < 16; j++)
#pragma unroll
                        for (int half = 0; half < 2; half++)
                        {
                            const int s = s0 + 2 * j + half;
                            const uint2 sc = s_symc [s];
                            const int n = (int) (sc.x & 0xffu), m = 64 - n;
                            const uint32_t inv_n = sc.x & 0xffff0000u, inv_m = sc.y;
                            int e = e0;
#pragma unroll
                            for (int ch = 0; ch < NCH; ch++)
                            {
                                uint32_t sf2 = half ? (r [16 * ch + j] >> 15) & 0x1fffeu : (r [16 * ch + j] << 1) & 0x1fffeu;
                                uint32_t sb2 = tot2 [ch] - sf2;
                                e += xt_term ((int) xt_mean (sf2, inv_n), (int) sf2, n)
                                    + xt_term ((int) xt_mean (sb2, inv_m), (int) sb2, m);
                            }
                            if (e < best_e)
                            {
                                best_e = e;
                                best_s = s;
                            }
                        }
                };
                if (opaque_warp)
                    sweep (std::integral_constant<int, 3> (), (int) tot_q + alpha_const);
                else
                    sweep (std::integral_constant<int, 4> (), (int) tot_q);
            }
            else
            {
                /* r [8 (side * 4 + ch) + j]: wide symbols 2 j, 2 j + 1 of this step */
                const int s0 = (step - n_chunks) * XT_WCHUNK;
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int half = 0; half < 2; half++)
                    {
                        const int s = s0 + 2 * j + half;
                        if (s >= n_syms2)
                            continue;       /* (a pair's averaged colours can score worse than padding would) */
                        const uint4 sc = s_symc2 [s];
                        const int na = (int) (sc.x & 0xffu), nb = (int) (sc.z & 0xffu);
                        const uint32_t inv_na = sc.x & 0xffff0000u, inv_nb = sc.z & 0xffff0000u;
                        uint32_t fg = 0, bg = 0;
                        uint32_t sfa [4], sfb [4];
#pragma unroll
                        for (int ch = 0; ch < 4; ch++)
                        {
                            sfa [ch] = half ? (r [8 * ch + j] >> 15) & 0x1fffeu : (r [8 * ch + j] << 1) & 0x1fffeu;
                            sfb [ch] = half ? (r [8 * (4 + ch) + j] >> 15) & 0x1fffeu : (r [8 * (4 + ch) + j] << 1) & 0x1fffeu;
                            uint32_t fa = xt_mean (sfa [ch], inv_na), fb = xt_mean (sfb [ch], inv_nb);
                            uint32_t ba = xt_mean (tot2 [ch] - sfa [ch], sc.y), bb = xt_mean (rtot2 [ch] - sfb [ch], sc.w);
                            fg |= ((fa >> 1) + (fb >> 1)) << (8 * ch);
                            bg |= ((ba >> 1) + (bb >> 1)) << (8 * ch);
                        }
                        int ea = (int) tot_q, eb = (int) rtot_q;
#pragma unroll
                        for (int ch = 0; ch < 4; ch++)
                        {
                            int f = (int) ((fg >> (8 * ch)) & 0xffu), b = (int) ((bg >> (8 * ch)) & 0xffu);
                            ea += xt_term (f, (int) sfa [ch], na) + xt_term (b, (int) (tot2 [ch] - sfa [ch]), 64 - na);
                            eb += xt_term (f, (int) sfb [ch], nb) + xt_term (b, (int) (rtot2 [ch] - sfb [ch]), 64 - nb);
                        }
                        if (ea + eb < best2_sum)
                        {
                            best2_sum = ea + eb;
                            best2_s = s;
                            best2_ea = ea;
                            best2_eb = eb;
                            best2_fg = fg;
                            best2_bg = bg;
                        }
                    }
            }
        }

        /* ---- results ---- */
        {
            DevCellEval out;
            out.c = ' ';
            out.fg = out.bg = 0;
            out.error = SYMBOL_ERROR_MAX;
            out.mean = tc_sums_to_color (tot, 64, s_invdiv);
            if (best_s >= 0 && best_e < SYMBOL_ERROR_MAX)
            {
                /* the winner's colours: its masked sums once more, from this thread's row of the planes */
                const uint2 sc = s_symc [best_s];
                const int n = (int) (sc.x & 0xffu);
                uint32_t s_fg [4] = { 0, 0, 0, 0 }, s_bg [4];
#pragma unroll
                for (int c = 0; c < 4; c++)
                {
                    uint4 mk = *(const uint4 *) (s_bn + ((size_t) c * npad + best_s) * 16);
                    const uint32_t mw [4] = { mk.x, mk.y, mk.z, mk.w };
#pragma unroll
                    for (int ch = 0; ch < 4; ch++)
                    {
                        uint4 pv = *(const uint4 *) (a_ptr + ch * XT_PLANE_BYTES + c * 2048 + t * 16);
                        const uint32_t pw [4] = { pv.x, pv.y, pv.z, pv.w };
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            s_fg [ch] = __dp4a (pw [i], mw [i], s_fg [ch]);
                    }
                }
#pragma unroll
                for (int ch = 0; ch < 4; ch++)
                    s_bg [ch] = tot [ch] - s_fg [ch];
                uint32_t fg = tc_sums_to_color (s_fg, n, s_invdiv), bg = tc_sums_to_color (s_bg, 64 - n, s_invdiv);
                out.error = best_e;
                out.c = __ldg (cv.syms.chars + best_s);
                tc_update_cell_colors (cx_colors, out.c, fg, bg, out.fg, out.bg);
            }
            if (cell_valid && t < TC_TILE_STRIDE)
                evals [cell] = out;
        }
        if (n_syms2 > 0)
        {
            DevWideEval out;
            out.c = 0;
            out.fg = out.bg = 0;
            out.error_a = out.error_b = SYMBOL_ERROR_MAX;
            if (t < 127 && best2_s >= 0 && best2_sum < 2 * SYMBOL_ERROR_MAX)
            {
                out.error_a = best2_ea;
                out.error_b = best2_eb;
                out.c = __ldg (cv.syms.chars2 + best2_s);
                tc_update_cell_colors (cx_colors, out.c, best2_fg, best2_bg, out.fg, out.bg);
            }
            if (pair_valid)
                wide_evals [cell + 1] = out;
        }

        /* the planes become the next tile's */
        group_barrier (group);
    }

    tc_fence_before ();
    __syncthreads ();
    if (warp == 0)
    {
        tc_fence_after ();
        asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "n"(TC_TMEM_COLS) : "memory");
    }
}

/* ---- launcher ------------------------------------------------------------------------- */

static const uint8_t *tc_pen_lut (const DevCanvas *cv, int device, cudaStream_t stream);

static size_t
tc_smem_bytes (const DevCanvas *cv, bool wide)
{
    return (size_t) cv->syms.npad * (64 + 8) + (wide ? (size_t) cv->syms.n2pad * (128 + 16) : 0)
        + (size_t) TC_GROUPS * TC_GROUP_BYTES;
}

bool
dev_match_tc_applies (const DevCanvas *cv, bool wide)
{
    if (cv->work_factor_int >= 8 || cv->fg_only || cv->use_quantized_error || cv->color_extractor != 0
        || !cv->extract_colors || !cv->consider_inverted)
        return false;
    if (cv->syms.n <= 0 || cv->syms.n >= 4096 || cv->syms.n2 >= 4096 || !cv->syms.bimage)
        return false;
    if ((cv->width_px & 7) != 0)
        return false;
    return tc_smem_bytes (cv, wide) <= 227 * 1024 - 1024;
}

/* Pen tables, one per (device, palette type, colour space), built on first use and kept for the
 * life of the process (16 MB each). */
static std::mutex g_lut_mutex;
static std::map<std::tuple<int, int, int>, const uint8_t *> g_luts;

static const uint8_t *
tc_pen_lut (const DevCanvas *cv, int device, cudaStream_t stream)
{
    if (cv->canvas_mode == MODE_TRUECOLOR)
        return nullptr;
    int type = cv->fg_pal_type;
    if (type != cv->bg_pal_type || cv->fg_pal_transparent != PAL_INDEX_TRANSPARENT
        || cv->bg_pal_transparent != PAL_INDEX_TRANSPARENT)
        return nullptr;
    if (type != PAL_FIXED_256 && type != PAL_FIXED_240 && type != PAL_FIXED_16 && type != PAL_FIXED_8)
        return nullptr;

    std::lock_guard<std::mutex> lock (g_lut_mutex);
    auto key = std::make_tuple (device, type, (int) cv->color_space);
    auto it = g_luts.find (key);
    if (it != g_luts.end ())
        return it->second;

    uint8_t *lut = nullptr;
    if (cudaMalloc (&lut, 1u << 24) != cudaSuccess)
    {
        cudaGetLastError ();
        return nullptr;
    }
    tc_build_pen_lut_kernel<<<(1u << 24) / 4 / 256, 256, 0, stream>>> (cv->fg_palette, cv->color_space, (uint32_t *) lut);
    dev_count_launch (1);
    /* other canvases on other streams may use the table as soon as it is published */
    if (cudaGetLastError () != cudaSuccess || cudaStreamSynchronize (stream) != cudaSuccess)
    {
        cudaFree (lut);
        return nullptr;
    }
    g_luts [key] = lut;
    return lut;
}

/* Build (or find) the pen table this canvas's matcher will read, ahead of the first draw */
void
dev_match_tc_prepare (const DevCanvas *cv, void *stream)
{
    int dev = 0;
    cudaGetDevice (&dev);
    (void) tc_pen_lut (cv, dev, (cudaStream_t) stream);
}

static size_t
xtc_smem_bytes (const DevCanvas *cv, bool wide, int *npad_out, int *n2pad_out)
{
    int npad = (cv->syms.n + XT_NCHUNK - 1) / XT_NCHUNK * XT_NCHUNK;
    int n2pad = wide ? (cv->syms.n2 + XT_WCHUNK - 1) / XT_WCHUNK * XT_WCHUNK : 0;
    if (npad_out) *npad_out = npad;
    if (n2pad_out) *n2pad_out = n2pad;
    return (size_t) npad * (64 + 8) + (size_t) n2pad * (128 + 16) + (size_t) XT_GROUPS * XT_GROUP_BYTES;
}

/* same domain as the table-driven exhaustive kernel (dev_match_exhaustive_applies), if the tables fit */
bool
dev_match_xtc_applies (const DevCanvas *cv, bool wide)
{
    if (cv->work_factor_int < 8 || cv->fg_only || cv->use_quantized_error || cv->color_extractor != 0)
        return false;
    if (cv->syms.n <= 0 || !cv->syms.bimage || (cv->width_px & 7) != 0)
        return false;
    return xtc_smem_bytes (cv, wide, nullptr, nullptr) <= 227 * 1024 - 2048;
}

int
dev_launch_match_xtc (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, DevWideEval *wide_evals,
                      void *stream)
{
    int n_cells = cv->width * cv->n_rows;
    if (n_cells <= 0)
        return 0;

    int n_tiles = (n_cells + TC_TILE_STRIDE - 1) / TC_TILE_STRIDE;
    int dev = 0;
    cudaGetDevice (&dev);
    const int n_sm = dev_sm_count ();

    int npad, n2pad;
    size_t smem = xtc_smem_bytes (cv, wide_evals != nullptr && cv->syms.n2 > 0, &npad, &n2pad);
    int err = dev_opt_in_smem ((const void *) match_xtc_kernel<0>, smem);
    if (err)
        return err;

    DevCanvas cvl = *cv;
    cvl.pen_lut = tc_pen_lut (cv, dev, (cudaStream_t) stream);

    int grid = std::min ((n_tiles + XT_GROUPS - 1) / XT_GROUPS, n_sm);
    match_xtc_kernel<0><<<grid, XT_THREADS, smem, (cudaStream_t) stream>>> (cvl, pixels, evals,
                                                                            cv->syms.n2 > 0 ? wide_evals : nullptr, n_tiles,
                                                                            npad, n2pad);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_match_tc (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, DevWideEval *wide_evals,
                     void *stream)
{
    int n_cells = cv->width * cv->n_rows;
    if (n_cells <= 0)
        return 0;

    int n_tiles = (n_cells + TC_TILE_STRIDE - 1) / TC_TILE_STRIDE;
    int dev = 0;
    cudaGetDevice (&dev);
    const int n_sm = dev_sm_count ();

    size_t smem = tc_smem_bytes (cv, wide_evals != nullptr);
    int kmax = cv->n_candidates <= 5 ? 5 : 8;
    bool wide_on = wide_evals != nullptr && cv->syms.n2 > 0;
    bool exact = cv->n_candidates == kmax && 2 * cv->syms.n >= kmax && (!wide_on || 2 * cv->syms.n2 >= kmax);
    auto kernel = kmax == 5 ? (exact ? match_tc_kernel<5, true> : match_tc_kernel<5, false>)
                            : (exact ? match_tc_kernel<8, true> : match_tc_kernel<8, false>);
    int err = dev_opt_in_smem ((const void *) kernel, smem);
    if (err)
        return err;

    DevCanvas cvl = *cv;
    cvl.pen_lut = tc_pen_lut (cv, dev, (cudaStream_t) stream);

    /* A block holds every register of its SM whatever the number of its groups that have a tile.
     * Canvases with fewer tiles than the device has groups get at least two tiles per block, so
     * that the kernels of other frames in flight find free SMs (C1, 127 tiles: 64 blocks instead of
     * 127; 47 000 instead of 39 000 frames/s with three frames in flight, 45 instead of 44 us for a
     * frame alone).  CHAFA_B200_MATCH_PACK overrides the number for measurements. */
    static const int pack_env = [] { const char *e = getenv ("CHAFA_B200_MATCH_PACK"); return e ? atoi (e) : 0; } ();
    int pack = pack_env > 0 ? pack_env : std::max (2, (n_tiles + n_sm - 1) / n_sm);
    pack = std::min (pack, TC_GROUPS);
    int grid = std::min ((n_tiles + pack - 1) / pack, n_sm);
    kernel<<<grid, TC_THREADS, smem, (cudaStream_t) stream>>> (cvl, pixels, evals, wide_evals, n_tiles);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
