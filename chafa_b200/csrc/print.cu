/* print.cu -- kernel K5: canvas cells -> UTF-8 text with terminal control sequences.
 *
 * The reference prints serially with a running attribute context
 * (chafa/internal/chafa-canvas-printer.c:56-763).  Every emit_attributes_* function
 * ends by setting that context to the current cell's own attributes, and each row
 * ends with a reset, so the bytes a cell contributes depend only on the cell itself
 * and on the nearest preceding cell of the row with a non-zero code point.  That
 * makes printing a map + exclusive scan:
 *
 *   measure:  one block per row; per cell, run the emitter in counting mode, block-scan
 *             the lengths into per-cell offsets, store the row length
 *   offsets:  one block scans the row lengths
 *   emit:     per cell, run the same emitter in writing mode at its final offset
 *
 * Sequences are the caller's ChafaTermInfo compiled to literal bytes + argument slots
 * (chafa/chafa-term-info.c:413-445); arguments are formatted as unsigned 8-bit
 * decimals (chafa/internal/chafa-string-util.h:33-37), including the reference's
 * truncation of pen numbers to 8 bits and its aixterm / +30 / +40 offsets
 * (chafa-term-info.c:1712-1746). */

#include <cuda_runtime.h>

#include <cstring>

#include "device.h"
#include "resolve.cuh"

#define MODE_TRUECOLOR 0
#define MODE_INDEXED_256 1
#define MODE_INDEXED_240 2
#define MODE_INDEXED_16 3
#define MODE_FGBG_BGFG 4
#define MODE_FGBG 5
#define MODE_INDEXED_8 6
#define MODE_INDEXED_16_8 7

#define OPT_REUSE_ATTRIBUTES 1

#define NONE 0xffffffffu   /* "no colour" in the printer's attribute state */

template <bool WRITE>
struct Out
{
    uint8_t *p;
    size_t n;

    __device__ __forceinline__ void put (uint8_t b)
    {
        if (WRITE)
            p [n] = b;
        n++;
    }
};

template <bool WRITE>
__device__ __forceinline__ void
put_u8_dec (Out<WRITE> &o, uint32_t v)
{
    v &= 0xff;
    if (v >= 100) o.put ((uint8_t) ('0' + v / 100));
    if (v >= 10) o.put ((uint8_t) ('0' + (v / 10) % 10));
    o.put ((uint8_t) ('0' + v % 10));
}

template <bool WRITE>
__device__ void
emit_seq (Out<WRITE> &o, const DevTermInfo *ti, int seq, const uint32_t *args, int n_args)
{
    const DevSeq &s = ti->seqs [seq];
    int ofs = 0, i;

    if (n_args == 0)
    {
        for (int k = 0; k < s.pre_len [0]; k++)
            o.put (s.str [k]);
        return;
    }
    if (s.arg_index [0] == 255)
        return;

    for (i = 0; i < n_args; i++)
    {
        for (int k = 0; k < s.pre_len [i]; k++)
            o.put (s.str [ofs + k]);
        ofs += s.pre_len [i];
        if (s.arg_index [i] < DEV_SEQ_ARGS_MAX)
            put_u8_dec (o, args [s.arg_index [i]]);
    }
    for (int k = 0; k < s.pre_len [i]; k++)
        o.put (s.str [ofs + k]);
}

template <bool WRITE>
__device__ __forceinline__ void
put_utf8 (Out<WRITE> &o, uint32_t c)
{
    if (c < 0x80) { o.put ((uint8_t) c); }
    else if (c < 0x800) { o.put ((uint8_t) (0xc0 | (c >> 6))); o.put ((uint8_t) (0x80 | (c & 0x3f))); }
    else if (c < 0x10000)
    {
        o.put ((uint8_t) (0xe0 | (c >> 12)));
        o.put ((uint8_t) (0x80 | ((c >> 6) & 0x3f)));
        o.put ((uint8_t) (0x80 | (c & 0x3f)));
    }
    else if (c < 0x200000)
    {
        o.put ((uint8_t) (0xf0 | (c >> 18)));
        o.put ((uint8_t) (0x80 | ((c >> 12) & 0x3f)));
        o.put ((uint8_t) (0x80 | ((c >> 6) & 0x3f)));
        o.put ((uint8_t) (0x80 | (c & 0x3f)));
    }
    else if (c < 0x4000000)
    {
        o.put ((uint8_t) (0xf8 | (c >> 24)));
        o.put ((uint8_t) (0x80 | ((c >> 18) & 0x3f)));
        o.put ((uint8_t) (0x80 | ((c >> 12) & 0x3f)));
        o.put ((uint8_t) (0x80 | ((c >> 6) & 0x3f)));
        o.put ((uint8_t) (0x80 | (c & 0x3f)));
    }
    else
    {
        o.put ((uint8_t) (0xfc | (c >> 30)));
        o.put ((uint8_t) (0x80 | ((c >> 24) & 0x3f)));
        o.put ((uint8_t) (0x80 | ((c >> 18) & 0x3f)));
        o.put ((uint8_t) (0x80 | ((c >> 12) & 0x3f)));
        o.put ((uint8_t) (0x80 | ((c >> 6) & 0x3f)));
        o.put ((uint8_t) (0x80 | (c & 0x3f)));
    }
}

/* Attribute state after a cell has been printed */
struct Attr
{
    uint32_t fg, bg;     /* truecolor: 0x00RRGGBB or NONE; indexed: pen or 256 */
    bool inverted, bold;
};

__device__ __forceinline__ Attr
reset_state (int mode)
{
    Attr a;
    a.fg = a.bg = (mode == MODE_TRUECOLOR) ? NONE : (uint32_t) PAL_INDEX_TRANSPARENT;
    a.inverted = a.bold = false;
    return a;
}

/* The (fg, bg, inverted[, bold]) a cell asks for
 * (printer.c:215-252, 344-378, 437-471, 533-567) */
__device__ __forceinline__ Attr
cell_attr (const DevPrint &p, const DevCell &cell, bool &both_transparent)
{
    Attr a;

    if (p.canvas_mode == MODE_TRUECOLOR)
    {
        uint32_t fg = ((int) (cell.fg >> 24) < p.alpha_threshold) ? NONE : (cell.fg & 0xffffffu);
        uint32_t bg = ((int) (cell.bg >> 24) < p.alpha_threshold) ? NONE : (cell.bg & 0xffffffu);
        both_transparent = fg == NONE && bg == NONE;
        if (fg == NONE && bg != NONE) { a.fg = bg; a.bg = fg; a.inverted = true; }
        else { a.fg = fg; a.bg = bg; a.inverted = false; }
        a.bold = false;
    }
    else
    {
        uint32_t fg = cell.fg, bg = cell.bg;
        const uint32_t T = PAL_INDEX_TRANSPARENT;
        both_transparent = fg == T && bg == T;
        if (fg == T && bg != T) { a.fg = bg; a.bg = fg; a.inverted = true; }
        else { a.fg = fg; a.bg = bg; a.inverted = false; }
        a.bold = (p.canvas_mode == MODE_INDEXED_16_8) && a.fg > 7 && a.fg < 256;
    }
    return a;
}

template <bool WRITE>
__device__ void
emit_colors (Out<WRITE> &o, const DevPrint &p, bool have_fg, bool have_bg, uint32_t fg, uint32_t bg)
{
    uint32_t args [6];
    int mode = p.canvas_mode;

    if (mode == MODE_TRUECOLOR)
    {
        if (have_fg && have_bg)
        {
            args [0] = (fg >> 16) & 0xff; args [1] = (fg >> 8) & 0xff; args [2] = fg & 0xff;
            args [3] = (bg >> 16) & 0xff; args [4] = (bg >> 8) & 0xff; args [5] = bg & 0xff;
            emit_seq (o, p.ti, DSEQ_SET_COLOR_FGBG_DIRECT, args, 6);
        }
        else
        {
            uint32_t c = have_fg ? fg : bg;
            args [0] = (c >> 16) & 0xff; args [1] = (c >> 8) & 0xff; args [2] = c & 0xff;
            emit_seq (o, p.ti, have_fg ? DSEQ_SET_COLOR_FG_DIRECT : DSEQ_SET_COLOR_BG_DIRECT, args, 3);
        }
    }
    else if (mode == MODE_INDEXED_256 || mode == MODE_INDEXED_240)
    {
        if (have_fg && have_bg) { args [0] = fg & 0xff; args [1] = bg & 0xff; emit_seq (o, p.ti, DSEQ_SET_COLOR_FGBG_256, args, 2); }
        else if (have_fg) { args [0] = fg & 0xff; emit_seq (o, p.ti, DSEQ_SET_COLOR_FG_256, args, 1); }
        else { args [0] = bg & 0xff; emit_seq (o, p.ti, DSEQ_SET_COLOR_BG_256, args, 1); }
    }
    else if (mode == MODE_INDEXED_16 || mode == MODE_INDEXED_8)
    {
        /* aixterm offsets on the 8-bit truncated pen */
        uint32_t f = fg & 0xff, b = bg & 0xff;
        uint32_t fa = f + (f < 8 ? 30 : 82), ba = b + (b < 8 ? 40 : 92);
        if (have_fg && have_bg) { args [0] = fa; args [1] = ba; emit_seq (o, p.ti, DSEQ_SET_COLOR_FGBG_16, args, 2); }
        else if (have_fg) { args [0] = fa; emit_seq (o, p.ti, DSEQ_SET_COLOR_FG_16, args, 1); }
        else { args [0] = ba; emit_seq (o, p.ti, DSEQ_SET_COLOR_BG_16, args, 1); }
    }
    else  /* MODE_INDEXED_16_8 */
    {
        uint32_t f = ((fg & 7) & 0xff) + 30, b = (bg & 0xff) + 40;
        if (have_fg && have_bg) { args [0] = f; args [1] = b; emit_seq (o, p.ti, DSEQ_SET_COLOR_FGBG_8, args, 2); }
        else if (have_fg) { args [0] = f; emit_seq (o, p.ti, DSEQ_SET_COLOR_FG_8, args, 1); }
        else { args [0] = b; emit_seq (o, p.ti, DSEQ_SET_COLOR_BG_8, args, 1); }
    }
}

__device__ __forceinline__ Attr state_before (const DevPrint &p, const DevCell *row, int x);

/* What a cell hands to the character queue (chafa-canvas-printer.c:93-111) */
struct Queued
{
    uint32_t c;
    int n;
};

/* Bytes of one cell given the attribute state left by its predecessor.  With @queued the
 * character itself is not written but returned: the caller run-length codes it. */
template <bool WRITE>
__device__ void
emit_cell (Out<WRITE> &o, const DevPrint &p, const DevCell &cell, bool next_is_zero, bool is_last_in_row,
           const Attr &prev, Queued *queued = nullptr)
{
    int mode = p.canvas_mode;
    const uint32_t none = (mode == MODE_TRUECOLOR) ? NONE : (uint32_t) PAL_INDEX_TRANSPARENT;

    if (mode == MODE_FGBG)
    {
        if (queued) { queued->c = cell.c; queued->n = 1; }
        else put_utf8 (o, cell.c);
        return;
    }

    if (mode == MODE_FGBG_BGFG)
    {
        uint32_t c = cell.c;
        bool invert = false;

        if (cell.fg == cell.bg && p.fgbg_blank_symbol != 0 && (is_last_in_row || !next_is_zero))
        {
            c = p.fgbg_blank_symbol;
            if (c == 0x2588)
                invert = true;
        }
        if (cell.bg == PAL_INDEX_FG)
            invert = !invert;

        if (p.optimizations & OPT_REUSE_ATTRIBUTES)
        {
            if (!prev.inverted && invert)
                emit_seq (o, p.ti, DSEQ_INVERT_COLORS, nullptr, 0);
            else if (prev.inverted && !invert)
                emit_seq (o, p.ti, DSEQ_RESET_ATTRIBUTES, nullptr, 0);
        }
        else
        {
            emit_seq (o, p.ti, invert ? DSEQ_INVERT_COLORS : DSEQ_RESET_ATTRIBUTES, nullptr, 0);
        }
        if (queued) { queued->c = c; queued->n = 1; }
        else put_utf8 (o, c);
        return;
    }

    bool both_transparent;
    Attr a = cell_attr (p, cell, both_transparent);
    Attr s = prev;

    if (p.optimizations & OPT_REUSE_ATTRIBUTES)
    {
        bool skip_attr_handling = p.fg_only && mode != MODE_TRUECOLOR;

        if (!skip_attr_handling)
        {
            if (!p.fg_only
                && ((s.inverted && !a.inverted)
                    || (s.bold && !a.bold)
                    || (s.fg != none && a.fg == none)
                    || (s.bg != none && a.bg == none)))
            {
                emit_seq (o, p.ti, DSEQ_RESET_ATTRIBUTES, nullptr, 0);
                s = reset_state (mode);
            }
            if (!s.inverted && a.inverted)
                emit_seq (o, p.ti, DSEQ_INVERT_COLORS, nullptr, 0);
            if (mode == MODE_INDEXED_16_8 && !s.bold && a.bold)
                emit_seq (o, p.ti, DSEQ_ENABLE_BOLD, nullptr, 0);
        }

        if (a.fg != s.fg)
        {
            if (a.bg != s.bg && a.bg != none)
                emit_colors (o, p, true, true, a.fg, a.bg);
            else if (a.fg != none)
                emit_colors (o, p, true, false, a.fg, a.bg);
        }
        else if (a.bg != s.bg && a.bg != none)
        {
            emit_colors (o, p, false, true, a.fg, a.bg);
        }
    }
    else
    {
        emit_seq (o, p.ti, DSEQ_RESET_ATTRIBUTES, nullptr, 0);
        if (a.inverted)
            emit_seq (o, p.ti, DSEQ_INVERT_COLORS, nullptr, 0);
        if (mode == MODE_INDEXED_16_8 && a.fg > 7)
            emit_seq (o, p.ti, DSEQ_ENABLE_BOLD, nullptr, 0);

        if (a.fg != none)
            emit_colors (o, p, true, a.bg != none, a.fg, a.bg);
        else if (a.bg != none)
            emit_colors (o, p, false, true, a.fg, a.bg);
    }

    if (both_transparent)
    {
        bool twice = !is_last_in_row && next_is_zero;
        if (queued) { queued->c = ' '; queued->n = twice ? 2 : 1; }
        else
        {
            o.put (' ');
            if (twice)
                o.put (' ');
        }
    }
    else
    {
        if (queued) { queued->c = cell.c; queued->n = 1; }
        else put_utf8 (o, cell.c);
    }
}

/* ---- repeat-character compression (chafa-canvas-printer.c:56-111) -------------------
 * The reference queues characters and flushes the queue before any control sequence and
 * at the end of a row; a flushed run of n equal characters becomes the character followed
 * by REPEAT_CHAR (n - 1) when that is shorter.  So a run starts at a cell that emits a
 * control sequence or whose character differs from the previous emitted one, and extends
 * over the following cells that do neither.  The run's head prints the whole run. */

__device__ __forceinline__ int
utf8_len (uint32_t c)
{
    return c < 0x80 ? 1 : c < 0x800 ? 2 : c < 0x10000 ? 3 : c < 0x200000 ? 4 : c < 0x4000000 ? 5 : 6;
}

template <bool WRITE>
__device__ void
put_dec_9999 (Out<WRITE> &o, uint32_t v)
{
    if (v > 9999) v = 9999;
    if (v >= 1000) o.put ((uint8_t) ('0' + v / 1000));
    if (v >= 100) o.put ((uint8_t) ('0' + (v / 100) % 10));
    if (v >= 10) o.put ((uint8_t) ('0' + (v / 10) % 10));
    o.put ((uint8_t) ('0' + v % 10));
}

/* One flushed run */
template <bool WRITE>
__device__ void
emit_run (Out<WRITE> &o, const DevPrint &p, uint32_t c, int n)
{
    int len = utf8_len (c);

    if (n > 1 && n * len > len + 4)
    {
        const DevSeq &s = p.ti->seqs [DSEQ_REPEAT_CHAR];
        put_utf8 (o, c);
        for (int k = 0; k < s.pre_len [0]; k++)
            o.put (s.str [k]);
        if (s.arg_index [0] != 255)
        {
            put_dec_9999 (o, (uint32_t) (n - 1));
            for (int k = 0; k < s.pre_len [1]; k++)
                o.put (s.str [s.pre_len [0] + k]);
        }
    }
    else
    {
        for (int i = 0; i < n; i++)
            put_utf8 (o, c);
    }
}

/* Attribute bytes, queued character and run-head flag of cell @x; returns false for cells
 * that print nothing (right halves of wide symbols) */
template <bool WRITE>
__device__ bool
rep_cell (Out<WRITE> &o, const DevPrint &p, const DevCell *row, int x, Queued &q, bool &is_head)
{
    DevCell cell = row [x];
    if (cell.c == 0)
        return false;

    bool next_zero = (x + 1 < p.width) && row [x + 1].c == 0;
    size_t before = o.n;
    emit_cell (o, p, cell, next_zero, x == p.width - 1, state_before (p, row, x), &q);
    bool has_attrs = o.n != before;

    /* previous printing cell's queued character */
    is_head = true;
    if (!has_attrs)
    {
        for (int j = x - 1; j >= 0; j--)
        {
            if (row [j].c == 0)
                continue;
            Out<false> dummy;
            dummy.p = nullptr;
            dummy.n = 0;
            Queued pq;
            bool pnz = (j + 1 < p.width) && row [j + 1].c == 0;
            emit_cell (dummy, p, row [j], pnz, j == p.width - 1, state_before (p, row, j), &pq);
            is_head = pq.c != q.c;
            break;
        }
    }
    return true;
}

/* Total number of queued characters of the run headed by cell @x */
__device__ int
rep_run_length (const DevPrint &p, const DevCell *row, int x, const Queued &q)
{
    int n = q.n;
    for (int j = x + 1; j < p.width; j++)
    {
        if (row [j].c == 0)
            continue;
        Out<false> dummy;
        dummy.p = nullptr;
        dummy.n = 0;
        Queued nq;
        bool nnz = (j + 1 < p.width) && row [j + 1].c == 0;
        emit_cell (dummy, p, row [j], nnz, j == p.width - 1, state_before (p, row, j), &nq);
        if (dummy.n != 0 || nq.c != q.c)
            break;
        n += nq.n;
    }
    return n;
}

/* Attribute state seen by cell @x of a row: that of the nearest cell to the left with
 * a non-zero code point, or the reset state. */
__device__ __forceinline__ Attr
state_before (const DevPrint &p, const DevCell *row, int x)
{
    int mode = p.canvas_mode;

    for (int j = x - 1; j >= 0; j--)
    {
        DevCell c = row [j];
        if (c.c == 0)
            continue;

        if (mode == MODE_FGBG_BGFG)
        {
            Attr a = reset_state (mode);
            bool next_zero = (j + 1 < p.width) && row [j + 1].c == 0;
            bool invert = false;
            if (c.fg == c.bg && p.fgbg_blank_symbol != 0 && (j == p.width - 1 || !next_zero))
                invert = p.fgbg_blank_symbol == 0x2588;
            if (c.bg == PAL_INDEX_FG)
                invert = !invert;
            a.inverted = invert;
            return a;
        }
        bool bt;
        return cell_attr (p, c, bt);
    }
    return reset_state (mode);
}

template <bool WRITE>
__device__ __forceinline__ void
emit_row_reset (Out<WRITE> &o, const DevPrint &p)
{
    if (p.canvas_mode == MODE_FGBG)
        return;
    emit_seq (o, p.ti, p.fg_only ? DSEQ_RESET_COLOR_FG : DSEQ_RESET_ATTRIBUTES, nullptr, 0);
}

#define PRINT_THREADS 512

__global__ void __launch_bounds__ (PRINT_THREADS)
print_measure_kernel (DevPrint p)
{
    __shared__ uint32_t s_warp [PRINT_THREADS / 32];
    __shared__ uint32_t s_carry;

    int r = blockIdx.x;
    int row_index = p.first_row + r;
    const DevCell *row = p.cells + (size_t) row_index * p.width;
    uint32_t *ofs = p.cell_len + (size_t) r * p.width;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads ();

    for (int base = 0; base < p.width; base += PRINT_THREADS)
    {
        int x = base + threadIdx.x;
        uint32_t len = 0;

        if (x < p.width)
        {
            DevCell cell = row [x];
            if (cell.c != 0 && p.have_repeat)
            {
                Out<false> o;
                o.p = nullptr;
                o.n = 0;
                Queued q;
                bool is_head;
                rep_cell (o, p, row, x, q, is_head);
                if (is_head)
                    emit_run (o, p, q.c, rep_run_length (p, row, x, q));
                len = (uint32_t) o.n;     /* followers of a run print nothing */
            }
            else if (cell.c != 0)
            {
                Out<false> o;
                o.p = nullptr;
                o.n = 0;
                bool next_zero = (x + 1 < p.width) && row [x + 1].c == 0;
                emit_cell (o, p, cell, next_zero, x == p.width - 1, state_before (p, row, x));
                len = (uint32_t) o.n;
            }
        }

        /* block exclusive scan */
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            uint32_t t = __shfl_up_sync (0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31)
            s_warp [warp] = incl;
        __syncthreads ();
        uint32_t warp_base = 0, total = 0;
        for (int w = 0; w < PRINT_THREADS / 32; w++)
        {
            if (w < warp) warp_base += s_warp [w];
            total += s_warp [w];
        }
        uint32_t carry = s_carry;
        if (x < p.width)
            ofs [x] = carry + warp_base + incl - len;
        __syncthreads ();
        if (threadIdx.x == 0)
            s_carry = carry + total;
        __syncthreads ();
    }

    if (threadIdx.x == 0)
    {
        Out<false> o;
        o.p = nullptr;
        o.n = 0;
        if (row_index == 0)
            emit_row_reset (o, p);     /* leading reset, first canvas row only */
        emit_row_reset (o, p);         /* trailing reset */
        size_t len = o.n + s_carry;
        if (row_index < p.total_height - 1)
            len++;                     /* newline */
        p.row_ofs [r + 1] = len;
        if (r == 0)
            p.row_ofs [0] = 0;
    }
}

/* In-place inclusive scan of row_ofs [1..n_rows], one warp */
__global__ void __launch_bounds__ (32)
print_offsets_kernel (DevPrint p)
{
    const int lane = threadIdx.x;
    uint64_t carry = 0;

    for (int base = 1; base <= p.n_rows; base += 32)
    {
        int r = base + lane;
        uint64_t v = r <= p.n_rows ? p.row_ofs [r] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            uint64_t o = __shfl_up_sync (0xffffffffu, v, d);
            if (lane >= d)
                v += o;
        }
        v += carry;
        if (r <= p.n_rows)
            p.row_ofs [r] = v;
        carry = __shfl_sync (0xffffffffu, v, 31);
    }
}

__global__ void __launch_bounds__ (PRINT_THREADS)
print_emit_kernel (DevPrint p)
{
    int r = blockIdx.x;
    int row_index = p.first_row + r;
    const DevCell *row = p.cells + (size_t) row_index * p.width;
    const uint32_t *ofs = p.cell_len + (size_t) r * p.width;
    uint64_t row_begin = p.row_ofs [r], row_end = p.row_ofs [r + 1];

    if (row_end > p.out_capacity)
        return;

    size_t prefix_len = 0;
    {
        Out<false> o;
        o.p = nullptr;
        o.n = 0;
        if (row_index == 0)
            emit_row_reset (o, p);
        prefix_len = o.n;
    }

    if (threadIdx.x == 0)
    {
        Out<true> o;
        o.p = p.out + row_begin;
        o.n = 0;
        if (row_index == 0)
            emit_row_reset (o, p);

        /* trailing reset and newline */
        Out<false> m;
        m.p = nullptr;
        m.n = 0;
        emit_row_reset (m, p);
        size_t tail = m.n + (row_index < p.total_height - 1 ? 1 : 0);
        Out<true> t;
        t.p = p.out + row_end - tail;
        t.n = 0;
        emit_row_reset (t, p);
        if (row_index < p.total_height - 1)
            t.put ('\n');
    }

    for (int x = threadIdx.x; x < p.width; x += PRINT_THREADS)
    {
        DevCell cell = row [x];
        if (cell.c == 0)
            continue;
        Out<true> o;
        o.p = p.out + row_begin + prefix_len + ofs [x];
        o.n = 0;
        if (p.have_repeat)
        {
            Queued q;
            bool is_head;
            rep_cell (o, p, row, x, q, is_head);
            if (is_head)
                emit_run (o, p, q.c, rep_run_length (p, row, x, q));
            continue;
        }
        bool next_zero = (x + 1 < p.width) && row [x + 1].c == 0;
        emit_cell (o, p, cell, next_zero, x == p.width - 1, state_before (p, row, x));
    }
}

/* ---- single-launch printer ------------------------------------------------------------
 * Block = one row, as above, but measure, offsets and emit are one kernel: a row publishes its
 * length as soon as it is measured and finds its start with a decoupled look-back over the rows
 * before it (status word = epoch | state | 40-bit value; state 1 = this row's length, 2 = inclusive
 * prefix).  The row's text is assembled in shared memory, at the same 16-byte phase as its place
 * in the output, and copied out with 128-bit stores.  Rows take their index from a ticket so that
 * a row only ever waits for rows that have already started. */

#define ST_SHIFT_EPOCH 42
#define ST_SHIFT_STATE 40
#define ST_VALUE_MASK ((1ull << 40) - 1)

/* MODE >= 0: specialised for that canvas mode without REPEAT_CHAR runs and without the
 * foreground-only switch (the kernel works on a copy of the parameter block in which those fields are
 * compile-time constants, so the emitter's mode switches fold away); MODE = -1: everything at run time. */
template <int MODE>
__global__ void __launch_bounds__ (PRINT_THREADS, 2)
print_fused_kernel (const __grid_constant__ DevPrint p_in, const __grid_constant__ DevResolveArgs rs)
{
    DevPrint p = p_in;
    if (MODE >= 0)
    {
        p.canvas_mode = MODE;
        p.have_repeat = 0;
        p.fg_only = 0;
    }

    extern __shared__ __align__ (16) uint8_t s_dyn [];
    __shared__ DevTermInfo s_ti;

    /* the control sequences are read byte by byte: keep them next to the SM */
    {
        const uint32_t *src = (const uint32_t *) p.ti;
        uint32_t *dst = (uint32_t *) &s_ti;
        for (int i = threadIdx.x; i < (int) (sizeof (DevTermInfo) / 4); i += PRINT_THREADS)
            dst [i] = src [i];
        p.ti = &s_ti;
    }
    __shared__ uint32_t s_warp [PRINT_THREADS / 32];
    __shared__ uint32_t s_carry;
    __shared__ int s_row;
    __shared__ unsigned long long s_begin;

    uint32_t *s_ofs = (uint32_t *) s_dyn;                                  /* width */
    uint8_t *s_text = s_dyn + (((size_t) p.width * 4 + 15) & ~(size_t) 15);  /* 16 + row bound */

    if (threadIdx.x == 0)
    {
        s_row = (int) atomicAdd (p.ctrl, 1u);
        s_carry = 0;
    }
    __syncthreads ();

    const int r = s_row;
    const int row_index = p.first_row + r;
    const DevCell *row = p.cells + (size_t) row_index * p.width;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long epoch = (unsigned long long) p.epoch;
    volatile unsigned long long *status = p.status;

    /* ---- the row walk of the draw, if it is still owed (resolve.cuh): this block's row ---- */
    if (rs.enabled)
    {
        resolve_row (rs.cv, rs.evals, rs.wide_evals, rs.cells, r);
        __syncthreads ();       /* the row's cells are read back below by other threads of the block */
    }

    /* ---- measure ---- */
    for (int base = 0; base < p.width; base += PRINT_THREADS)
    {
        int x = base + threadIdx.x;
        uint32_t len = 0;

        if (x < p.width)
        {
            DevCell cell = row [x];
            if (cell.c != 0 && p.have_repeat)
            {
                Out<false> o;
                o.p = nullptr;
                o.n = 0;
                Queued q;
                bool is_head;
                rep_cell (o, p, row, x, q, is_head);
                if (is_head)
                    emit_run (o, p, q.c, rep_run_length (p, row, x, q));
                len = (uint32_t) o.n;
            }
            else if (cell.c != 0)
            {
                Out<false> o;
                o.p = nullptr;
                o.n = 0;
                bool next_zero = (x + 1 < p.width) && row [x + 1].c == 0;
                emit_cell (o, p, cell, next_zero, x == p.width - 1, state_before (p, row, x));
                len = (uint32_t) o.n;
            }
        }

        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            uint32_t t = __shfl_up_sync (0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31)
            s_warp [warp] = incl;
        __syncthreads ();
        uint32_t warp_base = 0, total = 0;
        for (int w = 0; w < PRINT_THREADS / 32; w++)
        {
            if (w < warp) warp_base += s_warp [w];
            total += s_warp [w];
        }
        uint32_t carry = s_carry;
        if (x < p.width)
            s_ofs [x] = carry + warp_base + incl - len;
        __syncthreads ();
        if (threadIdx.x == 0)
            s_carry = carry + total;
        __syncthreads ();
    }

    /* leading reset (first canvas row only), trailing reset, newline */
    size_t prefix_len, tail_len;
    {
        Out<false> o;
        o.p = nullptr;
        o.n = 0;
        emit_row_reset (o, p);
        prefix_len = row_index == 0 ? o.n : 0;
        tail_len = o.n + (row_index < p.total_height - 1 ? 1 : 0);
    }
    const unsigned long long row_len = prefix_len + s_carry + tail_len;

    /* ---- where does this row start? ---- */
    if (warp == 0)
    {
        unsigned long long excl = 0;

        if (lane == 0)
        {
            unsigned long long st = (epoch << ST_SHIFT_EPOCH) | ((r == 0 ? 2ull : 1ull) << ST_SHIFT_STATE) | row_len;
            status [r] = st;
            __threadfence ();
        }
        if (r > 0)
        {
            int hi = r - 1;
            for (;;)
            {
                int idx = hi - lane;
                unsigned long long st = (epoch << ST_SHIFT_EPOCH) | (2ull << ST_SHIFT_STATE);   /* before row 0: prefix 0 */
                if (idx >= 0)
                {
                    do
                        st = status [idx];
                    while ((st >> ST_SHIFT_EPOCH) != epoch);
                }
                unsigned has_prefix = __ballot_sync (0xffffffffu, ((st >> ST_SHIFT_STATE) & 3ull) == 2ull);
                int first = has_prefix ? __ffs (has_prefix) - 1 : 31;
                unsigned long long v = lane <= first ? (st & ST_VALUE_MASK) : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1)
                    v += __shfl_xor_sync (0xffffffffu, v, d);
                excl += v;
                if (has_prefix)
                    break;
                hi -= 32;
            }
            if (lane == 0)
            {
                __threadfence ();
                status [r] = (epoch << ST_SHIFT_EPOCH) | (2ull << ST_SHIFT_STATE) | (excl + row_len);
            }
        }
        if (lane == 0)
        {
            s_begin = excl;
            p.row_ofs [r + 1] = excl + row_len;
            if (r == 0)
                p.row_ofs [0] = 0;
        }
    }
    __syncthreads ();

    const unsigned long long row_begin = s_begin, row_end = row_begin + row_len;
    const bool fits = row_end <= p.out_capacity;
    const bool staged = p.stage_bytes != 0;
    /* staged: byte i of the row goes to s_text [phase + i], phase = row_begin mod 16 */
    const uint32_t phase = (uint32_t) (row_begin & 15);
    uint8_t *dst = staged ? s_text + phase : p.out + row_begin;

    if (fits)
    {
        if (threadIdx.x == 0)
        {
            Out<true> o;
            o.p = dst;
            o.n = 0;
            if (row_index == 0)
                emit_row_reset (o, p);
            Out<true> t;
            t.p = dst + row_len - tail_len;
            t.n = 0;
            emit_row_reset (t, p);
            if (row_index < p.total_height - 1)
                t.put ('\n');
        }

        for (int x = threadIdx.x; x < p.width; x += PRINT_THREADS)
        {
            DevCell cell = row [x];
            if (cell.c == 0)
                continue;
            Out<true> o;
            o.p = dst + prefix_len + s_ofs [x];
            o.n = 0;
            if (p.have_repeat)
            {
                Queued q;
                bool is_head;
                rep_cell (o, p, row, x, q, is_head);
                if (is_head)
                    emit_run (o, p, q.c, rep_run_length (p, row, x, q));
                continue;
            }
            bool next_zero = (x + 1 < p.width) && row [x + 1].c == 0;
            emit_cell (o, p, cell, next_zero, x == p.width - 1, state_before (p, row, x));
        }

        if (staged)
        {
            __syncthreads ();
            /* bytes before the first and after the last whole 16-byte unit, then the units */
            const unsigned long long first_unit = (row_begin + 15) & ~15ull, last_unit = row_end & ~15ull;
            if (first_unit >= last_unit)
            {
                for (unsigned long long i = row_begin + threadIdx.x; i < row_end; i += PRINT_THREADS)
                    p.out [i] = s_text [phase + (i - row_begin)];
            }
            else
            {
                for (unsigned long long i = row_begin + threadIdx.x; i < first_unit; i += PRINT_THREADS)
                    p.out [i] = s_text [phase + (i - row_begin)];
                for (unsigned long long i = last_unit + threadIdx.x; i < row_end; i += PRINT_THREADS)
                    p.out [i] = s_text [phase + (i - row_begin)];
                const uint4 *s4 = (const uint4 *) (s_text + phase + (first_unit - row_begin));
                uint4 *g4 = (uint4 *) (p.out + first_unit);
                const unsigned long long n4 = (last_unit - first_unit) >> 4;
                for (unsigned long long i = threadIdx.x; i < n4; i += PRINT_THREADS)
                    g4 [i] = s4 [i];
            }
        }
    }

    /* the last row to finish rearms the ticket for the next launch */
    __syncthreads ();
    if (threadIdx.x == 0)
    {
        unsigned done = atomicAdd (p.ctrl + 1, 1u);
        if (done == (unsigned) p.n_rows - 1)
        {
            p.ctrl [0] = 0;
            p.ctrl [1] = 0;
        }
    }
}

size_t
dev_print_fused_smem_bytes (int width, size_t row_bound)
{
    return (((size_t) width * 4 + 15) & ~(size_t) 15) + 16 + ((row_bound + 15) & ~(size_t) 15) + 16;
}

/* One launch: measure, offsets, emit.  p->ctrl (2 words, zero before the first use), p->status
 * (n_rows words) and p->epoch (changes with every launch, never 0) drive the look-back;
 * p->stage_bytes = bound of one row's text, or 0 to write straight to the output. */
int
dev_launch_print_fused (const DevPrint *p, const DevResolveArgs *resolve, void *stream)
{
    static_assert (PRINT_THREADS == RESOLVE_THREADS, "the row walk runs on the printer's block");
    if (p->n_rows <= 0)
        return 0;
    DevResolveArgs rs;
    if (resolve && resolve->enabled)
        rs = *resolve;
    else
        memset (&rs, 0, sizeof (rs));
    size_t smem = p->stage_bytes ? dev_print_fused_smem_bytes (p->width, p->stage_bytes)
                                 : (((size_t) p->width * 4 + 15) & ~(size_t) 15) + 16;
    auto launch = [&] (auto kernel) -> cudaError_t
    {
        if (smem > 48 * 1024)
        {
            cudaError_t e = (cudaError_t) dev_opt_in_smem ((const void *) kernel, smem);
            if (e != cudaSuccess)
                return e;
        }
        kernel<<<p->n_rows, PRINT_THREADS, smem, (cudaStream_t) stream>>> (*p, rs);
        return cudaSuccess;
    };
    cudaError_t lerr;
    bool plain = !p->have_repeat && !p->fg_only;
    if (plain && p->canvas_mode == MODE_TRUECOLOR) lerr = launch (print_fused_kernel<MODE_TRUECOLOR>);
    else if (plain && p->canvas_mode == MODE_INDEXED_256) lerr = launch (print_fused_kernel<MODE_INDEXED_256>);
    else if (plain && p->canvas_mode == MODE_INDEXED_240) lerr = launch (print_fused_kernel<MODE_INDEXED_240>);
    else if (plain && p->canvas_mode == MODE_INDEXED_16) lerr = launch (print_fused_kernel<MODE_INDEXED_16>);
    else lerr = launch (print_fused_kernel<-1>);
    if (lerr != cudaSuccess)
        return (int) lerr;
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

int
dev_launch_print_measure (const DevPrint *p, void *stream)
{
    if (p->n_rows <= 0)
        return 0;
    print_measure_kernel<<<p->n_rows, PRINT_THREADS, 0, (cudaStream_t) stream>>> (*p);
    print_offsets_kernel<<<1, 32, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (2);
    return (int) cudaGetLastError ();
}

int
dev_launch_print_emit (const DevPrint *p, void *stream)
{
    if (p->n_rows <= 0)
        return 0;
    print_emit_kernel<<<p->n_rows, PRINT_THREADS, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
