/* host_scale.cpp -- scaler plan: per-dimension filter choice and sample tables.
 *
 * Host half of kernel K1.  Restates the set-up arithmetic of the reference's scaler
 * (chafa/internal/smolscale/smolscale.c: pick_filter_params 888-1016, init_dim
 * 1483-1556; smolscale-generic.c: precalc_* 33-269) so that the device kernel
 * samples exactly the source positions with exactly the weights the reference
 * uses.  All geometry is in 1/256 pixel units ("spx").  The plan is a few KB and
 * is uploaded once per draw; the per-pixel work is in scale.cu. */

#include "internal.h"
#include "device.h"

#include <algorithm>

#define SPX 256
#define SMALL_MUL 256u
#define BIG_MUL 65536u
#define BOXES_MULTIPLIER ((uint64_t) BIG_MUL * SMALL_MUL)
#define BILIN_MULTIPLIER ((uint64_t) BIG_MUL * BIG_MUL)
#define BILIN_BOX_CUTOFF 14
#define OPACITY_MAX 256

static inline int64_t spx_mod (int64_t n) { return ((n % SPX) + SPX) % SPX; }
static inline int64_t spx_to_px (int64_t spx) { return (spx + SPX - 1) / SPX; }

struct DimSetup
{
    uint32_t src_size_spx, src_size_px;
    uint32_t dest_size_spx, dest_size_px;
    int32_t placement_ofs_spx, placement_size_spx;
    uint32_t n_halvings = 0;
    uint32_t prehalving_px = 0, prehalving_spx = 0;
    int filter = SCALE_FILTER_COPY;
    uint16_t first_opacity = OPACITY_MAX, last_opacity = OPACITY_MAX;
    int32_t clear_before_px, clear_after_px, clip_before_px, clip_after_px;
    int32_t placement_ofs_px, placement_size_px;
};

static void
pick_filter (DimSetup *d, uint32_t dest_dim, bool nearest)
{
    uint32_t src_dim = d->src_size_px;
    uint32_t src_dim_spx = d->src_size_spx;
    int32_t dest_ofs_spx = d->placement_ofs_spx;
    uint32_t dest_dim_spx = (uint32_t) d->placement_size_spx;

    d->prehalving_px = dest_dim;
    d->first_opacity = (uint16_t) (spx_mod (-((int64_t) dest_ofs_spx) - 1) + 1);
    d->last_opacity = (uint16_t) (spx_mod ((int64_t) dest_ofs_spx + dest_dim_spx - 1) + 1);

    if (dest_dim == 1)
    {
        d->first_opacity = (uint16_t) dest_dim_spx;
        d->last_opacity = OPACITY_MAX;
    }

    if (dest_dim_spx == 0)
    {
        d->prehalving_spx = 0;
        d->n_halvings = 0;
        d->filter = SCALE_FILTER_ONE;
        return;
    }

    if (nearest)
    {
        d->prehalving_spx = dest_dim_spx;
        d->n_halvings = 0;
        d->first_opacity = d->last_opacity = OPACITY_MAX;
        if (src_dim <= 1)
            d->filter = SCALE_FILTER_ONE;
        else if ((dest_ofs_spx & 0xff) == 0 && src_dim_spx == dest_dim_spx)
            d->filter = SCALE_FILTER_COPY;
        else
            d->filter = SCALE_FILTER_NEAREST;
        return;
    }

    uint64_t floor_dim = std::max<uint64_t> (dest_dim_spx, SPX);

    if ((uint64_t) src_dim_spx > floor_dim * 255)
    {
        d->filter = SCALE_FILTER_BOX;
    }
    else if ((uint64_t) src_dim_spx >= floor_dim * BILIN_BOX_CUTOFF)
    {
        d->filter = SCALE_FILTER_BOX;
    }
    else if (src_dim <= 1)
    {
        d->filter = SCALE_FILTER_ONE;
    }
    else if ((dest_ofs_spx & 0xff) == 0 && src_dim_spx == dest_dim_spx)
    {
        d->filter = SCALE_FILTER_COPY;
        d->first_opacity = d->last_opacity = OPACITY_MAX;
    }
    else
    {
        uint32_t n_halvings = 0;
        uint64_t t = floor_dim;

        for (;;)
        {
            t *= 2;
            if (t >= src_dim_spx)
                break;
            n_halvings++;
        }
        n_halvings = std::min (n_halvings, 3u);

        d->prehalving_px = dest_dim << n_halvings;
        d->prehalving_spx = dest_dim_spx << n_halvings;
        d->filter = SCALE_FILTER_BILINEAR_0H + (int) n_halvings;
        d->n_halvings = n_halvings;
    }
}

static void
setup_dim (DimSetup *d, uint32_t src_size_spx, uint32_t dest_size_spx,
           int32_t placement_ofs_spx, int32_t placement_size_spx, bool nearest)
{
    int64_t placement_ofs_px, placement_size_px;

    d->src_size_spx = src_size_spx;
    d->src_size_px = (uint32_t) spx_to_px (src_size_spx);
    d->dest_size_spx = dest_size_spx;
    d->dest_size_px = (uint32_t) spx_to_px (dest_size_spx);
    d->placement_ofs_spx = placement_ofs_spx;
    d->placement_size_spx = placement_size_spx;

    if (placement_ofs_spx < 0)
        placement_ofs_px = ((int64_t) placement_ofs_spx - 255) / SPX;
    else
        placement_ofs_px = placement_ofs_spx / SPX;
    placement_size_px = spx_to_px ((int64_t) placement_size_spx + spx_mod (placement_ofs_spx));

    pick_filter (d, (uint32_t) placement_size_px, nearest);

    int64_t visible_first = std::min (std::max<int64_t> (placement_ofs_px, 0), (int64_t) d->dest_size_px);
    int64_t visible_end = std::max (std::min (placement_ofs_px + placement_size_px, (int64_t) d->dest_size_px),
                                    visible_first);

    d->clear_before_px = (int32_t) visible_first;
    d->clear_after_px = (int32_t) (d->dest_size_px - visible_end);
    d->clip_before_px = (int32_t) (visible_first - placement_ofs_px);
    d->clip_after_px = (int32_t) ((placement_ofs_px + placement_size_px) - visible_end);

    if (d->clip_before_px > 0)
        d->first_opacity = OPACITY_MAX;
    if (d->clip_after_px > 0)
        d->last_opacity = OPACITY_MAX;

    d->placement_ofs_px = (int32_t) visible_first;
    d->placement_size_px = (int32_t) (visible_end - visible_first);
}

/* Bilinear sample table: one (source offset, weight of the left/top sample) pair
 * per pre-halving sample inside the visible window. */
static void
linear_range (std::vector<uint32_t> &out, int64_t first_index, int64_t last_index,
              uint64_t first_sample_ofs, uint64_t sample_step, int sample_ofs_px_max,
              int64_t clip_first, int64_t clip_last)
{
    if (first_index < clip_first)
    {
        first_sample_ofs += sample_step * (uint64_t) (clip_first - first_index);
        first_index = clip_first;
    }
    if (last_index > clip_last)
        last_index = clip_last;

    uint64_t sample_ofs = first_sample_ofs;

    for (int64_t i = first_index; i < last_index; i++)
    {
        uint64_t sample_ofs_px = sample_ofs / BILIN_MULTIPLIER;

        if (sample_ofs_px >= (uint64_t) sample_ofs_px_max - 1)
        {
            /* Note: the running offset is deliberately not advanced on this branch */
            out.push_back ((uint32_t) (uint16_t) (sample_ofs_px_max - 2));
            continue;
        }

        uint32_t f = SMALL_MUL - (uint32_t) ((sample_ofs / (BILIN_MULTIPLIER / SMALL_MUL)) % SMALL_MUL);
        out.push_back ((uint32_t) (uint16_t) sample_ofs_px | ((uint32_t) (uint16_t) f << 16));
        sample_ofs += sample_step;
    }
}

static void
precalc_bilinear (std::vector<uint32_t> &out, const DimSetup *d)
{
    uint64_t src_dim_spx = d->src_size_spx;
    uint64_t dest_ofs_spx = (uint64_t) (int64_t) d->placement_ofs_spx;
    uint64_t dest_dim_spx = d->prehalving_spx;
    uint32_t prehalving_px = d->prehalving_px;
    unsigned n_halvings = d->n_halvings;
    uint32_t src_dim_px = (uint32_t) spx_to_px ((int64_t) src_dim_spx);
    uint64_t first_ofs [3], sample_step;
    int64_t clip_first = (int64_t) d->clip_before_px << n_halvings;
    int64_t clip_last = clip_first + ((int64_t) d->placement_size_px << n_halvings);

    dest_ofs_spx %= SPX;

    if (src_dim_spx > dest_dim_spx)
    {
        sample_step = (src_dim_spx * BILIN_MULTIPLIER) / dest_dim_spx;
        first_ofs [0] = (sample_step - BILIN_MULTIPLIER) / 2;
        first_ofs [1] = ((sample_step - BILIN_MULTIPLIER) / 2)
            + ((sample_step * (SPX - dest_ofs_spx) * (1u << n_halvings)) / SPX);
    }
    else
    {
        sample_step = ((src_dim_spx - SPX) * BILIN_MULTIPLIER)
            / (dest_dim_spx > SPX ? (dest_dim_spx - SPX) : 1);
        first_ofs [0] = 0;
        first_ofs [1] = (sample_step * (SPX - dest_ofs_spx) * (1u << n_halvings)) / SPX;
    }

    first_ofs [2] = (((src_dim_spx * BILIN_MULTIPLIER * 2) / SPX) + sample_step - BILIN_MULTIPLIER) / 2
        - sample_step * (1u << n_halvings);

    linear_range (out, 0, 1 << n_halvings, first_ofs [0], sample_step, (int) src_dim_px, clip_first, clip_last);

    if (prehalving_px > (1u << n_halvings))
    {
        linear_range (out, 1 << n_halvings, (int64_t) prehalving_px - (1 << n_halvings),
                      first_ofs [1], sample_step, (int) src_dim_px, clip_first, clip_last);
        linear_range (out, (int64_t) prehalving_px - (1 << n_halvings), prehalving_px,
                      first_ofs [2], sample_step, (int) src_dim_px, clip_first, clip_last);
    }
}

static void
precalc_boxes (std::vector<uint32_t> &out, uint32_t *span_step, uint32_t *span_mul, const DimSetup *d)
{
    uint32_t src_dim_spx = d->src_size_spx;
    uint32_t dest_dim = d->prehalving_px;
    uint32_t dest_ofs_spx = (uint32_t) d->placement_ofs_spx % SPX;
    uint32_t dest_dim_spx = (uint32_t) d->placement_size_spx;

    if (dest_dim_spx < 256)
        dest_dim_spx = 256;

    uint64_t frac_step = ((uint64_t) src_dim_spx * BIG_MUL) / (uint64_t) dest_dim_spx;
    uint64_t stride = frac_step / BIG_MUL;
    uint64_t f = (frac_step / SMALL_MUL) % SMALL_MUL;
    uint64_t a = BOXES_MULTIPLIER * 255;
    uint64_t b = (stride * 255) + ((f * 255) / 256);

    *span_step = (uint32_t) (frac_step / SMALL_MUL);
    *span_mul = (uint32_t) (a / (b + 1));

    uint64_t last_ofs = ((uint64_t) src_dim_spx * SMALL_MUL - frac_step) / SMALL_MUL;
    int64_t clip_first = d->clip_before_px;
    int64_t clip_last = clip_first + d->placement_size_px;

    if (clip_first == 0 && clip_last > 0)
        out.push_back (0);

    int64_t main_first = std::max<int64_t> (clip_first, 1);
    int64_t main_last = std::min<int64_t> (clip_last, (int64_t) dest_dim - 1);
    uint64_t frac = ((frac_step * (SPX - dest_ofs_spx)) / SPX) + (uint64_t) (main_first - 1) * frac_step;

    for (int64_t i = main_first; i < main_last; i++)
    {
        out.push_back ((uint32_t) std::min (frac / SMALL_MUL, last_ofs));
        frac += frac_step;
    }

    if (dest_dim > 1 && (int64_t) dest_dim - 1 >= clip_first && (int64_t) dest_dim - 1 < clip_last)
        out.push_back ((uint32_t) last_ofs);
}

static void
precalc_nearest (std::vector<uint32_t> &out, const DimSetup *d)
{
    uint32_t src_dim_px = d->src_size_px;
    int64_t clip_first = d->clip_before_px;
    int64_t clip_last = clip_first + d->placement_size_px;

    for (int64_t i = clip_first; i < clip_last; i++)
    {
        uint64_t ofs_px = (((uint64_t) i * 2 + 1) * d->src_size_spx) / ((uint64_t) d->placement_size_spx * 2);
        out.push_back ((uint32_t) (uint16_t) std::min<uint64_t> (ofs_px, (uint64_t) src_dim_px - 1));
    }
}

static void
finish_dim (ScaleDim *out, const DimSetup *d)
{
    out->src_size_px = d->src_size_px;
    out->dest_size_px = d->dest_size_px;
    out->filter = d->filter;
    out->n_halvings = (int) d->n_halvings;
    out->placement_ofs_px = d->placement_ofs_px;
    out->placement_size_px = d->placement_size_px;
    out->clip_before_px = d->clip_before_px;
    out->first_opacity = d->first_opacity;
    out->last_opacity = d->last_opacity;
    out->span_step = out->span_mul = 0;
    out->precalc.clear ();

    if (d->placement_size_px <= 0)
        return;

    switch (d->filter)
    {
        case SCALE_FILTER_ONE:
        case SCALE_FILTER_COPY:
            break;
        case SCALE_FILTER_NEAREST:
            precalc_nearest (out->precalc, d);
            break;
        case SCALE_FILTER_BOX:
            precalc_boxes (out->precalc, &out->span_step, &out->span_mul, d);
            break;
        default:
            precalc_bilinear (out->precalc, d);
            break;
    }
}

static int32_t
round_placement_ofs_spx (int32_t ofs_spx)
{
    int64_t n = (int64_t) ofs_spx + SPX / 2;
    int64_t px = (n >= 0) ? (n / SPX) : -((-n + SPX - 1) / SPX);
    return (int32_t) std::min<int64_t> (px * SPX, (int64_t) (INT32_MAX / SPX) * SPX);
}

static int32_t
round_placement_size_spx (int32_t size_spx)
{
    if (size_spx == 0)
        return 0;
    int64_t px = std::max<int64_t> (((int64_t) size_spx + SPX / 2) / SPX, 1);
    return (int32_t) std::min<int64_t> (px * SPX, (int64_t) (INT32_MAX / SPX) * SPX);
}

/* Arguments in whole pixels: Chafa always places on whole pixels
 * (chafa/internal/chafa-pixops.c:704-785). */
bool
scale_plan_build (ScalePlan *plan, int src_w, int src_h, int dest_w, int dest_h,
                  int place_x, int place_y, int place_w, int place_h, bool nearest)
{
    plan->valid = false;

    if (src_w < 1 || src_w > 65535 || src_h < 1 || src_h > 65535
        || dest_w < 1 || dest_w > 65535 || dest_h < 1 || dest_h > 65535)
        return false;

    int32_t px = place_x * SPX, py = place_y * SPX, pw = place_w * SPX, ph = place_h * SPX;

    if (pw <= 0 || ph <= 0)
    {
        px = py = pw = ph = 0;
    }
    else if (nearest)
    {
        px = round_placement_ofs_spx (px);
        py = round_placement_ofs_spx (py);
        pw = round_placement_size_spx (pw);
        ph = round_placement_size_spx (ph);
    }

    DimSetup h, v;
    setup_dim (&h, (uint32_t) src_w * SPX, (uint32_t) dest_w * SPX, px, pw, nearest);
    setup_dim (&v, (uint32_t) src_h * SPX, (uint32_t) dest_h * SPX, py, ph, nearest);

    if (h.placement_size_px == 0 || v.placement_size_px == 0)
    {
        setup_dim (&h, (uint32_t) src_w * SPX, (uint32_t) dest_w * SPX, 0, 0, nearest);
        setup_dim (&v, (uint32_t) src_h * SPX, (uint32_t) dest_h * SPX, 0, 0, nearest);
    }

    finish_dim (&plan->h, &h);
    finish_dim (&plan->v, &v);

    /* sRGB linearisation is switched off when the source is > 8191x the output in
     * either dimension (smolscale.c:1320-1336) */
    plan->linear = !((uint64_t) h.src_size_spx > std::max<uint64_t> ((uint64_t) h.placement_size_spx, SPX) * 8191
                     || (uint64_t) v.src_size_spx > std::max<uint64_t> ((uint64_t) v.placement_size_spx, SPX) * 8191);
    plan->valid = true;
    return true;
}

/* Source rows [*lo, *hi] that destination rows [dest_first, dest_first + dest_n) read through
 * the vertical filter: lets a row-band owner upload only its slice of the source. */
void
scale_plan_source_rows (const ScalePlan *plan, int dest_first, int dest_n, int *lo_out, int *hi_out)
{
    const ScaleDim &v = plan->v;
    int src_h = (int) v.src_size_px;
    int lo = src_h - 1, hi = 0;
    int yr0 = std::max (dest_first - v.placement_ofs_px, 0);
    int yr1 = std::min (dest_first + dest_n - v.placement_ofs_px, v.placement_size_px);

    for (int yr = yr0; yr < yr1; yr++)
    {
        int a = 0, b = 0;
        switch (v.filter)
        {
            case SCALE_FILTER_COPY:
                a = b = yr + v.clip_before_px;
                break;
            case SCALE_FILTER_ONE:
                a = b = 0;
                break;
            case SCALE_FILTER_NEAREST:
                a = b = (int) (v.precalc [yr] & 0xffff);
                break;
            case SCALE_FILTER_BOX:
            {
                uint32_t o0 = v.precalc [yr], o1 = o0 + v.span_step;
                a = (int) (o0 / 256);
                b = (int) (o1 / 256);
                break;
            }
            default:
            {
                int nh = v.n_halvings;
                a = src_h;
                b = 0;
                for (int i = 0; i < (1 << nh); i++)
                {
                    int ofs = (int) (v.precalc [(yr << nh) + i] & 0xffff);
                    a = std::min (a, ofs);
                    b = std::max (b, ofs + 1);
                }
                break;
            }
        }
        lo = std::min (lo, a);
        hi = std::max (hi, b);
    }
    if (yr0 >= yr1)
        lo = hi = 0;
    *lo_out = std::max (std::min (lo, src_h - 1), 0);
    *hi_out = std::max (std::min (hi, src_h - 1), *lo_out);
}

/* ------------------------------------------------------------------------ */
/* Tiling of the row-staged scaler (scale.cu, scale_tma_kernel): a block produces @rows_per_block
 * destination rows of @tile_w columns from a source tile that TMA brings into shared memory as
 * boxes of @box_w pixels x 8 rows.  Chosen here because the source extent of a destination tile
 * comes from the filter tables, which live on the host. */

static int
dim_src_first (const ScaleDim &d, int r)
{
    if (d.filter == SCALE_FILTER_COPY)
        return r + d.clip_before_px;
    return (int) (d.precalc [(size_t) r << d.n_halvings] & 0xffff);
}

static int
dim_src_last (const ScaleDim &d, int r)
{
    if (d.filter == SCALE_FILTER_COPY)
        return r + d.clip_before_px;
    return (int) (d.precalc [((size_t) r << d.n_halvings) + ((size_t) 1 << d.n_halvings) - 1] & 0xffff) + 1;
}

bool
scale_plan_tiles (const ScalePlan *plan, int dest_w, int first_row, int n_rows, int n_sms, ScaleTiles *out)
{
    const ScaleDim &h = plan->h, &v = plan->v;
    auto staged = [] (const ScaleDim &d)
    {
        return d.filter == SCALE_FILTER_COPY
            || (d.filter >= SCALE_FILTER_BILINEAR_0H && d.filter <= SCALE_FILTER_BILINEAR_3H);
    };
    memset (out, 0, sizeof (*out));
    if (!staged (h) || !staged (v) || n_rows <= 0 || dest_w <= 0)
        return false;
    if (h.filter != SCALE_FILTER_COPY && h.n_halvings > 2)
        return false;           /* eight samples per column: left to the gathering kernel */

    /* widest tile whose source columns fit a TMA box (256 elements) */
    int tile_w = 0, box_w = 0;
    for (int tw : { 128, 64, 32 })
    {
        int span_max = 0;
        for (int x0 = 0; x0 < dest_w; x0 += tw)
        {
            int a = std::max (x0 - h.placement_ofs_px, 0);
            int b = std::min (x0 + tw - h.placement_ofs_px, h.placement_size_px);
            if (a < b)
                span_max = std::max (span_max, dim_src_last (h, b - 1) - (dim_src_first (h, a) & ~3) + 1);   /* boxes start on 16 bytes */
        }
        int bw = std::max ((span_max + 3) & ~3, 4);
        if (bw <= 256)
        {
            tile_w = tw;
            box_w = bw;
            break;
        }
    }
    if (!tile_w)
        return false;

    /* rows per block: about eight blocks per SM, and a source tile of at most 64 rows */
    const int col_tiles = (dest_w + tile_w - 1) / tile_w;
    static const int per_sm = [] { const char *e = getenv ("CHAFA_B200_SCALE_BLOCKS_PER_SM"); return e && atoi (e) > 0 ? atoi (e) : 8; } ();
    int rows = std::max (1, (int) (((long long) n_rows * col_tiles + (long long) n_sms * per_sm - 1) / ((long long) n_sms * per_sm)));
    rows = std::min (rows, 64);
    int src_rows_max;
    for (;;)
    {
        src_rows_max = 1;
        for (int r0 = first_row; r0 < first_row + n_rows; r0 += rows)
        {
            int a = std::max (r0 - v.placement_ofs_px, 0);
            int b = std::min (std::min (r0 + rows, first_row + n_rows) - v.placement_ofs_px, v.placement_size_px);
            if (a < b)
                src_rows_max = std::max (src_rows_max, dim_src_last (v, b - 1) - dim_src_first (v, a) + 1);
        }
        if (src_rows_max <= SCALE_TILE_CHUNK_ROWS * SCALE_TILE_MAX_CHUNKS || rows == 1)
            break;
        rows = std::max (1, rows * 3 / 4);
    }
    if (src_rows_max > SCALE_TILE_CHUNK_ROWS * SCALE_TILE_MAX_CHUNKS)
        return false;

    out->tile_w = tile_w;
    out->box_w = box_w;
    out->rows_per_block = rows;
    out->n_chunks = (src_rows_max + SCALE_TILE_CHUNK_ROWS - 1) / SCALE_TILE_CHUNK_ROWS;
    return true;
}
