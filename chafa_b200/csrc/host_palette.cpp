/* host_palette.cpp -- dynamic palette construction for sixel mode, host side.
 *
 * The per-pixel work (sampling every n-th pixel into colour-cube bins, and later the
 * per-pixel pen lookups) runs on the device.  What is left is a serial algorithm over
 * at most 32768 bins that produces at most 255 colours, plus a 256-entry table build:
 * milliseconds of float work whose last bits decide integer results, so it stays on the
 * host and is compiled with the reference's own float flags (-O2 -ffast-math, see the
 * Makefile) and the same operation order.
 *
 * Restates (reference file:line):
 *   pairwise-nearest-neighbour merge      chafa/internal/chafa-palette.c:388-727
 *   duplicate removal, transparent pen    chafa/internal/chafa-palette.c:762-827
 *   quality table                         chafa/internal/chafa-palette.c:55-94, 912-953
 *   power-iteration PCA, 2 components     chafa/internal/chafa-pca.c:29-162
 *   projected, sorted pen table           chafa/internal/chafa-color-table.c:92-191, 288-313 */

#include "internal.h"
#include "device.h"

#include <cfloat>
#include <cmath>

/* ------------------------------------------------------------------------ */
/* Quality parameters                                                         */

struct QualityRow
{
    float min_quality;
    int n_samples;
    int bits_per_ch;
};

static const QualityRow k_quality [] =
{
    { .0f, 1 << 14, 3 }, { .1f, 1 << 15, 3 }, { .2f, 1 << 16, 4 }, { .3f, 1 << 17, 4 }, { .45f, 1 << 18, 4 },
    { .6f, 1 << 19, 5 }, { .7f, 1 << 20, 5 }, { .8f, 1 << 21, 5 }, { .95f, 1 << 26, 5 }
};

void
palette_quality_params (float quality, size_t n_pixels, size_t *step_out, int *bits_per_ch_out)
{
    if (quality < 0.0f)
        quality = 0.5f;
    quality = quality < 0.0f ? 0.0f : quality > 1.0f ? 1.0f : quality;

    const QualityRow *row = &k_quality [0];
    for (size_t i = 0; i < sizeof (k_quality) / sizeof (k_quality [0]) && quality >= k_quality [i].min_quality; i++)
        row = &k_quality [i];

    size_t step = n_pixels / (size_t) row->n_samples;
    if (step <= 4)
        step = 1;
    *step_out = step;
    *bits_per_ch_out = row->bits_per_ch;
}

/* ------------------------------------------------------------------------ */
/* Pairwise nearest neighbour                                                 */

struct Bin
{
    float accum [3];
    float err;
    float count;
    uint16_t nearest, next, prev, tm, mtm;
};

static const float k_luma [3] = { .299f, .587f, .114f };
static const float k_coeffs [3] [3] =
{
    { .299f, .587f, .114f },
    { -0.14713f, -0.28886f, 0.436f },
    { 0.615f, -0.51499f, -0.10001f }
};
#define PNN_RATIO .5f

static float
quantise_count (float count, int mode)
{
    if (mode > 0)
    {
        if (mode > 1)
            return (float) (int) sqrtf (count);
        return sqrtf (count);
    }
    if (mode < 0)
        return (float) (int) cbrtf (count);
    return count;
}

static void
nearest_bin (Bin *bins, uint16_t index, const float *weights)
{
    Bin *b1 = &bins [index];
    float err = FLT_MAX;
    uint16_t nearest = 0;

    for (uint16_t i = b1->next; i; i = bins [i].next)
    {
        float n2 = bins [i].count;
        float nerr2 = (b1->count * n2) / (b1->count + n2);
        float nerr = .0f;
        float t [3];

        if (nerr2 >= err)
            continue;

        for (int c = 0; c < 3; c++) t [c] = bins [i].accum [c] - b1->accum [c];
        for (int c = 0; c < 3; c++) t [c] = t [c] * t [c];
        for (int c = 0; c < 3; c++) t [c] = t [c] * weights [c];
        {
            float s = nerr2 * (1 - PNN_RATIO);
            for (int c = 0; c < 3; c++) t [c] = t [c] * s;
        }
        nerr += t [0] + t [1] + t [2];
        if (nerr >= err)
            continue;

        int j;
        for (j = 0; j < 3; j++)
        {
            for (int c = 0; c < 3; c++) t [c] = bins [i].accum [c] - b1->accum [c];
            for (int c = 0; c < 3; c++) t [c] = t [c] * k_coeffs [j] [c];
            for (int c = 0; c < 3; c++) t [c] = t [c] * t [c];
            {
                float s = nerr2 * PNN_RATIO;
                for (int c = 0; c < 3; c++) t [c] = t [c] * s;
            }
            nerr += t [0] + t [1] + t [2];
            if (nerr >= err)
                break;
        }

        if (nerr < err)
        {
            err = nerr;
            nearest = i;
        }
    }

    b1->err = err;
    b1->nearest = nearest;
}

/* @sums: per colour-cube bin { sum ch0, sum ch1, sum ch2, count } as exact integers
 * (the device accumulates them with integer atomics; the reference adds floats in
 * sample order, which is the same value as long as a sum stays below 2^24).
 * Writes up to @n_cols colours (ch0 | ch1 << 8 | ch2 << 16 | 0xff << 24) and returns
 * their number. */
int
pnn_palette_from_bins (const uint32_t *sums, int bits_per_ch, int n_cols, uint32_t *colors_out,
                       const uint32_t *float_bin_ids, const float *float_sums, int n_float_bins)
{
    float weights [3] = { k_luma [0], k_luma [1], k_luma [2] };
    int max_bins = 1 << (bits_per_ch * 3);
    std::vector<Bin> bins_v ((size_t) max_bins + 1);
    std::vector<uint16_t> heap_v ((size_t) max_bins + 1, 0);
    Bin *bins = bins_v.data ();
    uint16_t *heap = heap_v.data ();
    int n_bins = 0;
    int quan_rt = 1;

    memset (bins, 0, sizeof (Bin) * ((size_t) max_bins + 1));

    /* Active bins, averaged, packed to the front */
    for (int i = 0; i < max_bins; i++)
    {
        const uint32_t *s = sums + (size_t) i * 4;
        if (s [3] == 0)
            continue;

        Bin tb;
        memset (&tb, 0, sizeof (tb));
        float f [4] = { (float) s [0], (float) s [1], (float) s [2], (float) s [3] };
        /* sums that passed 2^24 were re-accumulated in float, in sample order, on the device */
        for (int k = 0; k < n_float_bins; k++)
            if (float_bin_ids [k] == (uint32_t) i)
                memcpy (f, float_sums + (size_t) k * 4, sizeof (f));
        tb.count = f [3];
        float inv = 1.0f / tb.count;
        tb.accum [0] = f [0] * inv;
        tb.accum [1] = f [1] * inv;
        tb.accum [2] = f [2] * inv;
        bins [n_bins++] = tb;
    }
    if (n_bins == 0)
        return 0;

    if (n_cols < 16)
        quan_rt = -1;

    float weight = std::min (0.9f, n_cols * 1.0f / n_bins);
    if (weight < .03f && weights [1] < 1.0f && weights [1] >= k_coeffs [0] [1])
    {
        weights [0] = weights [1] = weights [2] = 1.0f;
        if (n_cols >= 64)
            quan_rt = 0;
    }
    if (quan_rt > 0 && n_cols < 64)
        quan_rt = 2;

    int j;
    for (j = 0; j < n_bins - 1; j++)
    {
        bins [j].next = (uint16_t) (j + 1);
        bins [j + 1].prev = (uint16_t) j;
        bins [j].count = quantise_count (bins [j].count, quan_rt);
    }
    bins [j].count = quantise_count (bins [j].count, quan_rt);

    /* Heap keyed by each bin's distance to its nearest successor */
    for (int i = 0; i < n_bins; i++)
    {
        uint16_t h, l, l2;

        nearest_bin (bins, (uint16_t) i, weights);
        float err = bins [i].err;

        heap [0]++;
        for (l = heap [0]; l > 1; l = l2)
        {
            l2 = l >> 1;
            h = heap [l2];
            if (bins [h].err <= err)
                break;
            heap [l] = h;
        }
        heap [l] = (uint16_t) i;
    }

    /* Merge until n_cols bins remain */
    int extbins = n_bins - n_cols;
    for (int i = 0; i < extbins; )
    {
        Bin *tb;
        uint16_t b1;

        for (;;)
        {
            uint16_t h, l, l2;

            b1 = heap [1];
            tb = &bins [b1];

            if (tb->tm >= tb->mtm && bins [tb->nearest].mtm <= tb->tm)
                break;

            if (tb->mtm == 0xffff)
            {
                b1 = heap [1] = heap [heap [0]];
                heap [0]--;
            }
            else
            {
                nearest_bin (bins, b1, weights);
                tb->tm = (uint16_t) i;
            }

            float err = bins [b1].err;
            for (l = 1, l2 = 2; l2 <= heap [0]; l = l2, l2 = (uint16_t) (l * 2))
            {
                if (l2 < heap [0] && bins [heap [l2]].err > bins [heap [l2 + 1]].err)
                    l2++;
                h = heap [l2];
                if (err <= bins [h].err)
                    break;
                heap [l] = h;
            }
            heap [l] = b1;
        }

        tb = &bins [b1];
        Bin *nb = &bins [tb->nearest];
        float n1 = tb->count, n2 = nb->count;
        float d = 1.0f / (n1 + n2);
        float tv [3];

        for (int c = 0; c < 3; c++) tb->accum [c] = tb->accum [c] * n1;
        for (int c = 0; c < 3; c++) tv [c] = nb->accum [c] * n2;
        for (int c = 0; c < 3; c++) tb->accum [c] = tb->accum [c] + tv [c];
        for (int c = 0; c < 3; c++) tb->accum [c] = rintf (tb->accum [c]);
        for (int c = 0; c < 3; c++) tb->accum [c] = tb->accum [c] * d;

        tb->count += n2;
        tb->mtm = (uint16_t) ++i;

        bins [nb->prev].next = nb->next;
        bins [nb->next].prev = nb->prev;
        nb->mtm = 0xffff;
    }

    int k = 0;
    for (int i = 0; ; k++)
    {
        uint32_t c0 = (uint8_t) bins [i].accum [0], c1 = (uint8_t) bins [i].accum [1], c2 = (uint8_t) bins [i].accum [2];
        colors_out [k] = c0 | (c1 << 8) | (c2 << 16) | 0xff000000u;
        i = bins [i].next;
        if (!i)
            break;
    }
    return k + 1;
}

/* Drop colours that coincide on the sixel 0..100 scale, then place the transparent pen
 * (chafa-palette.c:762-827).  @colors holds 259 entries laid out like the reference's
 * palette (pens, then the three special colours); returns the new colour count. */
int
palette_clean_up (uint32_t *colors, int n_colors, int transparent_index)
{
    int best_diff = INT32_MAX, best_pair = 1;
    int i, j;

    for (i = 1, j = 1; i < n_colors; i++)
    {
        uint32_t a = colors [j - 1], b = colors [i];
        int diff = 0;

        for (int c = 0; c < 3; c++)
        {
            int t = (int) (((a >> (8 * c)) & 0xff) * 100) / 256 - (int) (((b >> (8 * c)) & 0xff) * 100) / 256;
            diff += t * t;
        }

        if (diff == 0)
            continue;
        if (diff < best_diff)
        {
            best_pair = j - 1;
            best_diff = diff;
        }
        colors [j++] = colors [i];
    }
    n_colors = j;

    if (transparent_index < 256)
    {
        if (n_colors < 256)
        {
            colors [n_colors] = colors [transparent_index];
            n_colors++;
        }
        else
        {
            colors [best_pair] = colors [transparent_index];
        }
    }
    return n_colors;
}

/* ------------------------------------------------------------------------ */
/* Pen table: PCA projection + sort                                           */

struct V3 { float v [3]; };

static inline float
dot3 (const V3 &a, const V3 &b)
{
    return a.v [0] * b.v [0] + a.v [1] * b.v [1] + a.v [2] * b.v [2];
}

static inline float
magnitude3 (const V3 &a)
{
    return sqrtf (a.v [0] * a.v [0] + a.v [1] * a.v [1] + a.v [2] * a.v [2]);
}

static inline void
normalize3 (V3 &out, const V3 &in)
{
    float mag = magnitude3 (in);
    if (mag == .0f)
    {
        out.v [0] = out.v [1] = out.v [2] = .0f;
        return;
    }
    float m = 1.0f / mag;
    out.v [0] = in.v [0] * m;
    out.v [1] = in.v [1] * m;
    out.v [2] = in.v [2] * m;
}

static float
power_iteration (const V3 *vecs, int n, V3 *eigenvector_out)
{
    V3 r = { { .11f, .23f, .51f } };
    float eigenvalue = 0.0f;

    normalize3 (r, r);

    for (int it = 0; it < 1000; it++)
    {
        V3 s = { { 0.0f, 0.0f, 0.0f } };
        V3 t;

        for (int i = 0; i < n; i++)
        {
            float u = dot3 (vecs [i], r);
            t.v [0] = vecs [i].v [0] * u;
            t.v [1] = vecs [i].v [1] * u;
            t.v [2] = vecs [i].v [2] * u;
            s.v [0] = s.v [0] + t.v [0];
            s.v [1] = s.v [1] + t.v [1];
            s.v [2] = s.v [2] + t.v [2];
        }

        eigenvalue = dot3 (r, s);

        t.v [0] = r.v [0] * eigenvalue;
        t.v [1] = r.v [1] * eigenvalue;
        t.v [2] = r.v [2] * eigenvalue;
        t.v [0] = t.v [0] - s.v [0];
        t.v [1] = t.v [1] - s.v [1];
        t.v [2] = t.v [2] - s.v [2];
        float err = magnitude3 (t);

        r = s;
        normalize3 (r, r);

        if (err < 0.0001f)
            break;
    }

    *eigenvector_out = r;
    return eigenvalue;
}

static void
fixed_from_float (int32_t *out, const V3 &in)
{
    out [0] = (int32_t) lrintf (in.v [0] * 32.0f);
    out [1] = (int32_t) lrintf (in.v [1] * 32.0f);
    out [2] = (int32_t) lrintf (in.v [2] * 32.0f);
}

static int
project_fixed (const int32_t *v, const int32_t *axis, int32_t mul)
{
    int64_t d = (int64_t) v [0] * axis [0] + (int64_t) v [1] * axis [1] + (int64_t) v [2] * axis [2];
    return (int) ((d * mul) / ((1 << 14) / 32));
}

static int
compare_entries (const void *a, const void *b)
{
    return ((const DevColorTableEntry *) a)->v [0] - ((const DevColorTableEntry *) b)->v [0];
}

/* @table->pens must be filled (0xffffffff = unused) */
void
color_table_build (DevColorTable *table)
{
    V3 v [256];
    int n = 0, n_entries = 0;

    memset (table->entries, 0, sizeof (table->entries));
    for (int i = 0; i < 256; i++)
    {
        if (table->pens [i] == 0xffffffffu)
            continue;
        table->entries [n_entries++].pen = i;
    }
    table->n_entries = n_entries;

    for (int i = 0; i < 256; i++)
    {
        uint32_t col = table->pens [i];
        if ((col & 0xff000000u) == 0xff000000u)
            continue;
        v [n].v [0] = (float) (col & 0xff) * 32.0f;
        v [n].v [1] = (float) ((col >> 8) & 0xff) * 32.0f;
        v [n].v [2] = (float) ((col >> 16) & 0xff) * 32.0f;
        n++;
    }

    /* average, recentre, two components with deflation */
    V3 average = { { 0.0f, 0.0f, 0.0f } };
    for (int i = 0; i < n; i++)
    {
        average.v [0] = average.v [0] + v [i].v [0];
        average.v [1] = average.v [1] + v [i].v [1];
        average.v [2] = average.v [2] + v [i].v [2];
    }
    {
        float s = 1.0f / (float) n;
        average.v [0] = average.v [0] * s;
        average.v [1] = average.v [1] * s;
        average.v [2] = average.v [2] * s;
    }
    {
        V3 t = { { 0.0f, 0.0f, 0.0f } };
        t.v [0] = t.v [0] - average.v [0];
        t.v [1] = t.v [1] - average.v [1];
        t.v [2] = t.v [2] - average.v [2];
        for (int i = 0; i < n; i++)
        {
            v [i].v [0] = v [i].v [0] + t.v [0];
            v [i].v [1] = v [i].v [1] + t.v [1];
            v [i].v [2] = v [i].v [2] + t.v [2];
        }
    }

    V3 eig [2];
    power_iteration (v, n, &eig [0]);
    for (int i = 0; i < n; i++)
    {
        float score = dot3 (v [i], eig [0]);
        v [i].v [0] = v [i].v [0] - eig [0].v [0] * score;
        v [i].v [1] = v [i].v [1] - eig [0].v [1] * score;
        v [i].v [2] = v [i].v [2] - eig [0].v [2] * score;
    }
    power_iteration (v, n, &eig [1]);

    fixed_from_float (table->eigenvectors [0], eig [0]);
    fixed_from_float (table->eigenvectors [1], eig [1]);
    fixed_from_float (table->average, average);

    for (int k = 0; k < 2; k++)
    {
        int32_t *e = table->eigenvectors [k];
        int m = e [0] * e [0] + e [1] * e [1] + e [2] * e [2];
        m = std::max (m, 1);
        table->eigen_mul [k] = (1 << 14) / m;
    }

    for (int i = 0; i < n_entries; i++)
    {
        DevColorTableEntry *entry = &table->entries [i];
        uint32_t col = table->pens [entry->pen];
        int32_t p [3] = { (int32_t) (col & 0xff) * 32 - table->average [0],
                          (int32_t) ((col >> 8) & 0xff) * 32 - table->average [1],
                          (int32_t) ((col >> 16) & 0xff) * 32 - table->average [2] };
        entry->v [0] = project_fixed (p, table->eigenvectors [0], table->eigen_mul [0]);
        entry->v [1] = project_fixed (p, table->eigenvectors [1], table->eigen_mul [1]);
    }

    /* The order among equal keys is whatever the C library's qsort leaves; the
     * reference calls the same function on the same 12-byte records. */
    qsort (table->entries, (size_t) n_entries, sizeof (DevColorTableEntry), compare_entries);
}
