/* host_sixel.cpp -- sixel pixel mode: orchestration of kernels K1, KP1-KP6 (palette) and
 * K6b-K6d.
 *
 * Replaces (reference file:line):
 *   chafa/internal/chafa-sixel-renderer.c:55-126, 422-522   renderer, header, palette text
 *   chafa/internal/chafa-indexed-image.c:336-491             scale -> palette -> quantise
 *   chafa/internal/chafa-palette.c:912-953                   chafa_palette_generate
 *
 * Per draw: the source is scaled onto the canvas pixmap (K1), the palette and its pen table
 * are generated on the device (palette_build.cu; the draw never waits for them: the colours
 * come back to the host asynchronously and are only needed when the text's palette header is
 * written), and the pixels are quantised (K6b) or error-diffused (K6c).  Per print: the band encoder
 * (K6d) measures, lays out and writes the sixel text; the host wraps it in the begin /
 * raster / palette header and the end sequence. */

#include "internal.h"
#include "device.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

struct ChafaCanvas;

/* host_canvas.cpp */
cudaStream_t canvas_stream (ChafaCanvas *c);
const ChafaCanvasConfig *canvas_config (const ChafaCanvas *c);
void **canvas_sixel_slot (ChafaCanvas *c);
int canvas_width_px (const ChafaCanvas *c);
int canvas_height_px (const ChafaCanvas *c);
void canvas_band_px (const ChafaCanvas *c, int *first_out, int *end_out);
const ChafaPlacement *canvas_placement (const ChafaCanvas *c);
uint32_t *canvas_pixels (ChafaCanvas *c);
const DevPalette *canvas_host_palette (const ChafaCanvas *c);
const DevPalette *canvas_device_palette (const ChafaCanvas *c);
void canvas_dither_params (const ChafaCanvas *c, int *mode, float *intensity, int *gws, int *ghs, int *tss, int *tsm,
                           const int32_t **d_texture);
bool canvas_run_scaler (ChafaCanvas *c, ChafaPixelType type, const void *d_src, int src_w, int src_h, int rowstride,
                        int px, int py, int pw, int ph, bool nearest, int first_row, int n_rows);
bool canvas_cuda_ok (int err, const char *what);

#define CK(call) canvas_cuda_ok ((int) (call), #call)

struct SixelState
{
    int width = 0, height = 0, image_height = 0;
    bool have_image = false;

    /* palette of the last draw, laid out like the reference's ChafaPalette.colors */
    uint32_t colors_rgb [PAL_INDEX_MAX];
    int n_colors = 0, first_color = 0, transparent_index = 255;
    bool dynamic = true;

    void *d_indexed = nullptr, *d_work = nullptr, *d_table = nullptr, *d_pal_colors = nullptr, *d_scratch = nullptr,
         *d_fixed_pal = nullptr, *d_lens = nullptr, *d_out = nullptr, *d_lut = nullptr, *d_exact = nullptr,
         *d_cache_scratch = nullptr, *d_row_batch = nullptr, *d_base_colors = nullptr, *d_result = nullptr;
    size_t cap_exact = 0, cap_cache_scratch = 0, cap_row_batch = 0, cap_work = 0;
    size_t cap_indexed = 0, cap_scratch = 0, cap_lens = 0, cap_out = 0;
    DevPaletteResult *h_result = nullptr;   /* pinned; filled asynchronously by the draw */
    cudaEvent_t palette_ready = nullptr;
    cudaEvent_t chain_done = nullptr;       /* blocking-sync: a thread waiting for a diffusion chain sleeps */
    bool palette_pending = false;
    void *h_text = nullptr;     /* pinned */
    size_t cap_h_text = 0;

    /* Row bands (multi-GPU stills).  own: the pixel rows this canvas quantises for print and
     * samples for the palette; calc_first: where the rows it has to compute begin (the lookup
     * cache of the reference runs per worker batch, so the pens of a band depend on the rows of
     * its first batch above it).  phase: 0 idle, 1 scaled + sampled, 2 float sums done. */
    int own_first = 0, own_end = 0, calc_first = 0, calc_end = 0;
    bool banded = false, reduce = false;
    int phase = 0;
    void *ext_head = nullptr;   /* caller-owned device buffer for the exchanged head of the workspace */

    /* Batched diffusion (sixel_launch_deferred_chains): the draw prepares everything but leaves the
     * chain to a launch shared with other canvases */
    bool defer_chain = false, chain_deferred = false;
    DevSixel deferred;
    void *d_batch = nullptr;
    size_t cap_batch = 0;
    cudaEvent_t batch_done = nullptr;
    DevPaletteBuild build;
    size_t n_pixels = 0;
};

static bool
grow (void **p, size_t *cap, size_t n)
{
    if (n <= *cap)
        return true;
    if (*p)
        b200_dev_free (*p);
    *p = nullptr;
    *cap = 0;
    size_t want = b200_dev_guard_mode () ? n : n + n / 8 + 256;
    if (!CK (b200_dev_malloc (p, want)))
        return false;
    *cap = want;
    return true;
}

void
sixel_state_free (void *state)
{
    SixelState *s = (SixelState *) state;
    if (!s)
        return;
    void *dev [] = { s->d_indexed, s->d_work, s->d_table, s->d_pal_colors, s->d_scratch, s->d_fixed_pal, s->d_lens,
                     s->d_out, s->d_lut, s->d_exact, s->d_cache_scratch, s->d_row_batch, s->d_base_colors,
                     s->d_result };
    for (void *p : dev)
        b200_dev_free (p);
    if (s->h_result)
        cudaFreeHost (s->h_result);
    if (s->palette_ready)
        cudaEventDestroy (s->palette_ready);
    if (s->chain_done)
        cudaEventDestroy (s->chain_done);
    if (s->batch_done)
        cudaEventDestroy (s->batch_done);
    b200_dev_free (s->d_batch);
    if (s->h_text)
        cudaFreeHost (s->h_text);
    delete s;
}

static SixelState *
state_of (ChafaCanvas *c)
{
    void **slot = canvas_sixel_slot (c);
    if (!*slot)
    {
        SixelState *s = new SixelState ();
        if (!CK (b200_dev_malloc (&s->d_table, sizeof (DevColorTable)))
            || !CK (b200_dev_malloc (&s->d_pal_colors, PAL_INDEX_MAX * 4))
            || !CK (b200_dev_malloc (&s->d_base_colors, PAL_INDEX_MAX * 4))
            || !CK (b200_dev_malloc (&s->d_result, sizeof (DevPaletteResult)))
            || !CK (b200_dev_malloc (&s->d_fixed_pal, sizeof (DevPalette)))
            || !CK (cudaMallocHost ((void **) &s->h_result, sizeof (DevPaletteResult)))
            || !CK (cudaEventCreateWithFlags (&s->palette_ready, cudaEventDisableTiming | cudaEventBlockingSync))
            || !CK (cudaEventCreateWithFlags (&s->chain_done, cudaEventDisableTiming | cudaEventBlockingSync)))
        {
            sixel_state_free (s);
            return nullptr;
        }
        *slot = s;
    }
    return (SixelState *) *slot;
}

/* How densely the image is sampled and how finely the colour cube is cut, by quality
 * (chafa-palette.c:55-94, 912-935): the last row whose threshold the quality reaches. */
static void
sampling_for_quality (float quality, size_t n_pixels, size_t *step_out, int *bits_per_ch_out)
{
    static const struct { float from; int log2_samples; int bits_per_ch; } rows [] =
    {
        { .0f, 14, 3 }, { .1f, 15, 3 }, { .2f, 16, 4 }, { .3f, 17, 4 }, { .45f, 18, 4 },
        { .6f, 19, 5 }, { .7f, 20, 5 }, { .8f, 21, 5 }, { .95f, 26, 5 }
    };
    if (quality < 0.0f)
        quality = 0.5f;
    quality = std::min (std::max (quality, 0.0f), 1.0f);

    int row = 0;
    for (int i = 0; i < (int) (sizeof (rows) / sizeof (rows [0])); i++)
    {
        if (!(quality >= rows [i].from))
            break;
        row = i;
    }
    size_t step = n_pixels >> rows [row].log2_samples;
    *step_out = step <= 4 ? 1 : step;       /* every cache line is fetched anyway: sample densely */
    *bits_per_ch_out = rows [row].bits_per_ch;
}

/* The generation of the dynamic palette from the scaled image (chafa_palette_generate), queued in
 * three steps so that band owners of a sharded still can exchange what the steps hand over:
 * sampling into colour-cube bins (own rows), float sums of the bins beyond float's exact range (own
 * rows, continuing from the owner above), and everything else from the complete bins.  Colours,
 * pen table and the colour-space copy all stay on the device; the colours travel to pinned host
 * memory behind an event for whoever prints later. */
static bool
palette_sample (ChafaCanvas *c, SixelState *s)
{
    const ChafaCanvasConfig *cfg = canvas_config (c);
    const DevPalette *hp = canvas_host_palette (c);
    cudaStream_t st = canvas_stream (c);
    DevPaletteBuild &b = s->build;

    memset (&b, 0, sizeof (b));
    sampling_for_quality (cfg->work_factor, s->n_pixels, &b.step, &b.bits_per_ch);
    if (!grow (&s->d_work, &s->cap_work, dev_palette_build_work_bytes (b.bits_per_ch)))
        return false;

    /* Start from the fixed table with the canvas's FG/BG pens, as the reference's palette does
     * (pageable source: staged by the runtime before the call returns) */
    if (!CK (cudaMemcpyAsync (s->d_base_colors, hp->colors_rgb, PAL_INDEX_MAX * 4, cudaMemcpyHostToDevice, st)))
        return false;

    b.pixels = canvas_pixels (c);
    b.n_pixels = s->n_pixels;
    b.alpha_threshold = cfg->alpha_threshold;
    b.color_space = cfg->color_space == CHAFA_COLOR_SPACE_DIN99D ? 1 : 0;
    b.transparent_index = s->transparent_index;
    b.base_colors = (const uint32_t *) s->d_base_colors;
    b.work = s->d_work;
    b.head = s->reduce ? s->ext_head : nullptr;
    b.sample_first = s->reduce ? (size_t) s->own_first * s->width : 0;
    b.sample_end = s->reduce ? (size_t) s->own_end * s->width : s->n_pixels;
    b.table_out = (DevColorTable *) s->d_table;
    b.pal_colors_out = (uint32_t *) s->d_pal_colors;
    b.result_out = (DevPaletteResult *) s->d_result;

    return CK (dev_launch_palette_sample (&b, st));
}

static bool
palette_finish (ChafaCanvas *c, SixelState *s)
{
    cudaStream_t st = canvas_stream (c);

    if (s->phase < 2 && !CK (dev_launch_palette_float_bins (&s->build, st)))
        return false;
    s->phase = 2;
    if (!CK (dev_launch_palette_finish (&s->build, st))
        || !CK (cudaMemcpyAsync (s->h_result, s->d_result, sizeof (DevPaletteResult), cudaMemcpyDeviceToHost, st))
        || !CK (cudaEventRecord (s->palette_ready, st)))
        return false;
    s->palette_pending = true;
    s->first_color = 0;
    return true;
}

/* The generated colours, on the host: waits for the draw's palette kernels if need be */
static bool
palette_on_host (SixelState *s)
{
    if (!s->palette_pending)
        return true;
    if (!CK (cudaEventSynchronize (s->palette_ready)))
        return false;
    for (int i = 0; i < PAL_INDEX_MAX; i++)
        s->colors_rgb [i] = s->h_result->colors_rgb [i];
    s->n_colors = s->h_result->n_colors;
    s->palette_pending = false;
    return true;
}

/* Which rows a draw computes.  Without a band: all of them.  With one (chafa_b200_canvas_set_band):
 * the band must start on a sixel band (6 pixel rows) and end on one or at the bottom; the rows it
 * computes begin with the reference's worker batch its first row falls into, because the lookup
 * cache the quantiser replays is per batch.  Diffusion is one chain over the image: a band owner
 * computes everything and keeps its rows ("replicas only"). */
static bool
plan_rows (ChafaCanvas *c, SixelState *s, int dither_mode, bool dynamic)
{
    int first, end;
    canvas_band_px (c, &first, &end);
    s->own_first = first;
    s->own_end = end;
    s->banded = first > 0 || end < s->height;
    s->calc_first = 0;
    s->calc_end = s->height;
    s->reduce = false;
    if (!s->banded)
        return true;
    if (first % 6 != 0 || (end % 6 != 0 && end != s->height) || end <= first)
    {
        b200_set_error ("sixel row band [%d, %d) px does not lie on sixel bands (6 pixel rows)", first, end);
        return false;
    }
    if (dither_mode == CHAFA_DITHER_MODE_DIFFUSION)
        return true;

    std::vector<uint16_t> row_batch ((size_t) s->height);
    reference_batch_of_rows (s->height, reference_batch_count (), row_batch.data ());
    int q = first;
    while (q > 0 && row_batch [q - 1] == row_batch [first])
        q--;
    s->calc_first = q;
    s->calc_end = end;
    s->reduce = dynamic;
    return true;
}

/* The source rows a banded draw reads (host sources upload only these) */
bool
sixel_source_rows (ChafaCanvas *c, int src_w, int src_h, int *lo_out, int *hi_out)
{
    const ChafaPlacement *pl = canvas_placement (c);
    const DevPalette *hp = canvas_host_palette (c);
    SixelState *s = state_of (c);
    int dither_mode, gws, ghs, tss, tsm;
    float intensity;
    const int32_t *tex;
    if (!s)
        return false;
    s->width = canvas_width_px (c);
    s->height = canvas_height_px (c);
    canvas_dither_params (c, &dither_mode, &intensity, &gws, &ghs, &tss, &tsm, &tex);
    if (!plan_rows (c, s, dither_mode, hp->type == PAL_DYNAMIC_256) || !s->banded)
        return false;

    int px, py, pw, ph;
    ScalePlan plan;
    tuck_and_align (src_w, src_h, s->width, s->height, pl ? pl->halign : CHAFA_ALIGN_START,
                    pl ? pl->valign : CHAFA_ALIGN_START, pl ? pl->tuck : CHAFA_TUCK_STRETCH, &px, &py, &pw, &ph);
    if (!scale_plan_build (&plan, src_w, src_h, s->width, s->height, px, py, pw, ph, false))
        return false;
    scale_plan_source_rows (&plan, s->calc_first, s->calc_end - s->calc_first, lo_out, hi_out);
    return true;
}

/* Scaling and the sampling half of the palette.  Returns false on error; *reduce_out says whether
 * the bins have to be summed across band owners before sixel_draw_finish (). */
bool
sixel_draw_begin (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int src_w, int src_h, int rowstride,
                  bool *reduce_out)
{
    const ChafaPlacement *pl = canvas_placement (c);
    const DevPalette *hp = canvas_host_palette (c);
    SixelState *s = state_of (c);
    if (reduce_out)
        *reduce_out = false;
    if (!s)
        return false;

    s->width = canvas_width_px (c);
    s->height = canvas_height_px (c);
    s->image_height = (s->height + 5) / 6 * 6;
    s->have_image = false;
    s->phase = 0;
    s->n_pixels = (size_t) s->width * s->height;
    s->dynamic = hp->type == PAL_DYNAMIC_256;
    s->transparent_index = 255;
    s->palette_pending = false;

    int dither_mode, gws, ghs, tss, tsm;
    float intensity;
    const int32_t *tex;
    canvas_dither_params (c, &dither_mode, &intensity, &gws, &ghs, &tss, &tsm, &tex);
    if (!plan_rows (c, s, dither_mode, s->dynamic))
        return false;

    int px, py, pw, ph;
    tuck_and_align (src_w, src_h, s->width, s->height, pl ? pl->halign : CHAFA_ALIGN_START,
                    pl ? pl->valign : CHAFA_ALIGN_START, pl ? pl->tuck : CHAFA_TUCK_STRETCH, &px, &py, &pw, &ph);
    if (!canvas_run_scaler (c, type, d_src, src_w, src_h, rowstride, px, py, pw, ph, false, s->calc_first,
                            s->calc_end - s->calc_first))
        return false;

    if (s->dynamic && !palette_sample (c, s))
        return false;
    s->phase = 1;
    if (reduce_out)
        *reduce_out = s->reduce;
    return true;
}

/* Row bands: the float sums of the bins beyond float's exact range, continued over this canvas's
 * rows from what the owners of the rows above left in the head (zero for the first).  Called once
 * per draw, in band order across the owners, after the bins have been summed. */
bool
sixel_draw_float_bins (ChafaCanvas *c)
{
    SixelState *s = (SixelState *) *canvas_sixel_slot (c);
    if (!s || s->phase != 1 || !s->dynamic)
        return false;
    if (!CK (dev_launch_palette_float_bins (&s->build, canvas_stream (c))))
        return false;
    s->phase = 2;
    return true;
}

/* The exchanged head of the palette workspace: device address, size in bytes, and how many leading
 * 32-bit words are integer sums (the rest are the float sums) */
void *
sixel_exchange_buffer (ChafaCanvas *c, size_t *n_bytes_out, size_t *n_sum_words_out)
{
    SixelState *s = (SixelState *) *canvas_sixel_slot (c);
    if (!s || s->phase < 1 || !s->dynamic)
        return nullptr;
    if (n_bytes_out) *n_bytes_out = dev_palette_build_head_bytes (s->build.bits_per_ch);
    if (n_sum_words_out) *n_sum_words_out = dev_palette_build_reduce_words (s->build.bits_per_ch);
    return s->build.head ? s->build.head : s->build.work;
}

/* Size of that head for this canvas's quality setting and geometry */
size_t
sixel_exchange_bytes (ChafaCanvas *c)
{
    const ChafaCanvasConfig *cfg = canvas_config (c);
    size_t step;
    int bits;
    sampling_for_quality (cfg->work_factor, (size_t) canvas_width_px (c) * canvas_height_px (c), &step, &bits);
    return dev_palette_build_head_bytes (bits);
}

void
sixel_set_exchange_buffer (ChafaCanvas *c, void *d_buffer)
{
    SixelState *s = state_of (c);
    if (s)
        s->ext_head = d_buffer;
}

bool
sixel_draw_finish (ChafaCanvas *c)
{
    const ChafaCanvasConfig *cfg = canvas_config (c);
    const DevPalette *hp = canvas_host_palette (c);
    cudaStream_t st = canvas_stream (c);
    SixelState *s = (SixelState *) *canvas_sixel_slot (c);
    if (!s || s->phase < 1)
        return false;

    if (s->dynamic)
    {
        if (!palette_finish (c, s))
            return false;
    }
    else
    {
        for (int i = 0; i < PAL_INDEX_MAX; i++)
            s->colors_rgb [i] = hp->colors_rgb [i];
        s->n_colors = hp->n_colors;
        s->first_color = hp->first_color;

        DevPalette fixed = *hp;
        fixed.transparent_index = 255;
        /* fixed palettes keep their own colour-space table for the error terms */
        if (!CK (cudaMemcpyAsync (s->d_fixed_pal, &fixed, sizeof (fixed), cudaMemcpyHostToDevice, st))
            || !CK (cudaMemcpyAsync (s->d_pal_colors, hp->colors, sizeof (hp->colors), cudaMemcpyHostToDevice, st)))
            return false;
    }
    s->phase = 0;

    if (!grow (&s->d_indexed, &s->cap_indexed, (size_t) s->width * s->image_height)
        || !grow (&s->d_scratch, &s->cap_scratch, (size_t) s->width * 2 * 8 + 64))
        return false;

    /* The rows computed: [calc_first, calc_end) of the image, as an image of its own whose row 0 is
     * image row calc_first (the whole image unless this is a band) */
    const int rows = s->calc_end - s->calc_first;
    const int rows_padded = s->calc_end == s->height ? s->image_height - s->calc_first : rows;
    const size_t px_ofs = (size_t) s->calc_first * s->width;

    DevSixel q;
    memset (&q, 0, sizeof (q));
    const int32_t *tex;
    canvas_dither_params (c, &q.dither_mode, &q.intensity, &q.grain_w_shift, &q.grain_h_shift, &q.tex_size_shift,
                          &q.tex_size_mask, &tex);
    q.texture = tex;
    q.pixels = canvas_pixels (c) + px_ofs;
    q.width = s->width;
    q.height = rows;
    q.image_height = rows_padded;
    q.y0 = s->calc_first;
    q.indexed = (uint8_t *) s->d_indexed + px_ofs;
    q.color_space = cfg->color_space;
    q.alpha_threshold = cfg->alpha_threshold;
    q.dynamic = s->dynamic;
    q.table = (const DevColorTable *) s->d_table;
    q.pal_colors = (const uint32_t *) s->d_pal_colors;
    q.fixed_palette = (const DevPalette *) s->d_fixed_pal;
    q.first_color = s->first_color;
    q.transparent_index = s->transparent_index;
    q.scratch = s->d_scratch;

    if (q.dither_mode == CHAFA_DITHER_MODE_DIFFUSION && hp->type != PAL_FIXED_FGBG)
    {
        /* Diffusion is one serial chain: take everything that is not on the chain out of it.
         * Colour-space conversion of the source pixels runs as a parallel pass, and the pen
         * of every possible colour is tabulated (16 MB) so the chain never walks the table. */
        /* 16 MB aligned to 16 MB (the chain ORs the entry's offset into the address), followed by
         * the per-brick summary (32 KB) */
        const size_t lut_bytes = (size_t) 1 << 24;
        if (!s->d_lut && !CK (b200_dev_malloc (&s->d_lut, 2 * lut_bytes + 32768 + 64)))
            return false;
        uint8_t *aligned = (uint8_t *) (((uintptr_t) s->d_lut + lut_bytes - 1) & ~(uintptr_t) (lut_bytes - 1));
        q.lut = aligned;
        q.brick_pens = aligned + lut_bytes;
        static const int policy = [] { const char *e = getenv ("CHAFA_B200_FS_BRICKS"); return e ? atoi (e) : 0; } ();
        q.brick_policy = policy;    /* measurement knob: 1 = always, 2 = never; default: the chain decides */
        if (cfg->color_space == CHAFA_COLOR_SPACE_DIN99D)
        {
            DevPrep pp;
            memset (&pp, 0, sizeof (pp));
            pp.pixels = canvas_pixels (c);
            pp.width = s->width;
            pp.height = s->height;
            pp.n_rows = s->height;
            pp.to_din99d = 1;
            if (!CK (dev_launch_prep_pass2 (&pp, st)))
                return false;
            q.preconverted = 1;
        }
    }
    s->chain_deferred = false;
    if (q.dither_mode == CHAFA_DITHER_MODE_DIFFUSION && s->defer_chain && dev_sixel_chain_applies (&q))
    {
        if (!CK (dev_launch_sixel_chain_tables (&q, st)))
            return false;
        s->deferred = q;
        s->chain_deferred = true;
        return true;            /* have_image is set once the shared launch has been queued */
    }
    if (q.dither_mode == CHAFA_DITHER_MODE_DIFFUSION)
    {
        if (!CK (dev_launch_sixel_quantise (&q, st)))
            return false;
    }
    else
    {
        /* Exact pens first, then the reference's lossy lookup cache replayed on top of them
         * for the batch split the reference would use with the current thread setting */
        const size_t n_calc = (size_t) s->width * rows;
        size_t scratch_bytes = dev_sixel_cache_scratch_bytes (n_calc);
        if (!grow (&s->d_exact, &s->cap_exact, (size_t) s->width * s->image_height)
            || !grow (&s->d_cache_scratch, &s->cap_cache_scratch, scratch_bytes)
            || !grow (&s->d_row_batch, &s->cap_row_batch, (size_t) s->height * 2 + 16))
            return false;
        std::vector<uint16_t> row_batch ((size_t) s->height);
        reference_batch_of_rows (s->height, reference_batch_count (), row_batch.data ());
        if (!CK (cudaMemcpyAsync (s->d_row_batch, row_batch.data (), row_batch.size () * 2, cudaMemcpyHostToDevice, st)))
            return false;
        q.indexed = (uint8_t *) s->d_exact + px_ofs;
        if (!CK (dev_launch_sixel_quantise (&q, st))
            || !CK (dev_launch_sixel_cache_replay (&q, (const uint16_t *) s->d_row_batch,
                                                   (const uint8_t *) s->d_exact + px_ofs,
                                                   (uint8_t *) s->d_indexed + px_ofs, s->d_cache_scratch,
                                                   s->cap_cache_scratch, st)))
            return false;
        /* row_batch is pageable: the runtime has staged it before cudaMemcpyAsync returned */
    }

    s->have_image = true;
    return true;
}

void
sixel_set_defer_chain (ChafaCanvas *c, bool defer)
{
    SixelState *s = state_of (c);
    if (s)
        s->defer_chain = defer;
}

/* Queue the diffusion chains that the draws of @canvases left open, as one launch per kind of
 * chain: the launch waits for every canvas's preparation and every canvas's stream waits for the
 * launch, so that printing a canvas afterwards is ordered as if the chain had run on its own
 * stream.  Returns the number of chains queued, -1 on error. */
int
sixel_launch_deferred_chains (ChafaCanvas **canvases, int n)
{
    int n_queued = 0;

    for (int unit = 0; unit < 2; unit++)
        for (int dynamic = 0; dynamic < 2; dynamic++)
        {
            std::vector<DevSixel> images;
            std::vector<ChafaCanvas *> owners;
            int max_width = 0;
            for (int i = 0; i < n; i++)
            {
                SixelState *s = canvases [i] ? (SixelState *) *canvas_sixel_slot (canvases [i]) : nullptr;
                if (!s || !s->chain_deferred || (s->deferred.intensity == 1.0f) != (unit != 0)
                    || (s->deferred.dynamic != 0) != (dynamic != 0))
                    continue;
                images.push_back (s->deferred);
                owners.push_back (canvases [i]);
                max_width = std::max (max_width, s->deferred.width);
            }
            if (images.empty ())
                continue;

            ChafaCanvas *lead = owners [0];
            SixelState *ls = (SixelState *) *canvas_sixel_slot (lead);
            cudaStream_t lst = canvas_stream (lead);
            if (!ls->batch_done && !CK (cudaEventCreateWithFlags (&ls->batch_done, cudaEventDisableTiming)))
                return -1;
            if (!grow (&ls->d_batch, &ls->cap_batch, images.size () * sizeof (DevSixel)))
                return -1;
            for (size_t i = 1; i < owners.size (); i++)
            {
                SixelState *s = (SixelState *) *canvas_sixel_slot (owners [i]);
                if (!CK (cudaEventRecord (s->chain_done, canvas_stream (owners [i])))
                    || !CK (cudaStreamWaitEvent (lst, s->chain_done, 0)))
                    return -1;
            }
            /* (pageable source: staged by the runtime before the call returns) */
            if (!CK (cudaMemcpyAsync (ls->d_batch, images.data (), images.size () * sizeof (DevSixel),
                                      cudaMemcpyHostToDevice, lst))
                || !CK (dev_launch_sixel_chain_batch ((const DevSixel *) ls->d_batch, (int) images.size (), max_width,
                                                      unit != 0, dynamic != 0, lst))
                || !CK (cudaEventRecord (ls->batch_done, lst)))
                return -1;
            for (size_t i = 0; i < owners.size (); i++)
            {
                SixelState *s = (SixelState *) *canvas_sixel_slot (owners [i]);
                if (i > 0 && !CK (cudaStreamWaitEvent (canvas_stream (owners [i]), ls->batch_done, 0)))
                    return -1;
                s->chain_deferred = false;
                s->have_image = true;
            }
            n_queued += (int) images.size ();
        }
    return n_queued;
}

bool
sixel_draw (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int src_w, int src_h, int rowstride)
{
    return sixel_draw_begin (c, d_src, type, src_w, src_h, rowstride, nullptr) && sixel_draw_finish (c);
}

/* Wait for the stream without spinning: a diffusion chain runs for a quarter of a second, and a
 * caller with many stills in flight has more waiting threads than the host has cores */
static bool
wait_sleeping (SixelState *s, cudaStream_t st)
{
    return CK (cudaEventRecord (s->chain_done, st)) && CK (cudaEventSynchronize (s->chain_done));
}

static char *
put_dec_u8 (char *p, unsigned v)
{
    v &= 0xff;
    if (v >= 100) *p++ = (char) ('0' + v / 100);
    if (v >= 10) *p++ = (char) ('0' + (v / 10) % 10);
    *p++ = (char) ('0' + v % 10);
    return p;
}

GString *
sixel_print (ChafaCanvas *c, ChafaTermInfo *ti)
{
    const ChafaCanvasConfig *cfg = canvas_config (c);
    cudaStream_t st = canvas_stream (c);
    SixelState *s = (SixelState *) *canvas_sixel_slot (c);
    if (!s || !s->have_image || !palette_on_host (s))
        return nullptr;

    /* header: begin-sixels (0;1;0), raster attributes, palette (chafa-sixel-renderer.c:422-457, 497-511) */
    std::string head;
    {
        char buf [CHAFA_TERM_SEQ_LENGTH_MAX * 2];
        unsigned args [3] = { 0, 1, 0 };
        char *e = term_info_emit (ti, buf, CHAFA_TERM_SEQ_BEGIN_SIXELS, args, 3, 1);
        head.append (buf, (size_t) (e - buf));
        snprintf (buf, CHAFA_TERM_SEQ_LENGTH_MAX, "\"1;1;%d;%d", s->width, s->height);
        head += buf;
        for (int pen = 0; pen < s->n_colors; pen++)
        {
            if (pen == s->transparent_index)
                continue;
            uint32_t col = s->colors_rgb [s->first_color + pen];
            char *p = buf;
            *p++ = '#';
            p = put_dec_u8 (p, (unsigned) pen);
            *p++ = ';';
            *p++ = '2';
            *p++ = ';';
            p = put_dec_u8 (p, ((col & 0xff) * 100) / 255);
            *p++ = ';';
            p = put_dec_u8 (p, (((col >> 8) & 0xff) * 100) / 255);
            *p++ = ';';
            p = put_dec_u8 (p, (((col >> 16) & 0xff) * 100) / 255);
            head.append (buf, (size_t) (p - buf));
        }
    }
    std::string tail;
    {
        char buf [CHAFA_TERM_SEQ_LENGTH_MAX * 2];
        char *e = term_info_emit (ti, buf, CHAFA_TERM_SEQ_END_SIXELS, nullptr, 0, 1);
        tail.assign (buf, (size_t) (e - buf));
    }

    /* A row band prints its own sixel bands; the header goes with the first band of the image and
     * the terminator with the last, so that the owners' texts concatenate to the full print */
    const int band0 = s->own_first / 6;
    const int band1 = s->own_end >= s->height ? s->image_height / 6 : s->own_end / 6;
    if (s->own_first > 0)
        head.clear ();
    if (s->own_end < s->height)
        tail.clear ();

    /* body: K6d */
    int n_bands = band1 - band0;
    size_t n_pairs = (size_t) n_bands * (size_t) std::max (s->n_colors, 1);
    size_t lens_bytes = n_pairs * 12 + (size_t) n_bands * 4 + ((size_t) n_bands + 1) * 8 + 64;
    if (!grow (&s->d_lens, &s->cap_lens, lens_bytes))
        return nullptr;

    DevSixelEncode e;
    memset (&e, 0, sizeof (e));
    uint8_t *base = (uint8_t *) s->d_lens;
    e.band_ofs = (unsigned long long *) base;
    base += ((size_t) n_bands + 1) * 8;
    e.body_len = (uint32_t *) base;
    base += n_pairs * 4;
    e.prefix_len = (uint32_t *) base;
    base += n_pairs * 4;
    e.pen_ofs = (uint32_t *) base;
    base += n_pairs * 4;
    e.band_len = (uint32_t *) base;
    e.indexed = (const uint8_t *) s->d_indexed + (size_t) band0 * 6 * s->width;
    e.width = s->width;
    e.height = s->height;
    e.n_bands = n_bands;
    e.band0 = band0;
    e.n_colors = s->n_colors;
    e.transparent_index = s->transparent_index;
    e.first_pen = s->transparent_index == 0 ? 1 : 0;

    unsigned long long body_len = 0;
    if (s->n_colors > 0 && n_bands > 0)
    {
        if (!CK (dev_launch_sixel_encode_measure (&e, st))
            || !CK (cudaMemcpyAsync (&body_len, e.band_ofs + n_bands, 8, cudaMemcpyDeviceToHost, st))
            || !wait_sleeping (s, st))
            return nullptr;
        if (!grow (&s->d_out, &s->cap_out, (size_t) body_len + 16))
            return nullptr;
        e.out = (uint8_t *) s->d_out;
        if (!CK (dev_launch_sixel_encode_emit (&e, st)))
            return nullptr;
    }

    /* D2H of the body into pinned staging */
    if (body_len)
    {
        if (body_len > s->cap_h_text)
        {
            if (s->h_text)
                cudaFreeHost (s->h_text);
            s->h_text = nullptr;
            s->cap_h_text = 0;
            if (!CK (cudaMallocHost (&s->h_text, (size_t) body_len + body_len / 4)))
                return nullptr;
            s->cap_h_text = (size_t) body_len + body_len / 4;
        }
        if (!CK (cudaMemcpyAsync (s->h_text, s->d_out, (size_t) body_len, cudaMemcpyDeviceToHost, st))
            || !CK (cudaStreamSynchronize (st)))
            return nullptr;
    }
    const char *body = (const char *) s->h_text;

    if (cfg->passthrough == CHAFA_PASSTHROUGH_NONE)
    {
        size_t total = head.size () + (size_t) body_len + tail.size ();
        char *buf = (char *) b200_malloc (total + 1);
        memcpy (buf, head.data (), head.size ());
        if (body_len)
            memcpy (buf + head.size (), body, (size_t) body_len);
        memcpy (buf + head.size () + body_len, tail.data (), tail.size ());
        buf [total] = '\0';
        return b200_string_new_take (buf, total, total + 1);
    }

    /* Multiplexer passthrough framing (chafa/internal/chafa-passthrough-encoder.c:27-181,
     * chafa-sixel-renderer.c:459-484): the stream is cut into packets, each wrapped in the
     * multiplexer's begin / end sequences -- 200 bytes for screen, 1 000 000 for tmux, which
     * also wants every ESC doubled; under screen each byte of the end-of-sixels sequence
     * travels in a packet of its own.  Pure framing of bytes the device produced. */
    const bool screen = cfg->passthrough == CHAFA_PASSTHROUGH_SCREEN;
    const size_t packet_max = screen ? 200 : 1000000;
    std::string pt_begin, pt_end;
    {
        char buf [CHAFA_TERM_SEQ_LENGTH_MAX * 2];
        char *end = term_info_emit (ti, buf, screen ? CHAFA_TERM_SEQ_BEGIN_SCREEN_PASSTHROUGH
                                                    : CHAFA_TERM_SEQ_BEGIN_TMUX_PASSTHROUGH, nullptr, 0, 1);
        pt_begin.assign (buf, (size_t) (end - buf));
        end = term_info_emit (ti, buf, screen ? CHAFA_TERM_SEQ_END_SCREEN_PASSTHROUGH
                                              : CHAFA_TERM_SEQ_END_TMUX_PASSTHROUGH, nullptr, 0, 1);
        pt_end.assign (buf, (size_t) (end - buf));
    }

    std::string out;
    out.reserve ((size_t) ((head.size () + body_len + tail.size ()) * (screen ? 1.2 : 1.05)) + 256);
    size_t packet_size = 0;
    auto packetized = [&] (const char *in, size_t len)
    {
        while (len > 0)
        {
            size_t remain = packet_max - packet_size;
            if (remain == 0)
            {
                out += pt_end;
                packet_size = 0;
                remain = packet_max;
            }
            if (packet_size == 0)
                out += pt_begin;
            size_t n = std::min (len, remain);
            out.append (in, n);
            len -= n;
            packet_size += n;
            in += n;
        }
    };
    auto flush = [&] ()
    {
        if (packet_size > 0)
        {
            out += pt_end;
            packet_size = 0;
        }
    };
    auto append = [&] (const char *in, size_t len)
    {
        if (screen)
        {
            packetized (in, len);
            return;
        }
        /* tmux: ESC doubled, in pieces of at most 1024 output bytes like the reference */
        char buf [1024];
        size_t j = 0;
        for (size_t i = 0; i < len; i++)
        {
            buf [j++] = in [i];
            if (in [i] == 0x1b)
                buf [j++] = 0x1b;
            if (j + 2 > sizeof (buf))
            {
                packetized (buf, j);
                j = 0;
            }
        }
        packetized (buf, j);
    };

    append (head.data (), head.size ());
    if (body_len)
        append (body, (size_t) body_len);
    if (screen)
    {
        for (size_t i = 0; i < tail.size (); i++)
        {
            flush ();
            packetized (tail.data () + i, 1);
        }
    }
    else
    {
        append (tail.data (), tail.size ());
    }
    flush ();
    return b200_string_new_copy (out.data (), out.size ());
}

/* Parity read-back: the indexed image and palette of the last sixel draw */
extern "C" int
chafa_b200_canvas_get_sixel_image (ChafaCanvas *canvas, uint8_t *pixels_out, int *width_out, int *height_out,
                                   int *n_colors_out, uint32_t *palette_out)
{
    if (!canvas)
        return 0;
    SixelState *s = (SixelState *) *canvas_sixel_slot (canvas);
    if (!s || !s->have_image || !palette_on_host (s))
        return 0;
    if (width_out) *width_out = s->width;
    if (height_out) *height_out = s->image_height;
    if (n_colors_out) *n_colors_out = s->n_colors;
    if (palette_out)
        for (int i = 0; i < 256; i++)
            palette_out [i] = s->colors_rgb [i];
    if (pixels_out)
    {
        cudaStream_t st = canvas_stream (canvas);
        if (!CK (cudaMemcpyAsync (pixels_out, s->d_indexed, (size_t) s->width * s->image_height,
                                  cudaMemcpyDeviceToHost, st))
            || !CK (cudaStreamSynchronize (st)))
            return 0;
    }
    return 1;
}
