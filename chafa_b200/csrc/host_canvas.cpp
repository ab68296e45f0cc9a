/* host_canvas.cpp -- ChafaCanvas: the drop-in entry points of the canvas rendering path.
 *
 * Replaces chafa/chafa-canvas.c of the reference (chafa_canvas_new 513-631,
 * draw_all_pixels 392-501 / 786-805, print 866-922, print_rows 951-1032, cell
 * accessors 1048-1402) with a device-resident design:
 *
 *   canvas creation   compile both symbol maps, palettes, dither texture and scaler
 *                     LUTs to flat arrays and upload them once
 *   draw_all_pixels   H2D copy of the borrowed source, then kernels K1 (scale +
 *                     composite), K2 (histogram / normalise / dither / DIN99d), K3/K3w
 *                     (cell matching) and K4 (row resolve) queued on the canvas's stream;
 *                     the call returns once the source has been consumed
 *   print             K5 (measure -> scan -> emit) and one D2H copy of the text
 *
 * Nothing here computes pixels or cells on the CPU: if CUDA is unavailable
 * chafa_canvas_new () fails loudly and returns NULL. */

#include "internal.h"
#include "device.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <mutex>

#include "gen_tables.inc"

/* ------------------------------------------------------------------------ */
/* CUDA plumbing                                                              */

static std::atomic<uint64_t> g_launch_count {0};

void
dev_count_launch (int n)
{
    g_launch_count += (uint64_t) n;
}

static bool
cuda_ok (cudaError_t err, const char *what)
{
    if (err == cudaSuccess)
        return true;
    b200_set_error ("%s: %s", what, cudaGetErrorString (err));
    fprintf (stderr, "chafa-b200: %s: %s\n", what, cudaGetErrorString (err));
    return false;
}

#define CU(call) cuda_ok ((call), #call)

#define SRC_RING_SLOTS 4
#define SRC_RING_CHUNK ((size_t) 4 << 20)

struct DevBuf
{
    void *p = nullptr;
    size_t cap = 0;

    bool ensure (size_t n)
    {
        if (n <= cap)
            return true;
        if (p)
            b200_dev_free (p);
        p = nullptr;
        cap = 0;
        size_t want = b200_dev_guard_mode () ? n : n + n / 8 + 256;     /* guard mode: no slack to hide in */
        if (!CU ((cudaError_t) b200_dev_malloc (&p, want)))
            return false;
        cap = want;
        return true;
    }
    void release ()
    {
        if (p)
            b200_dev_free (p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as () const { return (T *) p; }
};

struct PinnedBuf
{
    void *p = nullptr;
    size_t cap = 0;

    bool ensure (size_t n)
    {
        if (n <= cap)
            return true;
        if (p)
            cudaFreeHost (p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        if (!CU (cudaMallocHost (&p, want)))
            return false;
        cap = want;
        return true;
    }
    void release ()
    {
        if (p)
            cudaFreeHost (p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as () const { return (T *) p; }
};

/* Optional per-stage device timing with CUDA events on the canvas's stream
 * (chafa_b200_canvas_set_kernel_timing); used by bench.py for the roofline figures. */
enum
{
    ST_H2D = 0, ST_SCALE, ST_PREP, ST_MATCH, ST_MATCH_WIDE, ST_RESOLVE, ST_PRINT_MEASURE, ST_PRINT_EMIT, ST_D2H,
    ST_MAX
};

struct StageTimer
{
    bool enabled = false;
    cudaEvent_t ev [ST_MAX] [2] = {};
    bool pending [ST_MAX] = {};
    double ms [ST_MAX] = {};
    uint64_t count [ST_MAX] = {};

    void begin (int st, cudaStream_t s)
    {
        if (!enabled)
            return;
        if (pending [st])
            collect ();
        if (!ev [st] [0])
        {
            cudaEventCreate (&ev [st] [0]);
            cudaEventCreate (&ev [st] [1]);
        }
        cudaEventRecord (ev [st] [0], s);
    }
    void end (int st, cudaStream_t s)
    {
        if (!enabled)
            return;
        cudaEventRecord (ev [st] [1], s);
        pending [st] = true;
    }
    void collect ()
    {
        for (int i = 0; i < ST_MAX; i++)
        {
            if (!pending [i])
                continue;
            float t = 0.0f;
            cudaEventSynchronize (ev [i] [1]);
            if (cudaEventElapsedTime (&t, ev [i] [0], ev [i] [1]) == cudaSuccess)
            {
                ms [i] += t;
                count [i]++;
            }
            pending [i] = false;
        }
    }
    void release ()
    {
        for (int i = 0; i < ST_MAX; i++)
            for (int j = 0; j < 2; j++)
                if (ev [i] [j])
                {
                    cudaEventDestroy (ev [i] [j]);
                    ev [i] [j] = nullptr;
                }
    }
};

/* Scaler LUTs: one copy per device, shared by all canvases */
struct DeviceTables
{
    uint16_t *from_srgb = nullptr;
    uint8_t *to_srgb = nullptr;
    uint32_t *inv_div = nullptr;
};

static std::mutex g_tables_mutex;
static std::map<int, DeviceTables> g_tables;
static thread_local int t_device = -1;

static const DeviceTables *
device_tables (int device)
{
    std::lock_guard<std::mutex> lock (g_tables_mutex);
    auto it = g_tables.find (device);
    if (it != g_tables.end ())
        return &it->second;

    DeviceTables t;
    if (!CU (cudaMalloc ((void **) &t.from_srgb, sizeof (gen_from_srgb_lut)))
        || !CU (cudaMalloc ((void **) &t.to_srgb, sizeof (gen_to_srgb_lut)))
        || !CU (cudaMalloc ((void **) &t.inv_div, 4 * 256 * sizeof (uint32_t))))
        return nullptr;
    if (!CU (cudaMemcpy (t.from_srgb, gen_from_srgb_lut, sizeof (gen_from_srgb_lut), cudaMemcpyHostToDevice))
        || !CU (cudaMemcpy (t.to_srgb, gen_to_srgb_lut, sizeof (gen_to_srgb_lut), cudaMemcpyHostToDevice))
        || !CU (cudaMemcpy (t.inv_div, gen_inv_div_p8_lut, 1024, cudaMemcpyHostToDevice))
        || !CU (cudaMemcpy (t.inv_div + 256, gen_inv_div_p8l_lut, 1024, cudaMemcpyHostToDevice))
        || !CU (cudaMemcpy (t.inv_div + 512, gen_inv_div_p16_lut, 1024, cudaMemcpyHostToDevice))
        || !CU (cudaMemcpy (t.inv_div + 768, gen_inv_div_p16l_lut, 1024, cudaMemcpyHostToDevice)))
        return nullptr;
    g_tables [device] = t;
    return &g_tables [device];
}

/* Scaler tables of the device new canvases would use (for table-building helpers that run
 * a kernel without owning a canvas, e.g. glyph import) */
const void *
device_tables_for_current (const uint16_t **from_srgb, const uint8_t **to_srgb, const uint32_t **inv_div)
{
    int n = 0, device = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess || n <= 0)
    {
        cudaGetLastError ();
        return nullptr;
    }
    if (t_device >= 0)
        device = t_device;
    else if (cudaGetDevice (&device) != cudaSuccess)
        device = 0;
    if (!CU (cudaSetDevice (device)))
        return nullptr;
    const DeviceTables *t = device_tables (device);
    if (!t)
        return nullptr;
    *from_srgb = t->from_srgb;
    *to_srgb = t->to_srgb;
    *inv_div = t->inv_div;
    return t;
}

/* ------------------------------------------------------------------------ */
/* Host colour helpers                                                        */

/* chafa/internal/chafa-color.c:83-168, in double like the reference.  Used for the
 * handful of colours the host owns (default fg/bg, FG/BG pens); per-pixel conversion
 * is kernel K2. */
uint32_t
host_rgb_to_din99d (uint32_t color)
{
    double lin [3], xyz [3], fl [3];

    for (int i = 0; i < 3; i++)
    {
        double v = (double) ((color >> (8 * i)) & 0xff) / 255.0;
        lin [i] = v <= 0.04045 ? (v / 12.92) : pow ((v + 0.055) / 1.044, 2.4);
    }
    xyz [0] = 0.4124564 * lin [0] + 0.3575761 * lin [1] + 0.1804375 * lin [2];
    xyz [1] = 0.2126729 * lin [0] + 0.7151522 * lin [1] + 0.0721750 * lin [2];
    xyz [2] = 0.0193339 * lin [0] + 0.1191920 * lin [1] + 0.9503041 * lin [2];
    xyz [0] = 1.12 * xyz [0] - 0.12 * xyz [2];

    const double wp [3] = { 0.95047, 1.0, 1.08883 };
    for (int i = 0; i < 3; i++)
    {
        double v = xyz [i] / wp [i];
        fl [i] = v > (216.0 / 24389.0) ? cbrt (v) : ((24389.0 / 27.0) * v + 16.0) / 116.0;
    }
    double L = 116.0 * fl [1] - 16.0;
    double a = 500.0 * (fl [0] - fl [1]);
    double b = 200.0 * (fl [1] - fl [2]);
    double adj_L = 325.22 * log (1.0 + 0.0036 * L);
    double ee = 0.6427876096865393 * a + 0.766044443118978 * b;
    double f = 1.14 * (0.6427876096865393 * b - 0.766044443118978 * a);
    double G = sqrt (ee * ee + f * f);
    double C = 22.5 * log (1.0 + 0.06 * G);
    double h = atan2 (f, ee) + 0.8726646;
    while (h < 0.0) h += 6.283185;
    while (h > 6.283185) h -= 6.283185;

    uint32_t c0 = (uint32_t) (uint8_t) (adj_L * 2.5);
    uint32_t c1 = (uint32_t) (uint8_t) (C * cos (h) * 2.5 + 128.0);
    uint32_t c2 = (uint32_t) (uint8_t) (C * sin (h) * 2.5 + 128.0);
    return c0 | (c1 << 8) | (c2 << 16) | (color & 0xff000000u);
}

static void
init_palette (DevPalette *pal, int type, int alpha_threshold, uint32_t fg_rgb, uint32_t bg_rgb, int color_space)
{
    memset (pal, 0, sizeof (*pal));
    pal->type = type;
    pal->alpha_threshold = alpha_threshold;
    pal->transparent_index = PAL_INDEX_TRANSPARENT;

    switch (type)
    {
        case PAL_FIXED_FGBG: pal->first_color = PAL_INDEX_FG; pal->n_colors = 2; break;
        case PAL_FIXED_8: pal->n_colors = 8; break;
        case PAL_FIXED_16: pal->n_colors = 16; break;
        case PAL_FIXED_240: pal->first_color = 16; pal->n_colors = 240; break;
        case PAL_FIXED_256: pal->n_colors = 256; break;
        default: break;
    }

    const uint32_t *fixed = color_space == CHAFA_COLOR_SPACE_DIN99D ? gen_fixed_palette_din99d : gen_fixed_palette_rgb;
    for (int i = 0; i < PAL_INDEX_MAX; i++)
    {
        pal->fixed [i] = fixed [i];
        pal->colors [i] = fixed [i];
        pal->colors_rgb [i] = gen_fixed_palette_rgb [i];
    }

    /* FG pen opaque, BG pen transparent (chafa-canvas.c:226-230, 279-293) */
    uint32_t fg = color_from_packed_rgb (fg_rgb, 0xff), bg = color_from_packed_rgb (bg_rgb, 0x00);
    pal->colors_rgb [PAL_INDEX_FG] = fg;
    pal->colors_rgb [PAL_INDEX_BG] = bg;
    pal->colors [PAL_INDEX_FG] = color_space == CHAFA_COLOR_SPACE_DIN99D ? host_rgb_to_din99d (fg) : fg;
    pal->colors [PAL_INDEX_BG] = color_space == CHAFA_COLOR_SPACE_DIN99D ? host_rgb_to_din99d (bg) : bg;

    /* Nearest cube level per channel value (chafa-palette.c:226-237) */
    static const int levels [6] = { 0x00, 0x5f, 0x87, 0xaf, 0xd7, 0xff };
    int level = 0;
    for (int v = 0; v < 256; v++)
    {
        while (level < 5 && v >= (levels [level] + levels [level + 1]) / 2)
            level++;
        pal->cube_index [v] = (uint8_t) level;
    }
}

/* ------------------------------------------------------------------------ */
/* The canvas                                                                 */

struct ChafaCanvas
{
    std::atomic<int> refs {1};
    ChafaCanvasConfig config;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t src_consumed = nullptr;

    int width_px = 0, height_px = 0;
    int work_factor_int = 5;
    bool consider_inverted = true, extract_colors = true, use_quantized_error = false;
    std::vector<uint32_t> plan_tables_on_device;   /* scaler filter tables as last uploaded */
    ScalePlan plan_cached;                          /* scaler plan of plan_geom */
    std::vector<uint32_t> plan_filt;                /* its filter tables, h then v */
    int plan_geom [10] = { 0 };                     /* src w, h, dest w, h, placement x, y, w, h, nearest, valid */
    bool plan_uploaded_this_draw = false;
    bool prep_fused = false;        /* this draw: the scaler applied dither / colour-space conversion */
    bool scale_post_on = false;     /* handed to canvas_run_scaler by the symbols pipeline */
    DevPrepPixel scale_post;
    bool use_tensor_match = true;   /* K3t (match_tc.cu) where it applies; off = the warp-per-cell kernels */
    uint32_t blank_char = ' ', solid_char = 0;
    uint32_t default_fg = 0, default_bg = 0;
    ChafaPlacement *placement = nullptr;
    bool needs_clear = true;
    /* Pageable sources (what a drop-in caller hands over: a malloc'ed frame) are staged through a
     * ring of pinned chunks by the calling thread: the memcpy of chunk k + 1 overlaps the DMA of chunk
     * k, callers on several threads copy in parallel (the runtime's own staging of pageable memory
     * serialises them), and the call returns when the frame has been read, not when it has landed */
    PinnedBuf h_ring;
    cudaEvent_t ring_free [SRC_RING_SLOTS] = { nullptr, nullptr, nullptr, nullptr };
    bool ring_busy [SRC_RING_SLOTS] = { false, false, false, false };
    int ring_next = 0;

    /* The row walk (K4) of the last draw, owed until somebody needs the cells: the printer runs it
     * as the first phase of its own launch; anything else that reads cells runs it on its own */
    bool resolve_pending = false;
    DevResolveArgs resolve_args;
    ScaleTiles tiles = { 0, 0, 0, 0 };
    int tiles_key [3] = { 0, 0, 0 };
    int plan_serial = 0;            /* counts rebuilds of plan_cached */
    bool sixel_reduce = false;      /* two-phase sixel draw: the colour bins await summing across band owners */
    int band_first = 0, band_rows = 0;

    /* dither (chafa/internal/chafa-dither.c:55-84) */
    int dither_mode = 0;
    float dither_intensity = 1.0f;
    int grain_w_shift = 0, grain_h_shift = 0;
    int tex_size_shift = 0, tex_size_mask = 0;

    DevPalette h_fg_palette, h_bg_palette;

    /* device-resident tables */
    DevBuf d_tables;       /* symbols, palettes, texture: one allocation */
    DevSymbols d_syms;
    DevPalette *d_fg_palette = nullptr, *d_bg_palette = nullptr;
    int32_t *d_texture = nullptr;
    int32_t *d_hist = nullptr;   /* 2049 + 2 bounds */

    /* per-draw buffers */
    DevBuf d_src, d_pixels, d_evals, d_wide, d_cells, d_precalc, d_fs_scratch;
    PinnedBuf h_stage;
    bool have_pixels = false;
    bool src_event_pending = false;

    /* print buffers */
    DevBuf d_cell_len, d_row_ofs, d_out, d_ti;
    DevTermInfo ti_last;            /* what d_ti holds */
    bool ti_uploaded = false;
    DevBuf d_print_ctl;             /* single-launch printer: 2 control words + one status word per row */
    uint32_t print_epoch = 0;
    PinnedBuf h_print, h_text;
    const uint8_t *last_text = nullptr;         /* enqueue-only print: where its text and length are */
    const uint64_t *last_text_len = nullptr;

    /* host view of the cells for the accessor API */
    std::vector<DevCell> h_cells;
    bool h_cells_valid = true;    /* host copy is current */
    bool d_cells_valid = false;   /* device copy is current */

    StageTimer timer;

    /* sixel state lives in host_sixel.cpp */
    void *sixel = nullptr;
    /* Kitty / iTerm2 state lives in host_pixelimg.cpp */
    void *pixelimg = nullptr;
    int scale_composite = 0;        /* DevScale.composite for the next scaler run */
};

void sixel_state_free (void *state);
void pixelimg_state_free (void *state);

#define CHECK_CANVAS(c) do { RETURN_IF_FAIL ((c) != NULL); RETURN_IF_FAIL ((c)->refs.load () > 0); } while (0)
#define CHECK_CANVAS_VAL(c, v) do { RETURN_VAL_IF_FAIL ((c) != NULL, (v)); RETURN_VAL_IF_FAIL ((c)->refs.load () > 0, (v)); } while (0)

static int
grain_shift (int size)
{
    int s = 0;
    while (size >>= 1)
        s++;
    return s;
}

static bool
canvas_bind_device (const ChafaCanvas *canvas)
{
    return CU (cudaSetDevice (canvas->device));
}

static void
canvas_free_device (ChafaCanvas *c)
{
    cudaSetDevice (c->device);
    if (c->stream)
        cudaStreamSynchronize (c->stream);
    c->h_ring.release ();
    for (int i = 0; i < SRC_RING_SLOTS; i++)
        if (c->ring_free [i])
        {
            cudaEventDestroy (c->ring_free [i]);
            c->ring_free [i] = nullptr;
        }
    c->d_tables.release ();
    c->d_src.release ();
    c->d_pixels.release ();
    c->d_evals.release ();
    c->d_wide.release ();
    c->d_cells.release ();
    c->d_precalc.release ();
    c->d_fs_scratch.release ();
    c->d_cell_len.release ();
    c->d_row_ofs.release ();
    c->d_print_ctl.release ();
    c->d_out.release ();
    c->d_ti.release ();
    c->h_stage.release ();
    c->h_print.release ();
    c->h_text.release ();
    c->timer.release ();
    if (c->src_consumed)
        cudaEventDestroy (c->src_consumed);
    if (c->own_stream && c->stream)
        cudaStreamDestroy (c->stream);
    c->src_consumed = nullptr;
    c->stream = nullptr;
}

/* Flatten symbol maps, palettes and the dither texture into one device allocation */
static bool
canvas_upload_tables (ChafaCanvas *c)
{
    const ChafaSymbolMap &sm = c->config.symbol_map;
    const ChafaSymbolMap &fm = c->config.fill_symbol_map;
    size_t n = sm.symbols.size (), n2 = sm.symbols2.size (), nf = fm.symbols.size ();
    size_t tex_n = c->dither_mode == CHAFA_DITHER_MODE_ORDERED ? 256
        : c->dither_mode == CHAFA_DITHER_MODE_NOISE ? 64 * 64 * 3 : 0;

    std::vector<uint8_t> blob;
    auto reserve = [&] (size_t bytes, size_t align) -> size_t
    {
        size_t ofs = (blob.size () + align - 1) / align * align;
        blob.resize (ofs + bytes);
        return ofs;
    };

    size_t o_bitmaps = reserve (n * 8, 16);
    size_t o_bitmaps2 = reserve (n2 * 16, 16);
    size_t o_chars = reserve (n * 4, 16);
    size_t o_chars2 = reserve (n2 * 4, 16);
    size_t o_fill_chars = reserve (nf * 4, 16);
    size_t o_fg_pal = reserve (sizeof (DevPalette), 16);
    size_t o_bg_pal = reserve (sizeof (DevPalette), 16);
    size_t o_tex = reserve (tex_n * 4, 16);
    size_t o_hist = reserve ((2049 + 2) * 4, 16);
    size_t o_pop = reserve (n, 16);
    size_t o_pop2 = reserve (n2 * 2, 16);
    size_t o_fill_pop = reserve (nf, 16);
    size_t npad = (n + 127) / 128 * 128, n2pad = (n2 + 127) / 128 * 128;
    size_t o_bimage = reserve (npad * 64, 128);
    size_t o_bimage2 = reserve (n2pad * 128, 128);

    for (size_t i = 0; i < n; i++)
    {
        ((uint64_t *) &blob [o_bitmaps]) [i] = sm.symbols [i].bitmap;
        ((uint32_t *) &blob [o_chars]) [i] = sm.symbols [i].c;
        blob [o_pop + i] = (uint8_t) sm.symbols [i].popcount;
    }
    for (size_t i = 0; i < n2; i++)
    {
        ((uint64_t *) &blob [o_bitmaps2]) [2 * i] = sm.symbols2 [i].bitmap [0];
        ((uint64_t *) &blob [o_bitmaps2]) [2 * i + 1] = sm.symbols2 [i].bitmap [1];
        ((uint32_t *) &blob [o_chars2]) [i] = sm.symbols2 [i].c;
        uint64_t a = sm.symbols2 [i].bitmap [0], b = sm.symbols2 [i].bitmap [1];
        blob [o_pop2 + 2 * i] = (uint8_t) __builtin_popcountll (a);
        blob [o_pop2 + 2 * i + 1] = (uint8_t) __builtin_popcountll (b);
    }
    for (size_t i = 0; i < nf; i++)
    {
        ((uint32_t *) &blob [o_fill_chars]) [i] = fm.symbols [i].c;
        blob [o_fill_pop + i] = (uint8_t) fm.symbols [i].popcount;
    }
    /* operand images for the tensor-core matcher: K index = pixel index, bitmap MSB = pixel 0 */
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 64; k++)
            blob [o_bimage + ((size_t) (k >> 4) * npad + i) * 16 + (k & 15)]
                = ((sm.symbols [i].bitmap >> (63 - k)) & 1) ? 0xff : 0x01;
    for (size_t i = 0; i < n2; i++)
        for (int k = 0; k < 128; k++)
            blob [o_bimage2 + ((size_t) (k >> 4) * n2pad + i) * 16 + (k & 15)]
                = ((sm.symbols2 [i].bitmap [k >> 6] >> (63 - (k & 63))) & 1) ? 0xff : 0x01;
    memcpy (&blob [o_fg_pal], &c->h_fg_palette, sizeof (DevPalette));
    memcpy (&blob [o_bg_pal], &c->h_bg_palette, sizeof (DevPalette));

    if (c->dither_mode == CHAFA_DITHER_MODE_ORDERED)
        gen_bayer_matrix ((int *) &blob [o_tex], 16, c->dither_intensity);
    else if (c->dither_mode == CHAFA_DITHER_MODE_NOISE)
        gen_noise_matrix ((int *) &blob [o_tex], c->dither_intensity * 0.1f);

    if (!c->d_tables.ensure (blob.size ()))
        return false;
    if (!CU (cudaMemcpy (c->d_tables.p, blob.data (), blob.size (), cudaMemcpyHostToDevice)))
        return false;

    uint8_t *base = c->d_tables.as<uint8_t> ();
    c->d_syms.bitmaps = (const uint64_t *) (base + o_bitmaps);
    c->d_syms.chars = (const uint32_t *) (base + o_chars);
    c->d_syms.popcounts = base + o_pop;
    c->d_syms.n = (int32_t) n;
    c->d_syms.bitmaps2 = (const uint64_t *) (base + o_bitmaps2);
    c->d_syms.chars2 = (const uint32_t *) (base + o_chars2);
    c->d_syms.popcounts2 = base + o_pop2;
    c->d_syms.n2 = (int32_t) n2;
    c->d_syms.fill_chars = (const uint32_t *) (base + o_fill_chars);
    c->d_syms.fill_popcounts = base + o_fill_pop;
    c->d_syms.n_fill = (int32_t) nf;
    c->d_syms.bimage = base + o_bimage;
    c->d_syms.bimage2 = base + o_bimage2;
    c->d_syms.npad = (int32_t) npad;
    c->d_syms.n2pad = (int32_t) n2pad;
    c->d_fg_palette = (DevPalette *) (base + o_fg_pal);
    c->d_bg_palette = (DevPalette *) (base + o_bg_pal);
    c->d_texture = tex_n ? (int32_t *) (base + o_tex) : nullptr;
    c->d_hist = (int32_t *) (base + o_hist);
    return true;
}

static uint32_t
pick_blank_char (const ChafaCanvasConfig &cfg)
{
    Candidate cand [N_CANDIDATES_MAX];

    if (symbol_map_has_symbol (&cfg.symbol_map, 0x20) || symbol_map_has_symbol (&cfg.fill_symbol_map, 0x20))
        return 0x20;
    if (symbol_map_find_fill_candidates (&cfg.fill_symbol_map, 0, false, cand, N_CANDIDATES_MAX) > 0)
        return cfg.fill_symbol_map.symbols [cand [0].symbol_index].c;
    if (symbol_map_find_candidates (&cfg.symbol_map, 0, false, cand, N_CANDIDATES_MAX) > 0)
        return cfg.symbol_map.symbols [cand [0].symbol_index].c;
    return 0x20;
}

static uint32_t
pick_solid_char (const ChafaCanvasConfig &cfg)
{
    Candidate cand [N_CANDIDATES_MAX];

    if (symbol_map_has_symbol (&cfg.symbol_map, 0x2588) || symbol_map_has_symbol (&cfg.fill_symbol_map, 0x2588))
        return 0x2588;
    if (symbol_map_find_fill_candidates (&cfg.fill_symbol_map, 64, false, cand, N_CANDIDATES_MAX) > 0
        && cand [0].hamming_distance <= 32)
        return cfg.fill_symbol_map.symbols [cand [0].symbol_index].c;
    if (symbol_map_find_candidates (&cfg.symbol_map, ~(uint64_t) 0, false, cand, N_CANDIDATES_MAX) > 0
        && cand [0].hamming_distance <= 32)
        return cfg.symbol_map.symbols [cand [0].symbol_index].c;
    return 0;
}

static void
push_away (uint32_t &color, uint32_t ref, int min_diff)
{
    for (int i = 0; i < 3; i++)
    {
        int v = (int) ((color >> (8 * i)) & 0xff), r = (int) ((ref >> (8 * i)) & 0xff);
        int diff = v - r;
        if (diff >= -min_diff && diff <= 0)
            v = std::max (r - min_diff, 0);
        else if (diff <= min_diff && diff >= 0)
            v = std::min (r + min_diff, 255);
        color = (color & ~(0xffu << (8 * i))) | ((uint32_t) v << (8 * i));
    }
}

static bool canvas_prepare_device_luts (ChafaCanvas *c);

/* Derived state of a canvas from its (already copied) config */
static bool
canvas_setup (ChafaCanvas *c)
{
    ChafaCanvasConfig &cfg = c->config;

    if (cfg.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS)
    {
        c->width_px = cfg.width * 8;
        c->height_px = cfg.height * 8;
    }
    else
    {
        c->width_px = cfg.width * cfg.cell_width;
        c->height_px = cfg.height * cfg.cell_height;
    }

    c->work_factor_int = (int) (cfg.work_factor * 10 + 0.5f);
    c->consider_inverted = !(cfg.fg_only_enabled || cfg.canvas_mode == CHAFA_CANVAS_MODE_FGBG);
    c->extract_colors = !(cfg.canvas_mode == CHAFA_CANVAS_MODE_FGBG || cfg.canvas_mode == CHAFA_CANVAS_MODE_FGBG_BGFG);
    if (cfg.canvas_mode == CHAFA_CANVAS_MODE_FGBG)
        cfg.fg_only_enabled = true;
    c->use_quantized_error = cfg.canvas_mode == CHAFA_CANVAS_MODE_INDEXED_16_8 && !cfg.fg_only_enabled;

    symbol_map_prepare (&cfg.symbol_map);
    symbol_map_prepare (&cfg.fill_symbol_map);
    c->blank_char = pick_blank_char (cfg);
    c->solid_char = pick_solid_char (cfg);

    if (cfg.pixel_mode == CHAFA_PIXEL_MODE_KITTY || cfg.pixel_mode == CHAFA_PIXEL_MODE_ITERM2
        || (cfg.canvas_mode == CHAFA_CANVAS_MODE_TRUECOLOR && cfg.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS))
    {
        cfg.color_space = CHAFA_COLOR_SPACE_RGB;
        cfg.dither_mode = CHAFA_DITHER_MODE_NONE;
    }

    float base = 1.0f;
    if (cfg.dither_mode == CHAFA_DITHER_MODE_ORDERED)
    {
        switch (cfg.canvas_mode)
        {
            case CHAFA_CANVAS_MODE_TRUECOLOR:
            case CHAFA_CANVAS_MODE_INDEXED_256:
            case CHAFA_CANVAS_MODE_INDEXED_240: base = 0.1f; break;
            case CHAFA_CANVAS_MODE_INDEXED_16:
            case CHAFA_CANVAS_MODE_INDEXED_16_8: base = 0.25f; break;
            case CHAFA_CANVAS_MODE_INDEXED_8: base = 0.5f; break;
            default: base = 1.0f; break;
        }
    }
    c->dither_mode = cfg.dither_mode;
    c->dither_intensity = base * cfg.dither_intensity;
    c->grain_w_shift = grain_shift (cfg.dither_grain_width);
    c->grain_h_shift = grain_shift (cfg.dither_grain_height);
    if (c->dither_mode == CHAFA_DITHER_MODE_ORDERED)
    {
        c->tex_size_shift = 4;
        c->tex_size_mask = 15;
    }
    else if (c->dither_mode == CHAFA_DITHER_MODE_NOISE)
    {
        c->tex_size_shift = 6;
        c->tex_size_mask = 63;
    }
    else if (c->dither_mode == CHAFA_DITHER_MODE_DIFFUSION)
    {
        c->dither_intensity = std::min (c->dither_intensity, 1.0f);
    }

    /* Display colours in the canvas's colour space (chafa-canvas.c:153-200) */
    uint32_t fg = color_from_packed_rgb (cfg.fg_color_packed_rgb, 0xff);
    uint32_t bg = color_from_packed_rgb (cfg.bg_color_packed_rgb, 0xff);
    if (cfg.color_space == CHAFA_COLOR_SPACE_DIN99D)
    {
        fg = host_rgb_to_din99d (fg);
        bg = host_rgb_to_din99d (bg);
    }
    fg |= 0xff000000u;
    bg &= 0x00ffffffu;
    if (c->extract_colors && cfg.fg_only_enabled)
    {
        fg = 0xff7f7f7fu;
        push_away (bg, fg, 5);
    }
    c->default_fg = fg;
    c->default_bg = bg;

    int fg_type = PAL_DYNAMIC_256, bg_type = PAL_DYNAMIC_256;
    switch (cfg.canvas_mode)
    {
        case CHAFA_CANVAS_MODE_INDEXED_256: fg_type = bg_type = PAL_FIXED_256; break;
        case CHAFA_CANVAS_MODE_INDEXED_240: fg_type = bg_type = PAL_FIXED_240; break;
        case CHAFA_CANVAS_MODE_INDEXED_16: fg_type = bg_type = PAL_FIXED_16; break;
        case CHAFA_CANVAS_MODE_INDEXED_16_8: fg_type = PAL_FIXED_16; bg_type = PAL_FIXED_8; break;
        case CHAFA_CANVAS_MODE_INDEXED_8: fg_type = bg_type = PAL_FIXED_8; break;
        case CHAFA_CANVAS_MODE_FGBG_BGFG:
        case CHAFA_CANVAS_MODE_FGBG: fg_type = bg_type = PAL_FIXED_FGBG; break;
        default: break;
    }
    init_palette (&c->h_fg_palette, fg_type, cfg.alpha_threshold, cfg.fg_color_packed_rgb,
                  cfg.bg_color_packed_rgb, cfg.color_space);
    init_palette (&c->h_bg_palette, bg_type, cfg.alpha_threshold, cfg.fg_color_packed_rgb,
                  cfg.bg_color_packed_rgb, cfg.color_space);

    c->band_first = 0;
    c->band_rows = cfg.height;

    size_t n_cells = (size_t) cfg.width * cfg.height;
    c->h_cells.assign (n_cells, DevCell { ' ', 0, 0 });
    c->h_cells_valid = true;
    c->d_cells_valid = false;
    c->needs_clear = true;

    /* Device side */
    if (!canvas_bind_device (c))
        return false;
    if (!device_tables (c->device))
        return false;
    if (!c->stream)
    {
        if (!CU (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking)))
            return false;
        c->own_stream = true;
    }
    if (!c->src_consumed && !CU (cudaEventCreateWithFlags (&c->src_consumed, cudaEventDisableTiming)))
        return false;
    if (!canvas_upload_tables (c))
        return false;
    if (!c->d_cells.ensure (n_cells * sizeof (DevCell)))
        return false;
    return canvas_prepare_device_luts (c);
}

static void
fill_dev_canvas (const ChafaCanvas *c, DevCanvas *cv)
{
    const ChafaCanvasConfig &cfg = c->config;

    memset (cv, 0, sizeof (*cv));
    cv->width = cfg.width;
    cv->height = cfg.height;
    cv->width_px = c->width_px;
    cv->height_px = c->height_px;
    cv->canvas_mode = cfg.canvas_mode;
    cv->color_space = cfg.color_space;
    cv->color_extractor = cfg.color_extractor;
    cv->work_factor_int = c->work_factor_int;
    cv->n_candidates = std::min (std::max (c->work_factor_int, 1), N_CANDIDATES_MAX);
    cv->consider_inverted = c->consider_inverted;
    cv->extract_colors = c->extract_colors;
    cv->use_quantized_error = c->use_quantized_error;
    cv->fg_only = cfg.fg_only_enabled;
    cv->alpha_threshold = cfg.alpha_threshold;
    cv->optimizations = cfg.optimizations;
    cv->blank_char = c->blank_char;
    cv->solid_char = c->solid_char;
    cv->default_fg = c->default_fg;
    cv->default_bg = c->default_bg;
    cv->first_row = c->band_first;
    cv->n_rows = c->band_rows;
    cv->fg_palette = c->d_fg_palette;
    cv->bg_palette = c->d_bg_palette;
    cv->syms = c->d_syms;
    cv->fg_pal_type = c->h_fg_palette.type;
    cv->bg_pal_type = c->h_bg_palette.type;
    cv->fg_pal_transparent = c->h_fg_palette.transparent_index;
    cv->bg_pal_transparent = c->h_bg_palette.transparent_index;
}

/* The per-device lookup tables a canvas of this configuration will want (DIN99d of every colour,
 * nearest pen of every colour) are built when the canvas is set up, where waiting is expected:
 * their first use synchronises the stream, which a draw must not do (the device entry points
 * only queue work, and a draw may sit inside a stream capture). */
static bool
canvas_prepare_device_luts (ChafaCanvas *c)
{
    if (c->config.pixel_mode != CHAFA_PIXEL_MODE_SYMBOLS)
        return true;
    if (c->config.color_space == CHAFA_COLOR_SPACE_DIN99D && !dev_din99d_lut (c->device, c->stream))
        return false;
    DevCanvas cv;
    fill_dev_canvas (c, &cv);
    dev_match_tc_prepare (&cv, c->stream);
    return true;
}

static int
bytes_per_pixel (ChafaPixelType t)
{
    return (t == CHAFA_PIXEL_RGB8 || t == CHAFA_PIXEL_BGR8) ? 3 : 4;
}

static bool
ensure_resolved (ChafaCanvas *c)
{
    if (!c->resolve_pending)
        return true;
    c->resolve_pending = false;
    if (!canvas_bind_device (c))
        return false;
    c->timer.begin (ST_RESOLVE, c->stream);
    bool ok = CU ((cudaError_t) dev_launch_resolve (&c->resolve_args.cv, c->resolve_args.evals,
                                                    c->resolve_args.wide_evals, c->resolve_args.cells, c->stream));
    c->timer.end (ST_RESOLVE, c->stream);
    return ok;
}

static bool
sync_cells_to_host (ChafaCanvas *c)
{
    if (c->h_cells_valid)
        return true;
    if (!canvas_bind_device (c) || !ensure_resolved (c))
        return false;
    if (!CU (cudaMemcpyAsync (c->h_cells.data (), c->d_cells.p, c->h_cells.size () * sizeof (DevCell),
                              cudaMemcpyDeviceToHost, c->stream))
        || !CU (cudaStreamSynchronize (c->stream)))
        return false;
    c->h_cells_valid = true;
    return true;
}

static bool
sync_cells_to_device (ChafaCanvas *c)
{
    if (c->d_cells_valid)
        return true;
    if (!canvas_bind_device (c))
        return false;
    /* pageable source: the runtime stages it before returning */
    if (!CU (cudaMemcpyAsync (c->d_cells.p, c->h_cells.data (), c->h_cells.size () * sizeof (DevCell),
                              cudaMemcpyHostToDevice, c->stream)))
        return false;
    c->d_cells_valid = true;
    return true;
}

/* Placement of the image on the pixmap, snapped to the cell grid
 * (chafa/internal/chafa-pixops.c:700-732) */
static void
symbols_placement (const ChafaCanvas *c, int src_w, int src_h, int *px, int *py, int *pw, int *ph)
{
    const ChafaCanvasConfig &cfg = c->config;
    ChafaAlign halign = CHAFA_ALIGN_START, valign = CHAFA_ALIGN_START;
    ChafaTuck tuck = CHAFA_TUCK_STRETCH;
    int cw = cfg.cell_width, ch = cfg.cell_height;
    int x, y, w, h;

    if (c->placement)
    {
        halign = c->placement->halign;
        valign = c->placement->valign;
        tuck = c->placement->tuck;
    }

    tuck_and_align (src_w, src_h, cfg.width * cw, cfg.height * ch, halign, valign, tuck, &x, &y, &w, &h);

    x -= x % cw;
    y -= y % ch;
    w = (w + cw - 1) / cw * cw;
    h = (h + ch - 1) / ch * ch;

    *px = (x / cw) * 8;
    *py = (y / ch) * 8;
    *pw = (w / cw) * 8;
    *ph = (h / ch) * 8;
}

static void
fill_scale_dim (DevScaleDim *d, const ScaleDim &s, const uint32_t *precalc)
{
    d->filter = s.filter;
    d->n_halvings = s.n_halvings;
    d->src_size_px = (int32_t) s.src_size_px;
    d->placement_ofs_px = s.placement_ofs_px;
    d->placement_size_px = s.placement_size_px;
    d->clip_before_px = s.clip_before_px;
    d->first_opacity = s.first_opacity;
    d->last_opacity = s.last_opacity;
    d->span_step = s.span_step;
    d->span_mul = s.span_mul;
    d->precalc = precalc;
}

bool sixel_draw (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int w, int h, int rowstride);
bool sixel_draw_begin (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int w, int h, int rowstride,
                       bool *reduce_out);
bool sixel_draw_float_bins (ChafaCanvas *c);
bool sixel_draw_finish (ChafaCanvas *c);
bool sixel_source_rows (ChafaCanvas *c, int src_w, int src_h, int *lo_out, int *hi_out);
void *sixel_exchange_buffer (ChafaCanvas *c, size_t *n_bytes_out, size_t *n_sum_words_out);
void sixel_set_exchange_buffer (ChafaCanvas *c, void *d_buffer);
size_t sixel_exchange_bytes (ChafaCanvas *c);
void sixel_set_defer_chain (ChafaCanvas *c, bool defer);
int sixel_launch_deferred_chains (ChafaCanvas **canvases, int n);
GString *sixel_print (ChafaCanvas *c, ChafaTermInfo *ti);
bool pixelimg_draw (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int w, int h, int rowstride);
GString *pixelimg_print (ChafaCanvas *c, ChafaTermInfo *ti);

/* K1: scale @d_src onto the canvas pixmap (placement in whole destination pixels),
 * producing destination rows [first_row, first_row + n_rows).  Uploads the filter plan,
 * then records the event after which no host memory of this draw is read any more. */
bool
canvas_run_scaler (ChafaCanvas *c, ChafaPixelType type, const void *d_src, int src_w, int src_h, int rowstride,
                   int px, int py, int pw, int ph, bool nearest, int first_row, int n_rows)
{
    HOSTPROF ("draw.scaler");
    const ChafaCanvasConfig &cfg = c->config;
    const DeviceTables *tables = device_tables (c->device);
    if (!tables)
        return false;

    /* The plan depends on the geometry only: a stream of equally sized frames builds it once */
    const int geom [10] = { src_w, src_h, c->width_px, c->height_px, px, py, pw, ph, nearest ? 1 : 0, 1 };
    if (memcmp (geom, c->plan_geom, sizeof (geom)) != 0)
    {
        c->plan_geom [9] = 0;
        if (!scale_plan_build (&c->plan_cached, src_w, src_h, c->width_px, c->height_px, px, py, pw, ph, nearest))
        {
            b200_set_error ("unsupported scaling geometry %dx%d -> %dx%d", src_w, src_h, c->width_px, c->height_px);
            return false;
        }
        memcpy (c->plan_geom, geom, sizeof (geom));
        c->plan_serial++;
        c->plan_filt.assign (c->plan_cached.h.precalc.begin (), c->plan_cached.h.precalc.end ());
        c->plan_filt.insert (c->plan_filt.end (), c->plan_cached.v.precalc.begin (), c->plan_cached.v.precalc.end ());
    }
    const ScalePlan &plan = c->plan_cached;
    const std::vector<uint32_t> &filt = c->plan_filt;

    size_t nh = plan.h.precalc.size (), nv = plan.v.precalc.size ();
    void *precalc_before = c->d_precalc.p;
    if (!c->d_precalc.ensure ((nh + nv + 1) * 4) || !c->h_stage.ensure ((nh + nv + 1) * 4))
        return false;
    /* ... and uploads its filter tables once */
    c->plan_uploaded_this_draw = false;
    if (c->d_precalc.p != precalc_before || filt != c->plan_tables_on_device)
    {
        uint32_t *stage = c->h_stage.as<uint32_t> ();
        if (nh + nv)
        {
            memcpy (stage, filt.data (), (nh + nv) * 4);
            if (!CU (cudaMemcpyAsync (c->d_precalc.p, stage, (nh + nv) * 4, cudaMemcpyHostToDevice, c->stream)))
                return false;
        }
        c->plan_tables_on_device = filt;
        c->plan_uploaded_this_draw = true;
    }
    /* Everything that reads host memory has been queued: the caller's source buffer and
     * the staging area are free again once this event fires (the kernels are not waited for). */
    if (!CU (cudaEventRecord (c->src_consumed, c->stream)))
        return false;
    c->src_event_pending = true;

    if (!c->d_pixels.ensure ((size_t) c->width_px * c->height_px * 4))
        return false;

    DevScale sc;
    memset (&sc, 0, sizeof (sc));
    fill_scale_dim (&sc.h, plan.h, c->d_precalc.as<uint32_t> ());
    fill_scale_dim (&sc.v, plan.v, c->d_precalc.as<uint32_t> () + nh);
    sc.src = (const uint8_t *) d_src;
    sc.src_rowstride = rowstride;
    sc.src_pixel_type = type;
    sc.linear = plan.linear;
    sc.dest = c->d_pixels.as<uint32_t> ();
    sc.dest_w = c->width_px;
    sc.dest_h = c->height_px;
    sc.bg_rgb = color_from_packed_rgb (cfg.bg_color_packed_rgb, 0) & 0xffffffu;
    sc.first_row = first_row;
    sc.n_rows = n_rows;
    sc.composite = c->scale_composite;
    sc.post_on = c->scale_post_on;
    if (c->scale_post_on)
        sc.post = c->scale_post;
    sc.from_srgb = tables->from_srgb;
    sc.to_srgb = tables->to_srgb;
    sc.inv_div = tables->inv_div;
    {
        /* tiling of the row-staged kernel: geometry and band only, so kept with the plan */
        const int tkey [3] = { first_row, n_rows, c->plan_serial };
        if (memcmp (tkey, c->tiles_key, sizeof (tkey)) != 0)
        {
            int sms = 148;
            cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, c->device);
            if (!scale_plan_tiles (&plan, c->width_px, first_row, n_rows, sms, &c->tiles))
                memset (&c->tiles, 0, sizeof (c->tiles));
            memcpy (c->tiles_key, tkey, sizeof (tkey));
        }
        sc.tiles = c->tiles;
    }
    c->timer.begin (ST_SCALE, c->stream);
    {
        HOSTPROF ("draw.scaler.launch");
        if (!CU ((cudaError_t) dev_launch_scale (&sc, c->stream)))
            return false;
    }
    c->timer.end (ST_SCALE, c->stream);
    c->have_pixels = true;
    return true;
}

static void
fill_prep (const ChafaCanvas *c, DevPrep *pp, int first_px_row, int n_px_rows)
{
    const ChafaCanvasConfig &cfg = c->config;
    int pal_type = c->h_fg_palette.type;
    bool pal_16_8 = pal_type == PAL_FIXED_16 || pal_type == PAL_FIXED_8;
    bool do_norm = cfg.preprocessing_enabled && (pal_16_8 || pal_type == PAL_FIXED_FGBG);

    memset (pp, 0, sizeof (*pp));
    pp->pixels = c->d_pixels.as<uint32_t> ();
    pp->width = c->width_px;
    pp->height = c->height_px;
    pp->first_row = first_px_row;
    pp->n_rows = n_px_rows;
    pp->boost_saturation = cfg.preprocessing_enabled && pal_16_8;
    pp->build_histogram = do_norm;
    pp->hist = c->d_hist;
    pp->bounds = c->d_hist + 2049;
    pp->normalize = do_norm;
    pp->crop_pct = pal_type == PAL_FIXED_16 ? 5 : pal_type == PAL_FIXED_8 ? 10 : 20;
    pp->dither_mode = c->dither_mode;
    pp->grain_w_shift = c->grain_w_shift;
    pp->grain_h_shift = c->grain_h_shift;
    pp->tex_size_shift = c->tex_size_shift;
    pp->tex_size_mask = c->tex_size_mask;
    pp->texture = c->d_texture;
    pp->to_din99d = cfg.color_space == CHAFA_COLOR_SPACE_DIN99D;
    pp->din99d_lut = pp->to_din99d ? dev_din99d_lut (c->device, c->stream) : nullptr;
}

/* Rows of the pixmap this canvas works on: its band, or everything when error diffusion
 * (one chain over the whole image) is on */
static void
band_pixel_rows (const ChafaCanvas *c, int *first_px_row, int *n_px_rows)
{
    if (c->dither_mode == CHAFA_DITHER_MODE_DIFFUSION)
    {
        *first_px_row = 0;
        *n_px_rows = c->height_px;
    }
    else
    {
        *first_px_row = c->band_first * 8;
        *n_px_rows = c->band_rows * 8;
    }
}

/* Symbols pipeline, first half: scale and the per-pixel pass that feeds the intensity
 * histogram (chafa-pixops.c:477-566).  Source pixels are on the device. */
static bool
draw_symbols_phase1 (ChafaCanvas *c, ChafaPixelType type, const void *d_src, int src_w, int src_h, int rowstride)
{
    int px, py, pw, ph, first_px_row, n_px_rows;

    symbols_placement (c, src_w, src_h, &px, &py, &pw, &ph);
    band_pixel_rows (c, &first_px_row, &n_px_rows);

    DevPrep pp;
    fill_prep (c, &pp, first_px_row, n_px_rows);

    /* Without whole-image statistics (saturation boost, histogram, normalisation) and without
     * error diffusion the preprocessing is elementwise: the scaler applies it as it writes the
     * pixmap, and pass 2 is not launched. */
    c->prep_fused = !pp.boost_saturation && !pp.build_histogram && !pp.normalize
        && c->dither_mode != CHAFA_DITHER_MODE_DIFFUSION
        && (c->dither_mode != CHAFA_DITHER_MODE_NONE || pp.to_din99d);
    c->scale_post_on = c->prep_fused;
    if (c->prep_fused)
    {
        c->scale_post.dither_mode = pp.dither_mode;
        c->scale_post.grain_w_shift = pp.grain_w_shift;
        c->scale_post.grain_h_shift = pp.grain_h_shift;
        c->scale_post.tex_size_shift = pp.tex_size_shift;
        c->scale_post.tex_size_mask = pp.tex_size_mask;
        c->scale_post.to_din99d = pp.to_din99d;
        c->scale_post.texture = pp.texture;
        c->scale_post.din99d_lut = pp.din99d_lut;
    }
    bool scaled = canvas_run_scaler (c, type, d_src, src_w, src_h, rowstride, px, py, pw, ph, c->work_factor_int < 3,
                                     first_px_row, n_px_rows);
    c->scale_post_on = false;
    if (!scaled)
        return false;
    fill_prep (c, &pp, first_px_row, n_px_rows);      /* the pixmap may have been (re)allocated by the scaler */

    c->timer.begin (ST_PREP, c->stream);
    if (pp.boost_saturation || pp.build_histogram)
    {
        if (pp.build_histogram && !CU (cudaMemsetAsync (c->d_hist, 0, 2051 * 4, c->stream)))
            return false;
        if (!CU ((cudaError_t) dev_launch_prep_pass1 (&pp, c->stream)))
            return false;
    }
    return true;
}

/* Second half: crop bounds from the (possibly all-reduced) histogram, normalise / dither /
 * convert, match, resolve (chafa-pixops.c:568-670, chafa-symbol-renderer.c:797-917) */
static bool
draw_symbols_phase2 (ChafaCanvas *c)
{
    const ChafaCanvasConfig &cfg = c->config;
    int first_px_row, n_px_rows;
    DevPrep pp;

    band_pixel_rows (c, &first_px_row, &n_px_rows);
    fill_prep (c, &pp, first_px_row, n_px_rows);
    bool do_norm = pp.normalize != 0;

    size_t n_cells = (size_t) cfg.width * cfg.height;
    if (!c->d_evals.ensure (n_cells * sizeof (DevCellEval)))
        return false;
    bool use_wide = c->d_syms.n2 > 0 && cfg.width > 1;
    if (use_wide && !c->d_wide.ensure (n_cells * sizeof (DevWideEval)))
        return false;

    if (pp.build_histogram && !CU ((cudaError_t) dev_launch_prep_bounds (&pp, c->stream)))
        return false;

    bool need_pass2 = !c->prep_fused && (do_norm || pp.to_din99d || c->dither_mode != CHAFA_DITHER_MODE_NONE);
    if (need_pass2)
    {
        if (c->dither_mode == CHAFA_DITHER_MODE_DIFFUSION)
        {
            /* Normalisation and colour-space conversion first, then the serial
             * diffusion chain over the whole pixmap (chafa-pixops.c:568-625, 660-667).
             * Diffusion ignores bands: every band owner runs the full chain. */
            DevPrep p2 = pp;
            p2.dither_mode = CHAFA_DITHER_MODE_NONE;
            p2.first_row = 0;
            p2.n_rows = c->height_px;
            if ((do_norm || pp.to_din99d) && !CU ((cudaError_t) dev_launch_prep_pass2 (&p2, c->stream)))
                return false;

            DevFsDither fs;
            memset (&fs, 0, sizeof (fs));
            if (!c->d_fs_scratch.ensure (dev_fs_dither_symbols_scratch_bytes (c->width_px, c->grain_w_shift)))
                return false;
            fs.pixels = c->d_pixels.as<uint32_t> ();
            fs.width = c->width_px;
            fs.height = c->height_px;
            fs.grain_w_shift = c->grain_w_shift;
            fs.grain_h_shift = c->grain_h_shift;
            fs.intensity = c->dither_intensity;
            fs.color_space = cfg.color_space;
            fs.palette = c->d_fg_palette;
            fs.scratch = c->d_fs_scratch.p;
            if (!CU ((cudaError_t) dev_launch_fs_dither_symbols (&fs, c->stream)))
                return false;
        }
        else if (!CU ((cudaError_t) dev_launch_prep_pass2 (&pp, c->stream)))
            return false;
    }

    c->timer.end (ST_PREP, c->stream);

    DevCanvas cv;
    fill_dev_canvas (c, &cv);
    HOSTPROF ("draw.match+resolve");
    if (c->use_tensor_match && dev_match_tc_applies (&cv, use_wide))
    {
        HOSTPROF ("draw.match_tc.launch");
        c->timer.begin (ST_MATCH, c->stream);
        if (!CU ((cudaError_t) dev_launch_match_tc (&cv, c->d_pixels.as<uint32_t> (), c->d_evals.as<DevCellEval> (),
                                                    use_wide ? c->d_wide.as<DevWideEval> () : nullptr, c->stream)))
            return false;
        c->timer.end (ST_MATCH, c->stream);
    }
    else if (c->use_tensor_match && dev_match_xtc_applies (&cv, use_wide))
    {
        c->timer.begin (ST_MATCH, c->stream);
        if (!CU ((cudaError_t) dev_launch_match_xtc (&cv, c->d_pixels.as<uint32_t> (), c->d_evals.as<DevCellEval> (),
                                                     use_wide ? c->d_wide.as<DevWideEval> () : nullptr, c->stream)))
            return false;
        c->timer.end (ST_MATCH, c->stream);
    }
    else if (dev_match_exhaustive_applies (&cv))
    {
        c->timer.begin (ST_MATCH, c->stream);
        if (!CU ((cudaError_t) dev_launch_match_exhaustive (&cv, c->d_pixels.as<uint32_t> (),
                                                            c->d_evals.as<DevCellEval> (),
                                                            use_wide ? c->d_wide.as<DevWideEval> () : nullptr,
                                                            c->stream)))
            return false;
        c->timer.end (ST_MATCH, c->stream);
    }
    else
    {
        c->timer.begin (ST_MATCH, c->stream);
        if (!CU ((cudaError_t) dev_launch_match (&cv, c->d_pixels.as<uint32_t> (), c->d_evals.as<DevCellEval> (),
                                                 c->stream)))
            return false;
        c->timer.end (ST_MATCH, c->stream);
        if (use_wide)
        {
            c->timer.begin (ST_MATCH_WIDE, c->stream);
            if (!CU ((cudaError_t) dev_launch_match_wide (&cv, c->d_pixels.as<uint32_t> (),
                                                          c->d_wide.as<DevWideEval> (), c->stream)))
                return false;
            c->timer.end (ST_MATCH_WIDE, c->stream);
        }
    }
    if (!sync_cells_to_device (c))
        return false;
    /* The row walk is left to whoever reads the cells first: normally the printer, in its own launch */
    c->resolve_args.enabled = 1;
    c->resolve_args.cv = cv;
    c->resolve_args.evals = c->d_evals.as<DevCellEval> ();
    c->resolve_args.wide_evals = use_wide ? c->d_wide.as<DevWideEval> () : nullptr;
    c->resolve_args.cells = c->d_cells.as<DevCell> ();
    c->resolve_pending = true;
    if (c->timer.enabled && !ensure_resolved (c))       /* per-stage timing wants it as a stage of its own */
        return false;

    c->have_pixels = true;
    c->h_cells_valid = false;
    c->d_cells_valid = true;
    c->needs_clear = false;
    return true;
}

static bool
host_memory_is_pageable (const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes (&attr, p) != cudaSuccess)
    {
        cudaGetLastError ();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

/* @n bytes from pageable @src to device @d_dst on the canvas's stream, through the pinned ring.
 * Returns when @src has been read. */
static bool
copy_through_ring (ChafaCanvas *c, uint8_t *d_dst, const uint8_t *src, size_t n)
{
    if (!c->h_ring.ensure (SRC_RING_SLOTS * SRC_RING_CHUNK))
        return false;
    for (int i = 0; i < SRC_RING_SLOTS; i++)
        if (!c->ring_free [i] && !CU (cudaEventCreateWithFlags (&c->ring_free [i], cudaEventDisableTiming)))
            return false;

    for (size_t off = 0; off < n; off += SRC_RING_CHUNK)
    {
        const size_t len = std::min (SRC_RING_CHUNK, n - off);
        const int slot = c->ring_next;
        uint8_t *stage = (uint8_t *) c->h_ring.p + (size_t) slot * SRC_RING_CHUNK;
        if (c->ring_busy [slot] && !CU (cudaEventSynchronize (c->ring_free [slot])))
            return false;
        memcpy (stage, src + off, len);
        if (!CU (cudaMemcpyAsync (d_dst + off, stage, len, cudaMemcpyHostToDevice, c->stream))
            || !CU (cudaEventRecord (c->ring_free [slot], c->stream)))
            return false;
        c->ring_busy [slot] = true;
        c->ring_next = (slot + 1) % SRC_RING_SLOTS;
    }
    return true;
}

static void
draw_all_pixels (ChafaCanvas *c, ChafaPixelType type, const void *src, bool src_on_device,
                 int src_w, int src_h, int rowstride, bool two_phase = false)
{
    HOSTPROF ("draw.total");
    if (src_w == 0 || src_h == 0)
        return;
    if (rowstride < src_w * bytes_per_pixel (type) || src_w > 65535 || src_h > 65535)
    {
        b200_set_error ("unsupported source geometry");
        return;
    }
    if (!canvas_bind_device (c))
        return;

    const void *d_src = src;
    bool src_staged = false;        /* the caller's buffer has been read in full already */

    if (!src_on_device)
    {
        size_t n = (size_t) (src_h - 1) * rowstride + (size_t) src_w * bytes_per_pixel (type);
        size_t first_byte = 0;

        /* A row band needs only the source rows its vertical filter reads */
        if (c->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS && c->band_rows < c->config.height
            && c->dither_mode != CHAFA_DITHER_MODE_DIFFUSION)
        {
            ScalePlan plan;
            int px, py, pw, ph, lo, hi;
            symbols_placement (c, src_w, src_h, &px, &py, &pw, &ph);
            if (scale_plan_build (&plan, src_w, src_h, c->width_px, c->height_px, px, py, pw, ph,
                                  c->work_factor_int < 3))
            {
                scale_plan_source_rows (&plan, c->band_first * 8, c->band_rows * 8, &lo, &hi);
                first_byte = (size_t) lo * rowstride;
                size_t last_byte = std::min (n, (size_t) (hi + 1) * rowstride);
                n = last_byte > first_byte ? last_byte - first_byte : 0;
            }
        }

        else if (c->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS && c->band_rows < c->config.height)
        {
            int lo, hi;
            if (sixel_source_rows (c, src_w, src_h, &lo, &hi))
            {
                first_byte = (size_t) lo * rowstride;
                size_t last_byte = std::min (n, (size_t) (hi + 1) * rowstride);
                n = last_byte > first_byte ? last_byte - first_byte : 0;
            }
        }

        if (!c->d_src.ensure ((size_t) (src_h - 1) * rowstride + (size_t) src_w * bytes_per_pixel (type) + 16))
            return;
        c->timer.begin (ST_H2D, c->stream);
        if (n && host_memory_is_pageable (src))
        {
            if (!copy_through_ring (c, (uint8_t *) c->d_src.p + first_byte, (const uint8_t *) src + first_byte, n))
                return;
            src_staged = true;
        }
        else if (n && !CU (cudaMemcpyAsync ((uint8_t *) c->d_src.p + first_byte, (const uint8_t *) src + first_byte, n,
                                            cudaMemcpyHostToDevice, c->stream)))
            return;
        c->timer.end (ST_H2D, c->stream);
        d_src = c->d_src.p;
    }

    bool ok;
    if (c->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS)
        ok = draw_symbols_phase1 (c, type, d_src, src_w, src_h, rowstride) && (two_phase || draw_symbols_phase2 (c));
    else if (c->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS)
        ok = two_phase ? sixel_draw_begin (c, d_src, type, src_w, src_h, rowstride, &c->sixel_reduce)
                       : sixel_draw (c, d_src, type, src_w, src_h, rowstride);
    else
        ok = pixelimg_draw (c, d_src, type, src_w, src_h, rowstride);
    if (!ok && c->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS)
    {
        /* A draw that failed on the device (the reason is in chafa_b200_last_error) must not
         * leave the cells of an earlier frame to be printed as if they were this one's */
        c->h_cells.assign (c->h_cells.size (), DevCell { ' ', 0, 0 });
        c->h_cells_valid = true;
        c->d_cells_valid = false;
        c->have_pixels = false;
    }

    /* The source is borrowed for the duration of the call only, and the staging
     * buffer is reused by the next draw: wait until the copies have been consumed.
     * The kernels keep running. */
    if ((src_on_device || src_staged) && c->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS && !c->plan_uploaded_this_draw)
    {
        /* nothing of this draw reads host memory: no reason to wait for anything */
        c->src_event_pending = false;
        return;
    }
    if (!c->src_event_pending)
        CU (cudaEventRecord (c->src_consumed, c->stream));
    c->src_event_pending = false;
    CU (cudaEventSynchronize (c->src_consumed));
}

/* ------------------------------------------------------------------------ */
/* Printing                                                                   */

static const int k_dseq_map [DSEQ_MAX] =
{
    CHAFA_TERM_SEQ_RESET_ATTRIBUTES,
    CHAFA_TERM_SEQ_RESET_COLOR_FG,
    CHAFA_TERM_SEQ_INVERT_COLORS,
    CHAFA_TERM_SEQ_ENABLE_BOLD,
    CHAFA_TERM_SEQ_SET_COLOR_FG_DIRECT,
    CHAFA_TERM_SEQ_SET_COLOR_BG_DIRECT,
    CHAFA_TERM_SEQ_SET_COLOR_FGBG_DIRECT,
    CHAFA_TERM_SEQ_SET_COLOR_FG_256,
    CHAFA_TERM_SEQ_SET_COLOR_BG_256,
    CHAFA_TERM_SEQ_SET_COLOR_FGBG_256,
    CHAFA_TERM_SEQ_SET_COLOR_FG_16,
    CHAFA_TERM_SEQ_SET_COLOR_BG_16,
    CHAFA_TERM_SEQ_SET_COLOR_FGBG_16,
    CHAFA_TERM_SEQ_SET_COLOR_FG_8,
    CHAFA_TERM_SEQ_SET_COLOR_BG_8,
    CHAFA_TERM_SEQ_SET_COLOR_FGBG_8,
    CHAFA_TERM_SEQ_REPEAT_CHAR
};

/* Compile the sequences the printer needs; returns an upper bound on the bytes one
 * cell can produce */
static size_t
compile_term_info (const ChafaTermInfo *ti, DevTermInfo *out)
{
    size_t seq_max [DSEQ_MAX];

    memset (out, 0, sizeof (*out));
    for (int i = 0; i < DSEQ_MAX; i++)
    {
        const ParsedSeq &s = ti->seqs [k_dseq_map [i]];
        DevSeq &d = out->seqs [i];
        size_t total = 0;
        int n_args = 0;

        memcpy (d.str, s.str, DEV_SEQ_LEN_MAX);
        for (int a = 0; a < DEV_SEQ_ARGS_MAX; a++)
        {
            d.pre_len [a] = s.defined ? s.args [a].pre_len : 0;
            d.arg_index [a] = s.defined ? s.args [a].arg_index : 255;
            total += d.pre_len [a];
            if (d.arg_index [a] != 255)
                n_args++;
        }
        seq_max [i] = total + (size_t) n_args * 3;
    }

    size_t color_max = 0;
    for (int i = DSEQ_SET_COLOR_FG_DIRECT; i <= DSEQ_SET_COLOR_FGBG_8; i++)
        color_max = std::max (color_max, seq_max [i]);
    return seq_max [DSEQ_RESET_ATTRIBUTES] + seq_max [DSEQ_INVERT_COLORS] + seq_max [DSEQ_ENABLE_BOLD]
        + color_max + seq_max [DSEQ_REPEAT_CHAR] + 16;
}

struct PrintResult
{
    size_t len = 0;
    const uint8_t *d_out = nullptr;
    std::vector<uint64_t> row_ofs;    /* only when requested */
    const uint64_t *d_len = nullptr;  /* enqueue-only prints: device location of the length */
};

/* Queue measure + emit, learn the length; the text stays on the device */
static bool
print_symbols_device (ChafaCanvas *c, ChafaTermInfo *ti, PrintResult *res, bool want_row_ofs, bool enqueue_only = false)
{
    HOSTPROF ("print.total");
    const ChafaCanvasConfig &cfg = c->config;

    if (!canvas_bind_device (c) || !sync_cells_to_device (c))
        return false;

    int n_rows = c->band_rows;
    size_t n_cells = (size_t) cfg.width * n_rows;

    if (!c->h_print.ensure (sizeof (DevTermInfo) + 16 + (size_t) (n_rows + 1) * 8)
        || !c->d_ti.ensure (sizeof (DevTermInfo)))
        return false;

    DevTermInfo *h_ti = c->h_print.as<DevTermInfo> ();
    size_t cell_max;
    {
        HOSTPROF ("print.compile_term_info");
        cell_max = compile_term_info (ti, h_ti);
    }
    size_t reset_max = std::max (strlen ((const char *) h_ti->seqs [DSEQ_RESET_ATTRIBUTES].str),
                                 strlen ((const char *) h_ti->seqs [DSEQ_RESET_COLOR_FG].str));
    size_t capacity = n_cells * cell_max + (size_t) n_rows * (2 * reset_max + 1) + 16;

    if (!c->d_cell_len.ensure (n_cells * 4) || !c->d_row_ofs.ensure ((size_t) (n_rows + 1) * 8)
        || !c->d_out.ensure (capacity))
        return false;
    /* The compiled sequences rarely change between prints: upload them only when they do.  (The
     * staging copy is pageable-safe: the comparison copy lives in the canvas, not in the pinned area.) */
    if (!c->ti_uploaded || memcmp (&c->ti_last, h_ti, sizeof (DevTermInfo)) != 0)
    {
        if (!CU (cudaMemcpyAsync (c->d_ti.p, h_ti, sizeof (DevTermInfo), cudaMemcpyHostToDevice, c->stream))
            || !CU (cudaStreamSynchronize (c->stream)))
            return false;
        memcpy (&c->ti_last, h_ti, sizeof (DevTermInfo));
        c->ti_uploaded = true;
    }

    DevPrint p;
    memset (&p, 0, sizeof (p));
    p.cells = c->d_cells.as<DevCell> ();
    p.width = cfg.width;
    p.height = cfg.height;
    p.first_row = c->band_first;
    p.n_rows = n_rows;
    p.canvas_mode = cfg.canvas_mode;
    p.fg_only = cfg.fg_only_enabled;
    p.alpha_threshold = cfg.alpha_threshold;
    p.optimizations = cfg.optimizations;
    p.have_repeat = ti->seqs [CHAFA_TERM_SEQ_REPEAT_CHAR].defined
        && (cfg.optimizations & CHAFA_OPTIMIZATION_REPEAT_CELLS) != 0;
    p.fgbg_blank_symbol = symbol_map_has_symbol (&cfg.symbol_map, ' ') ? ' '
        : symbol_map_has_symbol (&cfg.symbol_map, 0x2588) ? 0x2588 : 0;
    p.total_height = cfg.height;
    p.ti = c->d_ti.as<DevTermInfo> ();
    p.cell_len = c->d_cell_len.as<uint32_t> ();
    p.row_ofs = c->d_row_ofs.as<uint64_t> ();
    p.out = c->d_out.as<uint8_t> ();
    p.out_capacity = c->d_out.cap;

    /* One launch (measure, look-back for the row offsets, emit).  Its control block is zeroed when
     * (re)allocated and whenever the 22-bit epoch wraps. */
    {
        void *before = c->d_print_ctl.p;
        if (!c->d_print_ctl.ensure (16 + (size_t) n_rows * 8))
            return false;
        c->print_epoch = (c->print_epoch + 1) & 0x3fffffu;
        if (c->d_print_ctl.p != before || c->print_epoch == 0)
        {
            if (!CU (cudaMemsetAsync (c->d_print_ctl.p, 0, c->d_print_ctl.cap, c->stream)))
                return false;
            c->print_epoch = 1;
        }
        p.ctrl = c->d_print_ctl.as<uint32_t> ();
        p.status = (unsigned long long *) (c->d_print_ctl.as<uint8_t> () + 16);
        p.epoch = c->print_epoch;
        size_t row_bound = (size_t) cfg.width * cell_max + 2 * reset_max + 1;
        p.stage_bytes = dev_print_fused_smem_bytes (cfg.width, row_bound) <= 100 * 1024 ? (uint32_t) row_bound : 0;
    }
    c->timer.begin (ST_PRINT_EMIT, c->stream);
    {
        HOSTPROF ("print.launch");
        const bool with_walk = c->resolve_pending;
        c->resolve_pending = false;
        if (!CU ((cudaError_t) dev_launch_print_fused (&p, with_walk ? &c->resolve_args : nullptr, c->stream)))
            return false;
    }
    c->timer.end (ST_PRINT_EMIT, c->stream);

    if (enqueue_only)
    {
        /* nothing is read back: the caller consumes text and length on the device, in stream order */
        res->len = 0;
        res->d_out = c->d_out.as<uint8_t> ();
        res->d_len = c->d_row_ofs.as<uint64_t> () + n_rows;
        return true;
    }

    uint64_t *h_ofs = (uint64_t *) (c->h_print.as<uint8_t> () + ((sizeof (DevTermInfo) + 15) & ~(size_t) 15));
    if (want_row_ofs)
    {
        if (!CU (cudaMemcpyAsync (h_ofs, c->d_row_ofs.p, (size_t) (n_rows + 1) * 8, cudaMemcpyDeviceToHost,
                                  c->stream)))
            return false;
    }
    else if (!CU (cudaMemcpyAsync (h_ofs + n_rows, c->d_row_ofs.as<uint64_t> () + n_rows, 8,
                                   cudaMemcpyDeviceToHost, c->stream)))
        return false;
    if (!CU (cudaStreamSynchronize (c->stream)))
        return false;

    res->len = (size_t) h_ofs [n_rows];
    res->d_out = c->d_out.as<uint8_t> ();
    if (res->len > c->d_out.cap)
    {
        b200_set_error ("printed output exceeded its bound");
        return false;
    }
    if (want_row_ofs)
        res->row_ofs.assign (h_ofs, h_ofs + n_rows + 1);
    return true;
}

static ChafaTermInfo *
resolve_term_info (ChafaTermInfo *ti)
{
    if (ti)
    {
        chafa_term_info_ref (ti);
        return ti;
    }
    return chafa_term_db_get_fallback_info (chafa_term_db_get_default ());
}

static GString *
empty_string ()
{
    return b200_string_new_copy ("", 0);
}

/* ------------------------------------------------------------------------ */
/* Internal accessors for host_sixel.cpp                                      */

cudaStream_t canvas_stream (ChafaCanvas *c) { return c->stream; }
uint32_t *canvas_pixels (ChafaCanvas *c) { return c->d_pixels.as<uint32_t> (); }
const DevPalette *canvas_host_palette (const ChafaCanvas *c) { return &c->h_fg_palette; }
const DevPalette *canvas_device_palette (const ChafaCanvas *c) { return c->d_fg_palette; }
bool canvas_cuda_ok (int err, const char *what) { return cuda_ok ((cudaError_t) err, what); }
const ChafaCanvasConfig *canvas_config (const ChafaCanvas *c) { return &c->config; }
void **canvas_sixel_slot (ChafaCanvas *c) { return &c->sixel; }
void **canvas_pixelimg_slot (ChafaCanvas *c) { return &c->pixelimg; }
void canvas_set_scale_composite (ChafaCanvas *c, int composite) { c->scale_composite = composite; }
int canvas_width_px (const ChafaCanvas *c) { return c->width_px; }
int canvas_height_px (const ChafaCanvas *c) { return c->height_px; }
/* the band as pixel rows [first, end) of the canvas */
void
canvas_band_px (const ChafaCanvas *c, int *first_out, int *end_out)
{
    int cell_h = c->config.height > 0 ? c->height_px / c->config.height : 1;
    *first_out = c->band_first * cell_h;
    *end_out = std::min (c->height_px, (c->band_first + c->band_rows) * cell_h);
}
int canvas_device (const ChafaCanvas *c) { return c->device; }
const ChafaPlacement *canvas_placement (const ChafaCanvas *c) { return c->placement; }
int canvas_work_factor_int (const ChafaCanvas *c) { return c->work_factor_int; }
void canvas_dither_params (const ChafaCanvas *c, int *mode, float *intensity, int *gws, int *ghs, int *tss, int *tsm,
                           const int32_t **d_texture)
{
    *mode = c->dither_mode;
    *intensity = c->dither_intensity;
    *gws = c->grain_w_shift;
    *ghs = c->grain_h_shift;
    *tss = c->tex_size_shift;
    *tsm = c->tex_size_mask;
    *d_texture = c->d_texture;
}
const void *canvas_scale_tables (const ChafaCanvas *c, const uint16_t **from_srgb, const uint8_t **to_srgb,
                                 const uint32_t **inv_div)
{
    const DeviceTables *t = device_tables (c->device);
    if (!t)
        return nullptr;
    *from_srgb = t->from_srgb;
    *to_srgb = t->to_srgb;
    *inv_div = t->inv_div;
    return t;
}

/* ------------------------------------------------------------------------ */
/* Public API                                                                 */

extern "C" {

int
chafa_b200_available (void)
{
    int n = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess || n <= 0)
    {
        cudaGetLastError ();
        return 0;
    }
    return 1;
}

int
chafa_b200_set_device (int device)
{
    int n = 0;
    if (cudaGetDeviceCount (&n) != cudaSuccess || device < 0 || device >= n)
    {
        cudaGetLastError ();
        return 0;
    }
    t_device = device;
    return 1;
}

int
chafa_b200_copy_from_device (void *host_dst, const void *device_src, size_t n_bytes)
{
    if (!n_bytes)
        return 1;
    return CU (cudaMemcpy (host_dst, device_src, n_bytes, cudaMemcpyDeviceToHost)) ? 1 : 0;
}

uint64_t
chafa_b200_staged_scale_count (void)
{
    return dev_scale_staged_launches ();
}

int
chafa_b200_canvas_synchronize (ChafaCanvas *canvas)
{
    CHECK_CANVAS_VAL (canvas, 0);
    if (!canvas_bind_device (canvas))
        return 0;
    return CU (cudaStreamSynchronize (canvas->stream)) ? 1 : 0;
}

void *
chafa_b200_device_alloc (size_t n_bytes)
{
    void *p = nullptr;
    if (!n_bytes || !CU ((cudaError_t) b200_dev_malloc (&p, n_bytes)))
        return nullptr;
    return p;
}

void
chafa_b200_device_free (void *d_ptr)
{
    if (d_ptr)
        b200_dev_free (d_ptr);
}

int
chafa_b200_copy_to_device (void *device_dst, const void *host_src, size_t n_bytes)
{
    if (!n_bytes)
        return 1;
    return CU (cudaMemcpy (device_dst, host_src, n_bytes, cudaMemcpyHostToDevice)) ? 1 : 0;
}

uint64_t
chafa_b200_launch_count (void)
{
    return g_launch_count.load ();
}

/* test hooks of the guard-band self-check (tests/test_guard_gpu.py) */
void *
chafa_b200_debug_cells_address (ChafaCanvas *canvas)
{
    if (canvas)
        ensure_resolved (canvas);
    return canvas ? canvas->d_cells.p : nullptr;
}

void
chafa_b200_debug_poke (void *d_ptr, int value)
{
    if (d_ptr)
        CU (cudaMemset (d_ptr, value, 1));
}

int
chafa_b200_get_din99d_table (uint32_t *host_out)
{
    int device = t_device >= 0 ? t_device : 0;
    if (!host_out || !CU (cudaSetDevice (device)))
        return 0;
    const uint32_t *lut = dev_din99d_lut (device, nullptr);
    if (!lut)
        return 0;
    return CU (cudaMemcpy (host_out, lut, (size_t) 4 << 24, cudaMemcpyDeviceToHost)) ? 1 : 0;
}

ChafaCanvas *
chafa_canvas_new (const ChafaCanvasConfig *config)
{
    if (config)
    {
        RETURN_VAL_IF_FAIL (config->width > 0, NULL);
        RETURN_VAL_IF_FAIL (config->height > 0, NULL);
    }

    if (!chafa_b200_available ())
    {
        b200_set_error ("no CUDA device: the B200 canvas renderer has no CPU fallback");
        fprintf (stderr, "chafa-b200: no CUDA device available; chafa_canvas_new () fails (no CPU fallback)\n");
        return NULL;
    }

    ChafaCanvas *c = new ChafaCanvas ();
    if (config)
    {
        ChafaCanvasConfig copy (*config);
        /* member-wise assignment keeps c->config.refs */
        c->config.width = copy.width;
        c->config.height = copy.height;
        c->config.cell_width = copy.cell_width;
        c->config.cell_height = copy.cell_height;
        c->config.canvas_mode = copy.canvas_mode;
        c->config.color_space = copy.color_space;
        c->config.dither_mode = copy.dither_mode;
        c->config.color_extractor = copy.color_extractor;
        c->config.pixel_mode = copy.pixel_mode;
        c->config.dither_grain_width = copy.dither_grain_width;
        c->config.dither_grain_height = copy.dither_grain_height;
        c->config.dither_intensity = copy.dither_intensity;
        c->config.fg_color_packed_rgb = copy.fg_color_packed_rgb;
        c->config.bg_color_packed_rgb = copy.bg_color_packed_rgb;
        c->config.alpha_threshold = copy.alpha_threshold;
        c->config.work_factor = copy.work_factor;
        c->config.symbol_map = copy.symbol_map;
        c->config.fill_symbol_map = copy.fill_symbol_map;
        c->config.preprocessing_enabled = copy.preprocessing_enabled;
        c->config.fg_only_enabled = copy.fg_only_enabled;
        c->config.optimizations = copy.optimizations;
        c->config.passthrough = copy.passthrough;
    }

    if (t_device >= 0)
        c->device = t_device;
    else if (cudaGetDevice (&c->device) != cudaSuccess)
        c->device = 0;

    if (!canvas_setup (c))
    {
        canvas_free_device (c);
        delete c;
        return NULL;
    }
    return c;
}

ChafaCanvas *
chafa_canvas_new_similar (ChafaCanvas *orig)
{
    RETURN_VAL_IF_FAIL (orig != NULL, NULL);
    int saved = t_device;
    t_device = orig->device;
    ChafaCanvas *c = chafa_canvas_new (&orig->config);
    t_device = saved;
    return c;
}

void
chafa_canvas_ref (ChafaCanvas *canvas)
{
    CHECK_CANVAS (canvas);
    canvas->refs++;
}

void
chafa_canvas_unref (ChafaCanvas *canvas)
{
    CHECK_CANVAS (canvas);
    if (--canvas->refs == 0)
    {
        if (canvas->placement)
            chafa_placement_unref (canvas->placement);
        if (canvas->sixel)
            sixel_state_free (canvas->sixel);
        if (canvas->pixelimg)
            pixelimg_state_free (canvas->pixelimg);
        canvas_free_device (canvas);
        delete canvas;
    }
}

const ChafaCanvasConfig *
chafa_canvas_peek_config (ChafaCanvas *canvas)
{
    CHECK_CANVAS_VAL (canvas, NULL);
    return &canvas->config;
}

void
chafa_canvas_set_placement (ChafaCanvas *canvas, ChafaPlacement *placement)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (placement != NULL);

    chafa_placement_ref (placement);
    if (canvas->placement)
        chafa_placement_unref (canvas->placement);
    canvas->placement = placement;

    ChafaImage *image = placement->image;
    if (!image || !image->frame)
        return;
    ChafaFrame *frame = image->frame;
    draw_all_pixels (canvas, frame->pixel_type, frame->data, false, frame->width, frame->height, frame->rowstride);
}

void
chafa_canvas_draw_all_pixels (ChafaCanvas *canvas, ChafaPixelType src_pixel_type, const guint8 *src_pixels,
                              gint src_width, gint src_height, gint src_rowstride)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (src_pixel_type < CHAFA_PIXEL_MAX);
    RETURN_IF_FAIL (src_pixels != NULL);
    RETURN_IF_FAIL (src_width >= 0);
    RETURN_IF_FAIL (src_height >= 0);

    draw_all_pixels (canvas, src_pixel_type, src_pixels, false, src_width, src_height, src_rowstride);
}

void
chafa_canvas_set_contents_rgba8 (ChafaCanvas *canvas, const guint8 *src_pixels, gint src_width, gint src_height,
                                 gint src_rowstride)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (src_pixels != NULL);
    draw_all_pixels (canvas, CHAFA_PIXEL_RGBA8_UNASSOCIATED, src_pixels, false, src_width, src_height,
                     src_rowstride);
}

void
chafa_b200_canvas_set_stream (ChafaCanvas *canvas, void *cuda_stream)
{
    CHECK_CANVAS (canvas);
    if (!canvas_bind_device (canvas))
        return;
    if (canvas->stream)
        cudaStreamSynchronize (canvas->stream);
    if (canvas->own_stream && canvas->stream)
        cudaStreamDestroy (canvas->stream);
    canvas->stream = (cudaStream_t) cuda_stream;
    canvas->own_stream = false;
}

void
chafa_b200_canvas_draw_all_pixels_device (ChafaCanvas *canvas, ChafaPixelType src_pixel_type,
                                          const void *d_src_pixels, int src_width, int src_height,
                                          int src_rowstride)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (src_pixel_type < CHAFA_PIXEL_MAX);
    RETURN_IF_FAIL (d_src_pixels != NULL);
    RETURN_IF_FAIL (src_width >= 0);
    RETURN_IF_FAIL (src_height >= 0);

    draw_all_pixels (canvas, src_pixel_type, d_src_pixels, true, src_width, src_height, src_rowstride);
}

void
chafa_b200_canvas_set_tensor_match (ChafaCanvas *canvas, int enabled)
{
    CHECK_CANVAS (canvas);
    canvas->use_tensor_match = enabled != 0;
}

void
chafa_b200_canvas_set_kernel_timing (ChafaCanvas *canvas, int enabled)
{
    CHECK_CANVAS (canvas);
    if (!canvas_bind_device (canvas))
        return;
    canvas->timer.collect ();
    canvas->timer.enabled = enabled != 0;
    for (int i = 0; i < ST_MAX; i++)
    {
        canvas->timer.ms [i] = 0.0;
        canvas->timer.count [i] = 0;
    }
}

int
chafa_b200_canvas_get_kernel_times (ChafaCanvas *canvas, double *ms_out, uint64_t *count_out, int max_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    if (!canvas_bind_device (canvas))
        return 0;
    canvas->timer.collect ();
    for (int i = 0; i < ST_MAX && i < max_out; i++)
    {
        if (ms_out) ms_out [i] = canvas->timer.ms [i];
        if (count_out) count_out [i] = canvas->timer.count [i];
    }
    return ST_MAX;
}

int
chafa_b200_canvas_draw_begin_device (ChafaCanvas *canvas, ChafaPixelType src_pixel_type, const void *d_src_pixels,
                                     int src_width, int src_height, int src_rowstride, void **d_histogram_out,
                                     size_t *n_int32_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (src_pixel_type < CHAFA_PIXEL_MAX, 0);
    RETURN_VAL_IF_FAIL (d_src_pixels != NULL, 0);
    RETURN_VAL_IF_FAIL (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS
                        || canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS, 0);

    if (d_histogram_out) *d_histogram_out = NULL;
    if (n_int32_out) *n_int32_out = 0;
    canvas->sixel_reduce = false;
    draw_all_pixels (canvas, src_pixel_type, d_src_pixels, true, src_width, src_height, src_rowstride, true);

    if (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS)
    {
        /* Generated palette, row band, no diffusion: the colour-cube bins of the band's samples
         * (chafa-palette.c:513-536) are to be summed with the other owners' */
        if (!canvas->sixel_reduce)
            return 0;
        size_t n_bytes, n_words;
        void *buf = sixel_exchange_buffer (canvas, &n_bytes, &n_words);
        if (d_histogram_out) *d_histogram_out = buf;
        if (n_int32_out) *n_int32_out = n_words;
        return buf ? 1 : 0;
    }

    DevPrep pp;
    fill_prep (canvas, &pp, 0, 0);
    /* With error diffusion every band owner processes the whole image: its histogram is
     * already the global one */
    bool reduce = pp.build_histogram && canvas->dither_mode != CHAFA_DITHER_MODE_DIFFUSION
        && canvas->band_rows < canvas->config.height;
    if (d_histogram_out) *d_histogram_out = reduce ? (void *) canvas->d_hist : NULL;
    if (n_int32_out) *n_int32_out = reduce ? 2049 : 0;
    return reduce ? 1 : 0;
}

static int
draw_batch (ChafaCanvas **canvases, int n_canvases, ChafaPixelType src_pixel_type, const void *const *src_pixels,
            bool src_on_device, int src_width, int src_height, int src_rowstride)
{
    RETURN_VAL_IF_FAIL (canvases != NULL && src_pixels != NULL && n_canvases >= 0, 0);
    RETURN_VAL_IF_FAIL (src_pixel_type < CHAFA_PIXEL_MAX, 0);

    for (int i = 0; i < n_canvases; i++)
    {
        CHECK_CANVAS_VAL (canvases [i], 0);
        RETURN_VAL_IF_FAIL (src_pixels [i] != NULL, 0);
        RETURN_VAL_IF_FAIL (canvases [i]->device == canvases [0]->device, 0);
    }
    for (int i = 0; i < n_canvases; i++)
    {
        const bool sixels = canvases [i]->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS;
        if (sixels)
            sixel_set_defer_chain (canvases [i], true);
        draw_all_pixels (canvases [i], src_pixel_type, src_pixels [i], src_on_device, src_width, src_height,
                         src_rowstride);
        if (sixels)
            sixel_set_defer_chain (canvases [i], false);
    }
    if (n_canvases > 0 && !canvas_bind_device (canvases [0]))
        return 0;
    return sixel_launch_deferred_chains (canvases, n_canvases) >= 0 ? 1 : 0;
}

int
chafa_b200_canvases_draw_batch_device (ChafaCanvas **canvases, int n_canvases, ChafaPixelType src_pixel_type,
                                       const void *const *d_src_pixels, int src_width, int src_height,
                                       int src_rowstride)
{
    return draw_batch (canvases, n_canvases, src_pixel_type, d_src_pixels, true, src_width, src_height, src_rowstride);
}

int
chafa_b200_canvases_draw_batch (ChafaCanvas **canvases, int n_canvases, ChafaPixelType src_pixel_type,
                                const void *const *src_pixels, int src_width, int src_height, int src_rowstride)
{
    return draw_batch (canvases, n_canvases, src_pixel_type, src_pixels, false, src_width, src_height, src_rowstride);
}

int
chafa_b200_canvas_draw_float_sums_device (ChafaCanvas *canvas, void **d_sums_out, size_t *n_bytes_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS, 0);
    if (d_sums_out) *d_sums_out = NULL;
    if (n_bytes_out) *n_bytes_out = 0;
    if (!canvas->sixel_reduce || !canvas_bind_device (canvas) || !sixel_draw_float_bins (canvas))
        return 0;
    size_t n_bytes, n_words;
    uint8_t *buf = (uint8_t *) sixel_exchange_buffer (canvas, &n_bytes, &n_words);
    if (!buf)
        return 0;
    if (d_sums_out) *d_sums_out = buf + n_words * 4;
    if (n_bytes_out) *n_bytes_out = n_bytes - n_words * 4;
    return 1;
}

size_t
chafa_b200_canvas_exchange_buffer_bytes (ChafaCanvas *canvas)
{
    CHECK_CANVAS_VAL (canvas, 0);
    if (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS)
        return sixel_exchange_bytes (canvas);
    return 2051 * sizeof (int32_t);
}

void
chafa_b200_canvas_set_histogram_buffer (ChafaCanvas *canvas, void *d_int32_2051)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (d_int32_2051 != NULL);
    if (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS)
    {
        sixel_set_exchange_buffer (canvas, d_int32_2051);
        return;
    }
    canvas->d_hist = (int32_t *) d_int32_2051;
}

void
chafa_b200_canvas_draw_finish (ChafaCanvas *canvas)
{
    CHECK_CANVAS (canvas);
    if (!canvas_bind_device (canvas))
        return;
    if (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS)
    {
        sixel_draw_finish (canvas);
        canvas->sixel_reduce = false;
        return;
    }
    draw_symbols_phase2 (canvas);
}

void
chafa_b200_canvas_set_band (ChafaCanvas *canvas, int first_row, int n_rows)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (first_row >= 0 && n_rows >= 0 && first_row + n_rows <= canvas->config.height);
    ensure_resolved (canvas);       /* a row walk still owed belongs to the band that was drawn */
    canvas->band_first = first_row;
    canvas->band_rows = n_rows;
}

GString *
chafa_canvas_print (ChafaCanvas *canvas, ChafaTermInfo *term_info)
{
    CHECK_CANVAS_VAL (canvas, NULL);

    ChafaTermInfo *ti = resolve_term_info (term_info);
    GString *str = nullptr;

    canvas_bind_device (canvas);

    if (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS)
    {
        PrintResult res;
        if (print_symbols_device (canvas, ti, &res, false))
        {
            /* D2H into pinned staging (a pageable destination would go through the
             * driver's bounce buffers, several times slower), then one memcpy into the
             * malloc'ed GString the caller will free. */
            char *buf = (char *) b200_malloc (res.len + 1);
            bool copied = res.len == 0;
            if (!copied && canvas->h_text.ensure (std::max (res.len, canvas->d_out.cap)))
            {
                canvas->timer.begin (ST_D2H, canvas->stream);
                copied = CU (cudaMemcpyAsync (canvas->h_text.p, res.d_out, res.len, cudaMemcpyDeviceToHost,
                                              canvas->stream));
                canvas->timer.end (ST_D2H, canvas->stream);
                copied = copied && CU (cudaStreamSynchronize (canvas->stream));
                if (copied)
                    memcpy (buf, canvas->h_text.p, res.len);
            }
            if (copied)
            {
                buf [res.len] = '\0';
                str = b200_string_new_take (buf, res.len, res.len + 1);
            }
            else
                b200_free (buf);
        }
    }
    else if (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SIXELS
             && ti->seqs [CHAFA_TERM_SEQ_BEGIN_SIXELS].defined && canvas->sixel)
    {
        str = sixel_print (canvas, ti);
    }
    else if ((canvas->config.pixel_mode == CHAFA_PIXEL_MODE_KITTY
              && ti->seqs [CHAFA_TERM_SEQ_BEGIN_KITTY_IMMEDIATE_IMAGE_V1].defined && canvas->pixelimg)
             || (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_ITERM2 && canvas->pixelimg))
    {
        str = pixelimg_print (canvas, ti);
    }

    if (!str)
        str = empty_string ();
    chafa_term_info_unref (ti);
    return str;
}

size_t
chafa_b200_canvas_print_device (ChafaCanvas *canvas, ChafaTermInfo *term_info, void **d_out_ptr)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (d_out_ptr != NULL, 0);
    RETURN_VAL_IF_FAIL (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS, 0);

    ChafaTermInfo *ti = resolve_term_info (term_info);
    PrintResult res;
    size_t len = 0;

    *d_out_ptr = NULL;
    if (print_symbols_device (canvas, ti, &res, false))
    {
        *d_out_ptr = (void *) res.d_out;
        len = res.len;
    }
    chafa_term_info_unref (ti);
    return len;
}

int
chafa_b200_canvas_print_device_enqueue (ChafaCanvas *canvas, ChafaTermInfo *term_info, void **d_out_ptr,
                                        const unsigned long long **d_len_ptr)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (d_out_ptr != NULL, 0);
    RETURN_VAL_IF_FAIL (d_len_ptr != NULL, 0);
    RETURN_VAL_IF_FAIL (canvas->config.pixel_mode == CHAFA_PIXEL_MODE_SYMBOLS, 0);

    ChafaTermInfo *ti = resolve_term_info (term_info);
    PrintResult res;
    int ok = 0;

    *d_out_ptr = NULL;
    *d_len_ptr = NULL;
    if (print_symbols_device (canvas, ti, &res, false, true))
    {
        *d_out_ptr = (void *) res.d_out;
        *d_len_ptr = (const unsigned long long *) res.d_len;
        canvas->last_text = res.d_out;
        canvas->last_text_len = res.d_len;
        ok = 1;
    }
    chafa_term_info_unref (ti);
    return ok;
}

size_t
chafa_b200_canvas_print_into (ChafaCanvas *canvas, ChafaTermInfo *term_info, void *d_dest, size_t capacity)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (d_dest != NULL, 0);

    void *d_text = NULL;
    size_t len = chafa_b200_canvas_print_device (canvas, term_info, &d_text);
    if (!len || len > capacity)
        return len;
    if (!CU (cudaMemcpyAsync (d_dest, d_text, len, cudaMemcpyDeviceToDevice, canvas->stream))
        || !CU (cudaStreamSynchronize (canvas->stream)))
        return 0;
    return len;
}

/* ---- one still rendered as row bands by several processes: assembling the text ---- */

int
chafa_b200_ipc_export (const void *d_ptr, void *handle_out_64)
{
    RETURN_VAL_IF_FAIL (d_ptr != NULL && handle_out_64 != NULL, 0);
    static_assert (sizeof (cudaIpcMemHandle_t) == 64, "IPC handle size");
    return CU (cudaIpcGetMemHandle ((cudaIpcMemHandle_t *) handle_out_64, (void *) d_ptr)) ? 1 : 0;
}

void *
chafa_b200_ipc_open (const void *handle_64)
{
    RETURN_VAL_IF_FAIL (handle_64 != NULL, NULL);
    int device = t_device >= 0 ? t_device : 0;
    void *p = nullptr;
    cudaIpcMemHandle_t h;
    memcpy (&h, handle_64, sizeof (h));
    if (!CU (cudaSetDevice (device)) || !CU (cudaIpcOpenMemHandle (&p, h, cudaIpcMemLazyEnablePeerAccess)))
        return nullptr;
    return p;
}

void
chafa_b200_ipc_close (void *d_peer_ptr)
{
    if (d_peer_ptr)
        CU (cudaIpcCloseMemHandle (d_peer_ptr));
}

int
chafa_b200_canvas_copy_print_length (ChafaCanvas *canvas, unsigned long long *d_len_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (d_len_out != NULL && canvas->last_text_len != NULL, 0);
    if (!canvas_bind_device (canvas))
        return 0;
    return CU (cudaMemcpyAsync (d_len_out, canvas->last_text_len, 8, cudaMemcpyDeviceToDevice, canvas->stream)) ? 1 : 0;
}

int
chafa_b200_canvas_band_scatter (ChafaCanvas *canvas, const unsigned long long *d_lengths, int n_bands, int band,
                                void *d_full_text, size_t capacity, unsigned long long *d_total_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (d_lengths != NULL && d_full_text != NULL && canvas->last_text != NULL, 0);
    RETURN_VAL_IF_FAIL (n_bands > 0 && band >= 0 && band < n_bands, 0);
    if (!canvas_bind_device (canvas))
        return 0;
    return CU ((cudaError_t) dev_launch_band_scatter (canvas->last_text, (const uint64_t *) d_lengths, n_bands, band,
                                                      (uint8_t *) d_full_text, capacity, (uint64_t *) d_total_out,
                                                      canvas->stream)) ? 1 : 0;
}

GString *
chafa_canvas_build_ansi (ChafaCanvas *canvas)
{
    CHECK_CANVAS_VAL (canvas, NULL);
    return chafa_canvas_print (canvas, NULL);
}

void
chafa_canvas_print_rows (ChafaCanvas *canvas, ChafaTermInfo *term_info, GString ***array_out, gint *array_len_out)
{
    CHECK_CANVAS (canvas);
    RETURN_IF_FAIL (array_out != NULL);

    if (canvas->config.pixel_mode != CHAFA_PIXEL_MODE_SYMBOLS)
    {
        GString **array = (GString **) b200_malloc (2 * sizeof (GString *));
        array [0] = chafa_canvas_print (canvas, term_info);
        array [1] = NULL;
        *array_out = array;
        if (array_len_out)
            *array_len_out = 1;
        return;
    }

    ChafaTermInfo *ti = resolve_term_info (term_info);
    PrintResult res;
    int n_rows = canvas->band_rows;
    GString **array = (GString **) b200_malloc ((size_t) (n_rows + 1) * sizeof (GString *));
    std::vector<char> text;

    bool ok = print_symbols_device (canvas, ti, &res, true);
    if (ok && res.len)
    {
        text.resize (res.len);
        ok = CU (cudaMemcpyAsync (text.data (), res.d_out, res.len, cudaMemcpyDeviceToHost, canvas->stream))
            && CU (cudaStreamSynchronize (canvas->stream));
    }

    for (int r = 0; r < n_rows; r++)
    {
        if (!ok)
        {
            array [r] = empty_string ();
            continue;
        }
        size_t begin = (size_t) res.row_ofs [r], end = (size_t) res.row_ofs [r + 1];
        /* Rows carry no separator here: drop the newline the joined form has */
        if (canvas->band_first + r < canvas->config.height - 1 && end > begin)
            end--;
        array [r] = b200_string_new_copy (text.data () + begin, end - begin);
    }
    array [n_rows] = NULL;
    *array_out = array;
    if (array_len_out)
        *array_len_out = n_rows;
    chafa_term_info_unref (ti);
}

gchar **
chafa_canvas_print_rows_strv (ChafaCanvas *canvas, ChafaTermInfo *term_info)
{
    CHECK_CANVAS_VAL (canvas, NULL);

    GString **gsa = nullptr;
    gint n = 0;
    chafa_canvas_print_rows (canvas, term_info, &gsa, &n);
    if (!gsa)
        return NULL;

    gchar **strv = (gchar **) b200_malloc ((size_t) (n + 1) * sizeof (gchar *));
    for (int i = 0; i < n; i++)
    {
        /* steal the character data, free the GString shell */
        strv [i] = gsa [i]->str;
        gsa [i]->str = NULL;
        chafa_b200_string_free (gsa [i]);
    }
    strv [n] = NULL;
    b200_free (gsa);
    return strv;
}

/* ---- cell accessors (chafa/chafa-canvas.c:1048-1402) ---- */

static DevCell *
cell_at (ChafaCanvas *canvas, int x, int y, bool for_write)
{
    if (!sync_cells_to_host (canvas))
        return nullptr;
    if (for_write)
        canvas->d_cells_valid = false;
    return &canvas->h_cells [(size_t) y * canvas->config.width + x];
}

#define CHECK_XY(canvas, x, y, v) \
    do { \
        RETURN_VAL_IF_FAIL ((x) >= 0 && (x) < (canvas)->config.width, v); \
        RETURN_VAL_IF_FAIL ((y) >= 0 && (y) < (canvas)->config.height, v); \
    } while (0)

#define CHECK_XY_VOID(canvas, x, y) \
    do { \
        RETURN_IF_FAIL ((x) >= 0 && (x) < (canvas)->config.width); \
        RETURN_IF_FAIL ((y) >= 0 && (y) < (canvas)->config.height); \
    } while (0)

gunichar
chafa_canvas_get_char_at (ChafaCanvas *canvas, gint x, gint y)
{
    CHECK_CANVAS_VAL (canvas, 0);
    CHECK_XY (canvas, x, y, 0);
    DevCell *cell = cell_at (canvas, x, y, false);
    return cell ? cell->c : 0;
}

gint
chafa_canvas_set_char_at (ChafaCanvas *canvas, gint x, gint y, gunichar c)
{
    CHECK_CANVAS_VAL (canvas, 0);
    CHECK_XY (canvas, x, y, 0);

    if (!uc_isprint (c) || uc_iszerowidth (c))
        return 0;
    int cwidth = uc_iswide (c) ? 2 : 1;
    if (x + cwidth > canvas->config.width)
        return 0;

    DevCell *cell = cell_at (canvas, x, y, true);
    if (!cell)
        return 0;

    cell [0].c = c;
    if (cwidth == 2)
    {
        cell [1].c = 0;
        cell [1].fg = cell [0].fg;
        cell [1].bg = cell [0].bg;
    }
    /* Overwriting the right half of a wide character blanks its left half */
    if (x > 0 && cell [-1].c != 0 && uc_iswide (cell [-1].c))
        cell [-1].c = canvas->blank_char;
    return cwidth;
}

static int
argb_to_rgb_or_none (const ChafaCanvas *canvas, uint32_t argb)
{
    if ((int) (argb >> 24) < canvas->config.alpha_threshold)
        return -1;
    return (int) (argb & 0xffffffu);
}

static uint32_t
rgb_or_none_to_argb (int rgb)
{
    return rgb < 0 ? 0x00808080u : (0xff000000u | ((uint32_t) rgb & 0xffffffu));
}

static int
pen_to_rgb_or_none (const ChafaCanvas *canvas, const DevPalette &pal, uint32_t pen)
{
    if (pen == PAL_INDEX_BG || pen == PAL_INDEX_TRANSPARENT || pen >= PAL_INDEX_MAX)
        return -1;
    uint32_t col = pal.colors_rgb [pen];
    if ((int) (col >> 24) < canvas->config.alpha_threshold)
        return -1;
    return (int) (((col & 0xff) << 16) | (col & 0xff00) | ((col >> 16) & 0xff));
}

static bool
is_indexed_mode (ChafaCanvasMode m)
{
    return m == CHAFA_CANVAS_MODE_INDEXED_256 || m == CHAFA_CANVAS_MODE_INDEXED_240
        || m == CHAFA_CANVAS_MODE_INDEXED_16 || m == CHAFA_CANVAS_MODE_INDEXED_16_8
        || m == CHAFA_CANVAS_MODE_INDEXED_8;
}

void
chafa_canvas_get_colors_at (ChafaCanvas *canvas, gint x, gint y, gint *fg_out, gint *bg_out)
{
    CHECK_CANVAS (canvas);
    CHECK_XY_VOID (canvas, x, y);
    DevCell *cell = cell_at (canvas, x, y, false);
    if (!cell)
        return;

    gint fg, bg;
    if (canvas->config.canvas_mode == CHAFA_CANVAS_MODE_TRUECOLOR)
    {
        fg = argb_to_rgb_or_none (canvas, cell->fg);
        bg = argb_to_rgb_or_none (canvas, cell->bg);
    }
    else
    {
        fg = pen_to_rgb_or_none (canvas, canvas->h_fg_palette, cell->fg);
        bg = pen_to_rgb_or_none (canvas, canvas->h_bg_palette, cell->bg);
    }
    *fg_out = fg;
    *bg_out = bg;
}

void
chafa_canvas_get_raw_colors_at (ChafaCanvas *canvas, gint x, gint y, gint *fg_out, gint *bg_out)
{
    CHECK_CANVAS (canvas);
    CHECK_XY_VOID (canvas, x, y);
    DevCell *cell = cell_at (canvas, x, y, false);
    if (!cell)
        return;

    gint fg = -1, bg = -1;
    ChafaCanvasMode m = canvas->config.canvas_mode;
    if (m == CHAFA_CANVAS_MODE_TRUECOLOR)
    {
        fg = argb_to_rgb_or_none (canvas, cell->fg);
        bg = argb_to_rgb_or_none (canvas, cell->bg);
    }
    else if (is_indexed_mode (m))
    {
        fg = cell->fg < 256 ? (gint) cell->fg : -1;
        bg = cell->bg < 256 ? (gint) cell->bg : -1;
    }
    else if (m == CHAFA_CANVAS_MODE_FGBG_BGFG)
    {
        fg = cell->fg == PAL_INDEX_FG ? 0 : -1;
        bg = cell->bg == PAL_INDEX_FG ? 0 : -1;
    }
    else
    {
        fg = 0;
        bg = -1;
    }
    if (fg_out) *fg_out = fg;
    if (bg_out) *bg_out = bg;
}

static void
spread_to_wide_halves (ChafaCanvas *canvas, DevCell *cell, int x)
{
    if (x > 0 && cell->c == 0)
    {
        cell [-1].fg = cell->fg;
        cell [-1].bg = cell->bg;
    }
    if (x < canvas->config.width - 1 && cell [1].c == 0)
    {
        cell [1].fg = cell->fg;
        cell [1].bg = cell->bg;
    }
}

static void
set_colors (ChafaCanvas *canvas, int x, int y, int fg, int bg, bool raw)
{
    DevCell *cell = cell_at (canvas, x, y, true);
    if (!cell)
        return;

    ChafaCanvasMode m = canvas->config.canvas_mode;
    if (m == CHAFA_CANVAS_MODE_TRUECOLOR)
    {
        cell->fg = rgb_or_none_to_argb (fg);
        cell->bg = rgb_or_none_to_argb (bg);
    }
    else if (is_indexed_mode (m))
    {
        if (raw)
        {
            cell->fg = fg >= 0 ? (uint32_t) fg : PAL_INDEX_TRANSPARENT;
            cell->bg = bg >= 0 ? (uint32_t) bg : PAL_INDEX_TRANSPARENT;
        }
        else
        {
            int cs = canvas->config.color_space;
            auto to_pen = [&] (const DevPalette &pal, int rgb) -> uint32_t
            {
                if (rgb < 0)
                    return PAL_INDEX_TRANSPARENT;
                /* The reference hands the RGB value to the lookup as it is, whatever
                 * the canvas's colour space */
                uint32_t col = color_from_packed_rgb ((uint32_t) rgb, 0xff);
                return (uint32_t) host_palette_lookup_nearest (&pal, cs, col);
            };
            cell->fg = to_pen (canvas->h_fg_palette, fg);
            cell->bg = to_pen (canvas->h_bg_palette, bg);
        }
    }
    else if (m == CHAFA_CANVAS_MODE_FGBG_BGFG)
    {
        cell->fg = fg >= 0 ? PAL_INDEX_FG : PAL_INDEX_TRANSPARENT;
        cell->bg = bg >= 0 ? PAL_INDEX_FG : PAL_INDEX_TRANSPARENT;
    }
    else
    {
        cell->fg = fg >= 0 ? (uint32_t) fg : PAL_INDEX_TRANSPARENT;
    }
    spread_to_wide_halves (canvas, cell, x);
}

void
chafa_canvas_set_colors_at (ChafaCanvas *canvas, gint x, gint y, gint fg, gint bg)
{
    CHECK_CANVAS (canvas);
    CHECK_XY_VOID (canvas, x, y);
    set_colors (canvas, x, y, fg, bg, false);
}

void
chafa_canvas_set_raw_colors_at (ChafaCanvas *canvas, gint x, gint y, gint fg, gint bg)
{
    CHECK_CANVAS (canvas);
    CHECK_XY_VOID (canvas, x, y);
    set_colors (canvas, x, y, fg, bg, true);
}

/* ---- parity read-backs ---- */

int
chafa_b200_canvas_get_cells (ChafaCanvas *canvas, void *cells_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    RETURN_VAL_IF_FAIL (cells_out != NULL, 0);
    if (!sync_cells_to_host (canvas))
        return 0;
    memcpy (cells_out, canvas->h_cells.data (), canvas->h_cells.size () * sizeof (DevCell));
    return (int) canvas->h_cells.size ();
}

int
chafa_b200_canvas_get_prepared_pixels (ChafaCanvas *canvas, void *pixels_out, int *width_px_out, int *height_px_out)
{
    CHECK_CANVAS_VAL (canvas, 0);
    if (width_px_out) *width_px_out = canvas->width_px;
    if (height_px_out) *height_px_out = canvas->height_px;
    if (!canvas->have_pixels || !pixels_out)
        return 0;
    if (!canvas_bind_device (canvas))
        return 0;
    size_t n = (size_t) canvas->width_px * canvas->height_px * 4;
    if (!CU (cudaMemcpyAsync (pixels_out, canvas->d_pixels.p, n, cudaMemcpyDeviceToHost, canvas->stream))
        || !CU (cudaStreamSynchronize (canvas->stream)))
        return 0;
    return 1;
}

void
chafa_b200_canvas_get_info (ChafaCanvas *canvas, int *out)
{
    CHECK_CANVAS (canvas);
    out [0] = canvas->width_px;
    out [1] = canvas->height_px;
    out [2] = canvas->work_factor_int;
    out [3] = (int) canvas->blank_char;
    out [4] = (int) canvas->solid_char;
    out [5] = canvas->consider_inverted;
    out [6] = canvas->extract_colors;
    out [7] = canvas->use_quantized_error;
    out [8] = (int) canvas->config.symbol_map.symbols.size ();
    out [9] = (int) canvas->config.symbol_map.symbols2.size ();
    out [10] = (int) canvas->config.fill_symbol_map.symbols.size ();
    out [11] = (int) canvas->config.fill_symbol_map.symbols2.size ();
}

}  /* extern "C" */
