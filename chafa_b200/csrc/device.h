/* device.h -- plain structs passed from the host objects to the CUDA kernels, and
 * the launcher entry points of each .cu file.  No CUDA types appear here so that
 * the host-only translation units can include it (a stream is a void *). */
#ifndef CHAFA_B200_DEVICE_H
#define CHAFA_B200_DEVICE_H

#include <cstddef>
#include <cstdint>

/* Layout of the reference's ChafaCanvasCell (chafa/internal/chafa-canvas-internal.h:30-37) */
struct DevCell
{
    uint32_t c;
    uint32_t fg;
    uint32_t bg;
};

/* Result of matching one cell against the narrow symbols (before row resolve) */
struct DevCellEval
{
    uint32_t c;      /* code point chosen */
    uint32_t fg;     /* cell colours after update_cell_colors (packed ARGB or pen) */
    uint32_t bg;
    int32_t error;
    uint32_t mean;   /* mean colour of the 64 pixels, ch0 | ch1<<8 | ch2<<16 | ch3<<24 */
};

/* Result of matching the cell pair (cx-1, cx) against the wide symbols; stored at cx */
struct DevWideEval
{
    uint32_t c;
    uint32_t fg;
    uint32_t bg;
    int32_t error_a;
    int32_t error_b;
};

enum
{
    PAL_DYNAMIC_256 = 0,
    PAL_FIXED_256,
    PAL_FIXED_240,
    PAL_FIXED_16,
    PAL_FIXED_8,
    PAL_FIXED_FGBG
};

#define PAL_INDEX_TRANSPARENT 256
#define PAL_INDEX_FG 257
#define PAL_INDEX_BG 258
#define PAL_INDEX_MAX 259

/* One entry of the sorted pen table of a dynamic palette
 * (role of chafa/internal/chafa-color-table.h:34-60) */
struct DevColorTableEntry
{
    int32_t v [2];
    int32_t pen;
};

struct DevColorTable
{
    DevColorTableEntry entries [256];
    uint32_t pens [256];      /* pen colour, ch0 | ch1<<8 | ch2<<16, 0xffffffff = unused */
    int32_t n_entries;
    int32_t eigenvectors [2] [3];
    int32_t average [3];
    int32_t eigen_mul [2];
};

struct DevPalette
{
    int32_t type;
    int32_t alpha_threshold;
    int32_t transparent_index;
    int32_t n_colors;
    int32_t first_color;
    /* Colours in the canvas's colour space, ch0 | ch1<<8 | ch2<<16 | ch3<<24.
     * fixed[] is the process-wide fixed table the pickers compare against;
     * colors[] is this palette's own copy (FG/BG pens set from the config). */
    uint32_t fixed [PAL_INDEX_MAX];
    uint32_t colors [PAL_INDEX_MAX];
    uint32_t colors_rgb [PAL_INDEX_MAX];
    uint8_t cube_index [256];
};

struct DevSymbols
{
    const uint64_t *bitmaps;     /* n */
    const uint32_t *chars;       /* n */
    const uint8_t *popcounts;    /* n */
    int32_t n;
    const uint64_t *bitmaps2;    /* 2 * n2: left, right */
    const uint32_t *chars2;
    const uint8_t *popcounts2;   /* per half: 2 * n2 */
    int32_t n2;
    const uint32_t *fill_chars;
    const uint8_t *fill_popcounts;
    int32_t n_fill;
    /* Tensor-core operand images (match_tc.cu): every symbol as 64 (narrow) or 128 (wide) bytes,
     * -1 where the symbol covers the pixel and +1 where it does not, rows padded with zero rows
     * to a multiple of 128.  16-byte K slice c of row r is at (c * npad + r) * 16. */
    const uint8_t *bimage;
    const uint8_t *bimage2;
    int32_t npad, n2pad;
};

/* Everything the symbol-mode kernels need to know about the canvas */
struct DevCanvas
{
    int32_t width, height;          /* cells */
    int32_t width_px, height_px;
    int32_t canvas_mode;
    int32_t color_space;
    int32_t color_extractor;
    int32_t work_factor_int;
    int32_t n_candidates;           /* clamp (wfi, 1, 8) */
    int32_t consider_inverted;
    int32_t extract_colors;
    int32_t use_quantized_error;
    int32_t fg_only;
    int32_t alpha_threshold;
    int32_t optimizations;
    uint32_t blank_char, solid_char;
    uint32_t default_fg, default_bg;   /* ch0 | ch1<<8 | ch2<<16 | ch3<<24 */
    int32_t first_row, n_rows;         /* band of cell rows to process */
    const DevPalette *fg_palette;      /* device pointers */
    const DevPalette *bg_palette;
    DevSymbols syms;
    /* host-side knowledge of the two palettes (their structs live in device memory) */
    int32_t fg_pal_type, bg_pal_type;
    int32_t fg_pal_transparent, bg_pal_transparent;
    const uint8_t *pen_lut;            /* set by the tensor matcher's launcher: 2^24 pens, or NULL */
};

/* ---- elementwise preprocessing (prep_pixel.cuh) ---- */

struct DevPrepPixel
{
    int32_t dither_mode;            /* 0 none, 1 ordered, 3 noise (2 = diffusion is not elementwise) */
    int32_t grain_w_shift, grain_h_shift;
    int32_t tex_size_shift, tex_size_mask;
    int32_t to_din99d;
    const int32_t *texture;
    const uint32_t *din99d_lut;     /* dev_din99d_lut (): DIN99d of each of the 2^24 colours, or NULL */
};

/* ---- scaler ---- */

struct DevScaleDim
{
    int32_t filter;
    int32_t n_halvings;
    int32_t src_size_px;
    int32_t placement_ofs_px;
    int32_t placement_size_px;
    int32_t clip_before_px;
    int32_t first_opacity, last_opacity;
    uint32_t span_step, span_mul;
    const uint32_t *precalc;        /* device pointer */
};

/* Tiling of the row-staged scaler (scale_tma_kernel); tile_w == 0: not applicable */
#define SCALE_TILE_CHUNK_ROWS 8
#define SCALE_TILE_MAX_CHUNKS 8
struct ScaleTiles
{
    int32_t tile_w;                 /* destination columns per block = threads per block (128 / 64 / 32) */
    int32_t box_w;                  /* source pixels per TMA box row (multiple of 4, <= 256) */
    int32_t rows_per_block;         /* destination rows per block */
    int32_t n_chunks;               /* boxes of SCALE_TILE_CHUNK_ROWS source rows a block may need */
};

struct DevScale
{
    DevScaleDim h, v;
    const uint8_t *src;             /* device pointer */
    int32_t src_rowstride;
    int32_t src_pixel_type;
    int32_t linear;                 /* sRGB linearisation on */
    uint32_t *dest;                 /* RGBA8, dest_w * dest_h */
    int32_t dest_w, dest_h;
    uint32_t bg_rgb;                /* ch0 | ch1<<8 | ch2<<16 (the compositing colour, alpha taken as 255) */
    int32_t first_row, n_rows;      /* dest rows to produce */
    const uint16_t *from_srgb;      /* device LUTs */
    const uint8_t *to_srgb;
    const uint32_t *inv_div;        /* 4 * 256: p8, p8l, p16, p16l */
    int32_t plain;                  /* 1: no compositing, premultiplied RGBA8 output (glyph import) */
    int32_t post_on;                /* 1: every pixel written goes through prep_dither_convert (post) */
    DevPrepPixel post;
    int32_t composite;              /* 0: over the colour, keeping the source alpha (symbols, sixels);
                                     * 1: over the opaque colour; 2: no colour -- both with unassociated
                                     * RGBA8 output (Kitty / iTerm2 images) */
    ScaleTiles tiles;               /* host_scale.cpp, scale_plan_tiles () */
};

/* ---- Kitty / iTerm2 image text ---- */

#define IMAGE_TEXT_PRE_MAX 16
#define IMAGE_TEXT_SUF_MAX 160
#define IMAGE_TEXT_FRAME_MAX 96

/* base64 text of the byte stream pre | image | suf, cut into chunks of chunk_bytes (a multiple
 * of 3; the last one may be short and is padded with '='), each chunk written as
 * frame_begin | base64 | frame_end. */
struct DevImageText
{
    const uint8_t *image;           /* device: RGBA8 pixmap */
    unsigned long long n_image;
    uint8_t pre [IMAGE_TEXT_PRE_MAX], suf [IMAGE_TEXT_SUF_MAX];
    int32_t n_pre, n_suf;
    uint32_t chunk_bytes;
    uint8_t frame_begin [IMAGE_TEXT_FRAME_MAX], frame_end [IMAGE_TEXT_FRAME_MAX];
    int32_t n_frame_begin, n_frame_end;
    uint8_t *out;                   /* device, 4-byte aligned */
};

unsigned long long dev_image_text_length (const DevImageText *p);
int dev_launch_image_text (const DevImageText *p, void *stream);

/* ---- preprocessing ---- */

struct DevPrep
{
    uint32_t *pixels;
    int32_t width, height;
    int32_t first_row, n_rows;
    int32_t boost_saturation;
    int32_t build_histogram;
    int32_t *hist;                  /* 2048 bins + [2048] = n_samples */
    int32_t normalize;              /* pass 2 */
    int32_t crop_pct;
    int32_t dither_mode;
    int32_t grain_w_shift, grain_h_shift;
    int32_t tex_size_shift, tex_size_mask;
    const int32_t *texture;
    int32_t to_din99d;
    const uint32_t *din99d_lut;
    int32_t *bounds;                /* device: [0] = min, [1] = max */
};

/* ---- Floyd-Steinberg diffusion, symbols mode ---- */

struct DevFsDither
{
    uint32_t *pixels;               /* canvas pixmap, already normalised / converted */
    int32_t width, height;
    int32_t grain_w_shift, grain_h_shift;
    float intensity;
    int32_t color_space;
    const DevPalette *palette;      /* device; fixed palettes only */
    void *scratch;                  /* dev_fs_dither_symbols_scratch_bytes () */
};

/* ---- sixel mode ---- */

struct DevSixel
{
    const uint32_t *pixels;         /* scaled image, width * height, RGBA8 */
    int32_t width, height;          /* canvas size in pixels */
    int32_t image_height;           /* height rounded up to a multiple of 6 */
    uint8_t *indexed;               /* width * image_height pens */
    int32_t color_space;
    int32_t alpha_threshold;
    int32_t dynamic;                /* 1: pen table of a generated palette; 0: fixed palette pickers */
    const DevColorTable *table;     /* device; in the canvas's colour space */
    const uint32_t *pal_colors;     /* device; 259 palette colours in the canvas's colour space */
    const DevPalette *fixed_palette;
    int32_t first_color;
    int32_t transparent_index;
    int32_t dither_mode;
    int32_t grain_w_shift, grain_h_shift;
    int32_t tex_size_shift, tex_size_mask;
    const int32_t *texture;
    float intensity;
    void *scratch;                  /* diffusion: 2 * width error records of 8 bytes */
    uint8_t *lut;                   /* diffusion: pen of each of the 2^24 colours (16 MB, aligned to 16 MB,
                                     * bricked: see sixel.cu), or NULL */
    uint8_t *brick_pens;            /* diffusion: summary of @lut per 8x8x8 colours (32768 bytes): the pen all
                                     * 512 entries share, else 0xff */
    int32_t brick_policy;           /* 0: the chain times both ways and decides; 1 / 2: brick table always / never */
    int32_t preconverted;           /* pixels are already in the canvas's colour space */
    int32_t y0;                     /* image row of pixels [0] (row bands: dither texture and batch lookups) */
};

/* Generated palette of a sixel draw, as the host needs it for the palette header of the text */
struct DevPaletteResult
{
    uint32_t colors_rgb [PAL_INDEX_MAX];
    int32_t n_colors;
    int32_t n_bins;                 /* occupied colour-cube bins the merge started from */
    int32_t n_partner_searches;     /* block-wide partner searches during the merge */
};

/* Input of dev_launch_palette_build (): everything is device memory */
struct DevPaletteBuild
{
    const uint32_t *pixels;         /* scaled image, RGBA8 */
    size_t n_pixels;
    size_t step;                    /* sample every step-th pixel */
    int32_t bits_per_ch;            /* 3..5: colour-cube resolution */
    int32_t alpha_threshold;
    int32_t color_space;            /* 1: the pen table and pal_colors_out are DIN99d */
    int32_t transparent_index;
    const uint32_t *base_colors;    /* PAL_INDEX_MAX colours the palette starts from */
    void *work;                     /* dev_palette_build_work_bytes (bits_per_ch) */
    void *head;                     /* NULL, or dev_palette_build_head_bytes () of the caller's for the bins,
                                     * counters and float sums (what band owners exchange) */
    size_t sample_first, sample_end;    /* the pixels this build samples: [0, n_pixels), or a band's */
    DevColorTable *table_out;
    uint32_t *pal_colors_out;       /* PAL_INDEX_MAX colours in the canvas's colour space */
    DevPaletteResult *result_out;
};

struct DevSixelEncode
{
    const uint8_t *indexed;
    int32_t width, height;          /* height: canvas pixels (not rounded) */
    int32_t n_bands;
    int32_t n_colors;
    int32_t transparent_index;
    int32_t first_pen;              /* the pen that carries the full-width workaround */
    int32_t band0;                  /* image band of band 0 here (row bands); @height stays the image's */
    uint32_t *body_len;             /* n_bands * n_colors */
    uint32_t *prefix_len;           /* n_bands * n_colors */
    uint32_t *pen_ofs;              /* n_bands * n_colors: offset inside the band */
    uint32_t *band_len;             /* n_bands */
    unsigned long long *band_ofs;   /* n_bands + 1 */
    uint8_t *out;
};

/* ---- printer ---- */

#define DEV_SEQ_LEN_MAX 96
#define DEV_SEQ_ARGS_MAX 8

enum
{
    DSEQ_RESET_ATTRIBUTES = 0,
    DSEQ_RESET_COLOR_FG,
    DSEQ_INVERT_COLORS,
    DSEQ_ENABLE_BOLD,
    DSEQ_SET_COLOR_FG_DIRECT,
    DSEQ_SET_COLOR_BG_DIRECT,
    DSEQ_SET_COLOR_FGBG_DIRECT,
    DSEQ_SET_COLOR_FG_256,
    DSEQ_SET_COLOR_BG_256,
    DSEQ_SET_COLOR_FGBG_256,
    DSEQ_SET_COLOR_FG_16,
    DSEQ_SET_COLOR_BG_16,
    DSEQ_SET_COLOR_FGBG_16,
    DSEQ_SET_COLOR_FG_8,
    DSEQ_SET_COLOR_BG_8,
    DSEQ_SET_COLOR_FGBG_8,
    DSEQ_REPEAT_CHAR,
    DSEQ_MAX
};

struct DevSeq
{
    uint8_t str [DEV_SEQ_LEN_MAX];
    uint8_t pre_len [DEV_SEQ_ARGS_MAX];    /* literal bytes before argument slot i; slot n_slots = tail */
    uint8_t arg_index [DEV_SEQ_ARGS_MAX];  /* 255 = end */
};

struct DevTermInfo
{
    DevSeq seqs [DSEQ_MAX];
};

struct DevPrint
{
    const DevCell *cells;
    int32_t width, height;
    int32_t first_row, n_rows;
    int32_t canvas_mode;
    int32_t fg_only;
    int32_t alpha_threshold;
    int32_t optimizations;
    int32_t have_repeat;
    uint32_t fgbg_blank_symbol;       /* emit_ansi_fgbg_bgfg's blank symbol, 0 if none */
    int32_t total_height;             /* canvas height (newline rule) */
    const DevTermInfo *ti;            /* device */
    uint32_t *cell_len;               /* width * height scratch */
    uint64_t *row_ofs;                /* height + 1 */
    uint8_t *out;
    size_t out_capacity;
    /* single-launch printer (print_fused_kernel) */
    uint32_t *ctrl;                   /* [0] row ticket, [1] rows finished; zero between launches */
    unsigned long long *status;       /* n_rows look-back words */
    uint32_t epoch;                   /* 1 .. 2^22 - 1, different from the previous launch's */
    uint32_t stage_bytes;             /* bound of one row's text if rows are staged in shared memory, else 0 */
};

/* The row walk (K4) as the first phase of the printer's launch: set when the cells of the last
 * draw have not been asked for in between, so that the walk need not be a launch of its own */
struct DevResolveArgs
{
    int32_t enabled;
    DevCanvas cv;
    const DevCellEval *evals;
    const DevWideEval *wide_evals;
    DevCell *cells;
};

/* ---- launchers (each returns 0 on success, else a cudaError_t value) ---- */

int dev_launch_scale (const DevScale *p, void *stream);
int dev_launch_prep_pass1 (const DevPrep *p, void *stream);
int dev_launch_prep_bounds (const DevPrep *p, void *stream);
int dev_launch_prep_pass2 (const DevPrep *p, void *stream);
/* RGB -> DIN99d of all 2^24 colours (64 MB), built on first use per device with the same double
 * arithmetic the per-pixel route uses; NULL if it cannot be allocated */
const uint32_t *dev_din99d_lut (int device, void *stream);
size_t dev_fs_dither_symbols_scratch_bytes (int width, int grain_w_shift);
int dev_launch_fs_dither_symbols (const DevFsDither *p, void *stream);
int dev_launch_match (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, void *stream);
bool dev_match_exhaustive_applies (const DevCanvas *cv);
int dev_launch_match_exhaustive (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals,
                                 DevWideEval *wide_evals, void *stream);
int dev_launch_match_wide (const DevCanvas *cv, const uint32_t *pixels, DevWideEval *wide_evals, void *stream);
bool dev_match_tc_applies (const DevCanvas *cv, bool wide);
void dev_match_tc_prepare (const DevCanvas *cv, void *stream);
bool dev_match_xtc_applies (const DevCanvas *cv, bool wide);
int dev_launch_match_xtc (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, DevWideEval *wide_evals,
                          void *stream);
int dev_launch_match_tc (const DevCanvas *cv, const uint32_t *pixels, DevCellEval *evals, DevWideEval *wide_evals,
                         void *stream);
int dev_launch_resolve (const DevCanvas *cv, const DevCellEval *evals, const DevWideEval *wide_evals,
                        DevCell *cells, void *stream);
unsigned long long dev_scale_staged_launches ();
int dev_opt_in_smem (const void *func, size_t smem);
int dev_sm_count ();
size_t dev_palette_build_work_bytes (int bits_per_ch);
size_t dev_palette_build_head_bytes (int bits_per_ch);
size_t dev_palette_build_reduce_words (int bits_per_ch);
int dev_launch_palette_build (const DevPaletteBuild *p, void *stream);
int dev_launch_palette_sample (const DevPaletteBuild *p, void *stream);
int dev_launch_palette_float_bins (const DevPaletteBuild *p, void *stream);
int dev_launch_palette_finish (const DevPaletteBuild *p, void *stream);
int dev_launch_sixel_quantise (const DevSixel *p, void *stream);
bool dev_sixel_chain_applies (const DevSixel *p);
int dev_launch_sixel_chain_tables (const DevSixel *p, void *stream);
int dev_launch_sixel_chain_batch (const DevSixel *d_images, int n, int max_width, bool unit_intensity, bool dynamic,
                                  void *stream);
size_t dev_sixel_cache_scratch_bytes (size_t n_px);
int dev_launch_sixel_cache_replay (const DevSixel *p, const uint16_t *d_row_batch, const uint8_t *exact, uint8_t *out,
                                   void *scratch, size_t scratch_bytes, void *stream);
int dev_launch_sixel_encode_measure (const DevSixelEncode *p, void *stream);
int dev_launch_sixel_encode_emit (const DevSixelEncode *p, void *stream);
int dev_launch_print_measure (const DevPrint *p, void *stream);
int dev_launch_print_emit (const DevPrint *p, void *stream);
size_t dev_print_fused_smem_bytes (int width, size_t row_bound);
int dev_launch_print_fused (const DevPrint *p, const DevResolveArgs *resolve, void *stream);

int dev_launch_band_scatter (const uint8_t *text, const uint64_t *lengths, int n_bands, int band, uint8_t *full_text,
                             size_t capacity, uint64_t *total_out, void *stream);

void dev_count_launch (int n);
int host_palette_lookup_nearest (const DevPalette *pal, int color_space, uint32_t color);

#endif /* CHAFA_B200_DEVICE_H */
