/* internal.h -- host-side object model shared by the C-ABI implementation files.
 *
 * Mirrors the *roles* of the reference's private structs (chafa/internal/
 * chafa-private.h:42-120, chafa-canvas-internal.h:30-88) but is laid out for a
 * device-resident pipeline: a symbol map compiles to flat arrays that are uploaded
 * once per canvas; a canvas owns device buffers and a stream instead of a pixel
 * heap and a thread pool. */
#ifndef CHAFA_B200_INTERNAL_H
#define CHAFA_B200_INTERNAL_H

#include "../../include/chafa.h"
#include "../../include/chafa_b200.h"

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#define N_CANDIDATES_MAX 8
#define SYMBOL_ERROR_MAX (INT32_MAX / 8)
#define CHAFA_PALETTE_INDEX_TRANSPARENT 256
#define CHAFA_PALETTE_INDEX_FG 257
#define CHAFA_PALETTE_INDEX_BG 258
#define CHAFA_PALETTE_INDEX_MAX 259
#define CHAFA_N_SYMBOLS_MAX 2048

/* ---- logging / argument guards (g_return_if_fail semantics: silent no-op) ---- */

bool b200_verbose ();
void b200_log (const char *fmt, ...) __attribute__ ((format (printf, 1, 2)));
void b200_set_error (const char *fmt, ...) __attribute__ ((format (printf, 1, 2)));

#define RETURN_IF_FAIL(e) do { if (!(e)) { if (b200_verbose ()) b200_log ("%s: assertion '%s' failed", __func__, #e); return; } } while (0)
#define RETURN_VAL_IF_FAIL(e, v) do { if (!(e)) { if (b200_verbose ()) b200_log ("%s: assertion '%s' failed", __func__, #e); return (v); } } while (0)

/* ---- unicode helpers (host_util.cpp; tables from tools/gen_unicode_tables.py) ---- */

bool uc_isprint (uint32_t c);
bool uc_isalpha (uint32_t c);
bool uc_isdigit (uint32_t c);
bool uc_ismark (uint32_t c);
bool uc_iszerowidth (uint32_t c);
bool uc_iswide (uint32_t c);
bool uc_iswide_cjk (uint32_t c);
bool uc_is_rtl_script (uint32_t c);
int uc_to_utf8 (uint32_t c, char *out);
uint32_t utf8_get_char (const char *p, int *len_out);

/* ---- GString helpers (malloc-backed, or GLib-backed when GLib is loaded) ---- */

GString *b200_string_new_take (char *data, size_t len, size_t alloc);
GString *b200_string_new_copy (const char *data, size_t len);
void *b200_malloc (size_t n);
/* device memory (host_devmem.cpp): cudaMalloc / cudaFree, framed by guard bands in guard mode */
int b200_dev_malloc (void **p, size_t n);
void b200_dev_free (void *p);
bool b200_dev_guard_mode ();
void b200_free (void *p);
void b200_set_gerror (GError **error, const char *domain, int code, const char *fmt, ...)
    __attribute__ ((format (printf, 4, 5)));

/* ---- symbol map ---- */

struct Selector
{
    bool is_range;
    bool additive;
    uint32_t tags;
    uint32_t first, last;
};

struct CompiledSymbol
{
    uint32_t c;
    uint32_t tags;
    uint64_t bitmap;
    int popcount;
};

struct CompiledSymbol2
{
    uint32_t c;
    uint32_t tags;
    uint64_t bitmap [2];
    int popcount;  /* both halves */
};

struct ChafaSymbolMap
{
    std::atomic<int> refs {1};
    bool need_rebuild = true;
    bool use_builtin_glyphs = true;
    std::vector<Selector> selectors;
    /* User glyphs keyed by code point; insertion order matters for the
     * reference's hash-table iteration order, so keep it. */
    std::vector<std::pair<uint32_t, uint64_t>> glyphs;
    std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> glyphs2;

    /* Populated by symbol_map_prepare () */
    std::vector<CompiledSymbol> symbols;
    std::vector<CompiledSymbol2> symbols2;

    ChafaSymbolMap () {}
    ChafaSymbolMap (const ChafaSymbolMap &o)
        : refs (1), need_rebuild (true), use_builtin_glyphs (o.use_builtin_glyphs),
          selectors (o.selectors), glyphs (o.glyphs), glyphs2 (o.glyphs2) {}
    ChafaSymbolMap &operator= (const ChafaSymbolMap &o)
    {
        need_rebuild = true;
        use_builtin_glyphs = o.use_builtin_glyphs;
        selectors = o.selectors;
        glyphs = o.glyphs;
        glyphs2 = o.glyphs2;
        symbols.clear ();
        symbols2.clear ();
        return *this;
    }
};

void symbol_map_prepare (ChafaSymbolMap *map);
bool symbol_map_has_symbol (const ChafaSymbolMap *map, uint32_t c);
uint32_t symbol_tags_for_char (uint32_t c);

bool glyph_import_device (ChafaPixelType pixel_format, const void *pixels, int width, int height, int rowstride,
                          bool wide, uint64_t *bitmaps_out);

struct Candidate
{
    int symbol_index;
    int hamming_distance;
    bool is_inverted;
};

/* Host restatements used only at canvas creation (blank/solid char choice) and by
 * the row resolver's tables; per-cell searches run on the device. */
int symbol_map_find_candidates (const ChafaSymbolMap *map, uint64_t bitmap, bool do_inverse,
                                Candidate *out, int n_max);
int symbol_map_find_fill_candidates (const ChafaSymbolMap *map, int popcount, bool do_inverse,
                                     Candidate *out, int n_max);

/* ---- canvas config ---- */

struct ChafaCanvasConfig
{
    std::atomic<int> refs {1};
    int width = 80, height = 24;
    int cell_width = 8, cell_height = 8;
    ChafaCanvasMode canvas_mode = CHAFA_CANVAS_MODE_TRUECOLOR;
    ChafaColorSpace color_space = CHAFA_COLOR_SPACE_RGB;
    ChafaDitherMode dither_mode = CHAFA_DITHER_MODE_NONE;
    ChafaColorExtractor color_extractor = CHAFA_COLOR_EXTRACTOR_AVERAGE;
    ChafaPixelMode pixel_mode = CHAFA_PIXEL_MODE_SYMBOLS;
    int dither_grain_width = 4, dither_grain_height = 4;
    float dither_intensity = 1.0f;
    uint32_t fg_color_packed_rgb = 0xffffff;
    uint32_t bg_color_packed_rgb = 0x000000;
    int alpha_threshold = 127;
    float work_factor = 0.5f;
    ChafaSymbolMap symbol_map;
    ChafaSymbolMap fill_symbol_map;
    bool preprocessing_enabled = true;
    bool fg_only_enabled = false;
    int optimizations = CHAFA_OPTIMIZATION_ALL;
    ChafaPassthrough passthrough = CHAFA_PASSTHROUGH_NONE;

    ChafaCanvasConfig ();
    ChafaCanvasConfig (const ChafaCanvasConfig &o);
};

/* ---- term info ---- */

struct SeqArg
{
    uint8_t pre_len;
    uint8_t arg_index;  /* 255 = sentinel */
    bool is_varargs;
};

struct ParsedSeq
{
    bool defined = false;
    std::string unparsed;
    char str [CHAFA_TERM_SEQ_LENGTH_MAX];
    SeqArg args [CHAFA_TERM_SEQ_ARGS_MAX];
};

struct ChafaTermInfo
{
    std::atomic<int> refs {1};
    ParsedSeq seqs [CHAFA_TERM_SEQ_MAX];
};

char *term_info_emit (const ChafaTermInfo *ti, char *out, int seq, const unsigned *argv, int n_args,
                      int type_size);

struct ChafaTermDb
{
    ChafaTermInfo *fallback;
};

/* ---- frame / image / placement ---- */

struct ChafaFrame
{
    std::atomic<int> refs {1};
    ChafaPixelType pixel_type;
    int width, height, rowstride;
    void *data;
    bool data_is_owned;
};

struct ChafaImage
{
    std::atomic<int> refs {1};
    ChafaFrame *frame = nullptr;
};

struct ChafaPlacement
{
    std::atomic<int> refs {1};
    ChafaImage *image = nullptr;
    int id = 0;
    ChafaAlign halign = CHAFA_ALIGN_START, valign = CHAFA_ALIGN_START;
    ChafaTuck tuck = CHAFA_TUCK_STRETCH;
};

/* ---- colour helpers (host) ---- */

/* Packed device colour: ch0 | ch1 << 8 | ch2 << 16 | ch3 << 24 (memory order of
 * the reference's ChafaColor on little-endian hosts). */
static inline uint32_t color_from_packed_rgb (uint32_t packed, uint8_t alpha)
{
    return ((packed >> 16) & 0xff) | (((packed >> 8) & 0xff) << 8) | ((packed & 0xff) << 16)
        | ((uint32_t) alpha << 24);
}

uint32_t host_rgb_to_din99d (uint32_t color);

/* ---- scaler plan (host_scale.cpp) ---- */

enum ScaleFilter
{
    SCALE_FILTER_COPY,
    SCALE_FILTER_ONE,
    SCALE_FILTER_BILINEAR_0H,
    SCALE_FILTER_BILINEAR_1H,
    SCALE_FILTER_BILINEAR_2H,
    SCALE_FILTER_BILINEAR_3H,
    SCALE_FILTER_BOX,
    SCALE_FILTER_NEAREST
};

struct ScaleDim
{
    uint32_t src_size_px = 0;
    uint32_t dest_size_px = 0;
    int filter = SCALE_FILTER_COPY;
    int n_halvings = 0;
    int placement_ofs_px = 0;     /* first visible dest pixel of the placement */
    int placement_size_px = 0;    /* visible placement size */
    int clip_before_px = 0;
    int first_opacity = 256, last_opacity = 256;
    uint32_t span_step = 0, span_mul = 0;
    /* Bilinear: pairs (offset, F) per sample; box: one uint32 per output pixel;
     * nearest: one offset per output pixel.  Stored as uint32 for the device. */
    std::vector<uint32_t> precalc;
};

struct ScalePlan
{
    ScaleDim h, v;
    bool linear = true;           /* sRGB linearisation on (off only at extreme reductions) */
    bool valid = false;
};

bool scale_plan_build (ScalePlan *plan, int src_w, int src_h, int dest_w, int dest_h,
                       int place_x, int place_y, int place_w, int place_h, bool nearest);

void scale_plan_source_rows (const ScalePlan *plan, int dest_first, int dest_n, int *lo_out, int *hi_out);
/* Host-side time per section of the draw / print calls, summed per label and printed at exit when
 * CHAFA_B200_HOST_PROFILE is set (a measurement aid: the device-resident path of short frames is
 * bound by what the calling thread spends queueing them) */
struct HostProf
{
    const char *label;
    long long t0;
    explicit HostProf (const char *label);
    ~HostProf ();
};
#define HOSTPROF_CAT2(a, b) a##b
#define HOSTPROF_CAT(a, b) HOSTPROF_CAT2 (a, b)
#define HOSTPROF(label) HostProf HOSTPROF_CAT (host_prof_, __LINE__) (label)

struct ScaleTiles;
bool scale_plan_tiles (const ScalePlan *plan, int dest_w, int first_row, int n_rows, int n_sms, ScaleTiles *out);

void tuck_and_align (int src_width, int src_height, int dest_width, int dest_height,
                     ChafaAlign halign, ChafaAlign valign, ChafaTuck tuck,
                     int *ofs_x, int *ofs_y, int *width, int *height);

/* ---- bayer / noise textures (host_objects.cpp) ---- */

int reference_batch_count ();
void reference_batch_of_rows (int n_rows, int n_batches, uint16_t *batch_out);

void gen_bayer_matrix (int *out, int matrix_size, float magnitude);
void gen_noise_matrix (int *out, float magnitude);

#endif /* CHAFA_B200_INTERNAL_H */
