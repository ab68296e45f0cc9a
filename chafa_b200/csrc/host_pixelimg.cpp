/* host_pixelimg.cpp -- Kitty and iTerm2 pixel modes: the scaler (K1) followed by the image
 * text kernel (K7); the host only assembles the few dozen bytes of framing around it.
 *
 * Replaces (reference file:line):
 *   chafa/internal/chafa-kitty-renderer.c:136-222    draw: scale, composite, unassociated RGBA8
 *   chafa/internal/chafa-kitty-renderer.c:227-474    immediate image, virtual image + unicode
 *                                                    placeholders under screen / tmux
 *   chafa/internal/chafa-iterm2-renderer.c:158-329   draw; TIFF wrapped in base64
 *   chafa/internal/chafa-passthrough-encoder.c       packet framing of the sequences
 *
 * Every Kitty chunk is wrapped in the same bytes, so the framing logic runs once on the host
 * for an empty payload to obtain the text before the first chunk, the text around each chunk
 * and the text after the last one; the device writes everything in between. */

#include "internal.h"
#include "device.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "gen_tables.inc"

struct ChafaCanvas;

/* host_canvas.cpp */
cudaStream_t canvas_stream (ChafaCanvas *c);
const ChafaCanvasConfig *canvas_config (const ChafaCanvas *c);
void **canvas_pixelimg_slot (ChafaCanvas *c);
int canvas_width_px (const ChafaCanvas *c);
int canvas_height_px (const ChafaCanvas *c);
const ChafaPlacement *canvas_placement (const ChafaCanvas *c);
uint32_t *canvas_pixels (ChafaCanvas *c);
void canvas_set_scale_composite (ChafaCanvas *c, int composite);
bool canvas_run_scaler (ChafaCanvas *c, ChafaPixelType type, const void *d_src, int src_w, int src_h, int rowstride,
                        int px, int py, int pw, int ph, bool nearest, int first_row, int n_rows);
bool canvas_cuda_ok (int err, const char *what);

#define CK(call) canvas_cuda_ok ((int) (call), #call)

struct PixelImgState
{
    bool have_image = false;
    void *d_out = nullptr;
    size_t cap_out = 0;
    void *h_text = nullptr;     /* pinned */
    size_t cap_h_text = 0;
};

void
pixelimg_state_free (void *state)
{
    PixelImgState *s = (PixelImgState *) state;
    if (!s)
        return;
    if (s->d_out)
        b200_dev_free (s->d_out);
    if (s->h_text)
        cudaFreeHost (s->h_text);
    delete s;
}

static PixelImgState *
state_of (ChafaCanvas *c)
{
    void **slot = canvas_pixelimg_slot (c);
    if (!*slot)
        *slot = new PixelImgState ();
    return (PixelImgState *) *slot;
}

/* chafa-kitty-renderer.c:136-222, chafa-iterm2-renderer.c:158-238 */
bool
pixelimg_draw (ChafaCanvas *c, const void *d_src, ChafaPixelType type, int src_w, int src_h, int rowstride)
{
    const ChafaCanvasConfig *cfg = canvas_config (c);
    const ChafaPlacement *pl = canvas_placement (c);
    PixelImgState *s = state_of (c);
    int w = canvas_width_px (c), h = canvas_height_px (c);

    s->have_image = false;

    /* The background is composited in only when the transparency threshold asks for a
     * fully flattened image (chafa-canvas.c:419-424); otherwise alpha travels to the terminal. */
    bool flatten_alpha = cfg->alpha_threshold < 1;

    int px, py, pw, ph;
    tuck_and_align (src_w, src_h, w, h, pl ? pl->halign : CHAFA_ALIGN_START, pl ? pl->valign : CHAFA_ALIGN_START,
                    pl ? pl->tuck : CHAFA_TUCK_STRETCH, &px, &py, &pw, &ph);
    canvas_set_scale_composite (c, flatten_alpha ? 1 : 2);
    bool ok = canvas_run_scaler (c, type, d_src, src_w, src_h, rowstride, px, py, pw, ph, false, 0, h);
    canvas_set_scale_composite (c, 0);
    if (!ok)
        return false;
    s->have_image = true;
    return true;
}

/* ---- framing ---------------------------------------------------------------------------- */

namespace {

/* The reference's packet encoder, writing to whichever string is current */
struct Framer
{
    ChafaPassthrough mode;
    const ChafaTermInfo *ti;
    std::string *out;
    size_t packet_size = 0;

    std::string seq (int id, const unsigned *args = nullptr, int n_args = 0, int type_size = 4) const
    {
        char buf [CHAFA_TERM_SEQ_LENGTH_MAX * 2];
        char *e = term_info_emit (ti, buf, id, args, n_args, type_size);
        return std::string (buf, (size_t) (e - buf));
    }

    size_t packet_max () const { return mode == CHAFA_PASSTHROUGH_SCREEN ? 200 : 1000000; }

    void packet_begin ()
    {
        if (mode == CHAFA_PASSTHROUGH_SCREEN)
            *out += seq (CHAFA_TERM_SEQ_BEGIN_SCREEN_PASSTHROUGH);
        else if (mode == CHAFA_PASSTHROUGH_TMUX)
            *out += seq (CHAFA_TERM_SEQ_BEGIN_TMUX_PASSTHROUGH);
    }

    void packet_end ()
    {
        if (mode == CHAFA_PASSTHROUGH_SCREEN)
            *out += seq (CHAFA_TERM_SEQ_END_SCREEN_PASSTHROUGH);
        else if (mode == CHAFA_PASSTHROUGH_TMUX)
            *out += seq (CHAFA_TERM_SEQ_END_TMUX_PASSTHROUGH);
    }

    void packetized (const char *in, size_t len)
    {
        while (len > 0)
        {
            size_t remain = packet_max () - packet_size;
            if (remain == 0)
            {
                packet_end ();
                packet_size = 0;
                remain = packet_max ();
            }
            if (packet_size == 0)
                packet_begin ();
            size_t n = std::min (len, remain);
            out->append (in, n);
            len -= n;
            packet_size += n;
            in += n;
        }
    }

    void append (const std::string &in)
    {
        if (mode == CHAFA_PASSTHROUGH_NONE)
            *out += in;
        else if (mode == CHAFA_PASSTHROUGH_SCREEN)
            packetized (in.data (), in.size ());
        else
        {
            std::string esc;
            for (char ch : in)
            {
                esc += ch;
                if (ch == 0x1b)
                    esc += ch;
            }
            packetized (esc.data (), esc.size ());
        }
    }

    void flush ()
    {
        if (packet_size > 0)
        {
            packet_end ();
            packet_size = 0;
        }
    }

    void reset () { packet_size = 0; }

    /* chafa-kitty-renderer.c:227-253 */
    void end_passthrough ()
    {
        if (mode == CHAFA_PASSTHROUGH_SCREEN)
        {
            std::string e = seq (CHAFA_TERM_SEQ_END_SCREEN_PASSTHROUGH);
            for (char ch : e)
            {
                flush ();
                packetized (&ch, 1);
            }
        }
        else if (mode == CHAFA_PASSTHROUGH_TMUX)
        {
            flush ();
            *out += seq (CHAFA_TERM_SEQ_END_TMUX_PASSTHROUGH);
        }
        flush ();
    }
};

void
append_utf8 (std::string &s, uint32_t c)
{
    if (c < 0x80)
        s += (char) c;
    else if (c < 0x800)
    {
        s += (char) (0xc0 | (c >> 6));
        s += (char) (0x80 | (c & 0x3f));
    }
    else if (c < 0x10000)
    {
        s += (char) (0xe0 | (c >> 12));
        s += (char) (0x80 | ((c >> 6) & 0x3f));
        s += (char) (0x80 | (c & 0x3f));
    }
    else
    {
        s += (char) (0xf0 | (c >> 18));
        s += (char) (0x80 | ((c >> 12) & 0x3f));
        s += (char) (0x80 | ((c >> 6) & 0x3f));
        s += (char) (0x80 | (c & 0x3f));
    }
}

/* Cells that show a virtual image: the placeholder character with the row and column written
 * as combining marks, the image id in the foreground colour (chafa-kitty-renderer.c:315-408) */
void
unicode_placement (const Framer &f, std::string &out, int width_cells, int height_cells, int placement_id)
{
    const bool screen = f.mode == CHAFA_PASSTHROUGH_SCREEN;

    width_cells = std::min (width_cells, 296);
    height_cells = std::min (height_cells, 296);

    for (int i = 0; i < height_cells; i++)
    {
        if (i > 0)
        {
            /* screen moves one cell too far after three of the marks */
            unsigned left = (unsigned) width_cells + ((screen && (i == 35 || i == 61 || i == 62)) ? 1 : 0);
            out += f.seq (CHAFA_TERM_SEQ_CURSOR_LEFT, &left, 1);
            out += f.seq (CHAFA_TERM_SEQ_CURSOR_DOWN_SCROLL);
        }
        unsigned id = (unsigned) placement_id;
        out += f.seq (CHAFA_TERM_SEQ_SET_COLOR_FG_256, &id, 1, 1);

        for (int j = 0; j < width_cells; j++)
        {
            append_utf8 (out, 0x10eeeeu);
            if (!screen || j == 0)
                append_utf8 (out, gen_kitty_diacritics [i]);
            if (!screen)
                append_utf8 (out, gen_kitty_diacritics [j]);
        }
    }
    out += f.seq (CHAFA_TERM_SEQ_RESET_COLOR_FG);
}

void
put_le16 (uint8_t *&p, unsigned v)
{
    *p++ = (uint8_t) v;
    *p++ = (uint8_t) (v >> 8);
}

void
put_le32 (uint8_t *&p, uint32_t v)
{
    put_le16 (p, v & 0xffff);
    put_le16 (p, v >> 16);
}

void
put_tiff_tag (uint8_t *&p, unsigned tag, unsigned type, uint32_t count, uint32_t data)
{
    put_le16 (p, tag);
    put_le16 (p, type);
    put_le32 (p, count);
    put_le32 (p, data);
}

}  /* namespace */

GString *
pixelimg_print (ChafaCanvas *c, ChafaTermInfo *ti)
{
    const ChafaCanvasConfig *cfg = canvas_config (c);
    const ChafaPlacement *pl = canvas_placement (c);
    cudaStream_t st = canvas_stream (c);
    PixelImgState *s = (PixelImgState *) *canvas_pixelimg_slot (c);
    if (!s || !s->have_image)
        return nullptr;

    const int w = canvas_width_px (c), h = canvas_height_px (c);
    const unsigned long long image_size = (unsigned long long) w * (unsigned long long) h * 4;

    DevImageText t;
    memset (&t, 0, sizeof (t));
    t.image = (const uint8_t *) canvas_pixels (c);
    t.n_image = image_size;

    std::string head, frame_begin, frame_end, tail;

    if (cfg->pixel_mode == CHAFA_PIXEL_MODE_KITTY)
    {
        ChafaPassthrough mode = cfg->passthrough;
        Framer f { mode, ti, &head };
        unsigned args [6] = { 32, (unsigned) w, (unsigned) h, (unsigned) cfg->width, (unsigned) cfg->height, 0 };
        int placement_id = pl ? pl->id : -1;

        if (mode == CHAFA_PASSTHROUGH_NONE)
        {
            /* chafa-kitty-renderer.c:291-312 */
            f.append (f.seq (CHAFA_TERM_SEQ_BEGIN_KITTY_IMMEDIATE_IMAGE_V1, args, 5));
            f.flush ();
        }
        else
        {
            /* chafa-kitty-renderer.c:410-474: ids cycle through 1..255 */
            if (placement_id < 1)
                placement_id = 1;
            else if (placement_id > 255)
                placement_id = 1 + (placement_id % 255);
            args [5] = (unsigned) placement_id;
            f.append (f.seq (CHAFA_TERM_SEQ_BEGIN_KITTY_IMMEDIATE_VIRT_IMAGE_V1, args, 6));
            f.reset ();
            f.end_passthrough ();
        }

        /* chafa-kitty-renderer.c:255-289 */
        t.chunk_bytes = mode == CHAFA_PASSTHROUGH_SCREEN ? 63 : 510;
        f.out = &frame_begin;
        f.append (f.seq (CHAFA_TERM_SEQ_BEGIN_KITTY_IMAGE_CHUNK));
        f.out = &frame_end;
        f.append (f.seq (CHAFA_TERM_SEQ_END_KITTY_IMAGE_CHUNK));
        f.reset ();
        f.end_passthrough ();

        f.out = &tail;
        f.append (f.seq (CHAFA_TERM_SEQ_END_KITTY_IMAGE));
        f.reset ();
        f.end_passthrough ();
        if (mode != CHAFA_PASSTHROUGH_NONE)
        {
            f.end_passthrough ();
            f.flush ();
            unicode_placement (f, tail, cfg->width, cfg->height, placement_id);
        }
    }
    else
    {
        /* iTerm2: an uncompressed little-endian TIFF, strip first, directory after it
         * (chafa-iterm2-renderer.c:251-329) */
        Framer f { CHAFA_PASSTHROUGH_NONE, ti, &head };
        unsigned args [2] = { (unsigned) cfg->width, (unsigned) cfg->height };
        head = f.seq (CHAFA_TERM_SEQ_BEGIN_ITERM2_IMAGE, args, 2);
        tail = f.seq (CHAFA_TERM_SEQ_END_ITERM2_IMAGE);

        uint8_t *p = t.pre;
        put_le32 (p, 0x002a4949u);
        put_le32 (p, (uint32_t) (image_size + 8));
        t.n_pre = 8;

        p = t.suf;
        put_le16 (p, 11);
        put_tiff_tag (p, 256, 4, 1, (uint32_t) w);                              /* ImageWidth */
        put_tiff_tag (p, 257, 4, 1, (uint32_t) h);                              /* ImageLength */
        put_tiff_tag (p, 258, 3, 4, (uint32_t) (image_size + 8 + 2 + 12 * 11 + 4));   /* BitsPerSample -> end of file */
        put_tiff_tag (p, 262, 3, 1, 2);                                         /* PhotometricInterpretation RGB */
        put_tiff_tag (p, 273, 4, 1, 8);                                         /* StripOffsets */
        put_tiff_tag (p, 274, 3, 1, 1);                                         /* Orientation top-left */
        put_tiff_tag (p, 277, 3, 1, 4);                                         /* SamplesPerPixel */
        put_tiff_tag (p, 278, 4, 1, (uint32_t) h);                              /* RowsPerStrip */
        put_tiff_tag (p, 279, 4, 1, (uint32_t) image_size);                     /* StripByteCounts */
        put_tiff_tag (p, 284, 3, 1, 1);                                         /* PlanarConfiguration contiguous */
        put_tiff_tag (p, 338, 3, 1, 2);                                         /* ExtraSamples: unassociated alpha */
        put_le32 (p, 0);
        for (int i = 0; i < 4; i++)
            put_le16 (p, 8);
        t.n_suf = (int32_t) (p - t.suf);

        /* one chunk over everything */
        unsigned long long n_stream = image_size + (unsigned long long) t.n_pre + (unsigned long long) t.n_suf;
        t.chunk_bytes = (uint32_t) ((n_stream + 2) / 3 * 3);
    }

    if (frame_begin.size () > IMAGE_TEXT_FRAME_MAX || frame_end.size () > IMAGE_TEXT_FRAME_MAX)
    {
        b200_set_error ("image chunk framing of %zu + %zu bytes exceeds the device limit", frame_begin.size (),
                        frame_end.size ());
        return nullptr;
    }
    memcpy (t.frame_begin, frame_begin.data (), frame_begin.size ());
    memcpy (t.frame_end, frame_end.data (), frame_end.size ());
    t.n_frame_begin = (int32_t) frame_begin.size ();
    t.n_frame_end = (int32_t) frame_end.size ();

    unsigned long long body_len = dev_image_text_length (&t);
    if (body_len)
    {
        size_t need = (size_t) body_len + 16;
        if (need > s->cap_out)
        {
            if (s->d_out)
                b200_dev_free (s->d_out);
            s->d_out = nullptr;
            s->cap_out = 0;
            if (!CK (b200_dev_malloc (&s->d_out, need)))
                return nullptr;
            s->cap_out = need;
        }
        if (need > s->cap_h_text)
        {
            if (s->h_text)
                cudaFreeHost (s->h_text);
            s->h_text = nullptr;
            s->cap_h_text = 0;
            if (!CK (cudaMallocHost (&s->h_text, need)))
                return nullptr;
            s->cap_h_text = need;
        }
        t.out = (uint8_t *) s->d_out;
        if (!CK (dev_launch_image_text (&t, st))
            || !CK (cudaMemcpyAsync (s->h_text, s->d_out, (size_t) body_len, cudaMemcpyDeviceToHost, st))
            || !CK (cudaStreamSynchronize (st)))
            return nullptr;
    }

    size_t total = head.size () + (size_t) body_len + tail.size ();
    char *buf = (char *) b200_malloc (total + 1);
    memcpy (buf, head.data (), head.size ());
    if (body_len)
        memcpy (buf + head.size (), s->h_text, (size_t) body_len);
    memcpy (buf + head.size () + body_len, tail.data (), tail.size ());
    buf [total] = '\0';
    return b200_string_new_take (buf, total, total + 1);
}
