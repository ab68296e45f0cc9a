/* host_objects.cpp -- small host-side objects and helpers around the canvas path:
 * frame / image / placement wrappers (chafa/chafa-frame.c, chafa-image.c,
 * chafa-placement.c), feature and thread-count queries (chafa/chafa-features.c),
 * geometry helpers (chafa/chafa-util.c:61-159, chafa/internal/chafa-math-util.c:25-117)
 * and the dither textures (chafa/chafa-util.c:162-231, chafa/internal/chafa-noise.c:227-240). */

#include "internal.h"

#include <algorithm>
#include <thread>

#include <cmath>

#include "gen_tables.inc"

/* ---- frame ---- */

static size_t
frame_bytes (ChafaPixelType t, int height, int rowstride)
{
    (void) t;
    return (size_t) height * (size_t) rowstride;
}

extern "C" {

ChafaFrame *
chafa_frame_new (gconstpointer data, ChafaPixelType pixel_type, gint width, gint height, gint rowstride)
{
    RETURN_VAL_IF_FAIL ((unsigned) pixel_type < CHAFA_PIXEL_MAX, NULL);
    RETURN_VAL_IF_FAIL (width > 0 && height > 0 && rowstride > 0, NULL);
    ChafaFrame *f = new ChafaFrame ();
    f->pixel_type = pixel_type;
    f->width = width;
    f->height = height;
    f->rowstride = rowstride;
    f->data = b200_malloc (frame_bytes (pixel_type, height, rowstride));
    memcpy (f->data, data, frame_bytes (pixel_type, height, rowstride));
    f->data_is_owned = true;
    return f;
}

ChafaFrame *
chafa_frame_new_steal (gpointer data, ChafaPixelType pixel_type, gint width, gint height, gint rowstride)
{
    RETURN_VAL_IF_FAIL ((unsigned) pixel_type < CHAFA_PIXEL_MAX, NULL);
    RETURN_VAL_IF_FAIL (width > 0 && height > 0 && rowstride > 0, NULL);
    ChafaFrame *f = new ChafaFrame ();
    f->pixel_type = pixel_type;
    f->width = width;
    f->height = height;
    f->rowstride = rowstride;
    f->data = data;
    f->data_is_owned = true;
    return f;
}

ChafaFrame *
chafa_frame_new_borrow (gpointer data, ChafaPixelType pixel_type, gint width, gint height, gint rowstride)
{
    ChafaFrame *f = chafa_frame_new_steal (data, pixel_type, width, height, rowstride);
    if (f)
        f->data_is_owned = false;
    return f;
}

void
chafa_frame_ref (ChafaFrame *frame)
{
    RETURN_IF_FAIL (frame != NULL);
    RETURN_IF_FAIL (frame->refs.load () > 0);
    frame->refs++;
}

void
chafa_frame_unref (ChafaFrame *frame)
{
    RETURN_IF_FAIL (frame != NULL);
    RETURN_IF_FAIL (frame->refs.load () > 0);
    if (--frame->refs == 0)
    {
        if (frame->data_is_owned)
            free (frame->data);
        delete frame;
    }
}

/* ---- image ---- */

ChafaImage *
chafa_image_new (void)
{
    return new ChafaImage ();
}

void
chafa_image_ref (ChafaImage *image)
{
    RETURN_IF_FAIL (image != NULL);
    RETURN_IF_FAIL (image->refs.load () > 0);
    image->refs++;
}

void
chafa_image_unref (ChafaImage *image)
{
    RETURN_IF_FAIL (image != NULL);
    RETURN_IF_FAIL (image->refs.load () > 0);
    if (--image->refs == 0)
    {
        if (image->frame)
            chafa_frame_unref (image->frame);
        delete image;
    }
}

void
chafa_image_set_frame (ChafaImage *image, ChafaFrame *frame)
{
    RETURN_IF_FAIL (image != NULL);
    if (frame)
        chafa_frame_ref (frame);
    if (image->frame)
        chafa_frame_unref (image->frame);
    image->frame = frame;
}

/* ---- placement ---- */

ChafaPlacement *
chafa_placement_new (ChafaImage *image, gint id)
{
    RETURN_VAL_IF_FAIL (image != NULL, NULL);
    if (id < 1)
        id = -1;
    ChafaPlacement *p = new ChafaPlacement ();
    chafa_image_ref (image);
    p->image = image;
    p->id = id;
    return p;
}

void
chafa_placement_ref (ChafaPlacement *placement)
{
    RETURN_IF_FAIL (placement != NULL);
    RETURN_IF_FAIL (placement->refs.load () > 0);
    placement->refs++;
}

void
chafa_placement_unref (ChafaPlacement *placement)
{
    RETURN_IF_FAIL (placement != NULL);
    RETURN_IF_FAIL (placement->refs.load () > 0);
    if (--placement->refs == 0)
    {
        chafa_image_unref (placement->image);
        delete placement;
    }
}

ChafaTuck
chafa_placement_get_tuck (ChafaPlacement *placement)
{
    RETURN_VAL_IF_FAIL (placement != NULL, CHAFA_TUCK_FIT);
    return placement->tuck;
}

void
chafa_placement_set_tuck (ChafaPlacement *placement, ChafaTuck tuck)
{
    RETURN_IF_FAIL (placement != NULL);
    RETURN_IF_FAIL ((unsigned) tuck < CHAFA_TUCK_MAX);
    placement->tuck = tuck;
}

ChafaAlign
chafa_placement_get_halign (ChafaPlacement *placement)
{
    RETURN_VAL_IF_FAIL (placement != NULL, CHAFA_ALIGN_START);
    return placement->halign;
}

void
chafa_placement_set_halign (ChafaPlacement *placement, ChafaAlign align)
{
    RETURN_IF_FAIL (placement != NULL);
    RETURN_IF_FAIL ((unsigned) align < CHAFA_ALIGN_MAX);
    placement->halign = align;
}

ChafaAlign
chafa_placement_get_valign (ChafaPlacement *placement)
{
    RETURN_VAL_IF_FAIL (placement != NULL, CHAFA_ALIGN_START);
    return placement->valign;
}

void
chafa_placement_set_valign (ChafaPlacement *placement, ChafaAlign align)
{
    RETURN_IF_FAIL (placement != NULL);
    RETURN_IF_FAIL ((unsigned) align < CHAFA_ALIGN_MAX);
    placement->valign = align;
}

/* ---- features / threads ---- */

/* The CPU SIMD leaves of the reference do not exist here; the kernels implement the
 * arithmetic of its AVX2 build, which is what these report. */
ChafaFeatures
chafa_get_builtin_features (void)
{
    return (ChafaFeatures) (CHAFA_FEATURE_MMX | CHAFA_FEATURE_SSE41 | CHAFA_FEATURE_POPCNT | CHAFA_FEATURE_AVX2);
}

ChafaFeatures
chafa_get_supported_features (void)
{
    return chafa_get_builtin_features ();
}

gchar *
chafa_describe_features (ChafaFeatures features)
{
    std::string s;
    if (features & CHAFA_FEATURE_MMX) s += "mmx ";
    if (features & CHAFA_FEATURE_SSE41) s += "sse4.1 ";
    if (features & CHAFA_FEATURE_POPCNT) s += "popcnt ";
    if (features & CHAFA_FEATURE_AVX2) s += "avx2 ";
    if (!s.empty ())
        s.pop_back ();
    return strdup (s.c_str ());
}

static std::atomic<int> n_threads_setting { -1 };

gint
chafa_get_n_threads (void)
{
    return n_threads_setting.load ();
}

void
chafa_set_n_threads (gint n)
{
    RETURN_IF_FAIL (n >= -1);
    n_threads_setting.store (n);
}

gint
chafa_get_n_actual_threads (void)
{
    /* The reference's answer for this host and setting (chafa-features.c:266-282): callers size
     * their own pipelines by it.  No worker threads exist here -- a canvas is driven by the calling
     * thread and the work runs as kernels -- but the number still decides the batch split the
     * sixel lookup cache replay reproduces. */
    return reference_batch_count ();
}

}  /* extern "C" */

/* Number of worker batches the reference would split a pass into on this host with the
 * current chafa_set_n_threads () setting (chafa/chafa-features.c:266-282: auto = number of
 * processors, at most 24).  The GPU path has no worker threads; the number only matters
 * where the reference's output depends on its batch split (the sixel lookup cache). */
int
reference_batch_count ()
{
    int n = n_threads_setting.load ();
    if (n < 0)
    {
        n = (int) std::thread::hardware_concurrency ();
        n = std::min (n, 24);
    }
    return n <= 0 ? 1 : n;
}

/* Row -> batch, as chafa_process_batches () divides @n_rows single-row units among
 * @n_batches (chafa/internal/chafa-batch.c:97-150), float accumulation included */
void
reference_batch_of_rows (int n_rows, int n_batches, uint16_t *batch_out)
{
    float units_per_batch = (float) n_rows / (float) n_batches;
    units_per_batch = std::max (units_per_batch, 1.0f);
    float next_unit_ofs_f = .0f;
    int unit_ofs [2] = { 0, 0 };

    for (int r = 0; r < n_rows; r++)
        batch_out [r] = 0;
    for (int i = 0; i < n_batches; i++)
    {
        do
        {
            next_unit_ofs_f += units_per_batch;
            unit_ofs [1] = (int) next_unit_ofs_f;
        }
        while (unit_ofs [0] == unit_ofs [1]);

        int r0 = unit_ofs [0], r1 = unit_ofs [1];
        if (i == n_batches - 1 || r1 > n_rows)
            r1 = n_rows;
        if (r0 >= r1)
            break;
        for (int r = r0; r < r1; r++)
            batch_out [r] = (uint16_t) i;
        unit_ofs [0] = unit_ofs [1];
    }
}

extern "C" {

/* ---- geometry ---- */

void
chafa_calc_canvas_geometry (gint src_width, gint src_height, gint *dest_width_inout, gint *dest_height_inout,
                            gfloat font_ratio, gboolean zoom, gboolean stretch)
{
    int dest_width = -1, dest_height = -1;
    double fr = (double) font_ratio;

    RETURN_IF_FAIL (src_width >= 0);
    RETURN_IF_FAIL (src_height >= 0);
    RETURN_IF_FAIL (font_ratio > 0.0f);

    if (dest_width_inout) dest_width = *dest_width_inout;
    if (dest_height_inout) dest_height = *dest_height_inout;

    if (src_width == 0 || src_height == 0 || dest_width == 0 || dest_height == 0)
    {
        if (dest_width_inout) *dest_width_inout = 0;
        if (dest_height_inout) *dest_height_inout = 0;
        return;
    }

    if (dest_width < 0 && dest_height < 0)
    {
        if (dest_width_inout)
            *dest_width_inout = std::max ((src_width + 7) / 8, 1);
        if (dest_height_inout)
            *dest_height_inout = std::max ((int) (((src_height + 7) / 8) * fr + 0.5), 1);
        return;
    }

    if (!zoom)
    {
        dest_width = std::min (dest_width, src_width);
        dest_height = std::min (dest_height, src_height);
    }

    if (!stretch || dest_width < 0 || dest_height < 0)
    {
        double src_aspect = src_width / (double) src_height;
        double dest_aspect = (dest_width / (double) dest_height) * fr;

        if (dest_width < 1)
            dest_width = (int) ceil (dest_height * (src_aspect / fr));
        else if (dest_height < 1)
            dest_height = (int) ceil ((dest_width / src_aspect) * fr);
        else if (src_aspect > dest_aspect)
            dest_height = (int) ceil (dest_width * (fr / src_aspect));
        else
            dest_width = (int) ceil (dest_height * (src_aspect / fr));
    }

    dest_width = std::max (dest_width, 1);
    dest_height = std::max (dest_height, 1);

    if (dest_width_inout && *dest_width_inout > 0)
        dest_width = std::min (dest_width, *dest_width_inout);
    if (dest_height_inout && *dest_height_inout > 0)
        dest_height = std::min (dest_height, *dest_height_inout);

    if (dest_width_inout) *dest_width_inout = dest_width;
    if (dest_height_inout) *dest_height_inout = dest_height;
}

void
chafa_free_gstring_array (GString **gsa)
{
    if (!gsa)
        return;
    for (int i = 0; gsa [i]; i++)
        chafa_b200_string_free (gsa [i]);
    free (gsa);
}

} /* extern "C" */

static int
align_dim (ChafaAlign align, int src_size, int dest_size)
{
    if (src_size > dest_size)
        return 0;
    switch (align)
    {
        case CHAFA_ALIGN_CENTER: return (dest_size - src_size) / 2;
        case CHAFA_ALIGN_END: return dest_size - src_size;
        default: return 0;
    }
}

void
tuck_and_align (int src_width, int src_height, int dest_width, int dest_height,
                ChafaAlign halign, ChafaAlign valign, ChafaTuck tuck,
                int *ofs_x, int *ofs_y, int *width, int *height)
{
    bool fit = (tuck == CHAFA_TUCK_FIT);

    *ofs_x = *ofs_y = 0;

    if (tuck == CHAFA_TUCK_STRETCH)
    {
        *width = dest_width;
        *height = dest_height;
    }
    else if (tuck == CHAFA_TUCK_SHRINK_TO_FIT && src_width <= dest_width && src_height <= dest_height)
    {
        *width = src_width;
        *height = src_height;
    }
    else
    {
        fit = true;
    }

    if (fit)
    {
        float rx = (float) dest_width / (float) src_width;
        float ry = (float) dest_height / (float) src_height;
        float r = rx < ry ? rx : ry;
        *width = (int) ceilf (src_width * r);
        *height = (int) ceilf (src_height * r);
    }

    *width = std::min (*width, dest_width);
    *height = std::min (*height, dest_height);
    *ofs_x = align_dim (halign, *width, dest_width);
    *ofs_y = align_dim (valign, *height, dest_height);
}

/* ---- dither textures ---- */

static void
fill_bayer_r (int *m, int size, int sub, int x, int y, int value, int step)
{
    if (sub == 1)
    {
        m [x + y * size] = value;
        return;
    }
    int half = sub / 2;
    fill_bayer_r (m, size, half, x, y, value, step * 4);
    fill_bayer_r (m, size, half, x + half, y + half, value + step, step * 4);
    fill_bayer_r (m, size, half, x + half, y, value + step * 2, step * 4);
    fill_bayer_r (m, size, half, x, y + half, value + step * 3, step * 4);
}

void
gen_bayer_matrix (int *out, int matrix_size, float magnitude)
{
    int maxval = matrix_size * matrix_size;
    float maxval_f = (float) maxval;

    fill_bayer_r (out, matrix_size, matrix_size, 0, 0, 0, 1);
    for (int i = 0; i < maxval; i++)
        out [i] = (int) (((float) out [i] - maxval_f / 2.0f) * (256.0f / maxval_f) * magnitude + 0.5f);
}

void
gen_noise_matrix (int *out, float magnitude)
{
    for (int i = 0; i < 64 * 64 * 3; i++)
        out [i] = (int) (((float) gen_blue_noise [i] - 128.0f) * magnitude + 0.5f);
}
