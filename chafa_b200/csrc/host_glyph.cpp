/* host_glyph.cpp -- user glyph import for chafa_symbol_map_add_glyph ().
 *
 * Replaces glyph_to_bitmap / glyph_to_bitmap_wide (chafa/chafa-symbol-map.c:295-354): the
 * glyph image is scaled to 8x8 (16x8 for wide code points) by kernel K1 on its plain route
 * (premultiplied RGBA8 out, no compositing -- what smol_scale_simple does), then the 64 / 128
 * coverage values are sharpened with the reference's 3x3 kernel and thresholded into the
 * bitmap (chafa-symbol-map.c:172-248).  The scaling runs on the device; the 3x3 pass over
 * 64 values is table building and stays on the host like the rest of the symbol map. */

#include "internal.h"
#include "device.h"

#include <cuda_runtime.h>

#include <algorithm>

const void *device_tables_for_current (const uint16_t **from_srgb, const uint8_t **to_srgb, const uint32_t **inv_div);

static bool
ok (cudaError_t err, const char *what)
{
    if (err == cudaSuccess)
        return true;
    b200_set_error ("%s: %s", what, cudaGetErrorString (err));
    fprintf (stderr, "chafa-b200: %s: %s\n", what, cudaGetErrorString (err));
    return false;
}

static uint64_t
coverage_to_bitmap (const uint8_t *cov, int rowstride)
{
    uint64_t bitmap = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
        {
            bitmap <<= 1;
            if (cov [y * rowstride + x] > 127)
                bitmap |= 1;
        }
    return bitmap;
}

bool
glyph_import_device (ChafaPixelType pixel_format, const void *pixels, int width, int height, int rowstride,
                     bool wide, uint64_t *bitmaps_out)
{
    const int dw = wide ? 16 : 8, dh = 8;
    const int bpp = (pixel_format == CHAFA_PIXEL_RGB8 || pixel_format == CHAFA_PIXEL_BGR8) ? 3 : 4;
    const uint16_t *from_srgb;
    const uint8_t *to_srgb;
    const uint32_t *inv_div;
    ScalePlan plan;

    if (width < 1 || height < 1 || rowstride < width * bpp)
        return false;
    if (!device_tables_for_current (&from_srgb, &to_srgb, &inv_div))
    {
        fprintf (stderr, "chafa-b200: chafa_symbol_map_add_glyph needs a CUDA device (no CPU fallback)\n");
        return false;
    }
    if (!scale_plan_build (&plan, width, height, dw, dh, 0, 0, dw, dh, false))
        return false;

    size_t src_bytes = (size_t) (height - 1) * rowstride + (size_t) width * bpp;
    size_t nh = plan.h.precalc.size (), nv = plan.v.precalc.size ();
    uint8_t *d_src = nullptr;
    uint32_t *d_dest = nullptr, *d_pre = nullptr;
    uint32_t scaled [16 * 8];
    bool good = false;

    if (ok (cudaMalloc ((void **) &d_src, src_bytes + 16), "cudaMalloc")
        && ok (cudaMalloc ((void **) &d_dest, sizeof (scaled)), "cudaMalloc")
        && ok (cudaMalloc ((void **) &d_pre, (nh + nv + 1) * 4), "cudaMalloc")
        && ok (cudaMemcpy (d_src, pixels, src_bytes, cudaMemcpyHostToDevice), "cudaMemcpy")
        && (!nh || ok (cudaMemcpy (d_pre, plan.h.precalc.data (), nh * 4, cudaMemcpyHostToDevice), "cudaMemcpy"))
        && (!nv || ok (cudaMemcpy (d_pre + nh, plan.v.precalc.data (), nv * 4, cudaMemcpyHostToDevice), "cudaMemcpy")))
    {
        DevScale sc;
        memset (&sc, 0, sizeof (sc));
        const ScaleDim *dims [2] = { &plan.h, &plan.v };
        DevScaleDim *out [2] = { &sc.h, &sc.v };
        for (int i = 0; i < 2; i++)
        {
            out [i]->filter = dims [i]->filter;
            out [i]->n_halvings = dims [i]->n_halvings;
            out [i]->src_size_px = (int32_t) dims [i]->src_size_px;
            out [i]->placement_ofs_px = dims [i]->placement_ofs_px;
            out [i]->placement_size_px = dims [i]->placement_size_px;
            out [i]->clip_before_px = dims [i]->clip_before_px;
            out [i]->first_opacity = dims [i]->first_opacity;
            out [i]->last_opacity = dims [i]->last_opacity;
            out [i]->span_step = dims [i]->span_step;
            out [i]->span_mul = dims [i]->span_mul;
        }
        sc.h.precalc = d_pre;
        sc.v.precalc = d_pre + nh;
        sc.src = d_src;
        sc.src_rowstride = rowstride;
        sc.src_pixel_type = pixel_format;
        sc.linear = plan.linear;
        sc.dest = d_dest;
        sc.dest_w = dw;
        sc.dest_h = dh;
        sc.first_row = 0;
        sc.n_rows = dh;
        sc.from_srgb = from_srgb;
        sc.to_srgb = to_srgb;
        sc.inv_div = inv_div;
        sc.plain = 1;

        if (ok ((cudaError_t) dev_launch_scale (&sc, nullptr), "scale kernel")
            && ok (cudaMemcpy (scaled, d_dest, sizeof (uint32_t) * dw * dh, cudaMemcpyDeviceToHost), "cudaMemcpy"))
            good = true;
    }
    cudaFree (d_src);
    cudaFree (d_dest);
    cudaFree (d_pre);
    if (!good)
        return false;

    /* coverage: alpha, or the mean of the colour channels for formats without alpha */
    uint8_t cov [16 * 8], sharp [16 * 8];
    for (int i = 0; i < dw * dh; i++)
    {
        uint32_t px = scaled [i];
        if (bpp == 3)
            cov [i] = (uint8_t) (((px & 0xff) + ((px >> 8) & 0xff) + ((px >> 16) & 0xff)) / 3);
        else
            cov [i] = (uint8_t) (px >> 24);
    }

    /* sharpen + boost contrast: centre 6, direct neighbours -1, edges cloned */
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++)
        {
            auto at = [&] (int xx, int yy) -> int
            {
                xx = std::min (std::max (xx, 0), dw - 1);
                yy = std::min (std::max (yy, 0), dh - 1);
                return cov [xx + yy * dw];
            };
            int sum = 6 * at (x, y) - at (x - 1, y) - at (x + 1, y) - at (x, y - 1) - at (x, y + 1);
            sharp [x + y * dw] = (uint8_t) std::min (std::max (sum, 0), 255);
        }

    bitmaps_out [0] = coverage_to_bitmap (sharp, dw);
    bitmaps_out [1] = wide ? coverage_to_bitmap (sharp + 8, dw) : 0;
    return true;
}
