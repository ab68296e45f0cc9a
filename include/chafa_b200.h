/* chafa_b200.h -- extra C-ABI entry points of the B200-native canvas renderer.
 *
 * These have no counterpart in the reference: they expose what a GPU-resident
 * pipeline can do that a CPU library cannot (source pixels and printed output
 * staying in HBM, a caller-supplied CUDA stream, row bands for multi-GPU
 * stills, batches of frames) plus read-backs used by the parity tests.  Plain
 * pointers and sizes only; no torch / CUDA types in the signatures (a stream is
 * passed as a void * holding a cudaStream_t).
 *
 * The reference-facing drop-in API is in chafa.h.
 */
#ifndef CHAFA_B200_H
#define CHAFA_B200_H

#include "chafa.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Library / device state ------------------------------------------------------ */

/* 1 if a CUDA device with an sm_100-class kernel image is usable, else 0 (then
 * chafa_canvas_new returns NULL). */
CHAFA_API int chafa_b200_available (void);
/* Select the CUDA device used by canvases created afterwards on this thread. */
CHAFA_API int chafa_b200_set_device (int device);
/* Last error text (thread-local), "" if none. */
CHAFA_API const char *chafa_b200_last_error (void);
/* Number of kernels this library has launched since load (all canvases). */
CHAFA_API uint64_t chafa_b200_launch_count (void);
/* Of those, launches of the row-staged scaler (source tiles brought to shared memory by TMA;
 * scale.cu): lets tests and benchmarks tell which scaler a geometry ran on. */
CHAFA_API uint64_t chafa_b200_staged_scale_count (void);
/* Synchronous device-to-host copy of @n_bytes (e.g. of the text returned by
 * chafa_b200_canvas_print_device ()), for callers without a CUDA runtime of their own. */
CHAFA_API int chafa_b200_copy_from_device (void *host_dst, const void *device_src, size_t n_bytes);
CHAFA_API int chafa_b200_copy_to_device (void *device_dst, const void *host_src, size_t n_bytes);
/* Device memory on the device selected with chafa_b200_set_device (), for callers that want a
 * source image or an exchange buffer in HBM without linking a CUDA runtime of their own. */
CHAFA_API void *chafa_b200_device_alloc (size_t n_bytes);
CHAFA_API void chafa_b200_device_free (void *d_ptr);
/* Free a GString / buffer returned by this library when GLib is not linked. */
CHAFA_API void chafa_b200_string_free (GString *string);
CHAFA_API void chafa_b200_free (void *p);

/* Self-check for out-of-bounds writes (where compute-sanitizer is not available): in guard mode
 * every device allocation made afterwards is exact-sized and framed by 4 KB pattern bands;
 * chafa_b200_check_guards () synchronises the device, reads the bands of all live allocations back
 * and returns how many allocations -- live ones, and ones checked when they were freed -- had a band
 * written to since the library was loaded (-1 on a CUDA error). */
CHAFA_API void chafa_b200_set_guard_mode (int enabled);
CHAFA_API int chafa_b200_check_guards (void);
/* Test hooks of that self-check: the device address of a canvas's cell records, and a one-byte
 * device store. */
CHAFA_API void *chafa_b200_debug_cells_address (ChafaCanvas *canvas);
CHAFA_API void chafa_b200_debug_poke (void *d_ptr, int value);

/* Canvas: device-resident operation ---------------------------------------------- */

/* Run this canvas's kernels on @cuda_stream (a cudaStream_t; NULL = default stream). */
CHAFA_API void chafa_b200_canvas_set_stream (ChafaCanvas *canvas, void *cuda_stream);

/* Wait until everything queued on the canvas's stream has run (the device entry points below only
 * queue work).  Returns 1 on success. */
CHAFA_API int chafa_b200_canvas_synchronize (ChafaCanvas *canvas);

/* Same as chafa_canvas_draw_all_pixels () but @d_src_pixels is DEVICE memory
 * (reference entry point it extends: chafa/chafa-canvas.c:786-805).  Asynchronous:
 * returns once the work is enqueued on the canvas's stream. */
CHAFA_API void chafa_b200_canvas_draw_all_pixels_device (ChafaCanvas *canvas,
                                                         ChafaPixelType src_pixel_type,
                                                         const void *d_src_pixels,
                                                         int src_width, int src_height,
                                                         int src_rowstride);

/* Draw @n_canvases canvases (all on one device), canvas i from the device image @d_src_pixels [i];
 * all sources share one geometry and pixel type.  Equivalent to calling
 * chafa_b200_canvas_draw_all_pixels_device () on each, and the results are the same bytes; the
 * difference is in how sixel canvases with Floyd-Steinberg diffusion are queued.  A diffusion pass
 * is one dependency chain over the image (chafa/internal/chafa-indexed-image.c:172-304: serpentine
 * rows, every pixel needs its predecessor's error), so it occupies one SM for a quarter of a second,
 * and chains queued one per stream stop overlapping beyond the 32 hardware queues of a process.
 * Here the chains of all the canvases go out as ONE launch, a block per image -- up to one chain per
 * SM at once.  Every canvas's stream is ordered after that launch, so each canvas is printed as
 * usual afterwards.  Returns 1 on success. */
CHAFA_API int chafa_b200_canvases_draw_batch_device (ChafaCanvas **canvases, int n_canvases,
                                                     ChafaPixelType src_pixel_type,
                                                     const void *const *d_src_pixels,
                                                     int src_width, int src_height, int src_rowstride);

/* The same for HOST images (what chafa_canvas_draw_all_pixels () takes): canvas i is drawn from
 * @src_pixels [i]; the sources are read before the call returns. */
CHAFA_API int chafa_b200_canvases_draw_batch (ChafaCanvas **canvases, int n_canvases,
                                              ChafaPixelType src_pixel_type,
                                              const void *const *src_pixels,
                                              int src_width, int src_height, int src_rowstride);

/* Same as chafa_canvas_print () but the bytes stay in HBM: *d_out_ptr receives a
 * device pointer owned by the canvas (valid until the next print), the return
 * value is the length in bytes.  Synchronises the stream once (to learn the
 * length). @term_info may be NULL. */
CHAFA_API size_t chafa_b200_canvas_print_device (ChafaCanvas *canvas, ChafaTermInfo *term_info,
                                                 void **d_out_ptr);

/* As chafa_b200_canvas_print_device (), but nothing is waited for or read back: the printer is
 * queued on the canvas's stream and the call returns.  *d_out receives the device address of the
 * text, *d_len the device address of its length in bytes (one 64-bit word); both are valid for
 * work queued on the same stream after this call and until the canvas draws or prints again.
 * For device-side consumers that render a stream of frames without a host round trip per frame.
 * Returns 1 on success. */
CHAFA_API int chafa_b200_canvas_print_device_enqueue (ChafaCanvas *canvas, ChafaTermInfo *term_info,
                                                      void **d_out, const unsigned long long **d_len);

/* As chafa_b200_canvas_print_device (), but the text is copied into the caller's DEVICE
 * buffer @d_dest (e.g. a slot of a gather buffer).  Returns the length; nothing is copied
 * if it exceeds @capacity. */
CHAFA_API size_t chafa_b200_canvas_print_into (ChafaCanvas *canvas, ChafaTermInfo *term_info, void *d_dest,
                                               size_t capacity);

/* Per-stage device timing (CUDA events on the canvas's stream).  Stage indices:
 * 0 H2D source copy, 1 scale, 2 preprocess, 3 match, 4 match-wide, 5 resolve,
 * 6 print-measure, 7 print-emit, 8 D2H text copy.  Enabling resets the sums. */
CHAFA_API void chafa_b200_canvas_set_kernel_timing (ChafaCanvas *canvas, int enabled);
CHAFA_API int chafa_b200_canvas_get_kernel_times (ChafaCanvas *canvas, double *ms_out, uint64_t *count_out,
                                                  int max_out);

/* Symbol matching runs on the thread-per-cell tensor-core kernel (match_tc.cu) wherever it applies
 * (default).  0 selects the warp-per-cell kernels (match.cu) instead; both are byte-identical to the
 * reference, the switch exists for A/B measurements and tests. */
CHAFA_API void chafa_b200_canvas_set_tensor_match (ChafaCanvas *canvas, int enabled);

/* Parity read-backs ------------------------------------------------------------- */

/* Copy the canvas's cell records to @cells_out: width*height records of
 * { uint32 c, uint32 fg_color, uint32 bg_color } (layout of the reference's
 * ChafaCanvasCell, chafa/internal/chafa-canvas-internal.h:30-37). */
CHAFA_API int chafa_b200_canvas_get_cells (ChafaCanvas *canvas, void *cells_out);
/* Copy the prepared pixmap (output of the scale/preprocess kernels; RGBA8,
 * width_px*height_px) of the last draw to @pixels_out.  Returns 0 if none. */
CHAFA_API int chafa_b200_canvas_get_prepared_pixels (ChafaCanvas *canvas, void *pixels_out,
                                                     int *width_px_out, int *height_px_out);
/* Compiled symbol table of a map, in evaluation order (parity with
 * chafa/chafa-symbol-map.c:384-470).  Returns the count. */
CHAFA_API int chafa_b200_symbol_map_get_compiled (const ChafaSymbolMap *map, int wide, int max_out,
                                                  uint32_t *c_out, uint64_t *bitmap0_out,
                                                  uint64_t *bitmap1_out);
/* Internal canvas facts for tests: out[0..11] = width_px, height_px,
 * work_factor_int, blank_char, solid_char, consider_inverted, extract_colors,
 * use_quantized_error, n_symbols, n_symbols2, n_fill_symbols, n_fill_symbols2. */
CHAFA_API void chafa_b200_canvas_get_info (ChafaCanvas *canvas, int *out);

/* The device's table of DIN99d colours (2^24 entries: entry [b << 16 | g << 8 | r] holds the
 * converted colour in the same byte order, alpha byte 0), copied to @host_out (64 MB); parity with
 * chafa_color_rgb_to_din99d (), chafa/internal/chafa-color.c:130-168. */
CHAFA_API int chafa_b200_get_din99d_table (uint32_t *host_out);

/* Sixel read-back: indexed image of the last sixel draw (width*height pen bytes). */
CHAFA_API int chafa_b200_canvas_get_sixel_image (ChafaCanvas *canvas, uint8_t *pixels_out,
                                                 int *width_out, int *height_out,
                                                 int *n_colors_out, uint32_t *palette_out);

/* Multi-GPU stills: render only cell rows [first_row, first_row + n_rows) of the
 * canvas from the full source image (row-band sharding).
 * The printed string of a band is the band's rows exactly as they appear in the
 * full print, including the row-0 leading reset and the newline rule. */
CHAFA_API void chafa_b200_canvas_set_band (ChafaCanvas *canvas, int first_row, int n_rows);

/* Assembling the text of a band-sharded still without a host round trip.  Every band owner
 * prints with chafa_b200_canvas_print_device_enqueue (), the band lengths are exchanged on the
 * device (e.g. ncclAllGather of the words chafa_b200_canvas_copy_print_length () delivers), and
 * chafa_b200_canvas_band_scatter () queues a kernel that copies this canvas's text to its place
 * -- the sum of the lengths of the bands before it -- in @d_full_text, which may be memory of
 * another GPU of the box mapped with chafa_b200_ipc_open () (the stores then travel over NVLink).
 * @d_total_out (may be NULL) receives the sum of all lengths.  Nothing is copied if the total
 * exceeds @capacity.  All calls only queue work on the canvas's stream. */
CHAFA_API int chafa_b200_canvas_copy_print_length (ChafaCanvas *canvas, unsigned long long *d_len_out);
CHAFA_API int chafa_b200_canvas_band_scatter (ChafaCanvas *canvas, const unsigned long long *d_lengths, int n_bands,
                                              int band, void *d_full_text, size_t capacity,
                                              unsigned long long *d_total_out);
/* CUDA IPC for the above: a 64-byte handle of a device allocation of this process, and the
 * mapping of another process's allocation into this one (on the device selected with
 * chafa_b200_set_device ()). */
CHAFA_API int chafa_b200_ipc_export (const void *d_ptr, void *handle_out_64);
CHAFA_API void *chafa_b200_ipc_open (const void *handle_64);
CHAFA_API void chafa_b200_ipc_close (void *d_peer_ptr);

/* Row bands of a sixel canvas (chafa_b200_canvas_set_band (); the band must start on a multiple of
 * six pixel rows and end on one or at the bottom of the canvas).  The printed string of a band is
 * its sixel bands as they appear in the full print; the header (raster attributes, palette) goes
 * with the band that holds row 0, the terminator with the band that holds the last row.
 *
 * With a generated palette the colour-cube bins of chafa_palette_generate ()
 * (chafa/internal/chafa-palette.c:513-536: per bin the sums of the three channels and the sample
 * count) have to be those of the whole image, so the draw is two-phase:
 *
 *   chafa_b200_canvas_draw_begin_device ()     scales the band and samples its rows; returns 1 and
 *                                              the device address of the bins (*n_int32_out words
 *                                              of exact integer sums, to be summed across the band
 *                                              owners: ncclAllReduce, ncclSum, 32-bit integers)
 *   chafa_b200_canvas_draw_float_sums_device () once per owner, in band order from the top: bins
 *                                              with 65 794 samples or more leave the range in which
 *                                              the reference's float sums are exact, so those sums
 *                                              are accumulated sample by sample in image order,
 *                                              each owner continuing from the sums the owner above
 *                                              handed on (*d_sums_out, *n_bytes_out: the bytes to
 *                                              pass to the next owner; after the last owner, to
 *                                              everybody -- e.g. one broadcast per owner, in order)
 *   chafa_b200_canvas_draw_finish ()           merges the bins into the palette (every owner
 *                                              computes the same one), quantises the band
 *
 * The buffer can be the caller's (chafa_b200_canvas_set_histogram_buffer (), at least
 * chafa_b200_canvas_exchange_buffer_bytes () bytes).  Error diffusion is one chain over the image:
 * a band owner then computes the whole image and prints its rows, and begin returns 0. */
CHAFA_API int chafa_b200_canvas_draw_float_sums_device (ChafaCanvas *canvas, void **d_sums_out, size_t *n_bytes_out);
CHAFA_API size_t chafa_b200_canvas_exchange_buffer_bytes (ChafaCanvas *canvas);

/* Two-phase draw for row bands of canvas modes that normalise against a whole-image
 * intensity histogram (16 / 8 colours, FGBG; chafa/internal/chafa-pixops.c:90-140, 519-566):
 * begin queues scaling and the histogram pass over this canvas's band and hands out the
 * device histogram (int32 [2049]: 2048 bins + sample count); the caller sums it across the
 * band owners (e.g. ncclAllReduce on the canvas's stream) and calls finish, which derives
 * the crop bounds from the summed histogram and queues the rest of the pipeline.
 * Returns 1 if a reduction is needed (otherwise *d_histogram_out is NULL; finish must
 * still be called).  Symbols mode, and sixel mode as described above. */
CHAFA_API int chafa_b200_canvas_draw_begin_device (ChafaCanvas *canvas, ChafaPixelType src_pixel_type,
                                                   const void *d_src_pixels, int src_width, int src_height,
                                                   int src_rowstride, void **d_histogram_out,
                                                   size_t *n_int32_out);
CHAFA_API void chafa_b200_canvas_draw_finish (ChafaCanvas *canvas);
/* Keep the histogram (and the two crop bounds after it) in a caller-owned DEVICE buffer of
 * 2051 int32, e.g. a tensor the caller's collective library can reduce in place. */
CHAFA_API void chafa_b200_canvas_set_histogram_buffer (ChafaCanvas *canvas, void *d_int32_2051);

#ifdef __cplusplus
}
#endif

#endif /* CHAFA_B200_H */
