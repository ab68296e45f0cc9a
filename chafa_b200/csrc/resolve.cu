/* resolve.cu -- kernel K4 as a launch of its own (the row walk itself is in resolve.cuh; the
 * printer runs it as the first phase of its own launch when the cells have not been read in
 * between, see print.cu). */

#include "resolve.cuh"

__global__ void __launch_bounds__ (RESOLVE_THREADS)
resolve_rows_kernel (DevCanvas cv, const DevCellEval *__restrict__ evals,
                     const DevWideEval *__restrict__ wide_evals, DevCell *__restrict__ cells)
{
    resolve_row (cv, evals, wide_evals, cells, (int) blockIdx.x);
}

int
dev_launch_resolve (const DevCanvas *cv, const DevCellEval *evals, const DevWideEval *wide_evals,
                    DevCell *cells, void *stream)
{
    if (cv->n_rows <= 0)
        return 0;
    resolve_rows_kernel<<<cv->n_rows, RESOLVE_THREADS, 0, (cudaStream_t) stream>>> (*cv, evals, wide_evals, cells);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}

/* Host-side call of the same picker code, for the cell colour setters of the public
 * API (chafa/chafa-canvas.c:118-131); not on the rendering path. */
int
host_palette_lookup_nearest (const DevPalette *pal, int color_space, uint32_t color)
{
    return palette_lookup_nearest (pal, color_space, color, nullptr);
}
