/* band_gather.cu -- the one exchange step of a band-sharded still: every band owner copies its
 * printed text to its place in the assembled text on the gathering GPU.
 *
 * The destination is peer memory (another GPU of the box, mapped through CUDA IPC), so the copy is
 * plain 16-byte stores that travel over NVLink; the place is the sum of the lengths of the bands
 * before this one, read from the all-gathered length words -- no host is involved and nothing
 * waits for anything: printer, all-gather and this kernel are queued back to back per frame.
 * (The reference concatenates its row batches in an ordered post step, chafa/internal/chafa-batch.c:
 * 150-178; this is that step across GPUs.) */

#include <cuda_runtime.h>

#include "device.h"

__global__ void __launch_bounds__ (256)
band_scatter_kernel (const uint8_t *__restrict__ text, const uint64_t *__restrict__ lengths, int n_bands, int band,
                     uint8_t *__restrict__ full_text, size_t capacity, uint64_t *__restrict__ total_out)
{
    uint64_t before = 0, total = 0;
    for (int i = 0; i < n_bands; i++)
    {
        if (i < band)
            before += lengths [i];
        total += lengths [i];
    }
    if (total_out && blockIdx.x == 0 && threadIdx.x == 0)
        *total_out = total;
    if (total > capacity)
        return;

    const uint64_t len = lengths [band];
    uint8_t *dst = full_text + before;
    const size_t tid = (size_t) blockIdx.x * blockDim.x + threadIdx.x, n_threads = (size_t) gridDim.x * blockDim.x;

    /* head up to the destination's 16-byte phase, body in 16-byte stores, tail */
    size_t head = (16 - ((uintptr_t) dst & 15)) & 15;
    if (head > len)
        head = len;
    for (size_t i = tid; i < head; i += n_threads)
        dst [i] = text [i];
    const size_t n_vec = (len - head) / 16;
    const bool src_aligned = (((uintptr_t) (text + head)) & 15) == 0;
    for (size_t v = tid; v < n_vec; v += n_threads)
    {
        const uint8_t *src = text + head + v * 16;
        uint4 w;
        if (src_aligned)
            w = *(const uint4 *) src;
        else
        {
            uint32_t t [4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                t [k] = (uint32_t) src [4 * k] | ((uint32_t) src [4 * k + 1] << 8) | ((uint32_t) src [4 * k + 2] << 16)
                    | ((uint32_t) src [4 * k + 3] << 24);
            w = make_uint4 (t [0], t [1], t [2], t [3]);
        }
        *(uint4 *) (dst + head + v * 16) = w;
    }
    for (size_t i = head + n_vec * 16 + tid; i < len; i += n_threads)
        dst [i] = text [i];
}

int
dev_launch_band_scatter (const uint8_t *text, const uint64_t *lengths, int n_bands, int band, uint8_t *full_text,
                         size_t capacity, uint64_t *total_out, void *stream)
{
    band_scatter_kernel<<<148, 256, 0, (cudaStream_t) stream>>> (text, lengths, n_bands, band, full_text, capacity,
                                                                total_out);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
