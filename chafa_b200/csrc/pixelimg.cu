/* pixelimg.cu -- kernel K7: the text of a Kitty or iTerm2 image.
 *
 * Both protocols carry the scaled RGBA8 pixmap as base64: Kitty in chunks of 510 bytes
 * (63 under GNU Screen), every chunk between a begin and an end sequence
 * (chafa/internal/chafa-kitty-renderer.c:255-289); iTerm2 as one run over an 8-byte TIFF
 * header, the pixmap and a 146-byte directory (chafa-iterm2-renderer.c:251-329).  The output
 * position of every byte is known in closed form, so one thread writes four output bytes:
 * it finds its chunk, decides between frame bytes and base64 characters, and reads the
 * three source bytes a character depends on (chafa-base64.c:47-115).
 *
 * HBM-bound: reads the pixmap once (4 B/pixel), writes 16/3 B/pixel plus framing. */

#include <cuda_runtime.h>

#include "device.h"

struct ImageTextGeometry
{
    unsigned long long n_stream;      /* pre + image + suf */
    unsigned long long n_chunks;
    unsigned long long chunk_text;    /* bytes one full chunk occupies in the output */
    unsigned long long total;
    uint32_t chunk_chars;             /* base64 characters of a full chunk */
};

static __host__ __device__ __forceinline__ ImageTextGeometry
image_text_geometry (const DevImageText &p)
{
    ImageTextGeometry g;
    g.n_stream = (unsigned long long) p.n_pre + p.n_image + (unsigned long long) p.n_suf;
    g.n_chunks = (g.n_stream + p.chunk_bytes - 1) / p.chunk_bytes;
    g.chunk_chars = p.chunk_bytes / 3 * 4;
    g.chunk_text = (unsigned long long) p.n_frame_begin + g.chunk_chars + (unsigned long long) p.n_frame_end;
    if (g.n_chunks == 0)
    {
        g.total = 0;
        return g;
    }
    unsigned long long last_bytes = g.n_stream - (g.n_chunks - 1) * p.chunk_bytes;
    unsigned long long last_chars = (last_bytes + 2) / 3 * 4;
    g.total = (g.n_chunks - 1) * g.chunk_text + p.n_frame_begin + last_chars + p.n_frame_end;
    return g;
}

__device__ __forceinline__ uint32_t
stream_byte (const DevImageText &p, unsigned long long i)
{
    if (i < (unsigned long long) p.n_pre)
        return p.pre [i];
    i -= p.n_pre;
    if (i < p.n_image)
        return __ldg (p.image + i);
    i -= p.n_image;
    return p.suf [i];
}

__device__ __forceinline__ uint32_t
base64_char (uint32_t v)
{
    /* A-Z a-z 0-9 + / */
    return v < 26 ? 'A' + v : v < 52 ? 'a' + (v - 26) : v < 62 ? '0' + (v - 52) : v == 62 ? '+' : '/';
}

__device__ __forceinline__ uint32_t
output_byte (const DevImageText &p, const ImageTextGeometry &g, unsigned long long o)
{
    unsigned long long chunk = o / g.chunk_text;
    uint32_t r = (uint32_t) (o - chunk * g.chunk_text);

    if (r < (uint32_t) p.n_frame_begin)
        return p.frame_begin [r];
    r -= p.n_frame_begin;

    unsigned long long first = chunk * p.chunk_bytes;
    unsigned long long n_bytes = g.n_stream - first;
    if (n_bytes > p.chunk_bytes)
        n_bytes = p.chunk_bytes;
    uint32_t n_chars = (uint32_t) ((n_bytes + 2) / 3 * 4);

    if (r >= n_chars)
        return p.frame_end [r - n_chars];

    uint32_t group = r >> 2, k = r & 3;
    unsigned long long b = first + (unsigned long long) group * 3;
    uint32_t have = (uint32_t) (n_bytes - (unsigned long long) group * 3);   /* >= 1 */
    uint32_t b0 = stream_byte (p, b);
    uint32_t b1 = have > 1 ? stream_byte (p, b + 1) : 0;
    uint32_t b2 = have > 2 ? stream_byte (p, b + 2) : 0;
    uint32_t v = (b0 << 16) | (b1 << 8) | b2;

    if ((k == 2 && have < 2) || (k == 3 && have < 3))
        return '=';
    return base64_char ((v >> (18 - 6 * k)) & 0x3f);
}

__global__ void __launch_bounds__ (256)
image_text_kernel (const __grid_constant__ DevImageText p)
{
    const ImageTextGeometry g = image_text_geometry (p);
    const unsigned long long n_words = (g.total + 3) / 4;

    for (unsigned long long w = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; w < n_words;
         w += (unsigned long long) gridDim.x * blockDim.x)
    {
        unsigned long long o = w * 4;
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (o + k < g.total)
                word |= output_byte (p, g, o + k) << (8 * k);
        ((uint32_t *) p.out) [w] = word;
    }
}

unsigned long long
dev_image_text_length (const DevImageText *p)
{
    return image_text_geometry (*p).total;
}

int
dev_launch_image_text (const DevImageText *p, void *stream)
{
    unsigned long long total = image_text_geometry (*p).total;
    if (total == 0)
        return 0;

    int dev = 0, sms = 148;
    cudaGetDevice (&dev);
    cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);

    unsigned long long n_words = (total + 3) / 4;
    unsigned long long blocks = (n_words + 255) / 256;
    unsigned long long cap = (unsigned long long) sms * 16;
    image_text_kernel<<<(unsigned) (blocks < cap ? blocks : cap), 256, 0, (cudaStream_t) stream>>> (*p);
    dev_count_launch (1);
    return (int) cudaGetLastError ();
}
