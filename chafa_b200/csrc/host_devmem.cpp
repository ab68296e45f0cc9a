/* host_devmem.cpp -- device allocations of the library, with optional guard bands.
 *
 * compute-sanitizer is not available on every pool this library is tested on, so the library can
 * check itself for out-of-bounds writes: in guard mode (chafa_b200_set_guard_mode (1)) every
 * device allocation is made exactly as large as asked for, framed by two 4 KB bands filled with a
 * pattern, and chafa_b200_check_guards () reads all bands back and counts the allocations whose
 * bands were written to.  The parity tests run a cross-section of configurations this way
 * (tests/test_guard_gpu.py). */

#include "internal.h"

#include <cuda_runtime.h>

#include <mutex>

#define GUARD_BYTES 4096
#define GUARD_PATTERN 0xa5

static std::mutex g_mutex;
static bool g_guard_mode = false;
static std::map<void *, size_t> g_guarded;      /* user pointer -> user size */
static int g_violations_freed = 0;              /* allocations found damaged when they were freed */

/* 1 if a band of the allocation was written to (-1 on a CUDA error); caller holds the lock and has
 * synchronised the device */
static int
bands_damaged (const void *user_ptr, size_t user_size)
{
    static std::vector<uint8_t> band (GUARD_BYTES);
    const uint8_t *user = (const uint8_t *) user_ptr;
    const uint8_t *bands [2] = { user - GUARD_BYTES, user + user_size };

    for (int k = 0; k < 2; k++)
    {
        if (cudaMemcpy (band.data (), bands [k], GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess)
            return -1;
        for (int i = 0; i < GUARD_BYTES; i++)
            if (band [i] != GUARD_PATTERN)
            {
                fprintf (stderr, "chafa-b200: guard band %s allocation %p (%zu bytes) overwritten at offset %d\n",
                         k ? "after" : "before", user_ptr, user_size, k ? i : i - GUARD_BYTES);
                return 1;
            }
    }
    return 0;
}

int
b200_dev_malloc (void **p, size_t n)
{
    std::lock_guard<std::mutex> lock (g_mutex);
    *p = nullptr;
    if (!g_guard_mode)
        return (int) cudaMalloc (p, n);

    /* the start stays 256-byte aligned; the end is exact up to 16 bytes (vector accesses) */
    size_t user = (n + 15) & ~(size_t) 15;
    uint8_t *base = nullptr;
    cudaError_t err = cudaMalloc ((void **) &base, user + 2 * GUARD_BYTES);
    if (err != cudaSuccess)
        return (int) err;
    cudaMemset (base, GUARD_PATTERN, GUARD_BYTES);
    cudaMemset (base + GUARD_BYTES + user, GUARD_PATTERN, GUARD_BYTES);
    *p = base + GUARD_BYTES;
    g_guarded [*p] = user;
    return 0;
}

void
b200_dev_free (void *p)
{
    if (!p)
        return;
    std::lock_guard<std::mutex> lock (g_mutex);
    auto it = g_guarded.find (p);
    if (it != g_guarded.end ())
    {
        cudaDeviceSynchronize ();
        if (bands_damaged (p, it->second) != 0)
            g_violations_freed++;
        g_guarded.erase (it);
        cudaFree ((uint8_t *) p - GUARD_BYTES);
        return;
    }
    cudaFree (p);
}

bool
b200_dev_guard_mode ()
{
    return g_guard_mode;
}

extern "C" {

void
chafa_b200_set_guard_mode (int enabled)
{
    std::lock_guard<std::mutex> lock (g_mutex);
    g_guard_mode = enabled != 0;
}

int
chafa_b200_check_guards (void)
{
    std::lock_guard<std::mutex> lock (g_mutex);
    int bad = g_violations_freed;

    if (cudaDeviceSynchronize () != cudaSuccess)
        return -1;
    for (auto &kv : g_guarded)
    {
        int r = bands_damaged (kv.first, kv.second);
        if (r < 0)
            return -1;
        bad += r;
    }
    return bad;
}

} /* extern "C" */
