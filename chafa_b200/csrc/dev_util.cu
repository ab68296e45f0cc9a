/* dev_util.cu -- small launch-side helpers shared by the kernel files. */

#include <cuda_runtime.h>

#include <mutex>
#include <unordered_map>

#include "device.h"

/* Opt a kernel in to @smem bytes of dynamic shared memory on the current device.  The attribute is
 * per (function, device) and only ever needs raising; asking the runtime costs about a microsecond
 * of host time per launch, which is a quarter of a launch on the short frames, so what has been
 * granted is remembered. */
int
dev_opt_in_smem (const void *func, size_t smem)
{
    static std::mutex mutex;
    static std::unordered_map<unsigned long long, size_t> granted;

    if (smem <= 48 * 1024)
        return 0;
    int dev = 0;
    cudaGetDevice (&dev);
    const unsigned long long key = (unsigned long long) (uintptr_t) func * 64u + (unsigned long long) (dev & 63);
    {
        std::lock_guard<std::mutex> lock (mutex);
        auto it = granted.find (key);
        if (it != granted.end () && it->second >= smem)
            return 0;
    }
    cudaError_t err = cudaFuncSetAttribute (func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (err != cudaSuccess)
        return (int) err;
    std::lock_guard<std::mutex> lock (mutex);
    size_t &g = granted [key];
    if (g < smem)
        g = smem;
    return 0;
}

/* Number of SMs of the current device */
int
dev_sm_count ()
{
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    cudaGetDevice (&dev);
    if (dev != cached_dev)
    {
        int sms = 0;
        cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev);
        cached_sms = sms > 0 ? sms : 148;
        cached_dev = dev;
    }
    return cached_sms;
}
