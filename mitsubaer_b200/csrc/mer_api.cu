/*
 * mer_api.cu — error plumbing, device checks and launch accounting of libmitsubaer_b200.so.
 * Replaces the Log(EError) -> std::runtime_error convention of src/libcore/logger.cpp:100-147
 * with status codes + mer_last_error() (the Mitsuba-side shim rethrows).
 */
#include "mer_internal.h"

namespace mer {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }

cudaError_t pool_malloc(void **p, size_t bytes) {
    static std::atomic<unsigned> configured{0}; /* bit per device */
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = nullptr;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    if (dev < 32 && !(configured.load() & (1u << dev))) {
        unsigned long long keep = ~0ull; /* never hand freed blocks back to the driver on a synchronisation */
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        configured.fetch_or(1u << dev);
    }
    e = cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t) 0);
    if (e == cudaErrorMemoryAllocation) { /* give the cached blocks back and try once more */
        cudaGetLastError();
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t) 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) 0); /* usable from any stream from here on */
    return e;
}

void pool_free(void *p) {
    if (!p) return;
    cudaDeviceSynchronize(); /* cudaFree's guarantee: nothing still uses the buffer */
    cudaFreeAsync(p, (cudaStream_t) 0);
}

int fail(int code, const std::string &msg) {
    t_error = msg;
    return code;
}

/* There is no CPU fallback anywhere in this library: without a Blackwell-class GPU every
 * compute entry point fails here, loudly. */
int check_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(MER_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                      "); libmitsubaer_b200 has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(MER_ERR_INVALID, "device index out of range");
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10)
        return fail(MER_ERR_CUDA, "device is not sm_100-class; this library is built for sm_100a only");
    return MER_OK;
}

RenderScratch &device_scratch(int device) {
    static RenderScratch scratch[64];
    return scratch[(device >= 0 && device < 64) ? device : 0];
}

} /* namespace mer */

extern "C" {

const char *mer_last_error(void) { return mer::t_error.c_str(); }

int mer_abi_version(void) { return MER_ABI_VERSION; }

int mer_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < count; d++) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
        if (major == 10) usable++;
    }
    return usable;
}

uint64_t mer_kernel_launch_count(void) { return mer::g_launches.load(); }

} /* extern "C" */
