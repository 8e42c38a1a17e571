/*
 * mer_api.cu — error plumbing, device checks and launch accounting of libmitsubaer_b200.so.
 * Replaces the Log(EError) -> std::runtime_error convention of src/libcore/logger.cpp:100-147
 * with status codes + mer_last_error() (the Mitsuba-side shim rethrows).
 */
#include <cstdlib>
#include <cstring>

#include "mer_internal.h"

namespace mer {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }

/* Grid-sized buffers come from a PRIVATE stream-ordered pool per device (cudaMemPoolCreate): its release threshold is lifted
 * so that a scene torn down and rebuilt reuses the memory, without touching the device's default pool, which the host
 * application (PyTorch, Mitsuba) shares.  mer_trim_memory() hands everything cached back to the driver. */
static std::mutex g_poolLock;
static cudaMemPool_t g_pools[64] = {nullptr};

static cudaError_t device_pool(int dev, cudaMemPool_t *out) {
    std::lock_guard<std::mutex> hold(g_poolLock);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!g_pools[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaError_t e = cudaMemPoolCreate(&g_pools[dev], &props);
        if (e != cudaSuccess) { g_pools[dev] = nullptr; return e; }
        unsigned long long keep = ~0ull; /* cached blocks stay with THIS pool across synchronisations */
        if (const char *env = getenv("MER_POOL_KEEP_BYTES")) keep = strtoull(env, nullptr, 10);
        cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = g_pools[dev];
    return cudaSuccess;
}

cudaError_t pool_malloc(void **p, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool = nullptr;
    e = device_pool(dev, &pool);
    if (e != cudaSuccess) return e;
    e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, (cudaStream_t) 0);
    if (e == cudaErrorMemoryAllocation) { /* give the cached blocks back and try once more */
        cudaGetLastError();
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, (cudaStream_t) 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t) 0); /* usable from any stream from here on */
    return e;
}

/* the caller has made sure nothing still uses the buffer (pool_quiesce once per handle, not once per buffer) */
void pool_free(void *p) {
    if (p) cudaFreeAsync(p, (cudaStream_t) 0);
}
void pool_quiesce() { cudaDeviceSynchronize(); }

void pool_trim(int dev) {
    std::lock_guard<std::mutex> hold(g_poolLock);
    if (dev >= 0 && dev < 64 && g_pools[dev]) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(g_pools[dev], 0);
    }
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

int mer_trim_memory(int device) {
    int rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    mer::array_cache_trim(device);
    mer::pool_trim(device);
    RenderScratch &S = mer::device_scratch(device);
    std::lock_guard<std::mutex> hold(S.lock);
    S.release();
    return MER_OK;
}

} /* extern "C" */
