/*
 * mer_internal.h — host-side internals shared by the .cu translation units of
 * libmitsubaer_b200.so (handle layouts, error plumbing, launch accounting).
 */
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "mer_device.cuh"
#include "mitsubaer_b200.h"

struct mer_rif {
    int device;
    int mode;
    mer_volume_desc desc;
    RifDev dev;
    float *d_coeff;
    float4 *d_packed;
    float4 *d_coeff8 = nullptr;     /* sector table for the load/store path (layout "coeff8") */
    cudaArray_t texArray = nullptr; /* atlas of the coefficient layers (texture storage) */
    size_t texW = 0, texH = 0;
    cudaTextureObject_t tex = 0;
};

struct mer_grid {
    int device;
    mer_volume_desc desc;
    GridDev dev;
    float *d_data;
};

/* per-DEVICE render scratch (path pool, counters, pinned read-back word), allocated on first use and reused by
 * every later mer_render* call on that device (cudaMalloc / cudaMallocHost / cudaFree cost 0.1-1 s per call);
 * lives until the process exits */
struct RenderScratch {
    std::mutex lock; /* mer_render* calls on one device are serialised */
    size_t poolBytes = 0;
    void *pool[16] = {nullptr};
    void *neeQ[3] = {nullptr, nullptr, nullptr}; /* direct-connection request queue */
    unsigned *neeCount = nullptr, *neePerm = nullptr, *neeHist = nullptr;
    unsigned char *neeKey = nullptr;
    float *neeOpl = nullptr;
    unsigned neeCap = 0;
    unsigned *nOut = nullptr;
    unsigned long long *counters = nullptr;
    unsigned long long *hostPinned = nullptr;
    void *paramsDev = nullptr, *paramsHost = nullptr; /* RenderParams of the pass in flight (global copy + pinned staging) */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evTail = nullptr;
    /* the step kernel's configuration the last frame on this device settled on, and the table it was measured on: the next
     * frame of the same scene starts from it instead of measuring from scratch */
    int stepChoice = 0;
    size_t stepChoiceTable = 0;
    enum { RING = 32 };
    cudaEvent_t ring[2 * RING] = {nullptr}; /* event pairs around the step kernel's launches (mer_render_stats.step_kernel_ms) */
    void release() {
        for (cudaEvent_t &e : ring) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (void *&p : pool) { cudaFree(p); p = nullptr; }
        for (void *&p : neeQ) { cudaFree(p); p = nullptr; }
        cudaFree(neeCount); neeCount = nullptr;
        cudaFree(neePerm); neePerm = nullptr;
        cudaFree(neeHist); neeHist = nullptr;
        cudaFree(neeKey); neeKey = nullptr;
        cudaFree(neeOpl); neeOpl = nullptr;
        neeCap = 0;
        cudaFree(nOut); nOut = nullptr;
        cudaFree(counters); counters = nullptr;
        if (hostPinned) cudaFreeHost(hostPinned);
        hostPinned = nullptr;
        cudaFree(paramsDev); paramsDev = nullptr;
        if (paramsHost) cudaFreeHost(paramsHost);
        paramsHost = nullptr;
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (evTail) cudaEventDestroy(evTail);
        evTail = nullptr;
        ev0 = ev1 = nullptr;
        poolBytes = 0;
    }
};

struct mer_medium {
    int device;
    mer_medium_desc desc; /* resolved */
    const mer_rif *rif;
    const mer_grid *grid;
    MediumDev dev;
};

namespace mer {

void set_error(const std::string &msg);
/* Grid-sized device buffers come from CUDA's stream-ordered pool of the current device with the release threshold lifted:
 * a scene that is torn down and rebuilt (the e2e leg of bench.py, a host that reloads volumes) reuses the memory instead of
 * paying cudaMalloc / cudaFree of half a gigabyte each time (measured: 0.2-0.8 s per rebuild of C2, depending on the box). */
cudaError_t pool_malloc(void **p, size_t bytes);
void pool_free(void *p);     /* stream-ordered free; call pool_quiesce() first, once per handle */
void pool_quiesce();
void pool_trim(int device);
/* CUDA arrays of the coefficient atlases are cached the same way (cudaMallocArray / cudaFreeArray are synchronising) */
cudaError_t array_acquire(int device, size_t W, size_t H, cudaArray_t *out);
void array_release(int device, cudaArray_t arr, size_t W, size_t H);
void array_cache_trim(int device);
int fail(int code, const std::string &msg);
extern std::atomic<uint64_t> g_launches;

/* RAII device switch that restores the caller's current device (PyTorch shares the process) */
struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_device(int device); /* MER_OK or MER_ERR_CUDA (no usable sm_100-class GPU) */
RenderScratch &device_scratch(int device);

/* RAII device buffer of the *_batch (host pointer) entry points: from the private stream-ordered pool, so that an entry
 * point that is called once per ray (the scalar virtuals of the Mitsuba binding) does not pay cudaMalloc / cudaFree, and
 * an early return on an error cannot leak the buffers allocated before it */
struct DevBuf {
    void *ptr = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { pool_free(ptr); }
    cudaError_t alloc(size_t bytes) { return pool_malloc(&ptr, bytes ? bytes : 1); }
    template <typename T> T *as() { return (T *) ptr; }
};

} /* namespace mer */

#define MER_CUDA(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            char _buf[512];                                                                             \
            snprintf(_buf, sizeof(_buf), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                                               \
            return mer::fail(_e == cudaErrorMemoryAllocation ? MER_ERR_OOM : MER_ERR_CUDA, _buf);       \
        }                                                                                               \
    } while (0)

/* every kernel launch goes through this so mer_kernel_launch_count() is exact */
#define MER_LAUNCH(kernel, grid, block, smem, stream, ...)                                              \
    do {                                                                                                \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                     \
        mer::g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
        MER_CUDA(cudaGetLastError());                                                                   \
    } while (0)

#define MER_REQUIRE(cond, msg)                                                                          \
    do {                                                                                                \
        if (!(cond)) return mer::fail(MER_ERR_INVALID, msg);                                            \
    } while (0)

static inline unsigned mer_blocks(size_t n, unsigned threads) { return (unsigned) ((n + threads - 1) / threads); }
