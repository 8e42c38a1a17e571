/*
 * mer_multi.cu — Integrator::render on all the GPUs of one box.
 *
 * What it replaces (paths relative to the MitsubaER tree): the scheduler behind SamplingIntegrator::render —
 *   src/librender/integrator.cpp:95-127  (BlockedRenderProcess, sched->schedule / wait)
 *   src/librender/renderproc.cpp:142-148 (worker results merged into the film under a mutex: Film::put(block))
 *   src/libcore/sched.cpp / sched_remote.cpp (LocalWorker threads, RemoteWorker TCP/SSH streams)
 * The reference hands 32x32 image blocks to one worker thread per core and adds the blocks into the film.  Here the
 * unit of work is the SAMPLE INDEX: GPU g of G renders, for every pixel, the samples s = begin + (g + k G) stride, with the
 * scene's grids replicated on every GPU (every GPU sees the whole image: equal load wherever the medium projects, no
 * filter-footprint halo), one host thread per GPU drives mer_render_device, and the per-GPU films — linear in
 * (sum of w RGB, sum of w) — are added on GPU 0 with ONE ncclReduce over NVLink (libnccl.so.2 is loaded at run time;
 * without it the films are copied peer to peer and added by a kernel).
 */
#include <dlfcn.h>

#include <cstring>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "mer_internal.h"

namespace {

__global__ void k_film_add(float *__restrict__ dst, const float *__restrict__ src, size_t n) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) dst[i] += src[i];
}

/* the few NCCL entry points used, resolved from libnccl.so.2 on first use (nccl.h is not needed to build) */
struct Nccl {
    typedef struct ncclComm *comm_t;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    Nccl() {
        /* The NCCL the process already has (PyTorch's, the host application's) is the one to use: a second libnccl.so.2
         * cannot be loaded next to it (same SONAME), and loading the system's first would hand an older library to a
         * torch imported LATER (its libtorch_cuda.so then fails on ncclDevCommCreate).  Only a process without NCCL loads
         * one here: MER_NCCL_LIB if set, else the loader's libnccl.so.2 — with RTLD_LOCAL. */
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) if (const char *e = getenv("MER_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        CommInitAll = (decltype(CommInitAll)) dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy)) dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart)) dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd)) dlsym(h, "ncclGroupEnd");
        Reduce = (decltype(Reduce)) dlsym(h, "ncclReduce");
        GetErrorString = (decltype(GetErrorString)) dlsym(h, "ncclGetErrorString");
        ok = CommInitAll && CommDestroy && GroupStart && GroupEnd && Reduce;
    }
};
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 }; /* ncclFloat32, ncclSum (nccl.h) */

struct Worker {
    int rc = MER_OK;
    std::string err;
    float *film = nullptr;
    cudaStream_t stream = nullptr;
    mer_render_stats st;
};

} /* namespace */

extern "C" int mer_render_multi(const mer_medium *const *media, int32_t ngpus, const mer_render_desc *r, float *film_host,
                                mer_render_stats *stats_out) {
    MER_REQUIRE(media && r && film_host && ngpus >= 1, "null argument or no GPU");
    MER_REQUIRE(r->width > 0 && r->height > 0 && r->sample_stride >= 1 && r->sample_begin >= 0, "bad film size or sample sharding");
    for (int g = 0; g < ngpus; g++) MER_REQUIRE(media[g], "null medium handle");
    const size_t count = (size_t) r->width * r->height * (3 * (size_t) ((r->frames > 1 && !r->modulation) ? r->frames : 1) + 2);
    const size_t bytes = count * sizeof(float);
    std::vector<Worker> W((size_t) ngpus);
    bool distinct = true; /* one communicator rank per device: NCCL needs distinct devices */
    for (int a = 0; a < ngpus; a++) for (int b = a + 1; b < ngpus; b++) if (media[a]->device == media[b]->device) distinct = false;

    /* ---- render: one host thread per GPU, sample indices interleaved */
    auto body = [&](int g) {
        Worker &w = W[(size_t) g];
        memset(&w.st, 0, sizeof(w.st));
        cudaError_t e = cudaSetDevice(media[g]->device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&w.film, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(w.film, 0, bytes, w.stream);
        if (e != cudaSuccess) { w.rc = MER_ERR_CUDA; w.err = cudaGetErrorString(e); return; }
        mer_render_desc d = *r;
        d.sample_begin = r->sample_begin + g * r->sample_stride;
        d.sample_stride = r->sample_stride * ngpus;
        w.rc = mer_render_device(media[g], &d, w.film, &w.st, w.stream);
        if (w.rc != MER_OK) w.err = mer_last_error();
        else if (cudaStreamSynchronize(w.stream) != cudaSuccess) { w.rc = MER_ERR_CUDA; w.err = "stream synchronisation failed"; }
    };
    {
        std::vector<std::thread> threads;
        for (int g = 1; g < ngpus; g++) threads.emplace_back(body, g);
        body(0);
        for (auto &t : threads) t.join();
    }
    int rc = MER_OK;
    std::string err;
    for (int g = 0; g < ngpus; g++) if (W[(size_t) g].rc != MER_OK && rc == MER_OK) { rc = W[(size_t) g].rc; err = "GPU " + std::to_string(media[g]->device) + ": " + W[(size_t) g].err; }

    /* ---- film reduce onto GPU 0 */
    if (rc == MER_OK && ngpus > 1) {
        bool reduced = false;
        const char *env = getenv("MER_NCCL"); /* MER_NCCL=0 forces the peer-copy path */
        const bool wantNccl = distinct && !(env && !strcmp(env, "0"));
        static Nccl *ncclLib = nullptr; /* loaded on the first reduce that can use it, never for shards of one device */
        static std::once_flag ncclOnce;
        if (wantNccl) std::call_once(ncclOnce, [] { ncclLib = new Nccl(); });
        if (wantNccl && ncclLib && ncclLib->ok) {
            Nccl &nccl = *ncclLib;
            std::vector<Nccl::comm_t> comms((size_t) ngpus, nullptr);
            std::vector<int> devs((size_t) ngpus);
            for (int g = 0; g < ngpus; g++) devs[(size_t) g] = media[g]->device;
            int nr = nccl.CommInitAll(comms.data(), ngpus, devs.data());
            if (nr == 0) {
                nr = nccl.GroupStart();
                for (int g = 0; g < ngpus && nr == 0; g++) {
                    cudaSetDevice(media[g]->device);
                    nr = nccl.Reduce(W[(size_t) g].film, W[(size_t) g].film, count, NCCL_FLOAT32, NCCL_SUM, 0, comms[(size_t) g], W[(size_t) g].stream);
                }
                const int ne = nccl.GroupEnd();
                if (nr == 0) nr = ne;
                for (int g = 0; g < ngpus; g++) { cudaSetDevice(media[g]->device); cudaStreamSynchronize(W[(size_t) g].stream); }
                reduced = nr == 0;
            }
            for (auto c : comms) if (c) nccl.CommDestroy(c);
            if (!reduced) cudaGetLastError();
        }
        if (!reduced) { /* peer copies + one add per GPU (also the path for two shards on one device) */
            cudaSetDevice(media[0]->device);
            float *tmp = nullptr;
            cudaError_t e = cudaMalloc(&tmp, bytes);
            for (int g = 1; g < ngpus && e == cudaSuccess; g++) {
                e = cudaMemcpyPeerAsync(tmp, media[0]->device, W[(size_t) g].film, media[g]->device, bytes, W[0].stream);
                if (e == cudaSuccess) {
                    k_film_add<<<148 * 8, 256, 0, W[0].stream>>>(W[0].film, tmp, count);
                    mer::g_launches.fetch_add(1);
                    e = cudaGetLastError();
                }
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(W[0].stream);
            cudaFree(tmp);
            if (e != cudaSuccess) { rc = MER_ERR_CUDA; err = std::string("film reduce: ") + cudaGetErrorString(e); }
        }
    }
    if (rc == MER_OK) {
        cudaSetDevice(media[0]->device);
        cudaError_t e = cudaMemcpy(film_host, W[0].film, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { rc = MER_ERR_CUDA; err = cudaGetErrorString(e); }
    }
    if (stats_out) {
        memset(stats_out, 0, sizeof(*stats_out));
        for (int g = 0; g < ngpus; g++) {
            const mer_render_stats &s = W[(size_t) g].st;
            stats_out->samples += s.samples; stats_out->ray_steps += s.ray_steps; stats_out->scatter_events += s.scatter_events;
            stats_out->null_collisions += s.null_collisions; stats_out->boundary_exits += s.boundary_exits;
            stats_out->nonfinite_dropped += s.nonfinite_dropped; stats_out->connections += s.connections;
            stats_out->connections_failed += s.connections_failed; stats_out->connection_steps += s.connection_steps;
            stats_out->kernel_launches += s.kernel_launches; stats_out->block_fetches += s.block_fetches;
            stats_out->step_launches += s.step_launches;
            stats_out->passes = std::max(stats_out->passes, s.passes);
            stats_out->device_ms = std::max(stats_out->device_ms, s.device_ms);         /* the GPUs run side by side */
            stats_out->step_kernel_ms = std::max(stats_out->step_kernel_ms, s.step_kernel_ms);
            stats_out->tail_ms = std::max(stats_out->tail_ms, s.tail_ms);
            if (g == 0) stats_out->step_lanes_per_sm = s.step_lanes_per_sm;
        }
    }
    for (int g = 0; g < ngpus; g++) {
        cudaSetDevice(media[g]->device);
        cudaFree(W[(size_t) g].film);
        if (W[(size_t) g].stream) cudaStreamDestroy(W[(size_t) g].stream);
    }
    if (rc != MER_OK) return mer::fail(rc, err);
    return MER_OK;
}
