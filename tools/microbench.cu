/*
 * tools/microbench.cu — the two ceilings SURVEY §8d asks the builder to MEASURE on the box instead of quoting a datasheet:
 *   (1) FP32 FMA issue peak   (the secondary bound of the tricubic stepper, ~1.0 kFLOP per ray-step)
 *   (2) L2 read bandwidth     (the fetch roofline when the coefficient table fits the 126 MB L2: C1)
 * plus an HBM read figure for cross-checking MEASURED_PEAKS.json.  Prints one JSON object.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_fma(float *out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (float) (threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 123.456f) out[0] = s; /* never true: keeps the chain alive */
}

/* the same chains with sm_100's packed FP32 FMA (fma.rn.f32x2: two FMAs per instruction on a 64-bit register pair) */
__global__ void __launch_bounds__(256) k_fma2(float *out, int iters, float a, float b) {
    unsigned long long x[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float lo = (float) (threadIdx.x + 2 * i) * 1e-3f, hi = (float) (threadIdx.x + 2 * i + 1) * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

/* every CTA streams the SAME window again and again: after the first sweep it is served by L2 (window << 126 MB)
 * or by HBM (window >> L2) */
__global__ void __launch_bounds__(256) k_read(const uint4 *__restrict__ buf, size_t nvec, int sweeps, unsigned *sink) {
    unsigned acc = 0;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (int s = 0; s < sweeps; s++) {
        /* rotate the starting point per sweep so that L1 cannot serve consecutive sweeps */
        size_t i = ((size_t) blockIdx.x * blockDim.x + threadIdx.x + (size_t) s * 977u * blockDim.x) % nvec;
        for (size_t k = 0; k < nvec / stride; k++) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc += v.x ^ v.y ^ v.z ^ v.w;
            i += stride;
            if (i >= nvec) i -= nvec;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *out; unsigned *sink;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&sink, 4));

    /* (1) FP32 FMA */
    const int iters = 1 << 15, blocks = sms * 8;
    k_fma<<<blocks, 256>>>(out, 1024, 1.0001f, 1e-6f);
    double best_fma = 0;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        k_fma<<<blocks, 256>>>(out, iters, 1.0001f, 1e-6f);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        double flops = 2.0 * 16 * (double) iters * blocks * 256;
        double tf = flops / (time_ms(e0, e1) * 1e-3) / 1e12;
        if (tf > best_fma) best_fma = tf;
    }

    /* (1b) packed FP32 FMA: same 16 chains per thread as 8 f32x2 chains */
    k_fma2<<<blocks, 256>>>(out, 1024, 1.0001f, 1e-6f);
    double best_fma2 = 0;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        k_fma2<<<blocks, 256>>>(out, iters, 1.0001f, 1e-6f);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        double flops = 2.0 * 16 * (double) iters * blocks * 256;
        double tf = flops / (time_ms(e0, e1) * 1e-3) / 1e12;
        if (tf > best_fma2) best_fma2 = tf;
    }

    /* (2) L2 and (3) HBM reads */
    auto read_bw = [&](size_t bytes, int sweeps) {
        uint4 *buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        const size_t nvec = bytes / 16;
        const int g = sms * 8;
        k_read<<<g, 256>>>(buf, nvec, 1, sink);
        double best = 0;
        for (int r = 0; r < 5; r++) {
            CK(cudaEventRecord(e0));
            k_read<<<g, 256>>>(buf, nvec, sweeps, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            const size_t per_sweep = (nvec / ((size_t) g * 256)) * ((size_t) g * 256) * 16;
            double gbs = (double) per_sweep * sweeps / (time_ms(e0, e1) * 1e-3) / 1e9;
            if (gbs > best) best = gbs;
        }
        CK(cudaFree(buf));
        return best;
    };
    const double l2_16 = read_bw((size_t) 16 << 20, 400), l2_32 = read_bw((size_t) 32 << 20, 200), l2_64 = read_bw((size_t) 64 << 20, 100);
    const double hbm = read_bw((size_t) 8 << 30, 2);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %d, \"fp32_fma_tflops\": %.2f, \"fp32_fma_f32x2_tflops\": %.2f, "
           "\"l2_read_gbs\": {\"16MiB\": %.1f, \"32MiB\": %.1f, \"64MiB\": %.1f}, \"hbm_read_gbs_8GiB\": %.1f, "
           "\"note\": \"best of 5 launches each, CUDA events; FMA: 16 independent chains/thread, 8 CTAs x 256 threads per SM; reads: ld.global.cg 128-bit, every CTA sweeps the whole window\"}\n",
           prop.name, sms, prop.clockRate / 1000, best_fma, best_fma2, l2_16, l2_32, l2_64, hbm);
    return 0;
}
