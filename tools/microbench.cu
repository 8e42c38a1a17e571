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

/* (4) incoherent 4x4x4-stencil-style gathers, every lane at its own random cell: 16 tld4 from a 2-D array (the atlas layout
 * of the stepper) and 8 ld.global.nc.v8.f32 sectors from a linear table (the coeff8 layout).  Window 64 MiB (L2-resident). */
__global__ void __launch_bounds__(128) k_tld4(cudaTextureObject_t tex, int dim, int iters, float *sink) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        const float u = (float) ((s >> 8) % (unsigned) (dim - 8) + 2), v = (float) ((s >> 3) % (unsigned) (dim - 40) + 2);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float vk = v + 8.0f * (float) k;
            const float4 a = tex2Dgather<float4>(tex, u, vk, 0), b = tex2Dgather<float4>(tex, u + 2.f, vk, 0);
            const float4 c = tex2Dgather<float4>(tex, u, vk + 2.f, 0), d = tex2Dgather<float4>(tex, u + 2.f, vk + 2.f, 0);
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}
__global__ void __launch_bounds__(128) k_ldg256(const float4 *__restrict__ tab, size_t nsec, int iters, float *sink) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        s = s * 1664525u + 1013904223u;
        const size_t base = (size_t) (s % (unsigned) (nsec - 8 * 4096));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float4 a, b;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                         : "l"(tab + 2 * (base + (size_t) k * 4096)));
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
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
    /* (4) gathers: GB/s of coefficients returned to registers */
    double tld4_gbs = 0, ldg256_gbs = 0;
    {
        const int dim = 4096, it4 = 2048, g = sms * 16;
        cudaArray_t arr;
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
        CK(cudaMallocArray(&arr, &cd, dim, dim, cudaArrayTextureGather));
        float *tmp;
        CK(cudaMalloc(&tmp, (size_t) dim * dim * 4));
        CK(cudaMemset(tmp, 0, (size_t) dim * dim * 4));
        CK(cudaMemcpy2DToArray(arr, 0, 0, tmp, (size_t) dim * 4, (size_t) dim * 4, dim, cudaMemcpyDeviceToDevice));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex;
        CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
        k_tld4<<<g, 128>>>(tex, dim, 64, out);
        for (int r = 0; r < 5; r++) {
            CK(cudaEventRecord(e0));
            k_tld4<<<g, 128>>>(tex, dim, it4, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            const double gbs = 256.0 * it4 * (double) g * 128 / (time_ms(e0, e1) * 1e-3) / 1e9;
            if (gbs > tld4_gbs) tld4_gbs = gbs;
        }
        const size_t nsec = ((size_t) 64 << 20) / 32;
        k_ldg256<<<g, 128>>>((const float4 *) tmp, nsec, 64, out);
        for (int r = 0; r < 5; r++) {
            CK(cudaEventRecord(e0));
            k_ldg256<<<g, 128>>>((const float4 *) tmp, nsec, it4, out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            const double gbs = 256.0 * it4 * (double) g * 128 / (time_ms(e0, e1) * 1e-3) / 1e9;
            if (gbs > ldg256_gbs) ldg256_gbs = gbs;
        }
        CK(cudaDestroyTextureObject(tex)); CK(cudaFreeArray(arr)); CK(cudaFree(tmp));
    }
    printf("{\"gather_tld4_gbs\": %.1f, \"gather_ldg256_gbs\": %.1f, ", tld4_gbs, ldg256_gbs);
    printf("\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_max\": %d, \"fp32_fma_tflops\": %.2f, \"fp32_fma_f32x2_tflops\": %.2f, "
           "\"l2_read_gbs\": {\"16MiB\": %.1f, \"32MiB\": %.1f, \"64MiB\": %.1f}, \"hbm_read_gbs_8GiB\": %.1f, "
           "\"note\": \"best of 5 launches each, CUDA events; FMA: 16 independent chains/thread, 8 CTAs x 256 threads per SM; reads: ld.global.cg 128-bit, every CTA sweeps the whole window\"}\n",
           prop.name, sms, prop.clockRate / 1000, best_fma, best_fma2, l2_16, l2_32, l2_64, hbm);
    return 0;
}
