/*
 * mer_medium.cu — Medium / PhaseFunction side of the path: the eikonal ray stepper
 * (trace, traceTillBoundary), sampleDistance, evalTransmittance and Henyey-Greenstein,
 * as batch kernels behind the C ABI.
 *
 * Reference (paths relative to the MitsubaER tree):
 *   HeterogeneousRefractiveMedium   src/medium/heterogeneousrefractive.cpp:201-297 (ctor), 393-400, 402-568,
 *                                   653-691, 707-776
 *   Medium base                     src/librender/medium.cpp:27-37
 *   HGPhaseFunction                 src/phase/hg.cpp:76-110
 */
#include <algorithm>
#include <cmath>
#include <functional>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "mer_internal.h"

/* ------------------------------------------------------------------ a8: trace() */
/* SDFSHAPE: the container is the signed-distance grid (MER_SHAPE_SDF); the test is a full lookup per step here (the batch
 * steppers are the parity surface, the renderer has the lazy version) */
template <int MODE, bool SDFSHAPE = false>
__device__ __forceinline__ bool trace_dev(const MediumDev &M, StencilCache<MODE> &S, float3 &p, float3 &v, float &n, float3 &G, float dist,
                                          float &distSurf, float &opl, int &count) {
    int steps;
    float rem;
    trace_split(dist, M.h, steps, rem);
    distSurf = 0.0f;
    const float h = M.h;
    for (int i = 0; i < steps; i++) {
        er_step_fused<MODE>(M.rif, S, p, v, n, G, h, opl);
        count++;
        if (!(SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) {
            er_step_fused<MODE>(M.rif, S, p, v, n, G, -h, opl); /* step back, :679 */
            count++;
            return false;
        }
        distSurf += h;
    }
    er_step_fused<MODE>(M.rif, S, p, v, n, G, rem, opl);
    count++;
    if (!(SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) {
        er_step_fused<MODE>(M.rif, S, p, v, n, G, -rem, opl);
        count++;
        return false;
    }
    distSurf += rem;
    return true;
}

/* ------------------------------------------------------------------ a9: traceTillBoundary() */
template <int MODE, bool SDFSHAPE = false>
__device__ __forceinline__ void trace_till_boundary_dev(const MediumDev &M, StencilCache<MODE> &S, float3 &p, float3 &v, float &n, float3 &G,
                                                        float &distSurf, float &opl, int &count) {
    distSurf = 0.0f;
    const float h = M.h;
    for (int i = 0; i < 100000; i++) { /* maxsteps 1e5, :746 */
        er_step_fused<MODE>(M.rif, S, p, v, n, G, h, opl);
        count++;
        if ((SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) {
            distSurf += h;
        } else {
            er_step_fused<MODE>(M.rif, S, p, v, n, G, -h, opl);
            count++;
            distSurf -= h; /* :761 */
            return;
        }
    }
}

/* ------------------------------------------------------------------ a10: aggressive_trace() (:697-704) */
template <int MODE>
__device__ __forceinline__ void aggressive_trace_dev(const MediumDev &M, StencilCache<MODE> &S, float3 &p, float3 &v, float &n,
                                                     float3 &G, float dist, float &opl, int &count) {
    int steps;
    float rem;
    trace_split(dist, M.h, steps, rem);
    for (int i = 0; i < steps; i++) {
        er_step_fused<MODE, false>(M.rif, S, p, v, n, G, M.h, opl);
        count++;
    }
    er_step_fused<MODE, false>(M.rif, S, p, v, n, G, rem, opl);
    count++;
}

/* Batch stepper behind mer_medium_trace_* (and the C4 step-size sweep).  Rays of a batch need very
 * different step counts (they leave the shape at random times), so a thread-per-ray loop leaves most
 * lanes idle waiting for the warp's longest ray (ncu r01e: 13.6 of 32 lanes active).  Instead every lane
 * is a small state machine over the SAME flat loop: it takes the next ray of its stride as soon as its
 * current one ends, and the warp executes one convergent step body (one lookup call site) per iteration.
 * Per-ray results are identical to trace() / traceTillBoundary() (:671-691, :742-776). */
enum TraceKind : int { T_FULL = 0, T_REM = 1, T_BACKF = 2, T_BACKR = 3, T_ENTRY = 4, T_IDLE = 5 };

template <int MODE, bool TILL_BOUNDARY, bool SDFSHAPE = false>
__global__ void __launch_bounds__(128, 3)
k_trace(const __grid_constant__ MediumDev M, size_t nRays, float *__restrict__ P, float *__restrict__ V,
        const float *__restrict__ dist, uint8_t *__restrict__ success, float *__restrict__ distSurfOut,
        float *__restrict__ oplOut, int32_t *__restrict__ nstepsOut, unsigned long long *__restrict__ fetchesOut) {
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t next = (size_t) blockIdx.x * blockDim.x + threadIdx.x, cur = 0;
    StencilCache<MODE> S;
    S.invalidate();
    unsigned nFetch = 0; /* coefficient blocks gathered (cell changes incl. speculative requests): the algorithmic traffic */
    float3 p = f3(0.f, 0.f, 0.f), v = p, G = p;
    float n = 1.0f, ds = 0.0f, opl = 0.0f, rem = 0.0f;
    int kind = T_IDLE, stepsLeft = 0, count = 0;
    const float h = M.h;
    while (true) {
        asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); /* see k_render_pass */
        if (kind == T_IDLE && next < nRays) {
            cur = next;
            next += stride;
            p = f3(P[3 * cur], P[3 * cur + 1], P[3 * cur + 2]);
            v = f3(V[3 * cur], V[3 * cur + 1], V[3 * cur + 2]);
            n = 1.0f; G = f3(0.f, 0.f, 0.f);
            ds = 0.0f; opl = 0.0f; count = 0;
            if (TILL_BOUNDARY) { stepsLeft = 100000; rem = 0.0f; }
            else trace_split(dist[cur], h, stepsLeft, rem);
            kind = T_ENTRY; /* zero-length step: fetches the field at the start point */
        }
        if (__ballot_sync(0xffffffffu, kind != T_IDLE) == 0u) break;
        if (kind != T_IDLE) {
            const int k = kind;
            const float hc = k == T_FULL ? h : (k == T_REM ? rem : (k == T_BACKF ? -h : (k == T_BACKR ? -rem : 0.0f)));
            /* measured (C4 sweep): speculation pays in the packed mode (+16 %) but costs the compute-bound
             * tricubic stepper 3-14 % */
            const int si = S.i, sj = S.j, sk = S.k;
            er_step_fused<MODE, MODE == MER_RIF_TRILINEAR_PACKED>(M.rif, S, p, v, n, G, hc, opl);
            nFetch += (S.i != si || S.j != sj || S.k != sk) ? 1u : 0u;
            const bool inside = SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p);
            bool done = false, ok = false;
            if (k == T_ENTRY) {
                opl = 0.0f;
                kind = (TILL_BOUNDARY || stepsLeft > 0) ? T_FULL : T_REM;
            } else {
                count++;
                if (k == T_FULL) {
                    if (inside) {
                        ds += h;
                        if (--stepsLeft == 0) {
                            if (TILL_BOUNDARY) done = true; /* 1e5 steps exhausted, :746 */
                            else kind = T_REM;
                        }
                    } else {
                        kind = T_BACKF;
                    }
                } else if (k == T_REM) {
                    if (inside) { ds += rem; done = true; ok = true; }
                    else kind = T_BACKR;
                } else {
                    if (TILL_BOUNDARY && k == T_BACKF) ds -= h; /* :761 */
                    done = true;
                }
            }
            if (done) {
                P[3 * cur] = p.x; P[3 * cur + 1] = p.y; P[3 * cur + 2] = p.z;
                V[3 * cur] = v.x; V[3 * cur + 1] = v.y; V[3 * cur + 2] = v.z;
                if (success) success[cur] = ok ? 1 : 0;
                if (distSurfOut) distSurfOut[cur] = ds;
                if (oplOut) oplOut[cur] = opl;
                if (nstepsOut) nstepsOut[cur] = count;
                kind = T_IDLE;
            }
        }
    }
    if (fetchesOut) {
        nFetch = __reduce_add_sync(0xffffffffu, nFetch);
        if ((threadIdx.x & 31u) == 0u && nFetch) atomicAdd(fetchesOut, (unsigned long long) nFetch);
    }
}

/* ------------------------------------------------------------------ a12: sampleDistance() */
struct SampleDistanceOut {
    uint8_t *success;
    float *t, *p, *d, *opl, *refRatioSq, *transmittance, *pdfSuccess, *pdfFailure, *sigmaS;
    int32_t *nsteps;
};

template <int MODE, bool SDFSHAPE = false>
__global__ void __launch_bounds__(128)
k_sample_distance(const __grid_constant__ MediumDev M, size_t nRays, const float *__restrict__ RO,
                  const float *__restrict__ RD, const float *__restrict__ mintIn, const float *__restrict__ xi,
                  SampleDistanceOut out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nRays; i += (size_t) gridDim.x * blockDim.x) {
        const float mint = mintIn ? mintIn[i] : 0.0f;
        float rnd = xi[2 * i], sampledDistance, sd = M.samplingDensity, pdfSampled = 0.0f;
        if (rnd < M.weight) {
            rnd = __fdiv_rn(rnd, M.weight);
            if (M.strategy == MER_STRATEGY_MAXIMUM) { /* :445 */
                sampledDistance = maxexp_sample(M, 1.0f - rnd, pdfSampled);
            } else {
                if (M.strategy == MER_STRATEGY_BALANCE) sd = M.sigmaT[min((int) (xi[2 * i + 1] * 3.0f), 2)];
                sampledDistance = __fdiv_rn(-fastlog_dev(1.0f - rnd), sd);
            }
        } else {
            sampledDistance = INFINITY;
        }
        float3 p = f3(RO[3 * i], RO[3 * i + 1], RO[3 * i + 2]), v = f3(RD[3 * i], RD[3 * i + 1], RD[3 * i + 2]);
        const float3 o = p;
        bool success = true;
        float distSurf = 0.0f, opl = 0.0f, refRatioSq = 0.0f, T[3] = {0.f, 0.f, 0.f}, ps = 1.0f, pf = 1.0f, t = 0.0f;
        int count = 0;
        if (!rif_inside_limits(M.hasSdf ? M.sdf : M.rif, p)) { /* :461-466 */
            success = false;
        } else {
            float n;
            float3 G;
            StencilCache<MODE> S;
            S.invalidate();
            rif_lookup_cached<MODE>(M.rif, p, S, n, G);
            const float refStart = n;
            refRatioSq = (float) (1.0 / (double) (refStart * refStart));
            v = f3(v.x * refStart, v.y * refStart, v.z * refStart);
            if (isfinite(sampledDistance)) {
                if (!M.aggressive) {
                    success = trace_dev<MODE, SDFSHAPE>(M, S, p, v, n, G, sampledDistance, distSurf, opl, count);
                } else { /* :476-493 */
                    float distLeft = sampledDistance, distTraced = 0.0f;
                    while (distLeft > MER_EPSILON) {
                        float sdfv;
                        float3 gUnused;
                        rif_tricubic(M.sdf, rif_to_volume(M.sdf, p), sdfv, gUnused); /* m_SDF->value(p) */
                        sdfv = __fsub_rn(-sdfv, M.maxSdfError);
                        if (sdfv < MER_EPSILON) break;
                        const float traceDist = fminf(sdfv, distLeft);
                        aggressive_trace_dev<MODE>(M, S, p, v, n, G, traceDist, opl, count);
                        distLeft = __fsub_rn(distLeft, traceDist);
                        distTraced = __fadd_rn(distTraced, traceDist);
                    }
                    success = trace_dev<MODE, SDFSHAPE>(M, S, p, v, n, G, distLeft, distSurf, opl, count);
                    distSurf = __fadd_rn(distSurf, distTraced);
                }
            } else {
                trace_till_boundary_dev<MODE, SDFSHAPE>(M, S, p, v, n, G, distSurf, opl, count);
                success = false;
            }
            refRatioSq *= n * n; /* refEnd = value at the final p, carried by the fused stepper */
            if (success) {
                t = sampledDistance + mint;
                if (p.x == o.x && p.y == o.y && p.z == o.z) success = false; /* no forward progress, :517-520 */
            } else {
                sampledDistance = distSurf;
                t = sampledDistance + mint;
            }
            float pdfFailure = 0.0f, pdfSuccess = 0.0f;
            if (M.strategy == MER_STRATEGY_MAXIMUM) { /* :534-536: pdfSuccess is what MaxExpDist::sample returned */
                pdfFailure = __fsub_rn(1.0f, maxexp_cdf(M, sampledDistance));
                pdfSuccess = pdfSampled;
            } else if (M.strategy == MER_STRATEGY_BALANCE) {
                for (int c = 0; c < 3; c++) {
                    float tmp = fastexp_dev(__fmul_rn(-M.sigmaT[c], sampledDistance));
                    pdfFailure = __fadd_rn(pdfFailure, tmp);
                    pdfSuccess = __fadd_rn(pdfSuccess, __fmul_rn(M.sigmaT[c], tmp));
                }
                pdfFailure = __fdiv_rn(pdfFailure, 3.0f);
                pdfSuccess = __fdiv_rn(pdfSuccess, 3.0f);
            } else {
                pdfFailure = fastexp_dev(__fmul_rn(-sd, sampledDistance));
                pdfSuccess = __fmul_rn(sd, pdfFailure);
            }
            float tmax = 0.0f;
            for (int c = 0; c < 3; c++) {
                T[c] = fastexp_dev(__fmul_rn(M.sigmaT[c], -sampledDistance));
                tmax = fmaxf(tmax, T[c]);
            }
            if (tmax < 1e-20f) T[0] = T[1] = T[2] = 0.0f;
            ps = __fmul_rn(pdfSuccess, M.weight);
            pf = __fadd_rn(__fmul_rn(M.weight, pdfFailure), 1.0f - M.weight);
        }
        if (out.success) out.success[i] = success ? 1 : 0;
        if (out.t) out.t[i] = t;
        if (out.p) { out.p[3 * i] = p.x; out.p[3 * i + 1] = p.y; out.p[3 * i + 2] = p.z; }
        if (out.d) { out.d[3 * i] = v.x; out.d[3 * i + 1] = v.y; out.d[3 * i + 2] = v.z; }
        if (out.opl) out.opl[i] = opl;
        if (out.refRatioSq) out.refRatioSq[i] = refRatioSq;
        if (out.transmittance) { out.transmittance[3 * i] = T[0]; out.transmittance[3 * i + 1] = T[1]; out.transmittance[3 * i + 2] = T[2]; }
        if (out.pdfSuccess) out.pdfSuccess[i] = ps;
        if (out.pdfFailure) out.pdfFailure[i] = pf;
        if (out.sigmaS) { out.sigmaS[3 * i] = M.sigmaS[0]; out.sigmaS[3 * i + 1] = M.sigmaS[1]; out.sigmaS[3 * i + 2] = M.sigmaS[2]; }
        if (out.nsteps) out.nsteps[i] = count;
    }
}

/* ------------------------------------------------------------------ a14-a16 */
__global__ void k_eval_transmittance(float s0, float s1, float s2, size_t n, const float *__restrict__ mint,
                                     const float *__restrict__ maxt, float *__restrict__ out) {
    const float sig[3] = {s0, s1, s2};
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float negLength = mint[i] - maxt[i];
        for (int c = 0; c < 3; c++) out[3 * i + c] = sig[c] != 0.0f ? fastexp_dev(__fmul_rn(sig[c], negLength)) : 1.0f;
    }
}

__global__ void k_hg_sample(float g, size_t n, const float *__restrict__ wi, const float *__restrict__ xi,
                            float *__restrict__ wo, float *__restrict__ pdf) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float3 w = f3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        float3 o = hg_sample_dev(g, w, xi[2 * i], xi[2 * i + 1]);
        wo[3 * i] = o.x; wo[3 * i + 1] = o.y; wo[3 * i + 2] = o.z;
        if (pdf) pdf[i] = hg_eval_dev(g, w, o);
    }
}

__global__ void k_hg_eval(float g, size_t n, const float *__restrict__ wi, const float *__restrict__ wo,
                          float *__restrict__ out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = hg_eval_dev(g, f3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), f3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]));
}

/* ===================================================================== host side */
namespace {

using mer::DevBuf; /* RAII device buffer for the *_batch (host pointer) entry points */

/* persistent-style sizing: every CTA the kernel can keep resident (occupancy query: 3 per SM for the
 * 168-register tricubic stepper, 6 for the packed one), times two so that the tail of one wave overlaps
 * the start of the next */
template <typename K> unsigned trace_grid(K kernel, size_t n) {
    int perSm = 3, sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, 128, 0) != cudaSuccess || perSm < 1) perSm = 3;
    size_t cap = (size_t) sms * (size_t) perSm * 2u;
    if (const char *e = getenv("MER_TRACE_CTAS_PER_SM")) cap = (size_t) sms * (size_t) atoi(e); /* tuning knob */
    size_t blocks = (n + 127) / 128;
    return (unsigned) (blocks < cap ? blocks : cap);
}

} /* namespace */

#define UP(buf, host, bytes)                                                      \
    do {                                                                          \
        MER_CUDA(buf.alloc(bytes));                                               \
        if (host) MER_CUDA(cudaMemcpy(buf.ptr, host, bytes, cudaMemcpyHostToDevice)); \
    } while (0)
#define DOWN(host, buf, bytes)                                                    \
    do {                                                                          \
        if (host) MER_CUDA(cudaMemcpy(host, buf.ptr, bytes, cudaMemcpyDeviceToHost)); \
    } while (0)

extern "C" {

int mer_medium_create(const mer_medium_desc *desc, const mer_rif *rif, const mer_grid *density, mer_medium **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    MER_REQUIRE(desc, "null medium descriptor");
    MER_REQUIRE(rif, "No RIF specified!"); /* heterogeneousrefractive.cpp:368-369 */
    MER_REQUIRE(!density || density->device == rif->device, "rif and density live on different devices");
    MER_REQUIRE(desc->stepsize > 0.0f, "stepsize must be positive");
    MER_REQUIRE(desc->hg_g > -1.0f && desc->hg_g < 1.0f,
                "The asymmetry parameter must lie in the interval (-1, 1)!"); /* hg.cpp:50-52 */
    MER_REQUIRE(desc->shape_type == MER_SHAPE_BOX || desc->shape_type == MER_SHAPE_SPHERE || desc->shape_type == MER_SHAPE_SDF, "unknown shape type");
    MER_REQUIRE(desc->boundary == MER_BOUNDARY_INDEX_MATCHED || desc->boundary == MER_BOUNDARY_HDIELECTRIC, "unknown boundary type");
    MER_REQUIRE(desc->radiance_scaling == MER_SCALING_REFERENCE || desc->radiance_scaling == MER_SCALING_PHYSICAL, "unknown radiance scaling");
    MER_REQUIRE(desc->strategy >= MER_STRATEGY_BALANCE && desc->strategy <= MER_STRATEGY_MAXIMUM,
                "Specified an unknown sampling strategy"); /* :296 */
    if (density) MER_REQUIRE(desc->density_scale > 0.0f, "density_scale must be positive when a density grid is attached");

    if (density) MER_REQUIRE(density->dev.channels == 1, "the density volume must have one channel (heterogeneous.cpp:269-271)");
    mer_medium *m = new mer_medium();
    memset(m, 0, sizeof(*m));
    m->device = rif->device;
    m->rif = rif;
    m->grid = density;
    m->desc = *desc;
    MediumDev &D = m->dev;
    D.rif = rif->dev;
    D.hasGrid = density ? 1 : 0;
    if (density) D.grid = density->dev;
    for (int i = 0; i < 3; i++) {
        D.sigmaA[i] = desc->sigma_a[i];
        D.sigmaS[i] = desc->sigma_s[i];
        D.sigmaT[i] = desc->sigma_a[i] + desc->sigma_s[i]; /* medium.cpp:36 */
        D.albedo[i] = desc->albedo[i];
    }
    D.h = desc->stepsize;
    /* mediumSamplingWeight default, :239-255 */
    float w = desc->medium_sampling_weight;
    if (w == -1.0f) {
        for (int i = 0; i < 3; i++) {
            float albedo = D.sigmaS[i] / D.sigmaT[i];
            if (albedo > w && D.sigmaT[i] != 0.0f) w = albedo;
        }
        if (w > 0.0f) w = std::max(w, 0.5f);
    }
    D.weight = w;
    D.strategy = desc->strategy;
    D.samplingDensity = 0.0f;
    if (desc->strategy == MER_STRATEGY_SINGLE) { /* :259-283 */
        int channel = 0;
        float smallest = std::numeric_limits<float>::infinity();
        for (int i = 0; i < 3; i++)
            if (D.sigmaT[i] < smallest) { smallest = D.sigmaT[i]; channel = i; }
        if (desc->channel >= 0) {
            if (desc->channel > 2) { delete m; return mer::fail(MER_ERR_INVALID, "channel out of range"); }
            channel = desc->channel;
        }
        D.samplingDensity = D.sigmaT[channel];
        m->desc.channel = channel;
    } else if (desc->strategy == MER_STRATEGY_MANUAL) {
        D.samplingDensity = desc->sampling_density;
    } else if (desc->strategy == MER_STRATEGY_MAXIMUM) { /* :287-291 + MaxExpDist::MaxExpDist, src/medium/maxexp.h:29-58, in Float */
        float sg[3] = {D.sigmaT[0], D.sigmaT[1], D.sigmaT[2]};
        std::sort(sg, sg + 3, std::greater<float>());
        float cdf[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < 3; i++) {
            if (i > 0 && sg[i] == sg[i - 1]) { delete m; return mer::fail(MER_ERR_INVALID, "Internal error: sigmaT must vary across channels"); }
            const float lower = i == 0 ? -1.0f : -std::pow(sg[i] / sg[i - 1], -sg[i] / (sg[i] - sg[i - 1]));
            const float upper = i == 2 ? 0.0f : -std::pow(sg[i + 1] / sg[i], -sg[i] / (sg[i + 1] - sg[i]));
            cdf[i + 1] = cdf[i] + (upper - lower);
            D.mxSigma[i] = sg[i];
            D.mxLower[i] = lower;
            D.mxStart[i] = i == 0 ? 0.0f : (float) std::log((double) (sg[i] / sg[i - 1])) / (sg[i] - sg[i - 1]);
        }
        D.mxNorm = cdf[3];
        D.mxInvNorm = 1.0f / D.mxNorm;
        for (int i = 0; i < 4; i++) D.mxCdf[i] = cdf[i] * D.mxInvNorm;
    }
    D.shapeType = desc->shape_type;
    for (int i = 0; i < 6; i++) D.shape[i] = desc->shape[i];
    D.g = desc->hg_g;
    D.boundary = desc->boundary;
    D.physicalScaling = desc->radiance_scaling == MER_SCALING_PHYSICAL ? 1 : 0;
    D.minExit2 = MER_EPSILON;
    D.densityScale = desc->density_scale;
    D.invMaxDensity = density ? 1.0f / (desc->density_scale * 1.0f) : 0.0f; /* heterogeneous.cpp:239-242 */
    m->desc.medium_sampling_weight = w;
    m->desc.sampling_density = D.samplingDensity;
    *out = m;
    return MER_OK;
}

void mer_medium_destroy(mer_medium *m) {
    if (!m) return;
    delete m;
}

int mer_medium_set_sdf(mer_medium *m, const mer_rif *sdf, int aggressive) {
    MER_REQUIRE(m, "null handle");
    if (!sdf) {
        MER_REQUIRE(!aggressive, "No SDF specified!"); /* heterogeneousrefractive.cpp:371-372 */
        m->dev.hasSdf = m->dev.aggressive = 0;
        return MER_OK;
    }
    MER_REQUIRE(sdf->device == m->device, "rif and sdf live on different devices");
    MER_REQUIRE(sdf->mode == MER_RIF_TRICUBIC, "the sdf volume must be a tricubic splinevolume");
    for (int i = 0; i < 3; i++) /* :374-377 */
        MER_REQUIRE(sdf->desc.bbox_min[i] == m->rif->desc.bbox_min[i] && sdf->desc.bbox_max[i] == m->rif->desc.bbox_max[i],
                    "The bounding boxes of rif and winding number do not match");
    m->dev.sdf = sdf->dev;
    m->dev.hasSdf = 1;
    m->dev.aggressive = aggressive ? 1 : 0;
    /* maxSDFError, splinevolume.cpp:282: sqrt(sum stride^2), stride = FLOAT(1/xres) */
    float s0 = (float) (1.0 / (double) sdf->dev.xres[0]), s1 = (float) (1.0 / (double) sdf->dev.xres[1]), s2 = (float) (1.0 / (double) sdf->dev.xres[2]);
    m->dev.maxSdfError = std::sqrt(s0 * s0 + s1 * s1 + s2 * s2);
    return MER_OK;
}

int mer_medium_set_albedo_grid(mer_medium *m, const mer_grid *albedo) {
    MER_REQUIRE(m, "null handle");
    if (!albedo) { m->dev.hasAlbedoGrid = 0; return MER_OK; }
    MER_REQUIRE(m->dev.hasGrid, "an albedo volume goes with a density volume (heterogeneous.cpp:229-236)");
    MER_REQUIRE(albedo->device == m->device, "medium and albedo volume live on different devices");
    MER_REQUIRE(albedo->dev.channels == 3, "the albedo volume must support spectrum lookups (heterogeneous.cpp:266-268)");
    m->dev.albedoGrid = albedo->dev;
    m->dev.hasAlbedoGrid = 1;
    return MER_OK;
}

int mer_medium_resolved(const mer_medium *m, mer_medium_desc *out, float *sampling_density_out) {
    MER_REQUIRE(m, "null handle");
    if (out) *out = m->desc;
    if (sampling_density_out) *sampling_density_out = m->dev.samplingDensity;
    return MER_OK;
}

int mer_medium_trace_device(const mer_medium *m, size_t n, float *p_dev, float *v_dev, const float *dist_dev,
                            uint8_t *success_dev, float *dist_surf_dev, float *opl_dev, int32_t *nsteps_dev,
                            void *stream) {
    return mer_medium_trace_counted_device(m, n, p_dev, v_dev, dist_dev, success_dev, dist_surf_dev, opl_dev, nsteps_dev, nullptr, stream);
}

int mer_medium_trace_counted_device(const mer_medium *m, size_t n, float *p_dev, float *v_dev, const float *dist_dev,
                                    uint8_t *success_dev, float *dist_surf_dev, float *opl_dev, int32_t *nsteps_dev,
                                    uint64_t *block_fetches_dev, void *stream) {
    MER_REQUIRE(m && (n == 0 || (p_dev && v_dev && dist_dev)), "null argument");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    if (m->dev.shapeType == MER_SHAPE_SDF) {
        MER_REQUIRE(m->dev.hasSdf, "shape type SDF needs the sdf volume (mer_medium_set_sdf)");
        if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "shape type SDF is built for the tricubic RIF mode");
        MER_LAUNCH((k_trace<MER_RIF_TRICUBIC, false, true>), trace_grid(k_trace<MER_RIF_TRICUBIC, false, true>, n), 128, 0, (cudaStream_t) stream, m->dev, n,
                   p_dev, v_dev, dist_dev, success_dev, dist_surf_dev, opl_dev, nsteps_dev, (unsigned long long *) block_fetches_dev);
    } else if (m->rif->mode == MER_RIF_TRICUBIC)
        MER_LAUNCH((k_trace<MER_RIF_TRICUBIC, false>), trace_grid(k_trace<MER_RIF_TRICUBIC, false>, n), 128, 0, (cudaStream_t) stream, m->dev, n, p_dev, v_dev,
                   dist_dev, success_dev, dist_surf_dev, opl_dev, nsteps_dev, (unsigned long long *) block_fetches_dev);
    else
        MER_LAUNCH((k_trace<MER_RIF_TRILINEAR_PACKED, false>), trace_grid(k_trace<MER_RIF_TRILINEAR_PACKED, false>, n), 128, 0, (cudaStream_t) stream, m->dev, n, p_dev,
                   v_dev, dist_dev, success_dev, dist_surf_dev, opl_dev, nsteps_dev, (unsigned long long *) block_fetches_dev);
    return MER_OK;
}

int mer_medium_trace_batch(const mer_medium *m, size_t n, float *p, float *v, const float *dist, uint8_t *success_out,
                           float *dist_surf_out, float *opl_out, int32_t *nsteps_out) {
    MER_REQUIRE(m && (n == 0 || (p && v && dist)), "null argument");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf dp, dv, dd, dok, dds, dopl, dns;
    UP(dp, p, n * 12); UP(dv, v, n * 12); UP(dd, dist, n * 4);
    UP(dok, (void *) nullptr, n); UP(dds, (void *) nullptr, n * 4); UP(dopl, (void *) nullptr, n * 4); UP(dns, (void *) nullptr, n * 4);
    int rc = mer_medium_trace_device(m, n, dp.as<float>(), dv.as<float>(), dd.as<float>(), dok.as<uint8_t>(),
                                     dds.as<float>(), dopl.as<float>(), dns.as<int32_t>(), nullptr);
    if (rc) return rc;
    MER_CUDA(cudaDeviceSynchronize());
    DOWN(p, dp, n * 12); DOWN(v, dv, n * 12); DOWN(success_out, dok, n); DOWN(dist_surf_out, dds, n * 4);
    DOWN(opl_out, dopl, n * 4); DOWN(nsteps_out, dns, n * 4);
    return MER_OK;
}

int mer_medium_trace_till_boundary_batch(const mer_medium *m, size_t n, float *p, float *v, float *dist_surf_out,
                                         float *opl_out, int32_t *nsteps_out) {
    MER_REQUIRE(m && (n == 0 || (p && v)), "null argument");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf dp, dv, dds, dopl, dns;
    UP(dp, p, n * 12); UP(dv, v, n * 12);
    UP(dds, (void *) nullptr, n * 4); UP(dopl, (void *) nullptr, n * 4); UP(dns, (void *) nullptr, n * 4);
    if (m->dev.shapeType == MER_SHAPE_SDF) {
        MER_REQUIRE(m->dev.hasSdf, "shape type SDF needs the sdf volume (mer_medium_set_sdf)");
        if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "shape type SDF is built for the tricubic RIF mode");
        MER_LAUNCH((k_trace<MER_RIF_TRICUBIC, true, true>), trace_grid(k_trace<MER_RIF_TRICUBIC, true, true>, n), 128, 0, 0, m->dev, n, dp.as<float>(),
                   dv.as<float>(), (const float *) nullptr, (uint8_t *) nullptr, dds.as<float>(), dopl.as<float>(), dns.as<int32_t>(), (unsigned long long *) nullptr);
    } else if (m->rif->mode == MER_RIF_TRICUBIC)
        MER_LAUNCH((k_trace<MER_RIF_TRICUBIC, true>), trace_grid(k_trace<MER_RIF_TRICUBIC, true>, n), 128, 0, 0, m->dev, n, dp.as<float>(), dv.as<float>(),
                   (const float *) nullptr, (uint8_t *) nullptr, dds.as<float>(), dopl.as<float>(), dns.as<int32_t>(), (unsigned long long *) nullptr);
    else
        MER_LAUNCH((k_trace<MER_RIF_TRILINEAR_PACKED, true>), trace_grid(k_trace<MER_RIF_TRILINEAR_PACKED, true>, n), 128, 0, 0, m->dev, n, dp.as<float>(), dv.as<float>(),
                   (const float *) nullptr, (uint8_t *) nullptr, dds.as<float>(), dopl.as<float>(), dns.as<int32_t>(), (unsigned long long *) nullptr);
    MER_CUDA(cudaDeviceSynchronize());
    DOWN(p, dp, n * 12); DOWN(v, dv, n * 12); DOWN(dist_surf_out, dds, n * 4); DOWN(opl_out, dopl, n * 4);
    DOWN(nsteps_out, dns, n * 4);
    return MER_OK;
}

int mer_medium_sample_distance_batch(const mer_medium *m, size_t n, const float *ray_o, const float *ray_d,
                                     const float *ray_mint, const float *xi, mer_medium_sampling_records *rec) {
    MER_REQUIRE(m && rec && (n == 0 || (ray_o && ray_d && xi)), "null argument");
    if (m->grid)
        return mer::fail(MER_ERR_UNSUPPORTED,
                         "sampleDistance with a density grid is only available inside mer_render (Woodcock on the curve)");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf dro, drd, dmint, dxi, dok, dt, dp, dd, dopl, drr, dT, dps, dpf, dss, dns;
    UP(dro, ray_o, n * 12); UP(drd, ray_d, n * 12); UP(dxi, xi, n * 8);
    if (ray_mint) UP(dmint, ray_mint, n * 4);
    UP(dok, (void *) nullptr, n); UP(dt, (void *) nullptr, n * 4); UP(dp, (void *) nullptr, n * 12);
    UP(dd, (void *) nullptr, n * 12); UP(dopl, (void *) nullptr, n * 4); UP(drr, (void *) nullptr, n * 4);
    UP(dT, (void *) nullptr, n * 12); UP(dps, (void *) nullptr, n * 4); UP(dpf, (void *) nullptr, n * 4);
    UP(dss, (void *) nullptr, n * 12); UP(dns, (void *) nullptr, n * 4);
    SampleDistanceOut o;
    o.success = dok.as<uint8_t>(); o.t = dt.as<float>(); o.p = dp.as<float>(); o.d = dd.as<float>();
    o.opl = dopl.as<float>(); o.refRatioSq = drr.as<float>(); o.transmittance = dT.as<float>();
    o.pdfSuccess = dps.as<float>(); o.pdfFailure = dpf.as<float>(); o.sigmaS = dss.as<float>();
    o.nsteps = dns.as<int32_t>();
    const float *mintDev = ray_mint ? dmint.as<float>() : nullptr;
    if (m->dev.shapeType == MER_SHAPE_SDF) {
        MER_REQUIRE(m->dev.hasSdf, "shape type SDF needs the sdf volume (mer_medium_set_sdf)");
        if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "shape type SDF is built for the tricubic RIF mode");
        MER_LAUNCH((k_sample_distance<MER_RIF_TRICUBIC, true>), trace_grid(k_sample_distance<MER_RIF_TRICUBIC, true>, n), 128, 0, 0, m->dev, n, dro.as<float>(),
                   drd.as<float>(), mintDev, dxi.as<float>(), o);
    } else if (m->rif->mode == MER_RIF_TRICUBIC)
        MER_LAUNCH(k_sample_distance<MER_RIF_TRICUBIC>, trace_grid(k_sample_distance<MER_RIF_TRICUBIC>, n), 128, 0, 0, m->dev, n, dro.as<float>(),
                   drd.as<float>(), mintDev, dxi.as<float>(), o);
    else
        MER_LAUNCH(k_sample_distance<MER_RIF_TRILINEAR_PACKED>, trace_grid(k_sample_distance<MER_RIF_TRILINEAR_PACKED>, n), 128, 0, 0, m->dev, n, dro.as<float>(),
                   drd.as<float>(), mintDev, dxi.as<float>(), o);
    MER_CUDA(cudaDeviceSynchronize());
    DOWN(rec->success, dok, n); DOWN(rec->t, dt, n * 4); DOWN(rec->p, dp, n * 12); DOWN(rec->d, dd, n * 12);
    DOWN(rec->optical_length, dopl, n * 4); DOWN(rec->ref_ratio_sq, drr, n * 4); DOWN(rec->transmittance, dT, n * 12);
    DOWN(rec->pdf_success, dps, n * 4); DOWN(rec->pdf_failure, dpf, n * 4); DOWN(rec->sigma_s, dss, n * 12);
    DOWN(rec->nsteps, dns, n * 4);
    return MER_OK;
}

int mer_medium_eval_transmittance_batch(const mer_medium *m, size_t n, const float *mint, const float *maxt,
                                        float *transmittance_out) {
    MER_REQUIRE(m && (n == 0 || (mint && maxt && transmittance_out)), "null argument");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf da, db, dout;
    UP(da, mint, n * 4); UP(db, maxt, n * 4); UP(dout, (void *) nullptr, n * 12);
    MER_LAUNCH(k_eval_transmittance, (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u), 256, 0, 0,
               m->dev.sigmaT[0], m->dev.sigmaT[1], m->dev.sigmaT[2], n, da.as<float>(), db.as<float>(), dout.as<float>());
    DOWN(transmittance_out, dout, n * 12);
    return MER_OK;
}

int mer_hg_sample_batch(int device, float g, size_t n, const float *wi, const float *xi, float *wo_out, float *pdf_out) {
    MER_REQUIRE(n == 0 || (wi && xi && wo_out), "null argument");
    MER_REQUIRE(g > -1.0f && g < 1.0f, "The asymmetry parameter must lie in the interval (-1, 1)!");
    int rc = mer::check_device(device);
    if (rc) return rc;
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(device);
    DevBuf dwi, dxi, dwo, dpdf;
    UP(dwi, wi, n * 12); UP(dxi, xi, n * 8); UP(dwo, (void *) nullptr, n * 12); UP(dpdf, (void *) nullptr, n * 4);
    MER_LAUNCH(k_hg_sample, (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u), 256, 0, 0, g, n, dwi.as<float>(),
               dxi.as<float>(), dwo.as<float>(), dpdf.as<float>());
    DOWN(wo_out, dwo, n * 12); DOWN(pdf_out, dpdf, n * 4);
    return MER_OK;
}

int mer_hg_eval_batch(int device, float g, size_t n, const float *wi, const float *wo, float *value_out) {
    MER_REQUIRE(n == 0 || (wi && wo && value_out), "null argument");
    MER_REQUIRE(g > -1.0f && g < 1.0f, "The asymmetry parameter must lie in the interval (-1, 1)!");
    int rc = mer::check_device(device);
    if (rc) return rc;
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(device);
    DevBuf dwi, dwo, dout;
    UP(dwi, wi, n * 12); UP(dwo, wo, n * 12); UP(dout, (void *) nullptr, n * 4);
    MER_LAUNCH(k_hg_eval, (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u), 256, 0, 0, g, n, dwi.as<float>(),
               dwo.as<float>(), dout.as<float>());
    DOWN(value_out, dout, n * 4);
    return MER_OK;
}

} /* extern "C" */
