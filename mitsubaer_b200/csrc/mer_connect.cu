/*
 * mer_connect.cu — curved direct connections, function level.
 *
 * What the reference does with Ceres (a shooting problem per connection) rests on three deterministic
 * pieces, which are built here and parity-tested in tests/ against a CPU restatement:
 *   Spline<3>::valueGradientAndHessian    include/mitsuba/core/basisspline.h:539-606 (+ splinevolume.cpp:371-377)
 *   er_derivativestep                      src/medium/heterogeneousrefractive.cpp:798-814
 *   computefdfBDPT                         :816-939 (closest approach by sign change + step halvings, boundary
 *                                          exit with Snell refraction: boundaryVelocityDerivative :1057-1074)
 * and on the solver (makeDirectConnections :1087-1163), for which the reference calls Ceres 1.14 BFGS line
 * search.  Ceres is not reproducible bit-wise (parity UNPINNED there); the solver here is a
 * Levenberg-Marquardt iteration on the same residual/Jacobian, validated by the residual it reaches.
 *
 * One thread per connection.  3x3 matrices are row-major float[9]; outer(a,b)_ij = a_i b_j
 * (include/mitsuba/core/matrix.h:584-588); premult(M, x) = M^T x (:765-769).
 */
#include <cmath>
#include <cstring>

#include "mer_internal.h"

#include "mer_connect.cuh"

using namespace merc;

namespace {

__global__ void k_rif_hessian(RifDev R, size_t n, const float *__restrict__ p, float *__restrict__ f, float *__restrict__ g,
                              float *__restrict__ H) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        Field F;
        rif_field(R, f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), F);
        if (f) f[i] = F.n;
        if (g) { g[3 * i] = F.G.x; g[3 * i + 1] = F.G.y; g[3 * i + 2] = F.G.z; }
        if (H) { M3 m = hess_full(F); for (int k = 0; k < 9; k++) H[9 * i + k] = m.m[k]; }
    }
}

__global__ void __launch_bounds__(128)
k_derivative_trace(const __grid_constant__ MediumDev M, size_t n, float *__restrict__ P, float *__restrict__ V,
                   const int32_t *__restrict__ nsteps, float *__restrict__ Aout, float *__restrict__ Bout) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float3 p = f3(P[3 * i], P[3 * i + 1], P[3 * i + 2]), v = f3(V[3 * i], V[3 * i + 1], V[3 * i + 2]);
        M3 A = m3_zero(), B = m3_identity();
        Field F;
        rif_field(M.rif, p, F);
        for (int k = 0; k < nsteps[i]; k++) er_derivativestep_fused(M.rif, p, v, A, B, F, M.h);
        P[3 * i] = p.x; P[3 * i + 1] = p.y; P[3 * i + 2] = p.z;
        V[3 * i] = v.x; V[3 * i + 1] = v.y; V[3 * i + 2] = v.z;
        for (int k = 0; k < 9; k++) { Aout[9 * i + k] = A.m[k]; Bout[9 * i + k] = B.m[k]; }
    }
}

__global__ void __launch_bounds__(128)
k_connection_residual(const __grid_constant__ MediumDev M, int precision, size_t n, const float *__restrict__ P1,
                      const float *__restrict__ P2, const float *__restrict__ V0, int isSensor, float *__restrict__ err,
                      float *__restrict__ derr, int32_t *__restrict__ status, int32_t *__restrict__ nsteps) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float3 e;
        M3 J;
        int count = 0;
        const float3 v0 = f3(V0[3 * i], V0[3 * i + 1], V0[3 * i + 2]), p1 = f3(P1[3 * i], P1[3 * i + 1], P1[3 * i + 2]), p2 = f3(P2[3 * i], P2[3 * i + 1], P2[3 * i + 2]);
        const bool refract = M.boundary == MER_BOUNDARY_HDIELECTRIC;
        const int st = M.shapeType == MER_SHAPE_SDF ? compute_fdf<true>(M, precision, v0, p1, p2, isSensor != 0, refract, e, J, count)
                                                    : compute_fdf<false>(M, precision, v0, p1, p2, isSensor != 0, refract, e, J, count);
        err[3 * i] = e.x; err[3 * i + 1] = e.y; err[3 * i + 2] = e.z;
        for (int k = 0; k < 9; k++) derr[9 * i + k] = J.m[k];
        if (status) status[i] = st;
        if (nsteps) nsteps[i] = count;
    }
}

struct ConnectOut {
    uint8_t *success;
    float *dirToP2, *revDir, *opl, *dist, *weight, *transmittance, *pdfSuccess, *pdfFailure;
    int32_t *evals;
};

/* makeDirectConnections (:1087-1163) with a Levenberg-Marquardt minimiser in place of Ceres BFGS, then eval()'s
 * pdfs / transmittance (:585-617).  The Sampler of connection i is the Philox stream (seed, i). */
__global__ void __launch_bounds__(128)
k_connect(const __grid_constant__ MediumDev M, int precision, float tol2, float rrweight, int maxIterations, size_t nConn,
          const float *__restrict__ P1, const float *__restrict__ P2, const float *__restrict__ D, int isSensor, int straightFirst,
          unsigned long long seed, ConnectOut out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nConn; i += (size_t) gridDim.x * blockDim.x) {
        const float3 p1 = f3(P1[3 * i], P1[3 * i + 1], P1[3 * i + 2]), p2 = f3(P2[3 * i], P2[3 * i + 1], P2[3 * i + 2]);
        const float3 din = f3(D[3 * i], D[3 * i + 1], D[3 * i + 2]);
        PathRng rng;
        rng.init(seed, (unsigned long long) i, 0u);
        ConnectResult R;
        if (M.shapeType == MER_SHAPE_SDF)
            connect_solve<true, true>(M, precision, tol2, rrweight, maxIterations, p1, p2, din, isSensor != 0, M.boundary == MER_BOUNDARY_HDIELECTRIC, straightFirst != 0, rng, R);
        else
            connect_solve<false, true>(M, precision, tol2, rrweight, maxIterations, p1, p2, din, isSensor != 0, M.boundary == MER_BOUNDARY_HDIELECTRIC, straightFirst != 0, rng, R);
        const bool success = R.success;
        const float weight = R.weight, opl = R.opl, dist = R.dist;
        const float3 dir = R.dir, rev = R.rev;
        const int evals = R.evals;
        float T[3] = {0.f, 0.f, 0.f}, ps = 1.0f, pf = 1.0f;
        if (success) { /* eval(), :585-617 */
            float pdfSuccess = 0.0f, pdfFailure = 0.0f;
            if (M.strategy == MER_STRATEGY_BALANCE) {
                for (int c = 0; c < 3; c++) {
                    const float t = fastexp_dev(__fmul_rn(-M.sigmaT[c], dist));
                    pdfSuccess = __fadd_rn(pdfSuccess, __fmul_rn(M.sigmaT[c], t));
                    pdfFailure = __fadd_rn(pdfFailure, t);
                }
                pdfSuccess = __fdiv_rn(pdfSuccess, 3.0f);
                pdfFailure = __fdiv_rn(pdfFailure, 3.0f);
            } else {
                const float t = fastexp_dev(__fmul_rn(-M.samplingDensity, dist));
                pdfSuccess = __fmul_rn(M.samplingDensity, t);
                pdfFailure = t;
            }
            float tmax = 0.0f;
            for (int c = 0; c < 3; c++) {
                T[c] = __fmul_rn(fastexp_dev(__fmul_rn(M.sigmaT[c], -dist)), weight);
                tmax = fmaxf(tmax, T[c]);
            }
            ps = __fmul_rn(pdfSuccess, M.weight);
            pf = __fadd_rn(__fmul_rn(pdfFailure, M.weight), 1.0f - M.weight);
            if (tmax < 1e-20f) T[0] = T[1] = T[2] = 0.0f;
        }
        out.success[i] = success ? 1 : 0;
        out.dirToP2[3 * i] = dir.x; out.dirToP2[3 * i + 1] = dir.y; out.dirToP2[3 * i + 2] = dir.z;
        out.revDir[3 * i] = rev.x; out.revDir[3 * i + 1] = rev.y; out.revDir[3 * i + 2] = rev.z;
        out.opl[i] = opl; out.dist[i] = dist; out.weight[i] = weight;
        out.transmittance[3 * i] = T[0]; out.transmittance[3 * i + 1] = T[1]; out.transmittance[3 * i + 2] = T[2];
        out.pdfSuccess[i] = ps; out.pdfFailure[i] = pf; out.evals[i] = evals;
    }
}

struct DevBuf {
    void *ptr = nullptr;
    ~DevBuf() { cudaFree(ptr); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&ptr, bytes ? bytes : 1); }
    template <typename T> T *as() { return (T *) ptr; }
};

} /* namespace */

#define UP(buf, host, bytes)                                                          \
    do {                                                                              \
        MER_CUDA(buf.alloc(bytes));                                                   \
        if (host) MER_CUDA(cudaMemcpy(buf.ptr, host, bytes, cudaMemcpyHostToDevice)); \
    } while (0)
#define DOWN(host, buf, bytes)                                                        \
    do {                                                                              \
        if (host) MER_CUDA(cudaMemcpy(host, buf.ptr, bytes, cudaMemcpyDeviceToHost)); \
    } while (0)

extern "C" {

int mer_rif_eval_hessian_batch(const mer_rif *r, size_t n, const float *p, float *value_out, float *grad_out, float *hess_out) {
    MER_REQUIRE(r && (n == 0 || p), "null argument");
    if (r->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "the Hessian needs the tricubic spline mode");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(r->device);
    DevBuf dp, df, dg, dh;
    UP(dp, p, n * 12); UP(df, (void *) nullptr, n * 4); UP(dg, (void *) nullptr, n * 12); UP(dh, (void *) nullptr, n * 36);
    MER_LAUNCH(k_rif_hessian, (unsigned) std::min<size_t>(mer_blocks(n, 128), 148u * 8u), 128, 0, 0, r->dev, n, dp.as<float>(), df.as<float>(),
               dg.as<float>(), dh.as<float>());
    DOWN(value_out, df, n * 4); DOWN(grad_out, dg, n * 12); DOWN(hess_out, dh, n * 36);
    return MER_OK;
}

int mer_medium_derivative_trace_batch(const mer_medium *m, size_t n, float *p, float *v, const int32_t *nsteps, float *dpdv0_out,
                                      float *dvdv0_out) {
    MER_REQUIRE(m && (n == 0 || (p && v && nsteps && dpdv0_out && dvdv0_out)), "null argument");
    if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "derivative steps need the tricubic spline mode");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf dp, dv, dn, da, db;
    UP(dp, p, n * 12); UP(dv, v, n * 12); UP(dn, nsteps, n * 4); UP(da, (void *) nullptr, n * 36); UP(db, (void *) nullptr, n * 36);
    MER_LAUNCH(k_derivative_trace, (unsigned) std::min<size_t>(mer_blocks(n, 128), 148u * 8u), 128, 0, 0, m->dev, n, dp.as<float>(),
               dv.as<float>(), dn.as<int32_t>(), da.as<float>(), db.as<float>());
    DOWN(p, dp, n * 12); DOWN(v, dv, n * 12); DOWN(dpdv0_out, da, n * 36); DOWN(dvdv0_out, db, n * 36);
    return MER_OK;
}

int mer_medium_connection_residual_batch(const mer_medium *m, int boundary_precision, size_t n, const float *p1, const float *p2,
                                         const float *v0, int is_sensor_sample, float *error_out, float *derror_out,
                                         int32_t *status_out, int32_t *nsteps_out) {
    MER_REQUIRE(m && (n == 0 || (p1 && p2 && v0 && error_out && derror_out)), "null argument");
    MER_REQUIRE(boundary_precision >= 0 && boundary_precision <= 9, "boundaryprecision out of range");
    if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "direct connections need the tricubic spline mode");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf d1, d2, dv, de, dj, ds, dn;
    UP(d1, p1, n * 12); UP(d2, p2, n * 12); UP(dv, v0, n * 12);
    UP(de, (void *) nullptr, n * 12); UP(dj, (void *) nullptr, n * 36); UP(ds, (void *) nullptr, n * 4); UP(dn, (void *) nullptr, n * 4);
    MER_LAUNCH(k_connection_residual, (unsigned) std::min<size_t>(mer_blocks(n, 128), 148u * 8u), 128, 0, 0, m->dev, boundary_precision, n,
               d1.as<float>(), d2.as<float>(), dv.as<float>(), is_sensor_sample, de.as<float>(), dj.as<float>(), ds.as<int32_t>(),
               dn.as<int32_t>());
    DOWN(error_out, de, n * 12); DOWN(derror_out, dj, n * 36); DOWN(status_out, ds, n * 4); DOWN(nsteps_out, dn, n * 4);
    return MER_OK;
}

int mer_medium_connect_batch(const mer_medium *m, const mer_connection_params *cp, size_t n, const float *p1, const float *p2,
                             const float *seed_dir, int is_sensor_sample, uint64_t seed, mer_connection_records *rec) {
    MER_REQUIRE(m && cp && rec && (n == 0 || (p1 && p2 && seed_dir)), "null argument");
    MER_REQUIRE(cp->tol2 > 0.0f && cp->rrweight > 0.0f && cp->rrweight < 1.0f, "tol2 must be positive and rrweight in (0, 1)");
    MER_REQUIRE(cp->boundary_precision >= 0 && cp->boundary_precision <= 9 && cp->max_iterations >= 1, "bad solver parameters");
    MER_REQUIRE(rec->success && rec->dir_to_p2 && rec->rev_dir_to_p1 && rec->optical_length && rec->distance && rec->weight &&
                rec->transmittance && rec->pdf_success && rec->pdf_failure && rec->evaluations, "every record buffer is required");
    if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "direct connections need the tricubic spline mode");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(m->device);
    DevBuf d1, d2, dd, dok, ddir, drev, dopl, ddist, dw, dT, dps, dpf, dev;
    UP(d1, p1, n * 12); UP(d2, p2, n * 12); UP(dd, seed_dir, n * 12);
    UP(dok, (void *) nullptr, n); UP(ddir, (void *) nullptr, n * 12); UP(drev, (void *) nullptr, n * 12); UP(dopl, (void *) nullptr, n * 4);
    UP(ddist, (void *) nullptr, n * 4); UP(dw, (void *) nullptr, n * 4); UP(dT, (void *) nullptr, n * 12); UP(dps, (void *) nullptr, n * 4);
    UP(dpf, (void *) nullptr, n * 4); UP(dev, (void *) nullptr, n * 4);
    ConnectOut o = {dok.as<uint8_t>(), ddir.as<float>(), drev.as<float>(), dopl.as<float>(), ddist.as<float>(), dw.as<float>(), dT.as<float>(),
                    dps.as<float>(), dpf.as<float>(), dev.as<int32_t>()};
    MER_LAUNCH(k_connect, (unsigned) std::min<size_t>(mer_blocks(n, 128), 148u * 8u), 128, 0, 0, m->dev, cp->boundary_precision, cp->tol2,
               cp->rrweight, cp->max_iterations, n, d1.as<float>(), d2.as<float>(), dd.as<float>(), is_sensor_sample,
               cp->start_mode == MER_START_STRAIGHT ? 1 : 0, (unsigned long long) seed, o);
    DOWN(rec->success, dok, n); DOWN(rec->dir_to_p2, ddir, n * 12); DOWN(rec->rev_dir_to_p1, drev, n * 12);
    DOWN(rec->optical_length, dopl, n * 4); DOWN(rec->distance, ddist, n * 4); DOWN(rec->weight, dw, n * 4);
    DOWN(rec->transmittance, dT, n * 12); DOWN(rec->pdf_success, dps, n * 4); DOWN(rec->pdf_failure, dpf, n * 4);
    DOWN(rec->evaluations, dev, n * 4);
    return MER_OK;
}

} /* extern "C" */
