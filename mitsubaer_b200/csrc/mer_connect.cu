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

namespace {

struct M3 { float m[9]; };

__device__ __forceinline__ M3 m3_identity() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1.0f : 0.0f; return r; }
__device__ __forceinline__ M3 m3_zero() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0.0f; return r; }
__device__ __forceinline__ M3 m3_mul(const M3 &A, const M3 &B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
__device__ __forceinline__ M3 m3_outer(float3 a, float3 b) {
    M3 C;
    const float av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.m[3 * i + j] = av[i] * bv[j];
    return C;
}
__device__ __forceinline__ float3 m3_premult(const M3 &A, float3 x) {
    return f3(x.x * A.m[0] + x.y * A.m[3] + x.z * A.m[6], x.x * A.m[1] + x.y * A.m[4] + x.z * A.m[7], x.x * A.m[2] + x.y * A.m[5] + x.z * A.m[8]);
}

/* symmetric Hessian: xx, yy, zz, xy, yz, zx */
struct Field { float n; float3 G; float H[6]; };
__device__ __forceinline__ M3 hess_full(const Field &F) {
    M3 r;
    r.m[0] = F.H[0]; r.m[1] = F.H[3]; r.m[2] = F.H[5];
    r.m[3] = F.H[3]; r.m[4] = F.H[1]; r.m[5] = F.H[4];
    r.m[6] = F.H[5]; r.m[7] = F.H[4]; r.m[8] = F.H[2];
    return r;
}

/* valueGradientAndHessian: the 64-tap stencil with the three kernels beta, beta', beta'' (basisspline.h:39-114),
 * separable contraction x -> y -> z; world-space rotation of gradient and Hessian (splinevolume.cpp:371-377) */
__device__ __forceinline__ void bs_weights2(float x, float fx, float w0[4], float w1[4], float w2[4]) {
    bs_weights(x, fx, w0, w1);
    const float a0 = x - (fx - 1.0f), a1 = x - fx, a2 = -(x - (fx + 1.0f)), a3 = -(x - (fx + 2.0f));
    w2[0] = 2.0f - a0;
    w2[1] = 3.0f * a1 - 2.0f;
    w2[2] = 3.0f * a2 - 2.0f;
    w2[3] = 2.0f - a3;
}

__device__ void rif_field(const RifDev &R, float3 pw, Field &F) {
    const float3 pv = rif_to_volume(R, pw);
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1], z = (pv.z - R.xmin[2]) * R.xres[2];
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int i0 = (int) fx, j0 = (int) fy, k0 = (int) fz;
    float wx0[4], wx1[4], wx2[4], wy0[4], wy1[4], wy2[4], wz0[4], wz1[4], wz2[4];
    bs_weights2(x, fx, wx0, wx1, wx2);
    bs_weights2(y, fy, wy0, wy1, wy2);
    bs_weights2(z, fz, wz0, wz1, wz2);
    const int N0 = R.N[0], N1 = R.N[1], N2 = R.N[2];
    const float4 *base = R.coeff8 + 2 * (size_t) clampi(i0, 0, N0 - 1);
    const size_t rowA = 2 * (size_t) clampi(j0 - 1, 0, N1 - 1) * (size_t) N0, rowB = 2 * (size_t) clampi(j0 + 1, 0, N1 - 1) * (size_t) N0;
    float f = 0, gx = 0, gy = 0, gz = 0, hxx = 0, hyy = 0, hzz = 0, hxy = 0, hyz = 0, hzx = 0;
#pragma unroll
    for (int dz = 0; dz < 4; dz++) {
        const size_t slab = 2 * (size_t) clampi(k0 - 1 + dz, 0, N2 - 1) * (size_t) N0 * (size_t) N1;
        float4 c[4];
        ldg256(base + slab + rowA, c[0], c[1]);
        ldg256(base + slab + rowB, c[2], c[3]);
        float b00 = 0, b10 = 0, b20 = 0, b01 = 0, b11 = 0, b02 = 0;
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            const float4 q = c[dy];
            const float a0 = q.x * wx0[0] + q.y * wx0[1] + q.z * wx0[2] + q.w * wx0[3];
            const float a1 = q.x * wx1[0] + q.y * wx1[1] + q.z * wx1[2] + q.w * wx1[3];
            const float a2 = q.x * wx2[0] + q.y * wx2[1] + q.z * wx2[2] + q.w * wx2[3];
            b00 = fmaf(a0, wy0[dy], b00); b10 = fmaf(a1, wy0[dy], b10); b20 = fmaf(a2, wy0[dy], b20);
            b01 = fmaf(a0, wy1[dy], b01); b11 = fmaf(a1, wy1[dy], b11); b02 = fmaf(a0, wy2[dy], b02);
        }
        f = fmaf(b00, wz0[dz], f); gx = fmaf(b10, wz0[dz], gx); gy = fmaf(b01, wz0[dz], gy); gz = fmaf(b00, wz1[dz], gz);
        hxx = fmaf(b20, wz0[dz], hxx); hyy = fmaf(b02, wz0[dz], hyy); hzz = fmaf(b00, wz2[dz], hzz);
        hxy = fmaf(b11, wz0[dz], hxy); hyz = fmaf(b01, wz1[dz], hyz); hzx = fmaf(b10, wz1[dz], hzx);
    }
    const float rx = R.xres[0], ry = R.xres[1], rz = R.xres[2];
    F.n = f;
    F.G = f3(gx * rx, gy * ry, gz * rz);
    F.H[0] = hxx * (rx * rx); F.H[1] = hyy * (ry * ry); F.H[2] = hzz * (rz * rz);
    F.H[3] = hxy * (rx * ry); F.H[4] = hyz * (ry * rz); F.H[5] = hzx * (rz * rx);
    if (R.hasXform) {
        F.G = rif_rot_t(R, F.G);
        M3 Rm, Rt, H = hess_full(F);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Rm.m[3 * i + j] = R.M[4 * i + j]; Rt.m[3 * i + j] = R.M[4 * j + i]; }
        H = m3_mul(m3_mul(Rt, H), Rm);
        F.H[0] = H.m[0]; F.H[1] = H.m[4]; F.H[2] = H.m[8]; F.H[3] = H.m[1]; F.H[4] = H.m[5]; F.H[5] = H.m[2];
    }
}

/* er_derivativestep (:798-814) with the field carried across steps like er_step_fused: on entry F is the field at
 * p, on exit the field at the new p.  p and v are updated with the same individually rounded operations as
 * er_step, so the trajectory is the one trace() follows. */
__device__ __forceinline__ void er_derivativestep_fused(const RifDev &R, float3 &p, float3 &v, M3 &A, M3 &B, Field &F, float h) {
    const float hs = __fmul_rn(0.5f, h);
    v = f3(__fadd_rn(v.x, __fmul_rn(hs, F.G.x)), __fadd_rn(v.y, __fmul_rn(hs, F.G.y)), __fadd_rn(v.z, __fmul_rn(hs, F.G.z)));
    M3 HA = m3_mul(hess_full(F), A);
#pragma unroll
    for (int i = 0; i < 9; i++) B.m[i] = fmaf(hs, HA.m[i], B.m[i]);
    const float recip = __frcp_rn(F.n);
    p = f3(__fadd_rn(p.x, __fmul_rn(__fmul_rn(h, v.x), recip)), __fadd_rn(p.y, __fmul_rn(__fmul_rn(h, v.y), recip)),
           __fadd_rn(p.z, __fmul_rn(__fmul_rn(h, v.z), recip)));
    rif_field(R, p, F);
    const float invn = __frcp_rn(F.n);
    M3 VGA = m3_mul(m3_outer(v, F.G), A);
#pragma unroll
    for (int i = 0; i < 9; i++) A.m[i] = fmaf(h, fmaf(-invn * invn, VGA.m[i], invn * B.m[i]), A.m[i]);
    v = f3(__fadd_rn(v.x, __fmul_rn(hs, F.G.x)), __fadd_rn(v.y, __fmul_rn(hs, F.G.y)), __fadd_rn(v.z, __fmul_rn(hs, F.G.z)));
    HA = m3_mul(hess_full(F), A);
#pragma unroll
    for (int i = 0; i < 9; i++) B.m[i] = fmaf(hs, HA.m[i], B.m[i]);
}

__device__ __forceinline__ float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

/* boundaryVelocityDerivative, :1057-1074 */
__device__ void boundary_velocity_derivative(float3 &v, M3 &B, float3 dtb, float3 dnb, float3 N, float ni, float ne) {
    const float dotp = dot3(v, N);
    float r = ne / ni;
    r = r * r - 1.0f;
    const float n2 = dot3(v, v);
    float sq = r * n2 + dotp * dotp;
    M3 NN = m3_outer(N, N), S = m3_outer(dnb, dtb), L;
    for (int i = 0; i < 9; i++) S.m[i] += B.m[i];
    if (sq < MER_EPSILON) {
        v = f3(2.0f * dotp * N.x - v.x, 2.0f * dotp * N.y - v.y, 2.0f * dotp * N.z - v.z);
        for (int i = 0; i < 9; i++) L.m[i] = 2.0f * NN.m[i] - ((i % 4 == 0) ? 1.0f : 0.0f);
        B = m3_mul(L, S);
        return;
    }
    sq = sqrtf(sq);
    const float sg = sgnf(dotp);
    const float3 w = f3((r * v.x + dotp * N.x) / sq, (r * v.y + dotp * N.y) / sq, (r * v.z + dotp * N.z) / sq);
    M3 NW = m3_outer(N, w);
    for (int i = 0; i < 9; i++) L.m[i] = ((i % 4 == 0) ? 1.0f : 0.0f) - NN.m[i] + sg * NW.m[i];
    B = m3_mul(L, S);
    v = f3(v.x - dotp * N.x + sg * sq * N.x, v.y - dotp * N.y + sg * sq * N.y, v.z - dotp * N.z + sg * sq * N.z);
}

/* computefdfBDPT, :816-939.  status: 0 closest approach inside the shape, 1 left the object (Snell + straight
 * extension; needs the sdf child), 2 degenerate (error = p1 - p2, Jacobian 0). */
__device__ int compute_fdf(const MediumDev &M, int precision, float3 vi, float3 p1, float3 p2, bool isSensorSample,
                           float3 &err, M3 &derr, int &count) {
    M3 A = m3_zero(), B = m3_identity();
    derr = m3_zero();
    err = p1 - p2;
    if (!rif_inside_limits(M.hasSdf ? M.sdf : M.rif, p1)) return 2;
    float h = M.h;
    int nBisect = (int) ceil((double) precision / log10(2.0));
    bool leftObject = false;
    float3 p = p1, v = vi, oldp, oldv;
    M3 oldA, oldB;
    bool signOld = signbit(dot3(p - p2, v)), signNew;
    Field F;
    rif_field(M.rif, p, F);
    { /* renormalise the launch velocity to |v| = n(p1) and chain B through it, :838-843 */
        const float r = F.n, n1 = sqrtf(dot3(vi, vi)), n2 = n1 * n1, n3 = n2 * n1;
        M3 P = m3_outer(v, v);
        for (int i = 0; i < 9; i++) P.m[i] = (r / n3) * (n2 * ((i % 4 == 0) ? 1.0f : 0.0f) - P.m[i]);
        B = m3_mul(P, B);
        const float recip = 1.0f / n1;
        v = f3((v.x * recip) * r, (v.y * recip) * r, (v.z * recip) * r);
    }
    Field oldF;
    for (int it = 0; it < 100000; it++) {
        oldp = p; oldv = v; oldA = A; oldB = B; oldF = F;
        er_derivativestep_fused(M.rif, p, v, A, B, F, h);
        count++;
        signNew = signbit(dot3(p - p2, v));
        if (signNew != signOld) {
            while (nBisect > 0) {
                nBisect--;
                p = oldp; v = oldv; A = oldA; B = oldB; F = oldF;
                h = h / 2;
                er_derivativestep_fused(M.rif, p, v, A, B, F, h);
                count++;
                signNew = signbit(dot3(p - p2, v));
                if (signNew == signOld) { oldp = p; oldv = v; oldA = A; oldB = B; oldF = F; }
            }
            break;
        } else if (!inside_shape(M, p)) {
            while (nBisect > 0) {
                nBisect--;
                p = oldp; v = oldv; A = oldA; B = oldB; F = oldF;
                h = h / 2;
                er_derivativestep_fused(M.rif, p, v, A, B, F, h);
                count++;
                if (inside_shape(M, p)) { oldp = p; oldv = v; oldA = A; oldB = B; oldF = F; }
            }
            const float3 dp1 = p - p1;
            if (dot3(dp1, dp1) < MER_EPSILON || !M.hasSdf) return M.hasSdf ? 2 : 1;
            const float nb = F.n;
            const float3 dnb = F.G;
            const float rn = 1.0f / nb;
            const float3 dpdtb = f3(v.x * rn, v.y * rn, v.z * rn);
            float sv;
            float3 N;
            rif_tricubic(M.sdf, rif_to_volume(M.sdf, p), sv, N);
            N = rif_rot_t(M.sdf, N);
            const float nl = 1.0f / sqrtf(dot3(N, N));
            N = f3(N.x * nl, N.y * nl, N.z * nl);
            float3 dtb = m3_premult(A, N);
            const float den = dot3(N, dpdtb);
            dtb = f3(-dtb.x / den, -dtb.y / den, -dtb.z / den);
            boundary_velocity_derivative(v, B, dtb, dnb, N, nb, 1.0f);
            const float extra_t = -dot3(v, p - p2) / dot3(v, v);
            leftObject = true;
            if (isSensorSample && extra_t < 0.0f) return 2;
            M3 O = m3_outer(dpdtb - v, dtb);
            for (int i = 0; i < 9; i++) A.m[i] += O.m[i] + extra_t * B.m[i];
            p = f3(p.x + extra_t * v.x, p.y + extra_t * v.y, p.z + extra_t * v.z);
            break;
        }
    }
    const float3 d = p - p2, a = m3_premult(A, v), b = m3_premult(B, d);
    float3 dpdt, dts;
    if (!leftObject) {
        /* :924 evaluates the field at the final p: that is the F carried by the fused stepper */
        const float rr = 1.0f / F.n;
        dpdt = f3(v.x * rr, v.y * rr, v.z * rr);
        const float den = dot3(v, dpdt) + dot3(d, F.G);
        dts = f3(-(a.x + b.x) / den, -(a.y + b.y) / den, -(a.z + b.z) / den);
    } else {
        dpdt = v;
        const float den = dot3(v, dpdt);
        dts = f3(-(a.x + b.x) / den, -(a.y + b.y) / den, -(a.z + b.z) / den);
    }
    M3 O = m3_outer(dpdt, dts);
    err = d;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) derr.m[3 * i + j] = A.m[3 * j + i] + O.m[3 * j + i]; /* transposed, :936-938 */
    return leftObject ? 1 : 0;
}

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
        int st = compute_fdf(M, precision, f3(V0[3 * i], V0[3 * i + 1], V0[3 * i + 2]), f3(P1[3 * i], P1[3 * i + 1], P1[3 * i + 2]),
                             f3(P2[3 * i], P2[3 * i + 1], P2[3 * i + 2]), isSensor != 0, e, J, count);
        err[3 * i] = e.x; err[3 * i + 1] = e.y; err[3 * i + 2] = e.z;
        for (int k = 0; k < 9; k++) derr[9 * i + k] = J.m[k];
        if (status) status[i] = st;
        if (nsteps) nsteps[i] = count;
    }
}

/* ------------------------------------------------------------------ makeDirectConnections / eval
 * computePathLengthsTillClosestP2 (:941-1030): re-trace with plain er_step, geometric length and the
 * midpoint-rule optical length, same closest-approach / boundary halvings. */
__device__ bool compute_path_lengths(const MediumDev &M, int precision, float tol2, float3 p1, float3 p2, float3 dirToP2, float3 &revDir,
                                     bool isSensorSample, float &opl, float &dist) {
    dist = 0.0f;
    opl = 0.0f;
    float h = M.h, n, oldn, dummy = 0.0f;
    int nBisect = (int) ceil((double) precision / log10(2.0));
    float3 p = p1, v = dirToP2, oldp, oldv, G, oldG;
    StencilCache<MER_RIF_TRICUBIC> S;
    S.invalidate();
    rif_lookup_cached<MER_RIF_TRICUBIC>(M.rif, p, S, n, G);
    bool signOld = signbit(dot3(p - p2, v)), signNew;
    for (int it = 0; it < 100000; it++) {
        oldp = p; oldv = v; oldn = n; oldG = G;
        er_step_fused<MER_RIF_TRICUBIC, false>(M.rif, S, p, v, n, G, h, dummy);
        signNew = signbit(dot3(p - p2, v));
        if (!inside_shape(M, p)) {
            if (!isSensorSample || !M.hasSdf) return false;
            while (nBisect > 0) {
                nBisect--;
                p = oldp; v = oldv; n = oldn; G = oldG;
                h = h / 2;
                er_step_fused<MER_RIF_TRICUBIC, false>(M.rif, S, p, v, n, G, h, dummy);
                if (inside_shape(M, p)) {
                    float nm;
                    float3 gm;
                    rif_lookup<MER_RIF_TRICUBIC>(M.rif, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z)), nm, gm);
                    dist += h;
                    opl += h * nm;
                    oldp = p; oldv = v; oldn = n; oldG = G;
                }
            }
            float sv;
            float3 N;
            rif_tricubic(M.sdf, rif_to_volume(M.sdf, p), sv, N);
            N = rif_rot_t(M.sdf, N);
            const float nl = 1.0f / sqrtf(dot3(N, N));
            N = f3(N.x * nl, N.y * nl, N.z * nl);
            { /* boundaryVelocity, :1036-1051, exterior index 1 */
                const float dotp = dot3(v, N);
                float r = 1.0f / n;
                r = r * r - 1.0f;
                float sq = r * dot3(v, v) + dotp * dotp;
                if (sq < MER_EPSILON) {
                    v = f3(2.0f * dotp * N.x - v.x, 2.0f * dotp * N.y - v.y, 2.0f * dotp * N.z - v.z);
                } else {
                    sq = sqrtf(sq);
                    const float sg = sgnf(dotp);
                    v = f3(v.x - dotp * N.x + sg * sq * N.x, v.y - dotp * N.y + sg * sq * N.y, v.z - dotp * N.z + sg * sq * N.z);
                }
            }
            const float extra_t = -dot3(v, p - p2) / dot3(v, v);
            if (extra_t < 0.0f) return false;
            p = f3(p.x + extra_t * v.x, p.y + extra_t * v.y, p.z + extra_t * v.z);
            opl += extra_t;
            break;
        }
        if (signNew != signOld) {
            while (nBisect > 0) {
                nBisect--;
                p = oldp; v = oldv; n = oldn; G = oldG;
                h = h / 2;
                er_step_fused<MER_RIF_TRICUBIC, false>(M.rif, S, p, v, n, G, h, dummy);
                signNew = signbit(dot3(p - p2, v));
                if (signNew == signOld) {
                    float nm;
                    float3 gm;
                    rif_lookup<MER_RIF_TRICUBIC>(M.rif, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z)), nm, gm);
                    dist += h;
                    opl += h * nm;
                    oldp = p; oldv = v; oldn = n; oldG = G;
                }
            }
            break;
        } else {
            float nm;
            float3 gm;
            rif_lookup<MER_RIF_TRICUBIC>(M.rif, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z)), nm, gm);
            dist += h;
            opl += h * nm;
        }
    }
    const float3 d = p - p2;
    if (dot3(d, d) > tol2) return false;
    const float vl = 1.0f / sqrtf(dot3(v, v));
    revDir = f3(-(v.x * vl), -(v.y * vl), -(v.z * vl));
    return true;
}

__device__ __forceinline__ bool solve3(const float *Mx, const float *b, float *x) {
    const float det = Mx[0] * (Mx[4] * Mx[8] - Mx[5] * Mx[7]) - Mx[1] * (Mx[3] * Mx[8] - Mx[5] * Mx[6]) + Mx[2] * (Mx[3] * Mx[7] - Mx[4] * Mx[6]);
    if (!(fabsf(det) > 0.0f)) return false;
    const float inv = 1.0f / det;
    x[0] = inv * (b[0] * (Mx[4] * Mx[8] - Mx[5] * Mx[7]) - Mx[1] * (b[1] * Mx[8] - Mx[5] * b[2]) + Mx[2] * (b[1] * Mx[7] - Mx[4] * b[2]));
    x[1] = inv * (Mx[0] * (b[1] * Mx[8] - Mx[5] * b[2]) - b[0] * (Mx[3] * Mx[8] - Mx[5] * Mx[6]) + Mx[2] * (Mx[3] * b[2] - b[1] * Mx[6]));
    x[2] = inv * (Mx[0] * (Mx[4] * b[2] - b[1] * Mx[7]) - Mx[1] * (Mx[3] * b[2] - b[1] * Mx[6]) + b[0] * (Mx[3] * Mx[7] - Mx[4] * Mx[6]));
    return true;
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
          const float *__restrict__ P1, const float *__restrict__ P2, const float *__restrict__ D, int isSensor, unsigned long long seed,
          ConnectOut out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nConn; i += (size_t) gridDim.x * blockDim.x) {
        const float3 p1 = f3(P1[3 * i], P1[3 * i + 1], P1[3 * i + 2]), p2 = f3(P2[3 * i], P2[3 * i + 1], P2[3 * i + 2]);
        const float3 din = f3(D[3 * i], D[3 * i + 1], D[3 * i + 2]);
        PathRng rng;
        rng.init(seed, (unsigned long long) i, 0u);
        bool success = false, converged = false;
        float weight = 1.0f, opl = 0.0f, dist = 0.0f, T[3] = {0.f, 0.f, 0.f}, ps = 1.0f, pf = 1.0f;
        float3 dir = f3(0.f, 0.f, 0.f), rev = f3(0.f, 0.f, 0.f);
        int evals = 0;
        if (rif_inside_limits(M.hasSdf ? M.sdf : M.rif, p1)) {
            float RIFp;
            float3 g0;
            rif_lookup<MER_RIF_TRICUBIC>(M.rif, p1, RIFp, g0);
            float x[3];
            while (true) {
                float3 ax, ay;
                coordinate_system(din, ax, ay);
                const float u1 = rng.next(), u2 = rng.next();
                const float z = u1, tmp = sqrtf(fmaxf(0.0f, 1.0f - z * z)), phi = (float) (2.0 * 3.14159265358979323846 * (double) u2);
                float sp, cp;
                sincosf(phi, &sp, &cp);
                const float lx = cp * tmp, ly = sp * tmp;
                x[0] = (lx * ax.x + ly * ay.x + z * din.x) * RIFp;
                x[1] = (lx * ax.y + ly * ay.y + z * din.y) * RIFp;
                x[2] = (lx * ax.z + ly * ay.z + z * din.z) * RIFp;
                float3 r;
                M3 Jt;
                int cnt = 0;
                compute_fdf(M, precision, f3(x[0], x[1], x[2]), p1, p2, isSensor != 0, r, Jt, cnt);
                evals++;
                float cost = 0.5f * dot3(r, r), lambda = 0.0f;
                int accepted = 0;
                for (int ev = 0; ev < 2 * maxIterations && accepted < maxIterations && !(cost < tol2); ev++) {
                    float JTJ[9], g[3], dx[3];
                    const float rv[3] = {r.x, r.y, r.z};
                    for (int a = 0; a < 3; a++) {
                        g[a] = -(Jt.m[3 * a] * rv[0] + Jt.m[3 * a + 1] * rv[1] + Jt.m[3 * a + 2] * rv[2]);
                        for (int b = 0; b < 3; b++) JTJ[3 * a + b] = Jt.m[3 * a] * Jt.m[3 * b] + Jt.m[3 * a + 1] * Jt.m[3 * b + 1] + Jt.m[3 * a + 2] * Jt.m[3 * b + 2];
                    }
                    if (lambda == 0.0f) lambda = 1e-3f * fmaxf(fmaxf(JTJ[0], JTJ[4]), fmaxf(JTJ[8], 1e-12f));
                    JTJ[0] += lambda; JTJ[4] += lambda; JTJ[8] += lambda;
                    if (!solve3(JTJ, g, dx)) break;
                    float3 rn;
                    M3 Jn;
                    compute_fdf(M, precision, f3(x[0] + dx[0], x[1] + dx[1], x[2] + dx[2]), p1, p2, isSensor != 0, rn, Jn, cnt);
                    evals++;
                    const float costn = 0.5f * dot3(rn, rn);
                    if (costn < cost) {
                        x[0] += dx[0]; x[1] += dx[1]; x[2] += dx[2];
                        r = rn; Jt = Jn; cost = costn;
                        lambda = fmaxf(lambda / 3.0f, 1e-15f);
                        accepted++;
                    } else {
                        lambda *= 4.0f;
                        if (lambda > 1e12f) break;
                    }
                }
                if (cost < tol2) { converged = true; break; }
                if (rng.next() < rrweight) weight = weight * (1.0f / rrweight);
                else break;
            }
            const float xl = 1.0f / sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
            dir = f3((x[0] * xl) * RIFp, (x[1] * xl) * RIFp, (x[2] * xl) * RIFp);
            if (converged && compute_path_lengths(M, precision, tol2, p1, p2, dir, rev, isSensor != 0, opl, dist)) {
                success = true;
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
               (unsigned long long) seed, o);
    DOWN(rec->success, dok, n); DOWN(rec->dir_to_p2, ddir, n * 12); DOWN(rec->rev_dir_to_p1, drev, n * 12);
    DOWN(rec->optical_length, dopl, n * 4); DOWN(rec->distance, ddist, n * 4); DOWN(rec->weight, dw, n * 4);
    DOWN(rec->transmittance, dT, n * 12); DOWN(rec->pdf_success, dps, n * 4); DOWN(rec->pdf_failure, dpf, n * 4);
    DOWN(rec->evaluations, dev, n * 4);
    return MER_OK;
}

} /* extern "C" */
