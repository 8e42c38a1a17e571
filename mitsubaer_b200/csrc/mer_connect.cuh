/*
 * mer_connect.cuh — device code of the curved direct connections, shared by the batch entry
 * points (mer_connect.cu) and by the integrator's next-event estimation (mer_render.cu).
 *
 *   Spline<3>::valueGradientAndHessian    include/mitsuba/core/basisspline.h:539-606 (+ splinevolume.cpp:371-377)
 *   er_derivativestep                      src/medium/heterogeneousrefractive.cpp:798-814
 *   computefdfBDPT                         :816-939 (closest approach by sign change + step halvings, boundary
 *                                          exit with Snell refraction: boundaryVelocityDerivative :1057-1074)
 *   computePathLengthsTillClosestP2        :941-1030
 *   makeDirectConnections                  :1087-1163 (Ceres 1.14 BFGS -> Levenberg-Marquardt on the same residual/Jacobian)
 *
 * 3x3 matrices are row-major float[9]; outer(a,b)_ij = a_i b_j (include/mitsuba/core/matrix.h:584-588);
 * premult(M, x) = M^T x (:765-769).
 */
#pragma once
#include "mer_device.cuh"

namespace merc {

struct M3 { float m[9]; };

static __device__ __forceinline__ M3 m3_identity() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = (i % 4 == 0) ? 1.0f : 0.0f; return r; }
static __device__ __forceinline__ M3 m3_zero() { M3 r; for (int i = 0; i < 9; i++) r.m[i] = 0.0f; return r; }
static __device__ __forceinline__ M3 m3_mul(const M3 &A, const M3 &B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
static __device__ __forceinline__ M3 m3_outer(float3 a, float3 b) {
    M3 C;
    const float av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.m[3 * i + j] = av[i] * bv[j];
    return C;
}
static __device__ __forceinline__ float3 m3_premult(const M3 &A, float3 x) {
    return f3(x.x * A.m[0] + x.y * A.m[3] + x.z * A.m[6], x.x * A.m[1] + x.y * A.m[4] + x.z * A.m[7], x.x * A.m[2] + x.y * A.m[5] + x.z * A.m[8]);
}

/* symmetric Hessian: xx, yy, zz, xy, yz, zx */
struct Field { float n; float3 G; float H[6]; };
static __device__ __forceinline__ M3 hess_full(const Field &F) {
    M3 r;
    r.m[0] = F.H[0]; r.m[1] = F.H[3]; r.m[2] = F.H[5];
    r.m[3] = F.H[3]; r.m[4] = F.H[1]; r.m[5] = F.H[4];
    r.m[6] = F.H[5]; r.m[7] = F.H[4]; r.m[8] = F.H[2];
    return r;
}

/* valueGradientAndHessian: the 64-tap stencil with the three kernels beta, beta', beta'' (basisspline.h:39-114),
 * separable contraction x -> y -> z; world-space rotation of gradient and Hessian (splinevolume.cpp:371-377) */
static __device__ __forceinline__ void bs_weights2(float x, float fx, float w0[4], float w1[4], float w2[4]) {
    bs_weights(x, fx, w0, w1);
    const float a0 = x - (fx - 1.0f), a1 = x - fx, a2 = -(x - (fx + 1.0f)), a3 = -(x - (fx + 2.0f));
    w2[0] = 2.0f - a0;
    w2[1] = 3.0f * a1 - 2.0f;
    w2[2] = 3.0f * a2 - 2.0f;
    w2[3] = 2.0f - a3;
}

static __device__ void rif_field(const RifDev &R, float3 pw, Field &F) {
    const float3 pv = rif_to_volume(R, pw);
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1], z = (pv.z - R.xmin[2]) * R.xres[2];
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int i0 = (int) fx, j0 = (int) fy, k0 = (int) fz;
    float wx0[4], wx1[4], wx2[4], wy0[4], wy1[4], wy2[4], wz0[4], wz1[4], wz2[4];
    bs_weights2(x, fx, wx0, wx1, wx2);
    bs_weights2(y, fy, wy0, wy1, wy2);
    bs_weights2(z, fz, wz0, wz1, wz2);
    const bool interior = rif_cell_interior(R, i0, j0, k0);
    float f = 0, gx = 0, gy = 0, gz = 0, hxx = 0, hyy = 0, hzz = 0, hxy = 0, hyz = 0, hzx = 0;
#pragma unroll
    for (int dz = 0; dz < 4; dz++) {
        float4 c[4];
        if (interior) rif_slab_interior<-1>(R, i0, j0, k0 - 1 + dz, c);
        else rif_slab_clamped(R, i0, j0, k0 - 1 + dz, c);
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
static __device__ __forceinline__ void er_derivativestep_fused(const RifDev &R, float3 &p, float3 &v, M3 &A, M3 &B, Field &F, float h) {
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

static __device__ __forceinline__ float sgnf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

/* boundaryVelocityDerivative, :1057-1074 */
static __device__ void boundary_velocity_derivative(float3 &v, M3 &B, float3 dtb, float3 dnb, float3 N, float ni, float ne) {
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

/* outward unit normal of the container at (or just inside) its surface: the normalised gradient of the `sdf` child when
 * there is one (:892-893), else the analytic normal of the box / sphere, which is the same thing for those shapes */
static __device__ float3 container_normal(const MediumDev &M, float3 p) {
    if (!M.hasSdf) return shape_normal(M, p);
    float sv;
    float3 N;
    rif_tricubic(M.sdf, rif_to_volume(M.sdf, p), sv, N);
    N = rif_rot_t(M.sdf, N);
    const float nl = 1.0f / sqrtf(dot3(N, N));
    return f3(N.x * nl, N.y * nl, N.z * nl);
}

/* computefdfBDPT, :816-939.  status: 0 closest approach inside the shape, 1 left the object (boundary + straight
 * extension), 2 degenerate (error = p1 - p2, Jacobian 0), 3 left the object by total internal reflection.  `refract` = the reference's behaviour (Snell to exterior
 * index 1, i.e. an hdielectric container); false = index-matched container: the velocity crosses unchanged. */
template <bool SDFSHAPE>
static __device__ __forceinline__ int compute_fdf(const MediumDev &M, int precision, float3 vi, float3 p1, float3 p2, bool isSensorSample,
                           bool refract, float3 &err, M3 &derr, int &count) {
    M3 A = m3_zero(), B = m3_identity();
    derr = m3_zero();
    err = p1 - p2;
    if (!rif_inside_limits(M.hasSdf ? M.sdf : M.rif, p1)) return 2;
    float h = M.h;
    int nBisect = (int) ceil((double) precision / log10(2.0));
    bool leftObject = false, tir = false;
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
        } else if (!(SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) {
            while (nBisect > 0) {
                nBisect--;
                p = oldp; v = oldv; A = oldA; B = oldB; F = oldF;
                h = h / 2;
                er_derivativestep_fused(M.rif, p, v, A, B, F, h);
                count++;
                if ((SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) { oldp = p; oldv = v; oldA = A; oldB = B; oldF = F; }
            }
            const float3 dp1 = p - p1;
            if (dot3(dp1, dp1) < M.minExit2) return 2;
            const float nb = F.n;
            const float3 dnb = F.G;
            const float rn = 1.0f / nb;
            const float3 dpdtb = f3(v.x * rn, v.y * rn, v.z * rn);
            const float3 N = container_normal(M, p);
            float3 dtb = m3_premult(A, N);
            const float den = dot3(N, dpdtb);
            dtb = f3(-dtb.x / den, -dtb.y / den, -dtb.z / den);
            if (refract) {
                const float dotp = dot3(v, N);
                float rr = 1.0f / nb;
                rr = rr * rr - 1.0f;
                tir = rr * dot3(v, v) + dotp * dotp < MER_EPSILON;
                boundary_velocity_derivative(v, B, dtb, dnb, N, nb, 1.0f);
            } else { /* d v(t_b) / d v0 = B + grad n (x) d t_b / d v0, v itself unchanged */
                const M3 S = m3_outer(dnb, dtb);
                for (int i = 0; i < 9; i++) B.m[i] += S.m[i];
            }
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
    return leftObject ? (tir ? 3 : 1) : 0; /* 3: total internal reflection at the boundary (the reference reflects and goes on, :1041-1044) */
}

/* ------------------------------------------------------------------ makeDirectConnections / eval
 * computePathLengthsTillClosestP2 (:941-1030): re-trace with plain er_step, geometric length and the
 * midpoint-rule optical length, same closest-approach / boundary halvings. */
struct ExitInfo { /* how a connection left the container (next-event estimation needs the boundary terms) */
    bool exited, tir;
    float tau;  /* optical depth of the density grid along the connection (midpoint rule on the re-trace's steps) */
    float nb;   /* RIF at the exit point */
    float cosI; /* cosine between the interior direction and the outward normal */
};

/* WANT_OPL = false skips the midpoint-rule optical length (a whole uncached 64-tap lookup per step of the re-trace) for
 * callers that only need the geometric length, the arrival direction and the boundary terms */
template <bool SDFSHAPE, bool WANT_OPL>
static __device__ __forceinline__ bool compute_path_lengths(const MediumDev &M, int precision, float tol2, float3 p1, float3 p2, float3 dirToP2, float3 &revDir,
                                     bool isSensorSample, bool refract, float &opl, float &dist, ExitInfo &ex) {
    dist = 0.0f;
    opl = 0.0f;
    ex.exited = ex.tir = false;
    ex.nb = 1.0f;
    ex.cosI = 1.0f;
    ex.tau = 0.0f;
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
        if (!(SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) {
            if (!isSensorSample) return false;
            while (nBisect > 0) {
                nBisect--;
                p = oldp; v = oldv; n = oldn; G = oldG;
                h = h / 2;
                er_step_fused<MER_RIF_TRICUBIC, false>(M.rif, S, p, v, n, G, h, dummy);
                if ((SDFSHAPE ? inside_shape_any(M, p) : inside_shape(M, p))) {
                    float nm = 0.0f;
                    float3 gm;
                    if (WANT_OPL) rif_lookup<MER_RIF_TRICUBIC>(M.rif, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z)), nm, gm);
                    dist += h;
                    opl += h * nm;
                    if (M.hasGrid) ex.tau += h * (grid_lookup(M.grid, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z))) * M.densityScale);
                    oldp = p; oldv = v; oldn = n; oldG = G;
                }
            }
            const float3 N = container_normal(M, p);
            ex.exited = true;
            ex.nb = n;
            ex.cosI = dot3(v, N) / sqrtf(dot3(v, v));
            if (refract) { /* boundaryVelocity, :1036-1051, exterior index 1 */
                const float dotp = dot3(v, N);
                float r = 1.0f / n;
                r = r * r - 1.0f;
                float sq = r * dot3(v, v) + dotp * dotp;
                if (sq < MER_EPSILON) {
                    ex.tir = true;
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
            opl += extra_t * sqrtf(dot3(v, v)); /* exterior index 1; |v| = 1 after Snell (:989), n_b when the boundary is index-matched */
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
                    float nm = 0.0f;
                    float3 gm;
                    if (WANT_OPL) rif_lookup<MER_RIF_TRICUBIC>(M.rif, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z)), nm, gm);
                    dist += h;
                    opl += h * nm;
                    if (M.hasGrid) ex.tau += h * (grid_lookup(M.grid, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z))) * M.densityScale);
                    oldp = p; oldv = v; oldn = n; oldG = G;
                }
            }
            break;
        } else {
            float nm = 0.0f;
            float3 gm;
            if (WANT_OPL) rif_lookup<MER_RIF_TRICUBIC>(M.rif, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z)), nm, gm);
            dist += h;
            opl += h * nm;
            if (M.hasGrid) ex.tau += h * (grid_lookup(M.grid, f3(0.5f * (p.x + oldp.x), 0.5f * (p.y + oldp.y), 0.5f * (p.z + oldp.z))) * M.densityScale);
        }
    }
    const float3 d = p - p2;
    if (dot3(d, d) > tol2) return false;
    const float vl = 1.0f / sqrtf(dot3(v, v));
    revDir = f3(-(v.x * vl), -(v.y * vl), -(v.z * vl));
    return true;
}

static __device__ __forceinline__ bool solve3(const float *Mx, const float *b, float *x) {
    const float det = Mx[0] * (Mx[4] * Mx[8] - Mx[5] * Mx[7]) - Mx[1] * (Mx[3] * Mx[8] - Mx[5] * Mx[6]) + Mx[2] * (Mx[3] * Mx[7] - Mx[4] * Mx[6]);
    if (!(fabsf(det) > 0.0f)) return false;
    const float inv = 1.0f / det;
    x[0] = inv * (b[0] * (Mx[4] * Mx[8] - Mx[5] * Mx[7]) - Mx[1] * (b[1] * Mx[8] - Mx[5] * b[2]) + Mx[2] * (b[1] * Mx[7] - Mx[4] * b[2]));
    x[1] = inv * (Mx[0] * (b[1] * Mx[8] - Mx[5] * b[2]) - b[0] * (Mx[3] * Mx[8] - Mx[5] * Mx[6]) + Mx[2] * (Mx[3] * b[2] - b[1] * Mx[6]));
    x[2] = inv * (Mx[0] * (Mx[4] * b[2] - b[1] * Mx[7]) - Mx[1] * (Mx[3] * b[2] - b[1] * Mx[6]) + b[0] * (Mx[3] * Mx[7] - Mx[4] * Mx[6]));
    return true;
}

/* makeDirectConnections (:1087-1163) with a Levenberg-Marquardt minimiser in place of Ceres BFGS, followed by the
 * re-trace of computePathLengthsTillClosestP2.  `rng` plays the Sampler (2 draws per start direction, 1 per failure). */
struct ConnectResult {
    bool success;
    float3 dir, rev; /* launch velocity (|dir| = n(p1)), unit reverse arrival direction */
    float opl, dist, weight, n1;
    int evals, steps; /* residual evaluations, Hessian-carrying leapfrog steps */
    ExitInfo exit;
    M3 J;        /* residual Jacobian at the accepted launch velocity x (as computefdf returns it) ... */
    float xnorm; /* ... and |x|: computefdf renormalises x to n(p1), so J scales like n(p1) / |x| */
};

template <bool SDFSHAPE, bool WANT_OPL>
static __device__ __forceinline__ void connect_solve(const MediumDev &M, int precision, float tol2, float rrweight, int maxIterations, float3 p1, float3 p2,
                                     float3 din, bool isSensor, bool refract, bool straightFirst, PathRng &rng, ConnectResult &R) {
    R.success = false;
    R.weight = 1.0f;
    R.opl = R.dist = 0.0f;
    R.n1 = 1.0f;
    R.dir = R.rev = f3(0.f, 0.f, 0.f);
    R.evals = R.steps = 0;
    R.exit.exited = R.exit.tir = false;
    R.exit.nb = R.exit.cosI = 1.0f;
    R.exit.tau = 0.0f;
    if (!rif_inside_limits(M.hasSdf ? M.sdf : M.rif, p1)) return;
    float RIFp;
    float3 g0;
    rif_lookup<MER_RIF_TRICUBIC>(M.rif, p1, RIFp, g0);
    R.n1 = RIFp;
    float x[3];
    bool converged = false;
    while (true) {
        float3 ax, ay;
        coordinate_system(din, ax, ay);
        if (straightFirst) { /* MER_START_STRAIGHT: the first guess is the seed direction itself ... */
            straightFirst = false;
            float3 g = din;
            if (refract && !M.hasSdf) {
                /* ... bent by Snell's law at the point where the straight line leaves an analytic container, as if the
                 * exterior direction were the seed direction: sin(theta_i) = sin(theta_seed) / n, never totally reflected */
                const float te = exit_distance(M, p1, din);
                const float3 pe = f3(p1.x + te * din.x, p1.y + te * din.y, p1.z + te * din.z);
                const float3 N = shape_normal(M, pe);
                const float c = dot3(din, N);
                if (c > 0.0f) {
                    const float inv = 1.0f / RIFp;
                    const float3 gt = f3((din.x - c * N.x) * inv, (din.y - c * N.y) * inv, (din.z - c * N.z) * inv);
                    const float gn = sqrtf(fmaxf(0.0f, 1.0f - dot3(gt, gt)));
                    g = f3(gt.x + gn * N.x, gt.y + gn * N.y, gt.z + gn * N.z);
                }
            }
            x[0] = g.x * RIFp; x[1] = g.y * RIFp; x[2] = g.z * RIFp;
        } else {
            const float u1 = rng.next(), u2 = rng.next();
            const float z = u1, tmp = sqrtf(fmaxf(0.0f, 1.0f - z * z)), phi = (float) (2.0 * 3.14159265358979323846 * (double) u2);
            float sp, cp;
            sincosf(phi, &sp, &cp);
            const float lx = cp * tmp, ly = sp * tmp;
            x[0] = (lx * ax.x + ly * ay.x + z * din.x) * RIFp;
            x[1] = (lx * ax.y + ly * ay.y + z * din.y) * RIFp;
            x[2] = (lx * ax.z + ly * ay.z + z * din.z) * RIFp;
        }
        float3 r;
        M3 Jt;
        int cnt = 0;
        /* evaluations that end degenerate or totally reflected carry no usable residual: infinite cost */
        int status = compute_fdf<SDFSHAPE>(M, precision, f3(x[0], x[1], x[2]), p1, p2, isSensor, refract, r, Jt, cnt);
        R.evals++;
        float cost = status >= 2 ? INFINITY : 0.5f * dot3(r, r), lambda = 0.0f;
        int accepted = 0;
        /* iterate to |r|^2 < tol2 / 4 so that the re-trace's |p - p2|^2 <= tol2 test (:1023-1027) is met with margin */
        for (int ev = 0; ev < 2 * maxIterations && accepted < maxIterations && !(cost < 0.125f * tol2) && cost < INFINITY; ev++) {
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
            status = compute_fdf<SDFSHAPE>(M, precision, f3(x[0] + dx[0], x[1] + dx[1], x[2] + dx[2]), p1, p2, isSensor, refract, rn, Jn, cnt);
            R.evals++;
            const float costn = status >= 2 ? INFINITY : 0.5f * dot3(rn, rn);
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
        R.steps += cnt;
        R.J = Jt;
        if (cost < tol2) { converged = true; break; }
        if (rng.next() < rrweight) R.weight = R.weight * (1.0f / rrweight);
        else break;
    }
    R.xnorm = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const float xl = 1.0f / R.xnorm;
    R.dir = f3((x[0] * xl) * RIFp, (x[1] * xl) * RIFp, (x[2] * xl) * RIFp);
    if (converged && compute_path_lengths<SDFSHAPE, WANT_OPL>(M, precision, tol2, p1, p2, R.dir, R.rev, isSensor, refract, R.opl, R.dist, R.exit)) R.success = true;
}

/* |d r_perp / d omega| of a solved connection: the area, in the plane perpendicular to the arriving ray, swept per unit
 * solid angle of launch direction at p1 (t^2 for a straight ray of length t).  The Jacobian of computefdf w.r.t. the
 * launch velocity has rank 2 (the radial direction is projected out, the image is perpendicular to the arrival
 * direction); the product of its two singular values is the Frobenius norm of its cofactor matrix, and a unit change of
 * direction is a change n(p1) of velocity; J taken at a launch velocity of length |x| is n(p1) / |x| times the one at
 * length n(p1), hence the factor |x|^2. */
static __device__ float connection_spread(const M3 &J, float xnorm) {
    const float *m = J.m;
    const float c[9] = {m[4] * m[8] - m[5] * m[7], m[5] * m[6] - m[3] * m[8], m[3] * m[7] - m[4] * m[6],
                        m[2] * m[7] - m[1] * m[8], m[0] * m[8] - m[2] * m[6], m[1] * m[6] - m[0] * m[7],
                        m[1] * m[5] - m[2] * m[4], m[2] * m[3] - m[0] * m[5], m[0] * m[4] - m[1] * m[3]};
    float s = 0.0f;
    for (int i = 0; i < 9; i++) s += c[i] * c[i];
    return xnorm * xnorm * sqrtf(s);
}

} /* namespace merc */
