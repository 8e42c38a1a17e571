/*
 * mer_device.cuh — device-side primitives of the eikonal hot path (sm_100a).
 *
 * Everything here is a from-scratch CUDA formulation of the reference's scalar C++
 * (citations relative to the MitsubaER tree):
 *   rif_lookup*      Spline<3>::valueAndGradient / gradient   include/mitsuba/core/basisspline.h:318-364, 438-471
 *                    + SplineDataSource wrappers              src/volume/splinevolume.cpp:319-360
 *   grid_lookup      GridDataSource::lookupFloat              src/volume/gridvolume.cpp:337-363
 *   er_step_fused    HeterogeneousRefractiveMedium::er_step   src/medium/heterogeneousrefractive.cpp:653-661
 *   inside_shape     insideShape / hackForSphere / hackForBox :707-726
 *   hg_sample/eval   HGPhaseFunction                          src/phase/hg.cpp:76-110
 *   Philox4x32-10    stands in for the per-thread Sampler     src/librender/renderjob.cpp:62-66
 *
 * Data layout in HBM:
 *   coeff   float  [z][y][x]   prefiltered B-spline coefficients, the reference's layout
 *   atlas   the same coefficients once more as a 2-D CUDA array (block-linear, texture path): layer z is tile
 *           (z & tileMask, z >> tileShift).  A 4x4x4 stencil is 16 texture gathers (tld4: one instruction returns
 *           a 2x2 block of texels at ANY alignment), so the table is 1x the grid and stays L2-resident up to
 *           ~300^3 (round 1 used an 8x-redundant table of x-tap tuples, coeff8: 512 MiB for 256^3, DRAM-bound)
 *   packed  float4 [z][y][x] = (n, dn/dx, dn/dy, dn/dz) sampled at the grid nodes (fast mode)
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mitsubaer_b200.h"

#define MER_EPSILON 1e-4f /* include/mitsuba/core/constants.h:28 */

struct RifDev {
    int mode;
    int N[3];
    float xmin[3], xres[3];
    float limLo[3], limHi[3];
    int hasXform;
    float M[12];
    const float *coeff;
    const float4 *packed;
    /* Texture storage of the cubic coefficients: a 2-D atlas of the z-layers in a CUDA array (tile (k & tileMask,
     * k >> tileShift) holds layer k), read with tld4 — one gather returns a 2x2 block of raw coefficients, 16 gathers a
     * 4x4x4 stencil, at any alignment. */
    unsigned long long tex;
    int tileShift, tileMask;
    /* Alternative storage for the load/store path: 2 x float4 per voxel = (c[x-1..x+2] of row y, c[x-1..x+2] of row y+1),
     * one aligned 32-byte sector.  A 4x4x4 stencil is 8 LDG.E.256.  null: the atlas is used. */
    const float4 *coeff8;
};

struct GridDev {
    int N[3];
    float G[12]; /* world -> grid, row-major 3x4 (gridvolume.cpp:190-195) */
    float aabbLo[3], aabbHi[3]; /* world AABB of the data box (gridvolume.cpp:199-201) */
    const float *data;
    int channels; /* 1: density (lookupFloat); 3: interleaved RGB albedo (lookupSpectrum, gridvolume.cpp:386-463) */
};

struct MediumDev {
    RifDev rif;
    RifDev sdf; /* <volume name="sdf"> (aggressive tracing only) */
    int hasSdf, aggressive;
    float maxSdfError;
    GridDev grid;
    int hasGrid;
    GridDev albedoGrid; /* <volume name="albedo"> (heterogeneous.cpp:262-268, looked up at :600 / :646) */
    int hasAlbedoGrid;
    float sigmaA[3], sigmaS[3], sigmaT[3];
    float h;
    float weight;          /* m_mediumSamplingWeight */
    int strategy;
    float samplingDensity; /* m_samplingDensity */
    int shapeType;
    int boundary; /* MER_BOUNDARY_*: what the container surface does to a ray */
    int physicalScaling; /* MER_SCALING_PHYSICAL: edges carry (n_start / n_end)^2 instead of the reference's refRatioSq */
    float minExit2; /* computefdfBDPT calls a connection degenerate when it leaves the shape within sqrt(minExit2) of p1 (:891: Epsilon) */
    float shape[6];
    float g;
    float densityScale, invMaxDensity;
    float albedo[3];
    /* strategy "maximum": MaxExpDist (src/medium/maxexp.h:29-58) built on the host by mer_medium_create — sigma_t sorted
     * in decreasing order, the discrete cdf of the pieces, where each piece starts, its lower integration bound */
    float mxSigma[3], mxCdf[4], mxStart[3], mxLower[3], mxNorm, mxInvNorm;
};

/* ------------------------------------------------------------------ small vector helpers */
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* ------------------------------------------------------------------ a1: B-spline kernels
 * basisspline.h:39-114.  `d` = x - index; the same piecewise polynomials, same branch
 * conditions.  (The reference evaluates kernel<1>'s centre branch through a double literal;
 * here it is single precision: <= 1 ulp apart.) */
__device__ __forceinline__ float bs_k0(float d) {
    float a = fabsf(d), t = 2.0f - a;
    float outer = 0.16666666666666666667f * t * t * t;
    float inner = 0.66666666666666666667f - a * a + 0.5f * a * a * a;
    float r = a > 1.0f ? outer : inner;
    return a > 2.0f ? 0.0f : r;
}
__device__ __forceinline__ float bs_k1(float d) {
    float a = fabsf(d), t = 2.0f - a;
    float s = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
    float outer = -0.5f * t * t;
    float inner = (1.5f * a - 2.0f) * a;
    float r = a > 1.0f ? outer : inner;
    return a > 2.0f ? 0.0f : s * r;
}

/* The four taps of a lookup sit at known distances: t0 in [1,2), t1 in [0,1), t2 in (-1,0], t3 in (-2,-1],
 * so each needs only ONE branch of the reference's piecewise kernel (same polynomial, same operation order
 * as bs_k0/bs_k1; at the shared end points both branches agree).  Halves the weight arithmetic. */
__device__ __forceinline__ void bs_weights(float x, float fx, float w0[4], float w1[4]) {
    const float d0 = x - (fx - 1.0f), d1 = x - fx, d2 = x - (fx + 1.0f), d3 = x - (fx + 2.0f);
    const float a0 = d0, a1 = d1, a2 = -d2, a3 = -d3; /* |d| */
    const float t0 = 2.0f - a0, t3 = 2.0f - a3;
    w0[0] = 0.16666666666666666667f * t0 * t0 * t0;
    w1[0] = -0.5f * t0 * t0;
    w0[1] = 0.66666666666666666667f - a1 * a1 + 0.5f * a1 * a1 * a1;
    w1[1] = (1.5f * a1 - 2.0f) * a1;
    w0[2] = 0.66666666666666666667f - a2 * a2 + 0.5f * a2 * a2 * a2;
    w1[2] = -((1.5f * a2 - 2.0f) * a2);
    w0[3] = 0.16666666666666666667f * t3 * t3 * t3;
    w1[3] = 0.5f * t3 * t3;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ float3 rif_to_volume(const RifDev &R, float3 p) {
    if (!R.hasXform) return p;
    return f3(R.M[0] * p.x + R.M[1] * p.y + R.M[2] * p.z + R.M[3], R.M[4] * p.x + R.M[5] * p.y + R.M[6] * p.z + R.M[7],
              R.M[8] * p.x + R.M[9] * p.y + R.M[10] * p.z + R.M[11]);
}
__device__ __forceinline__ float3 rif_rot_t(const RifDev &R, float3 g) { /* worldToVolume_RotT * g */
    if (!R.hasXform) return g;
    return f3(R.M[0] * g.x + R.M[4] * g.y + R.M[8] * g.z, R.M[1] * g.x + R.M[5] * g.y + R.M[9] * g.z,
              R.M[2] * g.x + R.M[6] * g.y + R.M[10] * g.z);
}

/* SplineDataSource::insideVolumeLimits, splinevolume.cpp:319-324 (strict inequalities) */
__device__ __forceinline__ bool rif_inside_limits(const RifDev &R, float3 pw) {
    float3 p = rif_to_volume(R, pw);
    return p.x > R.limLo[0] && p.x < R.limHi[0] && p.y > R.limLo[1] && p.y < R.limHi[1] && p.z > R.limLo[2] &&
           p.z < R.limHi[2];
}

/* ------------------------------------------------------------------ a4: tricubic value + gradient
 * One 4x4x4 stencil evaluation giving n and grad n (volume space, already scaled by dxres).
 * Separable contraction: 16 LDG.128 rows -> x (128 FMA) -> y (48 FMA) -> z (16 FMA).
 * Tap set = floor(x)-1 .. floor(x)+2 per axis, which is the reference's
 * ceil(x-2)..floor(x+2) range minus its zero-weight end taps.  Indices are clamped into the
 * grid (the reference reads out of bounds there). */
/* ---- stencil loads.  Interior cells (every tap inside the grid: always the case within insideVolumeLimits, whose margin
 * is two voxels) read the atlas: per layer four gathers, each the 2x2 block of texels around (u, v); component order
 * w,z,x,y = (i,j), (i+1,j), (i,j+1), (i+1,j+1) (checked against the coefficient array when the texture is built).
 * Anything else reads the linear array tap by tap with indices clamped into the grid (the reference reads out of bounds). */
__device__ __forceinline__ bool rif_cell_interior(const RifDev &R, int i0, int j0, int k0) {
    return (unsigned) (i0 - 1) <= (unsigned) (R.N[0] - 4) && (unsigned) (j0 - 1) <= (unsigned) (R.N[1] - 4) &&
           (unsigned) (k0 - 1) <= (unsigned) (R.N[2] - 4);
}
/* rows j0-1 .. j0+2 (x taps i0-1 .. i0+2 each) of layer k; (i0, j0) interior, k in the grid */
__device__ __forceinline__ void rif_slab_tex(const RifDev &R, int i0, int j0, int k, float4 c[4]) {
    const float u = (float) ((k & R.tileMask) * R.N[0] + i0), v = (float) ((k >> R.tileShift) * R.N[1] + j0);
    const cudaTextureObject_t t = (cudaTextureObject_t) R.tex;
    const float4 A = tex2Dgather<float4>(t, u, v, 0), B = tex2Dgather<float4>(t, u + 2.0f, v, 0);
    const float4 C = tex2Dgather<float4>(t, u, v + 2.0f, 0), D = tex2Dgather<float4>(t, u + 2.0f, v + 2.0f, 0);
    c[0] = make_float4(A.w, A.z, B.w, B.z);
    c[1] = make_float4(A.x, A.y, B.x, B.y);
    c[2] = make_float4(C.w, C.z, D.w, D.z);
    c[3] = make_float4(C.x, C.y, D.x, D.y);
}
/* one 32-byte sector per lane: LDG.E.256 (sm_100+), read-only path */
__device__ __forceinline__ void ldg256(const float4 *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}
/* the same four rows from the coeff8 table: two sectors */
__device__ __forceinline__ void rif_slab_lsu(const RifDev &R, int i0, int j0, int k, float4 c[4]) {
    const float4 *base = R.coeff8 + 2 * (((size_t) k * (size_t) R.N[1] + (size_t) (j0 - 1)) * (size_t) R.N[0] + (size_t) i0);
    ldg256(base, c[0], c[1]);
    ldg256(base + 4 * (size_t) R.N[0], c[2], c[3]);
}
/* LAYOUT: 0 atlas (texture gathers), 1 coeff8 (256-bit loads), -1 whichever the volume has (warp-uniform branch) */
template <int LAYOUT>
__device__ __forceinline__ void rif_slab_interior(const RifDev &R, int i0, int j0, int k, float4 c[4]) {
    if (LAYOUT == 1 || (LAYOUT < 0 && R.coeff8)) rif_slab_lsu(R, i0, j0, k, c);
    else rif_slab_tex(R, i0, j0, k, c);
}
__device__ __forceinline__ void rif_slab_clamped(const RifDev &R, int i0, int j0, int k, float4 c[4]) {
    const int N0 = R.N[0], N1 = R.N[1], N2 = R.N[2];
    const int xa = clampi(i0 - 1, 0, N0 - 1), xb = clampi(i0, 0, N0 - 1), xc = clampi(i0 + 1, 0, N0 - 1), xd = clampi(i0 + 2, 0, N0 - 1);
    const float *slab = R.coeff + (size_t) clampi(k, 0, N2 - 1) * (size_t) N0 * (size_t) N1;
#pragma unroll
    for (int dy = 0; dy < 4; dy++) {
        const float *row = slab + (size_t) clampi(j0 - 1 + dy, 0, N1 - 1) * (size_t) N0;
        c[dy] = make_float4(__ldg(row + xa), __ldg(row + xb), __ldg(row + xc), __ldg(row + xd));
    }
}

__device__ __forceinline__ void rif_tricubic(const RifDev &R, float3 pv, float &f, float3 &g) {
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1],
                z = (pv.z - R.xmin[2]) * R.xres[2];
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int i0 = (int) fx, j0 = (int) fy, k0 = (int) fz;
    float wx0[4], wx1[4], wy0[4], wy1[4], wz0[4], wz1[4];
    bs_weights(x, fx, wx0, wx1);
    bs_weights(y, fy, wy0, wy1);
    bs_weights(z, fz, wz0, wz1);
    float4 c[16];
    if (rif_cell_interior(R, i0, j0, k0)) {
        if (R.coeff8) {
#pragma unroll
            for (int dz = 0; dz < 4; dz++) rif_slab_lsu(R, i0, j0, k0 - 1 + dz, c + 4 * dz);
        } else {
#pragma unroll
            for (int dz = 0; dz < 4; dz++) rif_slab_tex(R, i0, j0, k0 - 1 + dz, c + 4 * dz);
        }
    } else {
#pragma unroll 1
        for (int dz = 0; dz < 4; dz++) rif_slab_clamped(R, i0, j0, k0 - 1 + dz, c + 4 * dz);
    }

    float accF = 0.f, accX = 0.f, accY = 0.f, accZ = 0.f;
#pragma unroll
    for (int dz = 0; dz < 4; dz++) {
        float b00 = 0.f, b10 = 0.f, b01 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            const float4 q = c[dz * 4 + dy];
            float a0 = q.x * wx0[0] + q.y * wx0[1] + q.z * wx0[2] + q.w * wx0[3];
            float a1 = q.x * wx1[0] + q.y * wx1[1] + q.z * wx1[2] + q.w * wx1[3];
            b00 = fmaf(a0, wy0[dy], b00);
            b10 = fmaf(a1, wy0[dy], b10);
            b01 = fmaf(a0, wy1[dy], b01);
        }
        accF = fmaf(b00, wz0[dz], accF);
        accX = fmaf(b10, wz0[dz], accX);
        accY = fmaf(b01, wz0[dz], accY);
        accZ = fmaf(b00, wz1[dz], accZ);
    }
    f = accF;
    g = f3(accX * R.xres[0], accY * R.xres[1], accZ * R.xres[2]);
}

/* ------------------------------------------------------------------ fast mode: packed trilinear
 * 8 aligned float4 {n, grad n} fetches, trilinear weights (gridvolume.cpp:337-363 arithmetic
 * applied to a 4-channel voxel).  Not the reference's interpolant: see SURVEY.md R1. */
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {
    return make_float4(fmaf(t, b.x - a.x, a.x), fmaf(t, b.y - a.y, a.y), fmaf(t, b.z - a.z, a.z),
                       fmaf(t, b.w - a.w, a.w));
}
__device__ __forceinline__ void rif_trilinear(const RifDev &R, float3 pv, float &f, float3 &g) {
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1],
                z = (pv.z - R.xmin[2]) * R.xres[2];
    const int N0 = R.N[0], N1 = R.N[1], N2 = R.N[2];
    const int i0 = clampi((int) floorf(x), 0, N0 - 2), j0 = clampi((int) floorf(y), 0, N1 - 2),
              k0 = clampi((int) floorf(z), 0, N2 - 2);
    const float tx = x - (float) i0, ty = y - (float) j0, tz = z - (float) k0;
    const float4 *b = R.packed + ((size_t) k0 * N1 + j0) * (size_t) N0 + i0;
    const size_t sy = N0, sz = (size_t) N0 * N1;
    float4 c000 = __ldg(b), c001 = __ldg(b + 1), c010 = __ldg(b + sy), c011 = __ldg(b + sy + 1),
           c100 = __ldg(b + sz), c101 = __ldg(b + sz + 1), c110 = __ldg(b + sz + sy), c111 = __ldg(b + sz + sy + 1);
    float4 r = lerp4(lerp4(lerp4(c000, c001, tx), lerp4(c010, c011, tx), ty),
                     lerp4(lerp4(c100, c101, tx), lerp4(c110, c111, tx), ty), tz);
    f = r.x;
    g = f3(r.y, r.z, r.w);
}

/* ------------------------------------------------------------------ stencil cache
 * A ray advances h << voxel pitch per step, so consecutive lookups mostly hit the SAME 4x4x4
 * coefficient block.  The block is kept in registers (16 float4) and re-fetched only when the cell
 * index changes: the L2->L1 traffic of a step drops from 512 B to 512 B x P(cell change) (ncu r01b:
 * the uncached kernel was bound by the L1/L2 pipe at a 10 % L1 hit rate, not by issue). */
/* Where the cached block lives.  Registers (MER_STENCIL_SMEM=0): no extra instructions, 168 regs -> 12
 * warps/SM.  Shared memory (MER_STENCIL_SMEM=1): [row][thread] float4, conflict-free LDS.128, ~100 regs ->
 * 20 warps/SM at the price of 16 LDS per step. */
#ifndef MER_STENCIL_SMEM
#define MER_STENCIL_SMEM 0
#endif
#define MER_STENCIL_BLOCK 128 /* threads per CTA of every kernel that uses a stencil cache */

template <int ROWS> struct StencilStore {
#if MER_STENCIL_SMEM
    float4 *base; /* &smem[threadIdx.x]; row r at base[r * MER_STENCIL_BLOCK] */
    __device__ __forceinline__ float4 get(int r) const { return base[r * MER_STENCIL_BLOCK]; }
    __device__ __forceinline__ void set(int r, float4 v) { base[r * MER_STENCIL_BLOCK] = v; }
    __device__ __forceinline__ void bind() {
        __shared__ float4 stencilSmem[ROWS * MER_STENCIL_BLOCK];
        base = stencilSmem + threadIdx.x;
    }
#else
    float4 c[ROWS];
    __device__ __forceinline__ float4 get(int r) const { return c[r]; }
    __device__ __forceinline__ void set(int r, float4 v) { c[r] = v; }
    __device__ __forceinline__ void bind() {}
#endif
};

template <int MODE> struct StencilCache;
template <> struct StencilCache<MER_RIF_TRICUBIC> : StencilStore<16> {
    int i, j, k;
    __device__ __forceinline__ void invalidate() { bind(); i = j = k = -0x7fffffff; }
    /* invalidate AND overwrite: tells the register allocator that the block is dead (used across a call) */
    __device__ __forceinline__ void clear() {
        i = j = k = -0x7fffffff;
#pragma unroll
        for (int r = 0; r < 16; r++) set(r, make_float4(0.f, 0.f, 0.f, 0.f));
    }
};
template <> struct StencilCache<MER_RIF_TRILINEAR_PACKED> : StencilStore<8> {
    int i, j, k;
    __device__ __forceinline__ void invalidate() { bind(); i = j = k = -0x7fffffff; }
    __device__ __forceinline__ void clear() {
        i = j = k = -0x7fffffff;
#pragma unroll
        for (int r = 0; r < 8; r++) set(r, make_float4(0.f, 0.f, 0.f, 0.f));
    }
};

/* load of the block of an INTERIOR cell (every tap inside the grid) */
template <int LAYOUT = -1>
__device__ __forceinline__ void rif_fetch_interior(const RifDev &R, StencilCache<MER_RIF_TRICUBIC> &S, int i0, int j0, int k0) {
#pragma unroll
    for (int dz = 0; dz < 4; dz++) {
        float4 c[4];
        rif_slab_interior<LAYOUT>(R, i0, j0, k0 - 1 + dz, c);
        S.set(dz * 4 + 0, c[0]); S.set(dz * 4 + 1, c[1]); S.set(dz * 4 + 2, c[2]); S.set(dz * 4 + 3, c[3]);
    }
    S.i = i0; S.j = j0; S.k = k0;
}
/* unconditional (re)load of the block of cell (i0, j0, k0), whatever the cell */
__device__ __forceinline__ void rif_fetch(const RifDev &R, StencilCache<MER_RIF_TRICUBIC> &S, int i0, int j0, int k0) {
    if (rif_cell_interior(R, i0, j0, k0)) {
        if (R.coeff8) rif_fetch_interior<1>(R, S, i0, j0, k0);
        else rif_fetch_interior<0>(R, S, i0, j0, k0);
    } else {
#pragma unroll
        for (int dz = 0; dz < 4; dz++) {
            float4 c[4];
            rif_slab_clamped(R, i0, j0, k0 - 1 + dz, c);
            S.set(dz * 4 + 0, c[0]); S.set(dz * 4 + 1, c[1]); S.set(dz * 4 + 2, c[2]); S.set(dz * 4 + 3, c[3]);
        }
        S.i = i0; S.j = j0; S.k = k0;
    }
}
__device__ __forceinline__ void stencil_ensure(const RifDev &R, StencilCache<MER_RIF_TRICUBIC> &S, int i0, int j0, int k0) {
    if (i0 != S.i || j0 != S.j || k0 != S.k) rif_fetch(R, S, i0, j0, k0);
}

/* sm_100's packed FP32 (fma.rn.f32x2: two FMAs on a 64-bit register pair in ONE issue slot; measured at the scalar FLOP
 * rate, tools/microbench.cu, so it saves issue slots, not pipe cycles).  The separable contraction is linear, so it can run
 * on pairs (x-taps {0,1} and {2,3} of a row side by side, which is how the 256-bit stencil loads leave them in registers)
 * and fold the two halves once at the very end: 7 packed instructions per stencil row instead of 11 scalar ones. */
/* MEASURED AND REJECTED (default off): k_trace shrinks from 864 to 808 instructions and every parity test passes, but C2
 * (39.6 vs 40.0-40.7 M samples/s) and the C4 sweep are unchanged: an FFMA2 holds the FMA pipe for two cycles, and the step
 * body is bound by the pipe plus dependent-issue latency, not by issue slots. */
#ifndef MER_FFMA2
#define MER_FFMA2 0
#endif
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fold2(f32x2_t a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return lo + hi; }

__device__ __forceinline__ void rif_tricubic_cached(const RifDev &R, float3 pv, StencilCache<MER_RIF_TRICUBIC> &S,
                                                    float &f, float3 &g) {
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1],
                z = (pv.z - R.xmin[2]) * R.xres[2];
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    stencil_ensure(R, S, (int) fx, (int) fy, (int) fz);
    float wx0[4], wx1[4], wy0[4], wy1[4], wz0[4], wz1[4];
    bs_weights(x, fx, wx0, wx1);
    bs_weights(y, fy, wy0, wy1);
    bs_weights(z, fz, wz0, wz1);
#if MER_FFMA2
    {
        const f32x2_t X0a = pk2(wx0[0], wx0[1]), X0b = pk2(wx0[2], wx0[3]), X1a = pk2(wx1[0], wx1[1]), X1b = pk2(wx1[2], wx1[3]);
        f32x2_t accF = 0ull, accX = 0ull, accY = 0ull, accZ = 0ull; /* (+0, +0) */
#pragma unroll
        for (int dz = 0; dz < 4; dz++) {
            f32x2_t b00 = 0ull, b10 = 0ull, b01 = 0ull;
#pragma unroll
            for (int dy = 0; dy < 4; dy++) {
                const float4 q = S.get(dz * 4 + dy);
                const f32x2_t qa = pk2(q.x, q.y), qb = pk2(q.z, q.w);
                const f32x2_t a0 = fma2(qb, X0b, mul2(qa, X0a)), a1 = fma2(qb, X1b, mul2(qa, X1a));
                const f32x2_t Y0 = pk2(wy0[dy], wy0[dy]), Y1 = pk2(wy1[dy], wy1[dy]);
                b00 = fma2(a0, Y0, b00);
                b10 = fma2(a1, Y0, b10);
                b01 = fma2(a0, Y1, b01);
            }
            const f32x2_t Z0 = pk2(wz0[dz], wz0[dz]), Z1 = pk2(wz1[dz], wz1[dz]);
            accF = fma2(b00, Z0, accF);
            accX = fma2(b10, Z0, accX);
            accY = fma2(b01, Z0, accY);
            accZ = fma2(b00, Z1, accZ);
        }
        f = fold2(accF);
        g = f3(fold2(accX) * R.xres[0], fold2(accY) * R.xres[1], fold2(accZ) * R.xres[2]);
        return;
    }
#endif
    float accF = 0.f, accX = 0.f, accY = 0.f, accZ = 0.f;
#pragma unroll
    for (int dz = 0; dz < 4; dz++) {
        float b00 = 0.f, b10 = 0.f, b01 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            const float4 q = S.get(dz * 4 + dy);
            float a0 = q.x * wx0[0] + q.y * wx0[1] + q.z * wx0[2] + q.w * wx0[3];
            float a1 = q.x * wx1[0] + q.y * wx1[1] + q.z * wx1[2] + q.w * wx1[3];
            b00 = fmaf(a0, wy0[dy], b00);
            b10 = fmaf(a1, wy0[dy], b10);
            b01 = fmaf(a0, wy1[dy], b01);
        }
        accF = fmaf(b00, wz0[dz], accF);
        accX = fmaf(b10, wz0[dz], accX);
        accY = fmaf(b01, wz0[dz], accY);
        accZ = fmaf(b00, wz1[dz], accZ);
    }
    f = accF;
    g = f3(accX * R.xres[0], accY * R.xres[1], accZ * R.xres[2]);
}

__device__ __forceinline__ void stencil_ensure(const RifDev &R, StencilCache<MER_RIF_TRILINEAR_PACKED> &S, int i0, int j0, int k0) {
    if (i0 != S.i || j0 != S.j || k0 != S.k) {
        const int N0 = R.N[0], N1 = R.N[1];
        const float4 *b = R.packed + ((size_t) k0 * N1 + j0) * (size_t) N0 + i0;
        const size_t sy = N0, sz = (size_t) N0 * N1;
        S.set(0, __ldg(b)); S.set(1, __ldg(b + 1)); S.set(2, __ldg(b + sy)); S.set(3, __ldg(b + sy + 1));
        S.set(4, __ldg(b + sz)); S.set(5, __ldg(b + sz + 1)); S.set(6, __ldg(b + sz + sy)); S.set(7, __ldg(b + sz + sy + 1));
        S.i = i0; S.j = j0; S.k = k0;
    }
}

__device__ __forceinline__ void rif_trilinear_cached(const RifDev &R, float3 pv, StencilCache<MER_RIF_TRILINEAR_PACKED> &S,
                                                     float &f, float3 &g) {
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1],
                z = (pv.z - R.xmin[2]) * R.xres[2];
    const int N0 = R.N[0], N1 = R.N[1], N2 = R.N[2];
    const int i0 = clampi((int) floorf(x), 0, N0 - 2), j0 = clampi((int) floorf(y), 0, N1 - 2),
              k0 = clampi((int) floorf(z), 0, N2 - 2);
    stencil_ensure(R, S, i0, j0, k0);
    const float tx = x - (float) i0, ty = y - (float) j0, tz = z - (float) k0;
    float4 r = lerp4(lerp4(lerp4(S.get(0), S.get(1), tx), lerp4(S.get(2), S.get(3), tx), ty),
                     lerp4(lerp4(S.get(4), S.get(5), tx), lerp4(S.get(6), S.get(7), tx), ty), tz);
    f = r.x;
    g = f3(r.y, r.z, r.w);
}

/* Speculative early refetch.  The dependency chain of a step is contraction_k -> p_{k+1} -> loads_{k+1} ->
 * contraction_{k+1}, so a cell change costs a full L2/HBM round trip with nothing to overlap (ncu r01:
 * long_scoreboard = 33 % of stall samples).  p_{k+1} is predictable to O(h^2 |grad n|) by linear
 * extrapolation as soon as contraction_k has consumed the cached block, so the block of the PREDICTED next
 * cell is fetched into the cache registers right there — a whole kick / containment test / loop turn /
 * drift / weight evaluation before it is needed.  The cache is keyed by the cell index, so a wrong
 * prediction only costs the ordinary refetch. */
#ifndef MER_SPECULATE
#define MER_SPECULATE 1
#endif
template <int MODE> __device__ __forceinline__ void rif_speculate(const RifDev &R, float3 pwNext, StencilCache<MODE> &S) {
#if MER_SPECULATE
    const float3 pv = rif_to_volume(R, pwNext);
    const float x = (pv.x - R.xmin[0]) * R.xres[0], y = (pv.y - R.xmin[1]) * R.xres[1], z = (pv.z - R.xmin[2]) * R.xres[2];
    if (MODE == MER_RIF_TRICUBIC) {
        stencil_ensure(R, S, (int) floorf(x), (int) floorf(y), (int) floorf(z));
    } else {
        stencil_ensure(R, S, clampi((int) floorf(x), 0, R.N[0] - 2), clampi((int) floorf(y), 0, R.N[1] - 2),
                       clampi((int) floorf(z), 0, R.N[2] - 2));
    }
#endif
}

template <int MODE>
__device__ __forceinline__ void rif_lookup_cached(const RifDev &R, float3 pw, StencilCache<MODE> &S, float &n, float3 &G);
template <>
__device__ __forceinline__ void rif_lookup_cached<MER_RIF_TRICUBIC>(const RifDev &R, float3 pw,
                                                                     StencilCache<MER_RIF_TRICUBIC> &S, float &n, float3 &G) {
    rif_tricubic_cached(R, rif_to_volume(R, pw), S, n, G);
    G = rif_rot_t(R, G);
}
template <>
__device__ __forceinline__ void rif_lookup_cached<MER_RIF_TRILINEAR_PACKED>(const RifDev &R, float3 pw,
                                                                             StencilCache<MER_RIF_TRILINEAR_PACKED> &S,
                                                                             float &n, float3 &G) {
    rif_trilinear_cached(R, rif_to_volume(R, pw), S, n, G);
    G = rif_rot_t(R, G);
}

/* ------------------------------------------------------------------ split lookup: cell -> fetch -> contract
 * The wavefront stepper (mer_render.cu) software-pipelines a step: the drift of step k+1 is applied right after the
 * second kick of step k, so the EXACT next cell is known a whole loop turn before its contraction and the block is
 * requested there (one fetch site, no prediction).  That needs the lookup in three pieces.  Same arithmetic as
 * rif_tricubic_cached / rif_trilinear_cached, operation for operation. */
struct CellPos {
    float x, y, z; /* continuous grid coordinates */
    int i, j, k;   /* cell index = key of the cached block */
};

/* XFORM = false: the caller knows the volume has no toWorld transform (the kernel variant is chosen on the host), so the
 * twelve matrix entries are not even loaded */
template <int MODE, bool XFORM = true> __device__ __forceinline__ CellPos rif_cell(const RifDev &R, float3 pw) {
    const float3 pv = XFORM ? rif_to_volume(R, pw) : pw;
    CellPos c;
    c.x = (pv.x - R.xmin[0]) * R.xres[0];
    c.y = (pv.y - R.xmin[1]) * R.xres[1];
    c.z = (pv.z - R.xmin[2]) * R.xres[2];
    if (MODE == MER_RIF_TRICUBIC) {
        c.i = (int) floorf(c.x); c.j = (int) floorf(c.y); c.k = (int) floorf(c.z);
    } else {
        c.i = clampi((int) floorf(c.x), 0, R.N[0] - 2);
        c.j = clampi((int) floorf(c.y), 0, R.N[1] - 2);
        c.k = clampi((int) floorf(c.z), 0, R.N[2] - 2);
    }
    return c;
}

/* can the hot loop fetch this cell's block with its branch-free loads?  Tricubic: every tap inside the grid (texture
 * gathers); packed trilinear: always (rif_cell clamps the cell into the grid) */
template <int MODE> __device__ __forceinline__ bool rif_cell_fast(const RifDev &R, const CellPos &c) {
    return MODE == MER_RIF_TRICUBIC ? rif_cell_interior(R, c.i, c.j, c.k) : true;
}

template <int MODE> __device__ __forceinline__ bool stencil_has(const StencilCache<MODE> &S, const CellPos &c) {
    return c.i == S.i && c.j == S.j && c.k == S.k;
}

__device__ __forceinline__ void rif_fetch(const RifDev &R, StencilCache<MER_RIF_TRILINEAR_PACKED> &S, int i0, int j0, int k0) {
    S.i = -0x7fffffff;
    stencil_ensure(R, S, i0, j0, k0);
}
template <int LAYOUT = -1>
__device__ __forceinline__ void rif_fetch_interior(const RifDev &R, StencilCache<MER_RIF_TRILINEAR_PACKED> &S, int i0, int j0, int k0) {
    rif_fetch(R, S, i0, j0, k0); /* indices are clamped into the grid by rif_cell */
}

/* contraction of the cached block at grid coordinates c; gradient in world space.  AFTER_X runs between the x stage
 * (the last reader of the 64 cached coefficients) and the y/z stages: the stepper requests the block of the predicted
 * next cell there, into the same registers, with the rest of the step still to overlap the loads. */
template <bool XFORM = true, typename AfterX>
__device__ __forceinline__ void rif_contract(const RifDev &R, StencilCache<MER_RIF_TRICUBIC> &S, const CellPos &c, float &f, float3 &G, AfterX afterX) {
    float wx0[4], wx1[4], wy0[4], wy1[4], wz0[4], wz1[4];
    bs_weights(c.x, floorf(c.x), wx0, wx1);
    bs_weights(c.y, floorf(c.y), wy0, wy1);
    bs_weights(c.z, floorf(c.z), wz0, wz1);
    float a0[16], a1[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const float4 q = S.get(r);
        a0[r] = q.x * wx0[0] + q.y * wx0[1] + q.z * wx0[2] + q.w * wx0[3];
        a1[r] = q.x * wx1[0] + q.y * wx1[1] + q.z * wx1[2] + q.w * wx1[3];
    }
    afterX(S);
    float accF = 0.f, accX = 0.f, accY = 0.f, accZ = 0.f;
#pragma unroll
    for (int dz = 0; dz < 4; dz++) {
        float b00 = 0.f, b10 = 0.f, b01 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 4; dy++) {
            b00 = fmaf(a0[dz * 4 + dy], wy0[dy], b00);
            b10 = fmaf(a1[dz * 4 + dy], wy0[dy], b10);
            b01 = fmaf(a0[dz * 4 + dy], wy1[dy], b01);
        }
        accF = fmaf(b00, wz0[dz], accF);
        accX = fmaf(b10, wz0[dz], accX);
        accY = fmaf(b01, wz0[dz], accY);
        accZ = fmaf(b00, wz1[dz], accZ);
    }
    f = accF;
    G = f3(accX * R.xres[0], accY * R.xres[1], accZ * R.xres[2]);
    if (XFORM) G = rif_rot_t(R, G);
}
template <bool XFORM = true, typename AfterX>
__device__ __forceinline__ void rif_contract(const RifDev &R, StencilCache<MER_RIF_TRILINEAR_PACKED> &S, const CellPos &c, float &f, float3 &G, AfterX afterX) {
    const float tx = c.x - (float) c.i, ty = c.y - (float) c.j, tz = c.z - (float) c.k;
    const float4 r0 = lerp4(S.get(0), S.get(1), tx), r1 = lerp4(S.get(2), S.get(3), tx), r2 = lerp4(S.get(4), S.get(5), tx), r3 = lerp4(S.get(6), S.get(7), tx);
    afterX(S);
    const float4 r = lerp4(lerp4(r0, r1, ty), lerp4(r2, r3, ty), tz);
    f = r.x;
    G = f3(r.y, r.z, r.w);
    if (XFORM) G = rif_rot_t(R, G);
}

/* SplineDataSource::valueAndGradient (splinevolume.cpp:352-358) in the handle's mode, world space */
template <int MODE> __device__ __forceinline__ void rif_lookup(const RifDev &R, float3 pw, float &n, float3 &G) {
    float3 pv = rif_to_volume(R, pw);
    if (MODE == MER_RIF_TRICUBIC)
        rif_tricubic(R, pv, n, G);
    else
        rif_trilinear(R, pv, n, G);
    G = rif_rot_t(R, G);
}

/* ------------------------------------------------------------------ a18: density lookup */
__device__ __forceinline__ float grid_lookup(const GridDev &D, float3 pw) {
    /* Transform::transformAffine: ((m0*x + m1*y) + m2*z) + m3, every operation rounded */
    const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(D.G[0], pw.x), __fmul_rn(D.G[1], pw.y)), __fmul_rn(D.G[2], pw.z)), D.G[3]);
    const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(D.G[4], pw.x), __fmul_rn(D.G[5], pw.y)), __fmul_rn(D.G[6], pw.z)), D.G[7]);
    const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(D.G[8], pw.x), __fmul_rn(D.G[9], pw.y)), __fmul_rn(D.G[10], pw.z)), D.G[11]);
    const int x1 = (int) floorf(px), y1 = (int) floorf(py), z1 = (int) floorf(pz);
    const int x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
    if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= D.N[0] || y2 >= D.N[1] || z2 >= D.N[2]) return 0.0f;
    const float fx = px - (float) x1, fy = py - (float) y1, fz = pz - (float) z1;
    const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
    const size_t rx = D.N[0], ry = D.N[1];
    const float *fd = D.data;
    const float d000 = __ldg(fd + ((size_t) z1 * ry + y1) * rx + x1), d001 = __ldg(fd + ((size_t) z1 * ry + y1) * rx + x2),
                d010 = __ldg(fd + ((size_t) z1 * ry + y2) * rx + x1), d011 = __ldg(fd + ((size_t) z1 * ry + y2) * rx + x2),
                d100 = __ldg(fd + ((size_t) z2 * ry + y1) * rx + x1), d101 = __ldg(fd + ((size_t) z2 * ry + y1) * rx + x2),
                d110 = __ldg(fd + ((size_t) z2 * ry + y2) * rx + x1), d111 = __ldg(fd + ((size_t) z2 * ry + y2) * rx + x2);
    /* same association as the reference, products and sums individually rounded */
    const float a0 = __fadd_rn(__fmul_rn(d000, gx), __fmul_rn(d001, fx)), a1 = __fadd_rn(__fmul_rn(d010, gx), __fmul_rn(d011, fx)),
                a2 = __fadd_rn(__fmul_rn(d100, gx), __fmul_rn(d101, fx)), a3 = __fadd_rn(__fmul_rn(d110, gx), __fmul_rn(d111, fx));
    const float b0 = __fadd_rn(__fmul_rn(a0, gy), __fmul_rn(a1, fy)), b1 = __fadd_rn(__fmul_rn(a2, gy), __fmul_rn(a3, fy));
    return __fadd_rn(__fmul_rn(b0, gz), __fmul_rn(b1, fz));
}

/* GridDataSource::lookupSpectrum, 3-channel grids (gridvolume.cpp:386-463): the float3 operators of the reference
 * act per channel, so each channel repeats lookupFloat's expression on the interleaved data */
__device__ __forceinline__ void grid_lookup3(const GridDev &D, float3 pw, float out[3]) {
    const float px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(D.G[0], pw.x), __fmul_rn(D.G[1], pw.y)), __fmul_rn(D.G[2], pw.z)), D.G[3]);
    const float py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(D.G[4], pw.x), __fmul_rn(D.G[5], pw.y)), __fmul_rn(D.G[6], pw.z)), D.G[7]);
    const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(D.G[8], pw.x), __fmul_rn(D.G[9], pw.y)), __fmul_rn(D.G[10], pw.z)), D.G[11]);
    const int x1 = (int) floorf(px), y1 = (int) floorf(py), z1 = (int) floorf(pz);
    const int x2 = x1 + 1, y2 = y1 + 1, z2 = z1 + 1;
    out[0] = out[1] = out[2] = 0.0f;
    if (x1 < 0 || y1 < 0 || z1 < 0 || x2 >= D.N[0] || y2 >= D.N[1] || z2 >= D.N[2]) return;
    const float fx = px - (float) x1, fy = py - (float) y1, fz = pz - (float) z1;
    const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
    const size_t rx = D.N[0], ry = D.N[1];
    const float *r00 = D.data + 3 * (((size_t) z1 * ry + y1) * rx + x1), *r01 = D.data + 3 * (((size_t) z1 * ry + y2) * rx + x1),
                *r10 = D.data + 3 * (((size_t) z2 * ry + y1) * rx + x1), *r11 = D.data + 3 * (((size_t) z2 * ry + y2) * rx + x1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        /* x2 = x1 + 1 is the next texel of the row: 3 floats further */
        const float a0 = __fadd_rn(__fmul_rn(__ldg(r00 + c), gx), __fmul_rn(__ldg(r00 + 3 + c), fx)),
                    a1 = __fadd_rn(__fmul_rn(__ldg(r01 + c), gx), __fmul_rn(__ldg(r01 + 3 + c), fx)),
                    a2 = __fadd_rn(__fmul_rn(__ldg(r10 + c), gx), __fmul_rn(__ldg(r10 + 3 + c), fx)),
                    a3 = __fadd_rn(__fmul_rn(__ldg(r11 + c), gx), __fmul_rn(__ldg(r11 + 3 + c), fx));
        const float b0 = __fadd_rn(__fmul_rn(a0, gy), __fmul_rn(a1, fy)), b1 = __fadd_rn(__fmul_rn(a2, gy), __fmul_rn(a3, fy));
        out[c] = __fadd_rn(__fmul_rn(b0, gz), __fmul_rn(b1, fz));
    }
}

/* ------------------------------------------------------------------ a11: containment */
/* value of the signed-distance spline; outlined: it is off the hot path (see inside_shape_lazy) */
/* (a compact loop over the scalar coefficient array: cicc 12.9 crashes on an outlined function around rif_tricubic's
 * unrolled 256-bit loads, and 64 scalar loads are fine for a lookup that is taken near the surface only) */
template <int UNUSED>
__device__ __noinline__ float sdf_value_outlined(const RifDev &S, float3 p) {
    const float3 pv = rif_to_volume(S, p);
    const float x = (pv.x - S.xmin[0]) * S.xres[0], y = (pv.y - S.xmin[1]) * S.xres[1], z = (pv.z - S.xmin[2]) * S.xres[2];
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const int i0 = (int) fx, j0 = (int) fy, k0 = (int) fz;
    float wx0[4], wx1[4], wy0[4], wy1[4], wz0[4], wz1[4];
    bs_weights(x, fx, wx0, wx1);
    bs_weights(y, fy, wy0, wy1);
    bs_weights(z, fz, wz0, wz1);
    const int N0 = S.N[0], N1 = S.N[1], N2 = S.N[2];
    float acc = 0.0f;
#pragma unroll 1
    for (int dz = 0; dz < 4; dz++) {
        const size_t slab = (size_t) clampi(k0 - 1 + dz, 0, N2 - 1) * (size_t) N1;
        float b = 0.0f;
#pragma unroll 1
        for (int dy = 0; dy < 4; dy++) {
            const float *row = S.coeff + (slab + (size_t) clampi(j0 - 1 + dy, 0, N1 - 1)) * (size_t) N0;
            const float a = row[clampi(i0 - 1, 0, N0 - 1)] * wx0[0] + row[clampi(i0, 0, N0 - 1)] * wx0[1] +
                            row[clampi(i0 + 1, 0, N0 - 1)] * wx0[2] + row[clampi(i0 + 2, 0, N0 - 1)] * wx0[3];
            b = fmaf(a, wy0[dy], b);
        }
        acc = fmaf(b, wz0[dz], acc);
    }
    return acc;
}
__device__ __forceinline__ float sdf_value(const RifDev &S, float3 p) { return sdf_value_outlined<0>(S, p); }

/* the analytic containers only: this is what the steppers' hot loops call (MER_SHAPE_SDF goes through inside_shape_any /
 * inside_shape_lazy, selected at compile time where the loop is hot) */
__device__ __forceinline__ bool inside_shape(const MediumDev &M, float3 p) {
    if (M.shapeType == MER_SHAPE_SPHERE) {
        float dx = p.x - M.shape[0], dy = p.y - M.shape[1], dz = p.z - M.shape[2];
        return (dx * dx + dy * dy + dz * dz) < M.shape[3] * M.shape[3];
    }
    return p.x >= M.shape[0] && p.x <= M.shape[3] && p.y >= M.shape[1] && p.y <= M.shape[4] && p.z >= M.shape[2] &&
           p.z <= M.shape[5];
}

/* The stepper's containment test for MER_SHAPE_SDF without a 64-tap lookup per step: |sdf| bounds the distance to the
 * surface from below (up to the interpolation error maxSdfError, as aggressive_trace assumes, :476-493), so after a
 * lookup that found the surface `safe` away the next lookups can wait until the ray has moved that far. */
__device__ __forceinline__ bool inside_shape_any(const MediumDev &M, float3 p) {
    if (M.shapeType == MER_SHAPE_SDF) return sdf_value(M.sdf, p) < 0.0f;
    return inside_shape(M, p);
}
__device__ __forceinline__ bool inside_shape_lazy(const MediumDev &M, float3 p, float moved, float &safe) {
    if (safe > moved) { safe -= moved; return true; }
    const float v = sdf_value(M.sdf, p);
    safe = -v - M.maxSdfError;
    return v < 0.0f;
}

/* ------------------------------------------------------------------ a7: leapfrog step
 * er_step (:653-661) with the two spline evaluations fused into one: the reference's
 * `gradient(p_new)` at the end of a step and `valueAndGradient(p)` at the start of the next
 * are evaluated at the same point, so (n, G) is carried across steps.  On entry (n, G) must be
 * the field at p; on exit they are the field at the new p.  `v/n` multiplies by 1/n as
 * TVector3::operator/ does. */
template <int MODE, bool SPECULATE = true>
__device__ __forceinline__ void er_step_fused(const RifDev &R, StencilCache<MODE> &S, float3 &p, float3 &v, float &n,
                                              float3 &G, float h, float &opl) {
    /* The leapfrog arithmetic is rounded operation by operation exactly like the reference's float
     * build (no fma contraction): on straight stretches the increment h*v/n is the same every step,
     * so a contracted multiply-add would turn a half-ulp rounding difference into a systematic
     * one-ulp-per-step drift against the reference.  Costs 9 extra instructions per step. */
    const float hs = __fmul_rn(0.5f, h);
    v = f3(__fadd_rn(v.x, __fmul_rn(hs, G.x)), __fadd_rn(v.y, __fmul_rn(hs, G.y)), __fadd_rn(v.z, __fmul_rn(hs, G.z)));
    const float recip = __frcp_rn(n);
    const float3 pOld = p;
    p = f3(__fadd_rn(p.x, __fmul_rn(__fmul_rn(h, v.x), recip)), __fadd_rn(p.y, __fmul_rn(__fmul_rn(h, v.y), recip)),
           __fadd_rn(p.z, __fmul_rn(__fmul_rn(h, v.z), recip)));
    opl = __fadd_rn(opl, __fmul_rn(h, n));
    rif_lookup_cached<MODE>(R, p, S, n, G);
    if (SPECULATE) rif_speculate<MODE>(R, f3(p.x + (p.x - pOld.x), p.y + (p.y - pOld.y), p.z + (p.z - pOld.z)), S);
    v = f3(__fadd_rn(v.x, __fmul_rn(hs, G.x)), __fadd_rn(v.y, __fmul_rn(hs, G.y)), __fadd_rn(v.z, __fmul_rn(hs, G.z)));
}

/* trace()'s decomposition of a distance (:673-675): bit-exact single-precision ops so the
 * step count matches the reference for every input. */
__device__ __forceinline__ void trace_split(float dist, float h, int &steps, float &rem) {
    float q = __fdiv_rn(dist, h);
    /* int conversion of huge / infinite quotients is undefined in C++; clamp */
    steps = q >= 2.0e9f ? 2000000000 : (int) q;
    rem = __fsub_rn(dist, __fmul_rn((float) steps, h));
}

/* ------------------------------------------------------------------ a15-a17: Henyey-Greenstein */
/* Every operation is individually rounded (no fma contraction), like the reference's x86 build:
 * cosTheta for |g| -> 1 and sinTheta = sqrt(1 - cos^2) are cancellation-prone, so contraction
 * would change the sampled direction by ~1e-5. */
__device__ __forceinline__ float hg_eval_dev(float g, float3 wi, float3 wo) {
    const float INV_FOURPI = 0.07957747154594766788f;
    const float gg = __fmul_rn(g, g);
    const float dotp = __fadd_rn(__fadd_rn(__fmul_rn(wi.x, wo.x), __fmul_rn(wi.y, wo.y)), __fmul_rn(wi.z, wo.z));
    const float temp = __fadd_rn(__fadd_rn(1.0f, gg), __fmul_rn(__fmul_rn(2.0f, g), dotp));
    return __fdiv_rn(__fmul_rn(INV_FOURPI, __fsub_rn(1.0f, gg)), __fmul_rn(temp, __fsqrt_rn(temp)));
}
__device__ __forceinline__ void coordinate_system(float3 a, float3 &b, float3 &c) { /* util.cpp:606-615 */
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.z, a.z))));
        c = f3(__fmul_rn(a.z, invLen), 0.0f, __fmul_rn(-a.x, invLen));
    } else {
        float invLen = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fmul_rn(a.y, a.y), __fmul_rn(a.z, a.z))));
        c = f3(0.0f, __fmul_rn(a.z, invLen), __fmul_rn(-a.y, invLen));
    }
    b = f3(__fsub_rn(__fmul_rn(c.y, a.z), __fmul_rn(c.z, a.y)), __fsub_rn(__fmul_rn(c.z, a.x), __fmul_rn(c.x, a.z)),
           __fsub_rn(__fmul_rn(c.x, a.y), __fmul_rn(c.y, a.x)));
}
static __device__ __noinline__ float3 hg_sample_dev(float g, float3 wi, float u1, float u2) {
    float cosTheta;
    if (fabsf(g) < MER_EPSILON) {
        cosTheta = __fsub_rn(1.0f, __fmul_rn(2.0f, u1));
    } else {
        const float gg = __fmul_rn(g, g);
        float sqrTerm = __fdiv_rn(__fsub_rn(1.0f, gg), __fadd_rn(__fsub_rn(1.0f, g), __fmul_rn(__fmul_rn(2.0f, g), u1)));
        cosTheta = __fdiv_rn(__fsub_rn(__fadd_rn(1.0f, gg), __fmul_rn(sqrTerm, sqrTerm)), __fmul_rn(2.0f, g));
    }
    float sinTheta = __fsqrt_rn(fmaxf(0.0f, __fsub_rn(1.0f, __fmul_rn(cosTheta, cosTheta))));
    /* `2*M_PI*sample.y` is a FLOAT product: constants.h:42-44,84-86 re-define M_PI as M_PI_FLT under -DSINGLE_PRECISION
     * (checked against src/phase/hg.cpp compiled verbatim: tests/golden/phase_ref.npz) */
    float phi = __fmul_rn(6.2831854820251464844f, u2);
    float sinPhi, cosPhi;
    sincosf(phi, &sinPhi, &cosPhi);
    float3 nrm = f3(-wi.x, -wi.y, -wi.z), s, t;
    coordinate_system(nrm, s, t);
    float lx = __fmul_rn(sinTheta, cosPhi), ly = __fmul_rn(sinTheta, sinPhi), lz = cosTheta;
    /* Frame::toWorld: s * v.x + t * v.y + n * v.z */
    return f3(__fadd_rn(__fadd_rn(__fmul_rn(s.x, lx), __fmul_rn(t.x, ly)), __fmul_rn(nrm.x, lz)),
              __fadd_rn(__fadd_rn(__fmul_rn(s.y, lx), __fmul_rn(t.y, ly)), __fmul_rn(nrm.y, lz)),
              __fadd_rn(__fadd_rn(__fmul_rn(s.z, lx), __fmul_rn(t.z, ly)), __fmul_rn(nrm.z, lz)));
}

/* fastlog / fastexp: double precision rounded to float on Linux x86-64, math.h:185-199 */
static __device__ __noinline__ float fastlog_dev(float x) { return (float) log((double) x); }
static __device__ __noinline__ float fastexp_dev(float x) { return (float) exp((double) x); }

/* ------------------------------------------------------------------ MaxExpDist, src/medium/maxexp.h:60-102
 * (std::lower_bound over three / four entries written out; every operation rounded like the reference's float build) */
__device__ __forceinline__ int maxexp_piece_of_t(const MediumDev &M, float t) { /* lower_bound(intervalStart, t) - 1, >= 0 */
    const int lb = (M.mxStart[0] >= t) ? 0 : ((M.mxStart[1] >= t) ? 1 : ((M.mxStart[2] >= t) ? 2 : 3));
    return max(0, lb - 1);
}
static __device__ __noinline__ float maxexp_sample(const MediumDev &M, float u, float &pdf) {
    const int lb = (M.mxCdf[0] >= u) ? 0 : ((M.mxCdf[1] >= u) ? 1 : ((M.mxCdf[2] >= u) ? 2 : ((M.mxCdf[3] >= u) ? 3 : 4)));
    const int index = min(max(0, lb - 1), 2);
    const float s = M.mxSigma[index];
    const float a = fastexp_dev(__fmul_rn(-M.mxStart[index], s));
    const float t = __fdiv_rn(-fastlog_dev(__fsub_rn(a, __fmul_rn(M.mxNorm, __fsub_rn(u, M.mxCdf[index])))), s);
    pdf = __fmul_rn(__fmul_rn(s, fastexp_dev(__fmul_rn(-s, t))), M.mxInvNorm);
    return t;
}
static __device__ __noinline__ float maxexp_pdf(const MediumDev &M, float t) {
    const float s = M.mxSigma[maxexp_piece_of_t(M, t)];
    return __fmul_rn(__fmul_rn(s, fastexp_dev(__fmul_rn(-s, t))), M.mxInvNorm);
}
static __device__ __noinline__ float maxexp_cdf(const MediumDev &M, float t) {
    const int index = maxexp_piece_of_t(M, t);
    const float upper = -fastexp_dev(__fmul_rn(-M.mxSigma[index], t));
    return __fadd_rn(M.mxCdf[index], __fmul_rn(__fsub_rn(upper, M.mxLower[index]), M.mxInvNorm));
}

/* ------------------------------------------------------------------ Philox4x32-10
 * key = seed, counter = (sample id lo, hi, block, 0); float = (u >> 8) * 2^-24.  The k-th
 * float of a sample's stream is word k%4 of block k/4, so the only per-path RNG state is k. */
static __device__ __noinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
struct PathRng {
    uint32_t k0, k1, s0, s1, k; /* key, sample id, draw index */
    uint32_t cachedBlock;       /* index of the block held in `cache` (0xffffffff: none) */
    uint4 cache;
    __device__ __forceinline__ void init(uint64_t seed, uint64_t sampleId, uint32_t draw) {
        k0 = (uint32_t) seed; k1 = (uint32_t) (seed >> 32);
        s0 = (uint32_t) sampleId; s1 = (uint32_t) (sampleId >> 32);
        k = draw;
        cachedBlock = 0xffffffffu;
    }
    __device__ __forceinline__ float next() {
        const uint32_t blk = k >> 2;
        if (blk != cachedBlock) {
            cache = philox4x32_10(s0, s1, blk, 0u, k0, k1);
            cachedBlock = blk;
        }
        const uint32_t w = (k & 3u) == 0 ? cache.x : ((k & 3u) == 1 ? cache.y : ((k & 3u) == 2 ? cache.z : cache.w));
        k++;
        return (float) (w >> 8) * (1.0f / 16777216.0f);
    }
};

/* ------------------------------------------------------------------ container surface */
/* fresnelDielectricExt, src/libcore/util.cpp:665-695 */
static __device__ __forceinline__ float fresnel_dielectric_ext(float cosThetaI_, float &cosThetaT_, float eta) {
    if (eta == 1.0f) { cosThetaT_ = -cosThetaI_; return 0.0f; }
    const float scale = (cosThetaI_ > 0.0f) ? __fdiv_rn(1.0f, eta) : eta;
    const float cosThetaTSqr = __fsub_rn(1.0f, __fmul_rn(__fsub_rn(1.0f, __fmul_rn(cosThetaI_, cosThetaI_)), __fmul_rn(scale, scale)));
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    const float cosThetaI = fabsf(cosThetaI_), cosThetaT = __fsqrt_rn(cosThetaTSqr);
    const float ect = __fmul_rn(eta, cosThetaT), eci = __fmul_rn(eta, cosThetaI);
    const float Rs = __fdiv_rn(__fsub_rn(cosThetaI, ect), __fadd_rn(cosThetaI, ect));
    const float Rp = __fdiv_rn(__fsub_rn(eci, cosThetaT), __fadd_rn(eci, cosThetaT));
    cosThetaT_ = (cosThetaI_ > 0.0f) ? -cosThetaT : cosThetaT;
    return __fmul_rn(0.5f, __fadd_rn(__fmul_rn(Rs, Rs), __fmul_rn(Rp, Rp)));
}

/* outward unit normal of the container at a surface point */
static __device__ __forceinline__ float3 shape_normal(const MediumDev &M, float3 p) {
    if (M.shapeType == MER_SHAPE_SPHERE) {
        float3 d = f3(p.x - M.shape[0], p.y - M.shape[1], p.z - M.shape[2]);
        float l = 1.0f / sqrtf(dot3(d, d));
        return f3(d.x * l, d.y * l, d.z * l);
    }
    const float pp[3] = {p.x, p.y, p.z};
    int axis = 0;
    float best = INFINITY, sign = 1.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float a = fabsf(pp[i] - M.shape[i]), b = fabsf(pp[i] - M.shape[3 + i]);
        if (a < best) { best = a; axis = i; sign = -1.0f; }
        if (b < best) { best = b; axis = i; sign = 1.0f; }
    }
    return f3(axis == 0 ? sign : 0.0f, axis == 1 ? sign : 0.0f, axis == 2 ? sign : 0.0f);
}

/* distance along a straight ray from a point inside the container to its surface (edge.cpp:45-67 re-finds the
 * surface point of a curved segment with a straight ray from the last interior point) */
#define MER_SDF_TRACE_STEPS 512
#define MER_SDF_TRACE_EPS 1e-4f

/* MER_SHAPE_SDF: sphere tracing from inside, advance by -sdf until the sign changes */
static __device__ __forceinline__ float exit_distance_sdf(const MediumDev &M, float3 o, float3 d) {
    float t = 0.0f;
    for (int i = 0; i < MER_SDF_TRACE_STEPS; i++) {
        const float v = sdf_value(M.sdf, f3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z));
        if (v >= 0.0f) break;
        t += fmaxf(-v, MER_SDF_TRACE_EPS);
    }
    return t;
}

static __device__ __forceinline__ float exit_distance(const MediumDev &M, float3 o, float3 d) {
    if (M.shapeType == MER_SHAPE_SPHERE) {
        float3 oc = f3(o.x - M.shape[0], o.y - M.shape[1], o.z - M.shape[2]);
        float b = dot3(oc, d), c = dot3(oc, oc) - M.shape[3] * M.shape[3];
        float disc = b * b - c;
        return disc > 0.0f ? fmaxf(-b + sqrtf(disc), 0.0f) : 0.0f;
    }
    float t1 = INFINITY;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (dd[i] == 0.0f) continue;
        float inv = 1.0f / dd[i];
        float ta = (M.shape[i] - oo[i]) * inv, tb = (M.shape[3 + i] - oo[i]) * inv;
        t1 = fminf(t1, fmaxf(ta, tb));
    }
    return fmaxf(t1, 0.0f);
}
