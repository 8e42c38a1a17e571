/*
 * mer_volume.cu — VolumeDataSource side of the path: .vol I/O, the B-spline prefilter on the
 * GPU, the RIF / density handles and their batch evaluation entry points.
 *
 * Reference (paths relative to the MitsubaER tree):
 *   Spline<3>::initialize / build1d / build3d   include/mitsuba/core/basisspline.h:124-138, 812-840, 865-890
 *   SplineDataSource                            src/volume/splinevolume.cpp:87-111, 204-360
 *   GridDataSource                              src/volume/gridvolume.cpp:188-199, 337-363
 *   .vol v3 format                              src/volume/splinevolume.cpp:39-75, mfiles/writeGridToVol.m
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mer_internal.h"

/* ===================================================================== prefilter (a3)
 * One thread per grid line; three passes (y, x, z — the reference's order).  Line l of a pass
 * starts at `base(l)` and advances by `stride`.  In the y and z passes adjacent threads own
 * adjacent x, so every load/store of a warp is one coalesced 128-byte line; in the x pass a
 * thread walks its own 128-byte lines (L1-resident between consecutive elements).
 *
 * Arithmetic follows build1d exactly: the initial causal coefficient is accumulated with the
 * products formed in double (C pow() semantics) and the running sum rounded to float after every
 * term; the recursions are single precision WITHOUT fma contraction (__fmul_rn/__fadd_rn), as
 * the reference's x86 build has none.  z1^e terms with e >= 64 (< 1e-36) are skipped: they
 * cannot change a float accumulator.
 */
enum { PASS_Y = 0, PASS_X = 1, PASS_Z = 2 };

__global__ void __launch_bounds__(128)
k_prefilter(const float *in, float *out, int N0, int N1, int N2, int pass) {
    const size_t n0 = N0, n01 = (size_t) N0 * N1;
    size_t nLines, stride, base;
    int len;
    const size_t l = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (pass == PASS_Y) {
        nLines = n0 * N2; len = N1; stride = n0;
        base = (l / n0) * n01 + (l % n0);
    } else if (pass == PASS_X) {
        nLines = (size_t) N1 * N2; len = N0; stride = 1;
        base = l * n0;
    } else {
        nLines = n01; len = N2; stride = n01;
        base = l;
    }
    if (l >= nLines) return;
    const float *src = in + base;
    float *dst = out + base;

    const float z1 = (float) (-2.0 + sqrt(3.0));
    const double z1d = (double) z1;
    const int CUT = 64;
    float cp = 0.0f;
    double zp = 1.0;
#pragma unroll 8
    for (int i = 0; i < len && i < CUT; i++) {
        cp = (float) ((double) cp + (double) src[(size_t) i * stride] * zp);
        zp *= z1d;
    }
    for (int i = len - 2; i > 0; i--) {
        int e = 2 * len - 2 - i;
        if (e >= CUT) break; /* e grows as i decreases */
        cp = (float) ((double) cp + (double) src[(size_t) i * stride] * pow(z1d, (double) e));
    }
    cp = (float) ((double) cp / (1.0 - pow(z1d, (double) (2 * len - 2))));

    /* The recursions are a dependent chain per line, so a thread that loads one element at a time keeps one 4-byte load in
     * flight: 2048 threads per SM x 4 bytes is a quarter of what the HBM latency asks for (measured: 1.55 TB/s at 1024^3).
     * Elements are therefore loaded PF at a time ahead of the chain (same operations in the same order: bit-identical). */
    constexpr int PF = 8;
    /* causal recursion, storing c+ in place */
    dst[0] = cp;
    float cpPrev2 = cp;
    for (int i0 = 1; i0 < len; i0 += PF) {
        float v[PF];
#pragma unroll
        for (int k = 0; k < PF; k++)
            if (i0 + k < len) v[k] = src[(size_t) (i0 + k) * stride];
#pragma unroll
        for (int k = 0; k < PF; k++)
            if (i0 + k < len) {
                cpPrev2 = cp;
                cp = __fadd_rn(v[k], __fmul_rn(z1, cp));
                dst[(size_t) (i0 + k) * stride] = cp;
            }
    }
    /* anti-causal recursion */
    const float gain = __fdiv_rn(z1, __fsub_rn(__fmul_rn(z1, z1), 1.0f));
    float cn = __fmul_rn(gain, __fadd_rn(cp, __fmul_rn(z1, cpPrev2)));
    dst[(size_t) (len - 1) * stride] = __fmul_rn(6.0f, cn);
    for (int i0 = len - 2; i0 >= 0; i0 -= PF) {
        float v[PF];
#pragma unroll
        for (int k = 0; k < PF; k++)
            if (i0 - k >= 0) v[k] = dst[(size_t) (i0 - k) * stride];
#pragma unroll
        for (int k = 0; k < PF; k++)
            if (i0 - k >= 0) {
                cn = __fmul_rn(z1, __fsub_rn(cn, v[k]));
                dst[(size_t) (i0 - k) * stride] = __fmul_rn(6.0f, cn);
            }
    }
}

/* The x pass, tiled: a thread still owns one line and runs build1d's recursions in the same order (bit-identical to
 * k_prefilter's PASS_X), but the line's elements travel through a shared-memory tile of 32 x-values by PFX_LINES lines, so
 * every global load and store is a warp reading 128 consecutive bytes of ONE line instead of 32 threads touching 32
 * different lines (k_prefilter's x pass ran at half the rate of its y and z passes: profiles/r02_launch_summary.txt).
 * Lines of a block are consecutive in memory (line l starts at l * N0).  Forward sweep over the tiles: c+ in place;
 * backward sweep: c- from c+, scaled by 6. */
enum { PFX_LINES = 128, PFX_TILE = 32 }; /* 4 warps, each moving every 4th line of a tile */

__global__ void __launch_bounds__(PFX_LINES)
k_prefilter_x(float *data, int N0, size_t nLines) {
    __shared__ float tile[PFX_LINES][PFX_TILE + 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t line0 = (size_t) blockIdx.x * PFX_LINES;
    const int linesHere = (int) min((size_t) PFX_LINES, nLines - line0);
    float *blockBase = data + line0 * (size_t) N0;
    const int len = N0;
    const int nTiles = (len + PFX_TILE - 1) / PFX_TILE;
    const bool mine = t < linesHere;

    auto loadTile = [&](int c) {
        const int x = c * PFX_TILE + lane;
#pragma unroll 8
        for (int r = warp; r < PFX_LINES; r += PFX_LINES / 32)
            if (r < linesHere && x < len) tile[r][lane] = blockBase[(size_t) r * len + x];
        __syncthreads();
    };
    auto storeTile = [&](int c) {
        __syncthreads();
        const int x = c * PFX_TILE + lane;
#pragma unroll 8
        for (int r = warp; r < PFX_LINES; r += PFX_LINES / 32)
            if (r < linesHere && x < len) blockBase[(size_t) r * len + x] = tile[r][lane];
        __syncthreads();
    };

    /* the next tile waits in registers while the block runs the recursion on the current one: the loads of a block overlap its
     * own arithmetic (a thread reads back exactly the addresses it stored in the other sweep, so program order is enough) */
    float nxt[PFX_LINES / 4];
    auto fetch = [&](int c) {
        const int x = c * PFX_TILE + lane;
#pragma unroll
        for (int k = 0; k < PFX_LINES / 4; k++) {
            const int r = warp + 4 * k;
            nxt[k] = (r < linesHere && x < len) ? blockBase[(size_t) r * len + x] : 0.0f;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int k = 0; k < PFX_LINES / 4; k++) tile[warp + 4 * k][lane] = nxt[k];
        __syncthreads();
    };

    const float z1 = (float) (-2.0 + sqrt(3.0));
    const double z1d = (double) z1;
    const int CUT = 64;
    /* initial causal coefficient: the first CUT elements (two tiles), then build1d's mirrored tail for lines shorter than CUT */
    float cp = 0.0f;
    double zp = 1.0;
    for (int c = 0; c < nTiles && c * PFX_TILE < CUT; c++) {
        loadTile(c);
        if (mine)
            for (int j = 0; j < PFX_TILE; j++) {
                const int i = c * PFX_TILE + j;
                if (i >= len || i >= CUT) break;
                cp = (float) ((double) cp + (double) tile[t][j] * zp);
                zp *= z1d;
            }
        __syncthreads();
    }
    if (mine) {
        const float *src = blockBase + (size_t) t * len;
        for (int i = len - 2; i > 0; i--) {
            int e = 2 * len - 2 - i;
            if (e >= CUT) break;
            cp = (float) ((double) cp + (double) src[i] * pow(z1d, (double) e));
        }
        cp = (float) ((double) cp / (1.0 - pow(z1d, (double) (2 * len - 2))));
    }

    /* causal recursion */
    float cpPrev2 = cp;
    fetch(0);
    for (int c = 0; c < nTiles; c++) {
        commit();
        if (c + 1 < nTiles) fetch(c + 1);
        if (mine)
#pragma unroll 4
            for (int j = 0; j < PFX_TILE; j++) {
                const int i = c * PFX_TILE + j;
                if (i >= len) break;
                if (i > 0) {
                    cpPrev2 = cp;
                    cp = __fadd_rn(tile[t][j], __fmul_rn(z1, cp));
                }
                tile[t][j] = cp;
            }
        storeTile(c);
    }
    /* anti-causal recursion */
    const float gain = __fdiv_rn(z1, __fsub_rn(__fmul_rn(z1, z1), 1.0f));
    float cn = 0.0f;
    fetch(nTiles - 1);
    for (int c = nTiles - 1; c >= 0; c--) {
        commit();
        if (c > 0) fetch(c - 1);
        if (mine)
#pragma unroll 4
            for (int j = PFX_TILE - 1; j >= 0; j--) {
                const int i = c * PFX_TILE + j;
                if (i >= len) continue;
                if (i == len - 1) cn = __fmul_rn(gain, __fadd_rn(cp, __fmul_rn(z1, cpPrev2)));
                else cn = __fmul_rn(z1, __fsub_rn(cn, tile[t][j]));
                tile[t][j] = __fmul_rn(6.0f, cn);
            }
        storeTile(c);
    }
}

/* coeff [z][y][x] -> coeff8 [z][y][x] = { (c[x-1], c[x], c[x+1], c[x+2]) of row y, the same of row y+1 }, indices clamped
 * (only interior cells are ever read through this table: rif_cell_interior) */
__global__ void k_expand_coeff8(const float *__restrict__ coeff, float4 *__restrict__ coeff8, int N0, int N1, size_t total) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const int x = (int) (i % (size_t) N0), y = (int) ((i / (size_t) N0) % (size_t) N1);
        const float *row = coeff + (i - x);
        const float *rowUp = y + 1 < N1 ? row + N0 : row;
        const int xm = max(x - 1, 0), x1 = min(x + 1, N0 - 1), x2 = min(x + 2, N0 - 1);
        coeff8[2 * i] = make_float4(row[xm], row[x], row[x1], row[x2]);
        coeff8[2 * i + 1] = make_float4(rowUp[xm], rowUp[x], rowUp[x1], rowUp[x2]);
    }
}

/* fast mode: sample the spline (value + gradient) at every grid node */
__global__ void k_build_packed(RifDev R, float4 *__restrict__ packed, size_t total) {
    const size_t n0 = R.N[0], n01 = (size_t) R.N[0] * R.N[1];
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t) gridDim.x * blockDim.x) {
        const int x = (int) (i % n0), y = (int) ((i / n0) % R.N[1]), z = (int) (i / n01);
        /* node position in volume space: xmin + index / xres */
        float3 pv = f3(R.xmin[0] + (float) x / R.xres[0], R.xmin[1] + (float) y / R.xres[1],
                       R.xmin[2] + (float) z / R.xres[2]);
        float f;
        float3 g;
        rif_tricubic(R, pv, f, g);
        packed[i] = make_float4(f, g.x, g.y, g.z);
    }
}

/* SplineDataSource::value / gradient / valueAndGradient over a batch */
template <int MODE>
__global__ void k_rif_eval(RifDev R, size_t n, const float *__restrict__ p, float *__restrict__ f,
                           float *__restrict__ g) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float val;
        float3 grad;
        rif_lookup<MODE>(R, f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), val, grad);
        if (f) f[i] = val;
        if (g) { g[3 * i] = grad.x; g[3 * i + 1] = grad.y; g[3 * i + 2] = grad.z; }
    }
}

__global__ void k_rif_inside(RifDev R, size_t n, const float *__restrict__ p, uint8_t *__restrict__ out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = rif_inside_limits(R, f3(p[3 * i], p[3 * i + 1], p[3 * i + 2])) ? 1 : 0;
}

__global__ void k_grid_lookup(GridDev D, size_t n, const float *__restrict__ p, float *__restrict__ out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        out[i] = grid_lookup(D, f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}

__global__ void k_grid_lookup3(GridDev D, size_t n, const float *__restrict__ p, float *__restrict__ out) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float rgb[3];
        grid_lookup3(D, f3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), rgb);
        out[3 * i] = rgb[0]; out[3 * i + 1] = rgb[1]; out[3 * i + 2] = rgb[2];
    }
}

/* ------------------------------------------------------------------ a19: straight-ray Woodcock tracking
 * heterogeneous.cpp:613-658 / :546-587; every operation individually rounded like the reference's float code. */
__device__ __forceinline__ bool aabb_ray_intersect(const GridDev &D, float3 o, float3 d, float &nearT, float &farT) {
    nearT = -INFINITY;
    farT = INFINITY;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (dd[i] == 0.0f) {
            if (oo[i] < D.aabbLo[i] || oo[i] > D.aabbHi[i]) return false;
        } else {
            const float rcp = __fdiv_rn(1.0f, dd[i]);
            float t1 = __fmul_rn(__fsub_rn(D.aabbLo[i], oo[i]), rcp), t2 = __fmul_rn(__fsub_rn(D.aabbHi[i], oo[i]), rcp);
            if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
            nearT = fmaxf(t1, nearT);
            farT = fminf(t2, farT);
            if (!(nearT <= farT)) return false;
        }
    }
    return true;
}
__device__ __forceinline__ float3 ray_at(float3 o, float3 d, float t) { /* Ray::operator(): o + t * d */
    return f3(__fadd_rn(o.x, __fmul_rn(t, d.x)), __fadd_rn(o.y, __fmul_rn(t, d.y)), __fadd_rn(o.z, __fmul_rn(t, d.z)));
}

__global__ void k_grid_woodcock(GridDev D, float scale, size_t n, const float *__restrict__ RO, const float *__restrict__ RD,
                                const float *__restrict__ rmint, const float *__restrict__ rmaxt, unsigned long long seed,
                                int evalTransmittance, uint8_t *__restrict__ success, float *__restrict__ tOut,
                                float *__restrict__ densOut, float *__restrict__ trOut) {
    const float invMax = 1.0f / (scale * 1.0f); /* heterogeneous.cpp:239-242 */
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float3 o = f3(RO[3 * i], RO[3 * i + 1], RO[3 * i + 2]), d = f3(RD[3 * i], RD[3 * i + 1], RD[3 * i + 2]);
        PathRng rng;
        rng.init(seed, (unsigned long long) i, 0u);
        float mint, maxt;
        const bool hit = aabb_ray_intersect(D, o, d, mint, maxt);
        mint = fmaxf(mint, rmint[i]);
        maxt = fminf(maxt, rmaxt[i]);
        if (!evalTransmittance) {
            bool ok = false;
            float t = mint, dens = 0.0f, tHit = 0.0f;
            while (hit) {
                t = __fsub_rn(t, __fmul_rn(fastlog_dev(1.0f - rng.next()), invMax));
                if (t >= maxt) break;
                dens = __fmul_rn(grid_lookup(D, ray_at(o, d, t)), scale);
                if (__fmul_rn(dens, invMax) > rng.next()) { ok = true; tHit = t; break; }
            }
            success[i] = ok ? 1 : 0;
            tOut[i] = tHit;
            densOut[i] = hit ? dens : 0.0f;
        } else {
            float result = 0.0f;
            if (!hit) {
                result = 2.0f;
            } else {
                for (int s = 0; s < 2; s++) {
                    float t = mint;
                    while (true) {
                        t = __fsub_rn(t, __fmul_rn(fastlog_dev(1.0f - rng.next()), invMax));
                        if (t >= maxt) { result += 1.0f; break; }
                        const float dens = __fmul_rn(grid_lookup(D, ray_at(o, d, t)), scale);
                        if (__fmul_rn(dens, invMax) > rng.next()) break;
                    }
                }
            }
            trOut[i] = result / 2.0f;
        }
    }
}

/* ===================================================================== host side */
namespace mer {
struct CachedArray { int device; size_t W, H; cudaArray_t arr; };
static std::mutex g_arrLock;
static std::vector<CachedArray> g_arrCache;

cudaError_t array_acquire(int device, size_t W, size_t H, cudaArray_t *out) {
    {
        std::lock_guard<std::mutex> hold(g_arrLock);
        for (size_t i = 0; i < g_arrCache.size(); i++)
            if (g_arrCache[i].device == device && g_arrCache[i].W == W && g_arrCache[i].H == H) {
                *out = g_arrCache[i].arr;
                g_arrCache.erase(g_arrCache.begin() + i);
                return cudaSuccess;
            }
    }
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float>();
    cudaError_t e = cudaMallocArray(out, &fmt, W, H, cudaArrayTextureGather);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        array_cache_trim(device);
        pool_trim(device);
        e = cudaMallocArray(out, &fmt, W, H, cudaArrayTextureGather);
    }
    return e;
}
void array_release(int device, cudaArray_t arr, size_t W, size_t H) {
    if (!arr) return;
    std::lock_guard<std::mutex> hold(g_arrLock);
    size_t mine = 0;
    for (const CachedArray &c : g_arrCache) mine += c.device == device;
    if (mine >= 2) { cudaFreeArray(arr); return; } /* keep at most two per device (a RIF and an SDF) */
    g_arrCache.push_back({device, W, H, arr});
}
void array_cache_trim(int device) {
    std::lock_guard<std::mutex> hold(g_arrLock);
    for (size_t i = 0; i < g_arrCache.size();)
        if (g_arrCache[i].device == device) { cudaFreeArray(g_arrCache[i].arr); g_arrCache.erase(g_arrCache.begin() + i); }
        else i++;
}
} /* namespace mer */

namespace {

int validate_desc(const mer_volume_desc *d, int minRes) {
    if (!d) return mer::fail(MER_ERR_INVALID, "null volume descriptor");
    for (int i = 0; i < 3; i++) {
        if (d->res[i] < minRes) return mer::fail(MER_ERR_INVALID, "volume resolution too small");
        if (!(d->bbox_max[i] > d->bbox_min[i])) return mer::fail(MER_ERR_INVALID, "invalid volume bounding box");
    }
    return MER_OK;
}

size_t voxels(const mer_volume_desc *d) { return (size_t) d->res[0] * d->res[1] * d->res[2]; }

void fill_rif_dev(mer_rif *r) {
    RifDev &D = r->dev;
    const mer_volume_desc &d = r->desc;
    D.mode = r->mode;
    for (int i = 0; i < 3; i++) {
        D.N[i] = d.res[i];
        D.xmin[i] = d.bbox_min[i];
        D.xres[i] = (float) (d.res[i] - 1) / (d.bbox_max[i] - d.bbox_min[i]); /* basisspline.h:130 */
        /* m_interpolatableLimits, splinevolume.cpp:280-281: 2*stride + Epsilon in double, stored single */
        float stride = (float) (1.0 / (double) D.xres[i]); /* getStride(), basisspline.h:622-624 */
        float margin = (float) (2.0 * (double) stride + (double) MER_EPSILON);
        D.limLo[i] = d.bbox_min[i] + margin;
        D.limHi[i] = d.bbox_max[i] + (-margin);
    }
    D.hasXform = d.has_transform != 0;
    for (int i = 0; i < 12; i++) D.M[i] = D.hasXform ? d.world_to_volume[i] : ((i == 0 || i == 5 || i == 10) ? 1.f : 0.f);
    D.coeff = r->d_coeff;
    D.packed = r->d_packed;
    D.tex = (unsigned long long) r->tex; /* tileShift / tileMask: rif_build_texture */
    D.coeff8 = r->d_coeff8;
}

/* does tld4 return (w, z, x, y) = (i0,j0), (i1,j0), (i0,j1), (i1,j1) and land on the texels it is asked for? */
__global__ void k_check_gather(RifDev R, int i, int j, int k, int *bad) {
    const float ox = (float) ((k & R.tileMask) * R.N[0]), oy = (float) ((k >> R.tileShift) * R.N[1]);
    const float4 g = tex2Dgather<float4>((cudaTextureObject_t) R.tex, ox + (float) (i + 1), oy + (float) (j + 1), 0);
    const size_t o = ((size_t) k * R.N[1] + j) * (size_t) R.N[0] + i;
    const float c00 = R.coeff[o], c10 = R.coeff[o + 1], c01 = R.coeff[o + R.N[0]], c11 = R.coeff[o + R.N[0] + 1];
    *bad = !(g.w == c00 && g.z == c10 && g.x == c01 && g.y == c11);
}

/* texture storage of the coefficients: layer k of coeff -> tile (k & mask, k >> shift) of a 2-D CUDA array (block-linear;
 * cudaArrayTextureGather).  The gather limit of sm_100 is 32768 x 32768 texels: 1024^3 fits exactly (32 x 32 tiles). */
int rif_build_texture(mer_rif *r, cudaStream_t s) {
    const mer_volume_desc &d = r->desc;
    fill_rif_dev(r);
    int maxW = 0, maxH = 0;
    MER_CUDA(cudaDeviceGetAttribute(&maxW, cudaDevAttrMaxTexture2DGatherWidth, r->device));
    MER_CUDA(cudaDeviceGetAttribute(&maxH, cudaDevAttrMaxTexture2DGatherHeight, r->device));
    /* tiles per atlas row: a power of two (tile origin = shift and mask), as square as the limits allow */
    int shift = 0;
    while ((1 << (2 * shift)) < d.res[2]) shift++;
    while (shift > 0 && (size_t) d.res[0] << shift > (size_t) maxW) shift--;
    while ((size_t) d.res[1] * (size_t) ((d.res[2] + (1 << shift) - 1) >> shift) > (size_t) maxH && ((size_t) d.res[0] << (shift + 1)) <= (size_t) maxW) shift++;
    const int TX = 1 << shift, TY = (d.res[2] + TX - 1) / TX;
    const size_t W = (size_t) d.res[0] * TX, H = (size_t) d.res[1] * TY;
    if (W > (size_t) maxW || H > (size_t) maxH)
        return mer::fail(MER_ERR_UNSUPPORTED, "volume exceeds the texture atlas (gather textures are limited to 32768 x 32768 texels: res_x * T <= 32768 and "
                                              "res_y * ceil(res_z / T) <= 32768 for a power of two T)");
    r->dev.tileShift = shift;
    r->dev.tileMask = TX - 1;
    MER_CUDA(mer::array_acquire(r->device, W, H, &r->texArray));
    r->texW = W; r->texH = H;
    for (int k = 0; k < d.res[2]; k++)
        MER_CUDA(cudaMemcpy2DToArrayAsync(r->texArray, (size_t) (k & (TX - 1)) * d.res[0] * sizeof(float), (size_t) (k >> shift) * d.res[1],
                                          r->d_coeff + (size_t) k * d.res[0] * d.res[1], (size_t) d.res[0] * sizeof(float),
                                          (size_t) d.res[0] * sizeof(float), (size_t) d.res[1], cudaMemcpyDeviceToDevice, s));
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = r->texArray;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    MER_CUDA(cudaCreateTextureObject(&r->tex, &res, &td, nullptr));
    r->dev.tex = (unsigned long long) r->tex;
    int *bad = nullptr, hbad = 1;
    MER_CUDA(cudaMalloc(&bad, sizeof(int)));
    MER_LAUNCH(k_check_gather, 1, 1, 0, s, r->dev, d.res[0] / 2, d.res[1] / 3, d.res[2] - 2, bad);
    cudaError_t e = cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(bad);
    if (e != cudaSuccess) return mer::fail(MER_ERR_CUDA, cudaGetErrorString(e));
    if (hbad) return mer::fail(MER_ERR_CUDA, "texture gather does not return the expected texels (component order / footprint)");
    return MER_OK;
}

int rif_build(mer_rif *r, const float *data_dev, cudaStream_t s) {
    const mer_volume_desc &d = r->desc;
    const size_t total = voxels(&d);
    MER_CUDA(mer::pool_malloc((void **) &r->d_coeff, total * sizeof(float)));
    const int N0 = d.res[0], N1 = d.res[1], N2 = d.res[2];
    const unsigned T = 128;
    MER_LAUNCH(k_prefilter, mer_blocks((size_t) N0 * N2, T), T, 0, s, data_dev, r->d_coeff, N0, N1, N2, (int) PASS_Y);
    if (getenv("MER_PREFILTER_X_UNTILED")) /* the one-thread-per-line x pass, kept for comparison (bit-identical) */
        MER_LAUNCH(k_prefilter, mer_blocks((size_t) N1 * N2, T), T, 0, s, r->d_coeff, r->d_coeff, N0, N1, N2, (int) PASS_X);
    else
        MER_LAUNCH(k_prefilter_x, mer_blocks((size_t) N1 * N2, PFX_LINES), PFX_LINES, 0, s, r->d_coeff, N0, (size_t) N1 * N2);
    MER_LAUNCH(k_prefilter, mer_blocks((size_t) N0 * N1, T), T, 0, s, r->d_coeff, r->d_coeff, N0, N1, N2, (int) PASS_Z);
    const unsigned G = (unsigned) std::min<size_t>(mer_blocks(total, 256), 148u * 16u);
    fill_rif_dev(r);
    /* Storage of the cubic coefficients for the 4x4x4 gathers.  "atlas" (default) = a 1x texture atlas read
     * with 16 tld4: 20 cycles of the SM's texture unit per lane and stencil (it returns one lane's four texels per 1.27
     * cycles), L2-resident up to ~300^3.  "coeff8" (MER_RIF_LAYOUT=coeff8) = an 8x table of 32-byte sectors read with 8
     * LDG.E.256: 10.6 cycles of the L1 per lane and stencil, but 512 MiB at 256^3 — every stencil comes from DRAM, and
     * with 16 warps per SM that latency is not covered (C2: 41 vs 50 M samples/s; profiles/README.md). */
    bool useCoeff8 = false;
    if (const char *e = getenv("MER_RIF_LAYOUT")) useCoeff8 = !strcmp(e, "coeff8") && r->mode == MER_RIF_TRICUBIC;
    if (useCoeff8) {
        cudaError_t e = mer::pool_malloc((void **) &r->d_coeff8, 2 * total * sizeof(float4));
        if (e != cudaSuccess) { cudaGetLastError(); r->d_coeff8 = nullptr; useCoeff8 = false; } /* no room: the atlas */
    }
    if (useCoeff8) {
        MER_LAUNCH(k_expand_coeff8, G, 256, 0, s, r->d_coeff, r->d_coeff8, N0, N1, total);
        r->dev.coeff8 = r->d_coeff8;
    } else {
        int rc = rif_build_texture(r, s);
        if (rc) return rc;
    }
    if (r->mode == MER_RIF_TRILINEAR_PACKED) {
        MER_CUDA(mer::pool_malloc((void **) &r->d_packed, total * sizeof(float4)));
        r->dev.packed = r->d_packed;
        MER_LAUNCH(k_build_packed, G, 256, 0, s, r->dev, r->d_packed, total);
        MER_CUDA(cudaStreamSynchronize(s));
        /* the cubic atlas was only needed to sample the spline at the nodes */
        cudaDestroyTextureObject(r->tex);
        mer::array_release(r->device, r->texArray, r->texW, r->texH);
        r->tex = 0;
        r->texArray = nullptr;
        r->dev.tex = 0;
    }
    MER_CUDA(cudaStreamSynchronize(s));
    return MER_OK;
}

/* .vol payloads this path reads: EFloat32 and EUInt8 (value / 255), one channel (densities, RIFs, SDFs) or three
 * interleaved channels (albedo grids) (gridvolume.cpp:251-262, 369-376, 401-460) */
static bool vol_encoding_supported(int32_t enc, int32_t ch) { return (ch == 1 || ch == 3) && (enc == 1 || enc == 3); }
/* the reference's own messages where it refuses a payload too (gridvolume.cpp:243-268) */
static int vol_unsupported(int32_t enc, int32_t ch) {
    char msg[160];
    if (enc == 2) return mer::fail(MER_ERR_UNSUPPORTED, "Error: float16 volumes are not yet supported!");
    if (enc == 4) return mer::fail(MER_ERR_UNSUPPORTED, "quantized-direction volumes feed the micro-flake phase functions, which are not on this path");
    snprintf(msg, sizeof(msg), "Encountered an unsupported %s volume data file (%i channels, only 1 and 3 are supported)", enc == 1 ? "float32" : "uint8", (int) ch);
    return mer::fail(MER_ERR_UNSUPPORTED, msg);
}

int read_vol_file(const char *path, mer_volume_desc *desc, std::vector<float> *data, int32_t *encoding,
                  int32_t *channels) {
    FILE *f = fopen(path, "rb");
    if (!f) return mer::fail(MER_ERR_INVALID, std::string("cannot open volume file: ") + path);
    unsigned char hdr[48];
    if (fread(hdr, 1, 48, f) != 48) { fclose(f); return mer::fail(MER_ERR_INVALID, "volume file too short"); }
    if (hdr[0] != 'V' || hdr[1] != 'O' || hdr[2] != 'L') {
        fclose(f);
        return mer::fail(MER_ERR_INVALID, "Encountered an invalid volume data file (incorrect header identifier)");
    }
    if (hdr[3] != 3) {
        fclose(f);
        return mer::fail(MER_ERR_INVALID, "Encountered an invalid volume data file (incorrect file version)");
    }
    int32_t enc, res[3], ch;
    float bb[6];
    memcpy(&enc, hdr + 4, 4);
    memcpy(res, hdr + 8, 12);
    memcpy(&ch, hdr + 20, 4);
    memcpy(bb, hdr + 24, 24);
    if (encoding) *encoding = enc;
    if (channels) *channels = ch;
    /* the header is untrusted input: sizes are checked before anything is multiplied or allocated */
    for (int i = 0; i < 3; i++)
        if (res[i] < 1 || res[i] > 65536) { fclose(f); return mer::fail(MER_ERR_INVALID, "Encountered an invalid volume data file (resolution out of range)"); }
    if (ch < 1 || ch > 4) { fclose(f); return mer::fail(MER_ERR_INVALID, "Encountered an invalid volume data file (channel count)"); }
    {
        const size_t bytesPer = enc == 1 ? 4 : (enc == 2 ? 2 : (enc == 3 ? 1 : (enc == 4 ? 2 : 0)));
        fseek(f, 0, SEEK_END);
        const long long fileBytes = (long long) ftell(f);
        fseek(f, 48, SEEK_SET);
        const unsigned long long need = 48ull + (unsigned long long) res[0] * (unsigned long long) res[1] * (unsigned long long) res[2] *
                                                  (unsigned long long) ch * (unsigned long long) bytesPer;
        if (bytesPer == 0) { fclose(f); return mer::fail(MER_ERR_INVALID, "Encountered a volume data file of unknown type"); } /* gridvolume.cpp:262 */
        if ((unsigned long long) fileBytes < need) { fclose(f); return mer::fail(MER_ERR_INVALID, "volume file truncated"); }
    }
    if (desc) {
        memset(desc, 0, sizeof(*desc));
        for (int i = 0; i < 3; i++) { desc->res[i] = res[i]; desc->bbox_min[i] = bb[i]; desc->bbox_max[i] = bb[3 + i]; }
        desc->world_to_volume[0] = desc->world_to_volume[5] = desc->world_to_volume[10] = 1.f;
    }
    if (data) {
        if (!vol_encoding_supported(enc, ch)) {
            fclose(f);
            return vol_unsupported(enc, ch);
        }
        size_t total = (size_t) res[0] * res[1] * res[2] * (size_t) ch;
        data->resize(total);
        if (enc == 1) {
            if (fread(data->data(), sizeof(float), total, f) != total) { fclose(f); return mer::fail(MER_ERR_INVALID, "volume file truncated"); }
        } else { /* EUInt8: m_densityMap[i] = i / 255 (gridvolume.cpp:369-376) */
            std::vector<unsigned char> bytes(total);
            if (fread(bytes.data(), 1, total, f) != total) { fclose(f); return mer::fail(MER_ERR_INVALID, "volume file truncated"); }
            for (size_t i = 0; i < total; i++) (*data)[i] = (float) bytes[i] / 255.0f;
        }
    }
    fclose(f);
    return MER_OK;
}

/* Streams the payload of a single-channel float32 (or uint8) .vol file into a fresh device float array: 64 MiB slabs through two pinned
 * staging buffers, the disk read of one overlapping the H2D copy of the other.  Peak host memory is 128 MiB whatever the
 * grid (a 1024^3 RIF is 4 GiB on disk); the reference reads the file through a memory map and prefilters on the CPU
 * (splinevolume.cpp:204-317), here the prefilter runs on the device array. */
int stream_vol_to_device(int device, const char *path, mer_volume_desc *desc, float **data_dev_out, int *channels_out = nullptr) {
    *data_dev_out = nullptr;
    int32_t enc = 0, ch = 0;
    int rc = read_vol_file(path, desc, nullptr, &enc, &ch);
    if (rc) return rc;
    if (!vol_encoding_supported(enc, ch)) return vol_unsupported(enc, ch);
    /* a caller that does not ask for the channel count takes scalar fields only (RIF, SDF) */
    if (!channels_out && ch != 1) return mer::fail(MER_ERR_INVALID, "this volume must have one channel (supportsFloatLookups)");
    if (channels_out) *channels_out = ch;
    rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    const size_t total = (size_t) desc->res[0] * desc->res[1] * desc->res[2] * (size_t) ch;
    FILE *f = fopen(path, "rb");
    if (!f) return mer::fail(MER_ERR_INVALID, std::string("cannot open volume file: ") + path);
    fseek(f, 48, SEEK_SET);
    const size_t slab = std::min<size_t>(total, (size_t) 16 << 20); /* floats per slab */
    float *dev = nullptr, *stage[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
    cudaError_t e = mer::pool_malloc((void **) &dev, total * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreate(&st);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaMallocHost(&stage[i], slab * sizeof(float));
        if (e == cudaSuccess) e = cudaEventCreate(&done[i]);
    }
    bool truncated = false;
    std::vector<unsigned char> bytes;
    for (size_t off = 0, k = 0; e == cudaSuccess && off < total; off += slab, k++) {
        const int b = (int) (k & 1);
        const size_t n = std::min(slab, total - off);
        if (k >= 2) e = cudaEventSynchronize(done[b]); /* the copy that last used this buffer */
        if (e != cudaSuccess) break;
        if (enc == 1) {
            if (fread(stage[b], sizeof(float), n, f) != n) { truncated = true; break; }
        } else { /* EUInt8 -> float on the way through the staging buffer */
            bytes.resize(n);
            if (fread(bytes.data(), 1, n, f) != n) { truncated = true; break; }
            for (size_t i = 0; i < n; i++) stage[b][i] = (float) bytes[i] / 255.0f;
        }
        e = cudaMemcpyAsync(dev + off, stage[b], n * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(done[b], st);
    }
    if (st) cudaStreamSynchronize(st);
    fclose(f);
    for (int i = 0; i < 2; i++) { if (stage[i]) cudaFreeHost(stage[i]); if (done[i]) cudaEventDestroy(done[i]); }
    if (st) cudaStreamDestroy(st);
    if (truncated || e != cudaSuccess) {
        mer::pool_free(dev);
        return truncated ? mer::fail(MER_ERR_INVALID, "volume file truncated") : mer::fail(e == cudaErrorMemoryAllocation ? MER_ERR_OOM : MER_ERR_CUDA, cudaGetErrorString(e));
    }
    *data_dev_out = dev;
    return MER_OK;
}

void apply_override(mer_volume_desc *d, const mer_volume_desc *ov) {
    if (!ov) return;
    /* `min`/`max` properties replace the file's bbox (splinevolume.cpp:93-98, 260-268); toWorld */
    if (ov->bbox_max[0] > ov->bbox_min[0]) {
        for (int i = 0; i < 3; i++) { d->bbox_min[i] = ov->bbox_min[i]; d->bbox_max[i] = ov->bbox_max[i]; }
    }
    d->has_transform = ov->has_transform;
    memcpy(d->world_to_volume, ov->world_to_volume, sizeof(d->world_to_volume));
}

} /* namespace */

extern "C" {

int mer_rif_create_device(int device, const mer_volume_desc *desc, const float *data_dev, int mode, mer_rif **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    int rc = validate_desc(desc, 4);
    if (rc) return rc;
    MER_REQUIRE(mode == MER_RIF_TRICUBIC || mode == MER_RIF_TRILINEAR_PACKED, "unknown RIF mode");
    MER_REQUIRE(data_dev != nullptr, "No RIF data specified!");
    rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    mer_rif *r = new mer_rif();
    memset(r, 0, sizeof(*r));
    r->device = device;
    r->mode = mode;
    r->desc = *desc;
    rc = rif_build(r, data_dev, 0);
    if (rc) { mer_rif_destroy(r); return rc; }
    *out = r;
    return MER_OK;
}

int mer_rif_create(int device, const mer_volume_desc *desc, const float *data, int mode, mer_rif **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    int rc = validate_desc(desc, 4);
    if (rc) return rc;
    MER_REQUIRE(data != nullptr, "No RIF data specified!");
    rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    float *d_data = nullptr;
    const size_t bytes = voxels(desc) * sizeof(float);
    MER_CUDA(mer::pool_malloc((void **) &d_data, bytes));
    cudaError_t e = cudaMemcpy(d_data, data, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { mer::pool_free(d_data); return mer::fail(MER_ERR_CUDA, cudaGetErrorString(e)); }
    rc = mer_rif_create_device(device, desc, d_data, mode, out);
    mer::pool_free(d_data);
    return rc;
}

int mer_rif_create_from_file(int device, const char *vol_path, const mer_volume_desc *override_or_null, int mode,
                             mer_rif **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    MER_REQUIRE(vol_path, "null path");
    mer_volume_desc d;
    float *raw = nullptr;
    int rc = stream_vol_to_device(device, vol_path, &d, &raw);
    if (rc) return rc;
    apply_override(&d, override_or_null);
    rc = mer_rif_create_device(device, &d, raw, mode, out);
    mer::DeviceGuard guard(device);
    mer::pool_free(raw);
    return rc;
}

void mer_rif_destroy(mer_rif *r) {
    if (!r) return;
    mer::DeviceGuard guard(r->device);
    mer::pool_quiesce();
    mer::pool_free(r->d_coeff);
    mer::pool_free(r->d_packed);
    mer::pool_free(r->d_coeff8);
    if (r->tex) cudaDestroyTextureObject(r->tex);
    mer::array_release(r->device, r->texArray, r->texW, r->texH);
    delete r;
}

int mer_rif_coefficients(const mer_rif *r, float *coeff_out_host) {
    MER_REQUIRE(r && coeff_out_host, "null argument");
    mer::DeviceGuard guard(r->device);
    MER_CUDA(cudaMemcpy(coeff_out_host, r->d_coeff, voxels(&r->desc) * sizeof(float), cudaMemcpyDeviceToHost));
    return MER_OK;
}

int mer_rif_desc(const mer_rif *r, mer_volume_desc *out, int *mode_out) {
    MER_REQUIRE(r, "null handle");
    if (out) *out = r->desc;
    if (mode_out) *mode_out = r->mode;
    return MER_OK;
}

int mer_rif_eval_device(const mer_rif *r, int what, size_t n, const float *p_dev, float *value_dev, float *grad_dev,
                        void *stream) {
    MER_REQUIRE(r && (n == 0 || p_dev), "null argument");
    MER_REQUIRE(what >= MER_EVAL_VALUE && what <= MER_EVAL_VALUE_AND_GRADIENT, "bad `what`");
    if (n == 0) return MER_OK;
    float *f = what == MER_EVAL_GRADIENT ? nullptr : value_dev;
    float *g = what == MER_EVAL_VALUE ? nullptr : grad_dev;
    MER_REQUIRE((what == MER_EVAL_GRADIENT || f) && (what == MER_EVAL_VALUE || g), "missing output buffer");
    mer::DeviceGuard guard(r->device);
    const unsigned G = (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u);
    if (r->mode == MER_RIF_TRICUBIC)
        MER_LAUNCH(k_rif_eval<MER_RIF_TRICUBIC>, G, 256, 0, (cudaStream_t) stream, r->dev, n, p_dev, f, g);
    else
        MER_LAUNCH(k_rif_eval<MER_RIF_TRILINEAR_PACKED>, G, 256, 0, (cudaStream_t) stream, r->dev, n, p_dev, f, g);
    return MER_OK;
}

int mer_rif_eval_batch(const mer_rif *r, int what, size_t n, const float *p, float *value_out, float *grad_out) {
    MER_REQUIRE(r && (n == 0 || p), "null argument");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(r->device);
    mer::DevBuf dp, df, dg;
    MER_CUDA(dp.alloc(n * 3 * sizeof(float)));
    MER_CUDA(df.alloc(n * sizeof(float)));
    MER_CUDA(dg.alloc(n * 3 * sizeof(float)));
    MER_CUDA(cudaMemcpy(dp.ptr, p, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    int rc = mer_rif_eval_device(r, what, n, dp.as<float>(), df.as<float>(), dg.as<float>(), nullptr);
    if (rc) return rc;
    MER_CUDA(cudaDeviceSynchronize());
    if (what != MER_EVAL_GRADIENT && value_out) MER_CUDA(cudaMemcpy(value_out, df.ptr, n * sizeof(float), cudaMemcpyDeviceToHost));
    if (what != MER_EVAL_VALUE && grad_out) MER_CUDA(cudaMemcpy(grad_out, dg.ptr, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return MER_OK;
}

int mer_rif_inside_limits_batch(const mer_rif *r, size_t n, const float *p, uint8_t *inside_out) {
    MER_REQUIRE(r && (n == 0 || (p && inside_out)), "null argument");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(r->device);
    mer::DevBuf dp, dout;
    MER_CUDA(dp.alloc(n * 3 * sizeof(float)));
    MER_CUDA(dout.alloc(n));
    MER_CUDA(cudaMemcpy(dp.ptr, p, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    MER_LAUNCH(k_rif_inside, (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u), 256, 0, 0, r->dev, n, dp.as<float>(), dout.as<uint8_t>());
    MER_CUDA(cudaMemcpy(inside_out, dout.ptr, n, cudaMemcpyDeviceToHost));
    return MER_OK;
}

/* ------------------------------------------------------------------ density grid */
static int grid_create_device(int device, const mer_volume_desc *desc, const float *data_dev, int channels, mer_grid **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    int rc = validate_desc(desc, 2);
    if (rc) return rc;
    MER_REQUIRE(channels == 1 || channels == 3, "a grid volume has one channel (density) or three (albedo)");
    MER_REQUIRE(data_dev != nullptr, "No density specified!");
    rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    mer_grid *g = new mer_grid();
    memset(g, 0, sizeof(*g));
    g->device = device;
    g->desc = *desc;
    g->dev.channels = channels;
    const size_t bytes = voxels(desc) * (size_t) channels * sizeof(float);
    cudaError_t e = mer::pool_malloc((void **) &g->d_data, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(g->d_data, data_dev, bytes, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { mer_grid_destroy(g); return mer::fail(MER_ERR_CUDA, cudaGetErrorString(e)); }
    for (int r = 0; r < 3; r++) {
        g->dev.N[r] = desc->res[r];
        /* m_worldToGrid = scale * translate(-min) * worldToVolume, gridvolume.cpp:190-195 */
        float scale = (float) (desc->res[r] - 1) / (desc->bbox_max[r] - desc->bbox_min[r]);
        for (int c = 0; c < 4; c++) {
            float w = desc->has_transform ? desc->world_to_volume[4 * r + c] : (r == c ? 1.0f : 0.0f);
            if (c == 3) w = w + (-desc->bbox_min[r]);
            g->dev.G[4 * r + c] = scale * w;
        }
    }
    g->dev.data = g->d_data;
    /* m_aabb: the data box's 8 corners through volumeToWorld (gridvolume.cpp:199-201) */
    if (!desc->has_transform) {
        for (int i = 0; i < 3; i++) { g->dev.aabbLo[i] = desc->bbox_min[i]; g->dev.aabbHi[i] = desc->bbox_max[i]; }
    } else {
        const float *w = desc->world_to_volume;
        /* inverse of the affine 3x4: R^-1 by cofactors, t' = -R^-1 t */
        double R[9] = {w[0], w[1], w[2], w[4], w[5], w[6], w[8], w[9], w[10]}, t[3] = {w[3], w[7], w[11]};
        double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (det == 0) { mer_grid_destroy(g); return mer::fail(MER_ERR_INVALID, "toWorld is singular"); }
        double I[9] = {(R[4] * R[8] - R[5] * R[7]) / det, (R[2] * R[7] - R[1] * R[8]) / det, (R[1] * R[5] - R[2] * R[4]) / det,
                       (R[5] * R[6] - R[3] * R[8]) / det, (R[0] * R[8] - R[2] * R[6]) / det, (R[2] * R[3] - R[0] * R[5]) / det,
                       (R[3] * R[7] - R[4] * R[6]) / det, (R[1] * R[6] - R[0] * R[7]) / det, (R[0] * R[4] - R[1] * R[3]) / det};
        for (int i = 0; i < 3; i++) { g->dev.aabbLo[i] = INFINITY; g->dev.aabbHi[i] = -INFINITY; }
        for (int c = 0; c < 8; c++) {
            double q[3] = {(c & 1 ? desc->bbox_max[0] : desc->bbox_min[0]) - t[0], (c & 2 ? desc->bbox_max[1] : desc->bbox_min[1]) - t[1],
                           (c & 4 ? desc->bbox_max[2] : desc->bbox_min[2]) - t[2]};
            for (int i = 0; i < 3; i++) {
                float pw = (float) (I[3 * i] * q[0] + I[3 * i + 1] * q[1] + I[3 * i + 2] * q[2]);
                g->dev.aabbLo[i] = std::min(g->dev.aabbLo[i], pw);
                g->dev.aabbHi[i] = std::max(g->dev.aabbHi[i], pw);
            }
        }
    }
    *out = g;
    return MER_OK;
}

int mer_grid_create_device(int device, const mer_volume_desc *desc, const float *data_dev, mer_grid **out) {
    return grid_create_device(device, desc, data_dev, 1, out);
}

static int grid_create_host(int device, const mer_volume_desc *desc, const float *data, int channels, mer_grid **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    int rc = validate_desc(desc, 2);
    if (rc) return rc;
    MER_REQUIRE(data != nullptr, "No density specified!");
    rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    float *d_data = nullptr;
    const size_t bytes = voxels(desc) * (size_t) channels * sizeof(float);
    MER_CUDA(mer::pool_malloc((void **) &d_data, bytes));
    cudaError_t e = cudaMemcpy(d_data, data, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { mer::pool_free(d_data); return mer::fail(MER_ERR_CUDA, cudaGetErrorString(e)); }
    rc = grid_create_device(device, desc, d_data, channels, out);
    mer::pool_free(d_data);
    return rc;
}

int mer_grid_create(int device, const mer_volume_desc *desc, const float *data, mer_grid **out) {
    return grid_create_host(device, desc, data, 1, out);
}

int mer_grid_create_spectrum(int device, const mer_volume_desc *desc, const float *rgb, mer_grid **out) {
    return grid_create_host(device, desc, rgb, 3, out);
}

int mer_grid_channels(const mer_grid *g) { return g ? g->dev.channels : 0; }

int mer_grid_create_from_file(int device, const char *vol_path, const mer_volume_desc *override_or_null,
                              mer_grid **out) {
    if (!out) return mer::fail(MER_ERR_INVALID, "null output handle");
    *out = nullptr;
    MER_REQUIRE(vol_path, "null path");
    mer_volume_desc d;
    float *raw = nullptr;
    int channels = 1;
    int rc = stream_vol_to_device(device, vol_path, &d, &raw, &channels);
    if (rc) return rc;
    apply_override(&d, override_or_null);
    rc = grid_create_device(device, &d, raw, channels, out);
    mer::DeviceGuard guard(device);
    mer::pool_free(raw);
    return rc;
}

void mer_grid_destroy(mer_grid *g) {
    if (!g) return;
    mer::DeviceGuard guard(g->device);
    mer::pool_quiesce();
    mer::pool_free(g->d_data);
    delete g;
}

int mer_grid_lookup_spectrum_batch(const mer_grid *g, size_t n, const float *p, float *rgb_out) {
    MER_REQUIRE(g && (n == 0 || (p && rgb_out)), "null argument");
    MER_REQUIRE(g->dev.channels == 3, "lookupSpectrum needs a 3-channel grid (supportsSpectrumLookups, gridvolume.cpp:579)");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(g->device);
    mer::DevBuf dp, dout;
    MER_CUDA(dp.alloc(n * 3 * sizeof(float)));
    MER_CUDA(dout.alloc(n * 3 * sizeof(float)));
    MER_CUDA(cudaMemcpy(dp.ptr, p, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    MER_LAUNCH(k_grid_lookup3, (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u), 256, 0, 0, g->dev, n, dp.as<float>(), dout.as<float>());
    MER_CUDA(cudaMemcpy(rgb_out, dout.ptr, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return MER_OK;
}

int mer_grid_lookup_batch(const mer_grid *g, size_t n, const float *p, float *value_out) {
    MER_REQUIRE(g && (n == 0 || (p && value_out)), "null argument");
    MER_REQUIRE(g->dev.channels == 1, "lookupFloat needs a single-channel grid (supportsFloatLookups, gridvolume.cpp:578)");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(g->device);
    mer::DevBuf dp, dout;
    MER_CUDA(dp.alloc(n * 3 * sizeof(float)));
    MER_CUDA(dout.alloc(n * sizeof(float)));
    MER_CUDA(cudaMemcpy(dp.ptr, p, n * 3 * sizeof(float), cudaMemcpyHostToDevice));
    MER_LAUNCH(k_grid_lookup, (unsigned) std::min<size_t>(mer_blocks(n, 256), 148u * 8u), 256, 0, 0, g->dev, n, dp.as<float>(), dout.as<float>());
    MER_CUDA(cudaMemcpy(value_out, dout.ptr, n * sizeof(float), cudaMemcpyDeviceToHost));
    return MER_OK;
}

static int grid_woodcock(const mer_grid *g, float scale, size_t n, const float *ro, const float *rd, const float *mint,
                         const float *maxt, uint64_t seed, int evalT, uint8_t *ok, float *t, float *dens, float *tr) {
    MER_REQUIRE(g && (n == 0 || (ro && rd && mint && maxt)), "null argument");
    MER_REQUIRE(g->dev.channels == 1, "the density volume must have one channel (heterogeneous.cpp:269-271)");
    MER_REQUIRE(scale > 0.0f, "scale must be positive");
    if (n == 0) return MER_OK;
    mer::DeviceGuard guard(g->device);
    mer::DevBuf dro, drd, dmi, dma, dt, dd, dtr, dok;
    MER_CUDA(dro.alloc(n * 12)); MER_CUDA(drd.alloc(n * 12)); MER_CUDA(dmi.alloc(n * 4)); MER_CUDA(dma.alloc(n * 4));
    MER_CUDA(dt.alloc(n * 4)); MER_CUDA(dd.alloc(n * 4)); MER_CUDA(dtr.alloc(n * 4)); MER_CUDA(dok.alloc(n));
    MER_CUDA(cudaMemcpy(dro.ptr, ro, n * 12, cudaMemcpyHostToDevice)); MER_CUDA(cudaMemcpy(drd.ptr, rd, n * 12, cudaMemcpyHostToDevice));
    MER_CUDA(cudaMemcpy(dmi.ptr, mint, n * 4, cudaMemcpyHostToDevice)); MER_CUDA(cudaMemcpy(dma.ptr, maxt, n * 4, cudaMemcpyHostToDevice));
    MER_LAUNCH(k_grid_woodcock, (unsigned) std::min<size_t>(mer_blocks(n, 128), 148u * 16u), 128, 0, 0, g->dev, scale, n, dro.as<float>(), drd.as<float>(),
               dmi.as<float>(), dma.as<float>(), (unsigned long long) seed, evalT, dok.as<uint8_t>(), dt.as<float>(), dd.as<float>(), dtr.as<float>());
    if (ok) MER_CUDA(cudaMemcpy(ok, dok.ptr, n, cudaMemcpyDeviceToHost));
    if (t) MER_CUDA(cudaMemcpy(t, dt.ptr, n * 4, cudaMemcpyDeviceToHost));
    if (dens) MER_CUDA(cudaMemcpy(dens, dd.ptr, n * 4, cudaMemcpyDeviceToHost));
    if (tr) MER_CUDA(cudaMemcpy(tr, dtr.ptr, n * 4, cudaMemcpyDeviceToHost));
    return MER_OK;
}

int mer_grid_sample_distance_batch(const mer_grid *g, float scale, size_t n, const float *ray_o, const float *ray_d,
                                   const float *ray_mint, const float *ray_maxt, uint64_t seed, uint8_t *success_out, float *t_out,
                                   float *density_at_t_out) {
    return grid_woodcock(g, scale, n, ray_o, ray_d, ray_mint, ray_maxt, seed, 0, success_out, t_out, density_at_t_out, nullptr);
}

int mer_grid_eval_transmittance_batch(const mer_grid *g, float scale, size_t n, const float *ray_o, const float *ray_d,
                                      const float *ray_mint, const float *ray_maxt, uint64_t seed, float *transmittance_out) {
    return grid_woodcock(g, scale, n, ray_o, ray_d, ray_mint, ray_maxt, seed, 1, nullptr, nullptr, nullptr, transmittance_out);
}

/* ------------------------------------------------------------------ .vol I/O (host only) */
int mer_vol_read_header(const char *path, mer_volume_desc *out, int32_t *encoding, int32_t *channels) {
    MER_REQUIRE(path, "null path");
    return read_vol_file(path, out, nullptr, encoding, channels);
}

int mer_vol_read_data(const char *path, float *data_out, size_t n_floats) {
    MER_REQUIRE(path && data_out, "null argument");
    mer_volume_desc d;
    std::vector<float> data;
    int rc = read_vol_file(path, &d, &data, nullptr, nullptr);
    if (rc) return rc;
    MER_REQUIRE(n_floats == data.size(), "buffer size does not match the volume resolution");
    memcpy(data_out, data.data(), n_floats * sizeof(float));
    return MER_OK;
}

static int vol_write(const char *path, const mer_volume_desc *desc, const float *data, int32_t ch) {
    MER_REQUIRE(path && desc && data, "null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return mer::fail(MER_ERR_INVALID, std::string("cannot create volume file: ") + path);
    unsigned char hdr[48];
    hdr[0] = 'V'; hdr[1] = 'O'; hdr[2] = 'L'; hdr[3] = 3;
    int32_t enc = 1;
    memcpy(hdr + 4, &enc, 4);
    memcpy(hdr + 8, desc->res, 12);
    memcpy(hdr + 20, &ch, 4);
    memcpy(hdr + 24, desc->bbox_min, 12);
    memcpy(hdr + 36, desc->bbox_max, 12);
    size_t total = voxels(desc) * (size_t) ch;
    bool ok = fwrite(hdr, 1, 48, f) == 48 && fwrite(data, sizeof(float), total, f) == total;
    fclose(f);
    if (!ok) return mer::fail(MER_ERR_INVALID, "short write to volume file");
    return MER_OK;
}

int mer_vol_write(const char *path, const mer_volume_desc *desc, const float *data) { return vol_write(path, desc, data, 1); }
int mer_vol_write_spectrum(const char *path, const mer_volume_desc *desc, const float *rgb) { return vol_write(path, desc, rgb, 3); }

} /* extern "C" */
