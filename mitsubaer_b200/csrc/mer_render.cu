/*
 * mer_render.cu — the integrator: a wavefront eikonal volumetric path tracer for sm_100a.
 *
 * What it replaces (paths relative to the MitsubaER tree):
 *   SamplingIntegrator::renderBlock      src/librender/integrator.cpp:140-190  (pixel/sample loop)
 *   PerspectiveCamera::sampleRay         src/sensors/perspective.cpp:126-157, 247-269
 *   VolumetricPathTracer::Li             src/integrators/path/volpath.cpp:84-343 (control flow, RR :326-336)
 *   PathVertex/PathEdge::sampleNext      src/libbidir/vertex.cpp:247-279, edge.cpp:26-103 (curved-walk semantics:
 *                                        wi = normalize(-mRec.d), weight *= refRatioSq, exit along normalize(mRec.d))
 *   Medium::sampleDistance               src/medium/heterogeneousrefractive.cpp:402-568 (free flight, pdfs)
 *   Woodcock tracking                    src/medium/heterogeneous.cpp:613-658 (density grid, along the curve: R2)
 *   ImageBlock::put                      include/mitsuba/render/imageblock.h:124-190
 *   ReconstructionFilter::configure      src/libcore/rfilter.cpp:37-55, src/rfilters/{gaussian,box}.cpp
 *   HDRFilm::develop                     src/films/hdrfilm.cpp:527-540
 *   HSmoothDielectric::sample            src/bsdfs/hdielectric.cpp:115-125, 244-300 (container surface whose eta is the
 *                                        RIF at the hit point; fresnelDielectricExt src/libcore/util.cpp:665-695)
 *
 * Structure: the CPU's recursive per-pixel loop becomes a WAVEFRONT over a pool
 * of path slots.  One launch ("pass") advances every live path by a bounded number of
 * leapfrog steps with the whole path state in registers; paths that finish are replaced
 * in-kernel by fresh camera samples (one atomic on a global sample counter) so lanes stay
 * busy; at the end of the pass the survivors are written out COMPACTED with a warp ballot +
 * one atomicAdd per warp, and the next pass runs over the compacted queue.  Inside a pass
 * the warp alternates between a convergent stepping phase (all lanes execute the same
 * 16xLDG.128 + separable contraction) and an event phase (scatter / exit / regenerate) that
 * is entered only when enough lanes are waiting, which bounds divergence.
 *
 * Next-event estimation along curved paths (desc.direct_connections): a scattering vertex appends a
 * 48-byte request to a device queue; between passes k_nee gives every request a thread that samples a point of the quad
 * emitter, solves the shooting problem of makeDirectConnections (mer_connect.cuh) and splats the contribution.  The
 * solver costs 1e3-1e5 Hessian-carrying steps per request, which is why it is its own kernel and not an event.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mer_internal.h"
#include "mer_connect.cuh"

#ifndef MER_RENDER_MIN_BLOCKS
#define MER_RENDER_MIN_BLOCKS 4 /* 127 registers, no spills: 16 warps per SM cover the gather latency better than 12 at 168 (C2: 49.6 -> 53.9 M samples/s) */
#endif
#ifndef MER_EVENT_MIN_BLOCKS
#define MER_EVENT_MIN_BLOCKS 4 /* the event kernel is latency-bound: 128 registers, 16 warps per SM */
#endif

namespace {

enum PathKind : int {
    K_FULL = 0,  /* next step: +h */
    K_REM = 1,   /* next step: +rem (remainder of trace()) */
    K_BACKF = 2, /* next step: -h   (step back after leaving the shape on a full step) */
    K_BACKR = 3, /* next step: -rem */
    K_ENTRY = 4, /* needs the field at p (camera ray just entered the container) */
    E_BEGIN = 5, /* start a path edge: sample the free-flight distance */
    E_REACHED = 6, /* trace() returned true */
    E_EXIT = 7,  /* trace()/traceTillBoundary() left the shape */
    E_NEW = 8,   /* needs a new camera sample */
    E_SURFACE = 9, /* at the container surface with the field known: dielectric interaction (hdielectric boundary) */
    E_SCATTER = 10, /* real collision accepted, throughput updated: direct connection request, phase sampling, roulette */
    K_DEAD = 11,
    E_FINISH = 12   /* the sample is finished, its splat is pending (inside one call of the event code only) */
};

enum {
    FLAG_TB = 1,
    FLAG_OUTWARD = 4, /* reached the surface from inside */
    FLAG_COVERED = 8, /* the quad's light along the current edge chain is estimated by a direct connection */
    FLAG_PARKED = 16, /* the request queue is full: wait for the next pass */
    FLAG_DRIFTED = 32, /* first kick + drift of the step `kind` are already applied (software-pipelined stepper) */
    FLAG_SLOW = 64     /* K_FULL whose 4x4x4 stencil touches the edge of the grid: stepped by the event code (clamped taps) */
};

enum { ST_SAMPLES = 0, ST_STEPS, ST_SCATTER, ST_NULL, ST_EXIT, ST_NONFINITE, ST_CONN, ST_CONNFAIL, ST_CONNSTEPS, ST_FETCH, ST_COUNT };

#define MER_NEE_SALT 0x5851F42D4C957F2DULL /* direct connections draw from their own Philox key */

/* the path pool: fixed slots, 128 bytes per path, structure of arrays of 16-byte quads (one LDG/STG.128 each) */
struct PathPool {
    float4 *h0; /* HOT (step kernel): p.xyz, v.x */
    float4 *h1; /* v.yz, n(p), grad n(p).x : the field at p carried by the fused stepper */
    float4 *h2; /* grad n(p).yz, distSurf, optical path length */
    uint2 *h3;  /* (int) stepsLeft, kind | flags << 8 */
    float4 *c0; /* COLD (event kernel only): throughput rgb, refStart */
    float4 *c1; /* segDist, sd, product of the surface BSDFs' relative indices (RR), rem */
    float4 *c2; /* origin of the current edge xyz, (int) depth */
    uint4 *c3;  /* rng draw index, pixel, sample index, unused */
};

struct RenderParams {
    MediumDev M;
    float camO[3], camLeft[3], camUp[3], camDir[3];
    float tanHalf, aspect, invW, invH;
    int W, H;
    int sppTotal, sampleBegin, sampleStride, sppLocal;
    unsigned long long seed, totalSamples;
    int maxDepth, rrDepth;
    float env[3];
    int hasQuad;
    float quadO[3], quadU[3], quadV[3], quadLe[3];
    float filterRadius, filterScale, filterValues[32];
    int frames, channels, calibrated; /* transient film: 3 * frames + 2 channels per pixel (bdpt_wr.cpp:52-56) */
    float minBound, binWidth;
    int modulation;                   /* continuous-wave ToF: contributions times correlationFunction(path length) */
    float lambda, phaseShift;         /* phaseShift = phase * lambda / (2 pi), pathlengthsampler.cpp:74 */
    int stepsPerPass;  /* steps of one visit of the step kernel to a path */
    int refillGate;    /* lanes of a warp that must be idle before the warp pops new slots */
    float *film;
    PathPool pool;
    unsigned nSlots;
    unsigned *stepHead; /* next slot of the round to hand out (reset per round) */
    unsigned *live;     /* slots that are not dead after the event kernel (reset per round) */
    unsigned long long *sampleCounter;
    unsigned long long *stats;
    /* direct connections */
    int nee, neePrecision, neeMaxIterations, neeStraightFirst;
    int neeMis; /* power heuristic between phase sampling and the curved connection (volpath.cpp:120-147, 164-173, 430-433) */
    float neeTol2, neeRRWeight;
    unsigned neeCap;
    unsigned *neeCount;
    float4 *neeQ0, *neeQ1; /* (p1.xyz, wi.x), (wi.yz, thr.rg) */
    uint4 *neeQ2;          /* thr.b, depth, pixel, sample */
    float *neeQ3;          /* optical path length at the vertex (transient film) */
    float *neeQ4;          /* HIT requests (neeMis): the pdf the walk sampled the emitted direction with */
    /* light tracing */
    int lightMode, emitterType;
    float beamO[3], beamD[3], beamPower[3];
    float lightScale;      /* 1 / (number of light paths * pixel area on the image plane at unit distance) */
    unsigned *neePerm;     /* requests ordered by expected length (null: queue order) */
    unsigned char *neeKey;
    unsigned *neeHist;     /* NEE_BINS counters, then offsets */
};

#define NEE_BINS 64
#define NEE_HIT 0x80000000u /* request flag in the depth word: a phase-sampled path that HIT the quad asks for its MIS weight */

struct Lane {
    float3 p, v;
    float3 o; /* ray.o of the current path edge ("no forward progress" test, heterogeneousrefractive.cpp:517-520) */
    float n;
    float3 G;
    float thr[3];
    float refStart, segDist, distSurf, rem, sd, etaPath, opl;
    float safe; /* MER_SHAPE_SDF: distance the ray may still move before the containment test needs a lookup (not persisted) */
    float phasePdf; /* pdf of the direction sampled at the last scattering vertex (MIS of emitter hits) */
    int stepsLeft, depth, kind, flags;
    PathRng rng;
    unsigned pixel, sample; /* sample id = pixel * sppTotal + sample */
};

__device__ __forceinline__ bool intersect_box(const float *box, float3 o, float3 d, float &tNear, float &tFar) {
    float t0 = 0.0f, t1 = INFINITY;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float inv = 1.0f / dd[i];
        float ta = (box[i] - oo[i]) * inv, tb = (box[3 + i] - oo[i]) * inv;
        float lo = ta, hi = tb;
        if (ta > tb) { lo = tb; hi = ta; }
        t0 = fmaxf(t0, lo);
        t1 = fminf(t1, hi);
    }
    tNear = t0;
    tFar = t1;
    return t0 < t1;
}

template <bool SDFSHAPE>
__device__ __forceinline__ bool intersect_shape(const MediumDev &M, float3 o, float3 d, float &tNear) {
    if (SDFSHAPE) {
        /* sphere tracing from where the ray enters the bounding box until the signed distance turns negative: the entry
         * point is inside the shape by at most MER_SDF_TRACE_EPS */
        float t, tFar;
        if (!intersect_box(M.shape, o, d, t, tFar)) return false;
        for (int i = 0; i < MER_SDF_TRACE_STEPS && t <= tFar; i++) {
            const float v = sdf_value(M.sdf, f3(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z));
            if (v < 0.0f) { tNear = t; return true; }
            t += fmaxf(v, MER_SDF_TRACE_EPS);
        }
        return false;
    }
    if (M.shapeType == MER_SHAPE_SPHERE) {
        float3 oc = f3(o.x - M.shape[0], o.y - M.shape[1], o.z - M.shape[2]);
        float b = dot3(oc, d), c = dot3(oc, oc) - M.shape[3] * M.shape[3];
        float disc = b * b - c;
        if (disc <= 0.0f) return false;
        float sq = sqrtf(disc), t0 = -b - sq, t1 = -b + sq;
        if (t1 <= 0.0f) return false;
        tNear = fmaxf(t0, 0.0f);
        return true;
    }
    float t0 = 0.0f, t1 = INFINITY;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float inv = 1.0f / dd[i];
        float ta = (M.shape[i] - oo[i]) * inv, tb = (M.shape[3 + i] - oo[i]) * inv;
        float lo = ta, hi = tb;
        if (ta > tb) { lo = tb; hi = ta; }
        t0 = fmaxf(t0, lo);
        t1 = fminf(t1, hi);
    }
    if (!(t0 < t1)) return false;
    tNear = t0;
    return true;
}

__device__ __forceinline__ bool intersect_quad(const RenderParams &P, float3 o, float3 d, float &t) {
    if (!P.hasQuad) return false;
    float3 u = f3(P.quadU[0], P.quadU[1], P.quadU[2]), v = f3(P.quadV[0], P.quadV[1], P.quadV[2]);
    float3 nrm = f3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
    float denom = dot3(d, nrm);
    if (denom == 0.0f) return false;
    float3 w = f3(P.quadO[0] - o.x, P.quadO[1] - o.y, P.quadO[2] - o.z);
    t = dot3(w, nrm) / denom;
    if (!(t > 0.0f)) return false;
    float3 q = f3(o.x + t * d.x - P.quadO[0], o.y + t * d.y - P.quadO[1], o.z + t * d.z - P.quadO[2]);
    float a = dot3(q, u) / dot3(u, u), b = dot3(q, v) / dot3(v, v);
    return a >= 0.0f && a <= 1.0f && b >= 0.0f && b <= 1.0f;
}

/* HSmoothDielectric::sample (hdielectric.cpp:244-300) in ERadiance mode with both components: d is the unit
 * direction of travel, N the outward normal, eta the RIF at the hit point.  Returns true for transmission. */
__device__ __forceinline__ bool hdielectric_sample(float3 d, float3 N, float eta, float u, bool radianceMode, float3 &dOut, float &weight,
                                                   float &etaScale) {
    const float wiN = -dot3(d, N); /* Frame::cosTheta(wi), wi = -d */
    float cosThetaT;
    const float F = fresnel_dielectric_ext(wiN, cosThetaT, eta);
    if (u <= F) {
        dOut = f3(d.x + 2.0f * wiN * N.x, d.y + 2.0f * wiN * N.y, d.z + 2.0f * wiN * N.z);
        weight = 1.0f;
        etaScale = 1.0f;
        return false;
    }
    const float invEta = __fdiv_rn(1.0f, eta), scale = -(cosThetaT < 0.0f ? invEta : eta);
    dOut = f3(scale * (-d.x - wiN * N.x) + cosThetaT * N.x, scale * (-d.y - wiN * N.y) + cosThetaT * N.y,
              scale * (-d.z - wiN * N.z) + cosThetaT * N.z);
    const float factor = radianceMode ? (cosThetaT < 0.0f ? invEta : eta) : 1.0f; /* radiance scaling only in ERadiance mode, :262-268 */
    weight = __fmul_rn(factor, factor);
    etaScale = cosThetaT < 0.0f ? eta : invEta;
    return true;
}

/* ImageBlock::put (imageblock.h:144-190) onto the global film with red.global.add.f32 */
/* frame of the transient film a path of optical length `len` falls into (bdpt_proc.cpp:446-449), -1: none */
__device__ __forceinline__ int path_frame(const RenderParams &P, float len) {
    if (P.frames <= 1) return 0;
    const float b = floorf((len - P.minBound) / P.binWidth);
    return (b >= 0.0f && b < (float) P.frames) ? (int) b : -1;
}

/* PathLengthSampler::correlationFunction, src/librender/pathlengthsampler.cpp:66-96; 1 without a modulation.  An emitter at
 * infinity (the environment) has no path length: it contributes nothing to a modulated image. */
__device__ __forceinline__ float path_weight(const RenderParams &P, float len) {
    if (P.modulation == MER_MODULATION_NONE) return 1.0f;
    if (!isfinite(len)) return 0.0f;
    float pl = len + P.phaseShift;
    if (P.modulation == MER_MODULATION_SINE) return cosf(pl * 6.283185307179586f / P.lambda);
    if (P.modulation == MER_MODULATION_SQUARE) return 4.0f / P.lambda * (fabsf(fmodf(pl, P.lambda) - P.lambda / 2.0f) - P.lambda / 4.0f);
    pl = fmodf(pl, P.lambda); /* Hamiltonian code */
    if (pl < P.lambda / 6.0f) return 6.0f * pl / P.lambda;
    if (pl < P.lambda / 2.0f) return 1.0f;
    if (pl < 2.0f * P.lambda / 3.0f) return 1.0f - (pl - P.lambda / 2.0f) * 6.0f / P.lambda;
    return 0.0f;
}

/* per-thread event counters (registers; reduced per warp at the end of the event kernel) */
struct EvStats { unsigned v[ST_COUNT]; };
#define ST_INC(st, i) ((st).v[i]++)
/* returns true when the sample is dropped because a value is not finite (imageblock.h:147-152) */
/* film accumulation: red.global.add.f32, no return value and no generic-address dispatch (a plain atomicAdd on the film
 * pointer compiled to ATOM.E.ADD.F32 plus a shared-memory CAS-spin fallback: the compiler cannot see that it is global) */
__device__ __forceinline__ void film_red(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __noinline__ bool film_put(const RenderParams &P, float sx, float sy, const float L[3], float alpha, float wgt, int frame) {
    const float value[5] = {L[0], L[1], L[2], alpha, wgt};
#pragma unroll
    for (int k = 0; k < 5; k++)
        if (!isfinite(value[k])) return true;
    const float px = sx - 0.5f, py = sy - 0.5f, r = P.filterRadius;
    const int x0 = max((int) ceilf(px - r), 0), y0 = max((int) ceilf(py - r), 0),
              x1 = min((int) floorf(px + r), P.W - 1), y1 = min((int) floorf(py + r), P.H - 1);
    for (int y = y0; y <= y1; ++y) {
        const float wy = P.filterValues[min((int) fabsf(((float) y - py) * P.filterScale), 31)];
        for (int x = x0; x <= x1; ++x) {
            const float w = P.filterValues[min((int) fabsf(((float) x - px) * P.filterScale), 31)] * wy;
            float *dest = P.film + ((size_t) y * P.W + x) * (size_t) P.channels;
            if (frame >= 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) film_red(dest + 3 * frame + k, w * value[k]);
            }
            film_red(dest + P.channels - 2, w * value[3]);
            film_red(dest + P.channels - 1, w * value[4]);
        }
    }
    return false;
}

/* the sample's film position is draw 0/1 of its Philox stream: recomputed, not stored */
__device__ __forceinline__ void sample_position(const RenderParams &P, unsigned pixel, unsigned sample, float &sx, float &sy) {
    const unsigned yy = pixel / (unsigned) P.W;
    int x = (int) (pixel - yy * (unsigned) P.W), y = (int) yy;
    const unsigned long long id = (unsigned long long) pixel * (unsigned long long) P.sppTotal + sample;
    uint4 b = philox4x32_10((uint32_t) id, (uint32_t) (id >> 32), 0u, 0u, (uint32_t) P.seed, (uint32_t) (P.seed >> 32));
    sx = (float) x + (float) (b.x >> 8) * (1.0f / 16777216.0f);
    sy = (float) y + (float) (b.y >> 8) * (1.0f / 16777216.0f);
}

/* EXTRAS = any of: transient film, direct connections, light tracing.  The plain camera walk is compiled without them. */
template <bool EXTRAS>
__device__ __forceinline__ void finish_sample(const RenderParams &P, Lane &L, const float rad[3], float alpha,
                                              EvStats &st, float pathLength = INFINITY) {
    L.kind = E_NEW;
    if (EXTRAS && P.lightMode) return; /* a light path that leaves or dies deposits nothing: only its connections do */
    float sx, sy;
    sample_position(P, L.pixel, L.sample, sx, sy);
    float out[3] = {rad[0], rad[1], rad[2]};
    int frame = 0;
    if (EXTRAS && P.modulation) {
        const float w = path_weight(P, pathLength);
        out[0] *= w; out[1] *= w; out[2] *= w;
    } else if (EXTRAS) {
        frame = path_frame(P, pathLength);
    }
    if (film_put(P, sx, sy, out, alpha, 1.0f, frame)) ST_INC(st, ST_NONFINITE);
}

__device__ __forceinline__ void begin_trace(const RenderParams &P, Lane &L, float dist) {
    L.segDist = dist;
    L.distSurf = 0.0f;
    L.flags &= ~(FLAG_TB | FLAG_DRIFTED | FLAG_SLOW);
    if (isfinite(dist)) {
        trace_split(dist, P.M.h, L.stepsLeft, L.rem);
        L.kind = L.stepsLeft > 0 ? K_FULL : K_REM;
    } else { /* traceTillBoundary, heterogeneousrefractive.cpp:742-776 */
        L.flags |= FLAG_TB;
        L.stepsLeft = 100000;
        L.rem = 0.0f;
        L.kind = K_FULL;
    }
}

/* K_FULL +h, K_REM +rem, K_BACKF -h, K_BACKR -rem, K_ENTRY 0: selects, no branches (bit 0 = remainder, bit 1 = backwards) */
__device__ __forceinline__ float step_length(int kind, float h, float rem) {
    const float mag = (kind & 1) ? rem : h;
    const float sgn = (kind & 2) ? -mag : mag;
    return kind == K_ENTRY ? 0.0f : sgn;
}

/* first half of er_step (:655-657): kick, drift with n at the OLD point, optical length.  Rounded operation by operation
 * like the reference's float build (see er_step_fused). */
template <bool EXTRAS, typename T>
__device__ __forceinline__ void lane_drift(T &L, float hc) {
    const float hs = __fmul_rn(0.5f, hc);
    L.v = f3(__fadd_rn(L.v.x, __fmul_rn(hs, L.G.x)), __fadd_rn(L.v.y, __fmul_rn(hs, L.G.y)), __fadd_rn(L.v.z, __fmul_rn(hs, L.G.z)));
    const float recip = __frcp_rn(L.n);
    L.p = f3(__fadd_rn(L.p.x, __fmul_rn(__fmul_rn(hc, L.v.x), recip)), __fadd_rn(L.p.y, __fmul_rn(__fmul_rn(hc, L.v.y), recip)),
             __fadd_rn(L.p.z, __fmul_rn(__fmul_rn(hc, L.v.z), recip)));
    if (EXTRAS) L.opl = __fadd_rn(L.opl, __fmul_rn(hc, L.n));
    L.flags |= FLAG_DRIFTED;
}
template <typename T> __device__ __forceinline__ void lane_kick(T &L, float hc) { /* second half of er_step (:659-660) */
    const float hs = __fmul_rn(0.5f, hc);
    L.v = f3(__fadd_rn(L.v.x, __fmul_rn(hs, L.G.x)), __fadd_rn(L.v.y, __fmul_rn(hs, L.G.y)), __fadd_rn(L.v.z, __fmul_rn(hs, L.G.z)));
}

/* exponential free-flight pdfs + transmittance at geometric length d (:533-562) */
__device__ __noinline__ void edge_weight(const MediumDev &M, float sd, float d, bool success, float edge[3]) {
    float pdfFailure = 0.0f, pdfSuccess = 0.0f;
    if (M.strategy == MER_STRATEGY_MAXIMUM) { /* :534-536; the density MaxExpDist::sample reported for d is MaxExpDist::pdf(d) */
        pdfFailure = __fsub_rn(1.0f, maxexp_cdf(M, d));
        pdfSuccess = success ? maxexp_pdf(M, d) : 0.0f;
    } else if (M.strategy == MER_STRATEGY_BALANCE) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float tmp = fastexp_dev(__fmul_rn(-M.sigmaT[c], d));
            pdfFailure = __fadd_rn(pdfFailure, tmp);
            pdfSuccess = __fadd_rn(pdfSuccess, __fmul_rn(M.sigmaT[c], tmp));
        }
        pdfFailure = __fdiv_rn(pdfFailure, 3.0f);
        pdfSuccess = __fdiv_rn(pdfSuccess, 3.0f);
    } else {
        pdfFailure = fastexp_dev(__fmul_rn(-sd, d));
        pdfSuccess = __fmul_rn(sd, pdfFailure);
    }
    float T[3], tmax = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        T[c] = fastexp_dev(__fmul_rn(M.sigmaT[c], -d));
        tmax = fmaxf(tmax, T[c]);
    }
    if (tmax < 1e-20f) T[0] = T[1] = T[2] = 0.0f;
    const float ps = __fmul_rn(pdfSuccess, M.weight), pf = __fadd_rn(__fmul_rn(M.weight, pdfFailure), 1.0f - M.weight);
#pragma unroll
    for (int c = 0; c < 3; c++)
        edge[c] = success ? __fdiv_rn(__fmul_rn(M.sigmaS[c], T[c]), ps) : __fdiv_rn(T[c], pf);
}

/* MIS (desc.direct_connections = 2): a phase-sampled path that reaches the quad after a vertex whose direct connection
 * was requested does not count the quad with weight 0 (round 1) but asks k_nee for its power-heuristic weight: the
 * connection vertex -> hit point is solved like any other, its Jacobian gives the solid-angle density p_nee the
 * next-event estimator would have had for that point, and the hit is splatted with p_phase^2 / (p_phase^2 + p_nee^2)
 * (volpath.cpp:164-173, miWeight :430-433).  false: the request queue is full (the caller parks the path). */
__device__ __forceinline__ bool nee_hit_request(const RenderParams &P, const Lane &L, float3 vertex, float3 y, const float rad[3], float length) {
    const unsigned slot = atomicAdd(P.neeCount, 1u);
    if (slot >= P.neeCap) return false;
    P.neeQ0[slot] = make_float4(vertex.x, vertex.y, vertex.z, y.x);
    P.neeQ1[slot] = make_float4(y.y, y.z, rad[0], rad[1]);
    P.neeQ2[slot] = make_uint4(__float_as_uint(rad[2]), (unsigned) L.depth | NEE_HIT, L.pixel, L.sample);
    P.neeQ3[slot] = length;
    P.neeQ4[slot] = L.phasePdf;
    return true;
}

/* Everything that is not a plain full step, ONE PASS over the kinds of event in the order a path runs through them:
 *     A  one er_step with a direct lookup (remainder step, step back, entry, edge-of-grid steps)
 *     R  end of a path edge: Woodcock test / edge weights; null collision -> next flight; exit -> radiance or surface
 *     C  scattering vertex: direct-connection request, phase sampling, Russian roulette
 *     S  dielectric container surface
 *     B  start of a path edge: free-flight distance
 *     F  film splat of a finished sample
 *     N  next camera sample (or emitter sample in light-tracing mode)
 * A collision runs A R [C] B in one call, an exit A R F N; what is left (the entry step of the new sample, a step back
 * after a remainder step that left the shape) waits for the next round's call.  One pass instead of a loop over "the
 * lane's current kind": every block is executed at most once per call by all the lanes that need it (the loop ran the
 * blocks in whatever order the lanes' kinds came up, 10 of 32 lanes: ncu r02z), and there is one splat site.
 * DIELECTRIC selects the container surface: false = index-matched null surface, true = hdielectric. */
template <int MODE, bool DIELECTRIC, bool EXTRAS, bool SDFSHAPE, bool XFORM>
__device__ __forceinline__ void handle_events(const RenderParams &P, Lane &L, EvStats &st) {
    const MediumDev &M = P.M;
    L.rng.cachedBlock = 0xffffffffu; /* the cached Philox block lives only inside one call (registers) */
    float frad[3] = {0.f, 0.f, 0.f}, falpha = 1.0f, flen = INFINITY; /* the finished sample's splat (kind E_FINISH) */
#define MER_FINISH(r0_, r1_, r2_, a_, len_) do { frad[0] = (r0_); frad[1] = (r1_); frad[2] = (r2_); falpha = (a_); flen = (len_); L.kind = E_FINISH; } while (0)

    /* ---- A: one er_step outside the step kernel (:653-661 inside trace() :674-686 / traceTillBoundary() :748-762): the
     * remainder step, the step back after leaving the shape, the zero-length step that fetches the field where a ray
     * enters the container, and full steps whose stencil touches the edge of the grid.  A direct 64-tap lookup (no
     * cached block), same arithmetic as the step kernel's. */
    if (L.kind <= K_ENTRY && (L.kind != K_FULL || (L.flags & FLAG_SLOW))) {
        const int kind = L.kind;
        const float hc = step_length(kind, M.h, L.rem);
        if (!(L.flags & FLAG_DRIFTED)) lane_drift<EXTRAS>(L, hc);
        L.flags &= ~(FLAG_DRIFTED | FLAG_SLOW);
        rif_lookup<MODE>(M.rif, L.p, L.n, L.G);
        lane_kick(L, hc);
        const bool inside = SDFSHAPE ? inside_shape_lazy(M, L.p, fabsf(hc), L.safe) : inside_shape(M, L.p);
        if (kind != K_ENTRY) ST_INC(st, ST_STEPS);
        ST_INC(st, ST_FETCH);
        if (kind == K_FULL) {
            if (inside) {
                L.distSurf += M.h;
                L.stepsLeft--;
                L.kind = L.stepsLeft > 0 ? K_FULL : ((L.flags & FLAG_TB) ? E_EXIT : K_REM);
                if (L.kind == K_FULL) { /* hand it back drifted if the next stencil is interior again */
                    lane_drift<EXTRAS>(L, M.h);
                    const CellPos cn = rif_cell<MODE, XFORM>(M.rif, L.p);
                    if (!rif_cell_fast<MODE>(M.rif, cn)) L.flags |= FLAG_SLOW;
                }
            } else {
                L.kind = K_BACKF;
            }
        } else if (kind == K_REM) {
            if (inside) L.distSurf += L.rem;
            L.kind = inside ? E_REACHED : K_BACKR;
        } else if (kind == K_ENTRY) {
            L.kind = DIELECTRIC ? E_SURFACE : E_BEGIN;
        } else {
            if (kind == K_BACKF && (L.flags & FLAG_TB)) L.distSurf -= M.h; /* :761 */
            L.kind = E_EXIT;
        }
    }

    /* ---- R: end of a path edge */
    if (L.kind == E_REACHED || L.kind == E_EXIT) {
        bool scatter = false, again = false;
        float edge[3] = {1.0f, 1.0f, 1.0f};
        if (L.kind == E_REACHED) {
            if (M.hasGrid) {
                /* Woodcock acceptance, heterogeneous.cpp:631-644 */
                float densityAtT = __fmul_rn(grid_lookup(M.grid, L.p), M.densityScale);
                if (__fmul_rn(densityAtT, M.invMaxDensity) > L.rng.next()) {
                    scatter = true;
                    /* sigmaS = albedo * densityAtT * scale over a transmittance of 1 / densityAtT: the edge weighs by the
                     * albedo, a constant or the `albedo` volume at the event (heterogeneous.cpp:646-649) */
                    if (M.hasAlbedoGrid) grid_lookup3(M.albedoGrid, L.p, edge);
                    else { edge[0] = M.albedo[0]; edge[1] = M.albedo[1]; edge[2] = M.albedo[2]; }
                } else {
                    ST_INC(st, ST_NULL);
                    begin_trace(P, L, __fmul_rn(-fastlog_dev(1.0f - L.rng.next()), M.invMaxDensity));
                    again = true;
                }
            } else if (L.p.x == L.o.x && L.p.y == L.o.y && L.p.z == L.o.z) { /* no forward progress, :517-520 */
                MER_FINISH(0.f, 0.f, 0.f, 1.0f, INFINITY);
                again = true;
            } else {
                scatter = true;
                edge_weight(M, L.sd, L.segDist, true, edge);
            }
        } else if (!M.hasGrid) {
            edge_weight(M, L.sd, L.distSurf, false, edge);
        }
        if (!again) {
            float rrs = (float) (1.0 / (double) (L.refStart * L.refStart)); /* :469 */
            rrs *= L.n * L.n;                                                /* :501, refEnd = n(p) */
            if (M.physicalScaling) rrs = 1.0f / rrs;
            if (EXTRAS && P.lightMode) rrs = 1.0f; /* weight[EImportance] carries no refRatioSq, edge.cpp:96-98 */
            if (scatter) {
                ST_INC(st, ST_SCATTER);
                if (P.maxDepth != -1 && L.depth >= P.maxDepth) {
                    MER_FINISH(0.f, 0.f, 0.f, 1.0f, INFINITY);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) L.thr[c] *= edge[c] * rrs;
                    L.kind = E_SCATTER;
                }
            } else {
                const float vinv = 1.0f / sqrtf(dot3(L.v, L.v));
                const float3 d = f3(L.v.x * vinv, L.v.y * vinv, L.v.z * vinv);
                const float thrNew[3] = {L.thr[0] * (edge[0] * rrs), L.thr[1] * (edge[1] * rrs), L.thr[2] * (edge[2] * rrs)};
                const bool deep = P.maxDepth != -1 && L.depth >= P.maxDepth;
                float tq = 0.0f;
                const bool hitsQuad = !DIELECTRIC && !deep && intersect_quad(P, L.p, d, tq);
                const bool dark = hitsQuad && EXTRAS && (L.flags & FLAG_COVERED);
                bool parked = false;
                if (EXTRAS && dark && P.neeMis) { /* the hit's share under the power heuristic comes from k_nee */
                    const float rad[3] = {thrNew[0] * P.quadLe[0], thrNew[1] * P.quadLe[1], thrNew[2] * P.quadLe[2]};
                    parked = !nee_hit_request(P, L, L.o, f3(L.p.x + tq * d.x, L.p.y + tq * d.y, L.p.z + tq * d.z), rad, L.opl + tq);
                }
                if (parked) { /* nothing has been committed: the exit is taken again next round */
                    L.flags |= FLAG_PARKED;
                } else {
                    ST_INC(st, ST_EXIT);
                    L.thr[0] = thrNew[0]; L.thr[1] = thrNew[1]; L.thr[2] = thrNew[2];
                    if (deep) {
                        MER_FINISH(0.f, 0.f, 0.f, 1.0f, INFINITY);
                    } else if (DIELECTRIC) { /* move to the surface point and fetch the field there, then E_SURFACE */
                        const float te = SDFSHAPE ? exit_distance_sdf(M, L.p, d) : exit_distance(M, L.p, d);
                        L.p = f3(L.p.x + te * d.x, L.p.y + te * d.y, L.p.z + te * d.z);
                        L.safe = 0.0f;
                        L.v = d;
                        L.flags |= FLAG_OUTWARD;
                        L.kind = K_ENTRY;
                    } else {
                        L.depth++;
                        const float *Le = hitsQuad ? P.quadLe : P.env;
                        const float k = dark ? 0.0f : 1.0f;
                        MER_FINISH(k * L.thr[0] * Le[0], k * L.thr[1] * Le[1], k * L.thr[2] * Le[2], 1.0f, hitsQuad ? L.opl + tq : INFINITY);
                    }
                }
            }
        }
    }

    /* ---- C: scattering vertex; L.v = arrival velocity, throughput already carries the edge */
    if (L.kind == E_SCATTER) {
        const float vinv = 1.0f / sqrtf(dot3(L.v, L.v));
        const float3 wi = f3(-L.v.x * vinv, -L.v.y * vinv, -L.v.z * vinv);
        bool parked = false;
        if (EXTRAS && P.nee) {
            if (P.maxDepth == -1 || L.depth + 1 < P.maxDepth) {
                const unsigned slot = atomicAdd(P.neeCount, 1u);
                if (slot >= P.neeCap) { /* the host clamps the count; retried next round */
                    L.flags |= FLAG_PARKED;
                    parked = true;
                } else {
                    P.neeQ0[slot] = make_float4(L.p.x, L.p.y, L.p.z, wi.x);
                    P.neeQ1[slot] = make_float4(wi.y, wi.z, L.thr[0], L.thr[1]);
                    P.neeQ2[slot] = make_uint4(__float_as_uint(L.thr[2]), (unsigned) L.depth, L.pixel, L.sample);
                    P.neeQ3[slot] = L.opl;
                }
            }
            if (!parked) L.flags |= FLAG_COVERED;
        }
        if (!parked) { /* phase sampling (wi = normalize(-mRec.d)) and the Russian roulette of volpath.cpp:326-336 */
            const float u1 = L.rng.next(), u2 = L.rng.next();
            L.v = hg_sample_dev(M.g, wi, u1, u2);
            if (EXTRAS && P.neeMis) L.phasePdf = hg_eval_dev(M.g, wi, L.v); /* HGPhaseFunction::sample's pdf (hg.cpp:100-105) */
            L.kind = E_BEGIN; /* field at p is still valid */
            if (L.depth++ >= P.rrDepth) {
                float q = fmaxf(L.thr[0], fmaxf(L.thr[1], L.thr[2]));
                if (DIELECTRIC) q *= L.etaPath * L.etaPath;
                q = fminf(q, 0.95f);
                if (L.rng.next() >= q) {
                    MER_FINISH(0.f, 0.f, 0.f, 1.0f, INFINITY);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) L.thr[c] /= q;
                }
            }
        }
    }

    /* ---- S: container surface with a dielectric BSDF; L.v = unit direction of travel, L.n = RIF at the hit point */
    if (DIELECTRIC && L.kind == E_SURFACE) {
        /* MER_SHAPE_SDF: normalised gradient of the signed distance; box / sphere: analytic */
        const float3 N = SDFSHAPE ? merc::container_normal(M, L.p) : shape_normal(M, L.p);
        const unsigned rngBefore = L.rng.k;
        const float u = L.rng.next();
        L.rng.next(); /* the BSDF sample is a Point2 */
        float3 dOut;
        float w, es;
        const bool transmitted = hdielectric_sample(L.v, N, L.n, u, !(EXTRAS && P.lightMode), dOut, w, es);
        const bool inside = (L.flags & FLAG_OUTWARD) ? !transmitted : transmitted;
        float tq = 0.0f;
        const bool hitsQuad = !inside && intersect_quad(P, L.p, dOut, tq);
        const bool dark = hitsQuad && EXTRAS && (L.flags & FLAG_COVERED);
        bool parked = false;
        if (EXTRAS && dark && P.neeMis) { /* the refracted exit of a covered chain reached the quad: MIS weight from k_nee */
            const float rad[3] = {L.thr[0] * w * P.quadLe[0], L.thr[1] * w * P.quadLe[1], L.thr[2] * w * P.quadLe[2]};
            parked = !nee_hit_request(P, L, L.o, f3(L.p.x + tq * dOut.x, L.p.y + tq * dOut.y, L.p.z + tq * dOut.z), rad, L.opl + tq);
        }
        if (parked) { /* nothing committed, the draws are given back: the surface event is taken again next round */
            L.rng.k = rngBefore;
            L.flags |= FLAG_PARKED;
        } else {
#pragma unroll
        for (int c = 0; c < 3; c++) L.thr[c] *= w;
        L.etaPath *= es;
        L.v = dOut;
        L.flags &= ~FLAG_OUTWARD;
        if (!inside) { /* rayIntersectAndLookForEmitter with a delta BSDF: MIS weight 1, volpath.cpp:300-320 */
            const float *Le = hitsQuad ? P.quadLe : P.env;
            const float k = dark ? 0.0f : 1.0f;
            MER_FINISH(k * L.thr[0] * Le[0], k * L.thr[1] * Le[1], k * L.thr[2] * Le[2], 1.0f, hitsQuad ? L.opl + tq : INFINITY);
        } else {
            L.flags &= ~FLAG_COVERED; /* an internal reflection starts a chain no direct connection accounts for */
            L.kind = E_BEGIN;
            if (L.depth++ >= P.rrDepth) { /* volpath.cpp:326-336 */
                const float q = fminf(fmaxf(L.thr[0], fmaxf(L.thr[1], L.thr[2])) * L.etaPath * L.etaPath, 0.95f);
                if (L.rng.next() >= q) {
                    MER_FINISH(0.f, 0.f, 0.f, 1.0f, INFINITY);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; c++) L.thr[c] /= q;
                }
            }
        }
        }
    }

    /* ---- B: Medium::sampleDistance prologue, heterogeneousrefractive.cpp:402-475; L.v = unit direction */
    if (L.kind == E_BEGIN) {
        if (!rif_inside_limits(M.rif, L.p)) {
            MER_FINISH(0.f, 0.f, 0.f, 1.0f, INFINITY);
        } else {
            L.refStart = L.n;
            L.o = L.p;
            L.v = f3(L.v.x * L.n, L.v.y * L.n, L.v.z * L.n);
            float dist;
            if (M.hasGrid) {
                dist = __fmul_rn(-fastlog_dev(1.0f - L.rng.next()), M.invMaxDensity);
            } else {
                float rnd = L.rng.next();
                L.sd = M.samplingDensity;
                if (rnd < M.weight) {
                    rnd = __fdiv_rn(rnd, M.weight);
                    if (M.strategy == MER_STRATEGY_MAXIMUM) {
                        float pdfUnused;
                        dist = maxexp_sample(M, 1.0f - rnd, pdfUnused);
                    } else {
                        if (M.strategy == MER_STRATEGY_BALANCE) L.sd = M.sigmaT[min((int) (L.rng.next() * 3.0f), 2)];
                        dist = __fdiv_rn(-fastlog_dev(1.0f - rnd), L.sd);
                    }
                } else {
                    dist = INFINITY;
                }
            }
            begin_trace(P, L, dist);
        }
    }

    /* ---- F: the one splat site */
    if (L.kind == E_FINISH) finish_sample<EXTRAS>(P, L, frad, falpha, st, flen);
#undef MER_FINISH

    /* ---- N: SamplingIntegrator::renderBlock: next (pixel, sample); a few samples that never reach the medium are
     * finished on the spot */
    for (int attempt = 0; attempt < 4 && L.kind == E_NEW; attempt++) {
        unsigned long long g = atomicAdd(P.sampleCounter, 1ULL);
        if (g >= P.totalSamples) { L.kind = K_DEAD; break; }
        ST_INC(st, ST_SAMPLES);
        unsigned pixel, k;
        if (P.totalSamples <= 0xffffffffULL) { /* warp-uniform: 32-bit division in the common case */
            pixel = (unsigned) g / (unsigned) P.sppLocal;
            k = (unsigned) g - pixel * (unsigned) P.sppLocal;
        } else {
            pixel = (unsigned) (g / (unsigned long long) P.sppLocal);
            k = (unsigned) (g - (unsigned long long) pixel * (unsigned long long) P.sppLocal);
        }
        L.pixel = pixel;
        L.sample = (unsigned) P.sampleBegin + k * (unsigned) P.sampleStride;
        L.rng.init(P.seed, (unsigned long long) pixel * (unsigned long long) P.sppTotal + L.sample, 0u);
        float3 o, d;
        float tBox;
        if (EXTRAS && P.lightMode) {
            /* ---- emitter-side walk: sample the emitter (its stream is keyed like a camera sample's) */
            if (P.emitterType == MER_EMITTER_COLLIMATED) { /* collimated.cpp:59-110: delta position and direction, weight = power */
                o = f3(P.beamO[0], P.beamO[1], P.beamO[2]);
                d = f3(P.beamD[0], P.beamD[1], P.beamD[2]);
                L.thr[0] = P.beamPower[0]; L.thr[1] = P.beamPower[1]; L.thr[2] = P.beamPower[2];
            } else { /* two-sided diffuse quad: uniform position, cosine-weighted direction => weight Le * pi * Area * 2 */
                const float u1 = L.rng.next(), u2 = L.rng.next(), u3 = L.rng.next(), u4 = L.rng.next(), u5 = L.rng.next();
                const float3 qu = f3(P.quadU[0], P.quadU[1], P.quadU[2]), qv = f3(P.quadV[0], P.quadV[1], P.quadV[2]);
                float3 Nq = f3(qu.y * qv.z - qu.z * qv.y, qu.z * qv.x - qu.x * qv.z, qu.x * qv.y - qu.y * qv.x);
                const float area = sqrtf(dot3(Nq, Nq)), side = u3 < 0.5f ? 1.0f : -1.0f;
                Nq = f3(Nq.x / area * side, Nq.y / area * side, Nq.z / area * side);
                float3 sa, ta;
                coordinate_system(Nq, sa, ta);
                const float rr = sqrtf(u4), lz = sqrtf(fmaxf(0.0f, 1.0f - u4));
                float sp, cp;
                sincosf(6.283185307179586f * u5, &sp, &cp);
                const float lx = rr * cp, ly = rr * sp;
                o = f3(P.quadO[0] + u1 * qu.x + u2 * qv.x, P.quadO[1] + u1 * qu.y + u2 * qv.y, P.quadO[2] + u1 * qu.z + u2 * qv.z);
                d = f3(sa.x * lx + ta.x * ly + Nq.x * lz, sa.y * lx + ta.y * ly + Nq.y * lz, sa.z * lx + ta.z * ly + Nq.z * lz);
                const float wgt = 6.283185307179586f * area;
                L.thr[0] = P.quadLe[0] * wgt; L.thr[1] = P.quadLe[1] * wgt; L.thr[2] = P.quadLe[2] * wgt;
            }
            if (!intersect_shape<SDFSHAPE>(M, o, d, tBox)) continue; /* misses the medium: next path */
            L.depth = 1;
            if (P.maxDepth != -1 && L.depth >= P.maxDepth) continue;
            L.opl = tBox; /* emitterPathlength counts every edge from the emitter, bdpt_proc.cpp:160-165 */
        } else {
            const unsigned py = pixel / (unsigned) P.W;
            const int x = (int) (pixel - py * (unsigned) P.W), y = (int) py;
            const float sx = (float) x + L.rng.next(), sy = (float) y + L.rng.next();
            /* PerspectiveCamera::sampleRay, closed form of m_sampleToCamera */
            float cx = (1.0f - 2.0f * (sx * P.invW)) * P.tanHalf, cy = (1.0f - 2.0f * (sy * P.invH)) * P.tanHalf / P.aspect;
            const float inv = 1.0f / sqrtf(cx * cx + cy * cy + 1.0f);
            cx *= inv; cy *= inv;
            const float cz = inv;
            d = f3(P.camLeft[0] * cx + P.camUp[0] * cy + P.camDir[0] * cz, P.camLeft[1] * cx + P.camUp[1] * cy + P.camDir[1] * cz,
                   P.camLeft[2] * cx + P.camUp[2] * cy + P.camDir[2] * cz);
            o = f3(P.camO[0], P.camO[1], P.camO[2]);
            float tQuad;
            const bool hitBox = intersect_shape<SDFSHAPE>(M, o, d, tBox), hitQuad = intersect_quad(P, o, d, tQuad);
            const float zero[3] = {0.f, 0.f, 0.f};
            if (hitQuad && (!hitBox || tQuad < tBox)) { finish_sample<EXTRAS>(P, L, P.quadLe, 1.0f, st, (EXTRAS && P.calibrated) ? 0.0f : tQuad); continue; }
            if (!hitBox) { finish_sample<EXTRAS>(P, L, P.env, 0.0f, st); continue; }
            L.depth = 1;
            if (P.maxDepth != -1 && L.depth >= P.maxDepth) { finish_sample<EXTRAS>(P, L, zero, 1.0f, st); continue; }
            L.opl = (EXTRAS && P.calibrated) ? 0.0f : tBox; /* bdpt_proc.cpp:163-171 */
            L.thr[0] = L.thr[1] = L.thr[2] = 1.0f;
        }
        if (!DIELECTRIC) L.depth = 2; /* index-matched container surface, volpath.cpp:287-296 */
        L.etaPath = 1.0f;
        L.p = f3(o.x + tBox * d.x, o.y + tBox * d.y, o.z + tBox * d.z);
        L.v = d;
        L.flags = 0;
        L.safe = 0.0f;
        L.n = 1.0f;
        L.G = f3(0.f, 0.f, 0.f);
        L.kind = K_ENTRY;
    }
}

/* ------------------------------------------------------------------ the two kernels of a round
 * The path pool is a set of FIXED slots in global memory (128 bytes per path, structure of arrays of 16-byte quads);
 * a slot renders one camera sample after the other until the frame runs out of samples.  A round is
 *
 *   k_event   one thread per slot whose path is at an EVENT: the remainder step and the step back of trace() (:674-686),
 *             the zero-length entry step, steps whose stencil touches the edge of the grid, free-flight sampling, Woodcock
 *             test, scattering, exits, film splat, next camera sample.  Runs until the path can take plain full steps.
 *   k_step    plain full steps only (kind K_FULL: +h, inside the container, interior stencil).  A persistent kernel:
 *             a lane POPS the next steppable slot from a global counter (one atomic per warp and refill), keeps the
 *             hot state (14 words) and the 4x4x4 coefficient block (64 words) in registers, steps until the path
 *             reaches an event or the visit's step budget, writes the hot state back and pops the next slot.
 *
 * Why two kernels (ncu r02s/r02u/r02v, profiles/README.md): with the event code inside the step kernel — inlined, or out
 * of line and lane-dense — its 54 KB of instructions were cycled through the 32 KB instruction cache between every few
 * turns of a 10 KB step loop (`no_instructions` the top stall reason, 34 %), it ran at 8 of 32 lanes because the kinds of
 * event diverge, and the step loop had to keep two paths per lane in shared memory to stay busy.  Here the step kernel's
 * whole code is the loop, lanes are refilled from the pool (no second path, no shared memory), and the event kernel
 * costs what it costs once per ~60 steps.
 *
 * A turn of the step loop: first kick + drift of step k (er_step :655-657) -> the cell of the drifted point is the EXACT
 * cell of the lookup: if it differs from the cached block the new block is requested (the one fetch site; the lane does
 * not sit the turn out, the SM's other warps cover the latency) -> contraction -> second kick -> containment test. */
__device__ __forceinline__ bool is_waiting(unsigned kf) {
    const int kind = (int) (kf & 0xffu), flags = (int) (kf >> 8);
    return (kind != K_FULL || (flags & FLAG_SLOW)) && kind != K_DEAD && !(flags & FLAG_PARKED);
}
__device__ __forceinline__ bool is_steppable(unsigned kf) { return (kf & 0xffu) == K_FULL && !((kf >> 8) & (FLAG_SLOW | FLAG_PARKED)); }

struct Hot { /* what a full step reads and writes */
    float3 p, v;
    float n;
    float3 G;
    float distSurf, opl, safe;
    int stepsLeft, flags;
};

__device__ __forceinline__ void hot_load(const PathPool &Q, unsigned slot, unsigned kf, int stepsLeft, Hot &H) {
    const float4 a = Q.h0[slot], b = Q.h1[slot], c = Q.h2[slot];
    H.p = f3(a.x, a.y, a.z);
    H.v = f3(a.w, b.x, b.y);
    H.n = b.z;
    H.G = f3(b.w, c.x, c.y);
    H.distSurf = c.z;
    H.opl = c.w;
    H.stepsLeft = stepsLeft;
    H.flags = (int) (kf >> 8);
    H.safe = 0.0f;
}
__device__ __forceinline__ void hot_store(const PathPool &Q, unsigned slot, const Hot &H, int kind) {
    Q.h0[slot] = make_float4(H.p.x, H.p.y, H.p.z, H.v.x);
    Q.h1[slot] = make_float4(H.v.y, H.v.z, H.n, H.G.x);
    Q.h2[slot] = make_float4(H.G.y, H.G.z, H.distSurf, H.opl);
    Q.h3[slot] = make_uint2((unsigned) H.stepsLeft, (unsigned) kind | ((unsigned) H.flags << 8));
}

/* the complete path <-> pool (event kernel) */
__device__ __forceinline__ void full_load(const RenderParams &P, unsigned slot, unsigned kf, int stepsLeft, Lane &L) {
    const PathPool &Q = P.pool;
    const float4 a = Q.h0[slot], b = Q.h1[slot], c = Q.h2[slot], d = Q.c0[slot], e = Q.c1[slot], f = Q.c2[slot];
    const uint4 g = Q.c3[slot];
    L.p = f3(a.x, a.y, a.z);
    L.v = f3(a.w, b.x, b.y);
    L.n = b.z;
    L.G = f3(b.w, c.x, c.y);
    L.distSurf = c.z;
    L.opl = c.w;
    L.stepsLeft = stepsLeft;
    L.kind = (int) (kf & 0xffu);
    L.flags = (int) (kf >> 8) & ~FLAG_PARKED;
    L.thr[0] = d.x; L.thr[1] = d.y; L.thr[2] = d.z; L.refStart = d.w;
    L.segDist = e.x; L.sd = e.y; L.etaPath = e.z; L.rem = e.w;
    L.o = f3(f.x, f.y, f.z);
    L.depth = __float_as_int(f.w);
    L.pixel = g.y; L.sample = g.z;
    L.phasePdf = __uint_as_float(g.w);
    L.rng.init(P.seed, (unsigned long long) g.y * (unsigned long long) P.sppTotal + g.z, g.x);
    L.safe = 0.0f;
}
__device__ __forceinline__ void full_store(const RenderParams &P, unsigned slot, const Lane &L) {
    const PathPool &Q = P.pool;
    Q.h0[slot] = make_float4(L.p.x, L.p.y, L.p.z, L.v.x);
    Q.h1[slot] = make_float4(L.v.y, L.v.z, L.n, L.G.x);
    Q.h2[slot] = make_float4(L.G.y, L.G.z, L.distSurf, L.opl);
    Q.h3[slot] = make_uint2((unsigned) L.stepsLeft, (unsigned) L.kind | ((unsigned) L.flags << 8));
    Q.c0[slot] = make_float4(L.thr[0], L.thr[1], L.thr[2], L.refStart);
    Q.c1[slot] = make_float4(L.segDist, L.sd, L.etaPath, L.rem);
    Q.c2[slot] = make_float4(L.o.x, L.o.y, L.o.z, __int_as_float(L.depth));
    Q.c3[slot] = make_uint4(L.rng.k, L.pixel, L.sample, __float_as_uint(L.phasePdf));
}

/* every slot starts by asking for a camera sample */
__global__ void k_pool_init(PathPool Q, unsigned nSlots) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nSlots; i += gridDim.x * blockDim.x)
        Q.h3[i] = make_uint2(0u, (unsigned) E_NEW);
}

template <int MODE, bool DIELECTRIC, bool EXTRAS, bool SDFSHAPE, bool XFORM>
__global__ void __launch_bounds__(128, MER_EVENT_MIN_BLOCKS)
k_event(const __grid_constant__ RenderParams P) {
    EvStats st;
#pragma unroll
    for (int i = 0; i < ST_COUNT; i++) st.v[i] = 0u;
    unsigned live = 0;
    for (unsigned slot = blockIdx.x * blockDim.x + threadIdx.x; slot < P.nSlots; slot += gridDim.x * blockDim.x) {
        const uint2 hk = P.pool.h3[slot];
        unsigned kf = hk.y & ~((unsigned) FLAG_PARKED << 8); /* a parked vertex tries the request queue again */
        if (is_waiting(kf)) {
            Lane L;
            full_load(P, slot, kf, (int) hk.x, L);
            handle_events<MODE, DIELECTRIC, EXTRAS, SDFSHAPE, XFORM>(P, L, st);
            full_store(P, slot, L);
            kf = (unsigned) L.kind;
        }
        if ((kf & 0xffu) != K_DEAD) live++;
    }
    /* one atomic per warp and counter that moved */
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int i = 0; i < ST_COUNT; i++) {
        const unsigned v = __reduce_add_sync(0xffffffffu, st.v[i]);
        if (lane == 0 && v) atomicAdd(P.stats + i, (unsigned long long) v);
    }
    live = __reduce_add_sync(0xffffffffu, live);
    if (lane == 0 && live) atomicAdd(P.live, live);
}

template <int MODE, int LAYOUT, bool EXTRAS, bool SDFSHAPE, bool XFORM>
__global__ void __launch_bounds__(MER_STENCIL_BLOCK, MER_RENDER_MIN_BLOCKS)
k_step(const __grid_constant__ RenderParams P) {
    const unsigned lane = threadIdx.x & 31u, ltMask = (1u << lane) - 1u;
    const MediumDev &M = P.M;
    const PathPool &Q = P.pool;
    const unsigned NONE = 0xffffffffu;
    Hot L;
    L.p = L.v = L.G = f3(0.f, 0.f, 0.f);
    L.n = 1.0f; L.distSurf = L.opl = L.safe = 0.0f; L.stepsLeft = L.flags = 0;
    unsigned cur = NONE; /* the slot whose hot state is in registers */
    int visit = 0;       /* steps left in this visit */
    StencilCache<MODE> S; /* registers only */
    S.invalidate();
    const float h = M.h;
    bool exhausted = false; /* warp-uniform: the round's slots are all handed out */
    unsigned nSteps = 0, nFetch = 0;
    while (true) {
        /* Explicit reconvergence (NVVM folds __syncwarp() into the following vote, so the barrier is spelled in PTX; and
         * no other warp barrier may appear in this kernel, or ptxas drops the converged-warp assumption and every vote
         * becomes BRA.DIV + WARPSYNC.COLLECTIVE, which synchronises without merging the groups: ncu r02u, 17 of 32 lanes) */
        asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
        unsigned need = __ballot_sync(0xffffffffu, cur == NONE);
        if (need == 0xffffffffu && exhausted) break;
        /* ---- refill: lanes without a path pop the next slots of the round, one atomic per warp and attempt */
        if (!exhausted && (__popc(need) >= P.refillGate || need == 0xffffffffu)) {
            while (need != 0u && !exhausted) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(P.stepHead, (unsigned) __popc(need));
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned idx = base + (unsigned) __popc(need & ltMask);
                bool got = false;
                if (((need >> lane) & 1u) && idx < P.nSlots) {
                    const uint2 hk = Q.h3[idx];
                    if (is_steppable(hk.y)) {
                        hot_load(Q, idx, hk.y, (int) hk.x, L);
                        cur = idx;
                        visit = P.stepsPerPass;
                        got = true;
                    }
                }
                exhausted = base + (unsigned) __popc(need) >= P.nSlots;
                need &= ~__ballot_sync(0xffffffffu, got);
            }
        }
        /* ---- first half of the step, then the one place where blocks are requested */
        CellPos c;
        bool step = false;
        if (cur != NONE) {
            if (!(L.flags & FLAG_DRIFTED)) lane_drift<EXTRAS>(L, h);
            c = rif_cell<MODE, XFORM>(M.rif, L.p);
            if (rif_cell_fast<MODE>(M.rif, c)) {
                if (!stencil_has(S, c)) {
                    rif_fetch_interior<LAYOUT>(M.rif, S, c.i, c.j, c.k);
                    nFetch++;
                }
                step = true;
            } else { /* the stencil touches the edge of the grid: a step for the event kernel */
                L.flags |= FLAG_SLOW;
                hot_store(Q, cur, L, K_FULL);
                cur = NONE;
            }
        }
        if (step) {
            /* ---------------- lookup and second half of er_step */
            rif_contract<XFORM>(M.rif, S, c, L.n, L.G, [](StencilCache<MODE> &) {});
            lane_kick(L, h);
            const bool inside = SDFSHAPE ? inside_shape_lazy(M, L.p, h, L.safe) : inside_shape(M, L.p);
            nSteps++;
            L.flags &= ~FLAG_DRIFTED;
            int next = K_BACKF;
            if (inside) {
                L.distSurf += h;
                L.stepsLeft--;
                next = L.stepsLeft > 0 ? K_FULL : ((L.flags & FLAG_TB) ? E_EXIT : K_REM);
            }
            if (next != K_FULL || --visit <= 0) { /* an event, or the end of the visit: back to the pool */
                hot_store(Q, cur, L, next);
                cur = NONE;
            }
        }
    }
    nSteps = __reduce_add_sync(0xffffffffu, nSteps);
    nFetch = __reduce_add_sync(0xffffffffu, nFetch);
    if (lane == 0 && nSteps) atomicAdd(P.stats + ST_STEPS, (unsigned long long) nSteps);
    if (lane == 0 && nFetch) atomicAdd(P.stats + ST_FETCH, (unsigned long long) nFetch);
}

/* A warp of k_nee costs what its longest connection costs, and the cost is roughly the number of steps from the vertex to
 * the container surface.  So the requests are counting-sorted by that distance (towards the centre of the quad, 64 bins,
 * or the pinhole; longest first) before the solver runs: three trivial kernels in front of one that costs 1e3-1e5 steps per thread. */
__global__ void k_nee_keys(const __grid_constant__ RenderParams P, unsigned nReq) {
    __shared__ unsigned hist[NEE_BINS];
    if (threadIdx.x < NEE_BINS) hist[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nReq) {
        const float4 a = P.neeQ0[i];
        const float3 p1 = f3(a.x, a.y, a.z);
        float3 d = P.lightMode ? f3(P.camO[0] - p1.x, P.camO[1] - p1.y, P.camO[2] - p1.z)
                               : f3(P.quadO[0] + 0.5f * (P.quadU[0] + P.quadV[0]) - p1.x, P.quadO[1] + 0.5f * (P.quadU[1] + P.quadV[1]) - p1.y,
                                    P.quadO[2] + 0.5f * (P.quadU[2] + P.quadV[2]) - p1.z);
        const float dl = 1.0f / sqrtf(dot3(d, d));
        d = f3(d.x * dl, d.y * dl, d.z * dl);
        const float te = exit_distance(P.M, p1, d);
        float extent;
        if (P.M.shapeType == MER_SHAPE_SPHERE) extent = 2.0f * P.M.shape[3];
        else extent = sqrtf((P.M.shape[3] - P.M.shape[0]) * (P.M.shape[3] - P.M.shape[0]) + (P.M.shape[4] - P.M.shape[1]) * (P.M.shape[4] - P.M.shape[1]) +
                            (P.M.shape[5] - P.M.shape[2]) * (P.M.shape[5] - P.M.shape[2]));
        const int bin = min(max((int) (te / extent * (float) NEE_BINS), 0), NEE_BINS - 1);
        const unsigned key = (unsigned) (NEE_BINS - 1 - bin); /* longest first */
        P.neeKey[i] = (unsigned char) key;
        atomicAdd(&hist[key], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NEE_BINS && hist[threadIdx.x]) atomicAdd(P.neeHist + threadIdx.x, hist[threadIdx.x]);
}

__global__ void k_nee_offsets(unsigned *hist) { /* exclusive prefix sum of NEE_BINS counters, one thread: 64 adds */
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned run = 0;
        for (int b = 0; b < NEE_BINS; b++) { const unsigned c = hist[b]; hist[b] = run; run += c; }
    }
}

__global__ void k_nee_scatter(const __grid_constant__ RenderParams P, unsigned nReq) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nReq) P.neePerm[atomicAdd(P.neeHist + P.neeKey[i], 1u)] = i;
}

/* Next-event estimation of the quad emitter from the queued scattering vertices, one thread per vertex.
 * Around the reference's shooting problem (makeDirectConnections, mer_connect.cuh) the estimator is
 *     thr * phase(wi, w) * exp(-sigma_t dist) * (n_b / n_1)^2 [* (1 - Fresnel) n_b^2 for hdielectric]
 *         * Le * |cos theta_y| * Area / |d r_perp / d omega|
 * i.e. the random walk's own exit-edge weights with the change of variables launch direction -> sampled point written
 * with the solver's Jacobian in place of 1 / distance^2.  The splat adds radiance only (filter weight 0): the sample's
 * weight is added once, by the walk. */
template <bool WANT_OPL, bool SDFSHAPE>
__global__ void __launch_bounds__(128, 2)
k_nee(const __grid_constant__ RenderParams P, unsigned nReq) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const MediumDev &M = P.M;
    unsigned st[3] = {0u, 0u, 0u}; /* connections, failed, steps */
    __shared__ unsigned nonfinite;
    if (threadIdx.x == 0) nonfinite = 0u;
    __syncthreads();
    if (tid < nReq) {
        const unsigned i = P.neePerm ? P.neePerm[tid] : tid;
        const float4 a = P.neeQ0[i], b = P.neeQ1[i];
        const uint4 c = P.neeQ2[i];
        const float3 p1 = f3(a.x, a.y, a.z), wi = f3(a.w, b.x, b.y);
        const float thr[3] = {b.z, b.w, __uint_as_float(c.x)};
        const bool isHit = (c.y & NEE_HIT) != 0u; /* MIS weight of a phase-sampled path that hit the quad */
        const unsigned depth = c.y & ~NEE_HIT, pixel = c.z, sample = c.w;
        const float oplVertex = (WANT_OPL || isHit) ? P.neeQ3[i] : 0.0f;
        PathRng nrng;
        nrng.init(P.seed ^ MER_NEE_SALT, (unsigned long long) pixel * (unsigned long long) P.sppTotal + sample, depth * 256u + (isHit ? 128u : 0u));
        float3 y, Nq = f3(0.f, 0.f, 1.f);
        float area = 1.0f;
        if (P.lightMode) { /* t = 1: the target is the pinhole */
            y = f3(P.camO[0], P.camO[1], P.camO[2]);
        } else if (isHit) { /* the target is the point the walk hit (stored in the wi slots) */
            const float3 qu = f3(P.quadU[0], P.quadU[1], P.quadU[2]), qv = f3(P.quadV[0], P.quadV[1], P.quadV[2]);
            Nq = f3(qu.y * qv.z - qu.z * qv.y, qu.z * qv.x - qu.x * qv.z, qu.x * qv.y - qu.y * qv.x);
            area = sqrtf(dot3(Nq, Nq));
            y = wi;
        } else {
            const float u = nrng.next(), w = nrng.next();
            const float3 qu = f3(P.quadU[0], P.quadU[1], P.quadU[2]), qv = f3(P.quadV[0], P.quadV[1], P.quadV[2]);
            Nq = f3(qu.y * qv.z - qu.z * qv.y, qu.z * qv.x - qu.x * qv.z, qu.x * qv.y - qu.y * qv.x);
            area = sqrtf(dot3(Nq, Nq));
            y = f3(P.quadO[0] + u * qu.x + w * qv.x, P.quadO[1] + u * qu.y + w * qv.y, P.quadO[2] + u * qu.z + w * qv.z);
        }
        float3 ds = f3(y.x - p1.x, y.y - p1.y, y.z - p1.z);
        const float dl = 1.0f / sqrtf(dot3(ds, ds));
        ds = f3(ds.x * dl, ds.y * dl, ds.z * dl);
        const bool refract = M.boundary == MER_BOUNDARY_HDIELECTRIC;
        merc::ConnectResult C;
        merc::connect_solve<SDFSHAPE, WANT_OPL>(M, P.neePrecision, P.neeTol2, P.neeRRWeight, P.neeMaxIterations, p1, y, ds, true, refract, P.neeStraightFirst != 0,
                            nrng, C);
        st[0] = 1u;
        const int steps = C.steps;
        bool ok = C.success && C.exit.exited && !C.exit.tir;
        if (ok || isHit) {
            const float spread = ok ? merc::connection_spread(C.J, C.xnorm) : 0.0f; /* the Jacobian of the solver's last accepted evaluation */
            ok = ok && spread > 0.0f;
            if (ok && P.lightMode) {
                /* C.rev is the direction in which the camera sees the vertex: pixel by the inverse of sampleRay's mapping,
                 * importance of the perspective sensor 1 / (A cos^3 theta) (src/sensors/perspective.cpp importance()),
                 * irradiance on the plane perpendicular to the arriving ray = intensity / spread (flux is conserved: no n^2) */
                const float3 cl = f3(P.camLeft[0], P.camLeft[1], P.camLeft[2]), cu = f3(P.camUp[0], P.camUp[1], P.camUp[2]),
                             cd = f3(P.camDir[0], P.camDir[1], P.camDir[2]);
                const float zc = dot3(C.rev, cd);
                ok = zc > 0.0f;
                if (ok) {
                    const float cx = dot3(C.rev, cl) / zc, cy = dot3(C.rev, cu) / zc;
                    const float sx = 0.5f * (float) P.W * (1.0f - cx / P.tanHalf), sy = 0.5f * (float) P.H * (1.0f - cy * P.aspect / P.tanHalf);
                    ok = sx >= 0.0f && sx < (float) P.W && sy >= 0.0f && sy < (float) P.H;
                    if (ok) {
                        const float inv1 = 1.0f / C.n1;
                        const float phase = hg_eval_dev(M.g, wi, f3(C.dir.x * inv1, C.dir.y * inv1, C.dir.z * inv1));
                        float bf = 1.0f;
                        if (refract) {
                            float cosT;
                            bf = 1.0f - fresnel_dielectric_ext(-C.exit.cosI, cosT, C.exit.nb);
                        }
                        const float g = P.lightScale / (spread * zc * zc * zc);
                        float rad[3];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const float T = fastexp_dev(M.hasGrid ? -C.exit.tau : M.sigmaT[k] * -C.dist);
                            rad[k] = thr[k] * phase * T * C.weight * bf * g;
                        }
                        /* bdpt_proc.cpp:352-357: the sensor connection's length is left out of a calibrated transient */
                        const float len = oplVertex + (P.calibrated ? 0.0f : C.opl);
                        if (WANT_OPL && P.modulation) {
                            const float w = path_weight(P, len);
#pragma unroll
                            for (int k = 0; k < 3; k++) rad[k] *= w;
                        }
                        const int frame = (WANT_OPL && !P.modulation) ? path_frame(P, len) : 0;
                        if (frame >= 0 && film_put(P, sx, sy, rad, 0.0f, 0.0f, frame)) atomicAdd(&nonfinite, 1u);
                    }
                }
            } else if (isHit) {
                /* the hit is already weighted by the walk; what is missing is the density p_nee (solid angle at the vertex)
                 * with which the next-event estimator samples this very point: spread / (Area cos theta_y) */
                float w = 1.0f;
                if (ok) {
                    const float cosY = fabsf(dot3(C.rev, Nq)) / area;
                    const float pNee = spread / fmaxf(cosY * area, 1e-20f), pPhase = P.neeQ4[i];
                    w = (pPhase * pPhase) / fmaxf(pPhase * pPhase + pNee * pNee, 1e-30f);
                }
                ok = true; /* an unsolvable connection means next-event estimation could not have found the point: weight 1 */
                float rad[3] = {thr[0] * w, thr[1] * w, thr[2] * w};
                float sx, sy;
                sample_position(P, pixel, sample, sx, sy);
                if (WANT_OPL && P.modulation) {
                    const float wm = path_weight(P, oplVertex);
#pragma unroll
                    for (int k = 0; k < 3; k++) rad[k] *= wm;
                }
                const int frame = (WANT_OPL && !P.modulation) ? path_frame(P, oplVertex) : 0;
                if (frame >= 0 && film_put(P, sx, sy, rad, 0.0f, 0.0f, frame)) atomicAdd(&nonfinite, 1u);
            } else if (ok) {
                const float cosY = fabsf(dot3(C.rev, Nq)) / area;
                const float inv1 = 1.0f / C.n1;
                const float phase = hg_eval_dev(M.g, wi, f3(C.dir.x * inv1, C.dir.y * inv1, C.dir.z * inv1));
                float scale = (float) (1.0 / (double) (C.n1 * C.n1)) * C.exit.nb * C.exit.nb;
                if (M.physicalScaling) scale = 1.0f / scale;
                if (refract) {
                    float cosT;
                    const float Fr = fresnel_dielectric_ext(-C.exit.cosI, cosT, C.exit.nb);
                    scale *= (1.0f - Fr) * (C.exit.nb * C.exit.nb);
                }
                float geom = cosY * area / spread;
                if (P.neeMis) { /* miWeight(pdf_emitter, pdf_phase), volpath.cpp:137-141: both densities per solid angle at the vertex */
                    const float pNee = 1.0f / fmaxf(geom, 1e-30f);
                    geom *= (pNee * pNee) / fmaxf(pNee * pNee + phase * phase, 1e-30f);
                }
                float rad[3];
#pragma unroll
                for (int k = 0; k < 3; k++) { /* homogeneous: exp(-sigma_t dist); density grid: exp(-optical depth along the curve) */
                    const float T = fastexp_dev(M.hasGrid ? -C.exit.tau : M.sigmaT[k] * -C.dist);
                    rad[k] = thr[k] * phase * T * C.weight * scale * P.quadLe[k] * geom;
                }
                float sx, sy;
                sample_position(P, pixel, sample, sx, sy);
                /* the connection's optical length: curved part (midpoint rule, :941-1030) + exterior segment */
                if (WANT_OPL && P.modulation) {
                    const float w = path_weight(P, oplVertex + C.opl);
#pragma unroll
                    for (int k = 0; k < 3; k++) rad[k] *= w;
                }
                const int frame = (WANT_OPL && !P.modulation) ? path_frame(P, oplVertex + C.opl) : 0;
                if (frame >= 0 && film_put(P, sx, sy, rad, 0.0f, 0.0f, frame)) atomicAdd(&nonfinite, 1u);
            }
        }
        if (!ok) st[1] = 1u;
        st[2] = (unsigned) steps;
    }
    const int slots[3] = {ST_CONN, ST_CONNFAIL, ST_CONNSTEPS};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        unsigned long long v = st[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(P.stats + slots[k], v);
    }
    __syncthreads();
    if (threadIdx.x == 0 && nonfinite) atomicAdd(P.stats + ST_NONFINITE, (unsigned long long) nonfinite);
}

/* a light-traced film has no camera samples: every pixel gets this GPU's share of a unit weight (and alpha), so that
 * develop() returns the sum of the splats, which are already normalised by the number of light paths */
__global__ void k_film_unit_weight(size_t nPixels, int channels, float share, float *__restrict__ film) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nPixels; i += (size_t) gridDim.x * blockDim.x) {
        film[i * channels + channels - 2] += share;
        film[i * channels + channels - 1] += share;
    }
}

/* HDRFilm::develop: ESpectrumAlphaWeight -> RGB */
__global__ void k_develop(size_t nPixels, int frames, const float *__restrict__ film, float *__restrict__ rgb) {
    const size_t C = 3 * (size_t) frames + 2, V = 3 * (size_t) frames;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < nPixels * V; i += (size_t) gridDim.x * blockDim.x) {
        const size_t px = i / V, k = i - px * V;
        const float w = film[C * px + C - 1], inv = w != 0.0f ? 1.0f / w : 0.0f;
        rgb[i] = film[C * px + k] * inv;
    }
}

/* ReconstructionFilter::configure (rfilter.cpp:37-55) for gaussian (stddev .5) and box (radius .5+1e-5) */
void configure_filter(int type, RenderParams &P) {
    const int RES = 31;
    const float stddev = 0.5f;
    const float radius = type == MER_FILTER_GAUSSIAN ? 4 * stddev : 0.5f + 1e-5f;
    float sum = 0;
    for (int i = 0; i < RES; i++) {
        float x = (radius * i) / RES, value;
        if (type == MER_FILTER_GAUSSIAN) {
            float alpha = -1.0f / (2.0f * stddev * stddev);
            value = std::max(0.0f, (float) std::exp((double) (alpha * x * x)) -
                                       (float) std::exp((double) (alpha * radius * radius)));
        } else {
            value = std::abs(x) <= radius ? 1.0f : 0.0f;
        }
        P.filterValues[i] = value;
        sum += value;
    }
    P.filterValues[RES] = 0.0f;
    P.filterScale = RES / radius;
    P.filterRadius = radius;
    sum *= 2 * radius / RES;
    float normalization = 1.0f / sum;
    for (int i = 0; i < RES; i++) P.filterValues[i] *= normalization;
}

void configure_camera(const mer_render_desc *r, RenderParams &P) {
    /* Transform::lookAt, src/libcore/transform.cpp:191-214 */
    float d[3], len = 0;
    for (int i = 0; i < 3; i++) { P.camO[i] = r->cam_origin[i]; d[i] = r->cam_target[i] - r->cam_origin[i]; len += d[i] * d[i]; }
    len = std::sqrt(len);
    for (int i = 0; i < 3; i++) P.camDir[i] = d[i] / len;
    const float *u = r->cam_up, *dir = P.camDir;
    float l[3] = {u[1] * dir[2] - u[2] * dir[1], u[2] * dir[0] - u[0] * dir[2], u[0] * dir[1] - u[1] * dir[0]};
    len = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int i = 0; i < 3; i++) P.camLeft[i] = l[i] / len;
    P.camUp[0] = dir[1] * P.camLeft[2] - dir[2] * P.camLeft[1];
    P.camUp[1] = dir[2] * P.camLeft[0] - dir[0] * P.camLeft[2];
    P.camUp[2] = dir[0] * P.camLeft[1] - dir[1] * P.camLeft[0];
    P.tanHalf = std::tan(0.5f * r->fov_deg * (float) (M_PI / 180.0));
    P.aspect = (float) r->width / (float) r->height;
    P.invW = 1.0f / r->width;
    P.invH = 1.0f / r->height;
}

} /* namespace */

extern "C" {

int mer_render_device(const mer_medium *m, const mer_render_desc *r, float *film_dev, mer_render_stats *stats_out,
                      void *stream_) {
    MER_REQUIRE(m && r && film_dev, "null argument");
    MER_REQUIRE(r->width > 0 && r->height > 0 && r->spp_total > 0, "film size and sample count must be positive");
    MER_REQUIRE(r->sample_stride >= 1 && r->sample_begin >= 0, "bad sample sharding");
    MER_REQUIRE(r->filter == MER_FILTER_BOX || r->filter == MER_FILTER_GAUSSIAN, "unknown reconstruction filter");
    if (m->dev.aggressive)
        return mer::fail(MER_ERR_UNSUPPORTED, "aggressivetracing is only available in mer_medium_sample_distance_batch");
    if (m->dev.shapeType == MER_SHAPE_SDF) {
        MER_REQUIRE(m->dev.hasSdf, "shape type SDF needs the sdf volume (mer_medium_set_sdf)");
        if (m->rif->mode != MER_RIF_TRICUBIC) return mer::fail(MER_ERR_UNSUPPORTED, "shape type SDF is built for the tricubic RIF mode");
    }
    if (r->light_tracing) {
        MER_REQUIRE(r->emitter_type == MER_EMITTER_QUAD || r->emitter_type == MER_EMITTER_COLLIMATED, "unknown emitter type");
        MER_REQUIRE(r->emitter_type != MER_EMITTER_QUAD || r->has_quad, "light tracing from the quad emitter needs the quad");
        if (m->rif->mode != MER_RIF_TRICUBIC)
            return mer::fail(MER_ERR_UNSUPPORTED, "light tracing needs the tricubic RIF mode (the sensor connection differentiates the spline twice)");
    }
    if (r->direct_connections && !r->light_tracing) {
        MER_REQUIRE(r->has_quad, "direct_connections needs the quad emitter");
        if (m->rif->mode != MER_RIF_TRICUBIC)
            return mer::fail(MER_ERR_UNSUPPORTED, "direct_connections needs the tricubic RIF mode (the solver differentiates the spline twice)");
    }
    mer::DeviceGuard guard(m->device);
    cudaStream_t stream = (cudaStream_t) stream_;

    RenderParams P;
    memset(&P, 0, sizeof(P));
    P.M = m->dev;
    configure_camera(r, P);
    configure_filter(r->filter, P);
    P.W = r->width; P.H = r->height;
    P.sppTotal = r->spp_total; P.sampleBegin = r->sample_begin; P.sampleStride = r->sample_stride;
    P.sppLocal = r->sample_begin < r->spp_total ? (r->spp_total - r->sample_begin + r->sample_stride - 1) / r->sample_stride : 0;
    P.seed = r->seed;
    P.totalSamples = (unsigned long long) r->width * r->height * (unsigned long long) P.sppLocal;
    P.maxDepth = r->max_depth; P.rrDepth = r->rr_depth;
    for (int i = 0; i < 3; i++) {
        P.env[i] = r->env_radiance[i];
        P.quadO[i] = r->quad_origin[i]; P.quadU[i] = r->quad_u[i]; P.quadV[i] = r->quad_v[i]; P.quadLe[i] = r->quad_radiance[i];
    }
    P.hasQuad = r->has_quad;
    P.stepsPerPass = r->steps_per_pass > 0 ? r->steps_per_pass : 128;
    if (const char *e = getenv("MER_VISIT")) P.stepsPerPass = std::max(atoi(e), 1); /* tuning knobs */
    P.refillGate = 4;
    if (const char *e = getenv("MER_GATE")) P.refillGate = atoi(e);
    P.refillGate = std::min(std::max(P.refillGate, 1), 32);
    P.film = film_dev;
    MER_REQUIRE(r->modulation >= MER_MODULATION_NONE && r->modulation <= MER_MODULATION_HAMILTONIAN, "unknown modulation");
    P.modulation = r->modulation;
    P.lambda = r->lambda;
    P.phaseShift = (float) ((double) r->phase_deg * M_PI / 180.0) * r->lambda * (float) (0.5 / M_PI);
    if (P.modulation) MER_REQUIRE(r->lambda > 0.0f, "modulation: lambda must be positive");
    P.frames = (r->frames > 1 && !P.modulation) ? r->frames : 1; /* one frame under a modulation, film.cpp:76-78 */
    P.channels = 3 * P.frames + 2;
    P.minBound = r->min_bound;
    P.binWidth = r->bin_width;
    P.calibrated = r->calibrated_transient ? 1 : 0;
    if (P.frames > 1) {
        MER_REQUIRE(r->bin_width > 0.0f, "transient film: bin_width must be positive");
        MER_REQUIRE((double) r->width * r->height * (3.0 * P.frames + 2.0) * sizeof(float) <= 64.0 * (double) (1ull << 30),
                    "transient film: width * height * (3 * frames + 2) floats exceed 64 GiB");
    }
    P.nee = (r->direct_connections || r->light_tracing) ? 1 : 0;
    P.neeMis = (r->direct_connections == 2 && !r->light_tracing) ? 1 : 0;
    P.lightMode = r->light_tracing ? 1 : 0;
    P.emitterType = r->emitter_type;
    {
        float dl = 0.0f;
        for (int i = 0; i < 3; i++) { P.beamO[i] = r->beam_origin[i]; P.beamPower[i] = r->beam_power[i]; dl += r->beam_direction[i] * r->beam_direction[i]; }
        dl = dl > 0.0f ? 1.0f / std::sqrt(dl) : 0.0f;
        for (int i = 0; i < 3; i++) P.beamD[i] = r->beam_direction[i] * dl;
        if (P.lightMode && P.emitterType == MER_EMITTER_COLLIMATED) MER_REQUIRE(dl > 0.0f, "collimated emitter: zero direction");
        const double aImg = (2.0 * P.tanHalf) * (2.0 * P.tanHalf / P.aspect); /* image plane at unit distance */
        const double nPaths = (double) r->width * r->height * (double) r->spp_total;
        P.lightScale = (float) (1.0 / (nPaths * aImg / ((double) r->width * r->height)));
    }
    /* the reference drops connections that leave the shape within sqrt(Epsilon) = 0.01 of p1; for next-event estimation that
     * would discard the brightest vertices (those next to the surface facing the light), so only exact degeneracy is dropped */
    P.M.minExit2 = 1e-10f;
    P.neeTol2 = r->connection.tol2 > 0.0f ? r->connection.tol2 : 1e-6f;
    P.neeRRWeight = r->connection.rrweight > 0.0f ? r->connection.rrweight : 1e-2f;
    P.neePrecision = r->connection.boundary_precision > 0 ? r->connection.boundary_precision : 3;
    P.neeMaxIterations = r->connection.max_iterations > 0 ? r->connection.max_iterations : 20;
    P.neeStraightFirst = r->connection.start_mode != MER_START_RANDOM;

    const unsigned TPB = 128;
    unsigned pool = r->pool_paths > 0 ? (unsigned) r->pool_paths : 148u * 16384u; /* 2.4 M slots, 310 MB: a round's tail is amortised over more visits */
    int l2Bytes = 0;
    cudaDeviceGetAttribute(&l2Bytes, cudaDevAttrL2CacheSize, m->device);
    const size_t tableBytes = (size_t) m->dev.rif.N[0] * m->dev.rif.N[1] * m->dev.rif.N[2] * (m->rif->mode == MER_RIF_TRICUBIC ? (m->dev.rif.coeff8 ? 32u : 4u) : 16u);
    if (r->pool_paths <= 0 && tableBytes > (size_t) std::max(l2Bytes, 1)) {
        /* A table larger than the L2: the samples in flight are consecutive pixels (pixel-major sample ids), i.e. a slab of the
         * volume about pool / (samples of this call) of the table thick.  With few samples per pixel on this GPU (an N-GPU
         * frame gives a rank 1/N of them) the default pool spreads over N times more image rows and the slab outgrows the
         * L2, and the frame's ramp and drain rounds, whose number does not shrink with the work, weigh N times more.
         * Keep the slab at about a third of the L2.  Measured on one rank's share of the 64-spp C5 frame (tools/c5_rank_probe.py):
         * 16 spp 9.4 -> 10.5 G ray steps/s with 0.6 M slots, 8 spp 8.2 -> 10.0 G with 0.3 M (64 spp, 2.4 M slots: 10.9 G). */
        const double fit = (double) l2Bytes / 3.0 * (double) P.totalSamples / (double) tableBytes;
        pool = (unsigned) std::min((double) pool, std::max(fit, 148.0 * 2048.0));
    }
    if (const char *e = getenv("MER_POOL")) pool = (unsigned) atol(e); /* tuning knob */
    if ((unsigned long long) pool > P.totalSamples) pool = (unsigned) P.totalSamples;
    pool = std::max(((pool + TPB - 1) / TPB) * TPB, TPB);

    RenderScratch &S = mer::device_scratch(m->device);
    std::lock_guard<std::mutex> hold(S.lock);
    /* the scratch is sized for the default pool at least: frames of one scene choose pools of different sizes (above), and
     * re-allocating between two of them costs more than the smaller frame (cudaFree synchronises: 0.6-1.5 s measured) */
    const size_t qBytes = (size_t) std::max(pool, r->pool_paths > 0 ? 0u : 148u * 16384u) * 16;
    if (S.poolBytes < qBytes) {
        S.release();
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaMalloc(&S.pool[i], qBytes);
        if (e == cudaSuccess) e = cudaMalloc(&S.nOut, 4 * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMalloc(&S.counters, (1 + ST_COUNT) * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMallocHost(&S.hostPinned, (4 + RenderScratch::RING) * sizeof(unsigned long long)); /* + the step counter after every step launch */
        if (e == cudaSuccess) e = cudaEventCreate(&S.ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&S.ev1);
        if (e == cudaSuccess) e = cudaEventCreate(&S.evTail);
        for (int i = 0; i < 2 * RenderScratch::RING && e == cudaSuccess; i++) e = cudaEventCreate(&S.ring[i]);
        if (e != cudaSuccess) { /* all or nothing: a later call must not find half of the scratch */
            S.release();
            return mer::fail(e == cudaErrorMemoryAllocation ? MER_ERR_OOM : MER_ERR_CUDA, cudaGetErrorString(e));
        }
        S.poolBytes = qBytes;
    }
    if (P.nee) {
        const unsigned cap = std::min(pool * 2u, 1u << 21);
        if (S.neeCap < cap) {
            for (int i = 0; i < 3; i++) { cudaFree(S.neeQ[i]); S.neeQ[i] = nullptr; }
            S.neeCap = 0;
            for (int i = 0; i < 3; i++) MER_CUDA(cudaMalloc(&S.neeQ[i], (size_t) cap * 16));
            cudaFree(S.neePerm); cudaFree(S.neeKey);
            S.neePerm = nullptr; S.neeKey = nullptr;
            MER_CUDA(cudaMalloc(&S.neePerm, (size_t) cap * sizeof(unsigned)));
            MER_CUDA(cudaMalloc(&S.neeKey, (size_t) cap));
            if (!S.neeHist) MER_CUDA(cudaMalloc(&S.neeHist, NEE_BINS * sizeof(unsigned)));
            cudaFree(S.neeOpl);
            S.neeOpl = nullptr;
            MER_CUDA(cudaMalloc(&S.neeOpl, 2 * (size_t) cap * sizeof(float))); /* optical lengths, then the HIT requests' phase pdfs */
            if (!S.neeCount) MER_CUDA(cudaMalloc(&S.neeCount, sizeof(unsigned)));
            S.neeCap = cap;
        }
        P.neeCap = cap;
        P.neeCount = S.neeCount;
        P.neeQ0 = (float4 *) S.neeQ[0]; P.neeQ1 = (float4 *) S.neeQ[1]; P.neeQ2 = (uint4 *) S.neeQ[2];
        P.neeQ3 = S.neeOpl;
        P.neeQ4 = S.neeOpl + cap;
        P.neeKey = S.neeKey;
        P.neeHist = S.neeHist;
        P.neePerm = m->dev.hasSdf ? nullptr : S.neePerm; /* the length estimate needs an analytic container */
        if (const char *e = getenv("MER_NEE_SORT")) if (atoi(e) == 0) P.neePerm = nullptr; /* tuning knob */
        MER_CUDA(cudaMemsetAsync(S.neeCount, 0, sizeof(unsigned), stream));
    }
    P.pool = PathPool{(float4 *) S.pool[0], (float4 *) S.pool[1], (float4 *) S.pool[2], (uint2 *) S.pool[3],
                      (float4 *) S.pool[4], (float4 *) S.pool[5], (float4 *) S.pool[6], (uint4 *) S.pool[7]};
    P.nSlots = pool;
    P.stepHead = S.nOut;
    P.live = S.nOut + 1;
    MER_CUDA(cudaMemsetAsync(S.counters, 0, (1 + ST_COUNT) * sizeof(unsigned long long), stream));
    P.sampleCounter = S.counters;
    P.stats = S.counters + 1;

    /* the step kernel is persistent: as many CTAs as fit the GPU at once (fewer for a small pool) */
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    const unsigned eventBlocks = pool / TPB;
    /* Lanes per SM of the step kernel.  WIDE (4 x 128) is what an L2-resident table wants: the texture unit's return path is
     * the limit and every lane helps to cover the gather latency.  A table far larger than the L2 (C5: 4 GiB) behaves the
     * other way round once the paths of a warp have scattered apart: a path re-gathers its whole 4x4x4 block on every cell
     * change and three quarters of it are the previous block, which only the L1 can serve; a block occupies ~9.6 128-byte
     * lines (8 x 4 texels of one layer of the atlas each), so beyond ~200 lanes per SM the L1 thrashes, every gather goes to
     * L2 / DRAM and the texture unit's miss queue backs up (tex_throttle 74 % of the stall samples).  Measured on C5, lanes per
     * SM -> G ray steps/s: 32: 2.7, 64: 5.2, 96: 7.3, 128: 9.3, 160: 10.5-10.8, 192: 10.0, 224: 7.6, 256: 6.6, 384: 5.6,
     * 512: 5.4 (profiles/r02_c5_lanes_sweep.json).  NARROW = 5 x 32; at 10.8 G steps/s the gathers return 2.7 TB/s, 0.76 of
     * what the same texture path delivers from the L2-resident C2 table: the narrow configuration is texture-bound again.
     * Which of the two a scene wants is MEASURED, not guessed, and measured again as the frame goes on (the first rounds
     * of a frame are coherent camera rays, which like WIDE whatever the table): when the table does not fit the L2, four
     * probe rounds (wide, narrow, wide, narrow: the same mix of paths) are compared by ray steps per ms of the step kernel
     * and the faster one renders on; the gap to the next probe doubles (32 -> 512 rounds) while the comparison keeps
     * confirming the choice.  C3 (512^3, h = pitch / 2) stays WIDE: 61.3 vs 40.4 M samples/s. */
    struct StepCfg { unsigned ctas, tpb; };
    const StepCfg stepCfgs[2] = {{(unsigned) MER_RENDER_MIN_BLOCKS, TPB}, {5u, 32u}};
    StepCfg stepFixed = stepCfgs[0];
    bool stepEnv = false;
    if (const char *e = getenv("MER_STEP_CTAS")) { stepFixed.ctas = (unsigned) std::min(std::max(atoi(e), 1), 16); stepEnv = true; } /* tuning knobs */
    if (const char *e = getenv("MER_STEP_TPB")) { stepFixed.tpb = (unsigned) std::min(std::max(atoi(e) / 32 * 32, 32), (int) TPB); stepEnv = true; }
    if (stepFixed.ctas * stepFixed.tpb > (unsigned) MER_RENDER_MIN_BLOCKS * TPB) stepFixed.ctas = (unsigned) MER_RENDER_MIN_BLOCKS * TPB / stepFixed.tpb;
    bool tuning = !stepEnv && !P.nee && tableBytes > (size_t) std::max(l2Bytes, 1);
    if (const char *e = getenv("MER_STEP_TUNE")) tuning = atoi(e) != 0 && !P.nee && !stepEnv;
    /* probes: four rounds (wide, narrow, wide, narrow) starting at probeAt; decided at the next look at the pool; the gap to
     * the next probe doubles (32 -> 512 rounds) every time the comparison confirms the current choice */
    /* the first rounds of a frame are coherent camera rays, which like WIDE whatever the table (a comparison made there flips
     * the choice for the next 40 rounds: measured on a 16-spp C5 frame, 0.83 instead of 0.9+ of the 64-spp rate), so the first
     * probe waits for the mix of paths to settle - and longer when the previous frame of the same table left its choice */
    const bool hinted = tuning && S.stepChoiceTable == tableBytes;
    unsigned long long probeAt = hinted ? 49 : 17, probeGap = 32;
    auto probe_of = [&](unsigned long long round) -> int { /* -1: not a probe round, else the candidate it runs */
        if (!tuning || round < probeAt || round >= probeAt + 4) return -1;
        return (int) ((round - probeAt) & 1ull);
    };
    int stepChoice = (tuning && S.stepChoiceTable == tableBytes) ? S.stepChoice : 0; /* index into stepCfgs */
    double tuneSteps[2] = {0.0, 0.0}, tuneMs[2] = {0.0, 0.0};
    unsigned long long roundsWith[2] = {0, 0};
    unsigned long long stepsSeen = 0; /* the device's step counter at the end of the previous step launch */
    int ringCfg[RenderScratch::RING];
    int syncEvery = 8; /* rounds between two looks at the live-slot counter (a host synchronisation) */
    if (const char *e = getenv("MER_SYNC_EVERY")) syncEvery = std::max(atoi(e), 1);

    unsigned long long rounds = 0, launches = 0, stepLaunches = 0;
    double stepMs = 0.0;
    bool tailMarked = false;
    int ringUsed = 0; /* event pairs around the step kernel's launches since the last synchronisation */
    auto drain_ring = [&]() {
        for (int i = 0; i < ringUsed; i++) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, S.ring[2 * i], S.ring[2 * i + 1]) == cudaSuccess) stepMs += ms;
            const unsigned long long seen = S.hostPinned[4 + i];
            if (ringCfg[i] >= 0) { /* a round of the tuning window */
                tuneSteps[ringCfg[i]] += (double) (seen - stepsSeen);
                tuneMs[ringCfg[i]] += ms;
            }
            stepsSeen = seen;
        }
        ringUsed = 0;
    };
    MER_CUDA(cudaEventRecord(S.ev0, stream));
    MER_LAUNCH(k_pool_init, std::min(eventBlocks, 148u * 8u), 256, 0, stream, P.pool, pool);
    launches++;
    const bool dielectric = m->desc.boundary == MER_BOUNDARY_HDIELECTRIC, extras = P.frames > 1 || P.modulation || P.nee || P.lightMode;
    const bool sdfShape = m->dev.shapeType == MER_SHAPE_SDF, xform = m->dev.rif.hasXform != 0, packed = m->rif->mode != MER_RIF_TRICUBIC;
    while (true) {
        MER_CUDA(cudaMemsetAsync(S.nOut, 0, 2 * sizeof(unsigned), stream)); /* stepHead, live */
        /* ---- events: every slot whose path cannot take a plain full step */
#define MER_EVENT(MODE_, D_, T_, S_)                                                                                     \
    do {                                                                                                                 \
        if (xform) MER_LAUNCH((k_event<MODE_, D_, T_, S_, true>), eventBlocks, TPB, 0, stream, P);                       \
        else MER_LAUNCH((k_event<MODE_, D_, T_, S_, false>), eventBlocks, TPB, 0, stream, P);                            \
    } while (0)
#define MER_EVENT2(MODE_, S_)                                                                                            \
    do {                                                                                                                 \
        if (dielectric) { if (extras) MER_EVENT(MODE_, true, true, S_); else MER_EVENT(MODE_, true, false, S_); }        \
        else { if (extras) MER_EVENT(MODE_, false, true, S_); else MER_EVENT(MODE_, false, false, S_); }                 \
    } while (0)
        if (sdfShape) MER_EVENT2(MER_RIF_TRICUBIC, true); /* tricubic only (checked above) */
        else if (!packed) MER_EVENT2(MER_RIF_TRICUBIC, false);
        else MER_EVENT2(MER_RIF_TRILINEAR_PACKED, false);
#undef MER_EVENT2
#undef MER_EVENT
        launches++;
        rounds++;
        const bool look = P.nee || rounds % (unsigned long long) syncEvery == 0 || rounds <= 2;
        if (look) MER_CUDA(cudaMemcpyAsync(S.hostPinned, P.live, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        if (P.nee) {
            MER_CUDA(cudaMemcpyAsync(S.hostPinned + 2, S.neeCount, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
            MER_CUDA(cudaStreamSynchronize(stream));
            const unsigned nReq = std::min(*(unsigned *) (S.hostPinned + 2), P.neeCap);
            if (nReq) {
                RenderParams Pn = P; /* this launch's view: sorted only when the sort ran */
                if (!(P.neePerm && nReq > 32u)) Pn.neePerm = nullptr;
                if (Pn.neePerm) {
                    MER_CUDA(cudaMemsetAsync(S.neeHist, 0, NEE_BINS * sizeof(unsigned), stream));
                    MER_LAUNCH(k_nee_keys, (nReq + 255u) / 256u, 256, 0, stream, P, nReq);
                    MER_LAUNCH(k_nee_offsets, 1, 32, 0, stream, S.neeHist);
                    MER_LAUNCH(k_nee_scatter, (nReq + 255u) / 256u, 256, 0, stream, P, nReq);
                    launches += 3;
                }
                const unsigned nb = (nReq + TPB - 1) / TPB;
                const bool wantOpl = P.frames > 1 || P.modulation; /* optical length: transient / modulated film only */
                if (sdfShape) { if (wantOpl) MER_LAUNCH((k_nee<true, true>), nb, TPB, 0, stream, Pn, nReq); else MER_LAUNCH((k_nee<false, true>), nb, TPB, 0, stream, Pn, nReq); }
                else { if (wantOpl) MER_LAUNCH((k_nee<true, false>), nb, TPB, 0, stream, Pn, nReq); else MER_LAUNCH((k_nee<false, false>), nb, TPB, 0, stream, Pn, nReq); }
                launches++;
                MER_CUDA(cudaMemsetAsync(S.neeCount, 0, sizeof(unsigned), stream));
            }
            if (*(unsigned *) S.hostPinned == 0u) break;
        }
        /* ---- steps: every slot whose path can */
        if (ringUsed == RenderScratch::RING) { MER_CUDA(cudaStreamSynchronize(stream)); drain_ring(); }
        const int probe = probe_of(rounds);
        const StepCfg cfg = stepEnv ? stepFixed : stepCfgs[probe >= 0 ? probe : stepChoice];
        ringCfg[ringUsed] = probe;
        roundsWith[probe >= 0 ? probe : stepChoice]++;
        const unsigned stepBlocks = std::min((unsigned) sms * cfg.ctas, (pool + cfg.tpb - 1) / cfg.tpb), stepTpb = cfg.tpb;
        MER_CUDA(cudaEventRecord(S.ring[2 * ringUsed], stream));
#define MER_STEP_K(K_) MER_LAUNCH(K_, stepBlocks, stepTpb, 0, stream, P) /* no shared memory: the SM's whole array is L1 (an explicit carve-out request changed nothing) */
#define MER_STEP(MODE_, L_, T_, S_)                                                                                      \
    do {                                                                                                                 \
        if (xform) MER_STEP_K((k_step<MODE_, L_, T_, S_, true>)); else MER_STEP_K((k_step<MODE_, L_, T_, S_, false>));    \
    } while (0)
#define MER_STEP2(L_)                                                                                                    \
    do {                                                                                                                 \
        if (sdfShape) { if (extras) MER_STEP(MER_RIF_TRICUBIC, L_, true, true); else MER_STEP(MER_RIF_TRICUBIC, L_, false, true); }    \
        else { if (extras) MER_STEP(MER_RIF_TRICUBIC, L_, true, false); else MER_STEP(MER_RIF_TRICUBIC, L_, false, false); }           \
    } while (0)
        if (packed) { if (extras) MER_STEP(MER_RIF_TRILINEAR_PACKED, 0, true, false); else MER_STEP(MER_RIF_TRILINEAR_PACKED, 0, false, false); }
        else if (m->dev.rif.coeff8) MER_STEP2(1);
        else MER_STEP2(0);
#undef MER_STEP_K
#undef MER_STEP2
#undef MER_STEP
        MER_CUDA(cudaEventRecord(S.ring[2 * ringUsed + 1], stream));
        MER_CUDA(cudaMemcpyAsync(S.hostPinned + 4 + ringUsed, P.stats + ST_STEPS, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        ringUsed++;
        launches++;
        stepLaunches++;
        if (look && !P.nee) {
            MER_CUDA(cudaStreamSynchronize(stream));
            drain_ring();
            if (tuning && tuneMs[0] > 0.0 && tuneMs[1] > 0.0 && rounds >= probeAt + 3) {
                /* both candidates have rendered two rounds of the same mix of paths; 3 % of hysteresis */
                const double wide = tuneSteps[0] / tuneMs[0], narrow = tuneSteps[1] / tuneMs[1];
                const int before = stepChoice;
                if (narrow > 1.03 * wide) stepChoice = 1;
                else if (wide > 1.03 * narrow) stepChoice = 0;
                probeGap = stepChoice == before ? std::min(probeGap * 2ull, 512ull) : 32ull;
                probeAt = rounds + probeGap;
                tuneSteps[0] = tuneSteps[1] = tuneMs[0] = tuneMs[1] = 0.0;
            }
            if (*(unsigned *) S.hostPinned == 0u) break;
            if (!tailMarked && *(unsigned *) S.hostPinned < pool / 2u) {
                MER_CUDA(cudaEventRecord(S.evTail, stream));
                tailMarked = true;
            }
        }
    }
    const unsigned long long passes = rounds;
    if (P.lightMode) {
        const size_t npx = (size_t) r->width * r->height;
        MER_LAUNCH(k_film_unit_weight, (unsigned) std::min<size_t>(mer_blocks(npx, 256), 148u * 8u), 256, 0, stream, npx, P.channels,
                   (float) P.sppLocal / (float) r->spp_total, film_dev);
        launches++;
    }
    MER_CUDA(cudaEventRecord(S.ev1, stream));
    MER_CUDA(cudaEventSynchronize(S.ev1));
    if (tuning) { S.stepChoice = stepChoice; S.stepChoiceTable = tableBytes; }
    if (stats_out) {
        unsigned long long hs[1 + ST_COUNT];
        MER_CUDA(cudaMemcpy(hs, S.counters, sizeof(hs), cudaMemcpyDeviceToHost));
        memset(stats_out, 0, sizeof(*stats_out));
        stats_out->samples = hs[1 + ST_SAMPLES];
        stats_out->ray_steps = hs[1 + ST_STEPS];
        stats_out->scatter_events = hs[1 + ST_SCATTER];
        stats_out->null_collisions = hs[1 + ST_NULL];
        stats_out->boundary_exits = hs[1 + ST_EXIT];
        stats_out->nonfinite_dropped = hs[1 + ST_NONFINITE];
        stats_out->connections = hs[1 + ST_CONN];
        stats_out->connections_failed = hs[1 + ST_CONNFAIL];
        stats_out->connection_steps = hs[1 + ST_CONNSTEPS];
        stats_out->passes = passes;
        stats_out->kernel_launches = launches;
        drain_ring();
        stats_out->step_kernel_ms = (float) stepMs;
        stats_out->step_launches = stepLaunches;
        if (tailMarked) cudaEventElapsedTime(&stats_out->tail_ms, S.evTail, S.ev1);
        stats_out->block_fetches = hs[1 + ST_FETCH];
        const int most = roundsWith[1] > roundsWith[0] ? 1 : 0; /* the configuration most rounds ran with */
        stats_out->step_lanes_per_sm = stepEnv ? stepFixed.ctas * stepFixed.tpb : stepCfgs[most].ctas * stepCfgs[most].tpb;
        float ms = 0;
        cudaEventElapsedTime(&ms, S.ev0, S.ev1);
        stats_out->device_ms = ms;
    }
    return MER_OK;
}

int mer_render(const mer_medium *m, const mer_render_desc *r, float *film_host, mer_render_stats *stats_out) {
    MER_REQUIRE(m && r && film_host, "null argument");
    MER_REQUIRE(r->width > 0 && r->height > 0, "film size must be positive");
    mer::DeviceGuard guard(m->device);
    float *film = nullptr;
    const size_t bytes = (size_t) r->width * r->height * (3 * (size_t) ((r->frames > 1 && !r->modulation) ? r->frames : 1) + 2) * sizeof(float);
    MER_CUDA(cudaMalloc(&film, bytes));
    cudaError_t e = cudaMemset(film, 0, bytes);
    int rc = e == cudaSuccess ? mer_render_device(m, r, film, stats_out, nullptr) : mer::fail(MER_ERR_CUDA, cudaGetErrorString(e));
    if (rc == MER_OK) {
        e = cudaMemcpy(film_host, film, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = mer::fail(MER_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(film);
    return rc;
}

int mer_film_develop_frames(int device, int32_t width, int32_t height, int32_t frames, const float *film, float *rgb_out) {
    MER_REQUIRE(film && rgb_out && width > 0 && height > 0 && frames > 0, "bad argument");
    int rc = mer::check_device(device);
    if (rc) return rc;
    mer::DeviceGuard guard(device);
    const size_t n = (size_t) width * height, C = 3 * (size_t) frames + 2, V = 3 * (size_t) frames;
    float *df = nullptr, *dr = nullptr;
    MER_CUDA(cudaMalloc(&df, n * C * sizeof(float)));
    MER_CUDA(cudaMalloc(&dr, n * V * sizeof(float)));
    cudaError_t e = cudaMemcpy(df, film, n * C * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_develop<<<(unsigned) std::min<size_t>(mer_blocks(n * V, 256), 148u * 8u), 256>>>(n, frames, df, dr);
        mer::g_launches.fetch_add(1);
        e = cudaMemcpy(rgb_out, dr, n * V * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(df);
    cudaFree(dr);
    if (e != cudaSuccess) return mer::fail(MER_ERR_CUDA, cudaGetErrorString(e));
    return MER_OK;
}

int mer_film_develop(int device, int32_t width, int32_t height, const float *film, float *rgb_out) {
    return mer_film_develop_frames(device, width, height, 1, film, rgb_out);
}

} /* extern "C" */
