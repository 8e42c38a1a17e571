/*
 * oracle/ref_trace.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles, VERBATIM and from where they lie under /root/reference (nothing is copied into this repo), the stepper of the
 * hot path:
 *     src/medium/heterogeneousrefractive.cpp   er_step (both overloads, :653-669), trace (:671-691), aggressive_trace
 *                                              (:697-704), insideShape / hackForSphere / hackForBox (:707-726),
 *                                              traceTillBoundary (:742-776), sampleDistance with its aggressive-tracing
 *                                              loop (:402-568), evalTransmittance (:393-400), enum ESamplingStrategy
 *     src/medium/maxexp.h                      MaxExpDist (strategy "maximum"), as it is
 *                                              er_derivativestep (:798-814), computefdfBDPT (:816-939),
 *                                              computePathLengthsTillClosestP2 (:941-1030), boundaryVelocity (:1036-1051), boundaryVelocityDerivative
 *                                              (:1057-1074), sgn (:163-165)   [the shooting problem's residual, its Jacobian
 *                                              and the connection's lengths: everything of a25 except the Ceres solver]
 *     src/volume/splinevolume.cpp              SplineDataSource's lookups (:319-377), in ref_volume.cpp
 *     src/libcore/aabb.cpp                     AABB::getCorner (:23-27), in ref_volume.cpp
 * The member-function bodies are cut out of the two .cpp files by oracle/Makefile (awk, by signature) into
 * hetref_extract.inc and splinevolume_extract.inc (a temporary directory, deleted after the build) and included into two structs (two translation units) that declare
 * exactly the data members those bodies use.  Underneath them the reference's own headers are compiled as they are:
 * include/mitsuba/core/{basisspline,transform,matrix,ray,aabb,vector,point,normal,math,constants,fwd,platform}.h
 * (oracle/shim_phase stands in for mitsuba.h / stream.h, oracle/shim_trace for the one Boost header matrix.h asks for).
 * Release flags of the reference: -DSINGLE_PRECISION -DSPECTRUM_SAMPLES=3, so FLOAT = Float = float.
 *
 * NOTE the reference's insideShape() is hackForSphere(): a hard-coded sphere, centre (-0.22827, 1.2, 0.152505), radius 0.3,
 * strict inequality (:707-718).  The parity tests put their medium in exactly that sphere.
 *
 * Built into oracle/_ref/libmer_reftrace.so; tests/test_oracle_cpu.py checks the restated er_step / trace /
 * traceTillBoundary of oracle/mer_oracle.cpp against it bit for bit, tests/golden/trace_ref.npz holds vectors generated
 * from it (tests/golden/make_trace_golden.py), and tests/test_gpu_medium.py compares the CUDA stepper with it.
 */
#include "ref_volume.h"
#include "ref_records.h"
#include <mitsuba/core/ray.h>        /* reference */
#include <mitsuba/core/spectrum.h>   /* reference */
#include <mitsuba/render/sampler.h>  /* oracle/shim_phase: next1D / next2D */
#include <medium/maxexp.h>           /* reference (src/medium/maxexp.h): MaxExpDist */
#include <mitsuba/core/properties.h> /* oracle/shim_phase */

namespace mitsuba {
#include "sgn_extract.inc" /* generated: template <typename T> FLOAT sgn(T) of heterogeneousrefractive.cpp:163-165 */

/* HeterogeneousRefractiveMedium (src/medium/heterogeneousrefractive.cpp) reduced to what its stepper uses */
struct RefHeterogeneousRefractiveMedium {
    RefVolume *m_rif, *m_SDF;
    FLOAT m_erstepsize, m_tol;
    int m_precision;
    bool m_aggressiveTracing;
    Spectrum m_sigmaA, m_sigmaS, m_sigmaT;
    Float m_samplingDensity, m_mediumSamplingWeight;
    MaxExpDist *m_maxExpDist;
#include "hetref_extract.inc" /* generated: enum ESamplingStrategy, evalTransmittance, sampleDistance, er_step x2, trace,
                                 aggressive_trace, insideShape, hackForSphere, hackForBox, traceTillBoundary, er_derivativestep,
                                 computefdfBDPT, computePathLengthsTillClosestP2, boundaryVelocity, boundaryVelocityDerivative */
    ESamplingStrategy m_strategy;
    /* the part of the constructor (:238-293) that resolves mediumSamplingWeight, the sampling strategy and its density */
    void resolve(const Properties &props, const std::string &strategy) {
#include "hetref_ctor_extract.inc" /* generated */
    }
};

}

using namespace mitsuba;

extern "C" {

void *ref_medium_create(const float *data, const int *N, const float *bmin, const float *bmax, float stepsize) {
    RefVolume *rif = ref_make_volume(data, N, bmin, bmax);
    RefHeterogeneousRefractiveMedium *m = new RefHeterogeneousRefractiveMedium();
    m->m_rif = rif;
    m->m_SDF = rif; /* only its insideVolumeLimits() is asked until ref_medium_configure() gives it a signed distance */
    m->m_erstepsize = stepsize;
    m->m_precision = 6;
    m->m_tol = 1e-6f;
    m->m_aggressiveTracing = false;
    m->m_maxExpDist = NULL;
    m->m_strategy = RefHeterogeneousRefractiveMedium::ESingle;
    m->m_samplingDensity = 1;
    m->m_mediumSamplingWeight = 0.5f;
    return m;
}

/* what the constructor resolves from the properties (heterogeneousrefractive.cpp:201-300), handed over resolved:
 * strategy 0 balance / 1 single / 2 manual / 3 maximum; sdf may be NULL (then no aggressive tracing) */
void ref_medium_configure(void *h, const float *sigmaA, const float *sigmaS, int strategy, float samplingDensity, float mediumSamplingWeight,
                          const float *sdf, const int *N, const float *bmin, const float *bmax, int aggressive) {
    RefHeterogeneousRefractiveMedium *m = (RefHeterogeneousRefractiveMedium *) h;
    for (int i = 0; i < 3; i++) { m->m_sigmaA[i] = sigmaA[i]; m->m_sigmaS[i] = sigmaS[i]; }
    m->m_sigmaT = m->m_sigmaA + m->m_sigmaS; /* :222 */
    m->m_strategy = (RefHeterogeneousRefractiveMedium::ESamplingStrategy) strategy;
    m->m_samplingDensity = samplingDensity;
    m->m_mediumSamplingWeight = mediumSamplingWeight;
    if (m->m_maxExpDist) { delete m->m_maxExpDist; m->m_maxExpDist = NULL; }
    if (strategy == 3) { /* :282-286 */
        std::vector<Float> coeffs(SPECTRUM_SAMPLES);
        for (int i = 0; i < SPECTRUM_SAMPLES; i++) coeffs[i] = m->m_sigmaT[i];
        m->m_maxExpDist = new MaxExpDist(coeffs);
    }
    if (m->m_SDF != m->m_rif) delete m->m_SDF;
    m->m_SDF = sdf ? ref_make_volume(sdf, N, bmin, bmax) : m->m_rif;
    m->m_aggressiveTracing = sdf && aggressive;
}

/* the constructor's resolution of the free-flight sampling (:238-293) from sigmaA / sigmaS and the properties
 * mediumSamplingWeight (-1 = automatic), strategy, samplingDensity (manual), channel (-1 = automatic) */
void ref_medium_resolve(void *h, const float *sigmaA, const float *sigmaS, const char *strategy, float mediumSamplingWeight, float samplingDensity,
                        int channel, float *weightOut, float *densityOut, int *strategyOut) {
    RefHeterogeneousRefractiveMedium *m = (RefHeterogeneousRefractiveMedium *) h;
    for (int i = 0; i < 3; i++) { m->m_sigmaA[i] = sigmaA[i]; m->m_sigmaS[i] = sigmaS[i]; }
    m->m_sigmaT = m->m_sigmaA + m->m_sigmaS; /* Medium::Medium, src/librender/medium.cpp:36 */
    m->m_samplingDensity = 0.0f;             /* initialiser list, :202 */
    if (m->m_maxExpDist) { delete m->m_maxExpDist; m->m_maxExpDist = NULL; }
    Properties props;
    if (mediumSamplingWeight != -1) props.floats["mediumSamplingWeight"] = mediumSamplingWeight;
    if (samplingDensity > 0) props.floats["samplingDensity"] = samplingDensity;
    if (channel >= 0) props.floats["channel"] = (float) channel;
    m->resolve(props, strategy);
    *weightOut = m->m_mediumSamplingWeight;
    *densityOut = m->m_samplingDensity;
    *strategyOut = (int) m->m_strategy;
}

/* the medium over a .vol file read by the reference's loader */
void *ref_medium_create_from_file(const char *path, float stepsize) {
    RefVolume *rif = ref_load_volume(path);
    RefHeterogeneousRefractiveMedium *m = new RefHeterogeneousRefractiveMedium();
    m->m_rif = rif;
    m->m_SDF = rif;
    m->m_erstepsize = stepsize;
    m->m_precision = 6;
    m->m_tol = 1e-6f;
    m->m_aggressiveTracing = false;
    m->m_maxExpDist = NULL;
    m->m_strategy = RefHeterogeneousRefractiveMedium::ESingle;
    m->m_samplingDensity = 1;
    m->m_mediumSamplingWeight = 0.5f;
    return m;
}

void ref_medium_free(void *h) {
    RefHeterogeneousRefractiveMedium *m = (RefHeterogeneousRefractiveMedium *) h;
    if (m->m_SDF != m->m_rif) delete m->m_SDF;
    delete m->m_rif;
    delete m->m_maxExpDist;
    delete m;
}

namespace {
struct ReplaySampler : public mitsuba::Sampler { /* next1D() replays the given numbers in order */
    const float *xi;
    int k;
    mitsuba::Float next1D() { return xi[k++]; }
    mitsuba::Point2 next2D() { mitsuba::Float a = xi[k++], b = xi[k++]; return mitsuba::Point2(a, b); }
};
}

/* sampleDistance (:402-568) over n rays; xi[n][2] are the numbers sampler->next1D() returns, in order */
void ref_sample_distance(void *h, size_t n, const float *ro, const float *rd, const float *mint, const float *xi, int *success, float *t,
                         float *p, float *d, float *opticalLength, float *refRatioSq, float *transmittance, float *pdfSuccess,
                         float *pdfFailure) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long) n; i++) {
        Ray ray(Point(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), Vector(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), 0.0f);
        ray.mint = mint[i];
        MediumSamplingRecord mRec;
        mRec.t = mRec.opticalLength = mRec.refRatioSq = 0;
        mRec.pdfSuccess = mRec.pdfSuccessRev = mRec.pdfFailure = 0;
        ReplaySampler s;
        s.xi = xi + 2 * i;
        s.k = 0;
        success[i] = m->sampleDistance(ray, mRec, &s) ? 1 : 0;
        t[i] = mRec.t;
        p[3 * i] = mRec.p.x; p[3 * i + 1] = mRec.p.y; p[3 * i + 2] = mRec.p.z;
        d[3 * i] = mRec.d.x; d[3 * i + 1] = mRec.d.y; d[3 * i + 2] = mRec.d.z;
        opticalLength[i] = mRec.opticalLength;
        refRatioSq[i] = mRec.refRatioSq;
        for (int c = 0; c < 3; c++) transmittance[3 * i + c] = mRec.transmittance[c];
        pdfSuccess[i] = mRec.pdfSuccess;
        pdfFailure[i] = mRec.pdfFailure;
    }
}

/* evalTransmittance (:393-400) */
void ref_eval_transmittance(void *h, size_t n, const float *mint, const float *maxt, float *out) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) {
        Ray ray(Point(0.0f), Vector(0.0f, 0.0f, 1.0f), 0.0f);
        ray.mint = mint[i];
        ray.maxt = maxt[i];
        Spectrum T = m->evalTransmittance(ray, NULL);
        for (int c = 0; c < 3; c++) out[3 * i + c] = T[c];
    }
}

int ref_medium_sizeof_float(void) { return (int) sizeof(FLOAT); }

/* the hard-coded container: centre and radius as the compiled code has them */
void ref_medium_container(float *centre_radius) {
    centre_radius[0] = -0.22827f; centre_radius[1] = 1.2f; centre_radius[2] = 0.152505f; centre_radius[3] = 0.3f;
}

/* SplineDataSource wrappers over n points */
void ref_rif_value_gradient(void *h, size_t n, const float *p, float *f, float *g) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) {
        FLOAT fv;
        VectorF G;
        m->m_rif->valueAndGradient(PointF(p[3 * i], p[3 * i + 1], p[3 * i + 2]), fv, G);
        f[i] = fv;
        g[3 * i] = G.x; g[3 * i + 1] = G.y; g[3 * i + 2] = G.z;
    }
}
void ref_rif_inside_limits(void *h, size_t n, const float *p, int *out) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) out[i] = m->m_rif->insideVolumeLimits(PointF(p[3 * i], p[3 * i + 1], p[3 * i + 2])) ? 1 : 0;
}
void ref_inside_shape(void *h, size_t n, const float *p, int *out) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) out[i] = m->insideShape(PointF(p[3 * i], p[3 * i + 1], p[3 * i + 2])) ? 1 : 0;
}

/* one er_step with the given step size (in place; opl accumulated) */
void ref_er_step(void *h, size_t n, float *p, float *v, const float *stepsize, float *opl) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) {
        PointF P(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        VectorF V(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        FLOAT o = opl[i];
        m->er_step(P, V, stepsize[i], o);
        p[3 * i] = P.x; p[3 * i + 1] = P.y; p[3 * i + 2] = P.z;
        v[3 * i] = V.x; v[3 * i + 1] = V.y; v[3 * i + 2] = V.z;
        opl[i] = o;
    }
}

/* trace (:671-691), in place */
void ref_trace(void *h, size_t n, float *p, float *v, const float *dist, float *dist_surf, float *opl, int *success) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long) n; i++) {
        PointF P(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        VectorF V(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        FLOAT ds = 0, o = opl[i];
        success[i] = m->trace(P, V, dist[i], ds, o) ? 1 : 0;
        p[3 * i] = P.x; p[3 * i + 1] = P.y; p[3 * i + 2] = P.z;
        v[3 * i] = V.x; v[3 * i + 1] = V.y; v[3 * i + 2] = V.z;
        dist_surf[i] = ds;
        opl[i] = o;
    }
}

/* traceTillBoundary (:742-776), in place */
void ref_trace_till_boundary(void *h, size_t n, float *p, float *v, float *dist_surf, float *opl) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long) n; i++) {
        PointF P(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        VectorF V(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        FLOAT ds = 0, o = opl[i];
        m->traceTillBoundary(P, V, ds, o);
        p[3 * i] = P.x; p[3 * i + 1] = P.y; p[3 * i + 2] = P.z;
        v[3 * i] = V.x; v[3 * i + 1] = V.y; v[3 * i + 2] = V.z;
        dist_surf[i] = ds;
        opl[i] = o;
    }
}

/* the medium's connection parameters: boundaryprecision and tol (heterogeneousrefractive.cpp:242-250) */
void ref_medium_set_connection(void *h, int precision, float tol) {
    RefHeterogeneousRefractiveMedium *m = (RefHeterogeneousRefractiveMedium *) h;
    m->m_precision = precision;
    m->m_tol = tol;
}

/* er_derivativestep (:798-814) nsteps times, in place; dpdv0 / dvdv0 row-major 3x3 */
void ref_derivative_trace(void *h, size_t n, float *p, float *v, float *dpdv0, float *dvdv0, int nsteps) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) {
        PointF P(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        VectorF V(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        Matrix3x3F A, B;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { A.m[r][c] = dpdv0[9 * i + 3 * r + c]; B.m[r][c] = dvdv0[9 * i + 3 * r + c]; }
        for (int k = 0; k < nsteps; k++) m->er_derivativestep(P, V, A, B, m->m_erstepsize);
        p[3 * i] = P.x; p[3 * i + 1] = P.y; p[3 * i + 2] = P.z;
        v[3 * i] = V.x; v[3 * i + 1] = V.y; v[3 * i + 2] = V.z;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { dpdv0[9 * i + 3 * r + c] = A.m[r][c]; dvdv0[9 * i + 3 * r + c] = B.m[r][c]; }
    }
}

/* computefdfBDPT (:816-939) exactly as DirectConnectionCostFunction::Evaluate (:1247-1280) calls it: dpdv0 = 0, dvdv0 = I
 * (makeDirectConnections :1088-1091); err[3], derr row-major 3x3 (= derror.m) */
void ref_connection_residual(void *h, size_t n, const float *p1, const float *p2, const float *v0, int isSensor, float *err, float *derr) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < (long) n; i++) {
        Matrix3x3F dpdv0((FLOAT)0);
        Matrix3x3F dvdv0((FLOAT)1, 0, 0,
                        0, 1, 0,
                        0, 0, 1);
        VectorF error(0.0);
        Matrix3x3F derror(0.0);
        bool isSensorSample = isSensor != 0;
        m->computefdfBDPT(VectorF(v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]), PointF(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]),
                          PointF(p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]), isSensorSample, dpdv0, dvdv0, error, derror);
        err[3 * i] = error.x; err[3 * i + 1] = error.y; err[3 * i + 2] = error.z;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) derr[9 * i + 3 * r + c] = derror.m[r][c];
    }
}

/* computePathLengthsTillClosestP2 (:941-1051) */
void ref_path_lengths(void *h, size_t n, const float *p1, const float *p2, const float *dirToP2, int isSensor, float *revDir, float *opticalDist,
                      float *dist, int *ok) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < (long) n; i++) {
        VectorF rev(0.0);
        FLOAT od = 0, d = 0;
        bool isSensorSample = isSensor != 0;
        ok[i] = m->computePathLengthsTillClosestP2(PointF(p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]), PointF(p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]),
                                                   VectorF(dirToP2[3 * i], dirToP2[3 * i + 1], dirToP2[3 * i + 2]), rev, isSensorSample, od, d) ? 1 : 0;
        revDir[3 * i] = rev.x; revDir[3 * i + 1] = rev.y; revDir[3 * i + 2] = rev.z;
        opticalDist[i] = od;
        dist[i] = d;
    }
}

/* aggressive_trace (:697-704), in place */
void ref_aggressive_trace(void *h, size_t n, float *p, float *v, const float *dist, float *opl) {
    const RefHeterogeneousRefractiveMedium *m = (const RefHeterogeneousRefractiveMedium *) h;
    for (size_t i = 0; i < n; i++) {
        PointF P(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        VectorF V(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        FLOAT o = opl[i];
        m->aggressive_trace(P, V, dist[i], o);
        p[3 * i] = P.x; p[3 * i + 1] = P.y; p[3 * i + 2] = P.z;
        v[3 * i] = V.x; v[3 * i + 1] = V.y; v[3 * i + 2] = V.z;
        opl[i] = o;
    }
}

} /* extern "C" */
