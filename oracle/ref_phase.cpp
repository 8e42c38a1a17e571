/*
 * oracle/ref_phase.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles, VERBATIM and from where they lie under /root/reference (nothing is copied into this repo):
 *     src/phase/hg.cpp                              HGPhaseFunction::sample / eval            
 *     include/mitsuba/core/frame.h                  Frame(n), Frame::toWorld                  (a17)
 *     include/mitsuba/core/{vector,point,normal,math,constants}.h    dot / cross / safe_sqrt / sincos / Epsilon ...
 *     src/medium/maxexp.h                           MaxExpDist (strategy "maximum")            (a12)
 *     src/libcore/util.cpp  coordinateSystem()      (a17)   } the two function bodies are cut out of util.cpp by
 *     src/bsdfs/hdielectric.cpp                     HSmoothDielectric::sample(bRec, p, sample) (:244-300), reflect (:86-88),
 *                                                   refract (:119-125), getEtaInvEta (:113-117)      (SURVEY f-3)
 *     src/libcore/util.cpp  fresnelDielectricExt()  (f-3)   } oracle/Makefile into util_extract.inc (temporary dir)
 * behind a C ABI, with the reference's release flags (-DSINGLE_PRECISION -DSPECTRUM_SAMPLES=3).  Built into
 * oracle/_ref/libmer_refphase.so; tests/test_oracle_cpu.py checks the restated HG / coordinateSystem / Fresnel of
 * oracle/mer_oracle.cpp against it bit for bit, and tests/golden/phase_ref.npz holds vectors generated from it.
 */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/frame.h>
#include <mitsuba/render/phase.h>
#include <mitsuba/render/sampler.h>

namespace mitsuba {
#include "util_extract.inc" /* generated: coordinateSystem, fresnelDielectricExt from src/libcore/util.cpp */
}

#include <phase/hg.cpp> /* -I/root/reference/src */
#include <medium/maxexp.h> /* MaxExpDist, the "maximum" free-flight strategy (heterogeneousrefractive.cpp:287-291, 437-445, 534-536) */

#include <mitsuba/core/spectrum.h> /* reference */
namespace mitsuba {
/* what HSmoothDielectric::sample names of the framework, reduced (the enumerators' values: include/mitsuba/render/bsdf.h:240-242,
 * include/mitsuba/render/common.h:38-40) */
enum { EDeltaReflection = 0x00020, EDeltaTransmission = 0x00040 };
enum ETransportMode { ERadiance = 0, EImportance = 1 };
struct Intersection {};
struct BSDFSamplingRecord {
    Intersection its;
    Vector wi, wo;
    Float eta;
    ETransportMode mode;
    unsigned int typeMask;
    int component, sampledComponent;
    unsigned int sampledType;
};
struct RefUnitTexture { Spectrum eval(const Intersection &) const { return Spectrum(1.0f); } };
struct RefInteriorMedium { Float rif; Float getRIF(const Point &) const { return rif; } };
struct RefShapeWithMedium { RefInteriorMedium medium; const RefInteriorMedium *getInteriorMedium() const { return &medium; } };
struct RefHSmoothDielectric {
    RefShapeWithMedium *m_shape;
    RefUnitTexture *m_specularReflectance, *m_specularTransmittance;
#include "hdielectric_extract.inc" /* generated: reflect, getEtaInvEta, refract, sample(bRec, p, sample) */
};
}

namespace {
struct FixedSampler : public mitsuba::Sampler {
    mitsuba::Float u1, u2;
    mitsuba::Float next1D() { return u1; }
    mitsuba::Point2 next2D() { return mitsuba::Point2(u1, u2); }
};
}

extern "C" {

/* HGPhaseFunction::sample for n (wi, xi) pairs; pdf = eval of the sampled direction (hg.cpp:100-105) */
void ref_hg_sample(float g, size_t n, const float *wi, const float *xi, float *wo, float *pdf) {
    mitsuba::Properties props;
    props.floats["g"] = g;
    mitsuba::HGPhaseFunction hg(props);
    hg.configure();
    FixedSampler s;
    for (size_t i = 0; i < n; i++) {
        mitsuba::PhaseFunctionSamplingRecord rec;
        rec.wi = mitsuba::Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        s.u1 = xi[2 * i];
        s.u2 = xi[2 * i + 1];
        mitsuba::Float p;
        hg.sample(rec, p, &s);
        wo[3 * i] = rec.wo.x; wo[3 * i + 1] = rec.wo.y; wo[3 * i + 2] = rec.wo.z;
        pdf[i] = p;
    }
}

void ref_hg_eval(float g, size_t n, const float *wi, const float *wo, float *out) {
    mitsuba::Properties props;
    props.floats["g"] = g;
    mitsuba::HGPhaseFunction hg(props);
    for (size_t i = 0; i < n; i++) {
        mitsuba::PhaseFunctionSamplingRecord rec;
        rec.wi = mitsuba::Vector(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        rec.wo = mitsuba::Vector(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        out[i] = hg.eval(rec);
    }
}

void ref_coordinate_system(size_t n, const float *a, float *b, float *c) {
    for (size_t i = 0; i < n; i++) {
        mitsuba::Vector bb, cc;
        mitsuba::coordinateSystem(mitsuba::Vector(a[3 * i], a[3 * i + 1], a[3 * i + 2]), bb, cc);
        b[3 * i] = bb.x; b[3 * i + 1] = bb.y; b[3 * i + 2] = bb.z;
        c[3 * i] = cc.x; c[3 * i + 1] = cc.y; c[3 * i + 2] = cc.z;
    }
}

void ref_fresnel_dielectric_ext(size_t n, const float *cosThetaI, const float *eta, float *F, float *cosThetaT) {
    for (size_t i = 0; i < n; i++) {
        mitsuba::Float ct;
        F[i] = mitsuba::fresnelDielectricExt(cosThetaI[i], ct, eta[i]);
        cosThetaT[i] = ct;
    }
}

/* HSmoothDielectric::sample for n rays arriving along d (world space, unit) at a surface with unit normal N and interior index eta;
 * u = sample.x; mode 0 radiance / 1 importance.  World <-> local through Frame(N) as Intersection::toLocal / toWorld do
 * (wi = toLocal(-d)).  Out: the sampled direction in world space, the returned weight (first channel), bRec.eta, transmitted? */
void ref_hdielectric_sample(size_t n, const float *d, const float *N, const float *eta, const float *u, int mode, float *dOut, float *weight,
                            float *etaScale, int *transmitted) {
    mitsuba::RefShapeWithMedium shape;
    mitsuba::RefUnitTexture one;
    mitsuba::RefHSmoothDielectric bsdf;
    bsdf.m_shape = &shape;
    bsdf.m_specularReflectance = bsdf.m_specularTransmittance = &one;
    for (size_t i = 0; i < n; i++) {
        shape.medium.rif = eta[i];
        mitsuba::Frame frame(mitsuba::Vector(N[3 * i], N[3 * i + 1], N[3 * i + 2]));
        mitsuba::BSDFSamplingRecord bRec;
        bRec.wi = frame.toLocal(mitsuba::Vector(-d[3 * i], -d[3 * i + 1], -d[3 * i + 2]));
        bRec.mode = mode ? mitsuba::EImportance : mitsuba::ERadiance;
        bRec.typeMask = 0xffffffffu; /* BSDF::EAll */
        bRec.component = -1;
        mitsuba::Spectrum w = bsdf.sample(bRec, mitsuba::Point(0.0f), mitsuba::Point2(u[i], 0.5f));
        mitsuba::Vector wo = frame.toWorld(bRec.wo);
        dOut[3 * i] = wo.x; dOut[3 * i + 1] = wo.y; dOut[3 * i + 2] = wo.z;
        weight[i] = w[0];
        etaScale[i] = bRec.eta;
        transmitted[i] = bRec.sampledType == (unsigned) mitsuba::EDeltaTransmission ? 1 : 0;
    }
}

/* MaxExpDist over a batch: what = 0 sample(u) -> (t, pdf), 1 pdf(t), 2 cdf(t) */
int ref_maxexp(const float sigmaT[3], int what, size_t n, const float *in, float *out0, float *out1) {
    std::vector<mitsuba::Float> st(sigmaT, sigmaT + 3);
    try {
        mitsuba::MaxExpDist mx(st);
        for (size_t i = 0; i < n; i++) {
            if (what == 0) out0[i] = mx.sample(in[i], out1[i]);
            else if (what == 1) out0[i] = mx.pdf(in[i]);
            else out0[i] = mx.cdf(in[i]);
        }
    } catch (const std::exception &) { return 1; }
    return 0;
}

} /* extern "C" */
