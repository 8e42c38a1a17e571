/*
 * oracle/ref_records.h - TEST INFRASTRUCTURE ONLY.  MediumSamplingRecord (include/mitsuba/render/medium.h:36-108) reduced to the
 * data members Medium::sampleDistance() of heterogeneousrefractive.cpp and heterogeneous.cpp write.
 */
#pragma once
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/spectrum.h> /* reference */

namespace mitsuba {
struct MediumSamplingRecord {
    Float t, opticalLength;
    Point p;
    Vector d;
    Float time;
    Vector orientation;
    Spectrum transmittance, sigmaA, sigmaS;
    Float pdfSuccess, pdfSuccessRev, pdfFailure;
    const void *medium;
    Float refRatioSq;
};
}
