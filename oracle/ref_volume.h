/*
 * oracle/ref_volume.h - TEST INFRASTRUCTURE ONLY.  The virtual lookups of VolumeDataSource (include/mitsuba/render/volume.h)
 * that HeterogeneousRefractiveMedium calls on its `rif` and `sdf` children, so that the medium's functions (ref_trace.cpp) and
 * SplineDataSource's (ref_volume.cpp) compile in separate translation units as they do in the reference:
 * heterogeneousrefractive.cpp defines its own `template <typename T> FLOAT sgn(T)` that would collide with basisspline.h's.
 */
#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba { using std::endl; } /* mitsuba.h brings it in; ray.h writes a bare `endl` */
#include <mitsuba/core/matrix.h> /* reference */

namespace mitsuba {
struct RefVolume {
    virtual ~RefVolume() {}
    virtual bool insideVolumeLimits(const PointF &p) const = 0;
    virtual Float maxSDFError() const = 0;
    virtual FLOAT value(const PointF &p) const = 0;
    virtual VectorF gradient(const PointF &p) const = 0;
    virtual Matrix3x3F hessian(const PointF &p) const = 0;
    virtual void valueAndGradient(const PointF &p, FLOAT &f, VectorF &v) const = 0;
    virtual void gradientAndHessian(const PointF &p, VectorF &v, Matrix3x3F &M) const = 0;
    virtual void valueGradientAndHessian(const PointF &p, FLOAT &f, VectorF &v, Matrix3x3F &M) const = 0;
};
/* SplineDataSource over a float grid [z][y][x] with an identity toWorld (ref_volume.cpp) */
RefVolume *ref_make_volume(const float *data, const int *N, const float *bmin, const float *bmax);
/* the same from a .vol file, through SplineDataSource::loadFromFile (ref_volume.cpp) */
RefVolume *ref_load_volume(const char *path);
}
