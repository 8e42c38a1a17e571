/*
 * oracle/ref_volume.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).  Part of oracle/_ref/libmer_reftrace.so.
 *
 * SplineDataSource's lookups (src/volume/splinevolume.cpp:319-377: insideVolumeLimits, maxSDFError, value, gradient, hessian,
 * valueAndGradient, gradientAndHessian, valueGradientAndHessian), cut out of the .cpp by oracle/Makefile into
 * oracle/_ref/splinevolume_extract.inc and compiled VERBATIM inside a struct that declares the data members they use, on top
 * of the reference's own basisspline.h / transform.h / matrix.h / aabb.h; AABB::getCorner comes from src/libcore/aabb.cpp
 * the same way.  Nothing is copied into this repo.
 */
#include "ref_volume.h"
namespace mitsuba { using std::endl; }
#include <mitsuba/core/basisspline.h>
namespace mitsuba { extern bool solveQuadratic(Float a, Float b, Float c, Float &x0, Float &x1); } /* util.h; bsphere.h mentions it */
#include <mitsuba/core/aabb.h>

namespace mitsuba {
#include "aabb_extract.inc" /* generated: AABB::getCorner from src/libcore/aabb.cpp */

struct RefSplineDataSource : public RefVolume {
    basisspline::Spline<3> m_spline;
    Transform m_worldToVolume;                 /* toWorld = identity: Transform() */
    Matrix3x3F m_worldToVolume_Rot, m_worldToVolume_RotT;
    AABB m_interpolatableLimits;
    Float m_maxSDFError;
#include "splinevolume_extract.inc" /* generated: the eight lookups; same signatures as RefVolume's, so they override them */
};

RefVolume *ref_make_volume(const float *data, const int *N, const float *bmin, const float *bmax) {
    RefSplineDataSource *rif = new RefSplineDataSource();
    FLOAT xmin[3], xmax[3];
    int n[3];
    for (int i = 0; i < 3; i++) { xmin[i] = (FLOAT) bmin[i]; xmax[i] = (FLOAT) bmax[i]; n[i] = N[i]; }
    rif->m_spline.initialize(xmin, xmax, n);
    /* splinevolume.cpp:280-282 */
    rif->m_interpolatableLimits = AABB(Point(xmin[0], xmin[1], xmin[2]) + Point( 2.0*rif->m_spline.getStride(0)+Epsilon,  2.0*rif->m_spline.getStride(1)+Epsilon,  2.0*rif->m_spline.getStride(2)+Epsilon),
                                       Point(xmax[0], xmax[1], xmax[2]) + Point(-2.0*rif->m_spline.getStride(0)-Epsilon, -2.0*rif->m_spline.getStride(1)-Epsilon, -2.0*rif->m_spline.getStride(2)-Epsilon));
    rif->m_maxSDFError = std::sqrt( rif->m_spline.getStride(0)*rif->m_spline.getStride(0) + rif->m_spline.getStride(1)*rif->m_spline.getStride(1) + rif->m_spline.getStride(2)*rif->m_spline.getStride(2));
    size_t total = (size_t) N[0] * N[1] * N[2];
    FLOAT *tmp = new FLOAT[total];
    for (size_t i = 0; i < total; i++) tmp[i] = (FLOAT) data[i]; /* splinevolume.cpp:284-287 */
    rif->m_spline.build(tmp);
    delete[] tmp;
    rif->m_worldToVolume_Rot.setIdentity();  /* :90-92 with an identity toWorld */
    rif->m_worldToVolume_RotT.setIdentity();
    return rif;
}
}
