/*
 * oracle/ref_volume.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).  Part of oracle/_ref/libmer_reftrace.so.
 *
 * SplineDataSource's lookups (src/volume/splinevolume.cpp:319-377: insideVolumeLimits, maxSDFError, value, gradient, hessian,
 * valueAndGradient, gradientAndHessian, valueGradientAndHessian), cut out of the .cpp by oracle/Makefile into
 * oracle/_ref/splinevolume_extract.inc and compiled VERBATIM inside a struct that declares the data members they use, on top
 * of the reference's own basisspline.h / transform.h / matrix.h / aabb.h; AABB::getCorner comes from src/libcore/aabb.cpp
 * the same way.  GridDataSource::lookupFloat (src/volume/gridvolume.cpp:337-388) likewise, with enum EVolumeType
 * (:101-106) and the three Transform functions its configure() uses (src/libcore/transform.cpp:28-65: operator*, translate,
 * scale).  Nothing is copied into this repo.
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

#include "transform_extract.inc" /* generated: Transform::operator*, Transform::translate, Transform::scale */

/* GridDataSource (src/volume/gridvolume.cpp) reduced to the data members lookupFloat() uses */
struct RefGridDataSource {
#include "gridvolume_extract.inc" /* generated: enum EVolumeType, lookupFloat */
    Transform m_worldToGrid;
    Vector3i m_res;
    EVolumeType m_volumeType;
    uint8_t *m_data;
    Float m_densityMap[256];
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

/* ---- C entry points of the density grid */
extern "C" {
/* type: 1 = float32 payload, 3 = uint8 payload (EVolumeType); data [z][y][x] */
void *ref_grid_create(const void *data, const int *N, const float *bmin, const float *bmax, int type) {
    RefGridDataSource *g = new RefGridDataSource();
    g->m_res = Vector3i(N[0], N[1], N[2]);
    g->m_volumeType = (RefGridDataSource::EVolumeType) type;
    const size_t total = (size_t) N[0] * N[1] * N[2], bytes = total * (type == 1 ? sizeof(float) : 1);
    g->m_data = new uint8_t[bytes];
    memcpy(g->m_data, data, bytes);
    /* configure(), gridvolume.cpp:188-195, with an identity toWorld */
    AABB m_dataAABB(Point(bmin[0], bmin[1], bmin[2]), Point(bmax[0], bmax[1], bmax[2]));
    Transform m_worldToVolume;
    Vector extents(m_dataAABB.getExtents());
    g->m_worldToGrid = Transform::scale(Vector(
            (g->m_res[0] - 1) / extents[0],
            (g->m_res[1] - 1) / extents[1],
            (g->m_res[2] - 1) / extents[2])
        ) * Transform::translate(-Vector(m_dataAABB.min)) * m_worldToVolume;
    for (int i=0; i<255; i++) g->m_densityMap[i] = i/255.0f; /* :210 */
    g->m_densityMap[255] = 1.0f;                             /* :214 */
    return g;
}
void ref_grid_free(void *h) {
    RefGridDataSource *g = (RefGridDataSource *) h;
    delete[] g->m_data;
    delete g;
}
void ref_grid_lookup(void *h, size_t n, const float *p, float *out) {
    const RefGridDataSource *g = (const RefGridDataSource *) h;
    for (size_t i = 0; i < n; i++) out[i] = g->lookupFloat(Point(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
}
}
