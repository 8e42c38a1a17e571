/*
 * oracle/ref_volume.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).  Part of oracle/_ref/libmer_reftrace.so.
 *
 * SplineDataSource's lookups (src/volume/splinevolume.cpp:319-377: insideVolumeLimits, maxSDFError, value, gradient, hessian,
 * valueAndGradient, gradientAndHessian, valueGradientAndHessian), cut out of the .cpp by oracle/Makefile into
 * splinevolume_extract.inc (a temporary directory, deleted after the build) and compiled VERBATIM inside a struct that declares the data members they use, on top
 * of the reference's own basisspline.h / transform.h / matrix.h / aabb.h; AABB::getCorner comes from src/libcore/aabb.cpp
 * the same way, and SplineDataSource::loadFromFile (:204-317, SURVEY a6: the .vol v3 header, the float -> FLOAT copy, the
 * prefilter) with enum EVolumeType, behind stand-ins for the framework plumbing it names (fs::path, the file resolver, the memory
 * map, MemoryStream).  GridDataSource::lookupFloat (src/volume/gridvolume.cpp:337-388) likewise, with enum EVolumeType
 * (:101-106) and the three Transform functions its configure() uses (src/libcore/transform.cpp:28-65: operator*, translate,
 * scale).  And the straight-ray medium over such a grid: HeterogeneousMedium::sampleDistance / evalTransmittance
 * (src/medium/heterogeneous.cpp:546-672, Woodcock tracking and Simpson quadrature) with integrateDensity, invertDensityIntegral,
 * lookupDensity and enum EIntegrationMethod.  Nothing is copied into this repo.
 */
#include "ref_volume.h"
#include "ref_records.h"
#include <mitsuba/core/ray.h>       /* reference */
#include <mitsuba/render/sampler.h> /* oracle/shim_phase */
#include <mitsuba/core/basisspline.h>
namespace mitsuba { extern bool solveQuadratic(Float a, Float b, Float c, Float &x0, Float &x1); } /* util.h; bsphere.h mentions it */
#include <mitsuba/core/aabb.h>

namespace mitsuba {
#include "aabb_extract.inc" /* generated: AABB::getCorner from src/libcore/aabb.cpp */

/* ---- what loadFromFile() names of the framework, reduced to what it does with it (TEST INFRASTRUCTURE) */
namespace fs {
struct path {
    std::string s;
    path() {}
    path(const std::string &x) : s(x) {}
    path(const char *x) : s(x) {}
    path filename() const { size_t k = s.find_last_of('/'); return path(k == std::string::npos ? s : s.substr(k + 1)); }
    const std::string &string() const { return s; }
};
}
template <typename T> struct ref {
    T *p;
    ref(T *q = NULL) : p(q) {}
    T *operator->() const { return p; }
    T *get() const { return p; }
};
struct FileResolver { fs::path resolve(const fs::path &p) const { return p; } };
struct Thread {
    static Thread *getThread() { static Thread t; return &t; }
    FileResolver *getFileResolver() { static FileResolver r; return &r; }
};
struct MemoryMappedFile { /* the whole file in memory */
    std::vector<char> buf;
    explicit MemoryMappedFile(const fs::path &p) {
        FILE *f = fopen(p.string().c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + p.string());
        fseek(f, 0, SEEK_END);
        buf.resize((size_t) ftell(f));
        fseek(f, 0, SEEK_SET);
        if (fread(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); throw std::runtime_error("short read"); }
        fclose(f);
    }
    void *getData() { return buf.data(); }
    size_t getSize() const { return buf.size(); }
};
struct MemoryStream { /* little-endian reads from a memory block (this host is little-endian) */
    const char *d;
    size_t n, pos;
    MemoryStream(void *data, size_t size) : d((const char *) data), n(size), pos(0) {}
    void setByteOrder(Stream::EByteOrder) {}
    void read(void *dst, size_t k) { if (pos + k > n) throw std::runtime_error("read past the end"); memcpy(dst, d + pos, k); pos += k; }
    int readInt() { int32_t v; read(&v, 4); return v; }
    float readSingle() { float v; read(&v, 4); return v; }
};
inline std::string memString(size_t) { return std::string(); }

struct RefSplineDataSource : public RefVolume {
#include "splinevolume_loader_extract.inc" /* generated: enum EVolumeType, loadFromFile */
    fs::path m_filename;
    ref<MemoryMappedFile> m_mmap;
    Vector3i m_res;
    int m_channels;
    EVolumeType m_volumeType;
    AABB m_dataAABB; /* default-constructed = invalid: the bounding box comes from the file */
    uint8_t *m_data;
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
#include "gridvolume_extract.inc" /* generated: enum EVolumeType, struct float3, lookupFloat, lookupSpectrum */
    Transform m_worldToGrid;
    Vector3i m_res;
    EVolumeType m_volumeType;
    uint8_t *m_data;
    Float m_densityMap[256];
};

/* what HeterogeneousMedium asks of its albedo / orientation children and of its phase function */
struct RefConstantSpectrumVolume {
    Spectrum value;
    Spectrum lookupSpectrum(const Point &) const { return value; }
    Vector lookupVector(const Point &) const { return Vector(0.0f); }
};
struct RefPhaseSigmaDir { Float sigmaDir(Float) const { return 1.0f; } };

/* HeterogeneousMedium (src/medium/heterogeneous.cpp) reduced to the data members its distance sampling uses */
struct RefHeterogeneousMedium {
#include "heterogeneous_extract.inc" /* generated: enum EIntegrationMethod, integrateDensity, invertDensityIntegral, evalTransmittance,
                                        sampleDistance, lookupDensity */
    EIntegrationMethod m_method;
    RefGridDataSource *m_density;
    RefConstantSpectrumVolume *m_albedo, *m_orientation;
    RefPhaseSigmaDir *m_phaseFunction;
    Float m_scale;
    bool m_anisotropicMedium;
    Float m_stepSize;
    AABB m_densityAABB;
    Float m_maxDensity, m_invMaxDensity;
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

/* SplineDataSource from a .vol file through the reference's own loadFromFile() */
RefVolume *ref_load_volume(const char *path) {
    RefSplineDataSource *rif = new RefSplineDataSource();
    rif->m_worldToVolume_Rot.setIdentity();
    rif->m_worldToVolume_RotT.setIdentity();
    rif->loadFromFile(fs::path(path));
    return rif;
}

/* ---- C entry points of the density grid */
extern "C" {
/* type: 1 = float32 payload, 3 = uint8 payload (EVolumeType); data [z][y][x][channels] */
void *ref_grid_create_channels(const void *data, const int *N, const float *bmin, const float *bmax, int type, int channels) {
    RefGridDataSource *g = new RefGridDataSource();
    g->m_res = Vector3i(N[0], N[1], N[2]);
    g->m_volumeType = (RefGridDataSource::EVolumeType) type;
    const size_t total = (size_t) N[0] * N[1] * N[2] * channels, bytes = total * (type == 1 ? sizeof(float) : 1);
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
void *ref_grid_create(const void *data, const int *N, const float *bmin, const float *bmax, int type) {
    return ref_grid_create_channels(data, N, bmin, bmax, type, 1);
}
void ref_grid_lookup_spectrum(void *h, size_t n, const float *p, float *out) {
    const RefGridDataSource *g = (const RefGridDataSource *) h;
    for (size_t i = 0; i < n; i++) {
        Spectrum s = g->lookupSpectrum(Point(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
        out[3 * i] = s[0]; out[3 * i + 1] = s[1]; out[3 * i + 2] = s[2];
    }
}
void ref_grid_free(void *h) {
    RefGridDataSource *g = (RefGridDataSource *) h;
    delete[] g->m_data;
    delete g;
}
/* HeterogeneousMedium over the grid: method 0 Simpson quadrature, 1 Woodcock tracking; configure() :221-262 resolved by the
 * caller: max_density = scale * (maximum of the grid), step_size for the quadrature */
namespace {
struct ReplaySampler : public mitsuba::Sampler {
    const float *xi;
    size_t k, n;
    mitsuba::Float next1D() { return k < n ? xi[k++] : 0.5f; }
    mitsuba::Point2 next2D() { mitsuba::Float a = next1D(), b = next1D(); return mitsuba::Point2(a, b); }
};
}
void *ref_hetmedium_create(void *grid, const float *bmin, const float *bmax, float scale, float maxDensity, int method, float stepSize, const float *albedo) {
    RefHeterogeneousMedium *m = new RefHeterogeneousMedium();
    m->m_density = (RefGridDataSource *) grid;
    m->m_albedo = new RefConstantSpectrumVolume();
    for (int c = 0; c < 3; c++) m->m_albedo->value[c] = albedo[c];
    m->m_orientation = NULL;
    m->m_phaseFunction = NULL;
    m->m_anisotropicMedium = false;
    m->m_scale = scale;
    m->m_method = (RefHeterogeneousMedium::EIntegrationMethod) method;
    m->m_stepSize = stepSize;
    m->m_densityAABB = AABB(Point(bmin[0], bmin[1], bmin[2]), Point(bmax[0], bmax[1], bmax[2]));
    m->m_maxDensity = maxDensity;          /* :243: m_scale * m_density->getMaximumFloatValue() */
    m->m_invMaxDensity = 1.0f / maxDensity; /* :244 */
    return m;
}
void ref_hetmedium_free(void *h) {
    RefHeterogeneousMedium *m = (RefHeterogeneousMedium *) h;
    delete m->m_albedo;
    delete m;
}
/* sampleDistance over n rays; xi: nxi numbers per ray that sampler->next1D() returns in order */
void ref_hetmedium_sample_distance(void *h, size_t n, const float *ro, const float *rd, const float *mint, const float *maxt, const float *xi, size_t nxi,
                                   int *success, float *t, float *sigmaS, float *transmittance) {
    const RefHeterogeneousMedium *m = (const RefHeterogeneousMedium *) h;
    for (size_t i = 0; i < n; i++) {
        Ray ray(Point(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), Vector(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), 0.0f);
        ray.mint = mint[i];
        ray.maxt = maxt[i];
        MediumSamplingRecord mRec;
        mRec.t = 0;
        ReplaySampler s;
        s.xi = xi + nxi * i; s.k = 0; s.n = nxi;
        success[i] = m->sampleDistance(ray, mRec, &s) ? 1 : 0;
        t[i] = mRec.t;
        for (int c = 0; c < 3; c++) { sigmaS[3 * i + c] = success[i] ? mRec.sigmaS[c] : 0.0f; transmittance[3 * i + c] = mRec.transmittance[c]; }
    }
}
void ref_hetmedium_eval_transmittance(void *h, size_t n, const float *ro, const float *rd, const float *mint, const float *maxt, const float *xi, size_t nxi,
                                      float *out) {
    const RefHeterogeneousMedium *m = (const RefHeterogeneousMedium *) h;
    for (size_t i = 0; i < n; i++) {
        Ray ray(Point(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), Vector(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), 0.0f);
        ray.mint = mint[i];
        ray.maxt = maxt[i];
        ReplaySampler s;
        s.xi = xi + nxi * i; s.k = 0; s.n = nxi;
        out[i] = m->evalTransmittance(ray, &s)[0];
    }
}
void ref_grid_lookup(void *h, size_t n, const float *p, float *out) {
    const RefGridDataSource *g = (const RefGridDataSource *) h;
    for (size_t i = 0; i < n; i++) out[i] = g->lookupFloat(Point(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
}
}
}
