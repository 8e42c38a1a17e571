/*
 * oracle/ref_film.cpp - TEST INFRASTRUCTURE ONLY (never linked into the product).  Part of oracle/_ref/libmer_reftrace.so.
 *
 * SURVEY a23, compiled VERBATIM from where it lies under /root/reference (nothing is copied into this repo):
 *     src/libcore/rfilter.cpp              ReconstructionFilter::configure (:37-55: the discretised, normalised filter table)
 *     include/mitsuba/core/rfilter.h       MTS_FILTER_RESOLUTION (:28), evalDiscretized (:75-77)
 *     src/rfilters/gaussian.cpp            GaussianFilter::eval (:52-57)
 *     src/rfilters/box.cpp                 BoxFilter::eval (:46-48)
 *     include/mitsuba/render/imageblock.h  ImageBlock::put(const Point2 &, const Float *) (:144-206) and the Spectrum overload
 *                                          (:124-131)
 * and SURVEY a22 (the pinhole sensor):
 *     src/sensors/perspective.cpp          PerspectiveCameraImpl::configure, the lines that build m_cameraToSample and invert
 *                                          it (:128-156), and sampleRay (:247-269)
 *     src/libcore/transform.cpp            Transform::perspective (:99-123), Transform::lookAt (:191-214) [operator*, translate,
 *                                          scale come with ref_volume.cpp]; include/mitsuba/core/matrix.inl's 4x4 inversion as it is
 * The bodies are cut out of those files by oracle/Makefile (awk, by signature) into film_*_extract.inc and camera_*_extract.inc in a
 * temporary directory (deleted after the build) and included into structs that declare exactly the data members they use.
 */
#include <mitsuba/mitsuba.h>
namespace mitsuba { using std::endl; }
#include <mitsuba/core/spectrum.h> /* reference */
#include <mitsuba/core/transform.h> /* reference (with matrix.h / matrix.inl / ray.h) */
#include "film_resolution_extract.inc" /* generated: #define MTS_FILTER_RESOLUTION 31 */

namespace mitsuba {

struct ReconstructionFilter { /* include/mitsuba/core/rfilter.h reduced to what configure() / evalDiscretized() / put() use */
    Float m_radius, m_scaleFactor;
    Float m_values[MTS_FILTER_RESOLUTION + 1];
    int m_borderSize;
    virtual ~ReconstructionFilter() {}
    virtual Float eval(Float x) const = 0;
    inline Float getRadius() const { return m_radius; }
    inline int getBorderSize() const { return m_borderSize; }
#include "film_discretized_extract.inc" /* generated: evalDiscretized */
    void configure();
};
#include "film_configure_extract.inc" /* generated: void ReconstructionFilter::configure() */

struct GaussianFilter : public ReconstructionFilter {
    Float m_stddev;
#include "film_gaussian_extract.inc" /* generated: eval */
};
struct BoxFilter : public ReconstructionFilter {
#include "film_box_extract.inc" /* generated: eval */
};

struct RefBitmap { /* the three accessors put() calls */
    int channels;
    Vector2i size;
    std::vector<Float> data;
    int getChannelCount() const { return channels; }
    const Vector2i &getSize() const { return size; }
    Float *getFloatData() { return data.data(); }
};

struct RefImageBlock { /* include/mitsuba/render/imageblock.h reduced to the data members put() uses */
    RefBitmap *m_bitmap;
    const ReconstructionFilter *m_filter;
    Point2i m_offset;
    int m_borderSize;
    Float *m_weightsX, *m_weightsY;
    bool m_warn;
#include "film_put_extract.inc" /* generated: both put() overloads */
};

/* VolumetricPathTracer (src/integrators/path/volpath.cpp) reduced to its MIS weight */
struct RefVolPath {
#include "volpath_miweight_extract.inc" /* generated: miWeight (volpath.cpp:430-433), the power heuristic of :137-141 and :164-173 */
};

#include "camera_util_extract.inc"      /* generated: degToRad (include/mitsuba/core/util.h:293) */
#include "camera_transform_extract.inc" /* generated: Transform::perspective, Transform::lookAt */

/* PerspectiveCameraImpl (src/sensors/perspective.cpp) reduced to what configure()'s projection lines and sampleRay() use */
struct RefFilmSize {
    Vector2i size;
    Point2i offset;
    const Vector2i &getSize() const { return size; }
    const Vector2i &getCropSize() const { return size; }
    const Point2i &getCropOffset() const { return offset; }
};
struct RefAnimatedTransform {
    Transform t;
    const Transform &eval(Float) const { return t; }
};
struct RefPerspectiveCamera {
    RefFilmSize *m_film;
    RefAnimatedTransform *m_worldTransform;
    Transform m_cameraToSample, m_sampleToCamera;
    Float m_aspect, m_xfov, m_nearClip, m_farClip;
    Vector2 m_invResolution;
    Float sampleTime(Float) const { return 0.0f; }
    void configureProjection() {
#include "camera_configure_extract.inc" /* generated: perspective.cpp:128-156 */
    }
#include "camera_sampleray_extract.inc" /* generated: sampleRay */
};

}

using namespace mitsuba;

static ReconstructionFilter *make_filter(int type) {
    ReconstructionFilter *f;
    if (type == 1) {
        GaussianFilter *g = new GaussianFilter();
        g->m_stddev = 0.5f;            /* gaussian.cpp:35 default */
        g->m_radius = 4 * g->m_stddev; /* :38 */
        f = g;
    } else {
        f = new BoxFilter();
        f->m_radius = 0.5f + 1e-5f; /* box.cpp:38 default */
    }
    f->configure();
    return f;
}

extern "C" {

/* volpath's miWeight(pdfA, pdfB) for n pairs */
void ref_mi_weight(size_t n, const float *pdfA, const float *pdfB, float *out) {
    RefVolPath v;
    for (size_t i = 0; i < n; i++) out[i] = v.miWeight(pdfA[i], pdfB[i]);
}

/* type: 0 box, 1 gaussian -> the 32 table values, radius, scale factor, border size */
void ref_filter_table(int type, float *values32, float *radius, float *scaleFactor, int *borderSize) {
    ReconstructionFilter *f = make_filter(type);
    for (int i = 0; i <= MTS_FILTER_RESOLUTION; i++) values32[i] = f->m_values[i];
    *radius = f->m_radius;
    *scaleFactor = f->m_scaleFactor;
    *borderSize = f->m_borderSize;
    delete f;
}

/* n samples (pos [n][2] in fractional pixel coordinates, values [n][channels]) put into ONE block that covers the whole W x H
 * film (offset 0, border = the filter's), in order; out: the block without its border, [H][W][channels]; ok[i] = put()'s result */
void ref_film_put(int type, int W, int H, int channels, size_t n, const float *pos, const float *values, float *out, int *ok) {
    ReconstructionFilter *f = make_filter(type);
    RefBitmap bmp;
    const int b = f->getBorderSize();
    bmp.channels = channels;
    bmp.size = Vector2i(W + 2 * b, H + 2 * b); /* imageblock.cpp: ImageBlock::ImageBlock allocates size + 2 * border */
    bmp.data.assign((size_t) bmp.size.x * bmp.size.y * channels, 0.0f);
    RefImageBlock blk;
    blk.m_bitmap = &bmp;
    blk.m_filter = f;
    blk.m_offset = Point2i(0, 0);
    blk.m_borderSize = b;
    const int wsize = (int) std::ceil(2 * f->getRadius()) + 1;
    std::vector<Float> wx(wsize), wy(wsize);
    blk.m_weightsX = wx.data();
    blk.m_weightsY = wy.data();
    blk.m_warn = true;
    for (size_t i = 0; i < n; i++) ok[i] = blk.put(Point2(pos[2 * i], pos[2 * i + 1]), values + (size_t) channels * i) ? 1 : 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            for (int k = 0; k < channels; k++)
                out[((size_t) y * W + x) * channels + k] = bmp.data[((size_t) (y + b) * bmp.size.x + (x + b)) * channels + k];
    delete f;
}


/* PerspectiveCamera::sampleRay for n pixel samples: <lookat origin target up>, fov along x (sensor.cpp:244-262, fovAxis "x"),
 * near / far clip 1e-2 / 1e4 (sensor.cpp defaults), uncropped W x H film */
void ref_camera_rays(const float *origin, const float *target, const float *up, float fov, int W, int H, size_t n, const float *samplePos,
                     float *o, float *d) {
    RefFilmSize film;
    film.size = Vector2i(W, H);
    film.offset = Point2i(0, 0);
    RefAnimatedTransform world;
    world.t = Transform::lookAt(Point(origin[0], origin[1], origin[2]), Point(target[0], target[1], target[2]), Vector(up[0], up[1], up[2]));
    RefPerspectiveCamera cam;
    cam.m_film = &film;
    cam.m_worldTransform = &world;
    cam.m_aspect = film.size.x / (Float) film.size.y;                      /* sensor.cpp ProjectiveCamera::configure */
    cam.m_xfov = fov;
    cam.m_nearClip = 1e-2f;
    cam.m_farClip = 1e4f;
    cam.m_invResolution = Vector2(1.0f / film.size.x, 1.0f / film.size.y);  /* sensor.cpp Sensor::configure */
    cam.configureProjection();
    for (size_t i = 0; i < n; i++) {
        Ray ray;
        cam.sampleRay(ray, Point2(samplePos[2 * i], samplePos[2 * i + 1]), Point2(0.5f, 0.5f), 0.5f);
        o[3 * i] = ray.o.x; o[3 * i + 1] = ray.o.y; o[3 * i + 2] = ray.o.z;
        d[3 * i] = ray.d.x; d[3 * i + 1] = ray.d.y; d[3 * i + 2] = ray.d.z;
    }
}

}
