/*
 * integration/mitsuba_plugins.cpp — the Mitsuba-side binding of libmitsubaer_b200.so.
 *
 * This is the stub a MitsubaER maintainer adds under src/ (one translation unit per plugin tag in
 * practice; shown together here).  It compiles ONLY inside a Mitsuba 0.5 tree (needs
 * <mitsuba/render/*.h>, Boost, ...), which this repository's image does not have, so it is not
 * built or tested here; the C ABI it calls is exercised by tests/ through ctypes instead.
 *
 * Plugin discovery is unchanged: `<medium type="heterogeneousrefractive">` dlopen()s
 * plugins/heterogeneousrefractive.so and calls CreateInstance (include/mitsuba/core/cobject.h:99-107,
 * src/libcore/plugin.cpp:71-96).  The classes below keep the reference's property and child names
 *  so existing scene XML parses unchanged, and turn a non-zero mer_status into
 * Log(EError, ...) (which throws, src/libcore/logger.cpp:100-147).
 */
#include <mitsuba/render/scene.h>
#include <mitsuba/render/volume.h>
#include <mitsuba/render/medium.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/plugin.h>

#include "mitsubaer_b200.h"

MTS_NAMESPACE_BEGIN

#define MER_CHECK(call) do { if ((call) != MER_OK) Log(EError, "mitsubaer_b200: %s", mer_last_error()); } while (0)

static void fillVolumeDesc(mer_volume_desc &d, const Transform &volumeToWorld, const AABB *aabbOverride) {
    memset(&d, 0, sizeof(d));
    Matrix4x4 inv = volumeToWorld.getInverseMatrix();
    d.has_transform = volumeToWorld.getMatrix().isIdentity() ? 0 : 1;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
            d.world_to_volume[4 * r + c] = (float) inv(r, c);
    if (aabbOverride) { /* `min` / `max` properties, splinevolume.cpp:93-98 */
        for (int i = 0; i < 3; ++i) { d.bbox_min[i] = aabbOverride->min[i]; d.bbox_max[i] = aabbOverride->max[i]; }
    }
}

/* ------------------------------------------------------------------ <volume type="splinevolume"> */
class B200SplineDataSource : public VolumeDataSource {
public:
    B200SplineDataSource(const Properties &props) : VolumeDataSource(props), m_handle(NULL) {
        m_volumeToWorld = props.getTransform("toWorld", Transform());
        AABB ov; bool hasOv = props.hasProperty("min") && props.hasProperty("max");
        if (hasOv) { ov.min = props.getPoint("min"); ov.max = props.getPoint("max"); }
        mer_volume_desc d;
        fillVolumeDesc(d, m_volumeToWorld, hasOv ? &ov : NULL);
        fs::path resolved = Thread::getThread()->getFileResolver()->resolve(props.getString("filename"));
        int mode = props.getString("fetch", "tricubic") == "trilinear_packed" ? MER_RIF_TRILINEAR_PACKED : MER_RIF_TRICUBIC;
        MER_CHECK(mer_rif_create_from_file(props.getInteger("device", 0), resolved.string().c_str(), &d, mode, &m_handle));
        mer_volume_desc out; int m;
        MER_CHECK(mer_rif_desc(m_handle, &out, &m));
        m_aabb = AABB(Point(out.bbox_min[0], out.bbox_min[1], out.bbox_min[2]), Point(out.bbox_max[0], out.bbox_max[1], out.bbox_max[2]));
    }
    virtual ~B200SplineDataSource() { mer_rif_destroy(m_handle); }

    /* scalar virtuals: one-element batches (correct, slow; the integrator never calls them) */
    FLOAT value(const PointF &p) const { float q[3] = {(float) p.x, (float) p.y, (float) p.z}, f;
        MER_CHECK(mer_rif_eval_batch(m_handle, MER_EVAL_VALUE, 1, q, &f, NULL)); return f; }
    VectorF gradient(const PointF &p) const { float q[3] = {(float) p.x, (float) p.y, (float) p.z}, g[3];
        MER_CHECK(mer_rif_eval_batch(m_handle, MER_EVAL_GRADIENT, 1, q, NULL, g)); return VectorF(g[0], g[1], g[2]); }
    void valueAndGradient(const PointF &p, FLOAT &f, VectorF &v) const { float q[3] = {(float) p.x, (float) p.y, (float) p.z}, fv, g[3];
        MER_CHECK(mer_rif_eval_batch(m_handle, MER_EVAL_VALUE_AND_GRADIENT, 1, q, &fv, g)); f = fv; v = VectorF(g[0], g[1], g[2]); }
    bool insideVolumeLimits(const PointF &p) const { float q[3] = {(float) p.x, (float) p.y, (float) p.z}; uint8_t in;
        MER_CHECK(mer_rif_inside_limits_batch(m_handle, 1, q, &in)); return in != 0; }
    bool supportsFloatLookups() const { return true; }
    Float getStepSize() const { return m_stepSize; }
    Float getMaximumFloatValue() const { return 1.0f; }
    mer_rif *handle() const { return m_handle; }
    MTS_DECLARE_CLASS()
private:
    Transform m_volumeToWorld; Float m_stepSize; mer_rif *m_handle;
};

/* ------------------------------------------------------------------ <medium type="heterogeneousrefractive"> */
class B200HeterogeneousRefractiveMedium : public Medium {
public:
    B200HeterogeneousRefractiveMedium(const Properties &props) : Medium(props), m_props(props), m_handle(NULL) { }
    virtual ~B200HeterogeneousRefractiveMedium() { mer_medium_destroy(m_handle); }

    void addChild(const std::string &name, ConfigurableObject *child) {
        if (child->getClass()->derivesFrom(MTS_CLASS(VolumeDataSource)) && name == "rif") m_rif = static_cast<B200SplineDataSource *>(child);
        else if (child->getClass()->derivesFrom(MTS_CLASS(VolumeDataSource)) && name == "density") m_density = static_cast<VolumeDataSource *>(child);
        else Medium::addChild(name, child); /* the phase function */
    }

    void configure() {
        Medium::configure();
        if (m_rif.get() == NULL) Log(EError, "No RIF specified!");
        mer_medium_desc d; memset(&d, 0, sizeof(d));
        for (int i = 0; i < 3; ++i) { d.sigma_a[i] = m_sigmaA[i]; d.sigma_s[i] = m_sigmaS[i]; }
        d.stepsize = m_props.getFloat("stepsize", 1e-3f);
        d.medium_sampling_weight = m_props.getFloat("mediumSamplingWeight", -1);
        std::string s = m_props.getString("strategy", "balance");
        d.strategy = s == "balance" ? MER_STRATEGY_BALANCE : s == "single" ? MER_STRATEGY_SINGLE : s == "manual" ? MER_STRATEGY_MANUAL : MER_STRATEGY_MAXIMUM;
        d.channel = m_props.getInteger("channel", -1);
        d.sampling_density = m_props.getFloat("samplingDensity", 0.0f);
        /* containment predicate: the interior shape's AABB (hackForBox form) or bounding sphere (hackForSphere form) */
        AABB box = m_shape->getAABB();
        d.shape_type = MER_SHAPE_BOX;
        for (int i = 0; i < 3; ++i) { d.shape[i] = box.min[i]; d.shape[3 + i] = box.max[i]; }
        /* the container's surface: hdielectric (eta from this medium's RIF) or anything index-matched */
        d.boundary = (m_shape->getBSDF() && m_shape->getBSDF()->getClass()->getName() == "HSmoothDielectric")
                         ? MER_BOUNDARY_HDIELECTRIC : MER_BOUNDARY_INDEX_MATCHED;
        d.hg_g = m_phaseFunction->getMeanCosine();
        d.density_scale = m_props.getFloat("scale", 1.0f);
        Spectrum albedo = m_props.getSpectrum("albedo", Spectrum(0.0f));
        for (int i = 0; i < 3; ++i) d.albedo[i] = albedo[i];
        MER_CHECK(mer_medium_create(&d, m_rif->handle(), NULL /* or the density grid handle */, &m_handle));
    }

    /* Medium::sampleDistance (include/mitsuba/render/medium.h:130-131): a one-ray batch that replays the
     * sampler draws the reference consumes (:404, :440) */
    bool sampleDistance(const Ray &ray, MediumSamplingRecord &mRec, Sampler *sampler) const {
        float o[3] = {ray.o.x, ray.o.y, ray.o.z}, dd[3] = {ray.d.x, ray.d.y, ray.d.z}, mint = ray.mint;
        float xi[2] = {sampler->next1D(), sampler->next1D()};
        uint8_t ok; float t, p[3], dv[3], opl, rr, T[3], ps, pf, ss[3]; int32_t ns;
        mer_medium_sampling_records r = {&ok, &t, p, dv, &opl, &rr, T, &ps, &pf, ss, &ns};
        MER_CHECK(mer_medium_sample_distance_batch(m_handle, 1, o, dd, &mint, xi, &r));
        mRec.t = t; mRec.p = Point(p[0], p[1], p[2]); mRec.d = Vector(dv[0], dv[1], dv[2]);
        mRec.opticalLength = opl; mRec.refRatioSq = rr; mRec.pdfSuccess = mRec.pdfSuccessRev = ps; mRec.pdfFailure = pf;
        for (int i = 0; i < 3; ++i) { mRec.transmittance[i] = T[i]; mRec.sigmaS[i] = ss[i]; mRec.sigmaA[i] = m_sigmaA[i]; }
        mRec.time = ray.time; mRec.medium = this;
        return ok != 0;
    }
    Spectrum evalTransmittance(const Ray &ray, Sampler *) const {
        float a = ray.mint, b = ray.maxt, T[3];
        MER_CHECK(mer_medium_eval_transmittance_batch(m_handle, 1, &a, &b, T));
        Spectrum s; for (int i = 0; i < 3; ++i) s[i] = T[i]; return s;
    }
    bool isHomogeneous() const { return false; }
    bool isheterogeneousrefractive() const { return true; }
    mer_medium *handle() const { return m_handle; }
    MTS_DECLARE_CLASS()
private:
    Properties m_props; ref<B200SplineDataSource> m_rif; ref<VolumeDataSource> m_density; mer_medium *m_handle;
};

/* ------------------------------------------------------------------ <integrator type="ervolpath">
 * The real entry point: Integrator::render (include/mitsuba/render/integrator.h:61-96) pulls the sensor,
 * film, filter, emitter and medium parameters out of the Scene, renders on the GPU(s) and hands the
 * [R,G,B,alpha,weight] film back with Film::setBitmap (src/films/hdrfilm.cpp:412-414). */
class EikonalVolPathIntegrator : public Integrator {
public:
    EikonalVolPathIntegrator(const Properties &props) : Integrator(props) {
        m_maxDepth = props.getInteger("maxDepth", -1);
        m_rrDepth = props.getInteger("rrDepth", 5);
        m_directConnections = props.getBoolean("directConnections", false); /* curved next-event estimation */
        m_lightTracing = props.getBoolean("lightTracing", false);           /* emitter-side walk + sensor connections */
        if (m_maxDepth == 0 || m_maxDepth < -1)
            Log(EError, "maxDepth must be set to -1 (infinite) or a value greater than zero!");
    }
    bool render(Scene *scene, RenderQueue *, const RenderJob *, int, int, int) {
        const Sensor *sensor = scene->getSensor();
        Film *film = const_cast<Film *>(sensor->getFilm());
        Vector2i size = film->getCropSize();
        mer_render_desc r; memset(&r, 0, sizeof(r));
        r.width = size.x; r.height = size.y;
        r.spp_total = (int) scene->getSampler()->getSampleCount();
        r.sample_begin = 0; r.sample_stride = 1;
        r.seed = 20201201;
        const Transform &toWorld = sensor->getWorldTransform()->eval(0);
        Point o = toWorld(Point(0.0f)); Vector dir = toWorld(Vector(0, 0, 1)), up = toWorld(Vector(0, 1, 0));
        for (int i = 0; i < 3; ++i) { r.cam_origin[i] = o[i]; r.cam_target[i] = o[i] + dir[i]; r.cam_up[i] = up[i]; }
        r.fov_deg = static_cast<const PerspectiveCamera *>(sensor)->getXFov();
        r.filter = film->getReconstructionFilter()->getRadius() > 1 ? MER_FILTER_GAUSSIAN : MER_FILTER_BOX;
        r.max_depth = m_maxDepth; r.rr_depth = m_rrDepth;
        r.direct_connections = m_directConnections ? 1 : 0;
        r.light_tracing = m_lightTracing ? 1 : 0;
        /* the fork's transient film (src/librender/film.cpp:56-78): frames, bounds, calibration */
        if (film->getDecompositionType() == Film::ETransient && film->getFrames() > 1) {
            r.frames = (int32_t) film->getFrames();
            r.min_bound = film->getDecompositionMinBound();
            r.bin_width = film->getDecompositionBinWidth();
            r.calibrated_transient = film->isCalibratedTransient() ? 1 : 0;
        }
        /* emitters: the rectangle area light (quad_*) and the collimated beam (beam_*) are read off scene->getEmitters() the
         * same way; left out of this stub for brevity */
        if (scene->hasEnvironmentEmitter()) {
            Spectrum L = scene->getEnvironmentEmitter()->evalEnvironment(RayDifferential(Point(0.0f), Vector(0, 0, 1), 0));
            for (int i = 0; i < 3; ++i) r.env_radiance[i] = L[i];
        }
        const B200HeterogeneousRefractiveMedium *medium = NULL;
        for (size_t i = 0; i < scene->getMedia().size(); ++i)
            if (scene->getMedia()[i]->isheterogeneousrefractive())
                medium = static_cast<const B200HeterogeneousRefractiveMedium *>(scene->getMedia()[i].get());
        if (!medium) Log(EError, "ervolpath needs a heterogeneousrefractive medium");
        ref<Bitmap> bitmap = r.frames > 1 ? new Bitmap(Bitmap::EMultiSpectrumAlphaWeight, Bitmap::EFloat32, size, 3 * r.frames + 2)
                                          : new Bitmap(Bitmap::ESpectrumAlphaWeight, Bitmap::EFloat32, size);
        mer_render_stats stats;
        MER_CHECK(mer_render(medium->handle(), &r, bitmap->getFloat32Data(), &stats));
        Log(EInfo, "ervolpath: %llu samples, %llu eikonal steps, %.1f ms on the GPU",
            (unsigned long long) stats.samples, (unsigned long long) stats.ray_steps, stats.device_ms);
        film->setBitmap(bitmap);
        return true;
    }
    void cancel() { }
    MTS_DECLARE_CLASS()
private:
    int m_maxDepth, m_rrDepth;
    bool m_directConnections, m_lightTracing;
};

MTS_IMPLEMENT_CLASS(B200SplineDataSource, false, VolumeDataSource)
MTS_IMPLEMENT_CLASS(B200HeterogeneousRefractiveMedium, false, Medium)
MTS_IMPLEMENT_CLASS(EikonalVolPathIntegrator, false, Integrator)
/* one MTS_EXPORT_PLUGIN per .so in a real tree: splinevolume.so, heterogeneousrefractive.so, ervolpath.so */
MTS_EXPORT_PLUGIN(EikonalVolPathIntegrator, "Eikonal volumetric path tracer (B200)");
MTS_NAMESPACE_END
