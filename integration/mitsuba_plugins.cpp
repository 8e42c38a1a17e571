/*
 * integration/mitsuba_plugins.cpp — the Mitsuba-side binding of libmitsubaer_b200.so.
 *
 * This is the file a MitsubaER maintainer adds under src/ and builds once per plugin tag:
 *
 *     -DMER_PLUGIN_SPLINEVOLUME            -> plugins/splinevolume.so            (<volume type="splinevolume">)
 *     -DMER_PLUGIN_GRIDVOLUME              -> plugins/gridvolume.so              (<volume type="gridvolume">, density child)
 *     -DMER_PLUGIN_HG                      -> plugins/hg.so                      (<phase type="hg">)
 *     -DMER_PLUGIN_HETEROGENEOUSREFRACTIVE -> plugins/heterogeneousrefractive.so (<medium type="heterogeneousrefractive">)
 *     -DMER_PLUGIN_ERVOLPATH               -> plugins/ervolpath.so               (<integrator type="ervolpath">, new)
 *
 * (all classes are compiled into every plugin, one MTS_EXPORT_PLUGIN each: the integrator needs the layout of the others'
 * handles).  It needs a Mitsuba 0.5 tree (mitsuba/render/ headers, Boost, ...), which this repository's image does not have:
 * here it is SYNTAX-CHECKED against integration/mitsuba_stub (declarations of exactly the members used, signatures as in
 * the MitsubaER headers) by tests/test_host_cpp.py, and the C ABI it calls is exercised by tests/ through ctypes / C / C++.
 *
 * Plugin discovery is unchanged: `<medium type="heterogeneousrefractive">` dlopen()s plugins/heterogeneousrefractive.so and
 * calls CreateInstance (include/mitsuba/core/cobject.h:99-107, src/libcore/plugin.cpp:71-96).  The classes keep the
 * reference's property and child names, its (Stream*, InstanceManager*) constructors and
 * serialize() (class.h:219; heterogeneousrefractive.cpp:343-359, 384-391), and turn a non-zero mer_status into
 * Log(EError, ...) (which throws, src/libcore/logger.cpp:100-147).
 */
#include <mitsuba/render/scene.h>
#include <mitsuba/render/volume.h>
#include <mitsuba/render/medium.h>
#include <mitsuba/render/phase.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/integrator.h>
#include <mitsuba/core/fresolver.h>
#include <mitsuba/core/plugin.h>

#include "mitsubaer_b200.h"

MTS_NAMESPACE_BEGIN

#define MER_CHECK(call) do { if ((call) != MER_OK) Log(EError, "mitsubaer_b200: %s", mer_last_error()); } while (0)

static void fillVolumeDesc(mer_volume_desc &d, const Transform &volumeToWorld, const AABB *aabbOverride) {
    memset(&d, 0, sizeof(d));
    const Matrix4x4 &inv = volumeToWorld.getInverseMatrix();
    d.has_transform = volumeToWorld.getMatrix().isIdentity() ? 0 : 1;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
            d.world_to_volume[4 * r + c] = (float) inv(r, c);
    if (aabbOverride) { /* `min` / `max` properties, splinevolume.cpp:93-98 */
        for (int i = 0; i < 3; ++i) { d.bbox_min[i] = aabbOverride->min[i]; d.bbox_max[i] = aabbOverride->max[i]; }
    }
}

/* what both volume plugins keep to be able to re-create themselves on a network worker (splinevolume.cpp:113-135) */
struct VolumeSource {
    std::string filename;
    Transform volumeToWorld;
    bool hasOverride;
    AABB override_;
    int device;
    void read(const Properties &props) {
        volumeToWorld = props.getTransform("toWorld", Transform());
        hasOverride = props.hasProperty("min") && props.hasProperty("max");
        if (hasOverride) { override_.min = props.getPoint("min"); override_.max = props.getPoint("max"); }
        filename = Thread::getThread()->getFileResolver()->resolve(fs::path(props.getString("filename"))).string();
        device = props.getInteger("device", 0);
    }
    void read(Stream *stream) {
        filename = stream->readString();
        device = stream->readInt();
        hasOverride = stream->readBool();
        for (int i = 0; i < 3 && hasOverride; ++i) { override_.min[i] = stream->readFloat(); override_.max[i] = stream->readFloat(); }
    }
    void write(Stream *stream) const {
        stream->writeString(filename);
        stream->writeInt(device);
        stream->writeBool(hasOverride);
        for (int i = 0; i < 3 && hasOverride; ++i) { stream->writeFloat(override_.min[i]); stream->writeFloat(override_.max[i]); }
    }
};

/* ------------------------------------------------------------------ <volume type="splinevolume"> (src/volume/splinevolume.cpp) */
class B200SplineDataSource : public VolumeDataSource {
public:
    B200SplineDataSource(const Properties &props) : VolumeDataSource(props), m_handle(NULL) {
        m_src.read(props);
        m_mode = props.getString("fetch", "tricubic") == "trilinear_packed" ? MER_RIF_TRILINEAR_PACKED : MER_RIF_TRICUBIC;
        load();
    }
    B200SplineDataSource(Stream *stream, InstanceManager *manager) : VolumeDataSource(stream, manager), m_handle(NULL) {
        m_src.read(stream);
        m_mode = stream->readInt();
        load();
    }
    virtual ~B200SplineDataSource() { mer_rif_destroy(m_handle); }
    void serialize(Stream *stream, InstanceManager *manager) const {
        VolumeDataSource::serialize(stream, manager);
        m_src.write(stream);
        stream->writeInt(m_mode);
    }

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
    Float getStepSize() const { return m_stepSize; }            /* splinevolume.cpp:183-185: half the smallest pitch */
    Float getMaximumFloatValue() const { return 1.0f; }         /* :677-679 */
    mer_rif *handle() const { return m_handle; }
    MTS_DECLARE_CLASS()
private:
    void load() {
        mer_volume_desc d;
        fillVolumeDesc(d, m_src.volumeToWorld, m_src.hasOverride ? &m_src.override_ : NULL);
        MER_CHECK(mer_rif_create_from_file(m_src.device, m_src.filename.c_str(), &d, m_mode, &m_handle));
        mer_volume_desc out; int m;
        MER_CHECK(mer_rif_desc(m_handle, &out, &m));
        m_aabb = AABB(Point(out.bbox_min[0], out.bbox_min[1], out.bbox_min[2]), Point(out.bbox_max[0], out.bbox_max[1], out.bbox_max[2]));
        m_stepSize = INFINITY;
        for (int i = 0; i < 3; ++i) m_stepSize = std::min(m_stepSize, 0.5f * (out.bbox_max[i] - out.bbox_min[i]) / (Float) (out.res[i] - 1));
    }
    VolumeSource m_src; int m_mode; Float m_stepSize; mer_rif *m_handle;
};

/* ------------------------------------------------------------------ <volume type="gridvolume"> (src/volume/gridvolume.cpp): the
 * density child of the medium, trilinear lookups */
class B200GridDataSource : public VolumeDataSource {
public:
    B200GridDataSource(const Properties &props) : VolumeDataSource(props), m_handle(NULL) { m_src.read(props); load(); }
    B200GridDataSource(Stream *stream, InstanceManager *manager) : VolumeDataSource(stream, manager), m_handle(NULL) { m_src.read(stream); load(); }
    virtual ~B200GridDataSource() { mer_grid_destroy(m_handle); }
    void serialize(Stream *stream, InstanceManager *manager) const { VolumeDataSource::serialize(stream, manager); m_src.write(stream); }
    bool supportsFloatLookups() const { return mer_grid_channels(m_handle) == 1; }    /* gridvolume.cpp:578 */
    bool supportsSpectrumLookups() const { return mer_grid_channels(m_handle) == 3; } /* gridvolume.cpp:579 */
    Float lookupFloat(const Point &p) const { float q[3] = {p.x, p.y, p.z}, v; /* gridvolume.cpp:337-363 */
        MER_CHECK(mer_grid_lookup_batch(m_handle, 1, q, &v)); return v; }
    Spectrum lookupSpectrum(const Point &p) const { float q[3] = {p.x, p.y, p.z}, v[3]; /* gridvolume.cpp:386-463 */
        MER_CHECK(mer_grid_lookup_spectrum_batch(m_handle, 1, q, v)); Spectrum s; s.fromLinearRGB(v[0], v[1], v[2]); return s; }
    Float getStepSize() const { return m_stepSize; }            /* gridvolume.cpp:581 */
    Float getMaximumFloatValue() const { return 1.0f; }
    mer_grid *handle() const { return m_handle; }
    MTS_DECLARE_CLASS()
private:
    void load() {
        mer_volume_desc d;
        fillVolumeDesc(d, m_src.volumeToWorld, m_src.hasOverride ? &m_src.override_ : NULL);
        MER_CHECK(mer_grid_create_from_file(m_src.device, m_src.filename.c_str(), &d, &m_handle));
        mer_volume_desc out; int32_t enc, ch;
        MER_CHECK(mer_vol_read_header(m_src.filename.c_str(), &out, &enc, &ch));
        if (m_src.hasOverride) for (int i = 0; i < 3; ++i) { out.bbox_min[i] = m_src.override_.min[i]; out.bbox_max[i] = m_src.override_.max[i]; }
        m_aabb = AABB(Point(out.bbox_min[0], out.bbox_min[1], out.bbox_min[2]), Point(out.bbox_max[0], out.bbox_max[1], out.bbox_max[2]));
        m_stepSize = INFINITY;
        for (int i = 0; i < 3; ++i) m_stepSize = std::min(m_stepSize, 0.5f * (out.bbox_max[i] - out.bbox_min[i]) / (Float) (out.res[i] - 1));
    }
    VolumeSource m_src; Float m_stepSize; mer_grid *m_handle;
};

/* ------------------------------------------------------------------ <phase type="hg"> (src/phase/hg.cpp:46-110) */
class B200HGPhaseFunction : public PhaseFunction {
public:
    B200HGPhaseFunction(const Properties &props) : PhaseFunction(props) {
        m_g = props.getFloat("g", 0.8f);
        if (m_g >= 1 || m_g <= -1) Log(EError, "The asymmetry parameter must lie in the interval (-1, 1)!");
    }
    B200HGPhaseFunction(Stream *stream, InstanceManager *manager) : PhaseFunction(stream, manager) { m_g = stream->readFloat(); }
    void serialize(Stream *stream, InstanceManager *manager) const { PhaseFunction::serialize(stream, manager); stream->writeFloat(m_g); }
    void configure() { PhaseFunction::configure(); m_type = EAngleDependence; }
    Float sample(PhaseFunctionSamplingRecord &pRec, Sampler *sampler) const {
        Point2 s(sampler->next2D());
        float wi[3] = {pRec.wi.x, pRec.wi.y, pRec.wi.z}, xi[2] = {s.x, s.y}, wo[3], pdf;
        MER_CHECK(mer_hg_sample_batch(0, m_g, 1, wi, xi, wo, &pdf));
        pRec.wo = Vector(wo[0], wo[1], wo[2]);
        return 1.0f;
    }
    Float sample(PhaseFunctionSamplingRecord &pRec, Float &pdf, Sampler *sampler) const { sample(pRec, sampler); pdf = eval(pRec); return 1.0f; }
    Float eval(const PhaseFunctionSamplingRecord &pRec) const {
        float wi[3] = {pRec.wi.x, pRec.wi.y, pRec.wi.z}, wo[3] = {pRec.wo.x, pRec.wo.y, pRec.wo.z}, v;
        MER_CHECK(mer_hg_eval_batch(0, m_g, 1, wi, wo, &v));
        return v;
    }
    Float getMeanCosine() const { return m_g; }
    MTS_DECLARE_CLASS()
private:
    Float m_g;
};

/* ------------------------------------------------------------------ <medium type="heterogeneousrefractive"> */
class B200HeterogeneousRefractiveMedium : public Medium {
public:
    B200HeterogeneousRefractiveMedium(const Properties &props) : Medium(props), m_handle(NULL) {
        /* heterogeneousrefractive.cpp:205-297 */
        m_stepSize = props.getFloat("stepsize", 1e-3f);
        m_mediumSamplingWeight = props.getFloat("mediumSamplingWeight", -1);
        m_strategy = props.getString("strategy", "balance");
        m_channel = props.getInteger("channel", -1);
        m_samplingDensity = props.getFloat("samplingDensity", 0.0f);
        m_scale = props.getFloat("scale", 1.0f);
        m_albedo = props.getSpectrum("albedo", Spectrum(0.0f));
        m_aggressive = props.getBoolean("aggressivetracing", false);
        m_connection.tol2 = props.getFloat("tol2", 1e-6f);
        m_connection.rrweight = props.getFloat("rrweight", 1e-2f);
        m_connection.boundary_precision = props.getInteger("boundaryprecision", 3);
        m_connection.max_iterations = props.getInteger("ceresmaxiterations", 20);
        m_connection.start_mode = MER_START_DEFAULT;
        if (props.getBoolean("monochromatic", false)) Log(EError, "monochromatic=true is not carried by this path");
        if (props.getBoolean("makesensordirectconnections", false)) Log(EError, "makesensordirectconnections: use the integrator's lightTracing");
    }
    B200HeterogeneousRefractiveMedium(Stream *stream, InstanceManager *manager) : Medium(stream, manager), m_handle(NULL) {
        /* the counterpart of :343-359: children first, then the scalars */
        m_rif = static_cast<B200SplineDataSource *>(manager->getInstance(stream));
        if (stream->readBool()) m_sdf = static_cast<B200SplineDataSource *>(manager->getInstance(stream));
        if (stream->readBool()) m_density = static_cast<B200GridDataSource *>(manager->getInstance(stream));
        if (stream->readBool()) m_albedoVolume = static_cast<B200GridDataSource *>(manager->getInstance(stream));
        m_stepSize = stream->readFloat(); m_mediumSamplingWeight = stream->readFloat(); m_strategy = stream->readString();
        m_channel = stream->readInt(); m_samplingDensity = stream->readFloat(); m_scale = stream->readFloat();
        for (int i = 0; i < 3; ++i) m_albedo[i] = stream->readFloat();
        m_aggressive = stream->readBool();
        m_connection.tol2 = stream->readFloat(); m_connection.rrweight = stream->readFloat();
        m_connection.boundary_precision = stream->readInt(); m_connection.max_iterations = stream->readInt();
        m_connection.start_mode = MER_START_DEFAULT;
        configure();
    }
    virtual ~B200HeterogeneousRefractiveMedium() { mer_medium_destroy(m_handle); }
    void serialize(Stream *stream, InstanceManager *manager) const { /* :384-391 */
        Medium::serialize(stream, manager);
        manager->serialize(stream, m_rif.get());
        stream->writeBool(m_sdf.get() != NULL); if (m_sdf.get()) manager->serialize(stream, m_sdf.get());
        stream->writeBool(m_density.get() != NULL); if (m_density.get()) manager->serialize(stream, m_density.get());
        stream->writeBool(m_albedoVolume.get() != NULL); if (m_albedoVolume.get()) manager->serialize(stream, m_albedoVolume.get());
        stream->writeFloat(m_stepSize); stream->writeFloat(m_mediumSamplingWeight); stream->writeString(m_strategy);
        stream->writeInt(m_channel); stream->writeFloat(m_samplingDensity); stream->writeFloat(m_scale);
        for (int i = 0; i < 3; ++i) stream->writeFloat(m_albedo[i]);
        stream->writeBool(m_aggressive);
        stream->writeFloat(m_connection.tol2); stream->writeFloat(m_connection.rrweight);
        stream->writeInt(m_connection.boundary_precision); stream->writeInt(m_connection.max_iterations);
    }

    void addChild(const std::string &name, ConfigurableObject *child) { /* :1177-1193 */
        if (child->getClass()->derivesFrom(MTS_CLASS(VolumeDataSource))) {
            if (name == "rif") m_rif = static_cast<B200SplineDataSource *>(child);
            else if (name == "sdf") m_sdf = static_cast<B200SplineDataSource *>(child);
            else if (name == "density") m_density = static_cast<B200GridDataSource *>(child); /* heterogeneous.cpp:262-281 */
            else if (name == "albedo") { /* heterogeneous.cpp:266-268 */
                m_albedoVolume = static_cast<B200GridDataSource *>(child);
                if (!m_albedoVolume->supportsSpectrumLookups()) Log(EError, "Medium: the albedo volume must have three channels");
            }
            else Log(EError, "Medium: Invalid child node! (\"%s\")", name.c_str());
        } else {
            Medium::addChild(name, child); /* the phase function */
        }
    }

    void configure() {
        Medium::configure();
        if (m_rif.get() == NULL) Log(EError, "No RIF specified!");
        mer_medium_desc d; memset(&d, 0, sizeof(d));
        for (int i = 0; i < 3; ++i) { d.sigma_a[i] = m_sigmaA[i]; d.sigma_s[i] = m_sigmaS[i]; d.albedo[i] = m_albedo[i]; }
        d.stepsize = m_stepSize;
        d.medium_sampling_weight = m_mediumSamplingWeight;
        d.strategy = m_strategy == "balance" ? MER_STRATEGY_BALANCE : m_strategy == "single" ? MER_STRATEGY_SINGLE
                   : m_strategy == "manual" ? MER_STRATEGY_MANUAL : m_strategy == "maximum" ? MER_STRATEGY_MAXIMUM : -1;
        if (d.strategy < 0) Log(EError, "Specified an unknown sampling strategy");
        d.channel = m_channel;
        d.sampling_density = m_samplingDensity;
        /* containment predicate: the signed-distance child when there is one (:376-380, :728-739), else the interior
         * shape's AABB (hackForBox form, :722-726) */
        AABB box = m_sdf.get() ? m_sdf->getAABB() : m_shape->getAABB();
        d.shape_type = m_sdf.get() ? MER_SHAPE_SDF : MER_SHAPE_BOX;
        for (int i = 0; i < 3; ++i) { d.shape[i] = box.min[i]; d.shape[3 + i] = box.max[i]; }
        /* the container's surface: hdielectric (eta from this medium's RIF) or anything index-matched */
        d.boundary = (m_shape->getBSDF() && m_shape->getBSDF()->getClass()->getName() == "HSmoothDielectric")
                         ? MER_BOUNDARY_HDIELECTRIC : MER_BOUNDARY_INDEX_MATCHED;
        d.hg_g = m_phaseFunction->getMeanCosine();
        d.density_scale = m_scale;
        if (m_handle) mer_medium_destroy(m_handle);
        MER_CHECK(mer_medium_create(&d, m_rif->handle(), m_density.get() ? m_density->handle() : NULL, &m_handle));
        if (m_sdf.get() || m_aggressive) MER_CHECK(mer_medium_set_sdf(m_handle, m_sdf.get() ? m_sdf->handle() : NULL, m_aggressive ? 1 : 0));
        if (m_albedoVolume.get()) MER_CHECK(mer_medium_set_albedo_grid(m_handle, m_albedoVolume->handle()));
    }

    /* Medium::sampleDistance (include/mitsuba/render/medium.h:130-131): a one-ray batch that replays the sampler draws the
     * reference consumes (:404, :440) */
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
    void eval(const Ray &, MediumSamplingRecord &) const { Log(EError, "eval(ray, mRec): use the direct-connection form (mer_medium_connect_batch)"); }
    bool isHomogeneous() const { return false; }
    bool isheterogeneousrefractive() const { return true; }
    mer_medium *handle() const { return m_handle; }
    const mer_connection_params &connection() const { return m_connection; }
    /* one more replica of this medium (same files, same properties) on another GPU, for mer_render_multi */
    mer_medium *replicate(int device, mer_rif **rifOut, mer_grid **gridOut, mer_grid **albedoOut) const {
        mer_volume_desc rd; int mode; mer_medium_desc md; float sd;
        MER_CHECK(mer_rif_desc(m_rif->handle(), &rd, &mode));
        MER_CHECK(mer_medium_resolved(m_handle, &md, &sd));
        std::vector<float> coeffFree; /* the replica prefilters the file again on its own GPU */
        MER_CHECK(mer_rif_create_from_file(device, rifFile().c_str(), &rd, mode, rifOut));
        *gridOut = NULL;
        if (m_density.get()) MER_CHECK(mer_grid_create_from_file(device, gridFile().c_str(), NULL, gridOut));
        mer_medium *m = NULL;
        MER_CHECK(mer_medium_create(&md, *rifOut, *gridOut, &m));
        *albedoOut = NULL;
        if (m_albedoVolume.get()) {
            MER_CHECK(mer_grid_create_from_file(device, m_albedoVolume->getProperties().getString("filename").c_str(), NULL, albedoOut));
            MER_CHECK(mer_medium_set_albedo_grid(m, *albedoOut));
        }
        return m;
    }
    MTS_DECLARE_CLASS()
private:
    std::string rifFile() const { return m_rif->getProperties().getString("filename"); }
    std::string gridFile() const { return m_density->getProperties().getString("filename"); }
    ref<B200SplineDataSource> m_rif, m_sdf; ref<B200GridDataSource> m_density, m_albedoVolume; mer_medium *m_handle;
    Float m_stepSize, m_mediumSamplingWeight, m_samplingDensity, m_scale; std::string m_strategy; int m_channel; Spectrum m_albedo;
    bool m_aggressive; mer_connection_params m_connection;
};

/* ------------------------------------------------------------------ <integrator type="ervolpath">
 * The real entry point: Integrator::render (include/mitsuba/render/integrator.h:61-96) pulls the sensor, film, filter,
 * emitter and medium parameters out of the Scene, renders on the GPU(s) and hands the [R,G,B,alpha,weight] film back with
 * Film::setBitmap (src/films/hdrfilm.cpp:412-414). */
class EikonalVolPathIntegrator : public Integrator {
public:
    EikonalVolPathIntegrator(const Properties &props) : Integrator(props) {
        m_maxDepth = props.getInteger("maxDepth", -1);
        m_rrDepth = props.getInteger("rrDepth", 5);
        m_directConnections = props.getBoolean("directConnections", false); /* curved next-event estimation */
        m_misConnections = props.getBoolean("misConnections", false);       /* ... with volpath's power heuristic (volpath.cpp:120-147) */
        m_lightTracing = props.getBoolean("lightTracing", false);           /* emitter-side walk + sensor connections */
        m_gpus = props.getInteger("gpus", 1);                               /* 0: every GPU of the box */
        if (m_maxDepth == 0 || m_maxDepth < -1)
            Log(EError, "maxDepth must be set to -1 (infinite) or a value greater than zero!");
    }
    EikonalVolPathIntegrator(Stream *stream, InstanceManager *manager) : Integrator(stream, manager) {
        m_maxDepth = stream->readInt(); m_rrDepth = stream->readInt();
        m_directConnections = stream->readBool(); m_misConnections = stream->readBool(); m_lightTracing = stream->readBool(); m_gpus = stream->readInt();
    }
    void serialize(Stream *stream, InstanceManager *manager) const {
        Integrator::serialize(stream, manager);
        stream->writeInt(m_maxDepth); stream->writeInt(m_rrDepth);
        stream->writeBool(m_directConnections); stream->writeBool(m_misConnections); stream->writeBool(m_lightTracing); stream->writeInt(m_gpus);
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
        Point o = toWorld(Point(0.0f)), ahead = toWorld(Point(0, 0, 1)), above = toWorld(Point(0, 1, 0));
        for (int i = 0; i < 3; ++i) { r.cam_origin[i] = o[i]; r.cam_target[i] = ahead[i]; r.cam_up[i] = above[i] - o[i]; }
        r.fov_deg = static_cast<const PerspectiveCamera *>(sensor)->getXFov();
        r.filter = film->getReconstructionFilter()->getRadius() > 1 ? MER_FILTER_GAUSSIAN : MER_FILTER_BOX;
        r.max_depth = m_maxDepth; r.rr_depth = m_rrDepth;
        r.direct_connections = m_misConnections ? 2 : (m_directConnections ? 1 : 0);
        r.light_tracing = m_lightTracing ? 1 : 0;
        /* the fork's transient film (src/librender/film.cpp:56-78): frames, bounds, calibration */
        if (film->getDecompositionType() == Film::ETransient && film->getFrames() > 1) {
            r.frames = (int32_t) film->getFrames();
            r.min_bound = film->getDecompositionMinBound();
            r.bin_width = film->getDecompositionBinWidth();
            r.calibrated_transient = film->isCalibratedTransient() ? 1 : 0;
        }
        /* ---- emitters: the constant environment, one rectangle area light (two-sided quad), one collimated beam */
        if (scene->hasEnvironmentEmitter()) {
            Spectrum L = scene->getEnvironmentEmitter()->evalEnvironment(RayDifferential(Point(0.0f), Vector(0, 0, 1), 0));
            for (int i = 0; i < 3; ++i) r.env_radiance[i] = L[i];
        }
        const ref_vector<Emitter> &emitters = scene->getEmitters();
        for (size_t e = 0; e < emitters.size(); ++e) {
            const Emitter *em = emitters[e].get();
            const std::string &type = em->getProperties().getPluginName();
            if (type == "area" && em->getShape() != NULL) { /* src/emitters/area.cpp on a <shape type="rectangle">: (-1,-1,0) .. (1,1,0) under toWorld */
                ref<TriMesh> mesh = const_cast<Shape *>(em->getShape())->createTriMesh();
                if (mesh->getVertexCount() != 4) Log(EError, "ervolpath: the area emitter must sit on a rectangle");
                const Point *v = mesh->getVertexPositions();
                Spectrum L = em->getProperties().getSpectrum("radiance", Spectrum(1.0f));
                r.has_quad = 1;
                for (int i = 0; i < 3; ++i) { r.quad_origin[i] = v[0][i]; r.quad_u[i] = v[1][i] - v[0][i]; r.quad_v[i] = v[3][i] - v[0][i]; r.quad_radiance[i] = L[i]; }
            } else if (type == "collimated") { /* src/emitters/collimated.cpp:59-110 */
                Transform t = em->getProperties().getTransform("toWorld", Transform());
                Point bo = t(Point(0.0f)), bz = t(Point(0, 0, 1));
                Spectrum P = em->getProperties().getSpectrum("power", Spectrum(1.0f));
                r.emitter_type = MER_EMITTER_COLLIMATED;
                for (int i = 0; i < 3; ++i) { r.beam_origin[i] = bo[i]; r.beam_direction[i] = bz[i] - bo[i]; r.beam_power[i] = P[i]; }
            } else if (!em->isEnvironmentEmitter()) {
                Log(EError, "ervolpath: emitter type \"%s\" is not on this path (constant environment, rectangle area light, collimated beam)", type.c_str());
            }
        }
        const B200HeterogeneousRefractiveMedium *medium = NULL;
        for (size_t i = 0; i < scene->getMedia().size(); ++i)
            if (scene->getMedia()[i]->isheterogeneousrefractive())
                medium = static_cast<const B200HeterogeneousRefractiveMedium *>(scene->getMedia()[i].get());
        if (!medium) Log(EError, "ervolpath needs a heterogeneousrefractive medium");
        r.connection = medium->connection();
        ref<Bitmap> bitmap = r.frames > 1 ? new Bitmap(Bitmap::EMultiSpectrumAlphaWeight, Bitmap::EFloat32, size, 3 * r.frames + 2)
                                          : new Bitmap(Bitmap::ESpectrumAlphaWeight, Bitmap::EFloat32, size);
        mer_render_stats stats;
        int gpus = m_gpus == 0 ? mer_device_count() : m_gpus;
        if (gpus <= 1) {
            MER_CHECK(mer_render(medium->handle(), &r, bitmap->getFloat32Data(), &stats));
        } else { /* the block scheduler (integrator.cpp:95-127) becomes one replica of the scene per GPU + one film reduce */
            std::vector<const mer_medium *> media(1, medium->handle());
            std::vector<mer_medium *> owned; std::vector<mer_rif *> rifs; std::vector<mer_grid *> grids;
            for (int g = 1; g < gpus; ++g) {
                mer_rif *rf; mer_grid *gr, *al;
                owned.push_back(medium->replicate(g, &rf, &gr, &al));
                rifs.push_back(rf); grids.push_back(gr); grids.push_back(al);
                media.push_back(owned.back());
            }
            int rc = mer_render_multi(&media[0], gpus, &r, bitmap->getFloat32Data(), &stats);
            for (size_t i = 0; i < owned.size(); ++i) { mer_medium_destroy(owned[i]); mer_rif_destroy(rifs[i]); }
            for (size_t i = 0; i < grids.size(); ++i) mer_grid_destroy(grids[i]);
            MER_CHECK(rc);
        }
        Log(EInfo, "ervolpath: %llu samples, %llu eikonal steps, %.1f ms on %d GPU(s)",
            (unsigned long long) stats.samples, (unsigned long long) stats.ray_steps, stats.device_ms, gpus);
        film->setBitmap(bitmap);
        return true;
    }
    void cancel() { }
    MTS_DECLARE_CLASS()
private:
    int m_maxDepth, m_rrDepth, m_gpus;
    bool m_directConnections, m_misConnections, m_lightTracing;
};

MTS_IMPLEMENT_CLASS_S(B200SplineDataSource, false, VolumeDataSource)
MTS_IMPLEMENT_CLASS_S(B200GridDataSource, false, VolumeDataSource)
MTS_IMPLEMENT_CLASS_S(B200HGPhaseFunction, false, PhaseFunction)
MTS_IMPLEMENT_CLASS_S(B200HeterogeneousRefractiveMedium, false, Medium)
MTS_IMPLEMENT_CLASS_S(EikonalVolPathIntegrator, false, Integrator)
/* one export per plugin .so (cobject.h:99-107): the build defines which tag this object file is */
#if defined(MER_PLUGIN_SPLINEVOLUME)
MTS_EXPORT_PLUGIN(B200SplineDataSource, "Cubic B-spline refractive-index volume (B200)");
#elif defined(MER_PLUGIN_GRIDVOLUME)
MTS_EXPORT_PLUGIN(B200GridDataSource, "Grid data source (B200)");
#elif defined(MER_PLUGIN_HG)
MTS_EXPORT_PLUGIN(B200HGPhaseFunction, "Henyey-Greenstein phase function (B200)");
#elif defined(MER_PLUGIN_HETEROGENEOUSREFRACTIVE)
MTS_EXPORT_PLUGIN(B200HeterogeneousRefractiveMedium, "Heterogeneous refractive medium (B200)");
#elif defined(MER_PLUGIN_ERVOLPATH)
MTS_EXPORT_PLUGIN(EikonalVolPathIntegrator, "Eikonal volumetric path tracer (B200)");
#endif
MTS_NAMESPACE_END
