/*
 * host/mer_host.hpp — C++ host side above the C ABI: a minimal mirror of the Mitsuba 0.5 plugin interface
 * for the eikonal path (same plugin tags, property names, child names, defaults and error messages), so that
 * a Mitsuba scene XML of the volumetric kind loads and renders without Mitsuba's framework (which cannot be
 * built in this image: Boost / Xerces / OpenEXR / SCons are absent).
 *
 * Mirrors (paths relative to the MitsubaER tree):
 *   Properties                      include/mitsuba/core/properties.h (typed getters with defaults, "unqueried" check)
 *   PluginManager::createObject     src/libcore/plugin.cpp:221-246      -> createObject() below (static registry)
 *   SceneHandler (Xerces SAX)       src/librender/scenehandler.cpp:210-219 ($name substitution), 712-777 (instantiation)
 *   SplineDataSource / GridDataSource / HGPhaseFunction / HeterogeneousRefractiveMedium ctor+addChild+configure
 *   Log(EError, ...) throwing       src/libcore/logger.cpp:100-147     -> std::runtime_error
 * Everything numerical happens in libmitsubaer_b200.so.
 */
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "mitsubaer_b200.h"

namespace merhost {

/* --dry-run: parse, instantiate and resolve everything but create no device handles (lets the XML layer be
 * tested on a machine without a GPU; the library itself has no CPU path) */
inline bool &dryRun() { static bool v = false; return v; }
/* the GPU a volume without a `device` property is created on: `mer_render --gpus N` loads the scene once per GPU */
inline int &defaultDevice() { static int v = 0; return v; }

[[noreturn]] inline void logError(const std::string &msg) { throw std::runtime_error(msg); } /* Log(EError, ...) */
inline void merCheck(int rc) { if (rc != MER_OK) logError(std::string("mitsubaer_b200: ") + mer_last_error()); }

struct Vec3 { double x = 0, y = 0, z = 0; };
struct Spectrum3 { float c[3] = {0, 0, 0}; };

/* 4x4 row-major affine transform, composed like Mitsuba's <transform> block (each element left-multiplies) */
struct Transform {
    double m[16];
    Transform() { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    static Transform translate(Vec3 t) { Transform r; r.m[3] = t.x; r.m[7] = t.y; r.m[11] = t.z; return r; }
    static Transform scale(Vec3 s) { Transform r; r.m[0] = s.x; r.m[5] = s.y; r.m[10] = s.z; return r; }
    static Transform rotate(Vec3 axis, double deg) {
        double len = std::sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
        if (len == 0) logError("rotate: zero axis");
        double x = axis.x / len, y = axis.y / len, z = axis.z / len, a = deg * M_PI / 180.0, c = std::cos(a), s = std::sin(a);
        Transform r;
        r.m[0] = x * x + (1 - x * x) * c; r.m[1] = x * y * (1 - c) - z * s; r.m[2] = x * z * (1 - c) + y * s;
        r.m[4] = x * y * (1 - c) + z * s; r.m[5] = y * y + (1 - y * y) * c; r.m[6] = y * z * (1 - c) - x * s;
        r.m[8] = x * z * (1 - c) - y * s; r.m[9] = y * z * (1 - c) + x * s; r.m[10] = z * z + (1 - z * z) * c;
        return r;
    }
    Transform operator*(const Transform &o) const {
        Transform r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += m[4 * i + k] * o.m[4 * k + j];
                r.m[4 * i + j] = s;
            }
        return r;
    }
    Vec3 point(Vec3 p) const { return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]}; }
    Vec3 vector(Vec3 v) const { return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z}; }
    bool isIdentity() const { Transform i; return std::memcmp(m, i.m, sizeof(m)) == 0; }
    bool inverse(Transform &out) const { /* affine inverse by Gauss-Jordan on the 4x4 */
        double a[4][8];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { a[i][j] = m[4 * i + j]; a[i][4 + j] = i == j; }
        for (int c = 0; c < 4; c++) {
            int piv = c;
            for (int r = c + 1; r < 4; r++) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
            if (std::fabs(a[piv][c]) < 1e-300) return false;
            for (int j = 0; j < 8; j++) std::swap(a[c][j], a[piv][j]);
            double d = a[c][c];
            for (int j = 0; j < 8; j++) a[c][j] /= d;
            for (int r = 0; r < 4; r++) if (r != c) { double f = a[r][c]; for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j]; }
        }
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out.m[4 * i + j] = a[i][4 + j];
        return true;
    }
};

/* ------------------------------------------------------------------ Properties */
class Properties {
public:
    std::string pluginName, id;
    bool has(const std::string &n) const { return m_str.count(n) || m_tr.count(n); }
    void setRaw(const std::string &n, const std::string &type, const std::string &v) {
        if (m_str.count(n)) logError("Property \"" + n + "\" was specified multiple times!");
        m_str[n] = {type, v};
    }
    void setTransform(const std::string &n, const Transform &t) { m_tr[n] = t; }
    std::string getString(const std::string &n, const std::string &def) const { return has(n) ? raw(n, "string") : def; }
    std::string getString(const std::string &n) const { need(n); return raw(n, "string"); }
    double getFloat(const std::string &n, double def) const { return m_str.count(n) ? num(raw(n, "float")) : def; }
    double getFloat(const std::string &n) const { need(n); return num(raw(n, "float")); }
    long getInteger(const std::string &n, long def) const { return m_str.count(n) ? (long) num(raw(n, "integer")) : def; }
    bool getBoolean(const std::string &n, bool def) const { return m_str.count(n) ? raw(n, "boolean") == "true" : def; }
    Spectrum3 getSpectrum(const std::string &n, float def) const {
        Spectrum3 s;
        if (!m_str.count(n)) { s.c[0] = s.c[1] = s.c[2] = def; return s; }
        std::vector<double> v = numbers(rawAny(n));
        if (v.size() == 1) v = {v[0], v[0], v[0]};
        if (v.size() != 3) logError("spectrum \"" + n + "\" must have 1 or 3 components (SPECTRUM_SAMPLES=3)");
        for (int i = 0; i < 3; i++) s.c[i] = (float) v[i];
        return s;
    }
    Vec3 getPoint(const std::string &n) const { need(n); std::vector<double> v = numbers(rawAny(n)); if (v.size() != 3) logError("point \"" + n + "\" needs x, y, z"); return {v[0], v[1], v[2]}; }
    Transform getTransform(const std::string &n, const Transform &def) const { auto it = m_tr.find(n); if (it == m_tr.end()) return def; m_q.insert(n); return it->second; }
    /* ConfigurableObject's "unqueried properties" warning is an error here: a typo must not be silently ignored */
    std::vector<std::string> unqueried() const {
        std::vector<std::string> r;
        for (auto &kv : m_str) if (!m_q.count(kv.first)) r.push_back(kv.first);
        for (auto &kv : m_tr) if (!m_q.count(kv.first)) r.push_back(kv.first);
        return r;
    }
    static std::vector<double> numbers(const std::string &s) {
        std::vector<double> out;
        std::string t = s;
        for (char &c : t) if (c == ',') c = ' ';
        std::istringstream is(t);
        double d;
        while (is >> d) out.push_back(d);
        return out;
    }
private:
    struct Entry { std::string type, value; };
    std::map<std::string, Entry> m_str;
    std::map<std::string, Transform> m_tr;
    mutable std::set<std::string> m_q;
    void need(const std::string &n) const { if (!m_str.count(n)) logError("Property \"" + n + "\" has not been specified!"); }
    const std::string &rawAny(const std::string &n) const { m_q.insert(n); return m_str.at(n).value; }
    const std::string &raw(const std::string &n, const std::string &type) const {
        const Entry &e = m_str.at(n);
        if (e.type != type && !(type == "float" && e.type == "integer"))
            logError("The property \"" + n + "\" has the wrong type (expected <" + type + ">).");
        m_q.insert(n);
        return e.value;
    }
    static double num(const std::string &s) {
        char *end = nullptr;
        double d = std::strtod(s.c_str(), &end);
        if (end == s.c_str()) logError("Could not parse \"" + s + "\" as a number");
        return d;
    }
};

/* ------------------------------------------------------------------ objects */
struct Object {
    Properties props;
    virtual ~Object() {}
    virtual const char *className() const = 0;
    virtual void addChild(const std::string &, std::shared_ptr<Object>) { logError(std::string(className()) + ": Invalid child node!"); }
    virtual void configure() {}
};
typedef std::shared_ptr<Object> ObjectRef;

inline void fillVolumeDesc(mer_volume_desc &d, const Properties &p) {
    std::memset(&d, 0, sizeof(d));
    Transform toWorld = p.getTransform("toWorld", Transform()), inv;
    if (!toWorld.inverse(inv)) logError("toWorld is singular");
    d.has_transform = toWorld.isIdentity() ? 0 : 1;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) d.world_to_volume[4 * r + c] = (float) inv.m[4 * r + c];
    if (p.has("min") && p.has("max")) { /* splinevolume.cpp:93-98 */
        Vec3 lo = p.getPoint("min"), hi = p.getPoint("max");
        d.bbox_min[0] = (float) lo.x; d.bbox_min[1] = (float) lo.y; d.bbox_min[2] = (float) lo.z;
        d.bbox_max[0] = (float) hi.x; d.bbox_max[1] = (float) hi.y; d.bbox_max[2] = (float) hi.z;
    }
}

struct SplineDataSource : Object { /* <volume type="splinevolume"> */
    mer_rif *handle = nullptr;
    mer_volume_desc desc;
    const char *className() const override { return "SplineDataSource"; }
    void configure() override {
        mer_volume_desc ov;
        fillVolumeDesc(ov, props);
        props.getBoolean("sendData", false);
        std::string fetch = props.getString("fetch", "tricubic");
        int mode = fetch == "tricubic" ? MER_RIF_TRICUBIC : fetch == "trilinear_packed" ? MER_RIF_TRILINEAR_PACKED : -1;
        if (mode < 0) logError("splinevolume: unknown fetch mode \"" + fetch + "\"");
        const int device = (int) props.getInteger("device", defaultDevice());
        const std::string file = props.getString("filename");
        if (dryRun()) { int32_t enc, ch; merCheck(mer_vol_read_header(file.c_str(), &desc, &enc, &ch)); return; }
        merCheck(mer_rif_create_from_file(device, file.c_str(), &ov, mode, &handle));
        merCheck(mer_rif_desc(handle, &desc, nullptr));
    }
    ~SplineDataSource() override { mer_rif_destroy(handle); }
};

struct GridDataSource : Object { /* <volume type="gridvolume"> */
    mer_grid *handle = nullptr;
    int channels = 1; /* the file header's channel count: 1 = density, 3 = albedo (gridvolume.cpp:251-262, 578-579) */
    const char *className() const override { return "GridDataSource"; }
    void configure() override {
        mer_volume_desc ov;
        fillVolumeDesc(ov, props);
        props.getBoolean("sendData", false);
        const int device = (int) props.getInteger("device", defaultDevice());
        const std::string file = props.getString("filename");
        { mer_volume_desc d; int32_t enc, ch; merCheck(mer_vol_read_header(file.c_str(), &d, &enc, &ch)); channels = ch; }
        if (dryRun()) return;
        merCheck(mer_grid_create_from_file(device, file.c_str(), &ov, &handle));
    }
    ~GridDataSource() override { mer_grid_destroy(handle); }
};

struct HGPhaseFunction : Object { /* <phase type="hg"> */
    float g = 0.8f;
    const char *className() const override { return "HGPhaseFunction"; }
    void configure() override {
        g = (float) props.getFloat("g", 0.8);
        if (g >= 1 || g <= -1) logError("The asymmetry parameter must lie in the interval (-1, 1)!"); /* hg.cpp:50-52 */
    }
};

struct SurfaceBsdf : Object { /* <bsdf> on the container: "null" (index-matched) or "hdielectric" (src/bsdfs/hdielectric.cpp) */
    const char *className() const override { return "SurfaceBsdf"; }
    void addChild(const std::string &, ObjectRef) override {}
    void configure() override { /* specularReflectance / specularTransmittance other than 1 are not carried */
        for (const char *n : {"specularReflectance", "specularTransmittance"})
            if (props.has(n)) { Spectrum3 v = props.getSpectrum(n, 1.0f); if (v.c[0] != 1 || v.c[1] != 1 || v.c[2] != 1) logError(std::string("hdielectric: ") + n + " must be 1 on this path"); }
    }
};

struct Shape : Object { /* the medium's container: <shape type="cube"|"sphere"> */
    int shapeType = MER_SHAPE_BOX;
    int boundary = MER_BOUNDARY_INDEX_MATCHED;
    float params[6] = {-1, -1, -1, 1, 1, 1};
    ObjectRef interior;
    const char *className() const override { return "Shape"; }
    void addChild(const std::string &name, ObjectRef child) override {
        if (name == "interior") interior = child;
        else if (auto b = std::dynamic_pointer_cast<SurfaceBsdf>(child)) {
            if (b->props.pluginName == "hdielectric") boundary = MER_BOUNDARY_HDIELECTRIC;
            else if (b->props.pluginName != "null") logError("Shape: the container's bsdf must be \"null\" or \"hdielectric\" on this path (got \"" + b->props.pluginName + "\")");
        }
        else if (name == "exterior" || std::string(child->className()) == "Ignored") { /* bsdf, exterior medium: not on this path */ }
        else logError("Shape: Invalid child node! (\"" + name + "\")");
    }
    void configure() override {
        Transform t = props.getTransform("toWorld", Transform());
        if (props.pluginName == "cube") {
            /* Mitsuba's cube is [-1,1]^3 in object space; the containment predicate is its world AABB (hackForBox form) */
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (int i = 0; i < 8; i++) {
                Vec3 c = t.point({(i & 1) ? 1.0 : -1.0, (i & 2) ? 1.0 : -1.0, (i & 4) ? 1.0 : -1.0});
                lo[0] = std::min(lo[0], c.x); lo[1] = std::min(lo[1], c.y); lo[2] = std::min(lo[2], c.z);
                hi[0] = std::max(hi[0], c.x); hi[1] = std::max(hi[1], c.y); hi[2] = std::max(hi[2], c.z);
            }
            shapeType = MER_SHAPE_BOX;
            for (int i = 0; i < 3; i++) { params[i] = (float) lo[i]; params[3 + i] = (float) hi[i]; }
        } else if (props.pluginName == "sphere") {
            Vec3 c = props.has("center") ? props.getPoint("center") : Vec3{0, 0, 0};
            double r = props.getFloat("radius", 1.0);
            Vec3 cw = t.point(c), sx = t.vector({r, 0, 0});
            shapeType = MER_SHAPE_SPHERE;
            params[0] = (float) cw.x; params[1] = (float) cw.y; params[2] = (float) cw.z;
            params[3] = (float) std::sqrt(sx.x * sx.x + sx.y * sx.y + sx.z * sx.z); params[4] = params[5] = 0;
        } else if (props.pluginName == "obj" || props.pluginName == "ply" || props.pluginName == "serialized") {
            shapeType = MER_SHAPE_SDF; /* the mesh itself is not read: the medium's sdf child volume describes it */
            props.getString("filename", "");
            for (const char *n : {"faceNormals", "flipNormals", "flipTexCoords", "collapse"}) props.getBoolean(n, false);
            props.getFloat("maxSmoothAngle", 0.0);
        } else {
            logError("shape type \"" + props.pluginName + "\" cannot contain a heterogeneousrefractive medium on this path (cube, sphere, or a mesh with an sdf volume)");
        }
    }
};

struct HeterogeneousRefractiveMedium : Object { /* <medium type="heterogeneousrefractive"> */
    std::shared_ptr<SplineDataSource> rif, sdf;
    std::shared_ptr<GridDataSource> density, albedo; /* heterogeneous.cpp:262-271 */
    std::shared_ptr<HGPhaseFunction> phase;
    mer_medium_desc desc;
    mer_connection_params connection;
    mer_medium *handle = nullptr;
    const char *className() const override { return "HeterogeneousRefractiveMedium"; }
    void addChild(const std::string &name, ObjectRef child) override {
        if (auto p = std::dynamic_pointer_cast<HGPhaseFunction>(child)) { if (phase) logError("Medium: phase function already set"); phase = p; }
        else if (name == "rif" && std::dynamic_pointer_cast<SplineDataSource>(child)) rif = std::dynamic_pointer_cast<SplineDataSource>(child);
        else if (name == "density" && std::dynamic_pointer_cast<GridDataSource>(child)) density = std::dynamic_pointer_cast<GridDataSource>(child);
        else if (name == "albedo" && std::dynamic_pointer_cast<GridDataSource>(child)) {
            albedo = std::dynamic_pointer_cast<GridDataSource>(child);
            if (albedo->channels != 3) logError("Medium: the albedo volume must support spectrum lookups (3 channels, heterogeneous.cpp:266-268)");
        }
        else if (name == "sdf" && std::dynamic_pointer_cast<SplineDataSource>(child)) sdf = std::dynamic_pointer_cast<SplineDataSource>(child); /* :376-380 */
        else logError("Medium: Invalid child node! (\"" + std::string(child->className()) + "\")");
    }
    void configure() override { /* properties are resolved here, the handle is created by attach() once the shape is known */
        std::memset(&desc, 0, sizeof(desc));
        double scale = props.getFloat("scale", 1.0);
        if (props.has("material")) logError("material presets are not carried by this path; give sigmaS/sigmaA or sigmaT/albedo");
        Spectrum3 albedo = props.getSpectrum("albedo", 0.0f);
        if (props.has("sigmaT")) { /* medium/materials.h:112-120 */
            if (!props.has("albedo")) logError("Medium: sigmaT needs albedo (materials.h:112-120)");
            Spectrum3 st = props.getSpectrum("sigmaT", 0.0f);
            for (int i = 0; i < 3; i++) { desc.sigma_s[i] = (float) (st.c[i] * scale * albedo.c[i]); desc.sigma_a[i] = (float) (st.c[i] * scale * (1 - albedo.c[i])); }
        } else {
            Spectrum3 ss = props.getSpectrum("sigmaS", 0.0f), sa = props.getSpectrum("sigmaA", 0.0f);
            for (int i = 0; i < 3; i++) { desc.sigma_s[i] = (float) (ss.c[i] * scale); desc.sigma_a[i] = (float) (sa.c[i] * scale); }
        }
        for (int i = 0; i < 3; i++) desc.albedo[i] = albedo.c[i];
        desc.stepsize = (float) props.getFloat("stepsize", 1e-3);
        desc.medium_sampling_weight = (float) props.getFloat("mediumSamplingWeight", -1);
        std::string s = props.getString("strategy", "balance");
        desc.strategy = s == "balance" ? MER_STRATEGY_BALANCE : s == "single" ? MER_STRATEGY_SINGLE : s == "manual" ? MER_STRATEGY_MANUAL : s == "maximum" ? MER_STRATEGY_MAXIMUM : -1;
        if (desc.strategy < 0) logError("Specified an unknown sampling strategy");
        desc.channel = (int) props.getInteger("channel", -1);
        desc.sampling_density = (float) props.getFloat("samplingDensity", 0.0);
        desc.density_scale = (float) scale;
        const std::string scaling = props.getString("radianceScaling", "reference");
        if (scaling != "reference" && scaling != "physical") logError("radianceScaling must be \"reference\" or \"physical\"");
        desc.radiance_scaling = scaling == "physical" ? MER_SCALING_PHYSICAL : MER_SCALING_REFERENCE;
        if (props.getBoolean("monochromatic", false)) /* librender/medium.cpp: collapses sigmaA / sigmaS to their average */
            logError("monochromatic=true is not carried by this path (give the same value for the three channels)");
        /* solver parameters of the direct connections (heterogeneousrefractive.cpp:208-219), used when the integrator asks for them */
        connection.tol2 = (float) props.getFloat("tol2", 1e-6);
        connection.rrweight = (float) props.getFloat("rrweight", 1e-2);
        connection.boundary_precision = (int) props.getInteger("boundaryprecision", 3);
        connection.max_iterations = (int) props.getInteger("ceresmaxiterations", 20);
        connection.start_mode = MER_START_DEFAULT;
        /* accepted for XML compatibility: Ceres-specific tolerances have no counterpart in the Levenberg-Marquardt solver */
        for (const char *n : {"ceresfunctiontolerance", "ceresgradienttolerance", "ceresparametertolerance"}) props.getFloat(n, 0);
        props.getBoolean("cerescheckgradients", false);
        /* these two change what the reference computes (:230, :476-493; edge.cpp:535-567): refuse rather than render differently */
        if (props.getBoolean("aggressivetracing", false))
            logError("aggressivetracing=true: the renderer steps without it (it is available in mer_medium_sample_distance_batch only)");
        if (props.getBoolean("makesensordirectconnections", false))
            logError("makesensordirectconnections=true belongs to the reference's bidirectional integrator; use the integrator's lightTracing here");
        if (!rif) logError("No RIF specified!");
    }
    void attach(const Shape &shape) {
        desc.shape_type = shape.shapeType;
        for (int i = 0; i < 6; i++) desc.shape[i] = shape.params[i];
        if (shape.shapeType == MER_SHAPE_SDF) { /* a mesh container: containment by the sdf child, bounded by that volume's box */
            if (!sdf) logError("shape type \"" + shape.props.pluginName + "\": a mesh can contain the medium only through an <volume name=\"sdf\"> child (signed-distance grid)");
            for (int i = 0; i < 3; i++) { desc.shape[i] = sdf->desc.bbox_min[i]; desc.shape[3 + i] = sdf->desc.bbox_max[i]; }
        }
        desc.boundary = shape.boundary;
        desc.hg_g = phase ? phase->g : 0.0f; /* Medium::configure: isotropic default */
        if (dryRun()) return;
        merCheck(mer_medium_create(&desc, rif->handle, density ? density->handle : nullptr, &handle));
        if (sdf) merCheck(mer_medium_set_sdf(handle, sdf->handle, 0));
        if (albedo) merCheck(mer_medium_set_albedo_grid(handle, albedo->handle));
    }
    ~HeterogeneousRefractiveMedium() override { mer_medium_destroy(handle); }
};

struct Ignored : Object { /* plugins that exist in volumetric scenes but are not on this path (bsdf, sampler internals) */
    const char *className() const override { return "Ignored"; }
    void addChild(const std::string &, ObjectRef) override {}
};

struct Generic : Object { /* sensor, film, rfilter, sampler, emitter, integrator: plain property bags read by Scene */
    std::vector<std::pair<std::string, ObjectRef>> children;
    const char *className() const override { return "Generic"; }
    void addChild(const std::string &name, ObjectRef c) override { children.push_back({name, c}); }
    std::shared_ptr<Generic> child(const std::string &tagType) const {
        for (auto &kv : children) if (auto g = std::dynamic_pointer_cast<Generic>(kv.second)) if (g->tag == tagType) return g;
        return nullptr;
    }
    std::string tag;
};

/* ------------------------------------------------------------------ scene */
struct Scene {
    mer_render_desc render;
    std::shared_ptr<HeterogeneousRefractiveMedium> medium;
    std::vector<ObjectRef> keepAlive;
    std::string integratorType;
};

Scene loadScene(const std::string &xmlPath, const std::map<std::string, std::string> &params);
void writePFM(const std::string &path, int w, int h, const float *rgb);

} /* namespace merhost */
