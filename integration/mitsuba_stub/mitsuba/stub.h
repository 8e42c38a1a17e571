/*
 * integration/mitsuba_stub/mitsuba/stub.h — DECLARATIONS ONLY, for a syntax check.
 *
 * A minimal stand-in for the Mitsuba 0.5 headers that integration/mitsuba_plugins.cpp includes, so that the binding can
 * be compiled with `g++ -fsyntax-only` in an image that has neither Mitsuba's framework nor Boost (tests/test_host_cpp.py
 * does that for every plugin tag).  Each class carries exactly the members the binding touches, with the signatures of
 * the MitsubaER headers cited next to them; nothing here is linked or shipped.  Inside a real Mitsuba tree this directory
 * is simply not on the include path.
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include <cmath>
#include <cstdarg>
#include <string>
#include <vector>

#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
#define MTS_DECLARE_CLASS() virtual const Class *getClass() const; static Class *m_theClass;
#define MTS_CLASS(x) x::m_theClass
#define MTS_IMPLEMENT_CLASS(name, abstract, super) Class *name::m_theClass = 0; const Class *name::getClass() const { return m_theClass; }
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super) MTS_IMPLEMENT_CLASS(name, abstract, super) \
    static Object *name##_unserialize(Stream *s, InstanceManager *m) { return new name(s, m); } /* class.h:219: needs the (Stream*, InstanceManager*) ctor */
/* include/mitsuba/core/cobject.h:99-107 */
#define MTS_EXPORT_PLUGIN(name, descr) extern "C" { void *CreateInstance(const Properties &props) { return new name(props); } \
    const char *GetDescription() { return descr; } }

namespace fs { struct path { std::string s; path() {} path(const std::string &v) : s(v) {} std::string string() const { return s; } }; }

MTS_NAMESPACE_BEGIN
typedef float Float;
#if defined(FLOATDEBUG) /* include/mitsuba/core/fwd.h:174-184 */
typedef double FLOAT;
#else
typedef float FLOAT;
#endif
enum ELogLevel { ETrace, EDebug, EInfo, EWarn, EError };
void Log(ELogLevel level, const char *fmt, ...); /* logger.h; EError throws (logger.cpp:100-147) */
#define SLog Log

template <typename T> struct TVec3 { T x, y, z; TVec3() : x(0), y(0), z(0) {} TVec3(T a, T b, T c) : x(a), y(b), z(c) {} explicit TVec3(T a) : x(a), y(a), z(a) {}
    T &operator[](int i) { return (&x)[i]; } const T &operator[](int i) const { return (&x)[i]; }
    TVec3 operator-(const TVec3 &o) const { return TVec3(x - o.x, y - o.y, z - o.z); } TVec3 operator+(const TVec3 &o) const { return TVec3(x + o.x, y + o.y, z + o.z); } };
typedef TVec3<Float> Vector; typedef TVec3<Float> Point; typedef TVec3<Float> Normal;
typedef TVec3<FLOAT> VectorF; typedef TVec3<FLOAT> PointF;
struct Point2 { Float x, y; Point2() : x(0), y(0) {} Point2(Float a, Float b) : x(a), y(b) {} };
struct Vector2i { int x, y; };
struct Spectrum { Float s[3]; Spectrum() { s[0] = s[1] = s[2] = 0; } explicit Spectrum(Float v) { s[0] = s[1] = s[2] = v; }
    Float &operator[](int i) { return s[i]; } const Float &operator[](int i) const { return s[i]; }
    void fromLinearRGB(Float r, Float g, Float b) { s[0] = r; s[1] = g; s[2] = b; } /* spectrum.h:793-797 (RGB build) */ };
struct Matrix4x4 { Float m[4][4]; Float operator()(int r, int c) const { return m[r][c]; } bool isIdentity() const; };
struct Transform { const Matrix4x4 &getMatrix() const; const Matrix4x4 &getInverseMatrix() const; Point operator()(const Point &p) const; };
struct AABB { Point min, max; AABB() {} AABB(const Point &a, const Point &b) : min(a), max(b) {} };
struct Ray { Point o; Float mint; Vector d; Float maxt; Float time; Ray() : mint(1e-4f), maxt(INFINITY), time(0) {} Ray(const Point &o_, const Vector &d_, Float t) : o(o_), mint(1e-4f), d(d_), maxt(INFINITY), time(t) {} };
typedef Ray RayDifferential;

class Stream { public: Float readFloat(); void writeFloat(Float); int readInt(); void writeInt(int); bool readBool(); void writeBool(bool);
    std::string readString(); void writeString(const std::string &); };
class Object; class ConfigurableObject; class Class;
class InstanceManager { public: Object *getInstance(Stream *); void serialize(Stream *, const Object *); };
class Class { public: bool derivesFrom(const Class *) const; const std::string &getName() const; };
class Object { public: virtual ~Object() {} virtual const Class *getClass() const; static Class *m_theClass; void incRef() const; void decRef() const; };
template <typename T> class ref { T *p; public: ref() : p(0) {} ref(T *q) : p(q) {} T *get() const { return p; } T *operator->() const { return p; } operator T *() const { return p; } ref &operator=(T *q) { p = q; return *this; } };
template <typename T> struct ref_vector : std::vector<ref<T> > {};

class Properties { public: /* include/mitsuba/core/properties.h */
    bool hasProperty(const std::string &) const; const std::string &getPluginName() const;
    Float getFloat(const std::string &, const Float &def) const; int getInteger(const std::string &, const int &def) const;
    bool getBoolean(const std::string &, const bool &def) const; std::string getString(const std::string &) const;
    std::string getString(const std::string &, const std::string &def) const; Transform getTransform(const std::string &, const Transform &def) const;
    Point getPoint(const std::string &) const; Vector getVector(const std::string &, const Vector &def) const;
    Spectrum getSpectrum(const std::string &, const Spectrum &def) const; };
class ConfigurableObject : public Object { public: /* include/mitsuba/core/cobject.h:48-77 */
    ConfigurableObject(const Properties &p) : m_properties(p) {} ConfigurableObject(Stream *, InstanceManager *) {}
    virtual void addChild(const std::string &name, ConfigurableObject *child); virtual void configure(); virtual void serialize(Stream *, InstanceManager *) const;
    inline const Properties &getProperties() const { return m_properties; } MTS_DECLARE_CLASS() protected: Properties m_properties; };

class FileResolver { public: fs::path resolve(const fs::path &) const; };
class Thread { public: static Thread *getThread(); FileResolver *getFileResolver(); };

class Sampler : public ConfigurableObject { public: Sampler(const Properties &p) : ConfigurableObject(p) {} virtual Float next1D() = 0; virtual Point2 next2D() = 0; size_t getSampleCount() const; MTS_DECLARE_CLASS() };
struct PhaseFunctionSamplingRecord { Vector wi, wo; }; /* phase.h:33-100 */
class PhaseFunction : public ConfigurableObject { public: PhaseFunction(const Properties &p) : ConfigurableObject(p) {} PhaseFunction(Stream *s, InstanceManager *m) : ConfigurableObject(s, m) {}
    virtual Float eval(const PhaseFunctionSamplingRecord &) const = 0; virtual Float sample(PhaseFunctionSamplingRecord &, Sampler *) const = 0;
    virtual Float sample(PhaseFunctionSamplingRecord &, Float &pdf, Sampler *) const = 0; virtual Float pdf(const PhaseFunctionSamplingRecord &) const;
    virtual Float getMeanCosine() const; enum { EAngleDependence = 2 }; MTS_DECLARE_CLASS() protected: unsigned m_type; };
class VolumeDataSource : public ConfigurableObject { public: /* volume.h:31-107 + the fork's value / gradient / insideVolumeLimits */
    VolumeDataSource(const Properties &p) : ConfigurableObject(p) {} VolumeDataSource(Stream *s, InstanceManager *m) : ConfigurableObject(s, m) {}
    inline const AABB &getAABB() const { return m_aabb; } virtual bool supportsFloatLookups() const; virtual Float lookupFloat(const Point &) const;
    virtual bool supportsSpectrumLookups() const; virtual Spectrum lookupSpectrum(const Point &) const; /* volume.h */
    virtual Float getStepSize() const = 0; virtual Float getMaximumFloatValue() const = 0; MTS_DECLARE_CLASS() protected: AABB m_aabb; };
class BSDF : public ConfigurableObject { public: BSDF(const Properties &p) : ConfigurableObject(p) {} MTS_DECLARE_CLASS() };
class Emitter; class TriMesh;
class Shape : public ConfigurableObject { public: Shape(const Properties &p) : ConfigurableObject(p) {} virtual AABB getAABB() const = 0; const BSDF *getBSDF() const;
    bool isEmitter() const; const Emitter *getEmitter() const; virtual ref<TriMesh> createTriMesh(); MTS_DECLARE_CLASS() };
class TriMesh : public Shape { public: TriMesh(const Properties &p) : Shape(p) {} size_t getVertexCount() const; const Point *getVertexPositions() const; AABB getAABB() const; };
class Emitter : public ConfigurableObject { public: Emitter(const Properties &p) : ConfigurableObject(p) {} bool isEnvironmentEmitter() const; const Shape *getShape() const;
    virtual Spectrum evalEnvironment(const RayDifferential &) const; MTS_DECLARE_CLASS() };
class Medium; /* medium.h:33-109 + the fork's additions */
struct MediumSamplingRecord { Float t, opticalLength; Point p; Float time; Spectrum transmittance, sigmaA, sigmaS; Float pdfSuccess, pdfSuccessRev, pdfFailure; const Medium *medium;
    Vector d, drev; Float refRatioSq, distance; };
class Medium : public ConfigurableObject { public: Medium(const Properties &p) : ConfigurableObject(p) {} Medium(Stream *s, InstanceManager *m) : ConfigurableObject(s, m) {}
    virtual bool sampleDistance(const Ray &, MediumSamplingRecord &, Sampler *) const = 0; virtual Spectrum evalTransmittance(const Ray &, Sampler *) const = 0;
    virtual void eval(const Ray &, MediumSamplingRecord &) const = 0; virtual bool isHomogeneous() const = 0; virtual bool isheterogeneousrefractive() const { return false; }
    const Shape *getShape() const { return m_shape; } MTS_DECLARE_CLASS() protected: Spectrum m_sigmaA, m_sigmaS, m_sigmaT; ref<PhaseFunction> m_phaseFunction; Shape *m_shape; };
class ReconstructionFilter : public ConfigurableObject { public: ReconstructionFilter(const Properties &p) : ConfigurableObject(p) {} Float getRadius() const; };
class Bitmap : public Object { public: enum EPixelFormat { ESpectrumAlphaWeight, EMultiSpectrumAlphaWeight }; enum EComponentFormat { EFloat32 };
    Bitmap(EPixelFormat, EComponentFormat, const Vector2i &size, int channels = -1); float *getFloat32Data(); };
class Film : public ConfigurableObject { public: Film(const Properties &p) : ConfigurableObject(p) {} enum EDecompositionType { ESteadyState, ETransient, EBounce };
    const Vector2i &getCropSize() const; virtual void setBitmap(const Bitmap *, Float multiplier = 1.0f) = 0; EDecompositionType getDecompositionType() const;
    Float getDecompositionMinBound() const; Float getDecompositionBinWidth() const; size_t getFrames() const; bool isCalibratedTransient() const; const ReconstructionFilter *getReconstructionFilter() const; };
struct AnimatedTransform { const Transform &eval(Float t) const; };
class Sensor : public ConfigurableObject { public: Sensor(const Properties &p) : ConfigurableObject(p) {} const Film *getFilm() const; const AnimatedTransform *getWorldTransform() const; MTS_DECLARE_CLASS() };
class PerspectiveCamera : public Sensor { public: PerspectiveCamera(const Properties &p) : Sensor(p) {} Float getXFov() const; MTS_DECLARE_CLASS() };
class Scene : public Object { public: const Sensor *getSensor() const; const Sampler *getSampler() const; const ref_vector<Medium> &getMedia() const; const ref_vector<Emitter> &getEmitters() const;
    const ref_vector<Shape> &getShapes() const; bool hasEnvironmentEmitter() const; const Emitter *getEnvironmentEmitter() const; };
class RenderQueue; class RenderJob;
class Integrator : public ConfigurableObject { public: Integrator(const Properties &p) : ConfigurableObject(p) {} Integrator(Stream *s, InstanceManager *m) : ConfigurableObject(s, m) {}
    virtual bool render(Scene *, RenderQueue *, const RenderJob *, int sceneResID, int sensorResID, int samplerResID) = 0; virtual void cancel() = 0; MTS_DECLARE_CLASS() };
MTS_NAMESPACE_END
