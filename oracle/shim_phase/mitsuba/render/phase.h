/* oracle shim — PhaseFunction / PhaseFunctionSamplingRecord reduced to the members src/phase/hg.cpp touches
 * (include/mitsuba/render/phase.h:33-186). TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/properties.h>
namespace mitsuba {
struct PhaseFunctionSamplingRecord {
    Vector wi, wo;
};
class PhaseFunction {
public:
    enum EPhaseFunctionType { EIsotropic = 1, EAngleDependence = 2, EAnisotropic = 4, ENonSymmetric = 8 };
    PhaseFunction(const Properties &) : m_type(0) {}
    PhaseFunction(Stream *, InstanceManager *) : m_type(0) {}
    virtual ~PhaseFunction() {}
    virtual void serialize(Stream *, InstanceManager *) const {}
    virtual void configure() {}
    virtual Float getMeanCosine() const { return 0; }
    virtual std::string toString() const { return ""; }
protected:
    unsigned int m_type;
};
}
