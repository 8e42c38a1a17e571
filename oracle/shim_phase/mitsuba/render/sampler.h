/* oracle shim — Sampler reduced to the calls on the path (include/mitsuba/render/sampler.h:124-127). TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <mitsuba/mitsuba.h>
namespace mitsuba {
class Sampler {
public:
    virtual ~Sampler() {}
    virtual Float next1D() = 0;
    virtual Point2 next2D() = 0;
};
}
