/* oracle shim — include/mitsuba/core/stream.h reduced to what vector.h / point.h / frame.h mention. TEST INFRASTRUCTURE ONLY. */
#pragma once
namespace mitsuba {
class Stream {
public:
    template <typename T> T readElement() { return T(); }
    template <typename T> void writeElement(T) {}
    float readFloat() { return 0.0f; }
    void writeFloat(float) {}
};
class InstanceManager {};
}
