/* oracle shim — include/mitsuba/core/stream.h reduced to what vector.h / point.h / frame.h mention. TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <cstddef>
namespace mitsuba {
class Stream {
public:
    enum EByteOrder { EBigEndian = 0, ELittleEndian = 1 }; /* stream.h: loadFromFile() of splinevolume.cpp names it */
    template <typename T> T readElement() { return T(); }
    template <typename T> void writeElement(T) {}
    float readFloat() { return 0.0f; }
    void writeFloat(float) {}
    template <typename T> void readArray(T *, size_t) {}        /* matrix.h (ref_trace.cpp) */
    template <typename T> void writeArray(const T *, size_t) {}
};
class InstanceManager {};
}
