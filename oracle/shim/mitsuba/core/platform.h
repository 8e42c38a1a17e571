/* oracle shim — stands in for include/mitsuba/core/platform.h so that the reference's
 * header-only basisspline.h compiles by itself.  TEST INFRASTRUCTURE ONLY.
 * FLOAT follows include/mitsuba/core/fwd.h:174-184 (double under -DFLOATDEBUG). */
#pragma once
#include <cmath>
#include <cstdlib>
#include <algorithm>
#ifndef MER_REF_FLOAT
#define MER_REF_FLOAT float
#endif
#define MTS_NAMESPACE_BEGIN namespace mitsuba {
#define MTS_NAMESPACE_END }
typedef unsigned int uint;
namespace mitsuba {
typedef MER_REF_FLOAT FLOAT;
struct VectorF {
    FLOAT x, y, z;
    VectorF() : x(0), y(0), z(0) {}
    VectorF(FLOAT v) : x(v), y(v), z(v) {}
    VectorF(FLOAT a, FLOAT b, FLOAT c) : x(a), y(b), z(c) {}
};
struct Matrix3x3F {
    FLOAT m[3][3];
    Matrix3x3F() {}
    Matrix3x3F(FLOAT a00, FLOAT a01, FLOAT a02, FLOAT a10, FLOAT a11, FLOAT a12, FLOAT a20, FLOAT a21,
               FLOAT a22) {
        m[0][0] = a00; m[0][1] = a01; m[0][2] = a02;
        m[1][0] = a10; m[1][1] = a11; m[1][2] = a12;
        m[2][0] = a20; m[2][1] = a21; m[2][2] = a22;
    }
};
}
