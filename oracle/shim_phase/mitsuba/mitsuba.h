/* oracle shim — stands in for include/mitsuba/mitsuba.h (which pulls in Boost) so that the reference's
 *   include/mitsuba/core/{platform,constants,fwd,math,vector,point,normal,frame}.h, src/phase/hg.cpp
 * and two functions of src/libcore/util.cpp compile VERBATIM from /root/reference (oracle/ref_phase.cpp).
 * Only the framework plumbing those files mention is stubbed here: logging macros and the class macros.
 * TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <mitsuba/core/platform.h> /* the reference's own */
#include <sstream>
#include <string>
#include <map>
#include <iostream>
#include <vector>
#include <cmath>
#include <algorithm>
#include <limits.h>
#include <stdio.h>
#include <string.h>
#include <stdexcept>
#include <limits>
#include <mitsuba/core/constants.h> /* reference */
#include <mitsuba/core/fwd.h>       /* reference */
#include <mitsuba/render/fwd.h>     /* reference */
#include <mitsuba/core/math.h>      /* reference */

/* logger.h: Log(EError) throws std::runtime_error (src/libcore/logger.cpp:100-147) */
namespace mitsuba {
enum ELogLevel { ETrace = 0, EDebug = 100, EInfo = 200, EWarn = 300, EError = 400 };
inline void shim_log(ELogLevel level, const char *fmt, ...) {
    if (level >= EError) throw std::runtime_error(fmt);
}
}
#define SLog(level, ...) ::mitsuba::shim_log(level, __VA_ARGS__)
#define Log(level, ...) ::mitsuba::shim_log(level, __VA_ARGS__)
#define Assert(cond) ((void) 0)
#define SAssert(cond) ((void) 0)
#define MTS_DECLARE_CLASS()
#define MTS_IMPLEMENT_CLASS_S(name, abstract, super)
#define MTS_EXPORT_PLUGIN(name, descr)

#include <mitsuba/core/vector.h> /* reference (its stream.h is the stub next to this file) */
#include <mitsuba/core/point.h>  /* reference */
#include <mitsuba/core/normal.h> /* reference */

/* util.h declarations of the two functions compiled from src/libcore/util.cpp */
namespace mitsuba {
extern void coordinateSystem(const Vector &a, Vector &b, Vector &c);
extern Float fresnelDielectricExt(Float cosThetaI, Float &cosThetaT, Float eta);
}
