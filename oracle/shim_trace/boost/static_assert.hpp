/* oracle shim — the one Boost header include/mitsuba/core/matrix.h asks for. TEST INFRASTRUCTURE ONLY. */
#pragma once
#define BOOST_STATIC_ASSERT(x) static_assert(x, "")
