/* oracle shim — the three constants basisspline.h uses; values as in
 * include/mitsuba/core/constants.h:69-71 (double) and :99-101 (single). */
#pragma once
#include <mitsuba/core/platform.h>
#define VHALF ((mitsuba::FLOAT) 0.5)
#define ONE_SIXTH ((mitsuba::FLOAT) 0.16666666666666666666666666666666667)
#define TWO_THIRD ((mitsuba::FLOAT) 0.66666666666666666666666666666666667)
