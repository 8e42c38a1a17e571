#pragma once
#include <mitsuba/core/platform.h>
