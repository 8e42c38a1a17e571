#include <mitsuba/stub.h> /* see stub.h: declarations only, for the syntax check of the binding */
