/* oracle shim — Properties reduced to the one getter src/phase/hg.cpp uses. TEST INFRASTRUCTURE ONLY. */
#pragma once
#include <map>
#include <string>
namespace mitsuba {
class Properties {
public:
    std::map<std::string, Float> floats;
    Float getFloat(const std::string &name, const Float &defVal) const {
        std::map<std::string, Float>::const_iterator it = floats.find(name);
        return it == floats.end() ? defVal : it->second;
    }
};
}
