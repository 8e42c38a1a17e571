/* oracle shim — Properties reduced to the getters src/phase/hg.cpp and the constructor of
 * src/medium/heterogeneousrefractive.cpp use. TEST INFRASTRUCTURE ONLY. */
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
    Float getFloat(const std::string &name) const { return floats.at(name); }
    int getInteger(const std::string &name, const int &defVal) const {
        std::map<std::string, Float>::const_iterator it = floats.find(name);
        return it == floats.end() ? defVal : (int) it->second;
    }
    bool getBoolean(const std::string &name, const bool &defVal) const {
        std::map<std::string, Float>::const_iterator it = floats.find(name);
        return it == floats.end() ? defVal : it->second != 0;
    }
};
}
