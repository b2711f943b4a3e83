// STAND-IN for include/settings/settings.h:11-101 (string map with typed accessors).
#pragma once
#include <cassert>
#include <cstdint>
#include <iostream>
#include <string>
#include <unordered_map>
namespace oka
{
class SettingsManager
{
    std::unordered_map<std::string, std::string> mMap;
public:
    template <typename T> void setAs(const char* name, const T& value) { mMap[name] = std::to_string(value); }
    void setAs(const char* name, const std::string& value) { mMap[name] = value; }
    template <typename T> T getAs(const char* name)
    {
        auto it = mMap.find(name);
        if (it == mMap.end())
        {
            std::cerr << "The setting " << name << " does not exist" << std::endl;
            assert(0);
            return T{};
        }
        return convert<T>(it->second);
    }
private:
    template <typename T> static T convert(const std::string& v);
};
template <> inline bool SettingsManager::convert<bool>(const std::string& v) { return v != "0" && v != "false" && !v.empty(); }
template <> inline uint32_t SettingsManager::convert<uint32_t>(const std::string& v) { return uint32_t(std::stoul(v)); }
template <> inline float SettingsManager::convert<float>(const std::string& v) { return std::stof(v); }
template <> inline std::string SettingsManager::convert<std::string>(const std::string& v) { return v; }
} // namespace oka
