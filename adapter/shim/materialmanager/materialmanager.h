// STAND-IN for include/materialmanager/materialmanager.h:33-48 (only MaterialManager::Param).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace oka
{
class MaterialManager
{
public:
    struct Param
    {
        enum class Type : uint32_t { eFloat = 0, eInt, eBool, eFloat2, eFloat3, eFloat4, eTexture };
        Type type;
        std::string name;
        std::vector<uint8_t> value;
    };
};
} // namespace oka
