// STAND-IN for include/render/common.h:10-35.
#pragma once
#include <cstddef>
#include <cstdint>
#include <settings/settings.h>
namespace oka
{
static constexpr int MAX_FRAMES_IN_FLIGHT = 3;
class Render;
struct SharedContext
{
    size_t mFrameNumber = 0;
    size_t mSubframeIndex = 0;
    SettingsManager* mSettingsManager = nullptr;
    Render* mRender = nullptr;
};
enum class Result : uint32_t { eOk, eFail, eOutOfMemory };
} // namespace oka
