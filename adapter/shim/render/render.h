// STAND-IN for include/render/render.h:9-63 (same interface).
#pragma once
#include "buffer.h"
#include "common.h"
#include <scene/scene.h>
namespace oka
{
enum class RenderType : int { eOptiX = 0, eMetal, eCompute };
class Render
{
public:
    virtual ~Render() = default;
    virtual void init() = 0;
    virtual void render(Buffer* output) = 0;
    virtual Buffer* createBuffer(const BufferDesc& desc) = 0;
    virtual void* getNativeDevicePtr() { return nullptr; }
    void setSharedContext(SharedContext* ctx) { mSharedCtx = ctx; }
    SharedContext& getSharedContext() { return *mSharedCtx; }
    void setScene(Scene* scene) { mScene = scene; }
    Scene* getScene() { return mScene; }
protected:
    SharedContext* mSharedCtx = nullptr;
    oka::Scene* mScene = nullptr;
};
class RenderFactory
{
public:
    static Render* createRender(RenderType type);
    static Render* createRender();
};
} // namespace oka
