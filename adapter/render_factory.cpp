// What src/render/render.cpp:10-35 becomes with the third backend registered.  In this
// repository it backs the shim's RenderFactory so the adapter test can go through the reference's own
// entry point RenderFactory::createRender(RenderType::eCompute).
#include "B200Render.h"

using namespace oka;

Render* RenderFactory::createRender(const RenderType type)
{
    if (type == RenderType::eCompute)
    {
        return new B200Render();
    }
    // eOptiX / eMetal live in the reference tree; unsupported here -> nullptr like render.cpp:17-18,25
    return nullptr;
}

Render* RenderFactory::createRender()
{
    return new B200Render();
}
