// Drives the backend exactly the way a Strelka client does (src/app/main.cpp:329-396): factory -> setScene ->
// setSharedContext -> init -> createBuffer -> render loop -> map -> getHostPointer.  Writes the float4 image
// to a raw file so that the Python test can compare it with the oracle.  Usage: adapter_test <scene.bin> <out.raw>
#include "B200Render.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace oka;

template <class T>
static bool readVec(FILE* f, std::vector<T>& v)
{
    uint64_t n = 0;
    if (std::fread(&n, 8, 1, f) != 1)
        return false;
    v.resize(n);
    return n == 0 || std::fread(v.data(), sizeof(T), n, f) == n;
}

int main(int argc, char** argv)
{
    if (argc < 3)
        return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f)
        return 2;
    Scene scene;
    uint32_t hdr[4]; // width, height, sppTotal, depth
    float cam[8]; // position xyz, orientation wxyz, fov
    if (std::fread(hdr, 4, 4, f) != 4 || std::fread(cam, 4, 8, f) != 8)
        return 2;
    struct RawInst { float t[16]; uint32_t type, geom, mat, light; };
    struct RawMat { uint32_t isMaterialX; float color[3]; float roughness, metallic; };
    std::vector<RawInst> inst;
    std::vector<RawMat> mats;
    bool ok = readVec(f, scene.mVertices) && readVec(f, scene.mIndices) && readVec(f, scene.mMeshes) && readVec(f, inst) &&
              readVec(f, scene.mLights) && readVec(f, mats);
    std::fclose(f);
    if (!ok)
        return 2;
    for (const RawInst& r : inst)
    {
        Instance i;
        std::memcpy(&i.transform, r.t, 64);
        i.type = r.type == 0 ? Instance::Type::eMesh : (r.type == 1 ? Instance::Type::eLight : Instance::Type::eCurve);
        i.mMeshId = r.geom;
        i.mMaterialId = r.mat;
        i.mLightId = r.light;
        scene.mInstances.push_back(i);
    }
    for (const RawMat& r : mats)
    {
        Scene::MaterialDescription d;
        d.type = r.isMaterialX ? Scene::MaterialDescription::Type::eMaterialX : Scene::MaterialDescription::Type::eMdl;
        d.file = "default.mdl";
        MaterialManager::Param p;
        p.type = MaterialManager::Param::Type::eFloat3;
        p.name = r.isMaterialX ? "diffuseColor" : "diffuse_color";
        p.value.resize(12);
        std::memcpy(p.value.data(), r.color, 12);
        d.params.push_back(p);
        if (r.isMaterialX)
        {
            MaterialManager::Param q;
            q.type = MaterialManager::Param::Type::eFloat;
            q.name = "roughness";
            q.value.resize(4);
            std::memcpy(q.value.data(), &r.roughness, 4);
            d.params.push_back(q);
            q.name = "metallic";
            std::memcpy(q.value.data(), &r.metallic, 4);
            d.params.push_back(q);
        }
        scene.addMaterial(d);
    }
    Camera& c = scene.getCamera(0);
    c.position = glm::float3(cam[0], cam[1], cam[2]);
    c.mOrientation = glm::quat{ cam[3], cam[4], cam[5], cam[6] };
    c.fov = cam[7];

    SettingsManager settings; // defaults of src/hdRunner/main.cpp:510-542
    settings.setAs<uint32_t>("render/pt/depth", hdr[3]);
    settings.setAs<uint32_t>("render/pt/sppTotal", hdr[2]);
    settings.setAs<uint32_t>("render/pt/spp", 1);
    settings.setAs<uint32_t>("render/pt/tonemapperType", 0);
    settings.setAs<uint32_t>("render/pt/debug", 0);
    settings.setAs<bool>("render/pt/enableAcc", true);
    settings.setAs<uint32_t>("render/pt/rectLightSamplingMethod", 0);
    settings.setAs<float>("render/post/tonemapper/filmIso", 100.0f);
    settings.setAs<float>("render/post/tonemapper/cm2_factor", 1.0f);
    settings.setAs<float>("render/post/tonemapper/fStop", 4.0f);
    settings.setAs<float>("render/post/tonemapper/shutterSpeed", 100.0f);
    settings.setAs<float>("render/post/gamma", 0.0f);
    settings.setAs<float>("render/pt/dev/shadowRayTmin", 0.0f);
    settings.setAs<float>("render/pt/dev/materialRayTmin", 0.0f);
    SharedContext ctx;
    ctx.mSettingsManager = &settings;

    Render* render = RenderFactory::createRender(RenderType::eCompute);
    render->setScene(&scene);
    render->setSharedContext(&ctx);
    render->init();
    Buffer* out = render->createBuffer(BufferDesc{ hdr[0], hdr[1], BufferFormat::FLOAT4 });
    for (uint32_t i = 0; i < hdr[2] + 2; ++i) // two extra frames: nothing left to render, image stays
        render->render(out);
    out->map();
    FILE* o = std::fopen(argv[2], "wb");
    std::fwrite(out->getHostPointer(), 1, out->getHostDataSize(), o);
    std::fclose(o);
    std::printf("subframe=%zu frames=%zu bytes=%zu\n", ctx.mSubframeIndex, ctx.mFrameNumber, out->getHostDataSize());
    const bool good = ctx.mSubframeIndex == hdr[2];
    delete out;
    delete render;
    return good ? 0 : 1;
}
