// The adapter against the reference's OWN headers and scene sources (include/render/render.h, buffer.h, common.h,
// include/scene/scene.h, camera.h, include/settings/settings.h, src/scene/scene.cpp, camera.cpp -- compiled where they
// lie under $STRELKA_REF_DIR, see adapter/Makefile target `ref`; only glm is substituted, adapter/shim_glm).
// The scene is built through the REAL oka::Scene API (createMesh / createInstance / createLight / createCurve /
// addMaterial), replaying the call journal a strelka_b200.Scene recorded, so that
//   dump   : what oka::Scene + B200Render::buildSceneView flatten it to can be compared with the Python mirror
//   render : the image through RenderFactory-style use of B200Render can be compared with the oracle (needs a GPU)
// Usage: adapter_ref_test dump|render|sharded <journal.bin> <out.bin>
#include "B200Render.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace oka;

namespace
{
struct Reader
{
    FILE* f;
    bool ok = true;
    template <class T>
    T get()
    {
        T v{};
        ok = ok && std::fread(&v, sizeof(T), 1, f) == 1;
        return v;
    }
    template <class T>
    std::vector<T> vec()
    {
        const uint64_t n = get<uint64_t>();
        std::vector<T> v(ok ? n : 0);
        ok = ok && (n == 0 || std::fread(v.data(), sizeof(T), n, f) == n);
        return v;
    }
};
template <class T>
void put(FILE* f, const T* p, uint64_t n)
{
    std::fwrite(&n, 8, 1, f);
    if (n)
        std::fwrite(p, sizeof(T), n, f);
}
} // namespace

int main(int argc, char** argv)
{
    if (argc < 4)
        return 2;
    const bool sharded = std::strcmp(argv[1], "sharded") == 0; // render through joinGroup / renderSharded (a group of one)
    const bool doRender = std::strcmp(argv[1], "render") == 0 || sharded;
    Reader r{ std::fopen(argv[2], "rb") };
    if (!r.f)
        return 2;
    const uint32_t width = r.get<uint32_t>(), height = r.get<uint32_t>(), sppTotal = r.get<uint32_t>(), depth = r.get<uint32_t>();
    float cam[8]; // position xyz, orientation wxyz, fov
    for (float& c : cam)
        c = r.get<float>();
    Scene scene;
    // journal: op code, payload ... until op 0
    for (;;)
    {
        const uint32_t op = r.get<uint32_t>();
        if (!r.ok || op == 0)
            break;
        if (op == 1) // createMesh
        {
            const std::vector<Scene::Vertex> vb = r.vec<Scene::Vertex>();
            const std::vector<uint32_t> ib = r.vec<uint32_t>();
            scene.createMesh(vb, ib);
        }
        else if (op == 2) // createInstance (mesh / curve; light instances come from createLight)
        {
            glm::mat4 t;
            for (int i = 0; i < 16; ++i)
                glm::value_ptr(t)[i] = r.get<float>();
            const uint32_t type = r.get<uint32_t>(), geom = r.get<uint32_t>(), mat = r.get<uint32_t>();
            scene.createInstance(type == 2 ? Instance::Type::eCurve : Instance::Type::eMesh, geom, mat, t);
        }
        else if (op == 3) // createLight
        {
            Scene::UniformLightDesc d{};
            d.type = r.get<int32_t>();
            for (int i = 0; i < 16; ++i)
                glm::value_ptr(d.xform)[i] = r.get<float>();
            d.useXform = true;
            d.color = glm::float3(0.0f);
            d.color.x = r.get<float>();
            d.color.y = r.get<float>();
            d.color.z = r.get<float>();
            d.intensity = r.get<float>();
            d.width = r.get<float>();
            d.height = r.get<float>();
            d.radius = r.get<float>();
            d.halfAngle = r.get<float>();
            scene.createLight(d);
        }
        else if (op == 4) // createCurve
        {
            const std::vector<uint32_t> counts = r.vec<uint32_t>();
            const std::vector<float> pts = r.vec<float>();
            const std::vector<float> widths = r.vec<float>();
            std::vector<glm::float3> p(pts.size() / 3);
            for (size_t i = 0; i < p.size(); ++i)
                p[i] = glm::float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
            scene.createCurve(Curve::Type::eCubic, counts, p, widths);
        }
        else if (op == 5) // addMaterial
        {
            const uint32_t isMaterialX = r.get<uint32_t>();
            float color[3] = { r.get<float>(), r.get<float>(), r.get<float>() };
            const float roughness = r.get<float>(), metallic = r.get<float>(), ior = r.get<float>(), clearcoat = r.get<float>(),
                        clearcoatRoughness = r.get<float>();
            Scene::MaterialDescription d;
            d.type = isMaterialX ? Scene::MaterialDescription::Type::eMaterialX : Scene::MaterialDescription::Type::eMdl;
            d.file = "default.mdl";
            d.name = "default_material";
            MaterialManager::Param p;
            p.type = MaterialManager::Param::Type::eFloat3;
            p.name = isMaterialX ? "diffuseColor" : "diffuse_color";
            p.value.resize(12);
            std::memcpy(p.value.data(), color, 12);
            d.params.push_back(p);
            if (isMaterialX)
            {
                MaterialManager::Param q;
                q.type = MaterialManager::Param::Type::eFloat;
                q.name = "roughness";
                q.value.resize(4);
                std::memcpy(q.value.data(), &roughness, 4);
                d.params.push_back(q);
                const char* names[] = { "metallic", "ior", "clearcoat", "clearcoatRoughness" }; // UsdPreviewSurface inputs
                const float values[] = { metallic, ior, clearcoat, clearcoatRoughness };
                for (int k = 0; k < 4; ++k)
                {
                    q.name = names[k];
                    std::memcpy(q.value.data(), &values[k], 4);
                    d.params.push_back(q);
                }
            }
            scene.addMaterial(d);
        }
        else
        {
            return 3;
        }
    }
    std::fclose(r.f);
    if (!r.ok)
        return 2;
    {
        Camera fresh; // the delegate adds the render camera (HdStrelka/Camera.cpp:65-106); a new Scene has none
        scene.addCamera(fresh);
    }
    Camera& c = scene.getCamera(0);
    c.type = Camera::CameraType::firstperson;
    c.position = glm::float3(cam[0], cam[1], cam[2]);
    c.mOrientation = glm::quat{ cam[3], cam[4], cam[5], cam[6] };
    c.fov = cam[7];

    B200Render render;
    render.setScene(&scene);
    FILE* o = std::fopen(argv[3], "wb");
    if (!o)
        return 2;
    if (!doRender)
    {
        // no device needed: what the real oka::Scene and the adapter flatten the journal to
        B200Render::SceneViewStorage st;
        const sb_scene_view v = render.buildSceneView(st);
        put(o, v.vertices, v.num_vertices);
        put(o, v.indices, v.num_indices);
        put(o, v.meshes, v.num_meshes);
        put(o, v.instances, v.num_instances);
        put(o, v.lights, v.num_lights);
        put(o, v.materials, v.num_materials);
        put(o, v.curves, v.num_curves);
        put(o, v.curve_points, v.num_curve_points * 3);
        put(o, v.curve_widths, v.num_curve_widths);
        put(o, v.curve_vertex_counts, v.num_curve_vertex_counts);
        c.updateAspectRatio(width / float(height));
        c.updateViewMatrix();
        put(o, glm::value_ptr(c.matrices.view), 16);
        put(o, glm::value_ptr(c.matrices.perspective), 16);
        std::fclose(o);
        return 0;
    }
    SettingsManager settings; // the keys OptiXRender::render reads; defaults of src/hdRunner/main.cpp:510-542
    settings.setAs<uint32_t>("render/pt/depth", depth);
    settings.setAs<uint32_t>("render/pt/sppTotal", sppTotal);
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
    render.setSharedContext(&ctx);
    render.init();
    Buffer* out = render.createBuffer(BufferDesc{ width, height, BufferFormat::FLOAT4 });
    if (sharded)
    {
        char id[SB_COMM_ID_BYTES];
        if (!B200Render::groupId(id) || !render.joinGroup(id, 0, 1))
            return 4;
    }
    for (uint32_t i = 0; i < sppTotal + 2; ++i) // two extra frames: nothing left to render, the image stays
    {
        if (sharded)
            render.renderSharded(out, 1);
        else
            render.render(out);
    }
    out->map();
    std::fwrite(out->getHostPointer(), 1, out->getHostDataSize(), o);
    std::fclose(o);
    std::printf("subframe=%zu frames=%zu bytes=%zu\n", size_t(ctx.mSubframeIndex), size_t(ctx.mFrameNumber), out->getHostDataSize());
    const bool good = ctx.mSubframeIndex == sppTotal;
    delete out;
    return good ? 0 : 1;
}
