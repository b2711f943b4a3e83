// See B200Render.h.  Mirrors the host flow of OptiXRender (src/render/optix/OptixRender.cpp):
//   init()   <- OptiXRender::init (:1059-1105): creates the device context, injects material 0
//   render() <- OptiXRender::render (:874-1057): frame-0 uploads + accel build, camera, settings, launch
// Errors: the reference assert(0)s; here they are logged to stderr and the frame is skipped.
#include "B200Render.h"

#include <cstdio>
#include <cstring>

namespace oka
{

static_assert(sizeof(Scene::Vertex) == sizeof(sb_vertex), "Scene::Vertex layout (scene.h:80-89)");
static_assert(sizeof(Mesh) == sizeof(sb_mesh), "oka::Mesh layout (scene.h:21-27)");
static_assert(sizeof(Curve) == sizeof(sb_curve), "oka::Curve layout (scene.h:29-42)");
static_assert(sizeof(Scene::Light) == sizeof(sb_light), "Scene::Light layout (scene.h:146-155)");

// ---- B200Buffer ------------------------------------------------------------------------------------
static uint32_t toSbFormat(BufferFormat f)
{
    switch (f)
    {
    case BufferFormat::UNSIGNED_BYTE4: return SB_FORMAT_UNSIGNED_BYTE4;
    case BufferFormat::FLOAT3: return SB_FORMAT_FLOAT3;
    default: return SB_FORMAT_FLOAT4;
    }
}

B200Buffer::B200Buffer(sb_ctx* ctx, const BufferDesc& desc)
{
    mWidth = desc.width;
    mHeight = desc.height;
    mFormat = desc.format;
    if (sb_buffer_create(ctx, desc.width, desc.height, toSbFormat(desc.format), &mHandle) != SB_OK)
        std::fprintf(stderr, "B200Buffer: %s\n", sb_last_error(ctx));
}
B200Buffer::~B200Buffer()
{
    sb_buffer_destroy(mHandle);
}
void B200Buffer::resize(uint32_t width, uint32_t height)
{
    if (sb_buffer_resize(mHandle, width, height) == SB_OK)
    {
        mWidth = width;
        mHeight = height;
    }
}
void* B200Buffer::map()
{
    sb_buffer_map(mHandle, nullptr);
    return nullptr; // OptixBuffer::map returns nullptr too (OptixBuffer.cpp:37-43); callers use getHostPointer()
}
void B200Buffer::unmap()
{
    sb_buffer_unmap(mHandle);
}
void* B200Buffer::getHostPointer()
{
    return sb_buffer_host_ptr(mHandle);
}
size_t B200Buffer::getHostDataSize()
{
    return sb_buffer_host_size(mHandle);
}
void* B200Buffer::getNativePtr()
{
    return sb_buffer_device_ptr(mHandle);
}

// ---- B200Render ------------------------------------------------------------------------------------
B200Render::~B200Render()
{
    sb_destroy(mCtx);
}

void B200Render::fail(const char* what)
{
    mError = std::string(what) + ": " + sb_last_error(mCtx);
    std::fprintf(stderr, "B200Render: %s\n", mError.c_str());
}

void B200Render::init()
{
    sb_device_cfg cfg = {};
    cfg.device = mDevice; // the reference uses the current device (cudaFree(0), OptixRender.cpp:166)
    if (sb_create(&cfg, &mCtx) != SB_OK)
    {
        fail("sb_create");
        return;
    }
    // material slot 0 = default.mdl::default_material, added by the backend before the delegate adds its
    // own (OptixRender.cpp:1091-1097) so that scene material indices keep their meaning
    if (mScene && mScene->getMaterials().empty())
    {
        Scene::MaterialDescription def;
        def.type = Scene::MaterialDescription::Type::eMdl;
        def.file = "default.mdl";
        def.name = "default_material";
        mScene->addMaterial(def);
    }
}

Buffer* B200Render::createBuffer(const BufferDesc& desc)
{
    return new B200Buffer(mCtx, desc); // raw owning pointer, deleted by the caller (RenderBuffer.cpp:122-126)
}

void* B200Render::getNativeDevicePtr()
{
    return nullptr;
}

static bool readFloats(const MaterialManager::Param& p, float* dst, size_t n)
{
    if (p.value.size() < n * sizeof(float))
        return false;
    std::memcpy(dst, p.value.data(), n * sizeof(float));
    return true;
}

sb_material B200Render::resolveMaterial(const Scene::MaterialDescription& desc)
{
    sb_material m;
    std::memset(&m, 0, sizeof(m));
    // eMaterialX descriptions carry a single ND_UsdPreviewSurface_surfaceshader (Material.cpp:173-177);
    // eMdl + default.mdl is the diffuse default material (RenderPass.cpp:222-245)
    m.model = (desc.type == Scene::MaterialDescription::Type::eMaterialX) ? SB_MATERIAL_USD_PREVIEW_SURFACE : SB_MATERIAL_DIFFUSE;
    m.base_color[0] = m.base_color[1] = m.base_color[2] = (m.model == SB_MATERIAL_DIFFUSE) ? 1.0f : 0.18f;
    m.roughness = 0.5f; // UsdPreviewSurface defaults (tests/materialmanager/test_materialmanager.cpp:31-46)
    m.metallic = 0.0f;
    m.ior = 1.5f;
    m.opacity = 1.0f;
    m.clearcoat = 0.0f;
    m.clearcoat_roughness = 0.01f;
    if (desc.hasColor)
    {
        m.base_color[0] = desc.color.x;
        m.base_color[1] = desc.color.y;
        m.base_color[2] = desc.color.z;
    }
    for (const MaterialManager::Param& p : desc.params)
    {
        float v[4] = { 0, 0, 0, 0 };
        if (p.name == "diffuse_color" || p.name == "diffuseColor" || p.name == "diffuse_color_constant")
            readFloats(p, m.base_color, 3);
        else if (p.name == "roughness" && readFloats(p, v, 1))
            m.roughness = v[0];
        else if (p.name == "metallic" && readFloats(p, v, 1))
            m.metallic = v[0];
        else if (p.name == "ior" && readFloats(p, v, 1))
            m.ior = v[0];
        else if (p.name == "opacity" && readFloats(p, v, 1))
            m.opacity = v[0];
        else if (p.name == "clearcoat" && readFloats(p, v, 1))
            m.clearcoat = v[0];
        else if (p.name == "clearcoatRoughness" && readFloats(p, v, 1))
            m.clearcoat_roughness = v[0];
        else if (p.name == "specularColor")
            readFloats(p, m.specular_color, 3);
        else if (p.name == "useSpecularWorkflow" && p.value.size() >= 4)
        {
            int32_t i = 0;
            std::memcpy(&i, p.value.data(), 4);
            m.use_specular_workflow = i != 0;
        }
    }
    return m;
}

sb_scene_view B200Render::buildSceneView(SceneViewStorage& st)
{
    Scene& s = *mScene;
    std::vector<sb_instance>& inst = st.instances;
    inst.resize(s.getInstances().size());
    for (size_t i = 0; i < inst.size(); ++i)
    {
        const Instance& in = s.getInstances()[i];
        std::memcpy(inst[i].transform, glm::value_ptr(in.transform), sizeof(float) * 16); // glm storage order
        inst[i].type = in.type == Instance::Type::eMesh ? SB_INSTANCE_MESH : (in.type == Instance::Type::eLight ? SB_INSTANCE_LIGHT : SB_INSTANCE_CURVE);
        inst[i].geom_id = in.mMeshId;
        inst[i].material_id = in.mMaterialId;
        inst[i].light_id = in.mLightId;
    }
    std::vector<sb_material>& mats = st.materials;
    mats.clear();
    st.textures.clear();
    st.texturePixels.clear();
    for (const Scene::MaterialDescription& d : s.getMaterials())
    {
        sb_material m = resolveMaterial(d);
        // texture inputs: "<node>_file" params of type eTexture (Material.cpp:120-136; OptixRender.cpp:1346-1387)
        for (const MaterialManager::Param& p : d.params)
        {
            if (p.type != MaterialManager::Param::Type::eTexture || !mTextureLoader)
                continue;
            const std::string path(reinterpret_cast<const char*>(p.value.data()), p.value.size());
            std::vector<uint8_t> px;
            uint32_t w = 0, h = 0;
            if (!mTextureLoader(path, px, w, h) || px.size() != size_t(w) * h * 4 || w == 0 || h == 0)
            {
                std::fprintf(stderr, "B200Render: unable to load texture %s\n", path.c_str());
                continue;
            }
            st.texturePixels.push_back(std::move(px));
            const uint32_t index1 = uint32_t(st.texturePixels.size());
            // which input the texture feeds is encoded in the node name the delegate prefixes (e.g. diffuseColor_texture_file)
            if (p.name.find("ormal") != std::string::npos)
                m.normal_texture = index1;
            else
                m.diffuse_texture = index1;
            sb_texture t;
            t.pixels = nullptr; // patched below: the vector may still move
            t.width = w;
            t.height = h;
            st.textures.push_back(t);
        }
        mats.push_back(m);
    }
    for (size_t i = 0; i < st.textures.size(); ++i)
        st.textures[i].pixels = st.texturePixels[i].data();
    sb_scene_view v;
    std::memset(&v, 0, sizeof(v));
    v.vertices = reinterpret_cast<const sb_vertex*>(s.getVertices().data());
    v.num_vertices = s.getVertices().size();
    v.indices = s.getIndices().data();
    v.num_indices = s.getIndices().size();
    v.meshes = reinterpret_cast<const sb_mesh*>(s.getMeshes().data());
    v.num_meshes = uint32_t(s.getMeshes().size());
    v.curves = reinterpret_cast<const sb_curve*>(s.getCurves().data());
    v.num_curves = uint32_t(s.getCurves().size());
    v.curve_points = reinterpret_cast<const float*>(s.getCurvesPoint().data());
    v.num_curve_points = s.getCurvesPoint().size();
    v.curve_widths = s.getCurvesWidths().data();
    v.num_curve_widths = s.getCurvesWidths().size();
    v.curve_vertex_counts = s.getCurvesVertexCounts().data();
    v.num_curve_vertex_counts = s.getCurvesVertexCounts().size();
    v.instances = inst.data();
    v.num_instances = uint32_t(inst.size());
    v.lights = reinterpret_cast<const sb_light*>(s.getLights().data());
    v.num_lights = uint32_t(s.getLights().size());
    v.materials = mats.data();
    v.num_materials = uint32_t(mats.size());
    v.textures = st.textures.empty() ? nullptr : st.textures.data();
    v.num_textures = uint32_t(st.textures.size());
    return v;
}

void B200Render::uploadScene()
{
    SceneViewStorage storage;
    const sb_scene_view v = buildSceneView(storage);
    if (sb_set_scene(mCtx, &v) != SB_OK)
        fail("sb_set_scene");
    else
        mSceneUploaded = true;
}

sb_settings B200Render::readSettings()
{
    SettingsManager& st = *getSharedContext().mSettingsManager;
    sb_settings s;
    sb_settings_default(&s);
    // the keys OptiXRender::render reads (OptixRender.cpp:910-1004)
    s.spp = st.getAs<uint32_t>("render/pt/spp");
    s.spp_total = st.getAs<uint32_t>("render/pt/sppTotal");
    s.depth = st.getAs<uint32_t>("render/pt/depth");
    s.enable_acc = st.getAs<bool>("render/pt/enableAcc") ? 1u : 0u;
    s.rect_light_sampling_method = st.getAs<uint32_t>("render/pt/rectLightSamplingMethod");
    s.debug = st.getAs<uint32_t>("render/pt/debug");
    s.shadow_ray_tmin = st.getAs<float>("render/pt/dev/shadowRayTmin");
    s.material_ray_tmin = st.getAs<float>("render/pt/dev/materialRayTmin");
    s.tonemapper_type = st.getAs<uint32_t>("render/pt/tonemapperType");
    s.gamma = st.getAs<float>("render/post/gamma");
    s.film_iso = st.getAs<float>("render/post/tonemapper/filmIso");
    s.cm2_factor = st.getAs<float>("render/post/tonemapper/cm2_factor");
    s.f_stop = st.getAs<float>("render/post/tonemapper/fStop");
    s.shutter_speed = st.getAs<float>("render/post/tonemapper/shutterSpeed");
    return s;
}

// frame-0 uploads, camera and settings of OptiXRender::render (OptixRender.cpp:874-1004); false = skip the frame
bool B200Render::prepareFrame(Buffer* output, sb_settings& s)
{
    if (!mCtx || !mScene || !output)
        return false;
    // frame 0: uploads + acceleration structure (OptixRender.cpp:876-888)
    if (getSharedContext().mFrameNumber == 0 || !mSceneUploaded)
    {
        uploadScene();
        if (!mSceneUploaded)
            return false;
    }
    Camera& camera = mScene->getCamera(0);
    camera.updateAspectRatio(output->width() / float(output->height()));
    camera.updateViewMatrix();
    if (sb_set_camera(mCtx, glm::value_ptr(camera.matrices.view), camera.fov) != SB_OK)
    {
        fail("sb_set_camera");
        return false;
    }
    s = readSettings();
    s.sample_offset = mRank; // sample sharding: this rank's stride of the global sample indices
    s.sample_stride = mWorld;
    if (sb_set_settings(mCtx, &s) != SB_OK)
    {
        fail("sb_set_settings");
        return false;
    }
    return true;
}

void B200Render::render(Buffer* output)
{
    sb_settings s;
    if (!prepareFrame(output, s))
        return;
    if (sb_render(mCtx, static_cast<B200Buffer*>(output)->handle()) != SB_OK)
        return fail("sb_render");
    getSharedContext().mSubframeIndex = sb_subframe_index(mCtx);
    output->unmap();
    getSharedContext().mFrameNumber++;
}

bool B200Render::joinGroup(const void* id, uint32_t rank, uint32_t world)
{
    if (!mCtx || sb_comm_init(mCtx, id, rank, world) != SB_OK)
    {
        fail("sb_comm_init");
        return false;
    }
    mRank = rank;
    mWorld = world;
    return true;
}

void B200Render::renderSharded(Buffer* output, uint32_t iterationsPerRank)
{
    sb_settings s;
    if (!prepareFrame(output, s))
        return;
    if (sb_render_sharded(mCtx, static_cast<B200Buffer*>(output)->handle(), iterationsPerRank) != SB_OK)
        return fail("sb_render_sharded");
    getSharedContext().mSubframeIndex = sb_subframe_index(mCtx); // this rank's samples
    output->unmap();
    getSharedContext().mFrameNumber += iterationsPerRank;
}

} // namespace oka
