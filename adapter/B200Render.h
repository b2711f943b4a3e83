// B200Render -- the third Strelka render backend (next to OptiXRender and MetalRender): a thin C++ host
// class implementing oka::Render (include/render/render.h:19-56) on top of the C ABI of
// include/sb/sb_api.h.  Inside a Strelka tree it is compiled against the real headers and registered
// under RenderType::eCompute; in this repository it is compiled against
// adapter/shim (minimal stand-ins) so that it can be built and tested without glm/OpenUSD/MDL.
#pragma once
#include <render/render.h>
#include <sb/sb_api.h>

#include <functional>
#include <string>
#include <vector>

namespace oka
{

class B200Buffer : public Buffer
{
public:
    B200Buffer(sb_ctx* ctx, const BufferDesc& desc);
    ~B200Buffer() override;
    void resize(uint32_t width, uint32_t height) override;
    void* map() override; // blocking D2H; like OptixBuffer::map it returns nullptr -- use getHostPointer()
    void unmap() override;
    void* getHostPointer() override;
    size_t getHostDataSize() override;
    void* getNativePtr();
    sb_buffer* handle() { return mHandle; }

private:
    sb_buffer* mHandle = nullptr;
};

class B200Render : public Render
{
public:
    B200Render() = default;
    ~B200Render() override;
    void init() override;
    void render(Buffer* output) override;
    Buffer* createBuffer(const BufferDesc& desc) override;
    void* getNativeDevicePtr() override;

    // pre-resolve an oka material description into one of the backend's closed-form models by
    // parameter name (precedent: MetalRender.cpp:84-89)
    static sb_material resolveMaterial(const Scene::MaterialDescription& desc);
    const std::string& lastError() const { return mError; }
    sb_ctx* context() { return mCtx; }

    // The flattened scene exactly as uploadScene() hands it to sb_set_scene: the oka::Scene arrays passed through, the
    // instances and materials converted.  `storage` owns the converted arrays and the decoded textures; the view
    // borrows from it and from the scene.  Public so that tools can inspect what a Strelka scene turns into.
    struct SceneViewStorage
    {
        std::vector<sb_instance> instances;
        std::vector<sb_material> materials;
        std::vector<sb_texture> textures;
        std::vector<std::vector<uint8_t>> texturePixels;
    };
    sb_scene_view buildSceneView(SceneViewStorage& storage);

    // UsdUVTexture inputs arrive as MaterialManager::Param::Type::eTexture params holding a file path (Material.cpp:
    // 120-136); OptiXRender decodes them with stb_image (OptixRender.cpp:1191-1200).  This adapter has no image
    // library of its own: the host hands it a decoder (path -> RGBA8, width, height).  Without one, or when decoding
    // fails, the material keeps its constant colour (the reference logs an error and binds an empty texture).
    // ---- multi-GPU (new: OptiXRender is single-GPU, OptixRender.cpp:163-189) ---------------------------------
    // One B200Render per GPU = one rank.  setDevice() before init(); rank 0 calls groupId() and hands the bytes to the
    // other ranks (MPI, socket ...); every rank then calls joinGroup() (collective) and renderSharded() instead of
    // render(): the rank renders sample indices rank, rank + world, ... of every pixel, the library sums the
    // accumulation buffers over NVLink (fused NVLS kernel or ncclAllReduce) and every rank's buffer receives the image.
    void setDevice(int device) { mDevice = device; }
    static bool groupId(char id[SB_COMM_ID_BYTES]) { return sb_comm_get_unique_id(id) == SB_OK; }
    bool joinGroup(const void* id, uint32_t rank, uint32_t world);
    void renderSharded(Buffer* output, uint32_t iterationsPerRank);
    const char* exchangePath() const { return sb_comm_exchange_path(mCtx); }

    using TextureLoader = std::function<bool(const std::string& path, std::vector<uint8_t>& rgba, uint32_t& width, uint32_t& height)>;
    void setTextureLoader(TextureLoader loader) { mTextureLoader = std::move(loader); }

private:
    void uploadScene();
    TextureLoader mTextureLoader;
    sb_settings readSettings();
    void fail(const char* what);

    bool prepareFrame(Buffer* output, sb_settings& s);
    sb_ctx* mCtx = nullptr;
    int mDevice = 0;
    uint32_t mRank = 0, mWorld = 1;
    bool mSceneUploaded = false;
    std::string mError;
};

} // namespace oka
