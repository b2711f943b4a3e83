// Implementation of the C ABI (include/sb/sb_api.h): the host side of the B200 backend.
//
// Mirrors the host logic of the reference's OptiX backend, without OptiX:
//   sb_create            <- OptiXRender::init            (OptixRender.cpp:1059-1105)
//   sb_set_scene         <- frame-0 block of render()    (OptixRender.cpp:876-888): uploads, accel build
//   sb_render            <- OptiXRender::render(Buffer*) (OptixRender.cpp:874-1057)
//   sb_buffer_*          <- OptixBuffer                  (OptixBuffer.cpp)
// No exception crosses the ABI; every failure is reported as sb_result + sb_last_error().
#include "../../include/sb/sb_api.h"
#include "kernels.h"
#include "scene_prep.h"
#include "host_math.h"
#include "sb_nvls.cuh"

#include <chrono>
#include <cstdio>
#include <dlfcn.h>
#include <nccl.h> // types only: the library is dlopen()ed when a communicator is first asked for (nccl_api below)
#include <cstring>
#include <string>
#include <vector>

using namespace sb;

static thread_local std::string g_createError;

struct sb_buffer
{
    sb_ctx* ctx = nullptr;
    void* dev = nullptr;
    void* host = nullptr; // pinned mirror
    uint32_t width = 0, height = 0, format = SB_FORMAT_FLOAT4;
    size_t bytes = 0;
    cudaEvent_t copied = nullptr; // sb_buffer_map_async: the copy into `host` has completed
    bool copyPending = false;
};

struct sb_ctx
{
    int device = 0;
    int numSms = 148;
    cudaStream_t stream = nullptr; // the stream work is issued on (own or caller-provided)
    cudaStream_t ownStream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    StageTimer timer;
    bool stageTimers = false;
    uint64_t launchCount = 0;
    std::string error;
    bool trackStats = false;
    bool fusedSmall = false;
    uint32_t maxBatchPaths = 32u << 20;
    uint32_t curveSplit = 8;

    SceneDev scene;
    bool haveScene = false;
    std::vector<cudaArray_t> texArrays; // one uchar4 array + one filtered texture object per sb_texture
    std::vector<cudaTextureObject_t> texObjects;
    uint32_t* instTriFirst = nullptr;
    SegInfo* segInfoUnsorted = nullptr;

    // camera
    float view[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 }; // glm column-major
    float fovY = 45.0f;
    float viewToWorld[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    float clipToViewRaw[16];
    bool rawMatrices = false;

    sb_settings settings;
    bool haveSettings = false;
    uint32_t subframe = 0; // samples accumulated by THIS context (SharedContext::mSubframeIndex)

    // per-resolution state
    uint32_t width = 0, height = 0, tilesX = 0, nPixPadded = 0, batchPaths = 0;
    size_t queuePaths = 0; // paths the queues are currently allocated for (grows on demand up to batchPaths)
    float4* S = nullptr; // accumulation: sum of T(L)
    float4* direct = nullptr; // non-accumulated launch result
    float4* aovD = nullptr; // diffuse / specular AOVs (debug views 2 / 3): count * A in xyz, count in w
    float4* aovS = nullptr;
    float4* aovScrD = nullptr; // AOV running sums of a launch larger than one wavefront batch (allocated on demand)
    float4* aovScrS = nullptr;
    Queues Q = {};
    StatCounters* stats = nullptr;
    uint32_t* sobolTab = nullptr;

    double buildMs = 0.0, renderMs = 0.0;

    // multi-GPU sample sharding (sb_comm_*): one context = one rank
    ncclComm_t comm = nullptr;
    int commRank = 0, commWorld = 1;
    float4* Sglobal = nullptr; // all-reduced copy of S (S itself stays this rank's partial sum)
    uint64_t shardRequested = 0; // iterations asked of sb_render_sharded since the last accumulation reset
    cudaEvent_t evXchgStart = nullptr, evXchgStop = nullptr; // brackets the exchange (all-reduce + resolve) of the last sharded render
    // NVLS path (sb_nvls.cuh): S and Sglobal as NCCL symmetric windows, device communicator with multimem + barriers
    bool nvlsReady = false, symBuffers = false;
    std::string nvlsStatus = "not a multi-GPU context";
    mutable std::string exchangeNote;
    int nvlsGrid = 0;
    size_t symPix = 0;
#if SB_HAVE_NCCL_DEVICE
    ncclDevComm devComm;
    ncclWindow_t winS = nullptr, winG = nullptr;
#endif
};

namespace
{

template <class T>
void dev_free(T*& p)
{
    if (p)
        cudaFree(p);
    p = nullptr;
}

// scene arrays come from (and return to) the stream-ordered pool, like the builder's temporaries (exec.h)
template <class T>
void pool_free(T*& p, cudaStream_t st)
{
    if (p)
        cudaFreeAsync(p, st);
    p = nullptr;
}

template <class T>
T* dev_upload(const T* src, size_t n, cudaStream_t st)
{
    T* d = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (cudaMallocAsync(&d, bytes, st) != cudaSuccess)
    {
        cudaGetLastError();
        throw std::bad_alloc();
    }
    if (n)
        SB_CUDA_CHECK(cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return d;
}

void free_scene(sb_ctx* c)
{
    SceneDev& s = c->scene;
    cudaStream_t st = c->stream; // the callers have synchronised it: nothing still reads these arrays
    pool_free(s.vertices, st);
    pool_free(s.indices, st);
    pool_free(s.meshes, st);
    pool_free(s.curves, st);
    pool_free(s.curvePoints, st);
    pool_free(s.curveRadii, st);
    pool_free(s.curveVertexCounts, st);
    pool_free(s.lights, st);
    pool_free(s.materials, st);
    pool_free(s.instances, st);
    pool_free(s.tris, st);
    pool_free(s.triShade, st);
    pool_free(s.segs, st);
    pool_free(s.segInfo, st);
    pool_free(s.triNodes, st);
    pool_free(s.segNodes, st);
    pool_free(c->instTriFirst, st);
    pool_free(c->segInfoUnsorted, st);
    pool_free(s.triUv, st);
    {
        unsigned long long* t = const_cast<unsigned long long*>(s.textures);
        pool_free(t, st);
    }
    for (cudaTextureObject_t o : c->texObjects)
        cudaDestroyTextureObject(o);
    for (cudaArray_t a : c->texArrays)
        cudaFreeArray(a);
    c->texObjects.clear();
    c->texArrays.clear();
    s = SceneDev();
    c->haveScene = false;
}

void free_queues(sb_ctx* c)
{
    for (int i = 0; i < 2; ++i)
    {
        dev_free(c->Q.rayO[i]);
        dev_free(c->Q.rayD[i]);
        dev_free(c->Q.thr[i]);
    }
    dev_free(c->Q.hitA);
    dev_free(c->Q.hitB);
    dev_free(c->Q.Lacc);
    dev_free(c->Q.shO);
    dev_free(c->Q.shD);
    dev_free(c->Q.shC);
    c->queuePaths = 0;
}

void release_sym(sb_ctx* c, bool keepFrame); // below (needs the NCCL loader)

void free_frame(sb_ctx* c)
{
    release_sym(c, false);
    dev_free(c->S);
    dev_free(c->direct);
    dev_free(c->aovD);
    dev_free(c->aovS);
    dev_free(c->aovScrD);
    dev_free(c->aovScrS);
    dev_free(c->Sglobal);
    free_queues(c);
    dev_free(c->Q.counts);
    c->width = c->height = 0;
}

template <class T>
T* dev_alloc(size_t n)
{
    T* d = nullptr;
    if (cudaMalloc(&d, (n ? n : 1) * sizeof(T)) != cudaSuccess)
        throw std::bad_alloc();
    return d;
}

// updatePathtracerParams (OptixRender.cpp:827-872): (re)allocate per-resolution state
void ensure_frame(sb_ctx* c, uint32_t w, uint32_t h)
{
    if (c->width == w && c->height == h)
        return;
    free_frame(c);
    c->width = w;
    c->height = h;
    c->tilesX = (w + 7) / 8;
    const uint32_t tilesY = (h + 3) / 4;
    c->nPixPadded = c->tilesX * tilesY * 32u;
    const uint32_t chunkMax = std::max<uint32_t>(1u, c->maxBatchPaths / c->nPixPadded);
    c->batchPaths = c->nPixPadded * chunkMax;
    c->S = dev_alloc<float4>(size_t(w) * h);
    c->direct = dev_alloc<float4>(size_t(w) * h);
    c->aovD = dev_alloc<float4>(size_t(w) * h);
    c->aovS = dev_alloc<float4>(size_t(w) * h);
    c->Q.counts = dev_alloc<uint32_t>(kNumCounts);
    c->Q.stats = c->stats;
    c->Q.sobolTab = c->sobolTab;
    SB_CUDA_CHECK(cudaMemsetAsync(c->S, 0, sizeof(float4) * size_t(w) * h, c->stream));
    SB_CUDA_CHECK(cudaMemsetAsync(c->direct, 0, sizeof(float4) * size_t(w) * h, c->stream));
    SB_CUDA_CHECK(cudaMemsetAsync(c->aovD, 0, sizeof(float4) * size_t(w) * h, c->stream));
    SB_CUDA_CHECK(cudaMemsetAsync(c->aovS, 0, sizeof(float4) * size_t(w) * h, c->stream));
    c->subframe = 0; // new dimensions reset rendering (OptixRender.cpp:834)
}

// Path-state queues for `np` paths in flight: 180 bytes per path, allocated for the largest batch actually
// rendered so far (one sample per render() call needs W*H paths; a batched render up to maxBatchPaths).
void ensure_queues(sb_ctx* c, size_t np)
{
    if (np <= c->queuePaths)
        return;
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream)); // nothing may still be using the old queues
    free_queues(c);
    for (int i = 0; i < 2; ++i)
    {
        c->Q.rayO[i] = dev_alloc<float4>(np);
        c->Q.rayD[i] = dev_alloc<float4>(np);
        c->Q.thr[i] = dev_alloc<float4>(np);
    }
    c->Q.hitA = dev_alloc<float4>(np);
    c->Q.hitB = dev_alloc<uint32_t>(np);
    c->Q.Lacc = dev_alloc<float4>(np);
    c->Q.shO = dev_alloc<float4>(np);
    c->Q.shD = dev_alloc<float4>(np);
    c->Q.shC = dev_alloc<float4>(np);
    c->queuePaths = np;
}

void fill_camera(const sb_ctx* c, float aspect, FrameParams& P)
{
    if (c->rawMatrices)
    {
        std::memcpy(P.clipToView, c->clipToViewRaw, sizeof(P.clipToView));
    }
    else
    {
        clip_to_view_from_fov(c->fovY, aspect, P.clipToView);
    }
    std::memcpy(P.viewToWorld, c->viewToWorld, sizeof(P.viewToWorld));
}

uint32_t local_sample_budget(const sb_settings& st)
{
    // number of global sample indices offset + k*stride below sppTotal
    const uint32_t stride = st.sample_stride ? st.sample_stride : 1u;
    if (st.sample_offset >= st.spp_total)
        return 0u;
    return (st.spp_total - st.sample_offset + stride - 1u) / stride;
}

// SharedContext::mSubframeIndex = 0: the next launch restarts the beauty and AOV accumulations
void reset_accum(sb_ctx* c)
{
    c->subframe = 0;
    c->shardRequested = 0;
    if (c->S)
    {
        const size_t bytes = sizeof(float4) * size_t(c->width) * c->height;
        SB_CUDA_CHECK(cudaMemsetAsync(c->S, 0, bytes, c->stream));
        SB_CUDA_CHECK(cudaMemsetAsync(c->aovD, 0, bytes, c->stream));
        SB_CUDA_CHECK(cudaMemsetAsync(c->aovS, 0, bytes, c->stream));
    }
}

void set_error(sb_ctx* c, const std::string& msg)
{
    if (c)
        c->error = msg;
    else
        g_createError = msg;
}

#define SB_API_BEGIN(ctxptr)                                                                                          \
    sb_ctx* ctx_ = (ctxptr);                                                                                          \
    try                                                                                                               \
    {                                                                                                                 \
        if (ctx_)                                                                                                     \
            cudaSetDevice(ctx_->device);
#define SB_API_END                                                                                                    \
    }                                                                                                                 \
    catch (const std::bad_alloc&)                                                                                     \
    {                                                                                                                 \
        set_error(ctx_, "out of memory");                                                                             \
        return SB_OUT_OF_MEMORY;                                                                                      \
    }                                                                                                                 \
    catch (const std::exception& e)                                                                                   \
    {                                                                                                                 \
        set_error(ctx_, e.what());                                                                                    \
        return SB_FAIL;                                                                                               \
    }                                                                                                                 \
    return SB_OK;

LaunchCfg launch_cfg(const sb_ctx* c)
{
    LaunchCfg l;
    l.stream = c->stream;
    l.numSms = c->numSms;
    l.timer = c->stageTimers ? const_cast<StageTimer*>(&c->timer) : nullptr;
    l.launchCount = const_cast<uint64_t*>(&c->launchCount);
    l.fusedSmall = c->fusedSmall;
    return l;
}

// One or more wavefront batches rendering `samples` consecutive local samples starting at c->subframe.
// mode 0/1/2 as in accumulate_pixel.
void render_samples(sb_ctx* c, uint32_t samples, uint32_t mode, bool debugNormals)
{
    const sb_settings& st = c->settings;
    FrameParams P;
    std::memset(&P, 0, sizeof(P));
    P.width = c->width;
    P.height = c->height;
    P.tilesX = c->tilesX;
    P.nPixPadded = c->nPixPadded;
    P.maxDepth = std::min<uint32_t>(st.depth, kMaxDepth - 1);
    P.sppTotal = st.spp_total;
    P.rectMethod = st.rect_light_sampling_method;
    P.debug = debugNormals ? 1u : (st.debug == 2u || st.debug == 3u ? st.debug : 0u);
    P.shadowTmin = st.shadow_ray_tmin;
    P.materialTmin = st.material_ray_tmin;
    fill_camera(c, float(c->width) / float(c->height), P);
    compute_exposure(st, P.exposure);
    P.sampleStride = st.sample_stride ? st.sample_stride : 1u;
    P.numLights = c->scene.numLights;
    const LaunchCfg cfg = launch_cfg(c);
    const uint32_t chunkMax = c->batchPaths / c->nPixPadded;
    uint32_t done = 0;
    // modes 1/2 form the linear mean of the whole launch (any render/pt/spp, like the reference's samples_per_launch
    // loop, OptixRender.cu:94-167): a launch larger than one wavefront batch keeps its running sum between batches
    const bool multiBatchAov = mode != 0u && samples > chunkMax && P.debug >= 2u;
    if (multiBatchAov && !c->aovScrD)
    {
        c->aovScrD = dev_alloc<float4>(size_t(c->width) * c->height);
        c->aovScrS = dev_alloc<float4>(size_t(c->width) * c->height);
    }
    const AccumTargets A = { c->S, c->direct, c->aovD, c->aovS, c->aovScrD, c->aovScrS };
    while (done < samples)
    {
        const uint32_t chunk = std::min(samples - done, chunkMax);
        P.chunk = chunk;
        ensure_queues(c, size_t(c->nPixPadded) * chunk);
        P.sampleBase = st.sample_offset + (c->subframe + done) * P.sampleStride;
        launch_wavefront_batch(cfg, P, c->scene, c->Q, c->trackStats);
        const uint32_t flags = (done == 0u ? kBatchFirst : 0u) | (done + chunk == samples ? kBatchLast : 0u);
        if (mode == 0u)
            launch_accumulate(cfg, P, c->Q, A, 0u, c->subframe + done, chunk, kBatchFirst | kBatchLast);
        else
            launch_accumulate(cfg, P, c->Q, A, mode, c->subframe, samples, flags);
        done += chunk;
    }
}

void write_output(sb_ctx* c, sb_buffer* out, bool fromDirect, bool post, uint32_t totalSamples)
{
    const LaunchCfg cfg = launch_cfg(c);
    const uint32_t npix = c->width * c->height;
    float e[3];
    compute_exposure(c->settings, e);
    const uint32_t dbg = c->settings.debug;
    if (fromDirect && dbg != 2u && dbg != 3u)
    {
        launch_copy_image(cfg, c->direct, out->dev, npix, out->format, e, post ? c->settings.tonemapper_type : 0u, post ? c->settings.gamma : 0.0f);
        return;
    }
    const float4* src = dbg == 2u ? c->aovD : (dbg == 3u ? c->aovS : c->S);
    launch_resolve(cfg, src, out->dev, npix, (dbg == 2u || dbg == 3u) ? 0xffffffffu : totalSamples, e, post ? c->settings.tonemapper_type : 0u,
                   post ? c->settings.gamma : 0.0f, out->format);
}

// OptiXRender::render (OptixRender.cpp:874-1057) for `iterations` consecutive calls
void render_impl(sb_ctx* c, sb_buffer* out, uint32_t iterations)
{
    if (!c->haveScene)
        throw std::runtime_error("sb_render: no scene set");
    if (!c->haveSettings)
        throw std::runtime_error("sb_render: no settings set");
    if (!out || out->ctx != c)
        throw std::runtime_error("sb_render: output buffer does not belong to this context");
    if (out->width == 0 || out->height == 0)
        throw std::runtime_error("sb_render: empty output buffer");
    ensure_frame(c, out->width, out->height);
    const sb_settings& st = c->settings;
    SB_CUDA_CHECK(cudaEventRecord(c->evStart, c->stream));
    const bool debugNormals = (st.debug == 1u);
    const bool acc = st.enable_acc != 0 && !debugNormals;
    bool renderedDirect = false;
    if (acc && st.spp == 1u)
    {
        // fast path: `iterations` launches of one sample each == one or more batches of many samples
        const uint32_t budget = local_sample_budget(st);
        const uint32_t left = budget > c->subframe ? budget - c->subframe : 0u;
        const uint32_t n = std::min(iterations, left);
        if (n)
            render_samples(c, n, 0u, false);
        c->subframe += n;
    }
    else
    {
        for (uint32_t it = 0; it < iterations; ++it)
        {
            uint32_t samples;
            if (acc)
            {
                const uint32_t budget = local_sample_budget(st);
                const uint32_t left = budget > c->subframe ? budget - c->subframe : 0u;
                samples = std::min(st.spp, left);
            }
            else
            {
                samples = debugNormals ? 1u : st.spp;
            }
            if (samples == 0u)
                break; // nothing left: the image is the accumulated estimate (OptixRender.cpp:1020-1030)
            if (acc)
            {
                render_samples(c, samples, 1u, false);
                c->subframe += samples;
            }
            else
            {
                c->subframe = 0;
                render_samples(c, samples, 2u, debugNormals);
                renderedDirect = true;
            }
        }
    }
    // debug == 1 skips the post-process (OptixRender.cpp:1045-1049)
    write_output(c, out, renderedDirect, !debugNormals, c->subframe);
    SB_CUDA_CHECK(cudaEventRecord(c->evStop, c->stream));
}

// ---- NCCL, loaded on demand ------------------------------------------------------------------------------------
// A single-GPU host never needs libnccl; a process that already holds one (torch's bundled copy) must share it.
// dlopen("libnccl.so.2") resolves to the copy already mapped under that SONAME, else to the system one;
// STRELKA_B200_NCCL names an explicit path.
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    // NCCL >= 2.28 only (symmetric memory + device API); null with an older library
    ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;
    ncclResult_t (*MemFree)(void*) = nullptr;
#if SB_HAVE_NCCL_DEVICE
    ncclResult_t (*CommWindowRegister)(ncclComm_t, void*, size_t, ncclWindow_t*, int) = nullptr;
    ncclResult_t (*CommWindowDeregister)(ncclComm_t, ncclWindow_t) = nullptr;
    ncclResult_t (*DevCommCreate)(ncclComm_t, ncclDevCommRequirements_t const*, ncclDevComm_t*) = nullptr;
    ncclResult_t (*DevCommDestroy)(ncclComm_t, ncclDevComm_t const*) = nullptr;
#endif
};
NcclApi& nccl_api()
{
    static NcclApi api;
    if (api.handle)
        return api;
    const char* env = getenv("STRELKA_B200_NCCL");
    void* h = dlopen(env && *env ? env : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
        throw std::runtime_error(std::string("multi-GPU rendering needs NCCL: ") + dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p)
            throw std::runtime_error(std::string("libnccl lacks ") + name);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.MemAlloc = reinterpret_cast<decltype(api.MemAlloc)>(dlsym(h, "ncclMemAlloc"));
    api.MemFree = reinterpret_cast<decltype(api.MemFree)>(dlsym(h, "ncclMemFree"));
#if SB_HAVE_NCCL_DEVICE
    api.CommWindowRegister = reinterpret_cast<decltype(api.CommWindowRegister)>(dlsym(h, "ncclCommWindowRegister"));
    api.CommWindowDeregister = reinterpret_cast<decltype(api.CommWindowDeregister)>(dlsym(h, "ncclCommWindowDeregister"));
    api.DevCommCreate = reinterpret_cast<decltype(api.DevCommCreate)>(dlsym(h, "ncclDevCommCreate"));
    api.DevCommDestroy = reinterpret_cast<decltype(api.DevCommDestroy)>(dlsym(h, "ncclDevCommDestroy"));
#endif
    api.handle = h;
    return api;
}
void nccl_check(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess)
        throw std::runtime_error(std::string(what) + ": " + nccl_api().GetErrorString(r));
}

// ---- NVLS path: S / Sglobal as NCCL symmetric windows (sb_nvls.cuh) ---------------------------------------------
// Give the symmetric buffers back; with keepFrame the accumulated S moves into a plain allocation so that the context
// keeps rendering (sb_comm_destroy), without it the frame is going away anyway (resize, sb_destroy).
void release_sym(sb_ctx* c, bool keepFrame)
{
    if (!c->symBuffers)
        return;
    cudaStreamSynchronize(c->stream);
    NcclApi& api = nccl_api();
#if SB_HAVE_NCCL_DEVICE
    if (c->comm && api.CommWindowDeregister)
    {
        if (c->winS)
            api.CommWindowDeregister(c->comm, c->winS);
        if (c->winG)
            api.CommWindowDeregister(c->comm, c->winG);
    }
    c->winS = c->winG = nullptr;
#endif
    float4* plain = nullptr;
    if (keepFrame && c->S && c->symPix)
    {
        plain = dev_alloc<float4>(c->symPix);
        cudaMemcpy(plain, c->S, sizeof(float4) * c->symPix, cudaMemcpyDeviceToDevice);
    }
    if (c->S)
        api.MemFree(c->S);
    if (c->Sglobal)
        api.MemFree(c->Sglobal);
    c->S = plain;
    c->Sglobal = nullptr;
    c->symBuffers = false;
    c->symPix = 0;
}

#if SB_HAVE_NCCL_DEVICE
// (re)create the symmetric S / G pair for the current frame size; collective over the group
void ensure_sym(sb_ctx* c)
{
    const size_t npix = size_t(c->width) * c->height;
    if (c->symBuffers && c->symPix == npix)
        return;
    NcclApi& api = nccl_api();
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    release_sym(c, true);
    void *ns = nullptr, *ng = nullptr;
    nccl_check(api.MemAlloc(&ns, sizeof(float4) * npix), "ncclMemAlloc");
    nccl_check(api.MemAlloc(&ng, sizeof(float4) * npix), "ncclMemAlloc");
    if (c->S)
    {
        SB_CUDA_CHECK(cudaMemcpy(ns, c->S, sizeof(float4) * npix, cudaMemcpyDeviceToDevice));
        cudaFree(c->S);
    }
    else
    {
        SB_CUDA_CHECK(cudaMemset(ns, 0, sizeof(float4) * npix));
    }
    if (c->Sglobal)
        cudaFree(c->Sglobal);
    c->S = static_cast<float4*>(ns);
    c->Sglobal = static_cast<float4*>(ng);
    c->symBuffers = true;
    c->symPix = npix;
    nccl_check(api.CommWindowRegister(c->comm, c->S, sizeof(float4) * npix, &c->winS, NCCL_WIN_COLL_SYMMETRIC), "ncclCommWindowRegister(S)");
    nccl_check(api.CommWindowRegister(c->comm, c->Sglobal, sizeof(float4) * npix, &c->winG, NCCL_WIN_COLL_SYMMETRIC), "ncclCommWindowRegister(G)");
}
#endif

// samples with global index < sppTotal that rank r of `world` owns (indices r, r + world, ...)
uint32_t shard_budget(uint32_t sppTotal, uint32_t r, uint32_t world)
{
    return r >= sppTotal ? 0u : (sppTotal - r + world - 1u) / world;
}

} // namespace

extern "C" {

uint32_t sb_abi_version(void)
{
    return SB_API_VERSION;
}

uint32_t sb_abi_struct_size(uint32_t which)
{
    static const uint32_t sizes[] = { sizeof(sb_settings), sizeof(sb_device_cfg), sizeof(sb_counters), sizeof(sb_scene_view),
                                      sizeof(sb_material), sizeof(sb_light),      sizeof(sb_instance), sizeof(sb_vertex),
                                      sizeof(sb_hit),      sizeof(sb_mesh),       sizeof(sb_curve),    sizeof(sb_texture) };
    return which < sizeof(sizes) / sizeof(sizes[0]) ? sizes[which] : 0u;
}

void sb_settings_default(sb_settings* s)
{
    std::memset(s, 0, sizeof(*s));
    s->spp = 1;
    s->spp_total = 64;
    s->depth = 4;
    s->enable_acc = 1;
    s->rect_light_sampling_method = 0;
    s->debug = 0;
    s->shadow_ray_tmin = 0.0f;
    s->material_ray_tmin = 0.0f;
    s->tonemapper_type = 0;
    s->gamma = 2.4f;
    s->film_iso = 100.0f;
    s->cm2_factor = 1.0f;
    s->f_stop = 4.0f;
    s->shutter_speed = 100.0f;
    s->sample_offset = 0;
    s->sample_stride = 1;
}

sb_result sb_create(const sb_device_cfg* cfg, sb_ctx** out)
{
    if (!out)
        return SB_FAIL;
    *out = nullptr;
    sb_ctx* c = nullptr;
    try
    {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
        {
            g_createError = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); this backend has no CPU fallback";
            return SB_FAIL;
        }
        const int dev = cfg ? cfg->device : 0;
        if (dev < 0 || dev >= count)
        {
            g_createError = "device ordinal out of range";
            return SB_FAIL;
        }
        SB_CUDA_CHECK(cudaSetDevice(dev));
        c = new sb_ctx();
        c->device = dev;
        cudaDeviceProp prop;
        SB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        c->numSms = prop.multiProcessorCount;
        {
            // keep freed blocks of the stream-ordered pool (BVH-build temporaries, exec.h) cached across rebuilds
            cudaMemPool_t pool = nullptr;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
            {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
        }
        SB_CUDA_CHECK(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
        c->stream = c->ownStream;
        c->stageTimers = cfg && (cfg->flags & SB_CFG_STAGE_TIMERS);
        SB_CUDA_CHECK(cudaEventCreate(&c->evStart));
        SB_CUDA_CHECK(cudaEventCreate(&c->evStop));
        c->trackStats = cfg && (cfg->flags & SB_CFG_TRAVERSAL_STATS);
        c->fusedSmall = cfg && (cfg->flags & SB_CFG_FUSED_SMALL);
        if (cfg && cfg->max_batch_paths)
            c->maxBatchPaths = cfg->max_batch_paths;
        if (cfg && cfg->curve_split)
            c->curveSplit = std::min<uint32_t>(cfg->curve_split, 64u);
        c->stats = dev_alloc<StatCounters>(1);
        SB_CUDA_CHECK(cudaMemsetAsync(c->stats, 0, sizeof(StatCounters), c->stream));
        c->sobolTab = upload_sobol_table(c->stream);
        sb_settings_default(&c->settings);
        c->haveSettings = true;
        *out = c;
        return SB_OK;
    }
    catch (const std::exception& e)
    {
        g_createError = e.what();
        delete c;
        return SB_FAIL;
    }
}

void sb_destroy(sb_ctx* c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm)
    {
        release_sym(c, false);
#if SB_HAVE_NCCL_DEVICE
        if (c->nvlsReady && nccl_api().DevCommDestroy)
            nccl_api().DevCommDestroy(c->comm, &c->devComm);
#endif
        nccl_api().CommDestroy(c->comm);
        c->comm = nullptr;
    }
    if (c->evXchgStart)
        cudaEventDestroy(c->evXchgStart);
    if (c->evXchgStop)
        cudaEventDestroy(c->evXchgStop);
    c->timer.release();
    free_scene(c);
    free_frame(c);
    dev_free(c->stats);
    dev_free(c->sobolTab);
    cudaEventDestroy(c->evStart);
    cudaEventDestroy(c->evStop);
    cudaStreamDestroy(c->ownStream);
    delete c;
}

sb_result sb_set_stream(sb_ctx* c, void* cudaStream)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stream = (cudaStream == SB_STREAM_PRIVATE) ? c->ownStream : static_cast<cudaStream_t>(cudaStream);
    SB_API_END
}

const char* sb_last_error(const sb_ctx* c)
{
    return c ? c->error.c_str() : g_createError.c_str();
}

sb_result sb_set_scene(sb_ctx* c, const sb_scene_view* v)
{
    if (!c || !v)
        return SB_FAIL;
    SB_API_BEGIN(c)
    const auto t0 = std::chrono::steady_clock::now();
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    free_scene(c);
    const char* traceEnv = getenv("STRELKA_B200_BUILD_TRACE");
    const bool trace = traceEnv && *traceEnv && *traceEnv != '0';
    auto tmark = std::chrono::steady_clock::now();
    auto hostMark = [&](const char* what) {
        if (!trace)
            return;
        cudaStreamSynchronize(c->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[sb build] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tmark).count());
        tmark = now;
    };
    hostMark("free previous scene");
    ScenePrep prep;
    prepare_scene(v, prep);
    hostMark("host prep (validate, instances)");
    const std::vector<InstDev>& inst = prep.inst;
    const std::vector<uint32_t>& triFirst = prep.triFirst;
    const std::vector<SegInfo>& segInfo = prep.segInfo;
    const uint64_t numTris = prep.numTris;
    const uint32_t numMaterials = prep.numMaterials;

    // ---- uploads (createVertexBuffer ... createLightBuffer, OptixRender.cpp:1117-1189) -----------------
    SceneDev& s = c->scene;
    cudaStream_t st = c->stream;
    s.vertices = dev_upload(v->vertices, v->num_vertices, st);
    s.indices = dev_upload(v->indices, v->num_indices, st);
    s.meshes = dev_upload(v->meshes, v->num_meshes, st);
    s.curves = dev_upload(v->curves, v->num_curves, st);
    s.curvePoints = dev_upload(v->curve_points, v->num_curve_points * 3, st);
    s.curveRadii = dev_upload(v->curve_widths, v->num_curve_widths, st);
    s.curveVertexCounts = dev_upload(v->curve_vertex_counts, v->num_curve_vertex_counts, st);
    s.lights = dev_upload(v->lights, v->num_lights, st);
    if (v->num_materials)
    {
        s.materials = dev_upload(v->materials, v->num_materials, st);
    }
    else
    {
        sb_material def;
        std::memset(&def, 0, sizeof(def));
        def.model = SB_MATERIAL_DIFFUSE; // default.mdl::default_material, injected slot 0 (OptixRender.cpp:1091-1097)
        def.base_color[0] = def.base_color[1] = def.base_color[2] = 1.0f;
        s.materials = dev_upload(&def, 1, st);
    }
    // ---- textures (OptiXRender::loadTextureFromFile, OptixRender.cpp:1191-1268: uchar4 array, wrap, linear filter,
    // normalised coordinates, normalised-float reads) ---------------------------------------------------------------
    bool anyTexturedMaterial = false;
    for (uint32_t i = 0; i < v->num_materials; ++i)
    {
        const sb_material& m = v->materials[i];
        if (m.diffuse_texture > v->num_textures || m.normal_texture > v->num_textures)
            throw std::runtime_error("sb_set_scene: material " + std::to_string(i) + " references a missing texture");
        anyTexturedMaterial = anyTexturedMaterial || m.diffuse_texture != 0u || m.normal_texture != 0u;
    }
    if (v->num_textures && v->textures)
    {
        for (uint32_t i = 0; i < v->num_textures; ++i)
        {
            const sb_texture& t = v->textures[i];
            if (!t.pixels || t.width == 0 || t.height == 0)
                throw std::runtime_error("sb_set_scene: texture " + std::to_string(i) + " is empty");
            const cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
            cudaArray_t arr = nullptr;
            SB_CUDA_CHECK(cudaMallocArray(&arr, &desc, t.width, t.height));
            c->texArrays.push_back(arr);
            SB_CUDA_CHECK(cudaMemcpy2DToArray(arr, 0, 0, t.pixels, size_t(t.width) * 4, size_t(t.width) * 4, t.height, cudaMemcpyHostToDevice));
            cudaResourceDesc res;
            std::memset(&res, 0, sizeof(res));
            res.resType = cudaResourceTypeArray;
            res.res.array.array = arr;
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof(td));
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeNormalizedFloat;
            td.normalizedCoords = 1;
            cudaTextureObject_t obj = 0;
            SB_CUDA_CHECK(cudaCreateTextureObject(&obj, &res, &td, nullptr));
            c->texObjects.push_back(obj);
        }
        static_assert(sizeof(cudaTextureObject_t) == sizeof(unsigned long long), "texture handles are passed as 64-bit integers");
        s.textures = dev_upload(reinterpret_cast<const unsigned long long*>(c->texObjects.data()), c->texObjects.size(), st);
        s.numTextures = v->num_textures;
        if (anyTexturedMaterial && numTris)
        {
            void* p = nullptr;
            if (cudaMallocAsync(&p, sizeof(uint4) * numTris, st) != cudaSuccess)
            {
                cudaGetLastError();
                throw std::bad_alloc();
            }
            s.triUv = static_cast<uint4*>(p);
        }
    }
    s.instances = dev_upload(inst.data(), inst.size(), st);
    s.numInstances = v->num_instances;
    s.numLights = v->num_lights;
    s.numMaterials = numMaterials;
    s.onlyRectLights = v->num_lights > 0;
    for (uint32_t i = 0; i < v->num_lights; ++i)
        s.onlyRectLights = s.onlyRectLights && v->lights[i].type == 0;
    s.anyPreviewMaterial = false;
    s.anyHairMaterial = false;
    for (uint32_t i = 0; i < v->num_materials; ++i)
    {
        s.anyPreviewMaterial = s.anyPreviewMaterial || v->materials[i].model == SB_MATERIAL_USD_PREVIEW_SURFACE;
        s.anyHairMaterial = s.anyHairMaterial || v->materials[i].model == SB_MATERIAL_HAIR;
    }
    s.numMeshes = v->num_meshes;
    s.numCurves = v->num_curves;
    s.numCurvePoints = v->num_curve_points;
    s.numCurveRadii = v->num_curve_widths;
    c->instTriFirst = dev_upload(triFirst.data(), triFirst.size(), st);
    c->segInfoUnsorted = dev_upload(segInfo.data(), segInfo.size(), st);
    SB_CUDA_CHECK(cudaStreamSynchronize(st)); // the host vectors above go out of scope
    hostMark("uploads");

    // ---- createAccelerationStructure (OptixRender.cpp:388-496), the B200 way ---------------------------
    ExecCuda ex;
    ex.stream = st;
    ex.numSms = c->numSms;
    ex.trace = trace;
    ex.traceT0 = std::chrono::steady_clock::now();
    try
    {
        build_scene_bvhs(ex, s, c->instTriFirst, uint32_t(numTris), nullptr, uint32_t(segInfo.size()), c->segInfoUnsorted, c->curveSplit);
    }
    catch (...)
    {
        ex.release();
        throw;
    }
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    ex.release();
    c->haveScene = true;
    reset_accum(c);
    c->buildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    SB_API_END
}

sb_result sb_reset_accumulation(sb_ctx* c)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    reset_accum(c);
    SB_API_END
}

sb_result sb_set_camera(sb_ctx* c, const float view[16], float fovY)
{
    if (!c || !view)
        return SB_FAIL;
    SB_API_BEGIN(c)
    const bool changed = c->rawMatrices || std::memcmp(c->view, view, sizeof(c->view)) != 0 || c->fovY != fovY;
    if (changed)
    {
        if (!view_to_world_from_view(view, c->viewToWorld))
            throw std::runtime_error("sb_set_camera: singular view matrix");
        std::memcpy(c->view, view, sizeof(c->view));
        c->fovY = fovY;
        c->rawMatrices = false;
        // a changed view/projection resets accumulation (OptixRender.cpp:903-908)
        reset_accum(c);
    }
    SB_API_END
}

sb_result sb_set_camera_matrices(sb_ctx* c, const float clipToView[16], const float viewToWorld[16])
{
    if (!c || !clipToView || !viewToWorld)
        return SB_FAIL;
    SB_API_BEGIN(c)
    // like the reference, which resets only when the matrices differ from the previous frame's (OptixRender.cpp:903-908):
    // a host that sets raw matrices every frame keeps accumulating while the camera stands still
    const bool changed = !c->rawMatrices || std::memcmp(c->clipToViewRaw, clipToView, sizeof(c->clipToViewRaw)) != 0 ||
                         std::memcmp(c->viewToWorld, viewToWorld, sizeof(c->viewToWorld)) != 0;
    if (changed)
    {
        std::memcpy(c->clipToViewRaw, clipToView, sizeof(c->clipToViewRaw));
        std::memcpy(c->viewToWorld, viewToWorld, sizeof(c->viewToWorld));
        c->rawMatrices = true;
        reset_accum(c);
    }
    SB_API_END
}

sb_result sb_set_settings(sb_ctx* c, const sb_settings* s)
{
    if (!c || !s)
        return SB_FAIL;
    SB_API_BEGIN(c)
    if (s->depth >= kMaxDepth)
        throw std::runtime_error("sb_set_settings: render/pt/depth must be < 32");
    if (s->spp_total == 0)
        throw std::runtime_error("sb_set_settings: render/pt/sppTotal must be > 0");
    // the changes that reset accumulation in the reference (OptixRender.cpp:913-934) + the sharding keys
    const sb_settings& o = c->settings;
    const bool reset = (o.rect_light_sampling_method != s->rect_light_sampling_method) || ((o.enable_acc != 0) != (s->enable_acc != 0)) ||
                       (o.spp_total > s->spp_total) || (o.sample_offset != s->sample_offset) || (o.sample_stride != s->sample_stride) ||
                       (o.depth != s->depth) || (o.debug != s->debug);
    c->settings = *s;
    if (c->settings.sample_stride == 0)
        c->settings.sample_stride = 1;
    c->haveSettings = true;
    if (reset)
    {
        reset_accum(c);
    }
    SB_API_END
}

uint32_t sb_subframe_index(const sb_ctx* c)
{
    return c ? c->subframe : 0u;
}

// ---- buffers ------------------------------------------------------------------------------------------
static size_t format_size(uint32_t f)
{
    return f == SB_FORMAT_FLOAT4 ? 16 : (f == SB_FORMAT_FLOAT3 ? 12 : 4);
}

sb_result sb_buffer_resize(sb_buffer* b, uint32_t w, uint32_t h)
{
    if (!b)
        return SB_FAIL;
    SB_API_BEGIN(b->ctx)
    SB_CUDA_CHECK(cudaStreamSynchronize(b->ctx->stream));
    if (b->dev)
        cudaFree(b->dev);
    if (b->host)
        cudaFreeHost(b->host);
    b->dev = b->host = nullptr;
    b->width = w;
    b->height = h;
    b->bytes = size_t(w) * h * format_size(b->format);
    if (b->bytes)
    {
        if (cudaMalloc(&b->dev, b->bytes) != cudaSuccess)
            throw std::bad_alloc();
        if (cudaMallocHost(&b->host, b->bytes) != cudaSuccess)
            throw std::bad_alloc();
        SB_CUDA_CHECK(cudaMemsetAsync(b->dev, 0, b->bytes, b->ctx->stream));
        std::memset(b->host, 0, b->bytes);
    }
    SB_API_END
}

sb_result sb_buffer_create(sb_ctx* c, uint32_t w, uint32_t h, uint32_t format, sb_buffer** out)
{
    if (!c || !out || format > SB_FORMAT_FLOAT3)
        return SB_FAIL;
    sb_buffer* b = new sb_buffer();
    b->ctx = c;
    b->format = format;
    const sb_result r = sb_buffer_resize(b, w, h);
    if (r != SB_OK)
    {
        delete b;
        return r;
    }
    *out = b;
    return SB_OK;
}

void sb_buffer_destroy(sb_buffer* b)
{
    if (!b)
        return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->dev)
        cudaFree(b->dev);
    if (b->host)
        cudaFreeHost(b->host);
    if (b->copied)
        cudaEventDestroy(b->copied);
    delete b;
}

sb_result sb_buffer_map(sb_buffer* b, void** hostPtr)
{
    if (!b)
        return SB_FAIL;
    SB_API_BEGIN(b->ctx)
    if (b->bytes)
    {
        SB_CUDA_CHECK(cudaMemcpyAsync(b->host, b->dev, b->bytes, cudaMemcpyDeviceToHost, b->ctx->stream));
        SB_CUDA_CHECK(cudaStreamSynchronize(b->ctx->stream));
    }
    if (hostPtr)
        *hostPtr = b->host;
    SB_API_END
}

sb_result sb_buffer_map_async(sb_buffer* b)
{
    if (!b)
        return SB_FAIL;
    SB_API_BEGIN(b->ctx)
    if (b->bytes)
    {
        if (!b->copied)
            SB_CUDA_CHECK(cudaEventCreateWithFlags(&b->copied, cudaEventDisableTiming));
        SB_CUDA_CHECK(cudaMemcpyAsync(b->host, b->dev, b->bytes, cudaMemcpyDeviceToHost, b->ctx->stream));
        SB_CUDA_CHECK(cudaEventRecord(b->copied, b->ctx->stream));
        b->copyPending = true;
    }
    SB_API_END
}

sb_result sb_buffer_map_wait(sb_buffer* b, void** hostPtr)
{
    if (!b)
        return SB_FAIL;
    SB_API_BEGIN(b->ctx)
    if (b->copyPending)
    {
        SB_CUDA_CHECK(cudaEventSynchronize(b->copied));
        b->copyPending = false;
    }
    if (hostPtr)
        *hostPtr = b->host;
    SB_API_END
}

sb_result sb_buffer_unmap(sb_buffer* b)
{
    return b ? SB_OK : SB_FAIL;
}
void* sb_buffer_host_ptr(sb_buffer* b)
{
    return b ? b->host : nullptr;
}
size_t sb_buffer_host_size(sb_buffer* b)
{
    return b ? b->bytes : 0;
}
void* sb_buffer_device_ptr(sb_buffer* b)
{
    return b ? b->dev : nullptr;
}
uint32_t sb_buffer_width(const sb_buffer* b)
{
    return b ? b->width : 0;
}
uint32_t sb_buffer_height(const sb_buffer* b)
{
    return b ? b->height : 0;
}

// ---- rendering ------------------------------------------------------------------------------------------
sb_result sb_render(sb_ctx* c, sb_buffer* out)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    render_impl(c, out, 1);
    SB_API_END
}

sb_result sb_render_iterations(sb_ctx* c, sb_buffer* out, uint32_t iterations)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    render_impl(c, out, iterations);
    SB_API_END
}

sb_result sb_synchronize(sb_ctx* c)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    SB_API_END
}

void* sb_accum_device_ptr(sb_ctx* c, uint64_t* numFloats)
{
    if (!c)
        return nullptr;
    if (numFloats)
        *numFloats = uint64_t(c->width) * c->height * 4u;
    return c->S;
}

sb_result sb_resolve(sb_ctx* c, sb_buffer* out, uint32_t totalSamples)
{
    if (!c || !out)
        return SB_FAIL;
    SB_API_BEGIN(c)
    if (!c->S || out->width != c->width || out->height != c->height)
        throw std::runtime_error("sb_resolve: nothing accumulated at this resolution");
    write_output(c, out, false, c->settings.debug != 1u, totalSamples);
    SB_API_END
}

sb_result sb_get_counters(sb_ctx* c, sb_counters* out)
{
    if (!c || !out)
        return SB_FAIL;
    SB_API_BEGIN(c)
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    StatCounters h;
    SB_CUDA_CHECK(cudaMemcpy(&h, c->stats, sizeof(h), cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof(*out));
    out->paths = h.paths;
    out->radiance_rays = h.radianceRays;
    out->shadow_rays = h.shadowRays;
    out->nodes_visited = h.nodes;
    out->tris_tested = h.tris;
    out->segs_tested = h.segs;
    out->stack_overflows = h.overflow;
    out->nodes_visited_shadow = h.nodesSh;
    out->tris_tested_shadow = h.trisSh;
    out->segs_tested_shadow = h.segsSh;
    out->kernel_launches = c->launchCount;
    c->timer.collect();
    for (int i = 0; i < kNumStages; ++i)
    {
        out->stage_ms[i] = c->timer.ms[i];
        out->stage_launches[i] = c->timer.launches[i];
    }
    out->num_triangles = c->scene.numTris;
    out->num_segments = c->scene.numSegs;
    out->bvh_nodes_tri = c->scene.numTriNodes;
    out->bvh_nodes_curve = c->scene.numSegNodes;
    out->build_ms = c->buildMs;
    out->bvh_depth_tri = c->scene.triDepth;
    out->bvh_depth_curve = c->scene.segDepth;
    out->exchange_nvls = c->nvlsReady ? 1u : 0u;
    if (c->evXchgStart)
    {
        float x = 0.0f;
        if (cudaEventElapsedTime(&x, c->evXchgStart, c->evXchgStop) == cudaSuccess)
            out->exchange_ms = x;
        else
            cudaGetLastError();
    }
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, c->evStart, c->evStop) == cudaSuccess)
        c->renderMs = ms;
    else
        cudaGetLastError();
    out->render_ms = c->renderMs;
    SB_API_END
}

sb_result sb_reset_counters(sb_ctx* c)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    SB_CUDA_CHECK(cudaMemsetAsync(c->stats, 0, sizeof(StatCounters), c->stream));
    SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->timer.reset();
    c->launchCount = 0;
    SB_API_END
}

// ---- multi-GPU: sample-stride sharding + one NCCL sum of S ---------------------------------------
sb_result sb_comm_get_unique_id(void* idOut)
{
    if (!idOut)
        return SB_FAIL;
    SB_API_BEGIN(nullptr)
    static_assert(SB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "sb_api.h and nccl.h disagree on the id size");
    ncclUniqueId id;
    nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(idOut, &id, sizeof(id));
    SB_API_END
}

sb_result sb_comm_init(sb_ctx* c, const void* idBytes, uint32_t rank, uint32_t world)
{
    if (!c || !idBytes || world == 0 || rank >= world)
        return SB_FAIL;
    SB_API_BEGIN(c)
    if (c->comm)
        throw std::runtime_error("sb_comm_init: this context already belongs to a group");
    ncclUniqueId id;
    std::memcpy(&id, idBytes, sizeof(id));
    nccl_check(nccl_api().CommInitRank(&c->comm, int(world), id, int(rank)), "ncclCommInitRank");
    c->commRank = int(rank);
    c->commWorld = int(world);
    // the fused NVLS all-reduce + resolve (sb_nvls.cuh) where library, headers and hardware allow it; every rank takes the
    // same decision (the inputs are properties of the group), else ncclAllReduce + k_resolve
    c->nvlsReady = false;
    if (world > 1)
    {
#if SB_HAVE_NCCL_DEVICE
        NcclApi& api = nccl_api();
        const char* env = getenv("STRELKA_B200_NVLS");
        if (env && *env == '0')
            c->nvlsStatus = "disabled by STRELKA_B200_NVLS=0";
        else if (!api.DevCommCreate || !api.CommWindowRegister || !api.MemAlloc || !api.MemFree)
            c->nvlsStatus = "libnccl older than 2.28: no device API";
        else
        {
            c->nvlsGrid = c->numSms * kNvlsCtasPerSm;
            ncclDevCommRequirements req;
            std::memset(&req, 0, sizeof(req));
            req.lsaMultimem = true;
            req.lsaBarrierCount = c->nvlsGrid;
            const ncclResult_t r = api.DevCommCreate(c->comm, &req, &c->devComm);
            if (r != ncclSuccess)
                c->nvlsStatus = std::string("ncclDevCommCreate: ") + api.GetErrorString(r);
            else if (c->devComm.lsaSize != int(world))
                c->nvlsStatus = "the ranks are not one NVLink (LSA) domain";
            else
            {
                c->nvlsReady = true;
                c->nvlsStatus = "NVLS multimem: fused all-reduce + resolve kernel";
            }
        }
#else
        c->nvlsStatus = "built without the NCCL device API headers (NCCL >= 2.28)";
#endif
    }
    // this rank renders the global sample indices rank, rank + world, ... (sampler indices unchanged)
    c->settings.sample_offset = rank;
    c->settings.sample_stride = world;
    reset_accum(c);
    SB_API_END
}

sb_result sb_comm_destroy(sb_ctx* c)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    if (c->comm)
    {
        SB_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        release_sym(c, true);
#if SB_HAVE_NCCL_DEVICE
        if (c->nvlsReady && nccl_api().DevCommDestroy)
            nccl_api().DevCommDestroy(c->comm, &c->devComm);
#endif
        c->nvlsReady = false;
        c->nvlsStatus = "not a multi-GPU context";
        nccl_api().CommDestroy(c->comm);
        c->comm = nullptr;
        c->commRank = 0;
        c->commWorld = 1;
        c->settings.sample_offset = 0;
        c->settings.sample_stride = 1;
        reset_accum(c);
    }
    SB_API_END
}

uint32_t sb_comm_world(const sb_ctx* c)
{
    return c ? uint32_t(c->commWorld) : 0u;
}

const char* sb_comm_exchange_path(const sb_ctx* c)
{
    if (!c)
        return "";
    return (c->commWorld > 1 && !c->nvlsReady) ? (c->exchangeNote = "ncclAllReduce + k_resolve (" + c->nvlsStatus + ")").c_str() : c->nvlsStatus.c_str();
}

sb_result sb_render_sharded(sb_ctx* c, sb_buffer* out, uint32_t iterations)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    const sb_settings& st = c->settings;
    if (st.spp != 1u || st.enable_acc == 0u || st.debug != 0u)
        throw std::runtime_error("sb_render_sharded: sample sharding needs render/pt/spp 1, enableAcc on and no debug view "
                                 "(only then is the accumulation a plain sum over samples, SURVEY.md 8e / quirk Q1)");
    if (st.sample_offset != uint32_t(c->commRank) || st.sample_stride != uint32_t(c->commWorld))
        throw std::runtime_error("sb_render_sharded: settings.sample_offset / sample_stride must be this context's rank / world size");
    if (!out || out->ctx != c || out->width == 0 || out->height == 0)
        throw std::runtime_error("sb_render_sharded: bad output buffer");
    ensure_frame(c, out->width, out->height);
#if SB_HAVE_NCCL_DEVICE
    if (c->commWorld > 1 && c->nvlsReady)
        ensure_sym(c); // S and Sglobal become NCCL symmetric windows (collective; once per resolution)
#endif
    render_impl(c, out, iterations); // this rank's share: stops at its own budget
    if (c->commWorld > 1)
    {
        if (!c->evXchgStart)
        {
            SB_CUDA_CHECK(cudaEventCreate(&c->evXchgStart));
            SB_CUDA_CHECK(cudaEventCreate(&c->evXchgStop));
        }
        SB_CUDA_CHECK(cudaEventRecord(c->evXchgStart, c->stream));
        // every rank is driven with the same call sequence, so the global sample count needs no exchange
        c->shardRequested += iterations;
        uint64_t total = 0;
        for (int r = 0; r < c->commWorld; ++r)
            total += std::min<uint64_t>(c->shardRequested, shard_budget(st.spp_total, uint32_t(r), uint32_t(c->commWorld)));
        const size_t npix = size_t(c->width) * c->height;
        const LaunchCfg cfg = launch_cfg(c);
        float e[3];
        compute_exposure(st, e);
        bool fused = false;
#if SB_HAVE_NCCL_DEVICE
        if (c->nvlsReady)
        {
            // ONE kernel: the NVSwitch sums the ranks' S in flight, the sums are multicast into every rank's Sglobal and
            // resolved in place (sb_nvls.cuh).  S itself stays this rank's partial sum.
            launch_allreduce_resolve_nvls(cfg, c->devComm, c->winS, c->winG, c->Sglobal, out->dev, uint32_t(npix), uint32_t(total), e,
                                          st.tonemapper_type, st.gamma, out->format, c->nvlsGrid);
            fused = true;
        }
#endif
        if (!fused)
        {
            if (!c->Sglobal)
                c->Sglobal = dev_alloc<float4>(npix);
            // out of place: S stays this rank's partial sum, so further calls keep accumulating correctly
            nccl_check(nccl_api().AllReduce(c->S, c->Sglobal, npix * 4, ncclFloat32, ncclSum, c->comm, c->stream), "ncclAllReduce");
            launch_resolve(cfg, c->Sglobal, out->dev, uint32_t(npix), uint32_t(total), e, st.tonemapper_type, st.gamma, out->format);
        }
        SB_CUDA_CHECK(cudaEventRecord(c->evXchgStop, c->stream));
        SB_CUDA_CHECK(cudaEventRecord(c->evStop, c->stream));
    }
    SB_API_END
}

// ---- test hooks ---------------------------------------------------------------------------------------------
sb_result sb_test_sampler(sb_ctx* c, uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxs,
                          const uint32_t* depth, const uint32_t* dim, float* out)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    cudaStream_t st = c->stream;
    uint32_t* d[6];
    const uint32_t* h[6] = { x, y, sample, maxs, depth, dim };
    for (int i = 0; i < 6; ++i)
        d[i] = dev_upload(h[i], n, st);
    float* dout = dev_alloc<float>(n);
    launch_test_sampler(launch_cfg(c), n, d[0], d[1], d[2], d[3], d[4], d[5], dout);
    SB_CUDA_CHECK(cudaMemcpyAsync(out, dout, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int i = 0; i < 6; ++i)
        cudaFree(d[i]);
    cudaFree(dout);
    SB_API_END
}

sb_result sb_test_light_sample(sb_ctx* c, uint32_t n, const sb_light* lights, const float* hp, const float* u, uint32_t method, float* out)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    cudaStream_t st = c->stream;
    sb_light* dl = dev_upload(lights, n, st);
    float* dh = dev_upload(hp, size_t(n) * 3, st);
    float* du = dev_upload(u, size_t(n) * 2, st);
    float* dout = dev_alloc<float>(size_t(n) * 12);
    launch_test_light_sample(launch_cfg(c), n, dl, dh, du, method, dout);
    SB_CUDA_CHECK(cudaMemcpyAsync(out, dout, sizeof(float) * 12 * n, cudaMemcpyDeviceToHost, st));
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dl);
    cudaFree(dh);
    cudaFree(du);
    cudaFree(dout);
    SB_API_END
}

sb_result sb_test_offset_ray(sb_ctx* c, uint32_t n, const float* p, const float* nrm, float* out)
{
    if (!c || !p || !nrm || !out)
        return SB_FAIL;
    SB_API_BEGIN(c)
    cudaStream_t st = c->stream;
    float* dp = dev_upload(p, size_t(n) * 3, st);
    float* dn = dev_upload(nrm, size_t(n) * 3, st);
    float* dout = dev_alloc<float>(size_t(n) * 3);
    launch_test_offset_ray(launch_cfg(c), n, dp, dn, dout);
    SB_CUDA_CHECK(cudaMemcpyAsync(out, dout, sizeof(float) * 3 * size_t(n), cudaMemcpyDeviceToHost, st));
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dp);
    cudaFree(dn);
    cudaFree(dout);
    SB_API_END
}

sb_result sb_test_texture(sb_ctx* c, uint32_t index, uint32_t n, const float* uv, float* out)
{
    if (!c || !uv || !out)
        return SB_FAIL;
    SB_API_BEGIN(c)
    if (!c->haveScene || index >= c->scene.numTextures)
        throw std::runtime_error("sb_test_texture: no such texture in the current scene");
    cudaStream_t st = c->stream;
    float* duv = dev_upload(uv, size_t(n) * 2, st);
    float* dout = dev_alloc<float>(size_t(n) * 4);
    launch_test_texture(launch_cfg(c), c->scene, index + 1u, n, duv, dout);
    SB_CUDA_CHECK(cudaMemcpyAsync(out, dout, sizeof(float) * 4 * size_t(n), cudaMemcpyDeviceToHost, st));
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(duv);
    cudaFree(dout);
    SB_API_END
}

sb_result sb_test_bsdf(sb_ctx* c, const sb_material* m, uint32_t n, const float* in, float* out)
{
    if (!c || !m || !in || !out)
        return SB_FAIL;
    SB_API_BEGIN(c)
    cudaStream_t st = c->stream;
    float* din = dev_upload(in, size_t(n) * 19, st);
    float* dout = dev_alloc<float>(size_t(n) * 15);
    launch_test_bsdf(launch_cfg(c), *m, n, din, dout);
    SB_CUDA_CHECK(cudaMemcpyAsync(out, dout, sizeof(float) * 15 * size_t(n), cudaMemcpyDeviceToHost, st));
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(din);
    cudaFree(dout);
    SB_API_END
}

sb_result sb_test_trace(sb_ctx* c, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits)
{
    if (!c)
        return SB_FAIL;
    SB_API_BEGIN(c)
    if (!c->haveScene)
        throw std::runtime_error("sb_test_trace: no scene set");
    cudaStream_t st = c->stream;
    if (mode > 3u)
        throw std::runtime_error("sb_test_trace: mode must be 0..3");
    float* dr = dev_upload(rays, size_t(n) * 8, st);
    sb_hit* dh = dev_alloc<sb_hit>(n);
    if (mode >= 2u)
    {
        // the production path: real queue records, the stage launchers of launch_wavefront_batch
        FrameParams P;
        std::memset(&P, 0, sizeof(P));
        P.materialTmin = n ? rays[3] : 0.0f;
        if (mode == 2u)
            for (uint32_t i = 0; i < n; ++i)
                if (rays[8 * size_t(i) + 3] != P.materialTmin || rays[8 * size_t(i) + 7] < 1e16f)
                    throw std::runtime_error("sb_test_trace mode 2: radiance rays share one tmin and have tmax >= 1e16 (OptixRender.cu:120-129)");
        ensure_queues(c, n);
        if (!c->Q.counts)
            c->Q.counts = dev_alloc<uint32_t>(kNumCounts);
        c->Q.stats = c->stats;
        c->Q.sobolTab = c->sobolTab;
        launch_test_trace_production(launch_cfg(c), P, c->scene, c->Q, c->instTriFirst, n, dr, mode, dh, c->trackStats);
    }
    else
        launch_test_trace(launch_cfg(c), c->scene, n, dr, mode, dh);
    SB_CUDA_CHECK(cudaMemcpyAsync(hits, dh, sizeof(sb_hit) * n, cudaMemcpyDeviceToHost, st));
    SB_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dr);
    cudaFree(dh);
    SB_API_END
}

} // extern "C"
