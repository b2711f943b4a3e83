/*
 * sb_api.h -- C ABI of the B200-native Strelka path-tracing backend ("sb" = Strelka/B200).
 *
 * This is the drop-in boundary for the reference's render backend interface.  Every entry
 * point cites the reference interface (file:line under arhix52/Strelka) it replaces.  The
 * ABI is plain C: opaque handles, POD structs, pointers and sizes; no C++/torch types, no
 * exceptions cross it.  All functions return sb_result unless stated otherwise and record a
 * human-readable message retrievable with sb_last_error().
 *
 * Threading contract (reference: render() is called from one thread, one call at a time,
 * OptixRender.cpp:874-1057): one sb_ctx per GPU, driven by one host thread at a time.
 *
 * There is NO CPU fallback behind this ABI: sb_create() fails if no CUDA device is usable.
 */
#ifndef SB_API_H
#define SB_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_API_VERSION 2

/* Version / layout handshake for bindings: SB_API_VERSION of the built library, and sizeof() of the ABI structs
 * (which: 0 sb_settings, 1 sb_device_cfg, 2 sb_counters, 3 sb_scene_view, 4 sb_material, 5 sb_light, 6 sb_instance,
 * 7 sb_vertex, 8 sb_hit, 9 sb_mesh, 10 sb_curve, 11 sb_texture; 0 for anything else). */
uint32_t sb_abi_version(void);
uint32_t sb_abi_struct_size(uint32_t which);

/* oka::Result, include/render/common.h:30-35 (declared but unused by the reference, which
 * assert(0)s instead, OptixRender.cpp:61-103; this ABI never aborts). */
typedef enum sb_result
{
    SB_OK = 0,
    SB_FAIL = 1,
    SB_OUT_OF_MEMORY = 2
} sb_result;

/* ------------------------------------------------------------------------------------------
 * Scene payload: glm-free POD mirrors of the arrays oka::Scene hands to a backend
 * (include/scene/scene.h).  Layouts are byte-identical to the reference's so an adapter can
 * pass std::vector<>::data() straight through.
 * ---------------------------------------------------------------------------------------- */

/* oka::Scene::Vertex, scene.h:80-89 == device Vertex, OptixRenderParams.h:19-28 (32 B).
 * tangent/normal: 10-10-10 bit packed (RenderPass.cpp:53-59), uv: 16-16 bit packed over
 * [-10,10] (RenderPass.cpp:61-67). */
typedef struct sb_vertex
{
    float pos[3];
    uint32_t tangent;
    uint32_t normal;
    uint32_t uv;
    float pad0;
    float pad1;
} sb_vertex;

/* oka::Mesh, scene.h:21-27.  Indices are mesh-local (closest_hit.cu:369-377). */
typedef struct sb_mesh
{
    uint32_t index;        /* first index in the index buffer */
    uint32_t count;        /* number of indices */
    uint32_t vb_offset;    /* first vertex in the vertex buffer */
    uint32_t vertex_count; /* number of vertices */
} sb_mesh;

/* oka::Curve, scene.h:29-42 (cubic B-spline strands; segment indexing OptixRender.cpp:232-245) */
typedef struct sb_curve
{
    uint32_t vertex_counts_start;
    uint32_t vertex_counts_count;
    uint32_t points_start;
    uint32_t points_count;
    uint32_t widths_start; /* (uint32_t)-1 if none */
    uint32_t widths_count;
} sb_curve;

/* oka::Instance::Type, scene.h:47-52 */
enum
{
    SB_INSTANCE_MESH = 0,
    SB_INSTANCE_LIGHT = 1,
    SB_INSTANCE_CURVE = 2
};

/* oka::Instance, scene.h:44-60.  transform = glm::mat4 storage order (column-major, the
 * translation lives in elements 12..14). material_id (uint32_t)-1 -> 0 (OptixRender.cpp:766). */
typedef struct sb_instance
{
    float transform[16];
    uint32_t type;
    uint32_t geom_id; /* mMeshId / mCurveId */
    uint32_t material_id;
    uint32_t light_id; /* (uint32_t)-1 for non-lights */
} sb_instance;

/* oka::Scene::Light == UniformLight, scene.h:146-155 / include/render/Lights.h:5-14 (112 B).
 * type: 0 rect, 1 disc, 2 sphere, 3 distant. */
typedef struct sb_light
{
    float points[4][4];
    float color[4];
    float normal[4];
    int32_t type;
    float half_angle;
    float pad0;
    float pad1;
} sb_light;

/* Materials.  The reference resolves every material through the (absent, closed) MDL SDK at
 * run time (materialmanager.cpp:524-609); this backend receives them pre-resolved into one of
 * the closed-form models of SURVEY.md Appendix D.  The adapter maps oka::Scene::
 * MaterialDescription (scene.h:65-78) params by name (precedent MetalRender.cpp:84-89). */
enum
{
    SB_MATERIAL_DIFFUSE = 0,            /* default.mdl::default_material: diffuse_color */
    SB_MATERIAL_USD_PREVIEW_SURFACE = 1, /* ND_UsdPreviewSurface_surfaceshader params */
    SB_MATERIAL_HAIR = 2                 /* df::chiang_hair_bsdf-style hair material */
};

typedef struct sb_material
{
    uint32_t model;
    float base_color[3]; /* diffuse_color / diffuseColor */
    float roughness;
    float metallic;
    float ior;
    float opacity; /* carried, not evaluated (the reference never evaluates cutout opacity) */
    float clearcoat;
    float clearcoat_roughness;
    float specular_color[3];
    uint32_t use_specular_workflow;
    /* hair (SB_MATERIAL_HAIR): */
    float hair_absorption[3]; /* sigma_a */
    float hair_roughness_lon; /* beta_m */
    float hair_roughness_azi; /* beta_n */
    float hair_cuticle_angle; /* alpha, radians */
    /* UsdUVTexture inputs of a UsdPreviewSurface network (MaterialManager::Param::Type::eTexture params
     * "<node>_file", Material.cpp:120-136; resolved by OptixRender.cpp:1346-1387): 1-based index into
     * sb_scene_view.textures, 0 = no texture (MDL's invalid texture index, texture_support_cuda.h:300-304).
     * diffuse_texture replaces base_color (rgb); normal_texture is a tangent-space normal map (rgb * 2 - 1). */
    uint32_t diffuse_texture;
    uint32_t normal_texture;
    float pad[2];
} sb_material; /* 96 B */

/* One 2D texture as the reference creates it (OptiXRender::loadTextureFromFile, OptixRender.cpp:1191-1268):
 * 8-bit RGBA texels (stbi STBI_rgb_alpha: row 0 = top of the image file), sampled with wrap addressing, linear
 * filtering, normalised coordinates and normalised-float reads.  Texture coordinates are the vertices' packed
 * st primvar with v already flipped by the delegate (RenderPass.cpp:109-114). */
typedef struct sb_texture
{
    const uint8_t* pixels; /* width * height * 4 bytes, borrowed for the duration of sb_set_scene */
    uint32_t width;
    uint32_t height;
} sb_texture;

/* Borrowed view of the flattened scene (valid only for the duration of sb_set_scene()).
 * Mirrors the getters the OptiX backend reads: getVertices/getIndices/getMeshes/getCurves/
 * getCurvesPoint/getCurvesWidths/getCurvesVertexCounts/getInstances/getLights
 * (OptixRender.cpp:1117-1189, 388-496, 719-825). */
typedef struct sb_scene_view
{
    const sb_vertex* vertices;
    uint64_t num_vertices;
    const uint32_t* indices;
    uint64_t num_indices;
    const sb_mesh* meshes;
    uint32_t num_meshes;
    const sb_curve* curves;
    uint32_t num_curves;
    const float* curve_points; /* 3 floats per point */
    uint64_t num_curve_points;
    const float* curve_widths; /* 1 float per point: the RADIUS at that control point, indexed like
                                  curve_points (BasisCurves.cpp:205-222 already halves the USD widths;
                                  OptiX reads widthBuffers with the point index, OptixRender.cpp:273-279) */
    uint64_t num_curve_widths;
    const uint32_t* curve_vertex_counts;
    uint64_t num_curve_vertex_counts;
    const sb_instance* instances;
    uint32_t num_instances;
    const sb_light* lights;
    uint32_t num_lights;
    const sb_material* materials;
    uint32_t num_materials;
    const sb_texture* textures; /* may be NULL */
    uint32_t num_textures;
} sb_scene_view;

/* ------------------------------------------------------------------------------------------
 * Settings: POD mirror of the SettingsManager keys the hot path reads (SURVEY.md section 5;
 * OptixRender.cpp:910-1004).  Field comments give the reference key.
 * ---------------------------------------------------------------------------------------- */
typedef struct sb_settings
{
    uint32_t spp;                        /* render/pt/spp            (samples per render() call) */
    uint32_t spp_total;                  /* render/pt/sppTotal       (== Params.maxSampleCount)  */
    uint32_t depth;                      /* render/pt/depth                                      */
    uint32_t enable_acc;                 /* render/pt/enableAcc                                  */
    uint32_t rect_light_sampling_method; /* render/pt/rectLightSamplingMethod (0 uniform, 1 SphQuad) */
    uint32_t debug;                      /* render/pt/debug (0 off, 1 normals, 2 diffuse AOV, 3 specular AOV) */
    float shadow_ray_tmin;               /* render/pt/dev/shadowRayTmin                          */
    float material_ray_tmin;             /* render/pt/dev/materialRayTmin                        */
    uint32_t tonemapper_type;            /* render/pt/tonemapperType (0 none, 1 Reinhard, 2 ACES, 3 Filmic) */
    float gamma;                         /* render/post/gamma (<=0: off)                         */
    float film_iso;                      /* render/post/tonemapper/filmIso                       */
    float cm2_factor;                    /* render/post/tonemapper/cm2_factor                    */
    float f_stop;                        /* render/post/tonemapper/fStop                         */
    float shutter_speed;                 /* render/post/tonemapper/shutterSpeed                  */
    /* --- extensions (no reference key) --- */
    uint32_t sample_offset; /* multi-GPU sample-stride sharding: this context renders sample   */
    uint32_t sample_stride; /* indices offset + k*stride (k = 0,1,...); default 0 / 1           */
    uint32_t reserved[4];
} sb_settings;

/* Defaults of src/hdRunner/main.cpp:510-542 (depth 4, spp 1, sppTotal 64, enableAcc 1, ISO 100,
 * cm2 1, fStop 4, shutter 100, gamma 2.4, tonemapper 0, tmins 0). */
void sb_settings_default(sb_settings* s);

/* ------------------------------------------------------------------------------------------
 * Context  (replaces class OptiXRender : oka::Render, OptixRender.h:63-161)
 * ---------------------------------------------------------------------------------------- */
typedef struct sb_ctx sb_ctx;
typedef struct sb_buffer sb_buffer;

typedef struct sb_device_cfg
{
    int32_t device;          /* CUDA device ordinal (reference: cudaFree(0), OptixRender.cpp:166) */
    uint32_t max_batch_paths; /* wavefront batch size cap, 0 = default (32 Mi paths)           */
    uint32_t flags;          /* SB_CFG_* */
    uint32_t curve_split;    /* BVH spans per cubic curve segment (tighter boxes for hair), 0 = default (8) */
} sb_device_cfg;

enum
{
    SB_CFG_TRAVERSAL_STATS = 1u, /* run the instrumented traversal kernels (counts nodes/prims per ray) */
    SB_CFG_STAGE_TIMERS = 2u,    /* bracket every kernel launch with CUDA events (per-stage device time) */
    SB_CFG_FUSED_SMALL = 4u      /* scenes of a few BVH nodes run whole paths in one kernel (path state in registers,
                                    path regeneration) instead of the wavefront queues.  Same images bit for bit;
                                    measured slower on B200, kept for A/B measurements */
};

/* RenderFactory::createRender(RenderType::eCompute) + Render::init()
 * (render.cpp:10-26, render.h:13, OptixRender.cpp:1059-1105). */
sb_result sb_create(const sb_device_cfg* cfg, sb_ctx** out_ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream) instead
 * of the context's own one, so that the caller's events and collectives are stream-ordered with the
 * render (reference: one created stream, OptixRender.cpp:168-170).  NULL is CUDA's legacy default
 * stream; SB_STREAM_PRIVATE restores the context's private non-blocking stream. */
#define SB_STREAM_PRIVATE ((void*)(intptr_t)-1)
sb_result sb_set_stream(sb_ctx* ctx, void* cuda_stream);
/* OptiXRender::~OptiXRender (OptixRender.cpp:159-161; the reference leaks, we free). */
void sb_destroy(sb_ctx* ctx);
/* Last error text of this context (or of sb_create when ctx == NULL). Never NULL. */
const char* sb_last_error(const sb_ctx* ctx);

/* Render::setScene + the frame-0 uploads and acceleration-structure build of
 * OptiXRender::render (OptixRender.cpp:876-888): copies the arrays to the device, flattens
 * instances to world space and builds the compressed 8-wide BVHs on the device.
 * Resets accumulation. */
sb_result sb_set_scene(sb_ctx* ctx, const sb_scene_view* scene);

/* Camera (oka::Camera, camera.h:16-95).  `view` is the glm view matrix storage (column-major
 * float[16]) as produced by Camera::updateViewMatrix (camera.cpp:10-23); fov_y_deg is
 * Camera::fov.  render() re-derives the projection from the OUTPUT buffer's aspect ratio
 * exactly like OptixRender.cpp:895-897 + camera.cpp:61-131 (near/far do not affect rays).
 * A changed camera resets accumulation (OptixRender.cpp:903-908). */
sb_result sb_set_camera(sb_ctx* ctx, const float view[16], float fov_y_deg);
/* Raw alternative: Params.clipToView / Params.viewToWorld, row-major (OptixRender.cpp:953-954).
 * Overrides the projection derivation until sb_set_camera is called again. */
sb_result sb_set_camera_matrices(sb_ctx* ctx, const float clip_to_view[16], const float view_to_world[16]);

/* SettingsManager snapshot for the next render() (OptixRender.cpp:910-1004).  Changes that
 * reset accumulation in the reference (rectLightSamplingMethod, enableAcc, sppTotal shrink)
 * do so here as well. */
sb_result sb_set_settings(sb_ctx* ctx, const sb_settings* settings);

/* SharedContext::mSubframeIndex = 0 (OptixRender.cpp:931-934). */
sb_result sb_reset_accumulation(sb_ctx* ctx);
/* SharedContext::mSubframeIndex (common.h:22-28): samples accumulated so far. */
uint32_t sb_subframe_index(const sb_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Buffers  (replaces class OptixBuffer : oka::Buffer, buffer.h:23-88, OptixBuffer.cpp)
 * ---------------------------------------------------------------------------------------- */
enum
{
    SB_FORMAT_UNSIGNED_BYTE4 = 0,
    SB_FORMAT_FLOAT4 = 1,
    SB_FORMAT_FLOAT3 = 2
}; /* oka::BufferFormat, buffer.h:9-14 */

/* Render::createBuffer(BufferDesc) (OptixRender.cpp:1107-1115). */
sb_result sb_buffer_create(sb_ctx* ctx, uint32_t width, uint32_t height, uint32_t format, sb_buffer** out);
void sb_buffer_destroy(sb_buffer* buf);
/* Buffer::resize (OptixBuffer.cpp:20-35). */
sb_result sb_buffer_resize(sb_buffer* buf, uint32_t width, uint32_t height);
/* Buffer::map: blocking device->host copy into the buffer's pinned host mirror
 * (OptixBuffer.cpp:37-43).  Unlike the reference (which returns nullptr and makes callers use
 * getHostPointer()), the host pointer is returned through *host_ptr when non-NULL. */
sb_result sb_buffer_map(sb_buffer* buf, void** host_ptr);
sb_result sb_buffer_unmap(sb_buffer* buf);
/* Non-blocking form for progressive display (the reference's map() blocks the app loop,
 * OptixBuffer.cpp:37-43, hdRunner/main.cpp:698-708): sb_buffer_map_async enqueues the device->host copy of
 * the buffer's current contents behind the rendering already issued and returns at once, so the next
 * sb_render() can be issued while the copy runs; sb_buffer_map_wait blocks until THAT copy has landed in
 * the pinned mirror (not until later work finishes) and returns the host pointer. */
sb_result sb_buffer_map_async(sb_buffer* buf);
sb_result sb_buffer_map_wait(sb_buffer* buf, void** host_ptr);
void* sb_buffer_host_ptr(sb_buffer* buf);   /* Buffer::getHostPointer, buffer.h:42-45 */
size_t sb_buffer_host_size(sb_buffer* buf); /* Buffer::getHostDataSize, buffer.h:46-49 */
void* sb_buffer_device_ptr(sb_buffer* buf); /* OptixBuffer::getNativePtr / Render::getNativeDevicePtr */
uint32_t sb_buffer_width(const sb_buffer* buf);
uint32_t sb_buffer_height(const sb_buffer* buf);

/* ------------------------------------------------------------------------------------------
 * Rendering  (replaces OptiXRender::render(Buffer*), OptixRender.cpp:874-1057)
 * ---------------------------------------------------------------------------------------- */

/* One reference render() call: min(spp, sppTotal - subframe) samples per pixel, accumulated,
 * output image = current estimate (+ tonemap/gamma post).  Asynchronous on the context's
 * stream; sb_buffer_map() / sb_synchronize() wait. */
sb_result sb_render(sb_ctx* ctx, sb_buffer* output);
/* Equivalent to calling sb_render() `iterations` times (the reference app loop,
 * hdRunner/main.cpp:663-708), but lets the backend pipeline several iterations per wavefront
 * batch.  Stops early at sppTotal like the reference. */
sb_result sb_render_iterations(sb_ctx* ctx, sb_buffer* output, uint32_t iterations);
/* cudaDeviceSynchronize of OptixRender.cpp:1012. */
sb_result sb_synchronize(sb_ctx* ctx);

/* Multi-GPU (new in this backend): the shardable state is
 * S = sum_k T(L_k) (float4 per pixel, w = 0) plus the sample count.  A harness all-reduces the
 * S buffers of the participating contexts (NCCL sum) and calls sb_resolve() with the global
 * sample count to write T^-1(S/n) (+post) into `output`. */
void* sb_accum_device_ptr(sb_ctx* ctx, uint64_t* num_floats);
sb_result sb_resolve(sb_ctx* ctx, sb_buffer* output, uint32_t total_samples);

/* Sample-sharded rendering over several GPUs behind the ABI (new: OptiXRender is single-GPU, OptixRender.cpp:163-189).
 * One sb_ctx per GPU = one rank; every rank holds a scene/BVH replica and renders the global sample indices
 * rank, rank + world, ... of every pixel with the reference's sampler indices unchanged (Params.maxSampleCount stays
 * the global sppTotal, RandomSampler.h:130-137).  The only exchange is one NCCL sum of S per call.
 *   sb_comm_get_unique_id : ncclGetUniqueId; called once (by rank 0), the SB_COMM_ID_BYTES bytes are handed to the other
 *                           ranks by the host application (file, socket, MPI, ...)
 *   sb_comm_init          : ncclCommInitRank on the context's device; collective over all ranks.  Sets the context's
 *                           sample_offset / sample_stride to rank / world (later sb_set_settings calls must keep them)
 *   sb_render_sharded     : sb_render_iterations for this rank's share of `iterations` samples PER RANK, then
 *                           ncclAllReduce(sum) of S on the context's stream into a separate buffer and the resolve
 *                           T^-1(S_global / n_global) (+post) into `output` on every rank.  S itself stays the rank's own
 *                           partial sum, so the call can be repeated for progressive display.  Collective: every rank
 *                           must issue the same sequence of calls.  Needs spp == 1, enableAcc, debug == 0.
 * libnccl.so.2 is dlopen()ed by the first of these calls; single-GPU use never loads it. */
#define SB_COMM_ID_BYTES 128
sb_result sb_comm_get_unique_id(void* id_out);
sb_result sb_comm_init(sb_ctx* ctx, const void* id, uint32_t rank, uint32_t world);
sb_result sb_comm_destroy(sb_ctx* ctx);
uint32_t sb_comm_world(const sb_ctx* ctx);
sb_result sb_render_sharded(sb_ctx* ctx, sb_buffer* output, uint32_t iterations);
/* How this context's group exchanges S: "NVLS multimem: fused all-reduce + resolve kernel" (one kernel: the NVSwitch
 * sums the ranks' accumulation buffers in flight -- multimem.ld_reduce / multimem.st over NCCL symmetric windows --
 * and the sums are resolved in place; needs NCCL >= 2.28 and NVLS hardware) or "ncclAllReduce + k_resolve (why)".
 * STRELKA_B200_NVLS=0 forces the latter.  Never NULL. */
const char* sb_comm_exchange_path(const sb_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Counters (new: the reference has no ray counters)
 * ---------------------------------------------------------------------------------------- */
typedef struct sb_counters
{
    uint64_t paths;          /* camera paths started */
    uint64_t radiance_rays;  /* closest-hit rays traced */
    uint64_t shadow_rays;    /* any-hit rays traced */
    /* SB_CFG_TRAVERSAL_STATS only (else 0): totals over the closest-hit (radiance) rays ... */
    uint64_t nodes_visited;  /* 80 B CWBVH nodes fetched */
    uint64_t tris_tested;    /* 48 B triangle records fetched */
    uint64_t segs_tested;    /* 64 B curve-segment records fetched */
    uint64_t stack_overflows;
    /* ... and over the any-hit (shadow) rays */
    uint64_t nodes_visited_shadow;
    uint64_t tris_tested_shadow;
    uint64_t segs_tested_shadow;
    /* geometry / build info */
    uint64_t num_triangles;  /* world-space triangles in the BVH */
    uint64_t num_segments;   /* curve segments in the BVH */
    uint64_t bvh_nodes_tri;
    uint64_t bvh_nodes_curve;
    double build_ms;         /* last sb_set_scene: upload + flatten + BVH build */
    double render_ms;        /* device time of the last render call (CUDA events) */
    uint64_t kernel_launches; /* kernels launched by render/resolve calls since the last reset */
    /* SB_CFG_STAGE_TIMERS only: device milliseconds and launch counts per stage since the last reset,
     * order: raygen, extend, shade, shadow, accumulate, resolve, fused path kernel, (reserved) */
    double stage_ms[8];
    uint64_t stage_launches[8];
    /* levels of the two wide BVHs; the builder refuses trees deeper than the traversal stack (sb_set_scene fails) */
    uint64_t bvh_depth_tri;
    uint64_t bvh_depth_curve;
    /* last sb_render_sharded of a group: device time of the exchange (all-reduce + resolve), and whether it was the
     * fused NVLS kernel (1) or ncclAllReduce + k_resolve (0) */
    double exchange_ms;
    uint64_t exchange_nvls;
} sb_counters;

sb_result sb_get_counters(sb_ctx* ctx, sb_counters* out); /* synchronizes */
sb_result sb_reset_counters(sb_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Test hooks: run the device implementations of the path's building blocks on caller-supplied
 * inputs so parity tests can compare them with the oracle / golden vectors one function at a
 * time.  All pointers are HOST pointers; the hooks copy, launch and copy back.
 * ---------------------------------------------------------------------------------------- */

/* RandomSampler.h:130-137,221-226: out[i] = random<dim[i]>(initSampler(x,y,.,sample,max,52), depth) */
sb_result sb_test_sampler(sb_ctx* ctx, uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample,
                          const uint32_t* max_samples, const uint32_t* depth, const uint32_t* dim, float* out);

/* Lights.h sampling: for each i sample light[i] (type decides the routine; rect uses `method`)
 * from hit_point[i] with u[i]; out 12 floats per item:
 * pointOnLight.xyz, pdf, normal.xyz, area, L.xyz, distToLight  (LightSampleData, Lights.h:16-26) */
sb_result sb_test_light_sample(sb_ctx* ctx, uint32_t n, const sb_light* lights, const float* hit_points,
                               const float* u, uint32_t method, float* out);

/* offset_ray (closest_hit.cu:218-233): out[i] = the ray origin for hit point p[i] pushed along normal[i] (3 floats each) */
sb_result sb_test_offset_ray(sb_ctx* ctx, uint32_t n, const float* p, const float* normal, float* out);

/* tex::lookup_float4 (texture_support_cuda.h:287-314) of texture `index` (0-based) of the current scene at n (u, v)
 * pairs -> n rgba quadruples: the hardware-filtered lookups the shade kernel performs for textured materials. */
sb_result sb_test_texture(sb_ctx* ctx, uint32_t index, uint32_t n, const float* uv, float* out);

/* The material protocol of closest_hit.cu:474-545 on caller inputs: per item mdlcode_sample(k1, xi) and
 * mdlcode_evaluate(k1, k2) of material `m` (the device implementations the shade kernel runs).
 * in : 19 floats per item: shading normal[3], geometric normal[3], tangent_u[3], k1[3], xi[4], k2 for evaluate[3]
 * out: 15 floats per item: sample k2[3], bsdf_over_pdf[3], pdf, event_type bits; evaluate bsdf_diffuse[3], bsdf_glossy[3]
 *      (cosine included), pdf */
sb_result sb_test_bsdf(sb_ctx* ctx, const sb_material* m, uint32_t n, const float* in, float* out);

/* Trace arbitrary rays against the current scene.  rays: 8 floats each (ox,oy,oz,tmin,dx,dy,dz,tmax).
 * mode 0: closest hit with mask 255; mode 1: any hit with mask 3 (shadow) -- one ray per thread, whole traversal.
 * mode 2 / 3: the same two queries through the PRODUCTION path: the rays are written into the wavefront queues and
 * traced by the stage launchers sb_render uses for secondary rays (the persistent dynamic-fetch k_extend / k_shadow
 * kernels on every scene of more than a few nodes).  Mode 2 requires one common tmin and tmax >= 1e16, like the
 * reference's radiance rays (OptixRender.cu:120-129).
 * hits: per ray {float t; float u; float v; uint32 prim; uint32 instance; uint32 kind}
 * kind 0 miss, 1 triangle, 2 curve (u = curve parameter). */
typedef struct sb_hit
{
    float t, u, v;
    uint32_t prim;
    uint32_t instance;
    uint32_t kind;
} sb_hit;
sb_result sb_test_trace(sb_ctx* ctx, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits);

#ifdef __cplusplus
}
#endif
#endif /* SB_API_H */
