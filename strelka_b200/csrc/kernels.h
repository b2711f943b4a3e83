// Launchers of the kernels in kernels.cu (host-callable).
#pragma once
#include "wavefront.cuh"
#include <vector>

namespace sb
{

enum Stage
{
    kStageRaygen = 0,
    kStageExtend,
    kStageShade,
    kStageShadow,
    kStageAccumulate,
    kStageResolve,
    kStagePathFused, // whole paths in one kernel (small scenes)
    kNumStages
};

// Optional per-launch CUDA-event bracketing (SB_CFG_STAGE_TIMERS): events are recorded on the launch
// stream and only read back (after a synchronize) by collect().
struct StageTimer
{
    struct Pending
    {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<cudaEvent_t> pool;
    std::vector<Pending> pending;
    double ms[kNumStages] = {};
    uint64_t launches[kNumStages] = {};
    cudaEvent_t get();
    void collect();
    void reset();
    void release();
};

struct LaunchCfg
{
    cudaStream_t stream;
    int numSms;
    StageTimer* timer; // may be null
    uint64_t* launchCount; // may be null
    bool fusedSmall; // scenes of a few BVH nodes run the single-kernel path tracer
};

uint32_t* upload_sobol_table(cudaStream_t stream); // also returns the device copy of the byte-sliced tables
// one wavefront batch: raygen, then per bounce extend -> shade -> shadow
void launch_wavefront_batch(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Q, bool stats);
// accumulation targets of one frame: S (beauty sum of T(L)), direct (non-accumulated launch result / running linear sum
// of a multi-batch launch), the two AOV sums and their multi-batch scratch (may be null when a launch fits one batch)
struct AccumTargets
{
    float4 *S, *direct, *aovD, *aovS, *scrD, *scrS;
};
// launchSamples / batchFlags: see accumulate_pixel (launches of more samples than one wavefront batch holds)
void launch_accumulate(const LaunchCfg& cfg, const FrameParams& P, const Queues& Q, const AccumTargets& A, uint32_t mode, uint32_t subframe,
                       uint32_t launchSamples, uint32_t batchFlags);
void launch_resolve(const LaunchCfg& cfg, const float4* S, void* out, uint32_t npix, uint32_t n, const float exposure[3], uint32_t tonemapper,
                    float gamma, uint32_t format);
void launch_copy_image(const LaunchCfg& cfg, const float4* src, void* out, uint32_t npix, uint32_t format, const float exposure[3],
                       uint32_t tonemapper, float gamma);
void launch_test_sampler(const LaunchCfg& cfg, uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxs,
                         const uint32_t* depth, const uint32_t* dim, float* out);
void launch_test_light_sample(const LaunchCfg& cfg, uint32_t n, const sb_light* lights, const float* hp, const float* u, uint32_t method, float* out);
void launch_test_trace(const LaunchCfg& cfg, const SceneDev& S, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits);
void launch_test_offset_ray(const LaunchCfg& cfg, uint32_t n, const float* p, const float* nrm, float* out);
void launch_test_texture(const LaunchCfg& cfg, const SceneDev& S, uint32_t index1, uint32_t n, const float* uv, float* out);
void launch_test_bsdf(const LaunchCfg& cfg, const sb_material& m, uint32_t n, const float* in, float* out);
// the stages of one bounce as launch_wavefront_batch dispatches them (persistent kernels for secondary rays of non-tiny scenes)
void launch_extend_stage(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Q, uint32_t depth, bool stats);
void launch_shadow_stage(const LaunchCfg& cfg, const SceneDev& S, const Queues& Q, uint32_t depth, bool stats);
// sb_test_trace modes 2 (closest hit) / 3 (any hit): caller rays through the real queues and the stage launchers above
void launch_test_trace_production(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Q, const uint32_t* instTriFirst,
                                  uint32_t n, const float* rays, uint32_t mode, sb_hit* hits, bool stats);

} // namespace sb
