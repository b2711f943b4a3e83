// Launchers of the kernels in kernels.cu (host-callable).
#pragma once
#include "wavefront.cuh"

namespace sb
{

struct LaunchCfg
{
    cudaStream_t stream;
    int numSms;
};

void upload_sobol_table(cudaStream_t stream);
// one wavefront batch: raygen, then per bounce extend -> shade -> shadow
void launch_wavefront_batch(const LaunchCfg& cfg, const FrameParams& P, const SceneDev& S, const Queues& Q, bool stats);
void launch_accumulate(const LaunchCfg& cfg, const FrameParams& P, const Queues& Q, float4* S, float4* direct, uint32_t mode, uint32_t subframe);
void launch_resolve(const LaunchCfg& cfg, const float4* S, void* out, uint32_t npix, uint32_t n, const float exposure[3], uint32_t tonemapper,
                    float gamma, uint32_t format);
void launch_copy_image(const LaunchCfg& cfg, const float4* src, void* out, uint32_t npix, uint32_t format);
void launch_test_sampler(const LaunchCfg& cfg, uint32_t n, const uint32_t* x, const uint32_t* y, const uint32_t* sample, const uint32_t* maxs,
                         const uint32_t* depth, const uint32_t* dim, float* out);
void launch_test_light_sample(const LaunchCfg& cfg, uint32_t n, const sb_light* lights, const float* hp, const float* u, uint32_t method, float* out);
void launch_test_trace(const LaunchCfg& cfg, const SceneDev& S, uint32_t n, const float* rays, uint32_t mode, sb_hit* hits);

} // namespace sb
