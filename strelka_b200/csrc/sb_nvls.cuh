// Fused all-reduce + resolve over NVLink SHARP (NVLS multimem) for the sample-sharded render (SURVEY.md 8e stretch):
// ONE kernel replaces ncclAllReduce(S) + k_resolve.  The accumulation buffers S of all ranks live in an NCCL symmetric
// window (ncclMemAlloc + ncclCommWindowRegister), so the NVSwitch can reduce them in flight:
//   phase 1  CTA j of rank r reduces sub-slice r of pixel chunk j with multimem.ld_reduce (the switch adds the N ranks'
//            values and returns the sum) and broadcasts the sum into every rank's copy of G with multimem.st
//   barrier  CTA j of every rank (NCCL device API: ncclLsaBarrierSession over the multicast inbox)
//   phase 2  CTA j resolves chunk j from its LOCAL G: image = T^-1(G / n) (+ post-process) into the output buffer
// Every GPU's S is read once over NVLink and every sum is written once per rank; no host round trip, no separate
// resolve pass.  Needs NCCL >= 2.28 (device API) and NVLS-capable hardware; sb_render_sharded falls back to
// ncclAllReduce + k_resolve otherwise (or with STRELKA_B200_NVLS=0).
#pragma once
#if SB_HAVE_NCCL_DEVICE
#include <nccl_device.h>

namespace sb
{

constexpr int kNvlsBlock = 256;
constexpr int kNvlsCtasPerSm = 2;

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// winS / winG: symmetric windows over the ranks' S (partial sums) and G (global sum) buffers, npix float4 each
__global__ void __launch_bounds__(kNvlsBlock) k_allreduce_resolve_nvls(ncclDevComm dc, ncclWindow_t winS, ncclWindow_t winG, const float4* localG, void* out,
                                                                      uint32_t npix, uint32_t nTotal, float3 e, uint32_t tonemapper, float gamma,
                                                                      uint32_t format)
{
    ncclCoopCta cta;
    ncclLsaBarrierSession<ncclCoopCta> bar(cta, dc, ncclTeamTagLsa(), blockIdx.x, /*multimem=*/true);
    // every rank's accumulation kernels have finished (stream order on each GPU) before anybody reads across NVLink
    bar.sync(cta, cuda::memory_order_acq_rel);
    const uint32_t chunk = (npix + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = min(blockIdx.x * chunk, npix), c1 = min(c0 + chunk, npix);
    const uint32_t sub = (c1 - c0 + uint32_t(dc.lsaSize) - 1) / uint32_t(dc.lsaSize);
    const uint32_t s0 = min(c0 + uint32_t(dc.lsaRank) * sub, c1), s1 = min(s0 + sub, c1);
    const float4* mcS = static_cast<const float4*>(ncclGetLsaMultimemPointer(winS, 0, dc));
    float4* mcG = static_cast<float4*>(ncclGetLsaMultimemPointer(winG, 0, dc));
    for (uint32_t i = s0 + threadIdx.x; i < s1; i += blockDim.x)
        multimem_st(mcG + i, multimem_ld_reduce_add(mcS + i));
    // the sums of chunk j have landed in every rank's G before anybody resolves it
    bar.sync(cta, cuda::memory_order_acq_rel);
    for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x)
    {
        const float4 c = resolve_pixel(localG[i], nTotal, e, tonemapper, gamma);
        if (format == SB_FORMAT_FLOAT4)
        {
            reinterpret_cast<float4*>(out)[i] = c;
        }
        else if (format == SB_FORMAT_FLOAT3)
        {
            float* o = reinterpret_cast<float*>(out) + 3 * size_t(i);
            o[0] = c.x;
            o[1] = c.y;
            o[2] = c.z;
        }
        else
        {
            uchar4 q;
            q.x = (unsigned char)(saturate(c.x) * 255.0f + 0.5f);
            q.y = (unsigned char)(saturate(c.y) * 255.0f + 0.5f);
            q.z = (unsigned char)(saturate(c.z) * 255.0f + 0.5f);
            q.w = 255;
            reinterpret_cast<uchar4*>(out)[i] = q;
        }
    }
}

inline void launch_allreduce_resolve_nvls(const LaunchCfg& cfg, const ncclDevComm& dc, ncclWindow_t winS, ncclWindow_t winG, const float4* localG,
                                          void* out, uint32_t npix, uint32_t nTotal, const float exposure[3], uint32_t tonemapper, float gamma,
                                          uint32_t format, int grid)
{
    if (cfg.launchCount)
        ++*cfg.launchCount;
    k_allreduce_resolve_nvls<<<grid, kNvlsBlock, 0, cfg.stream>>>(dc, winS, winG, localG, out, npix, nTotal,
                                                                 make_float3(exposure[0], exposure[1], exposure[2]), tonemapper, gamma, format);
    SB_CUDA_CHECK(cudaGetLastError());
}

} // namespace sb
#endif // SB_HAVE_NCCL_DEVICE
