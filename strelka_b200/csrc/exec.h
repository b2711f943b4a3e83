// Execution policy used by the BVH builder: the same orchestration code (bvh_build.h) runs
//   * as CUDA kernels + cub primitives on a stream        (ExecCuda -- the product), or
//   * as serial loops on the host                         (ExecHost -- tests/emul only, g++),
// selected at compile time by SB_HOST_EMUL.  The serial variant exists solely so that the CPU
// test-suite can exercise the builder/traversal logic without a GPU; it is never compiled into
// libstrelka_b200.so.
#pragma once
#include "hd.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#if !defined(SB_HOST_EMUL)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>
#endif

namespace sb
{

#if !defined(SB_HOST_EMUL)

#define SB_CUDA_CHECK(x)                                                                                              \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t e_ = (x);                                                                                         \
        if (e_ != cudaSuccess)                                                                                        \
            throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_));                                \
    } while (0)

template <class F>
__global__ void __launch_bounds__(256) k_parallel_for(size_t n, F f)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        f(i);
}

struct ExecCuda
{
    cudaStream_t stream = nullptr;
    int numSms = 148;
    void* cubTemp = nullptr;
    size_t cubTempBytes = 0;

    // Stream-ordered allocation from the device's default memory pool (its release threshold is raised by sb_create,
    // so freed blocks stay cached): a rebuild re-uses the blocks of the previous one instead of paying cudaMalloc /
    // cudaFree for every temporary -- measured on the 10 M-triangle scene those calls were up to 550 ms of a build
    // whose kernels take 80 ms.  Buffers that outlive the build (nodes, records) are later released with cudaFree,
    // which accepts pool allocations.
    template <class T>
    T* alloc(size_t n)
    {
        void* p = nullptr;
        if (n == 0)
            n = 1;
        cudaError_t e = cudaMallocAsync(&p, n * sizeof(T), stream);
        if (e != cudaSuccess)
        {
            cudaGetLastError();
            throw std::bad_alloc();
        }
        return static_cast<T*>(p);
    }
    void free(void* p)
    {
        if (p)
            cudaFreeAsync(p, stream);
    }
    template <class F>
    void pfor(size_t n, F f)
    {
        if (n == 0)
            return;
        const size_t blocks = (n + 255) / 256;
        const size_t cap = size_t(numSms) * 16; // grid-stride: a multiple of the SM count
        k_parallel_for<<<unsigned(blocks < cap ? blocks : cap), 256, 0, stream>>>(n, f);
        SB_CUDA_CHECK(cudaGetLastError());
    }
    void reserve_temp(size_t bytes)
    {
        if (bytes > cubTempBytes)
        {
            if (cubTemp)
                cudaFreeAsync(cubTemp, stream);
            SB_CUDA_CHECK(cudaMallocAsync(&cubTemp, bytes, stream));
            cubTempBytes = bytes;
        }
    }
    void exclusive_scan_u64(const uint64_t* in, uint64_t* out, size_t n)
    {
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, int(n), stream);
        reserve_temp(bytes);
        SB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(cubTemp, bytes, in, out, int(n), stream));
    }
    void sort_pairs_u64_u32(const uint64_t* keysIn, uint64_t* keysOut, const uint32_t* valsIn, uint32_t* valsOut, size_t n)
    {
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, keysIn, keysOut, valsIn, valsOut, int(n), 0, 64, stream);
        reserve_temp(bytes);
        SB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cubTemp, bytes, keysIn, keysOut, valsIn, valsOut, int(n), 0, 64, stream));
    }
    template <class T>
    T read(const T* p)
    {
        T v;
        SB_CUDA_CHECK(cudaMemcpyAsync(&v, p, sizeof(T), cudaMemcpyDeviceToHost, stream));
        SB_CUDA_CHECK(cudaStreamSynchronize(stream));
        return v;
    }
    template <class T>
    void write(T* p, const T& v)
    {
        SB_CUDA_CHECK(cudaMemcpyAsync(p, &v, sizeof(T), cudaMemcpyHostToDevice, stream));
        SB_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    void zero(void* p, size_t bytes)
    {
        SB_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream));
    }
    void release()
    {
        if (cubTemp)
            cudaFreeAsync(cubTemp, stream);
        cubTemp = nullptr;
        cubTempBytes = 0;
    }
    // STRELKA_B200_BUILD_TRACE=1: synchronise at every phase boundary of the builder and print the phase times
    bool trace = false;
    std::chrono::steady_clock::time_point traceT0;
    void mark(const char* what, long long n = -1)
    {
        if (!trace)
            return;
        cudaStreamSynchronize(stream);
        const auto now = std::chrono::steady_clock::now();
        if (what)
            fprintf(stderr, "[sb build] %-28s %9.3f ms%s\n", what, std::chrono::duration<double, std::milli>(now - traceT0).count(),
                    n >= 0 ? (" (" + std::to_string(n) + ")").c_str() : "");
        traceT0 = now;
    }
};
using Exec = ExecCuda;
#define SB_LAMBDA [=] __host__ __device__

#else // ---- SB_HOST_EMUL ------------------------------------------------------------------------------

struct ExecHost
{
    template <class T>
    T* alloc(size_t n)
    {
        return static_cast<T*>(std::calloc(n ? n : 1, sizeof(T)));
    }
    void free(void* p) { std::free(p); }
    void mark(const char*, long long = -1) {}
    template <class F>
    void pfor(size_t n, F f)
    {
        for (size_t i = 0; i < n; ++i)
            f(i);
    }
    void exclusive_scan_u64(const uint64_t* in, uint64_t* out, size_t n)
    {
        uint64_t acc = 0;
        for (size_t i = 0; i < n; ++i)
        {
            const uint64_t v = in[i];
            out[i] = acc;
            acc += v;
        }
    }
    void sort_pairs_u64_u32(const uint64_t* keysIn, uint64_t* keysOut, const uint32_t* valsIn, uint32_t* valsOut, size_t n)
    {
        std::vector<uint32_t> idx(n);
        for (size_t i = 0; i < n; ++i)
            idx[i] = uint32_t(i);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keysIn[a] < keysIn[b]; });
        for (size_t i = 0; i < n; ++i)
        {
            keysOut[i] = keysIn[idx[i]];
            valsOut[i] = valsIn[idx[i]];
        }
    }
    template <class T>
    T read(const T* p) { return *p; }
    template <class T>
    void write(T* p, const T& v) { *p = v; }
    void zero(void* p, size_t bytes) { std::memset(p, 0, bytes); }
    void release() {}
};
using Exec = ExecHost;
#define SB_LAMBDA [=]

#endif

} // namespace sb
