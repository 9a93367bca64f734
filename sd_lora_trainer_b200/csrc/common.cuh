// Shared helpers for the HBM-bound kernels: 16-byte bf16 vectors, warp/block reductions, error plumbing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

namespace b200 {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local message, integer status across the C ABI) -----------------
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define B200_CHECK_ARG(cond, ...)                    \
    do {                                             \
        if (!(cond)) return set_error(2, __VA_ARGS__); \
    } while (0)

#define B200_CHECK_LAUNCH(name)                                                           \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) return set_error(3, "%s: %s", name, cudaGetErrorString(e__)); \
        count_launch();                                                                   \
    } while (0)

constexpr int kNumSMs = 148;

}  // namespace b200
#include <cuda.h>
namespace b200 {
// 4-D bf16 TMA tensor map with 128B swizzle (gemm_host.cu); dims/strides innermost first, strides in elements.
int encode_map(CUtensorMap* map, const void* ptr, const long long dims[4], const long long strides[3], const int box[4]);
// general form: elem_bytes 2 (bf16) | 4 (fp32), swizzle_bytes 128 | 64 | 32 | 0
int encode_map_ex(CUtensorMap* map, const void* ptr, int elem_bytes, const long long dims[4], const long long strides[3],
                  const int box[4], int swizzle_bytes);


// ---- programmatic dependent launch -------------------------------------------------------------
// Every kernel of this library is launched with programmaticStreamSerialization and begins with
// pdl_launch(); pdl_wait(): the next kernel's CTAs may be scheduled (and run their prologue) while this grid drains,
// and no kernel touches global memory before its predecessor has fully completed and flushed.
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// B200_PDL=0 (profiling aid): launch without the programmatic-serialization attribute, so kernels run strictly one after
// the other and a profiler's per-kernel durations no longer include the time spent in griddepcontrol.wait.
inline int pdl_enabled() {
    static const int on = getenv("B200_PDL") ? atoi(getenv("B200_PDL")) : 1;
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same, for a kernel whose CTAs form clusters of `cluster_x` along x (CTA pairs for cta_group::2).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      unsigned cluster_x, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- bf16 helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float bfr(float x) {   // round-trip through bf16 (torch's per-op rounding)
    return __bfloat162float(__float2bfloat16_rn(x));
}

struct alignas(16) bf16x8 {
    __nv_bfloat162 h[4];
};

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
    const bf16x8 v = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __bfloat162float(v.h[i].x);
        f[2 * i + 1] = __bfloat162float(v.h[i].y);
    }
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
    bf16x8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<bf16x8*>(p) = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// Activations on the hardware special-function path.  Inputs and outputs are bf16 (relative resolution 4e-3); the kernels
// that use these were bound by the libm forms (expf / erff / IEEE division: ~40-50 instructions per element, measured
// 1.4 TB/s in gn_reduce_kernel<bwd, silu> - profiles/r02c), so:  sigmoid = rcp(1 + ex2(-x log2e)) with MUFU ex2 / rcp
// (~2 ulp), erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7) sharing ONE exponential with the Gaussian pdf.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float dsilu_f(float x) {
    const float s = sigmoid_f(x);
    return s * (1.f + x * (1.f - s));
}
// Phi(x) (standard normal cdf) and phi(x) (pdf) from one exponential
__device__ __forceinline__ void normal_cdf_pdf(float x, float& cdf, float& pdf) {
    const float e = __expf(-0.5f * x * x);                       // exp(-u^2) with u = |x| / sqrt(2)
    const float u = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.f, fmaf(0.3275911f, u, 1.f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    const float tail = 0.5f * (p * t) * e;                       // 0.5 * erfc(u)
    cdf = x >= 0.f ? 1.f - tail : tail;
    pdf = 0.39894228040143267794f * e;
}
__device__ __forceinline__ float gelu_f(float x) {
    float cdf, pdf;
    normal_cdf_pdf(x, cdf, pdf);
    return x * cdf;
}
__device__ __forceinline__ float dgelu_f(float x) {
    float cdf, pdf;
    normal_cdf_pdf(x, cdf, pdf);
    return fmaf(x, pdf, cdf);
}
// gelu(x) and gelu'(x) together (GEGLU backward needs both for the same gate value)
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
    float cdf, pdf;
    normal_cdf_pdf(x, cdf, pdf);
    g = x * cdf;
    dg = fmaf(x, pdf, cdf);
}

inline int grid_for(long long work_items, int threads, int max_blocks = kNumSMs * 16) {
    long long b = (work_items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return static_cast<int>(b);
}

}  // namespace b200
