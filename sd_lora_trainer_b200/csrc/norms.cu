// GroupNorm(+SiLU) and LayerNorm, forward and input-gradient, on NHWC bf16 activations ([rows, C], C contiguous).
// HBM-bound: every pass reads/writes whole 16-byte vectors with consecutive threads on consecutive addresses.
// Affine parameters are frozen in LoRA training, so no d(gamma)/d(beta) reductions exist anywhere.
#include "common.cuh"
#include "../../include/b200_lora.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// GroupNorm.  Pass 1 writes per-block partial sums of every (image, group) - combined inside the block in a fixed order,
// no atomics, so the statistics (and with them the whole forward pass) are bit-reproducible run to run; pass 2 adds the
// blocks' partials in block order and finalises (mean, rstd); pass 3 normalises.  A block owns `k` rows x all channels per iteration so each
// thread keeps a FIXED 8-channel vector and accumulates in registers.
// ---------------------------------------------------------------------------------------------
template <bool kBackward, bool kSilu>
__global__ void gn_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                 const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                 const float* __restrict__ stats, double* __restrict__ ws, long long hw, int C,
                                 int groups, int rows_per_block) {
    pdl_launch();
    pdl_wait();
    __shared__ float4 part[1024];              // one (lo0, lo1, hi0, hi1) per thread
    const int C8 = C >> 3, cpg = C / groups;
    const int b = blockIdx.y;
    const int v = threadIdx.x % C8, roff = threadIdx.x / C8, k = blockDim.x / C8;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, hw);
    float a0[8], a1[8], gm[8], bt[8], mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    if (kBackward) {
        load8(gamma + v * 8, gm);
        load8(beta + v * 8, bt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = (v * 8 + i) / cpg;
            mean[i] = stats[(b * groups + g) * 2];
            rstd[i] = stats[(b * groups + g) * 2 + 1];
        }
    }
    const bf16* xb = x + static_cast<long long>(b) * hw * C;
    const bf16* dyb = kBackward ? dy + static_cast<long long>(b) * hw * C : nullptr;
    for (long long r = r0 + roff; r < r1; r += k) {
        float f[8];
        load8(xb + r * C + v * 8, f);
        if (!kBackward) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a0[i] += f[i];
                a1[i] += f[i] * f[i];
            }
        } else {
            float d[8];
            load8(dyb + r * C + v * 8, d);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (f[i] - mean[i]) * rstd[i];
                float dz = d[i];
                if (kSilu) dz = bfr(dz * dsilu_f(bfr(xh * gm[i] + bt[i])));
                const float t = dz * gm[i];
                a0[i] += t;
                a1[i] += t * xh;
            }
        }
    }
    // ---- deterministic combine (no atomics): a thread's 8 channels touch its first group g0 and, when the group width is
    //      not a multiple of 8, g0 + 1; partials go to shared memory and one thread per (group, statistic) sums its
    //      contributors in a fixed order; the block's sums land in ITS slot of the partial buffer ----
    if (cpg >= 8) {
        const int g0 = (v * 8) / cpg;
        float lo0 = 0.f, lo1 = 0.f, hi0 = 0.f, hi1 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((v * 8 + i) / cpg == g0) {
                lo0 += a0[i];
                lo1 += a1[i];
            } else {
                hi0 += a0[i];
                hi1 += a1[i];
            }
        }
        part[threadIdx.x] = make_float4(lo0, lo1, hi0, hi1);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) {
            const int g = i >> 1, stat = i & 1;
            const int v_lo = (g * cpg) >> 3, v_hi = ((g + 1) * cpg - 1) >> 3;
            float acc = 0.f;
            for (int vv = v_lo; vv <= v_hi; ++vv) {
                const bool is_lo = (vv * 8) / cpg == g;            // else this vector's upper channels belong to g
                for (int rl = 0; rl < k; ++rl) {
                    const float4 pv = part[rl * C8 + vv];
                    acc += is_lo ? (stat ? pv.y : pv.x) : (stat ? pv.w : pv.z);
                }
            }
            ws[(static_cast<long long>(b) * gridDim.x + blockIdx.x) * groups * 2 + i] = static_cast<double>(acc);
        }
    } else {
        // narrow groups (fewer than 8 channels: only the scaled-down test nets): a vector spans several groups, so every
        // thread publishes all 16 channel sums (<= 256 threads here: 16 KiB) and the combine walks channels
        float* vals = reinterpret_cast<float*>(part);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            vals[threadIdx.x * 16 + i] = a0[i];
            vals[threadIdx.x * 16 + 8 + i] = a1[i];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * groups; i += blockDim.x) {
            const int g = i >> 1, stat = i & 1;
            float acc = 0.f;
            for (int c = g * cpg; c < (g + 1) * cpg; ++c)
                for (int rl = 0; rl < k; ++rl) acc += vals[(rl * C8 + (c >> 3)) * 16 + stat * 8 + (c & 7)];
            ws[(static_cast<long long>(b) * gridDim.x + blockIdx.x) * groups * 2 + i] = static_cast<double>(acc);
        }
    }
}

// Sums the per-block partials [batch, splits, groups, 2] in block order (deterministic) into sums [batch, groups, 2] and,
// for the forward, finalises (mean, rstd).
__global__ void gn_finalize_kernel(const double* __restrict__ partials, double* __restrict__ sums, float* __restrict__ stats,
                                   int batch, int groups, int splits, double count, float eps, int write_stats) {
    pdl_launch();
    pdl_wait();
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // one warp per (b, g)
    const int lane = threadIdx.x & 31;
    if (i >= batch * groups) return;
    const int b = i / groups, g = i - b * groups;
    double s0 = 0.0, s1 = 0.0;
    for (int sp = lane; sp < splits; sp += 32) {                     // lane-strided, then a fixed shuffle tree: deterministic
        const double* pp = partials + ((static_cast<long long>(b) * splits + sp) * groups + g) * 2;
        s0 += pp[0];
        s1 += pp[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane != 0) return;
    sums[2 * i] = s0;
    sums[2 * i + 1] = s1;
    if (write_stats) {
        const double mean = s0 / count;
        double var = s1 / count - mean * mean;
        if (var < 0) var = 0;
        stats[2 * i] = static_cast<float>(mean);
        stats[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
}

template <bool kSilu>
__global__ void gn_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma,
                                const bf16* __restrict__ beta, const float* __restrict__ stats, bf16* __restrict__ y,
                                long long hw, int C, int groups, long long total_vec) {
    pdl_launch();
    pdl_wait();
    const int C8 = C >> 3, cpg = C / groups;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total_vec;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(idx % C8);
        const int b = static_cast<int>(idx / (hw * C8));
        float f[8], gm[8], bt[8], o[8];
        load8(x + idx * 8, f);
        load8(gamma + v * 8, gm);
        load8(beta + v * 8, bt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = (v * 8 + i) / cpg;
            const float mean = stats[(b * groups + g) * 2], rstd = stats[(b * groups + g) * 2 + 1];
            float z = (f[i] - mean) * rstd * gm[i] + bt[i];
            if (kSilu) z = silu_f(bfr(z));
            o[i] = z;
        }
        store8(y + idx * 8, o);
    }
}

template <bool kSilu>
__global__ void gn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                    const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                    const float* __restrict__ stats, const double* __restrict__ ws,
                                    const bf16* __restrict__ dres, bf16* __restrict__ dx, long long hw, int C,
                                    int groups, long long total_vec) {
    pdl_launch();
    pdl_wait();
    const int C8 = C >> 3, cpg = C / groups;
    const float inv_n = 1.f / (static_cast<float>(hw) * cpg);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total_vec;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(idx % C8);
        const int b = static_cast<int>(idx / (hw * C8));
        float f[8], d[8], gm[8], bt[8], o[8];
        load8(x + idx * 8, f);
        load8(dy + idx * 8, d);
        load8(gamma + v * 8, gm);
        load8(beta + v * 8, bt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int g = (v * 8 + i) / cpg;
            const int sg = b * groups + g;
            const float mean = stats[sg * 2], rstd = stats[sg * 2 + 1];
            const float s1 = static_cast<float>(ws[sg * 2]) * inv_n, s2 = static_cast<float>(ws[sg * 2 + 1]) * inv_n;
            const float xh = (f[i] - mean) * rstd;
            float dz = d[i];
            if (kSilu) dz = bfr(dz * dsilu_f(bfr(xh * gm[i] + bt[i])));
            o[i] = rstd * (dz * gm[i] - s1 - xh * s2);
        }
        if (dres) {
            float rr[8];
            load8(dres + idx * 8, rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = bfr(o[i]) + rr[i];
        }
        store8(dx + idx * 8, o);
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the channel dim: one warp per row, the row stays in registers (C <= 2560).
// ---------------------------------------------------------------------------------------------
constexpr int kLnMaxVec = 10;   // 10 * 32 lanes * 8 = 2560 channels

template <bool kBackward, int NV>
__global__ void layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                 const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                 float* __restrict__ stats, bf16* __restrict__ out, long long rows, int C, float eps) {
    pdl_launch();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int C8 = C >> 3;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; row < rows; row += warps) {
        float f[NV][8];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int v = lane + j * 32;
            if (v < C8) {
                load8(x + row * C + v * 8, f[j]);
#pragma unroll
                for (int i = 0; i < 8; ++i) s += f[j][i];
            }
        }
        float mean, rstd;
        if (!kBackward) {
            // affine parameters are fetched BEFORE the two reductions: their L2 round trip overlaps the shuffles instead of
            // following them (the kernel is latency-bound: one row per warp, ~14 warps per SM)
            bf16x8 gr[NV], br[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = lane + j * 32;
                if (v < C8) {
                    gr[j] = *reinterpret_cast<const bf16x8*>(gamma + v * 8);
                    br[j] = *reinterpret_cast<const bf16x8*>(beta + v * 8);
                }
            }
            mean = warp_sum(s) / C;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                if (lane + j * 32 < C8) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float d = f[j][i] - mean;
                        q += d * d;
                    }
                }
            }
            rstd = rsqrtf(warp_sum(q) / C + eps);
            if (lane == 0) {
                stats[row * 2] = mean;
                stats[row * 2 + 1] = rstd;
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = lane + j * 32;
                if (v < C8) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        o[2 * i] = (f[j][2 * i] - mean) * rstd * __bfloat162float(gr[j].h[i].x) + __bfloat162float(br[j].h[i].x);
                        o[2 * i + 1] = (f[j][2 * i + 1] - mean) * rstd * __bfloat162float(gr[j].h[i].y) + __bfloat162float(br[j].h[i].y);
                    }
                    store8(out + row * C + v * 8, o);
                }
            }
        } else {
            mean = stats[row * 2];
            rstd = stats[row * 2 + 1];
            float t[NV][8];
            float s1 = 0.f, s2 = 0.f;
            bf16x8 rres[NV];                                  // residual-branch gradient, fetched ahead of the reductions
            if (beta) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const int v = lane + j * 32;
                    if (v < C8) rres[j] = *reinterpret_cast<const bf16x8*>(beta + row * C + v * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = lane + j * 32;
                if (v < C8) {
                    float gm[8], d[8];
                    load8(gamma + v * 8, gm);
                    load8(dy + row * C + v * 8, d);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        f[j][i] = (f[j][i] - mean) * rstd;   // x-hat
                        t[j][i] = d[i] * gm[i];
                        s1 += t[j][i];
                        s2 += t[j][i] * f[j][i];
                    }
                }
            }
            s1 = warp_sum(s1) / C;
            s2 = warp_sum(s2) / C;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = lane + j * 32;
                if (v < C8) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = rstd * (t[j][i] - s1 - f[j][i] * s2);
                    if (beta) {   // backward: `beta` carries the residual-branch gradient to add
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            o[2 * i] = bfr(o[2 * i]) + __bfloat162float(rres[j].h[i].x);
                            o[2 * i + 1] = bfr(o[2 * i + 1]) + __bfloat162float(rres[j].h[i].y);
                        }
                    }
                    store8(out + row * C + v * 8, o);
                }
            }
        }
    }
}

// stats layout (floats): [batch*groups*2 fp32 (mean, rstd)] [batch*groups*2 fp64 sums] [batch*splits*groups*2 fp64 partials]
static void gn_geometry(int batch, long long hw, int C8, int k, long long* splits_out, long long* rpb_out) {
    long long splits = (kNumSMs * 4 + batch - 1) / batch;
    long long rpb = (hw + splits - 1) / splits;
    if (rpb < k * 4) rpb = k * 4;
    splits = (hw + rpb - 1) / rpb;
    *splits_out = splits;
    *rpb_out = rpb;
}

static int gn_block_threads(int C8) {
    int k = 256 / C8;
    if (k < 1) k = 1;
    return C8 * k;
}

}  // namespace b200

using namespace b200;

extern "C" int64_t b200_groupnorm_stats_floats(int32_t batch, int64_t hw, int32_t C, int32_t groups) {
    if (batch < 1 || hw < 1 || C < 8 || groups < 1) return 0;
    const int C8 = C / 8, T = gn_block_threads(C8), k = T / C8;
    long long splits, rpb;
    gn_geometry(batch, hw, C8, k, &splits, &rpb);
    const long long bg = static_cast<long long>(batch) * groups;
    return bg * 2 + 2 * (bg * 2) + 2 * (bg * 2 * splits);
}

extern "C" int b200_groupnorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* stats,
                                  int32_t batch, int64_t hw, int32_t C, int32_t groups, float eps, int32_t silu,
                                  void* stream) {
    B200_CHECK_ARG(C % 8 == 0 && C % groups == 0 && groups <= 64 && C / 8 <= 1024, "groupnorm: unsupported C=%d groups=%d", C, groups);
    B200_CHECK_ARG(reinterpret_cast<uintptr_t>(stats) % 8 == 0, "groupnorm: stats must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C8 = C / 8, T = gn_block_threads(C8), k = T / C8;
    long long splits, rpb;
    gn_geometry(batch, hw, C8, k, &splits, &rpb);
    double* sums = reinterpret_cast<double*>(stats + static_cast<size_t>(batch) * groups * 2);
    double* partials = sums + static_cast<size_t>(batch) * groups * 2;
    dim3 grid(static_cast<unsigned>(splits), batch);
    launch_pdl(gn_reduce_kernel<false, false>, dim3(grid), dim3(T), 0, st, static_cast<const bf16*>(x), nullptr, nullptr, nullptr, nullptr,
                                                       partials, hw, C, groups, static_cast<int>(rpb));
    B200_CHECK_LAUNCH("gn_reduce");
    launch_pdl(gn_finalize_kernel, dim3((batch * groups + 7) / 8), dim3(256), 0, st, static_cast<const double*>(partials), sums, stats,
               static_cast<int>(batch), static_cast<int>(groups), static_cast<int>(splits),
               static_cast<double>(hw) * (C / groups), eps, 1);
    B200_CHECK_LAUNCH("gn_finalize");
    const long long total = static_cast<long long>(batch) * hw * C8;
    const int blocks = grid_for(total, 256);
    if (silu)
        launch_pdl(gn_apply_kernel<true>, dim3(blocks), dim3(256), 0, st, static_cast<const bf16*>(x), static_cast<const bf16*>(gamma),
                                                      static_cast<const bf16*>(beta), stats, static_cast<bf16*>(y), hw,
                                                      C, groups, total);
    else
        launch_pdl(gn_apply_kernel<false>, dim3(blocks), dim3(256), 0, st, static_cast<const bf16*>(x), static_cast<const bf16*>(gamma),
                                                       static_cast<const bf16*>(beta), stats, static_cast<bf16*>(y), hw,
                                                       C, groups, total);
    B200_CHECK_LAUNCH("gn_apply");
    return 0;
}

extern "C" int b200_groupnorm_bwd(const void* dy, const void* x, const void* gamma, const void* beta,
                                  const float* stats, const void* dres, void* dx, int32_t batch, int64_t hw, int32_t C,
                                  int32_t groups, int32_t silu, void* stream) {
    B200_CHECK_ARG(C % 8 == 0 && C % groups == 0 && groups <= 64 && C / 8 <= 1024, "groupnorm_bwd: unsupported C=%d groups=%d", C, groups);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C8 = C / 8, T = gn_block_threads(C8), k = T / C8;
    long long splits, rpb;
    gn_geometry(batch, hw, C8, k, &splits, &rpb);
    double* sums = reinterpret_cast<double*>(const_cast<float*>(stats) + static_cast<size_t>(batch) * groups * 2);
    double* partials = sums + static_cast<size_t>(batch) * groups * 2;
    dim3 grid(static_cast<unsigned>(splits), batch);
    const bf16 *xp = static_cast<const bf16*>(x), *dyp = static_cast<const bf16*>(dy);
    const bf16 *gp = static_cast<const bf16*>(gamma), *bp = static_cast<const bf16*>(beta);
    if (silu)
        launch_pdl(gn_reduce_kernel<true, true>, dim3(grid), dim3(T), 0, st, xp, dyp, gp, bp, stats, partials, hw, C, groups, static_cast<int>(rpb));
    else
        launch_pdl(gn_reduce_kernel<true, false>, dim3(grid), dim3(T), 0, st, xp, dyp, gp, bp, stats, partials, hw, C, groups, static_cast<int>(rpb));
    B200_CHECK_LAUNCH("gn_bwd_reduce");
    launch_pdl(gn_finalize_kernel, dim3((batch * groups + 7) / 8), dim3(256), 0, st, static_cast<const double*>(partials), sums,
               const_cast<float*>(stats), static_cast<int>(batch), static_cast<int>(groups), static_cast<int>(splits), 1.0, 0.f, 0);
    B200_CHECK_LAUNCH("gn_bwd_sum");
    const long long total = static_cast<long long>(batch) * hw * C8;
    const int blocks = grid_for(total, 256);
    const double* ws = sums;
    if (silu)
        launch_pdl(gn_bwd_apply_kernel<true>, dim3(blocks), dim3(256), 0, st, dyp, xp, gp, bp, stats, ws, static_cast<const bf16*>(dres), static_cast<bf16*>(dx), hw, C, groups, total);
    else
        launch_pdl(gn_bwd_apply_kernel<false>, dim3(blocks), dim3(256), 0, st, dyp, xp, gp, bp, stats, ws, static_cast<const bf16*>(dres), static_cast<bf16*>(dx), hw, C, groups, total);
    B200_CHECK_LAUNCH("gn_bwd_apply");
    return 0;
}

namespace b200 {
// ---------------------------------------------------------------------------------------------
// Affine-parameter gradients of GroupNorm(+SiLU) / LayerNorm for the dense (full fine-tune) backward, BASELINE config 5:
//   dgamma[c] += sum_rows dz * xhat      dbeta[c] += sum_rows dz,     dz = dy (* silu'(xhat*gamma+beta) when fused)
// kGroup: statistics per (sample, group) [GroupNorm] instead of per row [LayerNorm].  A block owns 32 channel vectors
// (256 channels) x `rows_per_block` rows: 8 row lanes accumulate in registers, are combined through shared memory and
// leave as one fp32 atomicAdd per channel and block.
// ---------------------------------------------------------------------------------------------
template <bool kGroup, bool kSilu>
__global__ void norm_param_grad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                       const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                       const float* __restrict__ stats, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, long long rows, long long hw, int C, int groups,
                                       long long rows_per_block) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[8][32][17];
    const int C8 = C >> 3, cpg = kGroup ? C / groups : 1;
    const int vl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int v = blockIdx.x * 32 + vl;
    const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
    const long long r1 = min(r0 + rows_per_block, rows);
    float ag[8], ab[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ag[i] = ab[i] = 0.f;
    if (v < C8) {
        float gm[8], bt[8];
        if (kSilu) {
            load8(gamma + v * 8, gm);
            load8(beta + v * 8, bt);
        }
        for (long long r = r0 + rl; r < r1; r += 8) {
            float f[8], d[8];
            load8(x + r * C + v * 8, f);
            load8(dy + r * C + v * 8, d);
            const long long b = kGroup ? r / hw : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long si = kGroup ? (b * groups + (v * 8 + i) / cpg) : r;
                const float xh = (f[i] - stats[si * 2]) * stats[si * 2 + 1];
                float dz = d[i];
                if (kSilu) dz = bfr(dz * dsilu_f(bfr(xh * gm[i] + bt[i])));
                ag[i] += dz * xh;
                ab[i] += dz;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        red[rl][vl][i] = ag[i];
        red[rl][vl][8 + i] = ab[i];
    }
    __syncthreads();
    // 32 vectors x 16 values = 512 sums per block; 256 threads take two each
    for (int t = threadIdx.x; t < 32 * 16; t += blockDim.x) {
        const int tv = t >> 4, j = t & 15;
        const int cv = blockIdx.x * 32 + tv;
        if (cv >= C8) continue;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += red[k][tv][j];
        if (j < 8) atomicAdd(dgamma + cv * 8 + j, sum);
        else atomicAdd(dbeta + cv * 8 + (j - 8), sum);
    }
}
}  // namespace b200

extern "C" int b200_norm_param_grad(const void* dy, const void* x, const void* gamma, const void* beta, const float* stats,
                                    float* dgamma, float* dbeta, int64_t rows, int64_t hw, int32_t C, int32_t groups,
                                    int32_t silu, void* stream) {
    B200_CHECK_ARG(dy && x && stats && dgamma && dbeta && rows >= 1 && C >= 8 && C % 8 == 0, "norm_param_grad: bad arguments");
    B200_CHECK_ARG(groups == 0 || (groups >= 1 && C % groups == 0 && hw >= 1 && rows % hw == 0),
                   "norm_param_grad: C=%d groups=%d", C, groups);
    B200_CHECK_ARG(!silu || (groups > 0 && gamma && beta), "norm_param_grad: the fused SiLU needs GroupNorm gamma / beta");
    const int cblocks = (C / 8 + 31) / 32;
    long long splits = (kNumSMs * 4 + cblocks - 1) / cblocks;
    long long rpb = (rows + splits - 1) / splits;
    if (rpb < 64) rpb = 64;
    splits = (rows + rpb - 1) / rpb;
    B200_CHECK_ARG(splits <= 65535, "norm_param_grad: too many row blocks");
    const dim3 grid(cblocks, static_cast<unsigned>(splits));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define NPG(G, S)                                                                                                     \
    launch_pdl(norm_param_grad_kernel<G, S>, grid, dim3(256), 0, st, static_cast<const bf16*>(dy), static_cast<const bf16*>(x), \
               static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), stats, dgamma, dbeta,                  \
               static_cast<long long>(rows), static_cast<long long>(hw), C, groups, rpb)
    if (groups > 0 && silu) NPG(true, true);
    else if (groups > 0) NPG(true, false);
    else NPG(false, false);
#undef NPG
    B200_CHECK_LAUNCH("norm_param_grad");
    return 0;
}

extern "C" int b200_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* stats,
                                  int64_t rows, int32_t C, float eps, void* stream) {
    B200_CHECK_ARG(C % 8 == 0 && C <= kLnMaxVec * 256, "layernorm: unsupported C=%d", C);
    const int blocks = grid_for(rows * 32, 256);
    const int nv = (C / 8 + 31) / 32;
#define LN_FWD(NV)                                                                                              \
    launch_pdl(layernorm_kernel<false, NV>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),                          \
        static_cast<const bf16*>(x), nullptr, static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta), stats, \
        static_cast<bf16*>(y), rows, C, eps)
    if (nv <= 2) LN_FWD(2); else if (nv <= 5) LN_FWD(5); else LN_FWD(10);
#undef LN_FWD
    B200_CHECK_LAUNCH("layernorm_fwd");
    return 0;
}

extern "C" int b200_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* stats,
                                  const void* dres, void* dx, int64_t rows, int32_t C, void* stream) {
    B200_CHECK_ARG(C % 8 == 0 && C <= kLnMaxVec * 256, "layernorm_bwd: unsupported C=%d", C);
    const int blocks = grid_for(rows * 32, 256);
    const int nv = (C / 8 + 31) / 32;
#define LN_BWD(NV)                                                                                              \
    launch_pdl(layernorm_kernel<true, NV>, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),                           \
        static_cast<const bf16*>(x), static_cast<const bf16*>(dy), static_cast<const bf16*>(gamma),             \
        static_cast<const bf16*>(dres),                                                                         \
        const_cast<float*>(stats), static_cast<bf16*>(dx), rows, C, 0.f)
    if (nv <= 2) LN_BWD(2); else if (nv <= 5) LN_BWD(5); else LN_BWD(10);
#undef LN_BWD
    B200_CHECK_LAUNCH("layernorm_bwd");
    return 0;
}
