// Step prologue (offset noise + DDPM add_noise), min-SNR weights, masked epsilon-MSE loss with fused dPred, the
// L1 |p| reduction, and the fused AdamW (+ L1 sign-gradient, + textual-inversion rows) update.
// All HBM-bound, one pass each.
#include <string.h>
#include "common.cuh"
#include "../../include/b200_lora.h"

namespace b200 {

// main.py:311-326 + diffusers DDPMScheduler.add_noise, with torch's per-op bf16 roundings reproduced:
//   x0 = bf16(latent); noise = bf16(noise + s*offset); a = bf16(acp[t]); sa = bf16(sqrt(a)); so = bf16(sqrt(bf16(1-a)))
//   noisy = bf16(bf16(sa*x0) + bf16(so*noise))
__global__ void noise_prologue_kernel(const float* __restrict__ latent, bf16* __restrict__ noise,
                                      const float* __restrict__ offset, float offset_scale,
                                      const float* __restrict__ acp, const long long* __restrict__ timesteps,
                                      bf16* __restrict__ noisy_nchw, bf16* __restrict__ noisy_nhwc8, int B, int C, int HW) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(B) * C * HW;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % HW);
        const int c = static_cast<int>((idx / HW) % C);
        const int b = static_cast<int>(idx / (static_cast<long long>(HW) * C));
        const float a = bfr(acp[timesteps[b]]);
        const float sa = bfr(sqrtf(a)), so = bfr(sqrtf(bfr(1.f - a)));
        const float x0 = bfr(latent[idx]);
        float nz = __bfloat162float(noise[idx]);
        if (offset) nz = bfr(nz + offset_scale * offset[b * C + c]);
        noise[idx] = __float2bfloat16_rn(nz);
        const bf16 out = __float2bfloat16_rn(bfr(sa * x0) + bfr(so * nz));
        noisy_nchw[idx] = out;
        if (noisy_nhwc8) noisy_nhwc8[(static_cast<long long>(b) * HW + p) * 8 + c] = out;
    }
}

// trainer/dataset.py:181-193 -> [3P] DiagonalGaussianDistribution.sample() * scaling_factor, fp32 (main.py:186 keeps the
// VAE in fp32):  x0 = (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps) * scaling_factor, the Gaussian draw eps injected.
__global__ void latent_sample_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                     const float* __restrict__ eps, float scaling_factor, float* __restrict__ out, long long n) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float lv = fminf(fmaxf(logvar[i], -30.f), 20.f);
        const float std = expf(0.5f * lv);
        // one rounding per op, like the three ATen kernels behind mean + std * randn (no FMA contraction)
        out[i] = __fmul_rn(__fadd_rn(mean[i], __fmul_rn(std, eps[i])), scaling_factor);
    }
}

// trainer/loss.py:83-106,145-161: w_b = (min(snr_b, gamma) / snr_b) / mean_b(...)   (epsilon prediction)
__global__ void snr_weights_kernel(const float* __restrict__ acp, const long long* __restrict__ timesteps,
                                   float snr_gamma, float* __restrict__ weights, int B) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    float w = 0.f;
    if (threadIdx.x < B) {
        const float a = acp[timesteps[threadIdx.x]];
        const float sa = sqrtf(a), so = sqrtf(1.f - a);
        const float r = sa / so;
        const float snr = r * r;
        w = fminf(snr, snr_gamma) / snr;
    }
    const float total = block_sum(w, red);
    if (threadIdx.x < B) weights[threadIdx.x] = w / (total / B);
}

// loss = (1/B) sum_b w_b * mean_{c,p}( bf16((bf16(pred - noise))^2) * mask ); dpred NHWC (ld_dpred) bf16.
__global__ void diffusion_loss_kernel(const bf16* __restrict__ pred, long long ld_pred, const bf16* __restrict__ noise,
                                      const float* __restrict__ mask, const float* __restrict__ weights,
                                      float loss_scale, float* __restrict__ loss_out, bf16* __restrict__ dpred,
                                      long long ld_dpred, int B, int C, int HW) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    const long long total = static_cast<long long>(B) * HW * C;
    const float inv = 1.f / (static_cast<float>(C) * HW * B);
    float acc = 0.f;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        const long long bp = idx / C;          // b*HW + p
        const int p = static_cast<int>(bp % HW);
        const int b = static_cast<int>(bp / HW);
        const long long nchw = (static_cast<long long>(b) * C + c) * HW + p;
        const float e = bfr(__bfloat162float(pred[bp * ld_pred + c]) - __bfloat162float(noise[nchw]));
        const float m = mask[nchw];
        const float w = weights[b];
        acc += bfr(e * e) * m * w;
        if (dpred) dpred[bp * ld_dpred + c] = __float2bfloat16_rn(loss_scale * inv * w * m * 2.f * e);
    }
    acc = block_sum(acc * inv, red);
    if (threadIdx.x == 0) atomicAdd(loss_out, acc);
}

__global__ void abs_sum_kernel(const bf16* __restrict__ p, long long n, float* __restrict__ out) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    float acc = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        acc += fabsf(__bfloat162float(p[i]));
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// torch.optim.AdamW on bf16 tensors, op for op (torch/optim/adam.py _single_tensor_adam/_multi_tensor_adam):
//   p  = bf16(p * (1 - lr*wd))
//   m  = bf16(m + (1-b1) * (g - m))                      (lerp, weight < 0.5 branch)
//   v  = bf16(bf16(v * b2) + (1-b2) * g * g)             (mul_, addcmul_)
//   d  = bf16(bf16(bf16(sqrt(v)) / sqrt(bc2)) + eps)     (sqrt, div, add)
//   p  = bf16(p - (lr/bc1) * (m / d))                    (addcdiv_)
// with g = bf16(bf16(grad*grad_scale) + l1_coeff*sign(p)) - the L1 penalty's autograd contribution (main.py:353-356).
struct AdamSeg {
    float decay;      // (float)(1 - lr*wd), 1.0 => skip the decay op (torch skips it when weight_decay == 0)
    float neg_step;   // (float)(-(lr / bias_correction1))
    float l1;
};
struct AdamHyper {    // 12 floats; also the layout of the device-resident copy used under CUDA graphs
    AdamSeg s0, s1;
    float one_minus_b1, beta2, one_minus_b2, eps, bc2_sqrt, grad_scale;
};

// x / c for a per-launch constant c with rc = RN(1 / c): q = x rc, r = x - q c (exact, one FMA), q + r rc - Markstein's
// correction, which yields the correctly rounded quotient for finite operands, so the result is the one `x / c` gives but
// without a MUFU.RCP per element (ncu: the XU pipe sat at 81 % with three IEEE special-function sequences per element).
__device__ __forceinline__ float div_by_const(float x, float c, float rc) {
    const float q = x * rc;
    return fmaf(fmaf(-q, c, x), rc, q);
}

// one element, registers in / registers out (shared by the 8-wide body and the scalar tail)
__device__ __forceinline__ void adam_elem(float& pv, float& gq, float& mv, float& vv, const AdamSeg s, const AdamHyper& h,
                                          const float rc_bc2) {
    float g = bfr(gq * h.grad_scale);
    if (s.l1 != 0.f) g = bfr(g + s.l1 * (pv > 0.f ? 1.f : (pv < 0.f ? -1.f : 0.f)));
    if (s.decay != 1.f) pv = bfr(pv * s.decay);
    mv = bfr(mv + h.one_minus_b1 * (g - mv));
    vv = bfr(vv * h.beta2);
    vv = bfr(vv + h.one_minus_b2 * (g * g));   // ATen foreach addcmul: self + scalar * (t1 * t2)
    float d = bfr(sqrtf(vv));
    d = bfr(div_by_const(d, h.bc2_sqrt, rc_bc2));
    d = bfr(d + h.eps);
    pv = bfr(pv + s.neg_step * (mv / d));
}

__device__ __forceinline__ void unpack8(const uint4 w, float (&f)[8]) {
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        f[2 * q] = __uint_as_float(u[q] << 16);
        f[2 * q + 1] = __uint_as_float(u[q] & 0xffff0000u);
    }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {   // inputs already hold bf16-representable values
    uint32_t u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = (__float_as_uint(f[2 * q]) >> 16) | (__float_as_uint(f[2 * q + 1]) & 0xffff0000u);
    return make_uint4(u[0], u[1], u[2], u[3]);
}

// 8 elements per thread and iteration: 16-byte loads / stores of p, m, v (bf16) and 2 x 16 bytes of the fp32 gradient -
// 14 algorithmic bytes per element (r+w of p, m, v; read + zero of g), all fully coalesced.  `vec` = every base pointer is
// 16-byte aligned (32 for grad); otherwise (odd slices) the scalar loop covers everything.
__global__ void __launch_bounds__(256) adamw_kernel(bf16* __restrict__ p, float* __restrict__ grad, bf16* __restrict__ m,
                                                    bf16* __restrict__ v, long long n, long long n_first, AdamHyper h,
                                                    const AdamHyper* __restrict__ h_dev, int zero_grad, int vec) {
    // (parameter-writing kernel: NO early launch_dependents - a dependent's pre-wait section, e.g. a weight prefetch,
    //  must not overlap these stores; the implicit trigger at grid completion applies)
    pdl_wait();
    if (h_dev) h = *h_dev;     // CUDA-graph replay: hyper-parameters live in device memory, refreshed by a memcpy
    const float rc_bc2 = 1.f / h.bc2_sqrt;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n8 = vec ? (n >> 3) : 0;
    for (long long c = tid; c < n8; c += nthreads) {
        const long long i0 = c << 3;
        float pf[8], mf[8], vf[8], gf[8];
        unpack8(reinterpret_cast<const uint4*>(p)[c], pf);
        unpack8(reinterpret_cast<const uint4*>(m)[c], mf);
        unpack8(reinterpret_cast<const uint4*>(v)[c], vf);
        const float4 g0 = reinterpret_cast<const float4*>(grad)[2 * c], g1 = reinterpret_cast<const float4*>(grad)[2 * c + 1];
        gf[0] = g0.x, gf[1] = g0.y, gf[2] = g0.z, gf[3] = g0.w, gf[4] = g1.x, gf[5] = g1.y, gf[6] = g1.z, gf[7] = g1.w;
        if (i0 + 8 <= n_first || i0 >= n_first) {
            const AdamSeg sg = i0 < n_first ? h.s0 : h.s1;
#pragma unroll
            for (int j = 0; j < 8; ++j) adam_elem(pf[j], gf[j], mf[j], vf[j], sg, h, rc_bc2);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) adam_elem(pf[j], gf[j], mf[j], vf[j], i0 + j < n_first ? h.s0 : h.s1, h, rc_bc2);
        }
        reinterpret_cast<uint4*>(p)[c] = pack8(pf);
        reinterpret_cast<uint4*>(m)[c] = pack8(mf);
        reinterpret_cast<uint4*>(v)[c] = pack8(vf);
        if (zero_grad) {
            reinterpret_cast<float4*>(grad)[2 * c] = make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(grad)[2 * c + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (long long i = (n8 << 3) + tid; i < n; i += nthreads) {
        float pv = __bfloat162float(p[i]), mv = __bfloat162float(m[i]), vv = __bfloat162float(v[i]), g = grad[i];
        adam_elem(pv, g, mv, vv, i < n_first ? h.s0 : h.s1, h, rc_bc2);
        p[i] = __float2bfloat16_rn(pv);
        m[i] = __float2bfloat16_rn(mv);
        v[i] = __float2bfloat16_rn(vv);
        if (zero_grad) grad[i] = 0.f;
    }
}

static inline int aligned16(const void* a, const void* b, const void* c, const void* g, const void* e = nullptr,
                            const void* f = nullptr) {
    const uintptr_t x = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                        reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f);
    return (x & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 31) == 0;
}

static AdamHyper make_hyper(double lr, double wd, double l1_coeff, double lr2, double wd2, double beta1, double beta2,
                            double eps, int step, double grad_scale) {
    // scalars are formed in double exactly as torch's Python code forms them, then narrowed once
    const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
    const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
    AdamHyper h;
    h.s0 = AdamSeg{wd != 0.0 ? static_cast<float>(1.0 - lr * wd) : 1.f, static_cast<float>(-(lr / bc1)),
                   static_cast<float>(l1_coeff)};
    h.s1 = AdamSeg{wd2 != 0.0 ? static_cast<float>(1.0 - lr2 * wd2) : 1.f, static_cast<float>(-(lr2 / bc1)), 0.f};
    h.one_minus_b1 = static_cast<float>(1.0 - beta1);
    h.beta2 = static_cast<float>(beta2);
    h.one_minus_b2 = static_cast<float>(1.0 - beta2);
    h.eps = static_cast<float>(eps);
    h.bc2_sqrt = static_cast<float>(sqrt(bc2));
    h.grad_scale = static_cast<float>(grad_scale);
    return h;
}


// ---------------------------------------------------------------------------------------------------------
// Prodigy (prodigyopt 1.0 as trainer/optimizer.py:22-34, 134-144 configures it: decouple, use_bias_correction,
// safeguard_warmup) on bf16 parameters with bf16 state and the flat fp32 gradient buffer.  Three launches per step:
//   accumulate: exp_avg / exp_avg_sq / s updates (op-for-op bf16 rounding) + the two global sums <g, p0 - p>, |s|_1
//   update_d  : one thread folds the sums into d_numerator, forms d_hat and advances d (double, like the Python code)
//   apply     : p -= decay*dlr*p ; p -= dlr * exp_avg / (sqrt(exp_avg_sq) + d*eps)
// scal (device doubles): [0] d  [1] d_max  [2] d_numerator  [3] sum|s|  [4] sum<g,p0-p>  [5] dlr of this step
//                        [6] skip flag (d_denom == 0: the package returns before touching anything)  [7] d_hat
// ---------------------------------------------------------------------------------------------------------
struct ProdigyHyper {   // 12 floats, device resident (refreshed by a 48-byte copy per step)
    float lr, beta1, beta2, beta3, eps, decay, d_coef, growth, d0, bias_corr, l1, grad_scale;
};

__device__ __forceinline__ void prodigy_acc_elem(float pv, float gq, float p0v, float& mv, float& vv, float& sv, float& dot,
                                                 float& den, const ProdigyHyper& h, float a_m, float a_v, float a_s) {
    float g = bfr(gq * h.grad_scale);
    if (h.l1 != 0.f) g = bfr(g + h.l1 * (pv > 0.f ? 1.f : (pv < 0.f ? -1.f : 0.f)));
    dot += g * bfr(p0v - pv);
    mv = bfr(mv * h.beta1);
    mv = bfr(mv + a_m * g);
    vv = bfr(vv * h.beta2);
    vv = bfr(vv + a_v * (g * g));
    sv = bfr(sv * h.beta3);
    sv = bfr(sv + a_s * g);
    den += fabsf(sv);
}

__global__ void __launch_bounds__(256) prodigy_accumulate_kernel(const bf16* __restrict__ p, float* __restrict__ grad,
                                                                 bf16* __restrict__ s, const bf16* __restrict__ p0,
                                                                 bf16* __restrict__ m, bf16* __restrict__ v, long long n,
                                                                 double* __restrict__ scal,
                                                                 const ProdigyHyper* __restrict__ h_dev, int zero_grad, int vec) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    const ProdigyHyper h = *h_dev;
    const double d = scal[0];
    const bool live = h.lr > 0.f;
    const float a_m = static_cast<float>(d * (1.0 - static_cast<double>(h.beta1)));
    const float a_v = static_cast<float>(d * d * (1.0 - static_cast<double>(h.beta2)));
    const float a_s = static_cast<float>((d / static_cast<double>(h.d0)) * d);
    float dot = 0.f, den = 0.f;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n8 = vec ? (n >> 3) : 0;
    for (long long c = tid; c < n8; c += nthreads) {
        if (live) {
            float pf[8], p0f[8], mf[8], vf[8], sf[8], gf[8];
            unpack8(reinterpret_cast<const uint4*>(p)[c], pf);
            unpack8(reinterpret_cast<const uint4*>(p0)[c], p0f);
            unpack8(reinterpret_cast<const uint4*>(m)[c], mf);
            unpack8(reinterpret_cast<const uint4*>(v)[c], vf);
            unpack8(reinterpret_cast<const uint4*>(s)[c], sf);
            const float4 g0 = reinterpret_cast<const float4*>(grad)[2 * c], g1 = reinterpret_cast<const float4*>(grad)[2 * c + 1];
            gf[0] = g0.x, gf[1] = g0.y, gf[2] = g0.z, gf[3] = g0.w, gf[4] = g1.x, gf[5] = g1.y, gf[6] = g1.z, gf[7] = g1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) prodigy_acc_elem(pf[j], gf[j], p0f[j], mf[j], vf[j], sf[j], dot, den, h, a_m, a_v, a_s);
            reinterpret_cast<uint4*>(m)[c] = pack8(mf);
            reinterpret_cast<uint4*>(v)[c] = pack8(vf);
            reinterpret_cast<uint4*>(s)[c] = pack8(sf);
        }
        if (zero_grad) {
            reinterpret_cast<float4*>(grad)[2 * c] = make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(grad)[2 * c + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (long long i = (n8 << 3) + tid; i < n; i += nthreads) {
        const float g = grad[i];
        if (zero_grad) grad[i] = 0.f;
        if (!live) continue;
        float mv = __bfloat162float(m[i]), vv = __bfloat162float(v[i]), sv = __bfloat162float(s[i]);
        prodigy_acc_elem(__bfloat162float(p[i]), g, __bfloat162float(p0[i]), mv, vv, sv, dot, den, h, a_m, a_v, a_s);
        m[i] = __float2bfloat16_rn(mv);
        v[i] = __float2bfloat16_rn(vv);
        s[i] = __float2bfloat16_rn(sv);
    }
    dot = block_sum(dot, red);
    __syncthreads();
    den = block_sum(den, red);
    if (threadIdx.x == 0 && live) {
        atomicAdd(scal + 4, static_cast<double>(dot));
        atomicAdd(scal + 3, static_cast<double>(den));
    }
}

__global__ void prodigy_update_d_kernel(double* __restrict__ scal, const ProdigyHyper* __restrict__ h_dev) {
    pdl_launch();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const ProdigyHyper h = *h_dev;
    double d = scal[0], d_max = scal[1];
    const double d0 = static_cast<double>(h.d0), den = scal[3];
    const double dlr = d * static_cast<double>(h.lr) * static_cast<double>(h.bias_corr);
    double num = scal[2] * static_cast<double>(h.beta3);
    if (h.lr > 0.f) num += (d / d0) * dlr * scal[4];
    scal[5] = dlr;
    if (den == 0.0) {
        scal[6] = 1.0;                      // nothing is written back, nothing is applied
    } else {
        double d_hat = d;
        if (h.lr > 0.f) {
            d_hat = static_cast<double>(h.d_coef) * num / den;
            if (d == d0) d = fmax(d, d_hat);
            d_max = fmax(d_max, d_hat);
            d = fmin(d_max, d * static_cast<double>(h.growth));
        }
        scal[0] = d;
        scal[1] = d_max;
        scal[2] = num;
        scal[6] = 0.0;
        scal[7] = d_hat;
    }
    scal[3] = 0.0;
    scal[4] = 0.0;
}

__device__ __forceinline__ float prodigy_apply_elem(float pv, float mv, float vv, const ProdigyHyper& h, float dlr, float eps_d,
                                                    float dec) {
    float denom = bfr(sqrtf(vv));
    denom = bfr(denom + eps_d);
    if (h.decay != 0.f) pv = bfr(pv + dec * pv);
    return bfr(pv - dlr * (mv / denom));
}

__global__ void __launch_bounds__(256) prodigy_apply_kernel(bf16* __restrict__ p, const bf16* __restrict__ m,
                                                            const bf16* __restrict__ v, long long n,
                                                            const double* __restrict__ scal,
                                                            const ProdigyHyper* __restrict__ h_dev, int vec) {
    // (parameter-writing kernel: NO early launch_dependents - a dependent's pre-wait section, e.g. a weight prefetch,
    //  must not overlap these stores; the implicit trigger at grid completion applies)
    pdl_wait();
    if (scal[6] != 0.0) return;
    const ProdigyHyper h = *h_dev;
    const float dlr = static_cast<float>(scal[5]);
    const float eps_d = static_cast<float>(scal[0] * static_cast<double>(h.eps));
    const float dec = static_cast<float>(-static_cast<double>(h.decay) * scal[5]);
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n8 = vec ? (n >> 3) : 0;
    for (long long c = tid; c < n8; c += nthreads) {
        float pf[8], mf[8], vf[8];
        unpack8(reinterpret_cast<const uint4*>(p)[c], pf);
        unpack8(reinterpret_cast<const uint4*>(m)[c], mf);
        unpack8(reinterpret_cast<const uint4*>(v)[c], vf);
#pragma unroll
        for (int j = 0; j < 8; ++j) pf[j] = prodigy_apply_elem(pf[j], mf[j], vf[j], h, dlr, eps_d, dec);
        reinterpret_cast<uint4*>(p)[c] = pack8(pf);
    }
    for (long long i = (n8 << 3) + tid; i < n; i += nthreads)
        p[i] = __float2bfloat16_rn(prodigy_apply_elem(__bfloat162float(p[i]), __bfloat162float(m[i]), __bfloat162float(v[i]), h,
                                                      dlr, eps_d, dec));
}

}  // namespace b200

using namespace b200;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int b200_noise_prologue(const float* latent, void* noise, const float* offset, float offset_scale,
                                   const float* alphas_cumprod, const int64_t* timesteps, void* noisy_nchw,
                                   void* noisy_nhwc8, int32_t B, int32_t C, int32_t HW, void* stream) {
    B200_CHECK_ARG(C <= 8, "noise_prologue: C > 8");
    launch_pdl(noise_prologue_kernel, dim3(grid_for(1LL * B * C * HW, 256)), dim3(256), 0, ST, 
        latent, static_cast<bf16*>(noise), offset, offset_scale, alphas_cumprod,
        reinterpret_cast<const long long*>(timesteps), static_cast<bf16*>(noisy_nchw), static_cast<bf16*>(noisy_nhwc8), B,
        C, HW);
    B200_CHECK_LAUNCH("noise_prologue");
    return 0;
}

extern "C" int b200_latent_sample(const float* mean, const float* logvar, const float* eps, float scaling_factor, float* out,
                                  int64_t n, void* stream) {
    B200_CHECK_ARG(n >= 1 && mean && logvar && eps && out, "latent_sample: bad arguments");
    launch_pdl(latent_sample_kernel, dim3(grid_for(n, 256)), dim3(256), 0, ST, mean, logvar, eps, scaling_factor, out,
               static_cast<long long>(n));
    B200_CHECK_LAUNCH("latent_sample");
    return 0;
}

extern "C" int b200_snr_weights(const float* alphas_cumprod, const int64_t* timesteps, float snr_gamma, float* weights,
                                int32_t B, void* stream) {
    B200_CHECK_ARG(B >= 1 && B <= 1024, "snr_weights: B out of range");
    launch_pdl(snr_weights_kernel, dim3(1), dim3(((B + 31) / 32) * 32), 0, ST, alphas_cumprod, reinterpret_cast<const long long*>(timesteps),
                                                           snr_gamma, weights, B);
    B200_CHECK_LAUNCH("snr_weights");
    return 0;
}

extern "C" int b200_diffusion_loss(const void* pred, int64_t ld_pred, const void* noise, const float* mask,
                                   const float* weights, float loss_scale, float* loss_out, void* dpred,
                                   int64_t ld_dpred, int32_t B, int32_t C, int32_t HW, void* stream) {
    launch_pdl(diffusion_loss_kernel, dim3(grid_for(1LL * B * C * HW, 256, kNumSMs * 4)), dim3(256), 0, ST, 
        static_cast<const bf16*>(pred), ld_pred, static_cast<const bf16*>(noise), mask, weights, loss_scale, loss_out,
        static_cast<bf16*>(dpred), ld_dpred, B, C, HW);
    B200_CHECK_LAUNCH("diffusion_loss");
    return 0;
}

extern "C" int b200_abs_sum(const void* p, int64_t n, float* out, void* stream) {
    launch_pdl(abs_sum_kernel, dim3(grid_for(n, 256, kNumSMs * 4)), dim3(256), 0, ST, static_cast<const bf16*>(p), n, out);
    B200_CHECK_LAUNCH("abs_sum");
    return 0;
}

extern "C" int b200_adamw(void* p, float* grad, void* m, void* v, int64_t n, int64_t n_first, double lr, double wd,
                          double l1_coeff, double lr2, double wd2, double beta1, double beta2, double eps, int32_t step,
                          double grad_scale, int32_t zero_grad, void* stream) {
    B200_CHECK_ARG(step >= 1, "adamw: step must be >= 1");
    const AdamHyper h = make_hyper(lr, wd, l1_coeff, lr2, wd2, beta1, beta2, eps, step, grad_scale);
    launch_pdl(adamw_kernel, dim3(grid_for((n + 7) / 8, 256, kNumSMs * 8)), dim3(256), 0, ST, static_cast<bf16*>(p), grad,
               static_cast<bf16*>(m), static_cast<bf16*>(v), static_cast<long long>(n), static_cast<long long>(n_first), h,
               static_cast<const AdamHyper*>(nullptr), static_cast<int>(zero_grad), aligned16(p, m, v, grad));
    B200_CHECK_LAUNCH("adamw");
    return 0;
}

extern "C" int b200_adamw_pack_hyper(double lr, double wd, double l1_coeff, double lr2, double wd2, double beta1,
                                     double beta2, double eps, int32_t step, double grad_scale, float* out_host12) {
    B200_CHECK_ARG(step >= 1 && out_host12 != nullptr, "adamw_pack_hyper: bad arguments");
    static_assert(sizeof(AdamHyper) == 12 * sizeof(float), "AdamHyper layout");
    const AdamHyper h = make_hyper(lr, wd, l1_coeff, lr2, wd2, beta1, beta2, eps, step, grad_scale);
    memcpy(out_host12, &h, sizeof(h));
    return 0;
}

extern "C" int b200_adamw_dev(void* p, float* grad, void* m, void* v, int64_t n, int64_t n_first,
                              const float* hyper_dev12, int32_t zero_grad, void* stream) {
    B200_CHECK_ARG(hyper_dev12 != nullptr, "adamw_dev: null hyper-parameter buffer");
    AdamHyper h;
    memset(&h, 0, sizeof(h));
    launch_pdl(adamw_kernel, dim3(grid_for((n + 7) / 8, 256, kNumSMs * 8)), dim3(256), 0, ST, static_cast<bf16*>(p), grad,
               static_cast<bf16*>(m), static_cast<bf16*>(v), static_cast<long long>(n), static_cast<long long>(n_first), h,
               reinterpret_cast<const AdamHyper*>(hyper_dev12), static_cast<int>(zero_grad), aligned16(p, m, v, grad));
    B200_CHECK_LAUNCH("adamw_dev");
    return 0;
}

extern "C" int b200_prodigy_pack_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, double d_coef,
                                       double growth_rate, double d0, int32_t k, int32_t use_bias_correction, double l1_coeff,
                                       double grad_scale, float* out_host12) {
    B200_CHECK_ARG(k >= 0 && out_host12 != nullptr && d0 > 0.0, "prodigy_pack_hyper: bad arguments");
    ProdigyHyper h;
    const double bc = use_bias_correction ? sqrt(1.0 - pow(beta2, k + 1.0)) / (1.0 - pow(beta1, k + 1.0)) : 1.0;
    h.lr = static_cast<float>(lr);
    h.beta1 = static_cast<float>(beta1);
    h.beta2 = static_cast<float>(beta2);
    h.beta3 = static_cast<float>(sqrt(beta2));
    h.eps = static_cast<float>(eps);
    h.decay = static_cast<float>(weight_decay);
    h.d_coef = static_cast<float>(d_coef);
    h.growth = static_cast<float>(growth_rate);     // inf stays inf
    h.d0 = static_cast<float>(d0);
    h.bias_corr = static_cast<float>(bc);
    h.l1 = static_cast<float>(l1_coeff);
    h.grad_scale = static_cast<float>(grad_scale);
    static_assert(sizeof(ProdigyHyper) == 12 * sizeof(float), "ProdigyHyper layout");
    memcpy(out_host12, &h, sizeof(h));
    return 0;
}

extern "C" int b200_prodigy_step(void* p, float* grad, void* s, const void* p0, void* exp_avg, void* exp_avg_sq, int64_t n,
                                 double* scal8, const float* hyper_dev12, int32_t zero_grad, void* stream) {
    B200_CHECK_ARG(n >= 1 && p && grad && s && p0 && exp_avg && exp_avg_sq && scal8 && hyper_dev12, "prodigy_step: bad arguments");
    const ProdigyHyper* hd = reinterpret_cast<const ProdigyHyper*>(hyper_dev12);
    const int vec = aligned16(p, exp_avg, exp_avg_sq, grad, s, p0);
    launch_pdl(prodigy_accumulate_kernel, dim3(grid_for((n + 7) / 8, 256, kNumSMs * 8)), dim3(256), 0, ST,
               static_cast<const bf16*>(p), grad, static_cast<bf16*>(s), static_cast<const bf16*>(p0), static_cast<bf16*>(exp_avg),
               static_cast<bf16*>(exp_avg_sq), static_cast<long long>(n), scal8, hd, static_cast<int>(zero_grad), vec);
    B200_CHECK_LAUNCH("prodigy_accumulate");
    launch_pdl(prodigy_update_d_kernel, dim3(1), dim3(32), 0, ST, scal8, hd);
    B200_CHECK_LAUNCH("prodigy_update_d");
    launch_pdl(prodigy_apply_kernel, dim3(grid_for((n + 7) / 8, 256, kNumSMs * 8)), dim3(256), 0, ST, static_cast<bf16*>(p),
               static_cast<const bf16*>(exp_avg), static_cast<const bf16*>(exp_avg_sq), static_cast<long long>(n),
               static_cast<const double*>(scal8), hd, vec);
    B200_CHECK_LAUNCH("prodigy_apply");
    return 0;
}
