// HBM-bound helpers of the step: row softmax fwd/bwd (unfused attention), GEGLU, SiLU, adds, NHWC layout helpers
// (nearest 2x upsample, strided 3x3 im2col / col2im, the 9-tap shift-stack for the conv-LoRA backward) and the
// sinusoidal timestep embedding.  16-byte vectors everywhere the layout allows it.
#include "common.cuh"
#include "../../include/b200_lora.h"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// softmax: one warp per row (cols <= 1024, <= 32 values per lane in registers) or one 128-thread
// block per row (cols <= 4096); rows of any other length take the 3-pass global-memory path.
// ------------------------------------------------------------------------------------------------
template <int kPerLane>
__global__ void softmax_fwd_warp_kernel(const float* __restrict__ S, bf16* __restrict__ P, long long rows, int cols,
                                        long long ld_s, long long ld_p) {
    pdl_launch();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; row < rows; row += warps) {
        const float* s = S + row * ld_s;
        float v[kPerLane];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            const int c = lane + j * 32;
            v[j] = (c < cols) ? s[c] : -INFINITY;
            mx = fmaxf(mx, v[j]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            v[j] = (lane + j * 32 < cols) ? __expf(v[j] - mx) : 0.f;
            sum += v[j];
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        bf16* p = P + row * ld_p;
#pragma unroll
        for (int j = 0; j < kPerLane; ++j) {
            const int c = lane + j * 32;
            if (c < ld_p) p[c] = __float2bfloat16_rn(v[j] * inv);   // columns in [cols, ld_p) get exact zeros
        }
    }
}

template <int kPerThread>
__global__ void softmax_fwd_block_kernel(const float* __restrict__ S, bf16* __restrict__ P, long long rows, int cols,
                                         long long ld_s, long long ld_p) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float* s = S + row * ld_s;
        float v[kPerThread];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            const int c = threadIdx.x + j * blockDim.x;
            v[j] = (c < cols) ? s[c] : -INFINITY;
            mx = fmaxf(mx, v[j]);
        }
        mx = block_max(mx, red);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            v[j] = (threadIdx.x + j * blockDim.x < cols) ? __expf(v[j] - mx) : 0.f;
            sum += v[j];
        }
        sum = block_sum(sum, red);
        const float inv = 1.f / sum;
        bf16* p = P + row * ld_p;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            const int c = threadIdx.x + j * blockDim.x;
            if (c < ld_p) p[c] = __float2bfloat16_rn(v[j] * inv);
        }
    }
}

__global__ void softmax_fwd_generic_kernel(const float* __restrict__ S, bf16* __restrict__ P, long long rows, int cols,
                                           long long ld_s, long long ld_p) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float* s = S + row * ld_s;
        float mx = -INFINITY;
        for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, s[c]);
        mx = block_max(mx, red);
        float sum = 0.f;
        for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf(s[c] - mx);
        sum = block_sum(sum, red);
        const float inv = 1.f / sum;
        bf16* p = P + row * ld_p;
        for (int c = threadIdx.x; c < ld_p; c += blockDim.x)
            p[c] = __float2bfloat16_rn(c < cols ? __expf(s[c] - mx) * inv : 0.f);
    }
}

// dS = P * (dP - sum_j P_j dP_j)
__global__ void softmax_bwd_kernel(const bf16* __restrict__ P, const float* __restrict__ dP, bf16* __restrict__ dS,
                                   long long rows, int cols, long long ld_p, long long ld_dp, int warp_rows) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    const int lane = threadIdx.x & 31;
    if (warp_rows) {
        const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
        for (long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; row < rows; row += warps) {
            const bf16* p = P + row * ld_p;
            const float* dp = dP + row * ld_dp;
            float dot = 0.f;
            for (int c = lane; c < cols; c += 32) dot += __bfloat162float(p[c]) * dp[c];
            dot = warp_sum(dot);
            bf16* ds = dS + row * ld_p;
            for (int c = lane; c < ld_p; c += 32)
                ds[c] = __float2bfloat16_rn(c < cols ? __bfloat162float(p[c]) * (dp[c] - dot) : 0.f);
        }
    } else {
        for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
            const bf16* p = P + row * ld_p;
            const float* dp = dP + row * ld_dp;
            float dot = 0.f;
            for (int c = threadIdx.x; c < cols; c += blockDim.x) dot += __bfloat162float(p[c]) * dp[c];
            dot = block_sum(dot, red);
            bf16* ds = dS + row * ld_p;
            for (int c = threadIdx.x; c < ld_p; c += blockDim.x)
                ds[c] = __float2bfloat16_rn(c < cols ? __bfloat162float(p[c]) * (dp[c] - dot) : 0.f);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GEGLU / SiLU / add
// ------------------------------------------------------------------------------------------------
// h = [value | gate] ([rows, 2*inner]).  il == 0: value in columns [0, inner), gate in [inner, 2*inner) (diffusers' chunk(2)).
// il > 0 (a multiple of 8 dividing inner; 128 in the step): blocks of il value columns alternate with blocks of il gate
// columns - the layout the FF up-projection writes when GEGLU is fused into its epilogue (one 256-wide tile = 128 value +
// the matching 128 gate columns): value column j sits at (j / il) * 2 il + j % il, its gate il further.
__device__ __forceinline__ long long geglu_value_col(int j, int il) { return il ? static_cast<long long>(j / il) * 2 * il + j % il : j; }

__global__ void geglu_fwd_kernel(const bf16* __restrict__ h, bf16* __restrict__ y, long long rows, int inner, int il) {
    pdl_launch();
    pdl_wait();
    const int I8 = inner >> 3;
    const long long total = rows * I8;
    const int goff = il ? il : inner;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = idx / I8;
        const int v = static_cast<int>(idx % I8);
        const long long vc = geglu_value_col(v * 8, il);
        float a[8], g[8], o[8];
        load8(h + r * 2 * inner + vc, a);
        load8(h + r * 2 * inner + vc + goff, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = a[i] * bfr(gelu_f(g[i]));
        store8(y + r * inner + v * 8, o);
    }
}

__global__ void geglu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ h, bf16* __restrict__ dh,
                                 long long rows, int inner, int il) {
    pdl_launch();
    pdl_wait();
    const int I8 = inner >> 3;
    const long long total = rows * I8;
    const int goff = il ? il : inner;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = idx / I8;
        const int v = static_cast<int>(idx % I8);
        const long long vc = geglu_value_col(v * 8, il);
        float a[8], g[8], d[8], da[8], dg[8];
        load8(h + r * 2 * inner + vc, a);
        load8(h + r * 2 * inner + vc + goff, g);
        load8(dy + r * inner + v * 8, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float gl, dgl;
            gelu_both(g[i], gl, dgl);
            da[i] = d[i] * bfr(gl);
            dg[i] = bfr(d[i] * a[i]) * dgl;
        }
        store8(dh + r * 2 * inner + vc, da);
        store8(dh + r * 2 * inner + vc + goff, dg);
    }
}

__global__ void silu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] = __float2bfloat16_rn(silu_f(__bfloat162float(x[i])));
}
__global__ void silu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx,
                                long long n) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * dsilu_f(__bfloat162float(x[i])));
}

// CLIP text-encoder MLP activations (transformers CLIPMLP, inference.py:131-177 path): kind 0 = exact (erf) GELU
// (OpenCLIP bigG), kind 1 = quick_gelu x * sigmoid(1.702 x) (CLIP-L).  fp32 math, one bf16 rounding.
template <int kKind>
__global__ void act_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8, long long n) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float f[8], o[8];
        load8(x + i * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = kKind == 0 ? gelu_f(f[j]) : silu_f(1.702f * f[j]) * (1.f / 1.702f);
        store8(y + i * 8, o);
    }
    if (blockIdx.x == 0) {
        for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
            const float f = __bfloat162float(x[i]);
            y[i] = __float2bfloat16_rn(kKind == 0 ? gelu_f(f) : silu_f(1.702f * f) * (1.f / 1.702f));
        }
    }
}
template <int kKind>
__global__ void act_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx,
                               long long n8, long long n) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float f[8], d[8], o[8];
        load8(x + i * 8, f);
        load8(dy + i * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = d[j] * (kKind == 0 ? dgelu_f(f[j]) : dsilu_f(1.702f * f[j]));
        store8(dx + i * 8, o);
    }
    if (blockIdx.x == 0) {
        for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
            const float f = __bfloat162float(x[i]);
            dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * (kKind == 0 ? dgelu_f(f) : dsilu_f(1.702f * f)));
        }
    }
}

// Head re-pitch for the fused attention kernel (head_dim 64 only): [rows, H * d_src] -> [rows, H * d_dst], per head the first
// min(d_src, d_dst) channels are copied and the rest zero-filled.  d = 40 (SD1.5's 64x64-latent layers) -> 64 makes
// QK^T, PV and every gradient exact (the padded channels are zero in q, k and v); 64 -> 40 drops them again.
// One 16-byte chunk per thread; d_src, d_dst multiples of 8.
__global__ void head_pad_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long rows, int H, int cs, int cd,
                                long long ld_src, long long ld_dst) {
    pdl_launch();
    pdl_wait();
    const long long per_row = static_cast<long long>(H) * cd;
    const long long total = rows * per_row;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / per_row;
        const int c = static_cast<int>(i - row * per_row);
        const int head = c / cd, j = c - head * cd;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (j < cs) w = *reinterpret_cast<const uint4*>(src + row * ld_src + (static_cast<long long>(head) * cs + j) * 8);
        *reinterpret_cast<uint4*>(dst + row * ld_dst + static_cast<long long>(c) * 8) = w;
    }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ c,
                           bf16* __restrict__ y, long long n8, long long n) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float fa[8], fb[8], o[8];
        load8(a + i * 8, fa);
        load8(b + i * 8, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fa[j] + fb[j];
        if (c) {
            float fc[8];
            load8(c + i * 8, fc);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = bfr(o[j]) + fc[j];
        }
        store8(y + i * 8, o);
    }
    // scalar tail
    if (blockIdx.x == 0) {
        for (long long i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
            float o = __bfloat162float(a[i]) + __bfloat162float(b[i]);
            if (c) o = bfr(o) + __bfloat162float(c[i]);
            y[i] = __float2bfloat16_rn(o);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NHWC layout helpers
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C8) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(N) * (2 * H) * (2 * W) * C8;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(idx % C8);
        long long p = idx / C8;
        const int wo = static_cast<int>(p % (2 * W));  p /= 2 * W;
        const int ho = static_cast<int>(p % (2 * H));
        const int n = static_cast<int>(p / (2 * H));
        const long long src = ((static_cast<long long>(n) * H + (ho >> 1)) * W + (wo >> 1)) * C8 + v;
        reinterpret_cast<uint4*>(y)[idx] = reinterpret_cast<const uint4*>(x)[src];
    }
}

__global__ void upsample2x_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int N, int H, int W, int C8) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(N) * H * W * C8;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(idx % C8);
        long long p = idx / C8;
        const int w = static_cast<int>(p % W);  p /= W;
        const int h = static_cast<int>(p % H);
        const int n = static_cast<int>(p / H);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float f[8];
                load8(dy + (((static_cast<long long>(n) * 2 * H + 2 * h + i) * 2 * W + 2 * w + j) * C8 + v) * 8, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        store8(dx + idx * 8, acc);
    }
}

__global__ void im2col3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int N, int H, int W, int C8,
                                 int stride, int Ho, int Wo) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(N) * Ho * Wo * 9 * C8;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(idx % C8);
        long long p = idx / C8;
        const int tap = static_cast<int>(p % 9);  p /= 9;
        const int wo = static_cast<int>(p % Wo);  p /= Wo;
        const int ho = static_cast<int>(p % Ho);
        const int n = static_cast<int>(p / Ho);
        const int h = ho * stride + tap / 3 - 1, w = wo * stride + tap % 3 - 1;
        uint4 val = zero;
        if (h >= 0 && h < H && w >= 0 && w < W)
            val = reinterpret_cast<const uint4*>(x)[((static_cast<long long>(n) * H + h) * W + w) * C8 + v];
        reinterpret_cast<uint4*>(col)[idx] = val;
    }
}

// gather form of the transposed im2col: dx[n,h,w,:] = sum over (tap, ho, wo) with ho*stride + kh - 1 == h, ...
__global__ void col2im3x3_kernel(const bf16* __restrict__ col, bf16* __restrict__ dx, int N, int H, int W, int C8,
                                 int stride, int Ho, int Wo) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(N) * H * W * C8;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(idx % C8);
        long long p = idx / C8;
        const int w = static_cast<int>(p % W);  p /= W;
        const int h = static_cast<int>(p % H);
        const int n = static_cast<int>(p / H);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int kh = 0; kh < 3; ++kh) {
            const int hn = h + 1 - kh;
            if (hn < 0 || hn % stride) continue;
            const int ho = hn / stride;
            if (ho >= Ho) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int wn = w + 1 - kw;
                if (wn < 0 || wn % stride) continue;
                const int wo = wn / stride;
                if (wo >= Wo) continue;
                float f[8];
                load8(col + ((((static_cast<long long>(n) * Ho + ho) * Wo + wo) * 9 + kh * 3 + kw) * C8 + v) * 8, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        }
        store8(dx + idx * 8, acc);
    }
}

// out[b, c] = sum_p x[b, p, c]   (d(time-embedding projection) = per-image column sum of d(conv1 output))
// out[b, c] = sum_p x[b, p, c]: 16-byte vector loads, rows split over gridDim.z, fp32 partial sums combined with
// red.global.add into a zeroed scratch row, converted to bf16 by colsum_finish_kernel.
__global__ void colsum_partial_kernel(const bf16* __restrict__ x, float* __restrict__ acc, long long hw, int C,
                                      int rows_per_block) {
    pdl_launch();
    pdl_wait();
    __shared__ float part[16][128 + 4];
    const int b = blockIdx.y;
    const int cg = threadIdx.x & 15;                 // 16 column groups of 8 channels = 128 channels per block
    const int ry = threadIdx.x >> 4;                 // 16 row lanes
    const int c0 = blockIdx.x * 128 + cg * 8;
    const long long p0 = static_cast<long long>(blockIdx.z) * rows_per_block;
    const long long p1 = min(hw, p0 + rows_per_block);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {
        const bf16* base = x + static_cast<long long>(b) * hw * C + c0;
        for (long long p = p0 + ry; p < p1; p += 16) {
            float f[8];
            load8(base + p * C, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += f[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) part[ry][cg * 8 + i] = a[i];
    __syncthreads();
    if (threadIdx.x < 128) {
        const int c = blockIdx.x * 128 + threadIdx.x;
        if (c < C) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) s += part[i][threadIdx.x];
            atomicAdd(acc + static_cast<long long>(b) * C + c, s);
        }
    }
}
__global__ void colsum_finish_kernel(const float* __restrict__ acc, bf16* __restrict__ out, long long n) {
    pdl_launch();
    pdl_wait();
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16_rn(acc[i]);
}

// K-major copies of every LoRA-B factor: bt[j, n] = B[n, j].  One launch for all layers (table rows: offset of B in
// the flat parameter buffer, offset of the copy, N, row stride rs); one CTA per layer.  The input-gradient GEMM reads
// the copy as its rank-r side operand: read in place (MN-major) that operand is a 64 x 32-byte box every CTA re-fetches
// at every k-block - an L2 hot spot that cost 3.7 us per launch.
__global__ void lora_bt_kernel(const bf16* __restrict__ params, bf16* __restrict__ bt, const long long* __restrict__ table) {
    // (parameter-writing kernel: NO early launch_dependents - a dependent's pre-wait section, e.g. a weight prefetch,
    //  must not overlap these stores; the implicit trigger at grid completion applies)
    pdl_wait();
    const long long* e = table + 4LL * blockIdx.x;
    const bf16* src = params + e[0];
    bf16* dst = bt + e[1];
    const int N = static_cast<int>(e[2]), rs = static_cast<int>(e[3]);
    for (int idx = threadIdx.x; idx < N * rs; idx += blockDim.x) {
        const int j = idx / N, n = idx - j * N;           // consecutive threads -> consecutive n: coalesced writes
        dst[idx] = src[static_cast<long long>(n) * rs + j];
    }
}

// Derived copies of LoRA-B factors for the fused q|k|v projection, one CTA per table row
// (src_off, dst_off, N, rs, dst_ld, transpose):   transpose == 0: dst[n * dst_ld + j] = B[n, j]   (a diagonal block of the
// block-diagonal [3C, 3r] factor of the forward);  transpose == 1: dst[j * dst_ld + n] = B[n, j]  (its K-major transpose,
// the side operand of the input-gradient GEMM).  Everything outside the blocks stays zero (written once at allocation).
__global__ void lora_pack_kernel(const bf16* __restrict__ params, bf16* __restrict__ dst_base, const long long* __restrict__ table) {
    // (parameter-writing kernel: NO early launch_dependents - a dependent's pre-wait section, e.g. a weight prefetch,
    //  must not overlap these stores; the implicit trigger at grid completion applies)
    pdl_wait();
    const long long* e = table + 6LL * blockIdx.x;
    const bf16* src = params + e[0];
    bf16* dst = dst_base + e[1];
    const int N = static_cast<int>(e[2]), rs = static_cast<int>(e[3]);
    const long long dst_ld = e[4];
    if (e[5] == 0) {
        for (int idx = threadIdx.x; idx < N * rs; idx += blockDim.x) {
            const int n = idx / rs, j = idx - n * rs;
            dst[n * dst_ld + j] = src[idx];
        }
    } else {
        for (int idx = threadIdx.x; idx < N * rs; idx += blockDim.x) {
            const int j = idx / N, n = idx - j * N;           // consecutive threads -> consecutive n: coalesced writes
            dst[j * dst_ld + n] = src[static_cast<long long>(n) * rs + j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Bicubic resize of channels-last maps (F.interpolate(mode="bicubic", align_corners=False, antialias=False) as the
// reference applies it to the captured cross-attention maps, ti_cross_attn_loss.py:262-266): A = -0.75, taps clamped
// to the image, fp32 arithmetic, one rounding to bf16.  x: [B, Hi, Wi, C] with pixel stride ld_in; y: [B, Ho, Wo, C].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
    w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
    w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
    w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void bicubic_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int Hi, int Wi, int Ho, int Wo,
                                   int C, long long ld_in, long long ld_out, float sh, float sw) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(B) * Ho * Wo * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        long long p = idx / C;
        const int ox = static_cast<int>(p % Wo);  p /= Wo;
        const int oy = static_cast<int>(p % Ho);
        const int b = static_cast<int>(p / Ho);
        const float ry = sh * (oy + 0.5f) - 0.5f, rx = sw * (ox + 0.5f) - 0.5f;
        const int iy = static_cast<int>(floorf(ry)), ix = static_cast<int>(floorf(rx));
        float wy[4], wx[4];
        cubic_coeffs(ry - iy, wy);
        cubic_coeffs(rx - ix, wx);
        const bf16* img = x + static_cast<long long>(b) * Hi * Wi * ld_in + c;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int yy = clampi(iy - 1 + i, 0, Hi - 1);
            float row = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = clampi(ix - 1 + j, 0, Wi - 1);
                row += __bfloat162float(img[(static_cast<long long>(yy) * Wi + xx) * ld_in]) * wx[j];
            }
            acc += row * wy[i];
        }
        y[((static_cast<long long>(b) * Ho + oy) * Wo + ox) * ld_out + c] = __float2bfloat16_rn(acc);
    }
}

// Gather form of the adjoint (no atomics): every input pixel sums the output pixels whose clamped taps land on it.
__global__ void bicubic_bwd_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int B, int Hi, int Wi, int Ho, int Wo,
                                   int C, long long ld_dy, long long ld_dx, float sh, float sw) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(B) * Hi * Wi * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        long long p = idx / C;
        const int xx = static_cast<int>(p % Wi);  p /= Wi;
        const int yy = static_cast<int>(p % Hi);
        const int b = static_cast<int>(p / Hi);
        // candidate outputs: |src(o) - in| <= 2 (+ clamped taps at the borders, which the tap test below catches)
        int oy0 = static_cast<int>(floorf((yy - 2.5f) / sh)) - 1, oy1 = static_cast<int>(ceilf((yy + 2.5f) / sh)) + 1;
        int ox0 = static_cast<int>(floorf((xx - 2.5f) / sw)) - 1, ox1 = static_cast<int>(ceilf((xx + 2.5f) / sw)) + 1;
        if (yy == 0) oy0 = 0;
        if (yy == Hi - 1) oy1 = Ho - 1;
        if (xx == 0) ox0 = 0;
        if (xx == Wi - 1) ox1 = Wo - 1;
        oy0 = max(oy0, 0);  oy1 = min(oy1, Ho - 1);
        ox0 = max(ox0, 0);  ox1 = min(ox1, Wo - 1);
        const bf16* g = dy + static_cast<long long>(b) * Ho * Wo * ld_dy + c;
        float acc = 0.f;
        for (int oy = oy0; oy <= oy1; ++oy) {
            const float ry = sh * (oy + 0.5f) - 0.5f;
            const int iy = static_cast<int>(floorf(ry));
            float wy[4];
            cubic_coeffs(ry - iy, wy);
            float cy = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (clampi(iy - 1 + i, 0, Hi - 1) == yy) cy += wy[i];
            if (cy == 0.f) continue;
            for (int ox = ox0; ox <= ox1; ++ox) {
                const float rx = sw * (ox + 0.5f) - 0.5f;
                const int ix = static_cast<int>(floorf(rx));
                float wx[4];
                cubic_coeffs(rx - ix, wx);
                float cx = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (clampi(ix - 1 + j, 0, Wi - 1) == xx) cx += wx[j];
                if (cx != 0.f) acc += cy * cx * __bfloat162float(g[(static_cast<long long>(oy) * Wo + ox) * ld_dy]);
            }
        }
        dx[((static_cast<long long>(b) * Hi + yy) * Wi + xx) * ld_dx + c] = __float2bfloat16_rn(acc);
    }
}

// U9[p, tap*r + j] = U[p - off(tap), j]
__global__ void shift_stack9_kernel(const bf16* __restrict__ U, bf16* __restrict__ U9, int N, int H, int W, int r,
                                    int ld_in, int ld_out) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(N) * H * W * 9 * r;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(idx % r);
        long long p = idx / r;
        const int tap = static_cast<int>(p % 9);  p /= 9;
        const int w = static_cast<int>(p % W);  p /= W;
        const int h = static_cast<int>(p % H);
        const int n = static_cast<int>(p / H);
        const int hs = h - (tap / 3 - 1), ws = w - (tap % 3 - 1);
        bf16 val = __float2bfloat16_rn(0.f);
        if (hs >= 0 && hs < H && ws >= 0 && ws < W) val = U[((static_cast<long long>(n) * H + hs) * W + ws) * ld_in + j];
        U9[((static_cast<long long>(n) * H + h) * W + w) * ld_out + tap * r + j] = val;
    }
}

// T[p, j] = bf16(alpha * sum_tap Z[p + off(tap), tap*r + j])  - the 3x3 conv-LoRA A product from ONE plain GEMM
// Z = X . A_taps^T ([M, 9r], fp32): gather-form adjoint of shift_stack9, zero padding at the image border.
__global__ void shift_sum9_kernel(const float* __restrict__ Z, bf16* __restrict__ T, int N, int H, int W, int r, int ld_z,
                                  int ld_t, float alpha) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(N) * H * W * r;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int j = static_cast<int>(idx % r);
        long long p = idx / r;
        const int w = static_cast<int>(p % W);
        const long long nh = p / W;
        const int h = static_cast<int>(nh % H);
        const long long n = nh / H;
        float acc = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int hs = h + tap / 3 - 1, ws = w + tap % 3 - 1;
            if (hs >= 0 && hs < H && ws >= 0 && ws < W) acc += Z[((n * H + hs) * W + ws) * ld_z + tap * r + j];
        }
        T[p * ld_t + j] = __float2bfloat16_rn(acc * alpha);
    }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, bf16* __restrict__ out, int n, int dim) {
    pdl_launch();
    pdl_wait();
    const int half = dim / 2;
    const int total = n * dim;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int i = idx / dim, j = idx % dim;
        const int k = j < half ? j : j - half;
        const float freq = expf(-9.210340371976184f * static_cast<float>(k) / static_cast<float>(half));
        const float a = t[i] * freq;
        out[idx] = __float2bfloat16_rn(j < half ? cosf(a) : sinf(a));
    }
}

}  // namespace b200

using namespace b200;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int b200_softmax_fwd(const float* S, void* P, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p,
                                void* stream) {
    B200_CHECK_ARG(cols >= 1 && ld_p >= cols && ld_s >= cols, "softmax: bad extents");
    bf16* p = static_cast<bf16*>(P);
    if (ld_p <= 128) {
        launch_pdl(softmax_fwd_warp_kernel<4>, dim3(grid_for(rows * 32, 256)), dim3(256), 0, ST, S, p, rows, cols, ld_s, ld_p);
    } else if (ld_p <= 1024) {
        launch_pdl(softmax_fwd_warp_kernel<32>, dim3(grid_for(rows * 32, 256)), dim3(256), 0, ST, S, p, rows, cols, ld_s, ld_p);
    } else if (ld_p <= 4096) {
        launch_pdl(softmax_fwd_block_kernel<32>, dim3(static_cast<int>(rows < kNumSMs * 16 ? rows : kNumSMs * 16)), dim3(128), 0, ST,
                   S, p, rows, cols, ld_s, ld_p);
    } else {
        launch_pdl(softmax_fwd_generic_kernel, dim3(static_cast<int>(rows < kNumSMs * 8 ? rows : kNumSMs * 8)), dim3(256), 0, ST,
                   S, p, rows, cols, ld_s, ld_p);
    }
    B200_CHECK_LAUNCH("softmax_fwd");
    return 0;
}

extern "C" int b200_softmax_bwd(const void* P, const float* dP, void* dS, int64_t rows, int32_t cols, int64_t ld_p,
                                int64_t ld_dp, void* stream) {
    B200_CHECK_ARG(cols >= 1 && ld_p >= cols && ld_dp >= cols, "softmax_bwd: bad extents");
    const int warp_rows = cols <= 512;
    const int blocks = warp_rows ? grid_for(rows * 32, 256) : static_cast<int>(rows < kNumSMs * 16 ? rows : kNumSMs * 16);
    launch_pdl(softmax_bwd_kernel, dim3(blocks), dim3(256), 0, ST, static_cast<const bf16*>(P), dP, static_cast<bf16*>(dS), rows, cols, ld_p,
                                               ld_dp, warp_rows);
    B200_CHECK_LAUNCH("softmax_bwd");
    return 0;
}

extern "C" int b200_geglu_fwd(const void* h, void* y, int64_t rows, int32_t inner, int32_t interleave, void* stream) {
    B200_CHECK_ARG(inner % 8 == 0 && interleave >= 0 && interleave % 8 == 0 && (interleave == 0 || inner % interleave == 0),
                   "geglu: inner %% 8 != 0 or a bad interleave");
    launch_pdl(geglu_fwd_kernel, dim3(grid_for(rows * (inner / 8), 256)), dim3(256), 0, ST, static_cast<const bf16*>(h),
               static_cast<bf16*>(y), static_cast<long long>(rows), static_cast<int>(inner), static_cast<int>(interleave));
    B200_CHECK_LAUNCH("geglu_fwd");
    return 0;
}
extern "C" int b200_geglu_bwd(const void* dy, const void* h, void* dh, int64_t rows, int32_t inner, int32_t interleave, void* stream) {
    B200_CHECK_ARG(inner % 8 == 0 && interleave >= 0 && interleave % 8 == 0 && (interleave == 0 || inner % interleave == 0),
                   "geglu: inner %% 8 != 0 or a bad interleave");
    launch_pdl(geglu_bwd_kernel, dim3(grid_for(rows * (inner / 8), 256)), dim3(256), 0, ST, static_cast<const bf16*>(dy),
               static_cast<const bf16*>(h), static_cast<bf16*>(dh), static_cast<long long>(rows), static_cast<int>(inner),
               static_cast<int>(interleave));
    B200_CHECK_LAUNCH("geglu_bwd");
    return 0;
}
extern "C" int b200_silu_fwd(const void* x, void* y, int64_t n, void* stream) {
    launch_pdl(silu_fwd_kernel, dim3(grid_for(n, 256)), dim3(256), 0, ST, static_cast<const bf16*>(x), static_cast<bf16*>(y), n);
    B200_CHECK_LAUNCH("silu_fwd");
    return 0;
}
extern "C" int b200_silu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream) {
    launch_pdl(silu_bwd_kernel, dim3(grid_for(n, 256)), dim3(256), 0, ST, static_cast<const bf16*>(dy), static_cast<const bf16*>(x),
                                                      static_cast<bf16*>(dx), n);
    B200_CHECK_LAUNCH("silu_bwd");
    return 0;
}
extern "C" int b200_act_fwd(const void* x, void* y, int64_t n, int32_t kind, void* stream) {
    B200_CHECK_ARG(n >= 1 && x && y && (kind == 0 || kind == 1), "act_fwd: bad arguments (kind %d)", kind);
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const long long n8 = aligned ? n / 8 : 0;
    const dim3 grid(grid_for(n8 > 0 ? n8 : 1, 256));
    if (kind == 0) launch_pdl(act_fwd_kernel<0>, grid, dim3(256), 0, ST, static_cast<const bf16*>(x), static_cast<bf16*>(y), n8, static_cast<long long>(n));
    else launch_pdl(act_fwd_kernel<1>, grid, dim3(256), 0, ST, static_cast<const bf16*>(x), static_cast<bf16*>(y), n8, static_cast<long long>(n));
    B200_CHECK_LAUNCH("act_fwd");
    return 0;
}
extern "C" int b200_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int32_t kind, void* stream) {
    B200_CHECK_ARG(n >= 1 && dy && x && dx && (kind == 0 || kind == 1), "act_bwd: bad arguments (kind %d)", kind);
    const bool aligned = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    const long long n8 = aligned ? n / 8 : 0;
    const dim3 grid(grid_for(n8 > 0 ? n8 : 1, 256));
    if (kind == 0) launch_pdl(act_bwd_kernel<0>, grid, dim3(256), 0, ST, static_cast<const bf16*>(dy), static_cast<const bf16*>(x), static_cast<bf16*>(dx), n8, static_cast<long long>(n));
    else launch_pdl(act_bwd_kernel<1>, grid, dim3(256), 0, ST, static_cast<const bf16*>(dy), static_cast<const bf16*>(x), static_cast<bf16*>(dx), n8, static_cast<long long>(n));
    B200_CHECK_LAUNCH("act_bwd");
    return 0;
}
extern "C" int b200_add(const void* a, const void* b, const void* c, void* y, int64_t n, void* stream) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const long long n8 = aligned ? n / 8 : 0;
    launch_pdl(add_kernel, dim3(grid_for(n8 > 0 ? n8 : 1, 256)), dim3(256), 0, ST, static_cast<const bf16*>(a), static_cast<const bf16*>(b),
                                                               static_cast<const bf16*>(c), static_cast<bf16*>(y), n8, n);
    B200_CHECK_LAUNCH("add");
    return 0;
}
extern "C" int b200_head_pad(const void* src, void* dst, int64_t rows, int32_t heads, int32_t d_src, int32_t d_dst,
                             int64_t ld_src, int64_t ld_dst, void* stream) {
    B200_CHECK_ARG(rows >= 1 && heads >= 1 && d_src >= 8 && d_dst >= 8 && d_src % 8 == 0 && d_dst % 8 == 0 &&
                       ld_src % 8 == 0 && ld_dst % 8 == 0 && ld_src >= 1LL * heads * d_src && ld_dst >= 1LL * heads * d_dst,
                   "head_pad: head dims / row strides must be multiples of 8 elements");
    B200_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "head_pad: unaligned base");
    launch_pdl(head_pad_kernel, dim3(grid_for(rows * heads * (d_dst / 8), 256)), dim3(256), 0, ST, static_cast<const bf16*>(src),
               static_cast<bf16*>(dst), static_cast<long long>(rows), static_cast<int>(heads), static_cast<int>(d_src / 8),
               static_cast<int>(d_dst / 8), static_cast<long long>(ld_src), static_cast<long long>(ld_dst));
    B200_CHECK_LAUNCH("head_pad");
    return 0;
}
extern "C" int b200_upsample2x_fwd(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
    B200_CHECK_ARG(C % 8 == 0, "upsample: C %% 8 != 0");
    launch_pdl(upsample2x_fwd_kernel, dim3(grid_for(4LL * N * H * W * (C / 8), 256)), dim3(256), 0, ST, 
        static_cast<const bf16*>(x), static_cast<bf16*>(y), N, H, W, C / 8);
    B200_CHECK_LAUNCH("upsample2x_fwd");
    return 0;
}
extern "C" int b200_upsample2x_bwd(const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
    B200_CHECK_ARG(C % 8 == 0, "upsample: C %% 8 != 0");
    launch_pdl(upsample2x_bwd_kernel, dim3(grid_for(1LL * N * H * W * (C / 8), 256)), dim3(256), 0, ST, 
        static_cast<const bf16*>(dy), static_cast<bf16*>(dx), N, H, W, C / 8);
    B200_CHECK_LAUNCH("upsample2x_bwd");
    return 0;
}
extern "C" int b200_im2col3x3(const void* x, void* col, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride,
                              void* stream) {
    B200_CHECK_ARG(C % 8 == 0 && (stride == 1 || stride == 2), "im2col: unsupported C/stride");
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    launch_pdl(im2col3x3_kernel, dim3(grid_for(9LL * N * Ho * Wo * (C / 8), 256)), dim3(256), 0, ST, 
        static_cast<const bf16*>(x), static_cast<bf16*>(col), N, H, W, C / 8, stride, Ho, Wo);
    B200_CHECK_LAUNCH("im2col3x3");
    return 0;
}
extern "C" int b200_col2im3x3(const void* col, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride,
                              void* stream) {
    B200_CHECK_ARG(C % 8 == 0 && (stride == 1 || stride == 2), "col2im: unsupported C/stride");
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    launch_pdl(col2im3x3_kernel, dim3(grid_for(1LL * N * H * W * (C / 8), 256)), dim3(256), 0, ST, 
        static_cast<const bf16*>(col), static_cast<bf16*>(dx), N, H, W, C / 8, stride, Ho, Wo);
    B200_CHECK_LAUNCH("col2im3x3");
    return 0;
}
extern "C" int b200_colsum(const void* x, void* out, float* scratch, int32_t batch, int64_t hw, int32_t C, void* stream) {
    B200_CHECK_ARG(C % 8 == 0, "colsum: C %% 8 != 0");
    B200_CHECK_ARG(scratch != nullptr, "colsum: needs a batch*C fp32 scratch buffer");
    const long long n = static_cast<long long>(batch) * C;
    cudaError_t e = cudaMemsetAsync(scratch, 0, n * sizeof(float), ST);
    if (e != cudaSuccess) return set_error(3, "colsum: memset: %s", cudaGetErrorString(e));
    const int cblocks = (C + 127) / 128;
    int nz = static_cast<int>((2LL * kNumSMs + static_cast<long long>(cblocks) * batch - 1) / (static_cast<long long>(cblocks) * batch));
    const int max_z = static_cast<int>((hw + 63) / 64);
    if (nz > max_z) nz = max_z;
    if (nz < 1) nz = 1;
    const int rows_per_block = static_cast<int>((hw + nz - 1) / nz);
    launch_pdl(colsum_partial_kernel, dim3(cblocks, batch, nz), dim3(256), 0, ST, static_cast<const bf16*>(x), scratch, hw, C,
               rows_per_block);
    B200_CHECK_LAUNCH("colsum_partial");
    launch_pdl(colsum_finish_kernel, dim3(static_cast<int>((n + 255) / 256)), dim3(256), 0, ST, static_cast<const float*>(scratch),
               static_cast<bf16*>(out), n);
    B200_CHECK_LAUNCH("colsum_finish");
    return 0;
}
extern "C" int b200_lora_transpose_b(const void* params, void* bt, const int64_t* table, int32_t n_entries, void* stream) {
    B200_CHECK_ARG(n_entries >= 1 && params && bt && table, "lora_transpose_b: bad arguments");
    launch_pdl(lora_bt_kernel, dim3(n_entries), dim3(256), 0, ST, static_cast<const bf16*>(params), static_cast<bf16*>(bt),
               reinterpret_cast<const long long*>(table));
    B200_CHECK_LAUNCH("lora_transpose_b");
    return 0;
}
extern "C" int b200_lora_pack(const void* params, void* dst, const int64_t* table, int32_t n_entries, void* stream) {
    B200_CHECK_ARG(n_entries >= 1 && params && dst && table, "lora_pack: bad arguments");
    launch_pdl(lora_pack_kernel, dim3(n_entries), dim3(256), 0, ST, static_cast<const bf16*>(params), static_cast<bf16*>(dst),
               reinterpret_cast<const long long*>(table));
    B200_CHECK_LAUNCH("lora_pack");
    return 0;
}
extern "C" int b200_bicubic_fwd(const void* x, void* y, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                                int64_t ld_in, int64_t ld_out, void* stream) {
    B200_CHECK_ARG(B >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1 && C >= 1, "bicubic: empty extents");
    B200_CHECK_ARG(ld_in >= C && ld_out >= C, "bicubic: pixel stride smaller than C");
    const float sh = static_cast<float>(Hi) / Ho, sw = static_cast<float>(Wi) / Wo;
    launch_pdl(bicubic_fwd_kernel, dim3(grid_for(1LL * B * Ho * Wo * C, 256)), dim3(256), 0, ST, static_cast<const bf16*>(x),
               static_cast<bf16*>(y), B, Hi, Wi, Ho, Wo, C, static_cast<long long>(ld_in), static_cast<long long>(ld_out), sh, sw);
    B200_CHECK_LAUNCH("bicubic_fwd");
    return 0;
}
extern "C" int b200_bicubic_bwd(const void* dy, void* dx, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                                int64_t ld_dy, int64_t ld_dx, void* stream) {
    B200_CHECK_ARG(B >= 1 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1 && C >= 1, "bicubic: empty extents");
    B200_CHECK_ARG(ld_dy >= C && ld_dx >= C, "bicubic: pixel stride smaller than C");
    const float sh = static_cast<float>(Hi) / Ho, sw = static_cast<float>(Wi) / Wo;
    launch_pdl(bicubic_bwd_kernel, dim3(grid_for(1LL * B * Hi * Wi * C, 256)), dim3(256), 0, ST, static_cast<const bf16*>(dy),
               static_cast<bf16*>(dx), B, Hi, Wi, Ho, Wo, C, static_cast<long long>(ld_dy), static_cast<long long>(ld_dx), sh, sw);
    B200_CHECK_LAUNCH("bicubic_bwd");
    return 0;
}
extern "C" int b200_shift_stack9(const void* U, void* U9, int32_t N, int32_t H, int32_t W, int32_t r, int32_t ld_in,
                                 int32_t ld_out, void* stream) {
    B200_CHECK_ARG(ld_in >= r && ld_out >= 9 * r, "shift_stack9: bad leading dims");
    launch_pdl(shift_stack9_kernel, dim3(grid_for(9LL * N * H * W * r, 256)), dim3(256), 0, ST, 
        static_cast<const bf16*>(U), static_cast<bf16*>(U9), N, H, W, r, ld_in, ld_out);
    B200_CHECK_LAUNCH("shift_stack9");
    return 0;
}
extern "C" int b200_shift_sum9(const float* Z, void* T, int32_t N, int32_t H, int32_t W, int32_t r, int32_t ld_z, int32_t ld_t,
                               float alpha, void* stream) {
    B200_CHECK_ARG(N >= 1 && H >= 1 && W >= 1 && r >= 1 && ld_z >= 9 * r && ld_t >= r, "shift_sum9: bad extents / leading dims");
    launch_pdl(shift_sum9_kernel, dim3(grid_for(1LL * N * H * W * r, 256)), dim3(256), 0, ST, Z, static_cast<bf16*>(T),
               static_cast<int>(N), static_cast<int>(H), static_cast<int>(W), static_cast<int>(r), static_cast<int>(ld_z),
               static_cast<int>(ld_t), alpha);
    B200_CHECK_LAUNCH("shift_sum9");
    return 0;
}
extern "C" int b200_timestep_embedding(const float* t, void* out, int32_t n, int32_t dim, void* stream) {
    B200_CHECK_ARG(dim % 2 == 0, "timestep_embedding: odd dim");
    launch_pdl(timestep_embedding_kernel, dim3(grid_for(1LL * n * dim, 256)), dim3(256), 0, ST, t, static_cast<bf16*>(out), n, dim);
    B200_CHECK_LAUNCH("timestep_embedding");
    return 0;
}
