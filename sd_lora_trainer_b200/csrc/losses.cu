// Token-attention regulariser (trainer/loss.py:10-80 of the reference) and token-std regulariser (loss.py:196-233, 291-297)
// as kernels, forward AND gradient - the two losses that were torch ops + autograd inside the step's graph.
//
// The token-attention loss sees the stacked cross-attention maps [layers, B, h, w, 77] only through (a) their mean over
// layers, LM[b, p, t], and (b) the mean over layers and pixels, per_tok[b, t].  So:
//   tal_layer_sum_kernel : LS[b, p, t] = sum_l maps_l[b, p, t] (fp32; maps_l read in place from the layers' score buffers,
//                          the 64x64 ones after the bicubic resize kernel)
//   tal_partial_kernel   : 64 pixel slices per caption: per-token sums, the heat-map regularisers' sums (partials, no atomics)
//   tal_loss_kernel      : one CTA per caption: combines the slices in a fixed order -> per_tok, the four regularisers
//   tal_grad_kernel      : the gradient map G[b, p, t] = d loss / d maps_l[b, p, t] (identical for every layer l), bf16,
//                          scaled by grad_scale, over (pixels / 32) x B CTAs
//   tal_finalize_kernel  : combines the per-caption partials into the scalar.
// Rounding points follow the reference's bf16 graph: per_tok, relu(.)^2, the per-caption mean, reg0 and LM are rounded to
// bf16 where torch rounds them; reg1..3 are fp32 like the reference (.float() heat-maps).
#include "common.cuh"
#include "../../include/b200_lora.h"

namespace b200 {

constexpr int kTalMaxLayers = 64;
constexpr int kTalMaxTok = 8;
constexpr int kTalMaxText = 80;

struct TalLayers {
    const bf16* ptr[kTalMaxLayers];
    long long ld[kTalMaxLayers];     // row stride (elements) of layer l: rows are (b, p)
    int n;
};

__global__ void tal_layer_sum_kernel(const __grid_constant__ TalLayers L, float* __restrict__ LS, long long rows, int n_text,
                                     int ld_ls) {
    pdl_launch();
    pdl_wait();
    const long long total = rows * n_text;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = idx / n_text;
        const int t = static_cast<int>(idx - row * n_text);
        float acc = 0.f;
        for (int l = 0; l < L.n; ++l) acc += __bfloat162float(L.ptr[l][row * L.ld[l] + t]);   // fixed order: deterministic
        LS[row * ld_ls + t] = acc;
    }
}

constexpr int kTalSlices = 64;                  // pixel slices per caption (one CTA each) in the reduction pass
constexpr int kTalPP = kTalMaxText + 2 + kTalMaxTok;            // per-slice partials: per-token sums, s1, s2, heat-map sums
constexpr int kTalAux = kTalMaxText + kTalMaxTok + 8;           // per-caption: g0[t], mu_j - mu_bar, {valid, nv}

// which captions hold every trainable token (the reference skips the others, loss.py:41-44)
__device__ __forceinline__ bool tal_valid(const long long* __restrict__ ti_pos, int B, int n_tok, int b, int* nv_out) {
    int nv = 0;
    bool valid = true;
    for (int bb = 0; bb < B; ++bb) {
        bool ok = true;
        for (int j = 0; j < n_tok; ++j) ok = ok && (ti_pos[bb * n_tok + j] >= 0);
        nv += ok ? 1 : 0;
        if (bb == b) valid = ok;
    }
    if (nv_out) *nv_out = nv;
    return valid;
}

__device__ __forceinline__ float tal_mask_at(const float* __restrict__ mask, long long mask_sb, int b, int p, int w, float sh, float sw,
                                             int Hm, int Wm) {
    const int y = p / w, x = p - y * w;                               // F.interpolate(mode="nearest")
    const int ys = min(static_cast<int>(floorf(y * sh)), Hm - 1), xs = min(static_cast<int>(floorf(x * sw)), Wm - 1);
    return mask[b * mask_sb + static_cast<long long>(ys) * Wm + xs];
}

// Pass 1, grid (kTalSlices, B): partial sums over one pixel slice of one caption, written (not accumulated) to
// pp[b][slice][kTalPP] so that pass 2 can combine them in a fixed order.
__global__ void __launch_bounds__(256) tal_partial_kernel(const float* __restrict__ LS, int ld_ls, int n_layers, int B, int hw, int h, int w,
                                                          int n_text, const float* __restrict__ mask, long long mask_sb, int Hm, int Wm,
                                                          const long long* __restrict__ ti_pos, int n_tok, float* __restrict__ pp) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    __shared__ float s_acc[3][kTalMaxText];
    __shared__ int s_pos[kTalMaxTok];
    const int b = blockIdx.y, slice = blockIdx.x, tid = threadIdx.x;
    const int p0 = static_cast<int>(static_cast<long long>(hw) * slice / kTalSlices);
    const int p1 = static_cast<int>(static_cast<long long>(hw) * (slice + 1) / kTalSlices);
    const float* ls = LS + static_cast<long long>(b) * hw * ld_ls;
    float* out = pp + (static_cast<long long>(b) * kTalSlices + slice) * kTalPP;
    const bool valid = tal_valid(ti_pos, B, n_tok, b, nullptr);
    if (tid < n_tok) s_pos[tid] = valid ? static_cast<int>(ti_pos[b * n_tok + tid]) : 0;
    // (1) per-token sums: thread <-> (pixel sub-slice, token), coalesced over tokens
    {
        const int sub = tid / kTalMaxText, t = tid - sub * kTalMaxText;
        if (sub < 3) {
            float a = 0.f;
            if (t < n_text)
                for (int p = p0 + sub; p < p1; p += 3) a += ls[static_cast<long long>(p) * ld_ls + t];
            s_acc[sub][t] = a;
        }
        __syncthreads();
        if (tid < kTalMaxText) out[tid] = s_acc[0][tid] + s_acc[1][tid] + s_acc[2][tid];
    }
    // (2) heat-maps of the trainable tokens (captions that hold all of them)
    float s1 = 0.f, s2 = 0.f;
    const float sh = static_cast<float>(Hm) / h, sw = static_cast<float>(Wm) / w;
    for (int j = 0; j < n_tok; ++j) {
        float a = 0.f;
        if (valid)
            for (int p = p0 + tid; p < p1; p += blockDim.x) {
                const float hm = bfr(ls[static_cast<long long>(p) * ld_ls + s_pos[j]] / n_layers);
                const float mk = tal_mask_at(mask, mask_sb, b, p, w, sh, sw, Hm, Wm);
                const float r1 = fmaxf(hm * mk, 0.f), r2 = fmaxf(hm * (1.f - mk) + 10.f, 0.f);
                s1 += r1 * r1;
                s2 += r2 * r2;
                a += hm;
            }
        a = block_sum(a, red);
        if (tid == 0) out[kTalMaxText + 2 + j] = a;
        __syncthreads();
    }
    s1 = block_sum(s1, red);
    __syncthreads();
    s2 = block_sum(s2, red);
    if (tid == 0) {
        out[kTalMaxText + 0] = s1;
        out[kTalMaxText + 1] = s2;
    }
}

// Pass 2, one CTA per caption: combines the slice partials in slice order, then the per-caption scalars.
// part[b][8]: 0 att_L2 (bf16 value), 1 sum relu(hm*mk)^2, 2 sum relu(hm*(1-mk)+10)^2, 3 var_j(mean_p hm), 4 valid flag
// aux[b][kTalAux]: [0,80) d reg0 / d maps per token (already / (layers * pixels)); [80,88) mu_j - mu_bar; 88 valid; 89 nv
__global__ void __launch_bounds__(128) tal_loss_kernel(const float* __restrict__ pp, int n_layers, int B, int hw, int n_text,
                                                       const long long* __restrict__ tok_len, const long long* __restrict__ ti_pos,
                                                       int n_tok, float* __restrict__ part, float* __restrict__ aux) {
    pdl_launch();
    pdl_wait();
    __shared__ float s_pertok[kTalMaxText];     // bf16-rounded per-token mean
    __shared__ float s_sc[2 + kTalMaxTok];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* src = pp + static_cast<long long>(b) * kTalSlices * kTalPP;
    int nv = 0;
    const bool valid = tal_valid(ti_pos, B, n_tok, b, &nv);
    if (tid < kTalPP) {
        float a = 0.f;
        for (int s = 0; s < kTalSlices; ++s) a += src[s * kTalPP + tid];
        if (tid < kTalMaxText) s_pertok[tid] = bfr(a / (static_cast<float>(n_layers) * hw));
        else s_sc[tid - kTalMaxText] = a;
    }
    __syncthreads();
    const int len = static_cast<int>(tok_len[b]);
    const int cnt = max(len - 2, 0);                              // tokens 1 .. len-2
    float* ax = aux + static_cast<long long>(b) * kTalAux;
    if (tid < kTalMaxText) {
        const float pt = s_pertok[tid];
        const bool sel = tid >= 1 && tid < len - 1 && tid < n_text;
        // reg0 = 5 * mean_b( mean_t relu(per_tok)^2 ):  d/d maps = 5/B * 1/cnt * 2 relu(per_tok) / (layers * pixels)
        ax[tid] = (sel && cnt > 0) ? 5.f / B / cnt * 2.f * fmaxf(pt, 0.f) / (static_cast<float>(n_layers) * hw) : 0.f;
    }
    if (tid == 0) {
        float sq_sum = 0.f;
        for (int t = 1; t < len - 1 && t < n_text; ++t) {
            const float r = fmaxf(s_pertok[t], 0.f);
            sq_sum += bfr(r * r);
        }
        part[b * 8 + 0] = bfr(sq_sum / cnt);                      // 0/0 = NaN for an empty caption, as in the reference
        part[b * 8 + 4] = valid ? 1.f : 0.f;
        float mu_bar = 0.f, var = 0.f;
        if (valid) {
            for (int j = 0; j < n_tok; ++j) mu_bar += s_sc[2 + j] / hw;
            mu_bar /= n_tok;
            for (int j = 0; j < n_tok; ++j) {
                const float d = s_sc[2 + j] / hw - mu_bar;
                var += d * d;
                ax[kTalMaxText + j] = d;
            }
            var /= (n_tok - 1);                                   // torch.var: unbiased (NaN for one token, as in the reference)
        }
        part[b * 8 + 1] = valid ? s_sc[0] : 0.f;
        part[b * 8 + 2] = valid ? s_sc[1] : 0.f;
        part[b * 8 + 3] = var;
        ax[kTalMaxText + kTalMaxTok + 0] = valid ? 1.f : 0.f;
        ax[kTalMaxText + kTalMaxTok + 1] = static_cast<float>(nv);
    }
}

// Pass 3, grid (ceil(hw / 32), B): gradient map G[b, p, t] = grad_scale * d loss / d maps_l[b, p, t], the same for every layer
__global__ void __launch_bounds__(256) tal_grad_kernel(const float* __restrict__ LS, int ld_ls, int n_layers, int hw, int h, int w, int n_text,
                                                       const float* __restrict__ mask, long long mask_sb, int Hm, int Wm,
                                                       const long long* __restrict__ ti_pos, int n_tok, float grad_scale,
                                                       const float* __restrict__ aux, bf16* __restrict__ G, long long ld_g) {
    pdl_launch();
    pdl_wait();
    __shared__ float s_g0[kTalMaxText];
    __shared__ float s_dmu[kTalMaxTok];
    __shared__ int s_pos[kTalMaxTok];
    const int b = blockIdx.y, tid = threadIdx.x;
    const float* ax = aux + static_cast<long long>(b) * kTalAux;
    const bool valid = ax[kTalMaxText + kTalMaxTok + 0] != 0.f;
    const int nv = static_cast<int>(ax[kTalMaxText + kTalMaxTok + 1]);
    if (tid < kTalMaxText) s_g0[tid] = ax[tid];
    if (tid < n_tok) {
        s_dmu[tid] = valid ? ax[kTalMaxText + tid] : 0.f;
        s_pos[tid] = valid ? static_cast<int>(ti_pos[b * n_tok + tid]) : -1;
    }
    __syncthreads();
    const float* ls = LS + static_cast<long long>(b) * hw * ld_ls;
    const float sh = static_cast<float>(Hm) / h, sw = static_cast<float>(Wm) / w;
    const float inv_l = 1.f / n_layers;
    const float denom = static_cast<float>(max(nv, 1)) * n_tok * hw;
    const bool live = nv > 0;                                     // nv == 0: the reference returns a grad-less 0
    const int p0 = blockIdx.x * 32, p1 = min(p0 + 32, hw);
    for (long long idx = static_cast<long long>(p0) * ld_g + tid; idx < static_cast<long long>(p1) * ld_g; idx += blockDim.x) {
        const int p = static_cast<int>(idx / ld_g), t = static_cast<int>(idx - static_cast<long long>(p) * ld_g);
        float gsum = 0.f;
        if (live && t < n_text) {
            gsum = s_g0[t];
            if (valid) {
                for (int j = 0; j < n_tok; ++j) {
                    if (s_pos[j] == t) {
                        const float hm = bfr(ls[static_cast<long long>(p) * ld_ls + t] / n_layers);
                        const float mk = tal_mask_at(mask, mask_sb, b, p, w, sh, sw, Hm, Wm);
                        float d = (2.f * fmaxf(hm * mk, 0.f) * mk + 4.f * fmaxf(hm * (1.f - mk) + 10.f, 0.f) * (1.f - mk)) / denom;
                        d += 2.f * s_dmu[j] / (n_tok - 1) / hw / nv;
                        gsum += d * inv_l;
                    }
                }
            }
        }
        G[(static_cast<long long>(b) * hw + p) * ld_g + t] = __float2bfloat16_rn(gsum * grad_scale);
    }
}

__global__ void tal_finalize_kernel(const float* __restrict__ part, int B, int n_tok, int hw, float* __restrict__ loss_out) {
    pdl_launch();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float a = 0.f, s1 = 0.f, s2 = 0.f, v = 0.f, nv = 0.f;
    for (int b = 0; b < B; ++b) {
        a += part[b * 8 + 0];
        s1 += part[b * 8 + 1];
        s2 += part[b * 8 + 2];
        v += part[b * 8 + 3];
        nv += part[b * 8 + 4];
    }
    if (nv == 0.f) {
        loss_out[0] = 0.f;
        return;
    }
    const float reg0 = bfr(5.f * bfr(a / B));
    const float denom = nv * n_tok * hw;
    loss_out[0] = reg0 + s1 / denom + 2.f * s2 / denom + v / nv;
}

// Token-std regulariser: loss_e = mean_rows( (mu_t - std(row))^2 / var_t ) per text encoder e (torch.std: unbiased, computed
// on the bf16 row and rounded to bf16), total = mean_e loss_e; gradient added to the flat fp32 buffer:
//   d/dx_i = coeff * (1/n_enc) * (1/n_rows) * (-2 (mu_t - s) / var_t) * (x_i - mean) / ((D - 1) s)
struct StdRegArgs {
    const bf16* rows[2];
    float* grads[2];
    float mu_t[2], var_t[2];
    int dim[2];
    int n_enc, n_rows;
    float coeff;            // 0.01 / grad-accumulation: what multiplies the loss in the step
    float* loss_out;        // += mean over encoders (unweighted, like losses['token_std_loss'])
};

__global__ void __launch_bounds__(256) token_std_kernel(const __grid_constant__ StdRegArgs a) {
    pdl_launch();
    pdl_wait();
    __shared__ float red[32];
    const int e = blockIdx.x / a.n_rows, r = blockIdx.x - e * a.n_rows;
    const int D = a.dim[e];
    const bf16* x = a.rows[e] + static_cast<long long>(r) * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s += __bfloat162float(x[i]);
    s = block_sum(s, red);
    const float mean = s / D;
    __syncthreads();
    float q = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float d = __bfloat162float(x[i]) - mean;
        q += d * d;
    }
    q = block_sum(q, red);
    const float sd_f = sqrtf(q / (D - 1));
    const float sd = bfr(sd_f);                                    // torch: std of a bf16 tensor is a bf16 value
    const float dev = bfr(a.mu_t[e] - sd);
    const float term = bfr(bfr(dev * dev) / a.var_t[e]);
    if (threadIdx.x == 0) atomicAdd(a.loss_out, term / a.n_rows / a.n_enc);
    if (a.grads[e] != nullptr && sd_f > 0.f) {
        const float k = a.coeff / a.n_enc / a.n_rows * (-2.f * dev / a.var_t[e]) / ((D - 1) * sd_f);
        float* g = a.grads[e] + static_cast<long long>(r) * D;
        for (int i = threadIdx.x; i < D; i += blockDim.x) g[i] += k * (__bfloat162float(x[i]) - mean);
    }
}

}  // namespace b200

using namespace b200;
#define ST static_cast<cudaStream_t>(stream)

extern "C" int64_t b200_token_attention_loss_floats(int32_t B, int32_t h, int32_t w, int32_t n_text) {
    return static_cast<int64_t>(B) * h * w * n_text + static_cast<int64_t>(B) * (8 + kTalAux + kTalSlices * kTalPP);
}

extern "C" int b200_token_attention_loss(const void* const* maps, const int64_t* lds, int32_t n_layers, int32_t B, int32_t h,
                                         int32_t w, int32_t n_text, const float* mask, int64_t mask_sb, int32_t Hm, int32_t Wm,
                                         const int64_t* tok_len, const int64_t* ti_pos, int32_t n_tok, float grad_scale,
                                         float* ws, int64_t ws_floats, float* loss_out, void* G, int64_t ld_g, void* stream) {
    B200_CHECK_ARG(maps && lds && n_layers >= 1 && n_layers <= kTalMaxLayers, "token_attention_loss: 1..%d layers", kTalMaxLayers);
    B200_CHECK_ARG(B >= 1 && h >= 1 && w >= 1 && n_text >= 1 && n_text <= kTalMaxText && n_tok >= 1 && n_tok <= kTalMaxTok,
                   "token_attention_loss: bad extents (text length <= %d, trainable tokens <= %d)", kTalMaxText, kTalMaxTok);
    B200_CHECK_ARG(mask && tok_len && ti_pos && ws && loss_out && Hm >= 1 && Wm >= 1, "token_attention_loss: null argument");
    const int hw = h * w;
    const int ld_ls = n_text;
    const long long need = b200_token_attention_loss_floats(B, h, w, n_text);
    B200_CHECK_ARG(ws_floats >= need, "token_attention_loss: workspace needs %lld floats", need);
    B200_CHECK_ARG(G == nullptr || ld_g >= n_text, "token_attention_loss: gradient row stride smaller than the text length");
    TalLayers L;
    memset(&L, 0, sizeof(L));
    L.n = n_layers;
    for (int l = 0; l < n_layers; ++l) {
        B200_CHECK_ARG(maps[l] != nullptr && lds[l] >= n_text, "token_attention_loss: layer %d: bad pointer / row stride", l);
        L.ptr[l] = static_cast<const bf16*>(maps[l]);
        L.ld[l] = lds[l];
    }
    float* LS = ws;
    float* part = ws + static_cast<long long>(B) * hw * ld_ls;
    const long long rows = static_cast<long long>(B) * hw;
    launch_pdl(tal_layer_sum_kernel, dim3(grid_for(rows * n_text, 256)), dim3(256), 0, ST, L, LS, rows, static_cast<int>(n_text), ld_ls);
    B200_CHECK_LAUNCH("tal_layer_sum");
    float* aux = part + 8LL * B;
    float* pp = aux + static_cast<long long>(B) * kTalAux;
    const long long* tp = reinterpret_cast<const long long*>(ti_pos);
    launch_pdl(tal_partial_kernel, dim3(kTalSlices, B), dim3(256), 0, ST, static_cast<const float*>(LS), ld_ls, static_cast<int>(n_layers),
               static_cast<int>(B), hw, static_cast<int>(h), static_cast<int>(w), static_cast<int>(n_text), mask, static_cast<long long>(mask_sb),
               static_cast<int>(Hm), static_cast<int>(Wm), tp, static_cast<int>(n_tok), pp);
    B200_CHECK_LAUNCH("tal_partial");
    launch_pdl(tal_loss_kernel, dim3(B), dim3(128), 0, ST, static_cast<const float*>(pp), static_cast<int>(n_layers), static_cast<int>(B), hw,
               static_cast<int>(n_text), reinterpret_cast<const long long*>(tok_len), tp, static_cast<int>(n_tok), part, aux);
    B200_CHECK_LAUNCH("tal_loss");
    if (G != nullptr) {
        launch_pdl(tal_grad_kernel, dim3((hw + 31) / 32, B), dim3(256), 0, ST, static_cast<const float*>(LS), ld_ls, static_cast<int>(n_layers),
                   hw, static_cast<int>(h), static_cast<int>(w), static_cast<int>(n_text), mask, static_cast<long long>(mask_sb),
                   static_cast<int>(Hm), static_cast<int>(Wm), tp, static_cast<int>(n_tok), grad_scale, static_cast<const float*>(aux),
                   static_cast<bf16*>(G), static_cast<long long>(ld_g));
        B200_CHECK_LAUNCH("tal_grad");
    }
    launch_pdl(tal_finalize_kernel, dim3(1), dim3(32), 0, ST, static_cast<const float*>(part), static_cast<int>(B),
               static_cast<int>(n_tok), hw, loss_out);
    B200_CHECK_LAUNCH("tal_finalize");
    return 0;
}

extern "C" int b200_token_std_loss(const void* rows0, const void* rows1, float* grads0, float* grads1, int32_t n_enc, int32_t n_rows,
                                   int32_t dim0, int32_t dim1, float mu_t0, float var_t0, float mu_t1, float var_t1, float coeff,
                                   float* loss_out, void* stream) {
    B200_CHECK_ARG((n_enc == 1 || n_enc == 2) && n_rows >= 1 && rows0 && loss_out && dim0 >= 2 && (n_enc == 1 || (rows1 && dim1 >= 2)),
                   "token_std_loss: bad arguments");
    StdRegArgs a;
    memset(&a, 0, sizeof(a));
    a.rows[0] = static_cast<const bf16*>(rows0);
    a.rows[1] = static_cast<const bf16*>(rows1);
    a.grads[0] = grads0;
    a.grads[1] = grads1;
    a.mu_t[0] = mu_t0;
    a.var_t[0] = var_t0;
    a.mu_t[1] = mu_t1;
    a.var_t[1] = var_t1;
    a.dim[0] = dim0;
    a.dim[1] = dim1;
    a.n_enc = n_enc;
    a.n_rows = n_rows;
    a.coeff = coeff;
    a.loss_out = loss_out;
    launch_pdl(token_std_kernel, dim3(n_enc * n_rows), dim3(256), 0, ST, a);
    B200_CHECK_LAUNCH("token_std_loss");
    return 0;
}
