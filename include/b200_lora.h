/* b200_lora.h - C ABI of the B200-native LoRA / textual-inversion training-step kernels.
 *
 * The reference (edenartlab/sd-lora-trainer) is pure Python and has NO FFI / plugin interface of its own
 * (SURVEY.md section 0.1, 8b); its "boundary" is the set of Python call sites listed next to each entry point
 * below.  The host-side mirror of those call sites lives in sd_lora_trainer_b200/trainer/ and reaches this
 * library through ctypes (sd_lora_trainer_b200/_lib.py); INTEGRATION.md shows the binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch types.  The caller (PyTorch) owns every buffer.
 *   - every function returns 0 on success, otherwise a non-zero code; b200_last_error() gives the message
 *     (thread-local).  No exception crosses the ABI.
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*); nothing synchronises.
 *   - bf16 tensors are `uint16_t`-sized elements (torch.bfloat16 storage); all strides are in ELEMENTS.
 */
#ifndef B200_LORA_H
#define B200_LORA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_LORA_ABI_VERSION 1

int b200_version(void);
const char* b200_last_error(void);
/* number of kernels this library has launched since load (bench.py's `gpu_launches` evidence) */
long long b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05 + TMA).  Replaces the cuBLASLt / cuDNN calls behind
 *   peft lora.Linear.forward / lora.Conv2d.forward            (trainer/optimizer.py:84-95 -> [3P] peft 0.10.0)
 *   attn.to_q/to_k/to_v/to_out[0], QK^T, the head-summed score  (trainer/ti_cross_attn_loss.py:167-220)
 *   diffusers FeedForward / proj_in / proj_out / ResnetBlock2D convs (main.py:329-336 -> [3P] diffusers 0.29.2)
 *   and their autograd backward (main.py:363): dX, and dA/dB of the LoRA factors only (no dW).
 *
 *   D[b1][b0][m, n] = alpha * sum_seg sum_k A_seg[m, k] * B_seg[n, k]  (+ bias[n]) (+ R[m, n])
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* ptr;     /* bf16 */
    int64_t rows;        /* extent of the strided dim  (K-major: M or N index;  MN-major: K index) */
    int64_t inner;       /* extent of the contiguous dim (K-major: K index;     MN-major: M or N index) */
    int64_t row_stride;  /* elements; multiple of 8 (16 B) */
    int64_t sb0, sb1;    /* batch strides in elements (multiples of 8); ignored when batched == 0 */
    int32_t mn_major;    /* 0: K-major, 1: MN-major */
    int32_t batched;     /* 0: shared by all batches (e.g. a weight) */
} b200_operand_t;

typedef struct {
    int32_t M, N;                 /* per-batch output extent */
    int32_t num_seg;              /* 1 or 2 K-segments accumulated into the same tile */
    int32_t K[2];                 /* reduction length per segment (conv: ignored for segment 0) */
    b200_operand_t A[2], B[2];
    int32_t nb0, nb1;             /* batch extents (>= 1) */
    int32_t splits;               /* split-K (>= 1); > 1 requires d_fp32 && d_atomic, num_seg == 1, conv == 0 */
    int32_t block_n;              /* 0: library picks the tile width for whole waves */
    int32_t pair_mode;            /* 0: library picks; 1: force the CTA-pair (cta_group::2, 256-row tiles) kernel;
                                     -1: never use it.  The pair kernel covers one K-major A segment, no conv / batch /
                                     split-K, row-major output (incl. the fused side path). */
    /* implicit 3x3/pad1/stride1 convolution on segment 0: A[0].ptr is NHWC [conv_N, conv_H, conv_W, conv_C],
       M must equal conv_N*conv_H*conv_W, B[0] is [N rows, 9*b_tap_k (+...)] K-major. */
    int32_t conv;
    int32_t conv_N, conv_H, conv_W, conv_C;
    int32_t b_tap_k, b_tap_n;     /* B coordinate advance per tap: k += tap*b_tap_k, n += tap*b_tap_n */
    /* output */
    void* D;
    int32_t d_fp32, d_atomic;
    int64_t d_sm, d_sn, d_sb0, d_sb1;
    float alpha;
    const void* bias;             /* bf16 or NULL */
    int32_t bias_rows;            /* rows sharing one bias vector (0: all rows) */
    int64_t bias_sb;
    const void* R;                /* bf16 residual or NULL */
    int64_t r_sm, r_sn, r_sb0, r_sb1;
    /* fused low-rank side path (num_seg == 1, no conv, no batch, no split-K):
         D += (side_alpha * A[0] . S^T) . B2^T ,  T_out[m, 0..side_r) = bf16(side_alpha * A[0] . S^T)
       forward : S = lora_A [r, K] (K-major),            B2 = lora_B [N, r] (K-major)
       dgrad   : S = lora_B [N_layer, r] (MN-major),     B2 = lora_A [r, K_layer] (MN-major)
       so  W.x + s.B.(A.x)  and its input gradient are ONE launch each, and T/U come out for the dA/dB GEMMs. */
    int32_t side;
    int32_t side_r;               /* rank (<= 32) */
    b200_operand_t S, B2;
    float side_alpha;
    void* T_out;                  /* bf16 [M, t_ld] or NULL */
    int64_t t_ld;
    /* group mode (num_seg == 2, fp32 atomic outputs): the two segments are two INDEPENDENT problems of the same
       M x N x K tiling sharing one launch - the dB = dY^T.T and dA = U^T.X weight-gradient GEMMs of one LoRA layer.
       Segment 0 accumulates into D (d_sm, d_sn), segment 1 into D2 (d2_sm, d2_sn). */
    int32_t group;
    void* D2;
    int64_t d2_sm, d2_sn;
    /* 1: the B operands (B[*], S, B2 - frozen weights and LoRA factors) are NOT written by the kernel that precedes
       this launch on the stream, so the pair kernel may fetch them before its programmatic-dependency wait. */
    int32_t b_static;
    /* GEGLU backward fused into the epilogue (CTA-pair kernel, bf16 row-major output, N % 32 == 0, no bias / residual /
       side path): the product is dy[M, N], the input gradient of diffusers' GEGLU output (FeedForward.net.2's dgrad); with
       geglu_h = the saved projection h = [value | gate] (bf16 [M, 2N], row stride geglu_h_ld) the kernel writes
       D[m, j] = dy * gelu(gate) and D[m, N + j] = dy * value * gelu'(gate) - D is [M, 2N] - and dy is never stored. */
    const void* geglu_h;
    int64_t geglu_h_ld;
    /* GEGLU forward fused into the epilogue (CTA-pair kernel, bf16 output, N % 256 == 0, no residual / side path): the B
       operand is FeedForward.net.0.proj's weight with its rows interleaved in blocks of 128 (128 value rows, the matching
       128 gate rows, ...; bias likewise), D [M, N] receives the projection in that interleaved column layout (kept for the
       backward) and geglu_y [M, N/2] (row stride geglu_y_ld) = value * gelu(gate), diffusers' GEGLU.forward. */
    void* geglu_y;
    int64_t geglu_y_ld;
} b200_gemm_t;

int b200_gemm(const b200_gemm_t* desc, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Row softmax for the (unfused, HBM-resident) attention path; replaces F.scaled_dot_product_attention's
 * softmax (trainer/ti_cross_attn_loss.py:197-199).  S is fp32 [rows, ld_s] already scaled; P is bf16
 * [rows, ld_p]; columns >= cols are written as 0.
 * --------------------------------------------------------------------------------------------------------- */
int b200_softmax_fwd(const float* S, void* P, int64_t rows, int32_t cols, int64_t ld_s, int64_t ld_p, void* stream);
/* dS = P * (dP - rowsum(P*dP)); dP fp32 [rows, ld_dp] -> dS bf16 [rows, ld_p] */
int b200_softmax_bwd(const void* P, const float* dP, void* dS, int64_t rows, int32_t cols, int64_t ld_p,
                     int64_t ld_dp, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Token-attention regulariser, forward and gradient (replaces compute_token_attention_loss, trainer/loss.py:10-80, and its
 * autograd backward): maps[l] -> bf16 [B, h*w, 77] cross-attention score maps of hooked layer l (row stride lds[l]
 * elements), already at the common (smallest) resolution h x w (process_and_stack_attention_scores,
 * trainer/ti_cross_attn_loss.py:239-268: larger maps go through b200_bicubic_fwd first).  mask: fp32 [B, Hm, Wm] with
 * batch stride mask_sb elements (mask[:, 0], resized in-kernel like F.interpolate(mode="nearest")); tok_len[b] = len(tokenizer.encode(caption_b)),
 * ti_pos[b, j] = position of trainable token j in caption b or -1.  loss_out[0] <- the regulariser; G (bf16
 * [B, h*w, ld_g], may be NULL) <- grad_scale * d loss / d maps[l] - the SAME map for every layer l (the loss sees only the
 * layer mean).  ws: fp32 workspace of b200_token_attention_loss_floats(B, h, w, n_text) floats.  n_layers <= 64, 77 <= 80 text
 * positions, <= 8 tokens.
 * --------------------------------------------------------------------------------------------------------- */
int64_t b200_token_attention_loss_floats(int32_t B, int32_t h, int32_t w, int32_t n_text);
int b200_token_attention_loss(const void* const* maps, const int64_t* lds, int32_t n_layers, int32_t B, int32_t h, int32_t w,
                              int32_t n_text, const float* mask, int64_t mask_sb, int32_t Hm, int32_t Wm, const int64_t* tok_len,
                              const int64_t* ti_pos, int32_t n_tok, float grad_scale, float* ws, int64_t ws_floats,
                              float* loss_out, void* G, int64_t ld_g, void* stream);
/* Token-std regulariser (ConditioningRegularizer.apply_regularization -> DistributionLoss.compute_std_loss,
 * trainer/loss.py:222-231, 291-297) over the n_rows trainable embedding rows of 1 or 2 text encoders:
 * loss_out[0] += mean_e mean_rows (mu_t[e] - std(row))^2 / var_t[e]; grads_e (fp32 [n_rows, dim_e], may be NULL) +=
 * coeff * d/d rows of that mean (coeff = 0.01 / gradient accumulation in the step). */
int b200_token_std_loss(const void* rows0, const void* rows1, float* grads0, float* grads1, int32_t n_enc, int32_t n_rows,
                        int32_t dim0, int32_t dim1, float mu_t0, float var_t0, float mu_t1, float var_t1, float coeff,
                        float* loss_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Batched LoRA weight gradients (replaces autograd's dW of every lora_A / lora_B nn.Linear of a transformer block,
 * main.py:363 -> [3P] peft lora.Linear): out_p[n, j] += sum_m X_p[m, n] * Y_p[m, j] for up to 32 independent problems in ONE
 * launch - dB = dY^T.T (X = dY [M, N], Y = T = s.x.A^T [M, r], out = dB [N, r]) and dA = U^T.x (X = x [M, K],
 * Y = U = s.dY.B [M, r], out[k, j] = dA[j, k]).  bf16 operands with 16-byte aligned rows, r <= 32; fp32 atomic output
 * addressed out[n*out_sn + j*out_sj].  No tensor maps: any operand addresses batch.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
    const void* X;                /* bf16 [M, Nout], row stride ld_x */
    const void* Y;                /* bf16 [M, >= roundup8(r)], row stride ld_y */
    float* out;                   /* fp32, accumulated in place */
    int64_t ld_x, ld_y, out_sn, out_sj;
    int32_t M, Nout, r;
} b200_wgrad_problem_t;
int b200_lora_wgrad_batch(const b200_wgrad_problem_t* problems, int32_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused attention for head_dim 64 (every SDXL head): replaces F.scaled_dot_product_attention forward and its
 * autograd backward (trainer/ti_cross_attn_loss.py:197-199, diffusers AttnProcessor2_0) without ever writing an
 * [L, Lk] tensor to HBM.  q: [B*L, ld], k/v: [B*Lk, ld] bf16 with head h at columns [h*64, h*64+64); o likewise.
 * lse: [B, H, L] fp32 (natural log), saved by the forward for the backward.
 * Row strides (elements, multiples of 8, >= H*64): forward ld for q / k / v and ld_o for o; backward ld_qkv for q / k / v,
 * ld_o for o and d_o, ld_d for dq / dk / dv - so q | k | v (and their gradients) may be column slices of one [rows, 3C]
 * buffer, the output of the fused q|k|v projection.
 * Backward workspaces (caller-owned): delta_ws fp32 [B*H*L], dq_acc_ws fp32 [B*L*H*64] (NULL when Lk <= 128).
 * split_ws (optional, may be NULL): fp32 [2*B*Lk*ld_d + B*H*ceil(Lk/128)] that is ZERO on entry and is left zero by the
 * kernel; with it, a problem whose (key blocks, H, B) grid would leave SMs idle (Lk <= 128: every cross-attention layer) or
 * quantise badly over the 148 SMs (self-attention at L = 1024) splits its query blocks over more CTAs (partial dK / dV
 * summed with fp32 atomics, the last CTA of a (b, h, key block) rounds them to bf16).
 * dsc (optional, may be NULL): bf16 [B, L, ld_dsc], the gradient of the head-summed pre-softmax scores captured by
 * DAAMLossAttnProcessor2_0 (trainer/ti_cross_attn_loss.py:201-212) - the score is sum_h scale*q_h.k_h, so its gradient
 * joins dS inside the kernel (dQ_h += scale*dsc.K_h, dK_h += scale*dsc^T.Q_h); columns [0, dsc_cols) of a row are read.
 * --------------------------------------------------------------------------------------------------------- */
int b200_flash_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t H,
                        int32_t L, int32_t Lk, int64_t ld, int64_t ld_o, float scale, void* stream);
int b200_flash_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                        float* delta_ws, float* dq_acc_ws, void* dq, void* dk, void* dv, int32_t B, int32_t H,
                        int32_t L, int32_t Lk, int64_t ld_qkv, int64_t ld_o, int64_t ld_d, float scale, float* split_ws,
                        int64_t split_ws_floats, const void* dsc, int64_t ld_dsc, int32_t dsc_cols, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Normalisation / activation kernels on NHWC ([rows, C]) bf16 activations; replace ATen GroupNorm / LayerNorm /
 * SiLU / GELU kernels under diffusers ResnetBlock2D, Transformer2DModel, BasicTransformerBlock, GEGLU.
 * Affine parameters are frozen, so the backward kernels emit dX only.
 * --------------------------------------------------------------------------------------------------------- */
/* stats: [batch, groups, 2] fp32 (mean, rstd) followed by fp64 scratch (the per-(image, group) sums and the per-block
   partial sums they are added from, in block order - no atomics, so the statistics are bit-reproducible); the caller
   allocates b200_groupnorm_stats_floats(batch, hw, C, groups) floats, 8-byte aligned; the backward reuses the scratch. */
int64_t b200_groupnorm_stats_floats(int32_t batch, int64_t hw, int32_t C, int32_t groups);
int b200_groupnorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* stats, int32_t batch,
                       int64_t hw, int32_t C, int32_t groups, float eps, int32_t silu, void* stream);
/* dx = dGroupNorm(dy) (+ dres when non-NULL: the gradient arriving through the residual/shortcut branch) */
int b200_groupnorm_bwd(const void* dy, const void* x, const void* gamma, const void* beta, const float* stats,
                       const void* dres, void* dx, int32_t batch, int64_t hw, int32_t C, int32_t groups, int32_t silu,
                       void* stream);
/* Affine-parameter gradients of a frozen-no-more norm layer (dense / full-UNet fine-tune backward, BASELINE config 5;
   main.py:143-148 `unet.requires_grad_(True)`): dgamma[c] += sum dz * xhat, dbeta[c] += sum dz (fp32, accumulating).
   groups > 0: GroupNorm over x [rows = batch*hw, C] with stats [(batch*groups) x (mean, rstd)] as groupnorm_fwd wrote
   them, silu = 1 when the layer fused SiLU (dz = dy * silu'(xhat*gamma+beta)); groups == 0: LayerNorm, stats per row. */
int b200_norm_param_grad(const void* dy, const void* x, const void* gamma, const void* beta, const float* stats,
                         float* dgamma, float* dbeta, int64_t rows, int64_t hw, int32_t C, int32_t groups, int32_t silu,
                         void* stream);
/* stats: [rows, 2] fp32 (mean, rstd) */
int b200_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* stats, int64_t rows,
                       int32_t C, float eps, void* stream);
int b200_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* stats, const void* dres, void* dx,
                       int64_t rows, int32_t C, void* stream);
/* h: [rows, 2*inner] = (value | gate); y = value * gelu(gate) (exact erf gelu, rounded like torch's op-by-op bf16) */
/* interleave = 0: h = [value (inner) | gate (inner)] as diffusers' GEGLU chunks it; interleave = il > 0: blocks of il value
   columns alternate with blocks of il gate columns (the layout the FF up-projection writes with its fused GEGLU epilogue) */
int b200_geglu_fwd(const void* h, void* y, int64_t rows, int32_t inner, int32_t interleave, void* stream);
int b200_geglu_bwd(const void* dy, const void* h, void* dh, int64_t rows, int32_t inner, int32_t interleave, void* stream);
int b200_silu_fwd(const void* x, void* y, int64_t n, void* stream);
int b200_silu_bwd(const void* dy, const void* x, void* dx, int64_t n, void* stream);
/* CLIP text-encoder MLP activation ([3P] transformers CLIPMLP.activation_fn on the get_conditioning_signals path,
   trainer/inference.py:131-177): kind 0 = exact (erf) GELU, kind 1 = quick_gelu x * sigmoid(1.702 x); bf16 in / out,
   n elements; the backward reads the pre-activation x:  dx = dy * act'(x). */
int b200_act_fwd(const void* x, void* y, int64_t n, int32_t kind, void* stream);
int b200_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int32_t kind, void* stream);
/* y = a + b (+ c), bf16 */
int b200_add(const void* a, const void* b, const void* c, void* y, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Layout helpers (NHWC).
 * --------------------------------------------------------------------------------------------------------- */
/* Head re-pitch around the fused attention kernel (replaces the head split of [3P] diffusers Attention.head_to_batch_dim
   on the AttnProcessor2_0 / DAAMLossAttnProcessor2_0 path, trainer/ti_cross_attn_loss.py:167-175, for head dims the
   tcgen05 kernel does not take natively): src [rows, heads * d_src] -> dst [rows, heads * d_dst]; per head the first
   min(d_src, d_dst) channels are copied, the rest zero-filled.  d_src, d_dst, ld_* multiples of 8 elements. */
int b200_head_pad(const void* src, void* dst, int64_t rows, int32_t heads, int32_t d_src, int32_t d_dst, int64_t ld_src,
                  int64_t ld_dst, void* stream);
int b200_upsample2x_fwd(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
int b200_upsample2x_bwd(const void* dy, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* 3x3 pad-1 im2col with stride: col[N*Ho*Wo, 9*C], tap-major (kh,kw,c) */
int b200_im2col3x3(const void* x, void* col, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, void* stream);
int b200_col2im3x3(const void* col, void* dx, int32_t N, int32_t H, int32_t W, int32_t C, int32_t stride, void* stream);
/* U9[p, tap*r + j] = U[p - offset(tap), j] (zero outside the image): lets the conv-LoRA backward run as plain GEMMs */
int b200_shift_stack9(const void* U, void* U9, int32_t N, int32_t H, int32_t W, int32_t r, int32_t ld_in, int32_t ld_out,
                      void* stream);
/* T[p, j] = bf16(alpha * sum_tap Z[p + off(tap), tap*r + j]): the 3x3 lora_A convolution of [3P] peft lora.Conv2d.forward
   (trainer/optimizer.py:84-95) from ONE plain GEMM Z = X . A_taps^T (fp32 [N*H*W, ld_z >= 9r]) instead of a 16-wide
   implicit convolution; zero padding at the image border.  T: bf16 [N*H*W, ld_t >= r]. */
int b200_shift_sum9(const float* Z, void* T, int32_t N, int32_t H, int32_t W, int32_t r, int32_t ld_z, int32_t ld_t,
                    float alpha, void* stream);
/* out[b, c] = sum_p x[b, p, c]  (x: [batch, hw, C] bf16, C % 8 == 0) - gradient of the per-image time-embedding bias.
 * scratch: batch*C fp32 (zeroed by the call; rows are split over CTAs and combined with fp32 reductions). */
int b200_colsum(const void* x, void* out, float* scratch, int32_t batch, int64_t hw, int32_t C, void* stream);
/* K-major copies of all LoRA-B factors in one launch: for every table row (off_B, off_Bt, N, rs) - offsets in
 * elements into the flat buffers - bt[off_Bt + j*N + n] = params[off_B + n*rs + j].  The fused input-gradient GEMM
 * (dX = dY.W + (s.dY.B).A, peft lora.Linear backward) reads the copy as its side operand. table: device int64 [n, 4]. */
int b200_lora_transpose_b(const void* params, void* bt, const int64_t* table, int32_t n_entries, void* stream);
/* Derived LoRA-B operands of the fused q|k|v projection (to_q / to_k / to_v of a self-attention block as ONE N = 3C GEMM
 * with a rank-3r side path, [3P] peft lora.Linear x3): per table row (off_B, off_dst, N, rs, dst_ld, transpose) - offsets
 * in elements - dst[off_dst + n*dst_ld + j] = B[n, j] (transpose 0: a diagonal block of the [3C, 3r] block-diagonal
 * factor) or dst[off_dst + j*dst_ld + n] = B[n, j] (transpose 1: its K-major transpose [3r, 3C] for the input gradient).
 * table: device int64 [n, 6]; one launch per step refreshes every block. */
int b200_lora_pack(const void* params, void* dst, const int64_t* table, int32_t n_entries, void* stream);
/* Bicubic resize of channels-last bf16 maps: F.interpolate(x, size=(Ho, Wo), mode="bicubic") exactly as the reference
 * applies it to the captured cross-attention maps (trainer/ti_cross_attn_loss.py:262-266: align_corners=False, no
 * antialias, A = -0.75, clamped taps).  x: [B, Hi, Wi, C] with pixel stride ld_in, y: [B, Ho, Wo, C] with pixel stride
 * ld_out (elements).  _bwd is the adjoint: dx[B, Hi, Wi, C] = J^T dy, written (not accumulated), gather form. */
int b200_bicubic_fwd(const void* x, void* y, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                     int64_t ld_in, int64_t ld_out, void* stream);
int b200_bicubic_bwd(const void* dy, void* dx, int32_t B, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, int32_t C,
                     int64_t ld_dy, int64_t ld_dx, void* stream);
/* sinusoidal timestep embedding (flip_sin_to_cos, shift 0): t fp32 [n] -> bf16 [n, dim] */
int b200_timestep_embedding(const float* t, void* out, int32_t n, int32_t dim, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Step prologue / loss / optimizer (main.py:311-326, trainer/loss.py:127-170, trainer/optimizer.py:265-275).
 * --------------------------------------------------------------------------------------------------------- */
/* PreprocessedDataset.__getitem__ (trainer/dataset.py:181-193): latent_dist.sample() * vae.config.scaling_factor on the
   cached DiagonalGaussianDistribution parameters, fp32; eps is the injected N(0,1) draw.
   out = (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps) * scaling_factor */
int b200_latent_sample(const float* mean, const float* logvar, const float* eps, float scaling_factor, float* out, int64_t n,
                       void* stream);
/* noise(bf16, in/out) += offset_scale * offset[b, c] ; noisy = sqrt(acp[t]) * bf16(latent) + sqrt(1 - acp[t]) * noise
   with acp rounded to bf16 before the sqrt, as DDPMScheduler.add_noise does.  latent fp32 NCHW, outputs NCHW bf16
   plus an NHWC(8-channel padded) copy of `noisy` for the conv_in TMA path. */
int b200_noise_prologue(const float* latent, void* noise, const float* offset, float offset_scale,
                        const float* alphas_cumprod, const int64_t* timesteps, void* noisy_nchw, void* noisy_nhwc8,
                        int32_t B, int32_t C, int32_t HW, void* stream);
/* min-SNR weights w[b] = (min(snr_b, gamma)/snr_b) / mean_b(.) from the fp32 alphas_cumprod table (loss.py:83-106,145-161) */
int b200_snr_weights(const float* alphas_cumprod, const int64_t* timesteps, float snr_gamma, float* weights, int32_t B,
                     void* stream);
/* loss_out[0] += (1/B) sum_b w[b] * mean_chw( (pred - noise)^2 * mask )  and, when dpred != NULL,
   dpred = loss_scale * dloss/dpred (bf16, NHWC with row stride ld_dpred).  pred: NHWC bf16 [B*HW, ld_pred];
   noise: NCHW bf16; mask: NCHW fp32.  The caller zeroes loss_out. */
int b200_diffusion_loss(const void* pred, int64_t ld_pred, const void* noise, const float* mask, const float* weights,
                        float loss_scale, float* loss_out, void* dpred, int64_t ld_dpred, int32_t B, int32_t C,
                        int32_t HW, void* stream);
/* out[0] += sum |p| over n bf16 elements */
int b200_abs_sum(const void* p, int64_t n, float* out, void* stream);
/* Fused AdamW over ONE flat buffer: bf16 params / fp32 grads / bf16 moments.  Elements [0, n_first) are the LoRA
   factors (lr, wd, and the L1 penalty's sign-gradient l1_coeff*sign(p) folded into g); elements [n_first, n) are the
   trainable textual-inversion embedding rows (lr2, wd2).  Reproduces torch.optim.AdamW on bf16 tensors op for op
   (every intermediate rounded to bf16 as ATen does), so a step from equal state and equal bf16 gradients is
   bit-identical to the reference's optimizer.  g = bf16(bf16(grad*grad_scale) + l1_coeff*sign(p)).
   Hyper-parameters are doubles because the reference forms them as Python floats.  Zeroes `grad` when zero_grad. */
int b200_adamw(void* p, float* grad, void* m, void* v, int64_t n, int64_t n_first, double lr, double wd,
               double l1_coeff, double lr2, double wd2, double beta1, double beta2, double eps, int32_t step,
               double grad_scale, int32_t zero_grad, void* stream);

/* CUDA-graph form of the same update: the 12 hyper-parameter floats live in DEVICE memory (refreshed by a
   memcpy between replays), packed on the host by b200_adamw_pack_hyper from the same doubles. */
int b200_adamw_pack_hyper(double lr, double wd, double l1_coeff, double lr2, double wd2, double beta1, double beta2,
                          double eps, int32_t step, double grad_scale, float* out_host12);
int b200_adamw_dev(void* p, float* grad, void* m, void* v, int64_t n, int64_t n_first, const float* hyper_dev12,
                   int32_t zero_grad, void* stream);

/* Prodigy ([3P] prodigyopt==1.0 as trainer/optimizer.py:22-34 (UNet) and 134-144 (textual inversion) configure it:
   decouple=True, use_bias_correction=True, safeguard_warmup=True, betas=(0.9, 0.99)) over n bf16 parameters with the
   fp32 flat gradient buffer; state s / p0 / exp_avg / exp_avg_sq are bf16 like the parameters (p0 = the parameters at
   construction, the others zero).  scal8: 8 device doubles, initialised {d0, d0, 0, 0, 0, 0, 0, d0}: [0] d, [1] d_max,
   [2] d_numerator, [3..4] the step's global sums, [5] dlr, [6] skip flag, [7] d_hat.  hyper_dev12: device copy of what
   b200_prodigy_pack_hyper packs on the host (k = optimizer steps taken so far; lr = the value main.py:286-291 writes
   into param_groups[0]['lr'] for the UNet, 1.0 for textual inversion).  Three launches: state update + global sums,
   the scalar d update (one thread, double), the parameter update.  Gradients are zeroed when zero_grad != 0. */
int b200_prodigy_pack_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, double d_coef,
                            double growth_rate, double d0, int32_t k, int32_t use_bias_correction, double l1_coeff,
                            double grad_scale, float* out_host12);
int b200_prodigy_step(void* p, float* grad, void* s, const void* p0, void* exp_avg, void* exp_avg_sq, int64_t n,
                      double* scal8, const float* hyper_dev12, int32_t zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_LORA_H */
