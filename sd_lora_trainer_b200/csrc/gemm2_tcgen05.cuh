// CTA-pair (cta_group::2) bf16 GEMM for sm_100a: the big, plain projections of the step.
//
//   D[m, n] = alpha * sum_k A[m, k] * B[n, k]  (+ bias[n]) (+ R[m, n])   (+ the fused LoRA side path, see below)
//
// Why a second kernel: with 128 x BN single-CTA tiles the K loop of gemm_tcgen05_kernel is bound by L2 -> SM traffic
// (profiles/r01a_gemm_ncu_full.md: 663 MB for 2048x10240x1280 = the ~11 TB/s fabric limit), not by the tensor pipe.
// Here two CTAs of a cluster (one TPC) own a 256 x BN tile: each loads ITS 128 rows of A and HALF of the B tile, one
// tcgen05.mma.cta_group::2 (M = 256) issued by the leader CTA reads both halves, so a 256 x 256 tile streams
// 64 KiB per k-block for 2x the FLOPs of the 48 KiB a 128 x 256 tile needs.
//
//   warp 0      TMA producer (both CTAs; transaction bytes complete on the LEADER's full barrier)
//   warp 1      MMA issuer (leader CTA only) + TMEM alloc/dealloc (both CTAs, cta_group::2)
//   warps 2-9   epilogue (each CTA drains its own 128 accumulator rows; two warps per TMEM lane quarter)
//
// Hot loops are deliberately lean: every kernel parameter they need is copied to registers up front, descriptors are
// formed from hoisted constants, and nothing but barrier waits, TMA / MMA issues and ring bookkeeping sits in them
// (the first kernel spent ~170 SASS instructions per k-block in each of these two warps).
//
// Fused LoRA side path (same contract as gemm_tcgen05_kernel): while the K loop accumulates A.B^T, a second
// cta_group::2 MMA per k-step accumulates A.S^T (N = 16 / 32) next to the main tile; the epilogue warps of BOTH CTAs
// scale + round their 128 rows of it to bf16 into a swizzled smem tile (and to T_out), the leader issues one more MMA
// T.B2^T into the main accumulator.
#pragma once
#include "ptx.cuh"
#include "common.cuh"
#include "gemm_tcgen05.cuh"

namespace b200 {

constexpr int k2Threads = 320;
constexpr int k2ABytes = 128 * 64 * 2;           // one CTA's A tile per stage
constexpr int k2RingBytes = 160 * 1024;
constexpr int k2TOff = k2RingBytes;              // T tile (16 KiB, 128B-swizzled K-major A operand of the final MMA)
constexpr int k2BarOff = k2TOff + 16384;
constexpr int k2EpiOff = k2BarOff + 1024;        // 1024-aligned: the TMA-store staging boxes are 64B-swizzled
constexpr int k2EpiBytes = 8 * 32 * kEpiLd * 4;  // >= 8 warps x 2 x 2 KiB TMA-store boxes
constexpr int k2SmemBytes = k2EpiOff + k2EpiBytes;
static_assert(k2SmemBytes <= 232448, "shared memory budget");

struct Gemm2Args {
    // scalars first: they fit in a few constant-cache lines and are read once, before griddepcontrol.wait
    int M, N, BN, tiles_n, total_tiles, kblocks, ktail16;
    int b_mn, side, side_mn, b2_mn, side_r16, side_r;
    // GEGLU backward fused into the epilogue (the input-gradient GEMM of FeedForward.net.2, diffusers GEGLU): the tile is
    // dy[m, j] (j < N); with h = [value | gate] ([M, 2N], row stride h_ld) the epilogue writes dh[m, j] = dy * gelu(gate) and
    // dh[m, N + j] = dy * value * gelu'(gate) into D ([M, 2N]) instead of dy itself
    // GEGLU forward fused into the epilogue (FeedForward.net.0.proj with its rows interleaved in blocks of 128: every 256-wide
    // tile holds 128 value columns and the matching 128 gate columns): D <- the projection (kept for the backward),
    // mapD2 ([M, N/2]) <- value * gelu(gate).  BN = 256 only.
    int geglu_fwd;
    int geglu_bwd;
    long long h_ld;
    const __nv_bfloat16* H;
    int a_mn;                 // segment 0's A operand is MN-major ([K, M] with M contiguous: dY^T of a weight gradient): two 64 x 64 boxes per CTA
    int d_accum;              // fp32 output: D += result (each tile element is owned by one CTA: plain read-modify-write)
    int stage_bytes, num_stages, side_off, acc_stages;
    int vec_ok, bias_rows;
    // implicit 3x3 / pad 1 / stride 1 convolution on segment 0's A operand (NHWC; mapA dims C, W, H, N)
    int conv, conv_cblocks, conv_W, conv_H, b_tap_k, b_tap_n;
    // optional second K segment accumulated into the same tile: plain K-major A2 [M, K2], B2 via mapB2 / b2_mn
    int nseg, kblocks2, ktail16_2;
    // B / S / B2 (weights, LoRA factors) are not written by the preceding kernel of the stream: the producer fetches
    // the first ring-full of them BEFORE griddepcontrol.wait, overlapping the predecessor's tail
    int b_static;
    // bf16 row-major output through TMA: each epilogue warp stages [32 rows x 32 columns] boxes (64B swizzle) and one
    // lane issues cp.async.bulk.tensor stores (mapD); needs 16-byte aligned rows of D (and R); M / N edges are clipped
    int tma_store;
    // stream-K: when a problem has fewer 256 x BN tiles than CTA pairs but a long K, the (tile, k-block) space is cut
    // into 74 equal contiguous ranges; every range's partial tile is added to D with a TMA reduce-add (D pre-set by the
    // host to the residual or zero; bias rides with the range that holds k-block 0).  Needs tma_store.
    int streamk;
    float alpha, side_alpha;
    long long d_sm, r_sm, t_ld, bias_sb;
    void* D;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* R;
    __nv_bfloat16* T_out;
    long long* dbg;           // developer probe: per-CTA globaltimer stamps [cta][16] (nullptr in production)
    CUtensorMap mapA, mapB, mapS, mapB2, mapA2, mapD, mapD2;
};

// The (tile, k-block range) work items of one CTA pair; producer, MMA and epilogue warps walk the same sequence.
struct WorkIter2 {
    int streamk, kblocks, total_tiles, npairs, tile;
    long long u, u_end;
    __device__ __forceinline__ WorkIter2(int streamk_, int kblocks_, int total_tiles_, int npairs_, int pair)
        : streamk(streamk_), kblocks(kblocks_), total_tiles(total_tiles_), npairs(npairs_), tile(pair) {
        const long long U = static_cast<long long>(total_tiles_) * kblocks_;
        u = U * pair / npairs_;
        u_end = U * (pair + 1) / npairs_;
    }
    __device__ __forceinline__ bool next(int& t, int& kb0, int& kb1) {
        if (!streamk) {
            if (tile >= total_tiles) return false;
            t = tile;
            kb0 = 0;
            kb1 = kblocks;
            tile += npairs;
            return true;
        }
        if (u >= u_end) return false;
        t = static_cast<int>(u / kblocks);
        kb0 = static_cast<int>(u - static_cast<long long>(t) * kblocks);
        const long long len = min(static_cast<long long>(kblocks - kb0), u_end - u);
        kb1 = kb0 + static_cast<int>(len);
        u += len;
        return true;
    }
};

struct Epi2 {
    void* D;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* R;
    long long d_sm, r_sm, bias_sb;
    int M, N, bias_rows, vec_ok, accum;
    float alpha;
};

// One 32-row x 32-column chunk of the tile: registers (one row per thread) -> padded smem patch -> 8 lanes per
// 128-byte row segment -> alpha / bias / residual -> vector stores.  Row-major outputs only.
template <int kEpi>
__device__ __forceinline__ void epi2_chunk(const Epi2& e, const uint32_t (&raw)[32], float* stage, int lane, int m_warp,
                                           int n_chunk, int cvalid) {
    float4* srow = reinterpret_cast<float4*>(stage + lane * kEpiLd);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        srow[j] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]), __uint_as_float(raw[4 * j + 2]),
                              __uint_as_float(raw[4 * j + 3]));
    __syncwarp();
    const int sub = lane >> 3, col = (lane & 7) * 4;
    const int n = n_chunk + col;
    const int nv = min(4, min(e.N - n, cvalid - col));
    const float alpha = e.alpha;
    const int m_first = m_warp + sub;
    const float* sp = stage + sub * kEpiLd + col;
    const long long d_step = 4 * e.d_sm, r_step = 4 * e.r_sm;
    const long long doff = static_cast<long long>(m_first) * e.d_sm + n;
    const __nv_bfloat16* rp = e.R ? e.R + static_cast<long long>(m_first) * e.r_sm + n : nullptr;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    const bool bias_per_row = e.bias != nullptr && e.bias_rows != 0;
    if (e.bias != nullptr && !bias_per_row && nv > 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < nv) bv[q] = __bfloat162float(e.bias[n + q]);
    }
    const bool interior = e.vec_ok && !bias_per_row && (m_warp + 32 <= e.M) && (cvalid == 32) && (n_chunk + 32 <= e.N);
    if (interior) {
        float4 q[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) q[it] = *reinterpret_cast<const float4*>(sp + it * 4 * kEpiLd);
        uint2 rr[8];
        if (rp) {
#pragma unroll
            for (int it = 0; it < 8; ++it) rr[it] = *reinterpret_cast<const uint2*>(rp + it * r_step);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            float v0 = fmaf(q[it].x, alpha, bv[0]), v1 = fmaf(q[it].y, alpha, bv[1]);
            float v2 = fmaf(q[it].z, alpha, bv[2]), v3 = fmaf(q[it].w, alpha, bv[3]);
            if (rp) {
                v0 += __uint_as_float(rr[it].x << 16);
                v1 += __uint_as_float(rr[it].x & 0xffff0000u);
                v2 += __uint_as_float(rr[it].y << 16);
                v3 += __uint_as_float(rr[it].y & 0xffff0000u);
            }
            if (kEpi == 1) {
                float4* dp4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.D) + doff + it * d_step);
                if (e.accum) {
                    const float4 old = *dp4;
                    v0 += old.x, v1 += old.y, v2 += old.z, v3 += old.w;
                }
                *dp4 = make_float4(v0, v1, v2, v3);
            } else {
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1);
                const __nv_bfloat162 h1 = __floats2bfloat162_rn(v2, v3);
                uint2 w2;
                w2.x = *reinterpret_cast<const uint32_t*>(&h0);
                w2.y = *reinterpret_cast<const uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.D) + doff + it * d_step) = w2;
            }
        }
    } else if (nv > 0) {
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const int m = m_first + it * 4;
            if (m >= e.M) break;
            const float* q = sp + it * 4 * kEpiLd;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < nv) {
                    float x = q[c] * alpha + bv[c];
                    if (bias_per_row) x += __bfloat162float(e.bias[(m / e.bias_rows) * e.bias_sb + n + c]);
                    if (rp) x += __bfloat162float(rp[it * r_step + c]);
                    if (kEpi == 1) {
                        float* dp1 = reinterpret_cast<float*>(e.D) + doff + it * d_step + c;
                        *dp1 = e.accum ? *dp1 + x : x;
                    } else {
                        reinterpret_cast<__nv_bfloat16*>(e.D)[doff + it * d_step + c] = __float2bfloat16_rn(x);
                    }
                }
            }
        }
    }
    __syncwarp();
}

// One [32 x 32] chunk, bf16 row-major output, through a TMA store: thread <-> row does alpha / bias / residual on its 32
// accumulators with 16-byte loads, packs to bf16, writes its 64-byte row into the 64B-swizzled staging box (conflict
// free) and lane 0 issues the bulk tensor store.  ~90 instructions per chunk and no st.global at all.
__device__ __forceinline__ void epi2_chunk_tma(const Epi2& e, const CUtensorMap* mapD, const uint32_t (&raw)[32], uint8_t* box,
                                               int lane, int m_warp, int n_chunk, bool reduce, bool with_bias) {
    const int m = m_warp + lane;
    float v[32];
    const float alpha = e.alpha;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * alpha;
    if (e.bias != nullptr && with_bias) {
        const int mb = e.bias_rows ? min(m, e.M - 1) / e.bias_rows : 0;
        const uint4* bp = reinterpret_cast<const uint4*>(e.bias + mb * e.bias_sb + n_chunk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 w = __ldg(bp + j);
            const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[8 * j + 2 * q] += __uint_as_float(u[q] << 16);
                v[8 * j + 2 * q + 1] += __uint_as_float(u[q] & 0xffff0000u);
            }
        }
    }
    if (e.R != nullptr && with_bias && m < e.M) {          // (stream-K: with the range that holds k-block 0, like the bias)
        const uint4* rp = reinterpret_cast<const uint4*>(e.R + static_cast<long long>(m) * e.r_sm + n_chunk);
        uint4 w4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w4[j] = rp[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t u[4] = {w4[j].x, w4[j].y, w4[j].z, w4[j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[8 * j + 2 * q] += __uint_as_float(u[q] << 16);
                v[8 * j + 2 * q + 1] += __uint_as_float(u[q] & 0xffff0000u);
            }
        }
    }
    uint8_t* rowp = box + lane * 64;
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 w;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
        w.x = *reinterpret_cast<const uint32_t*>(&h0);
        w.y = *reinterpret_cast<const uint32_t*>(&h1);
        w.z = *reinterpret_cast<const uint32_t*>(&h2);
        w.w = *reinterpret_cast<const uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(rowp + ((j ^ sw) * 16)) = w;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (reduce) tma_reduce_add_4d(mapD, box, n_chunk, m_warp, 0, 0);
        else tma_store_4d(mapD, box, n_chunk, m_warp, 0, 0);
        bulk_commit();
    }
    __syncwarp();
}

// One [32 x 32] chunk, fp32 output, through TMA: two [32 rows x 16 columns] boxes (64-byte rows, the same 64B-swizzled
// staging layout as the bf16 boxes).  accumulate = cp.reduce.async.bulk.tensor .add (the fp32 sum happens at the L2: the
// SMs never read the 10 GB of dense weight gradients back), else a plain bulk store.  alpha (+ bias) only; M / N edges
// are clipped by the tensor map.
__device__ __forceinline__ void epi2_chunk_tma_f32(const Epi2& e, const CUtensorMap* mapD, const uint32_t (&raw)[32], uint8_t* box,
                                                   int lane, int m_warp, int n_chunk, bool accumulate, bool with_bias) {
    float v[32];
    const float alpha = e.alpha;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]) * alpha;
    if (e.bias != nullptr && with_bias) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (n_chunk + i < e.N) v[i] += __bfloat162float(e.bias[n_chunk + i]);
    }
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        uint8_t* rowp = box + hb * 2048 + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(rowp + ((j ^ sw) * 16)) =
                make_float4(v[hb * 16 + 4 * j], v[hb * 16 + 4 * j + 1], v[hb * 16 + 4 * j + 2], v[hb * 16 + 4 * j + 3]);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (accumulate) {
            tma_reduce_add_4d(mapD, box, n_chunk, m_warp, 0, 0);
            if (n_chunk + 16 < e.N) tma_reduce_add_4d(mapD, box + 2048, n_chunk + 16, m_warp, 0, 0);
        } else {
            tma_store_4d(mapD, box, n_chunk, m_warp, 0, 0);
            if (n_chunk + 16 < e.N) tma_store_4d(mapD, box + 2048, n_chunk + 16, m_warp, 0, 0);
        }
        bulk_commit();
    }
    __syncwarp();
}

// GEGLU forward in the epilogue: the value chunk (columns n_val .. +32 of the tile's first half) and its gate chunk (128
// columns further) -> both stored to the projection output through the warp's two staging boxes, then
// y = bf16(value) * bf16(gelu(bf16(gate))) staged into the first box again and stored through mapD2.
__device__ __forceinline__ void epi2_pair_tma_geglu_fwd(const Epi2& e, const CUtensorMap* mapD, const CUtensorMap* mapY,
                                                        const uint32_t (&rv)[32], const uint32_t (&rg)[32], uint8_t* box,
                                                        uint8_t* ybox, int lane, int m_warp, int n_val, int y_col) {
    const float alpha = e.alpha;
    uint8_t* row_a = box + lane * 64;
    uint8_t* row_g = box + 2048 + lane * 64;
    uint8_t* row_y = ybox + lane * 64;             // third staging box (the idle T tile of the side path): no mid-chunk wait
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a[8], gt[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = __uint_as_float(rv[8 * j + i]) * alpha;
            gt[i] = __uint_as_float(rg[8 * j + i]) * alpha;
        }
        if (e.bias != nullptr) {
            const uint4 wa = __ldg(reinterpret_cast<const uint4*>(e.bias + n_val + 8 * j));
            const uint4 wg = __ldg(reinterpret_cast<const uint4*>(e.bias + n_val + 128 + 8 * j));
            const uint32_t ua[4] = {wa.x, wa.y, wa.z, wa.w}, ug[4] = {wg.x, wg.y, wg.z, wg.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[2 * q] += __uint_as_float(ua[q] << 16);
                a[2 * q + 1] += __uint_as_float(ua[q] & 0xffff0000u);
                gt[2 * q] += __uint_as_float(ug[q] << 16);
                gt[2 * q + 1] += __uint_as_float(ug[q] & 0xffff0000u);
            }
        }
        uint4 oa, og, oy;
        uint32_t* pa = reinterpret_cast<uint32_t*>(&oa);
        uint32_t* pg = reinterpret_cast<uint32_t*>(&og);
        uint32_t* py = reinterpret_cast<uint32_t*>(&oy);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 ha = __floats2bfloat162_rn(a[2 * q], a[2 * q + 1]);
            const __nv_bfloat162 hg = __floats2bfloat162_rn(gt[2 * q], gt[2 * q + 1]);
            pa[q] = *reinterpret_cast<const uint32_t*>(&ha);
            pg[q] = *reinterpret_cast<const uint32_t*>(&hg);
            // y from the ROUNDED projection values, as the stand-alone kernel reads them back
            const float y0 = __bfloat162float(ha.x) * bfr(gelu_f(__bfloat162float(hg.x)));
            const float y1 = __bfloat162float(ha.y) * bfr(gelu_f(__bfloat162float(hg.y)));
            const __nv_bfloat162 hy = __floats2bfloat162_rn(y0, y1);
            py[q] = *reinterpret_cast<const uint32_t*>(&hy);
        }
        *reinterpret_cast<uint4*>(row_a + ((j ^ sw) * 16)) = oa;
        *reinterpret_cast<uint4*>(row_g + ((j ^ sw) * 16)) = og;
        *reinterpret_cast<uint4*>(row_y + ((j ^ sw) * 16)) = oy;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tma_store_4d(mapD, box, n_val, m_warp, 0, 0);
        tma_store_4d(mapD, box + 2048, n_val + 128, m_warp, 0, 0);
        tma_store_4d(mapY, ybox, y_col, m_warp, 0, 0);
        bulk_commit();
    }
    __syncwarp();
}

// GEGLU backward in the epilogue: one [32 x 32] chunk of dy (registers, rounded to bf16 like the stored tensor it replaces)
// + the value / gate chunks of h -> two [32 x 32] chunks of dh (value half at column n_chunk, gate half at N + n_chunk),
// each through one of this warp's two staging boxes.
__device__ __forceinline__ void epi2_chunk_tma_geglu_bwd(const Epi2& e, const CUtensorMap* mapD, const uint32_t (&raw)[32],
                                                         uint8_t* box, int lane, int m_warp, int n_chunk,
                                                         const __nv_bfloat16* H, long long h_ld) {
    const int m = m_warp + lane;
    const bool rok = m < e.M;
    const __nv_bfloat16* hv = H + static_cast<long long>(rok ? m : 0) * h_ld + n_chunk;
    const __nv_bfloat16* hg = hv + e.N;
    uint8_t* row_a = box + lane * 64;
    uint8_t* row_g = box + 2048 + lane * 64;
    const int sw = (lane >> 1) & 3;
    const float alpha = e.alpha;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 wa = make_uint4(0u, 0u, 0u, 0u), wg = wa;
        if (rok) {
            wa = *reinterpret_cast<const uint4*>(hv + 8 * j);
            wg = *reinterpret_cast<const uint4*>(hg + 8 * j);
        }
        const uint32_t ua[4] = {wa.x, wa.y, wa.z, wa.w}, ug[4] = {wg.x, wg.y, wg.z, wg.w};
        float da[8], dg[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const int i = 2 * q + s2;
                const float a = s2 ? __uint_as_float(ua[q] & 0xffff0000u) : __uint_as_float(ua[q] << 16);
                const float gt = s2 ? __uint_as_float(ug[q] & 0xffff0000u) : __uint_as_float(ug[q] << 16);
                const float d = bfr(__uint_as_float(raw[8 * j + i]) * alpha);
                float gl, dgl;
                gelu_both(gt, gl, dgl);
                da[i] = d * bfr(gl);
                dg[i] = bfr(d * a) * dgl;
            }
        }
        uint4 oa, og;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(da[0], da[1]), t1 = __floats2bfloat162_rn(da[2], da[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(da[4], da[5]), t3 = __floats2bfloat162_rn(da[6], da[7]);
        oa.x = *reinterpret_cast<const uint32_t*>(&t0), oa.y = *reinterpret_cast<const uint32_t*>(&t1);
        oa.z = *reinterpret_cast<const uint32_t*>(&t2), oa.w = *reinterpret_cast<const uint32_t*>(&t3);
        t0 = __floats2bfloat162_rn(dg[0], dg[1]), t1 = __floats2bfloat162_rn(dg[2], dg[3]);
        t2 = __floats2bfloat162_rn(dg[4], dg[5]), t3 = __floats2bfloat162_rn(dg[6], dg[7]);
        og.x = *reinterpret_cast<const uint32_t*>(&t0), og.y = *reinterpret_cast<const uint32_t*>(&t1);
        og.z = *reinterpret_cast<const uint32_t*>(&t2), og.w = *reinterpret_cast<const uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(row_a + ((j ^ sw) * 16)) = oa;
        *reinterpret_cast<uint4*>(row_g + ((j ^ sw) * 16)) = og;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tma_store_4d(mapD, box, n_chunk, m_warp, 0, 0);
        tma_store_4d(mapD, box + 2048, e.N + n_chunk, m_warp, 0, 0);
        bulk_commit();
    }
    __syncwarp();
}

// Ragged chunk (fewer than 32 valid columns at the N edge) when the TMA-store path owns the staging memory: plain
// thread <-> row scalar stores.  Rare (none of the step's shapes has a ragged edge), so simplicity wins.
__device__ __forceinline__ void epi2_chunk_direct(const Epi2& e, const uint32_t (&raw)[32], int lane, int m_warp, int n_chunk,
                                                  int ncols, bool with_bias) {
    const int m = m_warp + lane;
    if (m >= e.M) return;
    const __nv_bfloat16* bp = (e.bias && with_bias) ? e.bias + (e.bias_rows ? (m / e.bias_rows) * e.bias_sb : 0) : nullptr;
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(e.D) + static_cast<long long>(m) * e.d_sm;
    const __nv_bfloat16* rp = (e.R && with_bias) ? e.R + static_cast<long long>(m) * e.r_sm : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = n_chunk + j;
        if (j < ncols && n < e.N) {
            float x = __uint_as_float(raw[j]) * e.alpha;
            if (bp) x += __bfloat162float(bp[n]);
            if (rp) x += __bfloat162float(rp[n]);
            dp[n] = __float2bfloat16_rn(x);
        }
    }
}

template <int kEpi>
__global__ void __launch_bounds__(k2Threads, 1) gemm2_kernel(const __grid_constant__ Gemm2Args g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + k2BarOff);   // [kMaxStages]  (used in the leader CTA)
    uint64_t* empty_bar = full_bar + kMaxStages;                           // [kMaxStages]  (both CTAs, multicast commit)
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;                      // [2]           (both CTAs, multicast commit)
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;                          // [2]           (leader; 8 warps x 2 CTAs)
    uint64_t* side_full_bar = tmem_empty_bar + 2;                          // [2]           (both CTAs, multicast commit)
    uint64_t* t_ready_bar = side_full_bar + 2;                             //               (leader; 4 warps x 2 CTAs)
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(t_ready_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // ---- everything the loops need, read once (constant bank), before the dependency wait ----
    const int BN = g.BN, tiles_n = g.tiles_n, total_tiles = g.total_tiles, kblocks = g.kblocks;
    const int num_stages = g.num_stages, stage_bytes = g.stage_bytes, acc_stages = g.acc_stages;
    const int side = g.side, side_off = g.side_off, r16 = g.side_r16;
    const int b_mn = g.b_mn, side_mn = g.side_mn, b2_mn = g.b2_mn, a_mn = g.a_mn;
    const int conv = g.conv, nseg = g.nseg, kblocks2 = g.kblocks2, streamk = g.streamk;
    const int npairs = static_cast<int>(gridDim.x >> 1), pair = static_cast<int>(blockIdx.x >> 1);
    long long* const dbg = g.dbg;
    if (threadIdx.x == 0) dbg_stamp(dbg, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.mapA);
        tma_prefetch_desc(&g.mapB);
        if (side) {
            tma_prefetch_desc(&g.mapS);
            tma_prefetch_desc(&g.mapB2);
        }
        if (nseg == 2) {
            tma_prefetch_desc(&g.mapA2);
            tma_prefetch_desc(&g.mapB2);
        }
        for (int i = 0; i < num_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 16);
            mbar_init(&side_full_bar[i], 1);
        }
        mbar_init(t_ready_bar, 8);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc2(tmem_base_ptr, 512);
        tmem_relinquish2();
    }
    pdl_launch();
    tc_fence_before();
    cluster_sync_all();       // the peer's barriers are initialised and its TMEM allocated before anything targets them
    tc_fence_after();
    if (threadIdx.x != 0) pdl_wait();         // thread 0 (the TMA producer) waits after its weight prefetch
    const uint32_t tmem_base = *tmem_base_ptr;

    if (warp == 0) {
        // ===================== TMA producer (one thread per CTA) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full0 = mapa_u32(smem_u32(full_bar), 0);      // the LEADER's full barriers
            const int bn_half = BN >> 1;
            const int b_boxes = bn_half >> 6;
            const uint32_t b_bytes = b_mn ? static_cast<uint32_t>(b_boxes) * 8192u : static_cast<uint32_t>(bn_half) * 128u;
            const int sr_half = r16 >> 1;
            const uint32_t s_bytes = side ? (side_mn ? 8192u : static_cast<uint32_t>(sr_half) * 128u) : 0u;
            const uint32_t tx = 2u * (static_cast<uint32_t>(k2ABytes) + b_bytes + s_bytes);
            const uint32_t b2_bytes = b2_mn ? static_cast<uint32_t>(b_boxes) * 8192u : static_cast<uint32_t>(bn_half) * 128u;
            const int s_row = static_cast<int>(rank) * sr_half;
            const int conv_H = g.conv_H, conv_W = g.conv_W, conv_cblocks = g.conv_cblocks;
            const int conv_tap_k = g.b_tap_k, conv_tap_n = g.b_tap_n;
            // B-side loads of k-block kb of a tile (weights + LoRA side factor) into ring slot st
            auto issue_b = [&](int st, int kb, int nh0) {
                uint8_t* sb = smem + st * stage_bytes + k2ABytes;
                const uint32_t fb = full0 + static_cast<uint32_t>(st) * 8u;
                if (conv) {
                    const int tap = kb / conv_cblocks, cb = kb - tap * conv_cblocks;
                    tma2_load_4d(sb, &g.mapB, fb, tap * conv_tap_k + cb * kBK, nh0 + tap * conv_tap_n, 0, 0);
                } else if (!b_mn) {
                    tma2_load_4d(sb, &g.mapB, fb, kb * kBK, nh0, 0, 0);
                } else {
                    for (int jb = 0; jb < b_boxes; ++jb) tma2_load_4d(sb + jb * 8192, &g.mapB, fb, nh0 + jb * 64, kb * kBK, 0, 0);
                }
                if (side) {
                    if (side_mn) tma2_load_4d(sb + side_off, &g.mapS, fb, s_row, kb * kBK, 0, 0);
                    else tma2_load_4d(sb + side_off, &g.mapS, fb, kb * kBK, s_row, 0, 0);
                }
            };
            auto issue_a = [&](int st, int kb, int m0) {
                uint8_t* sa = smem + st * stage_bytes;
                const uint32_t fb = full0 + static_cast<uint32_t>(st) * 8u;
                if (conv) {
                    const int hw = conv_H * conv_W;
                    const int cn0 = m0 / hw, ch0 = (m0 - cn0 * hw) / conv_W;
                    const int tap = kb / conv_cblocks, cb = kb - tap * conv_cblocks;
                    const int kh = tap / 3, kw = tap - kh * 3;
                    tma2_load_4d(sa, &g.mapA, fb, cb * kBK, kw - 1, ch0 + kh - 1, cn0);
                } else if (a_mn) {
                    tma2_load_4d(sa, &g.mapA, fb, m0, kb * kBK, 0, 0);
                    tma2_load_4d(sa + 8192, &g.mapA, fb, m0 + 64, kb * kBK, 0, 0);
                } else {
                    tma2_load_4d(sa, &g.mapA, fb, kb * kBK, m0, 0, 0);
                }
            };
            // ---- weight prefetch: the first ring-full of B tiles does not depend on the previous kernel ----
            int npre = 0;
            if (g.b_static && !streamk) {
                const int n_blk0 = pair % tiles_n;
                const int nh00 = n_blk0 * BN + static_cast<int>(rank) * bn_half;
                npre = kblocks < num_stages ? kblocks : num_stages;
                for (int st = 0; st < npre; ++st) {
                    if (rank == 0) mbar_expect_tx(&full_bar[st], tx);
                    issue_b(st, st, nh00);
                }
            }
            pdl_wait();
            dbg_stamp(dbg, 1);
            WorkIter2 wit(streamk, kblocks, total_tiles, npairs, pair);
            int tile, kb0, kb1;
            bool first_item = true;
            while (wit.next(tile, kb0, kb1)) {
                const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
                const int m0 = m_blk * 256 + static_cast<int>(rank) * 128;
                const int nh0 = n_blk * BN + static_cast<int>(rank) * bn_half;
                int kb_first = kb0;
                if (first_item && npre > 0) {
                    for (int st = 0; st < npre; ++st) issue_a(st, st, m0);
                    kb_first = npre;
                    stage = (npre == num_stages) ? 0 : npre;
                    phase = (npre == num_stages) ? 1u : 0u;
                }
                if (conv) {
                    // implicit im2col: one box per (tap, 64-channel block) with shifted coordinates; TMA zero-fills the halo
                    const int hw = conv_H * conv_W;
                    const int cn0 = m0 / hw, ch0 = (m0 - cn0 * hw) / conv_W;
                    int tap0 = kb_first / conv_cblocks;
                    int cb = kb_first - tap0 * conv_cblocks, kh = tap0 / 3;
                    int kw = tap0 - kh * 3, bk_tap = tap0 * conv_tap_k, bn_tap = tap0 * conv_tap_n;
                    for (int kb = kb_first; kb < kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
                        uint8_t* sa = smem + stage * stage_bytes;
                        const uint32_t fb = full0 + static_cast<uint32_t>(stage) * 8u;
                        tma2_load_4d(sa, &g.mapA, fb, cb * kBK, kw - 1, ch0 + kh - 1, cn0);
                        tma2_load_4d(sa + k2ABytes, &g.mapB, fb, bk_tap + cb * kBK, nh0 + bn_tap, 0, 0);
                        if (++cb == conv_cblocks) {
                            cb = 0;
                            bk_tap += conv_tap_k;
                            bn_tap += conv_tap_n;
                            if (++kw == 3) {
                                kw = 0;
                                ++kh;
                            }
                        }
                        advance_stage(stage, phase, num_stages);
                    }
                } else {
                    int k = kb_first * kBK;
                    for (int kb = kb_first; kb < kb1; ++kb, k += kBK) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
                        uint8_t* sa = smem + stage * stage_bytes;
                        uint8_t* sb = sa + k2ABytes;
                        const uint32_t fb = full0 + static_cast<uint32_t>(stage) * 8u;
                        if (a_mn) {
                            tma2_load_4d(sa, &g.mapA, fb, m0, k, 0, 0);
                            tma2_load_4d(sa + 8192, &g.mapA, fb, m0 + 64, k, 0, 0);
                        } else {
                            tma2_load_4d(sa, &g.mapA, fb, k, m0, 0, 0);
                        }
                        if (!b_mn) {
                            tma2_load_4d(sb, &g.mapB, fb, k, nh0, 0, 0);
                        } else {
                            for (int jb = 0; jb < b_boxes; ++jb) tma2_load_4d(sb + jb * 8192, &g.mapB, fb, nh0 + jb * 64, k, 0, 0);
                        }
                        if (side) {
                            if (side_mn) tma2_load_4d(sb + side_off, &g.mapS, fb, s_row, k, 0, 0);
                            else tma2_load_4d(sb + side_off, &g.mapS, fb, k, s_row, 0, 0);
                        }
                        advance_stage(stage, phase, num_stages);
                    }
                }
                if (nseg == 2 && kb1 == kblocks) {
                    // second K segment (conv-LoRA side products): plain K-major A2, B2 K- or MN-major
                    int k = 0;
                    for (int kb = 0; kb < kblocks2; ++kb, k += kBK) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (static_cast<uint32_t>(k2ABytes) + b2_bytes));
                        uint8_t* sa = smem + stage * stage_bytes;
                        uint8_t* sb = sa + k2ABytes;
                        const uint32_t fb = full0 + static_cast<uint32_t>(stage) * 8u;
                        tma2_load_4d(sa, &g.mapA2, fb, k, m0, 0, 0);
                        if (!b2_mn) {
                            tma2_load_4d(sb, &g.mapB2, fb, k, nh0, 0, 0);
                        } else {
                            for (int jb = 0; jb < b_boxes; ++jb) tma2_load_4d(sb + jb * 8192, &g.mapB2, fb, nh0 + jb * 64, k, 0, 0);
                        }
                        advance_stage(stage, phase, num_stages);
                    }
                }
                if (first_item) dbg_stamp(dbg, 2);           // every load of the first work item issued
                first_item = false;
                if (side) {
                    // one more ring slot per tile: this CTA's half of the B2 tile for the final rank-r MMA
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * b2_bytes);
                    uint8_t* sb = smem + stage * stage_bytes + k2ABytes;
                    const uint32_t fb = full0 + static_cast<uint32_t>(stage) * 8u;
                    if (!b2_mn) {
                        tma2_load_4d(sb, &g.mapB2, fb, 0, nh0, 0, 0);
                    } else {
                        for (int jb = 0; jb < b_boxes; ++jb) tma2_load_4d(sb + jb * 8192, &g.mapB2, fb, nh0 + jb * 64, 0, 0, 0);
                    }
                    advance_stage(stage, phase, num_stages);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA; warp-uniform loop, one elected lane issues) =====================
        if (rank == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, t_phase = 0;
            const uint32_t smem_base = smem_u32(smem);
            const bool leader = elect_one();
            const uint32_t idesc = umma2_idesc_bf16(BN, a_mn, b_mn);
            const uint32_t idesc_s = umma2_idesc_bf16(r16, 0, side_mn);
            const uint32_t idesc_2 = umma2_idesc_bf16(BN, 0, b2_mn);
            const uint64_t a_hi = umma_desc(0, 16, 1024);                 // K-major A (and the T tile of the side path)
            const uint64_t a0_hi = umma_desc(0, a_mn ? 8192u : 16u, 1024);   // segment 0's A: MN-major for weight gradients
            const uint64_t a_step = a_mn ? 128u : 2u;
            const uint64_t b_hi = umma_desc(0, b_mn ? 8192u : 16u, 1024);
            const uint64_t s_hi = umma_desc(0, side_mn ? 8192u : 16u, 1024);
            const uint64_t b2_hi = umma_desc(0, b2_mn ? 8192u : 16u, 1024);
            const uint64_t b_step = b_mn ? 128u : 2u, s_step = side_mn ? 128u : 2u, b2_step = b2_mn ? 128u : 2u;
            const int ktail = g.ktail16, ktail2 = g.ktail16_2;
            WorkIter2 wit(streamk, kblocks, total_tiles, npairs, pair);
            int tile, kb0, kb1;
            bool first_item = true;
            while (wit.next(tile, kb0, kb1)) {
                mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * 256);
                const uint32_t tmem_s = tmem_base + static_cast<uint32_t>(acc_stages == 1 ? 256 : acc * 256 + kSideCol);
                uint32_t accum = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (leader) {
                        if (dbg != nullptr && first_item && (kb == kb0 || kb == kb1 - 1)) dbg_stamp(dbg, kb == kb0 ? 3 : 4);
                        const uint32_t sa = smem_base + static_cast<uint32_t>(stage * stage_bytes);
                        const uint64_t ad = a0_hi | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
                        const uint64_t bd = b_hi | static_cast<uint64_t>(((sa + k2ABytes) & 0x3FFFF) >> 4);
                        const int n16 = (kb == kblocks - 1) ? ktail : 4;
                        if (side) {
                            const uint64_t sd = s_hi | static_cast<uint64_t>(((sa + k2ABytes + side_off) & 0x3FFFF) >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < n16) umma2_bf16(tmem_s, ad + a_step * k, sd + s_step * k, idesc_s, accum | (k > 0));
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < n16) umma2_bf16(tmem_d, ad + a_step * k, bd + b_step * k, idesc, accum | (k > 0));
                        umma2_commit_mc(&empty_bar[stage]);      // frees this ring slot in BOTH CTAs
                    }
                    __syncwarp();
                    accum = 1;
                    advance_stage(stage, phase, num_stages);
                }
                if (nseg == 2 && kb1 == kblocks) {
                    for (int kb = 0; kb < kblocks2; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t sa = smem_base + static_cast<uint32_t>(stage * stage_bytes);
                            const uint64_t ad = a_hi | static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
                            const uint64_t bd = b2_hi | static_cast<uint64_t>(((sa + k2ABytes) & 0x3FFFF) >> 4);
                            const int n16 = (kb == kblocks2 - 1) ? ktail2 : 4;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < n16) umma2_bf16(tmem_d, ad + 2u * k, bd + b2_step * k, idesc_2, 1);
                            umma2_commit_mc(&empty_bar[stage]);
                        }
                        __syncwarp();
                        advance_stage(stage, phase, num_stages);
                    }
                }
                if (side) {
                    if (leader) umma2_commit_mc(&side_full_bar[acc]);     // rank-r accumulator complete -> T-phase
                    __syncwarp();
                    mbar_wait(&full_bar[stage], phase);                   // both halves of the B2 tile landed
                    mbar_wait_cluster(t_ready_bar, t_phase);              // T staged in smem by both CTAs
                    t_phase ^= 1;
                    tc_fence_after();
                    if (leader && first_item) dbg_stamp(dbg, 9);
                    if (leader) {
                        const uint32_t sb2 = smem_base + static_cast<uint32_t>(stage * stage_bytes) + k2ABytes;
                        const uint64_t td = a_hi | static_cast<uint64_t>(((smem_base + k2TOff) & 0x3FFFF) >> 4);
                        const uint64_t b2d = b2_hi | static_cast<uint64_t>((sb2 & 0x3FFFF) >> 4);
                        for (int k = 0; k < (r16 >> 4); ++k) umma2_bf16(tmem_d, td + 2u * k, b2d + b2_step * k, idesc_2, 1);
                        umma2_commit_mc(&empty_bar[stage]);
                    }
                    __syncwarp();
                    advance_stage(stage, phase, num_stages);
                }
                if (leader) umma2_commit_mc(&tmem_full_bar[acc]);         // accumulator complete -> both epilogues
                __syncwarp();
                first_item = false;
                if (++acc == acc_stages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: each CTA drains its own 128 rows; 2 warps per TMEM lane quarter ============
        const int ew = warp - 2;
        const int half = ew >> 2;                     // which 32-column chunks this warp takes (even / odd)
        const int lane_base = (warp & 3) * 32;
        float* stage_buf = reinterpret_cast<float*>(smem + k2EpiOff) + ew * (32 * kEpiLd);
        // TMA-store staging: this warp's two 2 KiB boxes.  They alias the padded patches of the staged path, so a launch
        // uses one or the other (tma_store is per launch); ragged chunks of a TMA launch take epi2_chunk_direct.
        uint8_t* tma_box = smem + k2EpiOff + ew * 4096;
        const int tma_store = g.tma_store;
        int tma_buf = 0;
        const uint32_t tmem_empty_leader = mapa_u32(smem_u32(tmem_empty_bar), 0);
        const uint32_t t_ready_leader = mapa_u32(smem_u32(t_ready_bar), 0);
        Epi2 e;
        e.D = g.D;
        e.bias = g.bias;
        e.R = g.R;
        e.d_sm = g.d_sm;
        e.r_sm = g.r_sm;
        e.bias_sb = g.bias_sb;
        e.M = g.M;
        e.N = g.N;
        e.bias_rows = g.bias_rows;
        e.vec_ok = g.vec_ok;
        e.accum = g.d_accum;
        e.alpha = g.alpha;
        const float side_alpha = g.side_alpha;
        const int side_r = g.side_r;
        __nv_bfloat16* const T_out = g.T_out;
        const long long t_ld = g.t_ld;
        int acc = 0;
        uint32_t acc_phase = 0;
        WorkIter2 wit(streamk, kblocks, total_tiles, npairs, pair);
        int tile, kb0, kb1;
        bool first_item = true;
        while (wit.next(tile, kb0, kb1)) {
            const bool with_bias = !streamk || kb0 == 0;      // stream-K: the bias rides with the range holding k-block 0
            const int m_blk = tile / tiles_n, n_blk = tile - m_blk * tiles_n;
            const int m_warp = m_blk * 256 + static_cast<int>(rank) * 128 + lane_base;
            const int n0 = n_blk * BN;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + static_cast<uint32_t>(acc * 256);
            if (side && half == 0) {
                // ---- T-phase: this CTA's 128 rows of Tacc -> alpha, bf16 -> swizzled smem A operand (+ T_out) ----
                mbar_wait(&side_full_bar[acc], acc_phase);
                tc_fence_after();
                if (warp == 2 && lane == 0 && first_item) dbg_stamp(dbg, 8);
                const int row = lane_base + lane;
                uint8_t* trow = smem + k2TOff + row * 128;
                const int m = m_warp + lane;
                __nv_bfloat16* tp = (n_blk == 0 && T_out != nullptr && m < e.M) ? T_out + static_cast<long long>(m) * t_ld : nullptr;
                const bool t_vec = (side_r & 7) == 0 && (t_ld & 7) == 0;
                // up to 64 side columns (rank 48 = the fused q|k|v projection), 32 at a time.  Columns 32.. (wide ranks only)
                // go first, T_out copy included; columns 0..31 follow, and their T_out copy (the common rank-16 / 32 case)
                // comes AFTER the hand-off to the MMA warp, off the critical path - one 16-register chunk stays live.
                const uint32_t side_taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) +
                                            static_cast<uint32_t>(acc_stages == 1 ? 256 : acc * 256 + kSideCol);
                auto t_chunk = [&](int cc, uint32_t (&packed)[16]) {
                    uint32_t raw[32];
                    tmem_ld32(side_taddr + cc * 32, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(raw[2 * j]) * side_alpha,
                                                                        __uint_as_float(raw[2 * j + 1]) * side_alpha);
                        packed[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (cc * 32 + j * 8 < r16)
                            *reinterpret_cast<uint4*>(trow + (((cc * 4 + j) ^ (row & 7)) * 16)) =
                                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    }
                };
                auto t_out_chunk = [&](int cc, const uint32_t (&packed)[16]) {
                    if (tp == nullptr) return;
                    if (t_vec) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (cc * 32 + j * 8 < side_r)
                                *reinterpret_cast<uint4*>(tp + cc * 32 + j * 8) =
                                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    } else {
                        const __nv_bfloat16* pv = reinterpret_cast<const __nv_bfloat16*>(packed);
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (cc * 32 + j < side_r) tp[cc * 32 + j] = pv[j];
                    }
                };
                if (r16 > 32) {
                    uint32_t hi[16];
                    t_chunk(1, hi);
                    t_out_chunk(1, hi);
                }
                uint32_t lo16[16];
                t_chunk(0, lo16);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(t_ready_leader);      // release.cluster: the leader's MMA reads this smem
                t_out_chunk(0, lo16);
            }
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (warp == 2 && lane == 0 && first_item) dbg_stamp(dbg, 5);
            // chunks half, half+2, ... of 32 columns; the tcgen05.ld of the next chunk is in flight while this one is
            // converted / staged / stored (two register buffers)
            auto drain = [&](const uint32_t (&raw)[32], int c0) {
                if (n0 + c0 >= e.N) return;            // warp-uniform
                if (kEpi == 0 && tma_store && g.geglu_bwd) {
                    if (m_warp >= e.M) return;         // (the host admits the mode only for N % 32 == 0: no ragged chunk)
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    epi2_chunk_tma_geglu_bwd(e, &g.mapD, raw, tma_box, lane, m_warp, n0 + c0, g.H, g.h_ld);
                } else if (kEpi == 0 && tma_store) {
                    if (m_warp >= e.M) return;         // rows past the M edge (warp-uniform)
                    if (n0 + c0 + 32 <= e.N) {
                        // double-buffered staging box: the store issued two chunks ago must have read its smem
                        if (lane == 0) bulk_wait_read<1>();
                        __syncwarp();
                        epi2_chunk_tma(e, &g.mapD, raw, tma_box + (tma_buf & 1) * 2048, lane, m_warp, n0 + c0, streamk != 0,
                                       with_bias);
                        ++tma_buf;
                    } else {
                        // (the host only selects stream-K for problems without a ragged N edge)
                        epi2_chunk_direct(e, raw, lane, m_warp, n0 + c0, min(32, BN - c0), with_bias);
                    }
                } else if (kEpi == 1 && tma_store) {
                    if (m_warp >= e.M) return;         // rows past the M edge (warp-uniform)
                    // fp32 through TMA (dense weight gradients: reduce-add when accumulating): both staging boxes are one chunk
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    epi2_chunk_tma_f32(e, &g.mapD, raw, tma_box, lane, m_warp, n0 + c0, e.accum != 0, with_bias);
                } else {
                    epi2_chunk<kEpi>(e, raw, stage_buf, lane, m_warp, n0 + c0, min(32, BN - c0));
                }
            };
            if (kEpi == 0 && g.geglu_fwd) {
                // BN = 256: chunks 0..3 are value columns, 4..7 their gates; this warp pairs (half, half + 4), (half + 2, half + 6)
                if (m_warp < e.M) {
#pragma unroll 1
                    for (int p = half; p < 4; p += 2) {
                        uint32_t ra[32], rb[32];
                        tmem_ld32(taddr + p * 32, ra);
                        tmem_ld32(taddr + 128 + p * 32, rb);
                        tmem_ld_wait();
                        if (lane == 0) bulk_wait_read<0>();
                        __syncwarp();
                        epi2_pair_tma_geglu_fwd(e, &g.mapD, &g.mapD2, ra, rb, tma_box, smem + k2TOff + ew * 2048, lane, m_warp,
                                                n0 + p * 32, (n0 >> 1) + p * 32);
                    }
                }
            } else {
                uint32_t ra[32], rb[32];
                int c0 = half * 32;
                if (c0 < BN) tmem_ld32(taddr + c0, ra);
                for (; c0 < BN; c0 += 128) {
                    tmem_ld_wait();
                    const int c1 = c0 + 64;
                    if (c1 < BN) tmem_ld32(taddr + c1, rb);
                    drain(ra, c0);
                    if (c1 < BN) {
                        tmem_ld_wait();
                        if (c1 + 64 < BN) tmem_ld32(taddr + c1 + 64, ra);
                        drain(rb, c1);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (warp == 2 && lane == 0 && first_item) dbg_stamp(dbg, 6);
            first_item = false;
            if (lane == 0) mbar_arrive_cluster_nofence(tmem_empty_leader + static_cast<uint32_t>(acc) * 8u);
            if (++acc == acc_stages) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    if (warp >= 2 && lane == 0) bulk_wait<0>();       // this warp's TMA stores are complete before the grid is
    tc_fence_before();
    if (threadIdx.x == 0) dbg_stamp(dbg, 7);
    cluster_sync_all();       // no CTA leaves (or frees TMEM) while its peer may still read its smem / signal its barriers
    if (warp == 1) tmem_dealloc2(tmem_base, 512);
}

}  // namespace b200
