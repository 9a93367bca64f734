// Fused attention for head_dim 64 on sm_100a (the SDXL case: every head is 64 wide).
//
// Forward  (one CTA = 128 queries of one (batch, head); two CTAs per SM):
//     TMA: Q once, K/V in 128-key blocks through a 2-stage ring
//     tcgen05: S = Q.K^T into TMEM (128 lanes x 128 fp32 columns), O += P.V into TMEM (128 x 64)
//     softmax warps: one thread per query row reads its S row from TMEM (no shuffles), online max / exp2 / sum,
//     rescales O in TMEM, writes P as bf16 into a 128B-swizzled smem tile that is the next MMA's A operand.
//   Emits O (bf16, [B*L, C] with the head at column h*64) and LSE (fp32 [B, H, L]) for the backward.
//
// Backward (one CTA = 128 keys of one (batch, head); loops over 128-query blocks):
//     S = Q.K^T and dP = dO.V^T into TMEM; softmax warps rebuild P = exp(S*scale - LSE), dS = P*(dP - delta)*scale and
//     stage both as bf16 smem tiles; dV += P^T.dO, dK += dS^T.Q (the SAME smem tiles read MN-major) and
//     dQ_blk = dS.K, which leaves through fp32 red.global.add (it is summed over key blocks).
// No [L, L] tensor ever reaches HBM.
#pragma once
#include "ptx.cuh"
#include "common.cuh"

namespace b200 {

// Debug timelines (clock64 stamps of CTA 0, printed by the host wrappers when B200_FLASH_TIMELINE=1): compiled out unless this
// is set to true - profiles/r02m_flash_timelines.txt holds the ones the comments below quote.
#ifndef B200_FLASH_TIMELINE_BUILD
#define B200_FLASH_TIMELINE_BUILD 0
#endif
constexpr bool kFaTimeline = B200_FLASH_TIMELINE_BUILD != 0;

constexpr int kFaThreads = 320;                // TMA, MMA, 2 x 4 softmax warps (keys 0-63 / 64-127 of every block)
constexpr int kFaTile = 128 * 64 * 2;          // one [128 x 64] bf16 tile: 16 KiB
// forward smem map
constexpr int kFwdQ = 0;
constexpr int kFwdK = kFwdQ + kFaTile;         // 2 stages
constexpr int kFwdV = kFwdK + 2 * kFaTile;     // 2 stages
constexpr int kFwdP = kFwdV + 2 * kFaTile;     // [128 q x 128 k] bf16 = 2 column blocks of 16 KiB
constexpr int kFwdBar = kFwdP + 2 * kFaTile;
constexpr int kFwdSmem = kFwdBar + 128;

struct FlashFwdArgs {
    CUtensorMap mapQ, mapK, mapV;              // dims (64, rows, H, B); box (64, 128) K-major / (64, 64) for V
    __nv_bfloat16* O;
    float* LSE;                                // [B, H, L], natural-log domain
    int L, Lk, H;
    long long o_ld;                            // row stride of O in elements (= C)
    float scale;                               // 1/sqrt(d)
    long long* timeline;                       // debug (B200_FLASH_TIMELINE=1): clock64 stamps of CTA (0,0,0), [block][8]; else NULL
};

// 2^x on the MUFU pipe, denormal results flushed to zero (softmax weights below 2^-126 are zero for every purpose
// here); exp2f() wraps the same instruction in a range fix-up that costs three more instructions per element.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x on the FMA / ALU pipes (no MUFU): floor by the round-down magic-number add, degree-3 minimax polynomial of 2^f on
// [0, 1) (max relative error 8.8e-5 - P is rounded to bf16, 3.9e-3, right after), exponent spliced in with a shift-add.
// x <= ~100; results below 2^-126 come out as 2^-126 (negligible in a row sum whose largest term is >= 2^-8).
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -126.f);
    float t;
    asm("add.rm.ftz.f32 %0, %1, 0f4B400000;" : "=f"(t) : "f"(x));       // 1.5 * 2^23 + floor(x)
    const float f = x - (t - 12582912.f);
    const float pf = fmaf(fmaf(fmaf(0.077119089663028717f, f, 0.227564394474029541f), f, 0.695146143436431885f), f, 1.f);
    return __int_as_float(__float_as_int(pf) + (__float_as_int(t) << 23));
}

// two at a time, on FADD2 / FFMA2
__device__ __forceinline__ float2 poly_exp2x2(float2 x) {
    x = make_float2(fmaxf(x.x, -126.f), fmaxf(x.y, -126.f));
    float2 t;
    const float2 magic = make_float2(12582912.f, 12582912.f);
    asm("add.rm.ftz.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&t))
        : "l"(*reinterpret_cast<const unsigned long long*>(&x)), "l"(*reinterpret_cast<const unsigned long long*>(&magic)));
    const float2 fl = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __ffma2_rn(fl, make_float2(-1.f, -1.f), x);
    float2 pf = __ffma2_rn(make_float2(0.077119089663028717f, 0.077119089663028717f), f,
                           make_float2(0.227564394474029541f, 0.227564394474029541f));
    pf = __ffma2_rn(pf, f, make_float2(0.695146143436431885f, 0.695146143436431885f));
    pf = __ffma2_rn(pf, f, make_float2(1.f, 1.f));
    return make_float2(__int_as_float(__float_as_int(pf.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(pf.y) + (__float_as_int(t.y) << 23)));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Store 32 consecutive columns [c0, c0+32) of row `row` of a [128 x 128] bf16 tile kept as two 128B-swizzled
// K-major column blocks (the layout tcgen05.mma reads as an A operand, and - transposed - as an MN-major one).
__device__ __forceinline__ void store_p_chunk(uint8_t* tile, int row, int c0, const float (&v)[32]) {
    uint8_t* base = tile + (c0 >> 6) * kFaTile + row * 128;
    const int j0 = (c0 & 63) >> 3;             // first 16-byte chunk inside the 128-byte row
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int phys = (j0 + j) ^ (row & 7);
        uint4 q;
        q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
        q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(base + phys * 16) = q;
    }
}

// same, for 32 values already packed to bf16 pairs
__device__ __forceinline__ void store_packed_chunk(uint8_t* tile, int row, int c0, const uint32_t (&v)[16]) {
    uint8_t* base = tile + (c0 >> 6) * kFaTile + row * 128;
    const int j0 = (c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int phys = (j0 + j) ^ (row & 7);
        *reinterpret_cast<uint4*>(base + phys * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}

// same, 16 consecutive columns [c0, c0+16) (c0 a multiple of 16) as 8 packed pairs
__device__ __forceinline__ void store_packed16(uint8_t* tile, int row, int c0, const uint32_t (&v)[8]) {
    uint8_t* base = tile + (c0 >> 6) * kFaTile + row * 128;
    const int j0 = (c0 & 63) >> 3;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int phys = (j0 + j) ^ (row & 7);
        *reinterpret_cast<uint4*>(base + phys * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}

__global__ void __launch_bounds__(kFaThreads, 2) flash_fwd_kernel(const __grid_constant__ FlashFwdArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFwdBar);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;      // [2]
    uint64_t* v_full = bars + 3;      // [2]
    uint64_t* kv_empty = bars + 5;    // [2]
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;
    uint64_t* p_full = bars + 9;
    uint64_t* o_done = bars + 10;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (g.Lk + 127) >> 7;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.mapQ);
        tma_prefetch_desc(&g.mapK);
        tma_prefetch_desc(&g.mapV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 8);
        mbar_init(p_full, 8);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                   // nothing above touched global memory
    const uint32_t tmem = *tmem_ptr;
    // TMEM: S [128 x 128] at columns 0-127; TWO output accumulators O_a (keys 0-63 of every block) at 128-191 and
    // O_b (keys 64-127) at 192-255.  Each softmax warp group runs its own online softmax over its half of the keys
    // (own running reference and row sum, no per-block exchange); the halves are merged once, in the epilogue.
    const uint32_t tmem_S = tmem, tmem_O = tmem + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, kFaTile);
            tma_load_4d(smem + kFwdQ, &g.mapQ, q_full, 0, q0, h, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], kFaTile);
                tma_load_4d(smem + kFwdK + s * kFaTile, &g.mapK, &k_full[s], 0, j * 128, h, b);
                mbar_expect_tx(&v_full[s], kFaTile);
                tma_load_4d(smem + kFwdV + s * kFaTile, &g.mapV, &v_full[s], 0, j * 128, h, b);
                tma_load_4d(smem + kFwdV + s * kFaTile + 8192, &g.mapV, &v_full[s], 0, j * 128 + 64, h, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, 0, 0);     // S[128q x 128k] = Q (K-major) . K^T (K-major)
            const uint32_t idesc_o = umma_idesc_bf16(64, 0, 1);      // O[128q x 64d] += P (K-major) . V (MN-major)
            const uint32_t sq = smem_u32(smem + kFwdQ), sp = smem_u32(smem + kFwdP);
            auto issue_s = [&](int j) {
                const uint32_t sk = smem_u32(smem + kFwdK + (j & 1) * kFaTile);
                const uint64_t ad = umma_desc(sq, 16, 1024), bd = umma_desc(sk, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                if (j + 1 < nkv) {
                    mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                    mbar_wait(s_free, j & 1);              // the softmax warps have drained S_j out of TMEM
                    tc_fence_after();
                    issue_s(j + 1);
                }
                mbar_wait(p_full, j & 1);                   // P_j staged in smem, O rescaled
                mbar_wait(&v_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sv = smem_u32(smem + kFwdV + s * kFaTile);
                const uint64_t vd = umma_desc(sv, 8192, 1024);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t pd = umma_desc(sp + (k >> 2) * kFaTile, 16, 1024) + 2 * (k & 3);
                    umma_bf16(tmem_O + (k >> 2) * 64, pd, vd + 128 * k, idesc_o, (j > 0 || (k & 3) > 0) ? 1u : 0u);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(o_done);
            }
        }
        __syncwarp();
    } else {
        // ---- softmax / correction / epilogue: thread <-> (query row, half of the keys of every block) ----
        // One warp per scheduler could not keep up with the tensor pipe (issue-bound); two warp groups halve the work
        // per thread, and because each owns its own accumulator they never have to agree on a running maximum.
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const int half = (warp >= 6) ? 1 : 0;
        const uint32_t lane_off = static_cast<uint32_t>(lane_base) << 16;
        const uint32_t tmem_Oh = tmem_O + half * 64;
        const float sl2 = g.scale * 1.4426950408889634f;           // scores are used in the log2 domain
        // m: the reference maximum the exponentials are taken against.  It only moves when a row's running maximum
        // outgrows it by more than 2^8 (log2 domain), so O (in TMEM) is rescaled on a few blocks instead of on every one;
        // the final O / l and LSE = m.scale + ln(l) are exact whatever reference was used.
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kvalid = min(128, g.Lk - j * 128) - half * 64;   // valid keys among this group's 64 (may be <= 0)
            const bool full_blk = kvalid >= 64;                     // warp-uniform
            // pass 1: row maximum over this group's keys
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_S + lane_off + half * 64 + c * 32, raw);
                tmem_ld_wait();
                if (full_blk) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(raw[i]));
                }
            }
            const float m_cand = fmaxf(m, mx);
            const bool need = (m_cand - m) * sl2 > 8.f;             // true on the first block with a valid key (m = -inf)
            const bool any_need = __any_sync(0xffffffffu, need);
            if (j > 0) {
                // the previous P.V must have retired before O is rescaled and before sP is overwritten
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
            }
            if (any_need) {
                const float alpha = need ? fast_exp2((m - m_cand) * sl2) : 1.f;   // 0 when m was -inf
                if (need) m = m_cand;
                l *= alpha;
                if (j > 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tmem_Oh + lane_off + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tmem_Oh + lane_off + c * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            const float moff = m * sl2;
            // pass 2: exponentials, row sum, P -> smem (bf16)
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_S + lane_off + half * 64 + c * 32, raw);
                tmem_ld_wait();
                float p[32];
                if (full_blk) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        p[i] = fast_exp2(fmaf(__uint_as_float(raw[i]), sl2, -moff));
                        sum += p[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        p[i] = (c * 32 + i < kvalid) ? fast_exp2(fmaf(__uint_as_float(raw[i]), sl2, -moff)) : 0.f;
                        sum += p[i];
                    }
                }
                store_p_chunk(smem + kFwdP, row, half * 64 + c * 32, p);
            }
            l += sum;
            tc_fence_before();
            fence_proxy_async();                                    // generic-proxy smem writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(s_free);
                mbar_arrive(p_full);
            }
        }
        mbar_wait(o_done, (nkv - 1) & 1);
        tc_fence_after();
        // ---- merge the two halves: (m, l) of the other group through smem (the P tile is free now) ----
        float2* xch = reinterpret_cast<float2*>(smem + kFwdP);      // [2][128]
        xch[half * 128 + row] = make_float2(m, l);
        asm volatile("bar.sync 1, 256;" ::: "memory");              // the 8 softmax warps
        const float2 other = xch[(half ^ 1) * 128 + row];
        const float m_all = fmaxf(m, other.x);
        const float w_me = fast_exp2((m - m_all) * sl2), w_ot = fast_exp2((other.x - m_all) * sl2);   // 0 for an empty half
        const float l_all = l * w_me + other.y * w_ot;
        const float wa = half ? w_ot : w_me, wb = half ? w_me : w_ot;                                 // weights of O_a, O_b
        const int q = q0 + row;
        const float inv = 1.f / l_all;
        // this thread writes output columns [half*32, half*32 + 32) of its row
        __nv_bfloat16* op = g.O + (static_cast<long long>(b) * g.L + q) * g.o_ld + h * 64 + half * 32;
        {
            uint32_t oa[32], ob[32];
            tmem_ld32(tmem_O + lane_off + half * 32, oa);
            tmem_ld32(tmem_O + 64 + lane_off + half * 32, ob);
            tmem_ld_wait();
            if (q < g.L) {
                const float ca = wa * inv, cb = wb * inv;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(oa[8 * i + e]) * ca + __uint_as_float(ob[8 * i + e]) * cb;
                    uint4 w;
                    w.x = pack_bf16(v[0], v[1]);
                    w.y = pack_bf16(v[2], v[3]);
                    w.z = pack_bf16(v[4], v[5]);
                    w.w = pack_bf16(v[6], v[7]);
                    *reinterpret_cast<uint4*>(op + i * 8) = w;
                }
            }
        }
        if (half == 0 && q < g.L) g.LSE[(static_cast<long long>(b) * g.H + h) * g.L + q] = m_all * g.scale + logf(l_all);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

// -----------------------------------------------------------------------------------------------------------------
// Forward, P kept in tensor memory.  ncu on the kernel above: the shared-memory data pipe is the busiest unit (per
// 128-key block: 80 KB of MMA operand reads + 32 KB of P stores + 32 KB of TMA writes ~ 1260 clk of 128 B/clk, twice per
// SM with two resident CTAs), ahead of MUFU (1024 clk) and the tensor pipe (512 clk).  Here the softmax warps write P
// (bf16, two keys per 32-bit column) straight into TMEM with tcgen05.st and O += P.V reads its A operand from there:
// no P tile in shared memory, no generic->async proxy fence, and half the operand traffic of the P.V product.
// TMEM (256 columns): S [0,128), O [128,192), P [192,256).  One O accumulator means the two softmax warp groups (keys
// 0-63 / 64-127 of every block) must exponentiate against the same reference maximum: they swap their half-row maxima
// through a 2 KB shared-memory slot (double-buffered by block parity, one 64-thread named barrier per block).
// -----------------------------------------------------------------------------------------------------------------
constexpr int kFtsQ = 0;
constexpr int kFtsK = kFtsQ + kFaTile;         // 2 stages
constexpr int kFtsV = kFtsK + 2 * kFaTile;     // 2 stages
constexpr int kFtsX = kFtsV + 2 * kFaTile;     // [2 parities][2 halves][128 rows] fp32
constexpr int kFtsBar = kFtsX + 2048;
constexpr int kFtsSmem = kFtsBar + 128;

// 64-thread named barrier 2 + quarter with an immediate id (a register id makes ptxas reserve all 16 barriers of the CTA)
__device__ __forceinline__ void pair_bar_sync(int quarter) {
    switch (quarter) {
        case 0: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
    }
}

// kPoly: every kPoly-th exponential of a row goes to poly_exp2 instead of MUFU.EX2 (0: none).  Measured (scripts/gpu_r2ab.sh,
// B = 2, H = 20, L = 1024): all MUFU 30.2 us, a quarter on the FMA pipe 30.0 us, all on the FMA pipe 37.9 us, NO exponential
// at all 26.0 us - the exponentials are 14 % of the kernel; what bounds a block is the TMEM read port (64 B/clk per SM
// sub-partition: S is read twice, 2 x 256 clk per block and CTA) and the barrier / commit latencies around it.  With the packed
// arithmetic every 4th or 3rd pair on the FMA pipe measure the same (134.6 / 134.8 us at L = 4096), every 2nd is slower (142.5).
template <int kPoly>
__global__ void __launch_bounds__(kFaThreads, 2) flash_fwd_ts_kernel(const __grid_constant__ FlashFwdArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFtsBar);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;      // [2]
    uint64_t* v_full = bars + 3;      // [2]
    uint64_t* kv_empty = bars + 5;    // [2]
    uint64_t* s_full = bars + 7;
    uint64_t* s_free = bars + 8;
    uint64_t* p_full = bars + 9;
    uint64_t* o_done = bars + 10;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (g.Lk + 127) >> 7;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.mapQ);
        tma_prefetch_desc(&g.mapK);
        tma_prefetch_desc(&g.mapV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(s_free, 8);
        mbar_init(p_full, 8);
        mbar_init(o_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                   // nothing above touched global memory
    const uint32_t tmem = *tmem_ptr;
    const uint32_t tmem_S = tmem, tmem_O = tmem + 128, tmem_P = tmem + 192;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, kFaTile);
            tma_load_4d(smem + kFtsQ, &g.mapQ, q_full, 0, q0, h, b);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], kFaTile);
                tma_load_4d(smem + kFtsK + s * kFaTile, &g.mapK, &k_full[s], 0, j * 128, h, b);
                mbar_expect_tx(&v_full[s], kFaTile);
                tma_load_4d(smem + kFtsV + s * kFaTile, &g.mapV, &v_full[s], 0, j * 128, h, b);
                tma_load_4d(smem + kFtsV + s * kFaTile + 8192, &g.mapV, &v_full[s], 0, j * 128 + 64, h, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, 0, 0);     // S[128q x 128k] = Q (K-major) . K^T (K-major)
            const uint32_t idesc_o = umma_idesc_bf16(64, 0, 1);      // O[128q x 64d] += P (TMEM) . V (MN-major)
            const uint32_t sq = smem_u32(smem + kFtsQ);
            auto issue_s = [&](int j) {
                const uint32_t sk = smem_u32(smem + kFtsK + (j & 1) * kFaTile);
                const uint64_t ad = umma_desc(sq, 16, 1024), bd = umma_desc(sk, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, ad + 2 * k, bd + 2 * k, idesc_s, k > 0);
                umma_commit(s_full);
            };
            const bool tlm = kFaTimeline && g.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            if (tlm) g.timeline[7] = clock64();
            issue_s(0);
            for (int j = 0; j < nkv; ++j) {
                const int s = j & 1;
                if (j + 1 < nkv) {
                    mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
                    mbar_wait(s_free, j & 1);              // the softmax warps have pulled S_j out of TMEM
                    tc_fence_after();
                    if (tlm) g.timeline[j * 8 + 5] = clock64();
                    issue_s(j + 1);
                }
                mbar_wait(p_full, j & 1);                   // P_j written to TMEM, O rescaled
                mbar_wait(&v_full[s], (j >> 1) & 1);
                tc_fence_after();
                if (tlm) g.timeline[j * 8 + 6] = clock64();
                const uint32_t sv = smem_u32(smem + kFtsV + s * kFaTile);
                const uint64_t vd = umma_desc(sv, 8192, 1024);
#pragma unroll
                for (int k = 0; k < 8; ++k)                 // 16 keys = 8 packed columns of P per step
                    umma_bf16_ts(tmem_O, tmem_P + 8 * k, vd + 128 * k, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(&kv_empty[s]);
                umma_commit(o_done);
            }
        }
        __syncwarp();
    } else {
        // ---- softmax / correction / epilogue: thread <-> (query row, half of the keys of every block) ----
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const int half = (warp >= 6) ? 1 : 0;
        const int quarter = warp & 3;                               // the two warps of a TMEM lane quarter share a named barrier
        const uint32_t lane_off = static_cast<uint32_t>(lane_base) << 16;
        const float sl2 = g.scale * 1.4426950408889634f;           // scores are used in the log2 domain
        float* xch = reinterpret_cast<float*>(smem + kFtsX);
        // m: the reference maximum the exponentials are taken against (common to both halves of a row).  It only moves
        // when the row's running maximum outgrows it by more than 2^8 (log2 domain), so O (in TMEM) is rescaled on a few
        // blocks instead of on every one; the final O / l and LSE = m.scale + ln(l) are exact whatever reference was used.
        float m = -INFINITY, l = 0.f;
        const bool tl = kFaTimeline && g.timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            if (tl) g.timeline[j * 8 + 0] = clock64();
            const int kvalid = min(128, g.Lk - j * 128) - half * 64;   // valid keys among this group's 64 (may be <= 0)
            const bool full_blk = kvalid >= 64;                     // warp-uniform
            // pass 1: row maximum over this group's keys, then over both groups
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_S + lane_off + half * 64 + c * 32, raw);
                tmem_ld_wait();
                if (full_blk) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(raw[i]));
                }
            }
            if (tl) g.timeline[j * 8 + 1] = clock64();
            xch[((j & 1) * 2 + half) * 128 + row] = mx;
            pair_bar_sync(quarter);
            mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * 128 + row]);
            const float m_cand = fmaxf(m, mx);
            const bool need = (m_cand - m) * sl2 > 8.f;             // true on the first block (m = -inf)
            const bool any_need = __any_sync(0xffffffffu, need);
            // the previous P.V must have retired before O is rescaled (rare) and before P is overwritten (after the
            // exponentials: by then it has, and the wait costs nothing)
            if (j > 0 && any_need) {
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
            }
            if (tl) g.timeline[j * 8 + 2] = clock64();
            if (any_need) {
                const float alpha = need ? fast_exp2((m - m_cand) * sl2) : 1.f;   // 0 when m was -inf
                if (need) m = m_cand;
                l *= alpha;
                if (j > 0) {                                        // this group rescales its 32 columns of O
                    uint32_t o[32];
                    tmem_ld32(tmem_O + lane_off + half * 32, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st32(tmem_O + lane_off + half * 32, o);
                }
            }
            const float moff = m * sl2;
            // pass 2: exponentials, row sum, P -> TMEM (bf16 pairs)
            float sum = 0.f;
            uint32_t pk[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_S + lane_off + half * 64 + c * 32, raw);
                tmem_ld_wait();
                if (c == 1) {                                       // S_j is out of TMEM: the next Q.K^T may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                }
                float p[32];
                if (full_blk) {
                    // packed fp32x2 arithmetic (FFMA2 / FADD2): half the issue slots of the scalar form
                    const float2 sl2v = make_float2(sl2, sl2), noff = make_float2(-moff, -moff);
                    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 x = __ffma2_rn(make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])), sl2v, noff);
                        const bool on_fma = kPoly > 0 && ((i >> 1) % (kPoly > 0 ? kPoly : 1)) == (kPoly - 1);
                        const float2 pv = on_fma ? poly_exp2x2(x) : make_float2(fast_exp2(x.x), fast_exp2(x.y));
                        sum2 = __fadd2_rn(sum2, pv);
                        p[i] = pv.x;
                        p[i + 1] = pv.y;
                    }
                    sum += sum2.x + sum2.y;
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        p[i] = (c * 32 + i < kvalid) ? fast_exp2(fmaf(__uint_as_float(raw[i]), sl2, -moff)) : 0.f;
                        sum += p[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[c * 16 + i] = pack_bf16(p[2 * i], p[2 * i + 1]);
            }
            l += sum;
            if (tl) g.timeline[j * 8 + 3] = clock64();
            if (j > 0 && !any_need) {
                mbar_wait(o_done, (j - 1) & 1);
                tc_fence_after();
            }
            tmem_st32(tmem_P + lane_off + half * 32, pk);           // keys [half*64, +64) = packed columns [half*32, +32)
            tmem_st_wait();                                         // P (and a rescaled O) are in TMEM
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            if (tl) g.timeline[j * 8 + 4] = clock64();
        }
        mbar_wait(o_done, (nkv - 1) & 1);
        tc_fence_after();
        // ---- row sums of the two halves (same reference m) ----
        xch[(nkv & 1) * 256 + half * 128 + row] = l;
        pair_bar_sync(quarter);
        const float l_all = l + xch[(nkv & 1) * 256 + (half ^ 1) * 128 + row];
        const int q = q0 + row;
        const float inv = 1.f / l_all;
        // this thread writes output columns [half*32, half*32 + 32) of its row
        __nv_bfloat16* op = g.O + (static_cast<long long>(b) * g.L + q) * g.o_ld + h * 64 + half * 32;
        {
            uint32_t oa[32];
            tmem_ld32(tmem_O + lane_off + half * 32, oa);
            tmem_ld_wait();
            if (q < g.L) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 w;
                    w.x = pack_bf16(__uint_as_float(oa[8 * i + 0]) * inv, __uint_as_float(oa[8 * i + 1]) * inv);
                    w.y = pack_bf16(__uint_as_float(oa[8 * i + 2]) * inv, __uint_as_float(oa[8 * i + 3]) * inv);
                    w.z = pack_bf16(__uint_as_float(oa[8 * i + 4]) * inv, __uint_as_float(oa[8 * i + 5]) * inv);
                    w.w = pack_bf16(__uint_as_float(oa[8 * i + 6]) * inv, __uint_as_float(oa[8 * i + 7]) * inv);
                    *reinterpret_cast<uint4*>(op + i * 8) = w;
                }
            }
        }
        if (half == 0 && q < g.L) g.LSE[(static_cast<long long>(b) * g.H + h) * g.L + q] = m * g.scale + logf(l_all);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}


// =================================================================================================================
// backward
// =================================================================================================================
constexpr int kBwdThreads = 448;               // TMA, MMA, 4 softmax warps (keys 0-63), 4 dQ-epilogue warps, 4 softmax warps (keys 64-127)
constexpr int kBwdK = 0;
constexpr int kBwdV = kBwdK + kFaTile;
// Q / dO ring: a stage is held from its S / dP product until the dK / dV products of the same block retire, and a refill
// takes a TMA round trip (~1.5k clk) - with two stages the S / dP issue of block i+1 waited for that round trip after
// block i-1 retired (timeline, B200_FLASH_TIMELINE=1: ~1300 clk per block); with three the refill is a block ahead.
constexpr int kBwdStages = 3;
constexpr int kBwdQ = kBwdV + kFaTile;                   // kBwdStages stages
constexpr int kBwdDO = kBwdQ + kBwdStages * kFaTile;     // kBwdStages stages
constexpr int kBwdP = kBwdDO + kBwdStages * kFaTile;     // [128 q x 128 k] bf16
constexpr int kBwdDS = kBwdP + 2 * kFaTile;    // [128 q x 128 k] bf16
constexpr int kBwdDQ = kBwdDS + 2 * kFaTile;   // dQ block staged as fp32: two [128 q x 32 d] 128B-swizzled boxes (32 KiB)
constexpr int kBwdBar = kBwdDQ + 2 * kFaTile;
constexpr int kBwdSmem = kBwdBar + 256;

struct FlashBwdArgs {
    CUtensorMap mapQ, mapDO, mapK, mapV;       // dims (64, rows, H, B); box (64, 128)
    CUtensorMap mapDQ;                         // fp32 dQ accumulator, dims (ld, L, B); box (32, 128, 1): TMA reduce-add target
    const float* LSE;                          // [B, H, L]
    const float* Delta;                        // [B, H, L] = rowsum(dO * O)
    float* dQacc;                              // fp32 [B*L, ld] accumulated over key blocks (zeroed by the caller)
    __nv_bfloat16* dQ;                         // Lk <= 128 (one key block, e.g. every cross-attention layer): dQ has a
                                               // single contribution and is written here directly as bf16
    int dq_direct;
    __nv_bfloat16* dK;                         // [B*Lk, ld]
    __nv_bfloat16* dV;
    int L, Lk, H;
    long long ld;
    float scale;
    // Query split.  One CTA per (key block, h, b) that walks all L/128 query blocks leaves SMs idle when the grid is small
    // (cross-attention: 40 CTAs at SDXL's 1280-wide level) and quantises badly otherwise (self-attention at L = 1024: 320
    // CTAs on 148 SMs = 3 rounds for 2.16 rounds of work).  So the items of the last, partial round (or all of them, for a
    // small grid) are cut into nsplit contiguous query ranges, sized so that the pieces fill the machine once.  dQ blocks
    // are disjoint (written directly or reduce-added as before); every piece adds its partial dK / dV (fp32 atomics) into
    // dKVacc [2][B*Lk, ld], and the LAST piece of a (b, h, key block) to finish - counted in `counters` - rounds the sums
    // to bf16 into dK / dV and re-zeroes accumulators and counter, so the workspace is clean for the next launch.
    // Gradient of the head-summed pre-softmax scores (the DAAM hook, trainer/ti_cross_attn_loss.py:201-212): the hook's
    // score is sum_h scale * q_h . k_h, so its gradient dSc[b, q, key] (bf16, the same for every head) enters exactly where
    // dS does:  dQ_h = (dS_h + scale * dSc) . K_h,  dK_h = (dS_h + scale * dSc)^T . Q_h - no separate GEMMs.  Columns
    // [0, dsc_cols) of each row are readable (a multiple of 8; the padding beyond Lk is zero).  NULL: no hook.
    const __nv_bfloat16* dSc;
    long long ld_dsc;
    int dsc_cols;
    int nsplit;                                // query ranges per split item
    int n_unsplit;                             // leading work items that are not split
    int nkb;                                   // key blocks (counters are per (b, h, key block))
    int nbatch_rows;                           // B * Lk: rows of one accumulator plane
    float* dKVacc;
    int* counters;
    long long* timeline;                       // debug (B200_FLASH_TIMELINE=1): clock64 stamps of CTA 0, [query block][8]; else NULL
};

template <bool kHook>
__global__ void __launch_bounds__(kBwdThreads, 1) flash_bwd_kernel(const __grid_constant__ FlashBwdArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdBar);
    uint64_t* kv_full = bars + 0;
    uint64_t* qdo_full = bars + 1;     // [kBwdStages]
    uint64_t* qdo_empty = bars + 4;    // [kBwdStages]
    uint64_t* sdp_full = bars + 7;
    uint64_t* sdp_free = bars + 8;
    uint64_t* pds_full = bars + 9;
    uint64_t* pds_free = bars + 10;
    uint64_t* dq_full = bars + 11;
    uint64_t* dq_free = bars + 12;
    uint64_t* dkv_full = bars + 13;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq_all = (g.L + 127) >> 7;
    // 1-D grid over work items (key block fastest, then head, then batch).  Items [0, n_unsplit) are walked whole by one CTA
    // each; every later item is cut into g.nsplit contiguous query ranges (one CTA each) - the tail of the last, partial round
    // over the SMs, or every item when the whole grid is smaller than the machine.
    int item = static_cast<int>(blockIdx.x), qsp = 0, nsplit = 1;
    if (item >= g.n_unsplit) {
        const int t = item - g.n_unsplit;
        nsplit = g.nsplit;
        item = g.n_unsplit + t / nsplit;
        qsp = t - (t / nsplit) * nsplit;
    }
    const int kblk = item % g.nkb, h = (item / g.nkb) % g.H, b = item / (g.nkb * g.H);
    // query blocks [qb0, qb0 + nq) of this CTA; ring stages / barrier phases follow the LOCAL block index
    const int k0 = kblk * 128;
    const int qb0 = nsplit > 1 ? static_cast<int>((static_cast<long long>(nq_all) * qsp) / nsplit) : 0;
    const int nq = nsplit > 1 ? static_cast<int>((static_cast<long long>(nq_all) * (qsp + 1)) / nsplit) - qb0 : nq_all;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&g.mapQ);
        tma_prefetch_desc(&g.mapDO);
        tma_prefetch_desc(&g.mapK);
        tma_prefetch_desc(&g.mapV);
        mbar_init(kv_full, 1);
        for (int i = 0; i < kBwdStages; ++i) {
            mbar_init(&qdo_full[i], 1);
            mbar_init(&qdo_empty[i], 1);
        }
        mbar_init(sdp_full, 1);
        mbar_init(sdp_free, 8);
        mbar_init(pds_full, 8);
        mbar_init(pds_free, 1);
        mbar_init(dq_full, 1);
        mbar_init(dq_free, 4);
        mbar_init(dkv_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t t_S = tmem, t_dP = tmem + 128, t_dV = tmem + 256, t_dK = tmem + 320, t_dQ = tmem + 384;
    // dS once more, as bf16 pairs in the last 64 columns: the A operand of dQ = dS.K comes from tensor memory (TS-mode MMA), which
    // takes its 32 KB per block off the shared-memory pipe this kernel is bound by (the smem copy still feeds dK = dS^T.Q)
    const uint32_t t_dSp = tmem + 448;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * kFaTile);
            tma_load_4d(smem + kBwdK, &g.mapK, kv_full, 0, k0, h, b);
            tma_load_4d(smem + kBwdV, &g.mapV, kv_full, 0, k0, h, b);
            for (int i = 0; i < nq; ++i) {
                const int s = i % kBwdStages;
                mbar_wait(&qdo_empty[s], ((i / kBwdStages) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[s], 2 * kFaTile);
                tma_load_4d(smem + kBwdQ + s * kFaTile, &g.mapQ, &qdo_full[s], 0, (qb0 + i) * 128, h, b);
                tma_load_4d(smem + kBwdDO + s * kFaTile, &g.mapDO, &qdo_full[s], 0, (qb0 + i) * 128, h, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_kk = umma_idesc_bf16(128, 0, 0);   // [128q x 128k]: A K-major, B K-major
            const uint32_t id_mm = umma_idesc_bf16(64, 1, 1);    // [128k x 64d]:  A MN-major (P^T / dS^T), B MN-major
            const uint32_t id_km = umma_idesc_bf16(64, 0, 1);    // [128q x 64d]:  A K-major (dS), B MN-major (K)
            const uint32_t sk = smem_u32(smem + kBwdK), sv = smem_u32(smem + kBwdV);
            const uint32_t sp = smem_u32(smem + kBwdP), sds = smem_u32(smem + kBwdDS);
            mbar_wait(kv_full, 0);
            // Software-pipelined: S / dP of block i+1 are issued BEFORE dV / dK / dQ of block i, as soon as the softmax
            // warps have pulled S / dP of block i out of TMEM, so the softmax of block i+1 overlaps the three
            // gradient MMAs of block i (the loop used to be a strict MMA -> softmax -> MMA chain).
            auto issue_sdp = [&](int i) {
                const int s = i % kBwdStages;
                const uint32_t sq = smem_u32(smem + kBwdQ + s * kFaTile), sdo = smem_u32(smem + kBwdDO + s * kFaTile);
                mbar_wait(&qdo_full[s], (i / kBwdStages) & 1);
                mbar_wait(sdp_free, (i & 1) ^ 1);
                tc_fence_after();
                const uint64_t aq = umma_desc(sq, 16, 1024), bk = umma_desc(sk, 16, 1024);
                const uint64_t ado = umma_desc(sdo, 16, 1024), bv = umma_desc(sv, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(t_S, aq + 2 * k, bk + 2 * k, id_kk, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(t_dP, ado + 2 * k, bv + 2 * k, id_kk, k > 0);
                umma_commit(sdp_full);
            };
            const bool tlm = kFaTimeline && g.timeline != nullptr && blockIdx.x == 0;
            if (tlm) g.timeline[7] = clock64();
            issue_sdp(0);
            for (int i = 0; i < nq; ++i) {
                const int s = i % kBwdStages;
                const uint32_t sq = smem_u32(smem + kBwdQ + s * kFaTile), sdo = smem_u32(smem + kBwdDO + s * kFaTile);
                if (i + 1 < nq) {
                    issue_sdp(i + 1);
                    if (tlm) g.timeline[i * 8 + 4] = clock64();
                }
                mbar_wait(pds_full, i & 1);
                mbar_wait(dq_free, (i & 1) ^ 1);
                tc_fence_after();
                if (tlm) g.timeline[i * 8 + 5] = clock64();
                {
                    const uint64_t apT = umma_desc(sp, kFaTile, 1024), adsT = umma_desc(sds, kFaTile, 1024);
                    const uint64_t bdo = umma_desc(sdo, 8192, 1024), bq = umma_desc(sq, 8192, 1024);
                    const uint64_t bkm = umma_desc(sk, 8192, 1024);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16(t_dV, apT + 128 * k, bdo + 128 * k, id_mm, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16(t_dK, adsT + 128 * k, bq + 128 * k, id_mm, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 8; ++k) umma_bf16_ts(t_dQ, t_dSp + 8 * k, bkm + 128 * k, id_km, k > 0);
                }
                umma_commit(&qdo_empty[s]);
                umma_commit(pds_free);
                umma_commit(dq_full);
            }
            umma_commit(dkv_full);
        }
        __syncwarp();
    } else {
        const int lane_base = (warp & 3) * 32;
        const int row = lane_base + lane;
        const uint32_t lane_off = static_cast<uint32_t>(lane_base) << 16;
        const float sl2 = g.scale * 1.4426950408889634f;
        const int kvalid = min(128, g.Lk - k0);
        const long long stat_base = (static_cast<long long>(b) * g.H + h) * g.L;
        if (warp < 6 || warp >= 10) {
            // ---- softmax warps: rebuild P, form dS, stage both for the MMAs.  Thread <-> (query row, half of the
            //      128 keys): one warp per scheduler was the bottleneck of this kernel (issue-bound at ~1500
            //      instructions per row and block), so two warps share every TMEM lane quarter. ----
            const int chalf = (warp >= 10) ? 2 : 0;
            const bool tl = kFaTimeline && g.timeline != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0;
            for (int i = 0; i < nq; ++i) {
                const int q = (qb0 + i) * 128 + row;
                const bool qok = q < g.L;
                const float lse2 = qok ? g.LSE[stat_base + q] * 1.4426950408889634f : 0.f;
                const float delta = qok ? g.Delta[stat_base + q] : 0.f;
                mbar_wait(sdp_full, i & 1);
                tc_fence_after();
                if (tl) g.timeline[i * 8 + 0] = clock64();
                const bool full_blk = (kvalid == 128) && ((qb0 + i) * 128 + 128 <= g.L);     // warp-uniform
                // This thread's 64 keys go through in four 16-column steps.  The TMEM loads of step c+1 are issued before the
                // arithmetic of step c (tcgen05.wait::ld waits for every outstanding load, so the wait sits after it), the
                // first half of the P / dS row is stored as soon as the gradient MMAs of block i-1 have released the tiles
                // (they have, by then), and S / dP are handed back to the MMA warp before the last step's arithmetic.
                uint32_t pp[4][8], pd[4][8];                   // packed bf16 pairs
                uint32_t rs[2][16], rp[2][16];
                const int col_base = chalf * 32;               // first of this thread's 64 columns
                tmem_ld16(t_S + lane_off + col_base, rs[0]);
                tmem_ld16(t_dP + lane_off + col_base, rp[0]);
                tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c0 = col_base + cc * 16;         // columns [c0, c0 + 16) of the block
                    if (cc < 3) {
                        tmem_ld16(t_S + lane_off + c0 + 16, rs[(cc + 1) & 1]);
                        tmem_ld16(t_dP + lane_off + c0 + 16, rp[(cc + 1) & 1]);
                    }
                    const uint32_t(&xs)[16] = rs[cc & 1];
                    const uint32_t(&xp)[16] = rp[cc & 1];
                    float hk[kHook ? 16 : 1];                       // the hook instantiation only (cross-attention layers)
#pragma unroll
                    for (int e = 0; e < (kHook ? 16 : 1); ++e) hk[e] = 0.f;
                    if (kHook && qok) {
                        // this thread's slice of the hook gradient: columns k0 + c0 .. +16 of its query row
                        const int col0 = k0 + c0;
                        const __nv_bfloat16* hp = g.dSc + (static_cast<long long>(b) * g.L + q) * g.ld_dsc + col0;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (col0 + j * 8 + 8 <= g.dsc_cols) {
                                const uint4 w = *reinterpret_cast<const uint4*>(hp + j * 8);
                                const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                                for (int t2 = 0; t2 < 4; ++t2) {
                                    hk[kHook ? j * 8 + 2 * t2 : 0] = __uint_as_float(u[t2] << 16) * g.scale;
                                    hk[kHook ? j * 8 + 2 * t2 + 1 : 0] = __uint_as_float(u[t2] & 0xffff0000u) * g.scale;
                                }
                            }
                        }
                    }
                    if (full_blk) {
                        // packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2): the same roundings as the scalar form in half
                        // the issue slots
                        const float2 sl2v = make_float2(sl2, sl2), nlse = make_float2(-lse2, -lse2);
                        const float2 ndel = make_float2(-delta, -delta), scv = make_float2(g.scale, g.scale);
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            const float2 x = __ffma2_rn(make_float2(__uint_as_float(xs[e]), __uint_as_float(xs[e + 1])), sl2v, nlse);
                            // (moving exponentials to the FMA pipe as the forward does buys nothing here: measured equal at a
                            //  quarter, 3 % slower at a half - this loop is not MUFU-bound)
                            const float2 pv = make_float2(fast_exp2(x.x), fast_exp2(x.y));
                            const float2 t = __fadd2_rn(make_float2(__uint_as_float(xp[e]), __uint_as_float(xp[e + 1])), ndel);
                            const float2 u = __fmul2_rn(pv, t);
                            const float2 d = kHook ? __ffma2_rn(u, scv, make_float2(hk[kHook ? e : 0], hk[kHook ? e + 1 : 0]))
                                                   : __fmul2_rn(u, scv);
                            pp[cc][e >> 1] = pack_bf16(pv.x, pv.y);
                            pd[cc][e >> 1] = pack_bf16(d.x, d.y);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; e += 2) {
                            float pe[2], de[2];
#pragma unroll
                            for (int t2 = 0; t2 < 2; ++t2) {
                                const bool ok = qok && (c0 + e + t2 < kvalid);
                                pe[t2] = ok ? fast_exp2(fmaf(__uint_as_float(xs[e + t2]), sl2, -lse2)) : 0.f;
                                de[t2] = ok ? fmaf(pe[t2] * (__uint_as_float(xp[e + t2]) - delta), g.scale, hk[kHook ? e + t2 : 0]) : 0.f;
                            }
                            pp[cc][e >> 1] = pack_bf16(pe[0], pe[1]);
                            pd[cc][e >> 1] = pack_bf16(de[0], de[1]);
                        }
                    }
                    if (cc < 3) tmem_ld_wait();
                    if (cc == 1) {
                        if (tl) g.timeline[i * 8 + 1] = clock64();
                        mbar_wait(pds_free, (i & 1) ^ 1);     // the gradient MMAs of block i-1 have read the P / dS tiles
                        if (tl) g.timeline[i * 8 + 2] = clock64();
                        store_packed16(smem + kBwdP, row, col_base, pp[0]);
                        store_packed16(smem + kBwdP, row, col_base + 16, pp[1]);
                        store_packed16(smem + kBwdDS, row, col_base, pd[0]);
                        store_packed16(smem + kBwdDS, row, col_base + 16, pd[1]);
                        uint32_t w[16];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            w[e] = pd[0][e];
                            w[8 + e] = pd[1][e];
                        }
                        tmem_st16(t_dSp + lane_off + (col_base >> 1), w);     // keys [col_base, +32) = packed columns [col_base/2, +16)
                    }
                    if (cc == 2) {
                        // S / dP are out of TMEM: the MMA warp may start block i+1 while this block is still being finished
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(sdp_free);
                    }
                }
                store_packed16(smem + kBwdP, row, col_base + 32, pp[2]);
                store_packed16(smem + kBwdP, row, col_base + 48, pp[3]);
                store_packed16(smem + kBwdDS, row, col_base + 32, pd[2]);
                store_packed16(smem + kBwdDS, row, col_base + 48, pd[3]);
                {
                    uint32_t w[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        w[e] = pd[2][e];
                        w[8 + e] = pd[3][e];
                    }
                    tmem_st16(t_dSp + lane_off + (col_base >> 1) + 16, w);
                    tmem_st_wait();
                    tc_fence_before();
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(pds_full);
                if (tl) g.timeline[i * 8 + 3] = clock64();
            }
        } else {
            // ---- dQ epilogue warps: dQ_blk is summed over key blocks in an fp32 global accumulator.  It leaves through
            //      TMA reduce-add (cp.reduce.async.bulk.tensor .add.f32) from a swizzled smem staging tile: one bulk
            //      operation per [128 x 32] box instead of 2048 red.global.add.v4 per block, which had this kernel bound
            //      by the L2 atomic units. ----
            const bool issuer = (warp == 6 && lane == 0);
            uint8_t* stg = smem + kBwdDQ;
            const bool tlq = kFaTimeline && g.timeline != nullptr && blockIdx.x == 0 && warp == 6 && lane == 0;
            for (int i = 0; i < nq; ++i) {
                mbar_wait(dq_full, i & 1);
                tc_fence_after();
                if (tlq) g.timeline[i * 8 + 6] = clock64();
                if (g.dq_direct) {
                    // single key block: no accumulation over blocks -> straight to bf16, thread <-> query row
                    const int q = (qb0 + i) * 128 + row;
                    __nv_bfloat16* dst = g.dQ + (static_cast<long long>(b) * g.L + q) * g.ld + h * 64;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t r[32];
                        tmem_ld32(t_dQ + lane_off + c * 32, r);
                        tmem_ld_wait();
                        if (q < g.L) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                uint4 w;
                                w.x = pack_bf16(__uint_as_float(r[8 * e + 0]), __uint_as_float(r[8 * e + 1]));
                                w.y = pack_bf16(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]));
                                w.z = pack_bf16(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]));
                                w.w = pack_bf16(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]));
                                *reinterpret_cast<uint4*>(dst + c * 32 + e * 8) = w;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dq_free);
                    continue;
                }
                if (i > 0) {
                    if (issuer) bulk_wait_read<0>();            // the previous block's reductions have read the staging tile
                    __syncwarp();
                    named_bar_sync(2, 128);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_dQ + lane_off + c * 32, r);
                    tmem_ld_wait();
                    uint8_t* rowp = stg + c * kFaTile + row * 128;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        *reinterpret_cast<uint4*>(rowp + ((e ^ (row & 7)) * 16)) =
                            make_uint4(r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(dq_free);            // TMEM dQ drained: the next block's MMA may overwrite it
                named_bar_sync(2, 128);
                if (issuer) {
                    tma_reduce_add_4d(&g.mapDQ, stg, h * 64, (qb0 + i) * 128, b, 0);
                    tma_reduce_add_4d(&g.mapDQ, stg + kFaTile, h * 64 + 32, (qb0 + i) * 128, b, 0);
                    bulk_commit();
                }
                __syncwarp();
            }
            if (issuer) bulk_wait<0>();                         // every reduction performed before the grid completes
        }
        // ---- dV (softmax warps) / dK (epilogue warps): thread <-> key row ----
        mbar_wait(dkv_full, 0);
        tc_fence_after();
        const int key = k0 + row;
        const uint32_t src = (warp < 6) ? t_dV : t_dK;
        __nv_bfloat16* out = ((warp < 6) ? g.dV : g.dK) + (static_cast<long long>(b) * g.Lk + key) * g.ld + h * 64;
        float* acc = nsplit > 1 ? g.dKVacc + ((warp < 6) ? static_cast<long long>(g.nbatch_rows) * g.ld : 0LL) +
                                      (static_cast<long long>(b) * g.Lk + key) * g.ld + h * 64
                                : nullptr;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (warp >= 10) break;                 // the second softmax group has no output tile of its own (warp-uniform)
            uint32_t r[32];
            tmem_ld32(src + lane_off + c * 32, r);
            tmem_ld_wait();
            if (key < g.Lk) {
                if (acc != nullptr) {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        atomicAdd(reinterpret_cast<float4*>(acc + c * 32 + e * 4),
                                  make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]),
                                              __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3])));
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint4 w;
                        w.x = pack_bf16(__uint_as_float(r[8 * e + 0]), __uint_as_float(r[8 * e + 1]));
                        w.y = pack_bf16(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]));
                        w.z = pack_bf16(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]));
                        w.w = pack_bf16(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]));
                        *reinterpret_cast<uint4*>(out + c * 32 + e * 8) = w;
                    }
                }
            }
        }
        tc_fence_before();
        if (nsplit > 1) __threadfence();           // this thread's partial sums are visible before the CTA is counted
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
    if (nsplit > 1) {
        // the last CTA of this (b, h) to arrive owns the final rounding: fp32 sums -> bf16 dK / dV, workspace re-zeroed
        __shared__ int s_last;
        if (threadIdx.x == 0) {
            const int prev = atomicAdd(&g.counters[(b * g.H + h) * g.nkb + kblk], 1);
            s_last = (prev == nsplit - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const long long half = static_cast<long long>(g.nbatch_rows) * g.ld;
            const int kvalid2 = min(128, g.Lk - k0);               // keys of this block
            // <= 128 keys x 16 float4 x {dK, dV} = 4096 float4 over 448 threads: all of a thread's loads are issued before its
            // first store (a plain load-store loop exposes one L2 round trip per iteration at the very end of the kernel)
            constexpr int kPer = (128 * 16 * 2 + kBwdThreads - 1) / kBwdThreads;
            const int total = kvalid2 * 16 * 2;
            float4 v[kPer];
            long long offs[kPer];
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int idx = threadIdx.x + u * kBwdThreads;
                offs[u] = -1;
                if (idx < total) {
                    const int which = idx / (kvalid2 * 16);        // 0: dK, 1: dV
                    const int rem = idx - which * kvalid2 * 16;
                    const int key = k0 + (rem >> 4), c4 = (rem & 15) * 4;
                    offs[u] = which * half + (static_cast<long long>(b) * g.Lk + key) * g.ld + h * 64 + c4;
                    v[u] = __ldcg(reinterpret_cast<const float4*>(g.dKVacc + offs[u]));
                }
            }
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                if (offs[u] >= 0) {
                    __stcg(reinterpret_cast<float4*>(g.dKVacc + offs[u]), make_float4(0.f, 0.f, 0.f, 0.f));
                    const bool is_dv = offs[u] >= half;
                    uint2 w;
                    w.x = pack_bf16(v[u].x, v[u].y);
                    w.y = pack_bf16(v[u].z, v[u].w);
                    *reinterpret_cast<uint2*>((is_dv ? g.dV : g.dK) + (offs[u] - (is_dv ? half : 0))) = w;
                }
            }
            if (threadIdx.x == 0) g.counters[(b * g.H + h) * g.nkb + kblk] = 0;
        }
    }
}

// delta[b, h, l] = sum_d dO * O: eight lanes per (row, head), one 16-byte load of O and of dO each, a 3-step shuffle reduction
// (one thread per (row, head) walking 2 x 128 bytes was latency-bound: 7-9 us for 10 MB).  The same threads zero the fp32 dQ
// accumulator of the multi-key-block backward (dq_zero: [B*L, H*64] contiguous, or NULL), which used to be a memset node.
__global__ void flash_delta_kernel(const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO,
                                   float* __restrict__ delta, float* __restrict__ dq_zero, int B, int L, int H, long long ld) {
    pdl_launch();
    pdl_wait();
    const long long total = static_cast<long long>(B) * L * H * 8;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;        // a multiple of 32: groups of 8 stay together
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx - (threadIdx.x & 31) < total; idx += stride) {   // warp-uniform trip count (full-mask shuffles below)
        const bool live = idx < total;                     // total is a multiple of 8: a group is live or dead as a whole
        const int sub = static_cast<int>(idx & 7);
        const long long grp = idx >> 3;
        const int hh = static_cast<int>(grp % H);
        const long long bl = grp / H;
        float acc = 0.f;
        if (live) {
            const uint4 a = *reinterpret_cast<const uint4*>(O + bl * ld + hh * 64 + sub * 8);
            const uint4 c = *reinterpret_cast<const uint4*>(dO + bl * ld + hh * 64 + sub * 8);
            const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc += __bfloat162float(ha[e].x) * __bfloat162float(hc[e].x);
                acc += __bfloat162float(ha[e].y) * __bfloat162float(hc[e].y);
            }
            if (dq_zero != nullptr) {
                float4* z = reinterpret_cast<float4*>(dq_zero + (bl * H + hh) * 64 + sub * 8);
                z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (live && sub == 0) {
            const int l = static_cast<int>(bl % L);
            const int bb = static_cast<int>(bl / L);
            delta[(static_cast<long long>(bb) * H + hh) * L + l] = acc;
        }
    }
}

// x: contiguous fp32 rows of 4*row4 elements -> y: bf16 rows with stride ld_y (elements)
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4, int row4,
                                   long long ld_y) {
    pdl_launch();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        uint2 w;
        w.x = pack_bf16(v.x, v.y);
        w.y = pack_bf16(v.z, v.w);
        const long long row = i / row4;
        const int c4 = static_cast<int>(i - row * row4);
        *reinterpret_cast<uint2*>(y + row * ld_y + 4 * c4) = w;
    }
}

}  // namespace b200
