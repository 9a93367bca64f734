// Generic bf16 tensor-core GEMM for sm_100a: TMA -> smem (128B swizzle) -> tcgen05.mma -> TMEM -> epilogue.
//
//   D[b][m, n] = epi( sum_{seg} sum_k A_seg[b][m, k] * B_seg[b][n, k] )        (fp32 accumulate in TMEM)
//
// One persistent CTA per SM (6 warps: TMA producer warp, MMA issuer, 4 epilogue warps), 128 x BN output tiles with BN a
// RUNTIME multiple of 16 (<= 256) so the host can size the grid to a whole number of waves, 4-stage smem ring,
// double-buffered TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// What makes it the LoRA hot-path kernel rather than a plain GEMM:
//   * up to TWO K-segments accumulate into the SAME TMEM tile: segment 0 is the frozen W.x (or an implicit 3x3
//     convolution), segment 1 the low-rank side path (T = s.x.A^T against B), so  W.x + s.B.(A.x)  is one kernel;
//   * each operand is K-major or MN-major (descriptor + instruction-descriptor bits), so dX = dY.W, dA = U^T.X and
//     dB = dY^T.T read the SAME row-major tensors the forward used - no transposed weight copies, no dW ever;
//   * segment 0's A operand can be an NHWC activation walked as an implicit im2col: one 4-D TMA box per
//     (tap, 64-channel block), out-of-image rows/columns zero-filled by TMA;
//   * batched (two batch dims, per-operand) for the attention GEMMs, split-K with fp32 atomic accumulation for the
//     skinny LoRA weight-gradient GEMMs, and an epilogue with alpha, (per-image) bias and residual add.
#pragma once
#include "ptx.cuh"
#include "common.cuh"

namespace b200 {

constexpr int kBM = 128;            // tile rows (UMMA M)
constexpr int kBK = 64;             // K elements per pipeline stage (= one 128 B swizzle row of bf16)
constexpr int kStages = 4;            // ring slots at the widest tile (48 KiB each); narrower tiles get more slots
constexpr int kMaxStages = 12;
constexpr int kMaxBN = 256;
constexpr int kABytes = kBM * kBK * 2;      // 16 KiB
constexpr int kBBytes = kMaxBN * kBK * 2;   // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kGemmThreads = 192;
constexpr int kEpiLd = 36;                 // floats per staged row: 32 + 4 pad (144 B keeps float4 alignment, no bank conflicts)
constexpr int kEpiBytes = 4 * 32 * kEpiLd * 4;
constexpr int kSideMaxBN = 192;             // side accumulator sits at TMEM column 192 of an accumulator stage
constexpr int kSideCol = 192;               // TMEM column (inside an accumulator stage) of the rank-r side accumulator
constexpr int kTBytes = kBM * 128;          // T/U staged as a 128B-swizzled K-major A operand (only r columns used)
constexpr int kTOff = kStages * kStageBytes;            // 1024-byte aligned: it is read through a swizzled descriptor
constexpr int kBarOff = kTOff + kTBytes;
constexpr int kEpiOff = kBarOff + 512;
constexpr int kGemmSmemBytes = kEpiOff + kEpiBytes;
static_assert(kTOff % 1024 == 0 && kStageBytes % 1024 == 0, "swizzled tiles need 1024-byte alignment");

struct GemmArgs {
    CUtensorMap mapA[2];
    CUtensorMap mapB[2];
    int M, N;                 // per-batch output extent
    int tiles_m, tiles_n;
    int nb0, nb1;             // batch extents (b0 fastest)
    int splits;               // split-K factor (segment 0 only; requires num_seg == 1, conv == 0)
    int BN;
    int num_seg;
    int kblocks[2];           // 64-wide K blocks per segment (conv: 9 * conv_cblocks)
    int ktail16[2];           // 16-wide MMA steps in the LAST block of the segment (1..4)
    int a_mn[2], b_mn[2];     // operand is MN-major (else K-major)
    int a_batched[2], b_batched[2];
    // implicit 3x3 / pad 1 / stride 1 convolution on segment 0's A operand (NHWC, map dims C,W,H,N)
    int conv;
    int conv_cblocks;
    int conv_W, conv_H, conv_bh;   // conv_bh = image rows per box
    int b_tap_k, b_tap_n;          // per-tap offsets into B's (k, n) coordinates
    // epilogue
    void* D;
    int d_fp32;               // 0: bf16, 1: fp32
    int d_atomic;             // fp32 only: red.global.add instead of store
    long long d_sm, d_sn, d_sb0, d_sb1;
    float alpha;
    const __nv_bfloat16* bias;
    int bias_rows;            // rows sharing one bias vector (0: one vector for all rows)
    long long bias_sb;        // stride between bias vectors
    const __nv_bfloat16* R;   // residual, added after alpha/bias (nullptr: none)
    long long r_sm, r_sn, r_sb0, r_sb1;
    // fused low-rank side path (single segment only):  D += (side_alpha * A . S^T) . B2^T
    //   main loop : Tacc[128 x r16] += A_tile . S_tile   (S = LoRA A fwd / LoRA B bwd), in TMEM next to the main tile
    //   T-phase   : epilogue warps scale + round Tacc to bf16, stage it in smem as an A operand, n_blk 0 writes it out
    //   final MMA : main tile += T . B2_tile              (B2 = LoRA B fwd / LoRA A bwd)
    // smem ring geometry (host-computed): slot = A tile | B tile | side tile; as many slots as fit in 192 KiB
    int stage_bytes, num_stages, side_off;
    int side;
    int side_mn, b2_mn;       // operand majorness of S and B2
    int side_r16;             // rank rounded up to 16 (16 or 32): MMA N of the side accumulate, K of the final MMA
    int side_r;               // true rank (columns written to T_out)
    float side_alpha;
    __nv_bfloat16* T_out;     // [M, t_ld] bf16 (batch 0 only; the side path is not batched)
    long long t_ld;
    CUtensorMap mapS, mapB2;
    int vec_ok;               // epilogue may use 8/16-byte vector accesses (host-checked alignment)
    // group mode: the two "segments" are two INDEPENDENT problems of identical tiling (the dB and dA weight-gradient
    // GEMMs of one LoRA layer) sharing one launch; problem 1 writes D2 with its own strides
    int group;
    void* D2;
    long long d2_sm, d2_sn;
    int dbg_mode;             // developer probe: 1 = epilogue skips global stores, 2 = skips the smem read-back
    long long* dbg;           // developer probe: per-CTA globaltimer stamps [cta][8] (nullptr in production)
};

__device__ __forceinline__ void dbg_stamp(long long* dbg, int slot) {
    if (dbg) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[blockIdx.x * 16 + slot] = t;
    }
}

__device__ __forceinline__ void advance_stage(int& stage, uint32_t& phase, int num_stages) {
    if (++stage == num_stages) {
        stage = 0;
        phase ^= 1;
    }
}

// kEpi: 0 = bf16 row-major output, 1 = fp32 row-major output, 2 = transposed / strided / atomic output
template <int kEpi>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmArgs g) {
    // 128B swizzle atoms must sit on 1024 B boundaries.
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full_bar = empty_bar + kMaxStages;    // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
    uint64_t* side_full_bar = tmem_empty_bar + 2;     // [2]  side accumulator complete (MMA -> epilogue warps)
    uint64_t* t_ready_bar = side_full_bar + 2;        // T staged in smem (epilogue warps -> MMA)
    uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(t_ready_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) dbg_stamp(g.dbg, 0);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < g.num_seg; ++s) {
            tma_prefetch_desc(&g.mapA[s]);
            tma_prefetch_desc(&g.mapB[s]);
        }
        for (int i = 0; i < g.num_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 4);
            mbar_init(&side_full_bar[i], 1);
        }
        mbar_init(t_ready_bar, 4);
        if (g.side) {
            tma_prefetch_desc(&g.mapS);
            tma_prefetch_desc(&g.mapB2);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_base_ptr, 512);
        tmem_relinquish();
    }
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                   // barrier init / TMEM alloc above overlap the previous grid's tail
    if (threadIdx.x == 0) dbg_stamp(g.dbg, 1);
    const uint32_t tmem_base = *tmem_base_ptr;

    const int tiles_per_batch = g.tiles_m * g.tiles_n * g.splits;
    const int total_tiles = tiles_per_batch * g.nb0 * g.nb1 * (g.group ? 2 : 1);
    const int BN = g.BN;
    const int b_boxes_mn = (BN + 63) >> 6;
    const uint32_t b_bytes_k = static_cast<uint32_t>(BN) * 128u;
    const uint32_t b_bytes_mn = static_cast<uint32_t>(b_boxes_mn) * 8192u;

    if (warp == 0) {
        // ===================== TMA producer warp =====================
        // Issuing a cp.async.bulk.tensor costs the issuing thread a few hundred cycles, so the boxes of one ring slot
        // are issued by DIFFERENT lanes in parallel: lane 0 waits for the slot and arms the barrier, then
        // lane 0/1 -> A box(es), lanes 2..5 -> B box(es), lane 6 -> side tile.
        {
            int stage = 0;
            uint32_t phase = 0;
            const bool simple = (g.nb0 * g.nb1 * g.splits) == 1 && !g.group;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int n_blk, m_blk, split = 0, b0 = 0, b1 = 0;
                int seg_lo = 0, seg_hi = g.num_seg, tl = tile;
                if (g.group) {
                    seg_lo = tile / tiles_per_batch;
                    seg_hi = seg_lo + 1;
                    tl = tile - seg_lo * tiles_per_batch;
                }
                if (simple) {
                    m_blk = static_cast<int>(static_cast<unsigned>(tile) / static_cast<unsigned>(g.tiles_n));
                    n_blk = tile - m_blk * g.tiles_n;
                } else {
                    int t = tl;
                    n_blk = t % g.tiles_n;  t /= g.tiles_n;
                    m_blk = t % g.tiles_m;  t /= g.tiles_m;
                    split = t % g.splits;   t /= g.splits;
                    b0 = t % g.nb0;
                    b1 = t / g.nb0;
                }
                const int m0 = m_blk * kBM, n0 = n_blk * BN;
                int conv_n0 = 0, conv_h0 = 0;
                if (g.conv) {
                    const int hw = g.conv_H * g.conv_W;
                    conv_n0 = m0 / hw;
                    conv_h0 = (m0 % hw) / g.conv_W;
                }
                for (int seg = seg_lo; seg < seg_hi; ++seg) {
                    int kb_begin = 0, kb_end = g.kblocks[seg];
                    if (g.splits > 1) {
                        kb_begin = static_cast<int>((static_cast<long long>(kb_end) * split) / g.splits);
                        kb_end = static_cast<int>((static_cast<long long>(kb_end) * (split + 1)) / g.splits);
                    }
                    const int ab0 = g.a_batched[seg] ? b0 : 0, ab1 = g.a_batched[seg] ? b1 : 0;
                    const int bb0 = g.b_batched[seg] ? b0 : 0, bb1 = g.b_batched[seg] ? b1 : 0;
                    const uint32_t side_bytes = g.side ? (g.side_mn ? 8192u : static_cast<uint32_t>(g.side_r16) * 128u) : 0u;
                    const uint32_t tx = kABytes + (g.b_mn[seg] ? b_bytes_mn : b_bytes_k) + side_bytes;
                    const bool conv_seg = g.conv && seg == 0;
                    const int a_mn = g.a_mn[seg], b_mn = g.b_mn[seg];
                    for (int kb = kb_begin; kb < kb_end; ++kb) {
                        if (lane == 0) {
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            mbar_expect_tx(&full_bar[stage], tx);
                        }
                        __syncwarp();
                        uint8_t* sa = smem + stage * g.stage_bytes;
                        uint8_t* sb = sa + kABytes;
                        uint64_t* fb = &full_bar[stage];
                        int bk = kb * kBK, bn = n0;
                        int tap = 0, cb = 0;
                        if (conv_seg) {
                            tap = kb / g.conv_cblocks;
                            cb = kb - tap * g.conv_cblocks;
                            bk = tap * g.b_tap_k + cb * kBK;
                            bn = n0 + tap * g.b_tap_n;
                        }
                        if (lane == 0) {
                            if (conv_seg) {
                                const int kh = tap / 3, kw = tap - kh * 3;
                                tma_load_4d(sa, &g.mapA[0], fb, cb * kBK, kw - 1, conv_h0 + kh - 1, conv_n0);
                            } else if (a_mn) {
                                tma_load_4d(sa, &g.mapA[seg], fb, m0, kb * kBK, ab0, ab1);
                            } else {
                                tma_load_4d(sa, &g.mapA[seg], fb, kb * kBK, m0, ab0, ab1);
                            }
                        } else if (lane == 1) {
                            if (!conv_seg && a_mn) tma_load_4d(sa + 8192, &g.mapA[seg], fb, m0 + 64, kb * kBK, ab0, ab1);
                        } else if (lane < 6) {
                            const int jb = lane - 2;
                            if (b_mn) {
                                if (jb < b_boxes_mn) tma_load_4d(sb + jb * 8192, &g.mapB[seg], fb, bn + jb * 64, bk, bb0, bb1);
                            } else if (jb == 0) {
                                tma_load_4d(sb, &g.mapB[seg], fb, bk, bn, bb0, bb1);
                            }
                        } else if (lane == 6 && g.side) {
                            if (g.side_mn) tma_load_4d(sb + g.side_off, &g.mapS, fb, 0, kb * kBK, 0, 0);
                            else tma_load_4d(sb + g.side_off, &g.mapS, fb, kb * kBK, 0, 0, 0);
                        }
                        if (lane == 0 && kb == kb_begin && seg == seg_lo && tile == static_cast<int>(blockIdx.x)) dbg_stamp(g.dbg, 2);
                        advance_stage(stage, phase, g.num_stages);
                    }
                }
                if (g.side) {
                    // one more ring slot per tile: the B2 tile the final rank-r MMA multiplies T with
                    if (lane == 0) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], g.b2_mn ? b_bytes_mn : b_bytes_k);
                    }
                    __syncwarp();
                    uint8_t* sb = smem + stage * g.stage_bytes + kABytes;
                    if (lane >= 2 && lane < 6) {
                        const int jb = lane - 2;
                        if (g.b2_mn) {
                            if (jb < b_boxes_mn) tma_load_4d(sb + jb * 8192, &g.mapB2, &full_bar[stage], n0 + jb * 64, 0, 0, 0);
                        } else if (jb == 0) {
                            tma_load_4d(sb, &g.mapB2, &full_bar[stage], 0, n0, 0, 0);
                        }
                    }
                    advance_stage(stage, phase, g.num_stages);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer warp =====================
        // The whole warp runs the loop in lock-step (barrier waits, descriptor arithmetic stay warp-uniform and live in
        // uniform registers); only the tcgen05 instructions themselves are issued by one elected lane.
        {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0, t_phase = 0;
            const uint32_t smem_base = smem_u32(smem);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int seg_lo = 0, seg_hi = g.num_seg, tl = tile;
                if (g.group) {
                    seg_lo = tile / tiles_per_batch;
                    seg_hi = seg_lo + 1;
                    tl = tile - seg_lo * tiles_per_batch;
                }
                const int split = (g.splits > 1) ? (tl / (g.tiles_m * g.tiles_n)) % g.splits : 0;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kMaxBN);
                const uint32_t tmem_side = tmem_d + kSideCol;
                const uint32_t idesc_side = umma_idesc_bf16(g.side_r16, g.a_mn[0], g.side_mn);
                const uint32_t s_step = g.side_mn ? (2048u >> 4) : (32u >> 4);
                uint32_t accumulate = 0;
                for (int seg = seg_lo; seg < seg_hi; ++seg) {
                    int kb_begin = 0, kb_end = g.kblocks[seg];
                    const int kb_last = kb_end - 1;
                    if (g.splits > 1) {
                        kb_begin = static_cast<int>((static_cast<long long>(kb_end) * split) / g.splits);
                        kb_end = static_cast<int>((static_cast<long long>(kb_end) * (split + 1)) / g.splits);
                    }
                    const uint32_t idesc = umma_idesc_bf16(BN, g.a_mn[seg], g.b_mn[seg]);
                    // per-16-K-step advance of the descriptor start address (encoded >> 4)
                    const uint32_t a_step = g.a_mn[seg] ? (2048u >> 4) : (32u >> 4);
                    const uint32_t b_step = g.b_mn[seg] ? (2048u >> 4) : (32u >> 4);
                    const uint32_t a_lbo = g.a_mn[seg] ? 8192u : 16u, b_lbo = g.b_mn[seg] ? 8192u : 16u;
                    for (int kb = kb_begin; kb < kb_end; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + static_cast<uint32_t>(stage * g.stage_bytes);
                        const uint32_t sb = sa + kABytes;
                        const uint64_t adesc = umma_desc(sa, a_lbo, 1024);
                        const uint64_t bdesc = umma_desc(sb, b_lbo, 1024);
                        const int n16 = (kb == kb_last) ? g.ktail16[seg] : 4;
                        if (elect_one()) {
                            if (kb == kb_begin && seg == seg_lo && tile == static_cast<int>(blockIdx.x)) dbg_stamp(g.dbg, 3);
                            if (g.side) {
                                const uint64_t sdesc = umma_desc(sb + g.side_off, g.side_mn ? 8192u : 16u, 1024);
                                for (int k = 0; k < n16; ++k)
                                    umma_bf16(tmem_side, adesc + static_cast<uint64_t>(a_step * k),
                                              sdesc + static_cast<uint64_t>(s_step * k), idesc_side, accumulate | (k > 0));
                            }
                            if (g.dbg_mode == 3 && !g.b_mn[seg] && (BN % 32) == 0) {
                                // experiment: two independent accumulation chains (N halves) interleaved
                                const uint32_t idesc_h = umma_idesc_bf16(BN / 2, g.a_mn[seg], 0);
                                const uint64_t bdesc_h = umma_desc(sb + static_cast<uint32_t>(BN / 2) * 128u, 16, 1024);
                                for (int k = 0; k < n16; ++k) {
                                    umma_bf16(tmem_d, adesc + static_cast<uint64_t>(a_step * k),
                                              bdesc + static_cast<uint64_t>(b_step * k), idesc_h, accumulate | (k > 0));
                                    umma_bf16(tmem_d + BN / 2, adesc + static_cast<uint64_t>(a_step * k),
                                              bdesc_h + static_cast<uint64_t>(b_step * k), idesc_h, accumulate | (k > 0));
                                }
                            } else {
                                for (int k = 0; k < n16; ++k)
                                    umma_bf16(tmem_d, adesc + static_cast<uint64_t>(a_step * k),
                                              bdesc + static_cast<uint64_t>(b_step * k), idesc, accumulate | (k > 0));
                            }
                            umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                        }
                        __syncwarp();
                        accumulate = 1;
                        advance_stage(stage, phase, g.num_stages);
                    }
                }
                if (g.side) {
                    if (elect_one()) umma_commit(&side_full_bar[acc]);   // rank-r accumulator complete -> T-phase
                    __syncwarp();
                    mbar_wait(&full_bar[stage], phase);   // B2 tile landed
                    mbar_wait(t_ready_bar, t_phase);      // T (bf16) staged in smem
                    t_phase ^= 1;
                    tc_fence_after();
                    const uint32_t sb2 = smem_base + static_cast<uint32_t>(stage * g.stage_bytes) + kABytes;
                    const uint64_t tdesc = umma_desc(smem_base + kTOff, 16, 1024);
                    const uint64_t b2desc = umma_desc(sb2, g.b2_mn ? 8192u : 16u, 1024);
                    const uint32_t idesc2 = umma_idesc_bf16(BN, 0, g.b2_mn);
                    const uint32_t b2_step = g.b2_mn ? (2048u >> 4) : (32u >> 4);
                    if (elect_one()) {
                        for (int k = 0; k < (g.side_r16 >> 4); ++k)
                            umma_bf16(tmem_d, tdesc + static_cast<uint64_t>(2 * k), b2desc + static_cast<uint64_t>(b2_step * k),
                                      idesc2, 1);
                        umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    advance_stage(stage, phase, g.num_stages);
                }
                if (elect_one()) {
                    umma_commit(&tmem_full_bar[acc]);         // accumulator complete -> epilogue
                    if (tile == static_cast<int>(blockIdx.x)) dbg_stamp(g.dbg, 4);
                }
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps (TMEM -> registers -> smem transpose -> coalesced global) ==========
        // tcgen05.ld hands every thread ONE ROW of the tile; writing that straight out would touch 32 rows per store
        // instruction.  Each warp therefore bounces 32x32 fp32 sub-tiles through a padded smem patch and re-reads
        // them so that 8 consecutive lanes own one 128-byte row segment: alpha / bias / residual are applied in that
        // coalesced layout and leave as 16-byte (fp32) or 8-byte (bf16) vector stores.
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;    // TMEM lane quarter this warp may read
        float* stage = reinterpret_cast<float*>(smem + kEpiOff) + ew * (32 * kEpiLd);
        // vector stores / residual loads need 8-byte (bf16) or 16-byte (fp32) aligned rows: decided once, warp-uniform
        const bool vec_ok = g.vec_ok != 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int t = tile;
            void* Dp = g.D;
            long long dsm = g.d_sm, dsn = g.d_sn;
            if (g.group && tile >= tiles_per_batch) {
                t = tile - tiles_per_batch;
                Dp = g.D2;
                dsm = g.d2_sm;
                dsn = g.d2_sn;
            }
            const int n_blk = t % g.tiles_n;  t /= g.tiles_n;
            const int m_blk = t % g.tiles_m;  t /= g.tiles_m;
            t /= g.splits;
            const int b0 = t % g.nb0;
            const int b1 = t / g.nb0;
            const int m_warp = m_blk * kBM + lane_base;
            const int n0 = n_blk * BN;
            const long long d_boff = b1 * g.d_sb1 + b0 * g.d_sb0;
            const long long r_boff = b1 * g.r_sb1 + b0 * g.r_sb0;

            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + static_cast<uint32_t>(acc * kMaxBN);
            if (g.side) {
                // ---- T-phase: Tacc (fp32, TMEM) -> alpha, bf16 -> swizzled smem A operand (+ global copy from n_blk 0)
                mbar_wait(&side_full_bar[acc], acc_phase);
                tc_fence_after();
                uint32_t raw[32];
                tmem_ld32(taddr + kSideCol, raw);
                tmem_ld_wait();
                const int row = lane_base + lane;
                uint8_t* trow = smem + kTOff + row * 128;
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(raw[2 * j]) * g.side_alpha,
                                                                    __uint_as_float(raw[2 * j + 1]) * g.side_alpha);
                    packed[j] = *reinterpret_cast<const uint32_t*>(&hh);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j * 8 < g.side_r16)
                        *reinterpret_cast<uint4*>(trow + ((j ^ (row & 7)) * 16)) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                const int m = m_warp + row - lane_base;
                if (n_blk == 0 && g.T_out != nullptr && m < g.M) {
                    __nv_bfloat16* tp = g.T_out + static_cast<long long>(m) * g.t_ld;
                    const __nv_bfloat16* pv = reinterpret_cast<const __nv_bfloat16*>(packed);
                    if ((g.side_r & 7) == 0 && (g.t_ld & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j * 8 < g.side_r)
                                *reinterpret_cast<uint4*>(tp + j * 8) =
                                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < g.side_r) tp[j] = pv[j];
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_ready_bar);
            }
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (warp == 2 && lane == 0 && tile == static_cast<int>(blockIdx.x)) dbg_stamp(g.dbg, 5);
            // Code size matters here: 148 SMs enter this at once and an unrolled, branchy epilogue thrashes the
            // instruction cache (it cost more than the whole K loop).  Hence: one template instance per output kind,
            // rolled loops, and a single warp-uniform choice between the vector and the scalar body.
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t raw[32];
                tmem_ld32(taddr + c0, raw);
                tmem_ld_wait();
                const bool dbg0 = warp == 2 && lane == 0 && tile == static_cast<int>(blockIdx.x) && c0 == 0;
                if (dbg0) dbg_stamp(g.dbg, 8);
                if (n0 + c0 >= g.N) continue;          // warp-uniform
                const int cvalid = min(32, BN - c0);
                if (kEpi != 2) {
                    float4* srow = reinterpret_cast<float4*>(stage + lane * kEpiLd);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        srow[j] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                              __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
                    __syncwarp();
                    if (dbg0) dbg_stamp(g.dbg, 9);
                    const int sub = lane >> 3, col = (lane & 7) * 4;
                    const int n = n0 + c0 + col;
                    const int nv = min(4, min(g.N - n, cvalid - col));     // columns this lane owns (<= 0: none)
                    // One warp per scheduler runs this loop, so its latency is the epilogue's latency: keep the
                    // per-iteration dependency chain short (running pointers, no 64-bit multiplies) and let four
                    // iterations overlap.
                    const float alpha = g.alpha;
                    const int m_first = m_warp + sub;
                    const float* sp = stage + sub * kEpiLd + col;
                    const long long d_step = 4 * g.d_sm, r_step = 4 * g.r_sm;
                    long long doff = d_boff + static_cast<long long>(m_first) * g.d_sm + n;
                    const __nv_bfloat16* rp = g.R ? g.R + r_boff + static_cast<long long>(m_first) * g.r_sm + static_cast<long long>(n) * g.r_sn : nullptr;
                    float bv[4] = {0.f, 0.f, 0.f, 0.f};
                    const bool bias_per_row = g.bias != nullptr && g.bias_rows != 0;
                    if (g.bias != nullptr && !bias_per_row && nv > 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (e < nv) bv[e] = __bfloat162float(g.bias[n + e]);
                    }
                    // warp-uniform fast path: interior tile (all 32 rows and all 32 columns valid), vector accesses,
                    // no per-row bias.  ~12 instructions per 4 outputs, fully unrolled so the eight rows overlap.
                    const bool interior = vec_ok && !bias_per_row && (m_warp + 32 <= g.M) && (cvalid == 32) &&
                                          (n0 + c0 + 32 <= g.N);
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
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.D) + doff + it * d_step) = make_float4(v0, v1, v2, v3);
                            } else {
                                const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1);
                                const __nv_bfloat162 h1 = __floats2bfloat162_rn(v2, v3);
                                uint2 w2;
                                w2.x = *reinterpret_cast<const uint32_t*>(&h0);
                                w2.y = *reinterpret_cast<const uint32_t*>(&h1);
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.D) + doff + it * d_step) = w2;
                            }
                        }
                    } else if (vec_ok && nv == 4) {
#pragma unroll 1
                        for (int it = 0; it < 8; ++it) {
                            const int m = m_first + it * 4;
                            if (m < g.M) {
                                const float4 q = *reinterpret_cast<const float4*>(sp + it * 4 * kEpiLd);
                                float v0 = q.x * alpha + bv[0], v1 = q.y * alpha + bv[1], v2 = q.z * alpha + bv[2], v3 = q.w * alpha + bv[3];
                                if (bias_per_row) {
                                    const __nv_bfloat16* bp = g.bias + (m / g.bias_rows) * g.bias_sb + n;
                                    v0 += __bfloat162float(bp[0]);
                                    v1 += __bfloat162float(bp[1]);
                                    v2 += __bfloat162float(bp[2]);
                                    v3 += __bfloat162float(bp[3]);
                                }
                                if (rp) {
                                    const uint2 w2 = *reinterpret_cast<const uint2*>(rp + it * r_step);
                                    v0 += __uint_as_float(w2.x << 16);
                                    v1 += __uint_as_float(w2.x & 0xffff0000u);
                                    v2 += __uint_as_float(w2.y << 16);
                                    v3 += __uint_as_float(w2.y & 0xffff0000u);
                                }
                                if (kEpi == 1) {
                                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.D) + doff + it * d_step) = make_float4(v0, v1, v2, v3);
                                } else {
                                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1);
                                    const __nv_bfloat162 h1 = __floats2bfloat162_rn(v2, v3);
                                    uint2 w2;
                                    w2.x = *reinterpret_cast<const uint32_t*>(&h0);
                                    w2.y = *reinterpret_cast<const uint32_t*>(&h1);
                                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.D) + doff + it * d_step) = w2;
                                }
                            }
                        }
                    } else if (nv > 0) {
#pragma unroll 1
                        for (int it = 0; it < 8; ++it) {
                            const int m = m_first + it * 4;
                            if (m >= g.M) break;
                            const float* q = sp + it * 4 * kEpiLd;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (e < nv) {
                                    float x = q[e] * alpha + bv[e];
                                    if (bias_per_row) x += __bfloat162float(g.bias[(m / g.bias_rows) * g.bias_sb + n + e]);
                                    if (rp) x += __bfloat162float(rp[it * r_step + e * g.r_sn]);
                                    if (kEpi == 1) reinterpret_cast<float*>(g.D)[doff + it * d_step + e] = x;
                                    else reinterpret_cast<__nv_bfloat16*>(g.D)[doff + it * d_step + e] = __float2bfloat16_rn(x);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (dbg0) dbg_stamp(g.dbg, 10);
                } else {
                    // transposed (d_sm == 1), strided or atomic outputs: lane <-> row is already the coalesced direction
                    const int m = m_warp + lane;
                    if (m < g.M) {
                        const __nv_bfloat16* bias_row = g.bias ? g.bias + (g.bias_rows ? (m / g.bias_rows) * g.bias_sb : 0) : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + c0 + j;
                            if (j >= cvalid || n >= g.N) continue;
                            float v = __uint_as_float(raw[j]) * g.alpha;
                            if (bias_row) v += __bfloat162float(bias_row[n]);
                            if (g.R) v += __bfloat162float(g.R[r_boff + static_cast<long long>(m) * g.r_sm + static_cast<long long>(n) * g.r_sn]);
                            const long long doff = d_boff + static_cast<long long>(m) * dsm + static_cast<long long>(n) * dsn;
                            if (g.d_fp32) {
                                float* dp = reinterpret_cast<float*>(Dp) + doff;
                                if (g.d_atomic) atomicAdd(dp, v);
                                else *dp = v;
                            } else {
                                reinterpret_cast<__nv_bfloat16*>(Dp)[doff] = __float2bfloat16_rn(v);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (warp == 2 && lane == 0 && tile == static_cast<int>(blockIdx.x)) dbg_stamp(g.dbg, 6);
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) dbg_stamp(g.dbg, 7);
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace b200
