// Batched LoRA weight gradients: ONE launch for every  dB = dY^T . T  and  dA = U^T . X  of a transformer block.
//
//   out_p[n, j] += sum_{m < M} X_p[m, n] * Y_p[m, j]        X_p: [M, Nout] bf16 (dY or the layer input), Y_p: [M, r] (T or U)
//
// These problems are 16..32 columns wide: 84 MFLOP over 10 MB of operands each - pure data movement.  As single launches
// of the tcgen05 GEMM they ran at 1-2 % of the tensor peak and cost ~9 us of launch / prologue / drain latency apiece,
// 420 times a step (profiles/r01j_gemm_roofline_table.md).  Here a table of problems (pointers and extents, passed by
// value in the kernel parameters - no tensor maps, so ANY set of operand addresses batches) is cut into
// [m-chunk x 64 columns] work items, one CTA each; operands stream through a 3-stage cp.async ring (L2-resident: the
// tensors were produced microseconds earlier), the products run on mma.sync m16n8k16 (the rank-r output is far too narrow
// for a tcgen05 tile; the kernel is bound by L2 -> SM bytes, not by math), and every item adds its partial sums to the fp32
// gradient buffer with atomics (the flat buffer the optimizer and the data-parallel all-reduce read).
#include "common.cuh"
#include "../../include/b200_lora.h"

namespace b200 {

constexpr int kWgMaxProblems = 32;
constexpr int kWgTileN = 64;        // output rows (n) per CTA
constexpr int kWgTileM = 64;        // reduction rows per pipeline stage
constexpr int kWgStages = 3;
constexpr int kWgMaxR = 32;
constexpr int kWgThreads = 128;

struct WgProblem {
    const bf16* X;
    const bf16* Y;
    float* out;
    long long ld_x, ld_y, out_sn, out_sj;
    int M, Nout, r;
    int item0;          // first work item of this problem
    int msplit;         // m-chunks
    int mchunk;         // rows per chunk (multiple of kWgTileM)
};

struct WgArgs {
    int nproblems;
    int total_items;
    WgProblem p[kWgMaxProblems];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    const int sz = valid ? 16 : 0;                   // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tiles: X stage [64 m][64 n] bf16 = rows of 128 B, 16-byte chunk c of row m stored at chunk (c ^ (m & 7));
//             Y stage [64 m][32 j] bf16 = rows of 64 B,  chunk c of row m stored at chunk (c ^ ((m >> 1) & 3)).
// Both swizzles make the 8 row addresses of every ldmatrix 8x8 block fall into distinct bank groups.
__global__ void __launch_bounds__(kWgThreads) lora_wgrad_batch_kernel(const __grid_constant__ WgArgs g) {
    __shared__ __align__(128) uint8_t sX[kWgStages][kWgTileM * 128];
    __shared__ __align__(128) uint8_t sY[kWgStages][kWgTileM * 64];
    pdl_launch();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- which work item (problem table is in the constant bank; read before the dependency wait) ----
    const int item = blockIdx.x;
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < g.nproblems; ++i)
        if (item >= g.p[i].item0) pi = i;
    const WgProblem& P = g.p[pi];
    const int local = item - P.item0;
    const int n_tile = local / P.msplit, m_chunk = local - n_tile * P.msplit;
    const int n0 = n_tile * kWgTileN;
    const int m_begin = m_chunk * P.mchunk;
    const int m_end = min(P.M, m_begin + P.mchunk);
    const int Nout = P.Nout, r = P.r;
    const bf16* __restrict__ X = P.X;
    const bf16* __restrict__ Y = P.Y;
    const long long ld_x = P.ld_x, ld_y = P.ld_y;
    const int nsteps = (m_end - m_begin + kWgTileM - 1) / kWgTileM;
    const int jt = (r + 7) >> 3;                     // 8-wide output column tiles in use (<= 4)
    pdl_wait();

    auto load_stage = [&](int st, int step) {
        const int mb = m_begin + step * kWgTileM;
        // X: 64 rows x 8 chunks = 512 chunks, 4 per thread
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = tid + q * kWgThreads;
            const int row = c >> 3, ch = c & 7;
            const int m = mb + row, n = n0 + ch * 8;
            const bool ok = (m < m_end) && (n < Nout);
            cp_async16(&sX[st][row * 128 + ((ch ^ (row & 7)) << 4)], ok ? X + static_cast<long long>(m) * ld_x + n : X, ok);
        }
        // Y: 64 rows x 4 chunks = 256 chunks, 2 per thread (chunks past the rank are zero-filled)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = tid + q * kWgThreads;
            const int row = c >> 2, ch = c & 3;
            const int m = mb + row;
            const bool ok = (m < m_end) && (ch < jt);
            cp_async16(&sY[st][row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)], ok ? Y + static_cast<long long>(m) * ld_y + ch * 8 : Y, ok);
        }
    };

    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

#pragma unroll
    for (int s = 0; s < kWgStages - 1; ++s) {
        if (s < nsteps) load_stage(s, s);
        cp_async_commit();
    }
    // this warp owns output rows n0 + warp*16 .. +16; ldmatrix row-address roles of this lane
    const int mat = lane >> 3, rr = lane & 7;
    for (int step = 0; step < nsteps; ++step) {
        cp_async_wait<kWgStages - 2>();
        __syncthreads();
        {   // prefetch the stage that was consumed in the previous iteration
            const int nxt = step + kWgStages - 1;
            if (nxt < nsteps) load_stage(nxt % kWgStages, nxt);
            cp_async_commit();
        }
        const int st = step % kWgStages;
#pragma unroll
        for (int kk = 0; kk < kWgTileM / 16; ++kk) {
            // A fragment (16 n x 16 m) from X stored [m][n]: matrices (n 0-7, m 0-7), (n 8-15, m 0-7), (n 0-7, m 8-15), (n 8-15, m 8-15)
            uint32_t a[4];
            {
                const int row = kk * 16 + rr + ((mat & 2) ? 8 : 0);
                const int ch = warp * 2 + (mat & 1);
                ldsm_x4_t(a, &sX[st][row * 128 + ((ch ^ (row & 7)) << 4)]);
            }
            // B fragments (16 m x 8 j) from Y stored [m][j]: matrices (m 0-7, j-tile t), (m 8-15, j-tile t), (m 0-7, t+1), (m 8-15, t+1)
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
                if (t2 * 2 < jt) {
                    uint32_t b[4];
                    const int row = kk * 16 + rr + ((mat & 1) ? 8 : 0);
                    const int ch = t2 * 2 + (mat >> 1);
                    ldsm_x4_t(b, &sY[st][row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)]);
                    mma_bf16(acc[t2 * 2], a, b[0], b[1]);
                    if (t2 * 2 + 1 < jt) mma_bf16(acc[t2 * 2 + 1], a, b[2], b[3]);
                }
            }
        }
    }
    cp_async_wait<0>();
    // ---- partial sums -> fp32 gradient buffer (atomics; several m-chunks and the other data-parallel micro-steps add here) ----
    float* __restrict__ out = P.out;
    const long long sn = P.out_sn, sj = P.out_sj;
    const int n_a = n0 + warp * 16 + (lane >> 2), n_b = n_a + 8;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (t < jt) {
            const int j0 = t * 8 + (lane & 3) * 2;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j0 + e;
                if (j < r) {
                    if (n_a < Nout) atomicAdd(out + n_a * sn + j * sj, acc[t][e]);
                    if (n_b < Nout) atomicAdd(out + n_b * sn + j * sj, acc[t][2 + e]);
                }
            }
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_lora_wgrad_batch(const b200_wgrad_problem_t* problems, int32_t n, void* stream) {
    B200_CHECK_ARG(problems != nullptr && n >= 1 && n <= kWgMaxProblems, "lora_wgrad_batch: 1..%d problems per launch", kWgMaxProblems);
    WgArgs g;
    memset(&g, 0, sizeof(g));
    long long tiles = 0;
    for (int i = 0; i < n; ++i) {
        const b200_wgrad_problem_t& q = problems[i];
        B200_CHECK_ARG(q.X && q.Y && q.out && q.M >= 1 && q.Nout >= 1 && q.r >= 1 && q.r <= kWgMaxR,
                       "lora_wgrad_batch: problem %d: bad pointers / extents (rank <= %d)", i, kWgMaxR);
        B200_CHECK_ARG(q.ld_x % 8 == 0 && q.ld_y % 8 == 0 && q.Nout % 8 == 0 && q.ld_x >= q.Nout && q.ld_y >= ((q.r + 7) / 8) * 8 &&
                           reinterpret_cast<uintptr_t>(q.X) % 16 == 0 && reinterpret_cast<uintptr_t>(q.Y) % 16 == 0,
                       "lora_wgrad_batch: problem %d: operands need 16-byte aligned rows (extents / strides multiples of 8)", i);
        tiles += (q.Nout + kWgTileN - 1) / kWgTileN;
    }
    // m-chunks: enough work items for ~6 CTAs per SM, chunks of at least 256 rows
    int item = 0;
    for (int i = 0; i < n; ++i) {
        const b200_wgrad_problem_t& q = problems[i];
        WgProblem& w = g.p[i];
        w.X = static_cast<const bf16*>(q.X);
        w.Y = static_cast<const bf16*>(q.Y);
        w.out = q.out;
        w.ld_x = q.ld_x;
        w.ld_y = q.ld_y;
        w.out_sn = q.out_sn;
        w.out_sj = q.out_sj;
        w.M = q.M;
        w.Nout = q.Nout;
        w.r = q.r;
        long long want = (6LL * kNumSMs + tiles - 1) / tiles;
        const long long max_split = (q.M + 255) / 256;
        if (want > max_split) want = max_split;
        if (want < 1) want = 1;
        int mchunk = static_cast<int>((q.M + want - 1) / want);
        mchunk = (mchunk + kWgTileM - 1) / kWgTileM * kWgTileM;
        w.mchunk = mchunk;
        w.msplit = (q.M + mchunk - 1) / mchunk;
        w.item0 = item;
        item += ((q.Nout + kWgTileN - 1) / kWgTileN) * w.msplit;
    }
    g.nproblems = n;
    g.total_items = item;
    launch_pdl(lora_wgrad_batch_kernel, dim3(item), dim3(kWgThreads), 0, static_cast<cudaStream_t>(stream), g);
    B200_CHECK_LAUNCH("lora_wgrad_batch");
    return 0;
}
