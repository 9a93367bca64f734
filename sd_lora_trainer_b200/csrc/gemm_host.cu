// Host side of b200_gemm: validates the C descriptor, encodes the TMA tensor maps (driver entry point resolved
// at run time, so the library links against cudart only), picks the tile width for whole waves over 148 SMs and
// launches the persistent tcgen05 kernel.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm2_tcgen05.cuh"
#include "../../include/b200_lora.h"

namespace b200 {

std::string& last_error_ref() {
    static thread_local std::string s;
    return s;
}
int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 4-D tensor map.  dims/strides innermost first; strides in ELEMENTS (dims 1..3).  elem_bytes 2 = bf16, 4 = fp32;
// swizzle_bytes 128 | 64 | 32 | 0.
int encode_map_ex(CUtensorMap* map, const void* ptr, int elem_bytes, const long long dims[4], const long long strides[3],
                  const int box[4], int swizzle_bytes) {
    auto fn = get_encode_fn();
    if (!fn) return set_error(4, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t bdim[4], estr[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        gdim[i] = static_cast<cuuint64_t>(dims[i] < 1 ? 1 : dims[i]);
        bdim[i] = static_cast<cuuint32_t>(box[i]);
    }
    for (int i = 0; i < 3; ++i) {
        long long s = strides[i];
        if (s <= 0) s = 16 / elem_bytes;         // degenerate (extent-1) dims still need a legal stride
        gstr[i] = static_cast<cuuint64_t>(s) * elem_bytes;
        if (gstr[i] % 16) return set_error(2, "tensor map: stride %lld elements is not a multiple of 16 bytes", s);
    }
    if (reinterpret_cast<uintptr_t>(ptr) % 16) return set_error(2, "tensor map: base not 16-byte aligned");
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                    const_cast<void*>(ptr), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(4, "cuTensorMapEncodeTiled failed (%d): dims %lld,%lld,%lld,%lld box %d,%d,%d,%d", (int)r,
                         dims[0], dims[1], dims[2], dims[3], box[0], box[1], box[2], box[3]);
    return 0;
}
// 4-D bf16 tensor map, 128B swizzle (every GEMM / attention operand).
int encode_map(CUtensorMap* map, const void* ptr, const long long dims[4], const long long strides[3],
                      const int box[4]) {
    return encode_map_ex(map, ptr, 2, dims, strides, box, 128);
}

static int encode_operand(CUtensorMap* map, const b200_operand_t& op, int box_rows_kmajor, int nb0, int nb1) {
    // K-major : dims (K=inner, rows), box (64, box_rows).   MN-major: dims (MN=inner, K=rows), box (64, 64).
    // TMA zero-fills everything outside these extents, which is what handles the M/N/K tails.
    if (op.batched && ((nb0 > 1 && op.sb0 <= 0) || (nb1 > 1 && op.sb1 <= 0)))
        return set_error(2, "gemm: batched operand needs positive batch strides");
    const long long d[4] = {op.inner, op.rows, op.batched ? nb0 : 1, op.batched ? nb1 : 1};
    const long long strides[3] = {op.row_stride, op.batched ? op.sb0 : 0, op.batched ? op.sb1 : 0};
    const int box[4] = {64, op.mn_major ? 64 : box_rows_kmajor, 1, 1};
    return encode_map(map, op.ptr, d, strides, box);
}

// Tile width: whole waves over 148 persistent CTAs, tie-break towards wider tiles (fewer re-reads of A).
static int pick_block_n(long long tiles_m_total, int N, int max_bn) {
    const int n16 = ((N + 15) / 16) * 16;
    if (n16 <= max_bn && tiles_m_total * 1 >= kNumSMs / 2) return n16 < 16 ? 16 : n16;
    int best = 16;
    double best_cost = 1e30;
    const int hi = n16 < max_bn ? n16 : max_bn;
    for (int bn = hi; bn >= 32; bn -= 16) {
        const long long tiles = tiles_m_total * ((N + bn - 1) / bn);
        const long long waves = (tiles + kNumSMs - 1) / kNumSMs;
        const double cost = static_cast<double>(waves) * (bn + 24.0);   // 24 ~ per-tile fixed cost in columns
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_version(void) { return B200_LORA_ABI_VERSION; }
extern "C" const char* b200_last_error(void) { return last_error_ref().c_str(); }
extern "C" long long b200_launch_count(void) { return g_launches.load(); }



static long long* g_gemm_dbg = nullptr;
// developer probe (not part of the public header): per-CTA globaltimer stamps of the next GEMM launches
extern "C" void b200_debug_gemm_stamps(long long* dev_buf) { g_gemm_dbg = dev_buf; }

// ---------------------------------------------------------------------------------------------------------
// CTA-pair kernel (gemm2_tcgen05.cuh): eligibility, tile width, launch
// ---------------------------------------------------------------------------------------------------------
static int pair_k0(const b200_gemm_t* d) { return d->conv ? 9 * ((d->conv_C + 63) / 64) * 64 : d->K[0]; }

static bool pair_eligible(const b200_gemm_t* d) {
    if (d->nb0 != 1 || d->nb1 != 1 || d->splits != 1 || d->d_sn != 1) return false;
    if (d->d_atomic && !d->d_fp32) return false;           // accumulate mode: fp32 output, plain read-modify-write per tile
    if (d->num_seg < 1 || d->num_seg > 2 || (d->side && (d->num_seg != 1 || d->conv))) return false;
    if (d->conv) {
        const int W = d->conv_W, H = d->conv_H;
        if (d->B[0].mn_major || W < 1 || W > 128 || 128 % W) return false;
        int bh = 128 / W;
        if (bh > H) bh = H;
        if (H % bh || 128 % (W * bh)) return false;
        const int bimg = 128 / (W * bh);
        if (!(bimg == 1 || bh == H) || d->conv_C % 8) return false;
        if (static_cast<long long>(d->conv_N) * H * W != d->M) return false;
    } else if (d->A[0].mn_major && (d->side || d->num_seg != 1 || d->A[0].inner % 8)) {
        return false;                                      // MN-major A (weight gradients): one plain segment
    }
    if (d->num_seg == 2 && (d->A[1].mn_major || d->A[1].batched || d->B[1].batched)) return false;
    if (d->M < 256 || d->N < 64 || pair_k0(d) < 64) return false;
    if (d->R && d->r_sn != 1) return false;
    if (d->group) return false;
    return true;
}

// Estimated time of the pair kernel for tile width bn (seconds); drives the choice of bn and of the kernel.
static double pair_cost(const b200_gemm_t* d, int bn, bool side) {
    const long long tiles = static_cast<long long>((d->M + 255) / 256) * ((d->N + bn - 1) / bn);
    const int pairs = kNumSMs / 2;
    const long long waves = (tiles + pairs - 1) / pairs;
    const double kblocks = (pair_k0(d) + 63) / 64 + (d->num_seg == 2 ? (d->K[1] + 63) / 64 : 0);
    const double active = tiles < pairs ? tiles : pairs;
    const double t_mma = kblocks * 2.0 * bn / 1.8e9;                                    // 4 MMAs of bn/2 clocks per block
    const double t_l2 = kblocks * (32768.0 + 128.0 * bn) * active / 11.0e12;            // L2 -> SM fabric shared by the pairs
    const double t_tile = t_mma > t_l2 ? t_mma : t_l2;
    // drain of the queued MMAs + epilogue of the last tile: measured 2.1 us at bn = 160, 3.4 us at bn = 256
    // (profiles/r01f_gemm2_timeline.txt); a single accumulator stage (side path, bn > 192) serialises it on every wave
    const double t_epi = bn * 0.0135e-6;
    const bool single_acc = side && bn > kSideMaxBN;
    return waves * t_tile + (single_acc ? waves : 1) * t_epi + (single_acc ? 0.5e-6 : 0.0) + 2.0e-6;
}

static int pick_pair_bn(const b200_gemm_t* d, double* cost_out) {
    const bool side = d->side != 0;
    const bool need64 = d->B[0].mn_major || (side && d->B2.mn_major) || (d->num_seg == 2 && d->B[1].mn_major);
    int best = 0;
    double best_cost = 1e30;
    for (int bn = 256; bn >= 64; bn -= 32) {
        if (need64 && (bn % 128)) continue;
        if (bn > ((d->N + 31) / 32) * 32 && bn != 64 && !(need64 && bn == 128)) continue;   // wider than the problem
        const double c = pair_cost(d, bn, side);
        if (c < best_cost - 1e-12) {
            best_cost = c;
            best = bn;
        }
    }
    if (cost_out) *cost_out = best_cost;
    return best;
}

static int launch_gemm2(const b200_gemm_t* d, int bn, void* stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
        if (e != cudaSuccess) return set_error(3, "gemm2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const b200_operand_t& A = d->A[0];
    const b200_operand_t& B = d->B[0];
    const int K = d->K[0];
    const bool seg2 = d->num_seg == 2;
    B200_CHECK_ARG(bn % 16 == 0 && bn >= 32 && bn <= 256, "gemm2: block_n %d must be a multiple of 16 in [32, 256]", bn);
    B200_CHECK_ARG(!(B.mn_major || (d->side && d->B2.mn_major) || (seg2 && d->B[1].mn_major)) || bn % 128 == 0,
                   "gemm2: MN-major B needs block_n 128 or 256");
    Gemm2Args g;
    memset(&g, 0, sizeof(g));
    g.M = d->M;
    g.N = d->N;
    g.BN = bn;
    g.tiles_n = (d->N + bn - 1) / bn;
    g.total_tiles = ((d->M + 255) / 256) * g.tiles_n;
    g.b_mn = B.mn_major;
    g.nseg = d->num_seg;
    if (d->conv) {
        const int W = d->conv_W, H = d->conv_H, C = d->conv_C;
        int bh = 128 / W;
        if (bh > H) bh = H;
        const int bimg = 128 / (W * bh);
        const long long dims[4] = {C, W, H, d->conv_N};
        const long long strides[3] = {C, static_cast<long long>(W) * C, static_cast<long long>(H) * W * C};
        const int box[4] = {64, W, bh, bimg};
        if (int rc = encode_map(&g.mapA, A.ptr, dims, strides, box)) return rc;
        g.conv = 1;
        g.conv_cblocks = (C + kBK - 1) / kBK;
        g.conv_W = W;
        g.conv_H = H;
        g.b_tap_k = d->b_tap_k;
        g.b_tap_n = d->b_tap_n;
        g.kblocks = 9 * g.conv_cblocks;
        g.ktail16 = 4;                       // channel tails are zero-filled by TMA on the A side
    } else {
        B200_CHECK_ARG(A.mn_major ? A.rows >= K : A.inner >= K, "gemm2: A smaller than K");
        const long long dA[4] = {A.inner, A.rows, 1, 1}, sA[3] = {A.row_stride, 0, 0};
        const int boxA[4] = {64, A.mn_major ? 64 : 128, 1, 1};   // MN-major: dims (M, K), two 64 x 64 boxes per CTA and k-block
        if (int rc = encode_map(&g.mapA, A.ptr, dA, sA, boxA)) return rc;
        g.a_mn = A.mn_major;
        g.kblocks = (K + kBK - 1) / kBK;
        g.ktail16 = (K - (g.kblocks - 1) * kBK + 15) / 16;
    }
    {   // B: K-major box 64 x bn/2, MN-major box 64 x 64
        const long long dB[4] = {B.inner, B.rows, 1, 1}, sB[3] = {B.row_stride, 0, 0};
        const int boxB[4] = {64, B.mn_major ? 64 : bn / 2, 1, 1};
        if (int rc = encode_map(&g.mapB, B.ptr, dB, sB, boxB)) return rc;
    }
    int b_bytes = B.mn_major ? (bn / 128) * 8192 : (bn / 2) * 128;
    int side_bytes = 0;
    if (seg2) {
        const b200_operand_t& A2 = d->A[1];
        const b200_operand_t& Bs = d->B[1];
        const int K2 = d->K[1];
        B200_CHECK_ARG(K2 >= 1 && A2.inner >= K2, "gemm2: segment 1 A smaller than K");
        const long long dA2[4] = {A2.inner, A2.rows, 1, 1}, sA2[3] = {A2.row_stride, 0, 0};
        const int boxA2[4] = {64, 128, 1, 1};
        if (int rc = encode_map(&g.mapA2, A2.ptr, dA2, sA2, boxA2)) return rc;
        const long long dB2[4] = {Bs.inner, Bs.rows, 1, 1}, sB2[3] = {Bs.row_stride, 0, 0};
        const int boxB2[4] = {64, Bs.mn_major ? 64 : bn / 2, 1, 1};
        if (int rc = encode_map(&g.mapB2, Bs.ptr, dB2, sB2, boxB2)) return rc;
        g.b2_mn = Bs.mn_major;
        g.kblocks2 = (K2 + kBK - 1) / kBK;
        g.ktail16_2 = (K2 - (g.kblocks2 - 1) * kBK + 15) / 16;
        const int b2_bytes = Bs.mn_major ? (bn / 128) * 8192 : (bn / 2) * 128;
        if (b2_bytes > b_bytes) b_bytes = b2_bytes;
    }
    if (d->side) {
        B200_CHECK_ARG(d->side_r >= 1 && d->side_r <= 64, "gemm2: side rank %d out of range", d->side_r);
        g.side = 1;
        g.side_r = d->side_r;
        g.side_r16 = ((d->side_r + 15) / 16) * 16;
        g.side_mn = d->S.mn_major;
        g.b2_mn = d->B2.mn_major;
        g.side_alpha = d->side_alpha;
        g.T_out = static_cast<__nv_bfloat16*>(d->T_out);
        g.t_ld = d->t_ld;
        const long long dS[4] = {d->S.inner, d->S.rows, 1, 1}, sS[3] = {d->S.row_stride, 0, 0};
        const int boxS[4] = {64, d->S.mn_major ? 64 : g.side_r16 / 2, 1, 1};
        if (int rc = encode_map(&g.mapS, d->S.ptr, dS, sS, boxS)) return rc;
        const long long dB2[4] = {d->B2.inner, d->B2.rows, 1, 1}, sB2[3] = {d->B2.row_stride, 0, 0};
        const int boxB2[4] = {64, d->B2.mn_major ? 64 : bn / 2, 1, 1};
        if (int rc = encode_map(&g.mapB2, d->B2.ptr, dB2, sB2, boxB2)) return rc;
        const int b2_bytes = d->B2.mn_major ? (bn / 128) * 8192 : (bn / 2) * 128;
        if (b2_bytes > b_bytes) b_bytes = b2_bytes;
        side_bytes = d->S.mn_major ? 8192 : ((g.side_r16 / 2) * 128 + 1023) / 1024 * 1024;
    }
    b_bytes = (b_bytes + 1023) / 1024 * 1024;
    g.side_off = b_bytes;
    g.stage_bytes = k2ABytes + b_bytes + side_bytes;
    g.num_stages = k2RingBytes / g.stage_bytes;
    if (g.num_stages > kMaxStages) g.num_stages = kMaxStages;
    g.acc_stages = (d->side && bn > kSideMaxBN) ? 1 : 2;
    g.dbg = g_gemm_dbg;
    static const int prefetch_env = getenv("B200_PREFETCH_B") ? atoi(getenv("B200_PREFETCH_B")) : 0;   // measured: 80.7 vs 79.8 ms/step with it on (profiles/r01g_ab.txt)
    g.b_static = (d->b_static && prefetch_env) ? 1 : 0;
    g.D = d->D;
    g.d_sm = d->d_sm;
    g.d_accum = d->d_atomic ? 1 : 0;
    g.alpha = d->alpha;
    g.bias = static_cast<const __nv_bfloat16*>(d->bias);
    g.bias_rows = d->bias_rows;
    g.bias_sb = d->bias_sb;
    g.R = static_cast<const __nv_bfloat16*>(d->R);
    g.r_sm = d->r_sm;
    const int esz = d->d_fp32 ? 4 : 2;
    g.vec_ok = (reinterpret_cast<uintptr_t>(d->D) % (4 * esz) == 0) && (d->d_sm % 4 == 0) &&
               (!d->R || (reinterpret_cast<uintptr_t>(d->R) % 8 == 0 && d->r_sm % 4 == 0));
    static const int tma_epi_env = getenv("B200_TMA_EPI") ? atoi(getenv("B200_TMA_EPI")) : 1;
    auto al16 = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    if (tma_epi_env && !d->d_fp32 && bn % 32 == 0 && al16(d->D) && d->d_sm % 8 == 0 &&
        (!d->R || (al16(d->R) && d->r_sm % 8 == 0)) && (!d->bias || (al16(d->bias) && d->bias_sb % 8 == 0))) {
        const long long dD[4] = {d->N, d->M, 1, 1}, sD[3] = {d->d_sm, 0, 0};
        const int boxD[4] = {32, 32, 1, 1};
        if (int rc = encode_map_ex(&g.mapD, d->D, 2, dD, sD, boxD, 64)) return rc;
        g.tma_store = 1;
    }
    static const int tma_f32_env = getenv("B200_TMA_EPI_F32") ? atoi(getenv("B200_TMA_EPI_F32")) : 1;
    if (tma_f32_env && d->d_fp32 && !d->R && !d->bias_rows && bn % 32 == 0 && al16(d->D) && d->d_sm % 4 == 0 && d->N % 4 == 0) {
        // fp32 outputs (dense weight gradients, the conv-LoRA tap products) leave through TMA as [32 x 16] boxes; accumulate
        // mode uses the reduce-add form, so the SMs never read the old gradient values
        const long long dD[4] = {d->N, d->M, 1, 1}, sD[3] = {d->d_sm, 0, 0};
        const int boxD[4] = {16, 32, 1, 1};
        if (int rc = encode_map_ex(&g.mapD, d->D, 4, dD, sD, boxD, 64)) return rc;
        g.tma_store = 1;
    }
    if (d->geglu_y != nullptr) {
        B200_CHECK_ARG(g.tma_store && !d->d_fp32 && !d->side && !d->R && !d->bias_rows && bn == 256 && d->N % 256 == 0 &&
                           d->num_seg == 1 && d->geglu_h == nullptr && al16(d->geglu_y) && d->geglu_y_ld % 8 == 0 &&
                           d->geglu_y_ld >= d->N / 2,
                       "gemm2: fused GEGLU forward needs a plain bf16 TMA-store problem with 256-wide tiles and N %% 256 == 0");
        const long long dY[4] = {d->N / 2, d->M, 1, 1}, sY[3] = {d->geglu_y_ld, 0, 0};
        const int boxY[4] = {32, 32, 1, 1};
        if (int rc = encode_map_ex(&g.mapD2, d->geglu_y, 2, dY, sY, boxY, 64)) return rc;
        g.geglu_fwd = 1;
    }
    if (d->geglu_h != nullptr) {
        B200_CHECK_ARG(g.tma_store && !d->d_fp32 && !d->side && !d->R && !d->bias && d->N % 32 == 0 && d->num_seg == 1 &&
                           al16(d->geglu_h) && d->geglu_h_ld % 8 == 0 && d->geglu_h_ld >= 2LL * d->N && d->d_sm >= 2LL * d->N,
                       "gemm2: fused GEGLU backward needs a plain bf16 TMA-store problem with N %% 32 == 0 and [M, 2N] h / D");
        // the output tensor map spans both halves of dh
        const long long dD[4] = {2LL * d->N, d->M, 1, 1}, sD[3] = {d->d_sm, 0, 0};
        const int boxD[4] = {32, 32, 1, 1};
        if (int rc = encode_map_ex(&g.mapD, d->D, 2, dD, sD, boxD, 64)) return rc;
        g.geglu_bwd = 1;
        g.H = static_cast<const __nv_bfloat16*>(d->geglu_h);
        g.h_ld = d->geglu_h_ld;
    }
    int pairs = g.total_tiles < kNumSMs / 2 ? g.total_tiles : kNumSMs / 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // stream-K: few tiles, long K (FF down-projection and its input gradient, the deep convs): cut the (tile, k-block)
    // space into 74 equal ranges so every SM works; partial tiles are added to D by TMA reduce-add, so D is pre-set
    // here to the residual (or zero) and the kernel sees no residual.
    static const int streamk_env = getenv("B200_STREAMK") ? atoi(getenv("B200_STREAMK")) : 1;
    if (streamk_env && g.tma_store && !d->d_fp32 && !d->side && !g.geglu_bwd && !g.geglu_fwd && g.kblocks >= 32 && g.total_tiles * 4 <= (kNumSMs / 2) * 3 &&
        d->N % 32 == 0 && static_cast<long long>(g.total_tiles) * g.kblocks >= kNumSMs / 2) {
        const size_t row_bytes = static_cast<size_t>(d->N) * 2;
        cudaError_t e = cudaSuccess;
        // D starts at zero; a residual that lives elsewhere rides (like the bias) with the range that holds k-block 0 of a tile -
        // a memset instead of a device-to-device copy of the residual; R == D: D already holds it
        if (d->R != d->D || d->R == nullptr) e = cudaMemset2DAsync(d->D, static_cast<size_t>(d->d_sm) * 2, 0, row_bytes, d->M, st);
        if (e != cudaSuccess) return set_error(3, "gemm2: stream-K output initialisation: %s", cudaGetErrorString(e));
        g.streamk = 1;
        if (d->R == d->D) g.R = nullptr;
        pairs = kNumSMs / 2;
    }
    cudaError_t le;
    if (d->d_fp32) le = launch_pdl_cluster(gemm2_kernel<1>, dim3(2 * pairs), dim3(k2Threads), k2SmemBytes, st, 2u, g);
    else le = launch_pdl_cluster(gemm2_kernel<0>, dim3(2 * pairs), dim3(k2Threads), k2SmemBytes, st, 2u, g);
    if (le != cudaSuccess) return set_error(3, "gemm2 launch: %s", cudaGetErrorString(le));
    B200_CHECK_LAUNCH("gemm2_tcgen05");
    return 0;
}

extern "C" int b200_gemm(const b200_gemm_t* d, void* stream) {
    B200_CHECK_ARG(d != nullptr, "gemm: null descriptor");
    B200_CHECK_ARG(d->M >= 1 && d->N >= 1, "gemm: empty problem M=%d N=%d", d->M, d->N);
    B200_CHECK_ARG(d->num_seg == 1 || d->num_seg == 2, "gemm: num_seg must be 1 or 2");
    B200_CHECK_ARG(d->nb0 >= 1 && d->nb1 >= 1 && d->splits >= 1, "gemm: bad batch/split");
    B200_CHECK_ARG(d->D != nullptr, "gemm: null output");
    B200_CHECK_ARG(!(d->d_atomic && !d->d_fp32), "gemm: atomic accumulation needs an fp32 output");
    B200_CHECK_ARG(d->splits == 1 || (d->d_atomic && (d->num_seg == 1 || d->group) && !d->conv),
                   "gemm: split-K needs atomic fp32 output, one segment (or group mode), no conv");

    // CTA-pair kernel for the big plain projections (pair_mode: 0 auto, 1 force, -1 never)
    static const int pair_env = getenv("B200_GEMM2") ? atoi(getenv("B200_GEMM2")) : 1;
    if (d->pair_mode >= 0 && (pair_env || d->pair_mode > 0) && pair_eligible(d)) {
        int bn2 = d->block_n;
        if (d->geglu_y != nullptr) bn2 = 256;              // one tile = 128 value + 128 gate columns
        if (bn2 <= 0) bn2 = pick_pair_bn(d, nullptr);
        const bool bn_ok = bn2 >= 32 && bn2 % 16 == 0 && bn2 <= 256 &&
                           (!(d->B[0].mn_major || (d->side && d->B2.mn_major) || (d->num_seg == 2 && d->B[1].mn_major)) ||
                            bn2 % 128 == 0);
        if (bn_ok && (d->pair_mode > 0 || d->M * static_cast<long long>(d->N) >= 256LL * 512)) return launch_gemm2(d, bn2, stream);
        B200_CHECK_ARG(d->pair_mode <= 0, "gemm: pair_mode forced but block_n %d is not usable by the pair kernel", bn2);
    } else {
        B200_CHECK_ARG(d->pair_mode <= 0, "gemm: pair_mode forced on a problem the pair kernel does not cover");
    }

    B200_CHECK_ARG(d->geglu_h == nullptr && d->geglu_y == nullptr,
                   "gemm: the fused GEGLU epilogues exist only in the CTA-pair kernel (M >= 256, N >= 64, K >= 64)");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
        if (e != cudaSuccess) return set_error(3, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }

    if (d->group) {
        B200_CHECK_ARG(d->num_seg == 2 && !d->conv && !d->side && d->nb0 == 1 && d->nb1 == 1, "gemm: group mode needs two plain segments");
        B200_CHECK_ARG(d->d_fp32 && d->d_atomic && d->D2 != nullptr, "gemm: group mode needs fp32 atomic outputs D and D2");
        B200_CHECK_ARG(d->K[0] == d->K[1], "gemm: group mode needs equal reduction lengths");
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.group = d->group;
    g.D2 = d->D2;
    g.d2_sm = d->d2_sm;
    g.d2_sn = d->d2_sn;
    g.M = d->M;
    g.N = d->N;
    g.nb0 = d->nb0;
    g.nb1 = d->nb1;
    g.splits = d->splits;
    g.num_seg = d->num_seg;
    g.tiles_m = (d->M + kBM - 1) / kBM;
    int bn = d->block_n;
    if (d->side) {
        B200_CHECK_ARG(d->num_seg == 1 && !d->conv && d->nb0 == 1 && d->nb1 == 1 && d->splits == 1,
                       "gemm: the fused side path needs one segment, no conv, no batch, no split-K");
        B200_CHECK_ARG(d->side_r >= 1 && d->side_r <= 32, "gemm: side rank %d out of range", d->side_r);
    }
    if (bn <= 0) bn = pick_block_n(static_cast<long long>(g.tiles_m) * d->nb0 * d->nb1 * d->splits * (d->group ? 2 : 1), d->N,
                                   d->side ? kSideMaxBN : kMaxBN);
    if (d->side) B200_CHECK_ARG(bn <= kSideMaxBN, "gemm: block_n %d too wide for the side path", bn);
    B200_CHECK_ARG(bn % 16 == 0 && bn >= 16 && bn <= kMaxBN, "gemm: block_n %d must be a multiple of 16 in [16, 256]", bn);
    g.BN = bn;
    g.tiles_n = (d->N + bn - 1) / bn;

    for (int s = 0; s < d->num_seg; ++s) {
        const b200_operand_t& A = d->A[s];
        const b200_operand_t& B = d->B[s];
        g.a_mn[s] = A.mn_major;
        g.b_mn[s] = B.mn_major;
        g.a_batched[s] = A.batched;
        g.b_batched[s] = B.batched;
        if (s == 0 && d->conv) {
            B200_CHECK_ARG(!A.mn_major && !B.mn_major, "gemm: conv operands must be K-major");
            const int W = d->conv_W, H = d->conv_H, Nimg = d->conv_N, C = d->conv_C;
            B200_CHECK_ARG(W >= 1 && W <= 128 && 128 % W == 0, "gemm: conv width %d must divide 128", W);
            int bh = 128 / W;
            if (bh > H) bh = H;
            B200_CHECK_ARG(H % bh == 0 && 128 % (W * bh) == 0, "gemm: conv height %d unsupported for width %d", H, W);
            const int bimg = 128 / (W * bh);
            B200_CHECK_ARG(bimg == 1 || bh == H, "gemm: conv tile/image mismatch");
            B200_CHECK_ARG(static_cast<long long>(Nimg) * H * W == d->M, "gemm: conv M mismatch");
            B200_CHECK_ARG(C % 8 == 0, "gemm: conv channels %d not a multiple of 8", C);
            const long long dims[4] = {C, W, H, Nimg};
            const long long strides[3] = {C, static_cast<long long>(W) * C, static_cast<long long>(H) * W * C};
            const int box[4] = {64, W, bh, bimg};
            if (int rc = encode_map(&g.mapA[0], A.ptr, dims, strides, box)) return rc;
            g.conv = 1;
            g.conv_cblocks = (C + kBK - 1) / kBK;
            g.conv_W = W;
            g.conv_H = H;
            g.conv_bh = bh;
            g.b_tap_k = d->b_tap_k;
            g.b_tap_n = d->b_tap_n;
            g.kblocks[0] = 9 * g.conv_cblocks;
            // channel tails are zero-filled by TMA on the A side, so every block runs all four 16-wide MMA steps
            g.ktail16[0] = 4;
        } else {
            const int K = d->K[s];
            B200_CHECK_ARG(K >= 1, "gemm: K[%d] = %d", s, K);
            B200_CHECK_ARG(A.mn_major ? (A.rows >= K) : (A.inner >= K), "gemm: A[%d] smaller than K", s);
            if (int rc = encode_operand(&g.mapA[s], A, kBM, d->nb0, d->nb1)) return rc;
            g.kblocks[s] = (K + kBK - 1) / kBK;
            g.ktail16[s] = (K - (g.kblocks[s] - 1) * kBK + 15) / 16;
        }
        if (int rc = encode_operand(&g.mapB[s], B, bn, d->nb0, d->nb1)) return rc;
    }
    B200_CHECK_ARG(d->splits <= g.kblocks[0], "gemm: more splits (%d) than K blocks (%d)", d->splits, g.kblocks[0]);
    if (d->side) {
        g.side = 1;
        g.side_r = d->side_r;
        g.side_r16 = ((d->side_r + 15) / 16) * 16;
        g.side_mn = d->S.mn_major;
        g.b2_mn = d->B2.mn_major;
        g.side_alpha = d->side_alpha;
        g.T_out = static_cast<__nv_bfloat16*>(d->T_out);
        g.t_ld = d->t_ld;
        if (int rc = encode_operand(&g.mapS, d->S, g.side_r16, 1, 1)) return rc;
        if (int rc = encode_operand(&g.mapB2, d->B2, bn, 1, 1)) return rc;
    }

    g.dbg = g_gemm_dbg;
    static const int dbg_mode = getenv("B200_EPI_MODE") ? atoi(getenv("B200_EPI_MODE")) : 0;
    g.dbg_mode = dbg_mode;
    g.D = d->D;
    g.d_fp32 = d->d_fp32;
    g.d_atomic = d->d_atomic;
    g.d_sm = d->d_sm;
    g.d_sn = d->d_sn;
    g.d_sb0 = d->d_sb0;
    g.d_sb1 = d->d_sb1;
    g.alpha = d->alpha;
    g.bias = static_cast<const __nv_bfloat16*>(d->bias);
    g.bias_rows = d->bias_rows;
    g.bias_sb = d->bias_sb;
    g.R = static_cast<const __nv_bfloat16*>(d->R);
    g.r_sm = d->r_sm;
    g.r_sn = d->r_sn;
    g.r_sb0 = d->r_sb0;
    g.r_sb1 = d->r_sb1;

    // ring geometry: every slot holds the A tile, the widest B tile any segment / the B2 tile needs, and the side tile
    {
        int b_bytes = 0;
        for (int s = 0; s < d->num_seg; ++s) {
            const int bb = g.b_mn[s] ? ((bn + 63) / 64) * 8192 : bn * 128;
            b_bytes = bb > b_bytes ? bb : b_bytes;
        }
        if (g.side) {
            const int bb = g.b2_mn ? ((bn + 63) / 64) * 8192 : bn * 128;
            b_bytes = bb > b_bytes ? bb : b_bytes;
        }
        b_bytes = (b_bytes + 1023) / 1024 * 1024;
        const int side_bytes = g.side ? (g.side_mn ? 8192 : (g.side_r16 * 128 + 1023) / 1024 * 1024) : 0;
        g.side_off = b_bytes;
        g.stage_bytes = kABytes + b_bytes + side_bytes;
        g.num_stages = (kStages * kStageBytes) / g.stage_bytes;
        if (g.num_stages > kMaxStages) g.num_stages = kMaxStages;
        static const int cap = getenv("B200_GEMM_STAGES") ? atoi(getenv("B200_GEMM_STAGES")) : 0;   // tuning probe
        if (cap > 0 && g.num_stages > cap) g.num_stages = cap;
    }
    const long long total_tiles = static_cast<long long>(g.tiles_m) * g.tiles_n * g.splits * g.nb0 * g.nb1 * (g.group ? 2 : 1);
    const int grid = static_cast<int>(total_tiles < kNumSMs ? total_tiles : kNumSMs);
    const bool row_major = d->d_sn == 1 && !d->d_atomic;
    const int esz = d->d_fp32 ? 4 : 2;
    g.vec_ok = row_major && (reinterpret_cast<uintptr_t>(d->D) % (4 * esz) == 0) && (d->d_sm % 4 == 0) &&
               (d->d_sb0 % 4 == 0) && (d->d_sb1 % 4 == 0) &&
               (!d->R || (d->r_sn == 1 && reinterpret_cast<uintptr_t>(d->R) % 8 == 0 && d->r_sm % 4 == 0 &&
                          d->r_sb0 % 4 == 0 && d->r_sb1 % 4 == 0));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!row_major) launch_pdl(gemm_tcgen05_kernel<2>, dim3(grid), dim3(kGemmThreads), kGemmSmemBytes, st, g);
    else if (d->d_fp32) launch_pdl(gemm_tcgen05_kernel<1>, dim3(grid), dim3(kGemmThreads), kGemmSmemBytes, st, g);
    else launch_pdl(gemm_tcgen05_kernel<0>, dim3(grid), dim3(kGemmThreads), kGemmSmemBytes, st, g);
    B200_CHECK_LAUNCH("gemm_tcgen05");
    return 0;
}
