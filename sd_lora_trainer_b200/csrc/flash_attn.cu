// Host side of the fused attention kernels (head_dim 64): tensor maps over the [B*L, C] projection outputs with the
// head as a strided batch dimension, launch configuration, and the small helper kernels of the backward.
#include <stdlib.h>
#include "common.cuh"
#include "flash_attn.cuh"
#include "../../include/b200_lora.h"

using namespace b200;

static int head_map(CUtensorMap* map, const void* ptr, int rows, int H, int B, long long ld, int box_rows) {
    const long long dims[4] = {64, rows, H, B};
    const long long strides[3] = {ld, 64, static_cast<long long>(rows) * ld};
    const int box[4] = {64, box_rows, 1, 1};
    return encode_map(map, ptr, dims, strides, box);
}

extern "C" int b200_flash_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B,
                                   int32_t H, int32_t L, int32_t Lk, int64_t ld, int64_t ld_o, float scale, void* stream) {
    B200_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && Lk >= 1 && ld >= static_cast<int64_t>(H) * 64 && ld % 8 == 0 &&
                       ld_o >= static_cast<int64_t>(H) * 64 && ld_o % 8 == 0,
                   "flash_attn_fwd: bad extents");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
        if (e != cudaSuccess) return set_error(3, "flash_attn_fwd: %s", cudaGetErrorString(e));
        attr = true;
    }
    FlashFwdArgs g;
    memset(&g, 0, sizeof(g));
    if (int rc = head_map(&g.mapQ, q, L, H, B, ld, 128)) return rc;
    if (int rc = head_map(&g.mapK, k, Lk, H, B, ld, 128)) return rc;
    if (int rc = head_map(&g.mapV, v, Lk, H, B, ld, 64)) return rc;
    g.O = static_cast<__nv_bfloat16*>(o);
    g.LSE = lse;
    g.L = L;
    g.Lk = Lk;
    g.H = H;
    g.o_ld = ld_o;
    g.scale = scale;
    dim3 grid((L + 127) / 128, H, B);
    launch_pdl(flash_fwd_kernel, dim3(grid), dim3(kFaThreads), kFwdSmem, static_cast<cudaStream_t>(stream), g);
    B200_CHECK_LAUNCH("flash_fwd");
    return 0;
}

extern "C" int b200_flash_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                   const float* lse, float* delta_ws, float* dq_acc_ws, void* dq, void* dk, void* dv,
                                   int32_t B, int32_t H, int32_t L, int32_t Lk, int64_t ld_qkv, int64_t ld_o, int64_t ld,
                                   float scale, float* split_ws, int64_t split_ws_floats, const void* dsc, int64_t ld_dsc,
                                   int32_t dsc_cols, void* stream) {
    // ld_qkv: row stride of q / k / v; ld_o: of o and d_o; ld: of the outputs dq / dk / dv (the fused q|k|v projection hands
    // column slices of one [rows, 3C] buffer in and takes the three gradients back the same way)
    const int64_t C64 = static_cast<int64_t>(H) * 64;
    B200_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && Lk >= 1 && ld_qkv >= C64 && ld_o >= C64 && ld >= C64 && ld_qkv % 8 == 0 &&
                       ld_o % 8 == 0 && ld % 8 == 0,
                   "flash_attn_bwd: row strides must be multiples of 8 elements and at least H*64");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
        if (e != cudaSuccess) return set_error(3, "flash_attn_bwd: %s", cudaGetErrorString(e));
        attr = true;
    }
    const long long nq_elems = static_cast<long long>(B) * L * C64;        // fp32 dQ accumulator: contiguous [B*L, H*64]
    const bool dq_direct = Lk <= 128;          // one key block: dQ is written once, as bf16, by the kernel itself
    if (!dq_direct) cudaMemsetAsync(dq_acc_ws, 0, sizeof(float) * nq_elems, st);
    launch_pdl(flash_delta_kernel, dim3(grid_for(static_cast<long long>(B) * L * H, 256)), dim3(256), 0, st, 
        static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), delta_ws, B, L, H, ld_o);
    B200_CHECK_LAUNCH("flash_delta");
    FlashBwdArgs g;
    memset(&g, 0, sizeof(g));
    if (int rc = head_map(&g.mapQ, q, L, H, B, ld_qkv, 128)) return rc;
    if (int rc = head_map(&g.mapDO, d_o, L, H, B, ld_o, 128)) return rc;
    if (int rc = head_map(&g.mapK, k, Lk, H, B, ld_qkv, 128)) return rc;
    if (int rc = head_map(&g.mapV, v, Lk, H, B, ld_qkv, 128)) return rc;
    if (!dq_direct) {
        B200_CHECK_ARG(dq_acc_ws != nullptr, "flash_attn_bwd: Lk > 128 needs the fp32 dQ accumulator workspace");
        const long long dims[4] = {C64, L, B, 1};
        const long long strides[3] = {C64, static_cast<long long>(L) * C64, 0};
        const int box[4] = {32, 128, 1, 1};
        if (int rc = encode_map_ex(&g.mapDQ, dq_acc_ws, 4, dims, strides, box, 128)) return rc;
    }
    g.LSE = lse;
    g.Delta = delta_ws;
    g.dQacc = dq_acc_ws;
    g.dQ = static_cast<__nv_bfloat16*>(dq);
    g.dq_direct = dq_direct ? 1 : 0;
    g.dK = static_cast<__nv_bfloat16*>(dk);
    g.dV = static_cast<__nv_bfloat16*>(dv);
    g.L = L;
    g.Lk = Lk;
    g.H = H;
    g.ld = ld;
    g.scale = scale;
    if (dsc != nullptr) {
        B200_CHECK_ARG(ld_dsc % 8 == 0 && dsc_cols % 8 == 0 && dsc_cols >= 0 && dsc_cols <= ld_dsc &&
                           reinterpret_cast<uintptr_t>(dsc) % 16 == 0,
                       "flash_attn_bwd: the score-gradient rows must be 16-byte aligned with a multiple of 8 readable columns");
        g.dSc = static_cast<const __nv_bfloat16*>(dsc);
        g.ld_dsc = ld_dsc;
        g.dsc_cols = dsc_cols;
    }
    dim3 grid((Lk + 127) / 128, H, B);
    // Query split: (a) single key block (cross-attention) on a grid that would leave SMs idle; (b) several key blocks whose
    // CTA count quantises badly over the 148 SMs (one CTA per SM): pick the split that minimises rounds x blocks per CTA,
    // charging each extra CTA its fixed cost (K / V load, prologue, atomic epilogue ~ 1.5 query blocks).
    g.nsplit = 1;
    const int nq = (L + 127) / 128, nkb = (Lk + 127) / 128;
    g.nkb = nkb;
    const long long plane = static_cast<long long>(B) * Lk * ld;
    // (b) is opt-in (B200_FLASH_QSPLIT=1): with a uniform split the extra CTAs' fixed cost eats the rounds it saves (the model
    // below picks 1 for SDXL's L = 1024 layers); a non-uniform, stream-K-style cut of the (key block, query block) grid is
    // what would pay
    static const int split_env = getenv("B200_FLASH_QSPLIT") ? atoi(getenv("B200_FLASH_QSPLIT")) : 0;
    if (split_ws != nullptr && nq >= 2 && split_ws_floats >= 2 * plane + static_cast<long long>(B) * H * nkb) {
        int best = 1;
        if (dq_direct && B * H < kNumSMs) {
            best = kNumSMs / (B * H);
            if (best > nq) best = nq;
        } else if (!dq_direct && split_env) {
            double best_cost = 1e30;
            for (int ns = 1; ns <= 4 && ns <= nq; ++ns) {
                const long long ctas = static_cast<long long>(nkb) * H * B * ns;
                const long long rounds = (ctas + kNumSMs - 1) / kNumSMs;
                const double per_cta = (nq + ns - 1) / ns + (ns > 1 ? 1.5 : 0.5);
                const double cost = rounds * per_cta;
                if (cost < best_cost - 1e-9) {
                    best_cost = cost;
                    best = ns;
                }
            }
        }
        if (best >= 2) {
            B200_CHECK_ARG(reinterpret_cast<uintptr_t>(split_ws) % 16 == 0, "flash_attn_bwd: split workspace not 16-byte aligned");
            g.nsplit = best;
            g.nbatch_rows = B * Lk;
            g.dKVacc = split_ws;
            g.counters = reinterpret_cast<int*>(split_ws + 2 * plane);
            grid.x = nkb * best;
        }
    }
    if (g.dSc != nullptr) launch_pdl(flash_bwd_kernel<true>, dim3(grid), dim3(kBwdThreads), kBwdSmem, st, g);
    else launch_pdl(flash_bwd_kernel<false>, dim3(grid), dim3(kBwdThreads), kBwdSmem, st, g);
    B200_CHECK_LAUNCH("flash_bwd");
    if (!dq_direct) {
        launch_pdl(f32_to_bf16_kernel, dim3(grid_for(nq_elems / 4, 256)), dim3(256), 0, st, dq_acc_ws, static_cast<__nv_bfloat16*>(dq),
                   nq_elems / 4, static_cast<int>(C64 / 4), static_cast<long long>(ld));
        B200_CHECK_LAUNCH("flash_dq_convert");
    }
    return 0;
}
