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
    // B200_FLASH_FWD_TS=0: the variant that stages P in shared memory (kept for A/B measurements)
    static const bool ts_mode = !(getenv("B200_FLASH_FWD_TS") && atoi(getenv("B200_FLASH_FWD_TS")) == 0);
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_fwd_ts_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtsSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_fwd_ts_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFtsSmem);
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
    // B200_FLASH_TIMELINE=1 (debug, synchronises): clock64 stamps of CTA (0,0,0) printed to stderr after every launch -
    // per key block: s_full seen, row max done, half-row maxima exchanged, exponentials done, P published | MMA: S(j+1) issued, P.V(j) issued
    static const bool timeline = kFaTimeline && getenv("B200_FLASH_TIMELINE") && atoi(getenv("B200_FLASH_TIMELINE")) != 0;
    static long long* tl_buf = nullptr;
    const int nkv_dbg = (Lk + 127) / 128;
    if (timeline && ts_mode && nkv_dbg <= 64) {
        if (!tl_buf) cudaMalloc(&tl_buf, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(tl_buf, 0, 64 * 8 * sizeof(long long), static_cast<cudaStream_t>(stream));
        g.timeline = tl_buf;
    }
    // B200_FLASH_POLY=0: every exponential on MUFU (default: every 4th on the FMA pipe, see flash_fwd_ts_kernel)
    static const bool poly = !(getenv("B200_FLASH_POLY") && atoi(getenv("B200_FLASH_POLY")) == 0);
    const cudaStream_t st_f = static_cast<cudaStream_t>(stream);
    if (ts_mode && poly) launch_pdl(flash_fwd_ts_kernel<4>, dim3(grid), dim3(kFaThreads), kFtsSmem, st_f, g);
    else if (ts_mode) launch_pdl(flash_fwd_ts_kernel<0>, dim3(grid), dim3(kFaThreads), kFtsSmem, st_f, g);
    else launch_pdl(flash_fwd_kernel, dim3(grid), dim3(kFaThreads), kFwdSmem, static_cast<cudaStream_t>(stream), g);
    B200_CHECK_LAUNCH("flash_fwd");
    if (g.timeline != nullptr) {
        long long hbuf[64 * 8];
        cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
        cudaMemcpy(hbuf, tl_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
        const long long t0 = hbuf[7];
        fprintf(stderr, "flash_fwd timeline (clk since first S issue) B=%d H=%d L=%d Lk=%d\n", B, H, L, Lk);
        for (int j = 0; j < nkv_dbg; ++j)
            fprintf(stderr, "  blk %2d: s_full %6lld  max %6lld  xchg %6lld  exp %6lld  p_pub %6lld | S(j+1) issue %6lld  PV(j) issue %6lld\n", j,
                    hbuf[j * 8 + 0] - t0, hbuf[j * 8 + 1] - t0, hbuf[j * 8 + 2] - t0, hbuf[j * 8 + 3] - t0, hbuf[j * 8 + 4] - t0,
                    hbuf[j * 8 + 5] ? hbuf[j * 8 + 5] - t0 : 0, hbuf[j * 8 + 6] - t0);
    }
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
    B200_CHECK_ARG(dq_direct || dq_acc_ws != nullptr, "flash_attn_bwd: Lk > 128 needs the fp32 dQ accumulator workspace");
    launch_pdl(flash_delta_kernel, dim3(grid_for(static_cast<long long>(B) * L * H * 8, 256)), dim3(256), 0, st,
        static_cast<const __nv_bfloat16*>(o), static_cast<const __nv_bfloat16*>(d_o), delta_ws,
        dq_direct ? static_cast<float*>(nullptr) : dq_acc_ws, B, L, H, ld_o);    // also zeroes the dQ accumulator
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
    // Query split (see FlashBwdArgs): (a) a grid smaller than the machine - every item is cut so that the pieces fill it;
    // (b) a grid of several rounds whose last round is partial - only that round's items are cut (the full rounds cost the
    // same for every SM, so the pieces start together and the step ends one short round later instead of one full round
    // later).
    const int nq = (L + 127) / 128, nkb = (Lk + 127) / 128;
    const long long items = static_cast<long long>(nkb) * H * B;
    g.nsplit = 1;
    g.n_unsplit = static_cast<int>(items);
    g.nkb = nkb;
    const long long plane = static_cast<long long>(B) * Lk * ld;
    // B200_FLASH_TAILSPLIT: most pieces an item of the partial round is cut into (0 / 1: never cut)
    static const int tail_env = getenv("B200_FLASH_TAILSPLIT") ? atoi(getenv("B200_FLASH_TAILSPLIT")) : 8;
    long long n_ctas = items;
    if (split_ws != nullptr && nq >= 2 && split_ws_floats >= 2 * plane + items) {
        const long long full = items / kNumSMs * kNumSMs, tail = items - full;
        int ns = 1;
        if (tail > 0 && (full == 0 || (tail_env >= 2 && !dq_direct))) {
            ns = static_cast<int>(kNumSMs / tail);
            if (ns > nq) ns = nq;
            if (full > 0) {
                if (ns > tail_env) ns = tail_env;
                // Measured (scripts/gpu_r2v.sh, gpu_r2x.sh): a CTA's fixed cost (prologue, K / V load, pipeline fill, atomic
                // epilogue, last-arriver rounding) is worth ~4 query blocks, so pieces shorter than 8 blocks do not pay:
                // L = 1024 (8 blocks per item) is left whole, L = 4096 (32 blocks, 3 pieces of ~11) gains 6.5 %.
                if (ns > nq / 8) ns = nq / 8;
            }
        }
        if (ns >= 2) {
            B200_CHECK_ARG(reinterpret_cast<uintptr_t>(split_ws) % 16 == 0, "flash_attn_bwd: split workspace not 16-byte aligned");
            g.nsplit = ns;
            g.n_unsplit = static_cast<int>(full);
            g.nbatch_rows = B * Lk;
            g.dKVacc = split_ws;
            g.counters = reinterpret_cast<int*>(split_ws + 2 * plane);
            n_ctas = full + tail * ns;
        }
    }
    dim3 grid(static_cast<unsigned>(n_ctas));
    static const bool timeline = kFaTimeline && getenv("B200_FLASH_TIMELINE") && atoi(getenv("B200_FLASH_TIMELINE")) != 0;
    static long long* tl_buf = nullptr;
    if (timeline && nq <= 64) {                    // debug, synchronises: see b200_flash_attn_fwd
        if (!tl_buf) cudaMalloc(&tl_buf, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(tl_buf, 0, 64 * 8 * sizeof(long long), st);
        g.timeline = tl_buf;
    }
    if (g.dSc != nullptr) launch_pdl(flash_bwd_kernel<true>, dim3(grid), dim3(kBwdThreads), kBwdSmem, st, g);
    else launch_pdl(flash_bwd_kernel<false>, dim3(grid), dim3(kBwdThreads), kBwdSmem, st, g);
    B200_CHECK_LAUNCH("flash_bwd");
    if (g.timeline != nullptr) {
        long long hbuf[64 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(hbuf, tl_buf, sizeof(hbuf), cudaMemcpyDeviceToHost);
        const long long t0 = hbuf[7];
        fprintf(stderr, "flash_bwd timeline (clk since first S/dP issue) B=%d H=%d L=%d Lk=%d\n", B, H, L, Lk);
        for (int i = 0; i < nq && i < 10; ++i)
            fprintf(stderr, "  qblk %2d: sdp_full %6lld  first half done %6lld  pds_free %6lld  P/dS published %6lld | S/dP(i+1) issued %6lld  grads(i) issued %6lld | dq_full %6lld\n",
                    i, hbuf[i * 8 + 0] - t0, hbuf[i * 8 + 1] - t0, hbuf[i * 8 + 2] - t0, hbuf[i * 8 + 3] - t0,
                    hbuf[i * 8 + 4] ? hbuf[i * 8 + 4] - t0 : 0, hbuf[i * 8 + 5] - t0, hbuf[i * 8 + 6] - t0);
    }
    if (!dq_direct) {
        launch_pdl(f32_to_bf16_kernel, dim3(grid_for(nq_elems / 4, 256)), dim3(256), 0, st, dq_acc_ws, static_cast<__nv_bfloat16*>(dq),
                   nq_elems / 4, static_cast<int>(C64 / 4), static_cast<long long>(ld));
        B200_CHECK_LAUNCH("flash_dq_convert");
    }
    return 0;
}
