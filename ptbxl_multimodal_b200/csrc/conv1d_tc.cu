// bf16 tensor-core Conv1d(k=15, pad=7) as an implicit GEMM on tcgen05 / TMEM, operands staged by TMA.
//
// Replaces aten::convolution / convolution_backward (cuDNN) reached from nn.Conv1d at
// /root/reference/src/models/ecg_cnn.py:13 for the bf16 compute mode.
//
// Data layout (HBM): activations are "blocked channels-last" bf16  A[b][c/8][t][c%8]
// (16 bytes = 8 channels of one time step).  One TMA box {8 ch, 144 rows, C/8 chunks} lands in
// shared memory as [C/8][144][8] which IS the SWIZZLE_NONE core-matrix layout of tcgen05:
//   * as a K-major operand (K = channels) for forward / dgrad:  LBO = 144*16, SBO = 128
//   * as an MN-major operand (K = time) for wgrad:              SBO = 144*16, LBO = 128
// and because no swizzle is involved, tap k of the 15-tap stencil is the same tile viewed from
// start address + k*16 bytes: the input tile is loaded ONCE per output tile and reused by all taps.
// Zero padding at the sequence ends comes from TMA out-of-bounds fill (negative start row).
//
// GEMM per CTA (forward):  D[128 t x Co] = sum_{k<15} sum_{c} X[t+k-7, c] * W_k[c, o]
//   M = 128 time steps (TMEM lanes), N = Co (TMEM columns, fp32), K = 15 * Ci.
// Warp roles: warp 0 = TMA producer (input tile once, then a 4-stage ring of weight slabs),
// warp 1 = single-thread MMA issuer, warps 2..5 = epilogue (tcgen05.ld -> +bias -> bf16 -> HBM).
#include "tc_common.cuh"
#include <cstdlib>

// ---------------------------------------------------------------- host: tensor maps
ecg_tmap_encode_fn ecg_get_tmap_encode() {
    static ecg_tmap_encode_fn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (ecg_tmap_encode_fn)p;
    }
    return fn;
}

int ecg_make_act_tmap(CUtensorMap* m, const void* base, int B, int C, int L, int box_rows, int box_chunks) {
    ecg_tmap_encode_fn enc = ecg_get_tmap_encode();
    if (enc == nullptr) return ECGB200_EUNSUPPORTED;
    const cuuint64_t dims[4] = {8, (cuuint64_t)L, (cuuint64_t)(C / 8), (cuuint64_t)B};
    const cuuint64_t strides[3] = {16, (cuuint64_t)L * 16, (cuuint64_t)(C / 8) * L * 16};   // bytes, dims 1..3
    const cuuint32_t box[4] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_chunks, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ECGB200_EINVAL;
}

int ecg_make_act_tmap64(CUtensorMap* m, const void* base, int B, int C, int L, int box_rows, int box_chunks) {
    ecg_tmap_encode_fn enc = ecg_get_tmap_encode();
    if (enc == nullptr) return ECGB200_EUNSUPPORTED;
    if (box_rows <= 0 || box_rows > 128) return ECGB200_EINVAL;
    const cuuint64_t dims[3] = {(cuuint64_t)L * 2, (cuuint64_t)(C / 8), (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)L * 16, (cuuint64_t)(C / 8) * L * 16};         // bytes, dims 1..2
    const cuuint32_t box[3] = {(cuuint32_t)box_rows * 2, (cuuint32_t)box_chunks, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ECGB200_EINVAL;
}

// ---------------------------------------------------------------- layout conversion kernels
// x fp32 (B, Ci, T)  ->  xb bf16 [B][Cp/8][T][8], channels Ci..Cp-1 zero.   One thread per (b, chunk, t).
__global__ void pack_input_bf16_kernel(const float* __restrict__ x, uint4* __restrict__ xb,
                                       int B, int Ci, int Cp, int T) {
    const long long n = (long long)B * (Cp / 8) * T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const int cc = (int)((i / T) % (Cp / 8));
        const int b = (int)(i / ((long long)T * (Cp / 8)));
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cc * 8 + j;
            v[j] = c < Ci ? __ldg(x + ((size_t)b * Ci + c) * T + t) : 0.f;
        }
        xb[i] = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                           tc::pack_bf16(v[6], v[7]));
    }
}

// blocked bf16 [B][C/8][L][8] -> fp32 (B, C, L)   (debug / hooks / parity checks)
__global__ void unpack_act_bf16_kernel(const uint4* __restrict__ xb, float* __restrict__ x, int B, int C, int L) {
    const long long n = (long long)B * (C / 8) * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % L);
        const int cc = (int)((i / L) % (C / 8));
        const int b = (int)(i / ((long long)L * (C / 8)));
        const uint4 u = xb[i];
        const float2 a = tc::unpack_bf16(u.x), c = tc::unpack_bf16(u.y), d = tc::unpack_bf16(u.z), e = tc::unpack_bf16(u.w);
        float* o = x + ((size_t)b * C + cc * 8) * L + t;
        o[0] = a.x; o[(size_t)L] = a.y; o[(size_t)2 * L] = c.x; o[(size_t)3 * L] = c.y;
        o[(size_t)4 * L] = d.x; o[(size_t)5 * L] = d.y; o[(size_t)6 * L] = e.x; o[(size_t)7 * L] = e.y;
    }
}

// w fp32 (Co, Ci, 15) -> wf bf16 [15][Cip/8][Co][8]   (wf[k][c/8][o][c%8] = w[o][c][k], 0 for c >= Ci)
//                     -> wd bf16 [15][Co/8][Cip][8]   (wd[k][o/8][c][o%8] = w[o][c][14-k])   (may be NULL)
__global__ void prep_weights_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                         __nv_bfloat16* __restrict__ wd, int Co, int Ci, int Cip) {
    const int n = ECG_KS * Cip * Co;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // i indexes wf: [k][c/8][o][c%8]
        const int j = i & 7;
        const int o = (i >> 3) % Co;
        const int cc = (i / (8 * Co)) % (Cip / 8);
        const int k = i / (8 * Co * (Cip / 8));
        const int c = cc * 8 + j;
        const float v = c < Ci ? w[((size_t)o * Ci + c) * ECG_KS + k] : 0.f;
        wf[i] = __float2bfloat16(v);
        if (wd != nullptr)
            wd[(((size_t)(ECG_KS - 1 - k) * (Co / 8) + (o >> 3)) * Cip + c) * 8 + (o & 7)] = __float2bfloat16(v);
    }
}

extern "C" int ecgb200_pack_input_bf16(const float* x, void* xb, int B, int Ci, int T, void* stream) {
    if (!x || !xb || B <= 0 || Ci <= 0 || T <= 0) return ECGB200_EINVAL;
    const int Cp = (Ci + 15) / 16 * 16;
    const long long n = (long long)B * (Cp / 8) * T;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    pack_input_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (uint4*)xb, B, Ci, Cp, T);
    return ecg_launch_status();
}

extern "C" int ecgb200_unpack_act_bf16(const void* xb, float* x, int B, int C, int L, void* stream) {
    if (!x || !xb || B <= 0 || C <= 0 || (C & 7) || L <= 0) return ECGB200_EINVAL;
    const long long n = (long long)B * (C / 8) * L;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    unpack_act_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)xb, x, B, C, L);
    return ecg_launch_status();
}

extern "C" int ecgb200_conv1d_prep_weights_bf16(const float* w, void* wf, void* wd, int Co, int Ci, void* stream) {
    if (!w || !wf || Co <= 0 || Ci <= 0 || (Co & 7)) return ECGB200_EINVAL;
    const int Cip = (Ci + 15) / 16 * 16;
    const int n = ECG_KS * Cip * Co;
    prep_weights_bf16_kernel<<<ecg_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        w, (__nv_bfloat16*)wf, (__nv_bfloat16*)wd, Co, Ci, Cip);
    return ecg_launch_status();
}

// ---------------------------------------------------------------- forward / dgrad implicit GEMM
constexpr int TC_TILE_M = 128;      // output time steps per tile (TMEM lanes)
constexpr int TC_ROWS = 144;        // input rows staged: 128 + 14 halo, rounded to 8
constexpr int TC_HDR = 1024;        // barriers + TMEM slot
constexpr int C2_MAXST = 8;         // weight ring depth (max)
constexpr int C2_STATB = 8 * 2 * 256 * 4;   // per-epilogue-warp {sum, sumsq} x 256 channels

void ecg_set_timeout_conv(unsigned long long ns) { cudaMemcpyToSymbol(tc::g_mbar_timeout_ns, &ns, sizeof(ns)); }

extern "C" int ecgb200_debug_set_diag(unsigned long long* pinned_host) {
    cudaError_t e = cudaMemcpyToSymbol(tc::g_mbar_diag, &pinned_host, sizeof(pinned_host));
    return e == cudaSuccess ? 0 : (int)e;
}

// Debug timeline (clock64 stamps of CTA 0), enabled by ecgb200_debug_set_trace(ptr != NULL).
__device__ long long* g_conv_trace = nullptr;
#define CTR(id) do { if (trace != nullptr && blockIdx.x == 0) trace[id] = clock64(); } while (0)
extern "C" int ecgb200_debug_set_trace(long long* buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_conv_trace, &buf, sizeof(buf));
    return e == cudaSuccess ? 0 : (int)e;
}
// Per-CTA {first instruction, last instruction} %globaltimer stamps of the tcgen05 kernels (buf[2 * linear block id]).
__device__ unsigned long long* g_cta_span = nullptr;
extern "C" int ecgb200_debug_set_cta_span(unsigned long long* buf) {
    cudaError_t e = cudaMemcpyToSymbol(g_cta_span, &buf, sizeof(buf));
    return e == cudaSuccess ? 0 : (int)e;
}
#define CTA_SPAN(which) do { unsigned long long* sp_ = g_cta_span; if (sp_ != nullptr) { \
    sp_[2 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) + (which)] = tc::globaltimer_ns(); } } while (0)

struct Conv2Cfg {
    int Ci, Co, L, kch;              // kch = input channels per weight stage
    int R;                           // tiles per group (accumulators side by side, one weight pass)
    int AS;                          // accumulator stages (2 = epilogue overlaps the next group's MMAs)
    int NXB;                         // input-tile buffers (2 = next group's tiles stream in under the MMAs)
    int NST;                         // weight ring depth; 0 = whole weight tensor resident in shared memory
    int total_tiles, tiles_t, ngroups;
    uint32_t tmem_cols, xbytes_al;
    int wide;                        // input tiles through the 8-byte-element maps (per chunk: 2 KB + 256 B boxes)
    int Cn;                          // MODE 5: channels of the OUTPUT tensor (3 split planes of Co channels, padded)
};

// All MMAs of one weight stage: RC tiles x NJ K-steps, fully unrolled so that every descriptor is
// "uniform base + constant multiple of a uniform stride" (stays in the uniform datapath), issued four per
// asm statement with compile-time accumulate flags.  FIRST: this stage starts the accumulators (j == 0).
template <int NJ, int RC, bool FIRST>
__device__ __forceinline__ void conv_issue_stage(uint32_t acc0, uint64_t ad_k, uint64_t bd_w, uint32_t co,
                                                 uint32_t xal16, uint32_t bstep, uint32_t idesc, bool leader,
                                                 uint64_t* commit_bar) {
    constexpr int CNT = NJ * RC;
    uint32_t dd[CNT];
    uint64_t ad[CNT], bd[CNT];
#pragma unroll
    for (int r = 0; r < RC; ++r)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            dd[r * NJ + j] = acc0 + (uint32_t)r * co;
            ad[r * NJ + j] = ad_k + (uint64_t)((uint32_t)r * xal16 + (uint32_t)(j * 2 * TC_ROWS));
            bd[r * NJ + j] = bd_w + (uint64_t)((uint32_t)j * bstep);
        }
    (void)leader;
    constexpr int MASK = FIRST ? (NJ == 4 ? 0xE : (NJ == 2 ? 0xA : 0x0)) : 0xF;
#pragma unroll
    for (int c = 0; c + 4 <= CNT; c += 4) tc::mma_bf16_x4<MASK>(dd + c, ad + c, bd + c, idesc);
#pragma unroll
    for (int c = CNT & ~3; c < CNT; ++c) {
        if (!FIRST || (c % NJ) != 0) tc::mma_bf16_c<1>(dd[c], ad[c], bd[c], idesc);
        else tc::mma_bf16_c<0>(dd[c], ad[c], bd[c], idesc);
    }
    if (commit_bar != nullptr) tc::mma_commit(commit_bar);
}

// One whole group with the weight tensor resident in shared memory (small-channel layers: one channel
// group, no barrier inside): 15 taps fully unrolled, every descriptor = group base + compile-time multiple
// of a uniform stride.  The (NJ, RC) dispatch happens once per group, not per stage.
template <int NJ, int RC>
__device__ __forceinline__ void conv_issue_group_resident(uint32_t acc0, uint64_t ad_g, uint64_t bd_w, uint32_t co,
                                                          uint32_t xal16, uint32_t bstep, uint32_t stage16,
                                                          uint32_t idesc, bool leader) {
    conv_issue_stage<NJ, RC, true>(acc0, ad_g, bd_w, co, xal16, bstep, idesc, leader, nullptr);
#pragma unroll
    for (int k = 1; k < ECG_KS; ++k)
        conv_issue_stage<NJ, RC, false>(acc0, ad_g + (uint64_t)k, bd_w + (uint64_t)((uint32_t)k * stage16), co, xal16,
                                        bstep, idesc, leader, nullptr);
}

// One whole group with streamed weights: per stage, wait for the slab, issue RC x NJ MMAs, commit the slot.
// slot / phase are carried across groups.
template <int NJ, int RC>
__device__ __forceinline__ void conv_issue_group_stream(uint32_t acc0, uint64_t ad_g, uint64_t bd0, uint32_t co,
                                                        uint32_t xal16, uint32_t bstep, uint32_t stage16,
                                                        uint32_t gstep, int groups, uint32_t idesc, bool leader,
                                                        uint64_t* wfull, uint64_t* wempty, int nst, int& slot,
                                                        uint32_t& wphase) {
    bool first = true;
    for (int k = 0; k < ECG_KS; ++k) {
        uint64_t ad_k = ad_g + (uint64_t)k;                    // tap k = the same tile, k rows (16 B each) further
        for (int g = 0; g < groups; ++g, ad_k += gstep) {
            tc::mbar_wait(wfull + slot, wphase);
            tc::fence_after_sync();
            const uint64_t bd_w = bd0 + (uint64_t)((uint32_t)slot * stage16);
            if (first) conv_issue_stage<NJ, RC, true>(acc0, ad_k, bd_w, co, xal16, bstep, idesc, leader, wempty + slot);
            else conv_issue_stage<NJ, RC, false>(acc0, ad_k, bd_w, co, xal16, bstep, idesc, leader, wempty + slot);
            first = false;
            if (++slot == nst) { slot = 0; wphase ^= 1; }
        }
    }
}

// Persistent CTA (one per SM): loops over groups of R consecutive 128-step output tiles.
//   warp 0   TMA producer: input tiles (double buffered) + weights (resident once, or a ring of slabs)
//   warp 1   single-thread tcgen05.mma issuer, accumulators in TMEM (double buffered when they fit)
//   warps 2-9 epilogue: tcgen05.ld -> +bias -> bf16 -> HBM, and the per-channel {sum, sum of squares}
//            of the ROUNDED outputs for the train-mode BatchNorm that follows (one partial per CTA,
//            combined in a fixed order by the consumer => deterministic).  Two warps share each TMEM
//            lane quarter and split the (tile, 32-channel block) items between them: the epilogue is
//            bound by one warp's instruction latency, not by bandwidth.
//
// MODE 0 (training forward): the epilogue above.   MODE 3 (dgrad / plain conv): the same without the statistics.
// MODE 1 (inference): eval-mode BatchNorm folded into per-channel {scale, shift} (`bias` = scale), ReLU and
//        MaxPool1d(2) in the epilogue -- the two time steps of a pool pair are adjacent TMEM lanes = adjacent
//        threads, which swap half of their 32 channels with one shuffle each -- and only the POOLED bf16 rows go
//        to HBM (`y` = pooled output [B][Co/8][L/2][8]): the conv output itself never leaves the SM.
// MODE 5 (split-precision inference): as MODE 1, but the pooled fp32 value p is stored as THREE bf16 planes
//        [hi | lo | hi] with hi = bf16(p), lo = bf16(p - hi) (`y` = [B][Cn/8][L/2][8], plane p at channel p * Co):
//        the next conv, whose weights are laid out as [w_hi | w_hi | w_lo], then computes
//        x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation -- fp32-accurate to ~1e-5 (only lo*lo is dropped) at
//        three times the bf16 MMA count, on the same tcgen05 kernel.
// MODE 6 (split-precision Grad-CAM front end): the RAW conv output + bias in fp32, (B, Co, L) NCL (`stat_part` = the fp32
//        output): what the reference's forward hook on the 4th Conv1d captures, at fp32 accuracy from split inputs.
// MODE 2 (inference, last block): as MODE 1 but nothing is stored: the pooled rows are summed over time per
//        (tile, lane quarter) for AdaptiveAvgPool1d(1) (`stat_part` = gap_part[tile][4][Co], fixed order).
constexpr int C2_THREADS = 320;
template <int MODE>
__global__ void __launch_bounds__(C2_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap xmapA, const __grid_constant__ CUtensorMap xmapB,
               const __nv_bfloat16* __restrict__ wprep,
               const float* __restrict__ bias, const float* __restrict__ shift, __nv_bfloat16* __restrict__ y,
               float* __restrict__ stat_part, const Conv2Cfg P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem);           // [C2_MAXST]
    uint64_t* wempty = wfull + C2_MAXST;                            // [C2_MAXST]
    uint64_t* xfull = wempty + C2_MAXST;                            // [2]
    uint64_t* xempty = xfull + 2;                                   // [2]
    uint64_t* accfull = xempty + 2;                                 // [2]
    uint64_t* accempty = accfull + 2;                               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    float* statsh = reinterpret_cast<float*>(smem + TC_HDR);        // [4][2][256]
    uint8_t* xs = smem + TC_HDR + C2_STATB;
    uint8_t* wsm = xs + (size_t)P.NXB * P.R * P.xbytes_al;

    const int Ci = P.Ci, Co = P.Co, L = P.L, kch = P.kch, R = P.R;
    const uint32_t xbytes = (uint32_t)Ci * TC_ROWS * 2;
    const uint32_t stage_bytes = (uint32_t)kch * Co * 2;
    const int groups = Ci / kch;
    const int nstage = ECG_KS * groups;
    const bool resident = P.NST == 0;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // provably warp-uniform
    const int lane = threadIdx.x & 31;
    const int ngl = ((int)blockIdx.x < P.ngroups) ? (P.ngroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    long long* const trace = g_conv_trace;
    if (threadIdx.x == 0) { CTR(0); CTA_SPAN(0); }

    if (threadIdx.x == 0) {
        for (int i = 0; i < C2_MAXST; ++i) { tc::mbar_init(wfull + i, 1); tc::mbar_init(wempty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(xfull + i, 1); tc::mbar_init(xempty + i, 1);
            tc::mbar_init(accfull + i, 1); tc::mbar_init(accempty + i, 8);
        }
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&xmapA);
        tc::prefetch_tmap(&xmapB);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, P.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);        // uniform for the compiler too
    ecg_pdl_wait();                                    // everything above overlapped the previous kernel's tail
    if (threadIdx.x == 0) CTR(1);

    if (warp == 0) {
        if (lane == 0) {
            auto load_x = [&](int gi) {
                const int xb = gi % P.NXB;
                if (gi >= P.NXB) tc::mbar_wait(xempty + xb, ((gi / P.NXB) - 1) & 1);
                const int tile0 = ((int)blockIdx.x + gi * (int)gridDim.x) * R;
                const int rcount = min(R, P.total_tiles - tile0);
                tc::mbar_arrive_expect_tx(xfull + xb, xbytes * (uint32_t)rcount);
                for (int r = 0; r < rcount; ++r) {
                    const int tile = tile0 + r;
                    const int b = tile / P.tiles_t, t0 = (tile - b * P.tiles_t) * TC_TILE_M;
                    uint8_t* dst = xs + (size_t)(xb * R + r) * P.xbytes_al;
                    if (P.wide) {
                        // per 8-channel chunk: rows [t0-7, t0+121) as one 2 KB box, rows [t0+121, t0+137) as a 256 B box
                        for (int c = 0; c < Ci / 8; ++c, dst += TC_ROWS * 16) {
                            tc::tma_load_3d(dst, &xmapA, xfull + xb, 2 * (t0 - ECG_PAD), c, b);
                            tc::tma_load_3d(dst + 128 * 16, &xmapB, xfull + xb, 2 * (t0 - ECG_PAD + 128), c, b);
                        }
                    } else {
                        tc::tma_load_4d(dst, &xmapA, xfull + xb, 0, t0 - ECG_PAD, 0, b);       // one box {8, 144, Ci/8, 1}
                    }
                }
                if (gi < 4) CTR(8 + gi);                         // x load of group gi issued
            };
            // wide input tiles (2 x Ci/8 TMA instructions each) are issued by the epilogue warps, eight threads in parallel
            const bool xhere = !P.wide;
            if (ngl > 0 && xhere) load_x(0);
            if (resident) {
                const uint32_t wbytes = (uint32_t)nstage * stage_bytes;
                tc::mbar_arrive_expect_tx(wfull, wbytes);
                for (uint32_t off = 0; off < wbytes; off += 16384u) {
                    const uint32_t n = wbytes - off < 16384u ? wbytes - off : 16384u;
                    tc::bulk_load(wsm + off, reinterpret_cast<const uint8_t*>(wprep) + off, n, wfull);
                }
                if (xhere)
                    for (int gi = 1; gi < ngl; ++gi) load_x(gi);
            } else {
                int slot = 0;
                uint32_t ephase = 1;                             // first pass over the ring: slots start free
                for (int gi = 0; gi < ngl; ++gi) {
                    // the next group's input tiles are requested once this group's MMAs are under way
                    // (single input buffer: only after ALL of this group's weight stages are on their way -- the buffer is
                    // released by the group's last MMA, which needs those stages; asking earlier would deadlock the ring)
                    const int xat = P.NXB == 1 ? nstage - 1 : (nstage - 1 < P.NST ? nstage - 1 : P.NST);
                    for (int s = 0; s < nstage; ++s) {
                        if (!(gi == 0 && s < P.NST)) tc::mbar_wait(wempty + slot, ephase);
                        tc::mbar_arrive_expect_tx(wfull + slot, stage_bytes);
                        tc::bulk_load(wsm + (size_t)slot * stage_bytes,
                                      reinterpret_cast<const uint8_t*>(wprep) + (size_t)s * stage_bytes, stage_bytes,
                                      wfull + slot);
                        if (s == xat && gi + 1 < ngl && xhere) load_x(gi + 1);
                        if (++slot == P.NST) { slot = 0; ephase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: ONE elected thread runs everything (waits, tcgen05.mma, tcgen05.commit) inside a single
        // divergent region.  Two measured facts shape this code:
        //  * the issue loop is bound by the issuing thread's own instruction latency, not by the tensor pipe:
        //    130-420 cycles per MMA when descriptors / accumulate predicates are recomputed per instruction,
        //    against 47 (N<=64) / 64 (N=128) / 128 (N=256) cycles of tensor time.  So the (NJ, RC) structure
        //    is dispatched once per group into fully unrolled code whose descriptors are "base + constant",
        //    accumulate flags are compile-time, and MMAs leave four per asm statement;
        //  * tcgen05.commit must NOT sit in a small `if (leader)` of warp-converged code: ptxas 12.9 turns it
        //    into an unguarded warp-level UTCBAR fed by R2UR.BROADCAST, and the ring then deadlocked
        //    sporadically.  Hence no warp-uniform tricks here: everything below is single-thread code.
        const bool leader = tc::elect_one();
        if (leader) {
        const uint32_t idesc = tc::make_idesc_bf16(TC_TILE_M, Co, 0, 0);
        const uint64_t adesc0 = tc::make_desc(0, TC_ROWS * 16, 128);
        const uint64_t bdesc0 = tc::make_desc(0, (uint32_t)Co * 16, 128);
        const uint64_t alo0 = adesc0 + (uint64_t)(tc::smem_u32(xs) >> 4);        // address field: low 14 bits
        const uint64_t blo0 = bdesc0 + (uint64_t)(tc::smem_u32(wsm) >> 4);
        const uint32_t xal16 = P.xbytes_al >> 4, stage16 = stage_bytes >> 4;
        const uint32_t bstep = (uint32_t)(2 * Co);             // two 8-channel chunks of the weight slab
        const uint32_t gstep = (uint32_t)(kch / 8) * TC_ROWS;  // one channel group of the input tile
        const int nj = kch / 16;                               // 1, 2 or 4 K-steps per weight stage
        if (resident && ngl > 0) {
            tc::mbar_wait(wfull, 0);
            tc::fence_after_sync();
        }
        CTR(3);                                                // resident weights landed
        int slot = 0;
        uint32_t wphase = 0;
        for (int gi = 0; gi < ngl; ++gi) {
            const int xb = gi % P.NXB, as = gi % P.AS;
            const int tile0 = ((int)blockIdx.x + gi * (int)gridDim.x) * R;
            const int rcount = min(R, P.total_tiles - tile0);
            tc::mbar_wait(xfull + xb, (gi / P.NXB) & 1);
            if (gi < 4) CTR(16 + gi);                          // x tiles of group gi landed
            if (gi >= P.AS) tc::mbar_wait(accempty + as, ((gi / P.AS) - 1) & 1);
            if (gi < 4) CTR(24 + gi);                          // accumulator stage free
            tc::fence_after_sync();
            const uint32_t acc0 = tmem_base + (uint32_t)(as * R * Co);
            const uint64_t alo_g = alo0 + (uint64_t)((uint32_t)(xb * R) * xal16);
            if (resident) {
#define ECG_RES(NJ_, RC_) conv_issue_group_resident<NJ_, RC_>(acc0, alo_g, blo0, (uint32_t)Co, xal16, bstep, stage16, idesc, leader)
                if (nj == 1) { if (rcount == 4) ECG_RES(1, 4); else if (rcount == 3) ECG_RES(1, 3); else if (rcount == 2) ECG_RES(1, 2); else ECG_RES(1, 1); }
                else if (nj == 2) { if (rcount == 4) ECG_RES(2, 4); else if (rcount == 3) ECG_RES(2, 3); else if (rcount == 2) ECG_RES(2, 2); else ECG_RES(2, 1); }
                else { if (rcount == 4) ECG_RES(4, 4); else if (rcount == 3) ECG_RES(4, 3); else if (rcount == 2) ECG_RES(4, 2); else ECG_RES(4, 1); }
#undef ECG_RES
            } else {
#define ECG_STR(NJ_, RC_) conv_issue_group_stream<NJ_, RC_>(acc0, alo_g, blo0, (uint32_t)Co, xal16, bstep, stage16, gstep, groups, idesc, leader, wfull, wempty, P.NST, slot, wphase)
                if (nj == 4) { if (rcount == 4) ECG_STR(4, 4); else if (rcount == 3) ECG_STR(4, 3); else if (rcount == 2) ECG_STR(4, 2); else ECG_STR(4, 1); }
                else if (nj == 2) { if (rcount == 2) ECG_STR(2, 2); else ECG_STR(2, 1); }
                else { if (rcount == 2) ECG_STR(1, 2); else ECG_STR(1, 1); }
#undef ECG_STR
            }
            tc::mma_commit(xempty + xb);               // input tiles of this group are free again
            tc::mma_commit(accfull + as);              // accumulators of this group are complete
            if (gi < 4) CTR(32 + gi);                  // all MMAs of group gi issued
        }
        // programmatic dependent launch, triggered LATE: only this CTA's last epilogue is still to run, so the next
        // kernel's launch latency and prologue hide under the tail without its CTAs squatting on SM slots earlier
        // (triggering at the top of the kernel measured slower: the early dependents compete with the weight-gradient branch)
        ecg_pdl_launch_dependents();
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // which of the two warps of this quarter
        const int nblk = Co >> 5;
        const int row = 32 * q + lane;
        const size_t chunk_stride = (size_t)L * 8;     // elements between channel chunks
        const bool want_stats = MODE == 0 && stat_part != nullptr;
        float ssum[8], ssq[8];                         // lane j: channel 32*cb + j, over this warp's rows
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        // Column sums are first accumulated per thread (same row index of the R tiles of a group, 32 channels),
        // and only then reduced across the 32 rows by a transposing butterfly (31 shuffles leave lane j with
        // column j): one butterfly pair per (group, channel block) instead of per tile -- and for layers with
        // <= 64 channels, where a warp always owns the same block, ONE pair for the whole kernel.
        const bool small = nblk <= 2;
        float vs[32], qs[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { vs[i] = 0.f; qs[i] = 0.f; }
        auto butterfly = [&](int cb) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                    const float send = up ? vs[i] : vs[i + o], keep = up ? vs[i + o] : vs[i];
                    vs[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    const float send2 = up ? qs[i] : qs[i + o], keep2 = up ? qs[i + o] : qs[i];
                    qs[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, o);
                }
            }
            ssum[cb] += vs[0];
            ssq[cb] += qs[0];
        };
        // Wide input tiles (>= 128 input channels: per 8-channel chunk a 2 KB box for rows [t0-7, t0+121) and a 256 B box for
        // rows [t0+121, t0+137)): ONE thread gets a TMA instruction out only every ~50 cycles (operands through R2UR), 3-6 k
        // cycles per group.  Lane 0 of each of the eight epilogue warps takes every eighth chunk instead: the first NXB groups at
        // the start of the kernel, group g + NXB at the start of the epilogue of group g (its accumulators being complete means
        // that the MMAs have finished reading that input buffer).
        auto load_tiles = [&](int gi) {
            const int xb = gi % P.NXB;
            const int tile0 = ((int)blockIdx.x + gi * (int)gridDim.x) * R;
            const int rcount = min(R, P.total_tiles - tile0);
            if (warp == 2) tc::mbar_arrive_expect_tx(xfull + xb, xbytes * (uint32_t)rcount);
            for (int r = 0; r < rcount; ++r) {
                const int tile = tile0 + r;
                const int b = tile / P.tiles_t, t2 = 2 * ((tile - b * P.tiles_t) * TC_TILE_M - ECG_PAD);
                uint8_t* dst = xs + (size_t)(xb * R + r) * P.xbytes_al + (size_t)(warp - 2) * (TC_ROWS * 16);
                for (int c = warp - 2; c < Ci / 8; c += 8, dst += 8 * TC_ROWS * 16) {
                    tc::tma_load_3d(dst, &xmapA, xfull + xb, t2, c, b);
                    tc::tma_load_3d(dst + 128 * 16, &xmapB, xfull + xb, t2 + 256, c, b);
                }
            }
        };
        if (P.wide && lane == 0)
            for (int gi = 0; gi < P.NXB && gi < ngl; ++gi) load_tiles(gi);
        __syncwarp();
        for (int gi = 0; gi < ngl; ++gi) {
            const int as = gi % P.AS;
            const int tile0 = ((int)blockIdx.x + gi * (int)gridDim.x) * R;
            const int rcount = min(R, P.total_tiles - tile0);
            tc::mbar_wait(accfull + as, (gi / P.AS) & 1);
            if (gi < 4 && threadIdx.x == 64) CTR(40 + gi);       // accumulators of group gi complete
            tc::fence_after_sync();
            if (P.wide && lane == 0 && gi + P.NXB < ngl) load_tiles(gi + P.NXB);
            __syncwarp();
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                // work split between the two warps of a lane quarter: by channel block, or by tile when there is one block
                if (cb < nblk && (nblk == 1 || (cb & 1) == half)) {
                    const int c0 = cb * 32;
                    if (!small) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { vs[i] = 0.f; qs[i] = 0.f; }
                    }
                    float bv[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) bv[i] = bias != nullptr ? __ldg(bias + c0 + i) : 0.f;
                    float sh[(MODE == 1 || MODE == 2 || MODE == 5) ? 32 : 1];
                    if constexpr (MODE == 1 || MODE == 2 || MODE == 5) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sh[i] = __ldg(shift + c0 + i);
                    }
                    for (int r = 0; r < rcount; ++r) {
                        if (nblk == 1 && (r & 1) != half) continue;
                        const int tile = tile0 + r;
                        const int b = tile / P.tiles_t, t = (tile - b * P.tiles_t) * TC_TILE_M + row;
                        const bool live = t < L;
                        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((as * R + r) * Co + c0);
                        __nv_bfloat16* yrow = y + ((size_t)b * (Co / 8) * L + t) * 8;
                        float v[32];
                        tc::tmem_ld32(taddr, v);
                        tc::tmem_ld_wait();
                        if constexpr (MODE == 6) {
                            // lane = time step: for a fixed channel the warp's 32 lanes are 128 contiguous bytes of (b, c, :)
                            if (live) {
                                float* arow = stat_part + ((size_t)b * Co + c0) * L + t;
#pragma unroll
                                for (int i = 0; i < 32; ++i) arow[(size_t)i * L] = v[i] + bv[i];
                            }
                            continue;
                        }
                        if constexpr (MODE == 1 || MODE == 2 || MODE == 5) {
                            // relu(scale * conv + shift), then max over the pool pair (lanes 2p, 2p+1): the even lane
                            // keeps channels 0-15 of the block, the odd lane 16-31
                            const bool even = (lane & 1) == 0;
                            const int Lp = L >> 1, tp = t >> 1;
                            const bool plive = tp < Lp;
                            float m[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float a = fmaxf(fmaf(v[i], bv[i], sh[i]), 0.f);
                                const float c = fmaxf(fmaf(v[16 + i], bv[16 + i], sh[16 + i]), 0.f);
                                const float send = even ? c : a, keep = even ? a : c;
                                m[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
                            }
                            if constexpr (MODE == 5) {
                                if (plive) {
                                    uint32_t hi[8], lo[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        hi[j] = tc::pack_bf16(m[2 * j], m[2 * j + 1]);
                                        const float2 h = tc::unpack_bf16(hi[j]);
                                        lo[j] = tc::pack_bf16(m[2 * j] - h.x, m[2 * j + 1] - h.y);
                                    }
                                    const size_t cs = (size_t)Lp * 8;                    // elements between chunks
                                    __nv_bfloat16* prow = y + ((size_t)b * (P.Cn / 8) * Lp + tp) * 8 +
                                                          (size_t)(c0 / 8 + (even ? 0 : 2)) * cs;
                                    const uint4 h0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), h1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                                    const size_t plane = (size_t)(Co / 8) * cs;
                                    *reinterpret_cast<uint4*>(prow) = h0;
                                    *reinterpret_cast<uint4*>(prow + cs) = h1;
                                    *reinterpret_cast<uint4*>(prow + plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                                    *reinterpret_cast<uint4*>(prow + plane + cs) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                                    *reinterpret_cast<uint4*>(prow + 2 * plane) = h0;
                                    *reinterpret_cast<uint4*>(prow + 2 * plane + cs) = h1;
                                }
                            } else if constexpr (MODE == 1) {
                                if (plive) {
                                    __nv_bfloat16* prow = y + ((size_t)b * (Co / 8) * Lp + tp) * 8 +
                                                          (size_t)(c0 / 8 + (even ? 0 : 2)) * ((size_t)Lp * 8);
                                    *reinterpret_cast<uint4*>(prow) =
                                        make_uint4(tc::pack_bf16(m[0], m[1]), tc::pack_bf16(m[2], m[3]),
                                                   tc::pack_bf16(m[4], m[5]), tc::pack_bf16(m[6], m[7]));
                                    *reinterpret_cast<uint4*>(prow + (size_t)Lp * 8) =
                                        make_uint4(tc::pack_bf16(m[8], m[9]), tc::pack_bf16(m[10], m[11]),
                                                   tc::pack_bf16(m[12], m[13]), tc::pack_bf16(m[14], m[15]));
                                }
                            } else {
                                // sum over the 16 pool pairs of this warp: transposing butterfly over lane bits 4..1,
                                // after which lane l holds channel c0 + 16*(l&1) + (l>>1)
#pragma unroll
                                for (int i = 0; i < 16; ++i) m[i] = plive ? m[i] : 0.f;
#pragma unroll
                                for (int o = 8; o > 0; o >>= 1) {
                                    const bool up = (lane & (2 * o)) != 0;
#pragma unroll
                                    for (int i = 0; i < o; ++i) {
                                        const float send = up ? m[i] : m[i + o], keep = up ? m[i + o] : m[i];
                                        m[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * o);
                                    }
                                }
                                stat_part[((size_t)tile * 4 + q) * Co + c0 + 16 * (lane & 1) + (lane >> 1)] = m[0];
                            }
                            continue;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t pk[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = 8 * i + 2 * j;
                                pk[j] = tc::pack_bf16(v[c] + bv[c], v[c + 1] + bv[c + 1]);
                                if constexpr (MODE == 0) {
                                    const float2 rr = tc::unpack_bf16(pk[j]);  // the value the next kernels will read
                                    if (live) {
                                        vs[c] += rr.x; qs[c] = fmaf(rr.x, rr.x, qs[c]);
                                        vs[c + 1] += rr.y; qs[c + 1] = fmaf(rr.y, rr.y, qs[c + 1]);
                                    }
                                }
                            }
                            if (live)
                                *reinterpret_cast<uint4*>(yrow + (size_t)(c0 / 8 + i) * chunk_stride) =
                                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    if (want_stats && !small) butterfly(cb);
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(accempty + as);
            if (gi < 4 && threadIdx.x == 64) CTR(48 + gi);       // epilogue of group gi done
        }
        if (want_stats && small) {
            // this warp's only channel block: `half` when there are two blocks, block 0 otherwise
            if (nblk == 2 && half == 1) butterfly(1); else butterfly(0);
        }
        if (want_stats) {
            float* mine = statsh + (size_t)(half * 4 + q) * 2 * 256;
#pragma unroll
            for (int cb = 0; cb < 8; ++cb)
                if (cb < nblk) {
                    mine[cb * 32 + lane] = ssum[cb];
                    mine[256 + cb * 32 + lane] = ssq[cb];
                }
            asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight epilogue warps only
            const int e = threadIdx.x - 64;
            for (int c = e; c < Co; c += 256) {
                float s = 0.f, s2 = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) { s += statsh[(w * 2 + 0) * 256 + c]; s2 += statsh[(w * 2 + 1) * 256 + c]; }
                stat_part[((size_t)blockIdx.x * 2 + 0) * Co + c] = s;
                stat_part[((size_t)blockIdx.x * 2 + 1) * Co + c] = s2;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) CTR(2);
    if (warp == 2) {
        tc::tmem_dealloc(tmem_base, P.tmem_cols);
        if (lane == 0) CTA_SPAN(1);
    }
}

static uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

// SM count of the CURRENT device (cached per device index: one process may drive several GPUs)
static int ecg_num_sms() {
    static int cache[64] = {0};
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 148; }
    if (dev >= 0 && dev < 64 && cache[dev] > 0) return cache[dev];
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
        (void)cudaGetLastError();
        return 148;
    }
    if (dev >= 0 && dev < 64) cache[dev] = v;
    return v;
}

// Shape -> schedule.  Returns the grid size (number of persistent CTAs = number of stat partials).
static int conv2_cfg(int B, int Ci, int Co, int L, Conv2Cfg* P, size_t* smem_out) {
    const int nsm = ecg_num_sms();
    P->Ci = Ci; P->Co = Co; P->L = L;
    P->kch = Ci < 64 ? Ci : 64;
    P->tiles_t = ecg_cdiv(L, TC_TILE_M);
    P->total_tiles = B * P->tiles_t;
    P->xbytes_al = ((uint32_t)Ci * TC_ROWS * 2 + 1023u) & ~1023u;
    const size_t wbytes = (size_t)ECG_KS * Ci * Co * 2;
    const size_t stage = (size_t)P->kch * Co * 2;
    const bool resident = wbytes <= 64 * 1024;
    const size_t budget = 225 * 1024 - TC_HDR - C2_STATB;
    int bestR = 1, bestSpan = 1 << 30;
    bool found = false, single = false;
    const int rmax = Co >= 256 ? 2 : (512 / (2 * Co) < 4 ? 512 / (2 * Co) : 4);
    // second pass (single = true) only when no R fits with double-buffered input tiles: the 3x-wide inputs of the
    // split-precision (fp32-accurate) inference layers keep ONE input buffer (the next group's tiles load after this
    // group's MMAs have consumed it)
    for (int pass = 0; pass < 2 && !found; ++pass)
    for (int R = rmax; R >= 1; R >>= 1) {
        const int ng = ecg_cdiv(P->total_tiles, R);
        const int grid = ng < nsm ? ng : nsm;
        const int span = ecg_cdiv(ng, grid) * R;                     // tiles on the busiest CTA
        const int nxb = (ecg_cdiv(ng, grid) > 1 && pass == 0) ? 2 : 1;
        const size_t need = (size_t)nxb * R * P->xbytes_al + (resident ? wbytes : 3 * stage);
        if (need > budget) continue;
        found = true;
        single = pass == 1;
        // streamed weights: fewer, larger groups halve the L2->SM weight traffic, so prefer the larger R
        if (span < bestSpan || (!resident && span == bestSpan && R > bestR)) { bestSpan = span; bestR = R; }
        // streamed weights: keep the largest R unless a smaller one shortens the busiest CTA (small batches: with
        // fewer groups than SMs, one tile per CTA halves the serial chain); small-channel layers also need
        // R * (kch/16) >= 4 MMAs per issue batch
        if (resident ? R * (P->kch / 16) <= 4 : ng >= nsm) break;
    }
    if (bestR * Co > 512) bestR = 512 / Co;
    P->R = bestR;
    P->AS = 2 * bestR * Co <= 512 ? 2 : 1;
    P->ngroups = ecg_cdiv(P->total_tiles, bestR);
    const int grid = P->ngroups < nsm ? P->ngroups : nsm;
    if (!found) return -1;
    P->NXB = (ecg_cdiv(P->ngroups, grid) > 1 && !single) ? 2 : 1;
    const size_t xall = (size_t)P->NXB * bestR * P->xbytes_al;
    if (resident) {
        P->NST = 0;
        *smem_out = TC_HDR + C2_STATB + xall + wbytes;
    } else {
        int nst = (int)((budget - xall) / stage);
        if (nst > C2_MAXST) nst = C2_MAXST;
        if (nst < 2) return -1;
        P->NST = nst;
        *smem_out = TC_HDR + C2_STATB + xall + (size_t)nst * stage;
    }
    P->tmem_cols = tmem_cols_for(P->AS * bestR * Co);
    return grid;
}

#include "conv1d_tc2.cuh"

// A/B switch for the two-SM (cta_group::2) kernels of the wide layers: bit 0 = forward / dgrad, bit 1 = wgrad,
// bit 2 = also for shapes where the pair form does not pay (tests); default 3.
// Not for use between ecgb200_conv1d_stat_parts_bf16 and the launch it sizes.
extern "C" int ecgb200_debug_set_conv_pair(int mask) {
    g_conv_pair = mask & 7;
    return 0;
}

extern "C" int ecgb200_conv1d_stat_parts_bf16(int B, int Ci, int Co, int L) {
    Conv2Cfg P;
    size_t smem;
    if (B > 0 && L > 0) {
        const int gp = conv_pair_cfg(B, Ci, Co, L, &P, &smem);
        if (gp > 0) return gp;
    }
    if (B <= 0 || L <= 0 || Ci <= 0 || (Ci & 15) || Ci > 384 || (Ci > 64 && (Ci & 63)) || Co <= 0 || (Co & 31) || Co > 256) return 0;
    const int g = conv2_cfg(B, Ci, Co, L, &P, &smem);
    return g > 0 ? g : 0;
}

// xb [B][Ci/8][L][8] bf16 (Ci % 16 == 0), wprep [15][Ci/8][Co][8] bf16, bias fp32 (Co) or NULL,
// yb [B][Co/8][L][8] bf16.  Co % 32 == 0, Co <= 256, Ci <= 256.
// stat_part: NULL or float[parts][2][Co] (parts = ecgb200_conv1d_stat_parts_bf16) receiving per-CTA
// {sum, sum of squares} of the bf16-rounded outputs.
template <int MODE>
static int conv_tc_launch(const void* xb, const void* wprep, const float* bias, const float* shift, void* yb,
                          float* stat_part, int B, int Ci, int Co, int L, void* stream, int Cn = 0) {
    if (Ci <= 0 || (Ci & 15) || Ci > 384 || (Ci > 64 && (Ci & 63)) || Co <= 0 || (Co & 31) || Co > 256) return ECGB200_EUNSUPPORTED;
    if constexpr (MODE == 0 || MODE == 3 || MODE == 1 || MODE == 2) {
        // streamed-weight layers (blocks 3 and 4: training forward, dgrad, inference): the two-SM kernel
        Conv2Cfg PP;
        size_t smem2;
        const int gp = conv_pair_cfg(B, Ci, Co, L, &PP, &smem2);
        if (gp > 0)
            return conv_tc_pair_launch<MODE>(xb, wprep, bias, shift, yb, stat_part, B, Ci, Co, L, PP, gp, smem2, stream);
    }
    // Input tiles: one 4-D box of 16-byte rows per tile for the 16-channel stem (the copy engine moves ~one 16-byte row per
    // cycle); from 32 input channels on, two wide-row boxes per 8-channel chunk, issued by the eight epilogue threads in
    // parallel (kernel time at B=256, wide + parallel issue against the one-box form: 32->64 17.5 -> 16.6 us, 64->32 dgrad
    // 21.3 -> 20.0 us, 256->128 dgrad 29.7 -> 27.1 us; with ONE issuing thread the wide form only paid from 128 channels on)
    const int wide = Ci >= 32;
    CUtensorMap xmapA, xmapB;
    int rc = wide ? ecg_make_act_tmap64(&xmapA, xb, B, Ci, L, 128, 1) : ecg_make_act_tmap(&xmapA, xb, B, Ci, L, TC_ROWS, Ci / 8);
    if (rc) return rc;
    rc = ecg_make_act_tmap64(&xmapB, xb, B, Ci, L, TC_ROWS - 128, 1);
    if (rc) return rc;
    Conv2Cfg P;
    size_t smem;
    const int grid = conv2_cfg(B, Ci, Co, L, &P, &smem);
    if (grid <= 0) return ECGB200_EUNSUPPORTED;
    P.wide = wide;
    P.Cn = Cn > 0 ? Cn : Co;
    {   // the attribute is per device: set it on every call (cheap, legal during stream capture)
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return (int)e;
    }
    return ecg_launch_pdl(conv_tc_kernel<MODE>, dim3(grid), dim3(C2_THREADS), smem, (cudaStream_t)stream, xmapA, xmapB,
                          (const __nv_bfloat16*)wprep, bias, shift, (__nv_bfloat16*)yb, stat_part, P);
}

extern "C" int ecgb200_conv1d_fwd_stats_bf16(const void* xb, const void* wprep, const float* bias, void* yb,
                                             float* stat_part, int B, int Ci, int Co, int L, void* stream) {
    if (!xb || !wprep || !yb || B <= 0 || L <= 0) return ECGB200_EINVAL;
    // no statistics wanted (dgrad, plain forward): the instantiation without the 64 column-sum accumulators
    if (stat_part == nullptr) return conv_tc_launch<3>(xb, wprep, bias, nullptr, yb, nullptr, B, Ci, Co, L, stream);
    return conv_tc_launch<0>(xb, wprep, bias, nullptr, yb, stat_part, B, Ci, Co, L, stream);
}

// Inference block: pb = maxpool2(relu(scale * conv(xb) + shift)), the eval-mode BatchNorm (and the conv bias)
// folded into per-channel fp32 {scale, shift} (ecgb200_bn_fold_f32).  xb / wprep as above; pb [B][Co/8][L/2][8]
// bf16.  gap_part != NULL (last block): nothing is stored to pb (may be NULL); instead
// gap_part[B * ceil(L/128)][4][Co] receives the time sums of the pooled rows per (128-step tile, lane quarter):
// mean over time = (sum of a window's 4 * ceil(L/128) partials) / (L/2).
extern "C" int ecgb200_conv1d_bn_relu_pool_infer_bf16(const void* xb, const void* wprep, const float* scale,
                                                      const float* shift, void* pb, float* gap_part, int B, int Ci,
                                                      int Co, int L, void* stream) {
    if (!xb || !wprep || !scale || !shift || (!pb && !gap_part) || B <= 0 || L < 2) return ECGB200_EINVAL;
    if (gap_part != nullptr) return conv_tc_launch<2>(xb, wprep, scale, shift, nullptr, gap_part, B, Ci, Co, L, stream);
    return conv_tc_launch<1>(xb, wprep, scale, shift, pb, nullptr, B, Ci, Co, L, stream);
}

// ---------------------------------------------------------------- split-precision (fp32-accurate) inference
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi); the three products hi*hi + lo*hi + hi*lo are THREE TIMES THE INPUT
// CHANNELS of the same implicit GEMM: activations [x_hi | x_lo | x_hi], weights [w_hi | w_hi | w_lo], fp32 accumulation in
// TMEM.  Plane width = Ci rounded up to 16; the three planes are padded with zeros to a multiple of 64 (16 when <= 64).
extern "C" int ecgb200_split_channels(int Ci) {
    const int cpl = (Ci + 15) / 16 * 16, c3 = 3 * cpl;
    return c3 <= 64 ? 64 : (c3 + 63) / 64 * 64;
}

// x fp32 (B, Ci, T) -> xb bf16 [B][Ct/8][T][8], Ct = ecgb200_split_channels(Ci): planes [hi | lo | hi | 0]
__global__ void pack_input_split_bf16_kernel(const float* __restrict__ x, uint4* __restrict__ xb, int B, int Ci, int Cpl,
                                             int Ct, int T) {
    const long long n = (long long)B * (Ct / 8) * T;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const int cc = (int)((i / T) % (Ct / 8));
        const int b = (int)(i / ((long long)T * (Ct / 8)));
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int ch = cc * 8 + j, plane = ch / Cpl, c = ch - plane * Cpl;
            float val = 0.f;
            if (plane < 3 && c < Ci) {
                const float xv = __ldg(x + ((size_t)b * Ci + c) * T + t);
                const float hi = __bfloat162float(__float2bfloat16(xv));
                val = plane == 1 ? xv - hi : xv;          // rounded to bf16 below: hi for planes 0 / 2, lo for plane 1
            }
            v[j] = val;
        }
        xb[i] = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                           tc::pack_bf16(v[6], v[7]));
    }
}

// w fp32 (Co, Ci, 15) -> wf bf16 [15][Ct/8][Co][8]: planes [w_hi | w_hi | w_lo | 0] against the activations' [hi | lo | hi]
__global__ void prep_weights_split_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf, int Co, int Ci,
                                               int Cpl, int Ct) {
    const int n = ECG_KS * Ct * Co;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = i & 7;
        const int o = (i >> 3) % Co;
        const int cc = (i / (8 * Co)) % (Ct / 8);
        const int k = i / (8 * Co * (Ct / 8));
        const int ch = cc * 8 + j, plane = ch / Cpl, c = ch - plane * Cpl;
        float val = 0.f;
        if (plane < 3 && c < Ci) {
            const float wv = w[((size_t)o * Ci + c) * ECG_KS + k];
            const float hi = __bfloat162float(__float2bfloat16(wv));
            val = plane == 2 ? wv - hi : wv;
        }
        wf[i] = __float2bfloat16(val);
    }
}

extern "C" int ecgb200_pack_input_split_bf16(const float* x, void* xb, int B, int Ci, int T, void* stream) {
    if (!x || !xb || B <= 0 || Ci <= 0 || T <= 0) return ECGB200_EINVAL;
    const int Cpl = (Ci + 15) / 16 * 16, Ct = ecgb200_split_channels(Ci);
    const long long n = (long long)B * (Ct / 8) * T;
    const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    pack_input_split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (uint4*)xb, B, Ci, Cpl, Ct, T);
    return ecg_launch_status();
}

extern "C" int ecgb200_conv1d_prep_weights_split_bf16(const float* w, void* wf, int Co, int Ci, void* stream) {
    if (!w || !wf || Co <= 0 || Ci <= 0 || (Co & 7)) return ECGB200_EINVAL;
    const int Cpl = (Ci + 15) / 16 * 16, Ct = ecgb200_split_channels(Ci);
    const int n = ECG_KS * Ct * Co;
    prep_weights_split_bf16_kernel<<<ecg_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)wf, Co, Ci, Cpl, Ct);
    return ecg_launch_status();
}

// One ConvBlock in eval mode at fp32 accuracy: xb [B][Ct/8][L][8] split planes of the Ci input channels
// (Ct = ecgb200_split_channels(Ci)), wprep from ecgb200_conv1d_prep_weights_split_bf16.  pb != NULL: the pooled output as
// split planes [B][Cn/8][L/2][8], Cn = ecgb200_split_channels(Co) (zero the buffer once: the padding plane is never written);
// gap_part != NULL (last block): per-tile time sums as ecgb200_conv1d_bn_relu_pool_infer_bf16.
extern "C" int ecgb200_conv1d_bn_relu_pool_infer_split_bf16(const void* xb, const void* wprep, const float* scale,
                                                            const float* shift, void* pb, float* gap_part, int B,
                                                            int Ci, int Co, int L, void* stream) {
    if (!xb || !wprep || !scale || !shift || (!pb && !gap_part) || B <= 0 || L < 2 || Ci <= 0) return ECGB200_EINVAL;
    const int Ct = ecgb200_split_channels(Ci);
    if (gap_part != nullptr) return conv_tc_launch<2>(xb, wprep, scale, shift, nullptr, gap_part, B, Ct, Co, L, stream);
    return conv_tc_launch<5>(xb, wprep, scale, shift, pb, nullptr, B, Ct, Co, L, stream, ecgb200_split_channels(Co));
}

// Raw conv output + bias in fp32 (B, Co, L) from split-plane inputs (the Grad-CAM front end's 4th conv at fp32 accuracy).
extern "C" int ecgb200_conv1d_fwd_split_f32(const void* xb, const void* wprep, const float* bias, float* y, int B, int Ci,
                                            int Co, int L, void* stream) {
    if (!xb || !wprep || !y || B <= 0 || L <= 0 || Ci <= 0) return ECGB200_EINVAL;
    return conv_tc_launch<6>(xb, wprep, bias, nullptr, nullptr, y, B, ecgb200_split_channels(Ci), Co, L, stream);
}

extern "C" int ecgb200_conv1d_fwd_bf16(const void* xb, const void* wprep, const float* bias, void* yb,
                                       int B, int Ci, int Co, int L, void* stream) {
    return ecgb200_conv1d_fwd_stats_bf16(xb, wprep, bias, yb, nullptr, B, Ci, Co, L, stream);
}

// ---------------------------------------------------------------- wgrad implicit GEMM
// dW_k[o, c] = sum_{b,t} dY[b, t, o] * X[b, t+k-7, c]        (K = time, both operands MN-major)
//
// "Taps as N": for ONE 8-channel chunk of X, the 16-byte rows of the staged tile are
// [row][8 ch]; a B operand whose N-chunk stride (SBO) is 16 bytes -- one row -- makes N-chunk n
// the same chunk shifted by n rows, i.e. tap n.  So a single tcgen05.mma with N = 128 computes
// all 15 taps (+1 unused) of 8 input channels:  D[o][(k, c8)] += dY^T[o][t] * X[t + k][c8].
// That is 8x fewer MMA instructions than one instruction per (tap, 32 channels), which matters
// because a single thread issues them.  CTA tile: 128 output channels (TMEM lanes) x up to 4
// channel chunks (4 x 128 = 512 TMEM columns), accumulated IN TMEM over the CTA's whole share of
// (sample, 128-step time tile) work items; one epilogue per CTA writes a split-K partial
// part[z][o][c/8][16][8] that wgrad_tc_reduce_kernel sums in a fixed order (deterministic).
// Stage = dY tile [ochunks][128 rows][8] + X tile [ncc][144 rows][8] bf16, sized per layer; the ring is as deep as
// shared memory allows (<= WT_MAXST): a tile takes ~1.4 us to land, so the thin stem layers need 6-8 loads in
// flight to keep the issuer busy (3 stages: 1480 cycles per item on L1 against 1024 of MMA time).
constexpr int WT_MAXST = 8;
constexpr int WT_SMEM_BUDGET = 215 * 1024;
constexpr int WT_EPI_BYTES = 4 * 32 * 512;       // epilogue transposition scratch (re-uses the stage ring)

// All MMAs of one (sample, 128-step tile) work item: NCC channel chunks x 8 K-steps, fully unrolled so every
// descriptor is "item base + compile-time constant" (see conv_issue_stage for why).
template <int NCC>
__device__ __forceinline__ void wgrad_issue_item(uint32_t tmem_base, uint64_t alo, uint64_t blo, uint32_t idesc,
                                                 bool accum) {
#pragma unroll
    for (int i = 0; i < NCC; ++i) {
#pragma unroll
        for (int jb = 0; jb < TC_TILE_M / 16; jb += 4) {
            uint32_t dd[4];
            uint64_t al[4], bl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dd[e] = tmem_base + (uint32_t)(i * 128);
                al[e] = alo + (uint64_t)((jb + e) * 16);
                bl[e] = blo + (uint64_t)(i * TC_ROWS + (jb + e) * 16);
            }
            if (jb > 0 || accum) tc::mma_bf16_x4<0xF>(dd, al, bl, idesc);
            else tc::mma_bf16_x4<0xE>(dd, al, bl, idesc);
        }
    }
}

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmapA,
                const __grid_constant__ CUtensorMap xmapB,
                float* __restrict__ part, int Co, int Cip, int L, int B, int ncc, int ochunks, int nst) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WT_MAXST;
    uint64_t* accfull = empty + WT_MAXST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
    uint8_t* stages = smem + TC_HDR;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cb = blockIdx.x, ob = blockIdx.y, z = blockIdx.z, S = gridDim.z;
    long long* const trace = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_conv_trace : nullptr;
    if (threadIdx.x == 0) { CTR(0); CTA_SPAN(0); }
    const int tiles_t = (L + TC_TILE_M - 1) / TC_TILE_M;
    const int items = B * tiles_t;
    const int nloc = (items - z + S - 1) / S;            // items z, z+S, ...   (host guarantees >= 1)
    const uint32_t dybytes = (uint32_t)ochunks * 128 * 16;
    const uint32_t xbytes = (uint32_t)ncc * TC_ROWS * 16;
    const uint32_t stage_bytes = dybytes + xbytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WT_MAXST; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(accfull, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&dymap);
        tc::prefetch_tmap(&xmapA);
        tc::prefetch_tmap(&xmapB);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int slot = 0;
            uint32_t ephase = 1;                                 // fresh barriers: the first pass does not block
            int b = z / tiles_t, tt = z - b * tiles_t;           // item -> (sample, time tile), advanced by S per step
            const int db = S / tiles_t, dt = S - db * tiles_t;
            for (int n = 0; n < nloc; ++n) {
                tc::mbar_wait(empty + slot, ephase);
                uint8_t* st = stages + (size_t)slot * stage_bytes;
                tc::mbar_arrive_expect_tx(full + slot, dybytes + xbytes);
                tc::tma_load_3d(st, &dymap, full + slot, 2 * tt * TC_TILE_M, ob * 16, b);
                for (int c = 0; c < ncc; ++c) {
                    uint8_t* dst = st + dybytes + (size_t)c * (TC_ROWS * 16);
                    tc::tma_load_3d(dst, &xmapA, full + slot, 2 * (tt * TC_TILE_M - ECG_PAD), cb * ncc + c, b);
                    tc::tma_load_3d(dst + 128 * 16, &xmapB, full + slot, 2 * (tt * TC_TILE_M - ECG_PAD + 128), cb * ncc + c, b);
                }
                if (n < 8) CTR(8 + n);                           // loads of item n issued
                if (++slot == nst) { slot = 0; ephase ^= 1; }
                b += db; tt += dt;
                if (tt >= tiles_t) { tt -= tiles_t; ++b; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // lean issue loop (see conv_tc_kernel): descriptors built once, only the address field advances
            const uint32_t idesc = tc::make_idesc_bf16(128, 128, 1, 1);
            // A = dY^T: M (o) chunks 128*16 B apart, K (t) 8-row groups 128 B apart
            const uint64_t adesc0 = tc::make_desc(0, 128, 128 * 16);
            // B = X chunk: N chunk n = tap n = the chunk shifted by n rows (SBO = 16 B)
            const uint64_t bdesc0 = tc::make_desc(0, 128, 16);
            const uint64_t alo0 = adesc0 + (uint64_t)(tc::smem_u32(stages) >> 4);
            const uint64_t blo0 = bdesc0 + (uint64_t)((tc::smem_u32(stages) + dybytes) >> 4);
            int slot = 0;
            uint32_t fphase = 0, accum = 0;
            for (int n = 0; n < nloc; ++n) {
                tc::mbar_wait(full + slot, fphase);
                if (n < 8) CTR(16 + n);                          // item n landed
                tc::fence_after_sync();
                const uint64_t alo = alo0 + (uint64_t)((uint32_t)slot * (stage_bytes >> 4));
                const uint64_t blo = blo0 + (uint64_t)((uint32_t)slot * (stage_bytes >> 4));
                if (ncc == 4) wgrad_issue_item<4>(tmem_base, alo, blo, idesc, accum != 0);
                else if (ncc == 2) wgrad_issue_item<2>(tmem_base, alo, blo, idesc, accum != 0);
                else wgrad_issue_item<1>(tmem_base, alo, blo, idesc, accum != 0);
                accum = 1;
                tc::mma_commit(empty + slot);
                if (n < 8) CTR(24 + n);                          // MMAs of item n issued
                if (++slot == nst) { slot = 0; fphase ^= 1; }
            }
            tc::mma_commit(accfull);
            CTR(32);
        }
    } else {
        const int q = warp & 3;
        tc::mbar_wait(accfull, 0);
        if (threadIdx.x == 64) CTR(33);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
        // A thread owns one output-channel row (TMEM lane) whose 512 bytes per channel chunk are contiguous in the
        // partial layout, but neighbouring rows are kilobytes apart: storing straight from the registers makes every
        // store instruction 32 separate 16-byte transactions (measured: 19 k cycles of epilogue per CTA).  So each
        // warp transposes its 32 rows x 512 B through the (now idle) stage buffers, XOR-swizzled so that both the
        // row-wise writes and the column-wise reads are bank-conflict free, and every global store instruction
        // writes 512 contiguous bytes of one row.
        uint8_t* stg = stages + (size_t)(warp - 2) * (32 * 512);
        const int nchunk = ob * 128 + 32 * q < Co ? ncc : 0;         // a warp whose 32 rows lie beyond Co has nothing to store
        for (int i = 0; i < nchunk; ++i) {
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                float v[32];
                tc::tmem_ld32(taddr + (uint32_t)(i * 128 + g * 32), v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    *reinterpret_cast<float4*>(stg + lane * 512 + (((g * 8 + e) ^ (lane & 7)) << 4)) =
                        make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
            }
            __syncwarp();
            float* dst0 = part + (((size_t)z * Co + ob * 128 + 32 * q) * (Cip / 8) + cb * ncc + i) * 128 + lane * 4;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                const float4 t = *reinterpret_cast<const float4*>(stg + rr * 512 + ((lane ^ (rr & 7)) << 4));
                if (ob * 128 + 32 * q + rr < Co)
                    *reinterpret_cast<float4*>(dst0 + (size_t)rr * (Cip / 8) * 128) = t;
            }
            __syncwarp();
        }
    }
    if (threadIdx.x == 64) CTR(34);
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc::tmem_dealloc(tmem_base, 512);
        if (lane == 0) CTA_SPAN(1);
    }
}

// ---------------------------------------------------------------- thin layers: "taps as M"
// Stem layers (Co <= 64): with dY^T as the A operand only Co of the 128 MMA rows are real.  Swap the roles:
//   D[(k, c8)][o] += X[t + k][c8]^T * dY[t][o]
// A = one 8-channel chunk of the X tile, MN-major with an M-chunk stride of ONE ROW (16 B): M-chunk m = the chunk shifted
//     by m rows = tap m, so M = 16 taps x 8 channels = 128 rows, all real (tap 15 is computed and dropped);
// B = the dY tile [Co/8][128 rows][8], MN-major, N = Co (32 / 64): the 47-cycle instruction instead of the 64-cycle N = 128.
// Per (sample, 128-step tile) item: Cip/8 chunks x 8 K-steps MMAs, half the tensor time of the taps-as-N form, and the whole
// dY tile is read once per item (all chunks live in one CTA: Cip/8 * Co <= 256 TMEM columns).  The accumulator of chunk i
// sits in columns [i * Co, (i + 1) * Co): TMEM lane = (tap, c8), column = o, i.e. for a fixed o the 128 lanes are 512
// CONTIGUOUS bytes of the partial layout part[z][o][c/8][16][8] -- every store instruction of the epilogue is coalesced as is.
template <int NCC>
__device__ __forceinline__ void wgrad_thin_issue_item(uint32_t tmem_base, uint32_t co, uint64_t alo, uint64_t blo,
                                                      uint32_t idesc, bool accum) {
#pragma unroll
    for (int i = 0; i < NCC; ++i) {
#pragma unroll
        for (int jb = 0; jb < TC_TILE_M / 16; jb += 4) {
            uint32_t dd[4];
            uint64_t al[4], bl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dd[e] = tmem_base + (uint32_t)i * co;
                al[e] = alo + (uint64_t)(i * TC_ROWS + (jb + e) * 16);       // X chunk i, K-step = 16 time rows
                bl[e] = blo + (uint64_t)((jb + e) * 16);                      // dY tile, same K-step
            }
            if (jb > 0 || accum) tc::mma_bf16_x4<0xF>(dd, al, bl, idesc);
            else tc::mma_bf16_x4<0xE>(dd, al, bl, idesc);
        }
    }
}

__global__ void __launch_bounds__(192, 1)
wgrad_thin_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                  float* __restrict__ part, int Co, int Cip, int L, int B, int nst, uint32_t tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WT_MAXST;
    uint64_t* accfull = empty + WT_MAXST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
    uint8_t* stages = smem + TC_HDR;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int z = blockIdx.x, S = gridDim.x, ncc = Cip / 8;
    long long* const trace = blockIdx.x == 0 ? g_conv_trace : nullptr;
    if (threadIdx.x == 0) { CTR(0); CTA_SPAN(0); }
    const int tiles_t = (L + TC_TILE_M - 1) / TC_TILE_M;
    const int items = B * tiles_t;
    const int nloc = (items - z + S - 1) / S;            // items z, z+S, ...   (host guarantees >= 1)
    const uint32_t dybytes = (uint32_t)Co * 128 * 2;     // [Co/8][128 rows][8] bf16
    const uint32_t xbytes = (uint32_t)ncc * TC_ROWS * 16;
    const uint32_t stage_bytes = dybytes + xbytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WT_MAXST; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(accfull, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&dymap);
        tc::prefetch_tmap(&xmap);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int slot = 0;
            uint32_t ephase = 1;                                 // fresh barriers: the first pass does not block
            int b = z / tiles_t, tt = z - b * tiles_t;           // item -> (sample, time tile), advanced by S per step
            const int db = S / tiles_t, dt = S - db * tiles_t;
            for (int n = 0; n < nloc; ++n) {
                tc::mbar_wait(empty + slot, ephase);
                uint8_t* st = stages + (size_t)slot * stage_bytes;
                tc::mbar_arrive_expect_tx(full + slot, stage_bytes);
                tc::tma_load_3d(st, &dymap, full + slot, 2 * tt * TC_TILE_M, 0, b);          // one box: 128 rows x Co/8 chunks
                tc::tma_load_4d(st + dybytes, &xmap, full + slot, 0, tt * TC_TILE_M - ECG_PAD, 0, b);   // {8, 144, Cip/8, 1}
                if (n < 8) CTR(8 + n);
                if (++slot == nst) { slot = 0; ephase ^= 1; }
                b += db; tt += dt;
                if (tt >= tiles_t) { tt -= tiles_t; ++b; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, Co, 1, 1);
            // A = X chunk: M chunk m = tap m = the chunk shifted by m rows (SBO = 16 B), K 8-row groups 128 B apart
            const uint64_t adesc0 = tc::make_desc(0, 128, 16);
            // B = dY: N (o) chunks 128 rows * 16 B apart, K (t) 8-row groups 128 B apart
            const uint64_t bdesc0 = tc::make_desc(0, 128, 128 * 16);
            const uint64_t alo0 = adesc0 + (uint64_t)((tc::smem_u32(stages) + dybytes) >> 4);
            const uint64_t blo0 = bdesc0 + (uint64_t)(tc::smem_u32(stages) >> 4);
            int slot = 0;
            uint32_t fphase = 0, accum = 0;
            for (int n = 0; n < nloc; ++n) {
                tc::mbar_wait(full + slot, fphase);
                if (n < 8) CTR(16 + n);
                tc::fence_after_sync();
                const uint64_t alo = alo0 + (uint64_t)((uint32_t)slot * (stage_bytes >> 4));
                const uint64_t blo = blo0 + (uint64_t)((uint32_t)slot * (stage_bytes >> 4));
                if (ncc == 4) wgrad_thin_issue_item<4>(tmem_base, (uint32_t)Co, alo, blo, idesc, accum != 0);
                else if (ncc == 2) wgrad_thin_issue_item<2>(tmem_base, (uint32_t)Co, alo, blo, idesc, accum != 0);
                else wgrad_thin_issue_item<1>(tmem_base, (uint32_t)Co, alo, blo, idesc, accum != 0);
                accum = 1;
                tc::mma_commit(empty + slot);
                if (n < 8) CTR(24 + n);
                if (++slot == nst) { slot = 0; fphase ^= 1; }
            }
            tc::mma_commit(accfull);
            CTR(32);
        }
    } else {
        const int q = warp & 3;
        tc::mbar_wait(accfull, 0);
        if (threadIdx.x == 64) CTR(33);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
        for (int i = 0; i < ncc; ++i) {
            for (int c0 = 0; c0 < Co; c0 += 32) {
                float v[32];
                tc::tmem_ld32(taddr + (uint32_t)(i * Co + c0), v);
                tc::tmem_ld_wait();
                // lane (tap, c8) of output channel o = 512 contiguous bytes over the 128 lanes: coalesced as is
                float* dst = part + (((size_t)z * Co + c0) * ncc + i) * 128 + 32 * q + lane;
#pragma unroll
                for (int e = 0; e < 32; ++e) dst[(size_t)e * ncc * 128] = v[e];
            }
        }
        if (threadIdx.x == 64) CTR(34);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc::tmem_dealloc(tmem_base, tmem_cols);
        if (lane == 0) CTA_SPAN(1);
    }
}

// dW[o][c][k] = sum_z part[z][o][c/8][k][c%8]  (c < Ci, k < 15);  db[o] = sum_j db_part[o][j]
// Block = RC x 32 float4 columns of the partial layout x ZL z-lanes (ZL * RC = 32): every load is a coalesced
// 512-byte row; few splits (S <= 24) -> 8 z-lanes x 4 column sets per thread (independent loads in flight, one wave
// of blocks), many splits over a small weight tensor (the stem layers: S = 148) -> 32 z-lanes, so that no thread
// walks more than ~5 partials serially.  The z-lanes are combined through shared memory in a fixed order
// (deterministic for a given shape).
template <int ZL>                                          // z-lanes per block; RC = 32 / ZL column sets per thread
__global__ void __launch_bounds__(32 * ZL)
wgrad_tc_reduce_kernel(const float4* __restrict__ part, const float* __restrict__ db_part,
                       float* __restrict__ dw, float* __restrict__ db, int S, int Co, int Ci, int Cip,
                       int ndb, int nblk_w) {
    constexpr int RC = 32 / ZL;
    __shared__ float4 red[ZL][32 * RC];
    if ((int)blockIdx.x >= nblk_w) {                       // tail blocks: conv-bias gradient
        const int o = (blockIdx.x - nblk_w) * blockDim.x + threadIdx.x;
        if (o < Co && db != nullptr) {
            float s = 0.f;
            if (db_part != nullptr)
                for (int j = 0; j < ndb; ++j) s += db_part[(size_t)o * ndb + j];
            db[o] = s;
        }
        return;
    }
    const int n4 = Co * Cip * 4;                           // float4 columns per partial
    const int lane = threadIdx.x & 31, zl = threadIdx.x >> 5;
    const int col0 = blockIdx.x * (32 * RC) + lane;
    float4 acc[RC];
#pragma unroll
    for (int c = 0; c < RC; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int z = zl; z < S; z += ZL) {
        float4 v[RC];
#pragma unroll
        for (int c = 0; c < RC; ++c)
            v[c] = col0 + 32 * c < n4 ? __ldg(part + (size_t)z * n4 + col0 + 32 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < RC; ++c) { acc[c].x += v[c].x; acc[c].y += v[c].y; acc[c].z += v[c].z; acc[c].w += v[c].w; }
    }
#pragma unroll
    for (int c = 0; c < RC; ++c) red[zl][32 * c + lane] = acc[c];
    __syncthreads();
    if (zl < RC) {                                         // z-lane c finishes column set c
        const int col = col0 + 32 * zl;
        if (col < n4) {
            float4 t = red[0][32 * zl + lane];
#pragma unroll
            for (int i = 1; i < ZL; ++i) { const float4 v = red[i][32 * zl + lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
            const int idx = col * 4;                       // element index in [o][c/8][16][8]
            const int c8 = idx & 7, k = (idx >> 3) & 15;
            const int cc = (idx >> 7) % (Cip / 8), o = idx / (Cip * 16);
            if (k < ECG_KS) {
                const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = cc * 8 + c8 + e;
                    if (c < Ci) dw[((size_t)o * Ci + c) * ECG_KS + k] = v[e];
                }
            }
        }
    }
}

// thin layers: taps-as-M kernel (all chunks of the input in one CTA: Cip/8 * Co TMEM columns)
static bool wgrad_is_thin(int Cip, int Co) {
    return Co <= 64 && (Co & 31) == 0 && Cip <= 32 && Cip / 8 * Co <= 256;
}

static int wgrad_launch_reduce(const float4* pw, const float* db_part, float* dw, float* db, int S, int Co, int Ci, int Cip,
                               int ndb, cudaStream_t st) {
    if (S <= 24) {
        const int nblk_w = ecg_cdiv(Co * Cip * 4, 128);
        wgrad_tc_reduce_kernel<8><<<nblk_w + ecg_cdiv(Co, 256), 256, 0, st>>>(pw, db_part, dw, db, S, Co, Ci, Cip, ndb, nblk_w);
    } else if (S <= 80) {
        const int nblk_w = ecg_cdiv(Co * Cip * 4, 64);
        wgrad_tc_reduce_kernel<16><<<nblk_w + ecg_cdiv(Co, 512), 512, 0, st>>>(pw, db_part, dw, db, S, Co, Ci, Cip, ndb, nblk_w);
    } else {
        const int nblk_w = ecg_cdiv(Co * Cip * 4, 32);
        wgrad_tc_reduce_kernel<32><<<nblk_w + ecg_cdiv(Co, 1024), 1024, 0, st>>>(pw, db_part, dw, db, S, Co, Ci, Cip, ndb, nblk_w);
    }
    return ecg_launch_status();
}

#include "wgrad_tc2.cuh"

static void wgrad_tc_cfg(int B, int Cip, int Co, int L, int* ncc, int* S) {
    *ncc = Cip / 8 < 4 ? Cip / 8 : 4;
    if (wgrad_is_thin(Cip, Co)) *ncc = Cip / 8;
    // <= 64 output channels without the thin kernel: two chunks per CTA -- half the split-K partial bytes for the same MMA
    // count (measured at B=256: block 2 26.7 -> 23.4 us); wide layers are bound by re-reading the dY tile per chunk group
    // and stay at four (two chunks: block 4 41.5 -> 51.1 us, block 3 27.9 -> 30.0 us)
    else if (Co <= 64 && *ncc > 2) *ncc = 2;
    const int blocks_oc = (Cip / 8 / *ncc) * ecg_cdiv(Co, 128);
    const int items = B * ecg_cdiv(L, TC_TILE_M);
    int s = 148 / blocks_oc;
    if (s < 1) s = 1;
    if (s > items) s = items;
    *S = s;
}

extern "C" size_t ecgb200_conv1d_wgrad_bf16_ws_bytes(int B, int Ci, int Co, int L) {
    const int Cip = (Ci + 15) / 16 * 16;
    int ncc, S;
    wgrad_tc_cfg(B, Cip, Co, L, &ncc, &S);
    int pncc, pncb, pS;
    if (wgrad_pair_cfg(B, Cip, Co, L, &pncc, &pncb, &pS) && pS > S) S = pS;
    return (size_t)S * Co * Cip * 16 * sizeof(float);
}

// dyb [B][Co/8][L][8], xb [B][Cip/8][L][8] bf16 -> dw fp32 (Co, Ci, 15), db fp32 (Co) [NULL to skip].
// db_part: optional fp32 [Co][ndb] per-block sums of dy produced by the BN backward kernel.
extern "C" int ecgb200_conv1d_wgrad_bf16(const void* dyb, const void* xb, float* dw, float* db,
                                         const float* db_part, int ndb, void* ws, int B, int Ci, int Co,
                                         int L, void* stream) {
    if (!dyb || !xb || !dw || !ws || B <= 0 || Ci <= 0 || L <= 0) return ECGB200_EINVAL;
    const int Cip = (Ci + 15) / 16 * 16;
    if (Co <= 0 || (Co & 7) || Co > 256 || Cip > 256 || (Cip > 16 && (Cip & 31))) return ECGB200_EUNSUPPORTED;
    int ncc, S;
    wgrad_tc_cfg(B, Cip, Co, L, &ncc, &S);
    cudaStream_t st = (cudaStream_t)stream;
    if (wgrad_is_thin(Cip, Co)) {
        // taps as M: one CTA holds all Cip/8 chunks, the batch is split S ways over the SMs
        CUtensorMap dymap, xmap;
        int rc = ecg_make_act_tmap64(&dymap, dyb, B, Co, L, TC_TILE_M, Co / 8);
        if (rc) return rc;
        rc = ecg_make_act_tmap(&xmap, xb, B, Cip, L, TC_ROWS, Cip / 8);
        if (rc) return rc;
        const size_t stage = (size_t)Co * 256 + (size_t)(Cip / 8) * TC_ROWS * 16;
        int nst = (int)(WT_SMEM_BUDGET / stage);
        if (nst > WT_MAXST) nst = WT_MAXST;
        const size_t smem = TC_HDR + (size_t)nst * stage;
        cudaError_t e = cudaFuncSetAttribute(wgrad_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return (int)e;
        wgrad_thin_kernel<<<S, 192, smem, st>>>(dymap, xmap, (float*)ws, Co, Cip, L, B, nst, tmem_cols_for(Cip / 8 * Co));
        rc = ecg_launch_status();
        if (rc) return rc;
        return wgrad_launch_reduce((const float4*)ws, db_part, dw, db, S, Co, Ci, Cip, ndb, st);
    }
    int pncc, pncb, pS;
    if (wgrad_pair_cfg(B, Cip, Co, L, &pncc, &pncb, &pS)) {
        // wide layers: taps as M on CTA pairs, each SM staging half of the dY tile and its own chunks of the X tile
        CUtensorMap dymap, xmap;
        int rc = ecg_make_act_tmap64(&dymap, dyb, B, Co, L, TC_TILE_M, Co / 16);
        if (rc) return rc;
        rc = ecg_make_act_tmap(&xmap, xb, B, Cip, L, TC_ROWS, pncc);
        if (rc) return rc;
        const size_t stage = (size_t)(Co / 2) * 256 + (size_t)pncc * TC_ROWS * 16;
        int nst = (int)(WT_SMEM_BUDGET / stage);
        if (nst > WT_MAXST) nst = WT_MAXST;
        const size_t smem = TC_HDR + (size_t)nst * stage;
        cudaError_t e = cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return (int)e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pncb * pS); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, wgrad_pair_kernel, dymap, xmap, (float*)ws, Co, Cip, L, B, pncc, pncb, nst);
        if (e != cudaSuccess) return (int)e;
        rc = ecg_launch_status();
        if (rc) return rc;
        return wgrad_launch_reduce((const float4*)ws, db_part, dw, db, pS, Co, Ci, Cip, ndb, st);
    }
    const int ochunks = Co >= 128 ? 16 : Co / 8;
    CUtensorMap dymap, xmapA, xmapB;
    int rc = ecg_make_act_tmap64(&dymap, dyb, B, Co, L, TC_TILE_M, ochunks);
    if (rc) return rc;
    rc = ecg_make_act_tmap64(&xmapA, xb, B, Cip, L, 128, 1);
    if (rc) return rc;
    rc = ecg_make_act_tmap64(&xmapB, xb, B, Cip, L, TC_ROWS - 128, 1);
    if (rc) return rc;
    const size_t stage = (size_t)ochunks * 128 * 16 + (size_t)ncc * TC_ROWS * 16;
    // the A operand is always described as M = 128 rows = 16 chunks (rows >= Co are computed and discarded), so a
    // thin dY tile is read 2 KB x (16 - ochunks) past its end: keep that much slack behind the last stage
    const size_t slack = (size_t)(16 - ochunks) * 128 * 16;
    int nst = (int)((WT_SMEM_BUDGET - slack) / stage);
    if (nst > WT_MAXST) nst = WT_MAXST;
    size_t smem = TC_HDR + (size_t)nst * stage + slack;
    if (smem < TC_HDR + WT_EPI_BYTES) smem = TC_HDR + WT_EPI_BYTES;
    {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid(Cip / 8 / ncc, ecg_cdiv(Co, 128), S);
    wgrad_tc_kernel<<<grid, 192, smem, st>>>(dymap, xmapA, xmapB, (float*)ws, Co, Cip, L, B, ncc, ochunks, nst);
    rc = ecg_launch_status();
    if (rc) return rc;
    return wgrad_launch_reduce((const float4*)ws, db_part, dw, db, S, Co, Ci, Cip, ndb, st);
}
