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
constexpr int TC_TILE_M = 128;      // output time steps per CTA (TMEM lanes)
constexpr int TC_ROWS = 144;        // input rows staged: 128 + 14 halo, rounded to 8
constexpr int TC_NST = 4;           // weight ring depth
constexpr int TC_HDR = 1024;        // barriers + TMEM slot

__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __nv_bfloat16* __restrict__ wprep,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
               int Ci, int Co, int L, int kch, uint32_t tmem_cols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [TC_NST]
    uint64_t* empty = full + TC_NST;                                // [TC_NST]
    uint64_t* xfull = empty + TC_NST;
    uint64_t* accfull = xfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
    const uint32_t xbytes = (uint32_t)Ci * TC_ROWS * 2;
    const uint32_t xbytes_al = (xbytes + 1023u) & ~1023u;
    uint8_t* xs = smem + TC_HDR;
    uint8_t* wsm = xs + xbytes_al;
    const uint32_t stage_bytes = (uint32_t)kch * Co * 2;
    const int groups = Ci / kch;
    const int nstage = ECG_KS * groups;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * TC_TILE_M;
    const int b = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_NST; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(xfull, 1);
        tc::mbar_init(accfull, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&xmap);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(xfull, xbytes);
            tc::tma_load_4d(xs, &xmap, xfull, 0, t0 - ECG_PAD, 0, b);
            for (int s = 0; s < nstage; ++s) {
                const int slot = s % TC_NST;
                if (s >= TC_NST) tc::mbar_wait(empty + slot, ((s / TC_NST) - 1) & 1);
                tc::mbar_arrive_expect_tx(full + slot, stage_bytes);
                tc::bulk_load(wsm + (size_t)slot * stage_bytes,
                              reinterpret_cast<const uint8_t*>(wprep) + (size_t)s * stage_bytes, stage_bytes,
                              full + slot);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(TC_TILE_M, Co, 0, 0);
            const uint32_t xs_addr = tc::smem_u32(xs);
            const uint32_t ws_addr = tc::smem_u32(wsm);
            tc::mbar_wait(xfull, 0);
            tc::fence_after_sync();
            for (int s = 0; s < nstage; ++s) {
                const int slot = s % TC_NST;
                const int k = s / groups, g = s - k * groups;
                tc::mbar_wait(full + slot, (s / TC_NST) & 1);
                tc::fence_after_sync();
                const uint32_t wbase = ws_addr + slot * stage_bytes;
                const uint32_t xbase = xs_addr + (uint32_t)(g * (kch / 8)) * (TC_ROWS * 16) + (uint32_t)k * 16;
                for (int j = 0; j < kch / 16; ++j) {
                    const uint64_t ad = tc::make_desc(xbase + (uint32_t)(2 * j) * (TC_ROWS * 16), TC_ROWS * 16, 128);
                    const uint64_t bd = tc::make_desc(wbase + (uint32_t)(2 * j) * (Co * 16), (uint32_t)Co * 16, 128);
                    tc::mma_bf16(tmem_base, ad, bd, idesc, (s > 0 || j > 0) ? 1u : 0u);
                }
                tc::mma_commit(empty + slot);          // frees the weight slot when these MMAs finish
            }
            tc::mma_commit(accfull);                   // accumulator complete
        }
    } else {
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = 32 * q + lane;
        const int t = t0 + row;
        tc::mbar_wait(accfull, 0);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
        const size_t chunk_stride = (size_t)L * 8;     // elements between channel chunks
        __nv_bfloat16* yrow = y + ((size_t)b * (Co / 8) * L + t) * 8;
        for (int c0 = 0; c0 < Co; c0 += 32) {
            float v[32];
            tc::tmem_ld32(taddr + (uint32_t)c0, v);
            tc::tmem_ld_wait();
            if (t < L) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o[j] = v[8 * i + j] + (bias != nullptr ? __ldg(bias + c0 + 8 * i + j) : 0.f);
                    const uint4 pk = make_uint4(tc::pack_bf16(o[0], o[1]), tc::pack_bf16(o[2], o[3]),
                                                tc::pack_bf16(o[4], o[5]), tc::pack_bf16(o[6], o[7]));
                    *reinterpret_cast<uint4*>(yrow + (size_t)(c0 / 8 + i) * chunk_stride) = pk;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, tmem_cols);
}

static uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

// xb [B][Ci/8][L][8] bf16 (Ci % 16 == 0), wprep [15][Ci/8][Co][8] bf16, bias fp32 (Co) or NULL,
// yb [B][Co/8][L][8] bf16.  Co % 32 == 0, Co <= 256, Ci <= 256.
extern "C" int ecgb200_conv1d_fwd_bf16(const void* xb, const void* wprep, const float* bias, void* yb,
                                       int B, int Ci, int Co, int L, void* stream) {
    if (!xb || !wprep || !yb || B <= 0 || L <= 0) return ECGB200_EINVAL;
    if (Ci <= 0 || (Ci & 15) || Ci > 256 || Co <= 0 || (Co & 31) || Co > 256 || B > 65535) return ECGB200_EUNSUPPORTED;
    CUtensorMap xmap;
    int rc = ecg_make_act_tmap(&xmap, xb, B, Ci, L, TC_ROWS, Ci / 8);
    if (rc) return rc;
    const int kch = Ci < 64 ? Ci : 64;
    const uint32_t xbytes_al = ((uint32_t)Ci * TC_ROWS * 2 + 1023u) & ~1023u;
    const size_t smem = TC_HDR + xbytes_al + (size_t)TC_NST * kch * Co * 2;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    dim3 grid(ecg_cdiv(L, TC_TILE_M), B);
    conv_tc_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(xmap, (const __nv_bfloat16*)wprep, bias,
                                                              (__nv_bfloat16*)yb, Ci, Co, L, kch, tmem_cols_for(Co));
    return ecg_launch_status();
}
