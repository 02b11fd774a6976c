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

// Each CTA owns up to R output tiles (R accumulators side by side in TMEM) that share ONE pass
// over the weight ring: the L2 -> shared-memory weight traffic per FLOP drops by R, which is what
// bounds this kernel otherwise (a 128-row tile reuses each weight byte only 128 times).
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __nv_bfloat16* __restrict__ wprep,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
               int Ci, int Co, int L, int kch, uint32_t tmem_cols, int R, int total_tiles, int tiles_t) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [TC_NST]
    uint64_t* empty = full + TC_NST;                                // [TC_NST]
    uint64_t* xfull = empty + TC_NST;
    uint64_t* accfull = xfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
    const uint32_t xbytes = (uint32_t)Ci * TC_ROWS * 2;
    const uint32_t xbytes_al = (xbytes + 1023u) & ~1023u;
    uint8_t* xs = smem + TC_HDR;
    uint8_t* wsm = xs + (size_t)R * xbytes_al;
    const uint32_t stage_bytes = (uint32_t)kch * Co * 2;
    const int groups = Ci / kch;
    const int nstage = ECG_KS * groups;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile0 = blockIdx.x * R;
    const int rcount = min(R, total_tiles - tile0);

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
            tc::mbar_arrive_expect_tx(xfull, xbytes * (uint32_t)rcount);
            for (int r = 0; r < rcount; ++r) {
                const int tile = tile0 + r;
                const int b = tile / tiles_t, t0 = (tile - b * tiles_t) * TC_TILE_M;
                tc::tma_load_4d(xs + (size_t)r * xbytes_al, &xmap, xfull, 0, t0 - ECG_PAD, 0, b);
            }
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
                const uint32_t xoff = (uint32_t)(g * (kch / 8)) * (TC_ROWS * 16) + (uint32_t)k * 16;
                for (int r = 0; r < rcount; ++r) {
                    const uint32_t xbase = xs_addr + (uint32_t)r * xbytes_al + xoff;
                    for (int j = 0; j < kch / 16; ++j) {
                        const uint64_t ad = tc::make_desc(xbase + (uint32_t)(2 * j) * (TC_ROWS * 16), TC_ROWS * 16, 128);
                        const uint64_t bd = tc::make_desc(wbase + (uint32_t)(2 * j) * (Co * 16), (uint32_t)Co * 16, 128);
                        tc::mma_bf16(tmem_base + (uint32_t)(r * Co), ad, bd, idesc, (s > 0 || j > 0) ? 1u : 0u);
                    }
                }
                tc::mma_commit(empty + slot);          // frees the weight slot when these MMAs finish
            }
            tc::mma_commit(accfull);                   // all accumulators complete
        }
    } else {
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = 32 * q + lane;
        tc::mbar_wait(accfull, 0);
        tc::fence_after_sync();
        const size_t chunk_stride = (size_t)L * 8;     // elements between channel chunks
        for (int r = 0; r < rcount; ++r) {
            const int tile = tile0 + r;
            const int b = tile / tiles_t, t = (tile - b * tiles_t) * TC_TILE_M + row;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(r * Co);
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
    const size_t wring = (size_t)TC_NST * kch * Co * 2;
    const int tiles_t = ecg_cdiv(L, TC_TILE_M);
    const int total = B * tiles_t;
    // tiles per CTA: bounded by TMEM (512 fp32 columns), shared memory (~200 KB) and by keeping >= ~1 wave of CTAs
    int R = 512 / Co;
    if (R > 2) R = 2;
    if (const char* e = getenv("ECGB200_CONV_R")) { const int v = atoi(e); if (v >= 1 && v <= 16) R = v; }   // tuning
    if (R * Co > 512) R = 512 / Co;
    while (R > 1 && TC_HDR + (size_t)R * xbytes_al + wring > 225 * 1024) --R;
    while (R > 1 && ecg_cdiv(total, R) < 120) --R;
    const size_t smem = TC_HDR + (size_t)R * xbytes_al + wring;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_set = smem;
    }
    conv_tc_kernel<<<ecg_cdiv(total, R), 192, smem, (cudaStream_t)stream>>>(
        xmap, (const __nv_bfloat16*)wprep, bias, (__nv_bfloat16*)yb, Ci, Co, L, kch, tmem_cols_for(R * Co), R,
        total, tiles_t);
    return ecg_launch_status();
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
constexpr int WT_NST = 3;
constexpr int WT_DY_BYTES = 16 * 128 * 16;       // [<=16 chunks][128 rows][8] bf16
constexpr int WT_X_BYTES = 4 * TC_ROWS * 16;     // [<=4 chunks][144 rows][8] bf16
constexpr int WT_STAGE = WT_DY_BYTES + WT_X_BYTES;

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                float* __restrict__ part, int Co, int Cip, int L, int B, int ncc, int ochunks) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WT_NST;
    uint64_t* accfull = empty + WT_NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
    uint8_t* stages = smem + TC_HDR;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cb = blockIdx.x, ob = blockIdx.y, z = blockIdx.z, S = gridDim.z;
    const int tiles_t = (L + TC_TILE_M - 1) / TC_TILE_M;
    const int items = B * tiles_t;
    const int nloc = (items - z + S - 1) / S;            // items z, z+S, ...   (host guarantees >= 1)
    const uint32_t dybytes = (uint32_t)ochunks * 128 * 16;
    const uint32_t xbytes = (uint32_t)ncc * TC_ROWS * 16;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WT_NST; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(accfull, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&dymap);
        tc::prefetch_tmap(&xmap);
    }
    if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int n = 0; n < nloc; ++n) {
                const int it = z + n * S;
                const int b = it / tiles_t, t0 = (it - b * tiles_t) * TC_TILE_M;
                const int slot = n % WT_NST;
                if (n >= WT_NST) tc::mbar_wait(empty + slot, ((n / WT_NST) - 1) & 1);
                uint8_t* st = stages + (size_t)slot * WT_STAGE;
                tc::mbar_arrive_expect_tx(full + slot, dybytes + xbytes);
                tc::tma_load_4d(st, &dymap, full + slot, 0, t0, ob * 16, b);
                tc::tma_load_4d(st + WT_DY_BYTES, &xmap, full + slot, 0, t0 - ECG_PAD, cb * ncc, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, 128, 1, 1);
            const uint32_t st_addr = tc::smem_u32(stages);
            for (int n = 0; n < nloc; ++n) {
                const int slot = n % WT_NST;
                tc::mbar_wait(full + slot, (n / WT_NST) & 1);
                tc::fence_after_sync();
                const uint32_t dy_addr = st_addr + slot * WT_STAGE;
                const uint32_t x_addr = dy_addr + WT_DY_BYTES;
#pragma unroll 1
                for (int i = 0; i < ncc; ++i) {
#pragma unroll
                    for (int j = 0; j < TC_TILE_M / 16; ++j) {
                        // A = dY^T: M (o) chunks 128*16 B apart, K (t) 8-row groups 128 B apart
                        const uint64_t ad = tc::make_desc(dy_addr + (uint32_t)j * 256, 128, 128 * 16);
                        // B = X chunk i: N chunk n = tap n = the chunk shifted by n rows (SBO = 16 B)
                        const uint64_t bd = tc::make_desc(x_addr + (uint32_t)i * (TC_ROWS * 16) + (uint32_t)j * 256, 128, 16);
                        tc::mma_bf16(tmem_base + (uint32_t)(i * 128), ad, bd, idesc, (n > 0 || j > 0) ? 1u : 0u);
                    }
                }
                tc::mma_commit(empty + slot);
            }
            tc::mma_commit(accfull);
        }
    } else {
        const int q = warp & 3;
        const int o = ob * 128 + 32 * q + lane;
        tc::mbar_wait(accfull, 0);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
        for (int i = 0; i < ncc; ++i) {
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                float v[32];
                tc::tmem_ld32(taddr + (uint32_t)(i * 128 + g * 32), v);
                tc::tmem_ld_wait();
                if (o < Co) {
                    float* dst = part + ((((size_t)z * Co + o) * (Cip / 8) + cb * ncc + i) * 16 + 4 * g) * 8;
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

// dW[o][c][k] = sum_z part[z][o][c/8][k][c%8]  (c < Ci, k < 15);  db[o] = sum_j db_part[o][j]
// Block = 32 float4 columns of the partial layout x 8 z-lanes: every load is a coalesced 512-byte
// row, the 8 z-lanes are combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
wgrad_tc_reduce_kernel(const float4* __restrict__ part, const float* __restrict__ db_part,
                       float* __restrict__ dw, float* __restrict__ db, int S, int Co, int Ci, int Cip,
                       int ndb, int nblk_w) {
    __shared__ float4 red[8][32];
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
    const int col = blockIdx.x * 32 + lane;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < n4) {
        for (int z = zl; z < S; z += 8) {
            const float4 v = __ldg(part + (size_t)z * n4 + col);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    red[zl][lane] = acc;
    __syncthreads();
    if (zl == 0 && col < n4) {
        float4 t = red[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) { const float4 v = red[i][lane]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        const int idx = col * 4;                           // element index in [o][c/8][16][8]
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

static void wgrad_tc_cfg(int B, int Cip, int Co, int L, int* ncc, int* S) {
    *ncc = Cip / 8 < 4 ? Cip / 8 : 4;
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
    const int ochunks = Co >= 128 ? 16 : Co / 8;
    CUtensorMap dymap, xmap;
    int rc = ecg_make_act_tmap(&dymap, dyb, B, Co, L, TC_TILE_M, ochunks);
    if (rc) return rc;
    rc = ecg_make_act_tmap(&xmap, xb, B, Cip, L, TC_ROWS, ncc);
    if (rc) return rc;
    const size_t smem = TC_HDR + (size_t)WT_NST * WT_STAGE;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(Cip / 8 / ncc, ecg_cdiv(Co, 128), S);
    wgrad_tc_kernel<<<grid, 192, smem, st>>>(dymap, xmap, (float*)ws, Co, Cip, L, B, ncc, ochunks);
    rc = ecg_launch_status();
    if (rc) return rc;
    const int nblk_w = ecg_cdiv(Co * Cip * 4, 32);
    wgrad_tc_reduce_kernel<<<nblk_w + ecg_cdiv(Co, 256), 256, 0, st>>>((const float4*)ws, db_part, dw, db, S, Co, Ci,
                                                                      Cip, ndb, nblk_w);
    return ecg_launch_status();
}
