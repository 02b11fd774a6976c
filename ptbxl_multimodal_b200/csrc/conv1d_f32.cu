// fp32 (CUDA-core) Conv1d k=15 pad=7: forward / dgrad (same kernel, re-laid-out weights)
// and wgrad.  This is the exact-arithmetic path (fp32 FFMA, fp32 accumulate) that carries
// the <=1e-4 parity mode; the bf16 tcgen05 path lives in conv1d_tc.cu.
//
// Replaces aten::convolution / aten::convolution_backward reached from
// nn.Conv1d at /root/reference/src/models/ecg_cnn.py:13.
#include "common.cuh"

// ------------------------------------------------------------------ weight prep
__global__ void prep_weights_kernel(const float* __restrict__ w, float* __restrict__ w_fwd,
                                    float* __restrict__ w_dgr, int Co, int Ci) {
    const int n = Co * Ci * ECG_KS;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i % ECG_KS;
        const int c = (i / ECG_KS) % Ci;
        const int o = i / (ECG_KS * Ci);
        const float v = w[i];
        w_fwd[(c * ECG_KS + k) * Co + o] = v;
        if (w_dgr) w_dgr[(o * ECG_KS + (ECG_KS - 1 - k)) * Ci + c] = v;
    }
}

extern "C" int ecgb200_conv1d_prep_weights_f32(const float* w, float* w_fwd, float* w_dgr,
                                               int Co, int Ci, void* stream) {
    if (!w || !w_fwd || Co <= 0 || Ci <= 0) return ECGB200_EINVAL;
    const int n = Co * Ci * ECG_KS;
    prep_weights_kernel<<<ecg_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, w_fwd, w_dgr, Co, Ci);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ forward / dgrad
// Block tile: CO_TILE output channels x 128 time steps of one sample.
// Thread tile: 4 output channels x 8 consecutive time steps; per input channel the
// thread keeps a 22-sample sliding window of x in registers (15 taps + 7 extra outputs)
// so each smem word is read once per 4 output channels.
constexpr int CF_TT = 128;                 // time steps per block
constexpr int CF_CI = 8;                   // input channels staged per iteration
constexpr int CF_XS = 160;                 // padded row: idx(t) = t + 4*(t/32), t < 142

__device__ __forceinline__ int cf_xidx(int t) { return t + ((t >> 5) << 2); }

template <int CO_TILE>
__global__ void __launch_bounds__((CO_TILE / 4) * 16)
conv1d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                  const float* __restrict__ bias, float* __restrict__ y,
                  float* __restrict__ stat_part, int Ci, int Co, int L, int ntiles_total) {
    constexpr int NT = (CO_TILE / 4) * 16;
    __shared__ __align__(16) float xs[CF_CI * CF_XS];
    __shared__ __align__(16) float ws[CF_CI * ECG_KS * CO_TILE];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * CF_TT;
    const int co0 = blockIdx.y * CO_TILE;
    const int b = blockIdx.z;
    const float* xb = x + (size_t)b * Ci * L;
    const bool co_vec = ((Co & 3) == 0);

    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;

    for (int ci0 = 0; ci0 < Ci; ci0 += CF_CI) {
        // stage x[ci0:ci0+8, t0-7 : t0+135)
        for (int i = tid; i < CF_CI * (CF_TT + 2 * ECG_PAD); i += NT) {
            const int c = i / (CF_TT + 2 * ECG_PAD);
            const int tt = i - c * (CF_TT + 2 * ECG_PAD);
            const int t = t0 + tt - ECG_PAD;
            float v = 0.f;
            if (ci0 + c < Ci && t >= 0 && t < L) v = __ldg(xb + (size_t)(ci0 + c) * L + t);
            xs[c * CF_XS + cf_xidx(tt)] = v;
        }
        // stage wt[ci0:ci0+8, :, co0:co0+CO_TILE]
        if (co_vec) {
            for (int i = tid; i < CF_CI * ECG_KS * (CO_TILE / 4); i += NT) {
                const int o4 = i % (CO_TILE / 4);
                const int ck = i / (CO_TILE / 4);            // c*15 + k
                const int c = ck / ECG_KS;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int co = co0 + o4 * 4;
                if (ci0 + c < Ci && co < Co)
                    v = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(ci0 * ECG_KS + ck) * Co + co));
                *reinterpret_cast<float4*>(&ws[ck * CO_TILE + o4 * 4]) = v;
            }
        } else {
            for (int i = tid; i < CF_CI * ECG_KS * CO_TILE; i += NT) {
                const int o = i % CO_TILE;
                const int ck = i / CO_TILE;
                const int c = ck / ECG_KS;
                float v = 0.f;
                if (ci0 + c < Ci && co0 + o < Co) v = __ldg(wt + (size_t)(ci0 * ECG_KS + ck) * Co + co0 + o);
                ws[ck * CO_TILE + o] = v;
            }
        }
        __syncthreads();

        const int cmax = min(CF_CI, Ci - ci0);
#pragma unroll 1
        for (int c = 0; c < cmax; ++c) {
            float xr[24];
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                const float4 v = *reinterpret_cast<const float4*>(&xs[c * CF_XS + cf_xidx(8 * tx + 4 * m)]);
                xr[4 * m + 0] = v.x; xr[4 * m + 1] = v.y; xr[4 * m + 2] = v.z; xr[4 * m + 3] = v.w;
            }
#pragma unroll
            for (int k = 0; k < ECG_KS; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(&ws[(c * ECG_KS + k) * CO_TILE + 4 * ty]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[0][j] = fmaf(wv.x, xr[j + k], acc[0][j]);
                    acc[1][j] = fmaf(wv.y, xr[j + k], acc[1][j]);
                    acc[2][j] = fmaf(wv.z, xr[j + k], acc[2][j]);
                    acc[3][j] = fmaf(wv.w, xr[j + k], acc[3][j]);
                }
            }
        }
        __syncthreads();
    }

    // epilogue: bias, store, optional BatchNorm partial statistics
    const int tbase = t0 + 8 * tx;
    const int tile_cnt = min(CF_TT, L - t0);
    const int tile_id = b * gridDim.x + blockIdx.x;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const int co = co0 + 4 * ty + o;
        const bool co_ok = co < Co;
        const float bv = (bias != nullptr && co_ok) ? __ldg(bias + co) : 0.f;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc[o][j] += bv;
            if (tbase + j < L) s += acc[o][j];
        }
        if (co_ok) {
            float* yr = y + ((size_t)b * Co + co) * L + tbase;
            if ((L & 3) == 0 && tbase + 7 < L) {
                *reinterpret_cast<float4*>(yr) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
                *reinterpret_cast<float4*>(yr + 4) = make_float4(acc[o][4], acc[o][5], acc[o][6], acc[o][7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (tbase + j < L) yr[j] = acc[o][j];
            }
        }
        if (stat_part != nullptr) {
            // reduce over the 16 tx lanes that share (ty, o): lanes differ in bits 0..3
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            const float mean_t = s / (float)tile_cnt;
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (tbase + j < L) { const float d = acc[o][j] - mean_t; m2 = fmaf(d, d, m2); }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
            if (tx == 0 && co_ok) {
                stat_part[(size_t)co * ntiles_total + tile_id] = s;
                stat_part[((size_t)Co + co) * ntiles_total + tile_id] = m2;
            }
        }
    }
}

extern "C" int ecgb200_conv1d_stat_tiles(int B, int L) { return B * ecg_cdiv(L, CF_TT); }

extern "C" int ecgb200_conv1d_fwd_f32(const float* x, const float* wt, const float* bias, float* y,
                                      float* stat_part, int B, int Ci, int Co, int L, void* stream) {
    if (!x || !wt || !y || B <= 0 || Ci <= 0 || Co <= 0 || L <= 0) return ECGB200_EINVAL;
    if (B > 65535) return ECGB200_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = ecg_cdiv(L, CF_TT);
    const int ntot = B * tiles;
    if (Co % 64 == 0) {
        dim3 grid(tiles, Co / 64, B);
        conv1d_fwd_kernel<64><<<grid, 256, 0, st>>>(x, wt, bias, y, stat_part, Ci, Co, L, ntot);
    } else {
        dim3 grid(tiles, ecg_cdiv(Co, 32), B);
        conv1d_fwd_kernel<32><<<grid, 128, 0, st>>>(x, wt, bias, y, stat_part, Ci, Co, L, ntot);
    }
    return ecg_launch_status();
}

// ------------------------------------------------------------------ wgrad
// dW[o,c,k] = sum_{b,t} dy[b,o,t] x[b,c,t+k-7].  Block: 32 o x 16 c x 15 taps, looping over
// its share of (sample, 64-step time tile) work items; partial results go to scratch and a
// second kernel adds them in a fixed order (deterministic).
constexpr int WG_OT = 32, WG_CT = 16, WG_TT = 64;
constexpr int WG_DS = WG_TT + 4;               // dy smem row stride (floats)
constexpr int WG_XS = 84;                      // x smem row stride: 78 used, 84 % 32 == 20 -> LDS.128 conflict-free

__global__ void __launch_bounds__(128)
conv1d_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                    float* __restrict__ part_w, float* __restrict__ part_b,
                    int B, int Ci, int Co, int L) {
    __shared__ __align__(16) float dys[WG_OT * WG_DS];
    __shared__ __align__(16) float xs[WG_CT * WG_XS];
    const int tid = threadIdx.x;
    const int cl = tid & 15, og = tid >> 4;           // local c, o-group (4 channels)
    const int o0 = blockIdx.x * WG_OT, c0 = blockIdx.y * WG_CT;
    const int tiles_t = (L + WG_TT - 1) / WG_TT;
    const int items = B * tiles_t;

    float acc[4][ECG_KS];
    float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < ECG_KS; ++k) acc[o][k] = 0.f;

    for (int it = blockIdx.z; it < items; it += gridDim.z) {
        const int b = it / tiles_t;
        const int t0 = (it - b * tiles_t) * WG_TT;
        for (int i = tid; i < WG_OT * WG_TT; i += 128) {
            const int o = i / WG_TT, tt = i - o * WG_TT;
            float v = 0.f;
            if (o0 + o < Co && t0 + tt < L) v = __ldg(dy + ((size_t)b * Co + o0 + o) * L + t0 + tt);
            dys[o * WG_DS + tt] = v;
        }
        for (int i = tid; i < WG_CT * (WG_TT + 2 * ECG_PAD); i += 128) {
            const int c = i / (WG_TT + 2 * ECG_PAD), tt = i - c * (WG_TT + 2 * ECG_PAD);
            const int t = t0 + tt - ECG_PAD;
            float v = 0.f;
            if (c0 + c < Ci && t >= 0 && t < L) v = __ldg(x + ((size_t)b * Ci + c0 + c) * L + t);
            xs[c * WG_XS + tt] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int tc = 0; tc < WG_TT; tc += 8) {
            float xr[24], dr[4][8];
#pragma unroll
            for (int m = 0; m < 6; ++m) {
                // last float4 (m==5) reads 2 words past the 78 staged ones; still inside the row stride
                const float4 v = *reinterpret_cast<const float4*>(&xs[cl * WG_XS + tc + 4 * m]);
                xr[4 * m + 0] = v.x; xr[4 * m + 1] = v.y; xr[4 * m + 2] = v.z; xr[4 * m + 3] = v.w;
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float4 a = *reinterpret_cast<const float4*>(&dys[(4 * og + o) * WG_DS + tc]);
                const float4 c = *reinterpret_cast<const float4*>(&dys[(4 * og + o) * WG_DS + tc + 4]);
                dr[o][0] = a.x; dr[o][1] = a.y; dr[o][2] = a.z; dr[o][3] = a.w;
                dr[o][4] = c.x; dr[o][5] = c.y; dr[o][6] = c.z; dr[o][7] = c.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int k = 0; k < ECG_KS; ++k) {
                    acc[0][k] = fmaf(dr[0][j], xr[j + k], acc[0][k]);
                    acc[1][k] = fmaf(dr[1][j], xr[j + k], acc[1][k]);
                    acc[2][k] = fmaf(dr[2][j], xr[j + k], acc[2][k]);
                    acc[3][k] = fmaf(dr[3][j], xr[j + k], acc[3][k]);
                }
            if (cl == 0 && blockIdx.y == 0) {
#pragma unroll
                for (int o = 0; o < 4; ++o)
#pragma unroll
                    for (int j = 0; j < 8; ++j) accb[o] += dr[o][j];
            }
        }
        __syncthreads();
    }
    const size_t wsz = (size_t)Co * Ci * ECG_KS;
    if (c0 + cl < Ci) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int oc = o0 + 4 * og + o;
            if (oc < Co) {
                float* dst = part_w + blockIdx.z * wsz + ((size_t)oc * Ci + c0 + cl) * ECG_KS;
#pragma unroll
                for (int k = 0; k < ECG_KS; ++k) dst[k] = acc[o][k];
            }
        }
    }
    if (cl == 0 && blockIdx.y == 0) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int oc = o0 + 4 * og + o;
            if (oc < Co) part_b[(size_t)blockIdx.z * Co + oc] = accb[o];
        }
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b,
                                    float* __restrict__ dw, float* __restrict__ db,
                                    int nw, int Co, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nw) {
        float s = 0.f;
        for (int z = 0; z < S; ++z) s += part_w[(size_t)z * nw + i];
        dw[i] = s;
    } else if (i < nw + Co && db != nullptr) {
        const int o = i - nw;
        float s = 0.f;
        for (int z = 0; z < S; ++z) s += part_b[(size_t)z * Co + o];
        db[o] = s;
    }
}

static int wgrad_splits(int B, int Ci, int Co, int L) {
    const int blocks_oc = ecg_cdiv(Co, WG_OT) * ecg_cdiv(Ci, WG_CT);
    const int items = B * ecg_cdiv(L, WG_TT);
    int S = ecg_cdiv(148 * 4, blocks_oc);
    if (S > items) S = items;
    if (S < 1) S = 1;
    return S;
}

extern "C" size_t ecgb200_conv1d_wgrad_ws_bytes(int B, int Ci, int Co, int L) {
    const size_t S = (size_t)wgrad_splits(B, Ci, Co, L);
    return S * ((size_t)Co * Ci * ECG_KS + Co) * sizeof(float);
}

extern "C" int ecgb200_conv1d_wgrad_f32(const float* dy, const float* x, float* dw, float* db,
                                        void* ws, int B, int Ci, int Co, int L, void* stream) {
    if (!dy || !x || !dw || !ws || B <= 0 || Ci <= 0 || Co <= 0 || L <= 0) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int S = wgrad_splits(B, Ci, Co, L);
    const int nw = Co * Ci * ECG_KS;
    float* part_w = (float*)ws;
    float* part_b = part_w + (size_t)S * nw;
    dim3 grid(ecg_cdiv(Co, WG_OT), ecg_cdiv(Ci, WG_CT), S);
    conv1d_wgrad_kernel<<<grid, 128, 0, st>>>(dy, x, part_w, part_b, B, Ci, Co, L);
    int rc = ecg_launch_status();
    if (rc) return rc;
    wgrad_reduce_kernel<<<ecg_cdiv(nw + Co, 256), 256, 0, st>>>(part_w, part_b, dw, db, nw, Co, S);
    return ecg_launch_status();
}
