// Batched all-class Grad-CAM in closed form (no autograd, no backward pass), plus the
// per-lead z-score of the input pipeline.
//
// Replaces the per-(sample, class) forward + full backward of
// /root/reference/src/interpretability/grad_cam_1d.py:53-103 and its script clones
// (scripts/00_demo_inference.py:39-61, 12_grad_cam_ecg_demo.py:44-75, 13_grad_cam_af.py:51-76).
// In eval mode dScore_k/dA[ch,t] = v_k[ch] * s[ch] * mask[ch,t] / Lp, with
// mask = relu-active & first-argmax of its pool pair, so the channel weights are
// w_k[ch] = v_k[ch] * s[ch] * count[ch] / (Lp * L') and every class shares count[].
#include "common.cuh"

constexpr int GC_MAXK = 8;

__device__ __forceinline__ float upsample_at(const float* __restrict__ cam, int Lq, int d, float scale) {
    // F.interpolate(mode='linear', align_corners=False): src = max((d+0.5)*scale-0.5, 0)
    float src = fmaf((float)d + 0.5f, scale, -0.5f);
    src = fmaxf(src, 0.f);
    const int i0 = min((int)src, Lq - 1);
    const int i1 = min(i0 + 1, Lq - 1);
    const float lam = src - (float)i0;
    return (1.0f - lam) * cam[i0] + lam * cam[i1];
}

// (value, index) block arg-reductions; ties -> smallest index
__device__ __forceinline__ void block_minmax(float& mn, float& mx, int& amax, float* shf, int* shi) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float omn = __shfl_xor_sync(0xffffffffu, mn, o);
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oi = __shfl_xor_sync(0xffffffffu, amax, o);
        mn = fminf(mn, omn);
        if (omx > mx || (omx == mx && oi < amax)) { mx = omx; amax = oi; }
    }
    __syncthreads();
    if (lane == 0) { shf[w] = mn; shf[32 + w] = mx; shi[w] = amax; }
    __syncthreads();
    if (w == 0) {
        float a = lane < nw ? shf[lane] : INFINITY;
        float b = lane < nw ? shf[32 + lane] : -INFINITY;
        int ai = lane < nw ? shi[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oa = __shfl_xor_sync(0xffffffffu, a, o);
            const float ob = __shfl_xor_sync(0xffffffffu, b, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ai, o);
            a = fminf(a, oa);
            if (ob > b || (ob == b && oi < ai)) { b = ob; ai = oi; }
        }
        if (lane == 0) { shf[64] = a; shf[65] = b; shi[32] = ai; }
    }
    __syncthreads();
    mn = shf[64]; mx = shf[65]; amax = shi[32];
}

// One block (256 threads) per sample.  dynamic smem: wk[K][C] | cnt[C] | cam[K][Lq]
__global__ void __launch_bounds__(256)
gradcam_kernel(const float* __restrict__ A, const float* __restrict__ bn_state,
               const float* __restrict__ v, int v_per_sample, float* __restrict__ cam_lo,
               float* __restrict__ cam_hi, int32_t* __restrict__ argmax_out, int C, int Lq, int K,
               int T, int variant, float eps) {
    extern __shared__ float smem[];
    __shared__ float shf[66];
    __shared__ int shi[33];
    float* wk = smem;                  // K*C
    float* cnt = wk + K * C;           // C
    float* cam = cnt + C;              // K*Lq
    const int n = blockIdx.x;
    const float* An = A + (size_t)n * C * Lq;
    const int Lp = Lq / 2;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;

    // phase 1: per-channel count of positions that receive gradient through ReLU+MaxPool
    // (skipped when bn_state == NULL: v already holds the final channel weights mean_t dY/dA)
    for (int ch = w; bn_state != nullptr && ch < C; ch += nw) {
        const float sc = __ldg(bn_state + 2 * C + ch), sh = __ldg(bn_state + 3 * C + ch);
        const float* ar = An + (size_t)ch * Lq;
        int c = 0;
        for (int j = lane; j < Lp; j += 32) {
            const float r0 = fmaxf(fmaf(__ldg(ar + 2 * j), sc, sh), 0.f);
            const float r1 = fmaxf(fmaf(__ldg(ar + 2 * j + 1), sc, sh), 0.f);
            c += ((r0 >= r1 && r0 > 0.f) ? 1 : 0) + ((r1 > r0) ? 1 : 0);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[ch] = (float)c;
    }
    __syncthreads();
    // phase 2: channel weights per class
    const float* vn = v + (v_per_sample ? (size_t)n * K * C : 0);
    const float inv = 1.0f / ((float)Lp * (float)Lq);
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
        const int ch = i % C;
        wk[i] = bn_state != nullptr ? __ldg(vn + i) * __ldg(bn_state + 2 * C + ch) * cnt[ch] * inv
                                    : __ldg(vn + i);
    }
    __syncthreads();
    // phase 3: cam[k][t] = relu(sum_ch wk[k][ch] * A[ch][t])   (coalesced over t)
    for (int t = threadIdx.x; t < Lq; t += blockDim.x) {
        float acc[GC_MAXK];
#pragma unroll
        for (int k = 0; k < GC_MAXK; ++k) acc[k] = 0.f;
        for (int ch = 0; ch < C; ++ch) {
            const float a = __ldg(An + (size_t)ch * Lq + t);
#pragma unroll
            for (int k = 0; k < GC_MAXK; ++k)
                if (k < K) acc[k] = fmaf(wk[k * C + ch], a, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < GC_MAXK; ++k)
            if (k < K) cam[k * Lq + t] = fmaxf(acc[k], 0.f);
    }
    __syncthreads();
    // phase 4: normalise / upsample / argmax per class
    const bool up = (T > 0 && T != Lq);
    const float scale = up ? (float)Lq / (float)T : 1.f;
    for (int k = 0; k < K; ++k) {
        float* ck = cam + k * Lq;
        if (variant == 1) {
            float mn = INFINITY, mx = -INFINITY; int am = 0x7fffffff;
            for (int t = threadIdx.x; t < Lq; t += blockDim.x) {
                const float x = ck[t];
                mn = fminf(mn, x);
                if (x > mx) { mx = x; am = t; }
            }
            block_minmax(mn, mx, am, shf, shi);
            const float rng = mx - mn;
            for (int t = threadIdx.x; t < Lq; t += blockDim.x) {
                float x = ck[t] - mn;
                if (rng > 0.f) x = x / rng;
                ck[t] = x;
                if (cam_lo != nullptr) cam_lo[((size_t)n * K + k) * Lq + t] = x;
            }
            __syncthreads();
            if (up) {
                float mn2 = INFINITY, mx2 = -INFINITY; int am2 = 0x7fffffff;
                for (int d = threadIdx.x; d < T; d += blockDim.x) {
                    const float x = upsample_at(ck, Lq, d, scale);
                    if (cam_hi != nullptr) cam_hi[((size_t)n * K + k) * T + d] = x;
                    if (x > mx2) { mx2 = x; am2 = d; }
                    mn2 = fminf(mn2, x);
                }
                block_minmax(mn2, mx2, am2, shf, shi);
                am = am2;
            }
            if (threadIdx.x == 0 && argmax_out != nullptr) argmax_out[(size_t)n * K + k] = am;
        } else {
            // variant 2: upsample first, then (cam - min) / (max + eps) over the upsampled map
            const int Lout = up ? T : Lq;
            float mn = INFINITY, mx = -INFINITY; int am = 0x7fffffff;
            for (int d = threadIdx.x; d < Lout; d += blockDim.x) {
                const float x = up ? upsample_at(ck, Lq, d, scale) : ck[d];
                mn = fminf(mn, x);
                if (x > mx) { mx = x; am = d; }
            }
            block_minmax(mn, mx, am, shf, shi);
            const float den = (mx - mn) + eps;
            for (int d = threadIdx.x; d < Lout; d += blockDim.x) {
                const float x = ((up ? upsample_at(ck, Lq, d, scale) : ck[d]) - mn) / den;
                if (up) { if (cam_hi != nullptr) cam_hi[((size_t)n * K + k) * T + d] = x; }
                else if (cam_lo != nullptr) cam_lo[((size_t)n * K + k) * Lq + d] = x;
            }
            if (threadIdx.x == 0 && argmax_out != nullptr) argmax_out[(size_t)n * K + k] = am;
        }
        __syncthreads();
    }
}

extern "C" int ecgb200_gradcam_f32(const float* A, const float* bn_state, const float* v,
                                   int v_per_sample, float* cam_lo, float* cam_hi, int32_t* argmax,
                                   int B, int C, int Lq, int K, int T, int variant, float eps,
                                   void* stream) {
    if (!A || !v || B <= 0 || C <= 0 || Lq < 2 || K <= 0) return ECGB200_EINVAL;
    if (K > GC_MAXK || (variant != 1 && variant != 2)) return ECGB200_EUNSUPPORTED;
    const size_t smem = ((size_t)K * C + C + (size_t)K * Lq) * sizeof(float);
    if (smem > 200 * 1024) return ECGB200_EUNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(gradcam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gradcam_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(A, bn_state, v, v_per_sample, cam_lo, cam_hi,
                                                          argmax, C, Lq, K, T, variant, eps);
    return ecg_launch_status();
}

// ---------------------------------------------------------------- per-lead z-score (N1)
// out = (x - mean) / (std + 1e-6), population std over time; one block per (sample, lead) row.
// /root/reference/src/datasets/ptbxl.py:122-127.
__global__ void __launch_bounds__(256)
zscore_kernel(const float* __restrict__ x, float* __restrict__ out, int T) {
    __shared__ double sh[33];
    const float* xr = x + (size_t)blockIdx.x * T;
    float* orow = out + (size_t)blockIdx.x * T;
    double s = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) s += (double)xr[t];
    const double mean = block_sum_d(s, sh) / (double)T;
    double m2 = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) { const double d = (double)xr[t] - mean; m2 += d * d; }
    const double var = block_sum_d(m2, sh) / (double)T;
    const float mf = (float)mean;
    const float inv = 1.0f / ((float)sqrt(var) + 1e-6f);
    for (int t = threadIdx.x; t < T; t += blockDim.x) orow[t] = (xr[t] - mf) * inv;
}

extern "C" int ecgb200_zscore_f32(const float* x, float* out, int rows, int T, void* stream) {
    if (!x || !out || rows <= 0 || T <= 0) return ECGB200_EINVAL;
    zscore_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, out, T);
    return ecg_launch_status();
}

// out[r] = mean_t x[r, t]; one warp per row (channel weights = mean_t dY/dA, grad_cam_1d.py:85)
__global__ void row_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int L) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int t = lane; t < L; t += 32) s += __ldg(x + (size_t)row * L + t);
    s = warp_sum(s);
    if (lane == 0) out[row] = s / (float)L;
}

extern "C" int ecgb200_row_mean_f32(const float* x, float* out, int rows, int L, void* stream) {
    if (!x || !out || rows <= 0 || L <= 0) return ECGB200_EINVAL;
    row_mean_kernel<<<ecg_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, out, rows, L);
    return ecg_launch_status();
}

int g_ecg_pdl = 0;
extern "C" int ecgb200_set_pdl(int on) { const int old = g_ecg_pdl; g_ecg_pdl = on ? 1 : 0; return old; }
extern "C" int ecgb200_version(void) { return 100; }
extern "C" int ecgb200_arch(void) { return 1000; }

// ---------------------------------------------------------------- WFDB format-16 decode + per-lead z-score (N2)
// PTB-XL records are WFDB format 16: little-endian int16, sample-interleaved (one 2*n_leads-byte frame per time
// step).  The reference reads them with wfdb.rdsamp (float64 physical = (digital - baseline) / gain, -32768 = NaN),
// casts to float32, transposes to [leads, T] and z-scores each lead (src/datasets/ptbxl.py:25-29,122-127).  This
// kernel does all of it on the device from the raw bytes: one block per record, thread t walks frames t, t+256, ...
constexpr int WF_MAXL = 16;

// Per-lead statistics of one record from the RAW integers (exact): every thread sums its frames' samples (int64 sum and
// sum of squares per lead, -32768 = missing), one shuffle + shared-memory reduction per block, then in double
//   mean_d = S1 / T,  var_d = S2 / T - mean_d^2          (digital units; S2 < 2^53 is exact in double)
// and the affine map to physical units (x = (d - baseline) / gain) gives mean = (mean_d - baseline) / gain,
// std = sqrt(var_d) / |gain|.  Results per lead in shared memory: mean (physical, fp32), inv = 1 / (std + 1e-6),
// and for the fused pack kernel the integer / fractional split of mean_d with the combined scale 1 / (gain * (std + 1e-6)).
// This replaces three fp64 passes (sum, centred M2, output) with one integer pass: 63.7 -> ~4 us for 256 x 12 x 1000.
struct WfStats {
    float mean[WF_MAXL], inv[WF_MAXL];       // physical-unit mean, 1 / (std + 1e-6); NaN for a lead with a missing sample
    int mi[WF_MAXL];                         // round(mean_d)
    float mf[WF_MAXL], sc[WF_MAXL];          // mean_d - mi, 1 / (gain * (std + 1e-6))
};
template <typename LoadFrame>
__device__ __forceinline__ void wf_block_stats(LoadFrame load, int n_leads, int T, const float* __restrict__ gain,
                                               const int* __restrict__ baseline, WfStats* S,
                                               long long (*red)[2 * WF_MAXL + 1] /* [8][2*WF_MAXL+1] */) {
    long long s1[WF_MAXL], s2[WF_MAXL];
    int bad = 0;
    {
        int a1[WF_MAXL];                                 // |sum| <= 32768 * frames per thread: int32 up to 65535 frames
        unsigned long long a2[WF_MAXL];
#pragma unroll
        for (int l = 0; l < WF_MAXL; ++l) { a1[l] = 0; a2[l] = 0ull; }
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            short v[WF_MAXL];
            load(t, v);
#pragma unroll
            for (int l = 0; l < WF_MAXL; ++l)
                if (l < n_leads) {
                    const int d = (int)v[l];
                    bad |= (d == -32768) ? (1 << l) : 0;
                    a1[l] += d;
                    a2[l] += (unsigned)(d * d);
                }
        }
#pragma unroll
        for (int l = 0; l < WF_MAXL; ++l) { s1[l] = a1[l]; s2[l] = (long long)a2[l]; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int l = 0; l < WF_MAXL; ++l) {
        if (l < n_leads) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1[l] += __shfl_xor_sync(0xffffffffu, s1[l], o);
                s2[l] += __shfl_xor_sync(0xffffffffu, s2[l], o);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    if (lane == 0) {
#pragma unroll
        for (int l = 0; l < WF_MAXL; ++l) { red[w][2 * l] = s1[l]; red[w][2 * l + 1] = s2[l]; }
        red[w][2 * WF_MAXL] = bad;
    }
    __syncthreads();
    if (threadIdx.x < n_leads) {
        const int l = threadIdx.x, nw = blockDim.x >> 5;
        long long a = 0, b = 0, bd = 0;
        for (int j = 0; j < nw; ++j) { a += red[j][2 * l]; b += red[j][2 * l + 1]; bd |= red[j][2 * WF_MAXL]; }
        const double g = (double)gain[l];
        const double mean_d = (double)a / (double)T;
        double var_d = (double)b / (double)T - mean_d * mean_d;
        if (var_d < 0.0) var_d = 0.0;
        const float stdp = (float)(sqrt(var_d) / fabs(g));
        const float inv = 1.0f / (stdp + 1e-6f);
        const bool nan = ((bd >> l) & 1) != 0;
        const float qnan = __int_as_float(0x7fc00000);
        S->mean[l] = nan ? qnan : (float)((mean_d - (double)baseline[l]) / g);
        S->inv[l] = nan ? qnan : inv;
        const double mr = rint(mean_d);
        S->mi[l] = (int)mr;
        S->mf[l] = (float)(mean_d - mr);
        S->sc[l] = nan ? qnan : (float)((double)inv / g);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
wfdb16_zscore_kernel(const short* __restrict__ dat, const float* __restrict__ gain, const int* __restrict__ baseline,
                     float* __restrict__ out, int n_leads, int T, int normalize) {
    __shared__ long long red[8][2 * WF_MAXL + 1];
    __shared__ WfStats S;
    const short* rec = dat + (size_t)blockIdx.x * T * n_leads;
    float* orec = out + (size_t)blockIdx.x * n_leads * T;
    if (normalize)
        wf_block_stats([&](int t, short* v) {
#pragma unroll
            for (int l = 0; l < WF_MAXL; ++l) v[l] = l < n_leads ? rec[(size_t)t * n_leads + l] : (short)0;
        }, n_leads, T, gain, baseline, &S, red);
    for (int l = 0; l < n_leads; ++l) {
        const double g = (double)gain[l];
        const int bl = baseline[l];
        const float m = normalize ? S.mean[l] : 0.f, inv = normalize ? S.inv[l] : 1.f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const int d = (int)rec[(size_t)t * n_leads + l];
            // the decode itself stays the reference's: float32((digital - baseline) / gain) in double (bit-exact)
            const float ph = d == -32768 ? __int_as_float(0x7fc00000) : (float)((double)(d - bl) / g);
            orec[(size_t)l * T + t] = (ph - m) * inv;
        }
    }
}

// dat: (B, T, n_leads) int16 frames as stored in the .dat file; gain (ADC units per physical unit) and baseline per
// lead from the .hea header; out: (B, n_leads, T) fp32 = the tensor the reference's Dataset returns.
extern "C" int ecgb200_wfdb16_zscore_f32(const void* dat, const float* gain, const int* baseline, float* out, int B,
                                         int n_leads, int T, int normalize, void* stream) {
    if (!dat || !gain || !baseline || !out || B <= 0 || T <= 0) return ECGB200_EINVAL;
    if (n_leads <= 0 || n_leads > WF_MAXL) return ECGB200_EUNSUPPORTED;
    wfdb16_zscore_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const short*)dat, gain, baseline, out, n_leads, T, normalize);
    return ecg_launch_status();
}

// ---------------------------------------------------------------- raw frames -> the conv stack's input (N1 + N2 fused)
// The same decode + per-lead z-score, but the result goes straight to what the first tcgen05 conv reads: blocked
// channels-last bf16 xb[b][lead/8][t][8] with the leads zero-padded to a multiple of 16 -- the fp32 (B, leads, T) tensor
// of the reference's Dataset (src/datasets/ptbxl.py:122-127, 136-141) never exists.  One block per record: the raw frames
// are staged in shared memory once (coalesced 4-byte loads), every thread keeps the running sums of ALL leads of its
// frames (two passes in double: mean, then centred M2, as the oracle's numpy), one block reduction per pass.
#include <cuda_bf16.h>
// NL = compile-time lead count (12 for PTB-XL: the per-frame loops then carry no predicates), 0 = run-time n_leads.
template <int NL>
__global__ void __launch_bounds__(256)
wfdb16_zscore_pack_kernel(const short* __restrict__ dat, const float* __restrict__ gain, const int* __restrict__ baseline,
                          uint4* __restrict__ xb, int n_leads_rt, int Cp, int T) {
    extern __shared__ __align__(16) unsigned char wf_smem[];
    short* fr = reinterpret_cast<short*>(wf_smem);                       // [T][n_leads]
    __shared__ long long red[8][2 * WF_MAXL + 1];
    __shared__ WfStats S;
    const int n_leads = NL > 0 ? NL : n_leads_rt;
    const short* rec = dat + (size_t)blockIdx.x * T * n_leads;
    const int nbytes = T * n_leads * 2;
    if ((nbytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(rec) & 15) == 0)) {
        for (int i = threadIdx.x; i < (nbytes >> 4); i += blockDim.x)
            reinterpret_cast<uint4*>(fr)[i] = __ldg(reinterpret_cast<const uint4*>(rec) + i);
    } else {
        for (int i = threadIdx.x; i < (nbytes >> 2); i += blockDim.x)        // host guarantees T * n_leads even
            reinterpret_cast<int*>(fr)[i] = __ldg(reinterpret_cast<const int*>(rec) + i);
    }
    __syncthreads();
    wf_block_stats([&](int t, short* v) {
        if (NL == 12) {                                                      // one frame = 24 bytes = three 8-byte words
            const uint2* f2 = reinterpret_cast<const uint2*>(fr + t * 12);
            const uint2 a = f2[0], b = f2[1], c = f2[2];
            const unsigned w[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
#pragma unroll
            for (int l = 0; l < 6; ++l) { v[2 * l] = (short)(w[l] & 0xffffu); v[2 * l + 1] = (short)(w[l] >> 16); }
#pragma unroll
            for (int l = 12; l < WF_MAXL; ++l) v[l] = 0;
        } else {
#pragma unroll
            for (int l = 0; l < WF_MAXL; ++l) v[l] = l < n_leads ? fr[t * n_leads + l] : (short)0;
        }
    }, n_leads, T, gain, baseline, &S, red);
    // z = ((d - baseline) / gain - mean) / (std + 1e-6) = (d - mean_d) / (gain * (std + 1e-6)); d - mean_d is formed
    // exactly (integer part) + a small fraction, so fp32 keeps its full precision for large DC offsets
    const int nchunk = Cp / 8;
    uint4* out = xb + (size_t)blockIdx.x * nchunk * T;
    int mi[WF_MAXL];
    float mf[WF_MAXL], sc[WF_MAXL];
#pragma unroll
    for (int l = 0; l < WF_MAXL; ++l) {
        const bool on = l < n_leads;
        mi[l] = on ? S.mi[l] : 0; mf[l] = on ? S.mf[l] : 0.f; sc[l] = on ? S.sc[l] : 0.f;
    }
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        float v[WF_MAXL];
#pragma unroll
        for (int l = 0; l < WF_MAXL; ++l) {
            if (l < n_leads) {
                const int d = (int)fr[t * n_leads + l];
                v[l] = d == -32768 ? __int_as_float(0x7fc00000) : ((float)(d - mi[l]) - mf[l]) * sc[l];
            } else {
                v[l] = 0.f;
            }
        }
#pragma unroll
        for (int cc = 0; cc < WF_MAXL / 8; ++cc) {
            if (cc < nchunk) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * cc], v[8 * cc + 1]), h1 = __floats2bfloat162_rn(v[8 * cc + 2], v[8 * cc + 3]),
                               h2 = __floats2bfloat162_rn(v[8 * cc + 4], v[8 * cc + 5]), h3 = __floats2bfloat162_rn(v[8 * cc + 6], v[8 * cc + 7]);
                out[(size_t)cc * T + t] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                     *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            }
        }
    }
}

// dat (B, T, n_leads) int16 frames; xb [B][Cp/8][T][8] bf16 with Cp = n_leads rounded up to 16.
extern "C" int ecgb200_wfdb16_zscore_pack_bf16(const void* dat, const float* gain, const int* baseline, void* xb,
                                               int B, int n_leads, int T, void* stream) {
    if (!dat || !gain || !baseline || !xb || B <= 0 || T <= 0) return ECGB200_EINVAL;
    if (n_leads <= 0 || n_leads > WF_MAXL || ((T * n_leads) & 1) || (((uintptr_t)dat) & 3)) return ECGB200_EUNSUPPORTED;
    const size_t smem = (size_t)T * n_leads * sizeof(short);
    if (smem > 200 * 1024) return ECGB200_EUNSUPPORTED;
    const int Cp = (n_leads + 15) / 16 * 16;
    if (n_leads == 12) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(wfdb16_zscore_pack_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        wfdb16_zscore_pack_kernel<12><<<B, 256, smem, (cudaStream_t)stream>>>((const short*)dat, gain, baseline, (uint4*)xb,
                                                                             n_leads, Cp, T);
    } else {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(wfdb16_zscore_pack_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        wfdb16_zscore_pack_kernel<0><<<B, 256, smem, (cudaStream_t)stream>>>((const short*)dat, gain, baseline, (uint4*)xb,
                                                                            n_leads, Cp, T);
    }
    return ecg_launch_status();
}

