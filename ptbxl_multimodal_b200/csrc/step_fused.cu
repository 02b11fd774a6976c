// Fused per-step kernels of the bf16 TrainStep engine: they collapse the many tiny launches around
// the conv blocks (input packing + four weight re-layouts; GAP -> proj -> head -> BCE -> their
// backward; AdamW + step counter) into one launch each, because at batch 256 the step is bounded
// by launch latency, not by bytes or FLOPs.
//
// Replaces, from /root/reference: the aten::addmm pair of ECGCNN.forward (src/models/ecg_cnn.py:63-64),
// F.binary_cross_entropy_with_logits (src/training/loop.py:32) and their autograd backward
// (loop.py:33), and torch.optim.AdamW.step (loop.py:34).
#include "tc_common.cuh"

// ------------------------------------------------------------------ step prologue
struct PrepLayer {
    const float* w;            // fp32 master (Co, Ci, 15)
    __nv_bfloat16* wf;         // [15][Cip/8][Co][8]
    __nv_bfloat16* wd;         // [15][Co/8][Cip][8] tap-flipped transpose, or NULL
    int Co, Ci, Cip, n;        // n = 15 * Cip * Co
};
struct PrepArgs {
    const float* x;            // (B, Ci0, T) fp32
    uint4* xb;                 // [B][Cp0/8][T][8] bf16
    int B, Ci0, Cp0, T;
    long long n_pack;          // B * (Cp0/8) * T
    PrepLayer layer[4];
    const float* wp;           // proj weight (F, Cin) fp32, or NULL
    float* wpT;                // (Cin, F) transpose
    int F, Cin;
    int* step_ctr;             // incremented once per step (AdamW reads the new value)
};

// One launch: (a) x fp32 NCL -> blocked channels-last bf16 with the leads zero-padded to 16,
// (b) the four conv weights -> bf16 forward / dgrad operand layouts, (c) proj weight transpose for the
// fused head kernel, (d) optimizer step counter += 1.
__global__ void __launch_bounds__(256)
step_prep_kernel(const __grid_constant__ PrepArgs A) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0 && A.step_ctr != nullptr) A.step_ctr[0] += 1;
    for (long long i = tid; i < A.n_pack; i += nthreads) {
        const int t = (int)(i % A.T);
        const int cc = (int)((i / A.T) % (A.Cp0 / 8));
        const int b = (int)(i / ((long long)A.T * (A.Cp0 / 8)));
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cc * 8 + j;
            v[j] = c < A.Ci0 ? __ldg(A.x + ((size_t)b * A.Ci0 + c) * A.T + t) : 0.f;
        }
        A.xb[i] = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                             tc::pack_bf16(v[6], v[7]));
    }
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
        const PrepLayer& P = A.layer[l];
        if (P.w == nullptr) continue;
        // one thread per (o, c): its 15 taps are 60 contiguous bytes of the fp32 master; 8 neighbouring threads
        // (c % 8) fill one 16-byte unit of wf per tap
        const long long npair = (long long)P.Co * P.Cip;
        for (long long i = tid; i < npair; i += nthreads) {
            const int c = (int)(i % P.Cip), o = (int)(i / P.Cip);
            const float* src = P.w + ((size_t)o * P.Ci + c) * ECG_KS;
#pragma unroll
            for (int k = 0; k < ECG_KS; ++k) {
                const __nv_bfloat16 v = __float2bfloat16(c < P.Ci ? __ldg(src + k) : 0.f);
                P.wf[(((size_t)k * (P.Cip / 8) + (c >> 3)) * P.Co + o) * 8 + (c & 7)] = v;
                if (P.wd != nullptr)
                    P.wd[(((size_t)(ECG_KS - 1 - k) * (P.Co / 8) + (o >> 3)) * P.Cip + c) * 8 + (o & 7)] = v;
            }
        }
    }
    if (A.wp != nullptr) {
        const long long n = (long long)A.F * A.Cin;
        for (long long i = tid; i < n; i += nthreads) {          // i indexes wpT (coalesced writes)
            const int o = (int)(i % A.F), c = (int)(i / A.F);
            A.wpT[i] = __ldg(A.wp + (size_t)o * A.Cin + c);
        }
    }
}

// x (B,Ci0,T) fp32 -> xb; w[l] (co[l], ci[l], 15) -> wf[l] (+ wd[l] unless NULL) for l < nlayers <= 4;
// wp (F,Cin) -> wpT (Cin,F) unless NULL; *step_ctr += 1 unless NULL.  All arrays are HOST arrays.
extern "C" int ecgb200_step_prep_bf16(const float* x, void* xb, int B, int Ci0, int T, int nlayers,
                                      const float* const* w, void* const* wf, void* const* wd, const int* co,
                                      const int* ci, const float* wp, float* wpT, int F, int Cin,
                                      int* step_ctr, void* stream) {
    if (nlayers < 0 || nlayers > 4) return ECGB200_EINVAL;
    if (x != nullptr && (!xb || B <= 0 || Ci0 <= 0 || T <= 0)) return ECGB200_EINVAL;
    PrepArgs A;
    A.x = x; A.xb = (uint4*)xb; A.B = B; A.Ci0 = Ci0; A.Cp0 = (Ci0 + 15) / 16 * 16; A.T = T;
    A.n_pack = x != nullptr ? (long long)B * (A.Cp0 / 8) * T : 0;            // x == NULL: weights only
    long long work = A.n_pack > 0 ? A.n_pack : 1;
    for (int l = 0; l < 4; ++l) {
        PrepLayer& P = A.layer[l];
        if (l < nlayers) {
            if (!w || !wf || !wd || !co || !ci || !w[l] || !wf[l] || co[l] <= 0 || (co[l] & 7) || ci[l] <= 0) return ECGB200_EINVAL;
            P.w = w[l]; P.wf = (__nv_bfloat16*)wf[l]; P.wd = (__nv_bfloat16*)wd[l];
            P.Co = co[l]; P.Ci = ci[l]; P.Cip = (ci[l] + 15) / 16 * 16; P.n = ECG_KS * P.Cip * P.Co;
            if ((long long)P.Co * P.Cip > work) work = (long long)P.Co * P.Cip;
        } else {
            P.w = nullptr; P.wf = nullptr; P.wd = nullptr; P.Co = P.Ci = P.Cip = P.n = 0;
        }
    }
    A.wp = wp; A.wpT = wpT; A.F = F; A.Cin = Cin; A.step_ctr = step_ctr;
    if (wp != nullptr && (!wpT || F <= 0 || Cin <= 0)) return ECGB200_EINVAL;
    if (wp != nullptr && (long long)F * Cin > work) work = (long long)F * Cin;
    long long blocks = (work + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    step_prep_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ fused head: forward + loss + input gradients
// Per-window chain (no cross-window dependency): gap -> z = proj(gap) -> logits = head(z) -> BCE ->
// dlogits -> dz -> dgap.  One CTA handles HB windows with the 256-wide vectors in shared memory;
// weights are read coalesced (thread = output feature; proj forward uses the transposed copy).
constexpr int HB = 4;
constexpr int HMAXF = 256;      // feature widths up to 256 (the reference's feat_dim), labels up to 8
constexpr int HMAXL = 8;
constexpr int HT = 1024;        // threads: 4 K-quarters x 256 output features, every weight load in flight at once

// out[s][o] (+)= sum_{k in this thread's quarter} in[s][k] * W[k*ldw + o]; partials combined through shared memory
// `in` is stored window-minor ([k][HB]) so that the HB operands of one k are ONE 16-byte broadcast load.
__device__ __forceinline__ void head_gemv4(const float4* in, const float* __restrict__ W, int K, int N,
                                           float (*red)[HB][HMAXF], float* acc) {
    const int o = threadIdx.x & 255, kq = threadIdx.x >> 8;
    const int kper = (K + 3) >> 2, k0 = kq * kper, k1 = min(K, k0 + kper);
#pragma unroll
    for (int s = 0; s < HB; ++s) acc[s] = 0.f;
    if (o < N) {
#pragma unroll 16
        for (int k = k0; k < k1; ++k) {
            const float w = __ldg(W + (size_t)k * N + o);
            const float4 x = in[k];
            acc[0] = fmaf(x.x, w, acc[0]); acc[1] = fmaf(x.y, w, acc[1]);
            acc[2] = fmaf(x.z, w, acc[2]); acc[3] = fmaf(x.w, w, acc[3]);
        }
    }
#pragma unroll
    for (int s = 0; s < HB; ++s) red[kq][s][o] = acc[s];
    __syncthreads();
    if (kq == 0) {
#pragma unroll
        for (int s = 0; s < HB; ++s) acc[s] = red[0][s][o] + red[1][s][o] + red[2][s][o] + red[3][s][o];
    }
}

__global__ void __launch_bounds__(HT)
head_fwd_bwd_kernel(const float* __restrict__ gap, const float* __restrict__ wpT, const float* __restrict__ wp,
                    const float* __restrict__ bp, const float* __restrict__ wh, const float* __restrict__ bh,
                    const float* __restrict__ target, float* __restrict__ z, float* __restrict__ logits,
                    float* __restrict__ dlogits, float* __restrict__ dz, float* __restrict__ dgap,
                    float* __restrict__ loss_part, int B, int Cin, int F, int NL, float gscale) {
    static_assert(HB == 4, "head kernels keep the HB windows of one feature in a float4");
    __shared__ float4 gs4[HMAXF], dzs4[HMAXF];       // [feature][window]
    __shared__ float zs[HB][HMAXF];
    __shared__ float red[4][HB][HMAXF];
    __shared__ float dls[HB][HMAXL];
    __shared__ float lsum[HB * HMAXL];
    const int b0 = blockIdx.x * HB, nb = min(HB, B - b0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o = tid & 255;
    ecg_pdl_launch_dependents();
    ecg_pdl_wait();                                  // gap comes from the BatchNorm kernel launched just before
    for (int i = tid; i < HB * Cin; i += HT) {
        const int s = i / Cin, c = i - s * Cin;
        reinterpret_cast<float*>(gs4)[c * HB + s] = s < nb ? __ldg(gap + (size_t)(b0 + s) * Cin + c) : 0.f;
    }
    __syncthreads();
    float acc[HB];
    // z[s][o] = bp[o] + sum_c gap[s][c] * Wp[o][c]          (transposed copy: coalesced over o)
    head_gemv4(gs4, wpT, Cin, F, red, acc);
    if (tid < 256 && o < F) {
        const float bo = __ldg(bp + o);
#pragma unroll
        for (int s = 0; s < HB; ++s) {
            zs[s][o] = acc[s] + bo;
            if (s < nb) z[(size_t)(b0 + s) * F + o] = acc[s] + bo;
        }
    }
    __syncthreads();
    // logits, loss terms and dlogits: one warp per (window, label)
    const float inv_n = 1.0f / ((float)B * (float)NL);
    for (int idx = warp; idx < HB * NL; idx += HT / 32) {
        const int s = idx / NL, c = idx - s * NL;
        float a = 0.f;
        for (int k = lane; k < F; k += 32) a = fmaf(zs[s][k], __ldg(wh + (size_t)c * F + k), a);
        a = warp_sum(a);
        if (lane == 0) {
            float term = 0.f, dl = 0.f;
            if (s < nb) {
                const float x = a + __ldg(bh + c);
                const float y = __ldg(target + (size_t)(b0 + s) * NL + c);
                const float p = 1.0f / (1.0f + expf(-x));
                term = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
                dl = (p - y) * inv_n * gscale;
                logits[(size_t)(b0 + s) * NL + c] = x;
                dlogits[(size_t)(b0 + s) * NL + c] = dl;
            }
            dls[s][c] = dl;
            lsum[idx] = term;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < HB * NL; ++i) t += lsum[i];
        loss_part[blockIdx.x] = t;
    }
    // dz[s][o] = sum_c dlogits[s][c] * Wh[c][o]
    if (tid < 256 && o < F) {
#pragma unroll
        for (int s = 0; s < HB; ++s) acc[s] = 0.f;
        for (int c = 0; c < NL; ++c) {
            const float w = __ldg(wh + (size_t)c * F + o);
#pragma unroll
            for (int s = 0; s < HB; ++s) acc[s] = fmaf(dls[s][c], w, acc[s]);
        }
        dzs4[o] = make_float4(acc[0], acc[1], acc[2], acc[3]);
#pragma unroll
        for (int s = 0; s < HB; ++s)
            if (s < nb) dz[(size_t)(b0 + s) * F + o] = acc[s];
    }
    __syncthreads();
    // dgap[s][c] = sum_o dz[s][o] * Wp[o][c]                (coalesced over c)
    head_gemv4(dzs4, wp, F, Cin, red, acc);
    if (tid < 256 && o < Cin) {
#pragma unroll
        for (int s = 0; s < HB; ++s)
            if (s < nb) dgap[(size_t)(b0 + s) * Cin + o] = acc[s];
    }
}

extern "C" int ecgb200_head_loss_parts(int B) { return (B + HB - 1) / HB; }

// gap (B,Cin); wp (F,Cin), wpT (Cin,F) its transpose (ecgb200_step_prep_bf16), bp (F); wh (NL,F), bh (NL);
// target (B,NL).  Outputs z (B,F), logits / dlogits (B,NL), dz (B,F), dgap (B,Cin) and
// loss_part[ecgb200_head_loss_parts(B)] = per-CTA sums of the BCE terms (summed by head_wgrad).
extern "C" int ecgb200_head_fwd_bwd_f32(const float* gap, const float* wpT, const float* wp, const float* bp,
                                        const float* wh, const float* bh, const float* target, float* z,
                                        float* logits, float* dlogits, float* dz, float* dgap,
                                        float* loss_part, int B, int Cin, int F, int NL, float gscale,
                                        void* stream) {
    if (!gap || !wpT || !wp || !bp || !wh || !bh || !target || !z || !logits || !dlogits || !dz || !dgap ||
        !loss_part || B <= 0)
        return ECGB200_EINVAL;
    if (Cin <= 0 || Cin > HMAXF || F <= 0 || F > HMAXF || NL <= 0 || NL > HMAXL) return ECGB200_EUNSUPPORTED;
    return ecg_launch_pdl(head_fwd_bwd_kernel, dim3((B + HB - 1) / HB), dim3(HT), 0, (cudaStream_t)stream,
        gap, wpT, wp, bp, wh, bh, target, z, logits, dlogits, dz, dgap, loss_part, B, Cin, F, NL, gscale);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ fused head: weight gradients + loss
// Off the critical path (only AdamW consumes it): dWp = dz^T gap, dbp = colsum(dz), dWh = dlogits^T z,
// dbh = colsum(dlogits), loss = mean of the BCE terms.  Blocks [0, tiles): 32x32 tiles of dWp over
// K = B in a fixed order; block `tiles`: everything else.
__global__ void __launch_bounds__(256)
head_wgrad_kernel(const float* __restrict__ gap, const float* __restrict__ z, const float* __restrict__ dz,
                  const float* __restrict__ dlogits, const float* __restrict__ loss_part, int nparts,
                  float* __restrict__ dwp, float* __restrict__ dbp, float* __restrict__ dwh,
                  float* __restrict__ dbh, float* __restrict__ loss, int B, int Cin, int F, int NL,
                  int tiles_n) {
    __shared__ float As[32][33], Bs[32][33];
    const int tid = threadIdx.x;
    const int ntiles = tiles_n * ((F + 31) / 32);
    if ((int)blockIdx.x < ntiles) {
        const int o0 = ((int)blockIdx.x / tiles_n) * 32, c0 = ((int)blockIdx.x % tiles_n) * 32;
        const int tx = tid & 31, ty = tid >> 5;                  // 32 x 8: 4 outputs per thread
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float colsum = 0.f;                                      // dbp, by the blocks of tile column 0
        for (int k0 = 0; k0 < B; k0 += 32) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int k = k0 + ty + 8 * r;
                As[ty + 8 * r][tx] = (k < B && o0 + tx < F) ? __ldg(dz + (size_t)k * F + o0 + tx) : 0.f;
                Bs[ty + 8 * r][tx] = (k < B && c0 + tx < Cin) ? __ldg(gap + (size_t)k * Cin + c0 + tx) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) {
                const float bv = Bs[kk][tx];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[kk][ty + 8 * r], bv, acc[r]);
            }
            if (c0 == 0 && ty == 0)
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) colsum += As[kk][tx];
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int o = o0 + ty + 8 * r, c = c0 + tx;
            if (o < F && c < Cin) dwp[(size_t)o * Cin + c] = acc[r];
        }
        if (c0 == 0 && ty == 0 && o0 + tx < F) dbp[o0 + tx] = colsum;
        return;
    }
    // tail blocks: one per label (dWh row c), the last of them also dbh and the loss
    const int c = (int)blockIdx.x - ntiles;
    for (int o = tid; o < F; o += 256) {
        float acc = 0.f;
#pragma unroll 16
        for (int k = 0; k < B; ++k) acc = fmaf(__ldg(dlogits + (size_t)k * NL + c), __ldg(z + (size_t)k * F + o), acc);
        dwh[(size_t)c * F + o] = acc;
    }
    if (c == NL - 1) {
        if (tid < NL) {
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < B; ++k) acc += __ldg(dlogits + (size_t)k * NL + tid);
            dbh[tid] = acc;
        }
        if (tid == 32) {
            double t = 0.0;
#pragma unroll 8
            for (int i = 0; i < nparts; ++i) t += (double)__ldg(loss_part + i);
            *loss = (float)(t / ((double)B * (double)NL));
        }
    }
}

extern "C" int ecgb200_head_wgrad_f32(const float* gap, const float* z, const float* dz, const float* dlogits,
                                      const float* loss_part, float* dwp, float* dbp, float* dwh, float* dbh,
                                      float* loss, int B, int Cin, int F, int NL, void* stream) {
    if (!gap || !z || !dz || !dlogits || !loss_part || !dwp || !dbp || !dwh || !dbh || !loss || B <= 0)
        return ECGB200_EINVAL;
    if (Cin <= 0 || F <= 0 || NL <= 0 || NL > HMAXL) return ECGB200_EUNSUPPORTED;
    const int tiles_n = (Cin + 31) / 32;
    const int ntiles = tiles_n * ((F + 31) / 32);
    head_wgrad_kernel<<<ntiles + NL, 256, 0, (cudaStream_t)stream>>>(gap, z, dz, dlogits, loss_part,
                                                                    (B + HB - 1) / HB, dwp, dbp, dwh, dbh, loss,
                                                                    B, Cin, F, NL, tiles_n);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ AdamW over the flat parameter space
// Same update as ecgb200_adamw_f32 over ONE flat array, with t = *step_now (already incremented by
// ecgb200_step_prep_bf16), so the whole optimizer is a single launch in the captured step.
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, long long n, const float* __restrict__ hyper,
                  const int* __restrict__ step_now) {
    __shared__ float S[8];
    if (threadIdx.x == 0) {
        const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2], wd = hyper[4];
        const double step = (double)step_now[0];
        S[0] = (float)(1.0 - lr * wd);
        S[1] = (float)(1.0 - b1);
        S[2] = hyper[2];
        S[3] = (float)(1.0 - b2);
        S[4] = (float)sqrt(1.0 - pow(b2, step));
        S[5] = hyper[3];
        S[6] = (float)(lr / (1.0 - pow(b1, step)));
        S[7] = hyper[5];
    }
    __syncthreads();
    const float decay = S[0], one_m_b1 = S[1], b2 = S[2], one_m_b2 = S[3], bc2 = S[4], eps = S[5], ss = S[6],
                gscale = S[7];
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i],
               v4 = reinterpret_cast<float4*>(v)[i];
        const float gg[4] = {__fmul_rn(g4.x, gscale), __fmul_rn(g4.y, gscale), __fmul_rn(g4.z, gscale), __fmul_rn(g4.w, gscale)};
        float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
        const AdamK K = {decay, one_m_b1, b2, one_m_b2, bc2, eps, ss};
#pragma unroll
        for (int e = 0; e < 4; ++e) adamw_update(pp[e], mm[e], vv[e], gg[e], K);
        reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const AdamK K = {decay, one_m_b1, b2, one_m_b2, bc2, eps, ss};
        float pi = p[i], mi = m[i], vi = v[i];
        adamw_update(pi, mi, vi, __fmul_rn(g[i], gscale), K);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

// p, g, m, v: flat fp32 arrays of n elements, 16-byte aligned.  hyper as ecgb200_adamw_f32;
// step_now int[1] = the 1-based step index t (ecgb200_step_prep_bf16 increments it).
extern "C" int ecgb200_adamw_flat_f32(float* p, const float* g, float* m, float* v, int64_t n,
                                      const float* hyper, const int* step_now, void* stream) {
    if (!p || !g || !m || !v || n <= 0 || !hyper || !step_now) return ECGB200_EINVAL;
    if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return ECGB200_EINVAL;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;
    adamw_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper, step_now);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ fused head of the multimodal (FiLM) model
// Per-window chain of ECGMultimodal.forward (src/models/ecg_multimodal.py:88-99) + BCE + all input gradients:
//   h1 = relu(W0 d + b0); h2 = relu(W2 h1 + b2); film = Wf h2 + bf = [gamma_raw, beta];
//   z = Wp gap + bp; zc = (1 + tanh gamma_raw) * z + beta; logits = Wh zc + bh; BCE
//   dzc = Wh^T dl; dz = dzc (1 + tanh g); dfilm = [dzc z (1 - tanh^2 g), dzc];
//   dh2 = (Wf^T dfilm) [h2 > 0]; dh1 = (W2^T dh2) [h1 > 0]; dgap = Wp^T dz
// One CTA = HB windows; weight gradients are left to ecgb200_head_wgrad_multi_f32.
constexpr int MMH = 64;         // demo-encoder hidden width (<=)
constexpr int MMD = 8;          // demographic features (<=)

__global__ void __launch_bounds__(HT)
mm_head_fwd_bwd_kernel(const float* __restrict__ gap, const float* __restrict__ demo, const float* __restrict__ wpT,
                       const float* __restrict__ wp, const float* __restrict__ bp, const float* __restrict__ w0,
                       const float* __restrict__ b0v, const float* __restrict__ w2, const float* __restrict__ b2v,
                       const float* __restrict__ wf, const float* __restrict__ bfv, const float* __restrict__ wh,
                       const float* __restrict__ bh, const float* __restrict__ target, float* __restrict__ z,
                       float* __restrict__ h1g, float* __restrict__ h2g, float* __restrict__ filmg,
                       float* __restrict__ zcg, float* __restrict__ logits, float* __restrict__ dlogits,
                       float* __restrict__ dz, float* __restrict__ dfilmg, float* __restrict__ dh2g,
                       float* __restrict__ dh1g, float* __restrict__ dgap, float* __restrict__ loss_part, int B,
                       int Cin, int F, int D, int H, int NL, float gscale) {
    __shared__ float4 gs4[HMAXF], dzs4[HMAXF];
    __shared__ float zs[HB][HMAXF], zcs[HB][HMAXF];
    __shared__ float films[HB][2 * HMAXF];            // film, overwritten in place by dfilm in the backward half
    __shared__ float red[4][HB][HMAXF];
    __shared__ float ds[HB][MMD], h1s[HB][MMH], h2s[HB][MMH], dh2s[HB][MMH];
    __shared__ float dls[HB][HMAXL];
    __shared__ float lsum[HB * HMAXL];
    const int b0 = blockIdx.x * HB, nb = min(HB, B - b0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o = tid & 255;
    for (int i = tid; i < HB * Cin; i += HT) {
        const int s = i / Cin, c = i - s * Cin;
        reinterpret_cast<float*>(gs4)[c * HB + s] = s < nb ? __ldg(gap + (size_t)(b0 + s) * Cin + c) : 0.f;
    }
    if (tid < HB * D) {
        const int s = tid / D, i = tid - s * D;
        ds[s][i] = s < nb ? __ldg(demo + (size_t)(b0 + s) * D + i) : 0.f;
    }
    __syncthreads();
    // ---- demo encoder
    if (tid < HB * H) {
        const int s = tid / H, j = tid - s * H;
        float a = __ldg(b0v + j);
        for (int i = 0; i < D; ++i) a = fmaf(__ldg(w0 + (size_t)j * D + i), ds[s][i], a);
        a = fmaxf(a, 0.f);
        h1s[s][j] = a;
        if (s < nb) h1g[(size_t)(b0 + s) * H + j] = a;
    }
    __syncthreads();
    // h2 and film: one WARP per output row, lanes across the 64-wide reduction (a thread-per-row walk of the
    // (out, in) weight matrix touches 32 different lines per load instruction: measured 104 us for this kernel)
    for (int j = warp; j < H; j += HT / 32) {
        float a[HB] = {0.f, 0.f, 0.f, 0.f};
        for (int i = lane; i < H; i += 32) {
            const float w = __ldg(w2 + (size_t)j * H + i);
#pragma unroll
            for (int s = 0; s < HB; ++s) a[s] = fmaf(w, h1s[s][i], a[s]);
        }
#pragma unroll
        for (int s = 0; s < HB; ++s) a[s] = warp_sum(a[s]);
        if (lane < HB) {
            float v = 0.f;
#pragma unroll
            for (int s = 0; s < HB; ++s) if (s == lane) v = a[s];
            v = fmaxf(v + __ldg(b2v + j), 0.f);
            h2s[lane][j] = v;
            if (lane < nb) h2g[(size_t)(b0 + lane) * H + j] = v;
        }
    }
    __syncthreads();
    // ---- film = Wf h2 + bf   (four output rows per warp iteration: their weight loads are in flight together)
    for (int n0 = warp * 4; n0 < 2 * F; n0 += (HT / 32) * 4) {
        float a[4][HB];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int s = 0; s < HB; ++s) a[u][s] = 0.f;
        for (int k = lane; k < H; k += 32) {
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = (n0 + u < 2 * F) ? __ldg(wf + (size_t)(n0 + u) * H + k) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int s = 0; s < HB; ++s) a[u][s] = fmaf(w[u], h2s[s][k], a[u][s]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int s = 0; s < HB; ++s) a[u][s] = warp_sum(a[u][s]);
        if (lane < 4 * HB) {
            const int u = lane >> 2, sl = lane & 3;
            float v = 0.f;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu)
#pragma unroll
                for (int s = 0; s < HB; ++s) if (uu == u && s == sl) v = a[uu][s];
            const int n = n0 + u;
            if (n < 2 * F) {
                v += __ldg(bfv + n);
                films[sl][n] = v;
                if (sl < nb) filmg[(size_t)(b0 + sl) * 2 * F + n] = v;
            }
        }
    }
    // ---- z = Wp gap + bp
    float acc[HB];
    head_gemv4(gs4, wpT, Cin, F, red, acc);          // (contains a __syncthreads: films is complete after it)
    if (tid < 256 && o < F) {
        const float bo = __ldg(bp + o);
#pragma unroll
        for (int s = 0; s < HB; ++s) {
            const float zv = acc[s] + bo;
            const float th = tanhf(films[s][o]);
            const float zc = fmaf(1.0f + th, zv, films[s][F + o]);
            zs[s][o] = zv;
            zcs[s][o] = zc;
            if (s < nb) { z[(size_t)(b0 + s) * F + o] = zv; zcg[(size_t)(b0 + s) * F + o] = zc; }
        }
    }
    __syncthreads();
    // ---- logits, loss, dlogits
    const float inv_n = 1.0f / ((float)B * (float)NL);
    for (int idx = warp; idx < HB * NL; idx += HT / 32) {
        const int s = idx / NL, c = idx - s * NL;
        float a = 0.f;
        for (int k = lane; k < F; k += 32) a = fmaf(zcs[s][k], __ldg(wh + (size_t)c * F + k), a);
        a = warp_sum(a);
        if (lane == 0) {
            float term = 0.f, dl = 0.f;
            if (s < nb) {
                const float x = a + __ldg(bh + c);
                const float y = __ldg(target + (size_t)(b0 + s) * NL + c);
                const float p = 1.0f / (1.0f + expf(-x));
                term = fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
                dl = (p - y) * inv_n * gscale;
                logits[(size_t)(b0 + s) * NL + c] = x;
                dlogits[(size_t)(b0 + s) * NL + c] = dl;
            }
            dls[s][c] = dl;
            lsum[idx] = term;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < HB * NL; ++i) t += lsum[i];
        loss_part[blockIdx.x] = t;
    }
    // ---- dzc -> dz, dfilm
    if (tid < 256 && o < F) {
#pragma unroll
        for (int s = 0; s < HB; ++s) acc[s] = 0.f;
        for (int c = 0; c < NL; ++c) {
            const float w = __ldg(wh + (size_t)c * F + o);
#pragma unroll
            for (int s = 0; s < HB; ++s) acc[s] = fmaf(dls[s][c], w, acc[s]);
        }
        float dzv[HB];
#pragma unroll
        for (int s = 0; s < HB; ++s) {
            const float th = tanhf(films[s][o]);
            dzv[s] = acc[s] * (1.0f + th);
            const float dg = acc[s] * zs[s][o] * (1.0f - th * th);
            films[s][o] = dg;                           // (s, o) and (s, F + o) belong to this thread only
            films[s][F + o] = acc[s];
            if (s < nb) {
                dz[(size_t)(b0 + s) * F + o] = dzv[s];
                dfilmg[(size_t)(b0 + s) * 2 * F + o] = dg;
                dfilmg[(size_t)(b0 + s) * 2 * F + F + o] = acc[s];
            }
        }
        dzs4[o] = make_float4(dzv[0], dzv[1], dzv[2], dzv[3]);
    }
    __syncthreads();
    // ---- dh2 = (Wf^T dfilm) [h2 > 0]: 4 K-quarters x (HB x H) outputs, combined through `red`
    {
        const int q = tid >> 8, r = tid & 255;          // r -> (s, j)
        const int s = r / H, j = r - s * H;
        float a = 0.f;
        if (r < HB * H) {
            const int n0 = q * (2 * F / 4), n1 = n0 + 2 * F / 4;
#pragma unroll 16
            for (int n = n0; n < n1; ++n) a = fmaf(films[s][n], __ldg(wf + (size_t)n * H + j), a);
        }
        red[q][0][r] = a;
        __syncthreads();
        if (q == 0 && r < HB * H) {
            float t = red[0][0][r] + red[1][0][r] + red[2][0][r] + red[3][0][r];
            t = h2s[s][j] > 0.f ? t : 0.f;
            dh2s[s][j] = t;
            if (s < nb) dh2g[(size_t)(b0 + s) * H + j] = t;
        }
    }
    __syncthreads();
    // ---- dh1 = (W2^T dh2) [h1 > 0]
    if (tid < HB * H) {
        const int s = tid / H, i = tid - s * H;
        float a = 0.f;
#pragma unroll 8
        for (int j = 0; j < H; ++j) a = fmaf(dh2s[s][j], __ldg(w2 + (size_t)j * H + i), a);
        a = h1s[s][i] > 0.f ? a : 0.f;
        if (s < nb) dh1g[(size_t)(b0 + s) * H + i] = a;
    }
    // ---- dgap = Wp^T dz
    head_gemv4(dzs4, wp, F, Cin, red, acc);
    if (tid < 256 && o < Cin) {
#pragma unroll
        for (int s = 0; s < HB; ++s)
            if (s < nb) dgap[(size_t)(b0 + s) * Cin + o] = acc[s];
    }
}

extern "C" int ecgb200_mm_head_fwd_bwd_f32(const float* gap, const float* demo, const float* wpT, const float* wp,
                                           const float* bp, const float* w0, const float* b0, const float* w2,
                                           const float* b2, const float* wf, const float* bf, const float* wh,
                                           const float* bh, const float* target, float* z, float* h1, float* h2,
                                           float* film, float* zc, float* logits, float* dlogits, float* dz,
                                           float* dfilm, float* dh2, float* dh1, float* dgap, float* loss_part, int B,
                                           int Cin, int F, int D, int H, int NL, float gscale, void* stream) {
    if (!gap || !demo || !wpT || !wp || !bp || !w0 || !b0 || !w2 || !b2 || !wf || !bf || !wh || !bh || !target || !z ||
        !h1 || !h2 || !film || !zc || !logits || !dlogits || !dz || !dfilm || !dh2 || !dh1 || !dgap || !loss_part || B <= 0)
        return ECGB200_EINVAL;
    if (Cin <= 0 || Cin > HMAXF || F <= 0 || F > HMAXF || NL <= 0 || NL > HMAXL || D <= 0 || D > MMD || H <= 0 ||
        H > MMH || HB * H > 256)
        return ECGB200_EUNSUPPORTED;
    mm_head_fwd_bwd_kernel<<<(B + HB - 1) / HB, HT, 0, (cudaStream_t)stream>>>(
        gap, demo, wpT, wp, bp, w0, b0, w2, b2, wf, bf, wh, bh, target, z, h1, h2, film, zc, logits, dlogits, dz, dfilm,
        dh2, dh1, dgap, loss_part, B, Cin, F, D, H, NL, gscale);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ weight gradients of several small Linear layers
// dW_p[n][k] = sum_s A_p[s][n] * X_p[s][k],  db_p[n] = sum_s A_p[s][n]   for up to 6 problems in ONE launch
// (32x32 output tiles over K = B in a fixed order), plus loss = sum(loss_part) / (B * NL).
constexpr int HW_MAXP = 6;
struct HeadWgradProblems {
    const float* a[HW_MAXP];     // (B, N) upstream gradient
    const float* x[HW_MAXP];     // (B, K) layer input
    float* dw[HW_MAXP];          // (N, K)
    float* db[HW_MAXP];          // (N) or NULL
    int n[HW_MAXP], k[HW_MAXP], tile0[HW_MAXP + 1];
    int nprob;
};

__global__ void __launch_bounds__(256)
head_wgrad_multi_kernel(const __grid_constant__ HeadWgradProblems Q, const float* __restrict__ loss_part, int nparts,
                        float* __restrict__ loss, int B, int NL) {
    __shared__ float As[32][33], Bs[32][33];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x == Q.tile0[Q.nprob]) {               // extra block: the scalar loss
        if (tid == 0 && loss != nullptr) {
            double t = 0.0;
            for (int i = 0; i < nparts; ++i) t += (double)__ldg(loss_part + i);
            *loss = (float)(t / ((double)B * (double)NL));
        }
        return;
    }
    int p = 0;
    while (p + 1 < Q.nprob && (int)blockIdx.x >= Q.tile0[p + 1]) ++p;
    const int N = Q.n[p], K = Q.k[p];
    const int tiles_k = (K + 31) / 32;
    const int t = (int)blockIdx.x - Q.tile0[p];
    const int n0 = (t / tiles_k) * 32, k0 = (t % tiles_k) * 32;
    const float* __restrict__ A = Q.a[p];
    const float* __restrict__ X = Q.x[p];
    const int tx = tid & 31, ty = tid >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float colsum = 0.f;
    for (int s0 = 0; s0 < B; s0 += 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int s = s0 + ty + 8 * r;
            As[ty + 8 * r][tx] = (s < B && n0 + tx < N) ? __ldg(A + (size_t)s * N + n0 + tx) : 0.f;
            Bs[ty + 8 * r][tx] = (s < B && k0 + tx < K) ? __ldg(X + (size_t)s * K + k0 + tx) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const float bv = Bs[kk][tx];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[kk][ty + 8 * r], bv, acc[r]);
        }
        if (k0 == 0 && ty == 0)
#pragma unroll
            for (int kk = 0; kk < 32; ++kk) colsum += As[kk][tx];
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int n = n0 + ty + 8 * r, k = k0 + tx;
        if (n < N && k < K) Q.dw[p][(size_t)n * K + k] = acc[r];
    }
    if (k0 == 0 && ty == 0 && n0 + tx < N && Q.db[p] != nullptr) Q.db[p][n0 + tx] = colsum;
}

// a / x / dw / db / n / k: HOST arrays of nprob <= 6 entries.  loss_part may be NULL (then loss is not written).
extern "C" int ecgb200_head_wgrad_multi_f32(int nprob, const float* const* a, const float* const* x, float* const* dw,
                                            float* const* db, const int* n, const int* k, const float* loss_part,
                                            float* loss, int B, int NL, void* stream) {
    if (nprob <= 0 || nprob > HW_MAXP || !a || !x || !dw || !db || !n || !k || B <= 0) return ECGB200_EINVAL;
    HeadWgradProblems Q;
    int tiles = 0;
    for (int p = 0; p < HW_MAXP; ++p) {
        if (p < nprob) {
            if (!a[p] || !x[p] || !dw[p] || n[p] <= 0 || k[p] <= 0) return ECGB200_EINVAL;
            Q.a[p] = a[p]; Q.x[p] = x[p]; Q.dw[p] = dw[p]; Q.db[p] = db[p]; Q.n[p] = n[p]; Q.k[p] = k[p];
            Q.tile0[p] = tiles;
            tiles += ((n[p] + 31) / 32) * ((k[p] + 31) / 32);
        } else {
            Q.a[p] = Q.x[p] = nullptr; Q.dw[p] = Q.db[p] = nullptr; Q.n[p] = Q.k[p] = 0; Q.tile0[p] = tiles;
        }
    }
    Q.tile0[HW_MAXP] = tiles;
    for (int p = nprob; p <= HW_MAXP; ++p) Q.tile0[p] = tiles;
    Q.nprob = nprob;
    const int extra = (loss_part != nullptr && loss != nullptr) ? 1 : 0;
    head_wgrad_multi_kernel<<<tiles + extra, 256, 0, (cudaStream_t)stream>>>(Q, loss_part, (B + HB - 1) / HB, loss, B, NL);
    return ecg_launch_status();
}
