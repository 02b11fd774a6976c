// Inference-side kernels of the bf16 engine (InferStep): eval-mode BatchNorm folding, the proj-weight transpose
// and the fused head (GAP finalise -> proj -> [demo encoder -> FiLM] -> head -> sigmoid).
//
// Replaces, for model.eval() forward passes (/root/reference/src/training/loop.py:52-65, loop_demo.py:59-75,
// scripts/06_ecg_baseline_test.py:94-106): BatchNorm1d(eval) at src/models/ecg_cnn.py:14 (folded into the conv
// epilogue of conv_tc_kernel<1|2>), AdaptiveAvgPool1d + proj + head at ecg_cnn.py:61-64, DemoEncoder /
// film_gen / FiLM at src/models/ecg_multimodal.py:44-59,88-99 and torch.sigmoid at loop.py:63.
#include "common.cuh"

// scale = gamma / sqrt(running_var + eps);  shift = (conv_bias - running_mean) * scale + beta
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ conv_bias, float* __restrict__ scale,
                               float* __restrict__ shift, int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = gamma[c] / sqrtf(var[c] + eps);
    const float b = conv_bias != nullptr ? conv_bias[c] : 0.f;
    scale[c] = s;
    shift[c] = fmaf(b - mean[c], s, beta[c]);
}

extern "C" int ecgb200_bn_fold_f32(const float* gamma, const float* beta, const float* running_mean,
                                   const float* running_var, const float* conv_bias, float* scale, float* shift,
                                   int C, float eps, void* stream) {
    if (!gamma || !beta || !running_mean || !running_var || !scale || !shift || C <= 0) return ECGB200_EINVAL;
    bn_fold_kernel<<<ecg_cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var,
                                                                      conv_bias, scale, shift, C, eps);
    return ecg_launch_status();
}

// out[c][r] = in[r][c]
__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

extern "C" int ecgb200_transpose_f32(const float* in, float* out, int rows, int cols, void* stream) {
    if (!in || !out || rows <= 0 || cols <= 0) return ECGB200_EINVAL;
    transpose_f32_kernel<<<dim3(ecg_cdiv(cols, 32), ecg_cdiv(rows, 32)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
        in, out, rows, cols);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ fused inference head
// IH_W windows per CTA, 1024 threads (256 output features x 4 K-quarters).  Per window:
//   gap[c]   = inv_lp * sum_p gap_part[b][p][c]                       (fixed order: deterministic)
//   z[j]     = bp[j] + sum_k WpT[k][j] gap[k]                          (thread j, coalesced over j)
//   demo != NULL:  h1 = relu(W1 d + b1), h2 = relu(W2 h1 + b2), film = Wf h2 + bf   (w2, wf passed TRANSPOSED:
//                  (H_in, H_out) and (H, 2F), so that consecutive threads read consecutive addresses),
//                  z[j] <- (1 + tanh(film[j])) * z[j] + film[F + j]    (features out = un-modulated z, as the
//                  reference's ecg_backbone returns it)
//   logits[c] = bh[c] + sum_j Wh[c][j] z[j];   prob = sigmoid(logits)
constexpr int IH_W = 4;
constexpr int IH_MAXF = 256;     // C4, F <= 256
constexpr int IH_MAXH = 64;      // demo hidden width <= 64

struct InferHeadArgs {
    const float* gap_part; int nparts; float inv_lp;
    const float* wpT; const float* bp;
    const float* demo; const float* w1; const float* b1; const float* w2; const float* b2;
    const float* wf; const float* bf;
    const float* wh; const float* bh;
    float* z; float* logits; float* prob;
    int B, C4, F, D0, H, NL;
};

constexpr int IH_THREADS = 1024;   // 256 output features x 4 K-quarters

__global__ void __launch_bounds__(IH_THREADS) infer_head_kernel(const InferHeadArgs a) {
    __shared__ __align__(16) float gT[IH_MAXF][IH_W];      // gap
    __shared__ __align__(16) float zT[IH_MAXF][IH_W];      // features (then FiLM-modulated in place)
    __shared__ __align__(16) float part[4][IH_MAXF][IH_W]; // K-quarter partial sums of proj / film
    __shared__ __align__(16) float h1T[IH_MAXH][IH_W];
    __shared__ __align__(16) float h2T[IH_MAXH][IH_W];
    const int tid = threadIdx.x, b0 = blockIdx.x * IH_W;
    const int nw = min(IH_W, a.B - b0);
    const int j = tid & 255, kq = tid >> 8;                // output feature, K-quarter

    // gap: (channel, window) items over all threads; the partials of a window are summed in a fixed order
    for (int i = tid; i < a.C4 * IH_W; i += IH_THREADS) {
        const int c = i % a.C4, s = i / a.C4;
        float acc = 0.f;
        if (s < nw) {
            const float* p = a.gap_part + (size_t)(b0 + s) * a.nparts * a.C4 + c;
            for (int n = 0; n < a.nparts; ++n) acc += __ldg(p + (size_t)n * a.C4);
        }
        gT[c][s] = acc * a.inv_lp;
    }
    if (a.demo != nullptr) {
        for (int i = tid; i < a.H * IH_W; i += IH_THREADS) {
            const int s = i % IH_W, r = i / IH_W;
            float acc = a.b1[r];
            if (s < nw)
                for (int d = 0; d < a.D0; ++d) acc = fmaf(a.w1[r * a.D0 + d], a.demo[(size_t)(b0 + s) * a.D0 + d], acc);
            h1T[r][s] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    // proj: thread (j, kq) covers k in [kq*C4/4, (kq+1)*C4/4), weights fetched eight at a time (independent loads
    // in flight: the loop is bound by L2 latency, not by bandwidth or FLOPs)
    if (j < a.F) {
        float acc[IH_W] = {0.f, 0.f, 0.f, 0.f};
        const int kn = (a.C4 + 3) / 4, k0 = kq * kn, k1 = min(a.C4, k0 + kn);
        for (int k = k0; k < k1; k += 8) {
            float w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e] = k + e < k1 ? __ldg(a.wpT + (size_t)(k + e) * a.F + j) : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 g = *reinterpret_cast<const float4*>(gT[min(k + e, a.C4 - 1)]);
                acc[0] = fmaf(w[e], g.x, acc[0]); acc[1] = fmaf(w[e], g.y, acc[1]);
                acc[2] = fmaf(w[e], g.z, acc[2]); acc[3] = fmaf(w[e], g.w, acc[3]);
            }
        }
        *reinterpret_cast<float4*>(part[kq][j]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    if (a.demo != nullptr) {
        for (int i = tid; i < a.H * IH_W; i += IH_THREADS) {
            const int s = i % IH_W, r = i / IH_W;
            float acc = a.b2[r];
            for (int k = 0; k < a.H; ++k) acc = fmaf(__ldg(a.w2 + k * a.H + r), h1T[k][s], acc);
            h2T[r][s] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    if (kq == 0 && j < a.F) {
        const float bj = a.bp[j];
#pragma unroll
        for (int s = 0; s < IH_W; ++s) {
            const float v = bj + ((part[0][j][s] + part[1][j][s]) + (part[2][j][s] + part[3][j][s]));
            zT[j][s] = v;
            if (a.z != nullptr && s < nw) a.z[(size_t)(b0 + s) * a.F + j] = v;
        }
    }
    if (a.demo != nullptr) {
        __syncthreads();
        // film: thread (j, kq) -> gamma_j (kq 0,1) / beta_j (kq 2,3) over one half of H; wf is (H, 2F)
        if (j < a.F) {
            const int col = (kq >> 1) * a.F + j;
            const int hn = (a.H + 1) / 2, k0 = (kq & 1) * hn, k1 = min(a.H, k0 + hn);
            float acc[IH_W] = {0.f, 0.f, 0.f, 0.f};
            for (int k = k0; k < k1; k += 8) {
                float w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = k + e < k1 ? __ldg(a.wf + (size_t)(k + e) * 2 * a.F + col) : 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 h = *reinterpret_cast<const float4*>(h2T[min(k + e, a.H - 1)]);
                    acc[0] = fmaf(w[e], h.x, acc[0]); acc[1] = fmaf(w[e], h.y, acc[1]);
                    acc[2] = fmaf(w[e], h.z, acc[2]); acc[3] = fmaf(w[e], h.w, acc[3]);
                }
            }
            *reinterpret_cast<float4*>(part[kq][j]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
        if (kq == 0 && j < a.F) {
            const float bg = a.bf[j], bb = a.bf[a.F + j];
#pragma unroll
            for (int s = 0; s < IH_W; ++s) {
                const float ga = bg + (part[0][j][s] + part[1][j][s]);
                const float be = bb + (part[2][j][s] + part[3][j][s]);
                zT[j][s] = fmaf(1.0f + tanhf(ga), zT[j][s], be);
            }
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int item = warp; item < nw * a.NL; item += IH_THREADS / 32) {
        const int s = item % nw, c = item / nw;
        float acc = 0.f;
        for (int f = lane; f < a.F; f += 32) acc = fmaf(__ldg(a.wh + (size_t)c * a.F + f), zT[f][s], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            const float x = acc + a.bh[c];
            a.logits[(size_t)(b0 + s) * a.NL + c] = x;
            if (a.prob != nullptr) a.prob[(size_t)(b0 + s) * a.NL + c] = 1.0f / (1.0f + expf(-x));
        }
    }
}

extern "C" int ecgb200_infer_head_f32(const float* gap_part, int nparts, float inv_lp, const float* wpT,
                                      const float* bp, const float* demo, const float* w1, const float* b1,
                                      const float* w2, const float* b2, const float* wf, const float* bf,
                                      const float* wh, const float* bh, float* z, float* logits, float* prob,
                                      int B, int C4, int F, int D0, int H, int NL, void* stream) {
    if (!gap_part || !wpT || !bp || !wh || !bh || !logits || B <= 0 || nparts <= 0 || NL <= 0) return ECGB200_EINVAL;
    if (C4 <= 0 || C4 > IH_MAXF || F <= 0 || F > IH_MAXF) return ECGB200_EUNSUPPORTED;
    if (demo != nullptr) {
        if (!w1 || !b1 || !w2 || !b2 || !wf || !bf || D0 <= 0) return ECGB200_EINVAL;
        if (H <= 0 || H > IH_MAXH) return ECGB200_EUNSUPPORTED;
    }
    InferHeadArgs a{gap_part, nparts, inv_lp, wpT, bp, demo, w1, b1, w2, b2, wf, bf, wh, bh, z, logits, prob,
                    B, C4, F, D0, H, NL};
    infer_head_kernel<<<ecg_cdiv(B, IH_W), IH_THREADS, 0, (cudaStream_t)stream>>>(a);
    return ecg_launch_status();
}
