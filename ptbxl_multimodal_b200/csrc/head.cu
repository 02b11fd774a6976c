// Head-side kernels (all tiny, M = batch): Linear fwd/bwd, FiLM, sigmoid-BCE, AdamW.
//
// Replaces aten::addmm (nn.Linear, /root/reference/src/models/ecg_cnn.py:47,50;
// ecg_multimodal.py:52-54,85,86), tanh/mul/add (ecg_multimodal.py:93-96),
// binary_cross_entropy_with_logits (src/training/loop.py:32, loop_demo.py:33),
// sigmoid (loop.py:63) and torch.optim.AdamW.step (loop.py:34).
#include "common.cuh"

// ------------------------------------------------------------------ small strided SGEMM
// C[m][n] = act( sum_k A(m,k) * Bm(k,n) + bias[n] ), C row-major (ldc = N), with
// A(m,k) = A[m*sam + k*sak], Bm(k,n) = Bm[k*sbk + n*sbn]; each operand has one unit stride.
// 32x32 output tile, 256 threads (2x2 outputs each), K step 32.  Operand tiles are fetched
// with ONE 16-byte load per thread per operand (scalar fallback when shape/alignment forbid)
// and prefetched into registers one K-step ahead.  Deterministic (no split-K, fixed order).
constexpr int SG_T = 32, SG_K = 32;

// Tile fetch: `u` selects which index is unit-stride (1: the K index, 0: the M/N index).
// Thread t owns 4 consecutive elements along the unit-stride index.
struct SgTile { float v[4]; };

__device__ __forceinline__ SgTile sg_fetch(const float* __restrict__ P, int r0, int k0, int R, int K,
                                           long sr, long sk, int kunit, bool vec, int tid) {
    // r = row (M or N) index inside the tile, k = K index inside the tile
    SgTile t;
    int r, k;
    if (kunit) { r = tid >> 3; k = (tid & 7) * 4; } else { k = tid >> 3; r = (tid & 7) * 4; }
    const float* src = P + (long)(r0 + r) * sr + (long)(k0 + k) * sk;
    if (vec && r0 + r + (kunit ? 0 : 3) < R && k0 + k + (kunit ? 3 : 0) < K) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(src));
        t.v[0] = f.x; t.v[1] = f.y; t.v[2] = f.z; t.v[3] = f.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = r + (kunit ? 0 : i), kk = k + (kunit ? i : 0);
            t.v[i] = (r0 + rr < R && k0 + kk < K) ? __ldg(P + (long)(r0 + rr) * sr + (long)(k0 + kk) * sk) : 0.f;
        }
    }
    return t;
}

__device__ __forceinline__ void sg_store(float (*S)[SG_T + 1], const SgTile& t, int kunit, int tid) {
    int r, k;
    if (kunit) { r = tid >> 3; k = (tid & 7) * 4; } else { k = tid >> 3; r = (tid & 7) * 4; }
#pragma unroll
    for (int i = 0; i < 4; ++i) S[k + (kunit ? i : 0)][r + (kunit ? 0 : i)] = t.v[i];
}

__global__ void __launch_bounds__(256)
sgemm_small_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                   const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K,
                   long sam, long sak, long sbk, long sbn, int act, int avec, int bvec) {
    __shared__ float As[SG_K][SG_T + 1];
    __shared__ float Bs[SG_K][SG_T + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
    const int akunit = sak == 1, bkunit = sbk == 1;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    SgTile ta = sg_fetch(A, m0, 0, M, K, sam, sak, akunit, avec, tid);
    SgTile tb = sg_fetch(Bm, n0, 0, N, K, sbn, sbk, bkunit, bvec, tid);
    for (int k0 = 0; k0 < K; k0 += SG_K) {
        sg_store(As, ta, akunit, tid);
        sg_store(Bs, tb, bkunit, tid);
        __syncthreads();
        if (k0 + SG_K < K) {                       // prefetch the next K-step under the math
            ta = sg_fetch(A, m0, k0 + SG_K, M, K, sam, sak, akunit, avec, tid);
            tb = sg_fetch(Bm, n0, k0 + SG_K, N, K, sbn, sbk, bkunit, bvec, tid);
        }
#pragma unroll
        for (int kk = 0; kk < SG_K; ++kk) {
            const float a0 = As[kk][ty], a1 = As[kk][ty + 16];
            const float b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
            if (m < M && n < N) {
                float v = acc[i][j] + (bias != nullptr ? __ldg(bias + n) : 0.f);
                if (act == 1) v = fmaxf(v, 0.f);
                C[(size_t)m * N + n] = v;
            }
        }
}

static int sgemm_small(const float* A, const float* Bm, const float* bias, const float* mask_src,
                       float* C, int M, int N, int K, long sam, long sak, long sbk, long sbn,
                       int act, cudaStream_t st) {
    (void)mask_src;
    // 16-byte vector path: the non-unit stride must keep rows 16 B aligned
    const int avec = (((uintptr_t)A & 15) == 0) && ((sak == 1 ? sam : sak) % 4 == 0);
    const int bvec = (((uintptr_t)Bm & 15) == 0) && ((sbk == 1 ? sbn : sbk) % 4 == 0);
    dim3 grid(ecg_cdiv(N, SG_T), ecg_cdiv(M, SG_T));
    sgemm_small_kernel<<<grid, 256, 0, st>>>(A, Bm, bias, C, M, N, K, sam, sak, sbk, sbn, act, avec, bvec);
    return ecg_launch_status();
}

// in-place ReLU mask of dy, and column sums db[n] = sum_m dy[m][n] (one block per 32 columns)
__global__ void relu_mask_colsum_kernel(float* __restrict__ dy, const float* __restrict__ relu_out,
                                        float* __restrict__ db, int M, int N) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const int n = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (n < N) {
        for (int m = ty; m < M; m += 8) {
            float v = dy[(size_t)m * N + n];
            if (relu_out != nullptr) {
                if (!(relu_out[(size_t)m * N + n] > 0.f)) v = 0.f;
                dy[(size_t)m * N + n] = v;
            }
            s += v;
        }
    }
    sh[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N && db != nullptr) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][tx];
        db[n] = t;
    }
}

extern "C" int ecgb200_linear_fwd_f32(const float* x, const float* w, const float* b, float* y,
                                      int M, int K, int N, int act, void* stream) {
    if (!x || !w || !y || M <= 0 || K <= 0 || N <= 0) return ECGB200_EINVAL;
    // y[m][n] = sum_k x[m][k] * w[n][k]
    return sgemm_small(x, w, b, nullptr, y, M, N, K, K, 1, 1, K, act, (cudaStream_t)stream);
}

extern "C" int ecgb200_linear_bwd_f32(const float* x, const float* w, float* dy, const float* relu_out,
                                      float* dx, float* dw, float* db, int M, int K, int N, void* stream) {
    if (!x || !w || !dy || !dw || M <= 0 || K <= 0 || N <= 0) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (relu_out != nullptr || db != nullptr) {
        relu_mask_colsum_kernel<<<ecg_cdiv(N, 32), 256, 0, st>>>(dy, relu_out, db, M, N);
        int rc = ecg_launch_status();
        if (rc) return rc;
    }
    // dw[n][k] = sum_m dy[m][n] * x[m][k]      A(n, m) = dy[m*N + n], B(m, k) = x[m*K + k]
    int rc = sgemm_small(dy, x, nullptr, nullptr, dw, N, K, M, 1, N, K, 1, 0, st);
    if (rc) return rc;
    if (dx != nullptr) {
        // dx[m][k] = sum_n dy[m][n] * w[n][k]   A(m, n) = dy[m*N + n], B(n, k) = w[n*K + k]
        rc = sgemm_small(dy, w, nullptr, nullptr, dx, M, K, N, N, 1, K, 1, 0, st);
    }
    return rc;
}

// ------------------------------------------------------------------ FiLM
__global__ void film_fwd_kernel(const float* __restrict__ z, const float* __restrict__ film,
                                float* __restrict__ zc, int B, int F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * F) return;
    const int b = i / F, f = i - b * F;
    const float g = 1.0f + tanhf(film[(size_t)b * 2 * F + f]);
    zc[i] = fmaf(g, z[i], film[(size_t)b * 2 * F + F + f]);
}

__global__ void film_bwd_kernel(const float* __restrict__ z, const float* __restrict__ film,
                                const float* __restrict__ dzc, float* __restrict__ dz,
                                float* __restrict__ dfilm, int B, int F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * F) return;
    const int b = i / F, f = i - b * F;
    const float th = tanhf(film[(size_t)b * 2 * F + f]);
    const float d = dzc[i];
    if (dz != nullptr) dz[i] = d * (1.0f + th);
    dfilm[(size_t)b * 2 * F + f] = d * z[i] * (1.0f - th * th);
    dfilm[(size_t)b * 2 * F + F + f] = d;
}

extern "C" int ecgb200_film_fwd_f32(const float* z, const float* film, float* zc, int B, int F, void* stream) {
    if (!z || !film || !zc || B <= 0 || F <= 0) return ECGB200_EINVAL;
    film_fwd_kernel<<<ecg_cdiv(B * F, 256), 256, 0, (cudaStream_t)stream>>>(z, film, zc, B, F);
    return ecg_launch_status();
}

extern "C" int ecgb200_film_bwd_f32(const float* z, const float* film, const float* dzc, float* dz,
                                    float* dfilm, int B, int F, void* stream) {
    if (!z || !film || !dzc || !dfilm || B <= 0 || F <= 0) return ECGB200_EINVAL;
    film_bwd_kernel<<<ecg_cdiv(B * F, 256), 256, 0, (cudaStream_t)stream>>>(z, film, dzc, dz, dfilm, B, F);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ sigmoid + BCE
__global__ void __launch_bounds__(1024)
bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                  float* __restrict__ loss, float* __restrict__ dlogits, float* __restrict__ prob,
                  int n, float gscale) {
    __shared__ double sh[33];
    double s = 0.0;
    const float inv_n = 1.0f / (float)n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = logits[i];
        const float p = 1.0f / (1.0f + expf(-x));
        if (prob != nullptr) prob[i] = p;
        if (target != nullptr) {
            const float y = target[i];
            s += (double)(fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))));
            if (dlogits != nullptr) dlogits[i] = (p - y) * inv_n * gscale;
        }
    }
    if (loss != nullptr) {
        s = block_sum_d(s, sh);
        if (threadIdx.x == 0) *loss = (float)(s / (double)n);
    }
}

extern "C" int ecgb200_bce_logits_f32(const float* logits, const float* target, float* loss,
                                      float* dlogits, float* prob, int n, float gscale, void* stream) {
    if (!logits || n <= 0) return ECGB200_EINVAL;
    if ((loss || dlogits) && !target) return ECGB200_EINVAL;
    bce_logits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, target, loss, dlogits, prob, n, gscale);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ evaluation epilogue
// prob = sigmoid(logit) (loop.py:63), pred = prob >= threshold computed on the fp32 probability exactly as the
// reference does in numpy (scripts/06:127, metrics.py:37), and per-label confusion counts accumulated on the
// device: counts[c] = {tp, fp, fn, tn} (int32, integer atomics => deterministic), so that an evaluation epoch
// needs no per-batch host synchronisation for its thresholded metrics.
__global__ void __launch_bounds__(256)
eval_counts_kernel(const float* __restrict__ logits, const float* __restrict__ target, float* __restrict__ prob,
                   unsigned char* __restrict__ pred, int* __restrict__ counts, int rows, int C, float threshold) {
    __shared__ int sh[8][4];
    const int c = blockIdx.y;
    int tp = 0, fp = 0, fn = 0, tn = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        const float x = logits[(size_t)r * C + c];
        const float p = 1.0f / (1.0f + expf(-x));
        const bool yp = p >= threshold;
        if (prob != nullptr) prob[(size_t)r * C + c] = p;
        if (pred != nullptr) pred[(size_t)r * C + c] = yp ? 1 : 0;
        if (target != nullptr && counts != nullptr) {
            const bool yt = target[(size_t)r * C + c] > 0.5f;
            tp += (yp && yt); fp += (yp && !yt); fn += (!yp && yt); tn += (!yp && !yt);
        }
    }
    if (counts == nullptr || target == nullptr) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o);
        fn += __shfl_xor_sync(0xffffffffu, fn, o); tn += __shfl_xor_sync(0xffffffffu, tn, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[w][0] = tp; sh[w][1] = fp; sh[w][2] = fn; sh[w][3] = tn; }
    __syncthreads();
    if (threadIdx.x < 4) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        if (t) atomicAdd(counts + c * 4 + threadIdx.x, t);
    }
}

extern "C" int ecgb200_eval_counts_f32(const float* logits, const float* target, float* prob, unsigned char* pred,
                                       int* counts, int rows, int C, float threshold, void* stream) {
    if (!logits || rows <= 0 || C <= 0 || C > 65535) return ECGB200_EINVAL;
    int bx = ecg_cdiv(rows, 256);
    if (bx > 64) bx = 64;
    eval_counts_kernel<<<dim3(bx, C), 256, 0, (cudaStream_t)stream>>>(logits, target, prob, pred, counts, rows, C, threshold);
    return ecg_launch_status();
}

// ------------------------------------------------------------------ AdamW (multi-tensor, one launch)
constexpr int ADAM_MAXT = 64;
struct AdamTensors {
    float* p[ADAM_MAXT];
    const float* g[ADAM_MAXT];
    float* m[ADAM_MAXT];
    float* v[ADAM_MAXT];
    long long n[ADAM_MAXT];
};
struct AdamScalars { float decay, one_m_b1, b2, one_m_b2, bc2_sqrt, eps, step_size, gscale; };

// hyper (device): {lr, beta1, beta2, eps, weight_decay, gscale}; step_ctr (device): number of
// steps already taken.  Keeping both on the device makes the launch CUDA-graph replayable:
// the bias corrections are recomputed from the counter inside the kernel.
__global__ void __launch_bounds__(256)
adamw_kernel(const __grid_constant__ AdamTensors T, const float* __restrict__ hyper,
             const int* __restrict__ step_ctr) {
    __shared__ AdamScalars S;
    if (threadIdx.x == 0) {
        const double lr = hyper[0], b1 = hyper[1], b2 = hyper[2], wd = hyper[4];
        const double step = (double)(step_ctr[0] + 1);
        S.decay = (float)(1.0 - lr * wd);
        S.one_m_b1 = (float)(1.0 - b1);
        S.b2 = hyper[2];
        S.one_m_b2 = (float)(1.0 - b2);
        S.bc2_sqrt = (float)sqrt(1.0 - pow(b2, step));
        S.eps = hyper[3];
        S.step_size = (float)(lr / (1.0 - pow(b1, step)));
        S.gscale = hyper[5];
    }
    __syncthreads();
    const int t = blockIdx.y;
    const long long n = T.n[t];
    float* __restrict__ p = T.p[t];
    const float* __restrict__ g = T.g[t];
    float* __restrict__ m = T.m[t];
    float* __restrict__ v = T.v[t];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const AdamK K = {S.decay, S.one_m_b1, S.b2, S.one_m_b2, S.bc2_sqrt, S.eps, S.step_size};
        float pi = p[i], mi = m[i], vi = v[i];
        adamw_update(pi, mi, vi, __fmul_rn(g[i], S.gscale), K);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

__global__ void counter_inc_kernel(int* ctr) { ctr[0] += 1; }

extern "C" int ecgb200_adamw_f32(int nt, float* const* p, const float* const* g, float* const* m,
                                 float* const* v, const int64_t* numel, const float* hyper,
                                 int* step_ctr, void* stream) {
    if (nt <= 0 || !p || !g || !m || !v || !numel || !hyper || !step_ctr) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    for (int t0 = 0; t0 < nt; t0 += ADAM_MAXT) {
        AdamTensors T;
        const int cnt = nt - t0 < ADAM_MAXT ? nt - t0 : ADAM_MAXT;
        long long maxn = 0;
        for (int i = 0; i < cnt; ++i) {
            T.p[i] = p[t0 + i]; T.g[i] = g[t0 + i]; T.m[i] = m[t0 + i]; T.v[i] = v[t0 + i];
            T.n[i] = numel[t0 + i];
            if (T.n[i] > maxn) maxn = T.n[i];
        }
        long long bx = (maxn + 256 * 4 - 1) / (256 * 4);
        if (bx < 1) bx = 1;
        if (bx > 2048) bx = 2048;
        dim3 grid((unsigned)bx, cnt);
        adamw_kernel<<<grid, 256, 0, st>>>(T, hyper, step_ctr);
        int rc = ecg_launch_status();
        if (rc) return rc;
    }
    counter_inc_kernel<<<1, 1, 0, st>>>(step_ctr);
    return ecg_launch_status();
}
