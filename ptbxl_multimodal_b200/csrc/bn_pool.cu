// BatchNorm1d (train / eval) + ReLU + MaxPool1d(2) [+ global average pool], forward and
// backward, fp32.  Memory-bound: every kernel streams y once with coalesced accesses.
//
// Replaces aten::native_batch_norm(+_backward), relu_/threshold_backward,
// max_pool2d_with_indices(+_backward), adaptive_avg_pool1d reached from
// /root/reference/src/models/ecg_cnn.py:14-16,46,62.
#include "common.cuh"
#include <cstdlib>

// bn_state layout: [0:C) mean, [C:2C) rstd, [2C:3C) scale = gamma*rstd, [3C:4C) shift = beta - mean*scale

// ---------------------------------------------------------------- statistics
// One warp per (b, c) row: {sum, centred M2}.  Output layout [2][C][B] (tile == sample).
__global__ void bn_row_stats_kernel(const float* __restrict__ y, float* __restrict__ part,
                                    int B, int C, int L) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B * C) return;
    const int lane = threadIdx.x & 31;
    const float* yr = y + (size_t)row * L;
    float s = 0.f;
    for (int t = lane; t < L; t += 32) s += __ldg(yr + t);
    s = warp_sum(s);
    const float mean = s / (float)L;
    float m2 = 0.f;
    for (int t = lane; t < L; t += 32) { const float d = __ldg(yr + t) - mean; m2 = fmaf(d, d, m2); }
    m2 = warp_sum(m2);
    if (lane == 0) {
        const int b = row / C, c = row - b * C;
        part[(size_t)c * B + b] = s;
        part[((size_t)C + c) * B + b] = m2;
    }
}

// One block per channel: merge per-tile {sum, M2} (Chan et al.) in double, in a fixed order.
// cnt(tile) = min(tile_len, L - (tile % tiles_per_row) * tile_len).
__global__ void bn_finalize_kernel(const float* __restrict__ part, int ntiles, int tiles_per_row,
                                   int tile_len, int L, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                   float* __restrict__ bn_state, int C, float momentum, float eps) {
    __shared__ double sh[33];
    const int c = blockIdx.x;
    const float* ps = part + (size_t)c * ntiles;
    const float* pm = part + ((size_t)C + c) * ntiles;
    double s = 0.0;
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) s += (double)ps[i];
    s = block_sum_d(s, sh);
    const double n = (double)(ntiles / tiles_per_row) * (double)L;
    const double mean = s / n;
    double m2 = 0.0;
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) {
        const int tt = i % tiles_per_row;
        const double cnt = (double)min(tile_len, L - tt * tile_len);
        const double d = (double)ps[i] / cnt - mean;
        m2 += (double)pm[i] + cnt * d * d;
    }
    m2 = block_sum_d(m2, sh);
    if (threadIdx.x == 0) {
        const double var = m2 / n;                                   // biased
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float meanf = (float)mean;
        const float scale = gamma[c] * rstd;
        bn_state[c] = meanf;
        bn_state[C + c] = rstd;
        bn_state[2 * C + c] = scale;
        bn_state[3 * C + c] = beta[c] - meanf * scale;
        if (running_mean != nullptr) {
            const double unbiased = n > 1.0 ? m2 / (n - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * meanf;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
        if (c == 0 && nbt != nullptr) *nbt += 1;
    }
}

__global__ void bn_eval_state_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ rm, const float* __restrict__ rv,
                                     float* __restrict__ bn_state, int C, float eps) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float rstd = 1.0f / sqrtf(rv[c] + eps);
    const float scale = gamma[c] * rstd;
    bn_state[c] = rm[c];
    bn_state[C + c] = rstd;
    bn_state[2 * C + c] = scale;
    bn_state[3 * C + c] = beta[c] - rm[c] * scale;
}

extern "C" size_t ecgb200_bn_stats_ws_bytes(int B, int Co, int L) {
    (void)L;
    return (size_t)2 * Co * B * sizeof(float);
}

extern "C" int ecgb200_bn_train_stats_f32(const float* y, const float* stat_part, const float* gamma,
                                          const float* beta, float* running_mean, float* running_var,
                                          int64_t* nbt, float* bn_state, void* ws, int B, int Co,
                                          int L, float momentum, float eps, void* stream) {
    if (!gamma || !beta || !bn_state || B <= 0 || Co <= 0 || L <= 0) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (stat_part != nullptr) {
        const int tpr = ecg_cdiv(L, 128);
        bn_finalize_kernel<<<Co, 256, 0, st>>>(stat_part, B * tpr, tpr, 128, L, gamma, beta,
                                               running_mean, running_var, nbt, bn_state, Co, momentum, eps);
    } else {
        if (!y || !ws) return ECGB200_EINVAL;
        const int rows = B * Co;
        bn_row_stats_kernel<<<ecg_cdiv(rows, 8), 256, 0, st>>>(y, (float*)ws, B, Co, L);
        int rc = ecg_launch_status();
        if (rc) return rc;
        bn_finalize_kernel<<<Co, 256, 0, st>>>((const float*)ws, B, 1, L, L, gamma, beta, running_mean,
                                               running_var, nbt, bn_state, Co, momentum, eps);
    }
    return ecg_launch_status();
}

extern "C" int ecgb200_bn_eval_state_f32(const float* gamma, const float* beta, const float* running_mean,
                                         const float* running_var, float* bn_state, int Co, float eps,
                                         void* stream) {
    if (!gamma || !beta || !running_mean || !running_var || !bn_state || Co <= 0) return ECGB200_EINVAL;
    bn_eval_state_kernel<<<ecg_cdiv(Co, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean,
                                                                              running_var, bn_state, Co, eps);
    return ecg_launch_status();
}

// ---------------------------------------------------------------- forward apply
// grid (ceil(Lp/256), B*C): thread -> one pooled output (reads 2 inputs).
__global__ void bn_relu_pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ bn_state,
                                        float* __restrict__ p, int C, int L, int Lp, int vec) {
    const int row = blockIdx.y;
    const int c = row % C;
    const float sc = __ldg(bn_state + 2 * C + c), sh = __ldg(bn_state + 3 * C + c);
    const float* yr = y + (size_t)row * L;
    float* pr = p + (size_t)row * Lp;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Lp; j += gridDim.x * blockDim.x) {
        float a0, a1;
        if (vec) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(yr) + j);
            a0 = v.x; a1 = v.y;
        } else {
            a0 = __ldg(yr + 2 * j); a1 = __ldg(yr + 2 * j + 1);
        }
        const float r0 = fmaxf(fmaf(a0, sc, sh), 0.f), r1 = fmaxf(fmaf(a1, sc, sh), 0.f);
        pr[j] = fmaxf(r0, r1);
    }
}

// one warp per (b, c) row: gap[b,c] = mean_j pooled
__global__ void bn_relu_pool_gap_kernel(const float* __restrict__ y, const float* __restrict__ bn_state,
                                        float* __restrict__ p, float* __restrict__ gap,
                                        int rows, int C, int L, int Lp) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int c = row % C;
    const float sc = __ldg(bn_state + 2 * C + c), sh = __ldg(bn_state + 3 * C + c);
    const float* yr = y + (size_t)row * L;
    float s = 0.f;
    for (int j = lane; j < Lp; j += 32) {
        const float r0 = fmaxf(fmaf(__ldg(yr + 2 * j), sc, sh), 0.f);
        const float r1 = fmaxf(fmaf(__ldg(yr + 2 * j + 1), sc, sh), 0.f);
        const float m = fmaxf(r0, r1);
        if (p != nullptr) p[(size_t)row * Lp + j] = m;
        s += m;
    }
    s = warp_sum(s);
    if (lane == 0) gap[row] = s / (float)Lp;
}

extern "C" int ecgb200_bn_relu_pool_fwd_f32(const float* y, const float* bn_state, float* p, float* gap,
                                            int B, int Co, int L, void* stream) {
    if (!y || !bn_state || (!p && !gap) || B <= 0 || Co <= 0 || L < 2) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = L / 2, rows = B * Co;
    if (gap != nullptr) {
        bn_relu_pool_gap_kernel<<<ecg_cdiv(rows, 8), 256, 0, st>>>(y, bn_state, p, gap, rows, Co, L, Lp);
    } else {
        if (rows > 65535 * 32) return ECGB200_EUNSUPPORTED;
        // grid.y limit is 65535: launch in slabs of whole samples so the channel phase is kept
        const int threads = Lp >= 256 ? 256 : (Lp >= 128 ? 128 : 64);
        const int vec = ((L & 1) == 0) && (((uintptr_t)y & 7) == 0);
        // slabs aligned to C rows
        const int slab = (65535 / Co) * Co;
        for (int r0 = 0; r0 < rows; r0 += slab) {
            const int nr = rows - r0 < slab ? rows - r0 : slab;
            dim3 grid(ecg_cdiv(Lp, threads), nr);
            bn_relu_pool_fwd_kernel<<<grid, threads, 0, st>>>(y + (size_t)r0 * L, bn_state,
                                                              p + (size_t)r0 * Lp, Co, L, Lp, vec);
        }
    }
    return ecg_launch_status();
}

// ---------------------------------------------------------------- backward
// Routed gradient g at position 2j / 2j+1 of a pool pair (first index wins ties; ReLU mask r>0):
//   g0 = d if (r0 >= r1 && r0 > 0);  g1 = d if (r1 > r0)
// where d = dp[b,c,j]  (or dgap[b,c]/Lp).
struct PoolGrad { float g0, g1, xh0, xh1; };

__device__ __forceinline__ PoolGrad pool_grad(float a0, float a1, float d, float mean, float rstd,
                                              float sc, float sh) {
    PoolGrad r;
    const float r0 = fmaxf(fmaf(a0, sc, sh), 0.f), r1 = fmaxf(fmaf(a1, sc, sh), 0.f);
    r.g0 = (r0 >= r1 && r0 > 0.f) ? d : 0.f;
    r.g1 = (r1 > r0) ? d : 0.f;
    r.xh0 = (a0 - mean) * rstd;
    r.xh1 = (a1 - mean) * rstd;
    return r;
}

// one warp per row -> partial {sum g, sum g*xhat}; layout [2][C][B]
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ y, const float* __restrict__ bn_state,
                                     const float* __restrict__ dp, const float* __restrict__ dgap,
                                     float* __restrict__ part, int B, int C, int L, int Lp) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B * C) return;
    const int lane = threadIdx.x & 31;
    const int b = row / C, c = row - b * C;
    const float mean = __ldg(bn_state + c), rstd = __ldg(bn_state + C + c);
    const float sc = __ldg(bn_state + 2 * C + c), sh = __ldg(bn_state + 3 * C + c);
    const float* yr = y + (size_t)row * L;
    const float dconst = dgap != nullptr ? __ldg(dgap + row) / (float)Lp : 0.f;
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < Lp; j += 32) {
        const float d = dp != nullptr ? __ldg(dp + (size_t)row * Lp + j) : dconst;
        const PoolGrad r = pool_grad(__ldg(yr + 2 * j), __ldg(yr + 2 * j + 1), d, mean, rstd, sc, sh);
        s1 += r.g0 + r.g1;
        s2 = fmaf(r.g0, r.xh0, fmaf(r.g1, r.xh1, s2));
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        part[(size_t)c * B + b] = s1;
        part[((size_t)C + c) * B + b] = s2;
    }
}

// one block per channel: sums over the batch in double, fixed order
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       int B, int C) {
    __shared__ double sh[33];
    const int c = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        s1 += (double)part[(size_t)c * B + i];
        s2 += (double)part[((size_t)C + c) * B + i];
    }
    s1 = block_sum_d(s1, sh);
    s2 = block_sum_d(s2, sh);
    if (threadIdx.x == 0) {
        sums[c] = (float)s1;
        sums[C + c] = (float)s2;
        if (dbeta) dbeta[c] = (float)s1;
        if (dgamma) dgamma[c] = (float)s2;
    }
}

// grid (ceil(Lp/threads), rows): thread -> one pool pair (+ the odd tail element)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ bn_state,
                                    const float* __restrict__ dp, const float* __restrict__ dgap,
                                    const float* __restrict__ sums, float* __restrict__ dy,
                                    int C, int L, int Lp, float inv_n, int train, int vec) {
    const int row = blockIdx.y;
    const int c = row % C;
    const float mean = __ldg(bn_state + c), rstd = __ldg(bn_state + C + c);
    const float sc = __ldg(bn_state + 2 * C + c), sh = __ldg(bn_state + 3 * C + c);
    const float m1 = train ? __ldg(sums + c) * inv_n : 0.f;
    const float m2 = train ? __ldg(sums + C + c) * inv_n : 0.f;
    const float* yr = y + (size_t)row * L;
    float* dr = dy + (size_t)row * L;
    const float dconst = dgap != nullptr ? __ldg(dgap + row) / (float)Lp : 0.f;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Lp; j += gridDim.x * blockDim.x) {
        const float d = dp != nullptr ? __ldg(dp + (size_t)row * Lp + j) : dconst;
        const PoolGrad r = pool_grad(__ldg(yr + 2 * j), __ldg(yr + 2 * j + 1), d, mean, rstd, sc, sh);
        const float o0 = sc * (r.g0 - m1 - r.xh0 * m2);
        const float o1 = sc * (r.g1 - m1 - r.xh1 * m2);
        if (vec) {
            reinterpret_cast<float2*>(dr)[j] = make_float2(o0, o1);
        } else {
            dr[2 * j] = o0; dr[2 * j + 1] = o1;
        }
        if (j == Lp - 1 && (L & 1)) {
            const float xh = (__ldg(yr + L - 1) - mean) * rstd;
            dr[L - 1] = sc * (0.f - m1 - xh * m2);
        }
    }
}

extern "C" size_t ecgb200_bn_bwd_ws_bytes(int B, int Co) {
    return ((size_t)2 * Co * B + 2 * Co) * sizeof(float);
}

extern "C" int ecgb200_bn_relu_pool_bwd_f32(const float* y, const float* bn_state, const float* gamma,
                                            const float* dp, const float* dgap, float* dy, float* dgamma,
                                            float* dbeta, void* ws, int B, int Co, int L, int train,
                                            void* stream) {
    (void)gamma;
    if (!y || !bn_state || (!dp && !dgap) || !dy || !ws || B <= 0 || Co <= 0 || L < 2) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = L / 2, rows = B * Co;
    float* part = (float*)ws;
    float* sums = part + (size_t)2 * Co * B;
    bn_bwd_reduce_kernel<<<ecg_cdiv(rows, 8), 256, 0, st>>>(y, bn_state, dp, dgap, part, B, Co, L, Lp);
    int rc = ecg_launch_status();
    if (rc) return rc;
    bn_bwd_finalize_kernel<<<Co, 128, 0, st>>>(part, sums, dgamma, dbeta, B, Co);
    rc = ecg_launch_status();
    if (rc) return rc;
    const int threads = Lp >= 256 ? 256 : (Lp >= 128 ? 128 : 64);
    const float inv_n = 1.0f / ((float)B * (float)L);
    const int vec = ((L & 1) == 0) && (((uintptr_t)dy & 7) == 0);
    const int slab = (65535 / Co) * Co;
    for (int r0 = 0; r0 < rows; r0 += slab) {
        const int nr = rows - r0 < slab ? rows - r0 : slab;
        dim3 grid(ecg_cdiv(Lp, threads), nr);
        bn_bwd_apply_kernel<<<grid, threads, 0, st>>>(
            y + (size_t)r0 * L, bn_state, dp ? dp + (size_t)r0 * Lp : nullptr,
            dgap ? dgap + r0 : nullptr, sums, dy + (size_t)r0 * L, Co, L, Lp, inv_n, train, vec);
    }
    return ecg_launch_status();
}

// =====================================================================================
// bf16 "blocked channels-last" variants  A[b][c/8][t][c%8]  (16 B = 8 channels of one step).
// One block per (sample, channel chunk): every access is a coalesced 16-byte vector; the 8
// per-channel BN constants sit in registers.  Statistics / reductions stay fp32 (+double merge).
// =====================================================================================
#include <cuda_bf16.h>

__device__ __forceinline__ void bf8_unpack(const uint4 u, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 bf8_pack(const float* v) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return u;
}

// block-wide sum of NV floats per thread; result valid in thread 0 (and all of warp 0)
template <int NV>
__device__ __forceinline__ void block_sum_vec(float* v, float* sh /* [8][NV] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[w * NV + i] = v[i];
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = 0.f;
            for (int j = 0; j < nw; ++j) t += sh[j * NV + i];
            v[i] = t;
        }
    }
    __syncthreads();
}

// Work split: grid (C/8, NS); block (cc, s) owns samples [s*tile_b, min(B, (s+1)*tile_b)) of channel
// chunk cc and walks the flattened (sample, time) index space with 256 threads, so that every
// block has a few thousand 16-byte vectors to stream and exactly ONE block reduction at the end.
__host__ __device__ inline int bnb_tile_b(int B, int C, int blocks_per_sm = 4) {
    int ns = (148 * blocks_per_sm) / (C / 8);      // one full wave of resident blocks, never more
    if (ns > B) ns = B;
    if (ns < 1) ns = 1;
    return (B + ns - 1) / ns;
}
// backward kernels hold ~80 registers: 3 blocks of 256 threads per SM (held to 64 registers / 4 blocks they measured
// SLOWER, alone 18.5 -> 21.3 us per block and in the step 0.422 -> 0.441 ms: spills in the reduce pass, more partials)
extern "C" int ecgb200_bn_nsplit(int B, int C) { const int tb = bnb_tile_b(B, C, 3); return (B + tb - 1) / tb; }

// Merge `nparts` partial pairs laid out part[i][2][C] for the 8 channels of chunk cc with all 256
// threads (double, fixed order => deterministic): out[0..7] = first moments, out[8..15] = second.
__device__ __forceinline__ void merge_parts8(const float* __restrict__ part, int nparts, int C, int cc,
                                             double* shd /* [32*16] */, double* out /* [16] */) {
    const int ch = threadIdx.x & 7, grp = threadIdx.x >> 3;
    double a = 0.0, b = 0.0;
    for (int i = grp; i < nparts; i += 32) {
        a += (double)__ldg(part + ((size_t)i * 2 + 0) * C + cc * 8 + ch);
        b += (double)__ldg(part + ((size_t)i * 2 + 1) * C + cc * 8 + ch);
    }
    shd[grp * 16 + ch] = a;
    shd[grp * 16 + 8 + ch] = b;
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll 8
        for (int g = 0; g < 32; ++g) t += shd[g * 16 + threadIdx.x];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}

// Sum of v[0..7] over the 32 lanes with a transposing butterfly: 4 + 2 + 1 + 1 + 1 = 9 shuffles instead of 8 x 5.  Afterwards
// v[0] of lane l is the total of element 4 * bit4(l) + 2 * bit3(l) + bit2(l) (the same value in the four lanes l, l^1, l^2, l^3).
__device__ __forceinline__ float warp_sum8_transposed(float* v, int lane) {
#pragma unroll
    for (int o = 16, half = 4; o >= 4; o >>= 1, half >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half], keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// Block 4 (GAP gradient): the same two batch sums {sum g, sum g*a} from the forward pass's routing summary
// route[0][b][c] = routed pool pairs, route[1][b][c] = sum of the raw conv outputs at the routed positions, with g = dgap[b][c] / Lp
// for every routed position of (b, c).  Fixed order, double accumulation, every block for its own 8 channels.
__device__ __forceinline__ void merge_route8(const float* __restrict__ route, const float* __restrict__ dgap, int B, int C,
                                             int cc, float inv_lp, double* shd /* [32*16] */, double* out /* [16] */) {
    const int ch = threadIdx.x & 7, grp = threadIdx.x >> 3;
    double a = 0.0, b = 0.0;
    for (int i = grp; i < B; i += 32) {
        const size_t o = (size_t)i * C + cc * 8 + ch;
        const double g = (double)(__ldg(dgap + o) * inv_lp);
        a += g * (double)__ldg(route + o);
        b += g * (double)__ldg(route + (size_t)B * C + o);
    }
    shd[grp * 16 + ch] = a;
    shd[grp * 16 + 8 + ch] = b;
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll 8
        for (int g = 0; g < 32; ++g) t += shd[g * 16 + threadIdx.x];
        out[threadIdx.x] = t;
    }
    __syncthreads();
}

// {sum, centred M2} per (channel, split): part[c][s], part[C + c][s]
__global__ void __launch_bounds__(256)
bn_stats_bf16_kernel(const uint4* __restrict__ y, float* __restrict__ part, int B, int C, int L, int tile_b) {
    __shared__ float sh[8 * 8];
    __shared__ float mean_s[8];
    const int cc = blockIdx.x, NS = gridDim.y;
    const int b0 = blockIdx.y * tile_b, nb = min(tile_b, B - b0);
    const int n = nb * L;
    const size_t bstride = (size_t)(C / 8) * L;
    const uint4* y0 = y + ((size_t)b0 * (C / 8) + cc) * L;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int bl = idx / L, t = idx - bl * L;
        float v[8];
        bf8_unpack(__ldg(y0 + (size_t)bl * bstride + t), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += v[i];
    }
    block_sum_vec<8>(s, sh);
    if (threadIdx.x < 8) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (i == threadIdx.x) t = s[i];
        mean_s[threadIdx.x] = t / (float)n;
        part[(size_t)(cc * 8 + threadIdx.x) * NS + blockIdx.y] = t;
    }
    __syncthreads();
    float m[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { m[i] = mean_s[i]; q[i] = 0.f; }
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int bl = idx / L, t = idx - bl * L;
        float v[8];
        bf8_unpack(__ldg(y0 + (size_t)bl * bstride + t), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - m[i]; q[i] = fmaf(d, d, q[i]); }
    }
    block_sum_vec<8>(q, sh);
    if (threadIdx.x < 8) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (i == threadIdx.x) t = q[i];
        part[((size_t)C + cc * 8 + threadIdx.x) * NS + blockIdx.y] = t;
    }
}

extern "C" int ecgb200_bn_train_stats_bf16(const void* yb, const float* gamma, const float* beta,
                                           float* running_mean, float* running_var, int64_t* nbt,
                                           float* bn_state, void* ws, int B, int C, int L, float momentum,
                                           float eps, void* stream) {
    if (!yb || !gamma || !beta || !bn_state || !ws || B <= 0 || C <= 0 || (C & 7) || L <= 0) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int tile_b = bnb_tile_b(B, C), NS = (B + tile_b - 1) / tile_b;
    bn_stats_bf16_kernel<<<dim3(C / 8, NS), 256, 0, st>>>((const uint4*)yb, (float*)ws, B, C, L, tile_b);
    int rc = ecg_launch_status();
    if (rc) return rc;
    // tiles = sample ranges: "row" = all B*L samples of a channel, tile_len = tile_b*L
    bn_finalize_kernel<<<C, 128, 0, st>>>((const float*)ws, NS, NS, tile_b * L, B * L, gamma, beta, running_mean,
                                          running_var, nbt, bn_state, C, momentum, eps);
    return ecg_launch_status();
}

// p = maxpool2(relu(bn(y))) in the blocked layout (flattened (sample, pair) walk)
__global__ void __launch_bounds__(256)
bn_relu_pool_fwd_bf16_kernel(const uint4* __restrict__ y, const float* __restrict__ bn_state,
                             uint4* __restrict__ p, int B, int C, int L, int Lp, int tile_b) {
    const int cc = blockIdx.x;
    const int b0 = blockIdx.y * tile_b, nb = min(tile_b, B - b0);
    const int n = nb * Lp;
    float sc[8], sf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = __ldg(bn_state + 2 * C + cc * 8 + i);
        sf[i] = __ldg(bn_state + 3 * C + cc * 8 + i);
    }
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int bl = idx / Lp, j = idx - bl * Lp;
        const size_t row = (size_t)(b0 + bl) * (C / 8) + cc;
        float a0[8], a1[8], m[8];
        bf8_unpack(__ldg(y + row * L + 2 * j), a0);
        bf8_unpack(__ldg(y + row * L + 2 * j + 1), a1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float r0 = fmaxf(fmaf(a0[i], sc[i], sf[i]), 0.f), r1 = fmaxf(fmaf(a1[i], sc[i], sf[i]), 0.f);
            m[i] = fmaxf(r0, r1);
        }
        p[row * Lp + j] = bf8_pack(m);
    }
}

// variant with the global average pool: one warp per (sample, chunk) row at a time
__global__ void __launch_bounds__(256)
bn_relu_pool_gap_bf16_kernel(const uint4* __restrict__ y, const float* __restrict__ bn_state,
                             uint4* __restrict__ p, float* __restrict__ gap, int B, int C, int L, int Lp,
                             int tile_b) {
    const int cc = blockIdx.x;
    const int b0 = blockIdx.y * tile_b, nb = min(tile_b, B - b0);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float sc[8], sf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = __ldg(bn_state + 2 * C + cc * 8 + i);
        sf[i] = __ldg(bn_state + 3 * C + cc * 8 + i);
    }
    for (int bl = w; bl < nb; bl += nw) {
        const size_t row = (size_t)(b0 + bl) * (C / 8) + cc;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int j = lane; j < Lp; j += 32) {
            float a0[8], a1[8], m[8];
            bf8_unpack(__ldg(y + row * L + 2 * j), a0);
            bf8_unpack(__ldg(y + row * L + 2 * j + 1), a1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r0 = fmaxf(fmaf(a0[i], sc[i], sf[i]), 0.f), r1 = fmaxf(fmaf(a1[i], sc[i], sf[i]), 0.f);
                m[i] = fmaxf(r0, r1);
                acc[i] += m[i];
            }
            if (p != nullptr) p[row * Lp + j] = bf8_pack(m);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
        if (lane < 8) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) if (i == lane) t = acc[i];
            gap[(size_t)(b0 + bl) * C + cc * 8 + lane] = t / (float)Lp;
        }
    }
}

extern "C" int ecgb200_bn_relu_pool_fwd_bf16(const void* yb, const float* bn_state, void* pb, float* gap,
                                             int B, int C, int L, void* stream) {
    if (!yb || !bn_state || (!pb && !gap) || B <= 0 || C <= 0 || (C & 7) || L < 2) return ECGB200_EINVAL;
    const int Lp = L / 2;
    const int tile_b = bnb_tile_b(B, C), NS = (B + tile_b - 1) / tile_b;
    cudaStream_t st = (cudaStream_t)stream;
    if (gap != nullptr)
        bn_relu_pool_gap_bf16_kernel<<<dim3(C / 8, NS), 256, 0, st>>>((const uint4*)yb, bn_state, (uint4*)pb, gap,
                                                                     B, C, L, Lp, tile_b);
    else
        bn_relu_pool_fwd_bf16_kernel<<<dim3(C / 8, NS), 256, 0, st>>>((const uint4*)yb, bn_state, (uint4*)pb,
                                                                     B, C, L, Lp, tile_b);
    return ecg_launch_status();
}

// Train-mode forward in ONE pass over y: every block first turns the conv epilogue's per-CTA
// {sum, sum of squares} partials into mean / rstd / scale / shift for its 8 channels (identical in
// every block; block row 0 publishes bn_state and updates the running statistics), then applies
// BN + ReLU + MaxPool1d(2) (GAP variant: also the mean over time of the UNROUNDED pooled values).
// ROUTE (block 4, whose pooled output only feeds the time average): also leaves, per (window, channel), the number of pool
// pairs that route a gradient (ReLU-positive maximum) and the sum of the raw conv outputs at the routed positions --
// route[0][b][c], route[1][b][c].  The gradient of every pooled position of (b, c) is the same dgap[b][c] / Lp, so the two batch
// reductions of the BatchNorm backward become sums over B x C of dgap * route instead of a pass over y (ecgb200_bn_relu_pool_bwd_route_bf16).
template <bool GAP, bool ROUTE>
__global__ void __launch_bounds__(256, ROUTE ? 3 : 4)
bn_fwd_train_bf16_kernel(const uint4* __restrict__ y, const float* __restrict__ part, int nparts,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ running_mean, float* __restrict__ running_var,
                         int64_t* __restrict__ nbt, float* __restrict__ bn_state, uint4* __restrict__ p,
                         float* __restrict__ gap, int B, int C, int L, int Lp, int tile_b, float momentum,
                         float eps, int nrep, float* __restrict__ route) {
    __shared__ double shd[32 * 16], mom[16];
    __shared__ float scs[8], sfs[8];
    const int cc = blockIdx.x;
    const int b0 = blockIdx.y * tile_b, nb = min(tile_b, B - b0);
    ecg_pdl_launch_dependents();
    ecg_pdl_wait();
    merge_parts8(part, nparts, C, cc, shd, mom);
    if (threadIdx.x < 8) {
        const int c = cc * 8 + threadIdx.x;
        const double n = (double)B * (double)L * (double)nrep;   // nrep > 1: the partials cover nrep replicas (SyncBN)
        const double mean = mom[threadIdx.x] / n;
        double var = mom[8 + threadIdx.x] / n - mean * mean;          // biased
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float meanf = (float)mean;
        const float scale = __ldg(gamma + c) * rstd;
        const float shift = __ldg(beta + c) - meanf * scale;
        scs[threadIdx.x] = scale;
        sfs[threadIdx.x] = shift;
        if (blockIdx.y == 0) {
            bn_state[c] = meanf;
            bn_state[C + c] = rstd;
            bn_state[2 * C + c] = scale;
            bn_state[3 * C + c] = shift;
            if (running_mean != nullptr) {
                const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * meanf;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
            if (c == 0 && nbt != nullptr) *nbt += 1;
        }
    }
    __syncthreads();
    float sc[8], sf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = scs[i]; sf[i] = sfs[i]; }
    if (!GAP) {
        const int n = nb * Lp;
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int bl = idx / Lp, j = idx - bl * Lp;
            const size_t row = (size_t)(b0 + bl) * (C / 8) + cc;
            float a0[8], a1[8], m[8];
            bf8_unpack(__ldg(y + row * L + 2 * j), a0);
            bf8_unpack(__ldg(y + row * L + 2 * j + 1), a1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float r0 = fmaxf(fmaf(a0[i], sc[i], sf[i]), 0.f), r1 = fmaxf(fmaf(a1[i], sc[i], sf[i]), 0.f);
                m[i] = fmaxf(r0, r1);
            }
            p[row * Lp + j] = bf8_pack(m);
        }
    } else {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int bl = w; bl < nb; bl += nw) {
            const size_t row = (size_t)(b0 + bl) * (C / 8) + cc;
            float acc[8], cnt[ROUTE ? 8 : 1], sa[ROUTE ? 8 : 1];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            if constexpr (ROUTE) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { cnt[i] = 0.f; sa[i] = 0.f; }
            }
            for (int j0 = lane; j0 < Lp; j0 += 64) {
                uint4 u0[2], u1[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = min(j0 + 32 * e, Lp - 1);
                    u0[e] = __ldg(y + row * L + 2 * j);
                    u1[e] = __ldg(y + row * L + 2 * j + 1);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = j0 + 32 * e;
                    if (j >= Lp) break;
                    float a0[8], a1[8], m[8];
                    bf8_unpack(u0[e], a0);
                    bf8_unpack(u1[e], a1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float r0 = fmaxf(fmaf(a0[i], sc[i], sf[i]), 0.f), r1 = fmaxf(fmaf(a1[i], sc[i], sf[i]), 0.f);
                        m[i] = fmaxf(r0, r1);
                        acc[i] += m[i];
                        if constexpr (ROUTE) {                       // pool_sel: first index wins ties, ReLU mask
                            const bool s0 = (r0 >= r1) && (r0 > 0.f), s1 = r1 > r0;
                            cnt[i] += (s0 || s1) ? 1.f : 0.f;
                            sa[i] += s0 ? a0[i] : (s1 ? a1[i] : 0.f);
                        }
                    }
                    if (p != nullptr) p[row * Lp + j] = bf8_pack(m);
                }
            }
            if constexpr (ROUTE) {
                const float tg = warp_sum8_transposed(acc, lane), tc = warp_sum8_transposed(cnt, lane),
                            ts = warp_sum8_transposed(sa, lane);
                if ((lane & 3) == 0) {
                    const size_t o = (size_t)(b0 + bl) * C + cc * 8 + 4 * ((lane >> 4) & 1) + 2 * ((lane >> 3) & 1) + ((lane >> 2) & 1);
                    gap[o] = tg / (float)Lp;
                    route[o] = tc;
                    route[(size_t)B * C + o] = ts;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
                if (lane < 8) {
                    float t = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (i == lane) t = acc[i];
                    gap[(size_t)(b0 + bl) * C + cc * 8 + lane] = t / (float)Lp;
                }
            }
        }
    }
}

// yb blocked bf16; stat_part float[nparts][2][C] from ecgb200_conv1d_fwd_stats_bf16.
static int bn_fwd_train_launch(const void* yb, const float* stat_part, int nparts,
                                                   const float* gamma, const float* beta, float* running_mean,
                                                   float* running_var, int64_t* nbt, float* bn_state, void* pb,
                                                   float* gap, int B, int C, int L, float momentum, float eps,
                                                   int nrep, float* route, void* stream) {
    if (!yb || !stat_part || nparts <= 0 || !gamma || !beta || !bn_state || (!pb && !gap) || B <= 0 || C <= 0 ||
        (C & 7) || L < 2 || nrep < 1)
        return ECGB200_EINVAL;
    const int Lp = L / 2;
    const int tile_b = bnb_tile_b(B, C, 3), NS = (B + tile_b - 1) / tile_b;
    cudaStream_t st = (cudaStream_t)stream;
    if (gap != nullptr && route != nullptr)
        return ecg_launch_pdl(bn_fwd_train_bf16_kernel<true, true>, dim3(C / 8, NS), dim3(256), 0, st, (const uint4*)yb,
                              stat_part, nparts, gamma, beta, running_mean, running_var, nbt, bn_state, (uint4*)pb, gap,
                              B, C, L, Lp, tile_b, momentum, eps, nrep, route);
    if (gap != nullptr)
        return ecg_launch_pdl(bn_fwd_train_bf16_kernel<true, false>, dim3(C / 8, NS), dim3(256), 0, st, (const uint4*)yb,
                              stat_part, nparts, gamma, beta, running_mean, running_var, nbt, bn_state, (uint4*)pb, gap,
                              B, C, L, Lp, tile_b, momentum, eps, nrep, (float*)nullptr);
    return ecg_launch_pdl(bn_fwd_train_bf16_kernel<false, false>, dim3(C / 8, NS), dim3(256), 0, st, (const uint4*)yb,
                          stat_part, nparts, gamma, beta, running_mean, running_var, nbt, bn_state, (uint4*)pb,
                          (float*)nullptr, B, C, L, Lp, tile_b, momentum, eps, nrep, (float*)nullptr);
}

extern "C" int ecgb200_bn_relu_pool_fwd_train_bf16(const void* yb, const float* stat_part, int nparts,
                                                   const float* gamma, const float* beta, float* running_mean,
                                                   float* running_var, int64_t* nbt, float* bn_state, void* pb,
                                                   float* gap, int B, int C, int L, float momentum, float eps,
                                                   int nrep, void* stream) {
    return bn_fwd_train_launch(yb, stat_part, nparts, gamma, beta, running_mean, running_var, nbt, bn_state, pb, gap, B, C, L,
                               momentum, eps, nrep, nullptr, stream);
}

// Last block (gap != NULL) with the routing summary route[2][B][C] for ecgb200_bn_relu_pool_bwd_route_bf16.
extern "C" int ecgb200_bn_relu_pool_fwd_train_route_bf16(const void* yb, const float* stat_part, int nparts,
                                                         const float* gamma, const float* beta, float* running_mean,
                                                         float* running_var, int64_t* nbt, float* bn_state, void* pb,
                                                         float* gap, float* route, int B, int C, int L, float momentum,
                                                         float eps, int nrep, void* stream) {
    if (!gap || !route) return ECGB200_EINVAL;
    return bn_fwd_train_launch(yb, stat_part, nparts, gamma, beta, running_mean, running_var, nbt, bn_state, pb, gap, B, C, L,
                               momentum, eps, nrep, route, stream);
}

// Routing of one pool pair in terms of the raw conv outputs a0, a1 (first index wins ties, ReLU mask):
//   sel0 = relu(bn(a0)) >= relu(bn(a1)) && relu(bn(a0)) > 0 ;  sel1 = relu(bn(a1)) > relu(bn(a0))
__device__ __forceinline__ void pool_sel(float a0, float a1, float sc, float sf, bool& sel0, bool& sel1) {
    const float r0 = fmaxf(fmaf(a0, sc, sf), 0.f), r1 = fmaxf(fmaf(a1, sc, sf), 0.f);
    sel0 = (r0 >= r1) && (r0 > 0.f);
    sel1 = r1 > r0;
}

__device__ __forceinline__ void load_dgrad8(const uint4* __restrict__ dp, const float* __restrict__ dgap,
                                            size_t row, int Lp, int j, int b, int C, int cc, float inv_lp, float* d) {
    if (dp != nullptr) {
        bf8_unpack(__ldg(dp + row * Lp + j), d);
    } else {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(dgap + (size_t)b * C + cc * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(dgap + (size_t)b * C + cc * 8) + 1);
        d[0] = g0.x * inv_lp; d[1] = g0.y * inv_lp; d[2] = g0.z * inv_lp; d[3] = g0.w * inv_lp;
        d[4] = g1.x * inv_lp; d[5] = g1.y * inv_lp; d[6] = g1.z * inv_lp; d[7] = g1.w * inv_lp;
    }
}

// backward pass 1: partial {sum g, sum g*a} per (channel, split) -> part[c][s], part[C+c][s]
// (a = raw conv output; sum g*xhat = rstd * (sum g*a - mean * sum g) is formed when merging)
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_bf16_kernel(const uint4* __restrict__ y, const float* __restrict__ bn_state,
                          const uint4* __restrict__ dp, const float* __restrict__ dgap,
                          float* __restrict__ part, int B, int C, int L, int Lp, int tile_b) {
    __shared__ float sh[8 * 16];
    ecg_pdl_launch_dependents();            // the apply kernel may be scheduled (and run its prologue) under this one
    ecg_pdl_wait();                         // dp / dgap come from the kernel launched just before (no-op without PDL)
    const int cc = blockIdx.x, NS = gridDim.y;
    const int b0 = blockIdx.y * tile_b, nb = min(tile_b, B - b0);
    const float inv_lp = 1.0f / (float)Lp;
    float sc[8], sf[8], s[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = __ldg(bn_state + 2 * C + cc * 8 + i);
        sf[i] = __ldg(bn_state + 3 * C + cc * 8 + i);
        s[i] = 0.f; s[8 + i] = 0.f;
    }
    // flattened (sample, pool pair) walk, two pairs per iteration so that six 16-byte loads are in flight
    const int n = nb * Lp;
    for (int idx = threadIdx.x; idx < n; idx += 2 * blockDim.x) {
        const int idx2 = idx + blockDim.x;
        const bool has2 = idx2 < n;
        const int bl = idx / Lp, j = idx - bl * Lp;
        const int bl2 = has2 ? idx2 / Lp : bl, j2 = has2 ? idx2 - bl2 * Lp : j;
        const size_t row = (size_t)(b0 + bl) * (C / 8) + cc, row2 = (size_t)(b0 + bl2) * (C / 8) + cc;
        const uint4 u0 = __ldg(y + row * L + 2 * j), u1 = __ldg(y + row * L + 2 * j + 1);
        const uint4 w0 = __ldg(y + row2 * L + 2 * j2), w1 = __ldg(y + row2 * L + 2 * j2 + 1);
        float d[8], e[8];
        load_dgrad8(dp, dgap, row, Lp, j, b0 + bl, C, cc, inv_lp, d);
        load_dgrad8(dp, dgap, row2, Lp, j2, b0 + bl2, C, cc, inv_lp, e);
        float a0[8], a1[8];
        bf8_unpack(u0, a0);
        bf8_unpack(u1, a1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bool s0, s1;
            pool_sel(a0[i], a1[i], sc[i], sf[i], s0, s1);
            const float g = (s0 || s1) ? d[i] : 0.f;
            s[i] += g;
            s[8 + i] = fmaf(g, s0 ? a0[i] : a1[i], s[8 + i]);
        }
        if (has2) {
            bf8_unpack(w0, a0);
            bf8_unpack(w1, a1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bool s0, s1;
                pool_sel(a0[i], a1[i], sc[i], sf[i], s0, s1);
                const float g = (s0 || s1) ? e[i] : 0.f;
                s[i] += g;
                s[8 + i] = fmaf(g, s0 ? a0[i] : a1[i], s[8 + i]);
            }
        }
    }
    block_sum_vec<16>(s, sh);
    if (threadIdx.x < 8) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (i == threadIdx.x) { t1 = s[i]; t2 = s[8 + i]; }
        part[((size_t)blockIdx.y * 2 + 0) * C + cc * 8 + threadIdx.x] = t1;
        part[((size_t)blockIdx.y * 2 + 1) * C + cc * 8 + threadIdx.x] = t2;
    }
}

// backward pass 2: every block first merges the NS partials of its 8 channels (fixed order, double),
// block row 0 also publishes dgamma / dbeta; then
//   dy = scale*g - scale*m1 - scale*rstd*m2*(a - mean) = scale*g + A*a + Bc      (blocked bf16)
// and per-(channel, split) sums of dy for the conv-bias gradient.
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_bf16_kernel(const uint4* __restrict__ y, const float* __restrict__ bn_state,
                         const uint4* __restrict__ dp, const float* __restrict__ dgap,
                         const float* __restrict__ part, uint4* __restrict__ dy, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, float* __restrict__ db_part, int B, int C, int L, int Lp,
                         float inv_n, int train, int tile_b, int nparts, int local_idx, const float* __restrict__ route) {
    __shared__ float sh[8 * 8];
    __shared__ float cA[8], cB[8];
    __shared__ double shd[32 * 16], mom[16];
    const int cc = blockIdx.x, NS = gridDim.y;
    const int b0 = blockIdx.y * tile_b, nb = min(tile_b, B - b0);
    const float inv_lp = 1.0f / (float)Lp;
    ecg_pdl_wait();                         // `part` comes from the reduce kernel launched just before
    if (route != nullptr) merge_route8(route, dgap, B, C, cc, inv_lp, shd, mom);   // block 4: no reduce pass over y
    else merge_parts8(part, nparts > 0 ? nparts : NS, C, cc, shd, mom);     // nparts > 0: an exchanged (SyncBN) partial list
    if (threadIdx.x < 8) {
        const int c = cc * 8 + threadIdx.x;
        const double sg = mom[threadIdx.x], sga = mom[8 + threadIdx.x];
        const float mean = __ldg(bn_state + c), rstd = __ldg(bn_state + C + c), scl = __ldg(bn_state + 2 * C + c);
        const double sgx = (double)rstd * (sga - (double)mean * sg);       // sum g * xhat
        if (blockIdx.y == 0) {
            // SyncBN (local_idx >= 0): `part` holds one {sum g, sum g*a} pair per replica; the batch means below use all
            // of them, the affine gradients only this replica's own pair (the gradient exchange sums them over replicas)
            double lg = sg, lgx = sgx;
            if (local_idx >= 0) {
                lg = (double)__ldg(part + ((size_t)local_idx * 2 + 0) * C + c);
                lgx = (double)rstd * ((double)__ldg(part + ((size_t)local_idx * 2 + 1) * C + c) - (double)mean * lg);
            }
            if (dbeta != nullptr) dbeta[c] = (float)lg;
            if (dgamma != nullptr) dgamma[c] = (float)lgx;
        }
        const float m1 = train ? (float)sg * inv_n : 0.f, m2 = train ? (float)sgx * inv_n : 0.f;
        const float A = -scl * rstd * m2;
        cA[threadIdx.x] = A;
        cB[threadIdx.x] = -scl * m1 - A * mean;
    }
    __syncthreads();
    float sc[8], sf[8], A[8], Bc[8], sdy[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = __ldg(bn_state + 2 * C + cc * 8 + i);
        sf[i] = __ldg(bn_state + 3 * C + cc * 8 + i);
        A[i] = cA[i]; Bc[i] = cB[i];
        sdy[i] = 0.f;
    }
    int bl = 0, j = threadIdx.x;
    while (j >= Lp) { j -= Lp; ++bl; }
    while (bl < nb) {
        const size_t row = (size_t)(b0 + bl) * (C / 8) + cc;
        float a0[8], a1[8], d[8], o0[8], o1[8];
        const uint4 u0 = __ldg(y + row * L + 2 * j), u1 = __ldg(y + row * L + 2 * j + 1);
        load_dgrad8(dp, dgap, row, Lp, j, b0 + bl, C, cc, inv_lp, d);
        bf8_unpack(u0, a0);
        bf8_unpack(u1, a1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bool s0, s1;
            pool_sel(a0[i], a1[i], sc[i], sf[i], s0, s1);
            const float gd = sc[i] * d[i];
            o0[i] = fmaf(A[i], a0[i], Bc[i]) + (s0 ? gd : 0.f);
            o1[i] = fmaf(A[i], a1[i], Bc[i]) + (s1 ? gd : 0.f);
            sdy[i] += o0[i] + o1[i];
        }
        dy[row * L + 2 * j] = bf8_pack(o0);
        dy[row * L + 2 * j + 1] = bf8_pack(o1);
        if (j == Lp - 1 && (L & 1)) {
            float a[8], o[8];
            bf8_unpack(__ldg(y + row * L + L - 1), a);
#pragma unroll
            for (int i = 0; i < 8; ++i) { o[i] = fmaf(A[i], a[i], Bc[i]); sdy[i] += o[i]; }
            dy[row * L + L - 1] = bf8_pack(o);
        }
        j += blockDim.x;
        while (j >= Lp) { j -= Lp; ++bl; }
    }
    if (db_part != nullptr) {
        block_sum_vec<8>(sdy, sh);
        if (threadIdx.x < 8) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) if (i == threadIdx.x) t = sdy[i];
            db_part[(size_t)(cc * 8 + threadIdx.x) * NS + blockIdx.y] = t;
        }
    }
}

// yb, dpb, dyb blocked bf16; dgap fp32 (B,C) [layer 4]; db_part fp32 [C][ecgb200_bn_nsplit(B,C)] or
// NULL; ws: ecgb200_bn_bwd_ws_bytes(B,C).
extern "C" int ecgb200_bn_relu_pool_bwd_bf16(const void* yb, const float* bn_state, const void* dpb,
                                             const float* dgap, void* dyb, float* dgamma, float* dbeta,
                                             float* db_part, void* ws, int B, int C, int L, int train,
                                             void* stream) {
    if (!yb || !bn_state || (!dpb && !dgap) || !dyb || !ws || B <= 0 || C <= 0 || (C & 7) || L < 2) return ECGB200_EINVAL;
    if (dgap != nullptr && (((uintptr_t)dgap & 15) != 0)) return ECGB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int Lp = L / 2;
    const int tile_b = bnb_tile_b(B, C, 3), NS = (B + tile_b - 1) / tile_b;
    float* part = (float*)ws;
    int rc = ecg_launch_pdl(bn_bwd_reduce_bf16_kernel, dim3(C / 8, NS), dim3(256), 0, st, (const uint4*)yb, bn_state,
                            (const uint4*)dpb, dgap, part, B, C, L, Lp, tile_b);
    if (rc) return rc;
    const float inv_n = 1.0f / ((float)B * (float)L);
    // programmatic dependent launch of the apply pass: its blocks are scheduled while the reduce pass drains and
    // wait (griddepcontrol.wait = full completion + flush of the reduce kernel) before they read `part`
    return ecg_launch_pdl_if(true, bn_bwd_apply_bf16_kernel, dim3(C / 8, NS), dim3(256), 0, st, (const uint4*)yb, bn_state,
                             (const uint4*)dpb, dgap, (const float*)part, (uint4*)dyb, dgamma, dbeta, db_part, B, C, L, Lp,
                             inv_n, train, tile_b, 0, -1, (const float*)nullptr);
}

// Block 4 in ONE pass: the batch reductions come from the routing summary the forward pass left
// (ecgb200_bn_relu_pool_fwd_train_route_bf16), so y is read once (to write dy) instead of twice.
extern "C" int ecgb200_bn_relu_pool_bwd_route_bf16(const void* yb, const float* bn_state, const float* dgap,
                                                   const float* route, void* dyb, float* dgamma, float* dbeta,
                                                   float* db_part, int B, int C, int L, int train, void* stream) {
    if (!yb || !bn_state || !dgap || !route || !dyb || B <= 0 || C <= 0 || (C & 7) || L < 2) return ECGB200_EINVAL;
    if (((uintptr_t)dgap & 15) != 0) return ECGB200_EINVAL;
    const int Lp = L / 2;
    const int tile_b = bnb_tile_b(B, C, 3), NS = (B + tile_b - 1) / tile_b;
    const float inv_n = 1.0f / ((float)B * (float)L);
    return ecg_launch_pdl(bn_bwd_apply_bf16_kernel, dim3(C / 8, NS), dim3(256), 0, (cudaStream_t)stream, (const uint4*)yb,
                          bn_state, (const uint4*)nullptr, dgap, (const float*)nullptr, (uint4*)dyb, dgamma, dbeta, db_part, B, C,
                          L, Lp, inv_n, train, tile_b, 0, -1, route);
}

// The two passes as separate calls, for SyncBN under data parallel: pass 1 leaves this replica's partials
// part[ecgb200_bn_nsplit(B,C)][2][C] = {sum g, sum g*a}; the caller exchanges them (ecgb200_dp_bn_sync_f32 -> one pair per
// replica) and pass 2 takes the exchanged list: batch means over all `nrep` replicas' B*L samples, affine gradients from
// pair `local_idx` (this replica's own) only.
extern "C" int ecgb200_bn_relu_pool_bwd_reduce_bf16(const void* yb, const float* bn_state, const void* dpb,
                                                    const float* dgap, float* part, int B, int C, int L,
                                                    void* stream) {
    if (!yb || !bn_state || (!dpb && !dgap) || !part || B <= 0 || C <= 0 || (C & 7) || L < 2) return ECGB200_EINVAL;
    if (dgap != nullptr && (((uintptr_t)dgap & 15) != 0)) return ECGB200_EINVAL;
    const int Lp = L / 2;
    const int tile_b = bnb_tile_b(B, C, 3), NS = (B + tile_b - 1) / tile_b;
    bn_bwd_reduce_bf16_kernel<<<dim3(C / 8, NS), 256, 0, (cudaStream_t)stream>>>((const uint4*)yb, bn_state, (const uint4*)dpb,
                                                                                dgap, part, B, C, L, Lp, tile_b);
    return ecg_launch_status();
}
extern "C" int ecgb200_bn_relu_pool_bwd_apply_bf16(const void* yb, const float* bn_state, const void* dpb,
                                                   const float* dgap, const float* part, int nparts, int local_idx,
                                                   int nrep, void* dyb, float* dgamma, float* dbeta, float* db_part,
                                                   int B, int C, int L, int train, void* stream) {
    if (!yb || !bn_state || (!dpb && !dgap) || !part || nparts <= 0 || local_idx >= nparts || nrep < 1 || !dyb || B <= 0 ||
        C <= 0 || (C & 7) || L < 2)
        return ECGB200_EINVAL;
    if (dgap != nullptr && (((uintptr_t)dgap & 15) != 0)) return ECGB200_EINVAL;
    const int Lp = L / 2;
    const int tile_b = bnb_tile_b(B, C, 3), NS = (B + tile_b - 1) / tile_b;
    const float inv_n = 1.0f / ((float)B * (float)L * (float)nrep);
    return ecg_launch_pdl_if(false, bn_bwd_apply_bf16_kernel, dim3(C / 8, NS), dim3(256), 0, (cudaStream_t)stream,
                             (const uint4*)yb, bn_state, (const uint4*)dpb, dgap, part, (uint4*)dyb, dgamma,
                             dbeta, db_part, B, C, L, Lp, inv_n, train, tile_b, nparts, local_idx, (const float*)nullptr);
}
