// Data-parallel gradient exchange FUSED with the optimizer, over NVLink peer memory (one launch per rank).
//
// The reference is single-device (no DDP); the data-parallel contract it implies is "every rank
// applies torch.optim.AdamW (loop.py:34) to the MEAN of the per-rank gradients".  Instead of NCCL
// all-reduce followed by a replicated AdamW, every rank owns 1/world of the flat parameter space and one
// kernel does
//     barrier(all ranks' gradients final)                       [flags in peer memory, system scope]
//  -> g = sum over ranks of G_r[i]   (peer LOADS over NVLink, fixed rank order: deterministic)
//  -> AdamW on the owned shard (moments live only on the owner: optimizer state is sharded)
//  -> new p stored into EVERY rank's parameter buffer (peer STORES over NVLink)
//  -> barrier(all shards written)
// i.e. reduce-scatter + optimizer + all-gather in one pass: 2*(world-1)/world * 4 B/param cross the links
// (2.5 MB at 8 ranks: ~7 us at the measured 770 GB/s) and the optimizer touches 1/world of the state.
#include "common.cuh"

constexpr int DP_MAX_WORLD = 8;

struct DpPeers {
    float* p[DP_MAX_WORLD];                  // every rank's flat parameter buffer (peer mapped)
    const float* g[DP_MAX_WORLD];            // every rank's flat gradient buffer
    unsigned int* flags[DP_MAX_WORLD];       // every rank's flag pad: [0,W) entry flags, [W,2W) exit flags,
                                             // [2W] block counter, [2W+1] calls completed (the barrier epoch)
};

__device__ __forceinline__ void dp_st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int dp_ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dp_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Cross-rank waits: a rank may legitimately arrive much later than its peers (rank 0 runs validation or writes a
// checkpoint between epochs, a loader stalls), so the limit is minutes by default -- NCCL would simply wait -- and
// ecgb200_set_spin_timeout_ms(0) switches it off.  It exists so that a protocol bug or a dead peer ends in a trap
// instead of a GPU that spins for ever.
static __device__ unsigned long long g_dp_timeout_ns = 600000000000ull;
void ecg_set_timeout_dp(unsigned long long ns) { cudaMemcpyToSymbol(g_dp_timeout_ns, &ns, sizeof(ns)); }
// wait until flag >= epoch (monotonic epochs: no reset race)
__device__ __forceinline__ void dp_wait_flag(const unsigned int* f, unsigned int epoch) {
    unsigned long long t0 = 0;
    for (unsigned it = 0;; ++it) {
        if ((int)(dp_ld_acquire_sys(f) - epoch) >= 0) return;
        if ((it & 255u) == 255u) {
            const unsigned long long now = dp_globaltimer(), limit = g_dp_timeout_ns;
            if (t0 == 0) t0 = now;
            else if (limit != 0 && now - t0 > limit) __trap();
        }
    }
}

// 128 threads per block: the bucket-A exchange runs beside the tensor kernels of blocks 3..1, and 256 threads x 48
// registers did not fit next to a conv CTA (320 x 168) in the register file -- its blocks then took whole SMs away from
// the persistent dgrad grid while they spun on the entry barrier (dgrad_3: 29.6 -> 47.9 us in the traced schedule).
__global__ void __launch_bounds__(128)
dp_adamw_fused_kernel(const __grid_constant__ DpPeers Q, float* __restrict__ m, float* __restrict__ v, long long off,
                      long long n, int rank, int world, const float* __restrict__ hyper,
                      const int* __restrict__ step_now) {
    __shared__ float S[8];
    unsigned int* myflags = Q.flags[rank];
    // barrier epoch = number of this call (1-based), kept in the flag pad so that it does not depend on the
    // optimizer step (resume from a checkpoint); every rank makes the same number of calls
    const unsigned int epoch = myflags[2 * world + 1] + 1u;
    // ---- entry barrier: my gradients are final (previous kernels of this stream) -> tell every rank, wait for all
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        dp_st_release_sys(Q.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x < world) dp_wait_flag(myflags + threadIdx.x, epoch);
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
        S[7] = hyper[5];                                     // gscale = 1 / world
    }
    __syncthreads();
    const float decay = S[0], one_m_b1 = S[1], b2 = S[2], one_m_b2 = S[3], bc2 = S[4], eps = S[5], ss = S[6],
                gscale = S[7];
    // ---- owned shard of the bucket [off, off + n), in float4 units (off % 4 == 0, n % (4 * world) == 0)
    const long long n4 = n >> 2;
    const long long per = n4 / world;
    const long long lo = (off >> 2) + per * rank, hi = lo + per;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float* myp = Q.p[rank];
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r) {
            if (r < world) {
                const float4 t = __ldcv(reinterpret_cast<const float4*>(Q.g[r]) + i);      // peer load, not cached
                g4.x = __fadd_rn(g4.x, t.x); g4.y = __fadd_rn(g4.y, t.y); g4.z = __fadd_rn(g4.z, t.z); g4.w = __fadd_rn(g4.w, t.w);
            }
        }
        const float4 p4 = reinterpret_cast<const float4*>(myp)[i];
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        const float gg[4] = {__fmul_rn(g4.x, gscale), __fmul_rn(g4.y, gscale), __fmul_rn(g4.z, gscale), __fmul_rn(g4.w, gscale)};
        float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
        const AdamK K = {decay, one_m_b1, b2, one_m_b2, bc2, eps, ss};
#pragma unroll
        for (int e = 0; e < 4; ++e) adamw_update(pp[e], mm[e], vv[e], gg[e], K);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        const float4 out = make_float4(pp[0], pp[1], pp[2], pp[3]);
#pragma unroll
        for (int r = 0; r < DP_MAX_WORLD; ++r)
            if (r < world) reinterpret_cast<float4*>(Q.p[r])[i] = out;                      // peer store (all-gather)
    }
    // ---- exit barrier: when the LAST block of this rank is done, tell every rank; block 0 holds the kernel open
    // until every rank's shard has landed here (the next kernel of this stream reads the parameters)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(myflags + 2 * world, 1u) + 1u;
        if (done == epoch * gridDim.x) {                      // last block of this rank: every block has fenced its stores
            __threadfence();
            myflags[2 * world + 1] = epoch;
            for (int r = 0; r < world; ++r) dp_st_release_sys(Q.flags[r] + world + rank, epoch);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < world) dp_wait_flag(myflags + world + threadIdx.x, epoch);
}

extern "C" int ecgb200_dp_flag_words(int world) { return 2 * world + 2; }

// fixed grid per (bucket size, world): the exit barrier counts epoch * gridDim.x block arrivals, so a flag pad must
// always be used with the same bucket
static int dp_grid(long long n, int world) {
    long long b = (n / 4 / world + 255) / 256;             // ~2 float4 per thread
    return (int)(b < 8 ? 8 : (b > 120 ? 120 : b));
}

// p / g / flags: HOST arrays of `world` peer-mapped device pointers (this rank's own buffers at index `rank`).
// m, v: this rank's Adam moments (full length; only the owned shard of the bucket is used).  The bucket is the
// element range [off, off + n) of the flat buffers: off % 4 == 0, n % (4 * world) == 0; rank r owns
// [off + r*n/world, off + (r+1)*n/world).  flags: >= ecgb200_dp_flag_words(world) zero-initialised uint32 per rank,
// ONE PAD PER BUCKET (the call count on a pad is its barrier epoch).  hyper[5] must hold 1/world.  *step_now is the
// 1-based optimizer step index (bias correction), identical on every rank.  Every rank must make the same sequence of
// calls on a pad.
extern "C" int ecgb200_dp_adamw_fused_range_f32(float* const* p, const float* const* g, unsigned int* const* flags,
                                                float* m, float* v, int64_t off, int64_t n, int rank, int world,
                                                const float* hyper, const int* step_now, void* stream) {
    if (!p || !g || !flags || !m || !v || !hyper || !step_now || n <= 0 || off < 0) return ECGB200_EINVAL;
    if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return ECGB200_EUNSUPPORTED;
    if (n % (4LL * world) != 0 || (off & 3) != 0) return ECGB200_EINVAL;
    DpPeers Q;
    for (int r = 0; r < DP_MAX_WORLD; ++r) {
        Q.p[r] = r < world ? p[r] : nullptr;
        Q.g[r] = r < world ? g[r] : nullptr;
        Q.flags[r] = r < world ? flags[r] : nullptr;
        if (r < world && (!Q.p[r] || !Q.g[r] || !Q.flags[r])) return ECGB200_EINVAL;
        if (r < world && ((((uintptr_t)Q.p[r] | (uintptr_t)Q.g[r]) & 15) != 0)) return ECGB200_EINVAL;
    }
    if ((((uintptr_t)m | (uintptr_t)v) & 15) != 0) return ECGB200_EINVAL;
    dp_adamw_fused_kernel<<<dp_grid(n, world), 128, 0, (cudaStream_t)stream>>>(Q, m, v, (long long)off, (long long)n, rank,
                                                                             world, hyper, step_now);
    return ecg_launch_status();
}

// The whole flat space as one bucket.
extern "C" int ecgb200_dp_adamw_fused_f32(float* const* p, const float* const* g, unsigned int* const* flags,
                                          float* m, float* v, int64_t n, int rank, int world, const float* hyper,
                                          const int* step_now, void* stream) {
    return ecgb200_dp_adamw_fused_range_f32(p, g, flags, m, v, 0, n, rank, world, hyper, step_now, stream);
}

// ---------------------------------------------------------------- SyncBN statistics exchange
// Train-mode BatchNorm over the GLOBAL batch under data parallel (the reference is single-device: its statistics
// cover the whole batch, src/models/ecg_cnn.py:14): every replica reduces its own partial pairs
// local_part[nparts][2][C] (forward: {sum y, sum y^2} from the conv epilogue; backward: {sum g, sum g*a} from pass 1 of
// the BatchNorm backward) to ONE pair, publishes it in its exchange slot (peer memory), passes a cross-rank barrier and
// gathers every replica's pair in rank order into out[world][2][C] -- which the BatchNorm kernels then merge exactly
// like per-CTA partials (fixed order => every replica computes bit-identical statistics).
struct BnPeers {
    float* slot[DP_MAX_WORLD];               // every rank's exchange slot (>= 2*C floats, peer mapped)
    unsigned int* flags[DP_MAX_WORLD];       // every rank's flag pad (layout as DpPeers)
};

__global__ void __launch_bounds__(1024)
dp_bn_sync_kernel(const __grid_constant__ BnPeers Q, const float* __restrict__ local_part, int nparts, int C,
                  float* __restrict__ out, int rank, int world) {
    unsigned int* myflags = Q.flags[rank];
    const unsigned int epoch = myflags[2 * world + 1] + 1u;
    const int i = threadIdx.x;                       // i < 2*C: (which, channel)
    // local reduce: 1024 threads = up to 512 (which, channel) columns x >= 2 partial groups, independent loads in flight
    // (one thread per column walking 148 partials serially took 22 us); groups combined in a fixed order
    __shared__ double acc[1024];
    {
        const int ncol = 2 * C, ngrp = 1024 / ncol;               // C <= 256, power-of-two widths: ngrp >= 2
        const int col = threadIdx.x % ncol, grp = threadIdx.x / ncol;
        double s = 0.0;
        if (grp < ngrp) {
            const int which = col / C, c = col - which * C;
            for (int j = grp; j < nparts; j += ngrp) s += (double)__ldg(local_part + ((size_t)j * 2 + which) * C + c);
        }
        acc[threadIdx.x] = s;
        __syncthreads();
        if (i < ncol) {
            double t = 0.0;
            for (int g = 0; g < ngrp; ++g) t += acc[g * ncol + i];
            Q.slot[rank][i] = (float)t;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        dp_st_release_sys(Q.flags[threadIdx.x] + rank, epoch);
        dp_wait_flag(myflags + threadIdx.x, epoch);
    }
    __syncthreads();
    if (i < 2 * C) {
        const int which = i / C, c = i - which * C;
        for (int r = 0; r < world; ++r)
            out[((size_t)r * 2 + which) * C + c] = __ldcv(Q.slot[r] + i);          // peer load, not cached
    }
    if (threadIdx.x == 0) myflags[2 * world + 1] = epoch;
}

// slots / flags: HOST arrays of `world` peer-mapped pointers; a slot holds >= 2*C floats and must not be reused by
// another exchange before every rank has passed a later cross-rank barrier (the engine gives every (block, direction)
// its own slot; the optimizer exchange at the end of the step is that barrier).  flags: one pad
// (ecgb200_dp_flag_words) shared by all BatchNorm exchanges of the step.  C <= 256.
extern "C" int ecgb200_dp_bn_sync_f32(const float* local_part, int nparts, int C, float* const* slots,
                                      unsigned int* const* flags, float* out, int rank, int world, void* stream) {
    if (!local_part || nparts <= 0 || C <= 0 || !slots || !flags || !out) return ECGB200_EINVAL;
    if (C > 256 || (C & 7) || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return ECGB200_EUNSUPPORTED;
    BnPeers Q;
    for (int r = 0; r < DP_MAX_WORLD; ++r) {
        Q.slot[r] = r < world ? slots[r] : nullptr;
        Q.flags[r] = r < world ? flags[r] : nullptr;
        if (r < world && (!Q.slot[r] || !Q.flags[r])) return ECGB200_EINVAL;
    }
    dp_bn_sync_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(Q, local_part, nparts, C, out, rank, world);
    return ecg_launch_status();
}

// ================================================================ one-hop ("LL") exchange
// The barrier form above costs two cross-GPU barriers per exchange (fence + flag store + poll: ~5 us each, measured
// 21 us for an EMPTY bucket on two B200s) because data and "data is ready" travel separately.  Here every 4-byte payload
// travels WITH its flag in one 8-byte word {value, epoch} (8-byte stores are single-copy atomic, also over NVLink), so a
// receiver simply polls the word it needs: no fences, no barriers, one NVLink hop per phase.
//   phase 1  every rank PUSHES its gradients of shard j into owner j's inbox        gin[parity][sender][i]   (peer stores)
//   phase 2  the owner polls its inbox, sums in rank order (deterministic, bit-identical to the barrier form), applies
//            AdamW to its shard and PUSHES the new parameters to every rank's       pin[parity][i]           (peer stores)
//   phase 3  every rank polls its parameter inbox and unpacks it into its replica.
// Inboxes are double-buffered by epoch parity: a rank cannot start exchange e+2 before every rank has finished reading
// exchange e (it needs their parameters of e+1, which they send only after consuming e), so no word is overwritten early.
struct LlPeers {
    unsigned long long* gin[DP_MAX_WORLD];      // every rank's gradient inbox  [2][world][n / world] words
    unsigned long long* pin[DP_MAX_WORLD];      // every rank's parameter inbox [2][n] words
};

__device__ __forceinline__ void ll_store(unsigned long long* p, float v, unsigned int flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(flag) : "memory");
}
__device__ __forceinline__ float ll_wait(const unsigned long long* p, unsigned int flag) {
    unsigned int d, f;
    unsigned long long t0 = 0;
    for (unsigned it = 0;; ++it) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(d), "=r"(f) : "l"(p) : "memory");
        if (f == flag) return __uint_as_float(d);
        if ((it & 255u) == 255u) {
            const unsigned long long now = dp_globaltimer(), limit = g_dp_timeout_ns;
            if (t0 == 0) t0 = now;
            else if (limit != 0 && now - t0 > limit) __trap();
        }
    }
}

// U words `stride` apart, loads issued together (the polls of one index are a dependent L2 round trip each; four indices
// in flight per thread hide most of it); words that have not arrived yet fall back to the spinning wait
template <int U>
__device__ __forceinline__ void ll_wait_n(const unsigned long long* base, long long stride, int cnt, unsigned int flag,
                                          float* out) {
    unsigned int d[U], f[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (u < cnt)
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(d[u]), "=r"(f[u]) : "l"(base + u * stride) : "memory");
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (u < cnt) out[u] = f[u] == flag ? __uint_as_float(d[u]) : ll_wait(base + u * stride, flag);
}

constexpr int LL_U = 4;
__global__ void __launch_bounds__(128)
dp_adamw_ll_kernel(const __grid_constant__ LlPeers Q, float* __restrict__ p, const float* __restrict__ g,
                   float* __restrict__ m, float* __restrict__ v, unsigned int* __restrict__ ctr, long long off,
                   long long n, int rank, int world, const float* __restrict__ hyper, const int* __restrict__ step_now) {
    __shared__ float S[8];
    // ctr[0] = exchanges completed on this inbox pair (the epoch), ctr[1] = blocks finished
    const unsigned int epoch = ctr[0] + 1u;
    const long long per = n / world;
    const long long par = (long long)(epoch & 1u);
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    // ---- phase 1: my gradients of shard j -> owner j's inbox row `rank`
    for (int jj = 0; jj < world; ++jj) {
        const int j = (rank + jj) % world;                        // start with my own shard, then fan out round-robin
        unsigned long long* dst = Q.gin[j] + (par * world + rank) * per;
        const float* src = g + off + (long long)j * per;
        for (long long i = tid; i < per; i += nth) ll_store(dst + i, src[i], epoch);
    }
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
        S[7] = hyper[5];                                     // gscale = 1 / world
    }
    __syncthreads();
    const float decay = S[0], one_m_b1 = S[1], b2 = S[2], one_m_b2 = S[3], bc2 = S[4], eps = S[5], ss = S[6],
                gscale = S[7];
    // ---- phase 2: owned shard: gather the ranks' words in rank order, AdamW, push the new parameters
    const unsigned long long* inbox = Q.gin[rank] + par * world * per;
    const long long lo = off + per * rank;
    for (long long i0 = tid; i0 < per; i0 += LL_U * nth) {
        const int cnt = (int)((per - i0 + nth - 1) / nth < LL_U ? (per - i0 + nth - 1) / nth : LL_U);   // indices i0 + u * nth
        float gs[LL_U], pp[LL_U], mo[LL_U], vo[LL_U];
#pragma unroll
        for (int u = 0; u < LL_U; ++u) {
            gs[u] = 0.f;
            if (u < cnt) { pp[u] = p[lo + i0 + u * nth]; mo[u] = m[lo + i0 + u * nth]; vo[u] = v[lo + i0 + u * nth]; }
        }
        for (int r = 0; r < world; ++r) {                                     // rank order: deterministic
            float w[LL_U];
            ll_wait_n<LL_U>(inbox + (long long)r * per + i0, nth, cnt, epoch, w);
#pragma unroll
            for (int u = 0; u < LL_U; ++u) if (u < cnt) gs[u] = __fadd_rn(gs[u], w[u]);
        }
#pragma unroll
        for (int u = 0; u < LL_U; ++u) {
            if (u < cnt) {
                const long long i = i0 + u * nth;
                const AdamK K = {decay, one_m_b1, b2, one_m_b2, bc2, eps, ss};
                float pn = pp[u], mm = mo[u], vv = vo[u];
                adamw_update(pn, mm, vv, __fmul_rn(gs[u], gscale), K);
                m[lo + i] = mm;
                v[lo + i] = vv;
                p[lo + i] = pn;
                for (int jj = 1; jj < world; ++jj) {
                    const int j = (rank + jj) % world;
                    ll_store(Q.pin[j] + par * n + per * rank + i, pn, epoch);
                }
            }
        }
    }
    // ---- phase 3: the other shards' new parameters from my parameter inbox
    const unsigned long long* pbox = Q.pin[rank] + par * n;
    for (int jj = 1; jj < world; ++jj) {
        const int j = (rank + jj) % world;
        for (long long i0 = tid; i0 < per; i0 += LL_U * nth) {
            const int cnt = (int)((per - i0 + nth - 1) / nth < LL_U ? (per - i0 + nth - 1) / nth : LL_U);
            float w[LL_U];
            ll_wait_n<LL_U>(pbox + per * j + i0, nth, cnt, epoch, w);
#pragma unroll
            for (int u = 0; u < LL_U; ++u) if (u < cnt) p[off + per * j + i0 + u * nth] = w[u];
        }
    }
    // ---- epoch bookkeeping: the last block to finish publishes the new epoch for the next call
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(ctr + 1, 1u) + 1u;
        if (done == gridDim.x) { ctr[1] = 0u; __threadfence(); ctr[0] = epoch; }
    }
}

extern "C" size_t ecgb200_dp_ll_inbox_words(int64_t n) { return (size_t)(4 * n); }     // gin: 2*n words, pin: 2*n words

// One bucket [off, off + n) of the flat space through the one-hop exchange.  p, g, m, v: THIS rank's flat buffers (plain
// device memory); inbox: HOST array of `world` peer-mapped pointers to every rank's inbox of ecgb200_dp_ll_inbox_words(n)
// zero-initialised 8-byte words (gradient inbox first, parameter inbox behind it), one inbox per bucket; ctr: 2 zeroed
// uint32 in this rank's memory, one pair per bucket.  n % (4 * world) == 0, off % 4 == 0.  Same arithmetic, same
// summation order and therefore the same bits as ecgb200_dp_adamw_fused_range_f32.
extern "C" int ecgb200_dp_adamw_ll_f32(float* p, const float* g, float* m, float* v, void* const* inbox, unsigned int* ctr,
                                       int64_t off, int64_t n, int rank, int world, const float* hyper,
                                       const int* step_now, void* stream) {
    if (!p || !g || !m || !v || !inbox || !ctr || !hyper || !step_now || n <= 0 || off < 0) return ECGB200_EINVAL;
    if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return ECGB200_EUNSUPPORTED;
    if (n % (4LL * world) != 0 || (off & 3) != 0) return ECGB200_EINVAL;
    LlPeers Q;
    for (int r = 0; r < DP_MAX_WORLD; ++r) {
        Q.gin[r] = r < world ? (unsigned long long*)inbox[r] : nullptr;
        Q.pin[r] = r < world ? (unsigned long long*)inbox[r] + 2 * n : nullptr;
        if (r < world && (!inbox[r] || (((uintptr_t)inbox[r]) & 7) != 0)) return ECGB200_EINVAL;
    }
    // 128-thread blocks (they must fit next to a tensor-core CTA in the register file), ~4 words per thread and phase.
    // Cross-rank dependencies are block k <-> block k only (index i is always handled by thread i mod (grid * 128) on
    // every rank), so a partially resident grid cannot deadlock as long as blocks are dispatched in index order.
    long long b = (n / world + 511) / 512;
    const int grid = (int)(b < 8 ? 8 : (b > 148 ? 148 : b));
    dp_adamw_ll_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Q, p, g, m, v, ctr, (long long)off, (long long)n, rank, world,
                                                               hyper, step_now);
    return ecg_launch_status();
}

// SyncBN statistics exchange, one-hop form: every replica reduces its partial pairs to one pair and PUSHES it, as
// {value, epoch} words, into every replica's inbox row `rank`; each replica then polls its own inbox rows in rank order.
// inbox: HOST array of peer-mapped pointers to every rank's [2 parity][world][2 * 256] words for THIS exchange slot;
// ctr: 1 zeroed uint32 of this rank for this slot.
__global__ void __launch_bounds__(1024)
dp_bn_sync_ll_kernel(const __grid_constant__ LlPeers Q, const float* __restrict__ local_part, int nparts, int C,
                     float* __restrict__ out, unsigned int* __restrict__ ctr, int rank, int world) {
    __shared__ double acc[1024];
    const unsigned int epoch = ctr[0] + 1u;
    const int par = (int)(epoch & 1u);
    const int ncol = 2 * C, ngrp = 1024 / ncol;
    const int col = threadIdx.x % ncol, grp = threadIdx.x / ncol;
    double s = 0.0;
    if (grp < ngrp) {
        const int which = col / C, c = col - which * C;
        for (int j = grp; j < nparts; j += ngrp) s += (double)__ldg(local_part + ((size_t)j * 2 + which) * C + c);
    }
    acc[threadIdx.x] = s;
    __syncthreads();
    const int i = threadIdx.x;
    if (i < ncol) {
        double t = 0.0;
        for (int g = 0; g < ngrp; ++g) t += acc[g * ncol + i];
        const float val = (float)t;
        for (int jj = 0; jj < world; ++jj) {
            const int j = (rank + jj) % world;
            ll_store(Q.gin[j] + ((size_t)par * world + rank) * 512 + i, val, epoch);
        }
        const int which = i / C, c = i - which * C;
        const unsigned long long* inbox = Q.gin[rank] + (size_t)par * world * 512;
        for (int r = 0; r < world; ++r) out[((size_t)r * 2 + which) * C + c] = ll_wait(inbox + (size_t)r * 512 + i, epoch);
    }
    __syncthreads();
    if (threadIdx.x == 0) ctr[0] = epoch;
}

extern "C" int ecgb200_dp_bn_sync_ll_f32(const float* local_part, int nparts, int C, void* const* inbox,
                                         unsigned int* ctr, float* out, int rank, int world, void* stream) {
    if (!local_part || nparts <= 0 || C <= 0 || !inbox || !ctr || !out) return ECGB200_EINVAL;
    if (C > 256 || (C & 7) || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return ECGB200_EUNSUPPORTED;
    LlPeers Q;
    for (int r = 0; r < DP_MAX_WORLD; ++r) {
        Q.gin[r] = r < world ? (unsigned long long*)inbox[r] : nullptr;
        Q.pin[r] = nullptr;
        if (r < world && (!inbox[r] || (((uintptr_t)inbox[r]) & 7) != 0)) return ECGB200_EINVAL;
    }
    dp_bn_sync_ll_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(Q, local_part, nparts, C, out, ctr, rank, world);
    return ecg_launch_status();
}

