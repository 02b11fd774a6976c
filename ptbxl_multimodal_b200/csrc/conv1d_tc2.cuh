// Two-SM form (tcgen05 cta_group::2) of the streamed-weight Conv1d kernel: included by conv1d_tc.cu.
//
// A cluster of two CTAs on one TPC works on 2*R output tiles per group: CTA r owns tiles base + r*R .. base + r*R + R-1 (its
// own input tiles in its own shared memory, its own 128 TMEM lanes per tile), and ONE thread of the pair's rank-0 CTA issues
// every tcgen05.mma with M = 256.  The weight slab -- the B operand -- is split by output channel: each CTA stages only ITS HALF
// of the Co columns (wmap box {Co/2 rows of 16 B, 8 chunks}), and the hardware feeds both halves to both SMs.  Against the
// one-SM kernel that halves (a) the weight bytes each SM pulls from L2 per tile, (b) the B-operand shared-memory reads per
// MMA -- the shared-memory port, not the tensor pipe, paces the one-SM main loop -- and (c) the number of
// instructions the issuing thread has to get out.  It also makes one tile per CTA per group affordable for the 256-channel
// layer (R = 1, two accumulator stages in the 512 TMEM columns), so that the epilogue of group g runs under the MMAs of g+1.
//
// Barriers sit at the same shared-memory offsets in both CTAs:
//   wfull / xfull   leader's copy only: a thread of the leader arms it with the bytes of BOTH CTAs, both CTAs' TMA loads
//                   complete on it (cp.async.bulk.tensor ... .cta_group::2 with the leader's shared::cluster address)
//   wempty / accfull   both copies: tcgen05.commit ... multicast::cluster from the leader's MMA thread
//                   (accfull of group g also tells the epilogue warps that input buffer g % 2 may be refilled: they issue
//                   the input-tile loads, eight threads in parallel -- a single thread manages one TMA instruction per ~50 cycles)
//   accempty        leader's copy, 16 arrivals: the eight epilogue warps of each CTA (mbarrier.arrive on the mapa'd address)
// MODE 0: training forward (+ BatchNorm partial statistics, one partial per CTA); MODE 3: dgrad / plain conv;
// MODE 1 / 2: inference blocks (folded BatchNorm + ReLU + MaxPool in the epilogue; 2 = last block, time sums only) -- the
// epilogues of conv_tc_kernel.
#pragma once

constexpr int C2P_MAXST = 16;       // weight ring depth (max): half slabs are small, the ring has to cover the L2 latency

template <int RC, bool FIRST>
__device__ __forceinline__ void conv_pair_issue_stage(uint32_t acc0, uint64_t ad_k, uint64_t bd_w, uint32_t co, uint32_t xal16,
                                                      uint32_t bstep, uint32_t idesc, uint64_t* commit_bar) {
    constexpr int NJ = 4, CNT = NJ * RC;
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
    constexpr int MASK = FIRST ? 0xE : 0xF;
#pragma unroll
    for (int c = 0; c < CNT; c += 4) tc::mma_pair_bf16_x4<MASK>(dd + c, ad + c, bd + c, idesc);
    tc::mma_pair_commit(commit_bar);
}

template <int RC>
__device__ __forceinline__ void conv_pair_issue_group(uint32_t acc0, uint64_t ad_g, uint64_t bd0, uint32_t co, uint32_t xal16,
                                                      uint32_t bstep, uint32_t stage16, uint32_t gstep, int groups,
                                                      uint32_t idesc, uint64_t* wfull, uint64_t* wempty, int nst, int& slot,
                                                      uint32_t& wphase) {
    bool first = true;
    for (int k = 0; k < ECG_KS; ++k) {
        uint64_t ad_k = ad_g + (uint64_t)k;                    // tap k = the same tile, k rows (16 B each) further
        for (int g = 0; g < groups; ++g, ad_k += gstep) {
            tc::mbar_wait(wfull + slot, wphase);
            tc::fence_after_sync();
            const uint64_t bd_w = bd0 + (uint64_t)((uint32_t)slot * stage16);
            if (first) conv_pair_issue_stage<RC, true>(acc0, ad_k, bd_w, co, xal16, bstep, idesc, wempty + slot);
            else conv_pair_issue_stage<RC, false>(acc0, ad_k, bd_w, co, xal16, bstep, idesc, wempty + slot);
            first = false;
            if (++slot == nst) { slot = 0; wphase ^= 1; }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(C2_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap xmapA, const __grid_constant__ CUtensorMap xmapB,
                    const __grid_constant__ CUtensorMap wmap, const float* __restrict__ bias,
                    const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, float* __restrict__ stat_part,
                    const Conv2Cfg P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* wfull = reinterpret_cast<uint64_t*>(smem);           // [C2P_MAXST]
    uint64_t* wempty = wfull + C2P_MAXST;                           // [C2P_MAXST]
    uint64_t* xfull = wempty + C2P_MAXST;                           // [2]
    uint64_t* accfull = xfull + 2;                                  // [2]
    uint64_t* accempty = accfull + 2;                               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
    float* statsh = reinterpret_cast<float*>(smem + TC_HDR);        // [4][2][256]
    uint8_t* xs = smem + TC_HDR + C2_STATB;
    uint8_t* wsm = xs + (size_t)P.NXB * P.R * P.xbytes_al;

    const int Ci = P.Ci, Co = P.Co, L = P.L, R = P.R;
    constexpr int kch = 64;
    const uint32_t xbytes = (uint32_t)Ci * TC_ROWS * 2;
    const uint32_t half_bytes = (uint32_t)kch * (uint32_t)(Co / 2) * 2;      // this CTA's share of one weight stage
    const int groups = Ci / kch;
    const int nstage = ECG_KS * groups;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();                    // 0 = the CTA whose thread issues the MMAs
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int ngl = pair < P.ngroups ? (P.ngroups - 1 - pair) / npairs + 1 : 0;
    long long* const trace = g_conv_trace;
    if (threadIdx.x == 0) { CTR(0); CTA_SPAN(0); }

    if (threadIdx.x == 0) {
        for (int i = 0; i < C2P_MAXST; ++i) { tc::mbar_init(wfull + i, 1); tc::mbar_init(wempty + i, 1); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(xfull + i, 1);
            tc::mbar_init(accfull + i, 1); tc::mbar_init(accempty + i, 16);
        }
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&xmapA);
        tc::prefetch_tmap(&xmapB);
        tc::prefetch_tmap(&wmap);
    }
    if (warp == 2) tc::tmem_alloc_pair(tmem_slot, P.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();                                         // both CTAs' barriers initialised, both TMEM slices allocated
    tc::fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (threadIdx.x == 0) CTR(1);

    // tiles of group gi: [base, base + 2R) -- the first R belong to rank 0
    auto group_base = [&](int gi) { return (pair + gi * npairs) * 2 * R; };
    auto count_of = [&](int base, int r) { const int n = P.total_tiles - base - r * R; return n < 0 ? 0 : (n > R ? R : n); };

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t wfull_l = tc::mapa_u32(tc::smem_u32(wfull), 0);
            // weights only: the input tiles are issued by the epilogue warps (below)
            int slot = 0;
            uint32_t ephase = 1;                                 // first pass over the ring: slots start free
            for (int gi = 0; gi < ngl; ++gi) {
                for (int s = 0; s < nstage; ++s) {
                    if (!(gi == 0 && s < P.NST)) tc::mbar_wait(wempty + slot, ephase);
                    if (rank == 0) tc::mbar_arrive_expect_tx(wfull + slot, 2u * half_bytes);
                    const int k = s / groups, g = s - k * groups;
                    // this CTA's Co/2 columns of the 64-channel slab of tap k, channel group g: 8 rows of (Co/2) * 16 bytes
                    tc::tma_load_3d_pair(wsm + (size_t)slot * half_bytes, &wmap, wfull_l + 8u * (uint32_t)slot,
                                         (int)rank * Co, g * (kch / 8), k);
                    if (++slot == P.NST) { slot = 0; ephase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // single-thread issuer, rank 0 only (see conv_tc_kernel for why everything stays inside one divergent region)
        const bool leader = tc::elect_one();
        if (leader && rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(2 * TC_TILE_M, Co, 0, 0);
            const uint64_t adesc0 = tc::make_desc(0, TC_ROWS * 16, 128);
            const uint64_t bdesc0 = tc::make_desc(0, (uint32_t)(Co / 2) * 16, 128);
            const uint64_t alo0 = adesc0 + (uint64_t)(tc::smem_u32(xs) >> 4);
            const uint64_t blo0 = bdesc0 + (uint64_t)(tc::smem_u32(wsm) >> 4);
            const uint32_t xal16 = P.xbytes_al >> 4, stage16 = half_bytes >> 4;
            const uint32_t bstep = (uint32_t)Co;                   // two 8-channel chunks of the half slab: 2 * (Co/2) rows
            const uint32_t gstep = (uint32_t)(kch / 8) * TC_ROWS;
            CTR(3);
            int slot = 0;
            uint32_t wphase = 0;
            for (int gi = 0; gi < ngl; ++gi) {
                const int xb = gi % P.NXB, as = gi % P.AS;
                const int rcount = count_of(group_base(gi), 0);    // rank 0 never has fewer tiles than rank 1
                tc::mbar_wait(xfull + xb, (gi / P.NXB) & 1);
                if (gi < 4) CTR(16 + gi);
                if (gi >= P.AS) tc::mbar_wait(accempty + as, ((gi / P.AS) - 1) & 1);
                if (gi < 4) CTR(24 + gi);
                tc::fence_after_sync();
                const uint32_t acc0 = tmem_base + (uint32_t)(as * R * Co);
                const uint64_t alo_g = alo0 + (uint64_t)((uint32_t)(xb * R) * xal16);
#define ECG_PSTR(RC_) conv_pair_issue_group<RC_>(acc0, alo_g, blo0, (uint32_t)Co, xal16, bstep, stage16, gstep, groups, idesc, wfull, wempty, P.NST, slot, wphase)
                if (rcount == 4) ECG_PSTR(4); else if (rcount == 3) ECG_PSTR(3); else if (rcount == 2) ECG_PSTR(2); else ECG_PSTR(1);
#undef ECG_PSTR
                tc::mma_pair_commit(accfull + as);
                if (gi < 4) CTR(32 + gi);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // which of the two warps of this quarter
        const int nblk = Co >> 5;
        const int row = 32 * q + lane;
        const size_t chunk_stride = (size_t)L * 8;
        const bool want_stats = MODE == 0 && stat_part != nullptr;
        const uint32_t accempty_l = tc::mapa_u32(tc::smem_u32(accempty), 0);
        // Input tiles.  A tile is 2 x Ci/8 TMA instructions (per 8-channel chunk a 2 KB box for rows [t0-7, t0+121) and a 256 B
        // box for rows [t0+121, t0+137)), and ONE thread gets a TMA instruction out only every ~50 cycles (its operands travel
        // through R2UR): 3-6 k cycles per group from the producer thread, in front of the weight stages queued behind them.
        // So lane 0 of each of the eight epilogue warps takes every eighth chunk: groups 0 and 1 at the start of the kernel
        // (both buffers are free), group g+2 at the start of the epilogue of group g -- its accumulators being complete means
        // that the MMAs have finished reading buffer g % 2 -- so that the tiles land under the MMAs of group g+1.
        const uint32_t xfull_l = tc::mapa_u32(tc::smem_u32(xfull), 0);
        auto load_tiles = [&](int gi) {
            const int xb = gi % P.NXB;
            const int base = group_base(gi);
            const int rc0 = count_of(base, 0), rc1 = count_of(base, 1);
            if (rank == 0 && warp == 2) tc::mbar_arrive_expect_tx(xfull + xb, xbytes * (uint32_t)(rc0 + rc1));
            const int tile0 = base + (int)rank * R, rcount = rank ? rc1 : rc0;
            const uint32_t bar = xfull_l + 8u * (uint32_t)xb;
            for (int r = 0; r < rcount; ++r) {
                const int tile = tile0 + r;
                const int b = tile / P.tiles_t, t2 = 2 * ((tile - b * P.tiles_t) * TC_TILE_M - ECG_PAD);
                uint8_t* dst = xs + (size_t)(xb * R + r) * P.xbytes_al + (size_t)(warp - 2) * (TC_ROWS * 16);
                for (int c = warp - 2; c < Ci / 8; c += 8, dst += 8 * TC_ROWS * 16) {
                    tc::tma_load_3d_pair(dst, &xmapA, bar, t2, c, b);
                    tc::tma_load_3d_pair(dst + 128 * 16, &xmapB, bar, t2 + 256, c, b);
                }
            }
        };
        if (lane == 0) {
            if (ngl > 0) load_tiles(0);
            if (ngl > 1) load_tiles(1);                  // more than one group per pair: two input buffers (conv_pair_cfg)
        }
        __syncwarp();
        float ssum[8], ssq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
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
        for (int gi = 0; gi < ngl; ++gi) {
            const int as = gi % P.AS;
            const int base = group_base(gi);
            const int tile0 = base + (int)rank * R, rcount = count_of(base, (int)rank);
            tc::mbar_wait(accfull + as, (gi / P.AS) & 1);
            if (gi < 4 && threadIdx.x == 64) CTR(40 + gi);
            tc::fence_after_sync();
            if (lane == 0 && gi + 2 < ngl) load_tiles(gi + 2);
            __syncwarp();
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                if (cb < nblk && (nblk == 1 || (cb & 1) == half)) {
                    const int c0 = cb * 32;
                    if (!small) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) { vs[i] = 0.f; qs[i] = 0.f; }
                    }
                    float bv[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) bv[i] = bias != nullptr ? __ldg(bias + c0 + i) : 0.f;
                    float sh[(MODE == 1 || MODE == 2) ? 32 : 1];
                    if constexpr (MODE == 1 || MODE == 2) {
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
                        if constexpr (MODE == 1 || MODE == 2) {
                            // inference (see conv_tc_kernel): relu(scale * conv + shift) (`bias` = scale), max over the pool pair
                            // (lanes 2p, 2p+1): the even lane keeps channels 0-15 of the block, the odd lane 16-31
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
                            if constexpr (MODE == 1) {
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
                                // sum over the 16 pool pairs of this warp (transposing butterfly over lane bits 4..1):
                                // lane l ends with channel c0 + 16*(l&1) + (l>>1)
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
            if (lane == 0) tc::mbar_arrive_cluster(accempty_l + 8u * (uint32_t)as);
            if (gi < 4 && threadIdx.x == 64) CTR(48 + gi);
        }
        if (want_stats && small) {
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
    // neither CTA may leave (or free its TMEM) while the other one's loads still count on its barriers or the leader's MMAs
    // still write its accumulators
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    if (threadIdx.x == 0) CTR(2);
    if (warp == 2) {
        tc::tmem_dealloc_pair(tmem_base, P.tmem_cols);
        if (lane == 0) CTA_SPAN(1);
    }
}

// Shape -> schedule of the pair kernel.  Returns the grid size (2 x pairs; = number of stat partials), or 0 when the layer is
// not one the pair kernel takes (weights that fit in shared memory stay with the one-SM kernel: nothing to halve there).
static int g_conv_pair = 3;                        // A/B switch (bit 0: forward / dgrad, bit 1: wgrad, bit 2: even where it does not pay): ecgb200_debug_set_conv_pair
static int conv_pair_cfg(int B, int Ci, int Co, int L, Conv2Cfg* P, size_t* smem_out) {
    if (!(g_conv_pair & 1)) return 0;
    if (Ci < 64 || (Ci & 63) || Ci > 384 || Co < 64 || (Co & 63) || Co > 256) return 0;
    if ((size_t)ECG_KS * Ci * Co * 2 <= 64 * 1024) return 0;
    const int npmax = ecg_num_sms() / 2;
    P->Ci = Ci; P->Co = Co; P->L = L; P->kch = 64;
    P->tiles_t = ecg_cdiv(L, TC_TILE_M);
    P->total_tiles = B * P->tiles_t;
    P->xbytes_al = ((uint32_t)Ci * TC_ROWS * 2 + 1023u) & ~1023u;
    const size_t half = (size_t)64 * (Co / 2) * 2;
    const size_t budget = 225 * 1024 - TC_HDR - C2_STATB;
    // R tiles per CTA per group (two accumulator stages in the 512 TMEM columns).  Cost of a choice = rounds on the busiest
    // pair x (R + 1/2): every round streams the whole weight tensor again (at R = 1 a 128-channel layer pulls 32 B/clk per SM
    // out of L2 -- three quarters of what L2 delivers to 148 SMs) and pays its ramp; on a tie the larger R.
    int R = 0, best = 1 << 30, tiles_busiest = 0;
    for (int r = 256 / Co; r >= 1; r >>= 1) {
        const int npg = ecg_cdiv(P->total_tiles, 2 * r);
        const int rounds = ecg_cdiv(npg, npmax);
        if ((size_t)(rounds > 1 ? 2 : 1) * r * P->xbytes_al + 4 * half > budget) continue;
        const int cost = rounds * (2 * r + 1);
        if (cost < best) { best = cost; R = r; tiles_busiest = rounds * r; }
    }
    if (R < 1) return 0;
    // one tile per CTA: nothing for the second accumulator stage to overlap, and the pair's extra set-up / tear-down
    // (cluster barriers, multicast commits: ~1.5 us per kernel) costs more than the lighter main loop saves
    // (measured at batch 64: step 0.253 -> 0.267 ms)
    if (tiles_busiest < 2 && !(g_conv_pair & 4)) return 0;
    P->R = R; P->AS = 2;
    P->ngroups = ecg_cdiv(P->total_tiles, 2 * R);                     // pair groups
    const int npairs = P->ngroups < npmax ? P->ngroups : npmax;
    P->NXB = ecg_cdiv(P->ngroups, npairs) > 1 ? 2 : 1;
    const size_t xall = (size_t)P->NXB * R * P->xbytes_al;
    int nst = (int)((budget - xall) / half);
    if (nst > C2P_MAXST) nst = C2P_MAXST;
    if (nst < 2) return 0;
    // The ring has to cover the latency of a weight load under load (~2 us): slots x MMA time per slot.  The 256 -> 128
    // channel dgrad with two 72 KB input buffers keeps 9 slots of 256 cycles and starves (batch 1024: 114 us against 97 us
    // for the one-SM kernel); every other pair configuration of this network holds >= 4.6 k cycles.
    const int mma_cyc = Co <= 64 ? 47 : (Co <= 128 ? 64 : 128);
    if (nst * R * 4 * mma_cyc < 4000 && !(g_conv_pair & 4)) return 0;
    P->NST = nst;
    P->wide = 1; P->Cn = Co;
    P->tmem_cols = tmem_cols_for(2 * R * Co);
    *smem_out = TC_HDR + C2_STATB + xall + (size_t)nst * half;
    return 2 * npairs;
}

template <int MODE>
static int conv_tc_pair_launch(const void* xb, const void* wprep, const float* bias, const float* shift, void* yb,
                               float* stat_part, int B, int Ci, int Co, int L, const Conv2Cfg& P, int grid, size_t smem,
                               void* stream) {
    CUtensorMap xmapA, xmapB, wmap;
    int rc = ecg_make_act_tmap64(&xmapA, xb, B, Ci, L, 128, 1);
    if (rc) return rc;
    rc = ecg_make_act_tmap64(&xmapB, xb, B, Ci, L, TC_ROWS - 128, 1);
    if (rc) return rc;
    // the weight tensor [15][Ci/8][Co][8] has the shape of an activation tensor with 15 "windows" of Co "time steps"
    rc = ecg_make_act_tmap64(&wmap, wprep, ECG_KS, Ci, Co, Co / 2, 8);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(conv_tc_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel<MODE>, xmapA, xmapB, wmap, bias, shift, (__nv_bfloat16*)yb, stat_part, P);
    return e == cudaSuccess ? ecg_launch_status() : (int)e;
}
