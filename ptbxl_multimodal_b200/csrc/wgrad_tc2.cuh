// Two-SM form (tcgen05 cta_group::2) of the weight-gradient kernel for the wide layers (blocks 3 and 4): included by
// conv1d_tc.cu.  "Taps as M", as in wgrad_thin_kernel, on a CTA pair:
//   D[(k, c8)][o] += X[t + k][c8]^T * dY[t][o]          per 8-channel chunk of the input
//   A = one chunk of the X tile, MN-major, M-chunk stride ONE ROW (M-chunk m = the chunk shifted by m rows = tap m):
//       M = 16 taps x 8 channels per CTA, M = 256 per instruction -- CTA r of the pair contributes ITS OWN chunk;
//   B = the dY tile [Co/8][128 rows][8], MN-major, N = Co (128 / 256), split by output channel: each CTA stages only its half.
// Per MMA each SM reads 4 KB of A and N * 16 B of B from shared memory instead of 4 KB + N * 32 B (the one-SM taps-as-N
// kernel: 8 KB per 64-cycle instruction = the whole shared-memory port, 77-92 cycles per MMA measured), and the dY tile is
// fetched from L2 once per PAIR and item instead of once per CTA.  TMEM: chunk i of a CTA in columns [i * Co, (i + 1) * Co),
// lane = (tap, c8): 512 / Co chunks per CTA, twice that per pair; the pairs of one item tile the input channels
// (blockIdx -> (channel group, split)), the batch is split S ways.  Partials: part[z][o][c/8][16][8], the layout the other
// weight-gradient kernels write, so wgrad_tc_reduce_kernel finishes the job unchanged.
#pragma once

template <int NCC>
__device__ __forceinline__ void wgrad_pair_issue_item(uint32_t tmem_base, uint32_t co, uint64_t alo, uint64_t blo,
                                                      uint32_t idesc, bool accum) {
#pragma unroll
    for (int i = 0; i < NCC; ++i) {
#pragma unroll
        for (int jb = 0; jb < TC_TILE_M / 16; jb += 4) {
            uint32_t dd[4];
            uint64_t al[4], bl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dd[e] = tmem_base + (uint32_t)i * co;
                al[e] = alo + (uint64_t)(i * TC_ROWS + (jb + e) * 16);       // X chunk i, K-step = 16 time rows
                bl[e] = blo + (uint64_t)((jb + e) * 16);                      // dY half tile, same K-step
            }
            if (jb > 0 || accum) tc::mma_pair_bf16_x4<0xF>(dd, al, bl, idesc);
            else tc::mma_pair_bf16_x4<0xE>(dd, al, bl, idesc);
        }
    }
}

__global__ void __launch_bounds__(192, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                  float* __restrict__ part, int Co, int Cip, int L, int B, int ncc, int ncb, int nst) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + WT_MAXST;
    uint64_t* accfull = empty + WT_MAXST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
    uint8_t* stages = smem + TC_HDR;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int pair = (int)(blockIdx.x >> 1);
    const int cbp = pair % ncb, z = pair / ncb, S = (int)(gridDim.x >> 1) / ncb;
    const int chunk0 = (cbp * 2 + (int)rank) * ncc;                   // this CTA's first 8-channel chunk of the input
    long long* const trace = blockIdx.x == 0 ? g_conv_trace : nullptr;
    if (threadIdx.x == 0) { CTR(0); CTA_SPAN(0); }
    const int tiles_t = (L + TC_TILE_M - 1) / TC_TILE_M;
    const int items = B * tiles_t;
    const int nloc = (items - z + S - 1) / S;            // items z, z+S, ...   (host guarantees >= 1)
    const uint32_t dybytes = (uint32_t)(Co / 2) * 128 * 2;            // this CTA's half: [Co/16][128 rows][8] bf16
    const uint32_t xbytes = (uint32_t)ncc * TC_ROWS * 16;
    const uint32_t stage_bytes = dybytes + xbytes;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WT_MAXST; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(accfull, 1);
        tc::fence_barrier_init();
        tc::fence_proxy_async();
        tc::prefetch_tmap(&dymap);
        tc::prefetch_tmap(&xmap);
    }
    if (warp == 2) tc::tmem_alloc_pair(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t full_l = tc::mapa_u32(tc::smem_u32(full), 0);
            int slot = 0;
            uint32_t ephase = 1;                                 // fresh barriers: the first pass does not block
            int b = z / tiles_t, tt = z - b * tiles_t;           // item -> (sample, time tile), advanced by S per step
            const int db = S / tiles_t, dt = S - db * tiles_t;
            for (int n = 0; n < nloc; ++n) {
                tc::mbar_wait(empty + slot, ephase);
                uint8_t* st = stages + (size_t)slot * stage_bytes;
                if (rank == 0) tc::mbar_arrive_expect_tx(full + slot, 2u * stage_bytes);
                const uint32_t bar = full_l + 8u * (uint32_t)slot;
                // this CTA's Co/2 output channels of the dY tile: one box of 128 rows x Co/16 chunks
                tc::tma_load_3d_pair(st, &dymap, bar, 2 * tt * TC_TILE_M, (int)rank * (Co / 16), b);
                // this CTA's ncc chunks of the X tile: {8, 144, ncc, 1}
                tc::tma_load_4d_pair(st + dybytes, &xmap, bar, 0, tt * TC_TILE_M - ECG_PAD, chunk0, b);
                if (n < 8) CTR(8 + n);
                if (++slot == nst) { slot = 0; ephase ^= 1; }
                b += db; tt += dt;
                if (tt >= tiles_t) { tt -= tiles_t; ++b; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(2 * TC_TILE_M, Co, 1, 1);
            // A = X chunk: M chunk m = tap m = the chunk shifted by m rows (SBO = 16 B), K 8-row groups 128 B apart
            const uint64_t adesc0 = tc::make_desc(0, 128, 16);
            // B = dY half: N (o) chunks 128 rows * 16 B apart, K (t) 8-row groups 128 B apart
            const uint64_t bdesc0 = tc::make_desc(0, 128, 128 * 16);
            const uint64_t alo0 = adesc0 + (uint64_t)((tc::smem_u32(stages) + dybytes) >> 4);
            const uint64_t blo0 = bdesc0 + (uint64_t)(tc::smem_u32(stages) >> 4);
            int slot = 0;
            uint32_t fphase = 0, accum = 0;
            for (int n = 0; n < nloc; ++n) {
                tc::mbar_wait(full + slot, fphase);
                if (n < 8) CTR(16 + n);
                tc::fence_after_sync();
                const uint64_t alo = alo0 + (uint64_t)((uint32_t)slot * (stage_bytes >> 4));
                const uint64_t blo = blo0 + (uint64_t)((uint32_t)slot * (stage_bytes >> 4));
                if (ncc == 4) wgrad_pair_issue_item<4>(tmem_base, (uint32_t)Co, alo, blo, idesc, accum != 0);
                else wgrad_pair_issue_item<2>(tmem_base, (uint32_t)Co, alo, blo, idesc, accum != 0);
                accum = 1;
                tc::mma_pair_commit(empty + slot);
                if (n < 8) CTR(24 + n);
                if (++slot == nst) { slot = 0; fphase ^= 1; }
            }
            tc::mma_pair_commit(accfull);
            CTR(32);
        }
    } else {
        const int q = warp & 3;
        const int nct = Cip / 8;                                      // chunks of the whole input
        tc::mbar_wait(accfull, 0);
        if (threadIdx.x == 64) CTR(33);
        tc::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16);
        for (int i = 0; i < ncc; ++i) {
            for (int c0 = 0; c0 < Co; c0 += 32) {
                float v[32];
                tc::tmem_ld32(taddr + (uint32_t)(i * Co + c0), v);
                tc::tmem_ld_wait();
                // lane (tap, c8) of output channel o = 512 contiguous bytes over the 128 lanes: coalesced as is
                float* dst = part + (((size_t)z * Co + c0) * nct + chunk0 + i) * 128 + 32 * q + lane;
#pragma unroll
                for (int e = 0; e < 32; ++e) dst[(size_t)e * nct * 128] = v[e];
            }
        }
        if (threadIdx.x == 64) CTR(34);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();
    if (warp == 2) {
        tc::tmem_dealloc_pair(tmem_base, 512);
        if (lane == 0) CTA_SPAN(1);
    }
}

// Shape -> {chunks per CTA, channel groups of 2 * ncc chunks, splits}; false when the layer is not one this kernel takes
static bool wgrad_pair_cfg(int B, int Cip, int Co, int L, int* ncc, int* ncb, int* S) {
    if (!(g_conv_pair & 2)) return false;
    if (Co != 128 && Co != 256) return false;
    const int n = 512 / Co;                                           // 4 or 2 chunks per CTA
    if (Cip < 16 * n || (Cip / 8) % (2 * n)) return false;
    const int cb = Cip / 8 / (2 * n);
    const int npairs = ecg_num_sms() / 2;
    if (cb > npairs) return false;
    const int items = B * ecg_cdiv(L, TC_TILE_M);
    int s = npairs / cb;
    // the lighter main loop saves ~0.3 us per item of a pair's share; set-up and tear-down of a pair cost ~1.5 us more
    if (items < 6 * s && !(g_conv_pair & 4)) return false;
    if (s > items) s = items;
    *ncc = n; *ncb = cb; *S = s;
    return true;
}
