// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, M=128) from SWIZZLE_NONE shared-memory
// operands, as used by the conv kernels.  One CTA per SM, one issuing thread, clock64 around NMMA issues
// + commit + wait.
#include "../ptbxl_multimodal_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(320, 1)
mma_bench(int N, int nmma, int shift16, int nacc, int lbo, int sbo, int amn, int bmn, int a_stride, int b_stride,
          long long* out, int mode, int M, int blbo, int bsbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
    uint8_t* A = smem + 1024;
    uint8_t* Bm = smem + 1024 + 96 * 1024;
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_init(bar + 1, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(slot, 512);
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0x3c003c00u;
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *slot;
    if (mode == 2 && threadIdx.x < 32) {
        // warp-convergent issue: the whole warp runs the loop, one elected lane issues (CUTLASS style)
        const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
        const uint32_t a0 = tc::smem_u32(A), b0 = tc::smem_u32(Bm);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint64_t ad = tc::make_desc(a0 + (uint32_t)((i % 15) * shift16 * 16) + (uint32_t)((i & 3) * a_stride), lbo, sbo);
            const uint64_t bd = tc::make_desc(b0 + (uint32_t)((i & 3) * b_stride), (uint32_t)(bmn ? 128 : N * 16), bmn ? 16 : 128);
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
            if (pred) tc::mma_bf16(tmem + (uint32_t)((i % nacc) * N), ad, bd, idesc, i >= nacc ? 1u : 0u);
            __syncwarp();
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) {
            tc::mma_commit(bar);
            tc::mbar_wait(bar, 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    } else if (mode == 5) {
        const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // provably warp-uniform
        if (warp_u == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
            const uint64_t ad0 = tc::make_desc(tc::smem_u32(A), lbo, sbo);
            const uint64_t bd0 = tc::make_desc(tc::smem_u32(Bm), (uint32_t)(N * 16), 128);
            const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
            const uint32_t alo0 = (uint32_t)ad0, blo0 = (uint32_t)bd0;
            uint32_t leader;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
            const long long t0 = clock64();
            uint32_t accum = 0;
            int cnt = 0;
            while (cnt < nmma) {
                uint32_t wlo = blo0;
                for (int k = 0; k < 15; ++k) {
                    uint32_t alo_r = alo0 + (uint32_t)(k * shift16);
                    for (int r = 0; r < nacc; ++r, alo_r += (uint32_t)(a_stride >> 4)) {
                        const uint64_t ad = ((uint64_t)ahi << 32) | alo_r;
                        const uint64_t bd = ((uint64_t)bhi << 32) | wlo;
                        if (leader) tc::mma_bf16(tmem + (uint32_t)(r * N), ad, bd, idesc, accum);
                        ++cnt;
                    }
                    accum = 1;
                    wlo += (uint32_t)(b_stride >> 4);
                }
            }
            const long long t1 = clock64();
            if (leader) {
                tc::mma_commit(bar);
                tc::mbar_wait(bar, 0);
                const long long t2 = clock64();
                if (blockIdx.x == 0) { out[0] = (t1 - t0) * nmma / cnt; out[1] = (t2 - t0) * nmma / cnt; }
            }
        }
    } else if (mode >= 3 && mode <= 4 && threadIdx.x >= 64) {
        if (mode == 4) tc::mbar_wait(bar + 1, 0);               // spinning waiters like the epilogue warps
    } else if (mode >= 3 && mode <= 4 && threadIdx.x == 0) {
        // lean loop as in conv_tc_kernel: lo-word increments, tap shift, R accumulators, B walks stages
        const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
        const uint64_t ad0 = tc::make_desc(tc::smem_u32(A), lbo, sbo);
        const uint64_t bd0 = tc::make_desc(tc::smem_u32(Bm), (uint32_t)(N * 16), 128);
        const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
        const uint32_t alo0 = (uint32_t)ad0, blo0 = (uint32_t)bd0;
        const long long t0 = clock64();
        uint32_t accum = 0;
        int cnt = 0;
        while (cnt < nmma) {
            uint32_t wlo = blo0;
            for (int k = 0; k < 15; ++k) {
                uint32_t alo_r = alo0 + (uint32_t)(k * shift16);
                for (int r = 0; r < nacc; ++r, alo_r += (uint32_t)(a_stride >> 4)) {
                    const uint64_t ad = ((uint64_t)ahi << 32) | alo_r;
                    const uint64_t bd = ((uint64_t)bhi << 32) | wlo;
                    tc::mma_bf16(tmem + (uint32_t)(r * N), ad, bd, idesc, accum);
                    ++cnt;
                }
                accum = 1;
                wlo += (uint32_t)(b_stride >> 4);
            }
        }
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = (t1 - t0) * nmma / cnt; out[1] = (t2 - t0) * nmma / cnt; }
        tc::mbar_arrive(bar + 1);
    } else if (mode >= 10 && mode <= 13) {
        // issue-path experiments, whole warp 0 executes uniformly, elected lane issues
        const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
        if (warp_u == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
            const uint64_t ad0 = tc::make_desc(tc::smem_u32(A), lbo, sbo);
            const uint64_t bd0 = tc::make_desc(tc::smem_u32(Bm), (uint32_t)(N * 16), 128);
            const uint32_t ahi = (uint32_t)(ad0 >> 32), bhi = (uint32_t)(bd0 >> 32);
            uint32_t alo = (uint32_t)ad0, blo = (uint32_t)bd0;
            const uint32_t astep = (uint32_t)shift16, bstep = (uint32_t)(b_stride >> 4);
            uint32_t leader;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
            const long long t0 = clock64();
            if (mode == 10) {                      // one uniform add per MMA
#pragma unroll 8
                for (int i = 0; i < nmma; ++i) {
                    if (leader) tc::mma_bf16(tmem, ((uint64_t)ahi << 32) | alo, ((uint64_t)bhi << 32) | blo, idesc, 1u);
                    alo += astep; blo += bstep;
                    if ((i & 15) == 15) { alo -= 16 * astep; blo -= 16 * bstep; }
                }
            } else {                               // 11: batches of 4; 12: batches of 8; 13: batch 4, per-thread (vector) values
                const uint32_t vz = mode == 13 ? (threadIdx.x >> 6) : 0u;       // 0, but not provably uniform
                for (int i = 0; i < nmma; i += (mode == 12 ? 8 : 4)) {
                    uint32_t dd[4], al[4], bl[4], fl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { dd[e] = tmem + vz; al[e] = alo + (uint32_t)e * astep + vz; bl[e] = blo + (uint32_t)e * bstep; fl[e] = 1u; }
                    uint64_t a64[4], b64[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { a64[e] = ((uint64_t)ahi << 32) | al[e]; b64[e] = ((uint64_t)bhi << 32) | bl[e]; }
                    (void)fl;
                    if (leader) tc::mma_bf16_x4<0xF>(dd, a64, b64, idesc);
                    if (mode == 12) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) { a64[e] += 4 * astep; b64[e] += 4 * bstep; }
                        if (leader) tc::mma_bf16_x4<0xF>(dd, a64, b64, idesc);
                    }
                    alo += 4 * astep; blo += 4 * bstep;
                    if ((i & 15) == 12) { alo -= 16 * astep; blo -= 16 * bstep; }
                }
            }
            const long long t1 = clock64();
            if (leader) {
                tc::mma_commit(bar);
                tc::mbar_wait(bar, 0);
                const long long t2 = clock64();
                if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
            }
        }
    } else if (mode >= 6 && mode <= 9 && threadIdx.x == 0) {
        // which operand change costs?  6: alternate A, 7: alternate B, 8: alternate D, 9: alternate A+B
        const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
        const uint32_t a0 = tc::smem_u32(A), b0 = tc::smem_u32(Bm);
        const uint64_t adA = tc::make_desc(a0, lbo, sbo);
        const uint64_t adB = tc::make_desc(a0 + ((mode == 6 || mode == 9) ? a_stride : 0), lbo, sbo);
        const uint64_t bdA = tc::make_desc(b0, (uint32_t)(N * 16), 128);
        const uint64_t bdB = tc::make_desc(b0 + ((mode == 7 || mode == 9) ? b_stride : 0), (uint32_t)(N * 16), 128);
        const uint32_t dA = tmem, dB = tmem + (mode == 8 ? (uint32_t)N : 0u);
        const long long t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < nmma; i += 2) {
            tc::mma_bf16(dA, adA, bdA, idesc, 1u);
            tc::mma_bf16(dB, adB, bdB, idesc, 1u);
        }
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (mode == 1 && threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
        const uint32_t a0 = tc::smem_u32(A), b0 = tc::smem_u32(Bm);
        const uint64_t ad = tc::make_desc(a0, lbo, sbo);
        const uint64_t bd = tc::make_desc(b0, (uint32_t)(bmn ? 128 : N * 16), bmn ? 16 : 128);
        const long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < nmma; ++i) tc::mma_bf16(tmem, ad, bd, idesc, 1u);
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (mode == 0 && threadIdx.x == 0) {
        const uint32_t idesc = tc::make_idesc_bf16(M, N, amn, bmn);
        const uint32_t a0 = tc::smem_u32(A), b0 = tc::smem_u32(Bm);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint64_t ad = tc::make_desc(a0 + (uint32_t)((i % 15) * shift16 * 16) + (uint32_t)((i & 3) * a_stride), lbo, sbo);
            const uint64_t bd = tc::make_desc(b0 + (uint32_t)((i & 3) * b_stride), (uint32_t)(blbo ? blbo : (bmn ? 128 : N * 16)), (uint32_t)(bsbo ? bsbo : (bmn ? 16 : 128)));
            tc::mma_bf16(tmem + (uint32_t)((i % nacc) * N), ad, bd, idesc, i >= nacc ? 1u : 0u);
        }
        const long long t1 = clock64();
        tc::mma_commit(bar);
        tc::mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    const int smem = 1024 + 200 * 1024;
    cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nmma = 240;
    struct Case { const char* name; int N, shift, nacc, lbo, sbo, amn, bmn, astr, bstr; int mode = 0, M = 128, blbo = 0, bsbo = 0; };
    const Case cases[] = {
        {"fwd K-major N=256 aligned 1acc", 256, 0, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 256 * 16},
        {"fwd K-major N=256 shifted 1acc", 256, 1, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 256 * 16},
        {"fwd K-major N=256 shifted 2acc", 256, 1, 2, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 256 * 16},
        {"fwd K-major N=128 aligned 1acc", 128, 0, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 128 * 16},
        {"fwd K-major N=128 shifted 1acc", 128, 1, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 128 * 16},
        {"fwd K-major N=128 shifted 2acc", 128, 1, 2, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 128 * 16},
        {"fwd K-major N=128 shifted 4acc", 128, 1, 4, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 128 * 16},
        {"fwd K-major N=64  aligned 1acc", 64, 0, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 64 * 16},
        {"fwd K-major N=64  shifted 4acc", 64, 1, 4, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 64 * 16},
        {"fwd K-major N=32  aligned 1acc", 32, 0, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 32 * 16},
        {"fwd K-major N=32  shifted 1acc", 32, 1, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 32 * 16},
        {"fwd K-major N=32  shifted 4acc", 32, 1, 4, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 32 * 16},
        {"fwd K-major N=32  shifted 8acc", 32, 1, 8, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 32 * 16},
        {"invariant desc N=256", 256, 0, 1, 144 * 16, 128, 0, 0, 0, 0, 1, 128},
        {"invariant desc N=128", 128, 0, 1, 144 * 16, 128, 0, 0, 0, 0, 1, 128},
        {"invariant desc N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 0, 0, 1, 128},
        {"invariant desc N=16", 16, 0, 1, 144 * 16, 128, 0, 0, 0, 0, 1, 128},
        {"invariant desc M=64 N=256", 256, 0, 1, 144 * 16, 128, 0, 0, 0, 0, 1, 64},
        {"invariant desc M=64 N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 0, 0, 1, 64},
        {"elect-style N=256 shifted", 256, 1, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 256 * 16, 2, 128},
        {"elect-style N=32 shifted", 32, 1, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 32 * 16, 2, 128},
        {"elect-style M=64 N=32 shifted", 32, 1, 1, 144 * 16, 128, 0, 0, 2 * 144 * 16, 2 * 32 * 16, 2, 64},
        {"lean N=32 noshift 2acc", 32, 0, 2, 144 * 16, 128, 0, 0, 5120, 1024, 3, 128},
        {"lean N=32 shift 2acc", 32, 1, 2, 144 * 16, 128, 0, 0, 5120, 1024, 3, 128},
        {"lean N=32 shift 1acc", 32, 1, 1, 144 * 16, 128, 0, 0, 5120, 1024, 3, 128},
        {"lean N=32 shift 2acc + spinners", 32, 1, 2, 144 * 16, 128, 0, 0, 5120, 1024, 4, 128},
        {"lean N=128 shift 2acc", 128, 1, 2, 144 * 16, 128, 0, 0, 73728, 4096, 3, 128},
        {"lean N=128 noshift 2acc", 128, 0, 2, 144 * 16, 128, 0, 0, 73728, 4096, 3, 128},
        {"uniform-warp lean N=32 shift 2acc", 32, 1, 2, 144 * 16, 128, 0, 0, 5120, 1024, 5, 128},
        {"uniform-warp lean N=32 shift 1acc", 32, 1, 1, 144 * 16, 128, 0, 0, 5120, 1024, 5, 128},
        {"uniform-warp lean N=64 shift 2acc", 64, 1, 2, 144 * 16, 128, 0, 0, 5120, 2048, 5, 128},
        {"uniform-warp lean N=128 shift 2acc", 128, 1, 2, 144 * 16, 128, 0, 0, 36864, 4096, 5, 128},
        {"uniform add per MMA N=32", 32, 1, 1, 144 * 16, 128, 0, 0, 0, 1024, 10, 128},
        {"uniform batch4 N=32", 32, 1, 1, 144 * 16, 128, 0, 0, 0, 1024, 11, 128},
        {"vector batch4 N=32", 32, 1, 1, 144 * 16, 128, 0, 0, 0, 1024, 13, 128},
        {"uniform batch4 N=128", 128, 1, 1, 144 * 16, 128, 0, 0, 0, 4096, 11, 128},
        {"alt A (tile stride 5120) N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 5120, 1024, 6, 128},
        {"alt A (+16 B) N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 16, 1024, 6, 128},
        {"alt A (+128 B) N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 128, 1024, 6, 128},
        {"alt B (+1024 B) N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 5120, 1024, 7, 128},
        {"alt D (+N cols) N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 5120, 1024, 8, 128},
        {"alt A+B N=32", 32, 0, 1, 144 * 16, 128, 0, 0, 5120, 1024, 9, 128},
        {"alt A+B N=128", 128, 0, 1, 144 * 16, 128, 0, 0, 5120, 4096, 9, 128},
        {"alt A+B N=256", 256, 0, 1, 144 * 16, 128, 0, 0, 5120, 8192, 9, 128},
        // round 2 (mode 1 = invariant descriptors: pure tensor-pipe rate of the operand geometry)
        {"wgrad both MN-major, B taps-as-N N=128", 128, 0, 1, 128, 128 * 16, 1, 1, 256, 256, 1, 128, 0, 0},
        {"wgrad A K-major / B taps-as-N N=128", 128, 0, 1, 128 * 16, 128, 0, 1, 256, 256, 1, 128, 0, 0},
        {"taps-as-M A(sbo16) / B MN-major N=32", 32, 0, 1, 128, 16, 1, 1, 256, 256, 1, 128, 128, 2048},
        {"taps-as-M A(sbo16) / B MN-major N=64", 64, 0, 1, 128, 16, 1, 1, 256, 256, 1, 128, 128, 2048},
        {"taps-as-M A(sbo16) / B MN-major N=128", 128, 0, 1, 128, 16, 1, 1, 256, 256, 1, 128, 128, 2048},
        {"taps-as-M A(sbo16) / B MN-major N=256", 256, 0, 1, 128, 16, 1, 1, 256, 256, 1, 128, 128, 2048},
        {"MN-major A(sbo 2K) / MN-major B(sbo 2K) N=128", 128, 0, 1, 128, 128 * 16, 1, 1, 256, 256, 1, 128, 128, 2048},
        {"MN-major A(sbo 2K) / MN-major B(sbo 2K) N=32", 32, 0, 1, 128, 128 * 16, 1, 1, 256, 256, 1, 128, 128, 2048},
        {"K-major A / MN-major B(sbo 2K) N=128", 128, 0, 1, 128 * 16, 128, 0, 1, 256, 256, 1, 128, 128, 2048},
        {"MN-major A(sbo 2K) / K-major B N=128", 128, 0, 1, 128, 128 * 16, 1, 0, 256, 256, 1, 128, 0, 0},
        {"taps-as-M A(sbo16) / K-major B N=32", 32, 0, 1, 128, 16, 1, 0, 256, 256, 1, 128, 0, 0},
        {"taps-as-M A(sbo16) / K-major B N=64", 64, 0, 1, 128, 16, 1, 0, 256, 256, 1, 128, 0, 0},
        {"taps-as-M A(sbo16) / K-major B N=128", 128, 0, 1, 128, 16, 1, 0, 256, 256, 1, 128, 0, 0},
    };
    for (const Case& c : cases) {
        for (int rep = 0; rep < 2; ++rep)
            mma_bench<<<148, c.mode >= 3 ? 320 : 128, smem>>>(c.N, nmma, c.shift, c.nacc, c.lbo, c.sbo, c.amn, c.bmn, c.astr, c.bstr, out, c.mode, c.M, c.blbo, c.bsbo);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%-40s issue %6.1f cyc/mma   complete %6.1f cyc/mma   (%s)\n", c.name, (double)h[0] / nmma,
               (double)h[1] / nmma, cudaGetErrorString(e));
    }
    return 0;
}
