// sm_100a primitives used by the tensor-core (tcgen05 / TMEM / TMA) conv kernels:
// mbarrier, bulk + tensor TMA loads, TMEM alloc, UMMA descriptors, tcgen05.mma/commit/ld.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must abort the kernel (trap) instead of hanging the GPU.  If a diagnostic
// buffer was registered (ecgb200_debug_set_diag: pinned host memory), the stuck barrier is recorded first.
static __device__ unsigned long long* g_mbar_diag = nullptr;   // per translation unit
// deadlock limit of mbar_wait in ns (0 = wait for ever); set by ecgb200_set_spin_timeout_ms.  The default is long
// enough for time-slicing / preemption / a debugger to pause the CTA without a spurious trap.
static __device__ unsigned long long g_mbar_timeout_ns = 30000000000ull;
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    uint64_t t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if ((it & 1023u) == 1023u) {
            const uint64_t now = globaltimer_ns();
            const uint64_t limit = g_mbar_timeout_ns;
            if (t0 == 0) t0 = now;
            else if (limit == 0) continue;
            else if (now - t0 > limit) {                          // deadlock
                unsigned long long* d = g_mbar_diag;
                if (d != nullptr) {
                    d[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
                    d[2] = ((unsigned long long)addr << 32) | parity;
                    d[0] = 0xDEADull;
                    __threadfence_system();
                }
                __trap();
            } else if (now - t0 > limit / 2) {                    // half way: every stuck waiter leaves a note
                unsigned long long* d = g_mbar_diag;
                if (d != nullptr && (threadIdx.x & 31) == (threadIdx.x >> 5) % 32 * 0 + ((threadIdx.x & 31))) {
                    const unsigned w = threadIdx.x >> 5;
                    d[4 + 2 * w] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
                    d[5 + 2 * w] = ((unsigned long long)addr << 32) | parity;
                    __threadfence_system();
                }
            }
        }
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled tensor load, completes on an mbarrier with complete_tx::bytes
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// 3-D tiled tensor load (the 8-byte-element view of an activation tensor, see ecg_make_act_tmap64)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
          "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t cols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {           // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a converged warp (elect.sync): the lane that issues TMA / tcgen05 instructions while the
// whole warp executes the surrounding loop, so that every index / descriptor value stays WARP-UNIFORM and
// lives in the uniform register file the tensor-core instructions read (no per-MMA R2UR transfers).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, 16-byte units:
//   K-major : ((8,m),(T,2)) : ((1T,SBO),(1,LBO))   LBO = between the two 8-element K halves,
//                                                  SBO = between 8-row groups along M/N
//   MN-major: ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)) SBO = between 8-element chunks along M/N,
//                                                  LBO = between 8-row groups along K
// (cute/atom/mma_traits_sm100.hpp make_umma_desc; version field = 1 on sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // version = 1 (Blackwell)
    return d;                                     // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
// Instruction descriptor for kind::f16 with BF16 inputs, FP32 accumulate (mma_sm100_desc.hpp)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Four MMAs issued back to back from ONE asm statement.  Measured on B200: a tcgen05.mma whose descriptor
// registers were written just before it stalls the issuing thread for ~100+ cycles (130-420 cycles per MMA
// in a compute-descriptor / issue / compute / issue loop, against 47 (N<=64), 64 (N=128), 128 (N=256) cycles
// when the descriptors are ready), so all operands of a batch are formed first and the tensor pipe keeps
// draining the batch while the next one is being prepared.
//   alo/blo: low descriptor words (address field already added), ahi/bhi: shared high words,
//   d: TMEM accumulator addresses.  No per-slot enable predicate: a predicated tcgen05.mma makes ptxas
//   re-materialise its uniform operands in front of every instruction.
//   ACCMASK bit e = accumulate flag of slot e, a COMPILE-TIME constant: a run-time predicate (written by a
//   uniform-datapath instruction just before the batch) is what made the issue ~100 cycles per MMA slower.
template <int ACCMASK>
__device__ __forceinline__ void mma_bf16_x4(const uint32_t* d, const uint64_t* ad, const uint64_t* bd, uint32_t idesc) {
    // 64-bit descriptor operands formed in C++ (not packed inside the asm): ptxas then keeps the whole
    // descriptor arithmetic in the uniform datapath (UIADD3.64) instead of per-thread registers + R2UR.
    asm volatile(
        "{\n\t"
        ".reg .pred pa0, pa1, pa2, pa3;\n\t"
        "setp.ne.b32 pa0, %13, 0;\n\t"
        "setp.ne.b32 pa1, %14, 0;\n\t"
        "setp.ne.b32 pa2, %15, 0;\n\t"
        "setp.ne.b32 pa3, %16, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %8, %12, pa0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %5, %9, %12, pa1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%2], %6, %10, %12, pa2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%3], %7, %11, %12, pa3;\n\t"
        "}"
        ::"r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]),
          "l"(ad[0]), "l"(ad[1]), "l"(ad[2]), "l"(ad[3]),
          "l"(bd[0]), "l"(bd[1]), "l"(bd[2]), "l"(bd[3]),
          "r"(idesc),
          "n"(ACCMASK & 1), "n"((ACCMASK >> 1) & 1), "n"((ACCMASK >> 2) & 1), "n"((ACCMASK >> 3) & 1)
        : "memory");
}
// Single MMA with a compile-time accumulate flag (see mma_bf16_x4).
template <int ACC>
__device__ __forceinline__ void mma_bf16_c(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "n"(ACC)
        : "memory");
}
// mbarrier arrives when all previously issued MMAs of this thread have completed
// Same, predicated INSIDE the asm on a per-thread flag.  Use this in warp-uniform code: a bare
// `if (leader) mma_commit(bar)` after the warp has reconverged was compiled by ptxas 12.9 into an unguarded
// warp-level UTCBAR with the operand broadcast from the leader, and the barrier over-arrived (sporadic
// deadlock); with the predicate as an asm operand the instruction stays per-thread.
__device__ __forceinline__ void mma_commit_if(uint64_t* bar, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t"
        "@p tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}
// ---------------------------------------------------------------- CTA pair (cta_group::2, cluster of two CTAs on one TPC)
// Rank 0 of the pair issues every tcgen05.mma for both SMs: M = 256 (each CTA's own 128 rows of A and its own TMEM lanes),
// B split by N (each CTA's shared memory holds N/2 columns at the SAME offset).  Barriers live at the same offsets in both
// CTAs; `mapa` gives the shared::cluster address of the leader's copy.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {          // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {     // possibly the peer CTA's barrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 3-D tiled load into THIS CTA's shared memory whose bytes are counted on a barrier given by its shared::cluster address
// (the pair leader's): the .cta_group::2 form is what allows destination and barrier to sit in different CTAs of the pair
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr),
          "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr),
          "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_in_smem, uint32_t cols) {   // the same warp of BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {           // the same warp of BOTH CTAs
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int ACCMASK>
__device__ __forceinline__ void mma_pair_bf16_x4(const uint32_t* d, const uint64_t* ad, const uint64_t* bd, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred pa0, pa1, pa2, pa3;\n\t"
        "setp.ne.b32 pa0, %13, 0;\n\t"
        "setp.ne.b32 pa1, %14, 0;\n\t"
        "setp.ne.b32 pa2, %15, 0;\n\t"
        "setp.ne.b32 pa3, %16, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %4, %8, %12, pa0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], %5, %9, %12, pa1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%2], %6, %10, %12, pa2;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%3], %7, %11, %12, pa3;\n\t"
        "}"
        ::"r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]),
          "l"(ad[0]), "l"(ad[1]), "l"(ad[2]), "l"(ad[3]),
          "l"(bd[0]), "l"(bd[1]), "l"(bd[2]), "l"(bd[3]),
          "r"(idesc),
          "n"(ACCMASK & 1), "n"((ACCMASK >> 1) & 1), "n"((ACCMASK >> 2) & 1), "n"((ACCMASK >> 3) & 1)
        : "memory");
}
// all MMAs issued so far by this thread complete -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma_pair_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(h);
}

}  // namespace tc

// ---------------------------------------------------------------- host: tensor-map encoding
// cuTensorMapEncodeTiled is fetched from the driver at first use (cudaGetDriverEntryPoint) so that
// the library loads on machines without libcuda (CPU-only build / symbol checks).
typedef CUresult (*ecg_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);
ecg_tmap_encode_fn ecg_get_tmap_encode();

// Activation tensor in the blocked channels-last layout [B][C/8][L][8] (bf16):
// 4-D map {8, L, C/8, B}, box {8, box_rows, C/8 (or box_chunks), 1}, no swizzle, OOB rows read as 0.
int ecg_make_act_tmap(CUtensorMap* m, const void* base, int B, int C, int L, int box_rows, int box_chunks);
// The same tensor seen as 8-byte elements: 3-D map {2*L, C/8, B}, box {2*box_rows (<= 256), box_chunks, 1}.  The inner box
// row is then box_rows * 16 contiguous bytes instead of 16: a 16-byte inner row costs the copy engine ~1 cycle per ROW
// (measured: 4.7 k cycles for a 128-channel tile), a 2 KB one moves at full rate.  Coordinates: {2 * first row, chunk, b};
// out-of-bounds rows (negative or >= L) still read as 0.  A 144-row tile = a 128-row box + a 16-row box per chunk.
int ecg_make_act_tmap64(CUtensorMap* m, const void* base, int B, int C, int L, int box_rows, int box_chunks);
