// Shared helpers for the ecgb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ecgb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "ecgb200 kernels are written for sm_100a only"
#endif

#define ECG_KS 15
#define ECG_PAD 7

static inline int ecg_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

static inline int ecg_cdiv(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL): when enabled (ecgb200_set_pdl), the kernels of the step's critical path are
// launched with the programmatic-stream-serialization attribute, so a kernel's CTAs are scheduled -- and run their
// prologue (barrier init, TMEM allocation, index math) -- while the previous kernel drains.  Every such kernel
// calls ecg_pdl_wait() before it touches global memory, which blocks until the previous kernel has COMPLETED and
// flushed (full dependency semantics), and ecg_pdl_launch_dependents() at its top.
extern int g_ecg_pdl;
__device__ __forceinline__ void ecg_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ecg_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KP, typename... KA>
static inline int ecg_launch_pdl_if(bool on, void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    KA... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = on ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    return e == cudaSuccess ? ecg_launch_status() : (int)e;
}
template <typename... KP, typename... KA>
static inline int ecg_launch_pdl(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, KA... args) {
    return ecg_launch_pdl_if(g_ecg_pdl != 0, kernel, grid, block, smem, st, args...);
}

// One AdamW element update (torch.optim.AdamW order of operations, src/training/loop.py:34) with EVERY rounding pinned by
// intrinsics -- nothing is left to the compiler's FMA contraction, which may differ from kernel to kernel -- so that the
// multi-tensor, flat, barrier-exchange and one-hop-exchange kernels all produce the same bits from the same inputs.
struct AdamK { float decay, one_m_b1, b2, one_m_b2, bc2, eps, ss; };
__device__ __forceinline__ void adamw_update(float& p, float& m, float& v, float g, const AdamK& K) {
    const float pi = __fmul_rn(p, K.decay);                                          // decoupled weight decay
    m = __fadd_rn(m, __fmul_rn(K.one_m_b1, __fsub_rn(g, m)));                         // lerp
    v = __fadd_rn(__fmul_rn(v, K.b2), __fmul_rn(__fmul_rn(K.one_m_b2, g), g));        // mul + addcmul
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), K.bc2), K.eps);
    p = __fsub_rn(pi, __fmul_rn(K.ss, __fdiv_rn(m, denom)));                          // addcdiv
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of a double; every thread gets the result. blockDim.x <= 1024.
__device__ __forceinline__ double block_sum_d(double v, double* sh /* >= 33 doubles */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}
