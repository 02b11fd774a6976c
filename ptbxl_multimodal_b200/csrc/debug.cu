// Diagnostics that are not part of the hot path: device timestamps between the nodes of a captured step
// (the schedule the two-stream graph REALLY runs, without nsys) and the limit of the bounded spin waits.
#include "common.cuh"

void ecg_set_timeout_conv(unsigned long long ns);      // conv1d_tc.cu: mbarrier waits of the tcgen05 kernels
void ecg_set_timeout_dp(unsigned long long ns);        // dp_fused.cu: cross-rank flag waits

__global__ void stamp_kernel(unsigned long long* buf, int idx) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[idx] = t;
}

extern "C" int ecgb200_debug_stamp(unsigned long long* buf, int idx, void* stream) {
    if (!buf || idx < 0) return ECGB200_EINVAL;
    stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(buf, idx);
    return ecg_launch_status();
}

extern "C" int ecgb200_set_spin_timeout_ms(unsigned int mbarrier_ms, unsigned int peer_ms) {
    ecg_set_timeout_conv((unsigned long long)mbarrier_ms * 1000000ull);
    ecg_set_timeout_dp((unsigned long long)peer_ms * 1000000ull);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}
