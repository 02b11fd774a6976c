// Launch latency of a chain of small kernels inside a CUDA graph, with and without programmatic dependent launch.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_plain(float* p, int spin) {
    float v = p[threadIdx.x];
    for (int i = 0; i < spin; ++i) v = v * 1.0001f + 0.5f;
    p[threadIdx.x] = v;
}
__global__ void k_pdl(float* p, int spin, int late) {
    if (!late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float v = p[threadIdx.x];
    for (int i = 0; i < spin; ++i) v = v * 1.0001f + 0.5f;
    if (late) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    p[threadIdx.x] = v;
}
static float run(int mode, int grid, int spin, int n, float* d) {
    cudaStream_t s; cudaStreamCreate(&s);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < n; ++i) {
        if (mode == 0) k_plain<<<grid, 256, 0, s>>>(d, spin);
        else {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = a; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_pdl, d, spin, mode == 2 ? 1 : 0);
        }
    }
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    return ms * 1000.f / (10 * n);
}
int main() {
    float* d; cudaMalloc(&d, 1 << 20); cudaMemset(d, 0, 1 << 20);
    const char* names[3] = {"plain", "pdl trigger at top", "pdl trigger late"};
    for (int grid : {1, 148, 592}) for (int spin : {0, 2000, 20000})
        for (int mode = 0; mode < 3; ++mode)
            printf("grid %4d spin %6d  %-20s %7.2f us per kernel (chain of 24 in a graph)\n", grid, spin, names[mode], run(mode, grid, spin, 24, d));
    return 0;
}
