// launchbench.cu -- development microbenchmark: GPU-side cost per back-to-back kernel launch on
// this box (plain stream launches, CUDA graph, programmatic dependent launch).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void empty_kernel(uint32_t *out) { if (threadIdx.x == 9999) out[0] = 1; }
__global__ void pdl_kernel(uint32_t *out)
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    if (threadIdx.x == 9999) out[0] = 1;
}
__global__ void spin_kernel(uint32_t *out, long long cycles)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 9999) out[0] = 1;
}
__global__ void spin_pdl_kernel(uint32_t *out, long long cycles)
{
    cudaTriggerProgrammaticLaunchCompletion();     // let the next kernel's CTAs start as soon as they fit
    cudaGridDependencySynchronize();               // ... but order our work after the previous kernel
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 9999) out[0] = 1;
}

template <typename F> float time_it(F f, int iters, cudaStream_t st)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(10);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    f(iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / iters;
}

int main()
{
    uint32_t *o; cudaMalloc(&o, 64);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int threads : {32, 512}) for (int grid : {1, 148}) {
        float us = time_it([&](int n) { for (int i = 0; i < n; i++) empty_kernel<<<grid, threads, 0, st>>>(o); }, 500, st);
        printf("plain   empty <<<%3d,%3d>>>: %.2f us/launch\n", grid, threads, us);
    }
    {   // graph of 100 empty kernels
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 100; i++) empty_kernel<<<148, 512, 0, st>>>(o);
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        float us = time_it([&](int n) { for (int i = 0; i < n / 100; i++) cudaGraphLaunch(ge, st); }, 1000, st);
        printf("graph   empty <<<148,512>>>: %.2f us/kernel\n", us);
    }
    auto launch_pdl = [&](void (*k)(uint32_t *), int grid, int threads) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k, o);
    };
    {
        float us = time_it([&](int n) { for (int i = 0; i < n; i++) launch_pdl(pdl_kernel, 148, 512); }, 500, st);
        printf("PDL     empty <<<148,512>>>: %.2f us/launch  (%s)\n", us, cudaGetErrorString(cudaGetLastError()));
    }
    {   // graph + PDL
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 100; i++) launch_pdl(pdl_kernel, 148, 512);
        cudaStreamEndCapture(st, &g);
        cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
        float us = time_it([&](int n) { for (int i = 0; i < n / 100; i++) cudaGraphLaunch(ge, st); }, 1000, st);
        printf("graph+PDL empty <<<148,512>>>: %.2f us/kernel (%s)\n", us, cudaGetErrorString(e));
    }
    // kernels that run ~6 us each: how much of the launch cost hides behind execution?
    const long long cyc = 6 * 1900;   // ~6 us at 1.9 GHz
    {
        float us = time_it([&](int n) { for (int i = 0; i < n; i++) spin_kernel<<<148, 512, 0, st>>>(o, cyc); }, 300, st);
        printf("plain   spin6us: %.2f us/launch\n", us);
        float us1 = time_it([&](int n) { for (int i = 0; i < n; i++) spin_kernel<<<1, 32, 0, st>>>(o, cyc); }, 300, st);
        printf("plain   spin6us <<<1,32>>>: %.2f us/launch\n", us1);
    }
    {
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 100; i++) spin_kernel<<<148, 512, 0, st>>>(o, cyc);
        cudaStreamEndCapture(st, &g);
        cudaGraphInstantiate(&ge, g, 0);
        float us = time_it([&](int n) { for (int i = 0; i < n / 100; i++) cudaGraphLaunch(ge, st); }, 1000, st);
        printf("graph   spin6us: %.2f us/kernel\n", us);
    }
    {
        auto launch_spin_pdl = [&]() {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(148); cfg.blockDim = dim3(512); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, spin_pdl_kernel, o, cyc);
        };
        float us = time_it([&](int n) { for (int i = 0; i < n; i++) launch_spin_pdl(); }, 300, st);
        printf("PDL     spin6us: %.2f us/launch (%s)\n", us, cudaGetErrorString(cudaGetLastError()));
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 100; i++) launch_spin_pdl();
        cudaStreamEndCapture(st, &g);
        cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
        float us2 = time_it([&](int n) { for (int i = 0; i < n / 100; i++) cudaGraphLaunch(ge, st); }, 1000, st);
        printf("graph+PDL spin6us: %.2f us/kernel (%s)\n", us2, cudaGetErrorString(e));
    }
    return 0;
}
