// tmembench.cu -- development microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM as a
// function of warps per CTA, load width and loads in flight per wait.  (Not part of the product.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmembench tools/tmembench.cu && tools/tmembench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[X]);
template <> __device__ __forceinline__ void ld<16>(uint32_t a, uint32_t (&u)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(a) : "memory");
}
template <> __device__ __forceinline__ void ld<32>(uint32_t a, uint32_t (&u)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                   "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                   "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]) : "r"(a) : "memory");
}

template <int X, int U>
__global__ void __launch_bounds__(1024, 1) k(int iters, uint32_t *out, long long *cyc)
{
    __shared__ uint32_t s_t;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_t)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_t + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t v[U][X];
#pragma unroll
        for (int u = 0; u < U; u++) ld<X>(base + (uint32_t)(((it * U + u) * X) & 255), v[u]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < X; j++) acc ^= v[u][j];
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_t), "r"(512u) : "memory");
}

template <int X, int U>
void run(int warps)
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k<X, U><<<148, warps * 32>>>(10, out, cyc);
    k<X, U><<<148, warps * 32>>>(iters, out, cyc);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    const double bytes = (double)iters * U * X * 4 * 32 * warps;
    printf("warps=%2d x%-2d inflight=%d : %7.1f B/cycle/SM  (%lld cycles, %.1f cyc per wait round)  %s\n", warps, X, U, bytes / h[0], h[0],
           (double)h[0] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<16, 1>(w); run<16, 2>(w); run<16, 4>(w); run<32, 1>(w); run<32, 2>(w);
    }
    return 0;
}
