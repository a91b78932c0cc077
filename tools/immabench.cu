// immabench.cu -- development microbenchmark: cost of the legacy mma.sync m16n8k32 u8 x s8 (IMMA.16832) on sm_100a,
// as a dependent chain and as independent streams, for 1..4 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/immabench tools/immabench.cu && tools/immabench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(int iters, long long *out, int *sink)
{
    unsigned a[4] = {threadIdx.x, 2u, 3u, 4u};
    int d[CHAINS][4];
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) d[c][i] = c + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) imma(d[c], a, (unsigned)it, 7u);
    }
    const long long t1 = clock64();
    int s = 0;
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) s += d[c][i];
    if (s == 123456789) *sink = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main()
{
    long long *d_out; int *d_sink;
    cudaMalloc(&d_out, 1024 * 8); cudaMalloc(&d_sink, 4);
    const int iters = 4096;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int chains = 1; chains <= 4; chains *= 2) {
            long long h = 0;
            for (int rep = 0; rep < 2; rep++) {
                if (chains == 1) k<1><<<148, warps * 32>>>(iters, d_out, d_sink);
                else if (chains == 2) k<2><<<148, warps * 32>>>(iters, d_out, d_sink);
                else k<4><<<148, warps * 32>>>(iters, d_out, d_sink);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d (%d per SMSP), %d independent chains per warp: %.1f cycles per IMMA per warp, %.2f cycles per IMMA per SMSP\n",
                   warps, warps / 4, chains, (double)h / iters / chains, (double)h / iters / chains / (warps / 4));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
