// octetbench.cu -- development microbenchmark: cycles a consumer warp of the decode kernels needs per octet (8 quant
// blocks x 16 rows from shared memory -> IMMA -> scaled partial sums), alone and with the hand-off steps around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/octetbench tools/octetbench.cu && tools/octetbench
#include <cstdio>
#include <cuda_runtime.h>
#include "../vsim_b200/csrc/q4_kernels.cuh"
using namespace b200q4;

template <int QT, int V, int OPW>
__global__ void __launch_bounds__(576, 1) k(int iters, long long *out, float *sink)
{
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int CH = chunk_bytes(QT);
    uint8_t *stage = sm;                                   // 4 stages x 32 units
    uint8_t *s_aq = sm + 4 * 32 * CH;                      // 128 blocks x 32 B
    float *s_ap = reinterpret_cast<float *>(s_aq + 128 * 32);
    float *red = s_ap + 3 * 128;
    __shared__ uint64_t bar[8];
    for (int i = threadIdx.x; i < (4 * 32 * CH + 128 * 32 + 3 * 128 * 4 + 4 * 32 * 16 * 4) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(sm)[i] = (i * 2654435761u) >> 7 & 0x3f3f3f3fu;
    if (threadIdx.x < 8) mbar_init(&bar[threadIdx.x], 16 / OPW);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 16) return;
    float acc = 0.f;
    const int grp = warp / (16 / OPW), j = warp % (16 / OPW);
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int s = (it * OPW + grp) & 3;
        float v0[OPW], v2[OPW];
#pragma unroll
        for (int i = 0; i < OPW; i++) {
            const int oct = j * OPW + i;
            const uint8_t *chA = stage + (size_t)s * 32 * CH + (size_t)(2 * oct) * CH;
            compute_octet<QT, false>(chA, chA + CH, oct, s_aq, s_ap, 128, lane, v0[i], v2[i]);
        }
        if (V >= 1) {
#pragma unroll
            for (int i = 0; i < OPW; i++) {
                v0[i] += __shfl_xor_sync(0xffffffffu, v0[i], 1); v0[i] += __shfl_xor_sync(0xffffffffu, v0[i], 2);
                v2[i] += __shfl_xor_sync(0xffffffffu, v2[i], 1); v2[i] += __shfl_xor_sync(0xffffffffu, v2[i], 2);
            }
            const int g = lane >> 2, t = lane & 3;
#pragma unroll
            for (int i = 0; i < OPW; i++)
                if (t < 2) red[(s * 32 + 2 * (j * OPW + i)) * 16 + g + 8 * t] = t ? v2[i] : v0[i];
        } else {
#pragma unroll
            for (int i = 0; i < OPW; i++) acc += v0[i] + v2[i];
        }
        if (V >= 2) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[s]);
            if (V >= 3) {                                  // wait until the whole group has arrived (what a reducer would see)
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                                 : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(ph) : "memory");
                if (((it * OPW + grp) & 3) + OPW >= 4) ph ^= 1;
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 123.456f) *sink = acc;
    if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
}

// the 8-row kernel's consumer work on one stage: two units (4 quant blocks x 16 rows x 8 activation rows each) per warp,
// partials combined and stored as one float4 per lane
template <int QT, int V>
__global__ void __launch_bounds__(576, 1) k8(int iters, long long *out, float *sink)
{
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int CH = chunk_bytes(QT);
    uint8_t *stage = sm;                                   // 3 stages x 32 units
    uint8_t *s_aq = sm + 3 * 32 * CH;                      // 128 blocks x 8 rows x 32 B
    float *s_ap = reinterpret_cast<float *>(s_aq + 128 * 256);
    float *red = s_ap + 3 * 128 * 8;
    const int total = 3 * 32 * CH + 128 * 256 + 3 * 128 * 8 * 4 + 16 * 128 * 4;
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = (i * 2654435761u) >> 7 & 0x3f3f3f3fu;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 16) return;
    float sink_acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int s = it % 3;
        const uint8_t *chA = stage + (size_t)s * 32 * CH + (size_t)(2 * warp) * CH, *chB = chA + CH;
        Frag<QT> fA, fB;
        load_frag_smem<QT>(fA, chA, lane);
        load_frag_smem<QT>(fB, chB, lane);
        float aA[4] = {0.f, 0.f, 0.f, 0.f}, aB[4] = {0.f, 0.f, 0.f, 0.f};
        compute_unit_m8<QT>(fA, chA, 2 * warp, s_aq, s_ap, 128 * 8, lane, aA);
        compute_unit_m8<QT>(fB, chB, 2 * warp + 1, s_aq, s_ap, 128 * 8, lane, aB);
        aA[0] += aB[0]; aA[1] += aB[1]; aA[2] += aB[2]; aA[3] += aB[3];
        if (V >= 1) *reinterpret_cast<float4 *>(red + warp * 128 + lane * 4) = make_float4(aA[0], aA[1], aA[2], aA[3]);
        else sink_acc += aA[0] + aA[1] + aA[2] + aA[3];
    }
    const long long t1 = clock64();
    if (sink_acc == 123.456f) *sink = sink_acc;
    if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
}

template <int QT, int V>
static void run8(const char *what, long long *d_out, float *d_sink)
{
    const int iters = 2048;
    const size_t smem = 3 * 32 * chunk_bytes(QT) + 128 * 256 + 3 * 128 * 8 * 4 + 16 * 128 * 4;
    cudaFuncSetAttribute(k8<QT, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) { k8<QT, V><<<148, 576, smem>>>(iters, d_out, d_sink); cudaDeviceSynchronize(); }
    long long h[16];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 16; i++) mx = h[i] > mx ? h[i] : mx;
    printf("q4_%d M=8 %-40s: a 32-unit stage every %7.1f cycles per SM  (%s)\n", QT, what, (double)mx / iters, cudaGetErrorString(cudaGetLastError()));
}

template <int QT, int V, int OPW>
static void run(const char *what, long long *d_out, float *d_sink)
{
    const int iters = 2048;
    const size_t smem = 4 * 32 * chunk_bytes(QT) + 128 * 32 + 3 * 128 * 4 + 4 * 32 * 16 * 4;
    cudaFuncSetAttribute(k<QT, V, OPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; rep++) { k<QT, V, OPW><<<148, 576, smem>>>(iters, d_out, d_sink); cudaDeviceSynchronize(); }
    long long h[16];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 16; i++) mx = h[i] > mx ? h[i] : mx;
    printf("q4_%d %-46s %d octet(s)/warp/iter: %7.1f cycles per iteration per warp = %6.1f per octet; a 16-octet stage every %6.1f cycles per SM  (%s)\n",
           QT, what, OPW, (double)mx / iters, (double)mx / iters / OPW, (double)mx / iters / OPW, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    long long *d_out; float *d_sink;
    cudaMalloc(&d_out, 148 * 16 * 8); cudaMalloc(&d_sink, 4);
    run<0, 0, 1>("compute_octet only", d_out, d_sink);
    run<0, 1, 1>("+ quad shuffles + store", d_out, d_sink);
    run<0, 2, 1>("+ mbarrier arrive", d_out, d_sink);
    run<0, 3, 1>("+ wait for the group", d_out, d_sink);
    run<0, 0, 4>("compute_octet only", d_out, d_sink);
    run<0, 1, 4>("+ quad shuffles + store", d_out, d_sink);
    run<0, 2, 4>("+ mbarrier arrive", d_out, d_sink);
    run<0, 3, 4>("+ wait for the group", d_out, d_sink);
    run<1, 1, 1>("+ quad shuffles + store", d_out, d_sink);
    run<1, 1, 4>("+ quad shuffles + store", d_out, d_sink);
    run8<0, 0>("compute_unit_m8 x 2 only", d_out, d_sink);
    run8<0, 1>("+ store of the combined partial", d_out, d_sink);
    run8<1, 1>("+ store of the combined partial", d_out, d_sink);
    return 0;
}
