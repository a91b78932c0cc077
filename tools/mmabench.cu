// mmabench.cu -- development microbenchmark: cycles per tcgen05.mma (operands in shared memory, M = 128, K = 32 bytes)
// as the tile kernel issues them: one thread, e4m3 / bf16, K-major operand tiles without swizzle, N = 16 .. 128; with
// and without other warps loading the shared-memory pipe.  (Not part of the product.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vsim_b200/csrc -I include -o tools/mmabench tools/mmabench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "q4_tile.cuh"

using namespace b200q4;

// PAT 0: one e4m3 MMA per round; 1: the Q4_0 block (e4m3, e4m3 accumulate with the constant tile, bf16); 2: e4m3 + bf16
// LAYOUT: descriptor layout-type bits (0 = no swizzle, 2 = 128 B swizzle, 6 = 32 B swizzle: timing only, data is garbage)
template <int N, int PAT, int LAYOUT, int STORERS>
__global__ void __launch_bounds__(32 * (4 + 8), 1) k(int iters, long long *cyc, uint32_t *sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t s_t;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NA = 8;                                       // rotating operand tiles: A 4 KB each, B N x 32 B each
    uint8_t *sa = smem, *sb = smem + NA * 4096, *sx = sb + NA * 4096;
    for (int i = threadIdx.x; i < (NA * 8192 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&s_t, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_t;
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t i8 = umma_idesc_f8(N), i16 = umma_idesc_bf16(N);
            const uint64_t lay = (uint64_t)LAYOUT << 61;
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
                const int r = it & (NA - 1);
                const uint64_t ad = umma_desc(smem_u32(sa + r * 4096)) | lay, bd = umma_desc(smem_u32(sb + r * 4096)) | lay;
                const uint32_t d = tmem + (uint32_t)((it & 1) * 256);
                umma_f8(d, ad, bd, i8, 0u);
                if (PAT == 1) umma_f8(d, umma_desc(smem_u32(sa + 7 * 4096)) | lay, bd, i8, 1u);
                if (PAT >= 1) umma_bf16(d + N, umma_desc(smem_u32(sa + ((r + 3) & 7) * 4096)) | lay, umma_desc(smem_u32(sb + ((r + 3) & 7) * 4096)) | lay, i16, 0u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            cyc[blockIdx.x] = t1 - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + STORERS) {
        // the converters' traffic: 128-byte-per-warp stores into a scratch region, a 512-byte load per warp now and then
        uint32_t acc = 0;
        uint8_t *dst = sx + (warp - 4) * 4096 + lane * 4;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 16; i++) *reinterpret_cast<volatile uint32_t *>(dst + i * 128) = acc + i;
            uint4 q;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(smem_u32(sx + (warp - 4) * 4096 + lane * 16)) : "memory");
            acc += q.x + q.y + q.z + q.w;
        }
        if (acc == 0x1234567u) sink[threadIdx.x] = acc;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int PAT, int LAYOUT, int STORERS>
void run()
{
    long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4096);
    const int smem = 8 * 8192 + 32768 + 1024, iters = 4000;
    cudaFuncSetAttribute(k<N, PAT, LAYOUT, STORERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N, PAT, LAYOUT, STORERS><<<148, 32 * 12, smem>>>(64, cyc, sink);
    k<N, PAT, LAYOUT, STORERS><<<148, 32 * 12, smem>>>(iters, cyc, sink);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaGetLastError();
    const int per = PAT == 0 ? 1 : PAT == 1 ? 3 : 2;
    printf("N=%3d pattern=%d layout=%d storer warps=%d : %7.1f cycles per round (%d MMAs) = %6.1f per MMA   %s\n", N, PAT, LAYOUT, STORERS,
           (double)h[0] / iters, per, (double)h[0] / iters / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink);
}

// The tile kernel's TMEM hand-off in isolation: the issuer waits tm_empty[b], issues one block (3 MMAs), commits tm_full[b];
// EW "epilogue" warps wait tm_full[b] and arrive on tm_empty[b] (optionally after reading the buffer).  NBUF buffers.
// MODE 0: no hand-off at all, one commit per block to a barrier nobody waits on; 1: hand-off with try_wait; 2: hand-off
// with test_wait spinning; READ: the epilogue warps tcgen05.ld their part of the buffer before they release it.
template <int N, int NBUF, int MODE, int EW, int READ, int FENCE = 1, int G = 1>
__global__ void __launch_bounds__(32 * 12, 1) hs(int iters, long long *cyc, uint32_t *sink, long long *trace)
{
    constexpr int TR = 48;                                      // iterations traced (CTA 0): [4][TR] stamps
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t tm_full[4], tm_empty[4], done;
    __shared__ unsigned int rel[4];                              // MODE 6: releases of buffer b so far, all epilogue warps together
    __shared__ uint32_t s_t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NA = 8;
    uint8_t *sa = smem, *sb = smem + NA * 4096;
    for (int i = threadIdx.x; i < (NA * 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; i++) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], EW); }
        mbar_init(&done, 1);
        for (int i = 0; i < 4; i++) rel[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&s_t, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_t;
    constexpr int BUFC = 512 / NBUF;
    if (MODE == 8 && warp == 1) {                               // is a wait on an already completed barrier held up by a pending commit?
        if (elect_one()) {
            const uint32_t i8 = umma_idesc_f8(N), i16 = umma_idesc_bf16(N);
            mbar_arrive(&tm_empty[0]);                          // EW == 1 here: phase 0 of tm_empty[0] is complete for good
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
                const int r = it & (NA - 1), tb = it % NBUF;
                if (READ) mbar_wait(&tm_empty[0], 0u);
                const uint64_t ad = umma_desc(smem_u32(sa + r * 4096)), bd = umma_desc(smem_u32(sb + r * 4096));
                const uint32_t d = tmem + (uint32_t)(tb * BUFC);
                if (G) {
                    umma_f8(d, ad, bd, i8, 0u);
                    umma_f8(d, umma_desc(smem_u32(sa + 7 * 4096)), bd, i8, 1u);
                    umma_bf16(d + N, umma_desc(smem_u32(sa + ((r + 3) & 7) * 4096)), umma_desc(smem_u32(sb + ((r + 3) & 7) * 4096)), i16, 0u);
                }
                if (FENCE) umma_commit(&tm_full[tb]);
            }
            umma_commit(&done);
            mbar_wait(&done, 0);
            cyc[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    } else if (MODE == 7 && warp == 1) {                               // one issuer; the wait for step n + 1 comes BEFORE the commit of step n
        if (elect_one()) {
            const uint32_t i8 = umma_idesc_f8(N), i16 = umma_idesc_bf16(N);
            auto issue = [&](int it) {
                const int tb = it % NBUF;
                for (int g = 0; g < G; g++) {
                    const int r = (it + g) & (NA - 1);
                    const uint64_t ad = umma_desc(smem_u32(sa + r * 4096)), bd = umma_desc(smem_u32(sb + r * 4096));
                    const uint32_t d = tmem + (uint32_t)(tb * BUFC + g * 2 * N);
                    umma_f8(d, ad, bd, i8, 0u);
                    umma_f8(d, umma_desc(smem_u32(sa + 7 * 4096)), bd, i8, 1u);
                    umma_bf16(d + N, umma_desc(smem_u32(sa + ((r + 3) & 7) * 4096)), umma_desc(smem_u32(sb + ((r + 3) & 7) * 4096)), i16, 0u);
                }
            };
            const long long t0 = clock64();
            mbar_wait(&tm_empty[0], 1u);
            issue(0);
            for (int it = 0; it < iters; it++) {
                if (it + 1 < iters) { mbar_wait(&tm_empty[(it + 1) % NBUF], (uint32_t)(((it + 1) / NBUF) & 1) ^ 1u); tc_fence_after(); }
                umma_commit(&tm_full[it % NBUF]);
                if (it + 1 < iters) issue(it + 1);
            }
            umma_commit(&done);
            mbar_wait(&done, 0);
            cyc[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    } else if (MODE == 5 && (warp == 1 || warp == 2)) {               // two issuers, alternate blocks, each its own buffers
        if (elect_one()) {
            const uint32_t i8 = umma_idesc_f8(N), i16 = umma_idesc_bf16(N);
            const int w = warp - 1;
            const long long t0 = clock64();
            for (int it = w; it < iters; it += 2) {
                const int r = it & (NA - 1), tb = it % NBUF;
                mbar_wait(&tm_empty[tb], (uint32_t)((it / NBUF) & 1) ^ 1u);
                tc_fence_after();
                const uint64_t ad = umma_desc(smem_u32(sa + r * 4096)), bd = umma_desc(smem_u32(sb + r * 4096));
                const uint32_t d = tmem + (uint32_t)(tb * BUFC);
                umma_f8(d, ad, bd, i8, 0u);
                umma_f8(d, umma_desc(smem_u32(sa + 7 * 4096)), bd, i8, 1u);
                umma_bf16(d + N, umma_desc(smem_u32(sa + ((r + 3) & 7) * 4096)), umma_desc(smem_u32(sb + ((r + 3) & 7) * 4096)), i16, 0u);
                umma_commit(&tm_full[tb]);
            }
            if (w == 0) {
                umma_commit(&done);
                mbar_wait(&done, 0);
                cyc[blockIdx.x] = clock64() - t0;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t i8 = umma_idesc_f8(N), i16 = umma_idesc_bf16(N);
            RingPos tr(NBUF);
            const long long t0 = clock64();
            for (int it = 0; it < iters; it++) {
                const int r = it & (NA - 1);
                if (MODE == 1 || MODE == 4) mbar_wait(&tm_empty[tr.idx], tr.ph ^ 1u);
                if (MODE == 6) { const unsigned need = (unsigned)(it / NBUF) * EW; while (*reinterpret_cast<volatile unsigned int *>(&rel[tr.idx]) < need) { } }
                if (MODE == 2) mbar_spin(&tm_empty[tr.idx], tr.ph ^ 1u);
                if (trace && blockIdx.x == 0 && it >= 1000 && it < 1000 + TR) trace[it - 1000] = clock64();
                if (FENCE) tc_fence_after();
                for (int g = 0; g < G; g++) {
                const uint64_t ad = umma_desc(smem_u32(sa + r * 4096)), bd = umma_desc(smem_u32(sb + r * 4096));
                const uint32_t d = tmem + (uint32_t)(tr.idx * BUFC + g * 2 * N);
                umma_f8(d, ad, bd, i8, 0u);
                umma_f8(d, umma_desc(smem_u32(sa + 7 * 4096)), bd, i8, 1u);
                umma_bf16(d + N, umma_desc(smem_u32(sa + ((r + 3) & 7) * 4096)), umma_desc(smem_u32(sb + ((r + 3) & 7) * 4096)), i16, 0u);
                }
                umma_commit(&tm_full[tr.idx]);
                if (trace && blockIdx.x == 0 && it >= 1000 && it < 1000 + TR) trace[TR + it - 1000] = clock64();
                tr.next();
            }
            umma_commit(&done);
            mbar_wait(&done, 0);
            cyc[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + EW && MODE != 0 && MODE != 8) {
        RingPos tr(NBUF);
        const int ew = warp - 4;
        const uint32_t t_lane = tmem + ((uint32_t)((ew & 3) * 32) << 16) + (uint32_t)((ew >> 2) * (N / 2));
        float acc = 0.f;
        for (int it = 0; it < iters; it++) {
            if (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 6 || MODE == 7) mbar_wait(&tm_full[tr.idx], tr.ph); else if (MODE == 2) mbar_spin(&tm_full[tr.idx], tr.ph);
            tc_fence_after();
            if (trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it >= 1000 && it < 1000 + TR) trace[2 * TR + it - 1000] = clock64();
            if (READ) {
#pragma unroll
                for (int c = 0; c < N / 2; c += 16) {
                    float v[16], t[16];
                    tmem_ld<16>(t_lane + (uint32_t)(tr.idx * BUFC + c), v);
                    tmem_ld<16>(t_lane + (uint32_t)(tr.idx * BUFC + N + c), t);
                    tmem_wait_ld<16>(v, t);
#pragma unroll
                    for (int j = 0; j < 16; j++) acc = fmaf(v[j], t[j], acc);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && MODE != 3 && MODE != 6) mbar_arrive(&tm_empty[tr.idx]);
            if (lane == 0 && MODE == 6) atomicAdd(&rel[tr.idx], 1u);
            if (trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it >= 1000 && it < 1000 + TR) trace[3 * TR + it - 1000] = clock64();
            tr.next();
        }
        if (acc == 1.2345f) sink[threadIdx.x] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static bool g_trace = false;
template <int N, int NBUF, int MODE, int EW, int READ, int FENCE = 1, int G = 1>
void run_hs()
{
    long long *cyc; uint32_t *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4096);
    const int smem = 8 * 8192 + 1024, iters = 4000;
    cudaFuncSetAttribute(hs<N, NBUF, MODE, EW, READ, FENCE, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long *trace; cudaMalloc(&trace, 4 * 48 * 8); cudaMemset(trace, 0, 4 * 48 * 8);
    hs<N, NBUF, MODE, EW, READ, FENCE, G><<<148, 32 * 12, smem>>>(64, cyc, sink, nullptr);
    hs<N, NBUF, MODE, EW, READ, FENCE, G><<<148, 32 * 12, smem>>>(iters, cyc, sink, g_trace ? trace : nullptr);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaGetLastError();
    printf("hand-off N=%3d buffers=%d blocks/buffer=%d fence=%d mode=%d (%s) epilogue warps=%d read=%d : %7.1f cycles per block   %s\n", N, NBUF, G, FENCE, MODE,
           MODE == 0 ? "commit only" : MODE == 1 ? "try_wait" : MODE == 2 ? "test_wait spin" : MODE == 5 ? "try_wait, two issuers" : MODE == 6 ? "release counter in shared memory" : MODE == 7 ? "try_wait, wait(n+1) before commit(n)" : MODE == 8 ? "no partner; read=wait on a completed barrier, fence=commit, blocks=MMAs" : "-", EW, READ, (double)h[0] / iters / (G ? G : 1), e == cudaSuccess ? "" : cudaGetErrorString(e));
    if (g_trace) {
        long long t[4 * 48]; cudaMemcpy(t, trace, sizeof t, cudaMemcpyDeviceToHost);
        printf("   it: issuer got tm_empty | issuer committed | epilogue warp 0 saw tm_full | epilogue warp 0 arrived   (cycles since the first stamp)\n");
        for (int i = 0; i < 24; i++) printf("   %2d: %6lld %6lld %6lld %6lld\n", i, t[i] - t[0], t[48 + i] - t[0], t[96 + i] - t[0], t[144 + i] - t[0]);
    }
    cudaFree(cyc); cudaFree(sink); cudaFree(trace);
}

int main(int argc, char **argv)
{
    g_trace = argc > 1;                                          // any argument: stamp traces of three hand-off variants only
    if (g_trace) { run_hs<16, 4, 1, 8, 0>(); run_hs<128, 2, 1, 8, 0>(); run_hs<128, 2, 1, 8, 1>(); return 0; }
    printf("-- cycles per tcgen05.mma (M = 128, K = 32 bytes, operands in shared memory), one issuing thread, no hand-off\n");
    run<16, 0, 0, 0>(); run<32, 0, 0, 0>(); run<64, 0, 0, 0>(); run<128, 0, 0, 0>();
    run<16, 1, 0, 0>(); run<32, 1, 0, 0>(); run<64, 1, 0, 0>(); run<128, 1, 0, 0>();
    run<16, 2, 0, 0>(); run<64, 2, 0, 0>();
    run<16, 0, 2, 0>(); run<64, 0, 2, 0>(); run<16, 0, 6, 0>(); run<64, 0, 6, 0>();
    run<16, 1, 0, 4>(); run<64, 1, 0, 4>(); run<128, 1, 0, 4>(); run<16, 1, 0, 8>(); run<64, 1, 0, 8>();
    printf("-- what a wait costs the issuing thread (no partner warps)\n");
    run_hs<16, 4, 8, 1, 0, 1, 1>(); run_hs<16, 4, 8, 1, 1, 1, 1>(); run_hs<16, 4, 8, 1, 1, 0, 1>(); run_hs<16, 4, 8, 1, 1, 1, 0>(); run_hs<16, 4, 8, 1, 0, 1, 0>();
    printf("-- the TMEM hand-off (issuer: wait tm_empty, one block = 3 MMAs per block of the buffer, commit tm_full; epilogue warps: wait tm_full, arrive tm_empty)\n");
    run_hs<128, 2, 0, 8, 0>(); run_hs<128, 2, 1, 8, 0>(); run_hs<128, 2, 2, 8, 0>(); run_hs<128, 2, 1, 1, 0>(); run_hs<128, 2, 5, 8, 0>(); run_hs<128, 2, 6, 8, 0>(); run_hs<128, 2, 7, 8, 0>();
    run_hs<64, 4, 0, 8, 0>(); run_hs<64, 4, 1, 8, 0>(); run_hs<64, 4, 2, 8, 0>(); run_hs<64, 4, 5, 8, 0>(); run_hs<64, 4, 6, 8, 0>(); run_hs<64, 4, 7, 8, 0>();
    run_hs<16, 4, 0, 8, 0>(); run_hs<16, 4, 1, 8, 0>(); run_hs<16, 4, 1, 8, 0, 0>(); run_hs<16, 4, 2, 8, 0>(); run_hs<16, 4, 5, 8, 0>(); run_hs<16, 4, 6, 8, 0>(); run_hs<16, 4, 7, 8, 0>();
    run_hs<16, 4, 1, 8, 0, 1, 4>(); run_hs<16, 4, 7, 8, 0, 1, 4>();
    printf("-- the same with the epilogue warps reading their part of the buffer (unpipelined tcgen05.ld x16 pairs)\n");
    run_hs<128, 2, 1, 8, 1>(); run_hs<128, 2, 5, 8, 1>(); run_hs<64, 4, 1, 8, 1>(); run_hs<64, 2, 1, 8, 1>(); run_hs<64, 4, 5, 8, 1>(); run_hs<32, 4, 7, 8, 1, 1, 2>();
    return 0;
}
