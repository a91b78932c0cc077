// membench.cu -- development microbenchmark: how fast can 148 persistent CTAs stream a 42 MB
// tensor in 1280-byte units under different unit->warp assignments?  (Not part of the product.)
//   mode 0: one contiguous unit range per warp            (what q4_gemv_kernel r01 does)
//   mode 1: one contiguous range per CTA, warps interleaved unit by unit inside it
//   mode 2: global interleave: unit u -> warp u % Wtot at step u / Wtot
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u1(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

struct U { uint4 a, b; uint32_t c, d; };
__device__ __forceinline__ void load_unit(U &f, const uint8_t *chunk, int lane)
{
    f.a = ldg_stream_u4(chunk + lane * 16);
    f.b = ldg_stream_u4(chunk + 512 + lane * 16);
    f.c = ldg_stream_u1(chunk + 1024 + lane * 4);
    f.d = ldg_stream_u1(chunk + 1152 + lane * 4);
}

template <int MODE, int PD, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) stream_kernel(const uint8_t *wt, int U_total, uint32_t *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Wtot = gridDim.x * WARPS;
    const int wid = blockIdx.x * WARPS + warp;
    int first, count, stride;
    if (MODE == 0) {
        const int u0 = (int)(((long)wid * U_total) / Wtot), u1 = (int)(((long)(wid + 1) * U_total) / Wtot);
        first = u0; count = u1 - u0; stride = 1;
    } else if (MODE == 1) {
        const int c0 = (int)(((long)blockIdx.x * U_total) / gridDim.x), c1 = (int)(((long)(blockIdx.x + 1) * U_total) / gridDim.x);
        first = c0 + warp; count = (c1 - first + WARPS - 1) / WARPS; stride = WARPS;
        if (count < 0) count = 0;
    } else {
        first = wid; count = (U_total - first + Wtot - 1) / Wtot; stride = Wtot;
        if (count < 0) count = 0;
    }
    U buf[PD];
#pragma unroll
    for (int i = 0; i < PD; i++)
        if (i < count) load_unit(buf[i], wt + (size_t)(first + i * stride) * 1280, lane);
    uint32_t acc = 0;
    for (int b = 0; b < count; b += PD) {
#pragma unroll
        for (int i = 0; i < PD; i++) {
            const int k = b + i;
            if (k < count) {
                acc ^= buf[i].a.x ^ buf[i].a.y ^ buf[i].a.z ^ buf[i].a.w ^ buf[i].b.x ^ buf[i].b.y ^ buf[i].b.z ^ buf[i].b.w ^ buf[i].c ^ buf[i].d;
                if (k + PD < count) load_unit(buf[i], wt + (size_t)(first + (k + PD) * stride) * 1280, lane);
            }
        }
    }
    if (acc == 0x12345678u) out[wid] = acc;
}

// ---- per-warp shared-memory ring fed by cp.async.bulk (1280 B per copy), depth D -----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int D, int WARPS, int UPC>   // UPC = units per bulk copy
__global__ void __launch_bounds__(WARPS * 32, 1) bulk_kernel(const uint8_t *wt, int U_total, uint32_t *out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Wtot = gridDim.x * WARPS;
    const int wid = blockIdx.x * WARPS + warp;
    constexpr int SLOT = 1280 * UPC;
    uint8_t *ring = sm + (size_t)warp * D * SLOT;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + (size_t)WARPS * D * SLOT) + warp * D;
    const int u0 = (int)(((long)wid * U_total) / Wtot), u1 = (int)(((long)(wid + 1) * U_total) / Wtot);
    const int nslots = (u1 - u0 + UPC - 1) / UPC;
    if (lane == 0) {
        for (int i = 0; i < D; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < D && i < nslots; i++) {
            const int ub = u0 + i * UPC, n = min(UPC, u1 - ub);
            mbar_expect_tx(&bars[i], n * 1280);
            bulk_g2s(ring + i * SLOT, wt + (size_t)ub * 1280, n * 1280, &bars[i]);
        }
    }
    __syncwarp();
    uint32_t acc = 0;
    for (int k = 0; k < nslots; k++) {
        const int s = k % D;
        mbar_wait(&bars[s], (k / D) & 1);
        const int ub = u0 + k * UPC, n = min(UPC, u1 - ub);
        for (int j = 0; j < n; j++) {
            const uint8_t *c = ring + s * SLOT + j * 1280;
            const uint4 a = *reinterpret_cast<const uint4 *>(c + lane * 16), b = *reinterpret_cast<const uint4 *>(c + 512 + lane * 16);
            const uint32_t e = *reinterpret_cast<const uint32_t *>(c + 1024 + lane * 4), f = *reinterpret_cast<const uint32_t *>(c + 1152 + lane * 4);
            acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ e ^ f;
        }
        __syncwarp();
        if (lane == 0 && k + D < nslots) {
            const int ub2 = u0 + (k + D) * UPC, n2 = min(UPC, u1 - ub2);
            mbar_expect_tx(&bars[s], n2 * 1280);
            bulk_g2s(ring + s * SLOT, wt + (size_t)ub2 * 1280, n2 * 1280, &bars[s]);
        }
    }
    if (acc == 0x12345678u) out[wid] = acc;
}

// ---- mode0 register ring + L2 prefetch of the warp's range by cp.async.bulk.prefetch.L2 ----------
template <int PD, int WARPS, int AHEAD>   // AHEAD = 0: whole range at start; else sliding window of AHEAD units
__global__ void __launch_bounds__(WARPS * 32, 1) pf_kernel(const uint8_t *wt, int U_total, uint32_t *out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Wtot = gridDim.x * WARPS;
    const int wid = blockIdx.x * WARPS + warp;
    const int u0 = (int)(((long)wid * U_total) / Wtot), u1 = (int)(((long)(wid + 1) * U_total) / Wtot);
    const int count = u1 - u0;
    if (AHEAD == 0) {
        // 32 lanes split the range into 32 pieces (multiples of 16 bytes)
        const size_t total = (size_t)count * 1280, piece = ((total / 32) + 15) & ~(size_t)15;
        const size_t off = piece * lane;
        if (off < total) {
            const uint32_t sz = (uint32_t)min(piece, total - off);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(wt + (size_t)u0 * 1280 + off), "r"(sz) : "memory");
        }
    } else {
        if (lane < AHEAD && lane < count)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(wt + (size_t)(u0 + lane) * 1280), "r"(1280) : "memory");
    }
    U buf[PD];
#pragma unroll
    for (int i = 0; i < PD; i++)
        if (i < count) load_unit(buf[i], wt + (size_t)(u0 + i) * 1280, lane);
    uint32_t acc = 0;
    for (int b = 0; b < count; b += PD) {
#pragma unroll
        for (int i = 0; i < PD; i++) {
            const int k = b + i;
            if (k < count) {
                acc ^= buf[i].a.x ^ buf[i].a.y ^ buf[i].a.z ^ buf[i].a.w ^ buf[i].b.x ^ buf[i].b.y ^ buf[i].b.z ^ buf[i].b.w ^ buf[i].c ^ buf[i].d;
                if (k + PD < count) load_unit(buf[i], wt + (size_t)(u0 + k + PD) * 1280, lane);
                if (AHEAD > 0 && lane == 0 && k + AHEAD < count)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(wt + (size_t)(u0 + k + AHEAD) * 1280), "r"(1280) : "memory");
            }
        }
    }
    if (acc == 0x12345678u) out[wid] = acc;
}

template <typename L>
void time_graph(const char *name, L launch, int nbuf, size_t bytes)
{
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 100; i++) launch(i % nbuf, st);
    cudaStreamEndCapture(st, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 5; i++) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / 500;
    printf("[graph] %-40s %7.2f us  %7.1f GB/s  (%s / %s)\n", name, us, (double)bytes / (us * 1e-6) / 1e9, cudaGetErrorString(e), cudaGetErrorString(cudaGetLastError()));
}

template <int MODE, int PD, int WARPS>
void run(const char *name, uint8_t **bufs, int nbuf, int U_total, uint32_t *out, int grid)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 10; i++) stream_kernel<MODE, PD, WARPS><<<grid, WARPS * 32>>>(bufs[i % nbuf], U_total, out);
    cudaEventRecord(e0);
    const int iters = 100;
    for (int i = 0; i < iters; i++) stream_kernel<MODE, PD, WARPS><<<grid, WARPS * 32>>>(bufs[i % nbuf], U_total, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / iters, gbs = (double)U_total * 1280 / (us * 1e-6) / 1e9;
    printf("%-44s grid=%3d  %7.2f us  %7.1f GB/s  (%s)\n", name, grid, us, gbs, cudaGetErrorString(cudaGetLastError()));
}

__global__ void empty_kernel(uint32_t *out) { if (threadIdx.x == 9999) out[0] = 1; }

// classic grid-stride copy-style reader: many small CTAs, each thread 4 x uint4 in flight
__global__ void __launch_bounds__(256) gridstride_kernel(const uint4 *p, size_t n16, uint32_t *out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = ldg_stream_u4(p + i), b = ldg_stream_u4(p + i + stride), c = ldg_stream_u4(p + i + 2 * stride), d = ldg_stream_u4(p + i + 3 * stride);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n16; i += stride) { uint4 a = ldg_stream_u4(p + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

void run_gs(const char *name, uint8_t **bufs, int nbuf, size_t bytes, uint32_t *out, int grid)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 10; i++) gridstride_kernel<<<grid, 256>>>((const uint4 *)bufs[i % nbuf], bytes / 16, out);
    cudaEventRecord(e0);
    const int iters = 100;
    for (int i = 0; i < iters; i++) gridstride_kernel<<<grid, 256>>>((const uint4 *)bufs[i % nbuf], bytes / 16, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / iters;
    printf("%-44s grid=%5d  %7.2f us  %7.1f GB/s\n", name, grid, us, (double)bytes / (us * 1e-6) / 1e9);
}

int main()
{
    {
        uint32_t *o; cudaMalloc(&o, 64);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 10; i++) empty_kernel<<<148, 512>>>(o);
        cudaEventRecord(e0);
        for (int i = 0; i < 200; i++) empty_kernel<<<148, 512>>>(o);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("empty kernel back-to-back: %.2f us per launch\n", ms * 1e3 / 200);
    }
    const int U_total = 1024 * 32;                 // N=16384, K=4096, Q4_0
    const size_t bytes = (size_t)U_total * 1280;
    const int nbuf = 10;
    uint8_t *bufs[nbuf];
    for (int i = 0; i < nbuf; i++) { cudaMalloc(&bufs[i], bytes); cudaMemset(bufs[i], i + 1, bytes); }
    uint32_t *out; cudaMalloc(&out, 1 << 20);
    run<0, 4, 16>("mode0 per-warp ranges       PD=4 W=16", bufs, nbuf, U_total, out, 148);
    run<0, 8, 16>("mode0 per-warp ranges       PD=8 W=16", bufs, nbuf, U_total, out, 148);
    run<0, 4, 32>("mode0 per-warp ranges       PD=4 W=32", bufs, nbuf, U_total, out, 148);
    run<1, 4, 16>("mode1 per-CTA range         PD=4 W=16", bufs, nbuf, U_total, out, 148);
    run<1, 8, 16>("mode1 per-CTA range         PD=8 W=16", bufs, nbuf, U_total, out, 148);
    run<1, 4, 32>("mode1 per-CTA range         PD=4 W=32", bufs, nbuf, U_total, out, 148);
    run<2, 4, 16>("mode2 global interleave     PD=4 W=16", bufs, nbuf, U_total, out, 148);
    run<2, 8, 16>("mode2 global interleave     PD=8 W=16", bufs, nbuf, U_total, out, 148);
    run<2, 4, 32>("mode2 global interleave     PD=4 W=32", bufs, nbuf, U_total, out, 148);
    run<2, 2, 32>("mode2 global interleave     PD=2 W=32", bufs, nbuf, U_total, out, 148);
    run<2, 4, 8>("mode2 global interleave     PD=4 W=8 x2CTA", bufs, nbuf, U_total, out, 296);
    run<0, 4, 8>("mode0 per-warp ranges       PD=4 W=8 x2CTA", bufs, nbuf, U_total, out, 296);
    // ---- inside a CUDA graph (launch gap ~0.4 us instead of ~2.2 us) ----
    time_graph("mode0 PD=4 W=16", [&](int b, cudaStream_t st) { stream_kernel<0, 4, 16><<<148, 512, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("mode0 PD=4 W=32", [&](int b, cudaStream_t st) { stream_kernel<0, 4, 32><<<148, 1024, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("mode0 PD=2 W=32", [&](int b, cudaStream_t st) { stream_kernel<0, 2, 32><<<148, 1024, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("grid-stride 2368 CTAs", [&](int b, cudaStream_t st) { gridstride_kernel<<<2368, 256, 0, st>>>((const uint4 *)bufs[b], bytes / 16, out); }, nbuf, bytes);
    time_graph("L2pf whole-range PD=4 W=16", [&](int b, cudaStream_t st) { pf_kernel<4, 16, 0><<<148, 512, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("L2pf whole-range PD=2 W=16", [&](int b, cudaStream_t st) { pf_kernel<2, 16, 0><<<148, 512, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("L2pf whole-range PD=2 W=32", [&](int b, cudaStream_t st) { pf_kernel<2, 32, 0><<<148, 1024, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("L2pf ahead=8 PD=4 W=16", [&](int b, cudaStream_t st) { pf_kernel<4, 16, 8><<<148, 512, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("L2pf ahead=12 PD=2 W=16", [&](int b, cudaStream_t st) { pf_kernel<2, 16, 12><<<148, 512, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
    time_graph("L2pf whole-range PD=4 W=8 x2", [&](int b, cudaStream_t st) { pf_kernel<4, 8, 0><<<296, 256, 0, st>>>(bufs[b], U_total, out); }, nbuf, bytes);
#define BULK(D, W, UPC) { \
        const size_t smem = (size_t)W * D * 1280 * UPC + (size_t)W * D * 8; \
        cudaFuncSetAttribute(bulk_kernel<D, W, UPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        time_graph("bulk ring D=" #D " W=" #W " UPC=" #UPC, [&](int b, cudaStream_t st) { bulk_kernel<D, W, UPC><<<148, W * 32, smem, st>>>(bufs[b], U_total, out); }, nbuf, bytes); }
    BULK(4, 16, 1)
    BULK(8, 16, 1)
    BULK(4, 16, 2)
    BULK(4, 32, 1)
    BULK(2, 16, 4)
    BULK(4, 8, 4)
    BULK(3, 16, 3)
    run_gs("grid-stride 4x uint4/thread", bufs, nbuf, bytes, out, 148 * 8);
    run_gs("grid-stride 4x uint4/thread", bufs, nbuf, bytes, out, 148 * 16);
    run_gs("grid-stride 4x uint4/thread", bufs, nbuf, bytes, out, 148 * 32);
    run_gs("grid-stride 4x uint4/thread", bufs, nbuf, bytes, out, 148 * 64);
    {   // asymptote: one 420 MB region (the 10 buffers are separate allocations; use a fresh big one)
        uint8_t *big; const size_t bb = (size_t)1 << 30; cudaMalloc(&big, bb); cudaMemset(big, 3, bb);
        uint8_t *one[1] = { big };
        run_gs("grid-stride 1 GiB", one, 1, bb, out, 148 * 32);
        run<0, 4, 16>("mode0 PD=4 W=16  1 GiB", one, 1, (int)(bb / 1280), out, 148);
        run<0, 4, 32>("mode0 PD=4 W=32  1 GiB", one, 1, (int)(bb / 1280), out, 148);
        run<2, 4, 32>("mode2 PD=4 W=32  1 GiB", one, 1, (int)(bb / 1280), out, 148);
        uint8_t *q[1] = { big };
        run_gs("grid-stride 168 MB", q, 1, bytes * 4, out, 148 * 32);
        run<0, 4, 32>("mode0 PD=4 W=32  168 MB", q, 1, U_total * 4, out, 148);
    }
    return 0;
}
