// Development aid: NVLink peer-store / fence / flag latencies between two B200s of one box (single process).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2pbench tools/p2pbench.cu && tools/p2pbench
// Sizes the fused all-gather protocol of q4_kernels.cuh (what one round trip costs, what a release costs).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acq(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_rlx(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// mode 0: remote store + fence.sys; 1: remote red.release; 2: remote st.release after a remote data store;
// 3: local store + fence.sys (baseline); 4: remote store + fence.acq_rel.gpu
__global__ void k_fence(unsigned *remote, unsigned *local, long long *out, int iters, int mode)
{
    for (int i = 0; i < iters; i++) {
        const unsigned long long t0 = gtimer();
        if (mode == 0) { remote[i & 63] = i; __threadfence_system(); }
        else if (mode == 1) { asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(remote), "r"(1u) : "memory"); }
        else if (mode == 2) { remote[64 + (i & 63)] = i; asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(remote + 1), "r"((unsigned)i) : "memory"); __threadfence_system(); }
        else if (mode == 3) { local[i & 63] = i; __threadfence_system(); }
        else { remote[i & 63] = i; __threadfence(); }
        const unsigned long long t1 = gtimer();
        out[i] = (long long)(t1 - t0);
        __nanosleep(2000);
    }
}

// ping-pong on flags: side 0 starts.  proto 0: bare relaxed flag store; 1: data store + fence.sys + flag store;
// 2: data store + red.release.sys.add flag
__global__ void k_pingpong(unsigned *my_flag, unsigned *peer_flag, unsigned *peer_data, int side, int iters, int proto, long long *out)
{
    const unsigned long long t0 = gtimer();
    for (int i = 1; i <= iters; i++) {
        if (side == 1 || i > 1) {
            // wait for the peer's signal number (side 0 waits for i-1 echoes, side 1 for i)
            const unsigned want = side == 0 ? (unsigned)(i - 1) : (unsigned)i;
            while ((int)(ld_acq(my_flag) - want) < 0) { }
        }
        if (proto >= 1) { peer_data[i & 255] = i; }
        if (proto == 1) { __threadfence_system(); asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(peer_flag), "r"((unsigned)i) : "memory"); }
        else if (proto == 2) { asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(peer_flag), "r"(1u) : "memory"); }
        else { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(peer_flag), "r"((unsigned)i) : "memory"); }
    }
    if (side == 0) { while ((int)(ld_acq(my_flag) - (unsigned)iters) < 0) { } out[0] = (long long)(gtimer() - t0); }
}

int main()
{
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    unsigned *buf[2]; long long *out[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&buf[d], 4096)); CK(cudaMemset(buf[d], 0, 4096));
        CK(cudaMalloc(&out[d], 8 * 4096));
    }
    const int iters = 200;
    std::vector<long long> h(iters);
    const char *names[] = {"remote st + fence.sys", "remote red.release.sys.add (issue only)", "remote st; st.release.sys; fence.sys", "local st + fence.sys", "remote st + fence.gpu"};
    CK(cudaSetDevice(0));
    for (int mode = 0; mode < 5; mode++) {
        k_fence<<<1, 1>>>(buf[1], buf[0], out[0], iters, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), out[0], iters * 8, cudaMemcpyDeviceToHost));
        std::sort(h.begin() + 10, h.end());
        printf("%-44s median %6lld ns  min %6lld  p90 %6lld\n", names[mode], h[10 + (iters - 10) / 2], h[10], h[10 + (iters - 10) * 9 / 10]);
    }
    const char *pn[] = {"flag only", "data st + fence.sys + flag st", "data st + red.release.sys flag"};
    for (int proto = 0; proto < 3; proto++) {
        for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaMemset(buf[d], 0, 4096)); CK(cudaDeviceSynchronize()); }
        const int it = 2000;
        CK(cudaSetDevice(1));
        k_pingpong<<<1, 1>>>(buf[1], buf[0], buf[0] + 256, 1, it, proto, out[1]);
        CK(cudaSetDevice(0));
        k_pingpong<<<1, 1>>>(buf[0], buf[1], buf[1] + 256, 0, it, proto, out[0]);
        CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
        long long ns = 0;
        CK(cudaMemcpy(&ns, out[0], 8, cudaMemcpyDeviceToHost));
        printf("ping-pong %-34s one-way %7.0f ns\n", pn[proto], (double)ns / it / 2);
    }
    return 0;
}
