// q4_kernels.cuh -- hand-written sm_100a kernels for vsim's Q4_0/Q4_1 mul_mat path.
//
// Reference semantics (NAIST-Archlab/vsim): ggml.c:209-298 (activation-row
// quantizers), ggml.c:472-553 / imax.c:1205-1229 (q4 x q4 block dot products),
// ggml.c:4891-5453 (the operator).  Nothing here is translated from those
// loops; the formulation is:
//
//   Y[m][n] = sum_blk d0[n,blk] * d1[m,blk] * ISUM[m,n,blk]                 (Q4_0)
//   ISUM    = sum_{k in blk} (q0-8)(q1-8)                                    (exact, int32)
//   Y[m][n] = sum_blk d0*(d1*P + m1*S0) + m0*Z                              (Q4_1)
//   P = sum q0*q1, S0 = sum q0, Z = d1*sum q1 + 32*m1                       (P, S0 exact)
//
// i.e. the 32-element integer block sums are computed exactly on the tensor
// pipe (mma.sync m16n8k32 s8 x s8 -> s32: one instruction = 16 weight rows x
// one quant block x 8 activation rows) and only the per-block fp32 scale
// products are rounded -- tighter than the reference's own sequential fp32 sum.
//
// Device weight layout ("wtile", built once at upload by repack_kernel):
//   tile  = 16 consecutive weight rows, chunk = 4 consecutive quant blocks.
//   chunk (tile T, group c) lives at  (T*C + c) * CHUNK  bytes, C = ceil(nb/4):
//     [   0, 512)  lane l=(g=l/4,t=l%4): uint4 = word t (nibble bytes 4t..4t+3) of
//                  blocks 4c+0..3 of row 16T+g      -> one coalesced 512-B warp load
//     [ 512,1024)  same for row 16T+8+g
//     [1024,1280)  float d0[16 rows][4 blocks]
//     [1280,1536)  float m0[16 rows][4 blocks]                    (Q4_1 only)
//   1280 B / 1536 B per 16x4 blocks = exactly the reference's 20 / 24 B per block:
//   algorithmic bytes are unchanged, but every warp streams one contiguous
//   region with 128-bit loads and the registers it receives ARE the mma A
//   fragments (after two LOP3 per word to split nibbles).
//
// Activation pack ("apack"; built in shared memory by the GEMV prologue for M <= 2, by
// pack_act_kernel otherwise), MP = padded row count (2 or 8):
//   aq[blk][MP][32 B]: the 32 codes of block blk of row m as bytes (s8 = q-8 for
//                      Q4_0, u8 = q for Q4_1), ordered so that the 8 bytes at +8t
//                      are the mma B fragment of lane (g, t) for that row;
//   ap planes [blk][MP] floats: d1 (Q4_0); d1, m1, Z = d1*sum(q1)+32*m1 (Q4_1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

namespace b200q4 {

constexpr int QK = 32;
constexpr int TILE_ROWS = 16;
constexpr int CHUNK_BLKS = 4;
constexpr int CHUNK_Q4_0 = 1280;
constexpr int CHUNK_Q4_1 = 1536;
constexpr int GEMV_WARPS = 16;
constexpr int GEMV_THREADS = GEMV_WARPS * 32;

__host__ __device__ constexpr int chunk_bytes(int qt) { return qt == 0 ? CHUNK_Q4_0 : CHUNK_Q4_1; }
__host__ __device__ constexpr int ap_floats(int qt) { return qt == 0 ? 2 : 4; }

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const void *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// D(16x8,s32) = A(16x32,u8) * B(32x8,{s8|u8}) + C
template <bool B_SIGNED>
__device__ __forceinline__ void mma_16x8x32(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                            int c0, int c1, int c2, int c3)
{
    if (B_SIGNED)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c0), "r"(c1), "r"(c2), "r"(c3));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c0), "r"(c1), "r"(c2), "r"(c3));
}

// The reference's `(int8_t) round(v)` / `(uint8_t) round(v)` (ggml.c:239-240, 287-288):
// half-away-from-zero, then a C conversion that on the reference's x86-64 build
// yields INT_MIN for +-inf/NaN (cvttsd2si); CUDA's cvt would saturate instead.
__device__ __forceinline__ int round_to_int_ref(float v)
{
    const float r = roundf(v);
    return (fabsf(r) < 2147483648.0f) ? (int)r : INT_MIN;
}

// byte position of element e (0..31) of a block inside its 32-byte apack record:
// word 2t+(e&1) holds elements 8t+{0,2,4,6} (+1 for the odd word), t = e/8, so that
// the low-nibble plane of weight word t meets k-slots 4t..4t+3 and the high-nibble
// plane k-slots 16+4t..16+4t+3 of the m16n8k32 fragments.
__device__ __forceinline__ int apack_byte_of(int e)
{
    const int t = e >> 3, r = e & 7;
    return ((2 * t + (r & 1)) << 2) + (r >> 1);
}

// ---------------------------------------------------------------------------
// Activation-row quantization: one warp per (row m, block blk), lane = element.
// Produces the reference's codes bit-exactly (A.1/A.2 of SURVEY.md):
//   every fp32 op is an individually rounded IEEE op (__fdiv_rn/__fmul_rn/__fsub_rn,
//   no FMA contraction, no FTZ), roundf is half-away-from-zero.
// ---------------------------------------------------------------------------
struct BlockQ {
    int   code;      // this lane's 4-bit code (0..15)
    float d;         // block scale
    float mn;        // block min (Q4_1)
};

template <int QT>
__device__ __forceinline__ BlockQ quantize_block(float x)
{
    BlockQ q;
    if (QT == 0) {
        float amax = fabsf(x);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float d  = __fdiv_rn(amax, 7.0f);
        const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
        const float v  = __fmul_rn(x, id);
        q.code = (int)(uint8_t)((int8_t)round_to_int_ref(v) + 8);
        q.d = d;
        q.mn = 0.0f;
    } else {
        float mn = x, mx = x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        // the reference scans left to right with strict compares (ggml.c:268-274): of
        // several numerically equal extrema (+0.0 / -0.0) it keeps the FIRST one's bits.
        const unsigned first_mn = __ffs(__ballot_sync(0xffffffffu, x == mn)) - 1;
        const unsigned first_mx = __ffs(__ballot_sync(0xffffffffu, x == mx)) - 1;
        mn = __shfl_sync(0xffffffffu, x, first_mn);
        mx = __shfl_sync(0xffffffffu, x, first_mx);
        const float d  = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
        const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
        const float v  = __fmul_rn(__fsub_rn(x, mn), id);
        q.code = (int)(uint8_t)round_to_int_ref(v);
        q.d = d;
        q.mn = mn;
    }
    return q;
}

// Writes one block's apack record: 32 code bytes at aq_rec (even elements -- the low-nibble plane --
// pre-multiplied by 16, see "Fragment helpers") and the block parameters into the parameter planes:
//   Q4_0: plane 0 = d1, plane 1 = -128*sum(a) (int bits);   Q4_1: d1/16, m1/16, Z = d1*sum(q1)+32*m1.
// ap_rec points at this (block,row)'s slot of plane 0; `plane` is the plane stride in floats.
template <int QT>
__device__ __forceinline__ void emit_apack(const BlockQ &q, int lane, uint8_t *aq_rec, float *ap_rec, int plane)
{
    const int a = (QT == 0) ? (q.code & 15) - 8 : (q.code & 15);
    aq_rec[apack_byte_of(lane)] = (uint8_t)((lane & 1) ? a : a * 16);
    int s = a;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (QT == 0) {
            ap_rec[0] = q.d;
            ap_rec[plane] = __int_as_float(-128 * s);
        } else {
            ap_rec[0] = q.d * 0.0625f;
            ap_rec[plane] = q.mn * 0.0625f;
            ap_rec[2 * plane] = fmaf(q.d, (float)s, 32.0f * q.mn);     // Z = sum_k dequant(a_k)
        }
    }
}

// Quad-parallel quantizer used by the fused GEMV prologue: a lane owns 8 consecutive elements
// (element 8*tq .. 8*tq+7 of a block, tq = lane & 3), the 4 lanes of a quad own one block.  Same
// arithmetic as quantize_block (bit-identical codes and scales), 8x fewer steps per block; its 8
// codes are exactly words 2tq / 2tq+1 of the apack record, written with one 8-byte store.
// All 32 lanes must call it (shuffles); `valid` masks the stores.
template <int QT>
__device__ __forceinline__ void quantize8_emit(float4 lo, float4 hi, bool valid, int lane, uint8_t *aq_rec, float *ap_rec, int plane)
{
    const int tq = lane & 3, qbase = lane & 28;
    float v[8];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    int a[8];
    float d, mn = 0.0f;
    if (QT == 0) {
        float amax = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        d = __fdiv_rn(amax, 7.0f);
        const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = (int)(int8_t)round_to_int_ref(__fmul_rn(v[i], id));   // = code - 8
    } else {
        // strict left-to-right scans (ggml.c:268-274): first of several equal extrema wins
        float lmn = v[0], lmx = v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) { if (v[i] < lmn) lmn = v[i]; if (v[i] > lmx) lmx = v[i]; }
        mn = __shfl_sync(0xffffffffu, lmn, qbase);
        float mx = __shfl_sync(0xffffffffu, lmx, qbase);
#pragma unroll
        for (int j = 1; j < 4; j++) {
            const float omn = __shfl_sync(0xffffffffu, lmn, qbase + j), omx = __shfl_sync(0xffffffffu, lmx, qbase + j);
            if (omn < mn) mn = omn;
            if (omx > mx) mx = omx;
        }
        d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
        const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = (int)(uint8_t)round_to_int_ref(__fmul_rn(__fsub_rn(v[i], mn), id));
    }
    const uint32_t even = (uint32_t)((a[0] * 16) & 255) | ((uint32_t)((a[2] * 16) & 255) << 8) |
                          ((uint32_t)((a[4] * 16) & 255) << 16) | ((uint32_t)((a[6] * 16) & 255) << 24);
    const uint32_t odd  = (uint32_t)(a[1] & 255) | ((uint32_t)(a[3] & 255) << 8) | ((uint32_t)(a[5] & 255) << 16) | ((uint32_t)(a[7] & 255) << 24);
    if (valid) *reinterpret_cast<uint2 *>(aq_rec + tq * 8) = make_uint2(even, odd);
    int sum = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7];
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (valid && tq == 0) {
        if (QT == 0) {
            ap_rec[0] = d;
            ap_rec[plane] = __int_as_float(-128 * sum);
        } else {
            ap_rec[0] = d * 0.0625f;
            ap_rec[plane] = mn * 0.0625f;
            ap_rec[2 * plane] = fmaf(d, (float)sum, 32.0f * mn);
        }
    }
}

// fp32 rows -> apack [nbp][MP][32] + planes.  grid: nbp*MP warps.  Rows >= M and blocks >= nb
// (padding up to a multiple of 4) are zero records.
template <int QT>
__global__ void pack_act_kernel(const float *__restrict__ x, size_t x_stride, int M, int MP, int nb, int nbp,
                                uint8_t *__restrict__ aq, float *__restrict__ ap)
{
    // Programmatic dependent launch on both sides: the decode kernel that follows may start (and stream
    // weights) at once; this kernel itself waits for the previous one, which still reads aq/ap and may write X.
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long)nbp * MP) return;
    const int blk = (int)(w / MP), m = (int)(w - (long)blk * MP);
    const bool real = (blk < nb) && (m < M);
    const float xv = real ? x[(size_t)m * x_stride + (size_t)blk * QK + lane] : 0.0f;
    BlockQ q = quantize_block<QT>(xv);
    if (!real) { q.d = 0.0f; q.mn = 0.0f; q.code = (QT == 0) ? 8 : 0; }
    emit_apack<QT>(q, lane, aq + ((size_t)blk * MP + m) * 32, ap + (size_t)blk * MP + m, nbp * MP);
}

// reference-format quantized rows (params->wdata image) -> apack  (unpatched-ggml hook path)
template <int QT>
__global__ void pack_wdata_kernel(const uint8_t *__restrict__ wdata, int M, int MP, int nb, int nbp,
                                  uint8_t *__restrict__ aq, float *__restrict__ ap)
{
    cudaTriggerProgrammaticLaunchCompletion();           // see pack_act_kernel
    cudaGridDependencySynchronize();
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long)nbp * MP) return;
    const int blk = (int)(w / MP), m = (int)(w - (long)blk * MP);
    BlockQ q;
    q.d = 0.0f; q.mn = 0.0f; q.code = (QT == 0) ? 8 : 0;
    if (blk < nb && m < M) {
        if (QT == 0) {
            const uint8_t *b = wdata + ((size_t)m * nb + blk) * 20;
            q.d = *reinterpret_cast<const float *>(b);
            const uint8_t v = b[4 + (lane >> 1)];
            q.code = (lane & 1) ? (v >> 4) : (v & 15);
        } else {
            const uint8_t *row = wdata + (size_t)m * nb * 24;
            q.mn = reinterpret_cast<const float *>(row)[blk];
            q.d  = reinterpret_cast<const float *>(row)[nb + blk];
            const uint8_t v = row[(size_t)nb * 8 + (size_t)blk * 16 + (lane >> 1)];
            q.code = (lane & 1) ? (v >> 4) : (v & 15);
        }
    }
    emit_apack<QT>(q, lane, aq + ((size_t)blk * MP + m) * 32, ap + (size_t)blk * MP + m, nbp * MP);
}

// fp32 rows -> the reference's quantized row image, byte for byte
// (quantize_row_q4_0 ggml.c:209-251, quantize_row_q4_1 ggml.c:256-298).
template <int QT>
__global__ void quantize_rows_ref_kernel(const float *__restrict__ x, size_t x_stride, int M, int nb,
                                         uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long)nb * M) return;
    const int m = (int)(w / nb), blk = (int)(w - (long)m * nb);
    const BlockQ q = quantize_block<QT>(x[(size_t)m * x_stride + (size_t)blk * QK + lane]);
    // 8 lanes -> one 32-bit word of nibbles: element e sits at bit 4*(e%8)
    uint32_t word = (uint32_t)(q.code & 15) << (4 * (lane & 7));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if (QT == 0) {
        uint8_t *b = out + ((size_t)m * nb + blk) * 20;
        if (lane == 0) *reinterpret_cast<float *>(b) = q.d;
        if ((lane & 7) == 0) *reinterpret_cast<uint32_t *>(b + 4 + (lane >> 3) * 4) = word;
    } else {
        uint8_t *row = out + (size_t)m * nb * 24;
        if (lane == 0) {
            reinterpret_cast<float *>(row)[blk] = q.mn;
            reinterpret_cast<float *>(row)[nb + blk] = q.d;
        }
        if ((lane & 7) == 0)
            *reinterpret_cast<uint32_t *>(row + (size_t)nb * 8 + (size_t)blk * 16 + (lane >> 3) * 4) = word;
    }
}

// ---------------------------------------------------------------------------
// Offline WEIGHT quantizers: utils.cpp:424-481 ggml_quantize_q4_0 and
// utils.cpp:483-541 ggml_quantize_q4_1, what quantize_*.cpp runs over every 2-D tensor of a model.  Same row
// formats and rounding as the activation quantizers; q4_1 differs in ONE place that has to be kept to stay
// byte-identical: the running maximum starts at std::numeric_limits<float>::min() (the smallest positive
// normal float, utils.cpp:503), so a block with nothing above 1.17549435e-38 gets that as its max.
// Also fills the 16-bin histogram of emitted codes (hist, 64-bit counters).  One warp per (row, block).
// ---------------------------------------------------------------------------
template <int QT>
__global__ void quantize_weights_kernel(const float *__restrict__ x, size_t x_stride, long n_rows, int nb,
                                        uint8_t *__restrict__ out, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int s_hist[16];
    if (threadIdx.x < 16) s_hist[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w < n_rows * nb) {                               // (whole warps: w is warp-uniform)
        const long m = w / nb;
        const int blk = (int)(w - m * nb);
        const float xv = x[(size_t)m * x_stride + (size_t)blk * QK + lane];
        BlockQ q;
        if (QT == 0) {
            q = quantize_block<0>(xv);
        } else {
            float mn = xv, mx = xv;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            // strict left-to-right compares: of equal minima (+0.0 / -0.0) the FIRST one's bits are kept
            const unsigned first_mn = __ffs(__ballot_sync(0xffffffffu, xv == mn)) - 1;
            mn = __shfl_sync(0xffffffffu, xv, first_mn);
            mx = (mx > 1.17549435e-38f) ? mx : 1.17549435e-38f;      // numeric_limits<float>::min() start value
            const float d  = __fdiv_rn(__fsub_rn(mx, mn), 15.0f);
            const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
            q.code = (int)(uint8_t)round_to_int_ref(__fmul_rn(__fsub_rn(xv, mn), id));
            q.d = d;
            q.mn = mn;
        }
        atomicAdd(&s_hist[q.code & 15], 1u);
        uint32_t word = (uint32_t)(q.code & 15) << (4 * (lane & 7));
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        if (QT == 0) {
            uint8_t *b = out + ((size_t)m * nb + blk) * 20;
            if (lane == 0) *reinterpret_cast<float *>(b) = q.d;
            if ((lane & 7) == 0) *reinterpret_cast<uint32_t *>(b + 4 + (lane >> 3) * 4) = word;
        } else {
            uint8_t *row = out + (size_t)m * nb * 24;
            if (lane == 0) {
                reinterpret_cast<float *>(row)[blk] = q.mn;
                reinterpret_cast<float *>(row)[nb + blk] = q.d;
            }
            if ((lane & 7) == 0)
                *reinterpret_cast<uint32_t *>(row + (size_t)nb * 8 + (size_t)blk * 16 + (lane >> 3) * 4) = word;
        }
    }
    __syncthreads();
    if (threadIdx.x < 16 && s_hist[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)s_hist[threadIdx.x]);
}

// ---------------------------------------------------------------------------
// Weight re-tiling: reference row format (tightly packed rows on the device) -> wtile.
// One thread per 16-byte output slot.  Rows >= N and blocks >= nb are padding
// (scale 0, nibbles 8 / 0) so they contribute exactly 0.
// ---------------------------------------------------------------------------
template <int QT>
__global__ void repack_kernel(const uint8_t *__restrict__ rows, int N, int nb, int n_tiles, int C,
                              uint8_t *__restrict__ wt)
{
    constexpr int CH = chunk_bytes(QT);
    constexpr int SLOTS = CH / 16;
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)n_tiles * C * SLOTS;
    if (id >= total) return;
    const long chunk = id / SLOTS;
    const int slot = (int)(id - chunk * SLOTS);
    const int T = (int)(chunk / C), c = (int)(chunk - (long)T * C);
    const size_t rb = (size_t)nb * (QT == 0 ? 20 : 24);
    uint4 outv;
    if (slot < 64) {
        const int half = slot >> 5, lane = slot & 31, g = lane >> 2, t = lane & 3;
        const int row = T * TILE_ROWS + half * 8 + g;
        uint32_t wv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int blk = c * CHUNK_BLKS + i;
            uint32_t v = (QT == 0) ? 0x88888888u : 0u;
            if (row < N && blk < nb) {
                const uint8_t *src = (QT == 0) ? rows + (size_t)row * rb + (size_t)blk * 20 + 4 + 4 * t
                                               : rows + (size_t)row * rb + (size_t)nb * 8 + (size_t)blk * 16 + 4 * t;
                v = *reinterpret_cast<const uint32_t *>(src);
            }
            wv[i] = v;
        }
        outv = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    } else {
        const int s = slot - 64;            // 0..15 scales, 16..31 mins (Q4_1)
        const int r = s & 15;
        const bool is_min = s >= 16;
        const int row = T * TILE_ROWS + r;
        float fv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int blk = c * CHUNK_BLKS + i;
            float v = 0.0f;
            if (row < N && blk < nb) {
                if (QT == 0) v = *reinterpret_cast<const float *>(rows + (size_t)row * rb + (size_t)blk * 20);
                else {
                    const float *rf = reinterpret_cast<const float *>(rows + (size_t)row * rb);
                    v = is_min ? rf[blk] : rf[nb + blk];
                }
            }
            fv[i] = v;
        }
        outv = make_uint4(__float_as_uint(fv[0]), __float_as_uint(fv[1]), __float_as_uint(fv[2]), __float_as_uint(fv[3]));
    }
    *reinterpret_cast<uint4 *>(wt + chunk * CH + (size_t)slot * 16) = outv;
}

// wtile -> reference row format (b200q4_weight_download; round-trip check of the re-tiling)
template <int QT>
__global__ void unpack_kernel(const uint8_t *__restrict__ wt, int N, int nb, int C, uint8_t *__restrict__ rows)
{
    constexpr int CH = chunk_bytes(QT);
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (row, blk)
    if (id >= (long)N * nb) return;
    const int row = (int)(id / nb), blk = (int)(id - (long)row * nb);
    const int T = row / TILE_ROWS, r = row % TILE_ROWS, half = r >> 3, g = r & 7;
    const int c = blk / CHUNK_BLKS, i = blk % CHUNK_BLKS;
    const uint8_t *ch = wt + ((size_t)T * C + c) * CH;
    const size_t rb = (size_t)nb * (QT == 0 ? 20 : 24);
    uint32_t wv[4];
#pragma unroll
    for (int t = 0; t < 4; t++)
        wv[t] = reinterpret_cast<const uint32_t *>(ch + half * 512 + (g * 4 + t) * 16)[i];
    const float d = reinterpret_cast<const float *>(ch + 1024)[r * 4 + i];
    if (QT == 0) {
        uint8_t *dst = rows + (size_t)row * rb + (size_t)blk * 20;
        *reinterpret_cast<float *>(dst) = d;
#pragma unroll
        for (int t = 0; t < 4; t++) *reinterpret_cast<uint32_t *>(dst + 4 + 4 * t) = wv[t];
    } else {
        float *rf = reinterpret_cast<float *>(rows + (size_t)row * rb);
        rf[blk] = reinterpret_cast<const float *>(ch + 1280)[r * 4 + i];
        rf[nb + blk] = d;
        uint8_t *dst = rows + (size_t)row * rb + (size_t)nb * 8 + (size_t)blk * 16;
#pragma unroll
        for (int t = 0; t < 4; t++) *reinterpret_cast<uint32_t *>(dst + 4 * t) = wv[t];
    }
}

// ---------------------------------------------------------------------------
// Fragment helpers shared by the decode kernels.
//
// Nibbles become the u8 A fragment with ONE LOP3 per plane:
//     lo = w & 0x0F0F0F0F  (= q),      hi = w & 0xF0F0F0F0  (= 16 q)
// and the activation pack stores the even-element (lo-plane) codes pre-multiplied by 16, so both
// planes contribute 16*q0*a.  For Q4_0 (a = q1-8, s8) the -8 of the weight code is restored by
// starting the IMMA accumulator at -128*sum(a) (apack plane 1); the common factor 16 is removed
// exactly (x 0.0625) when Y is stored.  For Q4_1 the planes d1 and m1 are stored pre-divided by 16.
// ---------------------------------------------------------------------------
template <int QT> struct Frag;
// sa/sb: lane (g,t) keeps d0 of rows g / g+8 for block 4c+t only.
template <> struct Frag<0> { uint4 qa, qb; float sa, sb; };
template <> struct Frag<1> { uint4 qa, qb; float sa, sb, ma, mb; };

__device__ __forceinline__ uint32_t u4_get(const uint4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

__device__ __forceinline__ void make_a(uint32_t (&a)[4], uint32_t wa, uint32_t wb)
{
    a[0] = wa & 0x0F0F0F0Fu;
    a[1] = wb & 0x0F0F0F0Fu;
    a[2] = wa & 0xF0F0F0F0u;
    a[3] = wb & 0xF0F0F0F0u;
}

// D = A*B + C; A u8, B s8 (Q4_0) or u8 (Q4_1)
template <int QT>
__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const int (&c)[4])
{
    if constexpr (QT == 0)
        asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
    else
        asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

// ---- MP == 8: one IMMA per block, D column pair (2t, 2t+1) = activation rows; the scale of
// block i comes from lane (g,i) of the quad by shuffle.  apack [blk][8][32], planes [blk][8].
template <int QT>
__device__ __forceinline__ void compute_unit_m8(const Frag<QT> &f, const uint8_t *chunk, int c, const uint8_t *s_aq, const float *s_ap,
                                                int plane, int lane, float (&acc)[4])
{
    const int t = lane & 3, g = lane >> 2;
    // the four block scales of rows g / g+8: one 16-byte load each (the quad reads the same address)
    const float4 d0a4 = *reinterpret_cast<const float4 *>(chunk + 1024 + g * 16), d0b4 = *reinterpret_cast<const float4 *>(chunk + 1024 + 128 + g * 16);
    const float d0av[4] = {d0a4.x, d0a4.y, d0a4.z, d0a4.w}, d0bv[4] = {d0b4.x, d0b4.y, d0b4.z, d0b4.w};
    float m0av[4] = {0.f, 0.f, 0.f, 0.f}, m0bv[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (QT == 1) {
        const float4 a4 = *reinterpret_cast<const float4 *>(chunk + 1280 + g * 16), b4 = *reinterpret_cast<const float4 *>(chunk + 1280 + 128 + g * 16);
        m0av[0] = a4.x; m0av[1] = a4.y; m0av[2] = a4.z; m0av[3] = a4.w;
        m0bv[0] = b4.x; m0bv[1] = b4.y; m0bv[2] = b4.z; m0bv[3] = b4.w;
    }
#pragma unroll
    for (int i = 0; i < CHUNK_BLKS; i++) {
        uint32_t a[4];
        make_a(a, u4_get(f.qa, i), u4_get(f.qb, i));
        const int blk = c * 4 + i;
        const uint2 b = *reinterpret_cast<const uint2 *>(s_aq + blk * 256 + lane * 8);
        const float2 d1 = *reinterpret_cast<const float2 *>(s_ap + blk * 8 + 2 * t);
        const float d0a = d0av[i], d0b = d0bv[i];
        // int32 -> fp32 without the conversion pipe: the IMMA accumulator starts at the bit pattern of
        // 1.5 * 2^23, so the result register IS the float 12582912 + sum (|sum| < 2^18, exact); one FADD removes it.
        constexpr int MAGIC_I = 0x4B400000;
        constexpr float MAGIC_F = 12582912.0f;
        int d[4];
        if constexpr (QT == 0) {
            const int2 cr = *reinterpret_cast<const int2 *>(s_ap + plane + blk * 8 + 2 * t);
            const int ci[4] = {cr.x + MAGIC_I, cr.y + MAGIC_I, cr.x + MAGIC_I, cr.y + MAGIC_I};
            imma<QT>(d, a, b.x, b.y, ci);
            acc[0] = fmaf(d0a * d1.x, __int_as_float(d[0]) - MAGIC_F, acc[0]);
            acc[1] = fmaf(d0a * d1.y, __int_as_float(d[1]) - MAGIC_F, acc[1]);
            acc[2] = fmaf(d0b * d1.x, __int_as_float(d[2]) - MAGIC_F, acc[2]);
            acc[3] = fmaf(d0b * d1.y, __int_as_float(d[3]) - MAGIC_F, acc[3]);
        } else {
            const int magic[4] = {MAGIC_I, MAGIC_I, MAGIC_I, MAGIC_I};
            int s[4];
            imma<QT>(d, a, b.x, b.y, magic);
            imma<QT>(s, a, 0x10101010u, 0x01010101u, magic);     // 16 * sum of the block's weight codes
            const float2 m1 = *reinterpret_cast<const float2 *>(s_ap + plane + blk * 8 + 2 * t);
            const float2 z  = *reinterpret_cast<const float2 *>(s_ap + 2 * plane + blk * 8 + 2 * t);
            const float m0a = m0av[i], m0b = m0bv[i];
            const float sa0 = __int_as_float(s[0]) - MAGIC_F, sb0 = __int_as_float(s[2]) - MAGIC_F;
            // d0*(d1*P + m1*S0) + m0*Z   (d1, m1 stored /16; P, S0 computed x16)
            acc[0] += fmaf(d0a, fmaf(d1.x, __int_as_float(d[0]) - MAGIC_F, m1.x * sa0), m0a * z.x);
            acc[1] += fmaf(d0a, fmaf(d1.y, __int_as_float(d[1]) - MAGIC_F, m1.y * sa0), m0a * z.y);
            acc[2] += fmaf(d0b, fmaf(d1.x, __int_as_float(d[2]) - MAGIC_F, m1.x * sb0), m0b * z.x);
            acc[3] += fmaf(d0b, fmaf(d1.y, __int_as_float(d[3]) - MAGIC_F, m1.y * sb0), m0b * z.y);
        }
    }
}

// ---------------------------------------------------------------------------
// The decode kernel, pipelined form (the one b200q4_mul_mat uses):
//
//   producer warp   walks "tasks" (<= 64 consecutive units inside one 16-row tile, or a few whole
//                   tiles when a tile is shorter than 16 units) handed out by an atomic ticket
//                   counter -- dynamic scheduling: a slow SM simply takes fewer tasks -- and
//                   streams them into a ring of shared-memory stages, one cp.async.bulk (TMA 1-D
//                   bulk copy, 32 units = 40/48 KB) per stage, completion on an mbarrier.
//   16 consumer     warp w takes units 2w, 2w+1 of the stage: LDS for the A-fragment words and
//   warps           scales, one LOP3 per nibble plane, IMMA against the B fragments of the activation
//                   pack, scale FMAs, a quad reduction and one store of its partial sums.
//                   Stateless between stages.
//                   M = 1 ("octet"): the two units are 8 consecutive quant blocks; block j feeds
//                   MMA column j, so all eight IMMAs accumulate into one D fragment and lane (g,t)
//                   ends up with the integer sums of blocks 2t, 2t+1 for rows g, g+8.
//                   M = 2 ("chain4"): per unit, (block i, row m) -> column 2i+m.
//                   M <= 8 ("m8"):    per block one IMMA, columns = activation rows.
//   reducer warp    adds the per-unit partials of a stage in unit order (fixed -> bit-reproducible),
//                   carries the tile accumulator across stages, and at the end of a tile stores Y.
//
// Long rows (K > 4096, i.e. more than 32 units per tile) are K-SLICED: a tile is cut into PP pieces of
// TS <= 32 units and CTA b is bound to piece class b % PP -- it only ever needs the activations of its
// own K-slice, so the prologue quantizes (and shared memory holds) at most 128 quant blocks whatever K
// is, and the stage ring keeps its full depth.  Each class draws tiles from its own ticket counter.
// A finished piece is parked in a workspace slot with plain stores (no fence, no atomic: nothing sits
// between two stages of the stream, and the launch ends like an unsliced one); pipe_finalize_kernel,
// chained behind it with programmatic dependent launch, adds the PP slots of every tile in piece order
// (reproducible rounding) and stores Y while the NEXT decode launch is already filling its stage ring.
//
// Bytes in flight per SM = (stages-1) x 40 KB, independent of registers; weights never touch the
// register file before they are needed.  Programmatic dependent launch: the weight stream of launch
// i+1 starts while launch i drains; everything that depends on the previous kernel (X, Y, workspace,
// ticket counter) is touched only after cudaGridDependencySynchronize().
// ---------------------------------------------------------------------------
// NCW consumer warps per CTA (template parameter): 16 with one CTA per SM, or 8 with two CTAs per SM
// (half-size stages and ring).  Two co-resident CTAs let the next launch's CTA start streaming on
// an SM while the other CTA of the previous launch is still in its tail, so HBM never idles
// between back-to-back mul_mats.
constexpr int PIPE_TILE_SPLIT = 16;                     // tiles of >= this many units are scheduled as tile pieces
constexpr int PIPE_MAX_STAGES = 6;
constexpr int PIPE_BAR_DONE = 2;                        // named barriers 2 .. 2 + PIPE_MAX_STAGES - 1: consumers -> reducer, one per ring slot
constexpr int PIPE_STATIC_TASKS = 3;                    // per CTA, before the first ticket is drawn

constexpr int MAX_PEERS = 8;
// Fused all-gather state.  Sequencing lives on the device so that a launch (or a captured CUDA graph of
// launches) needs no per-call host values: *seq_ctr counts this rank's completed exchanges; exchange
// number s = *seq_ctr + 1 stores into Y buffer s & 1 (two buffers alternate: a fast rank may already be
// storing exchange s+1 into a peer that still reads exchange s as its next input).
// Arrival counters: slot `src` of rank r's flag array counts what rank `src` has delivered to r.  Every CTA,
// after ordering its own peer stores (fence.sys = one NVLink round trip), adds its share straight into the
// slot on every rank -- one more one-way trip; no "last CTA" hand-off, no second fence.  The shares of one
// exchange always sum to PEER_UNIT (CTA 0 adds the remainder), so exchange s has fully landed from `src`
// exactly when the slot reaches s * PEER_UNIT, whatever the grid size of the launches involved.
// A launch first waits until all of exchange *seq_ctr has landed locally -- its X is normally that exchange's Y.
constexpr uint32_t PEER_UNIT = 4096;                     // > any grid size used here
struct PeerOut {
    float    *y[2][MAX_PEERS];   // [buffer][rank]: rank r's full Y [M][y_stride] (y[.][rank] is local memory)
    uint32_t *flag[MAX_PEERS];   // rank r's arrival counters [world]
    const uint32_t *local_flags; // this rank's counters (= flag[rank])
    uint32_t *seq_ctr;           // local exchange counter
    int      *done;              // local CTA-completion counter (self-resetting); NULL: do not signal
    int       world, rank, col0; // col0: first output column of this rank's shard
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until rank r's arrival counter shows exchange `seq` complete (call with r < world)
__device__ __forceinline__ void peer_wait_one(const uint32_t *flags, int r, uint32_t seq)
{
    const uint32_t exp = seq * PEER_UNIT;
    for (uint32_t spins = 0;; spins++) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
        if ((int32_t)(v - exp) >= 0) break;
        // A peer that never arrives (a rank that died) must not hang this GPU for good: after roughly a minute of
        // polling the kernel traps, the stream reports an error and the process ends instead of spinning forever.
        if (spins > (1u << 26)) __trap();
        __nanosleep(20);
    }
}
// one output element: locally, or into every rank's Y (buffer b) at this shard's column offset
__device__ __forceinline__ void store_y(float *y, const PeerOut &po, int b, size_t idx, float v)
{
    if (po.world <= 1) { y[idx] = v; return; }
#pragma unroll 1
    for (int r = 0; r < po.world; r++) po.y[b][r][idx + po.col0] = v;
}
// this CTA's share of the exchange, added by thread r (< world) into rank r's counter for this source rank
__device__ __forceinline__ void peer_add_share(const PeerOut &po, int r)
{
    const uint32_t share = blockIdx.x == 0 ? PEER_UNIT - (gridDim.x - 1) : 1u;
    asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" :: "l"(po.flag[r] + po.rank), "r"(share) : "memory");
}
// local bookkeeping (off the critical path): the last CTA of the launch advances the exchange counter
__device__ __forceinline__ void peer_count_done(const PeerOut &po)
{
    if (atomicAdd(po.done, 1) == (int)gridDim.x - 1) {
        *po.done = 0;
        *po.seq_ctr = ld_volatile_u32(po.seq_ctr) + 1;
    }
}
// after a warp's last peer store (only the reducer warp stores Y)
__device__ __forceinline__ void peer_signal(const PeerOut &po, int lane)
{
    if (po.world <= 1 || !po.done) return;               // done == NULL: an earlier launch of a multi-launch call
    __threadfence_system();                              // every lane: its peer stores are performed system-wide ...
    __syncwarp();                                        // ... before any lane's signal
    if (lane < po.world) peer_add_share(po, lane);
    if (lane == 0) peer_count_done(po);
}
// holds the stream until exchange *seq_ctr has landed from every rank (lane r watches rank r)
static __global__ void wait_flags_kernel(const uint32_t *flags, int world, const uint32_t *seq_ctr)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    peer_wait_one(flags, r, ld_volatile_u32(seq_ctr));
}
// Publishes an exchange whose peer stores were issued by the PREVIOUS kernel in the stream (the decode kernel's
// epilogue).  PDL-chained: cudaGridDependencySynchronize returns once that kernel has completed and flushed its
// stores, then one system fence and one add of the whole exchange per rank.  Keeping the NVLink round trip of
// the fence out of the decode kernel lets its CTAs exit -- and the next launch start filling its stage ring --
// 1.7 us earlier (tools/p2pbench).  <<<1, 32>>>
static __global__ void peer_signal_kernel(PeerOut po)
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    __threadfence_system();
    __syncwarp();
    if ((int)threadIdx.x < po.world) peer_add_share(po, threadIdx.x);
    if (threadIdx.x == 0) peer_count_done(po);
}
// unfused variant for launches whose kernel has no peer epilogue (the tile kernel): copies this rank's
// [M][ncols] slice from a local buffer into every rank's Y, then signals like peer_signal
static __global__ void scatter_slice_kernel(const float *__restrict__ src, size_t src_stride, int M, int ncols, size_t y_stride, PeerOut po)
{
    const int b = (int)((ld_volatile_u32(po.seq_ctr) + 1) & 1u);
    const long total = (long)M * ncols;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int m = (int)(i / ncols), n = (int)(i - (long)m * ncols);
        const float v = src[(size_t)m * src_stride + n];
        for (int r = 0; r < po.world; r++) po.y[b][r][(size_t)m * y_stride + po.col0 + n] = v;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < po.world) peer_add_share(po, threadIdx.x);
    if (threadIdx.x == 0) peer_count_done(po);
}

// Geometry of the NEXT decode launch in the stream, when the caller knows it (b200q4_hint_next_weight; the ggml hook
// learns the order of a model's tensors): once a CTA has issued its last copy it asks L2 to prefetch the chunks the
// next launch's CTAs will request first, so the HBM channels keep working through this launch's tail and the next
// launch's first stages are L2 hits.  wt == NULL: no hint.
struct NextHint {
    const uint8_t *wt;
    int C, TS, PP, U, n_tiles, ntasks, grid, chunk, stages;   // stages: tasks per CTA to prefetch
};

struct PipeParams {
    const uint8_t *wt;
    const float   *x;        // fused quantization (M <= 2)
    const uint8_t *aq;       // global apack (non-fused): compact [nbp][MP][32]
    const float   *ap;       //   planes [NP][nbp][MP]
    float         *y;
    float         *partial;  // [ntasks][128] floats, used when a tile spans several tasks
    int           *ticket;   // [0] next task (unsliced), [1] CTAs done, [2 + c] next tile of piece class c   (self-resetting)
    size_t         x_stride, y_stride;   // floats
    int            U, N, M, C, n_tiles, nb;
    int            ntasks;   // tasks in the launch
    int            PP;       // K-slices (pieces) per tile; 1 = unsliced
    int            TS;       // units per task
    int            gnbp;     // quant blocks per row in the global apack (non-fused launches): plane stride = gnbp * rows
    int            nstatic;  // static tasks per CTA (0 = PIPE_STATIC_TASKS)
    int            flags;    // development knob: 1 = the reducer polls its barriers instead of try_wait (measured: no gain)
    int            stages;
    int            rslots;   // > 0 (8-row kernel): the reduction ring has its own `rslots` entries, apart from the weight ring, of one
                             // 16 x 8 partial per consumer WARP (its two units belong to one tile): 8 KB per entry instead of 16 KB per
                             // weight stage -- one more weight stage fits.  0: one entry of 2 slots per warp per weight stage
    unsigned long long *dbg; // optional timeline (B200Q4_TIMELINE builds), else NULL
    // Row-sharded launch with the all-gather fused into the epilogue (world > 1): this rank's output
    // columns are stored straight into every rank's Y over NVLink (peer-mapped pointers), then the last
    // CTA signals `seq` into slot `rank` of every rank's flag array.
    PeerOut        peer;
    NextHint       next;
};

__device__ __forceinline__ void l2_prefetch_bulk(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
// the chunks CTA b of the hinted launch will ask for first (its static tasks, same formulas as the producer below)
__device__ __forceinline__ void prefetch_next_launch(const NextHint &h, int b)
{
    if (!h.wt || b >= h.grid) return;
    const bool sliced = h.PP > 1;
    const int cls = sliced ? b % h.PP : 0;
    const int idx = sliced ? b / h.PP : b;
    const int ncls = sliced ? (h.grid - cls + h.PP - 1) / h.PP : h.grid;
    const int limit = sliced ? h.n_tiles : h.ntasks;
    for (int k = 0; k < h.stages; k++) {
        const int t = idx + k * ncls;
        if (t >= limit) break;
        long u0; int n;
        if (sliced) { u0 = (long)t * h.C + cls * h.TS; n = min(h.TS, h.C - cls * h.TS); }
        else if (h.C >= PIPE_TILE_SPLIT) { u0 = (long)t * h.C; n = h.C; }
        else { u0 = (long)t * h.TS; n = (int)min((long)h.TS, (long)h.U - u0); }
        l2_prefetch_bulk(h.wt + (size_t)u0 * h.chunk, (uint32_t)n * (uint32_t)h.chunk);
    }
}

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Timeline stamps are compiled in only with -DB200Q4_TIMELINE (tools/timeline.py builds that way).
#ifdef B200Q4_TIMELINE
#define PIPE_STAMP(slot) do { if (p.dbg) p.dbg[(size_t)blockIdx.x * 16 + (slot)] = gtimer(); } while (0)
#else
#define PIPE_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ bool mbar_test(uint64_t *b, uint32_t parity);
__device__ __forceinline__ void mbar_spin(uint64_t *b, uint32_t parity)      // busy polling, never suspends
{
    while (!mbar_test(b, parity)) { }
}
__device__ __forceinline__ bool mbar_test(uint64_t *b, uint32_t parity)      // non-blocking
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// The same with an L2 eviction-priority hint: a weight tensor is read once per token and is far larger than what else a
// decode step touches (activations, layer-norm weights, biases, the gelu / exp tables, the KV rows) -- streamed
// evict-first it does not push those out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *b, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)), "l"(policy) : "memory");
}

// First unit u0, unit count n, tile T and chunk index c0 of the first unit of task t.
// No integer division on the common path (the producer is one thread; a division costs ~150
// cycles of pure latency there).
__device__ __forceinline__ void task_range(const PipeParams &p, int t, int &u0, int &n, int &T, int &c0)
{
    if (p.C >= PIPE_TILE_SPLIT) {                        // one whole tile (unsliced: C <= 32)
        T = t;
        c0 = 0;
        u0 = T * p.C;
        n = p.C;
    } else {                                            // short tiles: a task = TS/C whole tiles
        u0 = t * p.TS;
        n = min(p.TS, p.U - u0);
        T = t * (p.TS / p.C);
        c0 = 0;
    }
}

template <int QT>
__device__ __forceinline__ void load_frag_smem(Frag<QT> &f, const uint8_t *chunk, int lane)
{
    f.qa = *reinterpret_cast<const uint4 *>(chunk + lane * 16);
    f.qb = *reinterpret_cast<const uint4 *>(chunk + 512 + lane * 16);
    f.sa = *reinterpret_cast<const float *>(chunk + 1024 + lane * 4);
    f.sb = *reinterpret_cast<const float *>(chunk + 1024 + 128 + lane * 4);
    if constexpr (QT == 1) {
        f.ma = *reinterpret_cast<const float *>(chunk + 1280 + lane * 4);
        f.mb = *reinterpret_cast<const float *>(chunk + 1280 + 128 + lane * 4);
    }
}

// ---- M == 2 ("chain4"): the 8 columns of the m16n8k32 tile are (block 4c+i, row m) -> column 2i+m.
// The four IMMAs of a unit accumulate into ONE D fragment and lane (g,t) ends up with the integer
// sums of block 4c+t for rows g/g+8 -- exactly the block whose scale d0[g][t] it loaded.
// B8: codes in the 8-column form [blk][8][32 B] (block 4c+i only in columns 2i, 2i+1): one 8-byte
// load per IMMA.  Otherwise the compact form [blk][2][32 B], read once and masked per block.
template <int QT, bool B8>
__device__ __forceinline__ void compute_unit_chain4c(const Frag<QT> &f, int c, const uint8_t *s_aq, const float *s_ap,
                                                     int plane, int lane, float (&acc)[4])
{
    const int g = lane >> 2, t = lane & 3;
    const uint8_t *bp = s_aq + (B8 ? c * 1024 : c * 256) + lane * 8;
    uint2 breg = make_uint2(0u, 0u);
    if constexpr (!B8) breg = *reinterpret_cast<const uint2 *>(bp);
    const float2 d1 = *reinterpret_cast<const float2 *>(s_ap + (c * 4 + t) * 2);
    int d[4] = {0, 0, 0, 0};
    int s[4] = {0, 0, 0, 0};
    if constexpr (QT == 0) {
        const int2 cr = *reinterpret_cast<const int2 *>(s_ap + plane + (c * 4 + t) * 2);
        d[0] = cr.x; d[1] = cr.y; d[2] = cr.x; d[3] = cr.y;
    }
#pragma unroll
    for (int i = 0; i < CHUNK_BLKS; i++) {
        uint32_t a[4];
        make_a(a, u4_get(f.qa, i), u4_get(f.qb, i));
        if constexpr (B8) {
            const uint2 b = *reinterpret_cast<const uint2 *>(bp + i * 256);
            imma<QT>(d, a, b.x, b.y, d);
        } else {
            const bool mine = (g >> 1) == i;
            imma<QT>(d, a, mine ? breg.x : 0u, mine ? breg.y : 0u, d);
        }
        if constexpr (QT == 1) {
            const bool col = g == 2 * i;                 // ones in column 2i only: 16 * S0 of block i
            imma<QT>(s, a, col ? 0x10101010u : 0u, col ? 0x01010101u : 0u, s);
        }
    }
    if constexpr (QT == 0) {
        acc[0] = (f.sa * d1.x) * (float)d[0];
        acc[1] = (f.sa * d1.y) * (float)d[1];
        acc[2] = (f.sb * d1.x) * (float)d[2];
        acc[3] = (f.sb * d1.y) * (float)d[3];
    } else {
        const float2 m1 = *reinterpret_cast<const float2 *>(s_ap + plane + (c * 4 + t) * 2);
        const float2 z  = *reinterpret_cast<const float2 *>(s_ap + 2 * plane + (c * 4 + t) * 2);
        const float sa0 = (float)s[0], sb0 = (float)s[2];
        acc[0] = fmaf(f.sa, fmaf(d1.x, (float)d[0], m1.x * sa0), f.ma * z.x);
        acc[1] = fmaf(f.sa, fmaf(d1.y, (float)d[1], m1.y * sa0), f.ma * z.y);
        acc[2] = fmaf(f.sb, fmaf(d1.x, (float)d[2], m1.x * sb0), f.mb * z.x);
        acc[3] = fmaf(f.sb, fmaf(d1.y, (float)d[3], m1.y * sb0), f.mb * z.y);
    }
}

// ---- M == 1 ("octet"): blocks 8o .. 8o+7 of a tile = the units with chunk index 2o (chE) and 2o+1
// (chO); either may be absent (NULL).  Block j feeds MMA column j, so lane (g,t)'s D registers are
// the integer sums of blocks 2t, 2t+1 for rows g (d[0], d[1]) and g+8 (d[2], d[3]); that lane
// reads those two blocks' scales from the unit that holds them (t < 2: chE, else chO).
// Codes: B8 form [blk][8][32 B] with block b's row in column b&7, or compact [blk][32 B];
// planes [blk] (one activation row).  Returns the lane's scaled sums for rows g and g+8.
template <int QT, bool B8>
__device__ __forceinline__ void compute_octet(const uint8_t *chE, const uint8_t *chO, int o, const uint8_t *s_aq,
                                              const float *s_ap, int plane, int lane, float &v0, float &v2)
{
    const int g = lane >> 2, t = lane & 3;
    const uint8_t *mine = (t >> 1) ? chO : chE;
    float2 sA = make_float2(0.f, 0.f), sB = sA, mA = sA, mB = sA;
    if (mine) {
        const uint8_t *sp = mine + 1024 + (g * 4 + 2 * (t & 1)) * 4;
        sA = *reinterpret_cast<const float2 *>(sp);
        sB = *reinterpret_cast<const float2 *>(sp + 128);
        if constexpr (QT == 1) {
            mA = *reinterpret_cast<const float2 *>(sp + 256);
            mB = *reinterpret_cast<const float2 *>(sp + 384);
        }
    }
    const float2 d1 = *reinterpret_cast<const float2 *>(s_ap + o * 8 + 2 * t);
    uint2 breg = make_uint2(0u, 0u);
    if constexpr (!B8) breg = *reinterpret_cast<const uint2 *>(s_aq + o * 256 + lane * 8);
    int dE[4] = {0, 0, 0, 0}, dO[4] = {0, 0, 0, 0};
    int sE[4] = {0, 0, 0, 0}, sO[4] = {0, 0, 0, 0};
    if constexpr (QT == 0) {
        const int2 cr = *reinterpret_cast<const int2 *>(s_ap + plane + o * 8 + 2 * t);
        dE[0] = cr.x; dE[1] = cr.y; dE[2] = cr.x; dE[3] = cr.y;       // own columns only matter; one copy suffices
    }
    if (chE) {
        const uint4 qa = *reinterpret_cast<const uint4 *>(chE + lane * 16), qb = *reinterpret_cast<const uint4 *>(chE + 512 + lane * 16);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t a[4];
            make_a(a, u4_get(qa, i), u4_get(qb, i));
            if constexpr (B8) {
                const uint2 b = *reinterpret_cast<const uint2 *>(s_aq + (o * 8 + i) * 256 + lane * 8);
                imma<QT>(dE, a, b.x, b.y, dE);
            } else {
                const bool col = g == i;
                imma<QT>(dE, a, col ? breg.x : 0u, col ? breg.y : 0u, dE);
            }
            if constexpr (QT == 1) {
                const bool col = g == i;
                imma<QT>(sE, a, col ? 0x10101010u : 0u, col ? 0x01010101u : 0u, sE);
            }
        }
    }
    if (chO) {
        const uint4 qa = *reinterpret_cast<const uint4 *>(chO + lane * 16), qb = *reinterpret_cast<const uint4 *>(chO + 512 + lane * 16);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t a[4];
            make_a(a, u4_get(qa, i), u4_get(qb, i));
            if constexpr (B8) {
                const uint2 b = *reinterpret_cast<const uint2 *>(s_aq + (o * 8 + 4 + i) * 256 + lane * 8);
                imma<QT>(dO, a, b.x, b.y, dO);
            } else {
                const bool col = g == 4 + i;
                imma<QT>(dO, a, col ? breg.x : 0u, col ? breg.y : 0u, dO);
            }
            if constexpr (QT == 1) {
                const bool col = g == 4 + i;
                imma<QT>(sO, a, col ? 0x10101010u : 0u, col ? 0x01010101u : 0u, sO);
            }
        }
    }
    // the two halves populate disjoint columns: D = dE + dO.  A lane whose half is absent has
    // zero scales (sA = sB = 0), so whatever its columns hold contributes nothing.
    const float p0 = (float)(dE[0] + dO[0]), p1 = (float)(dE[1] + dO[1]), p2 = (float)(dE[2] + dO[2]), p3 = (float)(dE[3] + dO[3]);
    if constexpr (QT == 0) {
        v0 = fmaf(sA.x * d1.x, p0, (sA.y * d1.y) * p1);
        v2 = fmaf(sB.x * d1.x, p2, (sB.y * d1.y) * p3);
    } else {
        const float2 m1 = *reinterpret_cast<const float2 *>(s_ap + plane + o * 8 + 2 * t);
        const float2 z  = *reinterpret_cast<const float2 *>(s_ap + 2 * plane + o * 8 + 2 * t);
        const float s0 = (float)(sE[0] + sO[0]), s1 = (float)(sE[1] + sO[1]), s2 = (float)(sE[2] + sO[2]), s3 = (float)(sE[3] + sO[3]);
        // d0*(d1*P + m1*S0) + m0*Z per block, two blocks per lane and row; an absent half has Z
        // multiplied by m0 = 0 and everything else by d0 = 0
        v0 = fmaf(sA.x, fmaf(d1.x, p0, m1.x * s0), mA.x * z.x) + fmaf(sA.y, fmaf(d1.y, p1, m1.y * s1), mA.y * z.y);
        v2 = fmaf(sB.x, fmaf(d1.x, p2, m1.x * s2), mB.x * z.x) + fmaf(sB.y, fmaf(d1.y, p3, m1.y * s3), mB.y * z.y);
    }
}

// Reducer: the sums of tile T are complete (acc): store Y.  Lane roles:
//   M1:      lane l < 16 carries row l (activation row 0);
//   MP == 2: lane (g, r) carries row g + 8*(r>>1), activation row m = r&1;
//   MP == 8: lane (g, t) carries rows g/g+8 x m = 2t, 2t+1 in acc[0..3].
template <int QT, int MP, bool M1>
__device__ __forceinline__ void pipe_store(const PipeParams &p, int T, int lane, const float (&acc)[4], int pb)
{
    const float post = (QT == 0) ? 0.0625f : 1.0f;       // undo the 16x of the A/B code scaling (exact)
    const int g = lane >> 2, t = lane & 3;
    if (M1) {
        const int n = T * TILE_ROWS + lane;
        if (lane < TILE_ROWS && n < p.N) store_y(p.y, p.peer, pb, n, acc[0] * post);
    } else if (MP == 2) {
        const int m = t & 1, n = T * TILE_ROWS + g + 8 * (t >> 1);
        if (m < p.M && n < p.N) store_y(p.y, p.peer, pb, (size_t)m * p.y_stride + n, acc[0] * post);
    } else {
        const int n0 = T * TILE_ROWS + g, n1 = n0 + 8;
        const int m0 = 2 * t, m1 = m0 + 1;
        if (m0 < p.M) {
            if (n0 < p.N) store_y(p.y, p.peer, pb, (size_t)m0 * p.y_stride + n0, acc[0] * post);
            if (n1 < p.N) store_y(p.y, p.peer, pb, (size_t)m0 * p.y_stride + n1, acc[2] * post);
        }
        if (m1 < p.M) {
            if (n0 < p.N) store_y(p.y, p.peer, pb, (size_t)m1 * p.y_stride + n0, acc[1] * post);
            if (n1 < p.N) store_y(p.y, p.peer, pb, (size_t)m1 * p.y_stride + n1, acc[3] * post);
        }
    }
}

// K-sliced launches, second half: Y[m][n] = sum over the PP parked pieces of tile n/16, in piece order
// (reproducible rounding).  One thread per output; the slot index of (m, row) mirrors the lane roles of
// pipe_store.  PDL on both sides: it waits for the decode kernel that parked the pieces, and the next decode
// launch may already be filling its stage ring while this runs (it needs no shared memory, so both fit on an SM).
// With peers (fused all-gather) it is this kernel that stores into every rank's Y and signals.
template <int QT, int MP, bool M1>
__global__ void pipe_finalize_kernel(const float *__restrict__ partial, int PP, int N, int M, float *y, size_t y_stride, PeerOut po)
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    const int pb = po.world > 1 ? (int)((ld_volatile_u32(po.seq_ctr) + 1) & 1u) : 0;
    const float post = (QT == 0) ? 0.0625f : 1.0f;       // undo the 16x of the A/B code scaling (exact)
    const long total = (long)M * N;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i - (long)m * N);
        const int T = n / TILE_ROWS, row = n % TILE_ROWS, g = row & 7, hi = row >> 3;
        int q;                                           // slot index = r * 32 + lane of the reducer that parked it
        if (M1) q = row;
        else if (MP == 2) q = 4 * g + (m + 2 * hi);
        else q = ((m & 1) + 2 * hi) * 32 + 4 * g + (m >> 1);
        const float *rs = partial + (size_t)T * PP * 128 + q;
        float a = 0.f;
        for (int k0 = 0; k0 < PP; k0 += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (k0 + k < PP) ? __ldcg(rs + (size_t)(k0 + k) * 128) : 0.f;   // loads first
#pragma unroll
            for (int k = 0; k < 8; k++) a += v[k];       // absent pieces add +0
        }
        store_y(y, po, pb, (size_t)m * y_stride + n, a * post);
    }
    if (po.world > 1 && po.done) {
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < po.world) peer_add_share(po, threadIdx.x);
        if (threadIdx.x == 0) peer_count_done(po);
    }
}

// M1: exactly one activation row, fused quantization (rows stored: 1).  MP: padded row count of the
// other modes (2 or 8).  B8: 256-byte-per-block code layout (no per-IMMA masking).
template <int QT, int MP, bool FUSED, bool M1, bool B8, int NCW>
__global__ void __launch_bounds__((NCW + 2) * 32, NCW == 8 ? 2 : 1) q4_gemv_pipe_kernel(const PipeParams p)
{
    constexpr int PIPE_CONSUMERS = NCW;                  // consumer warps
    constexpr int PIPE_UNITS = 2 * NCW;                  // units per stage
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int CH = chunk_bytes(QT);
    constexpr int NP = (QT == 0) ? 2 : 3;
    constexpr int STAGE_BYTES = PIPE_UNITS * CH;
    constexpr int ROWS = M1 ? 1 : MP;                          // activation rows stored per block
    constexpr int RW = (MP == 2) ? 32 : 128;                   // floats per unit in the reduction ring
    constexpr int AQ_BLK = (MP == 2 && B8) ? 256 : ROWS * 32;  // code bytes per block in shared memory
    // K-slicing: this CTA's piece class, the first chunk / quant block of its slice
    const bool sliced = p.PP > 1;
    const int cls = sliced ? (int)blockIdx.x % p.PP : 0;
    const int cbase = cls * p.TS, blk0 = cbase * CHUNK_BLKS;
    // blocks held in the shared-memory pack: whole chunks (of the slice); M1 reads whole octets (8 blocks), so pad
    // to 8 (TS is even, so a slice is whole octets already)
    const int nbp = sliced ? p.TS * CHUNK_BLKS : (M1 ? ((p.C * CHUNK_BLKS + 7) & ~7) : p.C * CHUNK_BLKS);
    const int nb_l = min(p.nb - blk0, nbp);              // quant blocks of the slice that exist (>= 1)
    const int plane = nbp * ROWS;
    // shared-memory map: [stages x STAGE_BYTES][red ring][apack codes][planes][barriers][desc]
    uint8_t *s_stage = smem;
    float   *s_red = reinterpret_cast<float *>(smem + (size_t)p.stages * STAGE_BYTES);
    constexpr bool rdec = MP != 2 && B8;                   // decoupled, compact reduction ring (PipeParams::rslots > 0; the 8-row
                                                           // kernel has no use for B8's own meaning, the 8-column code layout)
    const int  rent = rdec ? PIPE_CONSUMERS : PIPE_UNITS;   // slots per ring entry
    uint8_t *s_aq = reinterpret_cast<uint8_t *>(s_red + (size_t)(rdec ? p.rslots : p.stages) * rent * RW);
    float   *s_ap = reinterpret_cast<float *>(s_aq + (size_t)nbp * AQ_BLK);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_ap) + (((size_t)NP * plane * 4 + 15) & ~(size_t)15));
    uint64_t *bar_empty = bar_full + PIPE_MAX_STAGES;
    uint64_t *bar_done = bar_empty + PIPE_MAX_STAGES;       // consumers finished a stage's partials
    int4     *s_desc = reinterpret_cast<int4 *>(bar_done + PIPE_MAX_STAGES);   // {n units (<0 = stop), chunk index of first unit, tile of first unit, task}

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x < 3 * PIPE_MAX_STAGES) {             // one barrier per thread: a single round of shared-memory latency
        const int i = threadIdx.x % PIPE_MAX_STAGES, which = threadIdx.x / PIPE_MAX_STAGES;
        if (i < p.stages) {
            if (which == 0) mbar_init(&bar_full[i], 1);
            else if (which == 1) mbar_init(&bar_empty[i], 1);   // released by the reducer (which waited for every consumer)
            else mbar_init(&bar_done[i], 1);             // (decoupled reduction ring: entry i is free again; arrived by the reducer)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) PIPE_STAMP(0);
    cudaTriggerProgrammaticLaunchCompletion();           // let the next launch start streaming early

    if (warp == PIPE_CONSUMERS) {
        // ===================== producer =====================
        if (lane == 0) {
            int s = 0;                                   // stage index
            uint32_t ph = 1;                             // parity to wait for on bar_empty (first lap: passes at once)
            const uint64_t wpol = l2_policy_evict_first();   // the weight stream must not push the activations and tables out of L2
            auto issue = [&](int u0, int n, int T, int c0, int task) {
                for (int r = 0; r < n; r += PIPE_UNITS) {
                    mbar_wait(&bar_empty[s], ph);
                    const int cnt = min(PIPE_UNITS, n - r);
                    s_desc[s] = make_int4(cnt, c0, T, task);
                    c0 += cnt;
                    while (c0 >= p.C) { c0 -= p.C; ++T; }
                    mbar_arrive_expect_tx(&bar_full[s], (uint32_t)cnt * CH);
                    bulk_g2s_hint(s_stage + (size_t)s * STAGE_BYTES, p.wt + (size_t)(u0 + r) * CH, (uint32_t)cnt * CH, &bar_full[s], wpol);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            };
            // task t of this CTA's sequence: unsliced = task number, sliced = tile number (piece `cls` of it)
            auto issue_seq = [&](int t) {
                if (sliced) issue(t * p.C + cbase, min(p.TS, p.C - cbase), t, cbase, t * p.PP + cls);
                else {
                    int u0, n, T, c0;
                    task_range(p, t, u0, n, T, c0);
                    issue(u0, n, T, c0, t);
                }
            };
            // The first PIPE_STATIC_TASKS tasks of every CTA are static (task = b + k*grid; sliced: tile =
            // idx + k*ncls inside the CTA's class): they need nothing from the previous launch, so under
            // programmatic dependent launch the whole stage ring is already filling while the previous kernel drains.
            const int idx = sliced ? (int)blockIdx.x / p.PP : (int)blockIdx.x;
            const int ncls = sliced ? ((int)gridDim.x - cls + p.PP - 1) / p.PP : (int)gridDim.x;     // CTAs in this class
            const int limit = sliced ? p.n_tiles : p.ntasks;
            const int nstatic = p.nstatic > 0 ? p.nstatic : PIPE_STATIC_TASKS;
            for (int k = 0; k < nstatic; k++) {
                const int t = idx + k * ncls;
                if (t < limit) issue_seq(t);
            }
            PIPE_STAMP(1);
            cudaGridDependencySynchronize();
            PIPE_STAMP(2);
            // then dynamic tickets, fetched one task ahead so the atomic's latency hides behind the copies
            int *tk = sliced ? p.ticket + 2 + cls : p.ticket;
            const int base = nstatic * ncls;
            int cur = base + atomicAdd(tk, 1);
            while (cur < limit) {
                const int nxt = base + atomicAdd(tk, 1);
                issue_seq(cur);
                cur = nxt;
            }
            PIPE_STAMP(10);                              // last copy issued
            prefetch_next_launch(p.next, (int)blockIdx.x);
            mbar_wait(&bar_empty[s], ph);                // stop marker
            s_desc[s] = make_int4(-1, 0, 0, 0);
            mbar_arrive(&bar_full[s]);
            PIPE_STAMP(8);
            // the last CTA to run out of tasks re-arms the ticket counter for the next launch
            if (atomicAdd(p.ticket + 1, 1) == (int)gridDim.x - 1) {
                p.ticket[0] = 0; p.ticket[1] = 0;
                for (int c = 0; c < p.PP; c++) p.ticket[2 + c] = 0;
                __threadfence();
            }
        }
        return;
    }

    if (warp == PIPE_CONSUMERS + 1) {
        // ===================== reducer =====================
        cudaGridDependencySynchronize();
        const int pb = p.peer.world > 1 ? (int)((ld_volatile_u32(p.peer.seq_ctr) + 1) & 1u) : 0;   // this exchange's Y buffer
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int curT = -1, curTask = -1;
        int s = 0;
        uint32_t ph = 0;
        unsigned rq = 0;                                 // stages summed so far (decoupled reduction ring: entry rq % rslots)
        constexpr int NV = (MP == 2) ? 1 : 4;
        // sliced launches: a finished piece is parked in its workspace slot with plain stores -- no fence, no
        // atomic; pipe_finalize_kernel (the next launch in the stream) adds the pieces and stores Y
        auto emit = [&](int T, int task) {
            if (!sliced) { pipe_store<QT, MP, M1>(p, T, lane, acc, pb); return; }
            float *ws = p.partial + (size_t)task * 128;
#pragma unroll
            for (int r = 0; r < NV; r++) ws[r * 32 + lane] = acc[r];
        };
        for (;;) {
            // bar_done implies bar_full (every consumer passed it), so one wait per stage suffices;
            // the stop marker carries no consumer arrivals and is announced on bar_full alone.
            // (pure polling: the one reducer warp should notice a finished stage at once, not after a suspended try_wait)
            // Consumers -> reducer: named barrier PIPE_BAR_DONE + s (bar.arrive by the consumer threads, bar.sync here): the
            // barrier unit takes 16 warps' arrivals in a few cycles where 16 arrivals on one mbarrier cost every warp ~350
            // cycles of contention per stage (tools/octetbench.cu).  The stop marker has no arrivals: look at the stage first.
            if (p.flags & 1) mbar_spin(&bar_full[s], ph); else mbar_wait(&bar_full[s], ph);
            const int4 dsc = s_desc[s];
            if (dsc.x < 0) break;
            asm volatile("bar.sync %0, %1;" :: "r"(PIPE_BAR_DONE + s), "n"(PIPE_CONSUMERS * 32 + 32) : "memory");
            if (lane == 0) PIPE_STAMP(11);               // (timeline builds) the last stage's partials became visible
            const int task = dsc.w;
            int T = dsc.z;
            int left = p.C - dsc.y;                      // units of tile T from the stage's first unit on
            const int re = rdec ? (int)(rq % (unsigned)p.rslots) : s;
            const float *red = s_red + (size_t)re * rent * RW + (MP == 2 ? lane : lane * 4);
            int j = 0;
            while (j < dsc.x) {
                if (T != curT || task != curTask) {
                    if (curT >= 0) emit(curT, curTask);
                    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                    curT = T; curTask = task;
                }
                const int run = min(dsc.x - j, left);    // units of this stage that belong to tile T
                if (MP != 2 && rdec) {
                    // one slot per warp = unit pair (j and run are even here): pairs of slots, then a single one -- the order the
                    // unit-wise loop below gives when every other slot is zero
                    if constexpr (MP != 2) {
                        int k = j >> 1;
                        const int ke = (j + run + 1) >> 1;       // (an odd count only ends a whole tile of an odd number of units: its last warp has one unit)
                        for (; k + 2 <= ke; k += 2) {
                            const float4 a = *reinterpret_cast<const float4 *>(red + k * RW), c = *reinterpret_cast<const float4 *>(red + (k + 1) * RW);
                            acc[0] += a.x + c.x; acc[1] += a.y + c.y; acc[2] += a.z + c.z; acc[3] += a.w + c.w;
                        }
                        for (; k < ke; k++) {
                            const float4 v = *reinterpret_cast<const float4 *>(red + k * RW);
                            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                        }
                    }
                } else if (run == PIPE_UNITS && MP == 2) {
                    // common case, the whole stage is one tile: eight independent chains of four
                    // (fixed association -> bit-reproducible, short dependent latency)
                    float v[PIPE_UNITS];
#pragma unroll
                    for (int k = 0; k < PIPE_UNITS; k++) v[k] = red[k * RW];
                    float q[PIPE_UNITS / 4];
#pragma unroll
                    for (int h = 0; h < PIPE_UNITS / 4; h++) q[h] = (v[4 * h] + v[4 * h + 1]) + (v[4 * h + 2] + v[4 * h + 3]);
                    if constexpr (PIPE_UNITS == 32) acc[0] += ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
                    else                           acc[0] += (q[0] + q[1]) + (q[2] + q[3]);
                } else {
                    int k = j;
                    for (; k + 4 <= j + run; k += 4) {   // fixed grouping by four
                        if constexpr (MP == 2) {
                            acc[0] += (red[k * RW] + red[(k + 1) * RW]) + (red[(k + 2) * RW] + red[(k + 3) * RW]);
                        } else {
                            const float4 a = *reinterpret_cast<const float4 *>(red + k * RW), b = *reinterpret_cast<const float4 *>(red + (k + 1) * RW);
                            const float4 c = *reinterpret_cast<const float4 *>(red + (k + 2) * RW), d = *reinterpret_cast<const float4 *>(red + (k + 3) * RW);
                            acc[0] += (a.x + b.x) + (c.x + d.x); acc[1] += (a.y + b.y) + (c.y + d.y);
                            acc[2] += (a.z + b.z) + (c.z + d.z); acc[3] += (a.w + b.w) + (c.w + d.w);
                        }
                    }
                    for (; k < j + run; k++) {
                        if constexpr (MP == 2) acc[0] += red[k * RW];
                        else {
                            const float4 v = *reinterpret_cast<const float4 *>(red + k * RW);
                            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                        }
                    }
                }
                j += run;
                left -= run;
                if (left == 0) { ++T; left = p.C; }
            }
            __syncwarp();
            if (lane == 0) PIPE_STAMP(12);               // (timeline builds) the last stage is summed
            if (lane == 0) mbar_arrive(&bar_empty[s]);
            if (rdec) { if (lane == 0) mbar_arrive(&bar_done[re]); ++rq; }
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (lane == 0) PIPE_STAMP(13);                   // (timeline builds) stop marker seen
        if (curT >= 0) emit(curT, curTask);
        if (!sliced) peer_signal(p.peer, lane);
        if (lane == 0) PIPE_STAMP(7);
        return;
    }

    // ===================== consumers =====================
    if (threadIdx.x == 0) PIPE_STAMP(3);
    const int ctid = threadIdx.x;                        // 0 .. 511
    constexpr int CTHREADS = PIPE_CONSUMERS * 32;
    // where (block, row m)'s 32 code bytes live in shared memory
    auto rec_of = [&](int blk, int m) {
        if constexpr (MP == 2 && B8) return s_aq + (size_t)blk * 256 + (M1 ? (blk & 7) : 2 * (blk & 3) + m) * 32;
        else return s_aq + ((size_t)blk * ROWS + m) * 32;
    };
    if constexpr (FUSED) {
        // Everything quantization does not write must read as zero: the unused columns of the
        // 8-column form, the records of padded blocks (nb .. nbp: weight scale 0, but 0 x garbage
        // could be NaN) and row 1 of an M = 1 call in the two-row layout.  None of it depends on X, so it
        // is done BEFORE the grid dependency resolves (quantization writes disjoint bytes).
        for (int i = ctid; i < nbp * AQ_BLK / 8; i += CTHREADS) {
            const int blk = i / (AQ_BLK / 8), w8 = i - blk * (AQ_BLK / 8);      // 8-byte piece w8 of block blk
            const int col = w8 >> 2;                                            // 32-byte record index inside the block
            int m = col;
            if constexpr (MP == 2 && B8) m = M1 ? (col == (blk & 7) ? 0 : -1) : col - 2 * (blk & 3);
            const bool live = blk < nb_l && m >= 0 && m < p.M;
            if (!live) reinterpret_cast<uint2 *>(s_aq + (size_t)blk * AQ_BLK)[w8] = make_uint2(0u, 0u);
        }
        for (int pl = 0; pl < NP; pl++) {
            for (int i = nb_l * ROWS + ctid; i < nbp * ROWS; i += CTHREADS) s_ap[pl * plane + i] = 0.0f;   // padded blocks
            if (ROWS == 2 && p.M == 1)
                for (int blk = ctid; blk < nb_l; blk += CTHREADS) s_ap[pl * plane + blk * 2 + 1] = 0.0f;  // absent row
        }
    }
    cudaGridDependencySynchronize();                     // X may be the previous kernel's output
    if (p.peer.world > 1) {
        // ... or the previous exchange's Y: wait until every rank's slice of it has landed here
        if (ctid < p.peer.world) peer_wait_one(p.peer.local_flags, ctid, ld_volatile_u32(p.peer.seq_ctr));
        asm volatile("bar.sync 1, %0;" :: "n"(CTHREADS) : "memory");
    }
    if constexpr (FUSED) {
        // Issue the activation loads first (one item = 8 consecutive elements of one block; the
        // first FUSED_ITEMS items of a thread are held in registers) so that their L2 latency
        // overlaps the zero-fill below.
        constexpr int FUSED_ITEMS = 2;
        const int items = p.M * nb_l * 4;
        const float *xs = p.x + (size_t)blk0 * QK;       // this CTA's K-slice of every activation row
        float4 xlo[FUSED_ITEMS], xhi[FUSED_ITEMS];
#pragma unroll
        for (int it = 0; it < FUSED_ITEMS; it++) {
            const int id = it * CTHREADS + ctid;
            xlo[it] = make_float4(0.f, 0.f, 0.f, 0.f); xhi[it] = xlo[it];
            if (id < items) {
                const int bq = id >> 2, m = bq / nb_l, blk = bq - m * nb_l;
                const float4 *src = reinterpret_cast<const float4 *>(xs + (size_t)m * p.x_stride + (size_t)blk * QK + (lane & 3) * 8);
                xlo[it] = src[0]; xhi[it] = src[1];
            }
        }
#pragma unroll
        for (int it = 0; it < FUSED_ITEMS; it++) {
            const int id = it * CTHREADS + ctid;
            const bool valid = id < items;
            if (it * CTHREADS < items) {                 // warp-uniform: shuffles inside
                const int bq = valid ? (id >> 2) : 0;
                const int m = bq / nb_l, blk = bq - m * nb_l;
                quantize8_emit<QT>(xlo[it], xhi[it], valid, lane, rec_of(blk, m), s_ap + (size_t)blk * ROWS + m, plane);
            }
        }
        for (int base = FUSED_ITEMS * CTHREADS; base < items; base += CTHREADS) {      // long rows (K > 8192)
            const int id = base + ctid;
            const bool valid = id < items;
            const int bq = valid ? (id >> 2) : 0;
            const int m = bq / nb_l, blk = bq - m * nb_l;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            if (valid) {
                const float4 *src = reinterpret_cast<const float4 *>(xs + (size_t)m * p.x_stride + (size_t)blk * QK + (lane & 3) * 8);
                lo = src[0]; hi = src[1];
            }
            quantize8_emit<QT>(lo, hi, valid, lane, rec_of(blk, m), s_ap + (size_t)blk * ROWS + m, plane);
        }
    } else {
        // global apack: compact form [gnbp][ROWS][32] + planes [NP][gnbp * ROWS] over the whole row; this CTA
        // takes blocks blk0 .. blk0 + nbp (the host sizes gnbp so that every slice is fully backed)
        const uint4 *src = reinterpret_cast<const uint4 *>(p.aq) + (size_t)blk0 * ROWS * 2;
        if constexpr (MP == 2 && B8) {
            for (int i = ctid; i < nbp * 16; i += CTHREADS) {                   // 16-byte pieces of the 8-column form
                const int blk = i >> 4, col = (i >> 1) & 7, piece = i & 1;
                const int m = M1 ? (col == (blk & 7) ? 0 : -1) : col - 2 * (blk & 3);
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (m >= 0 && m < ROWS) v = src[(blk * ROWS + m) * 2 + piece];
                reinterpret_cast<uint4 *>(s_aq)[i] = v;
            }
        } else {
            uint4 *dst = reinterpret_cast<uint4 *>(s_aq);
            for (int i = ctid; i < nbp * ROWS * 2; i += CTHREADS) dst[i] = src[i];
        }
        const int pq = plane / 4;                                               // 16-byte pieces per plane here
        for (int i = ctid; i < NP * pq; i += CTHREADS) {
            const int pl = i / pq, j = i - pl * pq;
            reinterpret_cast<uint4 *>(s_ap)[i] = reinterpret_cast<const uint4 *>(p.ap + ((size_t)pl * p.gnbp + blk0) * ROWS)[j];
        }
    }
    asm volatile("bar.sync 1, %0;" :: "n"(CTHREADS) : "memory");
    if (threadIdx.x == 0) PIPE_STAMP(4);

    const int uA = 2 * warp, uB = uA + 1;                // this warp's units in every stage
    // chunk index (and tile) of unit u of a stage
    auto chunk_of = [&](const int4 &dsc, int u, int &T) { int c = dsc.y + u; T = dsc.z; while (c >= p.C) { c -= p.C; ++T; } return c; };
    // partial sums -> reduction ring
    auto put4 = [&](float *slot, float (&acc)[4]) {
        if constexpr (MP == 2) {
            // quad reduction here (the consumers have the slack, the single reducer warp does not):
            // afterwards every lane of the quad holds all four sums; lane t keeps value t
            // (value index = 2*(row half) + m, the order of acc[]).
#pragma unroll
            for (int r = 0; r < 4; r++) {
                acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 1);
                acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 2);
            }
            const int t = lane & 3;
            slot[lane] = t == 0 ? acc[0] : t == 1 ? acc[1] : t == 2 ? acc[2] : acc[3];
        } else {
            *reinterpret_cast<float4 *>(slot + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
    };
    auto put_zero = [&](float *slot) {
        if constexpr (MP == 2) slot[lane] = 0.f;
        else *reinterpret_cast<float4 *>(slot + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // M1: rows g (v0) and g+8 (v2) -> lanes 0..15 of the slot hold rows 0..15
    auto put_rows = [&](float *slot, float v0, float v2) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, 1); v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
        const int g = lane >> 2, t = lane & 3;
        if (t < 2) slot[g + 8 * t] = t ? v2 : v0;
    };

    int s = 0;
    uint32_t ph = 0;
    unsigned cq = 0;                                     // stages consumed so far (decoupled reduction ring)
    bool first = true;
    for (;;) {
        mbar_wait(&bar_full[s], ph);
        if (first && threadIdx.x == 0) { PIPE_STAMP(5); } first = false;
        if (threadIdx.x == 0) PIPE_STAMP(9);             // (timeline builds) ends up holding the arrival of the last stage
        const int4 dsc = s_desc[s];
        if (dsc.x < 0) break;
        const bool hA = uA < dsc.x, hB = uB < dsc.x;
        int TA = 0, TB = 0;
        const int cA = chunk_of(dsc, uA, TA), cB = chunk_of(dsc, uB, TB);
        const uint8_t *chA = s_stage + (size_t)s * STAGE_BYTES + (size_t)uA * CH, *chB = chA + CH;
        float *slotA = s_red + ((size_t)s * PIPE_UNITS + uA) * RW, *slotB = slotA + RW;
        if (rdec) {                                      // this warp's one slot of entry cq % rslots, once the reducer has read its previous content
            const unsigned re = cq % (unsigned)p.rslots, use = cq / (unsigned)p.rslots;
            if (dsc.x >= 0) mbar_wait(&bar_done[re], (use & 1u) ^ 1u);
            slotA = s_red + ((size_t)re * PIPE_CONSUMERS + warp) * RW;
            slotB = slotA;
            ++cq;
        }
        if constexpr (M1) {
            if (hA && hB && TA == TB && !(cA & 1)) {     // an aligned octet: the common case
                float v0, v2;
                compute_octet<QT, B8>(chA, chB, (cA - cbase) >> 1, s_aq, s_ap, plane, lane, v0, v2);
                put_rows(slotA, v0, v2);
                if (lane < TILE_ROWS) slotB[lane] = 0.f;
            } else {                                     // half octets (odd start, tile boundary, ragged tail)
                if (hA) {
                    float v0, v2;
                    compute_octet<QT, B8>((cA & 1) ? nullptr : chA, (cA & 1) ? chA : nullptr, (cA - cbase) >> 1, s_aq, s_ap, plane, lane, v0, v2);
                    put_rows(slotA, v0, v2);
                }
                if (hB) {
                    float v0, v2;
                    compute_octet<QT, B8>((cB & 1) ? nullptr : chB, (cB & 1) ? chB : nullptr, (cB - cbase) >> 1, s_aq, s_ap, plane, lane, v0, v2);
                    put_rows(slotB, v0, v2);
                }
            }
        } else {
            Frag<QT> fA, fB;
            float aA[4], aB[4];
            if (hA) load_frag_smem<QT>(fA, chA, lane);
            if (hB) load_frag_smem<QT>(fB, chB, lane);
            auto run_unit = [&](const Frag<QT> &f, const uint8_t *chunk, int c, float (&acc)[4]) {
                if constexpr (MP == 2) compute_unit_chain4c<QT, B8>(f, c, s_aq, s_ap, plane, lane, acc);
                else {
                    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                    compute_unit_m8<QT>(f, chunk, c, s_aq, s_ap, plane, lane, acc);
                }
            };
            if (hA && hB) {
                run_unit(fA, chA, cA - cbase, aA);
                run_unit(fB, chB, cB - cbase, aB);
                if (TA == TB) {                          // same tile: one partial for both units
                    aA[0] += aB[0]; aA[1] += aB[1]; aA[2] += aB[2]; aA[3] += aB[3];
                    put4(slotA, aA);
                    if (!rdec) put_zero(slotB);
                } else {
                    put4(slotA, aA);
                    put4(slotB, aB);
                }
            } else if (hA) {
                run_unit(fA, chA, cA - cbase, aA);
                put4(slotA, aA);
            }
        }
        asm volatile("bar.arrive %0, %1;" :: "r"(PIPE_BAR_DONE + s), "n"(PIPE_CONSUMERS * 32 + 32) : "memory");
        if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    if (threadIdx.x == 0) PIPE_STAMP(6);
}

// ---------------------------------------------------------------------------
// Exact mode (b200q4_set_exact / B200Q4_EXACT=1): the reference's OWN rounding sequence, so that Y is
// bit-identical to the CPU path and whole programs reproduce its logits and token streams exactly.
// ggml_vec_dot_q4_0 (ggml.c:484-510 == imax.c:1210-1228) is
//     sumf = 0;  for block i, for byte j:  f0 = d0*(lo0-8), f1 = d0*(hi0-8), f2 = d1*(lo1-8), f3 = d1*(hi1-8);
//                                          sumf += f0*f2 + f1*f3;
// with every * and + an individually rounded fp32 operation in exactly this order (the reference is built
// -msse3: no FMA), and ggml_vec_dot_q4_1 (ggml.c:525-552) the same with f = d*q + m.  That is one strictly
// sequential chain of K/2 additions per output: no tensor cores, no splitting of K, one thread per weight
// row with the dequantized activation rows (f2, f3: a function of the activations alone) read as broadcast
// loads.  Several times slower than the default kernels, whose block sums are exact integers and whose
// results are closer to the real-number value -- but different in the last bits, which the reference's own
// graph amplifies (its gelu goes through an fp16 table and every mul_mat re-quantizes its input to 4 bits:
// a rounding-level difference flips codes downstream).
// ---------------------------------------------------------------------------
// reference-format quantized rows -> the floats the reference's dot product forms from them
template <int QT>
__global__ void exact_dequant_act_kernel(const uint8_t *__restrict__ rows, int M, int nb, float *__restrict__ deq)
{
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)M * nb * QK;
    if (id >= total) return;
    const int e = (int)(id % QK);
    const long bm = id / QK;
    const int blk = (int)(bm % nb), m = (int)(bm / nb);
    float v;
    if (QT == 0) {
        const uint8_t *b = rows + ((size_t)m * nb + blk) * 20;
        const float d = *reinterpret_cast<const float *>(b);
        const uint8_t byte = b[4 + (e >> 1)];
        const int q = (e & 1) ? (byte >> 4) : (byte & 15);
        v = __fmul_rn(d, (float)(q - 8));
    } else {
        const uint8_t *row = rows + (size_t)m * nb * 24;
        const float mn = reinterpret_cast<const float *>(row)[blk], d = reinterpret_cast<const float *>(row)[nb + blk];
        const uint8_t byte = row[(size_t)nb * 8 + (size_t)blk * 16 + (e >> 1)];
        const int q = (e & 1) ? (byte >> 4) : (byte & 15);
        v = __fadd_rn(__fmul_rn(d, (float)q), mn);
    }
    deq[id] = v;                                         // [m][blk*32 + e]
}

// One 128-thread block per 16-row weight tile, MB activation rows per launch.  Per chunk (4 quant blocks = 64
// byte pairs) the block works in two steps:
//   produce  every thread (row r = tid / 8, pair group q = tid % 8) forms the eight terms s = f0*f2 + f1*f3 of
//            its row -- independent work, all the rounding of the reference but none of its ordering -- and
//            leaves them in shared memory, [m][pair][row];
//   chain    lane (m, row) of the first 16*M threads adds the 64 terms of the chunk to its running sum in pair
//            order: sumf = sumf + s, the reference's one serial dependency (one FADD per byte pair).
// The chain step is the floor (K/2 dependent additions per output); the produce step keeps everything else
// off it.
constexpr int EXACT_PAIR_STRIDE = 17;                    // floats per pair row in shared memory (16 + 1: spreads the banks)
template <int QT, int MB>
__global__ void __launch_bounds__(128) mul_mat_exact_kernel(const uint8_t *__restrict__ wt, int N, int C, int nb,
                                                            const float *__restrict__ deq, int M, int K,
                                                            float *__restrict__ y, size_t y_stride)
{
    constexpr int CH = chunk_bytes(QT);
    constexpr int MSTRIDE = 64 * EXACT_PAIR_STRIDE + 8;  // floats per activation row
    __shared__ float s_s[MB * MSTRIDE];
    const int T = blockIdx.x, tid = threadIdx.x;
    const int r = tid >> 3, q = tid & 7, half = r >> 3, g = r & 7;
    const int i = q >> 1, t0 = (q & 1) * 2;              // this thread's quant block of the chunk and its two nibble words
    const int cm = tid >> 4, cr = tid & 15;              // chain role: activation row, weight row
    const bool chain = cm < M && cm < MB;
    float sum = 0.0f;
    // this thread's weight words and scales of chunk c (loaded one chunk ahead: the global latency hides behind a whole step)
    uint32_t nw0 = 0, nw1 = 0;
    float nd0 = 0.f, nm0 = 0.f;
    auto fetch = [&](int c) {
        const uint8_t *chunk = wt + ((size_t)T * C + c) * CH;
        nw0 = __ldg(reinterpret_cast<const uint32_t *>(chunk + half * 512 + (g * 4 + t0) * 16) + i);
        nw1 = __ldg(reinterpret_cast<const uint32_t *>(chunk + half * 512 + (g * 4 + t0 + 1) * 16) + i);
        nd0 = __ldg(reinterpret_cast<const float *>(chunk + 1024) + r * 4 + i);
        if (QT == 1) nm0 = __ldg(reinterpret_cast<const float *>(chunk + 1280) + r * 4 + i);
    };
    fetch(0);
    for (int c = 0; c < C; c++) {
        const int blk = c * CHUNK_BLKS + i;
        const uint32_t w2[2] = {nw0, nw1};
        const float d0 = nd0, m0 = nm0;
        if (c + 1 < C) fetch(c + 1);
        if (blk < nb) {
            const float *a = deq + (size_t)blk * QK;
#pragma unroll
            for (int bw = 0; bw < 2; bw++) {
                const int t = t0 + bw;
                float4 fa[MB], fb[MB];                   // f2, f3 of byte pairs 4t .. 4t+3 (broadcast loads)
#pragma unroll
                for (int m = 0; m < MB; m++) {           // rows >= M recompute row M-1: no branches, the loads pipeline
                    const size_t mo = (size_t)(m < M ? m : M - 1) * K;
                    fa[m] = __ldg(reinterpret_cast<const float4 *>(a + mo + 8 * t));
                    fb[m] = __ldg(reinterpret_cast<const float4 *>(a + mo + 8 * t + 4));
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t byte = (w2[bw] >> (8 * b)) & 255u;
                    float f0, f1;
                    if (QT == 0) {
                        f0 = __fmul_rn(d0, (float)((int)(byte & 15u) - 8));
                        f1 = __fmul_rn(d0, (float)((int)(byte >> 4) - 8));
                    } else {
                        f0 = __fadd_rn(__fmul_rn(d0, (float)(byte & 15u)), m0);
                        f1 = __fadd_rn(__fmul_rn(d0, (float)(byte >> 4)), m0);
                    }
                    const int p = i * 16 + 4 * t + b;    // pair index inside the chunk = the reference's (block, j) order
#pragma unroll
                    for (int m = 0; m < MB; m++) {
                        const float4 &v = b < 2 ? fa[m] : fb[m];
                        const float f2 = (b & 1) ? v.z : v.x, f3 = (b & 1) ? v.w : v.y;
                        s_s[m * MSTRIDE + p * EXACT_PAIR_STRIDE + r] = __fadd_rn(__fmul_rn(f0, f2), __fmul_rn(f1, f3));
                    }
                }
            }
        }
        __syncthreads();
        if (chain) {
            const int np = min(64, (nb - c * CHUNK_BLKS) * 16);      // padding blocks are not part of the reference's sum
            const float *sp = s_s + cm * MSTRIDE + cr;
            for (int p0 = 0; p0 < np; p0 += 16) {        // np is a multiple of 16: the loads of a batch first, then the chain
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = sp[(p0 + k) * EXACT_PAIR_STRIDE];
#pragma unroll
                for (int k = 0; k < 16; k++) sum = __fadd_rn(sum, v[k]);
            }
        }
        __syncthreads();
    }
    const int n = T * TILE_ROWS + cr;
    if (chain && n < N) y[(size_t)cm * y_stride + n] = sum;
}

// ---------------------------------------------------------------------------
// Transposed-src0 branch (ggml.c:5118-5164 / 5391-5440, ggml_vec_mad_q4_*):
// Wt is K rows of N quantized values (reference row format, kept as is on the
// device), activations are NOT quantized.  Y[m][n] = sum_k deq(Wt[k][n]) * X[m][k],
// accumulated in k order like a single-thread reference run.  One thread per
// (m, n); consecutive threads read consecutive nibbles of the same Wt row.
// ---------------------------------------------------------------------------
template <int QT>
__global__ void mul_mat_transposed_kernel(const uint8_t *__restrict__ wt, int N, int K,
                                          const float *__restrict__ x, size_t x_stride, int M,
                                          float *__restrict__ y, size_t y_stride)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (n >= N || m >= M) return;
    const int nbn = N / QK, blk = n / QK, e = n % QK;
    const size_t rb = (size_t)nbn * (QT == 0 ? 20 : 24);
    float acc = 0.0f;
    for (int k = 0; k < K; k++) {
        const uint8_t *row = wt + (size_t)k * rb;
        float v;
        if (QT == 0) {
            const uint8_t *b = row + (size_t)blk * 20;
            const float d = *reinterpret_cast<const float *>(b);
            const uint8_t q = b[4 + (e >> 1)];
            const int c = (e & 1) ? (q >> 4) : (q & 15);
            v = __fmul_rn((float)(c - 8), d);
        } else {
            const float *rf = reinterpret_cast<const float *>(row);
            const float mn = rf[blk], d = rf[nbn + blk];
            const uint8_t q = row[(size_t)nbn * 8 + (size_t)blk * 16 + (e >> 1)];
            const int c = (e & 1) ? (q >> 4) : (q & 15);
            v = __fadd_rn(__fmul_rn(d, (float)c), mn);
        }
        acc = __fadd_rn(acc, __fmul_rn(v, x[(size_t)m * x_stride + k]));
    }
    y[(size_t)m * y_stride + n] = acc;
}

} // namespace b200q4
