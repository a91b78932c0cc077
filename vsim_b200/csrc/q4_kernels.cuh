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

// Writes one block's apack record: 32 code bytes at aq_rec and the block parameters into the
// parameter planes (plane 0: d1; Q4_1 also plane 1: m1, plane 2: Z).  ap_rec points at this
// (block,row)'s slot of plane 0; `plane` is the plane stride in floats.  Global or shared memory.
template <int QT>
__device__ __forceinline__ void emit_apack(const BlockQ &q, int lane, uint8_t *aq_rec, float *ap_rec, int plane)
{
    const int a = (QT == 0) ? (q.code & 15) - 8 : (q.code & 15);
    aq_rec[apack_byte_of(lane)] = (uint8_t)(int8_t)a;
    if (QT == 0) {
        if (lane == 0) ap_rec[0] = q.d;
    } else {
        int s = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            ap_rec[0] = q.d;
            ap_rec[plane] = q.mn;
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
__device__ __forceinline__ void quantize8_emit(const float *x8, bool valid, int lane, uint8_t *aq_rec, float *ap_rec, int plane)
{
    const int tq = lane & 3, qbase = lane & 28;
    float v[8];
    if (valid) {
        const float4 lo = *reinterpret_cast<const float4 *>(x8), hi = *reinterpret_cast<const float4 *>(x8 + 4);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.0f;
    }
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
    const uint32_t even = (uint32_t)(a[0] & 255) | ((uint32_t)(a[2] & 255) << 8) | ((uint32_t)(a[4] & 255) << 16) | ((uint32_t)(a[6] & 255) << 24);
    const uint32_t odd  = (uint32_t)(a[1] & 255) | ((uint32_t)(a[3] & 255) << 8) | ((uint32_t)(a[5] & 255) << 16) | ((uint32_t)(a[7] & 255) << 24);
    if (valid) *reinterpret_cast<uint2 *>(aq_rec + tq * 8) = make_uint2(even, odd);
    if (QT == 0) {
        if (valid && tq == 0) ap_rec[0] = d;
    } else {
        int sum = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7];
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (valid && tq == 0) {
            ap_rec[0] = d;
            ap_rec[plane] = mn;
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
// The decode kernel: dequant-fused skinny GEMM for M <= MP activation rows
// (MP = 2: decode GEMV, activation quantization fused into the prologue;
//  MP = 8: prompt chunks / small batches).
//
// Persistent grid (one 16-warp CTA per SM).  The weight stream is cut into
// "units" (one 1280/1536-byte chunk each = 16 rows x 4 blocks); the U units of
// the tensor, in memory order, are dealt out as one contiguous range per warp
// (+-1 unit), so every SM moves the same number of bytes whatever N and K are
// (N=4096 has only 256 row tiles for 148 SMs; whole-tile scheduling would idle
// 14% of the chip).  A warp whose range cuts a 16-row tile writes its partial
// sums to a workspace slot; the last warp to arrive at that tile (atomic
// counter) adds the slots in warp order -- a fixed order, so results are
// bit-reproducible run to run -- and stores Y.
//
// Per unit a lane issues 2 x LDG.128 (nibbles) + 2 (4 for Q4_1) x LDG.32
// (scales), PD units ahead of use (register ring).  Nibbles become the A
// fragment with 3 ALU ops per word:
//   Q4_0: lo = ((w<<4) & 0xF0F0F0F0) ^ 0x80808080, hi = (w & 0xF0F0F0F0) ^ 0x80808080
//         = 16*(q-8) as s8 (the XOR flips the nibble's top bit into a sign bit);
//         the factor 16 is removed exactly (x 0.0625) when Y is stored;
//   Q4_1: lo = w & 0x0F0F0F0F, hi = (w>>4) & 0x0F0F0F0F = q as u8.
// MP = 2 ("chain4"): the 8 columns of the m16n8k32 tile are used as
//   (block 4c+i, row m) -> column 2i+m; the four blocks' B fragments are
//   one-hot in their column pair, so the four IMMAs of a unit accumulate into
//   ONE D fragment and lane (g,t) ends up with the integer sums of block 4c+t
//   for rows g/g+8 -- exactly the block whose scale d0[g][t] it loaded.  One
//   I2F + FMUL + FFMA round per unit instead of four, no shuffles.
// MP = 8: one IMMA per block, D column pair (2t,2t+1) = activation rows; the
//   scale of block i comes from lane (g,i) of the quad by shuffle.
// ---------------------------------------------------------------------------
struct GemvParams {
    const uint8_t *wt;      // wtile
    const float   *x;       // fp32 activations (fused quantization, MP == 2)
    const uint8_t *aq;      // apack codes  [nbp][MP][32]        (non-fused)
    const float   *ap;      // apack planes [NP][nbp][MP] floats (non-fused)
    float         *y;
    float         *partial; // [2*Weff][4*32]
    int           *counters;// [n_tiles], zero between launches
    size_t         x_stride;// floats
    size_t         y_stride;// floats
    int            U;       // n_tiles * C  (< 2^31, checked on the host)
    int            N, M, C, n_tiles, nb;
    int            Weff;    // warps that own a non-empty unit range
};

template <int QT> struct Frag;
// sa/sb: lane (g,t) keeps d0 of rows g / g+8 for block 4c+t only.
template <> struct Frag<0> { uint4 qa, qb; float sa, sb; };
template <> struct Frag<1> { uint4 qa, qb; float sa, sb, ma, mb; };
__device__ __forceinline__ float ldg_stream_f1(const void *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

template <int QT>
__device__ __forceinline__ void load_frag(Frag<QT> &f, const uint8_t *chunk, int lane)
{
    f.qa = ldg_stream_u4(chunk + lane * 16);
    f.qb = ldg_stream_u4(chunk + 512 + lane * 16);
    f.sa = ldg_stream_f1(chunk + 1024 + lane * 4);          // d0[g][t]      (128 B per warp)
    f.sb = ldg_stream_f1(chunk + 1024 + 128 + lane * 4);    // d0[g+8][t]
    if constexpr (QT == 1) {
        f.ma = ldg_stream_f1(chunk + 1280 + lane * 4);
        f.mb = ldg_stream_f1(chunk + 1280 + 128 + lane * 4);
    }
}

__device__ __forceinline__ uint32_t u4_get(const uint4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// (a & b) ^ c in ONE LOP3 (immLut 0x6a); b and c are passed in registers on purpose -- with two
// immediates the compiler has to emit two LOP3s.
__device__ __forceinline__ uint32_t and_xor(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int QT>
__device__ __forceinline__ void make_a(uint32_t (&a)[4], uint32_t wa, uint32_t wb, uint32_t kmask, uint32_t ksign)
{
    if constexpr (QT == 0) {
        a[0] = and_xor(wa << 4, kmask, ksign);
        a[1] = and_xor(wb << 4, kmask, ksign);
        a[2] = and_xor(wa, kmask, ksign);
        a[3] = and_xor(wb, kmask, ksign);
    } else {
        a[0] = wa & 0x0F0F0F0Fu;
        a[1] = wb & 0x0F0F0F0Fu;
        a[2] = (wa >> 4) & 0x0F0F0F0Fu;
        a[3] = (wb >> 4) & 0x0F0F0F0Fu;
    }
}

// D = A*B + C; A s8 (Q4_0) or u8 (Q4_1), B s8 (Q4_0) or u8 (Q4_1)
template <int QT>
__device__ __forceinline__ void imma(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const int (&c)[4])
{
    if constexpr (QT == 0)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}

// ---- MP == 2: four blocks chained into one D fragment ------------------------------------------
// Shared-memory apack for this mode is the 8-column form [blk][8][32 B] in which block 4c+i only
// populates columns 2i (row 0) and 2i+1 (row 1); parameter planes are [blk][2].
template <int QT>
__device__ __forceinline__ void compute_unit_chain4(const Frag<QT> &f, int c, const uint8_t *s_aq, const float *s_ap,
                                                    int plane, int lane, uint32_t kmask, uint32_t ksign, float (&acc)[4])
{
    const int g = lane >> 2, t = lane & 3;
    const uint8_t *bp = s_aq + c * 1024 + lane * 8;
    const float2 d1 = *reinterpret_cast<const float2 *>(s_ap + (c * 4 + t) * 2);
    int d[4] = {0, 0, 0, 0};
    int s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < CHUNK_BLKS; i++) {
        uint32_t a[4];
        make_a<QT>(a, u4_get(f.qa, i), u4_get(f.qb, i), kmask, ksign);
        const uint2 b = *reinterpret_cast<const uint2 *>(bp + i * 256);
        imma<QT>(d, a, b.x, b.y, d);
        if constexpr (QT == 1) {           // S0 = sum of the block's weight codes: ones in column 2i only
            const uint32_t one = (g == 2 * i) ? 0x01010101u : 0u;
            imma<QT>(s, a, one, one, s);
        }
    }
    if constexpr (QT == 0) {
        acc[0] = fmaf(f.sa * d1.x, (float)d[0], acc[0]);
        acc[1] = fmaf(f.sa * d1.y, (float)d[1], acc[1]);
        acc[2] = fmaf(f.sb * d1.x, (float)d[2], acc[2]);
        acc[3] = fmaf(f.sb * d1.y, (float)d[3], acc[3]);
    } else {
        const float2 m1 = *reinterpret_cast<const float2 *>(s_ap + plane + (c * 4 + t) * 2);
        const float2 z  = *reinterpret_cast<const float2 *>(s_ap + 2 * plane + (c * 4 + t) * 2);
        const float sa0 = (float)s[0], sb0 = (float)s[2];
        // d0*(d1*P + m1*S0) + m0*Z
        acc[0] += fmaf(f.sa, fmaf(d1.x, (float)d[0], m1.x * sa0), f.ma * z.x);
        acc[1] += fmaf(f.sa, fmaf(d1.y, (float)d[1], m1.y * sa0), f.ma * z.y);
        acc[2] += fmaf(f.sb, fmaf(d1.x, (float)d[2], m1.x * sb0), f.mb * z.x);
        acc[3] += fmaf(f.sb, fmaf(d1.y, (float)d[3], m1.y * sb0), f.mb * z.y);
    }
}

// ---- MP == 8: one IMMA per block, columns = activation rows ---------------------------------------
template <int QT>
__device__ __forceinline__ void compute_unit_m8(const Frag<QT> &f, int c, const uint8_t *s_aq, const float *s_ap,
                                                int plane, int lane, uint32_t kmask, uint32_t ksign, float (&acc)[4])
{
    const int t = lane & 3, quad = lane & 28;
    const int zero[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < CHUNK_BLKS; i++) {
        uint32_t a[4];
        make_a<QT>(a, u4_get(f.qa, i), u4_get(f.qb, i), kmask, ksign);
        const int blk = c * 4 + i;
        const uint2 b = *reinterpret_cast<const uint2 *>(s_aq + blk * 256 + lane * 8);
        const float2 d1 = *reinterpret_cast<const float2 *>(s_ap + blk * 8 + 2 * t);
        const float d0a = __shfl_sync(0xffffffffu, f.sa, quad | i), d0b = __shfl_sync(0xffffffffu, f.sb, quad | i);
        int d[4];
        imma<QT>(d, a, b.x, b.y, zero);
        if constexpr (QT == 0) {
            acc[0] = fmaf(d0a * d1.x, (float)d[0], acc[0]);
            acc[1] = fmaf(d0a * d1.y, (float)d[1], acc[1]);
            acc[2] = fmaf(d0b * d1.x, (float)d[2], acc[2]);
            acc[3] = fmaf(d0b * d1.y, (float)d[3], acc[3]);
        } else {
            int s[4];
            imma<QT>(s, a, 0x01010101u, 0x01010101u, zero);
            const float2 m1 = *reinterpret_cast<const float2 *>(s_ap + plane + blk * 8 + 2 * t);
            const float2 z  = *reinterpret_cast<const float2 *>(s_ap + 2 * plane + blk * 8 + 2 * t);
            const float m0a = __shfl_sync(0xffffffffu, f.ma, quad | i), m0b = __shfl_sync(0xffffffffu, f.mb, quad | i);
            const float sa0 = (float)s[0], sb0 = (float)s[2];
            acc[0] += fmaf(d0a, fmaf(d1.x, (float)d[0], m1.x * sa0), m0a * z.x);
            acc[1] += fmaf(d0a, fmaf(d1.y, (float)d[1], m1.y * sa0), m0a * z.y);
            acc[2] += fmaf(d0b, fmaf(d1.x, (float)d[2], m1.x * sb0), m0b * z.x);
            acc[3] += fmaf(d0b, fmaf(d1.y, (float)d[3], m1.y * sb0), m0b * z.y);
        }
    }
}

__device__ __forceinline__ int unit_start(int w, int U, int Weff) { return (int)(((long)w * U) / Weff); }
__device__ __forceinline__ int warp_of_unit(int u, int U, int Weff) { return (int)((((long)u + 1) * Weff - 1) / U); }

// End of a warp's segment [s0,s1) of tile T: combine with the other warps that share the
// tile (fixed order) and store Y.  acc: MP==2: lane (g,t) holds rows g/g+8 x m=0/1 summed over
// blocks == t (mod 4) -> quad-reduced here; MP==8: rows g/g+8 x m=2t/2t+1.
// Arguments by value: keeps the caller's accumulators in registers (no local-memory round trip).
template <int QT, int MP>
__device__ __noinline__ void flush_tile(const GemvParams &p, int T, int s0, int s1, int wid, int lane,
                                        float a0, float a1, float a2, float a3)
{
    float acc[4] = {a0, a1, a2, a3};
    if (MP == 2) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 1);
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 2);
        }
    }
    const int tb = T * p.C, te = tb + p.C;
    if (!(s0 == tb && s1 == te)) {
        const int wf = warp_of_unit(tb, p.U, p.Weff), wl = warp_of_unit(te - 1, p.U, p.Weff);
        const int slot = 2 * wid + (tb <= unit_start(wid, p.U, p.Weff) ? 0 : 1);
        float *ws = p.partial + (size_t)slot * 128;
#pragma unroll
        for (int r = 0; r < 4; r++) ws[r * 32 + lane] = acc[r];
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(p.counters + T, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != wl - wf) return;                 // not the last contributor
        __threadfence();
#pragma unroll
        for (int r = 0; r < 4; r++) acc[r] = 0.0f;
        for (int w = wf; w <= wl; w++) {            // fixed order: reproducible rounding
            const int sl = 2 * w + (tb <= unit_start(w, p.U, p.Weff) ? 0 : 1);
            const float *rs = p.partial + (size_t)sl * 128;
#pragma unroll
            for (int r = 0; r < 4; r++) acc[r] += __ldcg(rs + r * 32 + lane);
        }
        if (lane == 0) p.counters[T] = 0;           // ready for the next launch
    }
    const float post = (QT == 0) ? 0.0625f : 1.0f;  // undo the 16x of the shifted-nibble A operand (exact)
    const int g = lane >> 2, t = lane & 3;
    const int n0 = T * TILE_ROWS + g, n1 = n0 + 8;
    int m0, m1;
    if (MP == 2) { if (t != 0) return; m0 = 0; m1 = 1; }
    else { m0 = 2 * t; m1 = m0 + 1; }
    if (m0 < p.M) {
        if (n0 < p.N) p.y[(size_t)m0 * p.y_stride + n0] = acc[0] * post;
        if (n1 < p.N) p.y[(size_t)m0 * p.y_stride + n1] = acc[2] * post;
    }
    if (m1 < p.M) {
        if (n0 < p.N) p.y[(size_t)m1 * p.y_stride + n0] = acc[1] * post;
        if (n1 < p.N) p.y[(size_t)m1 * p.y_stride + n1] = acc[3] * post;
    }
}

template <int QT, int MP, int PD, bool FUSED>
__global__ void __launch_bounds__(GEMV_THREADS, 1) q4_gemv_kernel(const GemvParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int CH = chunk_bytes(QT);
    constexpr int NP = (QT == 0) ? 1 : 3;
    const int nbp = p.C * CHUNK_BLKS;
    const int plane = nbp * MP;                         // floats per parameter plane
    const size_t aq_bytes = (size_t)nbp * 8 * 32;       // shared-memory codes are always the 8-column form
    uint8_t *s_aq = smem;
    float *s_ap = reinterpret_cast<float *>(smem + aq_bytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * GEMV_WARPS + warp;
    const bool active = wid < p.Weff;
    int u0 = 0, u1 = 0;
    Frag<QT> buf[PD];
    if (active) {                                    // start the weight stream before staging activations
        u0 = unit_start(wid, p.U, p.Weff);
        u1 = unit_start(wid + 1, p.U, p.Weff);
#pragma unroll
        for (int i = 0; i < PD; i++)
            if (u0 + i < u1) load_frag<QT>(buf[i], p.wt + (size_t)(u0 + i) * CH, lane);
    }
    if constexpr (MP == 2) {
        // zero the whole pack (columns of other blocks, padded rows/blocks must read as 0)
        const size_t total16 = (aq_bytes + (size_t)NP * plane * 4) / 16;
        for (size_t i = threadIdx.x; i < total16; i += GEMV_THREADS)
            reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if constexpr (FUSED) {
            // quantize the activation rows straight into shared memory (every CTA redundantly:
            // M*K*4 bytes from L2, no second kernel, no global round trip)
            const int items = p.M * p.nb * 4;                 // one item = 8 consecutive elements
            for (int base = 0; base < items; base += GEMV_THREADS) {
                const int id = base + (int)threadIdx.x;
                const bool valid = id < items;
                const int bq = valid ? (id >> 2) : 0;          // (m, blk) index
                const int m = bq / p.nb, blk = bq - m * p.nb;
                quantize8_emit<QT>(p.x + (size_t)m * p.x_stride + (size_t)blk * QK + (lane & 3) * 8, valid, lane,
                                   s_aq + ((size_t)blk * 8 + 2 * (blk & 3) + m) * 32, s_ap + (size_t)blk * 2 + m, plane);
            }
        } else {
            // global apack [nbp][2][32] -> 8-column form; planes copied as they are
            const uint4 *src = reinterpret_cast<const uint4 *>(p.aq);
            for (int i = threadIdx.x; i < nbp * 4; i += GEMV_THREADS) {       // 16-byte pieces
                const int blk = i >> 2, m = (i >> 1) & 1, piece = i & 1;
                reinterpret_cast<uint4 *>(s_aq + ((size_t)blk * 8 + 2 * (blk & 3) + m) * 32)[piece] = src[i];
            }
            const uint4 *src2 = reinterpret_cast<const uint4 *>(p.ap);
            uint4 *dst2 = reinterpret_cast<uint4 *>(s_ap);
            for (size_t i = threadIdx.x; i < (size_t)NP * plane * 4 / 16; i += GEMV_THREADS) dst2[i] = src2[i];
        }
    } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.aq);
        uint4 *dst = reinterpret_cast<uint4 *>(s_aq);
        for (size_t i = threadIdx.x; i < aq_bytes / 16; i += GEMV_THREADS) dst[i] = src[i];
        const uint4 *src2 = reinterpret_cast<const uint4 *>(p.ap);
        uint4 *dst2 = reinterpret_cast<uint4 *>(s_ap);
        for (size_t i = threadIdx.x; i < (size_t)NP * plane * 4 / 16; i += GEMV_THREADS) dst2[i] = src2[i];
    }
    __syncthreads();
    if (!active) return;

    // nibble-split constants kept in registers (see and_xor)
    uint32_t kmask = 0xF0F0F0F0u, ksign = 0x80808080u;
    asm volatile("" : "+r"(kmask), "+r"(ksign));

    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int T = u0 / p.C;
    int c = u0 - T * p.C;
    int seg0 = u0;
    const uint8_t *wp = p.wt + (size_t)(u0 + PD) * CH;
    for (int ub = u0; ub < u1; ub += PD) {
#pragma unroll
        for (int i = 0; i < PD; i++) {
            const int u = ub + i;
            if (u < u1) {
                if constexpr (MP == 2) compute_unit_chain4<QT>(buf[i], c, s_aq, s_ap, plane, lane, kmask, ksign, acc);
                else                   compute_unit_m8<QT>(buf[i], c, s_aq, s_ap, plane, lane, kmask, ksign, acc);
                if (u + PD < u1) load_frag<QT>(buf[i], wp, lane);
                wp += CH;
                ++c;
                if (c == p.C || u + 1 == u1) {
                    flush_tile<QT, MP>(p, T, seg0, u + 1, wid, lane, acc[0], acc[1], acc[2], acc[3]);
                    acc[0] = acc[1] = acc[2] = acc[3] = 0.0f;
                    if (c == p.C) { c = 0; ++T; }
                    seg0 = u + 1;
                }
            }
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
//                   bulk copy, 16 units = 20/24 KB) per stage, completion on an mbarrier.
//   16 consumer     warp w takes unit w of the stage: 4 LDS for the A-fragment words and scales,
//   warps           3 ALU/word nibble split, IMMA against the B fragments of the activation pack,
//                   scale FMAs, then ONE 16-byte store per lane of its per-unit partial sums.
//                   Stateless between stages.
//   reducer warp    adds the per-unit partials of a stage in unit order (fixed -> bit-reproducible),
//                   carries the tile accumulator across stages, and at the end of a tile (or tile
//                   piece) stores Y, or -- when a tile is split over several tasks (K > 8192) --
//                   parks the piece in a workspace slot; the last piece to arrive (atomic counter)
//                   adds the slots in piece order and stores Y.
//
// Bytes in flight per SM = (stages-1) x 20 KB, independent of registers; weights never touch the
// register file before they are needed.  Programmatic dependent launch: the weight stream of launch
// i+1 starts while launch i drains; everything that depends on the previous kernel (X, Y, workspace,
// ticket counter) is touched only after cudaGridDependencySynchronize().
// ---------------------------------------------------------------------------
constexpr int PIPE_CONSUMERS = 16;                      // consumer warps = units per stage
constexpr int PIPE_THREADS = (PIPE_CONSUMERS + 2) * 32; // + producer warp + reducer warp
constexpr int PIPE_MAX_STAGES = 8;

struct PipeParams {
    const uint8_t *wt;
    const float   *x;        // fused quantization (MP == 2)
    const uint8_t *aq;       // global apack (non-fused)
    const float   *ap;
    float         *y;
    float         *partial;  // [ntasks][128] floats, used when a tile spans several tasks
    int           *tile_counters;
    int           *ticket;   // [2]: next task, CTAs done   (self-resetting)
    size_t         x_stride, y_stride;   // floats
    int            U, N, M, C, n_tiles, nb;
    int            ntasks;   // tasks in the launch
    int            PP;       // pieces (tasks) per tile when C >= 16, else 1
    int            TS;       // units per task
    int            stages;
};

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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// first unit and unit count of task t
__device__ __forceinline__ void task_range(const PipeParams &p, int t, int &u0, int &n)
{
    if (p.C >= PIPE_CONSUMERS) {                         // PP pieces per tile, TS units each (last shorter)
        const int T = t / p.PP, piece = t - T * p.PP;
        const int off = piece * p.TS;
        u0 = T * p.C + off;
        n = min(p.TS, p.C - off);
    } else {                                            // short tiles: TS = whole tiles
        u0 = t * p.TS;
        n = min(p.TS, p.U - u0);
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

// MP == 2 on the compact pack [blk][2][32]: the B fragment of (block 4c + g/2, row g&1) is
// lane-linear (c*256 + lane*8); `bmask[i]` is all-ones in the lanes whose block is 4c+i, so
// block i's IMMA sees that block's codes in columns 2i, 2i+1 and zeros elsewhere.
template <int QT>
__device__ __forceinline__ void compute_unit_chain4c(const Frag<QT> &f, int c, const uint8_t *s_aq, const float *s_ap,
                                                     int plane, int lane, uint32_t kmask, uint32_t ksign,
                                                     const uint32_t (&bmask)[4], float (&acc)[4])
{
    const int g = lane >> 2, t = lane & 3;
    const uint2 breg = *reinterpret_cast<const uint2 *>(s_aq + c * 256 + lane * 8);
    const float2 d1 = *reinterpret_cast<const float2 *>(s_ap + (c * 4 + t) * 2);
    int d[4] = {0, 0, 0, 0};
    int s[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < CHUNK_BLKS; i++) {
        uint32_t a[4];
        make_a<QT>(a, u4_get(f.qa, i), u4_get(f.qb, i), kmask, ksign);
        imma<QT>(d, a, breg.x & bmask[i], breg.y & bmask[i], d);
        if constexpr (QT == 1) {
            const uint32_t one = (g == 2 * i) ? 0x01010101u : 0u;
            imma<QT>(s, a, one, one, s);
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

// Reducer: end of the run of units [.., u_last] of tile T inside task `task`.
template <int QT, int MP>
__device__ __forceinline__ void pipe_emit(const PipeParams &p, int T, int task, int lane, float (&acc)[4])
{
    if (MP == 2) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 1);
            acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 2);
        }
    }
    if (p.C >= PIPE_CONSUMERS && p.PP > 1) {            // the tile is split over PP tasks
        float *ws = p.partial + (size_t)task * 128;
#pragma unroll
        for (int r = 0; r < 4; r++) ws[r * 32 + lane] = acc[r];
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(p.tile_counters + T, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != p.PP - 1) return;
        __threadfence();
#pragma unroll
        for (int r = 0; r < 4; r++) acc[r] = 0.0f;
        for (int k = 0; k < p.PP; k++) {                // piece order: reproducible rounding
            const float *rs = p.partial + (size_t)(T * p.PP + k) * 128;
#pragma unroll
            for (int r = 0; r < 4; r++) acc[r] += __ldcg(rs + r * 32 + lane);
        }
        if (lane == 0) p.tile_counters[T] = 0;
    }
    const float post = (QT == 0) ? 0.0625f : 1.0f;
    const int g = lane >> 2, t = lane & 3;
    const int n0 = T * TILE_ROWS + g, n1 = n0 + 8;
    int m0, m1;
    if (MP == 2) { if (t != 0) return; m0 = 0; m1 = 1; }
    else { m0 = 2 * t; m1 = m0 + 1; }
    if (m0 < p.M) {
        if (n0 < p.N) p.y[(size_t)m0 * p.y_stride + n0] = acc[0] * post;
        if (n1 < p.N) p.y[(size_t)m0 * p.y_stride + n1] = acc[2] * post;
    }
    if (m1 < p.M) {
        if (n0 < p.N) p.y[(size_t)m1 * p.y_stride + n0] = acc[1] * post;
        if (n1 < p.N) p.y[(size_t)m1 * p.y_stride + n1] = acc[3] * post;
    }
}

template <int QT, int MP, bool FUSED>
__global__ void __launch_bounds__(PIPE_THREADS, 1) q4_gemv_pipe_kernel(const PipeParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int CH = chunk_bytes(QT);
    constexpr int NP = (QT == 0) ? 1 : 3;
    constexpr int STAGE_BYTES = PIPE_CONSUMERS * CH;
    const int nbp = p.C * CHUNK_BLKS;
    const int plane = nbp * MP;
    // shared-memory map: [stages x STAGE_BYTES][red: stages x 16 x 512 B][apack codes][planes][barriers][desc]
    uint8_t *s_stage = smem;
    float   *s_red = reinterpret_cast<float *>(smem + (size_t)p.stages * STAGE_BYTES);
    uint8_t *s_aq = reinterpret_cast<uint8_t *>(s_red) + (size_t)p.stages * PIPE_CONSUMERS * 512;
    float   *s_ap = reinterpret_cast<float *>(s_aq + (size_t)nbp * MP * 32);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_ap) + (((size_t)NP * plane * 4 + 15) & ~(size_t)15));
    uint64_t *bar_empty = bar_full + PIPE_MAX_STAGES;
    uint64_t *bar_done = bar_empty + PIPE_MAX_STAGES;       // consumers finished a stage's partials
    int4     *s_desc = reinterpret_cast<int4 *>(bar_done + PIPE_MAX_STAGES);   // {first unit, n units (<0 = stop), chunk index of first unit, task}

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; i++) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], PIPE_CONSUMERS + 1);
            mbar_init(&bar_done[i], PIPE_CONSUMERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cudaTriggerProgrammaticLaunchCompletion();           // let the next launch start streaming early

    if (warp == PIPE_CONSUMERS) {
        // ===================== producer =====================
        if (lane == 0) {
            int it = 0;
            int task = blockIdx.x;                       // first task is static: no dependency on the previous launch
            bool synced = false;
            while (task < p.ntasks) {
                int u0, n;
                task_range(p, task, u0, n);
                int c0 = u0 % p.C;
                for (int r = 0; r < n; r += PIPE_CONSUMERS, ++it) {
                    const int s = it % p.stages;
                    if (it >= p.stages) mbar_wait(&bar_empty[s], ((it / p.stages) - 1) & 1);
                    const int cnt = min(PIPE_CONSUMERS, n - r);
                    s_desc[s] = make_int4(u0 + r, cnt, c0, task);
                    c0 = (c0 + cnt) % p.C;
                    mbar_arrive_expect_tx(&bar_full[s], (uint32_t)cnt * CH);
                    bulk_g2s(s_stage + (size_t)s * STAGE_BYTES, p.wt + (size_t)(u0 + r) * CH, (uint32_t)cnt * CH, &bar_full[s]);
                }
                if (!synced) { cudaGridDependencySynchronize(); synced = true; }
                task = gridDim.x + atomicAdd(p.ticket, 1);
            }
            if (!synced) cudaGridDependencySynchronize();
            {   // stop marker
                const int s = it % p.stages;
                if (it >= p.stages) mbar_wait(&bar_empty[s], ((it / p.stages) - 1) & 1);
                s_desc[s] = make_int4(0, -1, 0, 0);
                mbar_arrive(&bar_full[s]);
            }
            // the last CTA to run out of tasks re-arms the ticket counter for the next launch
            if (atomicAdd(p.ticket + 1, 1) == (int)gridDim.x - 1) { p.ticket[0] = 0; p.ticket[1] = 0; __threadfence(); }
        }
        return;
    }

    if (warp == PIPE_CONSUMERS + 1) {
        // ===================== reducer =====================
        cudaGridDependencySynchronize();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int curT = -1, curTask = -1;
        for (int it = 0;; ++it) {
            const int s = it % p.stages;
            mbar_wait(&bar_full[s], (it / p.stages) & 1);
            const int4 dsc = s_desc[s];
            if (dsc.y < 0) break;
            const int task = dsc.w;
            int T = (dsc.x - dsc.z) / p.C;               // one division per stage
            int left = p.C - dsc.z;                      // units of tile T from dsc.x on
            mbar_wait(&bar_done[s], (it / p.stages) & 1);
            const float *red = s_red + (size_t)s * PIPE_CONSUMERS * 128 + lane * 4;
            int j = 0;
            while (j < dsc.y) {
                if (T != curT || task != curTask) {
                    if (curT >= 0) pipe_emit<QT, MP>(p, curT, curTask, lane, acc);
                    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                    curT = T; curTask = task;
                }
                const int run = min(dsc.y - j, left);    // units of this stage that belong to tile T
                if (run == PIPE_CONSUMERS) {             // common case: the whole stage is one tile
                    float4 v[PIPE_CONSUMERS];
#pragma unroll
                    for (int k = 0; k < PIPE_CONSUMERS; k++) v[k] = *reinterpret_cast<const float4 *>(red + k * 128);
#pragma unroll
                    for (int k = 0; k < PIPE_CONSUMERS; k++) { acc[0] += v[k].x; acc[1] += v[k].y; acc[2] += v[k].z; acc[3] += v[k].w; }
                } else {
                    for (int k = j; k < j + run; k++) {
                        const float4 v = *reinterpret_cast<const float4 *>(red + k * 128);
                        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                    }
                }
                j += run;
                left -= run;
                if (left == 0) { ++T; left = p.C; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[s]);
        }
        if (curT >= 0) pipe_emit<QT, MP>(p, curT, curTask, lane, acc);
        return;
    }

    // ===================== consumers =====================
    cudaGridDependencySynchronize();                     // X may be the previous kernel's output
    const int ctid = threadIdx.x;                        // 0 .. 511
    constexpr int CTHREADS = PIPE_CONSUMERS * 32;
    if constexpr (FUSED) {
        // padded blocks (nb .. nbp) and the unused row of an M=1 call must read as zero
        const size_t total16 = ((size_t)nbp * MP * 32 + (((size_t)NP * plane * 4 + 15) & ~(size_t)15)) / 16;
        for (size_t i = ctid; i < total16; i += CTHREADS) reinterpret_cast<uint4 *>(s_aq)[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("bar.sync 1, %0;" :: "n"(CTHREADS) : "memory");
        const int items = p.M * p.nb * 4;
        for (int base = 0; base < items; base += CTHREADS) {
            const int id = base + ctid;
            const bool valid = id < items;
            const int bq = valid ? (id >> 2) : 0;
            const int m = bq / p.nb, blk = bq - m * p.nb;
            quantize8_emit<QT>(p.x + (size_t)m * p.x_stride + (size_t)blk * QK + (lane & 3) * 8, valid, lane,
                               s_aq + ((size_t)blk * MP + m) * 32, s_ap + (size_t)blk * MP + m, plane);
        }
    } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.aq);
        uint4 *dst = reinterpret_cast<uint4 *>(s_aq);
        for (size_t i = ctid; i < (size_t)nbp * MP * 32 / 16; i += CTHREADS) dst[i] = src[i];
        const uint4 *src2 = reinterpret_cast<const uint4 *>(p.ap);
        uint4 *dst2 = reinterpret_cast<uint4 *>(s_ap);
        for (size_t i = ctid; i < (size_t)NP * plane * 4 / 16; i += CTHREADS) dst2[i] = src2[i];
    }
    asm volatile("bar.sync 1, %0;" :: "n"(CTHREADS) : "memory");

    uint32_t kmask = 0xF0F0F0F0u, ksign = 0x80808080u;
    asm volatile("" : "+r"(kmask), "+r"(ksign));
    uint32_t bmask[4];
#pragma unroll
    for (int i = 0; i < 4; i++) bmask[i] = ((lane >> 3) == i) ? 0xFFFFFFFFu : 0u;

    for (int it = 0;; ++it) {
        const int s = it % p.stages;
        mbar_wait(&bar_full[s], (it / p.stages) & 1);
        const int4 dsc = s_desc[s];
        if (dsc.y < 0) break;
        if (warp < dsc.y) {
            int c = dsc.z + warp;
            while (c >= p.C) c -= p.C;
            Frag<QT> f;
            load_frag_smem<QT>(f, s_stage + (size_t)s * STAGE_BYTES + (size_t)warp * CH, lane);
            float acc[4];
            if constexpr (MP == 2) compute_unit_chain4c<QT>(f, c, s_aq, s_ap, plane, lane, kmask, ksign, bmask, acc);
            else {
                acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
                compute_unit_m8<QT>(f, c, s_aq, s_ap, plane, lane, kmask, ksign, acc);
            }
            *reinterpret_cast<float4 *>(s_red + ((size_t)s * PIPE_CONSUMERS + warp) * 128 + lane * 4) =
                make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bar_done[s]); mbar_arrive(&bar_empty[s]); }
    }
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
