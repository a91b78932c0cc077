// q4_kernels.cuh -- hand-written sm_100a kernels for vsim's Q4_0/Q4_1 mul_mat path.
//
// Reference semantics (NAIST-Archlab/vsim): ggml.c:209-298 (activation-row
// quantizers), ggml.c:472-553 / imax.c:1205-1229 (q4 x q4 block dot products),
// ggml.c:4891-5453 (the operator).  Nothing here is translated from those
// loops; the formulation is:
//
//   Y[m][n] = sum_blk d0[n,blk] * d1[m,blk] * ISUM[m,n,blk]                 (Q4_0)
//   ISUM    = sum_{k in blk} (q0-8)(q1-8) = sum q0*a - 8*sum a,  a = q1-8    (exact, int32)
//   Y[m][n] = sum_blk d0*(d1*P + m1*S0) + m0*Z                              (Q4_1)
//   P = sum q0*q1, S0 = sum q0, Z = d1*sum q1 + 32*m1                       (P, S0 exact)
//
// i.e. the 32-element integer block sums are computed exactly on the tensor
// pipe (mma.sync m16n8k32 u8 x s8 -> s32: one instruction = 16 weight rows x
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
// Activation pack ("apack", rebuilt per call by pack_act_kernel, ~1 us):
//   aq[blk][m][32 B]: the 32 codes of block blk of row m as bytes (s8 = q-8 for
//                     Q4_0, u8 = q for Q4_1), ordered so that an 8-byte read at
//                     +8t is the mma B fragment of lane (g=m%8, t);
//   ap[blk][m]:       Q4_0 {d1, -8*sum(a) as int bits}; Q4_1 {d1, m1, Z, 0}.
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

// Writes one block's apack record (32 code bytes + parameters); aq/ap may be global or shared.
template <int QT>
__device__ __forceinline__ void emit_apack(const BlockQ &q, int lane, uint8_t *aq_rec, float *ap_rec)
{
    const int a = (QT == 0) ? (q.code & 15) - 8 : (q.code & 15);
    aq_rec[apack_byte_of(lane)] = (uint8_t)(int8_t)a;
    int s = a;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (QT == 0) {
            ap_rec[0] = q.d;
            ap_rec[1] = __int_as_float(-8 * s);
        } else {
            ap_rec[0] = q.d;
            ap_rec[1] = q.mn;
            ap_rec[2] = fmaf(q.d, (float)s, 32.0f * q.mn);     // Z = sum_k dequant(a_k)
            ap_rec[3] = 0.0f;
        }
    }
}

// fp32 rows -> apack.  grid: nbp*M warps.  Blocks >= nb (padding up to a multiple of 4) are zero.
template <int QT>
__global__ void pack_act_kernel(const float *__restrict__ x, size_t x_stride, int M, int nb, int nbp,
                                uint8_t *__restrict__ aq, float *__restrict__ ap)
{
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long)nbp * M) return;
    const int blk = (int)(w / M), m = (int)(w - (long)blk * M);
    const float xv = (blk < nb) ? x[(size_t)m * x_stride + (size_t)blk * QK + lane] : 0.0f;
    BlockQ q = quantize_block<QT>(xv);
    if (blk >= nb) { q.d = 0.0f; q.mn = 0.0f; q.code = (QT == 0) ? 8 : 0; }
    emit_apack<QT>(q, lane, aq + ((size_t)blk * M + m) * 32, ap + ((size_t)blk * M + m) * ap_floats(QT));
}

// reference-format quantized rows (params->wdata image) -> apack  (unpatched-ggml hook path)
template <int QT>
__global__ void pack_wdata_kernel(const uint8_t *__restrict__ wdata, int M, int nb, int nbp,
                                  uint8_t *__restrict__ aq, float *__restrict__ ap)
{
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long)nbp * M) return;
    const int blk = (int)(w / M), m = (int)(w - (long)blk * M);
    BlockQ q;
    q.d = 0.0f; q.mn = 0.0f; q.code = (QT == 0) ? 8 : 0;
    if (blk < nb) {
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
    emit_apack<QT>(q, lane, aq + ((size_t)blk * M + m) * 32, ap + ((size_t)blk * M + m) * ap_floats(QT));
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
// The decode kernel: dequant-fused skinny GEMM, M <= 8*NT activation rows.
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
// Per unit a lane issues 2 x LDG.128 (nibbles) + 2 (4 for Q4_1) x LDG.128
// (scales), PD units ahead of use (register ring), then per quant block:
// 4 LOP3/SHF to split nibbles into the u8 A fragment, one 8-byte LDS for the
// B fragment, one IMMA, and 4 scale FMAs.
// ---------------------------------------------------------------------------
struct GemvParams {
    const uint8_t *wt;      // wtile
    const uint8_t *aq;      // apack codes  [nbp][M][32]
    const float   *ap;      // apack params [nbp][M][2|4]
    float         *y;
    float         *partial; // [2*Weff][NT*4*32]
    int           *counters;// [n_tiles], zero between launches
    size_t         y_stride;// floats
    long           U;       // n_tiles * C
    int            N, M, C, n_tiles;
    int            Weff;    // warps that own a non-empty unit range
};

template <int QT> struct Frag;
// sa/sb: lane (g,t) keeps d0 of rows g / g+8 for block 4c+t only; the other three blocks'
// scales live in the neighbouring lanes of the quad and are fetched by shuffle at use.
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

__device__ __forceinline__ float f4_get(const float4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ uint32_t u4_get(const uint4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

template <int QT, int NT>
__device__ __forceinline__ void compute_unit(const Frag<QT> &f, int blk0, const uint8_t *s_aq, const float *s_ap,
                                             int M, int g, int t, float (&acc)[NT][4])
{
    const int quad = (g << 2);
#pragma unroll
    for (int i = 0; i < CHUNK_BLKS; i++) {
        const uint32_t wa = u4_get(f.qa, i), wb = u4_get(f.qb, i);
        const uint32_t a[4] = { wa & 0x0F0F0F0Fu, wb & 0x0F0F0F0Fu, (wa >> 4) & 0x0F0F0F0Fu, (wb >> 4) & 0x0F0F0F0Fu };
        const float d0a = __shfl_sync(0xffffffffu, f.sa, quad | i), d0b = __shfl_sync(0xffffffffu, f.sb, quad | i);
        const int blk = blk0 + i;
        int s0[4];
        if constexpr (QT == 1)      // S0[row] = sum of the row's 32 weight codes: A x ones
            mma_16x8x32<false>(s0, a, 0x01010101u, 0x01010101u, 0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < NT; j++) {
            const int mb = 8 * j + g;                 // activation row feeding this lane's B fragment
            uint2 b = make_uint2(0u, 0u);
            if (mb < M) b = *reinterpret_cast<const uint2 *>(s_aq + ((size_t)blk * M + mb) * 32 + t * 8);
            const int m0 = 8 * j + 2 * t, m1 = m0 + 1; // activation rows of this lane's D columns
            int d[4];
            if constexpr (QT == 0) {
                float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);
                if (m0 < M) p0 = *reinterpret_cast<const float2 *>(s_ap + ((size_t)blk * M + m0) * 2);
                if (m1 < M) p1 = *reinterpret_cast<const float2 *>(s_ap + ((size_t)blk * M + m1) * 2);
                const int c0 = __float_as_int(p0.y), c1 = __float_as_int(p1.y);
                mma_16x8x32<true>(d, a, b.x, b.y, c0, c1, c0, c1);
                acc[j][0] = fmaf(d0a * p0.x, (float)d[0], acc[j][0]);
                acc[j][1] = fmaf(d0a * p1.x, (float)d[1], acc[j][1]);
                acc[j][2] = fmaf(d0b * p0.x, (float)d[2], acc[j][2]);
                acc[j][3] = fmaf(d0b * p1.x, (float)d[3], acc[j][3]);
            } else {
                float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
                if (m0 < M) p0 = *reinterpret_cast<const float4 *>(s_ap + ((size_t)blk * M + m0) * 4);
                if (m1 < M) p1 = *reinterpret_cast<const float4 *>(s_ap + ((size_t)blk * M + m1) * 4);
                mma_16x8x32<false>(d, a, b.x, b.y, 0, 0, 0, 0);
                const float m0a = __shfl_sync(0xffffffffu, f.ma, quad | i), m0b = __shfl_sync(0xffffffffu, f.mb, quad | i);
                const float sa0 = (float)s0[0], sb0 = (float)s0[2];
                // d0*(d1*P + m1*S0) + m0*Z
                acc[j][0] += fmaf(d0a, fmaf(p0.x, (float)d[0], p0.y * sa0), m0a * p0.z);
                acc[j][1] += fmaf(d0a, fmaf(p1.x, (float)d[1], p1.y * sa0), m0a * p1.z);
                acc[j][2] += fmaf(d0b, fmaf(p0.x, (float)d[2], p0.y * sb0), m0b * p0.z);
                acc[j][3] += fmaf(d0b, fmaf(p1.x, (float)d[3], p1.y * sb0), m0b * p1.z);
            }
        }
    }
}

__device__ __forceinline__ long unit_start(long w, long U, int Weff) { return (w * U) / Weff; }
__device__ __forceinline__ int warp_of_unit(long u, long U, int Weff) { return (int)(((u + 1) * Weff - 1) / U); }

template <int NT>
__device__ __noinline__ void flush_tile(const GemvParams &p, int T, long s0, long s1, int wid, int lane,
                                        float (&acc)[NT][4])
{
    const long tb = (long)T * p.C, te = tb + p.C;
    if (!(s0 == tb && s1 == te)) {
        const int wf = warp_of_unit(tb, p.U, p.Weff), wl = warp_of_unit(te - 1, p.U, p.Weff);
        const int slot = 2 * wid + (tb <= unit_start(wid, p.U, p.Weff) ? 0 : 1);
        float *ws = p.partial + (size_t)slot * (NT * 4 * 32);
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) ws[(j * 4 + r) * 32 + lane] = acc[j][r];
        __threadfence();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(p.counters + T, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != wl - wf) return;                 // not the last contributor
        __threadfence();
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[j][r] = 0.0f;
        for (int w = wf; w <= wl; w++) {            // fixed order: reproducible rounding
            const int sl = 2 * w + (tb <= unit_start(w, p.U, p.Weff) ? 0 : 1);
            const float *rs = p.partial + (size_t)sl * (NT * 4 * 32);
#pragma unroll
            for (int j = 0; j < NT; j++)
#pragma unroll
                for (int r = 0; r < 4; r++) acc[j][r] += __ldcg(rs + (j * 4 + r) * 32 + lane);
        }
        if (lane == 0) p.counters[T] = 0;           // ready for the next launch
    }
    const int g = lane >> 2, t = lane & 3;
    const int n0 = T * TILE_ROWS + g, n1 = n0 + 8;
#pragma unroll
    for (int j = 0; j < NT; j++) {
        const int m0 = 8 * j + 2 * t, m1 = m0 + 1;
        if (m0 < p.M) {
            if (n0 < p.N) p.y[(size_t)m0 * p.y_stride + n0] = acc[j][0];
            if (n1 < p.N) p.y[(size_t)m0 * p.y_stride + n1] = acc[j][2];
        }
        if (m1 < p.M) {
            if (n0 < p.N) p.y[(size_t)m1 * p.y_stride + n0] = acc[j][1];
            if (n1 < p.N) p.y[(size_t)m1 * p.y_stride + n1] = acc[j][3];
        }
    }
}

template <int QT, int NT, int PD>
__global__ void __launch_bounds__(GEMV_THREADS, 1) q4_gemv_kernel(const GemvParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int CH = chunk_bytes(QT);
    constexpr int APF = ap_floats(QT);
    const int nbp = p.C * CHUNK_BLKS;
    const size_t aq_bytes = (size_t)nbp * p.M * 32;
    const size_t ap_bytes = (size_t)nbp * p.M * APF * 4;
    uint8_t *s_aq = smem;
    float *s_ap = reinterpret_cast<float *>(smem + aq_bytes);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wid = blockIdx.x * GEMV_WARPS + warp;
    const bool active = wid < p.Weff;
    long u0 = 0, u1 = 0;
    Frag<QT> buf[PD];
    if (active) {                                    // start the weight stream before staging activations
        u0 = unit_start(wid, p.U, p.Weff);
        u1 = unit_start(wid + 1, p.U, p.Weff);
#pragma unroll
        for (int i = 0; i < PD; i++)
            if (u0 + i < u1) load_frag<QT>(buf[i], p.wt + (size_t)(u0 + i) * CH, lane);
    }
    {   // apack: global (L2) -> shared, 16 B per thread per step
        const uint4 *src = reinterpret_cast<const uint4 *>(p.aq);
        uint4 *dst = reinterpret_cast<uint4 *>(s_aq);
        for (size_t i = threadIdx.x; i < aq_bytes / 16; i += GEMV_THREADS) dst[i] = src[i];
        const uint4 *src2 = reinterpret_cast<const uint4 *>(p.ap);
        uint4 *dst2 = reinterpret_cast<uint4 *>(s_ap);
        for (size_t i = threadIdx.x; i < ap_bytes / 16; i += GEMV_THREADS) dst2[i] = src2[i];
    }
    __syncthreads();
    if (!active) return;

    const int g = lane >> 2, t = lane & 3;
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
        for (int r = 0; r < 4; r++) acc[j][r] = 0.0f;

    int T = (int)(u0 / p.C);
    int c = (int)(u0 - (long)T * p.C);
    long seg0 = u0;
    for (long ub = u0; ub < u1; ub += PD) {
#pragma unroll
        for (int i = 0; i < PD; i++) {
            const long u = ub + i;
            if (u < u1) {
                const Frag<QT> cur = buf[i];
                if (u + PD < u1) load_frag<QT>(buf[i], p.wt + (size_t)(u + PD) * CH, lane);
                compute_unit<QT, NT>(cur, c * CHUNK_BLKS, s_aq, s_ap, p.M, g, t, acc);
                ++c;
                if (c == p.C || u + 1 == u1) {
                    flush_tile<NT>(p, T, seg0, u + 1, wid, lane, acc);
#pragma unroll
                    for (int j = 0; j < NT; j++)
#pragma unroll
                        for (int r = 0; r < 4; r++) acc[j][r] = 0.0f;
                    if (c == p.C) { c = 0; ++T; }
                    seg0 = u + 1;
                }
            }
        }
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
