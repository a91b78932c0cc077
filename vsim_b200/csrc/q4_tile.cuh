// q4_tile.cuh -- the batched (M >= 3) Q4_0/Q4_1 mul_mat kernel for sm_100a: TMA-staged weight tiles,
// nibble -> 8-bit expansion into shared memory, tcgen05 MMA with TMEM accumulators; the per-block
// fp32 scale products are ALSO formed on the tensor core, so the CUDA cores do one FFMA per
// (m, n, quant block) with both operands read from TMEM.
//
// Reference semantics: ggml_compute_forward_mul_mat_q4_{0,1}_f32 (ggml.c:4891-5453) with the block
// dot products of ggml.c:472-553 / imax.c:1205-1229.  Same exact-integer formulation as the decode
// kernel (q4_kernels.cuh):
//
//   Q4_0:  Y[m][n] = sum_blk (d0[n,blk] * d1[m,blk]) * I,    I = sum_k (q0-8)(q1-8)           (I exact)
//   Q4_1:  Y[m][n] = sum_blk (d0*d1) * I + (d0*S0)*m1 + m0*Z   with the codes CENTRED, d*q + m = d*(q-8) + (m + 8d):
//                    I as above, S0 = sum_k (q0-8), m0 := m0 + 8*d0, m1 := m1 + 8*d1, Z = d1*sum(q1-8) + 32*m1
//                    (I, S0 exact).  Centring keeps every term small; with raw codes d0*d1*P and d0*S0*m1 are
//                    two large numbers that nearly cancel in every block.
//
// Per quant block (K = 32) and per 128 weight rows x MT activation rows the tensor core runs:
//   1. V  = A x B      kind::f8f6f4, fresh accumulator.  A = the weight nibbles: a nibble q0 in [0,15]
//                      left as the byte 0x0q IS the e4m3 number q0 * 2^-9 (the e4m3 subnormals 1..7 * 2^-9
//                      continue linearly into the first binade 8..15 * 2^-9), so "dequantizing" is one AND
//                      (low nibbles) or shift + AND (high nibbles) per 4 elements.  B = the activation codes,
//                      written by our pack kernel as e4m3 bytes of (q1-8)*2^5.
//                      Every product is a small multiple of 2^-4; 32 of them sum exactly in fp32.
//   2. V += A8 x B     A8 = a constant tile of -8 * 2^-9: restores the "-8" of the weight codes,
//                      V = I / 16 exactly.
//   3. T  = A3 x B3    kind::f16 (bf16, K = 16): d0 = a1+a2+a3 and d1 = b1+b2+b3 split exactly into three
//                      bf16 pieces each; the nine K slots hold (a_i, b_j), so T[n][m] = d0[n] * d1[m] to
//                      fp32 accuracy (products exact, one fp32 accumulation), replicated over the tile by
//                      the MMA itself -- this is the "broadcast" that shared memory is too narrow for
//                      (8 B of scales per value per lane through a 128 B/cycle LSU bounded the first
//                      version of this kernel at 1/4 of its FMA rate).
//   4. R  = A4 x B4    (Q4_1 only) the rank-2 term (d0*S0)*m1' + m0*Z' with bf16 pieces in 12 K slots, a
//                      fresh tile per block, added by the epilogue (a running sum kept on the tensor core
//                      drifts: its accumulate step truncates).
// and the epilogue warps do   acc[m] = fma(T[m], V[m], acc[m]) (+ R[m])   (packed FFMA2), x16 at the store.
// The kernel is bounded by TMEM read bandwidth (8 B per value; ~400 B/cycle/SM measured with
// tools/tmembench.cu) for large M, by the weight stream (HBM) for small M.
//
// Warp roles (640 threads where TMEM holds four buffers, else 512 with the two issuers in warps 1 and 2; one CTA per SM,
// persistent; register budgets moved with setmaxnreg):
//   warp 0      producer: cp.async.bulk (TMA 1-D) of weight chunks, 8 x 1280/1536 B per 128 rows x 4
//               blocks straight from the decode kernel's "wtile" layout (one weight copy in HBM serves
//               both kernels); full/empty mbarriers.
//   warp 3      producer of the activation stages (codes + bf16 scale pieces, already in operand layout).
//   warp 2      TMEM allocator (512 columns).
//   warps 4-7   converters: nibble words -> bytes in the UMMA canonical layout (K-major, no swizzle: core
//               matrix = 8 rows x 16 B), bf16 pieces of d0 (and d0*S0, m0).
//   warps 8-15  epilogue: tcgen05.ld x16 double-buffered, FFMA2, final store / stream-K combine.
//   warps 16-19 MMA issuers (one elected thread each, one TMEM buffer each), tcgen05.commit to the epilogue and back to the rings.
#pragma once
#include "q4_kernels.cuh"

namespace b200q4 {

constexpr int TK_ROWS = 128;                 // weight rows per tile (MMA M)
constexpr int TK_T16 = TK_ROWS / TILE_ROWS;  // wtile tiles per tile
constexpr int TK_CONV_WARPS = 4;
constexpr int TK_EPI_WARPS = 8;
constexpr int TK_ISSUERS = 4;                // MMA issuer warps where TMEM holds four buffers: a warp group of their own behind the epilogue warps
constexpr int TK_THREADS_BASE = (4 + TK_CONV_WARPS + TK_EPI_WARPS) * 32;    // two buffers: the two issuers are warps 1 and 2
constexpr int TK_THREADS_MAX = TK_THREADS_BASE + TK_ISSUERS * 32;
constexpr int TK_MAX_RAW = 16, TK_MAX_ACT = 8, TK_MAX_CONV = 8, TK_MAX_TM = 4;
// register budgets after setmaxnreg: 640 threads (launched with 96): 128*(2*32 + 72 + 2*168) = 60416 <= 61440;
// 512 threads (launched with 128): 128*(40 + 72 + 2*200) = 65536
__host__ __device__ constexpr int tk_regs_misc(int nthr) { return nthr > TK_THREADS_BASE ? 32 : 40; }
__host__ __device__ constexpr int tk_regs_epi(int nthr) { return nthr > TK_THREADS_BASE ? 168 : 200; }
constexpr int TK_REGS_CONV = 72;
constexpr int TK_OP_TILE = TK_ROWS * 32;     // one 128-row operand tile of one block (fp8 K=32 or bf16 K=16)

__host__ __device__ constexpr int tk_nop(int qt) { return qt == 0 ? 2 : 3; }          // operand tiles per block and side
__host__ __device__ constexpr int tk_max_mt(int qt) { return qt == 0 ? 128 : 64; }
__host__ __device__ constexpr int tk_act_stage(int qt, int mt) { return CHUNK_BLKS * mt * 32 * tk_nop(qt); }
__host__ __device__ constexpr int tk_raw_stage(int qt) { return TK_T16 * chunk_bytes(qt); }
__host__ __device__ constexpr int tk_conv_slot(int qt) { return CHUNK_BLKS * TK_OP_TILE * tk_nop(qt); }
__host__ __device__ constexpr int tk_bar_bytes() { return (2 * TK_MAX_RAW + 2 * TK_MAX_ACT + 2 * TK_MAX_CONV + 2 * TK_MAX_TM + 2) * 8 + 64; }
__host__ __device__ inline size_t tk_smem_bytes(int qt, int mt, int nraw, int nact, int nconv)
{
    return (size_t)TK_OP_TILE + (size_t)nraw * tk_raw_stage(qt) + (size_t)nact * tk_act_stage(qt, mt) +
           (size_t)nconv * tk_conv_slot(qt) + tk_bar_bytes();
}
// blocks per TMEM buffer; NBUF buffers of G x [V | T (| R)] x MT columns
// TMEM map.  One block's tiles [V | T (| R)] take NT x MT columns; where a quarter of TMEM cannot hold that (Q4_0 at
// MT = 128, Q4_1 at MT = 64) the tile is issued as two half-width MMAs (SPLIT) into buffers of their own, so that there
// are always FOUR buffers in flight: the issuer -> epilogue -> issuer round trip (commit, mbarrier, tcgen05.ld, arrive,
// and an issuing thread's mbarrier waits do not return before its own MMAs have finished: tools/mmabench.cu) paces
// the kernel, and what hides it is the number of hand-offs in flight, not their size.
#ifdef TK_SPLIT
__host__ __device__ constexpr bool tk_split(int qt, int mt) { return tk_nop(qt) * mt > 128; }
#else
__host__ __device__ constexpr bool tk_split(int qt, int mt) { return false; }           // measured: no gain (profiles/r02_tile_kernel_experiments.txt)
#endif
__host__ __device__ constexpr int tk_sw(int qt, int mt) { return tk_split(qt, mt) ? mt / 2 : mt; }      // MMA N
__host__ __device__ constexpr int tk_g(int qt, int mt)                                                   // blocks per buffer
{
    const int g = 128 / (tk_nop(qt) * tk_sw(qt, mt));
    return g < 1 ? 1 : g > CHUNK_BLKS ? CHUNK_BLKS : g;
}
__host__ __device__ constexpr int tk_nbuf(int qt, int mt)
{
    const int n = 512 / (tk_g(qt, mt) * tk_nop(qt) * tk_sw(qt, mt));
    return n > TK_MAX_TM ? TK_MAX_TM : n;
}

// TMEM hand-offs per chunk and issuer groups: with fewer hand-offs per chunk than issuers, a chunk belongs to ONE group of
// issuers (chunk cc -> group cc % tk_groups); when the operand rings' depths are multiples of the group count a slot always
// belongs to the same group, whose issuers then see every phase of its barriers in order and the others need not touch it.
__host__ __device__ constexpr int tk_steps(int qt, int mt) { return tk_split(qt, mt) ? 2 * CHUNK_BLKS : CHUNK_BLKS / tk_g(qt, mt); }
__host__ __device__ constexpr int tk_groups(int qt, int mt) { return tk_steps(qt, mt) >= tk_nbuf(qt, mt) ? 1 : tk_nbuf(qt, mt) / tk_steps(qt, mt); }
__host__ __device__ constexpr int tk_threads(int qt, int mt) { return tk_nbuf(qt, mt) == 4 ? TK_THREADS_MAX : TK_THREADS_BASE; }

struct TileParams {
    const uint8_t *wt;       // wtile
    const uint8_t *tpack;    // [m_tiles][C] activation stages
    float         *y;
    float         *partial;  // stream-K pieces: [grid][2][MT][128]
    int           *counters; // one per tile (n tile x m tile), self-resetting
    size_t         y_stride; // floats
    int            N, M, C, n_t16, m_tiles, ntiles;
    int            nraw, nact, nconv;
    unsigned long long *dbg; // B200Q4_TIMELINE builds: per-CTA wait-cycle counters, else NULL
};

// profiling aid (-DB200Q4_TIMELINE, tools/tile_waits.py): cycles a role spends inside each kind of wait
#ifdef B200Q4_TIMELINE
#define TK_WAIT(bar, ph, slot) do { const long long t0_ = clock64(); mbar_wait((bar), (ph)); tk_prof[(slot)] += clock64() - t0_; } while (0)
#define TK_PROF_DECL long long tk_prof[16] = {0}; const long long tk_start = clock64()
#define TK_PROF_FLUSH(total_slot, cond) do { if (p.dbg && (cond)) { tk_prof[(total_slot)] = clock64() - tk_start; \
        for (int i_ = 0; i_ < 16; i_++) if (tk_prof[i_]) p.dbg[(size_t)blockIdx.x * 16 + i_] = (unsigned long long)tk_prof[i_]; } } while (0)
#else
#define TK_WAIT(bar, ph, slot) mbar_wait((bar), (ph))
#define TK_PROF_DECL do { } while (0)
#define TK_PROF_FLUSH(total_slot, cond) do { } while (0)
#endif

// second profiling aid (-DTK_TRACE, tools/tile_trace.py): CTA 0 stamps the first 32 chunk steps of every role
#ifdef TK_TRACE
#define TK_STAMP(ev, step, cond) do { if (p.dbg && blockIdx.x == 0 && (step) < 32 && (cond)) p.dbg[(ev) * 32 + (step)] = (unsigned long long)clock64(); } while (0)
#else
#define TK_STAMP(ev, step, cond) do { } while (0)
#endif

// ---- e4m3 encoding of v * 2^s for an integer |v| <= 15 (exact) ----------------------------------
__device__ __forceinline__ uint32_t e4m3_of_int(int v, int s)
{
    if (v == 0) return 0u;
    const uint32_t sign = v < 0 ? 0x80u : 0u;
    const uint32_t a = (uint32_t)(v < 0 ? -v : v);
    const int k = 31 - __clz((int)a);                         // a = 2^k * (1 + M/8)
    const uint32_t man = (a << (3 - k)) & 7u;
    return sign | ((uint32_t)(7 + s + k) << 3) | man;
}

// x = h1 + h2 + h3 exactly, each a bf16 (returned as its 16 bits): 8 + 8 + 8 significand bits
__device__ __forceinline__ void bf16_split3(float x, uint32_t &h1, uint32_t &h2, uint32_t &h3)
{
    const uint32_t u = __float_as_uint(x);
    h1 = u >> 16;
    const float r = __fsub_rn(x, __uint_as_float(u & 0xFFFF0000u));           // exact
    const uint32_t ur = __float_as_uint(r);
    h2 = ur >> 16;
    const float r2 = __fsub_rn(r, __uint_as_float(ur & 0xFFFF0000u));         // exact, <= 8 significant bits
    h3 = __float_as_uint(r2) >> 16;
}
// K slots 0..7 of a rank-1 product x*y hold the piece pairs (1,1) (1,2) (1,3) (2,1) (2,2) (2,3) (3,1) (3,2)
// -- one 16-byte store per operand row; (3,3) is below 2^-32 of the product and dropped; slots 8..15
// (the second core matrix) stay zero.
__device__ __forceinline__ uint4 bf16_slots_left(float x)     // x_1 x_1 x_1 x_2 x_2 x_2 x_3 x_3
{
    uint32_t a1, a2, a3;
    bf16_split3(x, a1, a2, a3);
    return make_uint4(a1 | (a1 << 16), a1 | (a2 << 16), a2 | (a2 << 16), a3 | (a3 << 16));
}
__device__ __forceinline__ uint4 bf16_slots_right(float y)    // y_1 y_2 y_3 y_1 y_2 y_3 y_1 y_2
{
    uint32_t b1, b2, b3;
    bf16_split3(y, b1, b2, b3);
    return make_uint4(b1 | (b2 << 16), b3 | (b1 << 16), b2 | (b3 << 16), b1 | (b2 << 16));
}
// the rank-2 term x*y + z*w in 12 slots: pairs (1,1) (1,2) (2,1) (2,2) (1,3) (3,1) of each product
// (the dropped pairs are below 2^-23 of the product)
__device__ __forceinline__ void bf16_slots2_left(float x, float z, uint4 &lo, uint4 &hi)
{
    uint32_t a1, a2, a3, c1, c2, c3;
    bf16_split3(x, a1, a2, a3);
    bf16_split3(z, c1, c2, c3);
    lo = make_uint4(a1 | (a1 << 16), a2 | (a2 << 16), a1 | (a3 << 16), c1 | (c1 << 16));
    hi = make_uint4(c2 | (c2 << 16), c1 | (c3 << 16), 0u, 0u);
}
__device__ __forceinline__ void bf16_slots2_right(float y, float w, uint4 &lo, uint4 &hi)
{
    uint32_t b1, b2, b3, e1, e2, e3;
    bf16_split3(y, b1, b2, b3);
    bf16_split3(w, e1, e2, e3);
    lo = make_uint4(b1 | (b2 << 16), b1 | (b2 << 16), b3 | (b1 << 16), e1 | (e2 << 16));
    hi = make_uint4(e1 | (e2 << 16), e3 | (e1 << 16), 0u, 0u);
}

// K position (byte inside the block's 32-byte operand row) of element e: the low nibbles (even
// elements) of nibble word t form bytes 4t..4t+3, the high nibbles (odd elements) bytes 16+4t..16+4t+3,
// so the converter's two 32-bit stores per word are each one contiguous 128 B per warp.
__device__ __forceinline__ int tk_kpos(int e) { return ((e & 1) << 4) + ((e >> 3) << 2) + ((e & 7) >> 1); }

// Byte offset of (row r, K byte p) inside one block's operand tile in the UMMA canonical K-major
// no-swizzle layout: core matrix = 8 rows x 16 B contiguous; K-adjacent core matrices 128 B apart
// (LBO), 8-row groups 256 B apart (SBO).
__device__ __forceinline__ int tk_operand_off(int r, int p) { return (r >> 3) * 256 + (p >> 4) * 128 + (r & 7) * 16 + (p & 15); }

// ---- activation pack for the tile kernel ----------------------------------------------------------
// Stage (m tile, chunk) = [4 blocks x codes tile][4 blocks x scale-piece tile]([4 blocks x rank-2 tile]),
// every tile MT rows x 32 B in operand layout.  One THREAD per (stage row m_local, block), padding included: the
// block's 32 floats arrive as eight float4, the scale arithmetic (two IEEE divisions) runs once per block instead of in
// all 32 lanes of a warp, and the operand row leaves as two 16-byte stores -- adjacent threads are adjacent rows, i.e.
// eight threads write one 128-byte core matrix.  Same arithmetic as quantize_block (q4_kernels.cuh): the codes are the
// reference's quantize_row_q4_0 / _q4_1 codes bit for bit.
// SRC_WDATA: rows already quantized by the reference's INIT phase (params->wdata image).
template <int QT, bool SRC_WDATA>
__global__ void __launch_bounds__(128) tile_pack_kernel(const float *__restrict__ x, size_t x_stride, const uint8_t *__restrict__ wdata,
                                                        int M, int MT, int nb, int C, int m_tiles, uint8_t *__restrict__ tpack)
{
    cudaTriggerProgrammaticLaunchCompletion();           // the tile kernel may start its weight stream now
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)m_tiles * C * CHUNK_BLKS * MT;
    if (t >= total) return;
    const int ml = (int)(t % MT);
    long r = t / MT;
    const int bi = (int)(r % CHUNK_BLKS); r /= CHUNK_BLKS;
    const int c = (int)(r % C);
    const int mt = (int)(r / C);
    const int m = mt * MT + ml, blk = c * CHUNK_BLKS + bi;
    const bool real = (m < M) && (blk < nb);
    float d = 0.0f, mn = 0.0f;
    int code[QK];                                        // padding: centred code 0, zero scales
#pragma unroll
    for (int e = 0; e < QK; e++) code[e] = 8;
    if (real) {
        if (SRC_WDATA) {
            const uint8_t *nib;
            if (QT == 0) {
                const uint8_t *b = wdata + ((size_t)m * nb + blk) * 20;
                d = *reinterpret_cast<const float *>(b);
                nib = b + 4;
            } else {
                const uint8_t *row = wdata + (size_t)m * nb * 24;
                mn = reinterpret_cast<const float *>(row)[blk];
                d  = reinterpret_cast<const float *>(row)[nb + blk];
                nib = row + (size_t)nb * 8 + (size_t)blk * 16;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t wv = reinterpret_cast<const uint32_t *>(nib)[i];
#pragma unroll
                for (int j = 0; j < 8; j++) code[8 * i + j] = (int)((wv >> (4 * j)) & 15u);
            }
        } else {
            float v[QK];
            const float4 *src = reinterpret_cast<const float4 *>(x + (size_t)m * x_stride + (size_t)blk * QK);
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
                for (int i = 0; i < QK / 4; i++) { const float4 f = __ldg(src + i); v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w; }
            } else {
#pragma unroll
                for (int e = 0; e < QK; e++) v[e] = x[(size_t)m * x_stride + (size_t)blk * QK + e];
            }
            if (QT == 0) {
                float amax = fabsf(v[0]);
#pragma unroll
                for (int e = 1; e < QK; e++) amax = fmaxf(amax, fabsf(v[e]));
                d = __fdiv_rn(amax, 7.0f);
                const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
                for (int e = 0; e < QK; e++) code[e] = (int)(uint8_t)((int8_t)round_to_int_ref(__fmul_rn(v[e], id)) + 8);
            } else {
                float lo = v[0], hi = v[0];
#pragma unroll
                for (int e = 1; e < QK; e++) { lo = fminf(lo, v[e]); hi = fmaxf(hi, v[e]); }
                // of several numerically equal extrema (+0.0 / -0.0) the reference keeps the FIRST one's bits (ggml.c:268-274)
                float lo_b = lo, hi_b = hi;
#pragma unroll
                for (int e = QK - 1; e >= 0; e--) { if (v[e] == lo) lo_b = v[e]; if (v[e] == hi) hi_b = v[e]; }
                mn = lo_b;
                d = __fdiv_rn(__fsub_rn(hi_b, mn), 15.0f);
                const float id = (d != 0.0f) ? __fdiv_rn(1.0f, d) : 0.0f;
#pragma unroll
                for (int e = 0; e < QK; e++) code[e] = (int)(uint8_t)round_to_int_ref(__fmul_rn(__fsub_rn(v[e], mn), id));
            }
        }
    }
    // operand row: K byte 4t + j = element 8t + 2j (the weight words' low nibbles), 16 + 4t + j = element 8t + 2j + 1 (tk_kpos)
    uint32_t wlo[4], whi[4];
    int s = 0;
#pragma unroll
    for (int tq = 0; tq < 4; tq++) {
        wlo[tq] = 0u; whi[tq] = 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int a0 = (code[8 * tq + 2 * j] & 15) - 8, a1 = (code[8 * tq + 2 * j + 1] & 15) - 8;
            wlo[tq] |= e4m3_of_int(a0, 5) << (8 * j);
            whi[tq] |= e4m3_of_int(a1, 5) << (8 * j);
            s += a0 + a1;
        }
    }
    uint8_t *stage = tpack + ((size_t)mt * C + c) * tk_act_stage(QT, MT);
    const size_t tile = (size_t)MT * 32, plane = (size_t)CHUNK_BLKS * tile;
    uint8_t *t1 = stage + (size_t)bi * tile;
    *reinterpret_cast<uint4 *>(t1 + tk_operand_off(ml, 0)) = make_uint4(wlo[0], wlo[1], wlo[2], wlo[3]);
    *reinterpret_cast<uint4 *>(t1 + tk_operand_off(ml, 16)) = make_uint4(whi[0], whi[1], whi[2], whi[3]);
    if (QT == 1) {                                       // rank-2 right factors: m1' = m1c/16, Z' = (d1*sum(q1-8) + 32*m1c)/16
        uint4 lo, hi;
        const float m1c = __fadd_rn(mn, 8.0f * d);
        bf16_slots2_right(m1c * 0.0625f, fmaf(d, (float)s, 32.0f * m1c) * 0.0625f, lo, hi);
        uint8_t *t4 = stage + 2 * plane + (size_t)bi * tile;
        *reinterpret_cast<uint4 *>(t4 + tk_operand_off(ml, 0)) = lo;
        *reinterpret_cast<uint4 *>(t4 + tk_operand_off(ml, 16)) = hi;
    }
    uint8_t *t3 = stage + plane + (size_t)bi * tile;     // the eight slots of d1 (the upper eight stay zero)
    *reinterpret_cast<uint4 *>(t3 + tk_operand_off(ml, 0)) = bf16_slots_right(d);
    *reinterpret_cast<uint4 *>(t3 + tk_operand_off(ml, 16)) = make_uint4(0u, 0u, 0u, 0u);
}

// The same stage, one WARP per (stage row m_local, block): more parallel, more instructions -- the faster of the two for the
// small grids of M <= 64 (3.2 us against 4.6), where it is launched as a programmatic dependent (DEP; see run_tile).
template <int QT, bool SRC_WDATA, bool DEP>
__global__ void tile_pack_warp_kernel(const float *__restrict__ x, size_t x_stride, const uint8_t *__restrict__ wdata,
                                 int M, int MT, int nb, int C, int m_tiles, uint8_t *__restrict__ tpack)
{
    cudaTriggerProgrammaticLaunchCompletion();           // the tile kernel may start its weight stream now
    if (DEP) cudaGridDependencySynchronize();            // launched as a programmatic dependent: X is ready, the previous call's tile kernel is through with tpack
    const int lane = threadIdx.x & 31;
    const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long total = (long)m_tiles * C * CHUNK_BLKS * MT;
    if (w >= total) return;
    const int ml = (int)(w % MT);
    long r = w / MT;
    const int bi = (int)(r % CHUNK_BLKS); r /= CHUNK_BLKS;
    const int c = (int)(r % C);
    const int mt = (int)(r / C);
    const int m = mt * MT + ml, blk = c * CHUNK_BLKS + bi;
    const bool real = (m < M) && (blk < nb);
    BlockQ q;
    q.d = 0.0f; q.mn = 0.0f; q.code = 8;                 // padding: centred code 0, zero scales
    if (SRC_WDATA) {
        if (real) {
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
    } else {
        const float xv = real ? x[(size_t)m * x_stride + (size_t)blk * QK + lane] : 0.0f;
        q = quantize_block<QT>(xv);
        if (!real) { q.d = 0.0f; q.mn = 0.0f; q.code = 8; }
    }
    const int a = (q.code & 15) - 8;
    uint8_t *stage = tpack + ((size_t)mt * C + c) * tk_act_stage(QT, MT);
    const size_t tile = (size_t)MT * 32, plane = (size_t)CHUNK_BLKS * tile;
    stage[(size_t)bi * tile + tk_operand_off(ml, tk_kpos(lane))] = (uint8_t)e4m3_of_int(a, 5);
    if (QT == 1) {
        int s = a;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 1) {                                 // rank-2 right factors: m1' = m1c/16, Z' = (d1*sum(q1-8) + 32*m1c)/16
            uint4 lo, hi;
            const float m1c = __fadd_rn(q.mn, 8.0f * q.d);
            bf16_slots2_right(m1c * 0.0625f, fmaf(q.d, (float)s, 32.0f * m1c) * 0.0625f, lo, hi);
            uint8_t *t4 = stage + 2 * plane + (size_t)bi * tile;
            *reinterpret_cast<uint4 *>(t4 + tk_operand_off(ml, 0)) = lo;
            *reinterpret_cast<uint4 *>(t4 + tk_operand_off(ml, 16)) = hi;
        }
    }
    if (lane == 0) {                                     // the eight slots of d1 (the upper eight stay zero)
        uint8_t *t3 = stage + plane + (size_t)bi * tile;
        *reinterpret_cast<uint4 *>(t3 + tk_operand_off(ml, 0)) = bf16_slots_right(q.d);
        *reinterpret_cast<uint4 *>(t3 + tk_operand_off(ml, 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ---- tcgen05 / TMEM helpers -------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)          // whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)            // whole warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle, LBO = 128 B, SBO = 256 B, descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}
// instruction descriptors: D = F32, both operands K-major, M = 128, N = n; A = B = E4M3 or BF16
__host__ __device__ constexpr uint32_t umma_idesc_f8(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TK_ROWS >> 4) << 24); }
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TK_ROWS >> 4) << 24); }

__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// NV 32-bit columns of this warp's 32 TMEM lanes -> v[0..NV).  (x16 is the fast shape: x32 measured
// 4x slower per byte, tools/tmembench.cu.)
template <int NV>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NV])
{
    static_assert(NV == 8 || NV == 16, "unsupported tcgen05.ld width");
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    if constexpr (NV == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                       "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(taddr) : "memory");
    }
}
// tcgen05.wait::ld, carrying the loaded registers as in/out operands: the compiler must not move
// their uses above the wait (it does not know the load is asynchronous).
template <int NV>
__device__ __forceinline__ void tmem_wait_ld(float (&a)[NV], float (&b)[NV])
{
    static_assert(NV == 8 || NV == 16, "unsupported width");
    uint32_t *u = reinterpret_cast<uint32_t *>(a), *w = reinterpret_cast<uint32_t *>(b);
    if constexpr (NV == 8)
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                     "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]) :: "memory");
    else
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                     "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]),
                     "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]),
                     "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]) :: "memory");
}

template <int NV>
__device__ __forceinline__ void tmem_wait_ld3(float (&a)[NV], float (&b)[NV], float (&c)[NV])
{
    tmem_wait_ld<NV>(a, b);
    uint32_t *u = reinterpret_cast<uint32_t *>(c);
#pragma unroll
    for (int i = 0; i < NV; i++) asm volatile("" : "+r"(u[i]));      // same ordering constraint for the third tile
}

// one lane of a converged warp (the compiler keeps what that lane computes from warp-uniform inputs in
// uniform registers, which is what the TMA / tcgen05 instructions take -- no per-lane "waterfall")
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(pred));
    return pred != 0;
}

struct RingPos {
    int idx; uint32_t ph; int n;
    __device__ __forceinline__ RingPos(int n_) : idx(0), ph(0), n(n_) {}
    __device__ __forceinline__ void next() { if (++idx == n) { idx = 0; ph ^= 1u; } }
};

// packed fp32 pairs (FFMA2)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(R)); }

// Stream-K work split: the launch's chunk steps (tile-major: tile = n tile x m tile with the m tile
// fastest, so CTAs that run side by side share their weight tile through L2) are cut into gridDim.x
// equal contiguous ranges.  A CTA whose range starts or ends inside a tile produces a PIECE of that
// tile; pieces are parked in a workspace and the owner of the tile's first chunk -- the last to get to
// the tile -- adds them in CTA order (fixed -> bit-reproducible) and stores Y.  One mechanism covers both the small-M case (few tiles, every SM
// must stream weights) and the large-M tail (tile count not a multiple of the SM count).
struct TkRange { int lo, hi; };                      // the host keeps ntiles * C below 2^31
__device__ __forceinline__ TkRange tk_range(const TileParams &p, int b, int G)
{
    const long T = (long)p.ntiles * p.C;
    TkRange r; r.lo = (int)(T * b / G); r.hi = (int)(T * (b + 1) / G);
    return r;
}
// the CTA whose range holds chunk step x
__device__ __forceinline__ int tk_owner(const TileParams &p, int x, int G)
{
    const long T = (long)p.ntiles * p.C;
    return (int)((((long)x + 1) * G + T - 1) / T) - 1;
}

// ---------------------------------------------------------------------------------------------------
// MT: activation rows per tile (MMA N): 16, 32, 64 (and 128 for Q4_0).
// TMEM: NBUF buffers of G blocks x [V: SW columns | T: SW columns (| R: SW columns, Q4_1)], SW = MT or MT / 2 (SPLIT).
// ---------------------------------------------------------------------------------------------------
template <int QT, int MT>
__global__ void __launch_bounds__(tk_threads(QT, MT), 1) q4_tile_kernel(const TileParams p)
{
    // (no early cudaTriggerProgrammaticLaunchCompletion here: measured slower at every M -- the next launch's CTAs only
    // compete with this launch's tail for the SMs' shared memory)
    constexpr int TK_THREADS = tk_threads(QT, MT);
    constexpr int TK_REGS_MISC = tk_regs_misc(TK_THREADS), TK_REGS_EPI = tk_regs_epi(TK_THREADS);
    static_assert(MT <= tk_max_mt(QT), "tile too wide for the TMEM map");
    constexpr int CH = chunk_bytes(QT);
    constexpr int RAW_STAGE = tk_raw_stage(QT), ACT_STAGE = tk_act_stage(QT, MT), CONV_SLOT = tk_conv_slot(QT);
    constexpr int B_TILE = MT * 32;                                           // one block's activation-side operand tile
    constexpr int A_PLANE = CHUNK_BLKS * TK_OP_TILE, B_PLANE = CHUNK_BLKS * B_TILE;
    constexpr bool SPLIT = tk_split(QT, MT);                                  // two half-width MMAs per block and operand
    constexpr int SW = tk_sw(QT, MT);                                         // MMA N = columns of one TMEM tile
    constexpr int G = tk_g(QT, MT);                                           // blocks per TMEM buffer
    constexpr int STEPS = SPLIT ? 2 * CHUNK_BLKS : CHUNK_BLKS / G;            // TMEM hand-offs per chunk
    constexpr int NBUF = tk_nbuf(QT, MT);
    constexpr int NT = tk_nop(QT);                                            // TMEM tiles per block: V, T (, R)
    constexpr int BUFC = G * NT * SW;                                         // columns per TMEM buffer
    static_assert(NBUF * BUFC <= 512 && (NBUF == 2 || NBUF == 4) && (!SPLIT || G == 1), "TMEM map: one issuer per buffer");
    constexpr int W = NBUF;                                                   // issuers at work
    constexpr int GROUPS = tk_groups(QT, MT);
    static_assert(STEPS == tk_steps(QT, MT), "hand-offs per chunk");
    const bool own_only = GROUPS > 1 && p.nconv % GROUPS == 0 && p.nact % GROUPS == 0;   // issuers touch their own group's chunks only
    const int n_arrive = own_only ? STEPS : W;                                // issuers arriving on a slot's empty barrier
    // hand-off k of this CTA's chunk number cc is hand-off cc * STEPS + k of the CTA: its TMEM buffer and the parity of that
    // buffer's use (compile-time functions of k wherever STEPS allows)
    auto tm_buf = [](int cc, int k) -> int {
        if constexpr (STEPS % NBUF == 0) return k % NBUF; else return (cc * STEPS + k) % NBUF;
    };
    auto tm_par = [](int cc, int k) -> uint32_t {
        if constexpr (STEPS % (2 * NBUF) == 0) return (uint32_t)(k / NBUF) & 1u;
        else if constexpr (STEPS % NBUF == 0) return (uint32_t)(cc * (STEPS / NBUF) + k / NBUF) & 1u;
        else return (uint32_t)((cc * STEPS + k) / NBUF) & 1u;
    };
    constexpr int HC = MT / 2;                                                // columns per epilogue thread
    constexpr int LW = HC >= 16 ? 16 : HC;                                    // columns per tcgen05.ld
    constexpr int NSB = HC / LW;                                              // loads per block and operand
    constexpr uint32_t IDESC8 = umma_idesc_f8(SW), IDESC16 = umma_idesc_bf16(SW);

    // (plain pointer arithmetic on the shared array only: an integer round trip would make every
    // access below a generic LD/ST instead of LDS/STS; the no-swizzle operand layout needs 16 B alignment)
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_a8 = smem;                                                     // the constant -8 * 2^-9 tile
    uint8_t *s_conv = s_a8 + TK_OP_TILE;                                      // [nconv][A | A3 (| A4)] x 4 blocks
    uint8_t *s_act = s_conv + (size_t)p.nconv * CONV_SLOT;                    // [nact][B | B3 (| B4)] x 4 blocks
    uint8_t *s_rawr = s_act + (size_t)p.nact * ACT_STAGE;                     // [nraw][RAW_STAGE]
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_rawr + (size_t)p.nraw * RAW_STAGE);
    uint64_t *raw_full = bars, *raw_empty = raw_full + TK_MAX_RAW;
    uint64_t *act_full = raw_empty + TK_MAX_RAW, *act_empty = act_full + TK_MAX_ACT;
    uint64_t *conv_full = act_empty + TK_MAX_ACT, *conv_empty = conv_full + TK_MAX_CONV;
    uint64_t *tm_full = conv_empty + TK_MAX_CONV, *tm_empty = tm_full + TK_MAX_TM;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(tm_empty + TK_MAX_TM);
    int      *s_flag = reinterpret_cast<int *>(s_tmem + 1);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform
    TK_STAMP(8, 7, threadIdx.x == 0);                                         // kernel entry
    const int G_ = (int)gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nraw; i++) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], TK_CONV_WARPS); }
        for (int i = 0; i < p.nact; i++) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], n_arrive); }
        for (int i = 0; i < p.nconv; i++) { mbar_init(&conv_full[i], TK_CONV_WARPS); mbar_init(&conv_empty[i], n_arrive); }
        for (int i = 0; i < NBUF; i++) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], SPLIT ? TK_EPI_WARPS / 2 : TK_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        for (int i = threadIdx.x; i < TK_OP_TILE / 16; i += TK_THREADS)
            reinterpret_cast<uint4 *>(s_a8)[i] = make_uint4(0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u);
        // K slots 8..15 of the scale-piece tiles (second core matrix of every 8-row group) are zero for good
        for (int i = threadIdx.x; i < p.nconv * CHUNK_BLKS * (TK_ROWS / 8) * 8; i += TK_THREADS) {
            const int piece = i & 7, grp = (i >> 3) % (TK_ROWS / 8), blk = (i >> 3) / (TK_ROWS / 8);     // blk over all slots
            const int slot = blk / CHUNK_BLKS, bi = blk % CHUNK_BLKS;
            reinterpret_cast<uint4 *>(s_conv + (size_t)slot * CONV_SLOT + A_PLANE + bi * TK_OP_TILE + grp * 256 + 128)[piece] = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    TK_STAMP(8, 8, threadIdx.x == 0);                                         // prologue done
    const TkRange rg = tk_range(p, (int)blockIdx.x, G_);

    auto issuer = [&](const int w) {
            // ===================== MMA issuers =====================
            // W = NBUF of them (one elected thread each); TMEM hand-off n = 0, 1, 2 ... of this CTA goes to buffer n % NBUF
            // and to issuer n % W, so every buffer has one issuer.  Why several: after tcgen05.mma / tcgen05.commit a
            // thread's next mbarrier wait takes 100-250 cycles even on a barrier that completed long ago, and does not return
            // before the thread's own MMAs have (tools/mmabench.cu, profiles/r02_mmabench.txt) -- one issuer alone leaves the
            // tensor pipe idle for most of every hand-off.  Unless a slot always belongs to the same issuer group (own_only, see
            // tk_groups), every issuer waits for the operand slots of EVERY chunk and arrives on their empty barriers, also
            // for chunks it issues nothing from: skipping chunks would let a parity wait run two phases ahead of its barrier.
            TK_PROF_DECL;
            RingPos ar(p.nact), cr(p.nconv);
            const uint32_t a8 = smem_u32(s_a8);
            for (int x = rg.lo; x < rg.hi && w < W; x++) {
                if (own_only && (x - rg.lo) % GROUPS != w / STEPS) { cr.next(); ar.next(); continue; }     // another group's chunk
                {
                    TK_WAIT(&conv_full[cr.idx], cr.ph, 5);
                    TK_STAMP(10, x - rg.lo, w == 0);
                    TK_WAIT(&act_full[ar.idx], ar.ph, 6);
                    TK_STAMP(4, x - rg.lo, w == 0);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(s_conv + (size_t)cr.idx * CONV_SLOT);
                    const uint32_t b_base = smem_u32(s_act + (size_t)ar.idx * ACT_STAGE);
                    const int cc = x - rg.lo;
#pragma unroll
                    for (int st = 0; st < STEPS; st++) {
                        const int tb = tm_buf(cc, st);                                        // W == NBUF: the buffer's issuer has its number
                        if (tb != w) continue;
                        const uint32_t tph = tm_par(cc, st);
                        TK_WAIT(&tm_empty[tb], tph ^ 1u, 7);
                        tc_fence_after();
#pragma unroll
                        for (int g = 0; g < G; g++) {
#ifdef TK_EXP_NOMMA
                            continue;
#endif
                            const int bi = SPLIT ? st >> 1 : st * G + g;                      // block of the chunk
                            const int bo = SPLIT ? (st & 1) * SW * 32 : 0;                    // activation rows [SW, 2 SW) of the tile
                            const uint32_t dv = tmem + (uint32_t)(tb * BUFC + g * NT * SW);
                            const uint64_t bd = umma_desc(b_base + bi * B_TILE + bo);
                            umma_f8(dv, umma_desc(a_base + bi * TK_OP_TILE), bd, IDESC8, 0u);
#ifndef TK_EXP_NOA8
                            umma_f8(dv, umma_desc(a8), bd, IDESC8, 1u);
#endif
                            umma_bf16(dv + SW, umma_desc(a_base + A_PLANE + bi * TK_OP_TILE), umma_desc(b_base + B_PLANE + bi * B_TILE + bo), IDESC16, 0u);
                            if constexpr (QT == 1)
                                umma_bf16(dv + 2 * SW, umma_desc(a_base + 2 * A_PLANE + bi * TK_OP_TILE),
                                          umma_desc(b_base + 2 * B_PLANE + bi * B_TILE + bo), IDESC16, 0u);
                        }
                        umma_commit(&tm_full[tb]);
                    }
                    umma_commit(&conv_empty[cr.idx]);
                    umma_commit(&act_empty[ar.idx]);
                    TK_STAMP(5, x - rg.lo, w == 0);
                    cr.next(); ar.next();
                }
            }
            TK_PROF_FLUSH(4, w == 0);
    };

    if (warp < 4) {
        reg_dec<TK_REGS_MISC>();
        if (warp == 0 && elect_one()) {
            // ===================== producer: weight chunks =====================
            TK_PROF_DECL;
            RingPos rr(p.nraw);
            const uint32_t raw0 = smem_u32(s_rawr);
            const size_t tstride = (size_t)p.C * CH;     // same chunk column, next 16-row tile
            int x = rg.lo;
            while (x < rg.hi) {
                const int tile = x / p.C, c0 = x - tile * p.C;
                const int nc = min(p.C - c0, rg.hi - x);
                const int nt = tile / p.m_tiles;
                const int ntl = min(TK_T16, p.n_t16 - nt * TK_T16);
                const uint8_t *src = p.wt + ((size_t)(nt * TK_T16) * p.C + c0) * CH;
                for (int c = 0; c < nc; c++, src += CH) {
                    TK_WAIT(&raw_empty[rr.idx], rr.ph ^ 1u, 1);
                    TK_STAMP(0, x + c - rg.lo, true);
                    const uint32_t bar = smem_u32(&raw_full[rr.idx]);
                    const uint32_t dst = raw0 + (uint32_t)rr.idx * RAW_STAGE;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)ntl * CH) : "memory");
                    if (ntl == TK_T16) {
#pragma unroll
                        for (int tl = 0; tl < TK_T16; tl++)
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"(dst + tl * CH), "l"(src + tl * tstride), "r"((uint32_t)CH), "r"(bar) : "memory");
                    } else {
                        for (int tl = 0; tl < ntl; tl++)
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"(dst + tl * CH), "l"(src + tl * tstride), "r"((uint32_t)CH), "r"(bar) : "memory");
                    }
                    rr.next();
                }
                x += nc;
            }
            TK_PROF_FLUSH(0, true);
        } else if (warp == 3 && elect_one()) {
            // ===================== producer: activation stages =====================
            TK_PROF_DECL;
            TK_STAMP(8, 0, true);
            cudaGridDependencySynchronize();             // the pack kernel's output
            TK_STAMP(8, 1, true);
            RingPos ar(p.nact);
            const uint32_t act0 = smem_u32(s_act);
            int x = rg.lo;
            while (x < rg.hi) {
                const int tile = x / p.C, c0 = x - tile * p.C;
                const int nc = min(p.C - c0, rg.hi - x);
                const int mt = tile % p.m_tiles;
                const uint8_t *src = p.tpack + ((size_t)mt * p.C + c0) * ACT_STAGE;
                for (int c = 0; c < nc; c++, src += ACT_STAGE) {
                    TK_WAIT(&act_empty[ar.idx], ar.ph ^ 1u, 3);
                    TK_STAMP(9, x + c - rg.lo, true);
                    const uint32_t bar = smem_u32(&act_full[ar.idx]);
#ifdef TK_EXP_NOACT
                    if (p.N > 0) { mbar_arrive(&act_full[ar.idx]); ar.next(); continue; }
#endif
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)ACT_STAGE) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(act0 + (uint32_t)ar.idx * ACT_STAGE), "l"(src), "r"((uint32_t)ACT_STAGE), "r"(bar) : "memory");
                    ar.next();
                }
                x += nc;
            }
            TK_PROF_FLUSH(2, true);
        } else if (TK_THREADS == TK_THREADS_BASE && (warp == 1 || warp == 2) && elect_one()) {
            issuer(warp - 1);
        }
        __syncwarp();
    } else if (warp < 4 + TK_CONV_WARPS) {
        // ===================== converters =====================
        reg_dec<TK_REGS_CONV>();
        const int cw = warp - 4;
        TK_PROF_DECL;
        RingPos rr(p.nraw), cr(p.nconv);
        for (int x = rg.lo; x < rg.hi; x++) {
            TK_WAIT(&raw_full[rr.idx], rr.ph, 9);
            TK_STAMP(1, x - rg.lo, cw == 0 && lane == 0);
            TK_WAIT(&conv_empty[cr.idx], cr.ph ^ 1u, 10);
            TK_STAMP(2, x - rg.lo, cw == 0 && lane == 0);
            const uint8_t *raw = s_rawr + (size_t)rr.idx * RAW_STAGE;
            uint8_t *conv = s_conv + (size_t)cr.idx * CONV_SLOT;
#ifdef TK_EXP_NOCONV
            if (p.N > 0) { __syncwarp(); if (lane == 0) { mbar_arrive(&conv_full[cr.idx]); mbar_arrive(&raw_empty[rr.idx]); } rr.next(); cr.next(); continue; }
#endif
            // nibble words: lane (g, t) = (row g of an 8-row group, word t); scale path: lane (g2, t2) =
            // (row g2, block t2) so that 8 consecutive lanes store 8 consecutive 16-byte operand rows
            const int g2 = lane & 7, t2 = lane >> 3;
            uint4 q[4];
            float d0[4], m0[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {                    // all loads first
                const uint8_t *ch = raw + (2 * cw + (u >> 1)) * CH;
                q[u] = *reinterpret_cast<const uint4 *>(ch + (u & 1) * 512 + lane * 16);
                d0[u] = *reinterpret_cast<const float *>(ch + 1024 + (((u & 1) * 8 + g2) * 4 + t2) * 4);
                if constexpr (QT == 1) m0[u] = *reinterpret_cast<const float *>(ch + 1280 + (((u & 1) * 8 + g2) * 4 + t2) * 4);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int rgp = (2 * cw + (u >> 1)) * 2 + (u & 1);           // 8-row group of this lane's row
                uint8_t *dst = conv + rgp * 256 + lane * 4;                  // = g * 16 + t * 4: one contiguous 128 B per warp and store
                int s0[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t wv = u4_get(q[u], i);
                    const uint32_t lo = wv & 0x0F0F0F0Fu, hi = (wv >> 4) & 0x0F0F0F0Fu;
                    *reinterpret_cast<uint32_t *>(dst + i * TK_OP_TILE) = lo;
                    *reinterpret_cast<uint32_t *>(dst + i * TK_OP_TILE + 128) = hi;
                    if constexpr (QT == 1) s0[i] = __dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, 0u));
                }
                *reinterpret_cast<uint4 *>(conv + A_PLANE + t2 * TK_OP_TILE + rgp * 256 + g2 * 16) = bf16_slots_left(d0[u]);
                if constexpr (QT == 1) {
                    int mine = 0;
#pragma unroll
                    for (int i = 0; i < 4; i++) {            // row sums of the codes of block i live in the quad of row g
                        s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], 1);
                        s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], 2);
                        const int v = __shfl_sync(0xffffffffu, s0[i], g2 * 4);
                        if (i == t2) mine = v;
                    }
                    uint4 l4, h4;
                    bf16_slots2_left(d0[u] * (float)(mine - 256), __fadd_rn(m0[u], 8.0f * d0[u]), l4, h4);   // centred: S0 - 8*32, m0 + 8*d0
                    uint8_t *d4 = conv + 2 * A_PLANE + t2 * TK_OP_TILE + rgp * 256 + g2 * 16;
                    *reinterpret_cast<uint4 *>(d4) = l4;
                    *reinterpret_cast<uint4 *>(d4 + 128) = h4;
                }
            }
            fence_proxy_async();                         // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(&conv_full[cr.idx]); mbar_arrive(&raw_empty[rr.idx]); }
            TK_STAMP(3, x - rg.lo, cw == 0 && lane == 0);
            rr.next(); cr.next();
        }
        TK_PROF_FLUSH(8, warp == 4 && lane == 0);
    } else if (warp < 4 + TK_CONV_WARPS + TK_EPI_WARPS) {
        // ===================== epilogue =====================
        reg_inc<TK_REGS_EPI>();
        const int ew = warp - (4 + TK_CONV_WARPS);
        const int q = ew & 3, h = ew >> 2;               // TMEM lane quarter (= warp id % 4), column half
        const int row = q * 32 + lane;                   // weight row inside the tile
        const int etid = ew * 32 + lane;
        // this thread's columns: all of half-tile h (SPLIT: a TMEM tile of its own), else columns [h HC, (h + 1) HC) of the tile
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(SPLIT ? 0 : h * HC);
        TK_PROF_DECL;
        cudaGridDependencySynchronize();                 // Y, the piece workspace and its counters belong to earlier launches until now
        TK_STAMP(8, 2, ew == 0 && lane == 0);
        int x = rg.lo;
        while (x < rg.hi) {
            const int tile = x / p.C, c0 = x - tile * p.C;
            const int nc = min(p.C - c0, rg.hi - x);
            const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
            f32x2 acc[HC / 2];
#pragma unroll
            for (int j = 0; j < HC / 2; j++) acc[j] = 0ull;
            for (int c = 0; c < nc; c++) {
                // software pipeline over the chunk's (block, 16-column) loads: V and T of load i+1 are
                // in flight while the FFMA2s of load i run
                constexpr int NL = CHUNK_BLKS * NSB;
                float v[2][LW], s[2][LW], r[2][QT == 1 ? LW : 1];
                const int cc = x + c - rg.lo;            // chunk number of this CTA
                // load i = (block i / NSB, column group i % NSB): its hand-off inside the chunk and first column
                auto step_of  = [&](int i) { const int bi = i / NSB; return SPLIT ? 2 * bi + h : bi / G; };
                auto first_of = [&](int i) { return i % NSB == 0 && (SPLIT || (i / NSB) % G == 0); };
                auto last_of  = [&](int i) { return i % NSB == NSB - 1 && (SPLIT || (i / NSB) % G == G - 1); };
                auto column   = [&](int i) { return (uint32_t)(tm_buf(cc, step_of(i)) * BUFC + (SPLIT ? 0 : ((i / NSB) % G) * NT * SW) + (i % NSB) * LW); };
#ifdef TK_EXP_NOEPI
                if (p.N > 0) {
                    for (int i = 0; i < NL; i++) if (last_of(i)) {
                        TK_WAIT(&tm_full[tm_buf(cc, step_of(i))], tm_par(cc, step_of(i)), 14);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tm_empty[tm_buf(cc, step_of(i))]);
                    }
                    continue;
                }
#endif
                {
                    TK_WAIT(&tm_full[tm_buf(cc, step_of(0))], tm_par(cc, step_of(0)), 14);
                    TK_STAMP(6, x + c - rg.lo, ew == 0 && lane == 0);
                    tc_fence_after();
                    const uint32_t col = column(0);
                    tmem_ld<LW>(t_lane + col, v[0]);
                    tmem_ld<LW>(t_lane + col + SW, s[0]);
                    if constexpr (QT == 1) tmem_ld<LW>(t_lane + col + 2 * SW, r[0]);
                }
#pragma unroll
                for (int i = 0; i < NL; i++) {
                    const int sb = i % NSB;
                    if constexpr (QT == 1) tmem_wait_ld3<LW>(v[i & 1], s[i & 1], r[i & 1]);
                    else                   tmem_wait_ld<LW>(v[i & 1], s[i & 1]);
                    if (last_of(i)) {                    // that was the last read of this TMEM buffer: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tm_empty[tm_buf(cc, step_of(i))]);
                    }
                    if (i + 1 < NL) {
                        if (first_of(i + 1)) {           // first read of the next buffer
                            TK_WAIT(&tm_full[tm_buf(cc, step_of(i + 1))], tm_par(cc, step_of(i + 1)), 14);
                            tc_fence_after();
                        }
                        const uint32_t col = column(i + 1);
                        tmem_ld<LW>(t_lane + col, v[(i + 1) & 1]);
                        tmem_ld<LW>(t_lane + col + SW, s[(i + 1) & 1]);
                        if constexpr (QT == 1) tmem_ld<LW>(t_lane + col + 2 * SW, r[(i + 1) & 1]);
                    }
#pragma unroll
                    for (int j = 0; j < LW / 2; j++) {
                        f32x2 a2 = fma2(pk2(s[i & 1][2 * j], s[i & 1][2 * j + 1]), pk2(v[i & 1][2 * j], v[i & 1][2 * j + 1]), acc[sb * (LW / 2) + j]);
                        if constexpr (QT == 1) a2 = add2(a2, pk2(r[i & 1][2 * j], r[i & 1][2 * j + 1]));
                        acc[sb * (LW / 2) + j] = a2;
                    }
                }
                TK_STAMP(7, x + c - rg.lo, ew == 0 && lane == 0);
            }
            TK_STAMP(8, 5, ew == 0 && lane == 0 && x + nc >= rg.hi);                  // last piece accumulated
            // ---- this CTA's share of the tile is done: store, or park the piece and let the last one combine ----
            const float post = 16.0f;
            const int n = nt * TK_ROWS + row;
            const int mbase = mt * MT + h * HC;
            const bool whole = (c0 == 0) && (nc == p.C);
            bool store = true;
            if (!whole) {
                // Stream-K fix-up.  The CTA whose range holds the tile's FIRST chunk gets to it at the END of its range: it is
                // the combiner -- it keeps its piece in registers, waits until the other owners have parked theirs (they did
                // so at the START of their ranges, long ago, or are finishing at the same moment), adds them in CTA order
                // (fixed -> bit-reproducible) and stores Y.  Every other owner parks its piece, bumps the tile's counter and
                // goes on at once: no atomic round trip, no "am I last" decision on anybody's critical path.
                const int b = (int)blockIdx.x;
                const int t0 = tile * p.C;
                const int bf = tk_owner(p, t0, G_), bl = tk_owner(p, t0 + p.C - 1, G_);
                if (c0 != 0) {                                // b != bf: park (slot 2 b; slot 2 b + 1 is no longer used)
                    float *ws = p.partial + ((size_t)(2 * b) * MT + (size_t)(h * HC)) * TK_ROWS + row;
#pragma unroll
                    for (int j = 0; j < HC / 2; j++) {
                        float a, bq;
                        upk2(acc[j], a, bq);
                        ws[(size_t)(2 * j) * TK_ROWS] = a;
                        ws[(size_t)(2 * j + 1) * TK_ROWS] = bq;
                    }
                    __threadfence();                          // every writer orders its own stores, then the barrier, then the count
                    asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");
                    if (etid == 0) atomicAdd(p.counters + tile, 1);
                    store = false;
                } else {                                      // b == bf
                    if (etid == 0) {
                        while (*reinterpret_cast<volatile int *>(p.counters + tile) < bl - bf) { }
                        p.counters[tile] = 0;                 // self-resetting: every other owner has counted
                    }
                    asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");
                    __threadfence();
                    float sum[HC];
#pragma unroll
                    for (int j = 0; j < HC / 2; j++) upk2(acc[j], sum[2 * j], sum[2 * j + 1]);
                    for (int bb = bf + 1; bb <= bl; bb++) {   // CTA order: reproducible rounding
                        const float *rsrc = p.partial + ((size_t)(2 * bb) * MT + (size_t)(h * HC)) * TK_ROWS + row;
#pragma unroll
                        for (int j = 0; j < HC; j++) sum[j] += __ldcg(rsrc + (size_t)j * TK_ROWS);
                    }
#pragma unroll
                    for (int j = 0; j < HC / 2; j++) acc[j] = pk2(sum[2 * j], sum[2 * j + 1]);
                }
            }
            if (store && n < p.N) {
#pragma unroll
                for (int j = 0; j < HC / 2; j++) {
                    float a, bq;
                    upk2(acc[j], a, bq);
                    if (mbase + 2 * j < p.M)     p.y[(size_t)(mbase + 2 * j) * p.y_stride + n] = a * post;
                    if (mbase + 2 * j + 1 < p.M) p.y[(size_t)(mbase + 2 * j + 1) * p.y_stride + n] = bq * post;
                }
            }
            x += nc;
        }
        TK_STAMP(8, 6, ew == 0 && lane == 0);                                     // last store / combine done
        TK_PROF_FLUSH(11, ew == 0 && lane == 0);
    } else {
        reg_dec<TK_REGS_MISC>();
        if (elect_one()) issuer(warp - (4 + TK_CONV_WARPS + TK_EPI_WARPS));
        __syncwarp();
    }
    TK_STAMP(8, 3, threadIdx.x == 0);
    tc_fence_before();
    __syncthreads();
    TK_STAMP(8, 4, threadIdx.x == 0);
    if (warp == 2) tmem_dealloc(tmem, 512);
}

} // namespace b200q4
