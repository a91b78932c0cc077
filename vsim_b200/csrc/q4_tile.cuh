// q4_tile.cuh -- the batched (M >= 3) Q4_0/Q4_1 mul_mat kernel for sm_100a: TMA-staged weight tiles,
// nibble -> 8-bit expansion into shared memory, tcgen05 MMA with TMEM accumulators, fp32 block
// scaling on the CUDA cores.
//
// Reference semantics: ggml_compute_forward_mul_mat_q4_{0,1}_f32 (ggml.c:4891-5453) with the block
// dot products of ggml.c:472-553 / imax.c:1205-1229.  Same exact-integer formulation as the decode
// kernel (q4_kernels.cuh):
//
//   Q4_0:  Y[m][n] = sum_blk d0[n,blk] * d1[m,blk] * ( P - 8*Sa ),   P = sum_k q0*a, a = q1-8, Sa = sum_k a
//   Q4_1:  Y[m][n] = sum_blk d0*d1*P + (d0*S0)*m1 + m0*Z,             P = sum_k q0*q1, S0 = sum_k q0,
//                                                                     Z = d1*sum(q1) + 32*m1
//
// The 32-element sums P are ONE tcgen05.mma each (kind::f8f6f4, M=128 weight rows x N=MT activation
// rows x K=32 = one quant block) and they are exact:
//   * a weight nibble q0 in [0,15], left as the byte 0x0q, IS the e4m3 number q0 * 2^-9 (e4m3
//     subnormals 0x01..0x07 = 1..7 * 2^-9 continue linearly into the first binade 0x08..0x0F =
//     8..15 * 2^-9).  So "dequantizing" the weights into the MMA operand is one AND (low nibbles) or
//     one shift + AND (high nibbles) per 4 elements -- no table, no arithmetic.
//   * the activation codes are written by our own pack kernel as e4m3 bytes of a*2^5 (Q4_0, |a|<=8)
//     or q1*2^4 (Q4_1, q1<=15): small integers times a power of two are exact in e4m3.
//   * every product is a multiple of 2^-4 / 2^-5 below 2^8; 32 of them sum exactly in fp32.
// So TMEM holds D = P/16 (Q4_0) or P/32 (Q4_1) exactly, one fresh accumulator tile per quant block,
// and the per-block fp32 scales -- different for every (n, blk) and (m, blk), which is what keeps a
// plain GEMM from doing this job -- are applied by the epilogue warps:
//   Q4_0:  acc[m] = fma(d0[n], fma(D, d1[m], e[m]), acc[m])        e = -0.5*d1*Sa        (x16 at the store)
//   Q4_1:  acc[m] = fma(m0[n], Z'[m], fma(w1[n], m1'[m], fma(d0[n], d1[m]*D, acc[m])))   (x32 at the store)
// Only those fp32 scale products are rounded.  The epilogue (2 or 4 FFMA per (m, n, block)) is what
// bounds the kernel once M is large; below M ~ 48 the weight stream (HBM) bounds it.
//
// Warp roles (512 threads, one CTA per SM, persistent over tasks):
//   warp 0      producer: cp.async.bulk (TMA 1-D) of weight chunks (8 x 1280/1536 B per 128 rows x 4
//               blocks, straight from the decode kernel's "wtile" layout -- one weight copy in HBM
//               serves both kernels) and of activation stages (codes already in the UMMA canonical
//               K-major layout + column scales), each ring with full/empty mbarriers.
//   warp 1      MMA issuer (one thread): 4 tcgen05.mma per chunk into alternating TMEM buffers,
//               tcgen05.commit to the epilogue and back to the rings.
//   warp 2      TMEM allocator (512 columns).
//   warps 4-7   converters: nibble words -> e4m3 bytes in the UMMA canonical layout (no swizzle: core
//               matrix = 8 rows x 16 B), row scales (d0, d0*S0, m0) into a small ring.
//   warps 8-15  epilogue: tcgen05.ld (32 lanes x 32 bit), scale FMAs, final store / split-K combine.
#pragma once
#include "q4_kernels.cuh"

namespace b200q4 {

constexpr int TK_ROWS = 128;                 // weight rows per tile (MMA M)
constexpr int TK_T16 = TK_ROWS / TILE_ROWS;  // wtile tiles per tile
constexpr int TK_CONV_WARPS = 4;
constexpr int TK_EPI_WARPS = 8;
constexpr int TK_THREADS = (4 + TK_CONV_WARPS + TK_EPI_WARPS) * 32;
constexpr int TK_MAX_RAW = 16, TK_MAX_ACT = 8, TK_MAX_CONV = 8;
// register budgets after setmaxnreg (launch: 512 threads x 128): 128*(40 + 72 + 2*200) = 65536
constexpr int TK_REGS_MISC = 40, TK_REGS_CONV = 72, TK_REGS_EPI = 200;

__host__ __device__ constexpr int tk_ncs(int qt) { return qt == 0 ? 2 : 4; }     // column-scale floats per (m, blk)
__host__ __device__ constexpr int tk_nrs(int qt) { return qt == 0 ? 1 : 3; }     // row-scale planes per (n, blk)
__host__ __device__ constexpr int tk_act_stage(int qt, int mt) { return CHUNK_BLKS * mt * 32 + CHUNK_BLKS * mt * tk_ncs(qt) * 4; }
__host__ __device__ constexpr int tk_raw_stage(int qt) { return TK_T16 * chunk_bytes(qt); }
__host__ __device__ constexpr int tk_conv_slot(int qt) { return CHUNK_BLKS * TK_ROWS * 32 + CHUNK_BLKS * tk_nrs(qt) * TK_ROWS * 4; }
__host__ __device__ constexpr int tk_bar_bytes() { return (2 * TK_MAX_RAW + 2 * TK_MAX_ACT + 2 * TK_MAX_CONV + 4) * 8 + 64; }
__host__ __device__ inline size_t tk_smem_bytes(int qt, int mt, int nraw, int nact, int nconv)
{
    return 1024 + (size_t)nraw * tk_raw_stage(qt) + (size_t)nact * tk_act_stage(qt, mt) +
           (size_t)nconv * tk_conv_slot(qt) + tk_bar_bytes();
}

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

// K position (byte inside the block's 32-byte operand row) of element e: the converter emits, per
// nibble word t (elements 8t .. 8t+7), the four low nibbles (even elements) then the four high ones.
__device__ __forceinline__ int tk_kpos(int e) { return (e & 24) + ((e & 1) << 2) + ((e & 7) >> 1); }

// Byte offset of (row r, K position p) inside one block's operand tile in the UMMA canonical K-major
// no-swizzle layout: core matrix = 8 rows x 16 B contiguous; K-adjacent core matrices 128 B apart
// (LBO), 8-row groups 256 B apart (SBO).
__device__ __forceinline__ int tk_operand_off(int r, int p) { return (r >> 3) * 256 + (p >> 4) * 128 + (r & 7) * 16 + (p & 15); }

// ---- activation pack for the tile kernel ----------------------------------------------------------
// One warp per (stage row m_local, block) of every (m tile, chunk) stage, padding included.
// SRC_WDATA: rows already quantized by the reference's INIT phase (params->wdata image).
template <int QT, bool SRC_WDATA>
__global__ void tile_pack_kernel(const float *__restrict__ x, size_t x_stride, const uint8_t *__restrict__ wdata,
                                 int M, int MT, int nb, int C, int m_tiles, uint8_t *__restrict__ tpack)
{
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
    q.d = 0.0f; q.mn = 0.0f; q.code = (QT == 0) ? 8 : 0;
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
        if (!real) { q.d = 0.0f; q.mn = 0.0f; q.code = (QT == 0) ? 8 : 0; }
    }
    const int a = (QT == 0) ? (q.code & 15) - 8 : (q.code & 15);
    uint8_t *stage = tpack + ((size_t)mt * C + c) * tk_act_stage(QT, MT);
    stage[(size_t)bi * MT * 32 + tk_operand_off(ml, tk_kpos(lane))] = (uint8_t)e4m3_of_int(a, QT == 0 ? 5 : 4);
    int s = a;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        // column scales, stored per PAIR of columns so that one LDS.128 feeds one packed FFMA2:
        //   Q4_0: {d1_a, d1_b, e_a, e_b}          Q4_1: {d1_a, d1_b, m1'_a, m1'_b, Z'_a, Z'_b, 0, 0}
        float *cs = reinterpret_cast<float *>(stage + (size_t)CHUNK_BLKS * MT * 32) +
                    ((size_t)bi * (MT / 2) + (ml >> 1)) * (2 * tk_ncs(QT)) + (ml & 1);
        if (QT == 0) {
            cs[0] = q.d;
            cs[2] = -0.5f * (q.d * (float)s);
        } else {
            cs[0] = q.d;
            cs[2] = q.mn * 0.03125f;
            cs[4] = fmaf(q.d, (float)s, 32.0f * q.mn) * 0.03125f;
            cs[6] = 0.0f;
        }
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
// instruction descriptor: D = F32, A = B = E4M3, both K-major, N = n, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TK_ROWS >> 4) << 24); }

__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// tcgen05.wait::ld, carrying the loaded registers as in/out operands: the compiler must not move
// their uses above the wait (it does not know the load is asynchronous).
template <int NV>
__device__ __forceinline__ void tmem_wait_ld(float (&v)[NV])
{
    static_assert(NV == 8 || NV == 16, "unsupported width");
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    if constexpr (NV == 8)
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]) :: "memory");
    else
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                     "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]) :: "memory");
}

// NV 32-bit columns of this warp's 32 TMEM lanes -> v[0..NV)
template <int NV>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NV])
{
    static_assert(NV == 8 || NV == 16 || NV == 32, "unsupported tcgen05.ld width");
    uint32_t *u = reinterpret_cast<uint32_t *>(v);
    if constexpr (NV == 8) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr) : "memory");
    } else if constexpr (NV == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                       "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(taddr) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                       "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                       "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                       "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]) : "r"(taddr) : "memory");
    }
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
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(R)); }

// Stream-K work split: the launch's chunk steps (tile-major: tile = n tile x m tile with the m tile
// fastest, so CTAs that run side by side share their weight tile through L2) are cut into gridDim.x
// equal contiguous ranges.  A CTA whose range starts or ends inside a tile produces a PIECE of that
// tile; pieces are parked in a workspace and the last one to arrive adds them in CTA order (fixed ->
// bit-reproducible) and stores Y.  One mechanism covers both the small-M case (few tiles, every SM
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
// MT: activation rows per tile (MMA N): 16, 32, 64, 128 or 256.  A TMEM buffer holds G = min(4, 256/MT)
// block accumulators (G * MT <= 256 columns); two buffers alternate.
// ---------------------------------------------------------------------------------------------------
template <int QT, int MT>
__global__ void __launch_bounds__(TK_THREADS, 1) q4_tile_kernel(const TileParams p)
{
    constexpr int CH = chunk_bytes(QT);
    constexpr int NCS = tk_ncs(QT), NRS = tk_nrs(QT);
    constexpr int RAW_STAGE = tk_raw_stage(QT), ACT_STAGE = tk_act_stage(QT, MT);
    constexpr int CONV_SLOT = tk_conv_slot(QT);
    constexpr int CONV_A = CHUNK_BLKS * TK_ROWS * 32;                         // operand bytes of a conv slot; row scales follow
    constexpr int G = (256 / MT) < CHUNK_BLKS ? (256 / MT) : CHUNK_BLKS;      // blocks per TMEM buffer
    constexpr int STEPS = CHUNK_BLKS / G;                                     // TMEM buffers per chunk
    constexpr int HC = MT / 2;                                                // columns per epilogue thread
    constexpr int LW = HC >= 16 ? 16 : HC;                                    // columns per tcgen05.ld
    constexpr int NSB = HC / LW;                                              // loads per block
    constexpr uint32_t IDESC = umma_idesc(MT);

    // (plain pointer arithmetic on the shared array only: an integer round trip would make every
    // access below a generic LD/ST instead of LDS/STS; the no-swizzle operand layout needs 16 B alignment)
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_conv = smem;                                                   // [nconv][4 blocks x 128 rows x 32 B | row scales [4][NRS][128]]
    uint8_t *s_act = s_conv + (size_t)p.nconv * CONV_SLOT;                    // [nact][ACT_STAGE]
    uint8_t *s_rawr = s_act + (size_t)p.nact * ACT_STAGE;                     // [nraw][RAW_STAGE]
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_rawr + (size_t)p.nraw * RAW_STAGE);
    uint64_t *raw_full = bars, *raw_empty = raw_full + TK_MAX_RAW;
    uint64_t *act_full = raw_empty + TK_MAX_RAW, *act_empty = act_full + TK_MAX_ACT;
    uint64_t *conv_full = act_empty + TK_MAX_ACT, *conv_empty = conv_full + TK_MAX_CONV;
    uint64_t *tm_full = conv_empty + TK_MAX_CONV, *tm_empty = tm_full + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(tm_empty + 2);
    int      *s_flag = reinterpret_cast<int *>(s_tmem + 1);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);    // provably warp-uniform
    const int G_ = (int)gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nraw; i++) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], TK_CONV_WARPS); }
        for (int i = 0; i < p.nact; i++) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], 1 + TK_EPI_WARPS); }
        for (int i = 0; i < p.nconv; i++) { mbar_init(&conv_full[i], TK_CONV_WARPS); mbar_init(&conv_empty[i], 1 + TK_EPI_WARPS); }
        for (int i = 0; i < 2; i++) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], TK_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const TkRange rg = tk_range(p, (int)blockIdx.x, G_);

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
                    const uint32_t bar = smem_u32(&act_full[ar.idx]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"((uint32_t)ACT_STAGE) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(act0 + (uint32_t)ar.idx * ACT_STAGE), "l"(src), "r"((uint32_t)ACT_STAGE), "r"(bar) : "memory");
                    ar.next();
                }
                x += nc;
            }
            TK_PROF_FLUSH(2, true);
        } else if (warp == 1 && elect_one()) {
            // ===================== MMA issuer =====================
            TK_PROF_DECL;
            RingPos ar(p.nact), cr(p.nconv), tr(2);
            for (int x = rg.lo; x < rg.hi; x++) {
                TK_WAIT(&conv_full[cr.idx], cr.ph, 5);
                TK_WAIT(&act_full[ar.idx], ar.ph, 6);
                tc_fence_after();
                const uint32_t a_base = smem_u32(s_conv + (size_t)cr.idx * CONV_SLOT);
                const uint32_t b_base = smem_u32(s_act + (size_t)ar.idx * ACT_STAGE);
#pragma unroll
                for (int st = 0; st < STEPS; st++) {
                    TK_WAIT(&tm_empty[tr.idx], tr.ph ^ 1u, 7);
                    tc_fence_after();
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        const int bi = st * G + g;
                        umma_f8(tmem + (uint32_t)(tr.idx * 256 + g * MT), umma_desc(a_base + bi * (TK_ROWS * 32)),
                                umma_desc(b_base + bi * (MT * 32)), IDESC);
                    }
                    umma_commit(&tm_full[tr.idx]);
                    tr.next();
                }
                umma_commit(&conv_empty[cr.idx]);
                umma_commit(&act_empty[ar.idx]);
                cr.next(); ar.next();
            }
            TK_PROF_FLUSH(4, true);
        }
        __syncwarp();
    } else if (warp < 4 + TK_CONV_WARPS) {
        // ===================== converters =====================
        reg_dec<TK_REGS_CONV>();
        const int cw = warp - 4, g = lane >> 2, t = lane & 3;
        TK_PROF_DECL;
        RingPos rr(p.nraw), cr(p.nconv);
        for (int x = rg.lo; x < rg.hi; x++) {
            TK_WAIT(&raw_full[rr.idx], rr.ph, 9);
            TK_WAIT(&conv_empty[cr.idx], cr.ph ^ 1u, 10);
            const uint8_t *raw = s_rawr + (size_t)rr.idx * RAW_STAGE;
            uint8_t *conv = s_conv + (size_t)cr.idx * CONV_SLOT;
            float *rs = reinterpret_cast<float *>(conv + CONV_A);
            uint4 q[4];
            float d0[4], m0[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {                    // all loads first
                const uint8_t *ch = raw + (2 * cw + (u >> 1)) * CH;
                q[u] = *reinterpret_cast<const uint4 *>(ch + (u & 1) * 512 + lane * 16);
                d0[u] = *reinterpret_cast<const float *>(ch + 1024 + (((u & 1) * 8 + g) * 4 + t) * 4);
                if constexpr (QT == 1) m0[u] = *reinterpret_cast<const float *>(ch + 1280 + (((u & 1) * 8 + g) * 4 + t) * 4);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int tile = 2 * cw + (u >> 1), half = u & 1;
                const int row = tile * 16 + half * 8 + g;
                uint8_t *dst = conv + (tile * 2 + half) * 256 + (t >> 1) * 128 + g * 16 + (t & 1) * 8;
                int s0[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t wv = u4_get(q[u], i);
                    const uint32_t lo = wv & 0x0F0F0F0Fu, hi = (wv >> 4) & 0x0F0F0F0Fu;
                    *reinterpret_cast<uint2 *>(dst + i * (TK_ROWS * 32)) = make_uint2(lo, hi);
                    if constexpr (QT == 1) s0[i] = __dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, 0u));
                }
                // lane (g, t) forwards the row scales of block t
                rs[(t * NRS + 0) * TK_ROWS + row] = d0[u];
                if constexpr (QT == 1) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], 1);
                        s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], 2);
                    }
                    const int mine = t == 0 ? s0[0] : t == 1 ? s0[1] : t == 2 ? s0[2] : s0[3];
                    rs[(t * NRS + 1) * TK_ROWS + row] = d0[u] * (float)mine;
                    rs[(t * NRS + 2) * TK_ROWS + row] = m0[u];
                }
            }
            fence_proxy_async();                         // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) { mbar_arrive(&conv_full[cr.idx]); mbar_arrive(&raw_empty[rr.idx]); }
            rr.next(); cr.next();
        }
        TK_PROF_FLUSH(8, warp == 4 && lane == 0);
    } else {
        // ===================== epilogue =====================
        reg_inc<TK_REGS_EPI>();
        const int ew = warp - (4 + TK_CONV_WARPS);
        const int q = ew & 3, h = ew >> 2;               // TMEM lane quarter (= warp id % 4), column half
        const int row = q * 32 + lane;                   // weight row inside the tile
        const int etid = ew * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * HC);
        TK_PROF_DECL;
        RingPos ar(p.nact), cr(p.nconv), tr(2);
        int x = rg.lo;
        while (x < rg.hi) {
            const int tile = x / p.C, c0 = x - tile * p.C;
            const int nc = min(p.C - c0, rg.hi - x);
            const int mt = tile % p.m_tiles, nt = tile / p.m_tiles;
            f32x2 acc[HC / 2];
#pragma unroll
            for (int j = 0; j < HC / 2; j++) acc[j] = 0ull;
            for (int c = 0; c < nc; c++) {
                TK_WAIT(&conv_full[cr.idx], cr.ph, 12);  // row scales written by the converters
                TK_WAIT(&act_full[ar.idx], ar.ph, 13);   // column scales landed (TMA)
                const float *rs = reinterpret_cast<const float *>(s_conv + (size_t)cr.idx * CONV_SLOT + CONV_A);
                const float *cs = reinterpret_cast<const float *>(s_act + (size_t)ar.idx * ACT_STAGE + CHUNK_BLKS * MT * 32) + (size_t)(h * HC / 2) * (2 * NCS);
                // software pipeline over the chunk's (block, 16-column) loads: load i+1 is in flight
                // while the FFMA2s of load i run
                constexpr int NL = CHUNK_BLKS * NSB;
                float v[2][LW];
                TK_WAIT(&tm_full[tr.idx], tr.ph, 14);
                tc_fence_after();
                tmem_ld<LW>(t_lane + (uint32_t)(tr.idx * 256), v[0]);
#pragma unroll
                for (int i = 0; i < NL; i++) {
                    const int bi = i / NSB, sb = i % NSB, g = bi % G;
                    tmem_wait_ld<LW>(v[i & 1]);
                    if (g == G - 1 && sb == NSB - 1) {   // that was the last read of this TMEM buffer: hand it back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tm_empty[tr.idx]);
                        tr.next();
                    }
                    if (i + 1 < NL) {
                        const int bn = (i + 1) / NSB, sn = (i + 1) % NSB, gn = bn % G;
                        if (gn == 0 && sn == 0) {        // first read of the next buffer
                            TK_WAIT(&tm_full[tr.idx], tr.ph, 14);
                            tc_fence_after();
                        }
                        tmem_ld<LW>(t_lane + (uint32_t)(tr.idx * 256 + gn * MT + sn * LW), v[(i + 1) & 1]);
                    }
                    const float d0 = rs[(bi * NRS + 0) * TK_ROWS + row];
                    const f32x2 d02 = pk2(d0, d0);
                    const float *csb = cs + ((size_t)bi * (MT / 2) + sb * (LW / 2)) * (2 * NCS);
                    if constexpr (QT == 0) {
#pragma unroll
                        for (int j = 0; j < LW / 2; j++) {
#ifdef TK_EXP_NOLDS
                            const float4 s2 = make_float4(d0, d0, 1.0f, 2.0f);
#else
                            const float4 s2 = *reinterpret_cast<const float4 *>(csb + j * 4);          // d1_a, d1_b, e_a, e_b
#endif
                            const f32x2 t2 = fma2(pk2(v[i & 1][2 * j], v[i & 1][2 * j + 1]), pk2(s2.x, s2.y), pk2(s2.z, s2.w));
                            acc[sb * (LW / 2) + j] = fma2(d02, t2, acc[sb * (LW / 2) + j]);
                        }
                    } else {
                        const float w1 = rs[(bi * NRS + 1) * TK_ROWS + row], m0 = rs[(bi * NRS + 2) * TK_ROWS + row];
                        const f32x2 w12 = pk2(w1, w1), m02 = pk2(m0, m0);
#pragma unroll
                        for (int j = 0; j < LW / 2; j++) {
                            const float4 sa = *reinterpret_cast<const float4 *>(csb + j * 8);          // d1_a, d1_b, m1'_a, m1'_b
                            const float2 sz = *reinterpret_cast<const float2 *>(csb + j * 8 + 4);      // Z'_a, Z'_b
                            const f32x2 t2 = mul2(pk2(v[i & 1][2 * j], v[i & 1][2 * j + 1]), pk2(sa.x, sa.y));
                            f32x2 a2 = fma2(d02, t2, acc[sb * (LW / 2) + j]);
                            a2 = fma2(w12, pk2(sa.z, sa.w), a2);
                            acc[sb * (LW / 2) + j] = fma2(m02, pk2(sz.x, sz.y), a2);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(&conv_empty[cr.idx]); mbar_arrive(&act_empty[ar.idx]); }
                cr.next(); ar.next();
            }
            // ---- this CTA's share of the tile is done: store, or park the piece and let the last one combine ----
            const float post = (QT == 0) ? 16.0f : 32.0f;
            const int n = nt * TK_ROWS + row;
            const int mbase = mt * MT + h * HC;
            const bool whole = (c0 == 0) && (nc == p.C);
            bool store = true;
            if (!whole) {
                const int b = (int)blockIdx.x;
                // piece slot: 0 if the tile began before this CTA's range (head piece), else 1 (tail piece)
                float *ws = p.partial + ((size_t)(2 * b + (c0 == 0 ? 1 : 0)) * MT + (size_t)(h * HC)) * TK_ROWS + row;
#pragma unroll
                for (int j = 0; j < HC / 2; j++) {
                    float a, bq;
                    upk2(acc[j], a, bq);
                    ws[(size_t)(2 * j) * TK_ROWS] = a;
                    ws[(size_t)(2 * j + 1) * TK_ROWS] = bq;
                }
                __threadfence();
                asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");
                const int t0 = tile * p.C;
                const int bf = tk_owner(p, t0, G_), bl = tk_owner(p, t0 + p.C - 1, G_);
                if (etid == 0) {
                    const int old = atomicAdd(p.counters + tile, 1);
                    const int last = (old == bl - bf);
                    *s_flag = last;
                    if (last) p.counters[tile] = 0;
                }
                asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");
                store = *s_flag != 0;
                asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");   // s_flag may be rewritten by the next tile
                if (store) {
                    __threadfence();
                    float sum[HC];
#pragma unroll
                    for (int j = 0; j < HC; j++) sum[j] = 0.0f;
                    for (int bb = bf; bb <= bl; bb++) {      // CTA order: reproducible rounding
                        const int slot = 2 * bb + (t0 < tk_range(p, bb, G_).lo ? 0 : 1);
                        const float *rsrc = p.partial + ((size_t)slot * MT + (size_t)(h * HC)) * TK_ROWS + row;
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
        TK_PROF_FLUSH(11, ew == 0 && lane == 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

} // namespace b200q4
