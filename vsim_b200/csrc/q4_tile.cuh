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

constexpr int TK_ROWS = 128;                 // weight rows per task (MMA M)
constexpr int TK_T16 = TK_ROWS / TILE_ROWS;  // wtile tiles per task
constexpr int TK_CONV_WARPS = 4;
constexpr int TK_EPI_WARPS = 8;
constexpr int TK_THREADS = (4 + TK_CONV_WARPS + TK_EPI_WARPS) * 32;
constexpr int TK_MAX_RAW = 16, TK_MAX_ACT = 8;
constexpr int TK_CONV_SLOTS = 2;             // chunk slots in the converted-A ring

__host__ __device__ constexpr int tk_ncs(int qt) { return qt == 0 ? 2 : 4; }     // column-scale floats per (m, blk)
__host__ __device__ constexpr int tk_nrs(int qt) { return qt == 0 ? 1 : 3; }     // row-scale planes per (n, blk)
__host__ __device__ constexpr int tk_act_stage(int qt, int mt) { return CHUNK_BLKS * mt * 32 + CHUNK_BLKS * mt * tk_ncs(qt) * 4; }
__host__ __device__ constexpr int tk_raw_stage(int qt) { return TK_T16 * chunk_bytes(qt); }
__host__ __device__ constexpr int tk_conv_slot() { return CHUNK_BLKS * TK_ROWS * 32; }
__host__ __device__ constexpr int tk_rs_slot(int qt) { return CHUNK_BLKS * tk_nrs(qt) * TK_ROWS * 4; }
__host__ __device__ constexpr int tk_bar_bytes() { return (2 * TK_MAX_RAW + 2 * TK_MAX_ACT + 2 * TK_CONV_SLOTS + 4) * 8 + 64; }
__host__ __device__ inline size_t tk_smem_bytes(int qt, int mt, int nraw, int nact)
{
    return 1024 + (size_t)nraw * tk_raw_stage(qt) + (size_t)nact * tk_act_stage(qt, mt) +
           (size_t)TK_CONV_SLOTS * (tk_conv_slot() + tk_rs_slot(qt)) + tk_bar_bytes();
}

struct TileParams {
    const uint8_t *wt;       // wtile
    const uint8_t *tpack;    // [m_tiles][C] activation stages
    float         *y;
    float         *partial;  // split-K: [ntasks][MT][128]
    int           *counters; // split-K: one per (n tile, m tile), self-resetting
    size_t         y_stride; // floats
    int            N, M, C, n_t16, m_tiles, PP, PC, ntasks, nraw, nact;
};

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
        float *cs = reinterpret_cast<float *>(stage + (size_t)CHUNK_BLKS * MT * 32) + ((size_t)bi * MT + ml) * tk_ncs(QT);
        if (QT == 0) {
            cs[0] = q.d;
            cs[1] = -0.5f * (q.d * (float)s);
        } else {
            cs[0] = q.d;
            cs[1] = q.mn * 0.03125f;
            cs[2] = fmaf(q.d, (float)s, 32.0f * q.mn) * 0.03125f;
            cs[3] = 0.0f;
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

struct RingPos {
    int idx; uint32_t ph; int n;
    __device__ __forceinline__ RingPos(int n_) : idx(0), ph(0), n(n_) {}
    __device__ __forceinline__ void next() { if (++idx == n) { idx = 0; ph ^= 1u; } }
};

// ---------------------------------------------------------------------------------------------------
// MT: activation rows per task (MMA N): 16, 32, 64 or 128.  A TMEM buffer holds G = min(4, 256/MT)
// block accumulators (G * MT <= 256 columns); two buffers alternate.
// ---------------------------------------------------------------------------------------------------
template <int QT, int MT>
__global__ void __launch_bounds__(TK_THREADS, 1) q4_tile_kernel(const TileParams p)
{
    constexpr int CH = chunk_bytes(QT);
    constexpr int NCS = tk_ncs(QT), NRS = tk_nrs(QT);
    constexpr int RAW_STAGE = tk_raw_stage(QT), ACT_STAGE = tk_act_stage(QT, MT);
    constexpr int CONV_SLOT = tk_conv_slot(), RS_SLOT = tk_rs_slot(QT);
    constexpr int G = (256 / MT) < CHUNK_BLKS ? (256 / MT) : CHUNK_BLKS;      // blocks per TMEM buffer
    constexpr int STEPS = CHUNK_BLKS / G;                                     // TMEM buffers per chunk
    constexpr int HC = MT / 2;                                                // columns per epilogue thread
    constexpr int LW = HC > 32 ? 32 : HC;                                     // columns per tcgen05.ld
    constexpr uint32_t IDESC = umma_idesc(MT);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *s_conv = smem;                                                   // [2 slots][4 blocks][128 rows x 32 B]
    uint8_t *s_act = s_conv + TK_CONV_SLOTS * CONV_SLOT;                      // [nact][ACT_STAGE]
    uint8_t *s_rawr = s_act + (size_t)p.nact * ACT_STAGE;                     // [nraw][RAW_STAGE]
    float   *s_rs = reinterpret_cast<float *>(s_rawr + (size_t)p.nraw * RAW_STAGE);   // [2 slots][4 blocks][NRS][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_rs) + TK_CONV_SLOTS * RS_SLOT);
    uint64_t *raw_full = bars, *raw_empty = raw_full + TK_MAX_RAW;
    uint64_t *act_full = raw_empty + TK_MAX_RAW, *act_empty = act_full + TK_MAX_ACT;
    uint64_t *conv_full = act_empty + TK_MAX_ACT, *conv_empty = conv_full + TK_CONV_SLOTS;
    uint64_t *tm_full = conv_empty + TK_CONV_SLOTS, *tm_empty = tm_full + 2;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(tm_empty + 2);
    int      *s_flag = reinterpret_cast<int *>(s_tmem + 1);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nraw; i++) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], TK_CONV_WARPS); }
        for (int i = 0; i < p.nact; i++) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], 1 + TK_EPI_WARPS); }
        for (int i = 0; i < TK_CONV_SLOTS; i++) { mbar_init(&conv_full[i], TK_CONV_WARPS); mbar_init(&conv_empty[i], 1 + TK_EPI_WARPS); }
        for (int i = 0; i < 2; i++) { mbar_init(&tm_full[i], 1); mbar_init(&tm_empty[i], TK_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    // task -> (n tile nt, m tile mt, first chunk c0, chunk count nc)
    auto task_of = [&](int task, int &nt, int &mt, int &c0, int &nc) {
        const int piece = task % p.PP, tile = task / p.PP;
        mt = tile % p.m_tiles; nt = tile / p.m_tiles;
        c0 = piece * p.PC;
        nc = min(p.PC, p.C - c0);
    };

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            RingPos rr(p.nraw), ar(p.nact);
            for (int task = blockIdx.x; task < p.ntasks; task += gridDim.x) {
                int nt, mt, c0, nc;
                task_of(task, nt, mt, c0, nc);
                const int ntl = min(TK_T16, p.n_t16 - nt * TK_T16);
                for (int c = c0; c < c0 + nc; c++) {
                    mbar_wait(&raw_empty[rr.idx], rr.ph ^ 1u);
                    mbar_arrive_expect_tx(&raw_full[rr.idx], (uint32_t)ntl * CH);
                    for (int tl = 0; tl < ntl; tl++)
                        bulk_g2s(s_rawr + (size_t)rr.idx * RAW_STAGE + tl * CH,
                                 p.wt + ((size_t)(nt * TK_T16 + tl) * p.C + c) * CH, CH, &raw_full[rr.idx]);
                    rr.next();
                    mbar_wait(&act_empty[ar.idx], ar.ph ^ 1u);
                    mbar_arrive_expect_tx(&act_full[ar.idx], ACT_STAGE);
                    bulk_g2s(s_act + (size_t)ar.idx * ACT_STAGE, p.tpack + ((size_t)mt * p.C + c) * ACT_STAGE, ACT_STAGE, &act_full[ar.idx]);
                    ar.next();
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            RingPos ar(p.nact), cr(TK_CONV_SLOTS), tr(2);
            for (int task = blockIdx.x; task < p.ntasks; task += gridDim.x) {
                int nt, mt, c0, nc;
                task_of(task, nt, mt, c0, nc);
                for (int c = 0; c < nc; c++) {
                    mbar_wait(&conv_full[cr.idx], cr.ph);
                    mbar_wait(&act_full[ar.idx], ar.ph);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(s_conv + (size_t)cr.idx * CONV_SLOT);
                    const uint32_t b_base = smem_u32(s_act + (size_t)ar.idx * ACT_STAGE);
#pragma unroll
                    for (int st = 0; st < STEPS; st++) {
                        mbar_wait(&tm_empty[tr.idx], tr.ph ^ 1u);
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
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 4 + TK_CONV_WARPS) {
        // ===================== converters =====================
        const int cw = warp - 4, g = lane >> 2, t = lane & 3;
        RingPos rr(p.nraw), cr(TK_CONV_SLOTS);
        for (int task = blockIdx.x; task < p.ntasks; task += gridDim.x) {
            int nt, mt, c0, nc;
            task_of(task, nt, mt, c0, nc);
            for (int c = 0; c < nc; c++) {
                mbar_wait(&raw_full[rr.idx], rr.ph);
                mbar_wait(&conv_empty[cr.idx], cr.ph ^ 1u);
                const uint8_t *raw = s_rawr + (size_t)rr.idx * RAW_STAGE;
                uint8_t *conv = s_conv + (size_t)cr.idx * CONV_SLOT;
                float *rs = s_rs + (size_t)cr.idx * (RS_SLOT / 4);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int tile = 2 * cw + (u >> 1), half = u & 1;
                    const uint8_t *ch = raw + tile * CH;
                    const uint4 q = *reinterpret_cast<const uint4 *>(ch + half * 512 + lane * 16);
                    const int row = tile * 16 + half * 8 + g;
                    uint8_t *dst = conv + (tile * 2 + half) * 256 + (t >> 1) * 128 + g * 16 + (t & 1) * 8;
                    int s0[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t wv = u4_get(q, i);
                        const uint32_t lo = wv & 0x0F0F0F0Fu, hi = (wv >> 4) & 0x0F0F0F0Fu;
                        *reinterpret_cast<uint2 *>(dst + i * (TK_ROWS * 32)) = make_uint2(lo, hi);
                        if constexpr (QT == 1) s0[i] = __dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, 0u));
                    }
                    // lane (g, t) forwards the row scales of block t
                    const float d0 = *reinterpret_cast<const float *>(ch + 1024 + ((half * 8 + g) * 4 + t) * 4);
                    rs[(t * NRS + 0) * TK_ROWS + row] = d0;
                    if constexpr (QT == 1) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], 1);
                            s0[i] += __shfl_xor_sync(0xffffffffu, s0[i], 2);
                        }
                        const int mine = t == 0 ? s0[0] : t == 1 ? s0[1] : t == 2 ? s0[2] : s0[3];
                        rs[(t * NRS + 1) * TK_ROWS + row] = d0 * (float)mine;
                        rs[(t * NRS + 2) * TK_ROWS + row] = *reinterpret_cast<const float *>(ch + 1280 + ((half * 8 + g) * 4 + t) * 4);
                    }
                }
                fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) { mbar_arrive(&conv_full[cr.idx]); mbar_arrive(&raw_empty[rr.idx]); }
                rr.next(); cr.next();
            }
        }
    } else if (warp >= 4 + TK_CONV_WARPS) {
        // ===================== epilogue =====================
        const int ew = warp - (4 + TK_CONV_WARPS);
        const int q = ew & 3, h = ew >> 2;               // TMEM lane quarter (= warp id % 4), column half
        const int row = q * 32 + lane;                   // weight row inside the task
        const int etid = ew * 32 + lane;
        RingPos ar(p.nact), cr(TK_CONV_SLOTS), tr(2);
        for (int task = blockIdx.x; task < p.ntasks; task += gridDim.x) {
            int nt, mt, c0, nc;
            task_of(task, nt, mt, c0, nc);
            float acc[HC];
#pragma unroll
            for (int j = 0; j < HC; j++) acc[j] = 0.0f;
            for (int c = 0; c < nc; c++) {
                mbar_wait(&conv_full[cr.idx], cr.ph);    // row scales written by the converters
                mbar_wait(&act_full[ar.idx], ar.ph);     // column scales landed (TMA)
                const float *rs = s_rs + (size_t)cr.idx * (RS_SLOT / 4);
                const float *cs = reinterpret_cast<const float *>(s_act + (size_t)ar.idx * ACT_STAGE + CHUNK_BLKS * MT * 32);
#pragma unroll
                for (int st = 0; st < STEPS; st++) {
                    mbar_wait(&tm_full[tr.idx], tr.ph);
                    tc_fence_after();
#pragma unroll
                    for (int g = 0; g < G; g++) {
                        const int bi = st * G + g;
                        const float d0 = rs[(bi * NRS + 0) * TK_ROWS + row];
                        float w1 = 0.f, m0 = 0.f;
                        if constexpr (QT == 1) { w1 = rs[(bi * NRS + 1) * TK_ROWS + row]; m0 = rs[(bi * NRS + 2) * TK_ROWS + row]; }
#pragma unroll
                        for (int l0 = 0; l0 < HC; l0 += LW) {
                            float v[LW];
                            tmem_ld<LW>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tr.idx * 256 + g * MT + h * HC + l0), v);
                            tmem_wait_ld();
                            if (g == G - 1 && l0 + LW >= HC) {       // last read of this TMEM buffer: hand it back
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&tm_empty[tr.idx]);
                            }
                            const float *csb = cs + ((size_t)bi * MT + h * HC + l0) * NCS;
#pragma unroll
                            for (int j = 0; j < LW; j += 2) {
                                if constexpr (QT == 0) {
                                    const float4 s2 = *reinterpret_cast<const float4 *>(csb + j * 2);      // d1, e, d1', e'
                                    acc[l0 + j]     = fmaf(d0, fmaf(v[j], s2.x, s2.y), acc[l0 + j]);
                                    acc[l0 + j + 1] = fmaf(d0, fmaf(v[j + 1], s2.z, s2.w), acc[l0 + j + 1]);
                                } else {
                                    const float4 sa = *reinterpret_cast<const float4 *>(csb + j * 4);      // d1, m1', Z', -
                                    const float4 sb = *reinterpret_cast<const float4 *>(csb + j * 4 + 4);
                                    acc[l0 + j]     = fmaf(m0, sa.z, fmaf(w1, sa.y, fmaf(d0, sa.x * v[j], acc[l0 + j])));
                                    acc[l0 + j + 1] = fmaf(m0, sb.z, fmaf(w1, sb.y, fmaf(d0, sb.x * v[j + 1], acc[l0 + j + 1])));
                                }
                            }
                        }
                    }
                    tr.next();
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(&conv_empty[cr.idx]); mbar_arrive(&act_empty[ar.idx]); }
                cr.next(); ar.next();
            }
            // ---- task done: store, or park the piece and let the last one combine ----
            const float post = (QT == 0) ? 16.0f : 32.0f;
            const int n = nt * TK_ROWS + row;
            const int mbase = mt * MT + h * HC;
            bool store = true;
            if (p.PP > 1) {
                float *ws = p.partial + (size_t)task * (MT * TK_ROWS) + (size_t)(h * HC) * TK_ROWS + row;
#pragma unroll
                for (int j = 0; j < HC; j++) ws[(size_t)j * TK_ROWS] = acc[j];
                __threadfence();
                asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");
                const int tile = task / p.PP;
                if (etid == 0) {
                    const int old = atomicAdd(p.counters + tile, 1);
                    *s_flag = (old == p.PP - 1);
                    if (old == p.PP - 1) p.counters[tile] = 0;
                }
                asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");
                store = *s_flag != 0;
                asm volatile("bar.sync 2, %0;" :: "n"(TK_EPI_WARPS * 32) : "memory");   // s_flag may be rewritten by the next task
                if (store) {
                    __threadfence();
#pragma unroll
                    for (int j = 0; j < HC; j++) acc[j] = 0.0f;
                    for (int k = 0; k < p.PP; k++) {         // piece order: reproducible rounding
                        const float *rsrc = p.partial + (size_t)(tile * p.PP + k) * (MT * TK_ROWS) + (size_t)(h * HC) * TK_ROWS + row;
#pragma unroll
                        for (int j = 0; j < HC; j++) acc[j] += __ldcg(rsrc + (size_t)j * TK_ROWS);
                    }
                }
            }
            if (store && n < p.N) {
#pragma unroll
                for (int j = 0; j < HC; j++)
                    if (mbase + j < p.M) p.y[(size_t)(mbase + j) * p.y_stride + n] = acc[j] * post;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

} // namespace b200q4
