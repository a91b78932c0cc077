// q4_prog.cuh -- the decode "program" kernel: ONE persistent launch walks a device-resident work-list of
// dependent M = 1 mul_mats (and the few non-mul_mat steps a GPT-NeoX decode token needs between them).
//
// Why.  Batch-1 decode is a chain of 169 (GPT-J-6B shapes) / 265 (GPT-NeoX-20B) dependent Q4 mul_mats per token
// (vsim.cpp:518-718 builds them, ggml.c:8543-8662 runs them one node at a time).  As one kernel launch per mul_mat
// every link of that chain pays a kernel boundary (~3 us: drain, launch, refill the weight pipeline) while the
// weights of a 4096 x 4096 tensor stream in 1.6 us.  Here the CTAs stay resident for the whole list:
//
//   producer warp   walks the list AHEAD of everyone else: for every mul_mat op it draws tasks (tile pieces, the
//                   scheduling of q4_gemv_pipe_kernel: static first tasks, then an atomic ticket per K-slice class)
//                   and streams them into the shared-memory stage ring with cp.async.bulk + mbarriers.  Weights
//                   depend on nothing, so the ring keeps filling -- HBM keeps streaming -- while the consumers
//                   wait for the previous op's outputs.  An END marker closes each op, a STOP marker the list.
//   16 consumer     take the stages in order.  At the first stage (or marker) of a new op they (1) wait until the
//   warps           op's dependency has been completed by ALL CTAs (an arrival counter in global memory),
//                   (2) build the op's activation pack in shared memory -- plain vector, sum of K-slice partials,
//                   gelu through the reference's fp16 table (ggml.c:819-827), or layer-norm (ggml.c:4246-4303) of
//                   the residual stream, which every CTA keeps for itself in shared memory and updates with the
//                   same arithmetic (no broadcast, no extra barrier) -- quantized with the bit-exact
//                   quantize_row_q4_0/_q4_1 arithmetic of quantize8_emit, then (3) run the octet IMMA code of the
//                   M = 1 decode kernel on their units.  Attention ops (rope + KV store + softmax(QK)V for the new
//                   token, ggml.c:6086-6153, 5825-5893, 4355-4553) are executed by the consumers between two mul_mat ops.
//   reducer warp    adds the per-unit partial sums of a stage in fixed order (bit-reproducible), stores Y
//                   (+ bias) or parks a K-slice partial, and at an END marker publishes "this CTA is done with op o":
//                   fence + one atomic add.  An op is complete when the counter reaches gridDim.x.
//
// Dependencies are real (op i+1 reads what op i wrote), but their latency -- last store -> fence -> atomic ->
// poll -> rebuild the activation pack, ~1.5 us -- is covered by the stage ring, which holds 160 KB per SM = 3.7 us
// of HBM time chip-wide.  Ops without mutual dependencies (Q, K, V and, with use_parallel_residual = 1, the MLP's
// first matrix, vsim.cpp:666-696) follow each other with no wait at all.
//
// Everything here is M = 1 (one activation row); prompts (M > 1) use the per-op kernels.
#pragma once
#include <cuda_fp16.h>
#include <cstddef>
#include "q4_kernels.cuh"

namespace b200q4 {

constexpr int PG_MAX_SRC = 4;
constexpr int PG_MAX_STAGES = 6;
constexpr int PG_UNITS = 2 * GEMV_WARPS;                 // units per stage (32)
constexpr int PG_RW = 16;                                // floats per unit in the reduction ring (16 weight rows)
constexpr int PG_CTHREADS = GEMV_WARPS * 32;             // consumer threads (512)
// "prep" warps that build the next op's activation pack while the consumers stream.  Measured (B200): with 2 warps a
// 4096-element pack takes 9-12 us (IEEE divisions, 8 items per thread, lowest issue priority) -- longer than the op it is
// meant to hide behind -- so the shipped configuration has none: the 512 consumer threads build every pack themselves
// (about 1.5 us + the wait for the inputs).  The code path is kept for wider-SM parts.
constexpr int PG_PWARPS = 0;
constexpr int PG_PTHREADS = PG_PWARPS * 32;
constexpr int PG_THREADS = (GEMV_WARPS + 2 + PG_PWARPS) * 32;
// Warp roles: [0, PG_PWARPS) prep, then GEMV_WARPS consumers, the producer, the reducer.  The SM's warp arbiter favours
// high warp ids: the prep warps, which may poll for a long time, must not take issue slots from anybody.
constexpr int PG_C0 = PG_PTHREADS;                       // first consumer thread
constexpr int PG_ATT_CHUNK = 64;                         // KV positions per attention work item
constexpr int PG_ATT_MAX_D = 512;                        // largest head size the attention step handles

enum { PG_MATVEC = 0, PG_ATTN = 1, PG_SUM = 2 };
enum { PG_IN_PLAIN = 0, PG_IN_LN = 1, PG_IN_GELU = 2 };
enum { PG_SRC_F32 = 0, PG_SRC_Q4ROW = 1, PG_SRC_TAG = 2 };
enum { PG_RES_KEEP = 0, PG_RES_ADD = 1, PG_RES_SET = 2, PG_RES_TEMP = 3 };

// One input vector:  v[k] = scale * sum_{p < np} ptr[p * stride + k]        (PG_SRC_F32: plain floats; np > 1: K-slice partials)
//                    the same over {value, tag} pairs                       (PG_SRC_TAG: the output of op `kind >> 8` of THIS launch)
//                    v[k] = dequantize_row(ptr, row *row)[k]                (PG_SRC_Q4ROW: ggml_get_rows on the Q4
//                           embedding matrix, ggml.c:300-362; np = quant type, stride = row length K)
// Tagged vectors are how one op hands its result to the next without a fence or a counter: the producer stores every
// value together with a tag (launch epoch, op index) in ONE 8-byte store, the consumer polls the pair until the tag is
// this launch's -- the data is its own flag (the idea of NCCL's LL protocol).  A dependent op starts one store-to-load
// latency after the last output was written, instead of store -> fence -> atomic -> poll -> load.
struct PgSrc {
    const void *ptr;
    const int  *row;
    int   np, stride;
    float scale;
    int   kind;              // low 8 bits: PG_SRC_*; PG_SRC_TAG: the producing op's index << 8
};
inline PgSrc pg_src_tag(const void *p, int op, int np = 1, int stride = 0, float scale = 1.0f) { return PgSrc{p, nullptr, np, stride, scale, PG_SRC_TAG | (op << 8)}; }

struct alignas(16) PgOp {
    int kind;                // PG_MATVEC / PG_ATTN / PG_SUM (y[n] = src[0] summed over its K-slice partials, n < N; spread over all CTAs)
    int dep;                 // index of the op that must be complete before this op's inputs are read (-1: none)
    // ---- PG_MATVEC:  Y[n] = sum_k W[n][k] * quantize(a)[k] (+ bias[n]) ---------------------------------------
    const uint8_t *wt;       // wtile
    int C, n_tiles, N, nb, U;
    int PP, TS, ntasks;      // K-slicing geometry (slice_geometry of q4_api.cu)
    int nstatic;             // tasks per CTA taken without a ticket: task = ((idx + rot) % ncls) + k * ncls, k < nstatic
    int rot;                 // rotation of the static assignment: consecutive small ops hand their odd tasks to different CTAs
    int dynamic;             // != 0: the tasks beyond nstatic * ncls are drawn from a ticket counter
    int tk_off;              // first ticket counter of this op in PgParams::ctr (one per K-slice class)
    int next_mv;             // index of the next PG_MATVEC op (n_ops: none) and what the producer needs of it ahead of time:
    int nx_tk_off, nx_PP, nx_dynamic;    //   its ticket counters (the first ticket is drawn one op early)
    int in_mode;             // a = v (PLAIN) | gelu(v) (GELU) | layernorm(residual)(LN), v = sum of the sources
    int res;                 // LN: KEEP = residual and its statistics as they are; ADD = residual += v first;
                             //     SET = residual = v; TEMP = normalise (residual + v) without storing it
    int nsrc, K;
    PgSrc src[PG_MAX_SRC];
    const float *ln_w, *ln_b;
    const uint16_t *gelu_tab;
    float *y;                // PP == 1: Y;  PP > 1: slice p of the raw sums at y[p * y_slice + n].  y_tag: {value, tag} pairs
    const float *bias;       // PP == 1 only (may be NULL)
    int y_slice, dep_mv;     // dep_mv: the op `dep` is a PG_MATVEC (its completion implies that of every earlier op)
    int y_tag, publish;      // publish: a later op names this one as its `dep`: count the CTAs that are done with it
    float *y_peer[8];        // world > 1: the same output vector on every rank (peer-mapped); this rank's rows go to all of them
    int row0, prebuilt;      // world > 1: first output row of this rank's shard inside the full vector.  prebuilt: this op's pack was
                             // prepared while the previous pack's ops were streaming (it is some earlier op's `early`); 2: and every
                             // consumer warp reads only the quant blocks it wrote itself (a stage is always the same 128 blocks
                             // of the pack: warp w has blocks 8w .. 8w+7), likewise for the pack that had the buffer before: no barrier
    float *y_plain;          // PP == 1, optional: an untagged copy of Y as well (what a caller outside the list reads)
    int cobuild, early;      // early (> 0): index of the LATER op whose pack every consumer warp prepares its share of before it starts
                             // on this op's stages (that op reads one plain vector nobody in the list writes: no pause at all);
                             // cobuild: consumer-built packs: index of a LATER op whose pack is built in the same pause (it needs nothing that
                             // is not there yet: the MLP's layer-norm of the residual stream next to the attention's), or -1
    int builder, wait_pack;  // who builds the pack: 0 the prep warps (ahead of time), 1 the consumers themselves (a residual update or gelu:
                             // big gathers, and the consumers would wait for the inputs anyway); wait_pack: for a prep-built layer-norm
                             // pack, the number of the consumer-built pack that updated the residual stream it reads (-1: none)
    int pack_of, pack_seq;   // index of the first op of a run of ops with the SAME activation pack (Q, K, V): built once;
                             // pack_seq: running number of that pack (its shared-memory buffer is pack_seq & 1)
    const float *early_x;    // early > 0: that op's input vector, tensor geometry and pack number (all pg_consume_op needs to
    int e_C, e_PP, e_TS, e_nb, e_seq;            // prepare its pack between two stages of this op)
    int hold_a, hold_H, pad9;    // hold_a (> 0): this op follows the attention op hold_a - 1 (hold_H heads) without needing its result.
                             // While few CTAs have attention work (items <= grid / 2), the producer of such a CTA draws no task of
                             // this op before its consumers are through the attention: the other CTAs stream the op meanwhile
                             // (all of its tasks come from the ticket counter), instead of everybody waiting for the few
    // (everything above is what a PG_MATVEC needs: the first PG_OPC_WORDS words are what the shared-memory op cache holds)
    // ---- PG_ATTN: the new token's attention for one layer (decode, one query row) -----------------------------
    const float *q, *k, *v;          // [H * d] {value, tag} pairs written by ops tq, tk, tv (biases already added)
    int tq, tk, tv, pad4;
    float *mem_k, *mem_v;            // this layer's KV memory [n_ctx][H * d], vsim.cpp:318-330
    float *att_out;                  // [H * d] merged heads (vsim.cpp:621-627), {value, tag} pairs
    float *att_peer[8];              // world > 1: att_out on every rank; H is then this rank's head count, head0 its first head
    int head0, pad8;
    float *att_part;                 // [H][max_chunks][d] weighted V sums per position chunk (several chunks per head only)
    float *att_sc;                   // [H][n_ctx] scaled scores, [H][max_chunks] chunk maxima behind them
    double *att_sum;                 // [H][max_chunks] chunk denominators
    const double2 *rope;             // [n_ctx][n_rot / 2] {cos, sin}(p * theta_i), host libm like ggml.c:6126-6129
    const uint16_t *exp_tab;         // ggml_init's fp16 exp table (ggml.c:1240-1255), what ggml_soft_max looks its exponentials up in
    int H, d, n_rot, max_chunks, cnt_off, n_ctx;   // cnt_off: 2 * H counters (scores published / partials published)
    float att_scale; int pad2;
};

constexpr int PG_OPC = 8;                                // mul_mat ops cached in shared memory (ring; > PG_MAX_STAGES + 1)
constexpr int PG_OPC_WORDS = 108;                        // 432 bytes per entry: one 16-byte load for each of 27 producer lanes (the PG_MATVEC part of PgOp)
static_assert(offsetof(PgOp, q) <= PG_OPC_WORDS * 4, "the PG_MATVEC part of PgOp must fit a cache entry");

struct PgParams {
    const PgOp *ops;
    int  n_ops, first_mv;
    int *ctr;                // [0] CTAs finished; [1 + o] CTAs done with op o, summed over all launches (never reset); tickets;
    int  n_ctr;              //   attention head counters (zero between launches: the last CTA to leave clears them).
    int *dyn;                // [0] token id (through PgSrc::row), [1] n_past, [2] launch epoch (the last CTA to leave advances it)
    int *err;                // set (op index + 1) before the kernel traps on a wait that can never end
    int  stages, e_cap, nbp_cap, att_cap;    // shared-memory sizing: residual floats, apack quant blocks, attention floats
    int  wt_evict_first;     // != 0: the weight stream carries an L2 evict-first hint (what else a token touches stays in L2)
    int  world, rank;        // row-sharded over `world` GPUs (1: single GPU)
    int *ctr_peer[8];        // world > 1: every rank's counter array (hint counters are added to on all ranks)
    unsigned long long *dbg; // optional timeline [grid][n_ops][8] of %globaltimer stamps (B200Q4_DBG=1), else NULL
};

__host__ __device__ constexpr size_t pg_stage_bytes(int qt) { return (size_t)PG_UNITS * chunk_bytes(qt) + (size_t)PG_UNITS * PG_RW * 4; }
__host__ __device__ constexpr size_t pg_fixed_bytes(int e_cap, int nbp_cap, int att_cap)
{
    return (size_t)e_cap * 4 + (size_t)nbp_cap * 32 + (size_t)nbp_cap * 12 + (size_t)att_cap * 4 + 40 * 8 +
           (size_t)nbp_cap * 44 /* the second pack buffer */ + 4 * 8 + 32 +
           3 * PG_MAX_STAGES * 8 + PG_MAX_STAGES * 16 + 64 + 128 + GEMV_WARPS * 16 + (PG_OPC + 1) * PG_OPC_WORDS * 4;
}

#ifndef Q4_PROG_TYPES_ONLY                               // (host-only includers stop here)

// Shared-memory map of q4_prog_kernel.  The out-of-line helpers recompute it from the parameter copy instead of taking
// pointers as arguments: whatever the main loop does not need stays out of its registers.
struct PgSmem {
    uint8_t *stage; float *red, *x; uint8_t *aq; float *ap, *att; double *stat;     // aq / ap: pack buffer 0; buffer 1 follows at pack_stride bytes
    uint64_t *bar_full, *bar_empty, *bar_done, *pack_full, *pack_empty; int4 *desc; int *dop, *flag, *wst, *geo; uint32_t *opc;
    int pack_stride;
};
extern __shared__ __align__(128) uint8_t pg_smem_base[];
__device__ __forceinline__ PgSmem pg_smem(const PgParams &p, int qt)
{
    PgSmem m;
    m.stage = pg_smem_base;
    m.red = reinterpret_cast<float *>(pg_smem_base + (size_t)p.stages * PG_UNITS * chunk_bytes(qt));
    m.x = m.red + (size_t)p.stages * PG_UNITS * PG_RW;
    m.pack_stride = p.nbp_cap * 44;                                    // one pack: codes [nbp_cap][32] + three planes [nbp_cap] floats
    m.aq = reinterpret_cast<uint8_t *>(m.x + p.e_cap);
    m.ap = reinterpret_cast<float *>(m.aq + (size_t)p.nbp_cap * 32);
    m.att = reinterpret_cast<float *>(m.aq + 2 * (size_t)m.pack_stride);
    m.stat = reinterpret_cast<double *>(m.att + p.att_cap);
    m.bar_full = reinterpret_cast<uint64_t *>(m.stat + 40);
    m.bar_empty = m.bar_full + PG_MAX_STAGES;
    m.bar_done = m.bar_empty + PG_MAX_STAGES;
    m.pack_full = m.bar_done + PG_MAX_STAGES;                          // [2] the prep warps have built pack buffer b
    m.pack_empty = m.pack_full + 2;                                    // [2] every consumer warp is done with pack buffer b
    m.desc = reinterpret_cast<int4 *>(m.pack_empty + 2);               // {units (0: none, -2: STOP), first chunk index, first tile, task}
    m.dop = reinterpret_cast<int *>(m.desc + PG_MAX_STAGES);           // op index of the stage | PG_END when it is the CTA's last of that op
    m.flag = m.dop + PG_MAX_STAGES + 2;                                // [0] scratch flag, [1] waited_mv (consumers), [2] launch epoch, [5] latest consumer-built pack, [6] tag base of this launch,
                                                                       // [7] waited_mv (prep warps)
    m.geo = m.flag + 8;                                                // [2][4] per pack buffer: C, first chunk of the K-slice, plane stride
    m.wst = m.geo + 8;                                                 // per consumer warp: slot, parity, current op, current pack
    m.opc = reinterpret_cast<uint32_t *>(m.wst + GEMV_WARPS * 4);      // op cache [PG_OPC][PG_OPC_WORDS]
    return m;
}

#ifdef B200Q4_TIMELINE
#define PG_STAMP(o, slot) do { if (p.dbg && (o) < p.n_ops) p.dbg[((size_t)blockIdx.x * p.n_ops + (o)) * 8 + (slot)] = gtimer(); } while (0)
// wait-time accumulators behind the stamps: [grid][16] cycles (0 producer waiting for a free slot, 1 consumer warp 0
// waiting for a full stage, 2 consumer warp 0 in pg_begin_op, 3 reducer waiting for the consumers, 4 reducer total)
#define PG_ACC(pp_, slot, cyc) do { if ((pp_)->dbg) atomicAdd((pp_)->dbg + (size_t)gridDim.x * (pp_)->n_ops * 8 + (size_t)blockIdx.x * 16 + (slot), (unsigned long long)(cyc)); } while (0)
#define PG_CLK() clock64()
#else
#define PG_STAMP(o, slot) do { } while (0)
#define PG_ACC(pp_, slot, cyc) do { } while (0)
#define PG_CLK() 0ll
#endif

__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *c >= target.  A counter that never gets there is a bug (or a dead peer): after ~2 s the kernel records
// where and traps instead of hanging the GPU.
__device__ __forceinline__ void pg_wait_ctr(const int *c, int target, int *err, int code)
{
    const long long t0 = clock64();
    for (unsigned spins = 0;; spins++) {
        if (ld_acquire_gpu(c) >= target) return;
        if (spins > 4096) __nanosleep(64);
        if ((spins & 1023) == 1023 && clock64() - t0 > 4000000000ll) {
            if (err) *err = code;
            __threadfence_system();
            __trap();
        }
    }
}
// "probably complete" hint: CTAs add to an op's counter after their last (tagged) stores WITHOUT a fence, one thread
// polls it with back-off, then everybody reads the tagged data (which is what guarantees correctness).  Gives up after
// ~1 ms: the tags decide.
__device__ __forceinline__ void pg_wait_hint(const int *c, int target)
{
    const long long t0 = clock64();
    for (;;) {
        int v;
        asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
        if (v >= target || clock64() - t0 > 2000000ll) return;
        __nanosleep(100);
    }
}
__device__ __forceinline__ void pg_add_hint(int *c)
{
    asm volatile("red.relaxed.sys.global.add.s32 [%0], 1;" :: "l"(c) : "memory");
}
// the hint of op o on every rank (sharded programs: a consumer waits for the CTAs of all ranks)
__device__ __forceinline__ void pg_add_hint_all(const PgParams &p, int o)
{
    if (p.world <= 1) { pg_add_hint(p.ctr + 1 + o); return; }
    for (int r = 0; r < p.world; r++) pg_add_hint(p.ctr_peer[r] + 1 + o);
}
__device__ __forceinline__ void pg_mbar_wait(uint64_t *b, uint32_t parity)
{
    const long long t0 = clock64();
    for (unsigned spins = 0;; spins++) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (done) return;
        if ((spins & 255) == 255 && clock64() - t0 > 8000000000ll) __trap();
    }
}
__device__ __forceinline__ void pg_bar() { asm volatile("bar.sync 1, %0;" :: "n"(PG_CTHREADS) : "memory"); }
__device__ __forceinline__ void pg_pbar() { asm volatile("bar.sync 2, %0;" :: "n"(PG_PTHREADS > 0 ? PG_PTHREADS : 32) : "memory"); }   // the prep warps' own barrier

__device__ __forceinline__ float4 ld_relaxed_f4(const void *p)
{
    float4 r;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ld_relaxed_f2(const void *p)
{
    float2 r;
    asm volatile("ld.relaxed.sys.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_tagged(float *y2, size_t idx, float v, int tag)
{
    asm volatile("st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" :: "l"(y2 + 2 * idx), "f"(v), "f"(__int_as_float(tag)) : "memory");
}
__device__ __noinline__ void pg_tag_timeout(int *err, int code)
{
    if (err) *err = code;
    __threadfence_system();
    __trap();
}
// one tagged element: spins until the pair carries this launch's tag
__device__ __forceinline__ float pg_ld_tag1(const float *base2, size_t idx, int tag, int *err)
{
    const long long t0 = clock64();
    for (unsigned spins = 0;; spins++) {
        const float2 v = ld_relaxed_f2(base2 + 2 * idx);
        if (__float_as_int(v.y) == tag) return v.x;
        if ((spins & 1023) == 1023 && clock64() - t0 > 4000000000ll) pg_tag_timeout(err, -tag);
    }
}
// four consecutive tagged elements (idx % 4 == 0)
__device__ __forceinline__ float4 pg_ld_tag4(const float *base2, size_t idx, int tag, int *err)
{
    const long long t0 = clock64();
    for (unsigned spins = 0;; spins++) {
        const float4 a = ld_relaxed_f4(base2 + 2 * idx), b = ld_relaxed_f4(base2 + 2 * idx + 4);
        if (__float_as_int(a.y) == tag && __float_as_int(a.w) == tag && __float_as_int(b.y) == tag && __float_as_int(b.w) == tag)
            return make_float4(a.x, a.z, b.x, b.z);
        if ((spins & 1023) == 1023 && clock64() - t0 > 4000000000ll) pg_tag_timeout(err, -tag);
    }
}

// four consecutive elements k .. k+3 of one source (k % 4 == 0)
__device__ __forceinline__ float4 pg_src4(const PgSrc &s, int k, int tagbase, int *err)
{
    const int kind = s.kind & 255;
    if (kind == PG_SRC_TAG) {
        const int tag = tagbase | ((s.kind >> 8) + 1);
        const float *b = static_cast<const float *>(s.ptr);
        float4 a = pg_ld_tag4(b, (size_t)k, tag, err);
        for (int p = 1; p < s.np; p++) {
            const float4 t = pg_ld_tag4(b, (size_t)p * s.stride + k, tag, err);
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        if (s.scale != 1.0f) { a.x *= s.scale; a.y *= s.scale; a.z *= s.scale; a.w *= s.scale; }
        return a;
    }
    if (kind == PG_SRC_F32) {
        const float *b = static_cast<const float *>(s.ptr) + k;
        float4 a = __ldcg(reinterpret_cast<const float4 *>(b));
        for (int p = 1; p < s.np; p++) {
            const float4 t = __ldcg(reinterpret_cast<const float4 *>(b + (size_t)p * s.stride));
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        if (s.scale != 1.0f) { a.x *= s.scale; a.y *= s.scale; a.z *= s.scale; a.w *= s.scale; }
        return a;
    }
    // dequantize_row_q4_0 (ggml.c:300-330): (q - 8) * d;  dequantize_row_q4_1 (:332-362): q * d + m
    const int type = s.np, nb = s.stride / QK, blk = k / QK, l = k % QK;
    const uint8_t *row = static_cast<const uint8_t *>(s.ptr) + (size_t)(*s.row) * nb * (type == 0 ? 20 : 24);
    float r[4];
    if (type == 0) {
        const uint8_t *b = row + (size_t)blk * 20;
        const float d = *reinterpret_cast<const float *>(b);
        const uint16_t two = *reinterpret_cast<const uint16_t *>(b + 4 + l / 2);      // l % 4 == 0: two bytes, 2-byte aligned
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = __fmul_rn((float)((int)((two >> (4 * i)) & 15) - 8), d);
    } else {
        const float m = reinterpret_cast<const float *>(row)[blk], d = reinterpret_cast<const float *>(row)[nb + blk];
        const uint16_t two = *reinterpret_cast<const uint16_t *>(row + (size_t)nb * 8 + (size_t)blk * 16 + l / 2);
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = __fadd_rn(__fmul_rn((float)((two >> (4 * i)) & 15), d), m);
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}
// sum of the op's sources in list order (fixed association: every CTA computes the same bits)
__device__ __forceinline__ float4 pg_sum_src4(const PgOp &op, int nsrc, int k, int tagbase, int *err)
{
    float4 a = pg_src4(op.src[0], k, tagbase, err);
    for (int j = 1; j < nsrc; j++) {
        const float4 t = pg_src4(op.src[j], k, tagbase, err);
        a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
    }
    return a;
}

// pg_sum_src4 for TWO quads of elements at once (kA, and kB when hasB): for every (source, partial) pair the loads of both
// quads are issued before either is looked at, which halves the number of memory round trips of a gather (a layer-norm
// after a K-sliced MLP reads 6 pairs per quad, two quads per thread).  Same additions in the same order: same bits.
__device__ __forceinline__ void pg_sum_src4x2(const PgOp &op, int nsrc, int kA, int kB, bool hasB, int tagbase, int *err, float4 &oA, float4 &oB)
{
    float4 accA = make_float4(0.f, 0.f, 0.f, 0.f), accB = accA;
    for (int j = 0; j < nsrc; j++) {
        const PgSrc &s = op.src[j];
        const int kind = s.kind & 255;
        float4 a, b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kind == PG_SRC_Q4ROW) {
            a = pg_src4(s, kA, tagbase, err);
            if (hasB) b = pg_src4(s, kB, tagbase, err);
        } else {
            const int tag = tagbase | ((s.kind >> 8) + 1);
            for (int p = 0; p < s.np; p++) {
                float4 tA, tB = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kind == PG_SRC_TAG) {
                    const float *base = static_cast<const float *>(s.ptr) + 2 * (size_t)p * s.stride;
                    const float4 a0 = ld_relaxed_f4(base + 2 * (size_t)kA), a1 = ld_relaxed_f4(base + 2 * (size_t)kA + 4);
                    float4 b0 = a0, b1 = a1;
                    if (hasB) { b0 = ld_relaxed_f4(base + 2 * (size_t)kB); b1 = ld_relaxed_f4(base + 2 * (size_t)kB + 4); }
                    if (__float_as_int(a0.y) == tag && __float_as_int(a0.w) == tag && __float_as_int(a1.y) == tag && __float_as_int(a1.w) == tag)
                        tA = make_float4(a0.x, a0.z, a1.x, a1.z);
                    else tA = pg_ld_tag4(base, (size_t)kA, tag, err);
                    if (hasB) {
                        if (__float_as_int(b0.y) == tag && __float_as_int(b0.w) == tag && __float_as_int(b1.y) == tag && __float_as_int(b1.w) == tag)
                            tB = make_float4(b0.x, b0.z, b1.x, b1.z);
                        else tB = pg_ld_tag4(base, (size_t)kB, tag, err);
                    }
                } else {
                    const float *base = static_cast<const float *>(s.ptr) + (size_t)p * s.stride;
                    tA = __ldcg(reinterpret_cast<const float4 *>(base + kA));
                    if (hasB) tB = __ldcg(reinterpret_cast<const float4 *>(base + kB));
                }
                if (p == 0) { a = tA; b = tB; }
                else {
                    a.x += tA.x; a.y += tA.y; a.z += tA.z; a.w += tA.w;
                    b.x += tB.x; b.y += tB.y; b.z += tB.z; b.w += tB.w;
                }
            }
            if (s.scale != 1.0f) {
                a.x *= s.scale; a.y *= s.scale; a.z *= s.scale; a.w *= s.scale;
                b.x *= s.scale; b.y *= s.scale; b.z *= s.scale; b.w *= s.scale;
            }
        }
        if (j == 0) { accA = a; accB = b; }
        else {
            accA.x = __fadd_rn(accA.x, a.x); accA.y = __fadd_rn(accA.y, a.y); accA.z = __fadd_rn(accA.z, a.z); accA.w = __fadd_rn(accA.w, a.w);
            accB.x = __fadd_rn(accB.x, b.x); accB.y = __fadd_rn(accB.y, b.y); accB.z = __fadd_rn(accB.z, b.z); accB.w = __fadd_rn(accB.w, b.w);
        }
    }
    oA = accA; oB = accB;
}

// ---------------------------------------------------------------------------------------------------------------
// Attention for the token at position n_past (decode): work items = (head, chunk of PG_ATT_CHUNK positions), dealt
// to the CTAs round-robin.  An item ropes Q (and, if its chunk holds the new position, ropes K and stores K and V into
// the KV memory), computes its chunk's scores, a local softmax (max, denominator) and the weighted V sum; with several
// chunks per head the last item of a head to finish merges the partials.  All 512 consumer threads take part.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pg_rope(float x_self, float x_pair, bool first_half, double2 cs)
{
    // ggml.c:6131-6138: r1 = cos*x1 - sin*x2 (first half), r2 = cos*x2 + sin*x1 (second half), in double
    return first_half ? (float)__dsub_rn(__dmul_rn(cs.x, (double)x_self), __dmul_rn(cs.y, (double)x_pair))
                      : (float)__dadd_rn(__dmul_rn(cs.x, (double)x_self), __dmul_rn(cs.y, (double)x_pair));
}

// Work decomposition: items = (head, chunk of CHK positions).  CHK is a multiple of PG_ATT_CHUNK chosen so that the
// items about fill the grid once (16 heads x 512 positions: 8 chunks of 64; 64 heads: 2 chunks of 256).
struct PgAttGeo { int nch, chk; };
__device__ __forceinline__ PgAttGeo pg_att_geo(int P, int H, int grid)
{
    int nch = (P + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK;
    const int cap = grid / H > 1 ? grid / H : 1;             // chunks per head that still give every item its own CTA
    if (nch > cap) nch = cap;
    int chk = (P + nch - 1) / nch;
    chk = (chk + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK * PG_ATT_CHUNK;
    PgAttGeo g;
    g.chk = chk; g.nch = (P + chk - 1) / chk;
    return g;
}

// scores of the chunk's positions [p_lo, p_lo + np) against s_q: LPP lanes share a position (each up to four float4 of
// the K row, all requested before the first product), float products summed in double (ggml.c:4484-4516), then ggml_scale
__device__ __forceinline__ void pg_att_scores(const float *mk, const float *s_q, const float *s_kn, float *s_sc, int p_lo, int np,
                                              int n_past, int E, int hd0, int d, float scale, int ctid)
{
    const int lane = ctid & 31, warp = ctid >> 5;
    if ((d & 3) == 0 && d <= 512) {
        const int d4 = d >> 2;
        int lpp = 1;
        while (lpp * 4 < d4) lpp <<= 1;                      // lanes per position: each lane at most 4 float4
        const int ppw = 32 / lpp, sub = lane & (lpp - 1), slot = lane / lpp;
        for (int pl0 = warp * ppw; pl0 < np; pl0 += GEMV_WARPS * ppw) {
            const int pl = pl0 + slot, pos = p_lo + pl;
            const bool live = pl < np;
            float4 kv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int f = sub + i * lpp;
                kv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && f < d4) {
                    if (pos == n_past) kv[i] = reinterpret_cast<const float4 *>(s_kn)[f];
                    else kv[i] = __ldcg(reinterpret_cast<const float4 *>(mk + (size_t)pos * E + hd0) + f);
                }
            }
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int f = sub + i * lpp;
                if (f < d4) {
                    const float4 q4 = reinterpret_cast<const float4 *>(s_q)[f];
                    sum = __dadd_rn(sum, __dadd_rn(__dadd_rn((double)__fmul_rn(kv[i].x, q4.x), (double)__fmul_rn(kv[i].y, q4.y)),
                                                   __dadd_rn((double)__fmul_rn(kv[i].z, q4.z), (double)__fmul_rn(kv[i].w, q4.w))));
                }
            }
            for (int of = lpp >> 1; of > 0; of >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, of));
            if (live && sub == 0) s_sc[pl] = __fmul_rn((float)sum, scale);
        }
    } else {
        for (int pl = warp; pl < np; pl += GEMV_WARPS) {
            const int pos = p_lo + pl;
            double sum = 0.0;
            if (pos == n_past) { for (int j = lane; j < d; j += 32) sum = __dadd_rn(sum, (double)__fmul_rn(s_kn[j], s_q[j])); }
            else {
                const float *kr = mk + (size_t)pos * E + hd0;
#pragma unroll 8
                for (int j = lane; j < d; j += 32) sum = __dadd_rn(sum, (double)__fmul_rn(__ldcg(kr + j), s_q[j]));
            }
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, of));
            if (lane == 0) s_sc[pl] = __fmul_rn((float)sum, scale);
        }
    }
}

// out[j] = sum over the chunk's positions of V[pos][j] * w[pos]  (ggml.c:4517-4553): position groups in parallel (float4
// columns, eight V rows requested at a time), merged in group order.  s_o: scratch [groups][d].  Result in s_o[0 .. d).
__device__ __forceinline__ void pg_att_pv(const float *mv, const float *s_vn, const float *s_w, float *s_o, int p_lo, int np,
                                          int n_past, int E, int hd0, int d, int ctid)
{
    if ((d & 3) == 0 && d <= 512) {
        const int d4 = d >> 2;
        int cols = 8;
        while (cols < d4) cols <<= 1;                        // threads per position group (a power of two >= d / 4)
        const int ng = PG_CTHREADS / cols, g = ctid / cols, c = ctid & (cols - 1);
        if (c < d4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int pl0 = g; pl0 < np; pl0 += 8 * ng) {
                float4 vv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int pl = pl0 + u * ng, pos = p_lo + pl;
                    vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pl < np) vv[u] = pos == n_past ? reinterpret_cast<const float4 *>(s_vn)[c]
                                                       : __ldcg(reinterpret_cast<const float4 *>(mv + (size_t)pos * E + hd0) + c);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int pl = pl0 + u * ng;
                    if (pl < np) {
                        const float w = s_w[pl];
                        a.x = __fadd_rn(a.x, __fmul_rn(vv[u].x, w)); a.y = __fadd_rn(a.y, __fmul_rn(vv[u].y, w));
                        a.z = __fadd_rn(a.z, __fmul_rn(vv[u].z, w)); a.w = __fadd_rn(a.w, __fmul_rn(vv[u].w, w));
                    }
                }
            }
            reinterpret_cast<float4 *>(s_o + (size_t)g * d)[c] = a;
        }
        pg_bar();
        if (ctid < d) {
            float a = s_o[ctid];
            for (int gg = 1; gg < ng; gg++) a = __fadd_rn(a, s_o[(size_t)gg * d + ctid]);
            s_o[ctid] = a;                                   // (thread ctid is the only reader of column ctid of the other groups)
        }
    } else {
        const int dj = (d + 31) & ~31, ng = PG_CTHREADS / dj, g = ctid / dj, j = ctid - g * dj;
        if (g < ng && j < d) {
            float a = 0.f;
#pragma unroll 8
            for (int pl = g; pl < np; pl += ng) {
                const int pos = p_lo + pl;
                const float vv = pos == n_past ? s_vn[j] : __ldcg(mv + (size_t)pos * E + hd0 + j);
                a = __fadd_rn(a, __fmul_rn(vv, s_w[pl]));
            }
            s_o[(size_t)g * dj + j] = a;
        }
        pg_bar();
        if (ctid < d) {
            float a = s_o[ctid];
            for (int gg = 1; gg < ng; gg++) a = __fadd_rn(a, s_o[(size_t)gg * dj + ctid]);
            s_o[ctid] = a;
        }
    }
    pg_bar();
}

__device__ __noinline__ void pg_attn(const PgParams *pp, int o, int qt)
{
    const PgParams &p = *pp;                                 // (a copy in shared memory: nothing here may live on the stack)
    const PgSmem sm = pg_smem(p, qt);
    float *s_att = sm.att;
    int *s_flag = sm.flag;
    const int n_past = p.dyn ? __ldg(p.dyn + 1) : 0;
    const PgOp &op = p.ops[o];
    const int ctid = threadIdx.x - PG_C0, lane = ctid & 31, warp = ctid >> 5;
    const int H = op.H, d = op.d, E = H * d, half = op.n_rot / 2, P = n_past + 1;
    const PgAttGeo ag = pg_att_geo(P, H, (int)gridDim.x);
    const int nch = ag.nch, CHK = ag.chk, items = H * nch;
    const int dpad = (d + 3) & ~3, scap = (op.n_ctx + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK * PG_ATT_CHUNK;
    float *s_q = s_att, *s_kn = s_q + dpad, *s_vn = s_kn + dpad, *s_sc = s_vn + dpad, *s_o = s_sc + scap + 8;
    const float *gq = op.q, *gk = op.k, *gv = op.v;
    float *mk = op.mem_k, *mv = op.mem_v;
    const float scale = op.att_scale;
    const uint16_t *etab = op.exp_tab;
    float *g_sc = op.att_sc, *g_mx = op.att_sc + (size_t)H * op.n_ctx;
    int *cnt1 = p.ctr + op.cnt_off, *cnt2 = cnt1 + H;
    const int tagbase = sm.flag[6];
    const int tag_q = tagbase | (op.tq + 1), tag_k = tagbase | (op.tk + 1), tag_v = tagbase | (op.tv + 1), tag_o = tagbase | (o + 1);
    // ggml_soft_max's exponential (ggml.c:5862-5870): fp16(x - max) -> table -> fp32
    auto texp = [&](float x) { return __half2float(__ushort_as_half(__ldg(etab + __half_as_ushort(__float2half_rn(x))))); };
    // the merged-heads vector on this rank, or (sharded: H is this rank's head count) on every rank
    auto att_store = [&](size_t idx, float v) {
        if (p.world <= 1) st_tagged(op.att_out, idx, v, tag_o);
        else for (int r = 0; r < p.world; r++) st_tagged(op.att_peer[r], idx, v, tag_o);
    };
    // warp 0: maximum of s_sc[0 .. np)
    auto chunk_max = [&](int np) {
        float m = -INFINITY;
        for (int i = lane; i < np; i += 32) m = fmaxf(m, s_sc[i]);
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));
        return m;
    };
    // warp 0: s_sc[i] = exp-table(s_sc[i] - M) in place; returns their sum (double, lane-strided partial sums added in a tree)
    auto chunk_exp = [&](int np, float M) {
        double l = 0.0;
        for (int i0 = lane; i0 < np; i0 += 128) {            // four table look-ups in flight per lane; added in index order
            float e[4];
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] = i0 + 32 * u < np ? texp(__fsub_rn(s_sc[i0 + 32 * u], M)) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + 32 * u < np) { s_sc[i0 + 32 * u] = e[u]; l = __dadd_rn(l, (double)e[u]); }
        }
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) l = __dadd_rn(l, __shfl_xor_sync(0xffffffffu, l, of));
        return l;
    };

    // ---- pass A: rope, KV store, the chunk's scaled scores and their maximum --------------------------------------
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int h = item / nch, c = item - h * nch;
        const int p_lo = c * CHK, np = min(CHK, P - p_lo);
        const bool has_new = p_lo + np == P;
        pg_bar();                                            // the previous item's scratch is no longer read
        if (ctid == 0) PG_STAMP(o, 0);
        for (int j = ctid; j < d; j += PG_CTHREADS) {
            // everything this element needs is requested at once (one memory round trip); a value whose tag is not this
            // launch's yet is polled for afterwards
            const bool roped = j < 2 * half, fh = j < half;
            const int jp = roped ? (fh ? j + half : j - half) : j, i0 = fh ? j : j - half;
            const size_t e = (size_t)h * d + j, ep = (size_t)h * d + jp;
            const float2 q2 = ld_relaxed_f2(gq + 2 * e), qp2 = ld_relaxed_f2(gq + 2 * ep);
            float2 k2 = make_float2(0.f, 0.f), kp2 = k2, v2 = k2;
            if (has_new) { k2 = ld_relaxed_f2(gk + 2 * e); kp2 = ld_relaxed_f2(gk + 2 * ep); v2 = ld_relaxed_f2(gv + 2 * e); }
            double2 cs = make_double2(1.0, 0.0);
            if (roped) cs = op.rope[(size_t)n_past * half + i0];
            float qv = __float_as_int(q2.y) == tag_q ? q2.x : pg_ld_tag1(gq, e, tag_q, p.err), kv = 0.f, vv = 0.f;
            if (has_new) {
                kv = __float_as_int(k2.y) == tag_k ? k2.x : pg_ld_tag1(gk, e, tag_k, p.err);
                vv = __float_as_int(v2.y) == tag_v ? v2.x : pg_ld_tag1(gv, e, tag_v, p.err);
            }
            if (roped) {
                qv = pg_rope(qv, __float_as_int(qp2.y) == tag_q ? qp2.x : pg_ld_tag1(gq, ep, tag_q, p.err), fh, cs);
                if (has_new) kv = pg_rope(kv, __float_as_int(kp2.y) == tag_k ? kp2.x : pg_ld_tag1(gk, ep, tag_k, p.err), fh, cs);
            }
            s_q[j] = qv;
            if (has_new) {
                s_kn[j] = kv; s_vn[j] = vv;
                mk[(size_t)n_past * E + h * d + j] = kv;     // cpy into the KV memory (vsim.cpp:557-558), K roped in place (:573-579)
                mv[(size_t)n_past * E + h * d + j] = vv;
            }
        }
        pg_bar();
        if (ctid == 0) PG_STAMP(o, 2);
        pg_att_scores(mk, s_q, s_kn, s_sc, p_lo, np, n_past, E, h * d, d, scale, ctid);
        pg_bar();
        if (ctid == 0) PG_STAMP(o, 4);
        if (nch > 1) {                                       // publish the chunk: the row maximum needs every chunk
            for (int i = ctid; i < np; i += PG_CTHREADS) g_sc[(size_t)h * op.n_ctx + p_lo + i] = s_sc[i];
            if (warp == 0) {
                const float m = chunk_max(np);
                if (lane == 0) g_mx[(size_t)h * op.max_chunks + c] = m;
            }
            __threadfence();
            pg_bar();
            if (ctid == 0) atomicAdd(cnt1 + h, 1);
            continue;
        }
        // ---- one chunk per head: soft_max and V^T . softmax right here (ggml.c:5825-5893, 4517-4553) ----------------
        if (warp == 0) {
            const float M = chunk_max(np);
            const double l = chunk_exp(np, M);
            const float inv = (float)(1.0 / l);
            __syncwarp();
            for (int i = lane; i < np; i += 32) s_sc[i] = __fmul_rn(s_sc[i], inv);
        }
        pg_bar();
        if (ctid == 0) PG_STAMP(o, 5);
        pg_att_pv(mv, s_vn, s_sc, s_o, p_lo, np, n_past, E, h * d, d, ctid);
        if (ctid == 0) PG_STAMP(o, 6);
        if (ctid < d) att_store((size_t)(op.head0 + h) * d + ctid, s_o[ctid]);
        if (ctid == 0) PG_STAMP(o, 7);
    }
    // ---- pass B (several chunks per head): exponentials against the ROW maximum, chunk denominators, weighted V sums; the
    //      head's last chunk to finish adds the pieces up.  Every CTA has published all of its chunks before it waits here.
    if (nch > 1) {
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int h = item / nch, c = item - h * nch;
            const int p_lo = c * CHK, np = min(CHK, P - p_lo);
            pg_bar();
            if (ctid == 0) pg_wait_ctr(cnt1 + h, nch, p.err, o + 1);
            pg_bar();
            for (int i = ctid; i < np; i += PG_CTHREADS) s_sc[i] = __ldcg(g_sc + (size_t)h * op.n_ctx + p_lo + i);
            if (p_lo + np == P)                              // the new position's V row (its K / V went to memory in pass A)
                for (int j = ctid; j < d; j += PG_CTHREADS) s_vn[j] = pg_ld_tag1(gv, (size_t)h * d + j, tag_v, p.err);
            pg_bar();
            if (warp == 0) {
                float M = -INFINITY;
                for (int cc = lane; cc < nch; cc += 32) M = fmaxf(M, __ldcg(g_mx + (size_t)h * op.max_chunks + cc));
#pragma unroll
                for (int of = 16; of > 0; of >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, of));
                const double l = chunk_exp(np, M);
                if (lane == 0) op.att_sum[(size_t)h * op.max_chunks + c] = l;
            }
            pg_bar();
            pg_att_pv(mv, s_vn, s_sc, s_o, p_lo, np, n_past, E, h * d, d, ctid);
            if (ctid < d) op.att_part[((size_t)h * op.max_chunks + c) * d + ctid] = s_o[ctid];
            __threadfence();
            pg_bar();
            if (ctid == 0) {
                const int last = atomicAdd(cnt2 + h, 1) == nch - 1;
                if (last) { cnt1[h] = 0; cnt2[h] = 0; }      // every chunk of the head is past both counters
                *s_flag = last;
            }
            pg_bar();
            if (*s_flag) {
                __threadfence();
                if (ctid < d) {
                    double S = 0.0;
                    for (int cc = 0; cc < nch; cc++) S = __dadd_rn(S, __ldcg(op.att_sum + (size_t)h * op.max_chunks + cc));
                    const float inv = (float)(1.0 / S);
                    float O = 0.f;
                    for (int cc = 0; cc < nch; cc++) O = __fadd_rn(O, __ldcg(op.att_part + ((size_t)h * op.max_chunks + cc) * d + ctid));
                    att_store((size_t)(op.head0 + h) * d + ctid, __fmul_rn(O, inv));
                }
            }
        }
    }
    if (op.publish) {
        __threadfence();
        pg_bar();
        if (ctid == 0) atomicAdd(p.ctr + 1 + o, 1);
    } else {
        pg_bar();
        if (ctid == 0) pg_add_hint_all(p, o);                // (no fence: only tells pollers that looking at the tags is worth it)
    }
}

// y[n] = sum of the K-slice partials of a sliced mul_mat (only where a caller wants the finished vector in memory;
// inside a program the next op adds the slices itself while it builds its activation pack)
__device__ __noinline__ void pg_sum(const PgParams *pp, int o, int qt)
{
    const PgParams &p = *pp;
    const PgSmem sm = pg_smem(p, qt);
    const int tagbase = sm.flag[6];
    const PgOp &op = p.ops[o];
    const int ctid = threadIdx.x - PG_C0;
    const int n4 = op.N / 4;
    for (int i = (int)blockIdx.x * PG_CTHREADS + ctid; i < n4; i += (int)gridDim.x * PG_CTHREADS)
        reinterpret_cast<float4 *>(op.y)[i] = pg_src4(op.src[0], i * 4, tagbase, p.err);
    for (int n = n4 * 4 + (int)blockIdx.x * PG_CTHREADS + ctid; n < op.N; n += (int)gridDim.x * PG_CTHREADS) {   // N % 4 tail
        const PgSrc &sr = op.src[0];
        float a = 0.f;
        for (int q = 0; q < sr.np; q++) {
            const size_t idx = (size_t)q * sr.stride + n;
            a += (sr.kind & 255) == PG_SRC_TAG ? pg_ld_tag1(static_cast<const float *>(sr.ptr), idx, tagbase | ((sr.kind >> 8) + 1), p.err)
                                               : __ldcg(static_cast<const float *>(sr.ptr) + idx);
        }
        op.y[n] = a * sr.scale;
    }
    if (op.publish) {
        __threadfence();
        pg_bar();
        if (ctid == 0) atomicAdd(p.ctr + 1 + o, 1);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The activation pack of a mul_mat op, built by the PG_PTHREADS prep threads into pack buffer b: dependency wait (tagged
// inputs are polled, a counted dependency is waited for), residual update / layer-norm / gelu, quantization.
// ---------------------------------------------------------------------------------------------------------------
// GROUP 0: the PG_PTHREADS prep threads (ahead of the consumers); GROUP 1: the PG_CTHREADS consumer threads themselves.
// waited_mv (the latest mul_mat op this CTA knows to be complete) lives in shared memory: sm.flag[7] / sm.flag[1].
template <int QT, int GROUP>
__device__ __noinline__ void pg_build_pack(const PgParams *pp, const PgOp *opp, int o, int b)
{
    constexpr int NP = (QT == 0) ? 2 : 3;
    constexpr int NT = GROUP == 0 ? PG_PTHREADS : PG_CTHREADS, NW = NT / 32, WFLAG = GROUP == 0 ? 7 : 1;
    auto gbar = [] { if (GROUP == 0) pg_pbar(); else pg_bar(); };
    const PgParams &p = *pp;                                 // in shared memory: nothing here may live on the stack
    const PgOp &op = *opp;
    const PgSmem sm = pg_smem(p, QT);
    float *s_x = sm.x;
    uint8_t *s_aq = sm.aq + (size_t)b * sm.pack_stride;
    float *s_ap = reinterpret_cast<float *>(s_aq + (size_t)p.nbp_cap * 32);
    double *s_stat = sm.stat;
    constexpr bool has_data = true;                          // (the prep warps build every pack: they cannot know the CTA's tasks ahead of time)
    const int ctid = GROUP == 0 ? (int)threadIdx.x : (int)threadIdx.x - PG_C0, lane = ctid & 31, warp = ctid >> 5;
    const bool sliced = op.PP > 1;
    const int cls = sliced ? (int)blockIdx.x % op.PP : 0;
    const int cbase = cls * op.TS, blk0 = cbase * CHUNK_BLKS;
    const int nbp = sliced ? op.TS * CHUNK_BLKS : ((op.C * CHUNK_BLKS + 7) & ~7);
    const int nb_l = min(op.nb - blk0, nbp);
    const int plane = nbp;
    const int in_mode = op.in_mode, res = op.res, nsrc = op.nsrc, K = op.K;
    // A CTA without a task of this op needs no activation pack; only an update of the residual stream has to happen
    // everywhere (every CTA carries its own copy).
    if (GROUP == 1) gbar();                                  // every consumer warp is done with this buffer's previous pack
    const int waited_mv = sm.flag[WFLAG];
    const int tagbase = sm.flag[6];
    int *err = p.err;
    if (has_data) {
        // records of padded blocks (weight scale 0, but 0 x garbage could be NaN) read as zero
        for (int i = nb_l * 4 + ctid; i < nbp * 4; i += NT) reinterpret_cast<uint2 *>(s_aq)[i] = make_uint2(0u, 0u);
        for (int pl = 0; pl < NP; pl++)
            for (int i = nb_l + ctid; i < nbp; i += NT) s_ap[pl * plane + i] = 0.0f;
    }
    const int dep = op.dep;
    if (dep >= 0 && dep > waited_mv) {
        if (ctid == 0) pg_wait_ctr(p.ctr + 1 + dep, (int)gridDim.x * sm.flag[2], p.err, o + 1);
        gbar();                                            // (everybody has read flag[1] by now)
        if (ctid == 0 && op.dep_mv) sm.flag[WFLAG] = dep;        // a complete mul_mat op implies every earlier op is complete
    }
    if (ctid == 0) {                                         // tagged inputs: one thread waits until they are probably all there
        for (int j = 0; j < nsrc; j++)
            if ((op.src[j].kind & 255) == PG_SRC_TAG) pg_wait_hint(p.ctr + 1 + (op.src[j].kind >> 8), (int)gridDim.x * p.world * sm.flag[2]);
        PG_STAMP(o, 2);
    }
    gbar();
    auto block_sum = [&](double v) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, of));
        if (lane == 0) s_stat[warp] = v;
        gbar();
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < NW; i++) t = __dadd_rn(t, s_stat[i]);
        gbar();
        return t;
    };
    const int items = nb_l * 4;                              // one item = 8 consecutive elements of one quant block
    if (in_mode == PG_IN_LN) {
        float4 *sx4 = reinterpret_cast<float4 *>(s_x);
        // the value layer-norm sees at float4 index k4
        auto ln_in = [&](int k4) {
            float4 a = sx4[k4];
            if (res == PG_RES_TEMP) {
                const float4 t = pg_sum_src4(op, nsrc, k4 * 4, tagbase, err);
                a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
            }
            return a;
        };
        if (res != PG_RES_KEEP) {
            // ggml_norm (ggml.c:4282-4298): mean and variance accumulated in double (tree order here), two passes
            double ls = 0.0;
            if (res == PG_RES_TEMP) {
                for (int k4 = ctid; k4 < K / 4; k4 += NT) {
                    const float4 a = ln_in(k4);
                    ls = __dadd_rn(ls, __dadd_rn(__dadd_rn((double)a.x, (double)a.y), __dadd_rn((double)a.z, (double)a.w)));
                }
            } else {
                for (int kb = ctid; kb < K / 4; kb += 2 * NT) {     // this thread's quads two at a time: their loads travel together
                    const bool second = kb + NT < K / 4;
                    float4 av[2];
                    pg_sum_src4x2(op, nsrc, kb * 4, (kb + NT) * 4, second, tagbase, err, av[0], av[1]);
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        if (u == 1 && !second) break;
                        const int k4 = kb + u * NT;
                        float4 a = av[u];
                        if (res == PG_RES_ADD) {
                            const float4 x = sx4[k4];
                            a.x = __fadd_rn(x.x, a.x); a.y = __fadd_rn(x.y, a.y); a.z = __fadd_rn(x.z, a.z); a.w = __fadd_rn(x.w, a.w);
                        }
                        sx4[k4] = a;
                        ls = __dadd_rn(ls, __dadd_rn(__dadd_rn((double)a.x, (double)a.y), __dadd_rn((double)a.z, (double)a.w)));
                    }
                }
            }
            const double mean = block_sum(ls) / (double)K;
            double l2 = 0.0;
            for (int k4 = ctid; k4 < K / 4; k4 += NT) {
                const float4 a = ln_in(k4);
                const double v0 = __dsub_rn((double)a.x, mean), v1 = __dsub_rn((double)a.y, mean);
                const double v2 = __dsub_rn((double)a.z, mean), v3 = __dsub_rn((double)a.w, mean);
                l2 = __dadd_rn(l2, __dadd_rn(__dadd_rn(__dmul_rn(v0, v0), __dmul_rn(v1, v1)), __dadd_rn(__dmul_rn(v2, v2), __dmul_rn(v3, v3))));
            }
            const double var = block_sum(l2) / (double)K;
            if (ctid == 0) {
                s_stat[32] = mean;
                s_stat[33] = (double)(float)(1.0 / sqrt(__dadd_rn(var, (double)1e-5f)));
            }
            gbar();
        }
        const double mean = s_stat[32];
        const float rstd = (float)s_stat[33];
        const float *lw = op.ln_w, *lb = op.ln_b;
        for (int base = 0; has_data && base < items; base += NT) {
            const int id = base + ctid;
            const bool valid = id < items;
            const int blk = valid ? id >> 2 : 0;
            const int k = (blk0 + blk) * QK + (lane & 3) * 8;
            float v[8];
            if (valid) {
                const float4 a0 = ln_in(k / 4), a1 = ln_in(k / 4 + 1);
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(lw + k)), w1 = __ldg(reinterpret_cast<const float4 *>(lw + k + 4));
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(lb + k)), b1 = __ldg(reinterpret_cast<const float4 *>(lb + k + 4));
                const float xs[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                // y = w * ((float)(x - mean) * scale) + b, every operation rounded on its own (ggml.c:4290-4301, vsim.cpp:523-533)
#pragma unroll
                for (int i = 0; i < 8; i++)
                    v[i] = __fadd_rn(__fmul_rn(ws[i], __fmul_rn((float)__dsub_rn((double)xs[i], mean), rstd)), bs[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = 0.f;
            }
            quantize8_emit<QT>(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), valid, lane,
                               s_aq + (size_t)blk * 32, s_ap + blk, plane);
        }
    } else if (GROUP == 0 && in_mode == PG_IN_PLAIN && nsrc == 1 && op.src[0].np == 1 && op.src[0].scale == 1.0f &&
               (op.src[0].kind & 255) != PG_SRC_Q4ROW) {
        // The prep warps' common case (one plain or tagged vector).  64 threads cover the slice with several items each:
        // the loads of a whole batch are issued before anything is looked at, so that a batch costs one memory round trip.
        constexpr int B = 4;                                 // items per thread and batch
        const PgSrc &sr = op.src[0];
        const bool tagged = (sr.kind & 255) == PG_SRC_TAG;
        const int tag = tagbase | ((sr.kind >> 8) + 1);
        const float *sp = static_cast<const float *>(sr.ptr);
        for (int base = 0; base < items; base += NT * B) {
            float4 r[B][4];                                  // tagged: 4 x {v, t, v, t};  plain: 2 x float4
            int kk[B];
#pragma unroll
            for (int i = 0; i < B; i++) {
                const int id = base + i * NT + ctid;
                const int blk = id < items ? id >> 2 : 0;
                kk[i] = (blk0 + blk) * QK + (lane & 3) * 8;
                if (id < items) {
                    if (tagged) {
#pragma unroll
                        for (int j = 0; j < 4; j++) r[i][j] = ld_relaxed_f4(sp + 2 * (size_t)kk[i] + 4 * j);
                    } else {
                        r[i][0] = __ldcg(reinterpret_cast<const float4 *>(sp + kk[i]));
                        r[i][1] = __ldcg(reinterpret_cast<const float4 *>(sp + kk[i] + 4));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < B; i++) {
                const int id = base + i * NT + ctid;
                if (base + i * NT >= items) break;           // (warp-uniform: whole warps stop together)
                const bool valid = id < items;
                const int blk = valid ? id >> 2 : 0;
                float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
                if (valid) {
                    if (tagged) {
                        bool ok = true;
#pragma unroll
                        for (int j = 0; j < 4; j++) ok = ok && __float_as_int(r[i][j].y) == tag && __float_as_int(r[i][j].w) == tag;
                        if (ok) {
                            lo = make_float4(r[i][0].x, r[i][0].z, r[i][1].x, r[i][1].z);
                            hi = make_float4(r[i][2].x, r[i][2].z, r[i][3].x, r[i][3].z);
                        } else {                             // not there yet after all: poll
                            lo = pg_ld_tag4(sp, (size_t)kk[i], tag, err);
                            hi = pg_ld_tag4(sp, (size_t)kk[i] + 4, tag, err);
                        }
                    } else { lo = r[i][0]; hi = r[i][1]; }
                }
                quantize8_emit<QT>(lo, hi, valid, lane, s_aq + (size_t)blk * 32, s_ap + blk, plane);
            }
        }
    } else {
        const uint16_t *tab = op.gelu_tab;
        for (int base = 0; base < items; base += NT) {
            const int id = base + ctid;
            const bool valid = id < items;
            const int blk = valid ? id >> 2 : 0;
            const int k = (blk0 + blk) * QK + (lane & 3) * 8;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            if (valid) {
                pg_sum_src4x2(op, nsrc, k, k + 4, true, tagbase, err, lo, hi);
                if (in_mode == PG_IN_GELU) {
                    // ggml_gelu (ggml.c:819-827, GGML_GELU_FP16): x -> fp16 -> table -> fp32
                    auto gl = [&](float x) { return __half2float(__ushort_as_half(__ldg(tab + __half_as_ushort(__float2half_rn(x))))); };
                    lo.x = gl(lo.x); lo.y = gl(lo.y); lo.z = gl(lo.z); lo.w = gl(lo.w);
                    hi.x = gl(hi.x); hi.y = gl(hi.y); hi.z = gl(hi.z); hi.w = gl(hi.w);
                }
            }
            quantize8_emit<QT>(lo, hi, valid, lane, s_aq + (size_t)blk * 32, s_ap + blk, plane);
        }
    }
    if (ctid == 0) { sm.geo[b * 4] = op.C; sm.geo[b * 4 + 1] = cbase; sm.geo[b * 4 + 2] = plane; if (GROUP == 1) sm.flag[5] = op.pack_seq; }
    if (ctid == 0) PG_STAMP(o, 3);
    gbar();                                               // every prep thread's part of the pack is written ...
    if (ctid == 0) mbar_arrive(&sm.pack_full[b]);            // ... before the consumers are told
}

// ===================== prep warps =====================
// Walk the mul_mat ops in list order, one pack ahead of the consumers: pack number q goes to buffer q & 1 as soon as
// every consumer warp has released that buffer (pack q - 2).  Packs the consumers build themselves are skipped.
template <int QT>
__device__ __noinline__ void pg_prep(const PgParams *pp)
{
    const PgParams &p = *pp;
    const PgSmem sm = pg_smem(p, QT);
    const int ptid = threadIdx.x;
    PgOp *my = reinterpret_cast<PgOp *>(sm.opc + (size_t)PG_OPC * PG_OPC_WORDS);      // the prep warps' own copy of the op they work on
    for (int o = p.first_mv; o < p.n_ops;) {
        pg_pbar();                                           // (the previous pack's readers of *my are done)
        if (ptid < PG_OPC_WORDS / 4) reinterpret_cast<uint4 *>(my)[ptid] = __ldg(reinterpret_cast<const uint4 *>(p.ops + o) + ptid);
        pg_pbar();
        const int nxt = my->next_mv;
        if (my->pack_of == o && my->builder == 0) {          // a new pack of ours (ops that reuse the previous one need nothing)
            const int q = my->pack_seq, b = q & 1, wq = my->wait_pack;
            if (ptid == 0) {
                pg_mbar_wait(&sm.pack_empty[b], (uint32_t)((q >> 1) & 1) ^ 1u);
                // a layer-norm pack reads the residual stream the consumers updated while they built pack wq
                if (wq >= 0) pg_mbar_wait(&sm.pack_full[wq & 1], (uint32_t)(wq >> 1) & 1u);
            }
            pg_pbar();
            pg_build_pack<QT, 0>(pp, my, o, b);
        }
        o = nxt;
    }
}

constexpr int PG_END = 1 << 30;                          // flag in a stage's op word: the CTA's last stage of that op
constexpr int PG_FAST = 1 << 29;                         // ... 32 units of one tile from an even chunk on: every warp has one whole, aligned octet
constexpr int PG_OPMASK = PG_FAST - 1;
// Consumers -> reducer hand-off of ring slot s: named barrier PG_BAR_DONE + s.  The 512 consumer threads arrive
// (bar.arrive, non-blocking), the reducer warp syncs on it: 16 arrivals on one mbarrier cost every warp ~350 cycles
// of contention per stage (tools/octetbench.cu), the barrier unit takes them in a few.
constexpr int PG_BAR_DONE = 3;
__device__ __forceinline__ void pg_done_arrive(int s) { asm volatile("bar.arrive %0, %1;" :: "r"(PG_BAR_DONE + s), "n"(PG_CTHREADS + 32) : "memory"); }
__device__ __forceinline__ void pg_done_sync(int s) { asm volatile("bar.sync %0, %1;" :: "r"(PG_BAR_DONE + s), "n"(PG_CTHREADS + 32) : "memory"); }

__device__ __forceinline__ const PgOp &pg_cached_op(const PgSmem &sm, int o)
{
    return *reinterpret_cast<const PgOp *>(sm.opc + (size_t)(o & (PG_OPC - 1)) * PG_OPC_WORDS);
}

// Each role is an out-of-line function that derives its pointers from the parameter copy in shared memory: every role
// gets its own register allocation, and nothing a role does not use stays live in another role's loop.

// ===================== producer =====================
// All 32 lanes keep the op cache filled (lane l < 21 carries 16 bytes of the NEXT mul_mat op while lane 0 issues the
// copies of the current one); lane 0 alone talks to the stage ring.
template <int QT>
__device__ __noinline__ void pg_producer(const PgParams *pp)
{
    constexpr int CH = chunk_bytes(QT);
    constexpr int STAGE_BYTES = PG_UNITS * CH;
    const PgParams &p = *pp;
    const PgSmem sm = pg_smem(p, QT);
    const int lane = threadIdx.x & 31;
    const int bid = (int)blockIdx.x, grid = (int)gridDim.x;
    int s = 0;
    uint32_t ph = 1;
    const uint64_t wpol = l2_policy_evict_first();
    auto put = [&](int cnt, int c0, int T, int task, int otag, const uint8_t *src) {
        const long long tw = PG_CLK();
        pg_mbar_wait(&sm.bar_empty[s], ph);
        PG_ACC(pp, 0, PG_CLK() - tw);
        sm.desc[s] = make_int4(cnt, c0, T, task);
        sm.dop[s] = otag;
#ifdef B200Q4_TIMELINE
        reinterpret_cast<long long *>(sm.stat + 8)[s] = clock64();   // (scratch doubles 8..13 are free while the consumers run stages)
#endif
        if (cnt > 0) {
            mbar_arrive_expect_tx(&sm.bar_full[s], (uint32_t)cnt * CH);
            if (p.wt_evict_first) bulk_g2s_hint(sm.stage + (size_t)s * STAGE_BYTES, src, (uint32_t)cnt * CH, &sm.bar_full[s], wpol);
            else bulk_g2s(sm.stage + (size_t)s * STAGE_BYTES, src, (uint32_t)cnt * CH, &sm.bar_full[s]);
        } else {
            mbar_arrive(&sm.bar_full[s]);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
    };
    auto fetch = [&](int o) {
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (o < p.n_ops && lane < PG_OPC_WORDS / 4) w = __ldg(reinterpret_cast<const uint4 *>(p.ops + o) + lane);
        return w;
    };
    int o = p.first_mv;
    uint4 cw = fetch(o);
    int pre_ticket = 0;                                      // first ticket of op o, drawn while the previous op was issued
    if (lane == 0 && o < p.n_ops) {
        const PgOp &g0 = p.ops[o];
        if (g0.dynamic) pre_ticket = atomicAdd(p.ctr + g0.tk_off + (g0.PP > 1 ? bid % g0.PP : 0), 1);
    }
    while (o < p.n_ops) {
        if (lane < PG_OPC_WORDS / 4) reinterpret_cast<uint4 *>(sm.opc + (size_t)(o & (PG_OPC - 1)) * PG_OPC_WORDS)[lane] = cw;
        __syncwarp();
        const PgOp &op = pg_cached_op(sm, o);
        const int nxt_o = op.next_mv;
        cw = fetch(nxt_o);                                   // lands while this op's tasks are issued
        if (lane == 0) {
            bool hold_mode = false;                         // (the same decision in every CTA: it depends on n_past alone)
            int items_h = 0;                                // hold mode: the CTAs 0 .. items_h - 1 have attention work
            if (op.hold_a > 0) {
                const PgAttGeo ag = pg_att_geo((p.dyn ? __ldg(p.dyn + 1) : 0) + 1, op.hold_H, grid);
                items_h = op.hold_H * ag.nch;
                hold_mode = items_h * 2 <= grid;
                if (!hold_mode) items_h = 0;
                if (hold_mode && bid < items_h && *reinterpret_cast<volatile int *>(sm.flag + 3) < op.hold_a) {
                    put(0, 0, 0, 0, o, nullptr);             // a bare entry of this op: what makes the consumers run the attention before it
                    const long long t0 = clock64();
                    while (*reinterpret_cast<volatile int *>(sm.flag + 3) < op.hold_a) {
                        __nanosleep(200);
                        if (clock64() - t0 > 8000000000ll) __trap();
                    }
                }
                if (op.dynamic) pre_ticket = atomicAdd(p.ctr + op.tk_off + (op.PP > 1 ? bid % op.PP : 0), 1);   // (none was drawn ahead)
            }
            PG_STAMP(o, 0);
            int nx_ticket = 0;
            if (nxt_o < p.n_ops && op.nx_dynamic) nx_ticket = atomicAdd(p.ctr + op.nx_tk_off + (op.nx_PP > 1 ? bid % op.nx_PP : 0), 1);
            const uint8_t *wt = op.wt;
            const int C = op.C, PP = op.PP, TS = op.TS, U = op.U;
            const bool sliced = PP > 1;
            const int cls = sliced ? bid % PP : 0, cbase = cls * TS;
            // hold mode: the static tasks are dealt to the CTAs without attention work only (hcls CTAs of this class have some:
            // the lowest block indices); whatever is left comes from the ticket counter, for everybody
            const int hcls = sliced ? (items_h > cls ? (items_h - cls + PP - 1) / PP : 0) : items_h;
            const int ncls = (sliced ? (grid - cls + PP - 1) / PP : grid) - hcls;
            const int nstatic = (hold_mode && bid < items_h) ? 0 : op.nstatic;
            int idx = (sliced ? bid / PP : bid) - hcls + (hold_mode ? 0 : op.rot);
            if (idx >= ncls) idx -= (idx / ncls) * ncls;
            const int limit = sliced ? op.n_tiles : op.ntasks;
            // tasks are issued one behind, so that the CTA's last stage of the op can carry the END flag
            int pend = -1;
            auto issue = [&](int u0, int n, int T, int c0, int task, bool last_task) {
                for (int r = 0; r < n; r += PG_UNITS) {
                    const int cnt = min(PG_UNITS, n - r);
                    const int fast = (cnt == PG_UNITS && !(c0 & 1) && c0 + cnt <= C) ? PG_FAST : 0;
                    put(cnt, c0, T, task, ((last_task && r + PG_UNITS >= n) ? (o | PG_END) : o) | fast, wt + (size_t)(u0 + r) * CH);
                    c0 += cnt;
                    while (c0 >= C) { c0 -= C; ++T; }
                }
            };
            auto issue_seq = [&](int t, bool last_task) {
                if (sliced) issue(t * C + cbase, min(TS, C - cbase), t, cbase, t * PP + cls, last_task);
                else if (C >= PIPE_TILE_SPLIT) issue(t * C, C, t, 0, t, last_task);               // one whole tile
                else issue(t * TS, min(TS, U - t * TS), t * (TS / C), 0, t, last_task);          // short tiles: TS / C whole tiles
            };
            auto push = [&](int t) { if (pend >= 0) issue_seq(pend, false); pend = t; };
            for (int k = 0; k < nstatic; k++) {
                const int t = idx + k * ncls;
                if (t < limit) push(t);
            }
            if (op.dynamic) {
                int *tk = p.ctr + op.tk_off + cls;
                const int base = op.nstatic * ncls;
                int cur = base + pre_ticket;
                while (cur < limit) {
                    const int nxt = base + atomicAdd(tk, 1);   // one task ahead: the atomic's latency hides behind the copies
                    push(cur);
                    cur = nxt;
                }
            }
            if (pend >= 0) issue_seq(pend, true);
            else put(0, 0, 0, 0, o | PG_END, nullptr);       // no task of this op for this CTA: a bare marker
            pre_ticket = nx_ticket;
        }
        __syncwarp();
        o = nxt_o;
    }
    if (lane == 0) put(-2, 0, 0, 0, p.n_ops, nullptr);
}

// ===================== reducer =====================
template <int QT>
__device__ __noinline__ void pg_reducer(const PgParams *pp)
{
    const PgParams &p = *pp;
    const PgSmem sm = pg_smem(p, QT);
    const int lane = threadIdx.x & 31;
    constexpr float post = (QT == 0) ? 0.0625f : 1.0f;       // undo the 16x of the A/B code scaling (exact)
    float acc = 0.f, bcur = 0.f;
    int curT = -1, curTask = -1, curOp = -1;
    float *y = nullptr, *y_plain = nullptr;
    const float *bias = nullptr;
    int N = 0, C = 1, PP = 1, y_slice = 0, ytag = 0, publish = 0, row0 = 0;
    const int world = p.world;
    const PgOp *cop = nullptr;
    const int rl = lane & 15;
    const int tagbase = sm.flag[6];
    auto emit = [&]() {
        if (curT < 0) return;
        const int n = curT * TILE_ROWS + lane;
        if (lane < TILE_ROWS && n < N) {
            const float v = (PP == 1 && bias) ? __fadd_rn(acc * post, bcur) : acc * post;
            const size_t idx = (PP == 1 ? (size_t)n : (size_t)(curTask - curT * PP) * y_slice + n) + row0;
            if (ytag && world > 1 && cop->y_peer[0]) {       // this rank's rows of the vector, on every rank
                for (int r = 0; r < world; r++) st_tagged(cop->y_peer[r], idx, v, ytag);
            } else if (ytag) st_tagged(y, idx, v, ytag);     // value and tag in one 8-byte store: the data is its own flag
            else y[idx] = v;
            if (y_plain) y_plain[n] = v;
        }
        curT = -1;
    };
    int s = 0;
    uint32_t ph = 0;
    const long long tr0 = PG_CLK();
    for (;;) {
        const long long tw = PG_CLK();
        pg_done_sync(s);                                     // every consumer warp has passed this slot (markers included)
        if (lane == 0) PG_ACC(pp, 3, PG_CLK() - tw);
        const int4 dsc = sm.desc[s];
        const int otag = sm.dop[s], o = otag & PG_OPMASK;
        if (dsc.x == -2) break;
        if (o != curOp) {
            const PgOp &op = pg_cached_op(sm, o);
            y = op.y; y_plain = op.y_plain; bias = op.bias; N = op.N; C = op.C; PP = op.PP; y_slice = op.y_slice; row0 = op.row0; cop = &op;
            ytag = op.y_tag ? (tagbase | (o + 1)) : 0; publish = op.publish;
            curOp = o;
        }
        if (dsc.x > 0) {
            const int task = dsc.w;
            int T = dsc.z;
            int left = C - dsc.y;
            const float *red = sm.red + (size_t)s * PG_UNITS * PG_RW + rl;
            int j = 0;
            while (j < dsc.x) {
                if (T != curT || task != curTask) {
                    emit();
                    acc = 0.f;
                    curT = T; curTask = task;
                    if (bias && PP == 1) {                   // asked for now, needed when the tile is complete
                        const int n = T * TILE_ROWS + lane;
                        bcur = (lane < TILE_ROWS && n < N) ? __ldg(bias + n) : 0.f;
                    }
                }
                const int run = min(dsc.x - j, left);
                if (otag & PG_FAST) {                        // aligned octets: the odd slots would all be zero (run == PG_UNITS)
                    float v[PG_UNITS / 2];
#pragma unroll
                    for (int k = 0; k < PG_UNITS / 2; k++) v[k] = red[2 * k * PG_RW];
                    float q[PG_UNITS / 4];
#pragma unroll
                    for (int h = 0; h < PG_UNITS / 4; h++) q[h] = v[2 * h] + v[2 * h + 1];
                    acc += ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
                } else if (run == PG_UNITS) {
                    float v[PG_UNITS];
#pragma unroll
                    for (int k = 0; k < PG_UNITS; k++) v[k] = red[k * PG_RW];
                    float q[PG_UNITS / 4];
#pragma unroll
                    for (int h = 0; h < PG_UNITS / 4; h++) q[h] = (v[4 * h] + v[4 * h + 1]) + (v[4 * h + 2] + v[4 * h + 3]);
                    acc += ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
                } else {
                    int k = j;
                    for (; k + 4 <= j + run; k += 4)
                        acc += (red[k * PG_RW] + red[(k + 1) * PG_RW]) + (red[(k + 2) * PG_RW] + red[(k + 3) * PG_RW]);
                    for (; k < j + run; k++) acc += red[k * PG_RW];
                }
                j += run;
                left -= run;
                if (left == 0) { ++T; left = C; }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_empty[s]);        // the slot goes back to the producer before anything slow happens
        if (otag & PG_END) {                                 // this CTA is done with op o: flush, and publish if somebody counts
            emit();
            if (publish) {
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(p.ctr + 1 + o, 1);
            } else if (ytag) {
                __syncwarp();
                if (lane == 0) pg_add_hint_all(p, o);        // (no fence: only tells pollers that looking at the tags is worth it)
            }
            if (lane == 0) PG_STAMP(o, 5);
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    if (lane == 0) PG_ACC(pp, 4, PG_CLK() - tr0);
    // the last CTA to get here clears every counter for the next launch (nobody polls or draws any more:
    // each CTA's consumers and producer are past their last use when its reducer sees STOP)
    __syncwarp();
    int last = 0;
    if (lane == 0) { __threadfence(); last = atomicAdd(p.ctr, 1) == (int)gridDim.x - 1; }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence();
        for (int i = 1 + p.n_ops + lane; i < p.n_ctr; i += 32) p.ctr[i] = 0;     // tickets, attention head counters
        if (lane == 0) { p.ctr[0] = 0; p.dyn[2] = p.dyn[2] + 1; }                 // next launch, next epoch: this launch's tags become stale
    }
}

// ===================== consumers =====================
// one stage: this warp's two units (an octet of quant blocks) against the activation pack
template <int QT>
__device__ __forceinline__ void pg_consume_stage(const PgSmem &sm, const uint8_t *s_aq, const float *s_ap, int gC, int gcbase, int gplane,
                                                 int s, int cnt, int c0, int T0, bool fast)
{
    constexpr int CH = chunk_bytes(QT);
    constexpr int STAGE_BYTES = PG_UNITS * CH;
    const int lane = threadIdx.x & 31, warp = (threadIdx.x - PG_C0) >> 5;
    const int uA = 2 * warp, uB = uA + 1;
    if (fast) {                                              // units 2 warp, 2 warp + 1 = one aligned octet of tile T0
        const uint8_t *chA = sm.stage + (size_t)s * STAGE_BYTES + (size_t)uA * CH;
        float v0, v2;
        compute_octet<QT, false>(chA, chA + CH, (c0 + uA - gcbase) >> 1, s_aq, s_ap, gplane, lane, v0, v2);
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, 1); v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
        const int g = lane >> 2, t = lane & 3;
        if (t < 2) sm.red[((size_t)s * PG_UNITS + uA) * PG_RW + g + 8 * t] = t ? v2 : v0;
        return;
    }
    auto put_rows = [&](float *slot, float v0, float v2) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, 1); v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
        const int g = lane >> 2, t = lane & 3;
        if (t < 2) slot[g + 8 * t] = t ? v2 : v0;
    };
    const bool hA = uA < cnt, hB = uB < cnt;
    int TA = T0, TB = T0;
    int cA = c0 + uA, cB = c0 + uB;
    while (cA >= gC) { cA -= gC; ++TA; }
    while (cB >= gC) { cB -= gC; ++TB; }
    const uint8_t *chA = sm.stage + (size_t)s * STAGE_BYTES + (size_t)uA * CH, *chB = chA + CH;
    float *slotA = sm.red + ((size_t)s * PG_UNITS + uA) * PG_RW, *slotB = slotA + PG_RW;
    if (hA && hB && TA == TB && !(cA & 1)) {                 // an aligned octet: the common case
        float v0, v2;
        compute_octet<QT, false>(chA, chB, (cA - gcbase) >> 1, s_aq, s_ap, gplane, lane, v0, v2);
        put_rows(slotA, v0, v2);
        if (lane < TILE_ROWS) slotB[lane] = 0.f;
    } else {                                                 // half octets (odd start, tile boundary, ragged tail)
        if (hA) {
            float v0, v2;
            compute_octet<QT, false>((cA & 1) ? nullptr : chA, (cA & 1) ? chA : nullptr, (cA - gcbase) >> 1, s_aq, s_ap, gplane, lane, v0, v2);
            put_rows(slotA, v0, v2);
        }
        if (hB) {
            float v0, v2;
            compute_octet<QT, false>((cB & 1) ? nullptr : chB, (cB & 1) ? chB : nullptr, (cB - gcbase) >> 1, s_aq, s_ap, gplane, lane, v0, v2);
            put_rows(slotB, v0, v2);
        }
    }
}

// The stages of ONE op, from the warp's current slot on, up to and including the CTA's last stage of that op (END
// flag) or the STOP marker: a loop without calls, so that its state stays in registers.  The slot / parity live in
// shared memory between calls (wst[0], wst[1]): no register is live across a call in the driver below, so nothing goes
// through the stack (local memory has no L1 to speak of next to 224 KB of shared memory: a reload is an L2 round trip).
// Returns 1 after the STOP marker.
template <int QT>
__device__ __noinline__ int pg_consume_op(const PgParams *pp)
{
    const PgSmem sm = pg_smem(*pp, QT);
    const int stages = pp->stages;
    const int lane = threadIdx.x & 31;
    int *wst = sm.wst + ((threadIdx.x - PG_C0) >> 5) * 4;
    int s = wst[0];
    uint32_t ph = (uint32_t)wst[1];
    const int b = wst[3] & 1;                                // the op's pack buffer and its geometry, left there by the prep warps
    const int gC = sm.geo[b * 4], gcbase = sm.geo[b * 4 + 1], gplane = sm.geo[b * 4 + 2];
    const uint8_t *s_aq = sm.aq + (size_t)b * sm.pack_stride;
    const float *s_ap = reinterpret_cast<const float *>(s_aq + (size_t)pp->nbp_cap * 32);
    // A later op whose input is a plain vector nobody in the list writes (PgOp::early): this thread's eight elements of it
    // are requested now and quantized into the OTHER pack buffer after the first stage -- no pause, no barrier here.
    float4 elo = make_float4(0.f, 0.f, 0.f, 0.f), ehi = elo;
    int e_items = -1, e_blk0 = 0, e_nbp = 0, e_seen = 0;
    {
        const PgOp &cop = pg_cached_op(sm, wst[2]);
        if (wst[2] < pp->n_ops && cop.early > 0) {          // (the STOP marker has no op record)
            const int ctid = (int)threadIdx.x - PG_C0;
            const bool sliced = cop.e_PP > 1;
            const int cbase = sliced ? ((int)blockIdx.x % cop.e_PP) * cop.e_TS : 0;
            e_blk0 = cbase * CHUNK_BLKS;
            e_nbp = sliced ? cop.e_TS * CHUNK_BLKS : ((cop.e_C * CHUNK_BLKS + 7) & ~7);
            e_items = min(cop.e_nb - e_blk0, e_nbp) * 4;                          // <= PG_CTHREADS (the list builder checks)
            if (ctid < e_items) {
                const float *xp = cop.early_x + (size_t)(e_blk0 + (ctid >> 2)) * QK + (lane & 3) * 8;
                elo = __ldcg(reinterpret_cast<const float4 *>(xp));
                ehi = __ldcg(reinterpret_cast<const float4 *>(xp + 4));
            }
        }
    }
    for (;;) {
        const long long tw = PG_CLK();
        pg_mbar_wait(&sm.bar_full[s], ph);
        if (threadIdx.x == PG_C0) PG_ACC(pp, 1, PG_CLK() - tw);
#ifdef B200Q4_TIMELINE
        if (threadIdx.x == PG_C0 && PG_CLK() - tw > 200) { PG_ACC(pp, 8, PG_CLK() - reinterpret_cast<long long *>(sm.stat + 8)[s]); PG_ACC(pp, 9, 1); }
#endif
        const int4 dsc = sm.desc[s];
        const int otag = sm.dop[s];
        const long long tc = PG_CLK();
        if (dsc.x > 0) pg_consume_stage<QT>(sm, s_aq, s_ap, gC, gcbase, gplane, s, dsc.x, dsc.y, dsc.z, (otag & PG_FAST) != 0);
        if (threadIdx.x == PG_C0 && dsc.x > 0) { PG_ACC(pp, 5, PG_CLK() - tc); PG_ACC(pp, 6, 1); }
        pg_done_arrive(s);
        // (once per op: after the third ring entry -- the input vector comes from HBM and takes longer than one stage
        // to arrive -- or the op's last entry if that comes first)
        if (e_items >= 0 && (++e_seen >= 3 || (otag & PG_END) || dsc.x == -2)) {
            constexpr int NP = (QT == 0) ? 2 : 3;
            const int ctid = (int)threadIdx.x - PG_C0;
            uint8_t *d_aq = sm.aq + (size_t)(b ^ 1) * sm.pack_stride;
            float *d_ap = reinterpret_cast<float *>(d_aq + (size_t)pp->nbp_cap * 32);
            for (int i = e_items + ctid; i < e_nbp * 4; i += PG_CTHREADS) reinterpret_cast<uint2 *>(d_aq)[i] = make_uint2(0u, 0u);
            for (int pl = 0; pl < NP; pl++)
                for (int i = (e_items >> 2) + ctid; i < e_nbp; i += PG_CTHREADS) d_ap[pl * e_nbp + i] = 0.0f;
            if ((ctid & ~31) < e_items) {
                const bool valid = ctid < e_items;
                const int blk = valid ? ctid >> 2 : 0;
                quantize8_emit<QT>(elo, ehi, valid, lane, d_aq + (size_t)blk * 32, d_ap + blk, e_nbp);
            }
            if (lane == 0) {                                 // (every warp: the next op may start without a barrier; same values)
                const PgOp &cop = pg_cached_op(sm, wst[2]);
                sm.geo[(b ^ 1) * 4] = cop.e_C; sm.geo[(b ^ 1) * 4 + 1] = e_blk0 / CHUNK_BLKS; sm.geo[(b ^ 1) * 4 + 2] = e_nbp; sm.flag[5] = cop.e_seq;
            }
            __syncwarp();
            e_items = -1;
        }
        const bool stop = dsc.x == -2;
        if (!stop && ++s == stages) { s = 0; ph ^= 1; }
        if (stop || (otag & PG_END)) {
#ifdef B200Q4_TIMELINE
            if (threadIdx.x == PG_C0 && !stop) { const PgParams &p = *pp; PG_STAMP(otag & PG_OPMASK, 4); }
#endif
            if (lane == 0) { wst[0] = s; wst[1] = (int)ph; }
            __syncwarp();
            return stop ? 1 : 0;
        }
    }
}

// side ops between two mul_mats (attention, final sums), in list order
template <int QT>
__device__ __noinline__ void pg_side_ops(const PgParams *pp, int from, int to)
{
    const PgParams &p = *pp;
    const PgSmem sm = pg_smem(p, QT);
    for (int a = from; a < to; a++) {
        const int kind = p.ops[a].kind;
        if (kind == PG_MATVEC) continue;
        const int dep = p.ops[a].dep;
        pg_bar();
        const int wmv = sm.flag[1];
        if (dep >= 0 && dep > wmv) {
            if (threadIdx.x == PG_C0) pg_wait_ctr(p.ctr + 1 + dep, (int)gridDim.x * sm.flag[2], p.err, a + 1);
            pg_bar();                                        // (everybody has read flag[1] by now)
            if (threadIdx.x == PG_C0 && p.ops[dep].kind == PG_MATVEC) sm.flag[1] = dep;
        }
        if (threadIdx.x == PG_C0) PG_STAMP(a, 1);
        if (kind == PG_ATTN) pg_attn(pp, a, QT);
        else pg_sum(pp, a, QT);
        if (threadIdx.x == PG_C0) { PG_STAMP(a, 3); *reinterpret_cast<volatile int *>(sm.flag + 3) = a + 1; }     // (what a holding producer waits for)
    }
}

template <int QT>
__device__ __noinline__ void pg_consumer(const PgParams *pp)
{
    {
        const PgSmem sm = pg_smem(*pp, QT);
        int *wst = sm.wst + ((threadIdx.x - PG_C0) >> 5) * 4;
        if ((threadIdx.x & 31) == 0) { wst[0] = 0; wst[1] = 0; wst[2] = -1; wst[3] = -1; }
        __syncwarp();
    }
    for (;;) {
        // peek: which op does the next stage belong to?  (every value below is re-read from shared memory after a call)
        {
            const PgSmem sm = pg_smem(*pp, QT);
            int *wst = sm.wst + ((threadIdx.x - PG_C0) >> 5) * 4;
            const int s = wst[0];
            const long long tw = PG_CLK();
            pg_mbar_wait(&sm.bar_full[s], (uint32_t)wst[1]);
            if (threadIdx.x == PG_C0) PG_ACC(pp, 7, PG_CLK() - tw);
            const int o = sm.dop[s] & PG_OPMASK, cur = wst[2];
            if (o > cur + 1) pg_side_ops<QT>(pp, cur + 1, o);     // (STOP carries o = n_ops: trailing side ops run too)
        }
        {
            const PgSmem sm = pg_smem(*pp, QT);
            int *wst = sm.wst + ((threadIdx.x - PG_C0) >> 5) * 4;
            const int s = wst[0];
            const int o = sm.dop[s] & PG_OPMASK, cnt = sm.desc[s].x;
            const int lane = threadIdx.x & 31;
            __syncwarp();
            if (lane == 0) wst[2] = o;
            if (cnt != -2) {
                // the op's activation pack: wait until the prep warps have built it; hand the previous buffer back first
                const int q = pg_cached_op(sm, o).pack_seq, cq = wst[3];
                if (q != cq) {
                    { const PgParams &p = *pp; if (threadIdx.x == PG_C0) PG_STAMP(o, 1); (void)p; }
                    __syncwarp();
                    if (lane == 0) wst[3] = q;
                    const long long tw = PG_CLK();
                    // consumer-built packs: all 512 threads; flag[5] = the latest pack built so far (a pack may have been built
                    // early, in the pause of an earlier op).  Prep-built packs (PG_PWARPS > 0) are waited for.
                    if (pg_cached_op(sm, o).builder == 1) {
                        const int co = pg_cached_op(sm, o).cobuild;
                        PgOp *my = reinterpret_cast<PgOp *>(sm.opc + (size_t)PG_OPC * PG_OPC_WORDS);
                        if (co >= 0 && (int)threadIdx.x - PG_C0 < PG_OPC_WORDS / 4)           // its op record arrives while the first pack is built
                            reinterpret_cast<uint4 *>(my)[threadIdx.x - PG_C0] = __ldg(reinterpret_cast<const uint4 *>(pp->ops + co) + (threadIdx.x - PG_C0));
                        if (pg_cached_op(sm, o).prebuilt) {
                            if (pg_cached_op(sm, o).prebuilt == 1) pg_bar();                        // every warp has stored its share of the pack (pg_consume_op of the previous
                                                             // pack's first op), and nobody reads the other buffer any more
                        } else if (q > sm.flag[5]) {
                            pg_build_pack<QT, 1>(pp, &pg_cached_op(sm, o), o, q & 1);
                            if (co >= 0) pg_build_pack<QT, 1>(pp, my, co, (q + 1) & 1);
                        }
                    } else pg_mbar_wait(&sm.pack_full[q & 1], (uint32_t)(q >> 1) & 1u);
                    if (threadIdx.x == PG_C0) PG_ACC(pp, 2, PG_CLK() - tw);
                    __syncwarp();
                }
            }
        }
        if (pg_consume_op<QT>(pp)) break;
    }
}

template <int QT>
__global__ void __launch_bounds__(PG_THREADS, 1) q4_prog_kernel(const PgParams p)
{
    __shared__ PgParams s_p;                                 // what the roles read (a by-reference kernel parameter would be copied to the stack)
    const PgSmem sm = pg_smem(p, QT);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x < 3 * PG_MAX_STAGES + 4) {
        const int i = threadIdx.x % PG_MAX_STAGES, which = threadIdx.x / PG_MAX_STAGES;
        if (which < 3) {
            if (i < p.stages) {
                if (which == 0) mbar_init(&sm.bar_full[i], 1);
                else if (which == 1) mbar_init(&sm.bar_empty[i], 1);
                else mbar_init(&sm.bar_done[i], GEMV_WARPS);
            }
        } else if (i < 2) mbar_init(&sm.pack_full[i], 1);
        else if (i < 4) mbar_init(&sm.pack_empty[i - 2], GEMV_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x == PG_THREADS - 1) {
        s_p = p;
        sm.flag[1] = -1; sm.flag[5] = -1; sm.flag[7] = -1; sm.flag[3] = 0;
        sm.flag[6] = (int)(((unsigned)p.dyn[2] & 0x7ffffu) << 12);           // tag = epoch << 12 | (op index + 1)
        sm.flag[2] = p.dyn[2];                                               // the launch epoch itself: arrival / hint counters are
                                                                             // never reset, op o is complete at epoch * (CTAs that add)
    }
    __syncthreads();
    if (PG_PWARPS > 0 && warp < PG_PWARPS) pg_prep<QT>(&s_p);
    else if (warp < PG_PWARPS + GEMV_WARPS) pg_consumer<QT>(&s_p);
    else if (warp == PG_PWARPS + GEMV_WARPS) pg_producer<QT>(&s_p);
    else pg_reducer<QT>(&s_p);
}

#endif // Q4_PROG_TYPES_ONLY

} // namespace b200q4
