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
#include "q4_kernels.cuh"

namespace b200q4 {

constexpr int PG_MAX_SRC = 4;
constexpr int PG_MAX_STAGES = 6;
constexpr int PG_UNITS = 2 * GEMV_WARPS;                 // units per stage (32)
constexpr int PG_RW = 16;                                // floats per unit in the reduction ring (16 weight rows)
constexpr int PG_CTHREADS = GEMV_WARPS * 32;             // consumer threads (512)
constexpr int PG_THREADS = (GEMV_WARPS + 2) * 32;
constexpr int PG_ATT_CHUNK = 64;                         // KV positions per attention work item
constexpr int PG_ATT_MAX_D = 512;                        // largest head size the attention step handles

enum { PG_MATVEC = 0, PG_ATTN = 1, PG_SUM = 2 };
enum { PG_IN_PLAIN = 0, PG_IN_LN = 1, PG_IN_GELU = 2 };
enum { PG_SRC_F32 = 0, PG_SRC_Q4ROW = 1 };
enum { PG_RES_KEEP = 0, PG_RES_ADD = 1, PG_RES_SET = 2, PG_RES_TEMP = 3 };

// One input vector:  v[k] = scale * sum_{p < np} ptr[p * stride + k]        (PG_SRC_F32; np > 1: K-slice partials)
//                    v[k] = dequantize_row(ptr, row *row)[k]                (PG_SRC_Q4ROW: ggml_get_rows on the Q4
//                           embedding matrix, ggml.c:300-362; np = quant type, stride = row length K)
struct PgSrc {
    const void *ptr;
    const int  *row;
    int   np, stride;
    float scale;
    int   kind;
};

struct PgOp {
    int kind;                // PG_MATVEC / PG_ATTN / PG_SUM (y[n] = src[0] summed over its K-slice partials, n < N; spread over all CTAs)
    int dep;                 // index of the op that must be complete before this op's inputs are read (-1: none)
    // ---- PG_MATVEC:  Y[n] = sum_k W[n][k] * quantize(a)[k] (+ bias[n]) ---------------------------------------
    const uint8_t *wt;       // wtile
    int C, n_tiles, N, nb, U;
    int PP, TS, ntasks;      // K-slicing geometry (slice_geometry of q4_api.cu)
    int nstatic;             // tasks per CTA taken without a ticket
    int tk_off;              // first ticket counter of this op in PgParams::ctr (one per K-slice class)
    int in_mode;             // a = v (PLAIN) | gelu(v) (GELU) | layernorm(residual)(LN), v = sum of the sources
    int res;                 // LN: KEEP = residual and its statistics as they are; ADD = residual += v first;
                             //     SET = residual = v; TEMP = normalise (residual + v) without storing it
    int nsrc, K;
    PgSrc src[PG_MAX_SRC];
    const float *ln_w, *ln_b;
    const uint16_t *gelu_tab;
    float *y;                // PP == 1: Y;  PP > 1: slice p of the raw sums at y[p * y_slice + n]
    const float *bias;       // PP == 1 only (may be NULL)
    int y_slice, pad0;
    // ---- PG_ATTN: the new token's attention for one layer (decode, one query row) -----------------------------
    const float *q, *k, *v;          // [H * d] (biases already added)
    float *mem_k, *mem_v;            // this layer's KV memory [n_ctx][H * d], vsim.cpp:318-330
    float *att_out;                  // [H * d] merged heads (vsim.cpp:621-627)
    float *att_part;                 // [H][max_chunks][d + 2] partial (sum, max, denominator) per position chunk
    const double2 *rope;             // [n_ctx][n_rot / 2] {cos, sin}(p * theta_i), host libm like ggml.c:6126-6129
    int H, d, n_rot, max_chunks, cnt_off, pad1;
    float att_scale; int pad2;
};

struct PgParams {
    const PgOp *ops;
    int  n_ops;
    int *ctr;                // [0] CTAs finished; [1 + o] CTAs done with op o; tickets; attention head counters.  All zero
    int  n_ctr;              //   between launches: the last CTA to leave clears them.
    const int *dyn;          // [0] token id (through PgSrc::row), [1] n_past
    int *err;                // set (op index + 1) before the kernel traps on a wait that can never end
    int  stages, e_cap, nbp_cap, att_cap;    // shared-memory sizing: residual floats, apack quant blocks, attention floats
};

__host__ __device__ constexpr size_t pg_stage_bytes(int qt) { return (size_t)PG_UNITS * chunk_bytes(qt) + (size_t)PG_UNITS * PG_RW * 4; }
__host__ __device__ constexpr size_t pg_fixed_bytes(int e_cap, int nbp_cap, int att_cap)
{
    return (size_t)e_cap * 4 + (size_t)nbp_cap * 32 + (size_t)nbp_cap * 12 + (size_t)att_cap * 4 + 40 * 8 +
           3 * PG_MAX_STAGES * 8 + PG_MAX_STAGES * 16 + 64 + 128;
}

#ifndef Q4_PROG_TYPES_ONLY                               // (host-only includers stop here)

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
        if (spins > 64) __nanosleep(40);
        if ((spins & 1023) == 1023 && clock64() - t0 > 4000000000ll) {
            if (err) *err = code;
            __threadfence_system();
            __trap();
        }
    }
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

// four consecutive elements k .. k+3 of one source (k % 4 == 0)
__device__ __forceinline__ float4 pg_src4(const PgSrc &s, int k)
{
    if (s.kind == PG_SRC_F32) {
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
__device__ __forceinline__ float4 pg_sum_src4(const PgOp &op, int nsrc, int k)
{
    float4 a = pg_src4(op.src[0], k);
    for (int j = 1; j < nsrc; j++) {
        const float4 t = pg_src4(op.src[j], k);
        a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
    }
    return a;
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

__device__ __noinline__ void pg_attn(const PgParams &p, int o, int n_past, float *s_att, int *s_flag)
{
    const PgOp &op = p.ops[o];
    const int ctid = threadIdx.x, lane = ctid & 31, warp = ctid >> 5;
    const int H = op.H, d = op.d, E = H * d, half = op.n_rot / 2, P = n_past + 1;
    const int nch = (P + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK, items = H * nch;
    const int dpad = (d + 3) & ~3;
    float *s_q = s_att, *s_kn = s_q + dpad, *s_vn = s_kn + dpad, *s_sc = s_vn + dpad, *s_o = s_sc + PG_ATT_CHUNK + 8;
    const float *gq = op.q, *gk = op.k, *gv = op.v;
    float *mk = op.mem_k, *mv = op.mem_v;
    const float scale = op.att_scale;
    const int dj = (d + 31) & ~31, ng = PG_CTHREADS / dj;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int h = item / nch, c = item - h * nch;
        const int p_lo = c * PG_ATT_CHUNK, np = min(PG_ATT_CHUNK, P - p_lo);
        const bool has_new = p_lo + np == P;
        pg_bar();                                            // the previous item's scratch is no longer read
        for (int j = ctid; j < d; j += PG_CTHREADS) {
            float qv = __ldcg(gq + h * d + j), kv = 0.f, vv = 0.f;
            if (has_new) { kv = __ldcg(gk + h * d + j); vv = __ldcg(gv + h * d + j); }
            if (j < 2 * half) {
                const bool fh = j < half;
                const int jp = fh ? j + half : j - half, i0 = fh ? j : j - half;
                const double2 cs = op.rope[(size_t)n_past * half + i0];
                qv = pg_rope(qv, __ldcg(gq + h * d + jp), fh, cs);
                if (has_new) kv = pg_rope(kv, __ldcg(gk + h * d + jp), fh, cs);
            }
            s_q[j] = qv;
            if (has_new) {
                s_kn[j] = kv; s_vn[j] = vv;
                mk[(size_t)n_past * E + h * d + j] = kv;     // cpy into the KV memory (vsim.cpp:557-558), K roped in place (:573-579)
                mv[(size_t)n_past * E + h * d + j] = vv;
            }
        }
        pg_bar();
        for (int pl = warp; pl < np; pl += GEMV_WARPS) {     // K . Q  (ggml.c:4484-4516), then ggml_scale
            const int pos = p_lo + pl;
            float sum = 0.f;
            if (pos == n_past) { for (int j = lane; j < d; j += 32) sum = fmaf(s_kn[j], s_q[j], sum); }
            else {
                const float *kr = mk + (size_t)pos * E + h * d;
                for (int j = lane; j < d; j += 32) sum = fmaf(__ldcg(kr + j), s_q[j], sum);
            }
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, of);
            if (lane == 0) s_sc[pl] = sum * scale;
        }
        pg_bar();
        if (warp == 0) {                                     // soft_max of the chunk (ggml.c:5825-5893), local max / denominator
            const float v0 = lane < np ? s_sc[lane] : -INFINITY, v1 = lane + 32 < np ? s_sc[lane + 32] : -INFINITY;
            float m = fmaxf(v0, v1);
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, of));
            const float e0 = lane < np ? expf(v0 - m) : 0.f, e1 = lane + 32 < np ? expf(v1 - m) : 0.f;
            float l = e0 + e1;
#pragma unroll
            for (int of = 16; of > 0; of >>= 1) l += __shfl_xor_sync(0xffffffffu, l, of);
            s_sc[lane] = e0; s_sc[lane + 32] = e1;
            if (lane == 0) { s_sc[PG_ATT_CHUNK] = m; s_sc[PG_ATT_CHUNK + 1] = l; }
        }
        pg_bar();
        {                                                    // V^T . softmax (ggml.c:4517-4553): position groups in parallel, merged in order
            const int g = ctid / dj, j = ctid - g * dj;
            if (g < ng && j < d) {
                float a = 0.f;
                for (int pl = g; pl < np; pl += ng) {
                    const int pos = p_lo + pl;
                    const float vv = pos == n_past ? s_vn[j] : __ldcg(mv + (size_t)pos * E + h * d + j);
                    a = fmaf(s_sc[pl], vv, a);
                }
                s_o[g * dj + j] = a;
            }
        }
        pg_bar();
        if (ctid < d) {
            float a = 0.f;
            for (int g = 0; g < ng; g++) a += s_o[g * dj + ctid];
            if (nch == 1) op.att_out[h * d + ctid] = a / s_sc[PG_ATT_CHUNK + 1];
            else {
                float *pt = op.att_part + ((size_t)h * op.max_chunks + c) * (d + 2);
                pt[ctid] = a;
                if (ctid == 0) { pt[d] = s_sc[PG_ATT_CHUNK]; pt[d + 1] = s_sc[PG_ATT_CHUNK + 1]; }
            }
        }
        if (nch > 1) {                                       // the head's last chunk to arrive merges the partials
            __threadfence();
            pg_bar();
            if (ctid == 0) {
                int *cnt = p.ctr + op.cnt_off + h;
                const int last = atomicAdd(cnt, 1) == nch - 1;
                if (last) *cnt = 0;
                *s_flag = last;
            }
            pg_bar();
            if (*s_flag) {
                __threadfence();
                if (ctid < d) {
                    const float *pt = op.att_part + (size_t)h * op.max_chunks * (d + 2);
                    float M = -INFINITY;
                    for (int cc = 0; cc < nch; cc++) M = fmaxf(M, __ldcg(pt + (size_t)cc * (d + 2) + d));
                    float L = 0.f, O = 0.f;
                    for (int cc = 0; cc < nch; cc++) {
                        const float w = expf(__ldcg(pt + (size_t)cc * (d + 2) + d) - M);
                        L = fmaf(__ldcg(pt + (size_t)cc * (d + 2) + d + 1), w, L);
                        O = fmaf(__ldcg(pt + (size_t)cc * (d + 2) + ctid), w, O);
                    }
                    op.att_out[h * d + ctid] = O / L;
                }
            }
        }
    }
    __threadfence();
    pg_bar();
    if (ctid == 0) atomicAdd(p.ctr + 1 + o, 1);
}

// y[n] = sum of the K-slice partials of a sliced mul_mat (only where a caller wants the finished vector in memory;
// inside a program the next op adds the slices itself while it builds its activation pack)
__device__ __noinline__ void pg_sum(const PgParams &p, int o)
{
    const PgOp &op = p.ops[o];
    const int ctid = threadIdx.x;
    const int n4 = op.N / 4;
    for (int i = (int)blockIdx.x * PG_CTHREADS + ctid; i < n4; i += (int)gridDim.x * PG_CTHREADS)
        reinterpret_cast<float4 *>(op.y)[i] = pg_src4(op.src[0], i * 4);
    for (int n = n4 * 4 + (int)blockIdx.x * PG_CTHREADS + ctid; n < op.N; n += (int)gridDim.x * PG_CTHREADS) {   // N % 4 tail
        const float *b = static_cast<const float *>(op.src[0].ptr) + n;
        float a = __ldcg(b);
        for (int q = 1; q < op.src[0].np; q++) a += __ldcg(b + (size_t)q * op.src[0].stride);
        op.y[n] = a * op.src[0].scale;
    }
    __threadfence();
    pg_bar();
    if (ctid == 0) atomicAdd(p.ctr + 1 + o, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// Start of a mul_mat op on the consumer side: dependency wait + activation pack.  Returns the op's geometry.
// ---------------------------------------------------------------------------------------------------------------
struct PgGeom { int C, cbase, plane; };

template <int QT>
__device__ __noinline__ PgGeom pg_begin_op(const PgParams &p, int o, int &waited_mv, float *s_x, uint8_t *s_aq, float *s_ap, double *s_stat)
{
    constexpr int NP = (QT == 0) ? 2 : 3;
    const PgOp &op = p.ops[o];
    const int ctid = threadIdx.x, lane = ctid & 31, warp = ctid >> 5;
    const bool sliced = op.PP > 1;
    const int cls = sliced ? (int)blockIdx.x % op.PP : 0;
    const int cbase = cls * op.TS, blk0 = cbase * CHUNK_BLKS;
    const int nbp = sliced ? op.TS * CHUNK_BLKS : ((op.C * CHUNK_BLKS + 7) & ~7);
    const int nb_l = min(op.nb - blk0, nbp);
    const int plane = nbp;
    const int in_mode = op.in_mode, res = op.res, nsrc = op.nsrc, K = op.K;
    pg_bar();                                                // every consumer warp is done with the previous op's pack
    // records of padded blocks (weight scale 0, but 0 x garbage could be NaN) read as zero
    for (int i = nb_l * 4 + ctid; i < nbp * 4; i += PG_CTHREADS) reinterpret_cast<uint2 *>(s_aq)[i] = make_uint2(0u, 0u);
    for (int pl = 0; pl < NP; pl++)
        for (int i = nb_l + ctid; i < nbp; i += PG_CTHREADS) s_ap[pl * plane + i] = 0.0f;
    const int dep = op.dep;
    if (dep >= 0 && dep > waited_mv) {
        if (ctid == 0) pg_wait_ctr(p.ctr + 1 + dep, (int)gridDim.x, p.err, o + 1);
        pg_bar();
        if (p.ops[dep].kind == PG_MATVEC) waited_mv = dep;   // a complete mul_mat op implies every earlier op is complete
    }
    auto block_sum = [&](double v) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, of));
        if (lane == 0) s_stat[warp] = v;
        pg_bar();
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < GEMV_WARPS; i++) t = __dadd_rn(t, s_stat[i]);
        pg_bar();
        return t;
    };
    const int items = nb_l * 4;                              // one item = 8 consecutive elements of one quant block
    if (in_mode == PG_IN_LN) {
        float4 *sx4 = reinterpret_cast<float4 *>(s_x);
        // the value layer-norm sees at float4 index k4
        auto ln_in = [&](int k4) {
            float4 a = sx4[k4];
            if (res == PG_RES_TEMP) {
                const float4 t = pg_sum_src4(op, nsrc, k4 * 4);
                a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
            }
            return a;
        };
        if (res != PG_RES_KEEP) {
            // ggml_norm (ggml.c:4282-4298): mean and variance accumulated in double (tree order here), two passes
            double ls = 0.0;
            for (int k4 = ctid; k4 < K / 4; k4 += PG_CTHREADS) {
                float4 a;
                if (res == PG_RES_TEMP) a = ln_in(k4);
                else {
                    a = pg_sum_src4(op, nsrc, k4 * 4);
                    if (res == PG_RES_ADD) {
                        const float4 x = sx4[k4];
                        a.x = __fadd_rn(x.x, a.x); a.y = __fadd_rn(x.y, a.y); a.z = __fadd_rn(x.z, a.z); a.w = __fadd_rn(x.w, a.w);
                    }
                    sx4[k4] = a;
                }
                ls = __dadd_rn(ls, __dadd_rn(__dadd_rn((double)a.x, (double)a.y), __dadd_rn((double)a.z, (double)a.w)));
            }
            const double mean = block_sum(ls) / (double)K;
            double l2 = 0.0;
            for (int k4 = ctid; k4 < K / 4; k4 += PG_CTHREADS) {
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
            pg_bar();
        }
        const double mean = s_stat[32];
        const float rstd = (float)s_stat[33];
        const float *lw = op.ln_w, *lb = op.ln_b;
        for (int base = 0; base < items; base += PG_CTHREADS) {
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
    } else {
        const uint16_t *tab = op.gelu_tab;
        for (int base = 0; base < items; base += PG_CTHREADS) {
            const int id = base + ctid;
            const bool valid = id < items;
            const int blk = valid ? id >> 2 : 0;
            const int k = (blk0 + blk) * QK + (lane & 3) * 8;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            if (valid) {
                lo = pg_sum_src4(op, nsrc, k);
                hi = pg_sum_src4(op, nsrc, k + 4);
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
    pg_bar();
    PgGeom gm;
    gm.C = op.C; gm.cbase = cbase; gm.plane = plane;
    return gm;
}

template <int QT>
__global__ void __launch_bounds__(PG_THREADS, 1) q4_prog_kernel(const PgParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int CH = chunk_bytes(QT);
    constexpr int STAGE_BYTES = PG_UNITS * CH;
    uint8_t *s_stage = smem;
    float   *s_red = reinterpret_cast<float *>(smem + (size_t)p.stages * STAGE_BYTES);
    float   *s_x = s_red + (size_t)p.stages * PG_UNITS * PG_RW;
    uint8_t *s_aq = reinterpret_cast<uint8_t *>(s_x + p.e_cap);
    float   *s_ap = reinterpret_cast<float *>(s_aq + (size_t)p.nbp_cap * 32);
    float   *s_att = s_ap + 3 * (size_t)p.nbp_cap;
    double  *s_stat = reinterpret_cast<double *>(s_att + p.att_cap);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(s_stat + 40);
    uint64_t *bar_empty = bar_full + PG_MAX_STAGES;
    uint64_t *bar_done = bar_empty + PG_MAX_STAGES;
    int4     *s_desc = reinterpret_cast<int4 *>(bar_done + PG_MAX_STAGES);   // {units (-1 END of op, -2 STOP), first chunk index, first tile, task}
    int      *s_dop = reinterpret_cast<int *>(s_desc + PG_MAX_STAGES);       // op index of the stage
    int      *s_flag = s_dop + PG_MAX_STAGES + 2;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bid = (int)blockIdx.x, grid = (int)gridDim.x;

    if (threadIdx.x < 3 * PG_MAX_STAGES) {
        const int i = threadIdx.x % PG_MAX_STAGES, which = threadIdx.x / PG_MAX_STAGES;
        if (i < p.stages) {
            if (which == 0) mbar_init(&bar_full[i], 1);
            else if (which == 1) mbar_init(&bar_empty[i], 1);
            else mbar_init(&bar_done[i], GEMV_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == GEMV_WARPS) {
        // ===================== producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;
            auto put = [&](int cnt, int c0, int T, int task, int o, const uint8_t *src) {
                pg_mbar_wait(&bar_empty[s], ph);
                s_desc[s] = make_int4(cnt, c0, T, task);
                s_dop[s] = o;
                if (cnt > 0) {
                    mbar_arrive_expect_tx(&bar_full[s], (uint32_t)cnt * CH);
                    bulk_g2s(s_stage + (size_t)s * STAGE_BYTES, src, (uint32_t)cnt * CH, &bar_full[s]);
                } else {
                    mbar_arrive(&bar_full[s]);
                }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            };
            for (int o = 0; o < p.n_ops; o++) {
                const PgOp &op = p.ops[o];
                if (op.kind != PG_MATVEC) continue;
                const uint8_t *wt = op.wt;
                const int C = op.C, PP = op.PP, TS = op.TS, U = op.U, nstatic = op.nstatic;
                const bool sliced = PP > 1;
                const int cls = sliced ? bid % PP : 0, cbase = cls * TS;
                const int idx = sliced ? bid / PP : bid;
                const int ncls = sliced ? (grid - cls + PP - 1) / PP : grid;
                const int limit = sliced ? op.n_tiles : op.ntasks;
                auto issue = [&](int u0, int n, int T, int c0, int task) {
                    for (int r = 0; r < n; r += PG_UNITS) {
                        const int cnt = min(PG_UNITS, n - r);
                        put(cnt, c0, T, task, o, wt + (size_t)(u0 + r) * CH);
                        c0 += cnt;
                        while (c0 >= C) { c0 -= C; ++T; }
                    }
                };
                auto issue_seq = [&](int t) {
                    if (sliced) issue(t * C + cbase, min(TS, C - cbase), t, cbase, t * PP + cls);
                    else if (C >= PIPE_TILE_SPLIT) issue(t * C, C, t, 0, t);                 // one whole tile
                    else issue(t * TS, min(TS, U - t * TS), t * (TS / C), 0, t);            // short tiles: TS / C whole tiles
                };
                for (int k = 0; k < nstatic; k++) {
                    const int t = idx + k * ncls;
                    if (t < limit) issue_seq(t);
                }
                int *tk = p.ctr + op.tk_off + cls;
                const int base = nstatic * ncls;
                int cur = base + atomicAdd(tk, 1);
                while (cur < limit) {
                    const int nxt = base + atomicAdd(tk, 1);       // one task ahead: the atomic's latency hides behind the copies
                    issue_seq(cur);
                    cur = nxt;
                }
                put(-1, 0, 0, 0, o, nullptr);
            }
            put(-2, 0, 0, 0, p.n_ops, nullptr);
        }
        return;
    }

    if (warp == GEMV_WARPS + 1) {
        // ===================== reducer =====================
        constexpr float post = (QT == 0) ? 0.0625f : 1.0f;   // undo the 16x of the A/B code scaling (exact)
        float acc = 0.f;
        int curT = -1, curTask = -1, curOp = -1;
        float *y = nullptr;
        const float *bias = nullptr;
        int N = 0, C = 1, PP = 1, y_slice = 0;
        const int rl = lane & 15;
        auto emit = [&]() {
            if (curT < 0) return;
            const int n = curT * TILE_ROWS + lane;
            if (lane < TILE_ROWS && n < N) {
                if (PP == 1) y[n] = bias ? __fadd_rn(acc * post, __ldg(bias + n)) : acc * post;
                else y[(size_t)(curTask - curT * PP) * y_slice + n] = acc * post;
            }
            curT = -1;
        };
        int s = 0;
        uint32_t ph = 0;
        for (;;) {
            pg_mbar_wait(&bar_done[s], ph);                  // every consumer warp has passed this slot (markers included)
            const int4 dsc = s_desc[s];
            const int o = s_dop[s];
            if (dsc.x == -2) break;
            if (o != curOp) {
                const PgOp &op = p.ops[o];
                y = op.y; bias = op.bias; N = op.N; C = op.C; PP = op.PP; y_slice = op.y_slice;
                curOp = o;
            }
            if (dsc.x < 0) {                                 // END of op o for this CTA: publish
                emit();
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(p.ctr + 1 + o, 1);
            } else {
                const int task = dsc.w;
                int T = dsc.z;
                int left = C - dsc.y;
                const float *red = s_red + (size_t)s * PG_UNITS * PG_RW + rl;
                int j = 0;
                while (j < dsc.x) {
                    if (T != curT || task != curTask) {
                        emit();
                        acc = 0.f;
                        curT = T; curTask = task;
                    }
                    const int run = min(dsc.x - j, left);
                    if (run == PG_UNITS) {
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
            if (lane == 0) mbar_arrive(&bar_empty[s]);
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        // the last CTA to get here clears every counter for the next launch (nobody polls or draws any more:
        // each CTA's consumers and producer are past their last use when its reducer sees STOP)
        __syncwarp();
        int last = 0;
        if (lane == 0) { __threadfence(); last = atomicAdd(p.ctr, 1) == grid - 1; }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            for (int i = lane; i < p.n_ctr; i += 32) p.ctr[i] = 0;
        }
        return;
    }

    // ===================== consumers =====================
    const int n_past = p.dyn ? p.dyn[1] : 0;
    int s = 0;
    uint32_t ph = 0;
    int cur = -1, waited_mv = -1;
    PgGeom gm;
    gm.C = 1; gm.cbase = 0; gm.plane = 0;
    auto run_attn_ops = [&](int from, int to) {
        for (int a = from; a < to; a++) {
            if (p.ops[a].kind == PG_MATVEC) continue;
            const int dep = p.ops[a].dep;
            if (dep >= 0 && dep > waited_mv) {
                pg_bar();
                if (threadIdx.x == 0) pg_wait_ctr(p.ctr + 1 + dep, grid, p.err, a + 1);
                pg_bar();
                if (p.ops[dep].kind == PG_MATVEC) waited_mv = dep;
            }
            if (p.ops[a].kind == PG_ATTN) pg_attn(p, a, n_past, s_att, s_flag);
            else pg_sum(p, a);
        }
    };
    const int uA = 2 * warp, uB = uA + 1;
    auto put_rows = [&](float *slot, float v0, float v2) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, 1); v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
        const int g = lane >> 2, t = lane & 3;
        if (t < 2) slot[g + 8 * t] = t ? v2 : v0;
    };
    for (;;) {
        pg_mbar_wait(&bar_full[s], ph);
        const int4 dsc = s_desc[s];
        const int o = s_dop[s];
        if (o != cur) {
            run_attn_ops(cur + 1, o);                        // (STOP carries o = n_ops: trailing attention ops run too)
            if (dsc.x != -2) gm = pg_begin_op<QT>(p, o, waited_mv, s_x, s_aq, s_ap, s_stat);
            cur = o;
        }
        if (dsc.x > 0) {
            const bool hA = uA < dsc.x, hB = uB < dsc.x;
            int TA = dsc.z, TB = dsc.z;
            int cA = dsc.y + uA, cB = dsc.y + uB;
            while (cA >= gm.C) { cA -= gm.C; ++TA; }
            while (cB >= gm.C) { cB -= gm.C; ++TB; }
            const uint8_t *chA = s_stage + (size_t)s * STAGE_BYTES + (size_t)uA * CH, *chB = chA + CH;
            float *slotA = s_red + ((size_t)s * PG_UNITS + uA) * PG_RW, *slotB = slotA + PG_RW;
            if (hA && hB && TA == TB && !(cA & 1)) {         // an aligned octet: the common case
                float v0, v2;
                compute_octet<QT, false>(chA, chB, (cA - gm.cbase) >> 1, s_aq, s_ap, gm.plane, lane, v0, v2);
                put_rows(slotA, v0, v2);
                if (lane < TILE_ROWS) slotB[lane] = 0.f;
            } else {                                         // half octets (odd start, tile boundary, ragged tail)
                if (hA) {
                    float v0, v2;
                    compute_octet<QT, false>((cA & 1) ? nullptr : chA, (cA & 1) ? chA : nullptr, (cA - gm.cbase) >> 1, s_aq, s_ap, gm.plane, lane, v0, v2);
                    put_rows(slotA, v0, v2);
                }
                if (hB) {
                    float v0, v2;
                    compute_octet<QT, false>((cB & 1) ? nullptr : chB, (cB & 1) ? chB : nullptr, (cB - gm.cbase) >> 1, s_aq, s_ap, gm.plane, lane, v0, v2);
                    put_rows(slotB, v0, v2);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_done[s]);
        if (dsc.x == -2) break;
        if (++s == p.stages) { s = 0; ph ^= 1; }
    }
}

#endif // Q4_PROG_TYPES_ONLY

} // namespace b200q4
