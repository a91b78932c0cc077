// neox_graph.cu -- device-resident GPT-NeoX forward pass (first slice).
//
// Mirrors gptneox_eval (vsim.cpp:470-747) for the graph the reference actually runs (parallel residual, F32 KV
// memory): token ids in, logits of the last token out, everything in between resident in HBM -- one H2D of N ints and
// one D2H of n_vocab floats per call instead of a PCIe round trip per mul_mat.  The Q4 mul_mats go through the public
// C ABI of this library (b200q4_mul_mat_dev: fast kernels, or the bit-identical ones in exact mode); the other ops are
// the kernels below.  They are CORRECTNESS-FIRST: each follows the arithmetic of the reference op it replaces exactly
// -- the order of every accumulation, double where ggml.c uses ggml_float, no FMA contraction, gelu and exp through
// fp16 tables built on the host with the same libm calls as ggml_init (ggml.c:1240-1255), rope angles from the host's
// pow/cos/sin -- so that with exact-mode mul_mats the logits equal the reference program's bit for bit
// (tests/test_gpu_neox_graph.py checks exactly that).
//
// Decode steps (one token, exact mode off) do not use these kernels at all: b200neox_load compiles the whole token --
// every mul_mat, layer-norm, gelu, the attention of every layer, the residual stream -- into ONE device-resident
// work-list executed by one persistent launch of q4_prog_kernel (q4_prog.cuh).  Prompts (N > 1) and exact mode run
// the per-op kernels below.
#include "../../include/b200neox.h"
#include "../../include/b200q4.h"
#define Q4_PROG_TYPES_ONLY                               // the work-list structures, not the kernel
#include "q4_prog_host.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

thread_local std::string g_nerr;
int nfail(int code, const std::string &msg) { g_nerr = msg; return code; }

#define NCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return nfail(B200Q4_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ---- kernels: one thread per output element / row, the reference's own operation order ----------------------------

// ggml_get_rows on a Q4 matrix (dequantize_row_q4_0 ggml.c:300-330: v = (q - 8) * d; _q4_1 :332-362: v = q * d + m)
__global__ void k_get_rows_q4(int type, const unsigned char *__restrict__ w, int K, const int *__restrict__ ids, int n_ids, float *__restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n_ids * K) return;
    const int r = (int)(i / K), e = (int)(i % K), blk = e / 32, l = e % 32, nb = K / 32;
    const unsigned char *row = w + (size_t)ids[r] * nb * (type == 0 ? 20 : 24);
    if (type == 0) {
        const unsigned char *b = row + (size_t)blk * 20;
        const float d = *reinterpret_cast<const float *>(b);
        const unsigned char v = b[4 + l / 2];
        const int q = (l & 1) ? (v >> 4) : (v & 15);
        y[i] = __fmul_rn((float)(q - 8), d);
    } else {
        const float m = reinterpret_cast<const float *>(row)[blk], d = reinterpret_cast<const float *>(row)[nb + blk];
        const unsigned char v = row[(size_t)nb * 8 + (size_t)blk * 16 + l / 2];
        const int q = (l & 1) ? (v >> 4) : (v & 15);
        y[i] = __fadd_rn(__fmul_rn((float)q, d), m);
    }
}

// ggml_norm (ggml.c:4246-4303) followed by mul(repeat(w), .) and add(., repeat(b)) (vsim.cpp:523-533): one thread per row
__global__ void k_layernorm(const float *__restrict__ x, int n_rows, int n, const float *__restrict__ w, const float *__restrict__ b,
                            float *__restrict__ y)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const float *xr = x + (size_t)r * n;
    float *yr = y + (size_t)r * n;
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean = __dadd_rn(mean, (double)xr[i]);
    mean = mean / (double)n;
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) {
        const double v = __dsub_rn((double)xr[i], mean);
        yr[i] = (float)v;
        sum2 = __dadd_rn(sum2, __dmul_rn(v, v));
    }
    const double eps = (double)1e-5f;
    const float scale = (float)(1.0 / sqrt(__dadd_rn(sum2 / (double)n, eps)));
    for (int i = 0; i < n; i++) yr[i] = __fadd_rn(__fmul_rn(w[i], __fmul_rn(yr[i], scale)), b[i]);
}

// The same operator for the fast mode (exact mode off): one 256-thread block per row, the two double accumulations as
// parallel tree reductions.  Every float operation is the one above; only the ORDER of the double additions differs,
// i.e. mean and variance may move by an ulp of a double (1e-16) before they are rounded to the float scale.
__global__ void __launch_bounds__(256) k_layernorm_fast(const float *__restrict__ x, int n, const float *__restrict__ w,
                                                         const float *__restrict__ b, float *__restrict__ y)
{
    __shared__ double sh[256];
    const float *xr = x + (size_t)blockIdx.x * n;
    float *yr = y + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x;
    auto block_sum = [&](double v) {
        sh[tid] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) sh[tid] = __dadd_rn(sh[tid], sh[tid + o]);
            __syncthreads();
        }
        const double t = sh[0];
        __syncthreads();
        return t;
    };
    double s = 0.0;
    for (int i = tid; i < n; i += 256) s = __dadd_rn(s, (double)xr[i]);
    const double mean = block_sum(s) / (double)n;
    double s2 = 0.0;
    for (int i = tid; i < n; i += 256) {
        const double v = __dsub_rn((double)xr[i], mean);
        s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    const double eps = (double)1e-5f;
    const float scale = (float)(1.0 / sqrt(__dadd_rn(block_sum(s2) / (double)n, eps)));
    for (int i = tid; i < n; i += 256) {
        const float v = (float)__dsub_rn((double)xr[i], mean);
        yr[i] = __fadd_rn(__fmul_rn(w[i], __fmul_rn(v, scale)), b[i]);
    }
}

// y[r][i] = a[r][i] + bias[i]   (ggml_add with a repeated bias; float addition commutes)
__global__ void k_add_bias(float *__restrict__ a, const float *__restrict__ bias, long total, int n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) a[i] = __fadd_rn(a[i], bias[i % n]);
}
__global__ void k_add(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ y, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) y[i] = __fadd_rn(a[i], b[i]);
}
__global__ void k_copy(const float *__restrict__ a, float *__restrict__ y, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) y[i] = a[i];
}

// ggml_gptneox_rope (ggml.c:6086-6153), in place on x[ne2][ne1][ne0]; cs[p][i0] = {cos, sin}(p * theta_i0) from the host
__global__ void k_rope(float *__restrict__ x, int ne0, int ne1, int ne2, int n_past, int n_dims, int mode, const double2 *__restrict__ cs)
{
    const int half = n_dims / 2;
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int first = mode == 0 ? 0 : n_past;
    if (id >= (long)(ne2 - first) * ne1 * half) return;
    const int i0 = (int)(id % half);
    const int i1 = (int)((id / half) % ne1);
    const int i2 = first + (int)(id / ((long)half * ne1));
    const int p = mode == 0 ? n_past + i2 : i2;
    const double2 c = cs[(size_t)p * half + i0];
    float *d = x + ((size_t)i2 * ne1 + i1) * ne0;
    const double x1 = d[i0], x2 = d[i0 + half];
    d[i0] = (float)__dsub_rn(__dmul_rn(c.x, x1), __dmul_rn(c.y, x2));
    d[i0 + half] = (float)__dadd_rn(__dmul_rn(c.x, x2), __dmul_rn(c.y, x1));
}

// K * Q (ggml.c:4484-4516 + vec_dot_f32 :399-434): float products accumulated in double, in order; then ggml_scale and
// ggml_diag_mask_inf.  kq[h][n][p], K = the memory view [P][H][d], Q [N][H][d]
__global__ void k_kq(const float *__restrict__ km, const float *__restrict__ q, int H, int N, int P, int d, int n_past, float scale, float *__restrict__ kq)
{
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long)H * N * P) return;
    const int p = (int)(id % P), n = (int)((id / P) % N), h = (int)(id / ((long)P * N));
    const float *kr = km + ((size_t)p * H + h) * d, *qr = q + ((size_t)n * H + h) * d;
    double s = 0.0;
    for (int i = 0; i < d; i++) s = __dadd_rn(s, (double)__fmul_rn(kr[i], qr[i]));
    const float v = __fmul_rn((float)s, scale);
    kq[id] = p > n_past + n ? -INFINITY : v;
}

// ggml_soft_max (ggml.c:5825-5893), in place: exp through the fp16 table, sum in double; one thread per row
__global__ void k_soft_max(float *__restrict__ p, int n_rows, int n, const unsigned short *__restrict__ t_exp)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float *pr = p + (size_t)r * n;
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) mx = pr[i] > mx ? pr[i] : mx;
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        if (pr[i] == -INFINITY) { pr[i] = 0.0f; continue; }
        const __half hx = __float2half_rn(__fsub_rn(pr[i], mx));
        const float val = __half2float(__ushort_as_half(t_exp[__half_as_ushort(hx)]));
        sum = __dadd_rn(sum, (double)val);
        pr[i] = val;
    }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; i++) pr[i] = __fmul_rn(pr[i], inv);
}

// V^T * softmax, ggml's transposed f32 mul_mat (ggml.c:4517-4553 + FINALIZE :4463-4481): thread t of nth owns positions
// [t*dc, (t+1)*dc), float accumulation in position order from +0, then dst = part[0]; dst += part[t].  Output written
// already merged: out[n][h*d + j]  (permute(0,2,1,3) + cpy, vsim.cpp:621-627)
__global__ void k_kqv(const float *__restrict__ vm, const float *__restrict__ kq, int H, int N, int P, int d, int nth, float *__restrict__ out)
{
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long)H * N * d) return;
    const int j = (int)(id % d), n = (int)((id / d) % N), h = (int)(id / ((long)d * N));
    const float *pr = kq + ((size_t)h * N + n) * P;
    const int dc = (P + nth - 1) / nth;
    float total = 0.0f;
    for (int t = 0; t < nth; t++) {
        float part = 0.0f;
        const int hi = min((t + 1) * dc, P);
        for (int ic = t * dc; ic < hi; ic++) part = __fadd_rn(part, __fmul_rn(vm[((size_t)ic * H + h) * d + j], pr[ic]));
        total = t == 0 ? part : __fadd_rn(total, part);
    }
    out[(size_t)n * H * d + (size_t)h * d + j] = total;
}

// ggml_gelu (ggml.c:819-827, GGML_GELU_FP16): x -> fp16 -> table -> fp32
__global__ void k_gelu(float *__restrict__ x, long total, const unsigned short *__restrict__ t_gelu)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) x[i] = __half2float(__ushort_as_half(t_gelu[__half_as_ushort(__float2half_rn(x[i]))]));
}

inline unsigned nblk(long total, int threads = 256) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace

struct NeoxLayer {
    b200q4_weight *wq, *wk, *wv, *wo, *w1, *w2;
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *bq, *bk, *bv, *bo, *b1, *b2;
};

struct b200neox {
    b200q4_model *file = nullptr;
    int n_vocab = 0, n_embd = 0, n_head = 0, n_layer = 0, n_rot = 0, par_res = 0, ftype = 0, n_ctx = 0;
    std::vector<NeoxLayer> layers;                       // every tensor resolved once, at load
    b200q4_weight *w_head = nullptr;
    const float *lnf_w = nullptr, *lnf_b = nullptr;
    std::vector<float *> f32;                            // device copies of the 1-D tensors (owned)
    unsigned char *d_wte = nullptr;                      // embedding matrix in the reference's row format (get_rows dequantizes it)
    int wte_type = 0;
    float *mem_k = nullptr, *mem_v = nullptr;            // [n_layer][n_ctx][n_embd]   (vsim.cpp:318-330)
    unsigned short *t_gelu = nullptr, *t_exp = nullptr;  // ggml_init's fp16 tables
    double2 *rope_cs = nullptr;                          // [n_ctx][n_rot/2] {cos, sin}
    float *x = nullptr, *cur = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *ff = nullptr, *ff2 = nullptr,
          *kq = nullptr, *merged = nullptr, *logits = nullptr;
    int *ids = nullptr;
    int max_n = 0;
    cudaStream_t st = nullptr;                           // the library's stream: all device work of this process is ordered on it
    // row-sharded over `world` GPUs (decode program only): this rank's shards, the peers' exchange pools
    int rank = 0, world = 1;
    std::vector<b200q4_weight *> shards;                 // owned (the file's own tensors are not uploaded when sharded)
    char *pool[8] = {nullptr};
    size_t pool_bytes = 0;
    // decode program (one persistent launch per token)
    b200q4::Prog prog;
    bool have_prog = false;
    std::string prog_why;                                // why no program was built
    std::vector<float *> prog_bufs;                      // its scratch vectors (owned)
};

namespace {

using namespace b200q4;

// Compiles one decode step (N = 1) into a work-list for q4_prog_kernel.  Graph: gptneox_eval, vsim.cpp:518-718.
// Exchange-pool layout of a sharded program: [counters][per layer: attention output, MLP hidden, MLP output (K-slice
// partials), attention projection][logits], every vector as {value, tag} pairs.  The same on every rank.
struct PoolLayout {
    size_t ctr_bytes, total;
    size_t att, hid, out, proj, layer_stride, logits;
    int pp_out, pp_hid, pp_proj, pp_head, st_E, st_H, st_V;
};
PoolLayout pool_layout(int E, int L, int V, int n_ops_bound)
{
    PoolLayout p{};
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    p.st_E = (E + 3) & ~3; p.st_H = (4 * E + 3) & ~3; p.st_V = (V + 3) & ~3;
    p.pp_proj = pg_geometry_dims(E, E).PP;                 // (the slicing depends on K only: a shard slices like the full tensor)
    p.pp_hid = pg_geometry_dims(4 * E, E).PP;
    p.pp_out = pg_geometry_dims(E, 4 * E).PP;
    p.pp_head = pg_geometry_dims(V, E).PP;
    p.ctr_bytes = al((size_t)n_ops_bound * 4);
    size_t off = 0;
    p.att = off; off += al((size_t)2 * p.st_E * 4);
    p.hid = off; off += al((size_t)2 * p.pp_hid * p.st_H * 4);
    p.out = off; off += al((size_t)2 * p.pp_out * p.st_E * 4);
    p.proj = off; off += al((size_t)2 * p.pp_proj * p.st_E * 4);
    p.layer_stride = off;
    p.logits = p.ctr_bytes + (size_t)L * p.layer_stride;
    p.total = p.logits + al((size_t)2 * p.pp_head * p.st_V * 4);
    return p;
}
int ops_bound(int L) { return 64 + 16 * (L * 8 + 4) + 8 * 64 * (L + 1); }      // counters: arrivals, tickets (<= 8 classes), head counters

// Compiles one decode step (N = 1) into a work-list for q4_prog_kernel.  Graph: gptneox_eval, vsim.cpp:518-718.
// world > 1: rank g owns rows [g N/G, (g+1) N/G) of every 2-D tensor (imax.c:1185-1188) and heads [g H/G, (g+1) H/G);
// Q, K, V and the KV memory stay local, every other vector is stored by its producers into EVERY rank's exchange pool.
int build_decode_program(b200neox *m)
{
    const int E = m->n_embd, H = m->n_head, d = E / H, L = m->n_layer, V = m->n_vocab, G = m->world, R = m->rank;
    if (d > PG_ATT_MAX_D) { m->prog_why = "head size above 512"; return 1; }
    if (m->wte_type != m->ftype - 2) { m->prog_why = "embedding matrix of another quant type"; return 1; }
    const int Hl = H / G, El = E / G;
    const PoolLayout pl = pool_layout(E, L, V, ops_bound(L));
    ProgBuild b;
    auto dev = [&](size_t floats) {
        float *p = nullptr;
        if (cudaMalloc(&p, (floats + 4) * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return (float *)nullptr; }
        cudaMemset(p, 0, (floats + 4) * sizeof(float));
        m->prog_bufs.push_back(p);
        return p;
    };
    auto pool_ptr = [&](int r, size_t off) { return reinterpret_cast<float *>(m->pool[r] + off); };
    // Every vector an op hands to a later op is a buffer of {value, tag} pairs (q4_prog.cuh): the reader polls the data.
    // The output of a mul_mat as a list of sources: the finished vector, or its K-slice partials plus the bias.
    struct Out { std::vector<PgSrc> src; int op; float *buf; };
    bool oom = false;
    // xoff: sharded and exchanged -> offset of the (full) vector in every rank's pool; (size_t)-1: a local vector
    auto matvec = [&](b200q4_weight *w, const float *bias, int in_mode, int res, const std::vector<PgSrc> &src,
                      const float *lw, const float *lb, float *plain_y, bool same_pack, size_t xoff, int stride_full) {
        Out o;
        o.buf = nullptr;
        const PgGeo g = pg_geometry(w);
        const int Nl = b200q4_weight_rows(w);
        const bool exch = G > 1 && xoff != (size_t)-1;
        const int row0 = exch ? R * Nl : 0;
        const int stride = exch ? stride_full : (Nl + 3) & ~3;
        float *y = plain_y;
        if (!y) y = exch ? pool_ptr(R, xoff) : dev(2 * (size_t)g.PP * stride);
        if (!y) { oom = true; o.op = -1; return o; }
        if (g.PP == 1) {
            o.op = b.add_matvec(w, -1, in_mode, res, src.data(), (int)src.size(), lw, lb, m->t_gelu, y, bias ? bias + row0 : nullptr, 0,
                                !plain_y, same_pack);
            o.src.push_back(pg_src_tag(y, o.op));
        } else {
            o.op = b.add_matvec(w, -1, in_mode, res, src.data(), (int)src.size(), lw, lb, m->t_gelu, y, nullptr, stride, true, same_pack);
            o.src.push_back(pg_src_tag(y, o.op, g.PP, stride));
            if (bias) o.src.push_back(pg_src_f32(bias));
        }
        if (o.op >= 0 && exch) {
            PgOp &op = b.ops[(size_t)o.op];
            op.row0 = row0;
            for (int r = 0; r < G; r++) op.y_peer[r] = pool_ptr(r, xoff);
        }
        o.buf = y;
        return o;
    };
    const size_t LOCAL = (size_t)-1;
    const size_t chunks = (size_t)(m->n_ctx + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK;
    float *part = dev((size_t)Hl * chunks * d), *att_sc = dev((size_t)Hl * (m->n_ctx + chunks));
    double *att_sum = reinterpret_cast<double *>(dev(2 * (size_t)Hl * chunks + 2));
    float *att_local = G == 1 ? dev(2 * (size_t)E) : nullptr;
    if (!part || !att_sc || !att_sum || (G == 1 && !att_local)) { m->prog_why = "device allocation failed"; return 1; }
    const float scale = (float)(1.0 / sqrt((double)((float)E / (float)H)));                  // vsim.cpp:590
    // what the next layer-norm op has to add to the residual stream first
    std::vector<PgSrc> pending;
    int pending_res = PG_RES_SET;
    PgSrc emb{m->d_wte, nullptr, m->wte_type, E, 1.0f, PG_SRC_Q4ROW};                         // ggml_get_rows(wte, embd), vsim.cpp:516
    pending.push_back(emb);
    const std::vector<PgSrc> none;
    bool bad = false;
    auto chk = [&](const Out &o) { if (o.op < 0) bad = true; return o; };
    for (int l = 0; l < L && !bad; l++) {
        const NeoxLayer &ly = m->layers[(size_t)l];
        float *mk = m->mem_k + (size_t)l * m->n_ctx * El, *mv = m->mem_v + (size_t)l * m->n_ctx * El;
        const size_t lo = pl.ctr_bytes + (size_t)l * pl.layer_stride;
        if ((int)pending.size() > PG_MAX_SRC) { m->prog_why = "too many residual sources"; return 1; }
        // (per-layer output buffers: a vector is only overwritten by ops that cannot start before all its readers are done)
        const Out oq = chk(matvec(ly.wq, ly.bq ? ly.bq + R * El : nullptr, PG_IN_LN, pending_res, pending, ly.ln1_w, ly.ln1_b, nullptr, false, LOCAL, 0));
        const Out ok = chk(matvec(ly.wk, ly.bk ? ly.bk + R * El : nullptr, PG_IN_LN, PG_RES_KEEP, none, ly.ln1_w, ly.ln1_b, nullptr, true, LOCAL, 0));
        const Out ov = chk(matvec(ly.wv, ly.bv ? ly.bv + R * El : nullptr, PG_IN_LN, PG_RES_KEEP, none, ly.ln1_w, ly.ln1_b, nullptr, true, LOCAL, 0));
        if (bad) break;
        if (oq.src.size() != 1 || oq.src[0].np != 1) { m->prog_why = "n_embd above 8192 (K-sliced Q/K/V)"; return 1; }
        pending.clear();
        float *att = G == 1 ? att_local : pool_ptr(R, lo + pl.att);
        auto attn = [&]() {
            const int oa = b.add_attn(-1, oq.buf, ok.buf, ov.buf, oq.op, ok.op, ov.op, mk, mv, att, part, att_sc, att_sum,
                                      m->rope_cs, m->t_exp, Hl, d, m->n_rot, m->n_ctx, scale);
            if (oa >= 0 && G > 1) {
                PgOp &op = b.ops[(size_t)oa];
                op.head0 = R * Hl;
                for (int r = 0; r < G; r++) op.att_peer[r] = pool_ptr(r, lo + pl.att);
            }
            return oa;
        };
        if (m->par_res == 1) {
            // vsim.cpp:666-696: the MLP reads the layer input, so it needs nothing from the attention.  List order: attention
            // right behind Q, K, V, then the MLP's first matrix, the attention's output projection, the MLP's second matrix --
            // no op reads what the op just before it wrote.  While only a few thread blocks have heads to work on (short
            // contexts, sharded models) their producers draw no task of the MLP's first matrix before their consumers are
            // through the attention (PgOp::hold_a): the other blocks stream it meanwhile.
            const int oa = attn();
            if (oa < 0) { bad = true; break; }
            const Out o1 = chk(matvec(ly.w1, ly.b1, PG_IN_LN, PG_RES_KEEP, none, ly.ln2_w, ly.ln2_b, nullptr, false, lo + pl.hid, pl.st_H));
            if (bad) break;
            if (!getenv("B200NEOX_NO_HOLD")) {
                PgOp &h = b.ops[(size_t)o1.op];
                h.hold_a = oa + 1; h.hold_H = Hl;
                h.dynamic = 1;                               // (in hold mode part of its tasks come from the ticket counter)
            }
            const std::vector<PgSrc> a_in{pg_src_tag(att, oa)};
            const Out oo = chk(matvec(ly.wo, ly.bo, PG_IN_PLAIN, PG_RES_KEEP, a_in, nullptr, nullptr, nullptr, false, lo + pl.proj, pl.st_E));
            if (bad) break;
            const Out o2 = chk(matvec(ly.w2, ly.b2, PG_IN_GELU, PG_RES_KEEP, o1.src, nullptr, nullptr, nullptr, false, lo + pl.out, pl.st_E));
            if (bad) break;
            pending = oo.src;                                                                // inpL = inpL + (cur + inpFF)
            pending.insert(pending.end(), o2.src.begin(), o2.src.end());
        } else {
            // vsim.cpp:636-665: inpFF = norm(cur + inpL) -> MLP; inpL = inpFF + inpL (the attention output only feeds the MLP)
            const int oa = attn();
            if (oa < 0) { bad = true; break; }
            const std::vector<PgSrc> a_in{pg_src_tag(att, oa)};
            const Out oo = chk(matvec(ly.wo, ly.bo, PG_IN_PLAIN, PG_RES_KEEP, a_in, nullptr, nullptr, nullptr, false, lo + pl.proj, pl.st_E));
            if (bad) break;
            const Out o1 = chk(matvec(ly.w1, ly.b1, PG_IN_LN, PG_RES_TEMP, oo.src, ly.ln2_w, ly.ln2_b, nullptr, false, lo + pl.hid, pl.st_H));
            if (bad) break;
            const Out o2 = chk(matvec(ly.w2, ly.b2, PG_IN_GELU, PG_RES_KEEP, o1.src, nullptr, nullptr, nullptr, false, lo + pl.out, pl.st_E));
            if (bad) break;
            pending = o2.src;
        }
        pending_res = PG_RES_ADD;
    }
    if (!bad) {
        if ((int)pending.size() > PG_MAX_SRC) { m->prog_why = "too many residual sources"; return 1; }
        const PgGeo g = pg_geometry(m->w_head);
        const bool plain = G == 1 && g.PP == 1;                                              // (sharded: the slices meet in the pool first)
        Out oh = chk(matvec(m->w_head, nullptr, PG_IN_LN, pending_res, pending, m->lnf_w, m->lnf_b, plain ? m->logits : nullptr, false,
                            pl.logits, pl.st_V));
        if (!bad && !plain && b.add_sum(-1, oh.src[0], m->logits, V) < 0) bad = true;
    }
    if (bad) { m->prog_why = oom ? "device allocation failed" : std::string("program builder: ") + b200q4_last_error(); return 1; }
    if (G > 1) {
        if (prog_counter_ints(b) * 4 > (int)pl.ctr_bytes) { m->prog_why = "counter area of the exchange pool too small"; return 1; }
        int *ctrs[8];
        for (int r = 0; r < G; r++) ctrs[r] = reinterpret_cast<int *>(m->pool[r]);
        if (prog_finalize_sharded(b, m->prog, R, G, ctrs)) { m->prog_why = std::string("program: ") + b200q4_last_error(); return 1; }
    } else if (prog_finalize(b, m->prog)) { m->prog_why = std::string("program: ") + b200q4_last_error(); return 1; }
    m->have_prog = true;
    return 0;
}

}  // namespace

extern "C" {

const char *b200neox_last_error(void) { return g_nerr.c_str(); }

void b200neox_free(b200neox *m)
{
    if (!m) return;
    if (m->st) cudaStreamSynchronize(m->st);
    if (m->have_prog) b200q4::prog_free(m->prog);
    for (float *p : m->prog_bufs) cudaFree(p);
    for (b200q4_weight *w : m->shards) b200q4_weight_free(w);
    for (float *p : m->f32) cudaFree(p);
    float *bufs[] = {m->mem_k, m->mem_v, m->x, m->cur, m->q, m->k, m->v, m->att, m->ff, m->ff2, m->kq, m->merged, m->logits};
    for (float *b : bufs) cudaFree(b);
    cudaFree(m->d_wte); cudaFree(m->t_gelu); cudaFree(m->t_exp); cudaFree(m->rope_cs); cudaFree(m->ids);
    if (m->file) b200q4_model_free(m->file);
    delete m;
}

static b200neox *neox_load_impl(const char *path, int n_ctx, int max_batch, int rank, int world, void *const *pool_peers, size_t pool_bytes)
{
    if (!path || n_ctx < 1 || max_batch < 1 || world < 1 || world > 8 || rank < 0 || rank >= world || (world > 1 && !pool_peers)) {
        nfail(B200Q4_EINVAL, "neox_load: bad argument");
        return nullptr;
    }
    b200q4_model *f = b200q4_model_load(path, 7, 0, world == 1 ? 1 : 0);   // vsim.cpp:119-138: 7 header ints, no vocabulary count
    if (!f) { nfail(B200Q4_EINVAL, std::string("neox_load: ") + b200q4_last_error()); return nullptr; }
    b200neox *m = new b200neox();
    m->file = f;
    m->rank = rank; m->world = world; m->pool_bytes = pool_bytes;
    for (int r = 0; r < world && pool_peers; r++) m->pool[r] = static_cast<char *>(pool_peers[r]);
    int *hp[] = {&m->n_vocab, &m->n_embd, &m->n_head, &m->n_layer, &m->n_rot, &m->par_res, &m->ftype};
    for (int i = 0; i < 7; i++) *hp[i] = b200q4_model_header(f, i);
    m->n_ctx = n_ctx; m->max_n = max_batch;
    auto bail = [&](const std::string &why) { nfail(B200Q4_EINVAL, "neox_load: " + why); b200neox_free(m); return (b200neox *)nullptr; };
    // hyper-parameters (the reference trusts the file; a bad one must not reach a kernel)
    if (m->n_vocab < 1 || m->n_embd < 1 || m->n_head < 1 || m->n_layer < 0 || m->n_rot < 0) return bail("bad hyper-parameters");
    if (m->n_embd % 32 || m->n_embd % m->n_head || m->n_rot % 2 || m->n_rot > m->n_embd / m->n_head) return bail("bad hyper-parameters");
    if (m->par_res != 0 && m->par_res != 1) return bail("use_parallel_residual must be 0 or 1 (vsim.cpp:697-700)");
    if (m->ftype != 2 && m->ftype != 3) return bail("only q4_0 / q4_1 model files are built (ftype 2 / 3)");
    if (world > 1) {
        if (m->n_head % world || m->n_embd % (16 * world) || m->n_vocab % world)
            return bail("the model does not shard evenly over this many GPUs (heads, 16-row tiles of n_embd, vocabulary rows)");
        for (int r = 0; r < world; r++) if (!m->pool[r]) return bail("NULL exchange pool");
        if (pool_bytes < pool_layout(m->n_embd, m->n_layer, m->n_vocab, ops_bound(m->n_layer)).total) return bail("exchange pool too small");
    }
    m->st = b200q4::dev_info().stream;
    if (!m->st) return bail("b200q4_init() has not succeeded");
    const int E = m->n_embd, V = m->n_vocab;
    // every tensor by name, with the shape the hyper-parameters imply ("has wrong shape in model file", vsim.cpp:375-448)
    std::map<std::string, int> index;
    const int nt = b200q4_model_tensor_count(f);
    for (int i = 0; i < nt; i++) {
        const char *name;
        b200q4_model_tensor_info(f, i, &name, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        index[name] = i;
    }
    std::string err;
    auto vec = [&](const std::string &name, int n) -> const float * {          // 1-D f32 tensor of n elements -> device copy
        auto it = index.find(name);
        if (it == index.end()) { if (err.empty()) err = "missing tensor " + name; return nullptr; }
        const char *nm; int nd, ft, ne0, ne1; size_t off, nb;
        b200q4_model_tensor_info(f, it->second, &nm, &nd, &ft, &ne0, &ne1, &off, &nb);
        if (nd != 1 || ft != 0 || ne0 != n) { if (err.empty()) err = "tensor '" + name + "' has wrong shape in model file"; return nullptr; }
        float *d = nullptr;
        if (cudaMalloc(&d, nb) != cudaSuccess || cudaMemcpy(d, b200q4_model_tensor_data(f, it->second), nb, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError(); cudaFree(d);
            if (err.empty()) err = "device copy of a 1-D tensor failed";
            return nullptr;
        }
        m->f32.push_back(d);
        return d;
    };
    auto mat = [&](const std::string &name, int K, int N, int *tensor_index = nullptr) -> b200q4_weight * {   // 2-D Q4 tensor [N rows][K]
        auto it = index.find(name);
        if (it == index.end()) { if (err.empty()) err = "missing tensor " + name; return nullptr; }
        const char *nm; int nd, ft, ne0, ne1; size_t off, nb;
        b200q4_model_tensor_info(f, it->second, &nm, &nd, &ft, &ne0, &ne1, &off, &nb);
        if (nd != 2 || ne0 != K || ne1 != N) { if (err.empty()) err = "tensor '" + name + "' has wrong shape in model file"; return nullptr; }
        if (ft != m->ftype) { if (err.empty()) err = "tensor '" + name + "' is not of the file's quant type"; return nullptr; }
        if (tensor_index) *tensor_index = it->second;
        if (world == 1 || tensor_index) return b200q4_model_weight(f, it->second);        // (tensor_index: the caller only wants the record)
        // sharded: this rank's rows [rank N/G, (rank+1) N/G) straight from the file mapping (imax.c:1185-1188 splits by rows)
        b200q4_weight *w = b200q4_weight_upload_rows(ft - 2, b200q4_model_tensor_data(f, it->second), N, K, 0, rank * (N / world), N / world);
        if (!w) { if (err.empty()) err = std::string("shard upload failed: ") + b200q4_last_error(); return nullptr; }
        m->shards.push_back(w);
        return w;
    };
    int wte_index = -1;
    mat("gpt_neox.embed_in.weight", E, V, &wte_index);
    m->lnf_w = vec("gpt_neox.final_layer_norm.weight", E);
    m->lnf_b = vec("gpt_neox.final_layer_norm.bias", E);
    m->w_head = mat("embed_out.weight", E, V);
    m->layers.resize((size_t)m->n_layer);
    for (int l = 0; l < m->n_layer && err.empty(); l++) {
        const std::string p = "gpt_neox.layers." + std::to_string(l) + ".";
        NeoxLayer &ly = m->layers[(size_t)l];
        ly.ln1_w = vec(p + "input_layernorm.weight", E);           ly.ln1_b = vec(p + "input_layernorm.bias", E);
        ly.ln2_w = vec(p + "post_attention_layernorm.weight", E);  ly.ln2_b = vec(p + "post_attention_layernorm.bias", E);
        ly.wq = mat(p + "attention.query.weight", E, E);           ly.bq = vec(p + "attention.query.bias", E);
        ly.wk = mat(p + "attention.key.weight", E, E);             ly.bk = vec(p + "attention.key.bias", E);
        ly.wv = mat(p + "attention.value.weight", E, E);           ly.bv = vec(p + "attention.value.bias", E);
        ly.wo = mat(p + "attention.dense.weight", E, E);           ly.bo = vec(p + "attention.dense.bias", E);
        ly.w1 = mat(p + "mlp.dense_h_to_4h.weight", E, 4 * E);     ly.b1 = vec(p + "mlp.dense_h_to_4h.bias", 4 * E);
        ly.w2 = mat(p + "mlp.dense_4h_to_h.weight", 4 * E, E);     ly.b2 = vec(p + "mlp.dense_4h_to_h.bias", E);
    }
    if (!err.empty()) return bail(err);
    {   // the embedding matrix is read by get_rows in the reference's row format
        const char *name; int nd, ft, ne0, ne1; size_t off, nb;
        b200q4_model_tensor_info(f, wte_index, &name, &nd, &ft, &ne0, &ne1, &off, &nb);
        m->wte_type = ft - 2;
        if (cudaMalloc(&m->d_wte, nb) != cudaSuccess ||
            cudaMemcpy(m->d_wte, b200q4_model_tensor_data(f, wte_index), nb, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail("device copy of the embedding matrix failed");
    }
    // ggml_init's tables (ggml.c:1240-1255), with this host's libm like the reference build
    std::vector<unsigned short> tg(1 << 16), te(1 << 16);
    for (unsigned i = 0; i < (1u << 16); i++) {
        __half_raw hr; hr.x = (unsigned short)i;
        const float fv = __half2float(__half(hr));
        const double GELU_COEF_A = 0.044715, SQRT_2_OVER_PI = 0.79788456;
        const float g = (float)(0.5 * fv * (1.0 + tanh(SQRT_2_OVER_PI * fv * (1.0 + GELU_COEF_A * fv * fv))));
        tg[i] = __half_raw(__float2half_rn(g)).x;
        te[i] = __half_raw(__float2half_rn((float)exp((double)fv))).x;
    }
    const int half = m->n_rot / 2;
    std::vector<double2> cs((size_t)n_ctx * half);
    for (int p = 0; p < n_ctx; p++)
        for (int i0 = 0; i0 < half; i0++) {
            const double theta = pow(10000.0, 2 * ((double)-i0) / m->n_rot);             // ggml.c:6126
            cs[(size_t)p * half + i0] = make_double2(cos(p * theta), sin(p * theta));
        }
    const size_t Es = (size_t)E, Nmax = (size_t)max_batch, kv = (size_t)m->n_layer * n_ctx * (Es / world) * 4 + 16;
    bool ok = cudaMalloc(&m->t_gelu, 1 << 17) == cudaSuccess && cudaMalloc(&m->t_exp, 1 << 17) == cudaSuccess &&
              cudaMalloc(&m->rope_cs, cs.size() * sizeof(double2) + 16) == cudaSuccess &&
              cudaMalloc(&m->mem_k, kv) == cudaSuccess && cudaMalloc(&m->mem_v, kv) == cudaSuccess &&
              cudaMalloc(&m->x, Nmax * Es * 4) == cudaSuccess && cudaMalloc(&m->cur, Nmax * Es * 4) == cudaSuccess &&
              cudaMalloc(&m->q, Nmax * Es * 4) == cudaSuccess && cudaMalloc(&m->k, Nmax * Es * 4) == cudaSuccess &&
              cudaMalloc(&m->v, Nmax * Es * 4) == cudaSuccess && cudaMalloc(&m->att, Nmax * Es * 4) == cudaSuccess &&
              cudaMalloc(&m->merged, Nmax * Es * 4) == cudaSuccess && cudaMalloc(&m->ff, Nmax * 4 * Es * 4) == cudaSuccess &&
              cudaMalloc(&m->ff2, Nmax * Es * 4) == cudaSuccess &&
              cudaMalloc(&m->kq, (size_t)m->n_head * Nmax * n_ctx * 4) == cudaSuccess &&
              cudaMalloc(&m->logits, Nmax * (size_t)m->n_vocab * 4 + 16) == cudaSuccess && cudaMalloc(&m->ids, Nmax * 4) == cudaSuccess;
    ok = ok && cudaMemset(m->mem_k, 0, kv) == cudaSuccess && cudaMemset(m->mem_v, 0, kv) == cudaSuccess &&
         cudaMemcpy(m->t_gelu, tg.data(), 1 << 17, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(m->t_exp, te.data(), 1 << 17, cudaMemcpyHostToDevice) == cudaSuccess &&
         (cs.empty() || cudaMemcpy(m->rope_cs, cs.data(), cs.size() * sizeof(double2), cudaMemcpyHostToDevice) == cudaSuccess);
    if (!ok) return bail(std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    const char *pe = getenv("B200NEOX_PROGRAM");
    if (pe && atoi(pe) == 0) m->prog_why = "disabled (B200NEOX_PROGRAM=0)";
    else build_decode_program(m);                        // failure is not fatal: the per-op kernels serve decode steps too
    return m;
}

b200neox *b200neox_load(const char *path, int n_ctx, int max_batch) { return neox_load_impl(path, n_ctx, max_batch, 0, 1, nullptr, 0); }

b200neox *b200neox_load_sharded(const char *path, int n_ctx, int rank, int world, void *const *pool_peers, size_t pool_bytes)
{
    b200neox *m = neox_load_impl(path, n_ctx, 1, rank, world, pool_peers, pool_bytes);
    if (m && world > 1 && !m->have_prog) {                // a sharded model has nothing but the decode program
        nfail(B200Q4_EINVAL, "neox_load_sharded: " + m->prog_why);
        b200neox_free(m);
        return nullptr;
    }
    return m;
}

size_t b200neox_sharded_pool_bytes(int n_vocab, int n_embd, int n_layer)
{
    if (n_vocab < 1 || n_embd < 32 || n_layer < 0) return 0;
    return pool_layout(n_embd, n_layer, n_vocab, ops_bound(n_layer)).total;
}

int b200neox_n_vocab(const b200neox *m) { return m ? m->n_vocab : B200Q4_EINVAL; }

/* 1 when decode steps run as one persistent launch; 0 (and why, through b200neox_last_error) when not */
int b200neox_decode_is_program(const b200neox *m)
{
    if (!m) return B200Q4_EINVAL;
    if (!m->have_prog) g_nerr = m->prog_why;
    return m->have_prog ? 1 : 0;
}

#ifdef B200Q4_TIMELINE
/* development aid (timeline builds, B200Q4_DBG=1): the decode program's [grid][n_ops][8] stamps of the last step, and the
 * kind of every op (0 mul_mat, 1 attention, 2 sum); returns the grid size, n_ops through *n_ops */
int b200neox_debug_timeline(const b200neox *m, unsigned long long *out, size_t cap_words, int *n_ops, int *kinds, int cap_kinds)
{
    if (!m || !m->have_prog || !m->prog.d_dbg) return B200Q4_EINVAL;
    const size_t n = (size_t)m->prog.grid * m->prog.n_ops * 8 + (size_t)m->prog.grid * 16;
    if (cap_words < n || cap_kinds < m->prog.n_ops) return B200Q4_EINVAL;
    cudaDeviceSynchronize();
    std::vector<b200q4::PgOp> ops((size_t)m->prog.n_ops);
    if (cudaMemcpy(ops.data(), m->prog.d_ops, ops.size() * sizeof(b200q4::PgOp), cudaMemcpyDeviceToHost) != cudaSuccess) return B200Q4_ECUDA;
    for (int i = 0; i < m->prog.n_ops; i++) kinds[i] = ops[(size_t)i].kind | (ops[(size_t)i].N << 4);
    *n_ops = m->prog.n_ops;
    return cudaMemcpy(out, m->prog.d_dbg, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? m->prog.grid : B200Q4_ECUDA;
}
#endif

int b200neox_eval(b200neox *m, const int *tokens, int N, int n_past, int n_threads_emulated, float *logits_out)
{
    if (!m || !tokens || !logits_out || N < 1 || N > m->max_n || n_past < 0 || n_past + N > m->n_ctx || n_threads_emulated < 1)
        return nfail(B200Q4_EINVAL, "neox_eval: bad argument");
    for (int i = 0; i < N; i++)
        if (tokens[i] < 0 || tokens[i] >= m->n_vocab) return nfail(B200Q4_EINVAL, "neox_eval: token id out of range");
    const int E = m->n_embd, H = m->n_head, d = E / H, P = n_past + N, nth = n_threads_emulated;
    cudaStream_t st = m->st;
    const bool exact = b200q4_exact_enabled() != 0;                                          // exact mode covers the graph ops too
    if (m->world > 1 && (N != 1 || exact || !m->have_prog))
        return nfail(B200Q4_EINVAL, "neox_eval: a sharded model runs decode steps (N = 1, exact mode off) only");
    if (N == 1 && !exact && m->have_prog) {
        // one token = one persistent launch: ids in, logits out
        b200q4::ApiLock lk;
        const int dyn[2] = {tokens[0], n_past};
        NCK(cudaMemcpyAsync(m->prog.d_dyn, dyn, sizeof dyn, cudaMemcpyHostToDevice, st));
        if (b200q4::prog_launch(m->prog, st)) return nfail(B200Q4_ECUDA, std::string("neox_eval: ") + b200q4_last_error());
        NCK(cudaMemcpyAsync(logits_out, m->logits, (size_t)m->n_vocab * 4, cudaMemcpyDeviceToHost, st));
        NCK(cudaStreamSynchronize(st));
        return B200Q4_OK;
    }
    auto mm = [&](b200q4_weight *w, const float *x, float *y) -> int {                       // ggml_mul_mat on a Q4 weight
        if (int rc = b200q4_mul_mat_dev(w, x, N, 0, y, 0, st)) return nfail(rc, std::string("neox_eval: ") + b200q4_last_error());
        return 0;
    };
    auto bias = [&](float *a, const float *b, int n) { k_add_bias<<<nblk((long)N * n), 256, 0, st>>>(a, b, (long)N * n, n); };
    auto ln = [&](const float *x, const float *w, const float *b, float *y) {
        if (exact) k_layernorm<<<nblk(N, 32), 32, 0, st>>>(x, N, E, w, b, y);               // the reference's own summation order
        else       k_layernorm_fast<<<N, 256, 0, st>>>(x, E, w, b, y);
    };
    NCK(cudaMemcpyAsync(m->ids, tokens, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    k_get_rows_q4<<<nblk((long)N * E), 256, 0, st>>>(m->wte_type, m->d_wte, E, m->ids, N, m->x);
    const float scale = (float)(1.0 / sqrt((double)((float)E / (float)H)));                 // vsim.cpp:590
    int rc;
    for (int l = 0; l < m->n_layer; l++) {
        const NeoxLayer &ly = m->layers[(size_t)l];
        float *mk = m->mem_k + (size_t)l * m->n_ctx * E, *mv = m->mem_v + (size_t)l * m->n_ctx * E;
        ln(m->x, ly.ln1_w, ly.ln1_b, m->cur);
        if ((rc = mm(ly.wq, m->cur, m->q))) return rc;
        bias(m->q, ly.bq, E);
        if ((rc = mm(ly.wk, m->cur, m->k))) return rc;
        bias(m->k, ly.bk, E);
        if ((rc = mm(ly.wv, m->cur, m->v))) return rc;
        bias(m->v, ly.bv, E);
        k_copy<<<nblk((long)N * E), 256, 0, st>>>(m->k, mk + (size_t)n_past * E, (long)N * E);          // cpy into the KV memory (:557-558)
        k_copy<<<nblk((long)N * E), 256, 0, st>>>(m->v, mv + (size_t)n_past * E, (long)N * E);
        if (m->n_rot > 0) {
            k_rope<<<nblk((long)N * H * (m->n_rot / 2)), 256, 0, st>>>(m->q, d, H, N, n_past, m->n_rot, 0, m->rope_cs);
            k_rope<<<nblk((long)N * H * (m->n_rot / 2)), 256, 0, st>>>(mk, d, H, P, n_past, m->n_rot, 1, m->rope_cs);   // in place on the memory view
        }
        k_kq<<<nblk((long)H * N * P), 256, 0, st>>>(mk, m->q, H, N, P, d, n_past, scale, m->kq);
        k_soft_max<<<nblk((long)H * N, 32), 32, 0, st>>>(m->kq, H * N, P, m->t_exp);
        k_kqv<<<nblk((long)H * N * d), 256, 0, st>>>(mv, m->kq, H, N, P, d, nth, m->merged);
        if ((rc = mm(ly.wo, m->merged, m->att))) return rc;
        bias(m->att, ly.bo, E);
        if (m->par_res == 0) {
            // vsim.cpp:636-665: inpFF = norm(cur + inpL) -> MLP;  inpL = inpFF + inpL
            k_add<<<nblk((long)N * E), 256, 0, st>>>(m->att, m->x, m->att, (long)N * E);
            ln(m->att, ly.ln2_w, ly.ln2_b, m->cur);
        } else {
            ln(m->x, ly.ln2_w, ly.ln2_b, m->cur);
        }
        if ((rc = mm(ly.w1, m->cur, m->ff))) return rc;
        bias(m->ff, ly.b1, 4 * E);
        k_gelu<<<nblk((long)N * 4 * E), 256, 0, st>>>(m->ff, (long)N * 4 * E, m->t_gelu);
        if ((rc = mm(ly.w2, m->ff, m->ff2))) return rc;
        bias(m->ff2, ly.b2, E);
        if (m->par_res == 0) {
            k_add<<<nblk((long)N * E), 256, 0, st>>>(m->ff2, m->x, m->x, (long)N * E);                   // inpL = inpFF + inpL
        } else {
            k_add<<<nblk((long)N * E), 256, 0, st>>>(m->att, m->ff2, m->ff2, (long)N * E);               // inpFF = cur + inpFF
            k_add<<<nblk((long)N * E), 256, 0, st>>>(m->x, m->ff2, m->x, (long)N * E);                   // inpL = inpL + inpFF
        }
    }
    ln(m->x, m->lnf_w, m->lnf_b, m->cur);
    if ((rc = mm(m->w_head, m->cur, m->logits))) return rc;
    NCK(cudaGetLastError());
    NCK(cudaMemcpyAsync(logits_out, m->logits + (size_t)(N - 1) * m->n_vocab, (size_t)m->n_vocab * 4, cudaMemcpyDeviceToHost, st));
    NCK(cudaStreamSynchronize(st));
    return B200Q4_OK;
}

}  // extern "C"
