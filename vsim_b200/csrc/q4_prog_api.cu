// q4_prog_api.cu -- host side of the decode program kernel (q4_prog.cuh) and the b200q4_chain_* entry points.
#include "../../include/b200q4.h"
#include "q4_prog_host.h"

#include <cstdio>
#include <cstring>

namespace b200q4 {

#define PCK(call)                                                                                      \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return api_fail(B200Q4_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

PgGeo pg_geometry(const b200q4_weight *w)
{
    PgGeo g{};
    const int C = w->C, n_tiles = w->n_tiles;
    if (C > 64) {                                        // K > 8192: slice, one stage or less per piece
        int PP = (C + 31) / 32;
        int TS = (C + PP - 1) / PP;
        TS = (TS + 1) & ~1;                              // even: unit pairs stay aligned to quant-block octets
        PP = (C + TS - 1) / TS;
        g.PP = PP; g.TS = TS; g.ntasks = n_tiles * PP;
    } else if (C >= PIPE_TILE_SPLIT) {                   // one task = one tile
        g.PP = 1; g.TS = C; g.ntasks = n_tiles;
    } else {                                             // short tiles: a task = TS / C whole tiles
        const int G = 32 / C > 0 ? 32 / C : 1;
        g.PP = 1; g.TS = G * C;
        g.ntasks = (int)(((long)n_tiles * C + g.TS - 1) / g.TS);
    }
    return g;
}

int ProgBuild::add_matvec(const b200q4_weight *w, int dep, int in_mode, int res, const PgSrc *src, int nsrc,
                          const float *ln_w, const float *ln_b, const uint16_t *gelu_tab, float *y, const float *bias, int y_slice)
{
    if (!w || w->transposed || !w->d_wt || w->N < 1 || w->K < QK) return api_fail(B200Q4_EINVAL, "program: bad weight");
    if (qt < 0) qt = w->type;
    if (w->type != qt) return api_fail(B200Q4_EINVAL, "program: every mul_mat of a program must have the same quant type");
    if (nsrc < 0 || nsrc > PG_MAX_SRC || dep >= (int)ops.size() || !y) return api_fail(B200Q4_EINVAL, "program: bad op");
    if ((in_mode != PG_IN_LN || res != PG_RES_KEEP) && nsrc < 1) return api_fail(B200Q4_EINVAL, "program: op without input");
    if ((long)w->n_tiles * w->C > 0x7fffffffL) return api_fail(B200Q4_EINVAL, "program: tensor too large");
    const PgGeo g = pg_geometry(w);
    if (g.PP > 1 && bias) return api_fail(B200Q4_EINVAL, "program: a K-sliced op cannot add a bias");
    const DevInfo di = dev_info();
    const int grid = di.sms > 0 ? di.sms : 148;
    if (g.PP > grid) return api_fail(B200Q4_EINVAL, "program: K=%d too large (more K-slices than SMs)", w->K);
    PgOp op;
    memset(&op, 0, sizeof op);
    op.kind = PG_MATVEC; op.dep = dep;
    op.wt = w->d_wt; op.C = w->C; op.n_tiles = w->n_tiles; op.N = w->N; op.nb = w->nb; op.U = w->n_tiles * w->C;
    op.PP = g.PP; op.TS = g.TS; op.ntasks = g.ntasks;
    // static tasks: all but (at least) the last task of a CTA; small ops are dealt out by tickets alone so that
    // whoever is ahead takes more
    const int ncls = g.PP > 1 ? grid / g.PP : grid;
    const int per = (g.PP > 1 ? w->n_tiles : g.ntasks) / (ncls > 0 ? ncls : 1);
    op.nstatic = per - 1 < 0 ? 0 : per - 1 > PIPE_STATIC_TASKS ? PIPE_STATIC_TASKS : per - 1;
    op.tk_off = n_tk; n_tk += g.PP;
    op.in_mode = in_mode; op.res = res; op.nsrc = nsrc; op.K = w->K;
    for (int i = 0; i < nsrc; i++) op.src[i] = src[i];
    op.ln_w = ln_w; op.ln_b = ln_b; op.gelu_tab = gelu_tab;
    op.y = y; op.bias = bias; op.y_slice = y_slice;
    const int nbp = g.PP > 1 ? g.TS * CHUNK_BLKS : ((w->C * CHUNK_BLKS + 7) & ~7);
    if (nbp > nbp_cap) nbp_cap = nbp;
    if (in_mode == PG_IN_LN && w->K > e_cap) e_cap = w->K;
    ops.push_back(op);
    return (int)ops.size() - 1;
}

int ProgBuild::add_attn(int dep, const float *q, const float *k, const float *v, float *mem_k, float *mem_v, float *att_out,
                        float *att_part, const double2 *rope, int H, int d, int n_rot, int n_ctx, float scale)
{
    if (H < 1 || d < 1 || d > PG_ATT_MAX_D || n_rot < 0 || n_rot % 2 || n_rot > d || n_ctx < 1 || dep >= (int)ops.size())
        return api_fail(B200Q4_EINVAL, "program: bad attention op (H=%d d=%d n_rot=%d)", H, d, n_rot);
    PgOp op;
    memset(&op, 0, sizeof op);
    op.kind = PG_ATTN; op.dep = dep;
    op.q = q; op.k = k; op.v = v; op.mem_k = mem_k; op.mem_v = mem_v; op.att_out = att_out; op.att_part = att_part; op.rope = rope;
    op.H = H; op.d = d; op.n_rot = n_rot; op.max_chunks = (n_ctx + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK;
    op.cnt_off = n_cnt; n_cnt += H;
    op.att_scale = scale;
    const int need = 3 * ((d + 3) & ~3) + PG_ATT_CHUNK + 8 + PG_CTHREADS;
    if (need > att_cap) att_cap = need;
    ops.push_back(op);
    return (int)ops.size() - 1;
}

int ProgBuild::add_sum(int dep, const PgSrc &src, float *y, int N)
{
    if (!y || N < 1 || dep >= (int)ops.size()) return api_fail(B200Q4_EINVAL, "program: bad sum op");
    PgOp op;
    memset(&op, 0, sizeof op);
    op.kind = PG_SUM; op.dep = dep; op.N = N; op.y = y; op.nsrc = 1; op.src[0] = src;
    ops.push_back(op);
    return (int)ops.size() - 1;
}

template <int QT>
static int set_attr(int smem_optin)
{
    PCK(cudaFuncSetAttribute(q4_prog_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
    return B200Q4_OK;
}

int prog_finalize(const ProgBuild &b, Prog &out)
{
    const DevInfo di = dev_info();
    if (di.device < 0) return api_fail(B200Q4_ENODEV, "b200q4_init() has not succeeded; no CPU fallback exists");
    if (b.ops.empty() || b.qt < 0) return api_fail(B200Q4_EINVAL, "program: no mul_mat op");
    Prog pg;
    pg.qt = b.qt; pg.grid = di.sms; pg.n_ops = (int)b.ops.size();
    pg.e_cap = (b.e_cap + 3) & ~3; pg.nbp_cap = (b.nbp_cap + 7) & ~7; pg.att_cap = (b.att_cap + 3) & ~3;
    const size_t fixed = pg_fixed_bytes(pg.e_cap, pg.nbp_cap, pg.att_cap), stage = pg_stage_bytes(pg.qt);
    if (fixed + 2 * stage > (size_t)di.smem_optin)
        return api_fail(B200Q4_EINVAL, "program: %zu bytes of per-CTA state leave no room for two weight stages", fixed);
    long st = (long)(((size_t)di.smem_optin - fixed) / stage);
    pg.stages = (int)(st > PG_MAX_STAGES ? PG_MAX_STAGES : st);
    pg.smem = fixed + (size_t)pg.stages * stage;
    for (const PgOp &op : b.ops) if (op.kind == PG_MATVEC) pg.n_matvec++;
    // counters: [0] exit, [1 + o] arrivals, tickets, attention head counters
    std::vector<PgOp> ops = b.ops;
    const int tk0 = 1 + pg.n_ops, cnt0 = tk0 + b.n_tk;
    for (PgOp &op : ops) { if (op.kind == PG_MATVEC) op.tk_off += tk0; if (op.kind == PG_ATTN) op.cnt_off += cnt0; }
    pg.n_ctr = cnt0 + b.n_cnt;
    if (int rc = (pg.qt == 0 ? set_attr<0>(di.smem_optin) : set_attr<1>(di.smem_optin))) return rc;
    PCK(cudaMalloc(&pg.d_ops, ops.size() * sizeof(PgOp)));
    if (cudaMalloc(&pg.d_ctr, (size_t)pg.n_ctr * sizeof(int)) != cudaSuccess || cudaMalloc(&pg.d_dyn, 16) != cudaSuccess ||
        cudaMalloc(&pg.d_err, 16) != cudaSuccess) {
        cudaGetLastError();
        prog_free(pg);
        return api_fail(B200Q4_ENOMEM, "program: device allocation failed");
    }
    cudaError_t e = cudaMemcpy(pg.d_ops, ops.data(), ops.size() * sizeof(PgOp), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(pg.d_ctr, 0, (size_t)pg.n_ctr * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(pg.d_dyn, 0, 16);
    if (e == cudaSuccess) e = cudaMemset(pg.d_err, 0, 16);
    if (e != cudaSuccess) { prog_free(pg); return api_fail(B200Q4_ECUDA, "program: upload failed: %s", cudaGetErrorString(e)); }
    out = pg;
    return B200Q4_OK;
}

int prog_launch(const Prog &pg, cudaStream_t st)
{
    PgParams p;
    p.ops = pg.d_ops; p.n_ops = pg.n_ops; p.ctr = pg.d_ctr; p.n_ctr = pg.n_ctr; p.dyn = pg.d_dyn; p.err = pg.d_err;
    p.stages = pg.stages; p.e_cap = pg.e_cap; p.nbp_cap = pg.nbp_cap; p.att_cap = pg.att_cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pg.grid); cfg.blockDim = dim3(PG_THREADS); cfg.dynamicSmemBytes = pg.smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;           // the CTAs wait on one another: all of them must be resident
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (pg.qt == 0) PCK(cudaLaunchKernelEx(&cfg, q4_prog_kernel<0>, p));
    else            PCK(cudaLaunchKernelEx(&cfg, q4_prog_kernel<1>, p));
    api_count_launch(1);
    return B200Q4_OK;
}

void prog_free(Prog &pg)
{
    cudaFree(pg.d_ops); cudaFree(pg.d_ctr); cudaFree(pg.d_dyn); cudaFree(pg.d_err);
    pg = Prog{};
}

} // namespace b200q4

using namespace b200q4;

// ---- b200q4_chain_*: n dependent M = 1 mul_mats in one persistent launch ---------------------------------------
struct b200q4_chain {
    Prog pg;
    std::vector<float *> partials;       // K-slice workspaces of sliced ops
};

extern "C" {

b200q4_chain *b200q4_chain_create(int n, const b200q4_weight *const *ws, const float *const *d_x, float *const *d_y, const int *deps)
{
    ApiLock lk;
    if (dev_info().device < 0) { api_fail(B200Q4_ENODEV, "b200q4_init() has not succeeded; no CPU fallback exists"); return nullptr; }
    if (n < 1 || !ws || !d_x || !d_y) { api_fail(B200Q4_EINVAL, "chain_create: bad argument"); return nullptr; }
    b200q4_chain *c = new b200q4_chain();
    ProgBuild b;
    struct Out { const float *y; int N, PP, stride; float *part; int op; };
    std::vector<Out> outs((size_t)n);
    auto bail = [&]() { for (float *p : c->partials) cudaFree(p); delete c; return (b200q4_chain *)nullptr; };
    for (int i = 0; i < n; i++) {
        const b200q4_weight *w = ws[i];
        if (!w || w->transposed || !d_x[i] || !d_y[i] || (deps && (deps[i] >= i || deps[i] < -1))) {
            api_fail(B200Q4_EINVAL, "chain_create: bad op %d", i);
            return bail();
        }
        if ((reinterpret_cast<uintptr_t>(d_x[i]) | reinterpret_cast<uintptr_t>(d_y[i])) & 15) {
            api_fail(B200Q4_EINVAL, "chain_create: op %d: x and y must be 16-byte aligned", i);
            return bail();
        }
        // the input: caller memory, or -- when it lies inside the output of an earlier K-sliced op -- that op's partials
        PgSrc src = pg_src_f32(d_x[i]);
        for (int j = 0; j < i; j++)
            if (outs[(size_t)j].PP > 1 && d_x[i] >= outs[(size_t)j].y && d_x[i] + w->K <= outs[(size_t)j].y + outs[(size_t)j].N)
                src = pg_src_f32(outs[(size_t)j].part + (d_x[i] - outs[(size_t)j].y), outs[(size_t)j].PP, outs[(size_t)j].stride);
        const PgGeo g = pg_geometry(w);
        Out o{d_y[i], w->N, g.PP, 0, nullptr, -1};
        float *y = d_y[i];
        if (g.PP > 1) {
            o.stride = (w->N + 3) & ~3;
            if (cudaMalloc(&o.part, (size_t)g.PP * o.stride * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                api_fail(B200Q4_ENOMEM, "chain_create: cudaMalloc failed");
                return bail();
            }
            c->partials.push_back(o.part);
            y = o.part;
        }
        const int dep = deps && deps[i] >= 0 ? outs[(size_t)deps[i]].op : -1;
        o.op = b.add_matvec(w, dep, PG_IN_PLAIN, PG_RES_KEEP, &src, 1, nullptr, nullptr, nullptr, y, nullptr, o.stride);
        if (o.op < 0) return bail();
        outs[(size_t)i] = o;
    }
    // the finished vectors of K-sliced ops, summed at the end of the list (nobody inside the list waits for them)
    for (int i = 0; i < n; i++)
        if (outs[(size_t)i].PP > 1 &&
            b.add_sum(outs[(size_t)i].op, pg_src_f32(outs[(size_t)i].part, outs[(size_t)i].PP, outs[(size_t)i].stride), d_y[i], outs[(size_t)i].N) < 0)
            return bail();
    if (prog_finalize(b, c->pg)) return bail();
    return c;
}

int b200q4_chain_run(b200q4_chain *c, void *stream)
{
    ApiLock lk;
    if (!c) return api_fail(B200Q4_EINVAL, "chain_run: NULL chain");
    return prog_launch(c->pg, stream ? (cudaStream_t)stream : dev_info().stream);
}

int b200q4_chain_info(const b200q4_chain *c, int *n_ops, int *stages, size_t *smem_bytes)
{
    if (!c) return B200Q4_EINVAL;
    if (n_ops) *n_ops = c->pg.n_ops;
    if (stages) *stages = c->pg.stages;
    if (smem_bytes) *smem_bytes = c->pg.smem;
    return B200Q4_OK;
}

void b200q4_chain_free(b200q4_chain *c)
{
    if (!c) return;
    ApiLock lk;
    cudaDeviceSynchronize();
    for (float *p : c->partials) cudaFree(p);
    prog_free(c->pg);
    delete c;
}

} // extern "C"
