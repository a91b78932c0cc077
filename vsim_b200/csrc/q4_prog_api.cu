// q4_prog_api.cu -- host side of the decode program kernel (q4_prog.cuh) and the b200q4_chain_* entry points.
#include "../../include/b200q4.h"
#include "q4_prog_host.h"

#include <cstdio>
#include <cstdlib>
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

PgGeo pg_geometry_dims(int N, int K)
{
    b200q4_weight w{};
    w.N = N; w.K = K; w.nb = K / QK;
    w.C = (w.nb + CHUNK_BLKS - 1) / CHUNK_BLKS;
    w.n_tiles = (N + TILE_ROWS - 1) / TILE_ROWS;
    return pg_geometry(&w);
}

int ProgBuild::add_matvec(const b200q4_weight *w, int dep, int in_mode, int res, const PgSrc *src, int nsrc,
                          const float *ln_w, const float *ln_b, const uint16_t *gelu_tab, float *y, const float *bias, int y_slice,
                          bool y_tag, bool same_pack, float *y_plain)
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
    // Scheduling.  Small ops (<= 3 tasks per CTA) are dealt out statically -- no atomics on their short critical path --
    // with a rotation that moves the CTAs holding one task more from op to op; large ops take their first tasks
    // statically and the tail from a ticket counter, so that whoever is ahead takes more.
    const int ncls = g.PP > 1 ? grid / g.PP : grid;      // CTAs per K-slice class (at least)
    const int limit = g.PP > 1 ? w->n_tiles : g.ntasks;
    const int per_up = (limit + ncls - 1) / ncls;
    if (per_up <= PIPE_STATIC_TASKS) {
        op.nstatic = per_up; op.dynamic = 0;
        op.rot = rot_acc % ncls;
        rot_acc += limit % ncls;
    } else {
        op.nstatic = PIPE_STATIC_TASKS < limit / ncls - 1 ? PIPE_STATIC_TASKS : limit / ncls - 1;
        op.dynamic = 1; op.rot = 0;
    }
    op.tk_off = n_tk; n_tk += g.PP;
    op.in_mode = in_mode; op.res = res; op.nsrc = nsrc; op.K = w->K;
    for (int i = 0; i < nsrc; i++) op.src[i] = src[i];
    op.ln_w = ln_w; op.ln_b = ln_b; op.gelu_tab = gelu_tab;
    op.y = y; op.bias = bias; op.y_slice = y_slice;
    op.dep_mv = dep >= 0 && ops[(size_t)dep].kind == PG_MATVEC;
    op.y_tag = y_tag ? 1 : 0;
    op.y_plain = g.PP == 1 ? y_plain : nullptr;
    if (ops.size() >= 4095) return api_fail(B200Q4_EINVAL, "program: too many ops");
    op.pack_of = (int)ops.size();
    op.pack_seq = n_packs;
    // who builds the pack (q4_prog.cuh): the consumers, when it is a residual update (they own the residual stream), a
    // layer-norm over a temporary sum, or gelu; the prep warps otherwise, ahead of time
    const bool res_upd = in_mode == PG_IN_LN && res != PG_RES_KEEP;
    op.builder = (PG_PWARPS == 0 || res_upd || in_mode == PG_IN_GELU) ? 1 : 0;
    op.cobuild = -1;
    op.wait_pack = -1;
    if (in_mode == PG_IN_LN && res == PG_RES_KEEP) {
        if (last_res_pack < 0) return api_fail(B200Q4_EINVAL, "program: layer-norm op before any residual update");
        op.wait_pack = last_res_pack;
    }
    if (in_mode == PG_IN_LN && (res == PG_RES_ADD || res == PG_RES_SET)) last_res_pack = n_packs;
    if (same_pack) {
        if (last_mv < 0) return api_fail(B200Q4_EINVAL, "program: same_pack without a previous mul_mat");
        const PgOp &pv = ops[(size_t)last_mv];
        if (pv.K != op.K || pv.PP != op.PP || pv.TS != op.TS || pv.C != op.C || res != PG_RES_KEEP)
            return api_fail(B200Q4_EINVAL, "program: same_pack ops must share their input geometry");
        op.pack_of = pv.pack_of;
        op.pack_seq = pv.pack_seq;
        op.builder = pv.builder;
        op.wait_pack = pv.wait_pack;
    } else {
        // a layer-norm pack that only re-reads the residual stream can be built in the pause of the pack that updated it,
        // if it is the very next pack (the two buffers alternate)
        bool cobuilt = false;
        if (PG_PWARPS == 0 && in_mode == PG_IN_LN && res == PG_RES_KEEP && last_res_pack == n_packs - 1 && last_res_op >= 0 &&
            ops[(size_t)last_res_op].early == 0) {
            ops[(size_t)last_res_op].cobuild = (int)ops.size();
            cobuilt = true;
        }
        // A pack made of one plain vector that nothing in the list writes needs no pause at all: every consumer warp
        // prepares its share between two stages of the previous pack's first op (q4_prog.cuh, pg_consume_op).  That op
        // must start with a barrier of its own (built there, or itself prepared ahead), and one item per thread must do.
        const int nbp_new = g.PP > 1 ? g.TS * CHUNK_BLKS : ((w->C * CHUNK_BLKS + 7) & ~7);
        const bool selfc = (g.PP == 1 && w->C == PG_UNITS) || (g.PP > 1 && g.TS == PG_UNITS && w->C % g.TS == 0);
        if (PG_PWARPS == 0 && in_mode == PG_IN_PLAIN && nsrc == 1 && (src[0].kind & 255) == PG_SRC_F32 && src[0].np == 1 &&
            src[0].scale == 1.0f && dep < 0 && last_pack_op >= 0 && !last_pack_cobuilt && ops[(size_t)last_pack_op].cobuild < 0 &&
            ops[(size_t)last_pack_op].early == 0 && nbp_new * 4 <= PG_CTHREADS && !getenv("B200Q4_PROG_NO_EARLY")) {
            PgOp &host = ops[(size_t)last_pack_op];
            host.early = (int)ops.size();
            host.early_x = static_cast<const float *>(src[0].ptr);
            host.e_C = w->C; host.e_PP = g.PP; host.e_TS = g.TS; host.e_nb = w->nb; host.e_seq = n_packs;
            op.prebuilt = 1;
            // no barrier at this op's start if nobody reads what another warp wrote: this pack is self-contained, and the
            // buffer it went into was free for sure -- the host op started with a barrier, or without one but then the
            // pack before it (the buffer's previous content) was self-contained as well
            const bool host_bar = host.prebuilt != 2;
            const bool prev_selfc = n_packs >= 2 && pack_selfc[(size_t)n_packs - 2];
            if (selfc && (host_bar || prev_selfc) && !getenv("B200Q4_PROG_EARLY_BARRIER")) op.prebuilt = 2;
        }
        last_pack_cobuilt = cobuilt;
        pack_selfc.push_back(selfc ? 1 : 0);
        last_pack_op = (int)ops.size();
        n_packs++;
    }
    if (in_mode == PG_IN_LN && (res == PG_RES_ADD || res == PG_RES_SET)) last_res_op = (int)ops.size();
    last_mv = (int)ops.size();
    const int nbp = g.PP > 1 ? g.TS * CHUNK_BLKS : ((w->C * CHUNK_BLKS + 7) & ~7);
    if (nbp > nbp_cap) nbp_cap = nbp;
    if (in_mode == PG_IN_LN && w->K > e_cap) e_cap = w->K;
    ops.push_back(op);
    return (int)ops.size() - 1;
}

int ProgBuild::add_attn(int dep, const float *q, const float *k, const float *v, int tq, int tk, int tv, float *mem_k, float *mem_v, float *att_out,
                        float *att_part, float *att_sc, double *att_sum, const double2 *rope, const uint16_t *exp_tab,
                        int H, int d, int n_rot, int n_ctx, float scale)
{
    if (H < 1 || d < 1 || d > PG_ATT_MAX_D || n_rot < 0 || n_rot % 2 || n_rot > d || n_ctx < 1 || dep >= (int)ops.size())
        return api_fail(B200Q4_EINVAL, "program: bad attention op (H=%d d=%d n_rot=%d)", H, d, n_rot);
    PgOp op;
    memset(&op, 0, sizeof op);
    op.kind = PG_ATTN; op.dep = dep;
    op.q = q; op.k = k; op.v = v; op.mem_k = mem_k; op.mem_v = mem_v; op.att_out = att_out; op.att_part = att_part; op.rope = rope;
    op.att_sc = att_sc; op.att_sum = att_sum; op.exp_tab = exp_tab; op.n_ctx = n_ctx;
    op.tq = tq; op.tk = tk; op.tv = tv;
    op.H = H; op.d = d; op.n_rot = n_rot; op.max_chunks = (n_ctx + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK;
    op.cnt_off = n_cnt; n_cnt += 2 * H;
    op.att_scale = scale;
    // q, new k, new v rows; the chunk's scores (a chunk is at most the whole context); the position groups' partial outputs
    const int groups = d % 4 == 0 ? PG_CTHREADS / (d / 4 <= 8 ? 8 : d / 4 <= 16 ? 16 : d / 4 <= 32 ? 32 : d / 4 <= 64 ? 64 : 128) : PG_CTHREADS / ((d + 31) & ~31);
    const int need = 3 * ((d + 3) & ~3) + (n_ctx + PG_ATT_CHUNK - 1) / PG_ATT_CHUNK * PG_ATT_CHUNK + 8 + groups * ((d + 31) & ~31) + 32;
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

// dynamic shared memory the kernel may use: the opt-in limit minus its static shared memory (the parameter copy)
template <int QT>
static int set_attr(int smem_optin, int *avail)
{
    cudaFuncAttributes fa;
    PCK(cudaFuncGetAttributes(&fa, q4_prog_kernel<QT>));
    *avail = smem_optin - (int)fa.sharedSizeBytes;
    PCK(cudaFuncSetAttribute(q4_prog_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, *avail));
    return B200Q4_OK;
}

int prog_counter_ints(const ProgBuild &b) { return 1 + (int)b.ops.size() + b.n_tk + b.n_cnt; }

int prog_finalize(const ProgBuild &b, Prog &out) { return prog_finalize_sharded(b, out, 0, 1, nullptr); }

int prog_finalize_sharded(const ProgBuild &b, Prog &out, int rank, int world, int *const *ctr_peers)
{
    const DevInfo di = dev_info();
    if (di.device < 0) return api_fail(B200Q4_ENODEV, "b200q4_init() has not succeeded; no CPU fallback exists");
    if (b.ops.empty() || b.qt < 0) return api_fail(B200Q4_EINVAL, "program: no mul_mat op");
    if (world < 1 || world > 8 || rank < 0 || rank >= world || (world > 1 && !ctr_peers)) return api_fail(B200Q4_EINVAL, "program: bad rank / world");
    Prog pg;
    pg.world = world; pg.rank = rank;
    pg.qt = b.qt; pg.grid = di.sms; pg.n_ops = (int)b.ops.size();
    pg.e_cap = (b.e_cap + 3) & ~3; pg.nbp_cap = (b.nbp_cap + 7) & ~7; pg.att_cap = (b.att_cap + 3) & ~3;
    int avail = 0;
    if (int rc = (pg.qt == 0 ? set_attr<0>(di.smem_optin, &avail) : set_attr<1>(di.smem_optin, &avail))) return rc;
    const size_t fixed = pg_fixed_bytes(pg.e_cap, pg.nbp_cap, pg.att_cap), stage = pg_stage_bytes(pg.qt);
    if (fixed + 2 * stage > (size_t)avail)
        return api_fail(B200Q4_EINVAL, "program: %zu bytes of per-CTA state leave no room for two weight stages", fixed);
    long st = (long)(((size_t)avail - fixed) / stage);
    pg.stages = (int)(st > PG_MAX_STAGES ? PG_MAX_STAGES : st);
    if (const char *e = getenv("B200Q4_PROG_STAGES")) {  // development: a shallower ring leaves the SM some L1
        const int v = atoi(e);
        if (v >= 2 && v < pg.stages) pg.stages = v;
    }
    pg.smem = fixed + (size_t)pg.stages * stage;
    for (const PgOp &op : b.ops) if (op.kind == PG_MATVEC) pg.n_matvec++;
    // counters: [0] exit, [1 + o] arrivals, tickets, attention head counters
    std::vector<PgOp> ops = b.ops;
    const int tk0 = 1 + pg.n_ops, cnt0 = tk0 + b.n_tk;
    for (PgOp &op : ops) { if (op.kind == PG_MATVEC) op.tk_off += tk0; if (op.kind == PG_ATTN) op.cnt_off += cnt0; }
    for (const PgOp &op : b.ops) if (op.dep >= 0) ops[(size_t)op.dep].publish = 1;     // only ops somebody waits for are counted
    int nxt = pg.n_ops;                                  // chain the mul_mat ops for the producer
    for (int i = pg.n_ops - 1; i >= 0; i--) {
        PgOp &op = ops[(size_t)i];
        op.next_mv = nxt;
        if (nxt < pg.n_ops) {
            op.nx_tk_off = ops[(size_t)nxt].tk_off; op.nx_PP = ops[(size_t)nxt].PP;
            op.nx_dynamic = ops[(size_t)nxt].hold_a > 0 ? 0 : ops[(size_t)nxt].dynamic;      // (a holding op draws its first ticket itself)
        }
        if (op.kind == PG_MATVEC) nxt = i;
    }
    pg.first_mv = nxt;
    pg.n_ctr = cnt0 + b.n_cnt;
    PCK(cudaMalloc(&pg.d_ops, ops.size() * sizeof(PgOp)));
    if (world > 1) {
        pg.own_ctr = false;
        for (int r = 0; r < world; r++) pg.ctr_peer[r] = ctr_peers[r];
        pg.d_ctr = ctr_peers[rank];
    }
    if ((world == 1 && cudaMalloc(&pg.d_ctr, (size_t)pg.n_ctr * sizeof(int)) != cudaSuccess) || cudaMalloc(&pg.d_dyn, 16) != cudaSuccess ||
        cudaMalloc(&pg.d_err, 16) != cudaSuccess) {
        cudaGetLastError();
        prog_free(pg);
        return api_fail(B200Q4_ENOMEM, "program: device allocation failed");
    }
    for (PgOp &op : ops)                                  // embedding rows are selected by the token id in dyn[0]
        for (int i = 0; i < op.nsrc; i++)
            if (op.src[i].kind == PG_SRC_Q4ROW && !op.src[i].row) op.src[i].row = pg.d_dyn;
    cudaError_t e = cudaMemcpy(pg.d_ops, ops.data(), ops.size() * sizeof(PgOp), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(pg.d_ctr, 0, (size_t)pg.n_ctr * sizeof(int));
    const int dyn0[4] = {0, 0, 1, 0};                    // epoch 1: zeroed buffers never carry a valid tag
    if (e == cudaSuccess) e = cudaMemcpy(pg.d_dyn, dyn0, 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(pg.d_err, 0, 16);
    if (e == cudaSuccess && getenv("B200Q4_DBG") && atoi(getenv("B200Q4_DBG")) != 0) {    // timeline builds only (-DB200Q4_TIMELINE)
        const size_t n = ((size_t)pg.grid * pg.n_ops * 8 + (size_t)pg.grid * 16) * sizeof(unsigned long long);
        e = cudaMalloc(&pg.d_dbg, n);
        if (e == cudaSuccess) e = cudaMemset(pg.d_dbg, 0, n);
    }
    if (e != cudaSuccess) { prog_free(pg); return api_fail(B200Q4_ECUDA, "program: upload failed: %s", cudaGetErrorString(e)); }
    out = pg;
    return B200Q4_OK;
}

int prog_launch(const Prog &pg, cudaStream_t st)
{
    PgParams p;
    p.ops = pg.d_ops; p.n_ops = pg.n_ops; p.first_mv = pg.first_mv; p.dbg = pg.d_dbg; p.ctr = pg.d_ctr; p.n_ctr = pg.n_ctr; p.dyn = pg.d_dyn; p.err = pg.d_err;
    p.stages = pg.stages; p.e_cap = pg.e_cap; p.nbp_cap = pg.nbp_cap; p.att_cap = pg.att_cap;
    { const char *e = getenv("B200Q4_PROG_EVICT_FIRST"); p.wt_evict_first = e ? atoi(e) : 1; }
    p.world = pg.world; p.rank = pg.rank;
    for (int r = 0; r < 8; r++) p.ctr_peer[r] = pg.ctr_peer[r];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pg.grid); cfg.blockDim = dim3(PG_THREADS); cfg.dynamicSmemBytes = pg.smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;           // the CTAs wait on one another: all of them must be resident
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    { static const bool nocoop = getenv("B200Q4_PROG_NOCOOP") && atoi(getenv("B200Q4_PROG_NOCOOP")) != 0;   // development: what the cooperative launch costs
      if (nocoop) cfg.numAttrs = 0; }
    if (pg.qt == 0) PCK(cudaLaunchKernelEx(&cfg, q4_prog_kernel<0>, p));
    else            PCK(cudaLaunchKernelEx(&cfg, q4_prog_kernel<1>, p));
    api_count_launch(1);
    return B200Q4_OK;
}

void prog_free(Prog &pg)
{
    cudaFree(pg.d_ops); if (pg.own_ctr) cudaFree(pg.d_ctr);
    cudaFree(pg.d_dyn); cudaFree(pg.d_err); cudaFree(pg.d_dbg);
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
    // An op whose output a later op of the chain reads keeps it in an internal buffer of {value, tag} pairs (the reader
    // polls the data itself: no fence, no counter); the caller's plain d_y is filled from it at the end of the list.
    struct Out { const float *y; int N, PP, stride; float *buf; bool tagged; int op; };
    std::vector<Out> outs((size_t)n);
    auto bail = [&]() { for (float *p : c->partials) cudaFree(p); delete c; return (b200q4_chain *)nullptr; };
    auto reads = [&](int i, int j) {                      // does op i read from the output of op j?
        return d_x[i] >= d_y[j] && d_x[i] + ws[i]->K <= d_y[j] + ws[j]->N;
    };
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
    }
    for (int i = 0; i < n; i++) {
        const b200q4_weight *w = ws[i];
        int from = -1;
        for (int j = i - 1; j >= 0 && from < 0; j--) if (reads(i, j)) from = j;         // the latest writer of that memory
        PgSrc src = pg_src_f32(d_x[i]);
        if (from >= 0) {
            const Out &f = outs[(size_t)from];
            src = pg_src_tag(f.buf + 2 * (d_x[i] - f.y), f.op, f.PP, f.stride);
        }
        // a dependency that is not simply "reads its output" still goes through the arrival counters
        const int dep = deps && deps[i] >= 0 && deps[i] != from ? outs[(size_t)deps[i]].op : -1;
        const PgGeo g = pg_geometry(w);
        bool read_later = false;
        for (int k = i + 1; k < n && !read_later; k++) if (ws[k] && reads(k, i)) read_later = true;
        Out o{d_y[i], w->N, g.PP, 0, nullptr, read_later || g.PP > 1, -1};
        float *y = d_y[i];
        if (o.tagged) {
            o.stride = (w->N + 3) & ~3;
            if (cudaMalloc(&o.buf, (size_t)g.PP * o.stride * 2 * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                api_fail(B200Q4_ENOMEM, "chain_create: cudaMalloc failed");
                return bail();
            }
            cudaMemset(o.buf, 0, (size_t)g.PP * o.stride * 2 * sizeof(float));
            c->partials.push_back(o.buf);
            y = o.buf;
        }
        // (an unsliced op stores the caller's plain copy itself, next to the tagged one)
        // consecutive ops on the very same input (Q, K, V) share one activation pack
        bool same_pack = false;
        if (i > 0 && d_x[i] == d_x[i - 1] && dep < 0 && ws[i - 1]->K == w->K && (src.kind & 255) == PG_SRC_F32) {
            const PgGeo gp = pg_geometry(ws[i - 1]);
            same_pack = gp.PP == g.PP && gp.TS == g.TS && ws[i - 1]->C == w->C;
        }
        o.op = b.add_matvec(w, dep, PG_IN_PLAIN, PG_RES_KEEP, &src, 1, nullptr, nullptr, nullptr, y, nullptr, o.stride, o.tagged, same_pack,
                            o.tagged ? d_y[i] : nullptr);
        if (o.op < 0) return bail();
        outs[(size_t)i] = o;
    }
    // the caller's copies of the internally kept outputs, at the end of the list (nobody inside the list waits for them)
    for (int i = 0; i < n; i++) {
        const Out &o = outs[(size_t)i];
        if (o.tagged && o.PP > 1 && b.add_sum(-1, pg_src_tag(o.buf, o.op, o.PP, o.stride), d_y[i], o.N) < 0) return bail();
    }
    if (prog_finalize(b, c->pg)) return bail();
    return c;
}

/* Row-sharded list (one process per GPU): op i is this rank's row shard of a tensor with world * N_i rows; its outputs go
 * into EVERY rank's pool as {value, tag} pairs at byte offset y_off[i] (rows rank * N_i ...), stored by the reducer warps
 * over NVLink -- the all-gather is part of the mul_mat, no collective call.  Inputs are plain local vectors.
 * The pool starts with the counter area (b200q4_chain_sharded_ctr_bytes(n) bytes, zeroed, at the same place on every rank). */
size_t b200q4_chain_sharded_ctr_bytes(int n) { return (((size_t)(64 + 16 * (n + 4)) * 4) + 255) & ~(size_t)255; }

b200q4_chain *b200q4_chain_create_sharded(int n, const b200q4_weight *const *ws, const float *const *d_x, int rank, int world,
                                          void *const *pool_peers, size_t pool_bytes, const size_t *y_off)
{
    ApiLock lk;
    if (dev_info().device < 0) { api_fail(B200Q4_ENODEV, "b200q4_init() has not succeeded; no CPU fallback exists"); return nullptr; }
    if (n < 1 || !ws || !d_x || !pool_peers || !y_off || world < 2 || world > 8 || rank < 0 || rank >= world) {
        api_fail(B200Q4_EINVAL, "chain_create_sharded: bad argument");
        return nullptr;
    }
    b200q4_chain *c = new b200q4_chain();
    ProgBuild b;
    const size_t ctr_bytes = b200q4_chain_sharded_ctr_bytes(n);
    for (int i = 0; i < n; i++) {
        const b200q4_weight *w = ws[i];
        if (!w || w->transposed || !d_x[i] || (reinterpret_cast<uintptr_t>(d_x[i]) & 15) || (y_off[i] & 15) || y_off[i] < ctr_bytes ||
            y_off[i] + (size_t)2 * 4 * w->N * world > pool_bytes) {
            api_fail(B200Q4_EINVAL, "chain_create_sharded: bad op %d", i);
            delete c;
            return nullptr;
        }
        if (pg_geometry(w).PP > 1) {
            api_fail(B200Q4_EINVAL, "chain_create_sharded: op %d: K > 8192 is not supported here", i);
            delete c;
            return nullptr;
        }
        const PgSrc src = pg_src_f32(d_x[i]);
        const bool same_pack = i > 0 && d_x[i] == d_x[i - 1] && ws[i - 1]->K == w->K && ws[i - 1]->C == w->C;
        float *y = reinterpret_cast<float *>(static_cast<char *>(pool_peers[rank]) + y_off[i]);
        const int o = b.add_matvec(w, -1, PG_IN_PLAIN, PG_RES_KEEP, &src, 1, nullptr, nullptr, nullptr, y, nullptr, 0, true, same_pack, nullptr);
        if (o < 0) { delete c; return nullptr; }
        PgOp &op = b.ops[(size_t)o];
        op.row0 = rank * w->N;
        for (int r = 0; r < world; r++) op.y_peer[r] = reinterpret_cast<float *>(static_cast<char *>(pool_peers[r]) + y_off[i]);
    }
    if ((size_t)prog_counter_ints(b) * 4 > ctr_bytes) { api_fail(B200Q4_EINVAL, "chain_create_sharded: counter area too small"); delete c; return nullptr; }
    int *ctrs[8] = {nullptr};
    for (int r = 0; r < world; r++) ctrs[r] = static_cast<int *>(pool_peers[r]);
    if (prog_finalize_sharded(b, c->pg, rank, world, ctrs)) { delete c; return nullptr; }
    return c;
}

int b200q4_chain_run(b200q4_chain *c, void *stream)
{
    ApiLock lk;
    if (!c) return api_fail(B200Q4_EINVAL, "chain_run: NULL chain");
    if (c->pg.d_dbg)                                      // (timeline builds: the accumulators cover one run)
        cudaMemsetAsync(c->pg.d_dbg + (size_t)c->pg.grid * c->pg.n_ops * 8, 0, (size_t)c->pg.grid * 16 * sizeof(unsigned long long),
                        stream ? (cudaStream_t)stream : dev_info().stream);
    return prog_launch(c->pg, stream ? (cudaStream_t)stream : dev_info().stream);
}

/* bench helper (b200q4_dev.h): `launches` runs of chain c, then one of `tail` (may be NULL), between two CUDA events recorded
 * on `stream` from C -- no interpreter between the first event and the first launch; returns the milliseconds or a
 * negative error code */
double b200q4_chain_time(b200q4_chain *c, b200q4_chain *tail, int launches, void *stream)
{
    ApiLock lk;
    if (!c || launches < 0) return api_fail(B200Q4_EINVAL, "chain_time: bad argument");
    cudaStream_t st = stream ? (cudaStream_t)stream : dev_info().stream;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return api_fail(B200Q4_ECUDA, "chain_time: event create failed");
    double ms = B200Q4_ECUDA;
    int rc = B200Q4_OK;
    if (cudaEventRecord(e0, st) == cudaSuccess) {
        for (int i = 0; i < launches && rc == B200Q4_OK; i++) rc = prog_launch(c->pg, st);
        if (rc == B200Q4_OK && tail) rc = prog_launch(tail->pg, st);
        if (rc == B200Q4_OK && cudaEventRecord(e1, st) == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess) {
            float f = 0.f;
            if (cudaEventElapsedTime(&f, e0, e1) == cudaSuccess) ms = f;
        } else if (rc != B200Q4_OK) ms = rc;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

/* development aid (timeline builds): copies the [grid][n_ops][8] stamps of the last run; returns grid or a negative error */
int b200q4_chain_timeline(const b200q4_chain *c, unsigned long long *out, size_t cap_words)
{
    if (!c || !c->pg.d_dbg) return B200Q4_EINVAL;
    const size_t n = (size_t)c->pg.grid * c->pg.n_ops * 8 + (size_t)c->pg.grid * 16;
    if (cap_words < n) return B200Q4_EINVAL;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, c->pg.d_dbg, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? c->pg.grid : B200Q4_ECUDA;
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
