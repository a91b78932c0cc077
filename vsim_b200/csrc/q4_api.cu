// q4_api.cu -- the C ABI of include/b200q4.h on top of q4_kernels.cuh.
// Host-side plumbing only: device selection, weight residency, scratch
// buffers, kernel selection and launch.  No CPU compute path exists here: if
// CUDA is unusable every entry point fails with B200Q4_ENODEV / B200Q4_ECUDA.
#include "../../include/b200q4.h"
#include "q4_kernels.cuh"
#include "q4_tile.cuh"

using b200q4::PIPE_MAX_STAGES;
using b200q4::PIPE_TILE_SPLIT;

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

using namespace b200q4;

#include "q4_internal.h"

namespace {

struct Ctx {
    int device = -1;
    int sms = 0;
    int smem_optin = 0;
    int smem_per_sm = 0;
    bool two_ctas = false;           // B200Q4_TWO_CTAS=1: two 8-consumer CTAs per SM when they fit (measured slower: 11.3 vs 9.2 us)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_aq = nullptr; size_t aq_cap = 0;
    float *d_ap = nullptr;   size_t ap_cap = 0;
    float *d_partial = nullptr; size_t partial_cap = 0;
    unsigned long long *d_dbg = nullptr;   // B200Q4_DBG=1: per-CTA timeline of the last decode launch
    int *d_ticket = nullptr;         // [2] task ticket + CTAs-done, self-resetting
    int *d_done = nullptr;           // fused all-gather: CTAs that finished their peer stores, self-resetting
    float *d_yloc = nullptr; size_t yloc_cap = 0;   // sharded tile-kernel launches: local slice before the scatter
    bool use_pdl = true;             // B200Q4_PDL=0 disables programmatic dependent launch
    float *d_x = nullptr;    size_t x_cap = 0;
    float *d_y = nullptr;    size_t y_cap = 0;
    uint8_t *d_tmp = nullptr; size_t tmp_cap = 0;
    uint8_t *d_tpack = nullptr; size_t tpack_cap = 0;           // tile kernel: activation stages
    float *d_tpartial = nullptr; size_t tpartial_cap = 0;        // tile kernel: split-K pieces
    int *d_tcounters = nullptr; size_t tcounters_cap = 0;        // tile kernel: split-K arrival counters (self-resetting)
    int fuse_max_k = 1 << 30;        // B200Q4_FUSE_MAX_K: longer rows are quantized once by a PDL-chained pack kernel instead of in the
                                     // CTAs' prologues (K-slicing bounds a prologue to 128 quant blocks, so fusing is the default for any K)
    int tile_min_m = 9;              // M >= this runs the tcgen05 tile kernel, below it the decode kernel (B200Q4_TILE_MIN_M overrides; 0 = never)
    bool zero_copy = true;           // B200Q4_ZERO_COPY=0: always stage host buffers through cudaMemcpy
    uint32_t *h_flag = nullptr, *d_flag = nullptr;   // mapped host word the stream writes when a zero-copy call is complete
    uint32_t hseq = 0;
    struct HostBlock { const char *base; size_t size; char *dev; };
    std::vector<HostBlock> host_blocks;              // b200q4_host_alloc'ed ranges: their device aliases are known without a driver query
    // cuStreamWriteValue32 (driver API, fetched through the runtime: no link-time dependency on libcuda): the
    // completion word is written by the stream's front end instead of a one-thread kernel
    int (*stream_write32)(cudaStream_t, unsigned long long, uint32_t, unsigned int) = nullptr;
    int  init_gen = 0;               // bumped by every successful b200q4_init (per-device function attributes are set again)
    bool exact = false;              // b200q4_set_exact / B200Q4_EXACT=1: the reference's own fp32 rounding sequence (bit-identical Y)
    uint8_t *d_exq = nullptr; size_t exq_cap = 0;       // exact mode: reference-format quantized activation rows
    float *d_exdeq = nullptr; size_t exdeq_cap = 0;     // exact mode: the floats the reference's dot product forms from them
    const b200q4_weight *hint_w = nullptr; int hint_m = 1;   // b200q4_hint_next_weight: consumed by the next decode launch
    int prefetch_stages = 2;         // B200Q4_PREFETCH_STAGES: tasks per CTA of the hinted next launch to pull into L2 (0 = off); measured: 2 fills
                                     // the ~2 us gap between two launches (12 MB), more only competes with the launch's own last stages
    int static_tasks = 0;            // B200Q4_STATIC_TASKS: tasks per CTA issued before the grid dependency resolves (0 = kernel default, 3)
    int b8_min_stages = 3;           // B200Q4_B8_MIN_STAGES: the 256-byte-per-block code layout is used when this many stages still fit (0 = never)
    int slice_min_c = 0;             // B200Q4_SLICE_MIN_C: tiles longer than this many units (4 quant blocks each) are K-sliced; 0 = default rule
    int pipe_flags = 0;              // B200Q4_PIPE_FLAGS: development knobs of the decode kernel (PipeParams::flags)
    uint64_t launches = 0;
    std::recursive_mutex mu;
    std::string err;
};
Ctx g;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g.err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(B200Q4_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

template <typename T>
int ensure(T *&p, size_t &cap, size_t need_bytes)
{
    if (need_bytes <= cap) return B200Q4_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = need_bytes + need_bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        return fail(B200Q4_ENOMEM, "cudaMalloc(%zu) failed", want);
    }
    cap = want;
    return B200Q4_OK;
}

int need_init()
{
    if (g.device < 0) return fail(B200Q4_ENODEV, "b200q4_init() has not succeeded; no CPU fallback exists");
    return B200Q4_OK;
}

// ---- pipelined decode kernel (q4_gemv_pipe_kernel) ---------------------------------------------
struct PipePlan { int MP, stages, PP, TS, ntasks, ncw, nbp, gnbp; bool b8, m1; size_t smem; int rslots; };

constexpr int PIPE_MAX_CLASSES = 62;                   // K-slices per tile the ticket array has room for (K <= 253,952)

// K-slicing geometry (see q4_gemv_pipe_kernel): PP pieces per tile of TS units each; PP = 1 -> unsliced
void slice_geometry(int C, int MP, int &PP, int &TS, int &ntasks, int n_tiles)
{
    // Measured (tools/quick_bench.py --neox): for M <= 2 a tile of up to two stages (K <= 8192) is faster whole --
    // slicing buys a shorter prologue but costs the finalize launch on the dependency chain; the 8-row kernel's
    // 256-byte-per-block pack makes slicing pay from K > 4096 on.
    const int min_c = g.slice_min_c > 0 ? g.slice_min_c : (MP == 8 ? 32 : 64);
    if (C > min_c) {                                    // slice: one stage or less per piece
        PP = (C + 31) / 32;
        TS = (C + PP - 1) / PP;
        TS = (TS + 1) & ~1;                            // even: unit pairs stay aligned to quant-block octets
        PP = (C + TS - 1) / TS;
        ntasks = n_tiles * PP;
    } else if (C >= PIPE_TILE_SPLIT) {                  // one task = one tile
        PP = 1; TS = C; ntasks = n_tiles;
    } else {                                            // short tiles: a task = TS/C whole tiles
        PP = 1;
        const int G = 32 / C > 0 ? 32 / C : 1;
        TS = G * C;
        ntasks = (int)(((long)n_tiles * C + TS - 1) / TS);
    }
}

// m1: a single activation row -> octet mode (codes quantized by the kernel itself when `fuse`, else by pack_*_kernel with one row)
PipePlan pipe_plan(const b200q4_weight *w, int M, bool fuse)
{
    PipePlan pp{};
    const int type = w->type;
    pp.MP = (M <= 2) ? 2 : 8;
    slice_geometry(w->C, pp.MP, pp.PP, pp.TS, pp.ntasks, w->n_tiles);
    const size_t optin = (size_t)(g.smem_optin > 0 ? g.smem_optin : 232448);
    auto stage_of = [&](int MP, int ncw) { return (size_t)2 * ncw * chunk_bytes(type) + (size_t)2 * ncw * (MP == 2 ? 128 : 512); };
    // quant blocks a CTA keeps in shared memory: its K-slice, or the whole (padded) row
    auto nbp_of = [&](bool m1) { return pp.PP > 1 ? pp.TS * CHUNK_BLKS : (m1 ? ((w->C * CHUNK_BLKS + 7) & ~7) : w->C * CHUNK_BLKS); };
    auto fixed = [&](int MP, bool m1, bool b8) {
        const int rows = m1 ? 1 : MP;
        const int nbp = nbp_of(m1);
        const size_t planes = ((size_t)(type == 0 ? 2 : 3) * nbp * rows * 4 + 15) & ~(size_t)15;
        return (size_t)nbp * (MP == 2 && b8 ? 256 : rows * 32) + planes + 3 * PIPE_MAX_STAGES * 8 + PIPE_MAX_STAGES * 16 + 128 + 256;
    };
    if (pp.MP == 8 && fixed(8, false, false) + 2 * stage_of(8, 16) > optin) {     // (cannot happen with K-slicing; kept as a guard)
        pp.MP = 2;
        slice_geometry(w->C, pp.MP, pp.PP, pp.TS, pp.ntasks, w->n_tiles);
    }
    pp.m1 = pp.MP == 2 && M == 1 && (fuse || !g.two_ctas);
    // Optional: two CTAs of 8 consumer warps per SM (each gets half of the SM's shared memory minus
    // the 1 KB the hardware reserves per CTA) with the 8-column code layout and >= 3 stages of ring.
    const size_t half = (size_t)(g.smem_per_sm > 0 ? g.smem_per_sm : 233472) / 2 - 1024;
    if (g.two_ctas && pp.MP == 2 && fixed(2, pp.m1, true) + 3 * stage_of(2, 8) <= half) {
        pp.ncw = 8; pp.b8 = true;
        long st = (long)((half - fixed(2, pp.m1, true)) / stage_of(2, 8));
        pp.stages = (int)(st > PIPE_MAX_STAGES ? PIPE_MAX_STAGES : st);
        pp.smem = fixed(2, pp.m1, true) + (size_t)pp.stages * stage_of(2, 8);
    } else {
        pp.ncw = 16;
        const size_t stage = stage_of(pp.MP, 16);
        // M = 2: the 256-byte-per-block code layout (one 8-byte load per IMMA) when >= 3 stages still fit.  M = 1 (octet
        // mode) keeps the compact layout: the extra stage of ring is worth more than the saved selects (measured:
        // 16384x4096 9.26 -> 9.16 us, lm_head 92.4 -> 95.6 %, 6144x24576 75.9 -> 79.1 % of the HBM peak).
        pp.b8 = pp.MP == 2 && !pp.m1 && g.b8_min_stages > 0 && fixed(2, pp.m1, true) + g.b8_min_stages * stage <= optin;
        if (fixed(pp.MP, pp.m1, pp.b8) + 2 * stage > optin) { pp.MP = 0; return pp; }
        long st = (long)((optin - fixed(pp.MP, pp.m1, pp.b8)) / stage);
        pp.stages = (int)(st > PIPE_MAX_STAGES ? PIPE_MAX_STAGES : st);
        pp.smem = fixed(pp.MP, pp.m1, pp.b8) + (size_t)pp.stages * stage;
        // 8-row kernel: a reduction ring of its own (2 entries of one 16 x 8 partial per consumer warp, 8 KB each) instead of
        // 16 KB of partials per weight stage, whenever a warp's two units always belong to one tile (whole tiles of >= 16
        // units, K-slices, or an even number of units per tile) and a weight stage more fits that way
        // (measured, 16384 x 4096, M = 8: Q4_1 2 -> 3 stages 21.5 -> 18.8 us; Q4_0 3 -> 4 stages 15.1 -> 16.1 us: with three stages
        // the ring is deep enough and the extra wait per stage costs more than the fourth stage buys)
        if (pp.MP == 8 && pp.stages < 3 && (pp.PP > 1 || w->C >= PIPE_TILE_SPLIT || w->C % 2 == 0) && !getenv("B200Q4_NO_RDEC")) {
            const size_t wstage = (size_t)2 * 16 * chunk_bytes(type), rring = (size_t)2 * 16 * 512;
            const long st2 = (long)((optin - fixed(8, false, false) - rring) / wstage);
            const int stages2 = (int)(st2 > PIPE_MAX_STAGES ? PIPE_MAX_STAGES : st2);
            if (stages2 > pp.stages) {
                pp.rslots = 2; pp.stages = stages2;
                pp.smem = fixed(8, false, false) + rring + (size_t)stages2 * wstage;
            }
        }
    }
    pp.nbp = nbp_of(pp.m1);
    pp.gnbp = pp.PP > 1 ? pp.PP * pp.TS * CHUNK_BLKS : pp.nbp;     // global pack: every slice fully backed
    return pp;
}

template <int QT, int MP, bool FUSED, bool M1, bool B8, int NCW>
int launch_pipe(const PipeParams &p, int grid, size_t smem, cudaStream_t st)
{
    static int attr_gen = -1;                                // (function attributes are per device and context: redone after every b200q4_init)
    if (attr_gen != g.init_gen) {
        CK(cudaFuncSetAttribute(q4_gemv_pipe_kernel<QT, MP, FUSED, M1, B8, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_optin));
        if (NCW == 8)
            CK(cudaFuncSetAttribute(q4_gemv_pipe_kernel<QT, MP, FUSED, M1, B8, NCW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_gen = g.init_gen;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3((NCW + 2) * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g.use_pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, q4_gemv_pipe_kernel<QT, MP, FUSED, M1, B8, NCW>, p));
    g.launches++;
    return B200Q4_OK;
}

// small helper kernels that take part in programmatic dependent launch (they call
// cudaGridDependencySynchronize themselves): launched with the PDL attribute so they do not serialise the stream
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), int blocks, int threads, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g.use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

int run_pipe(const b200q4_weight *w, int mc, const PipePlan &pl, bool fused, const float *d_x, size_t x_stride_f,
             float *d_y, size_t y_stride_f, cudaStream_t st, const PeerOut *peer)
{
    PipeParams p;
    if (peer) p.peer = *peer; else { p.peer = PeerOut{}; p.peer.world = 1; }
    p.wt = w->d_wt; p.x = d_x; p.aq = g.d_aq; p.ap = g.d_ap; p.y = d_y; p.partial = g.d_partial;
    p.ticket = g.d_ticket;
    p.x_stride = x_stride_f; p.y_stride = y_stride_f;
    p.U = w->n_tiles * w->C; p.N = w->N; p.M = mc; p.C = w->C; p.n_tiles = w->n_tiles; p.nb = w->nb;
    p.ntasks = pl.ntasks; p.PP = pl.PP; p.TS = pl.TS; p.stages = pl.stages; p.rslots = pl.rslots; p.gnbp = pl.gnbp; p.flags = g.pipe_flags; p.nstatic = g.static_tasks;
    p.dbg = g.d_dbg;
    p.next = NextHint{};
    if (g.hint_w && g.prefetch_stages > 0 && !g.hint_w->transposed && g.hint_w->wt_bytes &&
        !(g.tile_min_m > 0 && g.hint_m >= g.tile_min_m) && !g.exact) {
        const b200q4_weight *h = g.hint_w;
        const PipePlan hp = pipe_plan(h, g.hint_m, true);
        if (hp.MP > 0) {
            const int hslots = g.sms * (hp.ncw == 8 ? 2 : 1);
            p.next.wt = h->d_wt; p.next.C = h->C; p.next.TS = hp.TS; p.next.PP = hp.PP; p.next.U = h->n_tiles * h->C;
            p.next.n_tiles = h->n_tiles; p.next.ntasks = hp.ntasks; p.next.grid = hp.ntasks < hslots ? hp.ntasks : hslots;
            p.next.chunk = chunk_bytes(h->type); p.next.stages = g.prefetch_stages;
        }
    }
    g.hint_w = nullptr;                                  // one-shot
    const int slots = g.sms * (pl.ncw == 8 ? 2 : 1);
    const int grid = pl.ntasks < slots ? pl.ntasks : slots;
    const bool m1 = pl.m1 && mc == 1;
#define PIPE_DISPATCH(QT)                                                                                          \
    if (pl.ncw == 8) {                                 /* MP == 2, B8 */                                           \
        if (fused) return m1 ? launch_pipe<QT, 2, true, true, true, 8>(p, grid, pl.smem, st)                       \
                             : launch_pipe<QT, 2, true, false, true, 8>(p, grid, pl.smem, st);                     \
        return launch_pipe<QT, 2, false, false, true, 8>(p, grid, pl.smem, st);                                    \
    }                                                                                                              \
    if (pl.MP == 8) return pl.rslots > 0 ? launch_pipe<QT, 8, false, false, true, 16>(p, grid, pl.smem, st)        \
                                         : launch_pipe<QT, 8, false, false, false, 16>(p, grid, pl.smem, st);      \
    if (fused) {                                                                                                   \
        if (pl.b8) return m1 ? launch_pipe<QT, 2, true, true, true, 16>(p, grid, pl.smem, st)                      \
                             : launch_pipe<QT, 2, true, false, true, 16>(p, grid, pl.smem, st);                    \
        return m1 ? launch_pipe<QT, 2, true, true, false, 16>(p, grid, pl.smem, st)                                \
                  : launch_pipe<QT, 2, true, false, false, 16>(p, grid, pl.smem, st);                              \
    }                                                                                                              \
    if (m1) return pl.b8 ? launch_pipe<QT, 2, false, true, true, 16>(p, grid, pl.smem, st)                         \
                         : launch_pipe<QT, 2, false, true, false, 16>(p, grid, pl.smem, st);                       \
    return pl.b8 ? launch_pipe<QT, 2, false, false, true, 16>(p, grid, pl.smem, st)                                \
                 : launch_pipe<QT, 2, false, false, false, 16>(p, grid, pl.smem, st);
    if (w->type == 0) { PIPE_DISPATCH(0) }
    PIPE_DISPATCH(1)
#undef PIPE_DISPATCH
}


// ---- tcgen05 tile kernel (q4_tile_kernel) -------------------------------------------------------
struct TilePlan { int MT, m_tiles, n_tiles, ntiles, grid, nraw, nact, nconv; size_t smem; };

TilePlan tile_plan(const b200q4_weight *w, int M)
{
    TilePlan tp{};
    const int qt = w->type;
    const int cap = tk_max_mt(qt);                                            // 128 (Q4_0) / 64 (Q4_1): the TMEM map
    tp.MT = M <= 16 ? 16 : M <= 32 ? 32 : (M <= 64 || cap == 64) ? 64 : 128;
    tp.m_tiles = (M + tp.MT - 1) / tp.MT;
    tp.n_tiles = (w->n_tiles + TK_T16 - 1) / TK_T16;
    tp.ntiles = tp.n_tiles * tp.m_tiles;
    const long steps = (long)tp.ntiles * w->C;                                // stream-K: chunk steps are dealt out evenly
    const int sms = g.sms > 0 ? g.sms : 148;
    tp.grid = (int)(steps < sms ? steps : sms);
    // shared-memory rings: activation stages, converted-operand slots, and whatever is left for raw weight stages
    const size_t optin = (size_t)(g.smem_optin > 0 ? g.smem_optin : 232448);
    tp.nact = tp.MT >= 128 ? 2 : 3;
    tp.nconv = 3;
    auto raw_fit = [&]() { return ((long)optin - (long)tk_smem_bytes(qt, tp.MT, 0, tp.nact, tp.nconv)) / tk_raw_stage(qt); };
    long nraw = raw_fit();
    while (nraw < 3 && tp.nconv > 2) { tp.nconv--; nraw = raw_fit(); }
    while (nraw < 3 && tp.nact > 2) { tp.nact--; nraw = raw_fit(); }
    // fewer TMEM hand-offs per chunk than issuers (small MT): ring depths that are multiples of the issuer-group count let an
    // issuer skip the other groups' chunks (tk_groups, q4_tile.cuh) -- taken when at least four raw stages still fit
    const int groups = tk_groups(qt, tp.MT);
    if (groups > 1 && qt == 0) {                                              // (Q4_1, MT = 16: two 48 KB slots instead of three cost more than the waits)
        const int keep_act = tp.nact, keep_conv = tp.nconv;
        tp.nact = 4;
        tp.nconv = 4;
        if (raw_fit() >= 4) nraw = raw_fit();
        else { tp.nact = keep_act; tp.nconv = keep_conv; }
    }
    if (nraw > TK_MAX_RAW) nraw = TK_MAX_RAW;
    tp.nraw = (int)nraw;
    tp.smem = tk_smem_bytes(qt, tp.MT, tp.nraw, tp.nact, tp.nconv);
    return tp;
}

template <int QT, int MT>
int launch_tile(const TileParams &p, int grid, size_t smem, cudaStream_t st)
{
    static int attr_gen = -1;
    if (attr_gen != g.init_gen) {
        CK(cudaFuncSetAttribute(q4_tile_kernel<QT, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_optin));
        attr_gen = g.init_gen;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tk_threads(QT, MT)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g.use_pdl ? 1 : 0;     // starts streaming weights while the pack kernel runs
    CK(cudaLaunchKernelEx(&cfg, q4_tile_kernel<QT, MT>, p));
    g.launches++;
    return B200Q4_OK;
}

int run_tile(const b200q4_weight *w, const float *d_x, size_t x_stride_f, const uint8_t *d_wdata, int M,
             float *d_y, size_t y_stride_f, cudaStream_t st)
{
    const TilePlan tp = tile_plan(w, M);
    if (tp.nraw < 2 || tp.smem > (size_t)g.smem_optin) return fail(B200Q4_EINVAL, "tile kernel: shared memory too small");
    if ((long)tp.ntiles * w->C > 0x7fffffffL) return fail(B200Q4_EINVAL, "tile kernel: tensor too large");
    int rc;
    const size_t stage = (size_t)tk_act_stage(w->type, tp.MT);
    if ((rc = ensure(g.d_tpack, g.tpack_cap, (size_t)tp.m_tiles * w->C * stage))) return rc;
    if ((rc = ensure(g.d_tpartial, g.tpartial_cap, (size_t)tp.grid * 2 * tp.MT * TK_ROWS * sizeof(float)))) return rc;
    const size_t need = (size_t)tp.ntiles * sizeof(int);
    if (need > g.tcounters_cap) {
        if ((rc = ensure(g.d_tcounters, g.tcounters_cap, need))) return rc;
        CK(cudaMemsetAsync(g.d_tcounters, 0, g.tcounters_cap, st));
    }
    // activation pack: small grids (M <= 64) one warp per (stage row, block) as a programmatic dependent launch -- set up while
    // the kernel before it drains, waiting for that kernel's results (X; the previous call's tile kernel reading d_tpack) at
    // its first instruction; larger ones one thread per (stage row, block), launched plainly (the tile kernel launches when
    // every pack CTA has started, i.e. at the end of a grid that is not resident at once anyway)
    const long pack_units = (long)tp.m_tiles * w->C * CHUNK_BLKS * tp.MT;
    const bool pack_warp = M <= 64;
    const int blocks = pack_warp ? (int)((pack_units * 32 + 255) / 256) : (int)((pack_units + 127) / 128);
    static const int dev_skip = getenv("B200Q4_TILE_SKIP") ? atoi(getenv("B200Q4_TILE_SKIP")) : 0;    // development: 1 = no pack kernel, 2 = no tile kernel
    cudaLaunchConfig_t pcfg = {};
    pcfg.gridDim = dim3(blocks); pcfg.blockDim = dim3(256); pcfg.dynamicSmemBytes = 0; pcfg.stream = st;
    cudaLaunchAttribute pat[1];
    pat[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pat[0].val.programmaticStreamSerializationAllowed = 1;
    pcfg.attrs = pat; pcfg.numAttrs = 1;
#define TILE_PACK(QT)                                                                                                  \
    if (dev_skip == 1) { }                                                                                             \
    else if (pack_warp && g.use_pdl && d_wdata) CK(cudaLaunchKernelEx(&pcfg, tile_pack_warp_kernel<QT, true, true>, (const float *)nullptr, (size_t)0, d_wdata, M, tp.MT, w->nb, w->C, tp.m_tiles, g.d_tpack)); \
    else if (pack_warp && g.use_pdl)            CK(cudaLaunchKernelEx(&pcfg, tile_pack_warp_kernel<QT, false, true>, d_x, x_stride_f, (const uint8_t *)nullptr, M, tp.MT, w->nb, w->C, tp.m_tiles, g.d_tpack)); \
    else if (pack_warp && d_wdata) tile_pack_warp_kernel<QT, true, false><<<blocks, 256, 0, st>>>(nullptr, 0, d_wdata, M, tp.MT, w->nb, w->C, tp.m_tiles, g.d_tpack); \
    else if (pack_warp)            tile_pack_warp_kernel<QT, false, false><<<blocks, 256, 0, st>>>(d_x, x_stride_f, nullptr, M, tp.MT, w->nb, w->C, tp.m_tiles, g.d_tpack); \
    else if (d_wdata) tile_pack_kernel<QT, true><<<blocks, 128, 0, st>>>(nullptr, 0, d_wdata, M, tp.MT, w->nb, w->C, tp.m_tiles, g.d_tpack); \
    else              tile_pack_kernel<QT, false><<<blocks, 128, 0, st>>>(d_x, x_stride_f, nullptr, M, tp.MT, w->nb, w->C, tp.m_tiles, g.d_tpack);
    if (w->type == 0) { TILE_PACK(0) } else { TILE_PACK(1) }
#undef TILE_PACK
    g.launches++;
    CK(cudaGetLastError());
    TileParams p;
    p.wt = w->d_wt; p.tpack = g.d_tpack; p.y = d_y; p.partial = g.d_tpartial; p.counters = g.d_tcounters;
    p.y_stride = y_stride_f; p.N = w->N; p.M = M; p.C = w->C; p.n_t16 = w->n_tiles; p.m_tiles = tp.m_tiles;
    p.ntiles = tp.ntiles; p.nraw = tp.nraw; p.nact = tp.nact; p.nconv = tp.nconv;
    p.dbg = g.d_dbg;
    if (dev_skip == 2) return B200Q4_OK;
#define TILE_DISPATCH(QT)                                                              \
    switch (tp.MT) {                                                                   \
    case 16:  return launch_tile<QT, 16>(p, tp.grid, tp.smem, st);                     \
    case 32:  return launch_tile<QT, 32>(p, tp.grid, tp.smem, st);                     \
    default:  return launch_tile<QT, 64>(p, tp.grid, tp.smem, st);                     \
    }
    if (w->type == 0) {
        if (tp.MT == 128) return launch_tile<0, 128>(p, tp.grid, tp.smem, st);
        TILE_DISPATCH(0)
    }
    TILE_DISPATCH(1)
#undef TILE_DISPATCH
}

// ---- exact mode: bit-identical to ggml_vec_dot_q4_0/_q4_1 (see q4_kernels.cuh) ----------------------------
int run_exact(const b200q4_weight *w, const float *d_x, size_t x_stride_f, const uint8_t *d_wdata, int M,
              float *d_y, size_t y_stride_f, cudaStream_t st)
{
    const size_t rb = (size_t)w->nb * (w->type == 0 ? 20 : 24);
    int rc;
    if ((rc = ensure(g.d_exdeq, g.exdeq_cap, (size_t)M * w->K * sizeof(float)))) return rc;
    const uint8_t *rows = d_wdata;
    if (!rows) {                                         // byte-exact quantize_row_q4_0/_q4_1 image of X
        if ((rc = ensure(g.d_exq, g.exq_cap, rb * (size_t)M))) return rc;
        const long warps = (long)w->nb * M;
        const int blocks = (int)((warps * 32 + 255) / 256);
        if (w->type == 0) quantize_rows_ref_kernel<0><<<blocks, 256, 0, st>>>(d_x, x_stride_f, M, w->nb, g.d_exq);
        else              quantize_rows_ref_kernel<1><<<blocks, 256, 0, st>>>(d_x, x_stride_f, M, w->nb, g.d_exq);
        g.launches++;
        rows = g.d_exq;
    }
    {
        const long total = (long)M * w->K;
        const int blocks = (int)((total + 255) / 256);
        if (w->type == 0) exact_dequant_act_kernel<0><<<blocks, 256, 0, st>>>(rows, M, w->nb, g.d_exdeq);
        else              exact_dequant_act_kernel<1><<<blocks, 256, 0, st>>>(rows, M, w->nb, g.d_exdeq);
        g.launches++;
    }
    const int blocks = w->n_tiles;                       // one block per 16-row tile
    for (int m0 = 0; m0 < M; m0 += 8) {
        const int mc = M - m0 < 8 ? M - m0 : 8;
        const float *dq = g.d_exdeq + (size_t)m0 * w->K;
        float *yo = d_y + (size_t)m0 * y_stride_f;
#define EXACT(QT)                                                                                                             \
        if (mc == 1)      mul_mat_exact_kernel<QT, 1><<<blocks, 128, 0, st>>>(w->d_wt, w->N, w->C, w->nb, dq, mc, w->K, yo, y_stride_f); \
        else if (mc == 2) mul_mat_exact_kernel<QT, 2><<<blocks, 128, 0, st>>>(w->d_wt, w->N, w->C, w->nb, dq, mc, w->K, yo, y_stride_f); \
        else if (mc <= 4) mul_mat_exact_kernel<QT, 4><<<blocks, 128, 0, st>>>(w->d_wt, w->N, w->C, w->nb, dq, mc, w->K, yo, y_stride_f); \
        else              mul_mat_exact_kernel<QT, 8><<<blocks, 128, 0, st>>>(w->d_wt, w->N, w->C, w->nb, dq, mc, w->K, yo, y_stride_f);
        if (w->type == 0) { EXACT(0) } else { EXACT(1) }
#undef EXACT
        g.launches++;
    }
    CK(cudaGetLastError());
    return B200Q4_OK;
}

int ensure_scratch(int nbp, int MP)
{
    int rc;
    if ((rc = ensure(g.d_aq, g.aq_cap, (size_t)nbp * MP * 32))) return rc;
    if ((rc = ensure(g.d_ap, g.ap_cap, (size_t)nbp * MP * 12))) return rc;
    return B200Q4_OK;
}

int mul_mat_dev_locked(const b200q4_weight *w, const float *d_x, int M, size_t x_stride, float *d_y,
                       size_t y_stride, cudaStream_t st, const uint8_t *d_wdata, const PeerOut *peer = nullptr)
{
    if (!w || w->transposed || M < 0 || (!d_x && !d_wdata) || (!d_y && !peer)) return fail(B200Q4_EINVAL, "mul_mat: bad argument");
    if ((x_stride % 4) || (y_stride % 4)) return fail(B200Q4_EINVAL, "mul_mat: strides must be multiples of 4 bytes");
    if (M == 0 || w->N == 0) return B200Q4_OK;
    if ((long)w->n_tiles * w->C > 0x7fffffffL) return fail(B200Q4_EINVAL, "tensor too large");
    if (g.exact) {
        if (!peer) return run_exact(w, d_x, x_stride / 4, d_wdata, M, d_y, y_stride / 4, st);
        // sharded: the slice is computed locally, then stored into every rank's Y like the tile path below
        if (int rc = ensure(g.d_yloc, g.yloc_cap, (size_t)M * w->N * sizeof(float))) return rc;
        wait_flags_kernel<<<1, 32, 0, st>>>(peer->local_flags, peer->world, peer->seq_ctr);
        g.launches++;
        if (int rc = run_exact(w, d_x, x_stride / 4, d_wdata, M, g.d_yloc, (size_t)w->N, st)) return rc;
        const long total = (long)M * w->N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 4 * g.sms) blocks = 4 * g.sms;
        scatter_slice_kernel<<<blocks, 256, 0, st>>>(g.d_yloc, (size_t)w->N, M, w->N, y_stride / 4, *peer);
        g.launches++;
        CK(cudaGetLastError());
        return B200Q4_OK;
    }
    if (g.tile_min_m > 0 && M >= g.tile_min_m) {
        if (!peer) return run_tile(w, d_x, x_stride / 4, d_wdata, M, d_y, y_stride / 4, st);
        // sharded: wait for the previous exchange (X), let the tile kernel write this rank's slice locally,
        // then a copy kernel stores it into every rank's Y
        if (int rc = ensure(g.d_yloc, g.yloc_cap, (size_t)M * w->N * sizeof(float))) return rc;
        wait_flags_kernel<<<1, 32, 0, st>>>(peer->local_flags, peer->world, peer->seq_ctr);
        g.launches++;
        if (int rc = run_tile(w, d_x, x_stride / 4, d_wdata, M, g.d_yloc, (size_t)w->N, st)) return rc;
        const long total = (long)M * w->N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 4 * g.sms) blocks = 4 * g.sms;
        scatter_slice_kernel<<<blocks, 256, 0, st>>>(g.d_yloc, (size_t)w->N, M, w->N, y_stride / 4, *peer);
        g.launches++;
        CK(cudaGetLastError());
        return B200Q4_OK;
    }
    // the fused prologue reads X with 128-bit loads
    const bool fuse = !d_wdata && d_x && (reinterpret_cast<uintptr_t>(d_x) % 16 == 0) && (x_stride % 16 == 0) && w->K <= g.fuse_max_k;
    const PipePlan pp = pipe_plan(w, M, fuse);
    const int MP = pp.MP;
    if (MP < 1) return fail(B200Q4_EINVAL, "mul_mat: K=%d too large for the shared-memory activation pack", w->K);
    if (pp.PP > PIPE_MAX_CLASSES) return fail(B200Q4_EINVAL, "mul_mat: K=%d too large (more than %d K-slices)", w->K, PIPE_MAX_CLASSES);
    int rc;
    if ((rc = ensure_scratch(pp.gnbp, MP))) return rc;
    if (pp.PP > 1 && (rc = ensure(g.d_partial, g.partial_cap, (size_t)pp.ntasks * 128 * sizeof(float)))) return rc;
    const size_t rbw = (size_t)w->nb * (w->type == 0 ? 20 : 24);
    // global pack geometry = what the kernel's shared-memory copy expects: octet mode keeps ONE row per block and pads to whole octets
    const int prow = pp.m1 ? 1 : MP;
    const int nbp = pp.gnbp;
    for (int m0 = 0; m0 < M; m0 += MP) {
        const int mc = (M - m0 < MP) ? M - m0 : MP;
        const float *xs = d_x ? reinterpret_cast<const float *>(reinterpret_cast<const char *>(d_x) + (size_t)m0 * x_stride) : nullptr;
        float *yo = reinterpret_cast<float *>(reinterpret_cast<char *>(d_y) + (size_t)m0 * y_stride);
        const bool fused = (MP == 2) && fuse;
        if (!fused) {
            const long warps = (long)nbp * prow;
            const int blocks = (int)((warps * 32 + 255) / 256);
            if (d_wdata) {
                const uint8_t *src = d_wdata + (size_t)m0 * rbw;
                CK(launch_pdl(w->type == 0 ? pack_wdata_kernel<0> : pack_wdata_kernel<1>, blocks, 256, st, src, mc, prow, w->nb, nbp, g.d_aq, g.d_ap));
            } else {
                CK(launch_pdl(w->type == 0 ? pack_act_kernel<0> : pack_act_kernel<1>, blocks, 256, st, xs, x_stride / 4, mc, prow, w->nb, nbp, g.d_aq, g.d_ap));
            }
            g.launches++;
        }
        PeerOut po{};
        if (peer) {                                      // rows m0.. of every rank's Y; only the last launch signals
            po = *peer;
            for (int b = 0; b < 2; b++)
                for (int r = 0; r < po.world; r++)
                    po.y[b][r] = reinterpret_cast<float *>(reinterpret_cast<char *>(po.y[b][r]) + (size_t)m0 * y_stride);
            if (m0 + MP < M) po.done = nullptr;
        }
        PeerOut pk = po;                                 // the decode kernel stores to the peers but never signals
        pk.done = nullptr;
        if ((rc = run_pipe(w, mc, pp, fused, xs, x_stride / 4, yo, y_stride / 4, st, peer ? &pk : nullptr))) return rc;
        if (peer && pp.PP == 1 && po.done) {             // last launch of an unsliced call: publish the exchange
            CK(launch_pdl(peer_signal_kernel, 1, 32, st, po));
            g.launches++;
        }
        if (pp.PP > 1) {                                 // K-sliced: the pieces are parked, this adds them up and stores Y
            PeerOut fo;
            if (peer) fo = po; else { fo = PeerOut{}; fo.world = 1; }
            const long total = (long)mc * w->N;
            int blocks = (int)((total + 127) / 128);
            if (blocks > 8 * g.sms) blocks = 8 * g.sms;
            const bool m1 = pp.m1 && mc == 1;
            cudaError_t e;
#define FIN(QT) (m1 ? launch_pdl(pipe_finalize_kernel<QT, 2, true>, blocks, 128, st, (const float *)g.d_partial, pp.PP, w->N, mc, yo, y_stride / 4, fo) \
               : MP == 2 ? launch_pdl(pipe_finalize_kernel<QT, 2, false>, blocks, 128, st, (const float *)g.d_partial, pp.PP, w->N, mc, yo, y_stride / 4, fo) \
                         : launch_pdl(pipe_finalize_kernel<QT, 8, false>, blocks, 128, st, (const float *)g.d_partial, pp.PP, w->N, mc, yo, y_stride / 4, fo))
            e = w->type == 0 ? FIN(0) : FIN(1);
#undef FIN
            CK(e);
            g.launches++;
        }
    }
    return B200Q4_OK;
}

// ---- zero-copy host path (decode calls on pinned buffers) -------------------------------------------
// X rows are pulled over PCIe ONCE by a tiny PDL-chained kernel (the decode kernel that follows is already
// filling its stage ring meanwhile), Y is stored by the decode epilogue straight into the caller's buffer,
// and completion is a word in mapped host memory that the calling thread polls -- no copy-engine round trips,
// no stream-synchronize wake-up.
__global__ void fetch_rows_kernel(const float *__restrict__ hx, size_t x_stride_f, int M, int K, float *__restrict__ dx)
{
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
    const int k4 = K / 4;
    const long total = (long)M * k4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int m = (int)(i / k4), j = (int)(i - (long)m * k4);
        reinterpret_cast<float4 *>(dx)[i] = reinterpret_cast<const float4 *>(hx + (size_t)m * x_stride_f)[j];
    }
}
__global__ void signal_host_kernel(uint32_t *flag, uint32_t v)
{
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(flag) = v;
}
// device alias of a pinned (cudaHostAlloc / cudaHostRegister) host pointer, or NULL for pageable memory
void *pinned_alias(const void *p)
{
    for (const auto &hb : g.host_blocks)
        if ((const char *)p >= hb.base && (const char *)p < hb.base + hb.size) return hb.dev + ((const char *)p - hb.base);
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}
// wait for the stream to write `seq` into the mapped word; falls back to the driver now and then so that a
// faulted stream is reported instead of spinning forever
int wait_host_flag(uint32_t seq)
{
    volatile uint32_t *f = g.h_flag;
    for (unsigned spins = 0;; spins++) {
        if (*f == seq) return B200Q4_OK;
        if ((spins & 0x3fff) == 0x3fff) {
            cudaError_t e = cudaStreamQuery(g.stream);
            if (e == cudaSuccess) { if (*f == seq) return B200Q4_OK; CK(cudaStreamSynchronize(g.stream)); return B200Q4_OK; }
            if (e != cudaErrorNotReady) return fail(B200Q4_ECUDA, "stream failed: %s", cudaGetErrorString(e));
        }
    }
}

b200q4_weight *upload_impl(int type, const void *rows, int N, int K, size_t row_stride, bool transposed)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (need_init()) return nullptr;
    // for the transposed form the quantized axis is N (rows are K long in count, N wide in values)
    const int qlen = transposed ? N : K, nrows = transposed ? K : N;
    if ((type != 0 && type != 1) || !rows || N < 0 || K < 0 || qlen % QK != 0) {
        fail(B200Q4_EINVAL, "weight_upload: bad argument (type=%d N=%d K=%d)", type, N, K);
        return nullptr;
    }
    const int nb = qlen / QK;
    const size_t rb = (size_t)nb * (type == 0 ? 20 : 24);
    if (row_stride == 0) row_stride = rb;
    if (row_stride < rb) { fail(B200Q4_EINVAL, "weight_upload: row_stride < row bytes"); return nullptr; }
    b200q4_weight *w = new b200q4_weight();
    w->type = type; w->N = N; w->K = K; w->nb = nb; w->transposed = transposed;
    w->d_wt = nullptr; w->d_counters = nullptr;
    const size_t raw = rb * (size_t)nrows;
    auto bail = [&](int code, const char *what) { fail(code, "weight_upload: %s", what); b200q4_weight_free(w); return (b200q4_weight *)nullptr; };
    if (transposed) {
        w->C = 0; w->n_tiles = 0; w->wt_bytes = raw;
        if (raw && cudaMalloc(&w->d_wt, raw) != cudaSuccess) return bail(B200Q4_ENOMEM, "cudaMalloc failed");
        if (raw && cudaMemcpy2D(w->d_wt, rb, rows, row_stride, rb, nrows, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(B200Q4_ECUDA, "H2D copy failed");
        return w;
    }
    w->C = (nb + CHUNK_BLKS - 1) / CHUNK_BLKS;
    w->n_tiles = (N + TILE_ROWS - 1) / TILE_ROWS;
    w->wt_bytes = (size_t)w->n_tiles * w->C * chunk_bytes(type);
    if (w->wt_bytes == 0) return w;
    if (cudaMalloc(&w->d_wt, w->wt_bytes) != cudaSuccess) return bail(B200Q4_ENOMEM, "cudaMalloc(wtile) failed");
    if (cudaMalloc(&w->d_counters, sizeof(int) * (size_t)w->n_tiles) != cudaSuccess) return bail(B200Q4_ENOMEM, "cudaMalloc(counters) failed");
    if (cudaMemsetAsync(w->d_counters, 0, sizeof(int) * (size_t)w->n_tiles, g.stream) != cudaSuccess) return bail(B200Q4_ECUDA, "memset failed");
    uint8_t *d_raw = nullptr;
    if (cudaMalloc(&d_raw, raw) != cudaSuccess) return bail(B200Q4_ENOMEM, "cudaMalloc(staging) failed");
    if (cudaMemcpy2DAsync(d_raw, rb, rows, row_stride, rb, N, cudaMemcpyHostToDevice, g.stream) != cudaSuccess) {
        cudaFree(d_raw);
        return bail(B200Q4_ECUDA, "H2D copy failed");
    }
    const long slots = (long)w->n_tiles * w->C * (chunk_bytes(type) / 16);
    const int blocks = (int)((slots + 255) / 256);
    if (type == 0) repack_kernel<0><<<blocks, 256, 0, g.stream>>>(d_raw, N, nb, w->n_tiles, w->C, w->d_wt);
    else           repack_kernel<1><<<blocks, 256, 0, g.stream>>>(d_raw, N, nb, w->n_tiles, w->C, w->d_wt);
    g.launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    cudaFree(d_raw);
    if (e != cudaSuccess) return bail(B200Q4_ECUDA, cudaGetErrorString(e));
    return w;
}

} // namespace

namespace b200q4 {
DevInfo dev_info() { return DevInfo{g.device, g.sms, g.smem_optin, g.stream}; }
int api_fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g.err = buf;
    return code;
}
void api_count_launch(int n) { g.launches += (uint64_t)n; }
void api_lock() { g.mu.lock(); }
void api_unlock() { g.mu.unlock(); }
} // namespace b200q4

extern "C" {

int b200q4_abi_version(void) { return B200Q4_ABI_VERSION; }

const char *b200q4_last_error(void) { return g.err.c_str(); }

int b200q4_init(int device)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (g.device == device && g.device >= 0) return B200Q4_OK;
    if (g.device >= 0) return fail(B200Q4_EINVAL, "already initialised on device %d", g.device);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(B200Q4_ENODEV, "no CUDA device visible; this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(B200Q4_EINVAL, "device %d out of range (0..%d)", device, n - 1);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(B200Q4_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&g.ev0));
    CK(cudaEventCreate(&g.ev1));
    g.sms = prop.multiProcessorCount;
    g.smem_optin = (int)prop.sharedMemPerBlockOptin;
    g.smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    if (const char *e = getenv("B200Q4_TWO_CTAS")) g.two_ctas = atoi(e) != 0;
    CK(cudaMalloc(&g.d_ticket, (2 + PIPE_MAX_CLASSES) * sizeof(int)));
    CK(cudaMemset(g.d_ticket, 0, (2 + PIPE_MAX_CLASSES) * sizeof(int)));
    CK(cudaMalloc(&g.d_done, sizeof(int)));
    CK(cudaMemset(g.d_done, 0, sizeof(int)));
    if (const char *e = getenv("B200Q4_PDL")) g.use_pdl = atoi(e) != 0;
    if (const char *e = getenv("B200Q4_TILE_MIN_M")) g.tile_min_m = atoi(e);
    if (const char *e = getenv("B200Q4_FUSE_MAX_K")) g.fuse_max_k = atoi(e);
    if (const char *e = getenv("B200Q4_PIPE_FLAGS")) g.pipe_flags = atoi(e);
    if (const char *e = getenv("B200Q4_SLICE_MIN_C")) g.slice_min_c = atoi(e) < 16 ? 16 : atoi(e);
    if (const char *e = getenv("B200Q4_ZERO_COPY")) g.zero_copy = atoi(e) != 0;
    if (const char *e = getenv("B200Q4_B8_MIN_STAGES")) g.b8_min_stages = atoi(e);
    if (const char *e = getenv("B200Q4_STATIC_TASKS")) g.static_tasks = atoi(e);
    if (const char *e = getenv("B200Q4_PREFETCH_STAGES")) g.prefetch_stages = atoi(e);
    if (const char *e = getenv("B200Q4_EXACT")) g.exact = atoi(e) != 0;
    g.init_gen++;
    if (cudaHostAlloc(reinterpret_cast<void **>(&g.h_flag), 64, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void **>(&g.d_flag), g.h_flag, 0) == cudaSuccess) {
        *g.h_flag = 0; g.hseq = 0;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        const char *sw = getenv("B200Q4_STREAM_WRITE");
        if ((!sw || atoi(sw) != 0) && cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess && fn)
            g.stream_write32 = reinterpret_cast<int (*)(cudaStream_t, unsigned long long, uint32_t, unsigned int)>(fn);
        else
            cudaGetLastError();
    } else {
        cudaGetLastError();
        g.h_flag = nullptr; g.d_flag = nullptr;          // no mapped memory: the copy path serves every call
    }
    if (const char *e = getenv("B200Q4_DBG")) if (atoi(e) != 0) {
        CK(cudaMalloc(&g.d_dbg, 256 * 16 * sizeof(unsigned long long)));
        CK(cudaMemset(g.d_dbg, 0, 256 * 16 * sizeof(unsigned long long)));
    }
    g.device = device;
    g.launches = 0;
    return B200Q4_OK;
}

void b200q4_shutdown(void)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (g.device < 0) return;
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.stream);
    if (g.h_flag) cudaFreeHost(g.h_flag);
    g.h_flag = nullptr; g.d_flag = nullptr;
    cudaFree(g.d_ticket); g.d_ticket = nullptr;
    cudaFree(g.d_done); g.d_done = nullptr;
    cudaFree(g.d_yloc); g.d_yloc = nullptr; g.yloc_cap = 0;
    cudaFree(g.d_exq); g.d_exq = nullptr; g.exq_cap = 0;
    cudaFree(g.d_exdeq); g.d_exdeq = nullptr; g.exdeq_cap = 0;
    cudaFree(g.d_aq); cudaFree(g.d_ap); cudaFree(g.d_partial); cudaFree(g.d_x); cudaFree(g.d_y); cudaFree(g.d_tmp);
    cudaFree(g.d_tpack); cudaFree(g.d_tpartial); cudaFree(g.d_tcounters);
    g.d_tpack = nullptr; g.d_tpartial = nullptr; g.d_tcounters = nullptr;
    g.tpack_cap = g.tpartial_cap = g.tcounters_cap = 0;
    g.d_aq = nullptr; g.d_ap = nullptr; g.d_partial = nullptr; g.d_x = nullptr; g.d_y = nullptr; g.d_tmp = nullptr;
    g.aq_cap = g.ap_cap = g.partial_cap = g.x_cap = g.y_cap = g.tmp_cap = 0;
    cudaEventDestroy(g.ev0); cudaEventDestroy(g.ev1);
    cudaStreamDestroy(g.stream);
    g.stream = nullptr;
    g.device = -1;
}

int b200q4_device_info(char *name, size_t cap, int *sm_count, size_t *hbm_bytes)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g.device));
    if (name && cap) { strncpy(name, prop.name, cap - 1); name[cap - 1] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
    return B200Q4_OK;
}

b200q4_weight *b200q4_weight_upload(int type, const void *rows, int N, int K, size_t row_stride)
{
    return upload_impl(type, rows, N, K, row_stride, false);
}

b200q4_weight *b200q4_weight_upload_rows(int type, const void *rows, int N, int K, size_t row_stride, int row0, int nrows)
{
    if (K < 0 || K % QK != 0 || row0 < 0 || nrows < 0 || (long)row0 + nrows > N) {
        fail(B200Q4_EINVAL, "weight_upload_rows: bad shard [%d,+%d) of %d rows", row0, nrows, N);
        return nullptr;
    }
    const size_t rb = (size_t)(K / QK) * (type == 0 ? 20 : 24);
    if (row_stride == 0) row_stride = rb;
    return upload_impl(type, rows ? (const char *)rows + (size_t)row0 * row_stride : nullptr, nrows, K, row_stride, false);
}

b200q4_weight *b200q4_weight_upload_transposed(int type, const void *rows, int N, int K, size_t row_stride)
{
    return upload_impl(type, rows, N, K, row_stride, true);
}

void b200q4_weight_free(b200q4_weight *w)
{
    if (!w) return;
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (g.hint_w == w) g.hint_w = nullptr;
    if (g.device >= 0) {
        cudaStreamSynchronize(g.stream);
        cudaFree(w->d_wt);
        cudaFree(w->d_counters);
    }
    delete w;
}

int b200q4_weight_rows(const b200q4_weight *w) { return w ? w->N : B200Q4_EINVAL; }
int b200q4_weight_cols(const b200q4_weight *w) { return w ? w->K : B200Q4_EINVAL; }
int b200q4_weight_type(const b200q4_weight *w) { return w ? w->type : B200Q4_EINVAL; }
size_t b200q4_weight_device_bytes(const b200q4_weight *w) { return w ? w->wt_bytes : 0; }

int b200q4_weight_download(const b200q4_weight *w, void *rows_out, size_t row_stride)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!w || !rows_out) return fail(B200Q4_EINVAL, "weight_download: bad argument");
    const size_t rb = (size_t)w->nb * (w->type == 0 ? 20 : 24);
    if (row_stride == 0) row_stride = rb;
    const int nrows = w->transposed ? w->K : w->N;
    if (nrows == 0 || rb == 0) return B200Q4_OK;
    if (w->transposed) {
        CK(cudaMemcpy2D(rows_out, row_stride, w->d_wt, rb, rb, nrows, cudaMemcpyDeviceToHost));
        return B200Q4_OK;
    }
    if (int rc = ensure(g.d_tmp, g.tmp_cap, rb * (size_t)nrows)) return rc;
    const long n = (long)w->N * w->nb;
    const int blocks = (int)((n + 255) / 256);
    if (w->type == 0) unpack_kernel<0><<<blocks, 256, 0, g.stream>>>(w->d_wt, w->N, w->nb, w->C, g.d_tmp);
    else              unpack_kernel<1><<<blocks, 256, 0, g.stream>>>(w->d_wt, w->N, w->nb, w->C, g.d_tmp);
    g.launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(rows_out, row_stride, g.d_tmp, rb, rb, nrows, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return B200Q4_OK;
}

int b200q4_mul_mat_dev(const b200q4_weight *w, const float *d_x, int M, size_t x_stride, float *d_y,
                       size_t y_stride, void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!w) return fail(B200Q4_EINVAL, "mul_mat_dev: NULL weight");
    if (x_stride == 0) x_stride = (size_t)w->K * 4;
    if (y_stride == 0) y_stride = (size_t)w->N * 4;
    return mul_mat_dev_locked(w, d_x, M, x_stride, d_y, y_stride, stream ? (cudaStream_t)stream : g.stream, nullptr);
}

int b200q4_mul_mat_dev_allgather(const b200q4_weight *w, const float *d_x, int M, size_t x_stride,
                                 float *const *y_peers, int world, int rank, int col0, size_t y_stride,
                                 uint32_t *const *flag_peers, uint32_t *seq_counter, void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!w || !y_peers || !flag_peers || !seq_counter || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || col0 < 0)
        return fail(B200Q4_EINVAL, "mul_mat_dev_allgather: bad argument (world=%d rank=%d)", world, rank);
    if (x_stride == 0) x_stride = (size_t)w->K * 4;
    if (y_stride < (size_t)(col0 + w->N) * 4) return fail(B200Q4_EINVAL, "mul_mat_dev_allgather: y_stride smaller than col0 + shard rows");
    cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
    PeerOut po{};
    for (int r = 0; r < world; r++) {
        if (!y_peers[r] || !y_peers[world + r] || !flag_peers[r]) return fail(B200Q4_EINVAL, "mul_mat_dev_allgather: NULL peer pointer");
        po.y[0][r] = y_peers[r]; po.y[1][r] = y_peers[world + r]; po.flag[r] = flag_peers[r];
    }
    po.local_flags = flag_peers[rank]; po.seq_ctr = seq_counter;
    po.done = g.d_done; po.world = world; po.rank = rank; po.col0 = col0;
    if (M == 0 || w->N == 0 || world == 1) {
        // nothing to exchange: a single rank runs the plain kernel on the buffer the sequence number selects
        // (the host mirrors the counter: exchange s uses buffer s & 1), then publishes like any other launch
        if (world == 1 && M > 0 && w->N > 0) {
            uint32_t seq_now = 0;
            CK(cudaMemcpyAsync(&seq_now, seq_counter, sizeof seq_now, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (int rc = mul_mat_dev_locked(w, d_x, M, x_stride, y_peers[((seq_now + 1) & 1u) ? world : 0] + col0, y_stride, st, nullptr)) return rc;
        }
        scatter_slice_kernel<<<1, 32, 0, st>>>(nullptr, 0, 0, 0, y_stride / 4, po);
        g.launches++;
        CK(cudaGetLastError());
        return B200Q4_OK;
    }
    return mul_mat_dev_locked(w, d_x, M, x_stride, nullptr, y_stride, st, nullptr, &po);
}

int b200q4_wait_flags(const uint32_t *local_flags, int world, const uint32_t *seq_counter, void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!local_flags || !seq_counter || world < 1 || world > MAX_PEERS) return fail(B200Q4_EINVAL, "wait_flags: bad argument");
    wait_flags_kernel<<<1, 32, 0, stream ? (cudaStream_t)stream : g.stream>>>(local_flags, world, seq_counter);
    g.launches++;
    CK(cudaGetLastError());
    return B200Q4_OK;
}

int b200q4_mul_mat(const b200q4_weight *w, const float *x, int M, size_t x_stride, float *y, size_t y_stride)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!w || w->transposed || M < 0 || (M > 0 && (!x || !y))) return fail(B200Q4_EINVAL, "mul_mat: bad argument");
    if (x_stride == 0) x_stride = (size_t)w->K * 4;
    if (y_stride == 0) y_stride = (size_t)w->N * 4;
    if (x_stride < (size_t)w->K * 4 || y_stride < (size_t)w->N * 4) return fail(B200Q4_EINVAL, "mul_mat: stride smaller than a row");
    if (M == 0 || w->N == 0) return B200Q4_OK;
    if (int rc = ensure(g.d_x, g.x_cap, (size_t)M * w->K * 4)) return rc;
    const bool decode = !(g.tile_min_m > 0 && M >= g.tile_min_m);
    if (g.zero_copy && decode && g.h_flag && reinterpret_cast<uintptr_t>(x) % 16 == 0 && x_stride % 16 == 0 && y_stride % 4 == 0) {
        const float *xa = static_cast<const float *>(pinned_alias(x));
        float *ya = xa ? static_cast<float *>(pinned_alias(y)) : nullptr;
        if (xa && ya) {
            const long total = (long)M * (w->K / 4);
            const int blocks = (int)((total + 255) / 256);
            CK(launch_pdl(fetch_rows_kernel, blocks, 256, g.stream, xa, x_stride / 4, M, w->K, g.d_x));
            g.launches++;
            if (int rc = mul_mat_dev_locked(w, g.d_x, M, (size_t)w->K * 4, ya, y_stride, g.stream, nullptr)) return rc;
            const uint32_t seq = ++g.hseq;
            if (!g.stream_write32 || g.stream_write32(g.stream, (unsigned long long)(uintptr_t)g.d_flag, seq, 0) != 0) {
                g.stream_write32 = nullptr;              // not supported here: a one-thread kernel writes the word
                signal_host_kernel<<<1, 1, 0, g.stream>>>(g.d_flag, seq);
                g.launches++;
                CK(cudaGetLastError());
            }
            return wait_host_flag(seq);
        }
    }
    if (int rc = ensure(g.d_y, g.y_cap, (size_t)M * w->N * 4)) return rc;
    CK(cudaMemcpy2DAsync(g.d_x, (size_t)w->K * 4, x, x_stride, (size_t)w->K * 4, M, cudaMemcpyHostToDevice, g.stream));
    if (int rc = mul_mat_dev_locked(w, g.d_x, M, (size_t)w->K * 4, g.d_y, (size_t)w->N * 4, g.stream, nullptr)) return rc;
    CK(cudaMemcpy2DAsync(y, y_stride, g.d_y, (size_t)w->N * 4, (size_t)w->N * 4, M, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return B200Q4_OK;
}

int b200q4_mul_mat_wdata(const b200q4_weight *w, const void *wdata, int M, float *y, size_t y_stride)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!w || w->transposed || M < 0 || (M > 0 && (!wdata || !y))) return fail(B200Q4_EINVAL, "mul_mat_wdata: bad argument");
    if (y_stride == 0) y_stride = (size_t)w->N * 4;
    if (y_stride < (size_t)w->N * 4) return fail(B200Q4_EINVAL, "mul_mat_wdata: stride smaller than a row");
    if (M == 0 || w->N == 0) return B200Q4_OK;
    const size_t rb = (size_t)w->nb * (w->type == 0 ? 20 : 24);
    if (int rc = ensure(g.d_tmp, g.tmp_cap, rb * (size_t)M)) return rc;
    if (int rc = ensure(g.d_y, g.y_cap, (size_t)M * w->N * 4)) return rc;
    CK(cudaMemcpyAsync(g.d_tmp, wdata, rb * (size_t)M, cudaMemcpyHostToDevice, g.stream));
    if (int rc = mul_mat_dev_locked(w, nullptr, M, (size_t)w->K * 4, g.d_y, (size_t)w->N * 4, g.stream, g.d_tmp)) return rc;
    CK(cudaMemcpy2DAsync(y, y_stride, g.d_y, (size_t)w->N * 4, (size_t)w->N * 4, M, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return B200Q4_OK;
}

int b200q4_quantize_rows_dev(int type, const float *d_x, int M, int K, size_t x_stride, void *d_out, void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if ((type != 0 && type != 1) || M < 0 || K < 0 || K % QK != 0 || (M > 0 && K > 0 && (!d_x || !d_out)))
        return fail(B200Q4_EINVAL, "quantize_rows: bad argument");
    if (x_stride == 0) x_stride = (size_t)K * 4;
    if (x_stride % 4) return fail(B200Q4_EINVAL, "quantize_rows: stride must be a multiple of 4 bytes");
    if (M == 0 || K == 0) return B200Q4_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
    const int nb = K / QK;
    const long warps = (long)nb * M;
    const int blocks = (int)((warps * 32 + 255) / 256);
    if (type == 0) quantize_rows_ref_kernel<0><<<blocks, 256, 0, st>>>(d_x, x_stride / 4, M, nb, (uint8_t *)d_out);
    else           quantize_rows_ref_kernel<1><<<blocks, 256, 0, st>>>(d_x, x_stride / 4, M, nb, (uint8_t *)d_out);
    g.launches++;
    CK(cudaGetLastError());
    return B200Q4_OK;
}

int b200q4_quantize_rows(int type, const float *x, int M, int K, size_t x_stride, void *out)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if ((type != 0 && type != 1) || M < 0 || K < 0 || K % QK != 0 || (M > 0 && K > 0 && (!x || !out)))
        return fail(B200Q4_EINVAL, "quantize_rows: bad argument");
    if (M == 0 || K == 0) return B200Q4_OK;
    if (x_stride == 0) x_stride = (size_t)K * 4;
    const size_t rb = (size_t)(K / QK) * (type == 0 ? 20 : 24);
    if (int rc = ensure(g.d_x, g.x_cap, (size_t)M * K * 4)) return rc;
    if (int rc = ensure(g.d_tmp, g.tmp_cap, rb * (size_t)M)) return rc;
    CK(cudaMemcpy2DAsync(g.d_x, (size_t)K * 4, x, x_stride, (size_t)K * 4, M, cudaMemcpyHostToDevice, g.stream));
    if (int rc = b200q4_quantize_rows_dev(type, g.d_x, M, K, (size_t)K * 4, g.d_tmp, g.stream)) return rc;
    CK(cudaMemcpyAsync(out, g.d_tmp, rb * (size_t)M, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return B200Q4_OK;
}

int b200q4_quantize_weights(int type, const float *src, long long n_rows, int K, size_t src_stride, void *dst, long long *hist)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if ((type != 0 && type != 1) || n_rows < 0 || K < 0 || K % QK != 0 || (n_rows > 0 && K > 0 && (!src || !dst)))
        return fail(B200Q4_EINVAL, "quantize_weights: bad argument");
    if (src_stride == 0) src_stride = (size_t)K * 4;
    if (src_stride % 4 || src_stride < (size_t)K * 4) return fail(B200Q4_EINVAL, "quantize_weights: bad stride");
    unsigned long long *d_hist = nullptr;
    CK(cudaMalloc(&d_hist, 16 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_hist, 0, 16 * sizeof(unsigned long long), g.stream);
    const int nb = K / QK;
    const size_t rb = (size_t)nb * (type == 0 ? 20 : 24);
    int rc = B200Q4_OK;
    if (e == cudaSuccess && n_rows > 0 && K > 0) {
        long long slab = (64ll << 20) / K;               // rows per slab: <= 256 MB of fp32 in flight
        if (slab < 1) slab = 1;
        if (slab > n_rows) slab = n_rows;
        if (!(rc = ensure(g.d_x, g.x_cap, (size_t)slab * K * 4)) && !(rc = ensure(g.d_tmp, g.tmp_cap, rb * (size_t)slab))) {
            for (long long r0 = 0; r0 < n_rows && e == cudaSuccess; r0 += slab) {
                const long long nr = n_rows - r0 < slab ? n_rows - r0 : slab;
                e = cudaMemcpy2DAsync(g.d_x, (size_t)K * 4, (const char *)src + (size_t)r0 * src_stride, src_stride, (size_t)K * 4, (size_t)nr,
                                      cudaMemcpyHostToDevice, g.stream);
                if (e != cudaSuccess) break;
                const long long warps = nr * nb;
                const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
                if (type == 0) quantize_weights_kernel<0><<<blocks, 256, 0, g.stream>>>(g.d_x, (size_t)K, (long)nr, nb, g.d_tmp, d_hist);
                else           quantize_weights_kernel<1><<<blocks, 256, 0, g.stream>>>(g.d_x, (size_t)K, (long)nr, nb, g.d_tmp, d_hist);
                g.launches++;
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaMemcpyAsync((char *)dst + (size_t)r0 * rb, g.d_tmp, rb * (size_t)nr, cudaMemcpyDeviceToHost, g.stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);       // the staging buffers are reused by the next slab
            }
        }
    }
    unsigned long long h[16] = {0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_hist, sizeof h, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    cudaFree(d_hist);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(B200Q4_ECUDA, "quantize_weights: %s", cudaGetErrorString(e));
    if (hist) for (int i = 0; i < 16; i++) hist[i] += (long long)h[i];
    return B200Q4_OK;
}

int b200q4_mul_mat_transposed(const b200q4_weight *wt, const float *x, int M, size_t x_stride, float *y, size_t y_stride)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!wt || !wt->transposed || M < 0 || (M > 0 && (!x || !y))) return fail(B200Q4_EINVAL, "mul_mat_transposed: bad argument");
    const int N = wt->N, K = wt->K;
    if (x_stride == 0) x_stride = (size_t)K * 4;
    if (y_stride == 0) y_stride = (size_t)N * 4;
    if (M == 0 || N == 0) return B200Q4_OK;
    if (int rc = ensure(g.d_x, g.x_cap, (size_t)M * (K > 0 ? K : 1) * 4)) return rc;
    if (int rc = ensure(g.d_y, g.y_cap, (size_t)M * N * 4)) return rc;
    if (K > 0) CK(cudaMemcpy2DAsync(g.d_x, (size_t)K * 4, x, x_stride, (size_t)K * 4, M, cudaMemcpyHostToDevice, g.stream));
    dim3 grid((N + 127) / 128, M);
    if (wt->type == 0) mul_mat_transposed_kernel<0><<<grid, 128, 0, g.stream>>>(wt->d_wt, N, K, g.d_x, K, M, g.d_y, N);
    else               mul_mat_transposed_kernel<1><<<grid, 128, 0, g.stream>>>(wt->d_wt, N, K, g.d_x, K, M, g.d_y, N);
    g.launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(y, y_stride, g.d_y, (size_t)N * 4, (size_t)N * 4, M, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    return B200Q4_OK;
}

// ---- model files (row f-3, first half): the reference's on-disk format, Q4 tensors straight from the mapping to HBM ----
struct b200q4_model_tensor { std::string name; int n_dims, ftype, ne[2]; size_t offset, nbytes; b200q4_weight *w; };
struct b200q4_model {
    std::vector<int> header;
    std::vector<b200q4_model_tensor> tensors;
    void *map = nullptr; size_t map_size = 0;
};

b200q4_model *b200q4_model_load(const char *path, int n_header_ints, int vocab_has_count, int upload)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (!path || n_header_ints < 1 || n_header_ints > 32) { fail(B200Q4_EINVAL, "model_load: bad argument"); return nullptr; }
    if (upload && need_init()) return nullptr;
    const int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) { if (fd >= 0) close(fd); fail(B200Q4_EINVAL, "model_load: cannot open '%s'", path); return nullptr; }
    const size_t size = (size_t)st.st_size;
    void *map = size ? mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (!map || map == MAP_FAILED) { fail(B200Q4_EINVAL, "model_load: cannot map '%s'", path); return nullptr; }
    const unsigned char *b = static_cast<const unsigned char *>(map);
    b200q4_model *m = new b200q4_model();
    m->map = map; m->map_size = size;
    size_t off = 0;
    auto bail = [&](const char *what) { fail(B200Q4_EINVAL, "model_load: %s (offset %zu of '%s')", what, off, path); b200q4_model_free(m); return (b200q4_model *)nullptr; };
    auto rd32 = [&](int32_t &v) { if (off + 4 > size) return false; memcpy(&v, b + off, 4); off += 4; return true; };
    int32_t magic = 0;
    if (!rd32(magic) || magic != 0x67676d6c) return bail("bad magic");                  // vsim.cpp:121-126
    for (int i = 0; i < n_header_ints; i++) { int32_t v; if (!rd32(v)) return bail("truncated header"); m->header.push_back(v); }
    int32_t n_vocab = m->header[0];                                                        // first header int in every family
    if (vocab_has_count && (!rd32(n_vocab) || n_vocab != m->header[0])) return bail("bad vocab size");   // quantize_gptj.cpp:101-110
    if (n_vocab < 0) return bail("bad vocab size");
    for (int i = 0; i < n_vocab; i++) {                                                    // vsim.cpp:163-173
        int32_t len;
        if (!rd32(len) || len < 0 || off + (size_t)len > size) return bail("truncated vocabulary");
        off += (size_t)len;
    }
    while (off < size) {                                                                   // vsim.cpp:375-448
        int32_t n_dims, length, ftype;
        if (!rd32(n_dims) || !rd32(length) || !rd32(ftype)) return bail("truncated tensor header");
        if (n_dims < 1 || n_dims > 2 || length < 0 || ftype < 0 || ftype > 3) return bail("bad tensor header");
        b200q4_model_tensor t;
        t.n_dims = n_dims; t.ftype = ftype; t.ne[0] = t.ne[1] = 1; t.w = nullptr;
        for (int i = 0; i < n_dims; i++) {
            int32_t v;
            if (!rd32(v)) return bail("truncated tensor shape");
            if (v < 0) return bail("bad tensor shape");
            t.ne[i] = v;
        }
        if ((size_t)length > size - off) return bail("truncated tensor name");
        t.name.assign(reinterpret_cast<const char *>(b + off), (size_t)length);
        off += (size_t)length;
        const size_t nel = (size_t)t.ne[0] * (size_t)t.ne[1];                      // (both < 2^31: no wrap)
        if (ftype >= 2 && t.ne[0] % QK) return bail("quantized tensor whose row length is not a multiple of 32");
        if (nel > (size_t)-1 / 4) return bail("tensor too large");                  // nel * bytes per element must not wrap either
        t.nbytes = ftype == 0 ? nel * 4 : ftype == 1 ? nel * 2 : nel / QK * (ftype == 2 ? 20 : 24);
        t.offset = off;
        if (t.nbytes > size - off) return bail("truncated tensor data");
        off += t.nbytes;
        m->tensors.push_back(t);
    }
    if (upload)
        for (auto &t : m->tensors)
            if (t.n_dims == 2 && t.ftype >= 2) {
                t.w = b200q4_weight_upload(t.ftype - 2, b + t.offset, t.ne[1], t.ne[0], 0);   // rows = ne[1], K = ne[0]
                if (!t.w) { b200q4_model_free(m); return nullptr; }
            }
    return m;
}

int b200q4_model_header(const b200q4_model *m, int i) { return (m && i >= 0 && i < (int)m->header.size()) ? m->header[(size_t)i] : B200Q4_EINVAL; }
int b200q4_model_tensor_count(const b200q4_model *m) { return m ? (int)m->tensors.size() : B200Q4_EINVAL; }

int b200q4_model_tensor_info(const b200q4_model *m, int i, const char **name, int *n_dims, int *ftype, int *ne0, int *ne1,
                             size_t *file_offset, size_t *nbytes)
{
    if (!m || i < 0 || i >= (int)m->tensors.size()) return fail(B200Q4_EINVAL, "model_tensor_info: bad index");
    const b200q4_model_tensor &t = m->tensors[(size_t)i];
    if (name) *name = t.name.c_str();
    if (n_dims) *n_dims = t.n_dims;
    if (ftype) *ftype = t.ftype;
    if (ne0) *ne0 = t.ne[0];
    if (ne1) *ne1 = t.ne[1];
    if (file_offset) *file_offset = t.offset;
    if (nbytes) *nbytes = t.nbytes;
    return B200Q4_OK;
}

b200q4_weight *b200q4_model_weight(const b200q4_model *m, int i)
{
    return (m && i >= 0 && i < (int)m->tensors.size()) ? m->tensors[(size_t)i].w : nullptr;
}

const void *b200q4_model_tensor_data(const b200q4_model *m, int i)
{
    return (m && i >= 0 && i < (int)m->tensors.size()) ? static_cast<const unsigned char *>(m->map) + m->tensors[(size_t)i].offset : nullptr;
}

void b200q4_model_free(b200q4_model *m)
{
    if (!m) return;
    for (auto &t : m->tensors) if (t.w) b200q4_weight_free(t.w);
    if (m->map) munmap(m->map, m->map_size);
    delete m;
}

int b200q4_hint_next_weight(const b200q4_weight *next, int M_next)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    g.hint_w = next;
    g.hint_m = M_next > 0 ? M_next : 1;
    return B200Q4_OK;
}

int b200q4_set_exact(int on)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    const int old = g.exact ? 1 : 0;
    g.exact = on != 0;
    return old;
}

int b200q4_exact_enabled(void) { return g.exact ? 1 : 0; }

int b200q4_sync(void *stream)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    CK(cudaStreamSynchronize(stream ? (cudaStream_t)stream : g.stream));
    return B200Q4_OK;
}

void *b200q4_host_alloc(size_t bytes)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (need_init()) return nullptr;
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        fail(B200Q4_ENOMEM, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    void *dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, p, 0) == cudaSuccess && dev)
        g.host_blocks.push_back({(const char *)p, bytes ? bytes : 1, (char *)dev});
    else
        cudaGetLastError();
    return p;
}

void b200q4_host_free(void *p)
{
    if (!p) return;
    {
        std::lock_guard<std::recursive_mutex> lk(g.mu);
        for (size_t i = 0; i < g.host_blocks.size(); i++)
            if (g.host_blocks[i].base == (const char *)p) { g.host_blocks.erase(g.host_blocks.begin() + (long)i); break; }
    }
    cudaFreeHost(p);
}

// development aid (not part of b200q4.h): copies the [256][16] timeline of the last decode launch
int b200q4_debug_timeline(unsigned long long *out)
{
    if (!g.d_dbg) return B200Q4_EINVAL;
    cudaStreamSynchronize(g.stream);
    return cudaMemcpy(out, g.d_dbg, 256 * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : B200Q4_ECUDA;
}

// development / test knob (not part of b200q4.h): smallest M that runs the tcgen05 tile kernel
// (0 = never); returns the previous value.
int b200q4_debug_set_tile_min_m(int m)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    const int old = g.tile_min_m;
    g.tile_min_m = m;
    return old;
}

// development / test knob (not part of b200q4.h): longest row (K) whose activations the decode kernel
// quantizes in its own prologue; longer rows go through the pack kernel.  Returns the previous value.
int b200q4_debug_set_fuse_max_k(int k)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    const int old = g.fuse_max_k;
    g.fuse_max_k = k;
    return old;
}

// development / test knob (not part of b200q4.h): tiles longer than this many units are K-sliced
// (0 = the default rule: 64 for M <= 2, 32 for 3 <= M <= 8).  Returns the previous value.
int b200q4_debug_set_slice_min_c(int c)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    const int old = g.slice_min_c;
    g.slice_min_c = c;
    return old;
}

// development / test aid (not part of b200q4.h): the decode launch plan for a (type, N, K, M) problem, computed on the host
// alone (no device needed: without one the shared-memory limits default to the B200's).  out[0..11] = MP, PP, TS, ntasks,
// grid, stages, nbp, gnbp, b8, m1, C, n_tiles; returns the dynamic shared-memory bytes (or a negative error).
long long b200q4_debug_plan(int type, int N, int K, int M, int *out)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if ((type != 0 && type != 1) || N < 1 || K < QK || K % QK || M < 1 || M > 8 || !out) return B200Q4_EINVAL;
    b200q4_weight w{};
    w.type = type; w.N = N; w.K = K; w.nb = K / QK; w.transposed = false;
    w.C = (w.nb + CHUNK_BLKS - 1) / CHUNK_BLKS;
    w.n_tiles = (N + TILE_ROWS - 1) / TILE_ROWS;
    const PipePlan pp = pipe_plan(&w, M, true);
    if (pp.MP < 1) return B200Q4_EINVAL;
    const int slots = (g.sms > 0 ? g.sms : 148) * (pp.ncw == 8 ? 2 : 1);
    const int vals[12] = {pp.MP, pp.PP, pp.TS, pp.ntasks, pp.ntasks < slots ? pp.ntasks : slots, pp.stages, pp.nbp, pp.gnbp,
                          pp.b8 ? 1 : 0, pp.m1 ? 1 : 0, w.C, w.n_tiles};
    for (int i = 0; i < 12; i++) out[i] = vals[i];
    return (long long)pp.smem;
}

uint64_t b200q4_launch_count(void) { return g.launches; }
void b200q4_launch_count_reset(void) { g.launches = 0; }

const char *b200q4_kernel_for(int type, int M, int N, int K)
{
    (void)N;
    if ((type != 0 && type != 1) || K <= 0 || K % QK) return "invalid";
    if (g.exact) return type == 0 ? "q4_0_exact" : "q4_1_exact";
    if (g.tile_min_m > 0 && M >= g.tile_min_m) return type == 0 ? "q4_0_tile_tcgen05" : "q4_1_tile_tcgen05";
    const int mp = (M <= 2) ? 2 : 8;
    return type == 0 ? (mp == 2 ? "q4_0_gemv_pipe_chain4" : "q4_0_gemv_pipe_m8") : (mp == 2 ? "q4_1_gemv_pipe_chain4" : "q4_1_gemv_pipe_m8");
}

double b200q4_time_mul_mat_dev(b200q4_weight *const *ws, int nw, const float *d_x, int M, float *d_y, int iters)
{
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (int rc = need_init()) return rc;
    if (!ws || nw < 1 || iters < 1) return fail(B200Q4_EINVAL, "time_mul_mat_dev: bad argument");
    if (cudaEventRecord(g.ev0, g.stream) != cudaSuccess) return fail(B200Q4_ECUDA, "event record failed");
    for (int i = 0; i < iters; i++) {
        const b200q4_weight *w = ws[i % nw];
        g.hint_w = ws[(i + 1) % nw]; g.hint_m = M;      // the caller's rotation is known: hint the next tensor
        if (int rc = mul_mat_dev_locked(w, d_x, M, (size_t)w->K * 4, d_y, (size_t)w->N * 4, g.stream, nullptr)) return rc;
    }
    if (cudaEventRecord(g.ev1, g.stream) != cudaSuccess) return fail(B200Q4_ECUDA, "event record failed");
    if (cudaEventSynchronize(g.ev1) != cudaSuccess) return fail(B200Q4_ECUDA, "event sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g.ev0, g.ev1);
    return (double)ms / iters;
}

} // extern "C"
