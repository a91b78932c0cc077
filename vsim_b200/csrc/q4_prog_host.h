// q4_prog_host.h -- host side of the decode program kernel (q4_prog.cuh): building a work-list, launching it.
// Internal to libb200q4.so; the C ABI on top of it is b200q4_chain_* (include/b200q4.h) and b200neox_eval.
#pragma once
#include "q4_internal.h"
#include "q4_prog.cuh"

#include <vector>

namespace b200q4 {

// K-slicing geometry of one M = 1 mul_mat inside a program (the rule q4_api.cu's slice_geometry applies for M <= 2)
struct PgGeo { int PP, TS, ntasks; };
PgGeo pg_geometry(const b200q4_weight *w);
PgGeo pg_geometry_dims(int N, int K);               // the same from the tensor's dimensions alone

struct ProgBuild {
    int qt = -1;                         // quant type shared by every mul_mat of the program
    std::vector<PgOp> ops;
    int n_tk = 0, n_cnt = 0;             // ticket counters / attention head counters handed out so far
    int rot_acc = 0;                     // running rotation of the static task assignment
    int last_mv = -1;                    // previous mul_mat op (pack reuse)
    int n_packs = 0;                     // activation packs so far
    int last_pack_op = -1;               // first op of the latest activation pack
    bool last_pack_cobuilt = false;      // ... which is built in the pause of an earlier op (no barrier at its own start)
    std::vector<char> pack_selfc;        // per pack: every stage of its ops covers the same 128 quant blocks (warp w reads blocks 8w .. 8w+7)
    int last_res_pack = -1, last_res_op = -1;   // pack / op during whose construction the residual stream was last updated
    int e_cap = 4, nbp_cap = 8, att_cap = 4;
    // Appends Y = W . quantize(in(sources)); returns the op index or a negative error.  For a K-sliced weight
    // (pg_geometry(w).PP > 1) y receives PP slices of raw sums, y_slice elements apart, and bias must be NULL.
    // y_tag: y holds {value, tag} pairs (8 bytes per element) that later ops of the list read as PG_SRC_TAG sources;
    // same_pack: the activation is that of the previous mul_mat op (same input, mode, layer-norm weights): built once.
    int add_matvec(const b200q4_weight *w, int dep, int in_mode, int res, const PgSrc *src, int nsrc,
                   const float *ln_w, const float *ln_b, const uint16_t *gelu_tab, float *y, const float *bias, int y_slice,
                   bool y_tag = false, bool same_pack = false, float *y_plain = nullptr);
    // att_part [H][chunks][d], att_sc [H][n_ctx + chunks], att_sum [H][chunks] (chunks = ceil(n_ctx / PG_ATT_CHUNK)): scratch
    // q, k, v: tagged vectors written by ops tq, tk, tv; att_out: tagged
    int add_attn(int dep, const float *q, const float *k, const float *v, int tq, int tk, int tv, float *mem_k, float *mem_v, float *att_out,
                 float *att_part, float *att_sc, double *att_sum, const double2 *rope, const uint16_t *exp_tab,
                 int H, int d, int n_rot, int n_ctx, float scale);
    int add_sum(int dep, const PgSrc &src, float *y, int N);
};

struct Prog {
    PgOp *d_ops = nullptr;
    int   n_ops = 0, n_matvec = 0, first_mv = 0;
    int  *d_ctr = nullptr;
    int   n_ctr = 0;
    int  *d_dyn = nullptr;               // [0] token id, [1] n_past, [2] launch epoch
    int  *d_err = nullptr;
    unsigned long long *d_dbg = nullptr; // timeline (development builds)
    int   world = 1, rank = 0;           // row-sharded programs: this rank of `world`
    int  *ctr_peer[8] = {nullptr};       // ... and every rank's counter array (peer-mapped; entry `rank` is d_ctr, not owned)
    bool  own_ctr = true;
    int   qt = 0, stages = 0, grid = 0, e_cap = 0, nbp_cap = 0, att_cap = 0;
    size_t smem = 0;
};

inline PgSrc pg_src_f32(const float *p, int np = 1, int stride = 0, float scale = 1.0f) { return PgSrc{p, nullptr, np, stride, scale, PG_SRC_F32}; }

int  prog_finalize(const ProgBuild &b, Prog &out);      // uploads the list, sizes shared memory, allocates the counters
// sharded: the counters live in caller-provided peer-mapped memory (ctr_peers[r], prog_counter_ints(b) ints each, zeroed)
int  prog_finalize_sharded(const ProgBuild &b, Prog &out, int rank, int world, int *const *ctr_peers);
int  prog_counter_ints(const ProgBuild &b);
int  prog_launch(const Prog &pg, cudaStream_t st);      // one cooperative launch of q4_prog_kernel
void prog_free(Prog &pg);

} // namespace b200q4
