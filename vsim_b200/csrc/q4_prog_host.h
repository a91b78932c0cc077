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

struct ProgBuild {
    int qt = -1;                         // quant type shared by every mul_mat of the program
    std::vector<PgOp> ops;
    int n_tk = 0, n_cnt = 0;             // ticket counters / attention head counters handed out so far
    int e_cap = 4, nbp_cap = 8, att_cap = 4;
    // Appends Y = W . quantize(in(sources)); returns the op index or a negative error.  For a K-sliced weight
    // (pg_geometry(w).PP > 1) y receives PP slices of raw sums, y_slice floats apart, and bias must be NULL.
    int add_matvec(const b200q4_weight *w, int dep, int in_mode, int res, const PgSrc *src, int nsrc,
                   const float *ln_w, const float *ln_b, const uint16_t *gelu_tab, float *y, const float *bias, int y_slice);
    int add_attn(int dep, const float *q, const float *k, const float *v, float *mem_k, float *mem_v, float *att_out,
                 float *att_part, const double2 *rope, int H, int d, int n_rot, int n_ctx, float scale);
    int add_sum(int dep, const PgSrc &src, float *y, int N);
};

struct Prog {
    PgOp *d_ops = nullptr;
    int   n_ops = 0, n_matvec = 0;
    int  *d_ctr = nullptr;
    int   n_ctr = 0;
    int  *d_dyn = nullptr;               // [0] token id, [1] n_past
    int  *d_err = nullptr;
    int   qt = 0, stages = 0, grid = 0, e_cap = 0, nbp_cap = 0, att_cap = 0;
    size_t smem = 0;
};

inline PgSrc pg_src_f32(const float *p, int np = 1, int stride = 0, float scale = 1.0f) { return PgSrc{p, nullptr, np, stride, scale, PG_SRC_F32}; }

int  prog_finalize(const ProgBuild &b, Prog &out);      // uploads the list, sizes shared memory, allocates the counters
int  prog_launch(const Prog &pg, cudaStream_t st);      // one cooperative launch of q4_prog_kernel
void prog_free(Prog &pg);

} // namespace b200q4
