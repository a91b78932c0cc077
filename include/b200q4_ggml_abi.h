/*
 * b200q4_ggml_abi.h -- layout mirror of the two reference structs that cross
 * the accelerator hook, for builds that do not have the reference's ggml.h on
 * the include path.  Nothing here is copied code: it states the x86-64 ABI of
 *
 *   struct ggml_tensor          ggml.h:275-305   (160 bytes: type@0 n_dims@4
 *                               ne@8 nb@24 op@56 src0@72 src1@80 n_tasks@120 data@144)
 *   struct ggml_compute_params  ggml.c:1052-1060, re-declared at imax.c:1115-1121
 *   enum ggml_task_type         ggml.c:1046-1050, imax.c:1109-1113
 *
 * The reference itself shares the second struct between ggml.c and imax.c by
 * re-declaring it, so a re-declaration is the reference's own convention.
 * tests/test_abi.py checks these offsets against the real header through
 * oracle/_ref/libggml_ref.so (ref_tensor_layout), and imax_b200.c
 * static-asserts them when it is compiled with -DB200Q4_HAVE_GGML_H.
 */
#ifndef B200Q4_GGML_ABI_H
#define B200Q4_GGML_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#define B200Q4_GGML_MAX_DIMS 4
#define B200Q4_GGML_MAX_OPT  4

enum b200q4_ggml_task_type {
    B200Q4_GGML_TASK_INIT = 0,
    B200Q4_GGML_TASK_COMPUTE,
    B200Q4_GGML_TASK_FINALIZE,
};

struct b200q4_ggml_compute_params {
    int    type;            /* enum ggml_task_type */
    int    ith, nth;
    size_t wsize;
    void  *wdata;           /* INIT phase leaves the quantized src1 rows here */
};

struct b200q4_ggml_tensor {
    int     type;           /* enum ggml_type: 0 = Q4_0, 1 = Q4_1, 6 = F32 */
    int     n_dims;
    int     ne[B200Q4_GGML_MAX_DIMS];
    size_t  nb[B200Q4_GGML_MAX_DIMS];
    int     op;
    bool    is_param;
    void   *grad;
    void   *src0;
    void   *src1;
    void   *opt[B200Q4_GGML_MAX_OPT];
    int     n_tasks;
    int     perf_runs;
    int64_t perf_cycles;
    int64_t perf_time_us;
    void   *data;
    char    padding[8];
};

#endif
