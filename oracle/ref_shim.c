/*
 * ref_shim.c -- glue that lets the UNMODIFIED reference sources under
 * /root/reference be linked into a shared library / binary in this image.
 * TEST INFRASTRUCTURE ONLY (see q4_oracle.c header).  This file is ours; it
 * contains no reference code.  It is compiled together with the reference's
 * ggml.c / monitor.c / imax.c by oracle/Makefile, outputs go to oracle/_ref/.
 *
 *  - the seven X11 entry points of xdisp.c (no X11 headers in this image;
 *    the viewer is disabled at vsim.cpp:962 anyway),
 *  - ggml_time_init / ggml_time_us, declared at ggml.h:339-341 but defined
 *    nowhere in the fork (only the offline quantize_*.cpp tools call them),
 *  - with -DREF_SHIM_LIB: `int NTHREAD` (monitor.c:32 expects the program to
 *    define it, vsim.cpp:922 does) and ref_mul_mat(), a harness that drives
 *    the reference's public ggml.h API (ggml_new_tensor_2d, ggml_mul_mat,
 *    ggml_build_forward, ggml_graph_compute) so tests and the CPU baseline
 *    run the reference's own operator, scheduler and hook.
 */
#include <signal.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

void x11_open(int m) { (void)m; }
void x11_update(void) {}
void BGR_to_X(void) {}
void BOX_to_X(void) {}
int  x11_checkevent(void) { return 0; }
void x11_softu64_dist(float x, float y) { (void)x; (void)y; }
void x11_softu64_update(void) {}

#ifndef REF_SHIM_NO_TIME
void ggml_time_init(void) {}
int64_t ggml_time_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (int64_t)ts.tv_sec * 1000000 + ts.tv_nsec / 1000;
}
#endif

#ifdef REF_SHIM_LIB
#include "ggml.h"

int NTHREAD = 1;

/* One Q4 mul_mat through the reference's own graph executor.
 * type: GGML_TYPE_Q4_0 (0) / GGML_TYPE_Q4_1 (1).  W: N rows of quantized bytes.
 * X: [M][K] fp32 contiguous.  Y: [M][N] fp32 contiguous.  reps >= 1 repeats
 * ggml_graph_compute on the same graph (for timing); returns the best
 * wall time of one graph_compute in microseconds, or -1 on error. */
double ref_mul_mat(int type, const void *W, int N, int K,
                   const float *X, int M, float *Y, int n_threads, int reps)
{
    const size_t wbytes = (size_t)N * (K / 32) * (type == 0 ? 20 : 24);
    const size_t mem = wbytes + (size_t)M * K * 4 + (size_t)M * N * 4
                     + (size_t)M * (K / 32) * 24 + 64 * 64 + (1u << 20);
    struct ggml_init_params ip = { mem, NULL };
    struct ggml_context *ctx = ggml_init(ip);
    if (!ctx) return -1.0;
    struct ggml_tensor *w = ggml_new_tensor_2d(ctx, (enum ggml_type)type, K, N);
    struct ggml_tensor *x = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, M);
    memcpy(w->data, W, wbytes);
    memcpy(x->data, X, (size_t)M * K * 4);
    struct ggml_tensor *y = ggml_mul_mat(ctx, w, x);
    struct ggml_cgraph gf = ggml_build_forward(y);
    gf.n_threads = n_threads;
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); r++) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        ggml_graph_compute(ctx, &gf);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double us = (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) / 1e3;
        if (us < best) best = us;
    }
    memcpy(Y, y->data, (size_t)M * N * 4);
    ggml_free(ctx);
    return best;
}

/* Timing form for bench.py --impl reference: the tensors and the graph are built once (as vsim.cpp
 * loads weights once), then `warmup` untimed and `steps` timed ggml_graph_compute calls; returns
 * the wall time of the timed calls in microseconds (or -1). */
double ref_mul_mat_steps(int type, const void *W, int N, int K, const float *X, int M, float *Y,
                         int n_threads, int warmup, int steps)
{
    const size_t wbytes = (size_t)N * (K / 32) * (type == 0 ? 20 : 24);
    const size_t mem = wbytes + (size_t)M * K * 4 + (size_t)M * N * 4
                     + (size_t)M * (K / 32) * 24 + 64 * 64 + (1u << 20);
    struct ggml_init_params ip = { mem, NULL };
    struct ggml_context *ctx = ggml_init(ip);
    if (!ctx) return -1.0;
    struct ggml_tensor *w = ggml_new_tensor_2d(ctx, (enum ggml_type)type, K, N);
    struct ggml_tensor *x = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, M);
    memcpy(w->data, W, wbytes);
    memcpy(x->data, X, (size_t)M * K * 4);
    struct ggml_tensor *y = ggml_mul_mat(ctx, w, x);
    struct ggml_cgraph gf = ggml_build_forward(y);
    gf.n_threads = n_threads;
    for (int r = 0; r < warmup; r++) ggml_graph_compute(ctx, &gf);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < steps; r++) ggml_graph_compute(ctx, &gf);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    memcpy(Y, y->data, (size_t)M * N * 4);
    ggml_free(ctx);
    return (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) / 1e3;
}

/* Same, but src0 is the transposed view of a [N][K]-shaped quantized tensor
 * stored as K rows of N values (ggml_transpose -> nb01 < nb00 branch,
 * ggml.c:5118-5164 / 5391-5440). */
double ref_mul_mat_transposed(int type, const void *Wt, int N, int K,
                              const float *X, int M, float *Y, int n_threads)
{
    const size_t wbytes = (size_t)K * (N / 32) * (type == 0 ? 20 : 24);
    const size_t mem = wbytes + (size_t)M * K * 4 + (size_t)M * N * 4 * (n_threads + 2)
                     + 64 * 64 * 4 + (1u << 20);
    struct ggml_init_params ip = { mem, NULL };
    struct ggml_context *ctx = ggml_init(ip);
    if (!ctx) return -1.0;
    struct ggml_tensor *wt = ggml_new_tensor_2d(ctx, (enum ggml_type)type, N, K);
    struct ggml_tensor *x  = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, K, M);
    memcpy(wt->data, Wt, wbytes);
    memcpy(x->data, X, (size_t)M * K * 4);
    struct ggml_tensor *y = ggml_mul_mat(ctx, ggml_transpose(ctx, wt), x);
    struct ggml_cgraph gf = ggml_build_forward(y);
    gf.n_threads = n_threads;
    ggml_graph_compute(ctx, &gf);
    memcpy(Y, y->data, (size_t)M * N * 4);
    ggml_free(ctx);
    return 0.0;
}

/* Layout probe used by tests/test_abi.py to check include/b200q4_ggml_abi.h
 * against the real struct ggml_tensor. */
#include <stddef.h>
void ref_tensor_layout(size_t out[8])
{
    out[0] = sizeof(struct ggml_tensor);
    out[1] = offsetof(struct ggml_tensor, type);
    out[2] = offsetof(struct ggml_tensor, n_dims);
    out[3] = offsetof(struct ggml_tensor, ne);
    out[4] = offsetof(struct ggml_tensor, nb);
    out[5] = offsetof(struct ggml_tensor, src0);
    out[6] = offsetof(struct ggml_tensor, src1);
    out[7] = offsetof(struct ggml_tensor, data);
}

/* ---- one-op graphs of the NON-hot ops of the NeoX graph (row f-1 groundwork: pins oracle/neox_ops_oracle.c) ----
 * op: 0 norm, 1 gelu, 2 soft_max, 3 gptneox_rope(n_past = a, n_dims = b, mode = c) on x[ne2][ne1][ne0],
 *     4 diag_mask_inf(n_past = a).  x is copied into a fresh tensor, y receives the result tensor's data. */
int ref_op(int op, const float *x, int ne0, int ne1, int ne2, int a, int b, int c, float *y, int n_threads)
{
    const size_t n = (size_t)ne0 * ne1 * ne2;
    struct ggml_init_params ip = { n * 4 * 4 + (1u << 20), NULL };
    struct ggml_context *ctx = ggml_init(ip);
    if (!ctx) return -1;
    struct ggml_tensor *t = ggml_new_tensor_3d(ctx, GGML_TYPE_F32, ne0, ne1, ne2);
    memcpy(t->data, x, n * 4);
    struct ggml_tensor *r = NULL;
    switch (op) {
    case 0: r = ggml_norm(ctx, t); break;
    case 1: r = ggml_gelu(ctx, t); break;
    case 2: r = ggml_soft_max(ctx, t); break;
    case 3: r = ggml_gptneox_rope(ctx, t, a, b, c); break;
    case 4: r = ggml_diag_mask_inf(ctx, t, a); break;
    default: ggml_free(ctx); return -2;
    }
    struct ggml_cgraph gf = ggml_build_forward(r);
    gf.n_threads = n_threads;
    ggml_graph_compute(ctx, &gf);
    memcpy(y, r->data, n * 4);
    ggml_free(ctx);
    return 0;
}

int ref_get_rows(int type, const void *W, int N, int K, const int *ids, int n_ids, float *y)
{
    const size_t wbytes = (size_t)N * (K / 32) * (type == 0 ? 20 : 24);
    struct ggml_init_params ip = { wbytes + (size_t)n_ids * K * 4 + (size_t)n_ids * 4 + (1u << 20), NULL };
    struct ggml_context *ctx = ggml_init(ip);
    if (!ctx) return -1;
    struct ggml_tensor *w = ggml_new_tensor_2d(ctx, (enum ggml_type)type, K, N);
    struct ggml_tensor *idt = ggml_new_tensor_1d(ctx, GGML_TYPE_I32, n_ids);
    memcpy(w->data, W, wbytes);
    memcpy(idt->data, ids, (size_t)n_ids * 4);
    struct ggml_tensor *r = ggml_get_rows(ctx, w, idt);
    struct ggml_cgraph gf = ggml_build_forward(r);
    gf.n_threads = 1;
    ggml_graph_compute(ctx, &gf);
    memcpy(y, r->data, (size_t)n_ids * K * 4);
    ggml_free(ctx);
    return 0;
}
#endif /* REF_SHIM_LIB */
