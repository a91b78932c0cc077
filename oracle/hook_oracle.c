/* hook_oracle.c -- TEST INFRASTRUCTURE ONLY.  The reference's accelerator hook (imax.c:1133-1139) served by the CPU
 * oracle (q4_oracle.c), linked with the same patched ggml.c as vsim-b200 into oracle/_ref/vsim-oracle.  It bisects
 * program-level differences: vsim-oracle vs vsim-ubuntu isolates the ggml.c patch + hook contract, vsim-oracle vs
 * vsim-b200 isolates the GPU back-end.  B200Q4_DUMP_DIR dumps every call like imax_b200.c does. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/b200q4_ggml_abi.h"
void q4o_quantize_row_q4_0(const float *x, void *y, int k);
void q4o_quantize_row_q4_1(const float *x, void *y, int k);
float q4o_vec_dot_q4_0(int n, const void *w, const void *a);
float q4o_vec_dot_q4_1(int n, const void *w, const void *a);

void init_xmax(void) {}

static void hook(int type, const struct b200q4_ggml_compute_params *params, const struct b200q4_ggml_tensor *src0,
                 const struct b200q4_ggml_tensor *src1, struct b200q4_ggml_tensor *dst)
{
    if (params->type != 1 || params->ith != 0) return;
    const int K = src0->ne[0], N = src0->ne[1], M = src1->ne[1];
    const size_t rb = (size_t)(K / 32) * (type == 0 ? 20 : 24);
    unsigned char *q = malloc(rb);
    for (int m = 0; m < M; m++) {
        const float *x = (const float *)((const char *)src1->data + (size_t)m * src1->nb[1]);
        float *y = (float *)((char *)dst->data + (size_t)m * dst->nb[1]);
        if (type == 0) q4o_quantize_row_q4_0(x, q, K); else q4o_quantize_row_q4_1(x, q, K);
        for (int n = 0; n < N; n++) {
            const void *w = (const char *)src0->data + (size_t)n * src0->nb[1];
            y[n] = type == 0 ? q4o_vec_dot_q4_0(K, w, q) : q4o_vec_dot_q4_1(K, w, q);
        }
    }
    free(q);
    static int dump_n = 0;
    const char *dir = getenv("B200Q4_DUMP_DIR");
    if (dir && dump_n < 400) {
        char path[512];
        snprintf(path, sizeof path, "%s/call_%04d.bin", dir, dump_n++);
        FILE *f = fopen(path, "wb");
        if (f) {
            const int hdr[4] = {type, N, K, M};
            fwrite(hdr, sizeof hdr, 1, f);
            for (int m = 0; m < M; m++) fwrite((const char *)src1->data + (size_t)m * src1->nb[1], sizeof(float), (size_t)K, f);
            for (int m = 0; m < M; m++) fwrite((const char *)dst->data + (size_t)m * dst->nb[1], sizeof(float), (size_t)N, f);
            fclose(f);
        }
    }
}
void imax_ggml_compute_forward_mul_mat_q4_0_f32(int T, int L, const struct b200q4_ggml_compute_params *p, const struct b200q4_ggml_tensor *a,
                                                const struct b200q4_ggml_tensor *b, struct b200q4_ggml_tensor *d) { (void)T; (void)L; hook(0, p, a, b, d); }
void imax_ggml_compute_forward_mul_mat_q4_1_f32(int T, int L, const struct b200q4_ggml_compute_params *p, const struct b200q4_ggml_tensor *a,
                                                const struct b200q4_ggml_tensor *b, struct b200q4_ggml_tensor *d) { (void)T; (void)L; hook(1, p, a, b, d); }
