/*
 * q4_oracle.c -- CPU restatement of vsim's Q4_0 / Q4_1 mul_mat hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vsim_b200/ (the product) may call,
 * link or import this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs.  The product path is the CUDA library behind include/b200q4.h.
 *
 * Parity status: PINNED.  The reference ships no golden vectors of its own
 * (SURVEY.md section 4), so this restatement is pinned against the reference
 * itself compiled here from /root/reference by oracle/Makefile
 * (oracle/_ref/libggml_ref.so); tests/test_oracle.py compares the two byte
 * for byte whenever oracle/_ref exists, and against tests/golden/*.npz (vectors
 * produced by that compiled reference, generator: tests/golden/make_golden.py)
 * everywhere else.
 *
 * Arithmetic contract (must be built WITHOUT fast-math and WITHOUT FMA
 * contraction: -O2 -msse3 -ffp-contract=off, as Makefile-ubuntu:5 implies):
 *   every float multiply, add, subtract and divide below is one individually
 *   rounded IEEE-754 binary32 operation, evaluated in the order written.
 *
 * Each function names the reference lines it follows.
 */
#include <math.h>
#include <float.h>
#include <limits.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define Q4O_QK 32            /* ggml.c:204  "#define QK 32" */
#define Q4O_BS0 20           /* ggml.c:895-915: sizeof(float) + QK/2     */
#define Q4O_BS1 24           /* ggml.c:895-915: 2*sizeof(float) + QK/2   */

/* Reference: `(int8_t) round(v)` and `(uint8_t) round(v)` (ggml.c:239-240,
 * 287-288).  For finite in-range v this is the obvious conversion.  For
 * v = +-inf / NaN (only reachable when 1/d overflows, i.e. a block whose
 * amax is a denormal) C leaves the conversion undefined; the x86-64 build of
 * the reference goes through cvttsd2si which returns INT_MIN, whose low byte
 * is 0.  We state that behaviour explicitly so that the oracle does not
 * depend on the compiler ("out-of-domain corner"). */
static inline int q4o_round_to_int(float v)
{
    double r = round((double)v);           /* half away from zero */
    if (!(fabs(r) < 2147483648.0))
        return INT_MIN;
    return (int)r;
}

/* ggml.c:209-251  quantize_row_q4_0.
 * Block i of 32 floats -> { float d; uint8_t qs[16] } at y + 20*i. */
void q4o_quantize_row_q4_0(const float *x, void *y, int k)
{
    const int nblk = k / Q4O_QK;
    uint8_t *out = (uint8_t *)y;

    for (int i = 0; i < nblk; i++) {
        const float *xb = x + (size_t)i * Q4O_QK;
        float amax = 0.0f;
        for (int l = 0; l < Q4O_QK; l++) {
            float a = fabsf(xb[l]);
            amax = (amax > a) ? amax : a;               /* MAX macro, ggml.c:62 */
        }
        const float d  = amax / 7.0f;                    /* (1<<3)-1, ggml.c:229 */
        const float id = (d != 0.0f) ? 1.0f / d : 0.0f;  /* ggml.c:230 */

        memcpy(out, &d, sizeof(float));
        uint8_t *qs = out + sizeof(float);
        for (int l = 0; l < Q4O_QK; l += 2) {
            const float v0 = xb[l]     * id;
            const float v1 = xb[l + 1] * id;
            const uint8_t c0 = (uint8_t)((int8_t)q4o_round_to_int(v0) + 8);
            const uint8_t c1 = (uint8_t)((int8_t)q4o_round_to_int(v1) + 8);
            qs[l / 2] = (uint8_t)(c0 | (c1 << 4));       /* ggml.c:245 */
        }
        out += Q4O_BS0;
    }
}

/* ggml.c:256-298  quantize_row_q4_1.
 * Row layout is struct-of-arrays: float min[nblk]; float d[nblk]; u8 qs[nblk][16]. */
void q4o_quantize_row_q4_1(const float *x, void *y, int k)
{
    const int nblk = k / Q4O_QK;
    float   *pm = (float *)y;
    float   *pd = pm + nblk;
    uint8_t *pb = (uint8_t *)(pd + nblk);

    for (int i = 0; i < nblk; i++) {
        const float *xb = x + (size_t)i * Q4O_QK;
        float mn = FLT_MAX, mx = -FLT_MAX;               /* ggml.c:268-269 */
        for (int l = 0; l < Q4O_QK; l++) {
            const float v = xb[l];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        const float d  = (mx - mn) / 15.0f;              /* (1<<4)-1, ggml.c:277 */
        const float id = (d != 0.0f) ? 1.0f / d : 0.0f;
        pm[i] = mn;
        pd[i] = d;
        for (int l = 0; l < Q4O_QK; l += 2) {
            const float v0 = (xb[l]     - mn) * id;      /* subtract, then multiply */
            const float v1 = (xb[l + 1] - mn) * id;
            const uint8_t c0 = (uint8_t)q4o_round_to_int(v0);
            const uint8_t c1 = (uint8_t)q4o_round_to_int(v1);
            pb[i * 16 + l / 2] = (uint8_t)(c0 | (c1 << 4));
        }
    }
}

/* ggml.c:301-334  dequantize_row_q4_0:  y = (code - 8) * d */
void q4o_dequantize_row_q4_0(const void *x, float *y, int k)
{
    const int nblk = k / Q4O_QK;
    const uint8_t *in = (const uint8_t *)x;
    for (int i = 0; i < nblk; i++, in += Q4O_BS0) {
        float d;
        memcpy(&d, in, sizeof(float));
        const uint8_t *qs = in + sizeof(float);
        for (int j = 0; j < 16; j++) {
            const int8_t lo = (int8_t)(qs[j] & 0xf);
            const int8_t hi = (int8_t)(qs[j] >> 4);
            y[i * Q4O_QK + 2 * j]     = (float)(lo - 8) * d;
            y[i * Q4O_QK + 2 * j + 1] = (float)(hi - 8) * d;
        }
    }
}

/* ggml.c:336-367  dequantize_row_q4_1:  y = code * d + min */
void q4o_dequantize_row_q4_1(const void *x, float *y, int k)
{
    const int nblk = k / Q4O_QK;
    const float   *pm = (const float *)x;
    const float   *pd = pm + nblk;
    const uint8_t *pb = (const uint8_t *)(pd + nblk);
    for (int i = 0; i < nblk; i++) {
        const float m = pm[i], d = pd[i];
        for (int j = 0; j < 16; j++) {
            const int8_t lo = (int8_t)(pb[i * 16 + j] & 0xf);
            const int8_t hi = (int8_t)(pb[i * 16 + j] >> 4);
            y[i * Q4O_QK + 2 * j]     = (float)lo * d + m;
            y[i * Q4O_QK + 2 * j + 1] = (float)hi * d + m;
        }
    }
}

/* ggml.c:472-511 ggml_vec_dot_q4_0, restated verbatim by the live x86 loop
 * imax.c:1205-1229: sequential fp32 sum over blocks, then bytes; every
 * product and sum individually rounded: sumf += (f0*f2) + (f1*f3). */
float q4o_vec_dot_q4_0(int n, const void *w, const void *a)
{
    const int nblk = n / Q4O_QK;
    const uint8_t *pw = (const uint8_t *)w;
    const uint8_t *pa = (const uint8_t *)a;
    float sumf = 0.0f;
    for (int i = 0; i < nblk; i++, pw += Q4O_BS0, pa += Q4O_BS0) {
        float d0, d1;
        memcpy(&d0, pw, sizeof(float));
        memcpy(&d1, pa, sizeof(float));
        const uint8_t *q0 = pw + sizeof(float);
        const uint8_t *q1 = pa + sizeof(float);
        for (int j = 0; j < 16; j++) {
            const uint8_t v0 = q0[j], v1 = q1[j];
            const float f0 = d0 * (float)((int8_t)(v0 & 0xf) - 8);
            const float f1 = d0 * (float)((int8_t)(v0 >> 4)  - 8);
            const float f2 = d1 * (float)((int8_t)(v1 & 0xf) - 8);
            const float f3 = d1 * (float)((int8_t)(v1 >> 4)  - 8);
            const float p0 = f0 * f2;
            const float p1 = f1 * f3;
            const float s  = p0 + p1;
            sumf = sumf + s;
        }
    }
    return sumf;
}

/* ggml.c:513-553 ggml_vec_dot_q4_1:  f = d*code + m on both sides. */
float q4o_vec_dot_q4_1(int n, const void *w, const void *a)
{
    const int nblk = n / Q4O_QK;
    const float *m0p = (const float *)w, *m1p = (const float *)a;
    const float *d0p = m0p + nblk,      *d1p = m1p + nblk;
    const uint8_t *q0 = (const uint8_t *)(d0p + nblk);
    const uint8_t *q1 = (const uint8_t *)(d1p + nblk);
    float sumf = 0.0f;
    for (int i = 0; i < nblk; i++) {
        const float m0 = m0p[i], m1 = m1p[i], d0 = d0p[i], d1 = d1p[i];
        for (int j = 0; j < 16; j++) {
            const uint8_t v0 = q0[i * 16 + j], v1 = q1[i * 16 + j];
            const float t0 = d0 * (float)(v0 & 0xf);
            const float t1 = d0 * (float)(v0 >> 4);
            const float t2 = d1 * (float)(v1 & 0xf);
            const float t3 = d1 * (float)(v1 >> 4);
            const float f0 = t0 + m0, f1 = t1 + m0, f2 = t2 + m1, f3 = t3 + m1;
            const float p0 = f0 * f2;
            const float p1 = f1 * f3;
            const float s  = p0 + p1;
            sumf = sumf + s;
        }
    }
    return sumf;
}

size_t q4o_row_bytes(int type, int k)
{
    return (size_t)(k / Q4O_QK) * (type == 0 ? Q4O_BS0 : Q4O_BS1);
}

/* ------------------------------------------------------------------------
 * The operator: ggml_compute_forward_mul_mat_q4_0_f32 (ggml.c:4891-5165) and
 * _q4_1_f32 (ggml.c:5167-5453), non-transposed src0 (nb01 >= nb00) case.
 *   INIT    (ggml.c:5024-5041 / 5302-5320): quantize every src1 row into wdata
 *   COMPUTE (imax.c:1182-1230 / ggml.c:5350-5390): thread ith owns weight rows
 *           [ith*ceil(N/nth), ...), Y[m][n] = dot(W[n], wdata[m]).
 * The result does not depend on nth.
 * ------------------------------------------------------------------------ */
struct q4o_job {
    int type, ith, nth, N, K, M;
    const uint8_t *W;
    const uint8_t *wdata;
    float *Y;
    size_t y_stride;            /* in floats */
};

static void *q4o_worker(void *arg)
{
    const struct q4o_job *j = (const struct q4o_job *)arg;
    const size_t rb = q4o_row_bytes(j->type, j->K);
    const int dr  = (j->N + j->nth - 1) / j->nth;       /* imax.c:1185 */
    const int ir0 = dr * j->ith;
    const int ir1 = (ir0 + dr < j->N) ? ir0 + dr : j->N;
    for (int ir = ir0; ir < ir1; ir++) {
        const uint8_t *wrow = j->W + (size_t)ir * rb;
        for (int ic = 0; ic < j->M; ic++) {
            const uint8_t *arow = j->wdata + (size_t)ic * rb;
            j->Y[(size_t)ic * j->y_stride + ir] = (j->type == 0)
                ? q4o_vec_dot_q4_0(j->K, wrow, arow)
                : q4o_vec_dot_q4_1(j->K, wrow, arow);
        }
    }
    return NULL;
}

/* type: 0 = Q4_0, 1 = Q4_1 (enum ggml_type, ggml.h:217-226).
 * W: N rows of reference-format bytes; X: M rows of K floats, row stride
 * x_stride floats; Y: M rows of N floats, row stride y_stride floats.
 * wdata_out (optional, may be NULL): receives the M quantized activation rows
 * exactly as the reference leaves them in params->wdata.
 * Returns 0, or -1 on bad arguments. */
int q4o_mul_mat(int type, const void *W, int N, int K,
                const float *X, int M, size_t x_stride,
                float *Y, size_t y_stride, int nth, void *wdata_out)
{
    if ((type != 0 && type != 1) || K <= 0 || (K % Q4O_QK) != 0 || N < 0 || M < 0)
        return -1;
    if (nth < 1) nth = 1;
    if (nth > 64) nth = 64;
    const size_t rb = q4o_row_bytes(type, K);
    uint8_t *wdata = (uint8_t *)malloc(rb * (size_t)(M > 0 ? M : 1));
    if (!wdata) return -1;
    for (int m = 0; m < M; m++) {                        /* INIT phase */
        if (type == 0) q4o_quantize_row_q4_0(X + (size_t)m * x_stride, wdata + (size_t)m * rb, K);
        else           q4o_quantize_row_q4_1(X + (size_t)m * x_stride, wdata + (size_t)m * rb, K);
    }
    if (wdata_out) memcpy(wdata_out, wdata, rb * (size_t)M);

    struct q4o_job jobs[64];
    pthread_t th[64];
    for (int t = 0; t < nth; t++) {
        jobs[t] = (struct q4o_job){ type, t, nth, N, K, M, (const uint8_t *)W, wdata, Y, y_stride };
        if (t > 0) pthread_create(&th[t], NULL, q4o_worker, &jobs[t]);
    }
    q4o_worker(&jobs[0]);
    for (int t = 1; t < nth; t++) pthread_join(th[t], NULL);
    free(wdata);
    return 0;
}

/* ------------------------------------------------------------------------
 * Transposed-src0 branch (nb01 < nb00): ggml.c:5118-5164 / 5391-5440 with
 * ggml_vec_mad_q4_0 / _q4_1 (ggml.c:672-741).  src0 is stored as K rows of N
 * quantized values (quant blocks run along N); activations stay fp32.
 *   Y[m][:] = sum over k (in order) of dequant(Wt[k][:]) * X[m][k]
 * The reference gives every thread a private partial Y (columns of src1 split
 * contiguously) and adds the partials in thread order in FINALIZE
 * (ggml.c:5048-5071); nth therefore changes the rounding, so it is a
 * parameter here too.
 * ------------------------------------------------------------------------ */
static void q4o_vec_mad(int type, int n, float *y, const void *x, float v)
{
    const int nblk = n / Q4O_QK;
    if (type == 0) {
        const uint8_t *in = (const uint8_t *)x;
        for (int i = 0; i < nblk; i++, in += Q4O_BS0) {
            float d; memcpy(&d, in, sizeof(float));
            const uint8_t *qs = in + sizeof(float);
            for (int j = 0; j < 16; j++) {
                const float v0 = (float)((int8_t)(qs[j] & 0xf) - 8) * d;
                const float v1 = (float)((int8_t)(qs[j] >> 4)  - 8) * d;
                const float p0 = v0 * v, p1 = v1 * v;
                y[i * Q4O_QK + 2 * j]     = y[i * Q4O_QK + 2 * j]     + p0;
                y[i * Q4O_QK + 2 * j + 1] = y[i * Q4O_QK + 2 * j + 1] + p1;
            }
        }
    } else {
        const float *pm = (const float *)x, *pd = pm + nblk;
        const uint8_t *pb = (const uint8_t *)(pd + nblk);
        for (int i = 0; i < nblk; i++) {
            const float m = pm[i], d = pd[i];
            for (int j = 0; j < 16; j++) {
                const float t0 = d * (float)(pb[i * 16 + j] & 0xf);
                const float t1 = d * (float)(pb[i * 16 + j] >> 4);
                const float v0 = t0 + m, v1 = t1 + m;
                const float p0 = v0 * v, p1 = v1 * v;
                y[i * Q4O_QK + 2 * j]     = y[i * Q4O_QK + 2 * j]     + p0;
                y[i * Q4O_QK + 2 * j + 1] = y[i * Q4O_QK + 2 * j + 1] + p1;
            }
        }
    }
}

int q4o_mul_mat_transposed(int type, const void *Wt, int N, int K,
                           const float *X, int M, size_t x_stride,
                           float *Y, size_t y_stride, int nth)
{
    if ((type != 0 && type != 1) || N <= 0 || (N % Q4O_QK) != 0 || K < 0 || M < 0)
        return -1;
    if (nth < 1) nth = 1;
    const size_t rb = q4o_row_bytes(type, N);
    float *part = (float *)calloc((size_t)nth * (size_t)M * (size_t)N + 1, sizeof(float));
    if (!part) return -1;
    const int dc = (K + nth - 1) / nth;                   /* ggml.c:5123 */
    for (int t = 0; t < nth; t++) {
        const int ic0 = dc * t;
        const int ic1 = (ic0 + dc < K) ? ic0 + dc : K;
        float *py = part + (size_t)t * M * N;
        for (int m = 0; m < M; m++)
            for (int ic = ic0; ic < ic1; ic++)
                q4o_vec_mad(type, N, py + (size_t)m * N,
                            (const uint8_t *)Wt + (size_t)ic * rb, X[(size_t)m * x_stride + ic]);
    }
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            float acc = part[(size_t)m * N + n];          /* cpy thread 0, then acc 1.. */
            for (int t = 1; t < nth; t++)
                acc = acc + part[((size_t)t * M + m) * N + n];
            Y[(size_t)m * y_stride + n] = acc;
        }
    free(part);
    return 0;
}

/* ---- offline WEIGHT quantizers ------------------------------------------
 * utils.cpp:424-481 ggml_quantize_q4_0 and utils.cpp:483-541 ggml_quantize_q4_1, the functions quantize_*.cpp
 * run over every 2-D tensor of a model file.  Same block formats and rounding as the activation quantizers
 * above, with two differences: (1) a 16-bin histogram of the emitted codes, (2) q4_1 starts its running
 * maximum at std::numeric_limits<float>::min() -- the smallest POSITIVE normal float, not the most negative
 * one (utils.cpp:503) -- so a block whose values are all <= 1.17549435e-38 gets max = 1.17549435e-38.
 * n_rows rows of k floats, tightly packed; returns the bytes written. */
size_t q4o_quantize_weights(int type, const float *src, long n_rows, int k, void *dst, int64_t *hist)
{
    const int nblk = k / Q4O_QK;
    const size_t row_size = (size_t)nblk * (type == 0 ? Q4O_BS0 : 24);
    for (long row = 0; row < n_rows; row++) {
        const float *x = src + (size_t)row * k;
        uint8_t *out = (uint8_t *)dst + (size_t)row * row_size;
        if (type == 0) {
            for (int i = 0; i < nblk; i++) {
                const float *xb = x + (size_t)i * Q4O_QK;
                float amax = 0.0f;
                for (int l = 0; l < Q4O_QK; l++) {
                    const float a = fabsf(xb[l]);
                    amax = (amax < a) ? a : amax;        /* std::max(amax, fabsf(v)), utils.cpp:447 */
                }
                const float d  = amax / 7.0f;
                const float id = (d != 0.0f) ? 1.0f / d : 0.0f;
                uint8_t *blk = out + (size_t)i * Q4O_BS0;
                memcpy(blk, &d, sizeof(float));
                for (int l = 0; l < Q4O_QK; l += 2) {
                    const float v0 = xb[l] * id, v1 = xb[l + 1] * id;
                    const uint8_t c0 = (uint8_t)((int8_t)q4o_round_to_int(v0) + 8);
                    const uint8_t c1 = (uint8_t)((int8_t)q4o_round_to_int(v1) + 8);
                    hist[c0 & 15]++;
                    hist[c1 & 15]++;
                    blk[4 + l / 2] = (uint8_t)(c0 | (c1 << 4));
                }
            }
        } else {
            float   *pm = (float *)out;
            float   *pd = pm + nblk;
            uint8_t *pb = (uint8_t *)(pd + nblk);
            for (int i = 0; i < nblk; i++) {
                const float *xb = x + (size_t)i * Q4O_QK;
                float mn = FLT_MAX, mx = FLT_MIN;        /* utils.cpp:502-503: numeric_limits<float>::max() / ::min() */
                for (int l = 0; l < Q4O_QK; l++) {
                    const float v = xb[l];
                    if (v < mn) mn = v;
                    if (v > mx) mx = v;
                }
                const float d  = (mx - mn) / 15.0f;
                const float id = (d != 0.0f) ? 1.0f / d : 0.0f;
                pm[i] = mn;
                pd[i] = d;
                for (int l = 0; l < Q4O_QK; l += 2) {
                    const float v0 = (xb[l] - mn) * id, v1 = (xb[l + 1] - mn) * id;
                    const uint8_t c0 = (uint8_t)q4o_round_to_int(v0);
                    const uint8_t c1 = (uint8_t)q4o_round_to_int(v1);
                    hist[c0 & 15]++;
                    hist[c1 & 15]++;
                    pb[i * 16 + l / 2] = (uint8_t)(c0 | (c1 << 4));
                }
            }
        }
    }
    return (size_t)n_rows * row_size;
}
