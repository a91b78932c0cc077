/* neox_ops_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the NON-hot ops the reference's GPT-NeoX graph executes between its Q4 mul_mats
 * (SURVEY.md section 8 row f-1, appendix D): the groundwork a device-resident graph will be checked against.
 * Every function cites the reference lines it follows and is pinned bit for bit to the compiled reference
 * (oracle/_ref/libggml_ref.so, one-op graphs through the public ggml.h API) by tests/test_ops_oracle.py.
 * The parity traps live here: gelu and soft_max go through fp16 lookup tables, norm accumulates in double,
 * rope works in double and -- in mode 1 -- only touches the rows past n_past, in place.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

void q4o_dequantize_row_q4_0(const void *x, float *y, int k);
void q4o_dequantize_row_q4_1(const void *x, float *y, int k);

/* ---- fp16 <-> fp32 (ggml.c:81-143: software conversion, IEEE round-to-nearest-even; the build has no F16C) ---- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float nxo_fp16_to_fp32(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
    if (exp == 0x1F) return u2f(sign | 0x7F800000u | (man << 13));            /* inf / nan */
    if (exp != 0) return u2f(sign | ((exp + 112u) << 23) | (man << 13));     /* normal */
    if (man == 0) return u2f(sign);                                          /* +-0 */
    const float v = (float)man * 0x1.0p-24f;                                 /* subnormal: man * 2^-24, exact */
    return u2f(sign | f2u(v));
}

uint16_t nxo_fp32_to_fp16(float f)
{
    const uint32_t w = f2u(f);
    const uint16_t sign = (uint16_t)((w >> 16) & 0x8000u);
    const uint32_t a = w & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);                   /* nan -> the quiet nan ggml emits */
    if (a >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);                  /* |f| >= 65536 (incl. inf) -> inf */
    if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                  /* rounds up to 65536 -> inf */
    if (a < 0x33000000u) return sign;                                         /* |f| < 2^-25 -> 0 (ties at 2^-25 go to even = 0) */
    /* round |f| to a multiple of the fp16 ulp of its binade, to nearest even, in double (exact arithmetic) */
    const double x = (double)u2f(a);
    int e;
    (void)frexp(x, &e);                                                       /* x = m * 2^e, m in [0.5, 1) */
    int ulp_exp = e - 11;                                                     /* fp16 keeps 11 significant bits */
    if (ulp_exp < -24) ulp_exp = -24;                                         /* subnormal range: fixed ulp 2^-24 */
    const double q = nearbyint(ldexp(x, -ulp_exp));                           /* default rounding mode: nearest even */
    const double r = ldexp(q, ulp_exp);
    if (r >= 65536.0) return (uint16_t)(sign | 0x7C00u);
    if (r < 0x1.0p-14) return (uint16_t)(sign | (uint16_t)ldexp(r, 24));      /* subnormal: mantissa = r / 2^-24 */
    int er;
    const double mr = frexp(r, &er);                                          /* r = mr * 2^er */
    const uint32_t he = (uint32_t)(er - 1 + 15);
    const uint32_t hm = (uint32_t)ldexp(mr * 2.0 - 1.0, 10);
    return (uint16_t)(sign | (he << 10) | hm);
}

/* ---- the lookup tables of ggml_init (ggml.c:1240-1255) ---- */
static uint16_t t_gelu[1 << 16], t_exp[1 << 16];
static int tables_ready = 0;

static float gelu_f32(float x)                                               /* ggml.c:795-800: double arithmetic, float result */
{
    const double GELU_COEF_A = 0.044715, SQRT_2_OVER_PI = 0.79788456;
    return (float)(0.5 * x * (1.0 + tanh(SQRT_2_OVER_PI * x * (1.0 + GELU_COEF_A * x * x))));
}

void nxo_init_tables(void)
{
    if (tables_ready) return;
    for (uint32_t i = 0; i < (1u << 16); i++) {
        const float f = nxo_fp16_to_fp32((uint16_t)i);
        t_gelu[i] = nxo_fp32_to_fp16(gelu_f32(f));
        t_exp[i]  = nxo_fp32_to_fp16((float)exp((double)f));
    }
    tables_ready = 1;
}

/* ggml.c:4246-4303 ggml_compute_forward_norm_f32: mean and variance in double, eps = 1e-5f, then a float scale */
void nxo_norm_rows(const float *x, int n_rows, int n, float *y)
{
    const double eps = 1e-5f;
    for (int r = 0; r < n_rows; r++) {
        const float *xr = x + (size_t)r * n;
        float *yr = y + (size_t)r * n;
        double mean = 0.0;
        for (int i = 0; i < n; i++) mean += xr[i];
        mean /= n;
        double sum2 = 0.0;
        for (int i = 0; i < n; i++) {
            const double v = xr[i] - mean;
            yr[i] = (float)v;
            sum2 += v * v;
        }
        const float scale = (float)(1.0 / sqrt(sum2 / n + eps));
        for (int i = 0; i < n; i++) yr[i] *= scale;
    }
}

/* ggml.c:819-827 ggml_vec_gelu_f32 (GGML_GELU_FP16): x -> fp16 -> table -> fp32 */
void nxo_gelu(const float *x, long n, float *y)
{
    nxo_init_tables();
    for (long i = 0; i < n; i++) y[i] = nxo_fp16_to_fp32(t_gelu[nxo_fp32_to_fp16(x[i])]);
}

/* ggml.c:5825-5893 ggml_compute_forward_soft_max_f32 (in place): exp through the fp16 table, sum in double */
void nxo_soft_max_rows(float *p, int n_rows, int n)
{
    nxo_init_tables();
    for (int r = 0; r < n_rows; r++) {
        float *pr = p + (size_t)r * n;
        float mx = -INFINITY;
        for (int i = 0; i < n; i++) mx = pr[i] > mx ? pr[i] : mx;
        double sum = 0.0;
        for (int i = 0; i < n; i++) {
            if (pr[i] == -INFINITY) { pr[i] = 0.0f; continue; }
            const float val = nxo_fp16_to_fp32(t_exp[nxo_fp32_to_fp16(pr[i] - mx)]);
            sum += val;
            pr[i] = val;
        }
        const float inv = (float)(1.0 / sum);
        for (int i = 0; i < n; i++) pr[i] *= inv;
    }
}

/* ggml.c:6086-6153 ggml_compute_forward_gptneox_rope_f32, in place on x[ne2][ne1][ne0] (position, head, dim):
 * mode 0 rotates every row i2 with position n_past + i2; mode 1 rotates only the rows i2 >= n_past, with position i2
 * (vsim.cpp:573-579 applies it to the whole K memory view every step: earlier rows keep what earlier steps wrote). */
void nxo_rope_neox(float *x, int ne0, int ne1, int ne2, int n_past, int n_dims, int mode)
{
    for (int i2 = (mode == 0 ? 0 : n_past); i2 < ne2; i2++) {
        const int p = (mode == 0 ? n_past + i2 : i2);
        for (int i1 = 0; i1 < ne1; i1++) {
            float *d = x + ((size_t)i2 * ne1 + i1) * ne0;
            for (int i0 = 0; i0 < n_dims / 2; i0++) {
                const double theta = pow(10000.0, 2 * ((double)-i0) / n_dims);
                const double c = cos(p * theta), s = sin(p * theta);
                const double x1 = d[i0], x2 = d[i0 + n_dims / 2];
                d[i0] = (float)(c * x1 - s * x2);
                d[i0 + n_dims / 2] = (float)(c * x2 + s * x1);
            }
        }
    }
}

/* ggml.c:5764-5798 ggml_compute_forward_diag_mask_inf_f32, in place on x[nz][nr][nc] */
void nxo_diag_mask_inf(float *x, int nc, int nr, int nz, int n_past)
{
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < nr; j++)
            for (int i = n_past; i < nc; i++)
                if (i > n_past + j) x[((size_t)k * nr + j) * nc + i] = -INFINITY;
}

/* ggml.c get_rows on a Q4 matrix: row ids[i] dequantized into y[i][K] (dequantize_row_q4_0/_q4_1, ggml.c:300-362) */
void nxo_get_rows_q4(int type, const void *W, int K, const int *ids, int n_ids, float *y)
{
    const size_t rb = (size_t)(K / 32) * (type == 0 ? 20 : 24);
    for (int i = 0; i < n_ids; i++) {
        const void *row = (const char *)W + (size_t)ids[i] * rb;
        if (type == 0) q4o_dequantize_row_q4_0(row, y + (size_t)i * K, K);
        else           q4o_dequantize_row_q4_1(row, y + (size_t)i * K, K);
    }
}
