// neox_graph.cu -- device-resident GPT-NeoX forward pass (first slice).
//
// Mirrors gptneox_eval (vsim.cpp:470-747) for the graph the reference actually runs (parallel residual, F32 KV
// memory): token ids in, logits of the last token out, everything in between resident in HBM -- one H2D of N ints and
// one D2H of n_vocab floats per call instead of a PCIe round trip per mul_mat.  The Q4 mul_mats go through the public
// C ABI of this library (b200q4_mul_mat_dev: fast kernels, or the bit-identical ones in exact mode); the other ops are
// the kernels below.  They are CORRECTNESS-FIRST: each follows the arithmetic of the reference op it replaces exactly
// -- the order of every accumulation, double where ggml.c uses ggml_float, no FMA contraction, gelu and exp through
// fp16 tables built on the host with the same libm calls as ggml_init (ggml.c:1240-1255), rope angles from the host's
// pow/cos/sin -- so that with exact-mode mul_mats the logits equal the reference program's bit for bit
// (tests/test_gpu_neox_graph.py checks exactly that).  Making them fast (fused, parallel reductions where the order allows) is the
// next step; at batch 1 they are microseconds next to the weight stream.
#include "../../include/b200neox.h"
#include "../../include/b200q4.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

thread_local std::string g_nerr;
int nfail(int code, const std::string &msg) { g_nerr = msg; return code; }

#define NCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return nfail(B200Q4_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ---- kernels: one thread per output element / row, the reference's own operation order ----------------------------

// ggml_get_rows on a Q4 matrix (dequantize_row_q4_0 ggml.c:300-330: v = (q - 8) * d; _q4_1 :332-362: v = q * d + m)
__global__ void k_get_rows_q4(int type, const unsigned char *__restrict__ w, int K, const int *__restrict__ ids, int n_ids, float *__restrict__ y)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n_ids * K) return;
    const int r = (int)(i / K), e = (int)(i % K), blk = e / 32, l = e % 32, nb = K / 32;
    const unsigned char *row = w + (size_t)ids[r] * nb * (type == 0 ? 20 : 24);
    if (type == 0) {
        const unsigned char *b = row + (size_t)blk * 20;
        const float d = *reinterpret_cast<const float *>(b);
        const unsigned char v = b[4 + l / 2];
        const int q = (l & 1) ? (v >> 4) : (v & 15);
        y[i] = __fmul_rn((float)(q - 8), d);
    } else {
        const float m = reinterpret_cast<const float *>(row)[blk], d = reinterpret_cast<const float *>(row)[nb + blk];
        const unsigned char v = row[(size_t)nb * 8 + (size_t)blk * 16 + l / 2];
        const int q = (l & 1) ? (v >> 4) : (v & 15);
        y[i] = __fadd_rn(__fmul_rn((float)q, d), m);
    }
}

// ggml_norm (ggml.c:4246-4303) followed by mul(repeat(w), .) and add(., repeat(b)) (vsim.cpp:523-533): one thread per row
__global__ void k_layernorm(const float *__restrict__ x, int n_rows, int n, const float *__restrict__ w, const float *__restrict__ b,
                            float *__restrict__ y)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const float *xr = x + (size_t)r * n;
    float *yr = y + (size_t)r * n;
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean = __dadd_rn(mean, (double)xr[i]);
    mean = mean / (double)n;
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) {
        const double v = __dsub_rn((double)xr[i], mean);
        yr[i] = (float)v;
        sum2 = __dadd_rn(sum2, __dmul_rn(v, v));
    }
    const double eps = (double)1e-5f;
    const float scale = (float)(1.0 / sqrt(__dadd_rn(sum2 / (double)n, eps)));
    for (int i = 0; i < n; i++) yr[i] = __fadd_rn(__fmul_rn(w[i], __fmul_rn(yr[i], scale)), b[i]);
}

// The same operator for the fast mode (exact mode off): one 256-thread block per row, the two double accumulations as
// parallel tree reductions.  Every float operation is the one above; only the ORDER of the double additions differs,
// i.e. mean and variance may move by an ulp of a double (1e-16) before they are rounded to the float scale.
__global__ void __launch_bounds__(256) k_layernorm_fast(const float *__restrict__ x, int n, const float *__restrict__ w,
                                                         const float *__restrict__ b, float *__restrict__ y)
{
    __shared__ double sh[256];
    const float *xr = x + (size_t)blockIdx.x * n;
    float *yr = y + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x;
    auto block_sum = [&](double v) {
        sh[tid] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) sh[tid] = __dadd_rn(sh[tid], sh[tid + o]);
            __syncthreads();
        }
        const double t = sh[0];
        __syncthreads();
        return t;
    };
    double s = 0.0;
    for (int i = tid; i < n; i += 256) s = __dadd_rn(s, (double)xr[i]);
    const double mean = block_sum(s) / (double)n;
    double s2 = 0.0;
    for (int i = tid; i < n; i += 256) {
        const double v = __dsub_rn((double)xr[i], mean);
        s2 = __dadd_rn(s2, __dmul_rn(v, v));
    }
    const double eps = (double)1e-5f;
    const float scale = (float)(1.0 / sqrt(__dadd_rn(block_sum(s2) / (double)n, eps)));
    for (int i = tid; i < n; i += 256) {
        const float v = (float)__dsub_rn((double)xr[i], mean);
        yr[i] = __fadd_rn(__fmul_rn(w[i], __fmul_rn(v, scale)), b[i]);
    }
}

// y[r][i] = a[r][i] + bias[i]   (ggml_add with a repeated bias; float addition commutes)
__global__ void k_add_bias(float *__restrict__ a, const float *__restrict__ bias, long total, int n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) a[i] = __fadd_rn(a[i], bias[i % n]);
}
__global__ void k_add(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ y, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) y[i] = __fadd_rn(a[i], b[i]);
}
__global__ void k_copy(const float *__restrict__ a, float *__restrict__ y, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) y[i] = a[i];
}

// ggml_gptneox_rope (ggml.c:6086-6153), in place on x[ne2][ne1][ne0]; cs[p][i0] = {cos, sin}(p * theta_i0) from the host
__global__ void k_rope(float *__restrict__ x, int ne0, int ne1, int ne2, int n_past, int n_dims, int mode, const double2 *__restrict__ cs)
{
    const int half = n_dims / 2;
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int first = mode == 0 ? 0 : n_past;
    if (id >= (long)(ne2 - first) * ne1 * half) return;
    const int i0 = (int)(id % half);
    const int i1 = (int)((id / half) % ne1);
    const int i2 = first + (int)(id / ((long)half * ne1));
    const int p = mode == 0 ? n_past + i2 : i2;
    const double2 c = cs[(size_t)p * half + i0];
    float *d = x + ((size_t)i2 * ne1 + i1) * ne0;
    const double x1 = d[i0], x2 = d[i0 + half];
    d[i0] = (float)__dsub_rn(__dmul_rn(c.x, x1), __dmul_rn(c.y, x2));
    d[i0 + half] = (float)__dadd_rn(__dmul_rn(c.x, x2), __dmul_rn(c.y, x1));
}

// K * Q (ggml.c:4484-4516 + vec_dot_f32 :399-434): float products accumulated in double, in order; then ggml_scale and
// ggml_diag_mask_inf.  kq[h][n][p], K = the memory view [P][H][d], Q [N][H][d]
__global__ void k_kq(const float *__restrict__ km, const float *__restrict__ q, int H, int N, int P, int d, int n_past, float scale, float *__restrict__ kq)
{
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long)H * N * P) return;
    const int p = (int)(id % P), n = (int)((id / P) % N), h = (int)(id / ((long)P * N));
    const float *kr = km + ((size_t)p * H + h) * d, *qr = q + ((size_t)n * H + h) * d;
    double s = 0.0;
    for (int i = 0; i < d; i++) s = __dadd_rn(s, (double)__fmul_rn(kr[i], qr[i]));
    const float v = __fmul_rn((float)s, scale);
    kq[id] = p > n_past + n ? -INFINITY : v;
}

// ggml_soft_max (ggml.c:5825-5893), in place: exp through the fp16 table, sum in double; one thread per row
__global__ void k_soft_max(float *__restrict__ p, int n_rows, int n, const unsigned short *__restrict__ t_exp)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    float *pr = p + (size_t)r * n;
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) mx = pr[i] > mx ? pr[i] : mx;
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        if (pr[i] == -INFINITY) { pr[i] = 0.0f; continue; }
        const __half hx = __float2half_rn(__fsub_rn(pr[i], mx));
        const float val = __half2float(__ushort_as_half(t_exp[__half_as_ushort(hx)]));
        sum = __dadd_rn(sum, (double)val);
        pr[i] = val;
    }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; i++) pr[i] = __fmul_rn(pr[i], inv);
}

// V^T * softmax, ggml's transposed f32 mul_mat (ggml.c:4517-4553 + FINALIZE :4463-4481): thread t of nth owns positions
// [t*dc, (t+1)*dc), float accumulation in position order from +0, then dst = part[0]; dst += part[t].  Output written
// already merged: out[n][h*d + j]  (permute(0,2,1,3) + cpy, vsim.cpp:621-627)
__global__ void k_kqv(const float *__restrict__ vm, const float *__restrict__ kq, int H, int N, int P, int d, int nth, float *__restrict__ out)
{
    const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long)H * N * d) return;
    const int j = (int)(id % d), n = (int)((id / d) % N), h = (int)(id / ((long)d * N));
    const float *pr = kq + ((size_t)h * N + n) * P;
    const int dc = (P + nth - 1) / nth;
    float total = 0.0f;
    for (int t = 0; t < nth; t++) {
        float part = 0.0f;
        const int hi = min((t + 1) * dc, P);
        for (int ic = t * dc; ic < hi; ic++) part = __fadd_rn(part, __fmul_rn(vm[((size_t)ic * H + h) * d + j], pr[ic]));
        total = t == 0 ? part : __fadd_rn(total, part);
    }
    out[(size_t)n * H * d + (size_t)h * d + j] = total;
}

// ggml_gelu (ggml.c:819-827, GGML_GELU_FP16): x -> fp16 -> table -> fp32
__global__ void k_gelu(float *__restrict__ x, long total, const unsigned short *__restrict__ t_gelu)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) x[i] = __half2float(__ushort_as_half(t_gelu[__half_as_ushort(__float2half_rn(x[i]))]));
}

inline unsigned nblk(long total, int threads = 256) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace

struct b200neox {
    b200q4_model *file = nullptr;
    int n_vocab = 0, n_embd = 0, n_head = 0, n_layer = 0, n_rot = 0, par_res = 0, ftype = 0, n_ctx = 0;
    std::map<std::string, int> index;                    // tensor name -> index in the file
    std::map<std::string, float *> f32;                  // 1-D tensors, device copies
    unsigned char *d_wte = nullptr;                      // embedding matrix in the reference's row format (get_rows dequantizes it)
    float *mem_k = nullptr, *mem_v = nullptr;            // [n_layer][n_ctx][n_embd]   (vsim.cpp:318-330)
    unsigned short *t_gelu = nullptr, *t_exp = nullptr;  // ggml_init's fp16 tables
    double2 *rope_cs = nullptr;                          // [n_ctx][n_rot/2] {cos, sin}
    float *x = nullptr, *cur = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *ff = nullptr, *ff2 = nullptr,
          *kq = nullptr, *merged = nullptr, *logits = nullptr;
    int *ids = nullptr;
    int max_n = 0;
    cudaStream_t st = nullptr;
};

extern "C" {

const char *b200neox_last_error(void) { return g_nerr.c_str(); }

void b200neox_free(b200neox *m)
{
    if (!m) return;
    if (m->st) cudaStreamSynchronize(m->st);
    for (auto &kv : m->f32) cudaFree(kv.second);
    float *bufs[] = {m->mem_k, m->mem_v, m->x, m->cur, m->q, m->k, m->v, m->att, m->ff, m->ff2, m->kq, m->merged, m->logits};
    for (float *b : bufs) cudaFree(b);
    cudaFree(m->d_wte); cudaFree(m->t_gelu); cudaFree(m->t_exp); cudaFree(m->rope_cs); cudaFree(m->ids);
    if (m->st) cudaStreamDestroy(m->st);
    if (m->file) b200q4_model_free(m->file);
    delete m;
}

b200neox *b200neox_load(const char *path, int n_ctx, int max_batch)
{
    if (!path || n_ctx < 1 || max_batch < 1) { nfail(B200Q4_EINVAL, "neox_load: bad argument"); return nullptr; }
    b200q4_model *f = b200q4_model_load(path, 7, 0, 1);           // vsim.cpp:119-138: 7 header ints, no vocabulary count
    if (!f) { nfail(B200Q4_EINVAL, std::string("neox_load: ") + b200q4_last_error()); return nullptr; }
    b200neox *m = new b200neox();
    m->file = f;
    int *hp[] = {&m->n_vocab, &m->n_embd, &m->n_head, &m->n_layer, &m->n_rot, &m->par_res, &m->ftype};
    for (int i = 0; i < 7; i++) *hp[i] = b200q4_model_header(f, i);
    m->n_ctx = n_ctx; m->max_n = max_batch;
    auto bail = [&](const std::string &why) { nfail(B200Q4_EINVAL, "neox_load: " + why); b200neox_free(m); return (b200neox *)nullptr; };
    if (m->par_res != 1) return bail("only use_parallel_residual == 1 graphs are built (what the reference's models use)");
    if (m->n_embd % m->n_head || m->n_rot % 2 || m->n_rot > m->n_embd / m->n_head) return bail("bad hyper-parameters");
    if (cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed");
    const int nt = b200q4_model_tensor_count(f);
    for (int i = 0; i < nt; i++) {
        const char *name; int nd, ft, ne0, ne1; size_t off, nb;
        b200q4_model_tensor_info(f, i, &name, &nd, &ft, &ne0, &ne1, &off, &nb);
        m->index[name] = i;
        if (nd == 1) {
            if (ft != 0) return bail(std::string("1-D tensor '") + name + "' is not f32");
            float *d = nullptr;
            if (cudaMalloc(&d, nb) != cudaSuccess || cudaMemcpy(d, b200q4_model_tensor_data(f, i), nb, cudaMemcpyHostToDevice) != cudaSuccess)
                return bail("device copy of a 1-D tensor failed");
            m->f32[name] = d;
        } else if (ft < 2) return bail(std::string("2-D tensor '") + name + "' is not Q4 (f16/f32 model files are not built)");
    }
    {   // the embedding matrix is read by get_rows in the reference's row format
        auto it = m->index.find("gpt_neox.embed_in.weight");
        if (it == m->index.end()) return bail("no gpt_neox.embed_in.weight");
        const char *name; int nd, ft, ne0, ne1; size_t off, nb;
        b200q4_model_tensor_info(f, it->second, &name, &nd, &ft, &ne0, &ne1, &off, &nb);
        if (cudaMalloc(&m->d_wte, nb) != cudaSuccess ||
            cudaMemcpy(m->d_wte, b200q4_model_tensor_data(f, it->second), nb, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail("device copy of the embedding matrix failed");
    }
    // ggml_init's tables (ggml.c:1240-1255), with this host's libm like the reference build
    std::vector<unsigned short> tg(1 << 16), te(1 << 16);
    for (unsigned i = 0; i < (1u << 16); i++) {
        __half_raw hr; hr.x = (unsigned short)i;
        const float fv = __half2float(__half(hr));
        const double GELU_COEF_A = 0.044715, SQRT_2_OVER_PI = 0.79788456;
        const float g = (float)(0.5 * fv * (1.0 + tanh(SQRT_2_OVER_PI * fv * (1.0 + GELU_COEF_A * fv * fv))));
        tg[i] = __half_raw(__float2half_rn(g)).x;
        te[i] = __half_raw(__float2half_rn((float)exp((double)fv))).x;
    }
    const int half = m->n_rot / 2;
    std::vector<double2> cs((size_t)n_ctx * half);
    for (int p = 0; p < n_ctx; p++)
        for (int i0 = 0; i0 < half; i0++) {
            const double theta = pow(10000.0, 2 * ((double)-i0) / m->n_rot);             // ggml.c:6126
            cs[(size_t)p * half + i0] = make_double2(cos(p * theta), sin(p * theta));
        }
    const size_t E = (size_t)m->n_embd, Nmax = (size_t)max_batch, kv = (size_t)m->n_layer * n_ctx * E * 4;
    bool ok = cudaMalloc(&m->t_gelu, 1 << 17) == cudaSuccess && cudaMalloc(&m->t_exp, 1 << 17) == cudaSuccess &&
              cudaMalloc(&m->rope_cs, cs.size() * sizeof(double2) + 16) == cudaSuccess &&
              cudaMalloc(&m->mem_k, kv) == cudaSuccess && cudaMalloc(&m->mem_v, kv) == cudaSuccess &&
              cudaMalloc(&m->x, Nmax * E * 4) == cudaSuccess && cudaMalloc(&m->cur, Nmax * E * 4) == cudaSuccess &&
              cudaMalloc(&m->q, Nmax * E * 4) == cudaSuccess && cudaMalloc(&m->k, Nmax * E * 4) == cudaSuccess &&
              cudaMalloc(&m->v, Nmax * E * 4) == cudaSuccess && cudaMalloc(&m->att, Nmax * E * 4) == cudaSuccess &&
              cudaMalloc(&m->merged, Nmax * E * 4) == cudaSuccess && cudaMalloc(&m->ff, Nmax * 4 * E * 4) == cudaSuccess &&
              cudaMalloc(&m->ff2, Nmax * E * 4) == cudaSuccess &&
              cudaMalloc(&m->kq, (size_t)m->n_head * Nmax * n_ctx * 4) == cudaSuccess &&
              cudaMalloc(&m->logits, Nmax * (size_t)m->n_vocab * 4) == cudaSuccess && cudaMalloc(&m->ids, Nmax * 4) == cudaSuccess;
    ok = ok && cudaMemset(m->mem_k, 0, kv) == cudaSuccess && cudaMemset(m->mem_v, 0, kv) == cudaSuccess &&
         cudaMemcpy(m->t_gelu, tg.data(), 1 << 17, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(m->t_exp, te.data(), 1 << 17, cudaMemcpyHostToDevice) == cudaSuccess &&
         (cs.empty() || cudaMemcpy(m->rope_cs, cs.data(), cs.size() * sizeof(double2), cudaMemcpyHostToDevice) == cudaSuccess);
    if (!ok) return bail(std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    return m;
}

int b200neox_n_vocab(const b200neox *m) { return m ? m->n_vocab : B200Q4_EINVAL; }

int b200neox_eval(b200neox *m, const int *tokens, int N, int n_past, int n_threads_emulated, float *logits_out)
{
    if (!m || !tokens || !logits_out || N < 1 || N > m->max_n || n_past < 0 || n_past + N > m->n_ctx || n_threads_emulated < 1)
        return nfail(B200Q4_EINVAL, "neox_eval: bad argument");
    const int E = m->n_embd, H = m->n_head, d = E / H, P = n_past + N, nth = n_threads_emulated;
    cudaStream_t st = m->st;
    const bool exact = b200q4_exact_enabled() != 0;                                          // exact mode covers the graph ops too
    auto W = [&](const std::string &name) { auto it = m->index.find(name); return it == m->index.end() ? nullptr : b200q4_model_weight(m->file, it->second); };
    auto F = [&](const std::string &name) { auto it = m->f32.find(name); return it == m->f32.end() ? nullptr : it->second; };
    auto mm = [&](const std::string &name, const float *x, int rows_in, float *y) -> int {   // ggml_mul_mat on a Q4 weight
        b200q4_weight *w = W(name);
        if (!w) return nfail(B200Q4_EINVAL, "neox_eval: missing tensor " + name);
        if (int rc = b200q4_mul_mat_dev(w, x, rows_in, 0, y, 0, st)) return nfail(rc, std::string("neox_eval: ") + b200q4_last_error());
        return 0;
    };
    auto bias = [&](float *a, const std::string &name, int n) -> int {
        const float *b = F(name);
        if (!b) return nfail(B200Q4_EINVAL, "neox_eval: missing tensor " + name);
        k_add_bias<<<nblk((long)N * n), 256, 0, st>>>(a, b, (long)N * n, n);
        return 0;
    };
    auto ln = [&](const float *x, const std::string &name, float *y) -> int {
        const float *w = F(name + ".weight"), *b = F(name + ".bias");
        if (!w || !b) return nfail(B200Q4_EINVAL, "neox_eval: missing tensor " + name);
        if (exact) k_layernorm<<<nblk(N, 32), 32, 0, st>>>(x, N, E, w, b, y);       // the reference's own summation order
        else       k_layernorm_fast<<<N, 256, 0, st>>>(x, E, w, b, y);
        return 0;
    };
    NCK(cudaMemcpyAsync(m->ids, tokens, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    k_get_rows_q4<<<nblk((long)N * E), 256, 0, st>>>(m->ftype - 2, m->d_wte, E, m->ids, N, m->x);
    const float scale = (float)(1.0 / sqrt((double)((float)E / (float)H)));                 // vsim.cpp:590
    int rc;
    for (int l = 0; l < m->n_layer; l++) {
        const std::string p = "gpt_neox.layers." + std::to_string(l) + ".";
        float *mk = m->mem_k + (size_t)l * m->n_ctx * E, *mv = m->mem_v + (size_t)l * m->n_ctx * E;
        if ((rc = ln(m->x, p + "input_layernorm", m->cur))) return rc;
        if ((rc = mm(p + "attention.query.weight", m->cur, N, m->q)) || (rc = bias(m->q, p + "attention.query.bias", E))) return rc;
        if ((rc = mm(p + "attention.key.weight", m->cur, N, m->k)) || (rc = bias(m->k, p + "attention.key.bias", E))) return rc;
        if ((rc = mm(p + "attention.value.weight", m->cur, N, m->v)) || (rc = bias(m->v, p + "attention.value.bias", E))) return rc;
        k_copy<<<nblk((long)N * E), 256, 0, st>>>(m->k, mk + (size_t)n_past * E, (long)N * E);          // cpy into the KV memory (:557-558)
        k_copy<<<nblk((long)N * E), 256, 0, st>>>(m->v, mv + (size_t)n_past * E, (long)N * E);
        k_rope<<<nblk((long)N * H * (m->n_rot / 2)), 256, 0, st>>>(m->q, d, H, N, n_past, m->n_rot, 0, m->rope_cs);
        k_rope<<<nblk((long)N * H * (m->n_rot / 2)), 256, 0, st>>>(mk, d, H, P, n_past, m->n_rot, 1, m->rope_cs);   // in place on the memory view
        k_kq<<<nblk((long)H * N * P), 256, 0, st>>>(mk, m->q, H, N, P, d, n_past, scale, m->kq);
        k_soft_max<<<nblk((long)H * N, 32), 32, 0, st>>>(m->kq, H * N, P, m->t_exp);
        k_kqv<<<nblk((long)H * N * d), 256, 0, st>>>(mv, m->kq, H, N, P, d, nth, m->merged);
        if ((rc = mm(p + "attention.dense.weight", m->merged, N, m->att)) || (rc = bias(m->att, p + "attention.dense.bias", E))) return rc;
        if ((rc = ln(m->x, p + "post_attention_layernorm", m->cur))) return rc;
        if ((rc = mm(p + "mlp.dense_h_to_4h.weight", m->cur, N, m->ff)) || (rc = bias(m->ff, p + "mlp.dense_h_to_4h.bias", 4 * E))) return rc;
        k_gelu<<<nblk((long)N * 4 * E), 256, 0, st>>>(m->ff, (long)N * 4 * E, m->t_gelu);
        if ((rc = mm(p + "mlp.dense_4h_to_h.weight", m->ff, N, m->ff2)) || (rc = bias(m->ff2, p + "mlp.dense_4h_to_h.bias", E))) return rc;
        k_add<<<nblk((long)N * E), 256, 0, st>>>(m->att, m->ff2, m->ff2, (long)N * E);                   // inpFF = cur + inpFF
        k_add<<<nblk((long)N * E), 256, 0, st>>>(m->x, m->ff2, m->x, (long)N * E);                       // inpL = inpL + inpFF
    }
    if ((rc = ln(m->x, "gpt_neox.final_layer_norm", m->cur))) return rc;
    if ((rc = mm("embed_out.weight", m->cur, N, m->logits))) return rc;
    NCK(cudaGetLastError());
    NCK(cudaMemcpyAsync(logits_out, m->logits + (size_t)(N - 1) * m->n_vocab, (size_t)m->n_vocab * 4, cudaMemcpyDeviceToHost, st));
    NCK(cudaStreamSynchronize(st));
    return B200Q4_OK;
}

}  // extern "C"
