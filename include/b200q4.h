/*
 * b200q4.h -- C ABI of the B200 (sm_100a) back-end for vsim's Q4_0 / Q4_1
 * mul_mat hot path.  Plain C, plain pointers and sizes; no CUDA, torch or C++
 * types cross this boundary.  Implemented by libb200q4.so
 * (vsim_b200/csrc/*.cu).  There is NO CPU fallback behind any entry point: if
 * no sm_100 device is usable every call fails loudly (see error convention).
 *
 * What it replaces in the reference (NAIST-Archlab/vsim; file:line):
 *   - imax.c:52-142     init_xmax()            -> b200q4_init()
 *   - imax.c:1133-2292  imax_ggml_compute_forward_mul_mat_q4_0_f32()
 *                                              -> b200q4_mul_mat() behind the same-named
 *                                                 hook in vsim_b200/csrc/imax_b200.c
 *   - imax.c:2145-2148  per-call xmax_cpyin of the whole weight tensor
 *                                              -> b200q4_weight_upload(): once per tensor
 *   - ggml.c:209-298    quantize_row_q4_0 / quantize_row_q4_1 (INIT phase,
 *                       ggml.c:5024-5041, 5302-5320) -> fused into b200q4_mul_mat();
 *                       b200q4_quantize_rows() exposes the byte-exact row image
 *   - ggml.c:472-553    ggml_vec_dot_q4_0 / _q4_1 (COMPUTE phase, ggml.c:5350-5390,
 *                       imax.c:1205-1229)      -> the GEMV / tile kernels
 *   - ggml.c:5118-5164, 5391-5440  transposed-src0 branch (ggml_vec_mad_q4_*)
 *                                              -> b200q4_mul_mat_transposed()
 *
 * Data formats are the reference's, unchanged (ggml.c:204-298, 895-915):
 *   Q4_0 row  = K/32 blocks of { float d; uint8_t qs[16] }            (20 B each)
 *   Q4_1 row  = float min[K/32]; float d[K/32]; uint8_t qs[K/32][16]  (24 B/block, SoA)
 *   byte j of qs holds element 2j in the low nibble, 2j+1 in the high nibble.
 * Weights are handed over in that format and re-tiled ON THE DEVICE at upload;
 * the host copy is never modified.
 *
 * Error convention.  The reference's hook has no error channel (printf +
 * exit(1), imax.c:2042-2049).  Here every int-returning function returns
 * B200Q4_OK (0) or a negative B200Q4_E* code and records a message readable
 * with b200q4_last_error(); the ggml hook in imax_b200.c turns any failure
 * into fprintf(stderr) + exit(1) exactly like the reference.
 *
 * Threading: calls are serialised internally (one device context per process,
 * one stream); the ggml hook may be entered from any pool thread.  The
 * device-resident entry points (*_dev, chains) enqueue on the stream they are
 * given but share ONE set of device scratch buffers: work enqueued through them
 * must be on a single stream at a time (or the caller orders the streams with
 * events) -- b200neox_eval and the host-buffer entry points use the library's
 * own stream.
 */
#ifndef B200Q4_H
#define B200Q4_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200Q4_ABI_VERSION 1

/* enum ggml_type values (ggml.h:217-226) */
#define B200Q4_TYPE_Q4_0 0
#define B200Q4_TYPE_Q4_1 1
#define B200Q4_QK 32

#define B200Q4_OK          0
#define B200Q4_EINVAL     -1   /* bad argument (shape, type, NULL, stride)           */
#define B200Q4_ENODEV     -2   /* no CUDA device / not sm_100 / init not called       */
#define B200Q4_ECUDA      -3   /* a CUDA runtime call or kernel failed                */
#define B200Q4_ENOMEM     -4   /* device or pinned-host allocation failed             */

typedef struct b200q4_weight b200q4_weight;   /* opaque: one weight tensor resident in HBM */

/* ---- life cycle ---------------------------------------------------------- */
int  b200q4_abi_version(void);
/* Select `device` (>= 0), require compute capability 10.x, create the stream
 * and scratch buffers.  Idempotent for the same device.  (imax.c:52 init_xmax) */
int  b200q4_init(int device);
void b200q4_shutdown(void);
const char *b200q4_last_error(void);
/* Fills name (<= cap bytes), SM count and total HBM bytes of the active device. */
int  b200q4_device_info(char *name, size_t cap, int *sm_count, size_t *hbm_bytes);

/* ---- weights --------------------------------------------------------------
 * rows: N rows in the reference's row format for `type`, `row_stride` bytes
 * apart (src0->nb[1]; pass 0 for tightly packed).  K % 32 == 0.  The tensor is
 * copied to the device once and re-tiled there.  Returns NULL on failure.    */
b200q4_weight *b200q4_weight_upload(int type, const void *rows, int N, int K, size_t row_stride);
/* Row shard [row0, row0+nrows) of the same host tensor (multi-GPU row split,
 * the split the reference's threads use at imax.c:1185-1188).               */
b200q4_weight *b200q4_weight_upload_rows(int type, const void *rows, int N, int K,
                                         size_t row_stride, int row0, int nrows);
void   b200q4_weight_free(b200q4_weight *w);
int    b200q4_weight_rows(const b200q4_weight *w);      /* N of this (shard of the) tensor */
int    b200q4_weight_cols(const b200q4_weight *w);      /* K */
int    b200q4_weight_type(const b200q4_weight *w);
size_t b200q4_weight_device_bytes(const b200q4_weight *w);
/* Reads the tensor back from HBM in the reference's row format (round-trip check). */
int    b200q4_weight_download(const b200q4_weight *w, void *rows_out, size_t row_stride);

/* ---- the operator: HOST buffers in, HOST buffers out ----------------------
 * Y[m][n] = sum_k dequant(W[n][k]) * dequant(quantize_row(X[m])[k]),
 * m < M, n < N: ggml_compute_forward_mul_mat_q4_{0,1}_f32 for non-transposed
 * src0 (INIT + COMPUTE phases).  x: M rows of K floats, x_stride bytes apart
 * (src1->nb[1]); y: M rows of N floats, y_stride bytes apart (dst->nb[1]).
 * y is completely written when the call returns.                            */
int b200q4_mul_mat(const b200q4_weight *w, const float *x, int M, size_t x_stride,
                   float *y, size_t y_stride);
/* Same, but the activations arrive already quantized by the reference's INIT
 * phase: wdata = M rows in the row format of the weight's type (params->wdata,
 * ggml.c:5033-5034).  This is what an UNPATCHED ggml.c hands to the hook.    */
int b200q4_mul_mat_wdata(const b200q4_weight *w, const void *wdata, int M,
                         float *y, size_t y_stride);

/* quantize_row_q4_0 / _q4_1 over M rows: out receives M * (K/32)*{20|24} bytes,
 * byte-identical to the reference (scales as raw fp32 bits, then nibbles).  */
int b200q4_quantize_rows(int type, const float *x, int M, int K, size_t x_stride, void *out);

/* Transposed src0 (nb01 < nb00; ggml.c:5118-5164 / 5391-5440): wt holds K rows
 * of N quantized values (N % 32 == 0), activations stay fp32.
 * Y[m][n] = sum_k dequant(Wt[k][n]) * X[m][k].                               */
b200q4_weight *b200q4_weight_upload_transposed(int type, const void *rows, int N, int K, size_t row_stride);
int b200q4_mul_mat_transposed(const b200q4_weight *wt, const float *x, int M, size_t x_stride,
                              float *y, size_t y_stride);

/* ---- device-resident variants (pointers are DEVICE pointers) ---------------
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 * library's own stream).  Used by callers that keep activations in HBM
 * (bench.py `value`, the multi-GPU path).                                   */
int b200q4_mul_mat_dev(const b200q4_weight *w, const float *d_x, int M, size_t x_stride,
                       float *d_y, size_t y_stride, void *stream);
int b200q4_quantize_rows_dev(int type, const float *d_x, int M, int K, size_t x_stride,
                             void *d_out, void *stream);

/* ---- chains: several dependent decode (M = 1) mul_mats in ONE persistent launch ----------
 * Batch-1 decode is a chain of dependent mul_mats (ggml_graph_compute runs them node by
 * node, ggml.c:8543-8662); launched one kernel each, every link pays a kernel boundary
 * that is longer than the weight stream of a small tensor.  A chain is a device-resident
 * work-list: op i computes d_y[i][0..N_i) = W_i . quantize_row(d_x[i][0..K_i)) exactly like
 * b200q4_mul_mat_dev with M = 1 (same kernels' arithmetic, bit-identical Y), and may start
 * only when op deps[i] (an index < i, or -1) is complete -- normally the op that wrote its
 * d_x.  One launch of q4_prog_kernel walks the list: the thread blocks stay resident, the
 * weights of op i+1 are already streaming into shared memory while op i's outputs are
 * being published.  All weights of a chain must have the same quant type; d_x / d_y are
 * 16-byte aligned device pointers that stay valid for the life of the chain.           */
typedef struct b200q4_chain b200q4_chain;
b200q4_chain *b200q4_chain_create(int n, const b200q4_weight *const *ws, const float *const *d_x,
                                  float *const *d_y, const int *deps);
/* The same for row shards on several GPUs (one process per GPU): op i is this rank's shard (N_i rows)
 * of a tensor with world * N_i rows, the split of imax.c:1185-1188; its outputs are stored by the kernel into EVERY rank's
 * pool (pool_peers[r]: rank r's pool mapped into this process, CUDA IPC / symmetric memory; pool_bytes each, zeroed, same
 * layout everywhere) at byte offset y_off[i] as {value, tag} pairs (8 bytes per element, 2 * world * N_i floats per op):
 * the all-gather of the output slices is part of the mul_mat, no collective call.  The pool begins with
 * b200q4_chain_sharded_ctr_bytes(n) bytes of counters.  Inputs d_x[i] are local fp32 vectors (K <= 8192).              */
size_t b200q4_chain_sharded_ctr_bytes(int n);
b200q4_chain *b200q4_chain_create_sharded(int n, const b200q4_weight *const *ws, const float *const *d_x, int rank, int world,
                                          void *const *pool_peers, size_t pool_bytes, const size_t *y_off);
int  b200q4_chain_run(b200q4_chain *c, void *stream);       /* asynchronous on `stream` */
int  b200q4_chain_info(const b200q4_chain *c, int *n_ops, int *stages, size_t *smem_bytes);
void b200q4_chain_free(b200q4_chain *c);

/* ---- multi-GPU: row shards with the all-gather fused into the kernel ---------
 * The reference splits a mul_mat by output rows across its threads
 * (imax.c:1185-1188: thread ith owns rows [ith*ceil(N/nth), ...)); here rank
 * `rank` of `world` GPUs owns the shard `w` (b200q4_weight_upload_rows), X is
 * replicated, and the kernel's epilogue stores this rank's output columns
 * [col0, col0 + rows(w)) straight into EVERY rank's Y over NVLink.
 *   y_peers[b*world + r]  rank r's Y buffer b (b = 0, 1; M rows, y_stride bytes
 *                         apart) mapped into this process (CUDA IPC / symmetric
 *                         memory; entry `rank` is local).  Two buffers alternate.
 *   flag_peers[r]         rank r's arrival counters (zero-initialised uint32[world]),
 *                         likewise mapped: slot `src` counts what rank `src` has
 *                         delivered to r.
 *   seq_counter           a zero-initialised uint32 in THIS rank's memory.  The
 *                         sequencing is on the device (no per-call host value, so
 *                         calls can be captured in a CUDA graph): exchange number
 *                         s = *seq_counter + 1 goes to buffer s & 1.  Every thread
 *                         block orders its own peer stores (one system fence) and
 *                         then adds its share to flag_peers[r][rank] on all r; the
 *                         shares of one exchange sum to 4096, so exchange s has
 *                         landed from `rank` when that slot reaches 4096 * s.
 * Before it reads X a call waits until exchange *seq_counter (the previous one,
 * normally the producer of its X) has landed from every rank.  The consumer of
 * the LAST exchange of a chain calls b200q4_wait_flags(), which holds `stream`
 * until flag_peers[rank][0..world) have all reached 4096 * *seq_counter.
 * No reduction crosses GPUs: the result equals the single-GPU one bit for bit.
 * world <= 8.  Decode shapes (M <= 8) store from the GEMV epilogue itself;
 * larger M computes the slice locally and a copy kernel scatters it.        */
int b200q4_mul_mat_dev_allgather(const b200q4_weight *w, const float *d_x, int M, size_t x_stride,
                                 float *const *y_peers, int world, int rank, int col0, size_t y_stride,
                                 uint32_t *const *flag_peers, uint32_t *seq_counter, void *stream);
int b200q4_wait_flags(const uint32_t *local_flags, int world, const uint32_t *seq_counter, void *stream);
int b200q4_sync(void *stream);

/* ---- model files (row f-3, first half) -----------------------------------------------
 * The reference's on-disk format (vsim.cpp:108-458): magic 0x67676d6c,
 * n_header_ints int32 hyper-parameters (7 for the GPT-NeoX files vsim.cpp loads: n_vocab, n_embd,
 * n_head, n_layer, n_rot, use_parallel_residual, ftype; 6 for the files of quantize_{gptj,gpt2,
 * bloom}.cpp), the vocabulary (with a leading count when vocab_has_count != 0, as the quantize_*
 * tools write it; vsim.cpp:155 reads none), then tensor records {n_dims, name length, ftype
 * (0 f32, 1 f16, 2 q4_0, 3 q4_1), dims, name, data}.  The file is memory-mapped; with upload != 0
 * every 2-D Q4 tensor goes from the mapping straight to HBM (no second host copy) and stays
 * resident until b200q4_model_free().  upload == 0 only indexes the file (needs no GPU).     */
typedef struct b200q4_model b200q4_model;
b200q4_model *b200q4_model_load(const char *path, int n_header_ints, int vocab_has_count, int upload);
int  b200q4_model_header(const b200q4_model *m, int i);
int  b200q4_model_tensor_count(const b200q4_model *m);
int  b200q4_model_tensor_info(const b200q4_model *m, int i, const char **name, int *n_dims, int *ftype,
                              int *ne0, int *ne1, size_t *file_offset, size_t *nbytes);
b200q4_weight *b200q4_model_weight(const b200q4_model *m, int i);      /* resident Q4 tensor or NULL */
const void *b200q4_model_tensor_data(const b200q4_model *m, int i);    /* the tensor's bytes in the mapping */
void b200q4_model_free(b200q4_model *m);

/* ---- next-weight hint ------------------------------------------------------------
 * A decode graph visits its weight tensors in a fixed order.  Telling the library which
 * tensor the NEXT decode call (M_next rows) will use lets the current launch, once it has
 * requested its own last bytes, ask L2 to prefetch the chunks the next launch will read
 * first: HBM keeps streaming through the tail of one launch and the head of the next.
 * One-shot (consumed by the next device-side decode launch); purely a performance hint --
 * a wrong or missing hint changes nothing but timing.  The ggml hook learns the order by
 * itself (the tensor that followed this one last time).                              */
int b200q4_hint_next_weight(const b200q4_weight *next, int M_next);

/* ---- offline weight quantizers (row f-4) -----------------------------------------
 * utils.cpp:424-481 ggml_quantize_q4_0 / utils.cpp:483-541 ggml_quantize_q4_1 -- what
 * quantize_gptneox.cpp & co. run over every 2-D tensor of a model file -- on the device:
 * n_rows rows of K floats at src (host, src_stride bytes apart, 0 = tight) -> the
 * reference row image at dst (host, tight), BYTE FOR BYTE including q4_1's
 * numeric_limits<float>::min() start value of the running maximum (utils.cpp:503);
 * hist (may be NULL) is the reference's 16-bin code histogram and is ADDED to, as the
 * reference does across the tensors of a file.  Streams the input in 256 MB slabs.  */
int b200q4_quantize_weights(int type, const float *src, long long n_rows, int K, size_t src_stride,
                            void *dst, long long *hist);

/* ---- exact mode ----------------------------------------------------------------
 * on != 0: every mul_mat entry point computes ggml_vec_dot_q4_0/_q4_1's own rounding
 * sequence (ggml.c:484-510, 525-552 == imax.c:1210-1228: one strictly sequential
 * fp32 chain per output, every product and sum individually rounded, no FMA), so Y
 * is BIT-IDENTICAL to the reference CPU path and whole programs reproduce its logits
 * and greedy token streams exactly.  Several times slower than the default kernels
 * (whose exact-integer block sums are closer to the real-number result but differ in
 * the last bits -- which the reference's fp16-table gelu and 4-bit activation
 * re-quantization amplify downstream).  Returns the previous setting.  Also
 * selectable with the environment variable B200Q4_EXACT=1 at b200q4_init().      */
int b200q4_set_exact(int on);
int b200q4_exact_enabled(void);

/* Pinned host memory for callers that want zero-staging H2D/D2H copies.      */
void *b200q4_host_alloc(size_t bytes);
void  b200q4_host_free(void *p);

/* ---- introspection used by bench.py / tests -------------------------------- */
/* Number of kernels this library has launched since init (or the last reset). */
uint64_t b200q4_launch_count(void);
void     b200q4_launch_count_reset(void);
/* Name of the kernel variant b200q4_mul_mat would pick for (type, M, N, K).  */
const char *b200q4_kernel_for(int type, int M, int N, int K);
/* (bench / test / development helpers -- kernel timing loops, launch-plan dumps, dispatch overrides, timelines -- are
 * declared in b200q4_dev.h: exported by the same library, not part of the product ABI.)                              */

#ifdef __cplusplus
}
#endif
#endif /* B200Q4_H */
