/*
 * b200neox.h -- device-resident GPT-NeoX forward pass (first slice).
 *
 * Replaces, for the graph the reference runs, gptneox_eval (vsim.cpp:470-747) + ggml_graph_compute (ggml.c:8245-8700):
 * token ids in, logits of the last token out, every tensor in between resident in HBM.  The Q4 mul_mats are this
 * library's own (include/b200q4.h: the fast kernels, or -- after b200q4_set_exact(1) -- the bit-identical ones); the
 * other ops follow the reference's arithmetic exactly (ggml.c:4246-4303 norm, :819-827 gelu through the fp16 table,
 * :5825-5893 soft_max, :6086-6153 gptneox_rope, :4355-4553 the two f32 mul_mat branches of attention).  With exact-mode
 * mul_mats the logits equal the reference program's bit for bit, INCLUDING their dependence on ggml's thread count
 * (the per-thread partial sums of V^T x softmax, ggml.c:4463-4481): n_threads_emulated selects which --threads value
 * of the reference to reproduce.
 *
 * Decode steps (N = 1, exact mode off) are ONE persistent kernel launch: b200neox_load compiles the token's graph --
 * every mul_mat, both layer-norms, gelu, rope + KV store + attention of every layer, the residual stream -- into a
 * device-resident work-list for q4_prog_kernel (vsim_b200/csrc/q4_prog.cuh); the weights stream through shared memory
 * without a kernel boundary between dependent mul_mats.  Prompts (N > 1) and exact mode run one kernel per op.
 *
 * Model files: the reference's GPT-NeoX format with Q4_0 or Q4_1 weights (vsim.cpp:108-458), use_parallel_residual = 0
 * (vsim.cpp:636-665) or 1 (vsim.cpp:666-696).  Every tensor's shape is checked against the hyper-parameters at load
 * ("has wrong shape in model file", vsim.cpp:375-448).  All device work is ordered on the library's own stream.
 * Plain C ABI; every function returns B200Q4_OK (0) / a negative B200Q4_E* code or NULL and records a message for
 * b200neox_last_error().  Needs b200q4_init() first.  No CPU fallback.
 */
#ifndef B200NEOX_H
#define B200NEOX_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200neox b200neox;

/* n_ctx: positions of KV memory (the reference uses 512, vsim.cpp:758); max_batch: most tokens one eval may carry */
b200neox *b200neox_load(const char *path, int n_ctx, int max_batch);
int  b200neox_n_vocab(const b200neox *m);
/* gptneox_eval(model, n_threads, n_past, tokens[N]) -> logits of the last token (host, n_vocab floats) */
int  b200neox_eval(b200neox *m, const int *tokens, int N, int n_past, int n_threads_emulated, float *logits_out);
/* ---- row-sharded decode over the GPUs of one NVLink domain ---------------------------------------
 * One process per GPU.  Rank `rank` of `world` (<= 8) owns rows [rank N/world, (rank+1) N/world) of every 2-D tensor -- the
 * split the reference's threads use (imax.c:1185-1188) -- and heads [rank H/world, ...): Q, K, V and the KV memory stay
 * local; every other vector a later op reads (attention output, MLP hidden and output, logits) is stored by the thread
 * blocks that compute it straight into EVERY rank's exchange pool over NVLink, as {value, tag} pairs the readers poll: no
 * collective call, no fence, no reduction across GPUs (results equal the single-GPU program's bit for bit).
 *   pool_peers[r]  rank r's exchange pool mapped into this process (CUDA IPC / symmetric memory; entry `rank` is local),
 *                  b200neox_sharded_pool_bytes() bytes, zero-initialised, at the same layout on every rank.
 * Every rank calls b200neox_eval with the same token and n_past (N = 1, exact mode off); each gets the full logits.     */
size_t b200neox_sharded_pool_bytes(int n_vocab, int n_embd, int n_layer);
b200neox *b200neox_load_sharded(const char *path, int n_ctx, int rank, int world, void *const *pool_peers, size_t pool_bytes);
/* 1 when decode steps run as one persistent launch, 0 when they fall back to one kernel per op (b200neox_last_error says why) */
int  b200neox_decode_is_program(const b200neox *m);
void b200neox_free(b200neox *m);
const char *b200neox_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
