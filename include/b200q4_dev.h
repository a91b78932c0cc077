/* b200q4_dev.h -- bench / test / development helpers exported by libb200q4.so next to the product ABI (b200q4.h).
 * Nothing a host integrating the library needs: bench.py's kernel-only timing loop, the tests' launch-plan dump and
 * dispatch overrides, the timeline readers of -DB200Q4_TIMELINE builds.  They may change without notice.              */
#ifndef B200Q4_DEV_H
#define B200Q4_DEV_H

#include "b200q4.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Times `iters` back-to-back device-resident mul_mats with CUDA events on the library stream, rotating over `nw` weight
 * handles (L2 defeat); returns the mean milliseconds per call or a negative error code.                              */
double b200q4_time_mul_mat_dev(b200q4_weight *const *ws, int nw, const float *d_x, int M, float *d_y, int iters);
/* Times `launches` runs of program list c followed by one of `tail` (may be NULL) with CUDA events recorded from C on
 * `stream` (NULL = the library stream) and waits for them; returns the milliseconds or a negative error code.         */
double b200q4_chain_time(b200q4_chain *c, b200q4_chain *tail, int launches, void *stream);
/* Host-only: the launch plan b200q4_mul_mat would use for (type, N, K, M): out[0..] = grid, stages, static tasks, K-slices
 * ...; returns the dynamic shared memory in bytes (tests/test_plan_host.py).                                          */
long long b200q4_debug_plan(int type, int N, int K, int M, int *out);
/* Dispatch overrides for tests (each returns the previous value): smallest M that takes the tcgen05 tile kernel, largest
 * K whose activation quantization is fused into the decode kernel, smallest chunk count that is K-sliced.             */
int b200q4_debug_set_tile_min_m(int m);
int b200q4_debug_set_fuse_max_k(int k);
int b200q4_debug_set_slice_min_c(int c);
/* B200Q4_DBG=1: the per-CTA %globaltimer stamps of the last decode launch (256 x 16 words).                          */
int b200q4_debug_timeline(unsigned long long *out);
/* -DB200Q4_TIMELINE builds, B200Q4_DBG=1: stamps of the last run of a program list (tools/prog_timeline.py).          */
int b200q4_chain_timeline(const b200q4_chain *c, unsigned long long *out, size_t cap_words);

#ifdef __cplusplus
}
#endif
#endif /* B200Q4_DEV_H */
