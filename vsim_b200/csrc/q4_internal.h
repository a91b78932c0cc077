// q4_internal.h -- declarations shared by the translation units of libb200q4.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

struct b200q4_weight {
    int type, N, K, nb, C, n_tiles;
    bool transposed;
    uint8_t *d_wt;       // wtile (or raw reference rows when transposed)
    size_t wt_bytes;
    int *d_counters;     // one per 16-row tile
};

namespace b200q4 {

struct DevInfo { int device, sms, smem_optin; cudaStream_t stream; };
DevInfo dev_info();                                     // device < 0: b200q4_init() has not succeeded
int  api_fail(int code, const char *fmt, ...);          // records the message for b200q4_last_error(), returns code
void api_count_launch(int n);
void api_lock();
void api_unlock();
struct ApiLock { ApiLock() { api_lock(); } ~ApiLock() { api_unlock(); } };

} // namespace b200q4
