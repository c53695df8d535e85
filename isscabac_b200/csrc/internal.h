// internal.h -- helpers shared by the translation units of libisscabac.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stddef.h>

namespace isscabac_internal {
extern thread_local char g_err[512];
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
size_t smem_limit();
int sm_count();
// zeroed u32 on the device for a persistent kernel's work queue (cleared on `st`, no allocation per call)
int work_counter(cudaStream_t st, uint32_t** out);
// make cudaMallocAsync/cudaFreeAsync on the default pool cheap (release threshold = max), once per device
int keep_pool_cached();
// out[0..n] = exclusive scan of in[0..n) (out[n] = total); scratch: cabac_compact_scratch_bytes(n)
int exclusive_scan_u32_u64(const uint32_t* d_in, uint64_t* d_out, uint64_t n, void* d_scratch, cudaStream_t st);
}  // namespace isscabac_internal

#define CK(call)                                                          \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) return isscabac_internal::cuda_fail(e__, #call); \
  } while (0)
