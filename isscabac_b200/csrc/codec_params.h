// codec_params.h -- the argument block of the op-array kernels (kernels.cu: wide / split / general kernels,
// kernels_lat.cu: latency kernels) and the launcher of the latency kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace isscabac_internal {

struct CodecParams {
  uint32_t n_streams, n_ctx;
  int per_stream_init;
  const uint64_t* op_off;
  const void* ops;
  const uint8_t* ctx_init;
  uint8_t* ctx_scratch;  // CtxGmem only
  // encode
  uint8_t* slab;
  uint64_t slab_stride;
  uint32_t* lengths;
  uint32_t* overflow;
  // decode
  const uint64_t* byte_off;
  const uint8_t* bytes;
  uint8_t* bins;
  uint8_t* finish_ok;
  // launch-dependent choices made by the host (run_codec)
  uint32_t prefetch_ops;   // wide decoder: ask the op stream into L1 four blocks ahead (pays with few tiles per SM only)
  uint32_t ho_eighths;     // hand-over encoder: the giving warps code this many eighths of the lockstep blocks (k_encode_ops_wide_ho)
};

// Latency kernels (cabac_spec.cuh) for the u8 op format.  done = false: the geometry does not fit (too many contexts for
// shared memory); the caller falls back to the wide kernels.  tiles_per_sm_max: use them only up to this many 32-stream
// tiles per SM (0 = always).
int launch_lat_codec(bool encode, const CodecParams& P, cudaStream_t st, bool& done);

}  // namespace isscabac_internal
