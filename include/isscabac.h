/* isscabac.h -- C ABI of the B200-native many-stream CABAC engine.
 *
 * One shared library (isscabac_b200/libisscabac.so), plain C types only.  It is
 * the drop-in boundary for the reference's CABAC path:
 *
 *   reference interface (paths relative to the reference repo)      replaced by
 *   --------------------------------------------------------------  ------------------------
 *   CABAC_ArithmeticEncoder::start/encodeBin/encodeBinEP/            cabac_encode_ops*        (batch)
 *     encodeBinsEP/encodeBinTrm/finish                               simplecabac_* handle API (1 stream)
 *     (CABAC/CABAC_ArithmeticEncoder.h:52-72, .cpp:54-412)
 *   CABAC_ArithmeticDecoder::start/decodeBin/decodeBinEP/            cabac_decode_ops*        (batch)
 *     decodeBinsEP/decodeBinTrm/finish                               simplecabac_* handle API
 *     (CABAC/CABAC_ArithmeticDecoder.h:46-61, .cpp:54-472)
 *   CABAC_BitstreamFile (byte sink/source, .h:56-73)                 slab / payload + offset table
 *   CABAC_ContextModels::initContextModelsByMpsState/ByP0Prob        cabac_ctx_from_prob / _from_state
 *     (CABAC/CABAC_ContextModelsInit.cpp:51-148)
 *   mexFunction command protocol (CABAC/SimpleCABACMex.cpp:100-472)  simplecabac_dispatch
 *   cabacBinarizer.m / cabacDebinarizer.m /                          cabac_binarize_symbols,
 *     cabacDecodeSymbolFinished.m / cabacContextSelection.m /        cabac_encode_symbols,
 *     cabacEncode.m:45-70 / cabacDecode.m:29-55 / cabacDemo.m:101-186  cabac_decode_symbols
 *   cabacInitContextModel.m:15-129, cabacEncode.m:30-31              cabac_iss_ctx_stats / _from_counters
 *   cabacEncode.m:40-67 (ctxHist, ctxCost, H), ContextModel.cpp:97-134,  cabac_ctx_trace_ops, simplecabac_get_stats,
 *     SimpleCABACMex.cpp:231-241,356-466 (trace statistics)              cabac_encode_symbols(.., d_bits_after_symbol)
 *   one file per stream + .mat side info (SimpleCABACMex.cpp:195,288,    cabac_container_*
 *     ISS.m:197-201)
 *   quantizeWrapper.m:1-88, quantizeLloyd :91-176, quantize.m:59-84     cabac_quantize_matrices
 *
 * Conventions
 *   - every function returns 0 (ISSCABAC_OK) or a negative ISSCABAC_ERR_* code; no
 *     exceptions cross the ABI; isscabac_last_error() gives a thread-local detail string.
 *   - "d_" pointers are DEVICE pointers on the current CUDA device, "h_" pointers are host
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Device entry points are stream-ordered and do not synchronise unless stated.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with ISSCABAC_ERR_CUDA.
 *
 * Op format (one op per bin; engine-level API of the reference, one call = one op)
 *   op = (code << 1) | bin
 *   u8  ops (op_width 1): code 0..124 = context index, 125 = terminate bin, 126 = bypass bin
 *   u16 ops (op_width 2): code 0..998 = context index, 0x7FFD terminate, 0x7FFE bypass
 *   A code >= n_ctx of the call that is not the terminate code is coded as a BYPASS bin -- defined behaviour for
 *   malformed op arrays, identical in every kernel formulation and in both directions.
 * Context state byte: (state << 1) | mps, state 0..63 (CABAC/ContextModel.h:78-80).
 */
#ifndef ISSCABAC_H
#define ISSCABAC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built with -fvisibility=hidden */
#endif

#define ISSCABAC_VERSION 100

#define ISSCABAC_OK 0
#define ISSCABAC_ERR_INVALID (-1)     /* bad argument */
#define ISSCABAC_ERR_CUDA (-2)        /* CUDA runtime error / no device */
#define ISSCABAC_ERR_OVERFLOW (-3)    /* an output buffer was too small */
#define ISSCABAC_ERR_NOMEM (-4)
#define ISSCABAC_ERR_UNSUPPORTED (-5)
#define ISSCABAC_ERR_STATE (-6)       /* handle used out of order (e.g. encodeBin before encodeStart) */
#define ISSCABAC_ERR_IO (-7)          /* bitstream file could not be opened */
#define ISSCABAC_ERR_CORRUPT (-8)     /* decodeFinish: terminate bin / stop bit check failed */

#define ISSCABAC_OP8_TRM 125u
#define ISSCABAC_OP8_EP 126u
#define ISSCABAC_OP16_TRM 0x7FFDu
#define ISSCABAC_OP16_EP 0x7FFEu
#define ISSCABAC_MAX_CTX 999u         /* RWTH_MAX_NUM_CONTEXTS-1, CABAC/CommonDef.h:56 */

/* binarization methods (CABAC/cabacBinarizer.m:12-27) */
enum { ISSCABAC_BIN_TU = 0, ISSCABAC_BIN_EG0 = 1, ISSCABAC_BIN_EG1 = 2, ISSCABAC_BIN_EG2 = 3,
       ISSCABAC_BIN_FL32 = 4,
       /* truncated Rice (cabacBinarizer.m:39-54): binarize / encode only, as upstream -- the reference's decode loops have
        * no case for it (cabacDecodeSymbolFinished.m:10-32), cabac_decode_symbols returns ISSCABAC_ERR_UNSUPPORTED.
        * A symbol costs (v >> k) + 1 + k bins: keep values below 2^20. */
       ISSCABAC_BIN_TR0 = 5, ISSCABAC_BIN_TR1 = 6, ISSCABAC_BIN_TR2 = 7 };
/* context-selection profiles */
enum { ISSCABAC_PROFILE_DEMO = 0,       /* CABAC/cabacDemo.m:113-121 (3 contexts)                 */
       ISSCABAC_PROFILE_ISS = 1,        /* ISS/+coder/cabacContextSelection.m:24-67 (7*Nlbp+2)    */
       ISSCABAC_PROFILE_FLAT = 2,       /* neighbour-free restriction of ISS (2*Nlbp+2 contexts)  */
       ISSCABAC_PROFILE_FLAT_EPSUF = 3  /* FLAT prefix contexts, suffix bins bypass (Nlbp+1)      */ };
/* ISS cmTypes mask (ISS/ISS.m:51) */
enum { ISSCABAC_CM_COND0 = 1, ISSCABAC_CM_COND1 = 2, ISSCABAC_CM_CONDBINLFT = 4,
       ISSCABAC_CM_CONDS0 = 8, ISSCABAC_CM_CONDS1 = 16 };

typedef struct {
  int32_t profile;   /* ISSCABAC_PROFILE_* */
  int32_t method;    /* ISSCABAC_BIN_* */
  uint32_t Nq;       /* number of quantisation levels (symbols are 0..Nq-1) */
  int32_t Nlbp;      /* last bin position modelled with its own context (ISS.m:52) */
  uint32_t types;    /* ISSCABAC_CM_* mask (ISS profile) */
  uint32_t rows;     /* ISS profile: matrix rows of the column-major stream; 0 = one column */
} isscabac_symcfg;

/* ---- library ------------------------------------------------------------- */
int isscabac_version(void);
const char* isscabac_strerror(int code);
const char* isscabac_last_error(void);
/* sm_count / cc = properties of the current device */
int isscabac_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- context initialisation (host, double arithmetic like the reference) -- */
/* CABAC_ContextModelsInit.cpp:124-148 (xMapProbabilityToState) over n probabilities p(0) */
int cabac_ctx_from_prob(const double* p0, uint32_t n, uint8_t* ctx_out);
/* CABAC_ContextModelsInit.cpp:51-80: n triples [ctxIdx mps state]; ctxIdx is ignored */
int cabac_ctx_from_state(const double* triples, uint32_t n, uint8_t* ctx_out);
/* number of contexts a profile uses */
int cabac_profile_num_ctx(int profile, int Nlbp);

/* ---- batch encode / decode over op arrays (device pointers) --------------- */
/* Stream s codes ops[op_off[s] .. op_off[s+1]) as start(); ops...; finish().
 * d_ctx_init: n_ctx state bytes shared by all streams, or n_streams*n_ctx when
 * per_stream_init != 0.  Stream s is written to d_slab + s*slab_stride
 * (slab_stride % 16 == 0, d_slab 16-byte aligned); d_lengths[s] = its byte length.
 * A stream longer than slab_stride sets bit 0 of *d_overflow (optional) and its
 * length is still reported, so the caller can retry with a larger stride.
 * cabac_slab_stride_bound() gives a stride that can never overflow. */
int cabac_encode_ops(uint32_t n_streams, const uint64_t* d_op_off, const void* d_ops, int op_width,
                     const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                     uint8_t* d_slab, uint64_t slab_stride, uint32_t* d_lengths,
                     uint32_t* d_overflow, void* stream);
uint64_t cabac_slab_stride_bound(uint64_t max_ops_per_stream);
/* The kernel formulation cabac_encode_ops picks for a call of this shape with u8 ops (one warp per 32 streams; the same with one
 * hand-over per SM when an SM holds 4k + 2 tiles; two warps per 32 streams when there are few tiles per SM; ...): the name a
 * profiler shows for it.  Informational (bench.py labels its per-kernel numbers with it); "" without a device. */
const char* cabac_encode_ops_kernel(uint32_t n_streams, uint32_t n_ctx);
const char* cabac_decode_ops_kernel(uint32_t n_streams, uint32_t n_ctx);

/* Stream s is d_bytes[byte_off[s] .. byte_off[s+1]); the op array gives the kind of
 * every bin (bit 0 ignored); d_bins[i] = decoded bin of op i.  d_finish_ok[s]
 * (optional) = 1 when Decoder::finish()'s two checks hold (Decoder.cpp:75-81).
 * The decoders (this one and cabac_decode_symbols) read d_bytes in ALIGNED 32-bit words:
 * the word that holds the last payload byte is read whole, so up to 3 bytes behind
 * byte_off[n_streams] must be readable (their values are never used; every CUDA
 * allocation is a multiple of 256 bytes, so a device buffer of exactly the payload's
 * size qualifies -- compute-sanitizer's memcheck wants the size rounded up to 4). */
int cabac_decode_ops(uint32_t n_streams, const uint64_t* d_byte_off, const uint8_t* d_bytes,
                     const uint64_t* d_op_off, const void* d_ops, int op_width,
                     const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                     uint8_t* d_bins, uint8_t* d_finish_ok, void* stream);

/* ---- length scan + compaction --------------------------------------------- */
/* d_byte_off[0..n] = exclusive scan of d_lengths (u64); streams copied back to back
 * into d_payload (capacity payload_cap bytes; overflow -> bit 1 of *d_overflow).
 * d_scratch: cabac_compact_scratch_bytes(n_streams) bytes. */
size_t cabac_compact_scratch_bytes(uint32_t n_streams);
int cabac_compact(uint32_t n_streams, const uint8_t* d_slab, uint64_t slab_stride,
                  const uint32_t* d_lengths, uint8_t* d_payload, uint64_t payload_cap,
                  uint64_t* d_byte_off, void* d_scratch, uint32_t* d_overflow, void* stream);

/* ---- multi-GPU: streams sharded over the GPUs of one box ------------------------------------------------ */
/* One process per GPU; rank r codes the contiguous global stream range [h_first[r], h_first[r+1]) with the
 * single-GPU entry points above -- the coding loop never communicates.  What the ranks exchange afterwards is
 * what turns N local bitstreams into one: per-stream byte lengths (-> the global offset table on every rank) and,
 * when one contiguous bitstream is wanted on every rank, the payload bytes.  NCCL over NVLink / NVSwitch; the
 * library binds libnccl.so.2 at run time (the one already loaded in the process, e.g. PyTorch's, else the system
 * one), so single-GPU users have no NCCL dependency.
 * What it replaces in the reference: nothing is parallel there -- one matrix = one stream = one FILE
 * (ISS/+coder/cabacEncode.m:34-37,97; CABAC/SimpleCABACMex.cpp:195 writes it, :288 reads it), the "hand-off" is the
 * file system.  SURVEY.md 8(b).3 / 8(e) name these entry points.
 * All calls are collective over the communicator's ranks and stream-ordered on `stream`. */
typedef struct isscabac_mgpu isscabac_mgpu;
#define ISSCABAC_MGPU_ID_BYTES 128
/* rank 0: ncclGetUniqueId; ship the 128 bytes to the other ranks by any means (torch.distributed, MPI, a file) */
int cabac_multi_gpu_unique_id(uint8_t* h_id);
/* every rank, on its own current device: ncclCommInitRank */
int cabac_multi_gpu_init(const uint8_t* h_id, int rank, int world, isscabac_mgpu** out);
/* or wrap a communicator the application already has (ncclComm_t passed as void*; not destroyed by _destroy) */
int cabac_multi_gpu_attach(void* nccl_comm, int rank, int world, isscabac_mgpu** out);
int cabac_multi_gpu_destroy(isscabac_mgpu* mg);
int cabac_multi_gpu_info(const isscabac_mgpu* mg, int* rank, int* world);
/* Length exchange + global offset table, no host synchronisation: all-gather-v of the per-stream lengths (one grouped
 * ncclBroadcast per rank; the counts are the host-known partition h_first[0..world]) into d_all_lengths[n_total],
 * then the device-wide exclusive scan -> d_byte_off[n_total + 1] (u64), the same on every rank.
 * d_scratch: cabac_compact_scratch_bytes(n_total) bytes. */
int cabac_multi_gpu_gather_table(isscabac_mgpu* mg, const uint32_t* h_first, const uint32_t* d_local_lengths,
                                 uint32_t* d_all_lengths, uint64_t* d_byte_off, void* d_scratch, void* stream);
/* Payload exchange: rank r's compacted local payload lands at d_byte_off[h_first[r]] of d_payload on EVERY rank (one
 * grouped ncclBroadcast per rank).  NCCL wants the byte counts on the host: the world + 1 boundary offsets are read
 * back first (8 * (world + 1) bytes, one stream synchronisation -- the only one of the multi-GPU path); they are
 * returned in h_rank_byte_first[0..world] when that is not NULL.  ISSCABAC_ERR_OVERFLOW if the total exceeds payload_cap. */
int cabac_multi_gpu_assemble(isscabac_mgpu* mg, const uint32_t* h_first, const uint64_t* d_byte_off,
                             const uint8_t* d_local_payload, uint8_t* d_payload, uint64_t payload_cap,
                             uint64_t* h_rank_byte_first, void* stream);
/* Compaction FUSED with the payload exchange over NVLink peer memory, no host synchronisation at all: every rank owns
 * one buffer of a symmetric allocation (cudaMalloc + CUDA IPC, opened by every peer); the compaction kernel reads each
 * local slab row once and stores it at its GLOBAL offset d_byte_off[h_first[rank] + s] into the buffer of EVERY rank
 * (peer stores through NVSwitch), so the assembled bitstream exists on all ranks when the kernels and the closing
 * barrier (cabac_multi_gpu_barrier) have run.  cap = the size given to _symmetric_alloc (bit 1 of *d_overflow on excess). */
int cabac_multi_gpu_symmetric_alloc(isscabac_mgpu* mg, uint64_t bytes, uint8_t** d_local);
int cabac_multi_gpu_symmetric_free(isscabac_mgpu* mg);
int cabac_multi_gpu_compact_p2p(isscabac_mgpu* mg, const uint32_t* h_first, const uint8_t* d_slab, uint64_t slab_stride,
                                const uint32_t* d_local_lengths, const uint64_t* d_byte_off, uint32_t* d_overflow,
                                void* stream);
/* all ranks' work enqueued on `stream` before this call is complete on every rank when work after it starts */
int cabac_multi_gpu_barrier(isscabac_mgpu* mg, void* stream);

/* ---- symbol level: binarizer + context selection on device ---------------- */
/* symbols -> ops (u8 op format).  Two calls: with d_ops == NULL only d_op_off[0..n_streams]
 * (u64, exclusive scan of bins per stream) is produced; then call again with a d_ops buffer of
 * at least op_off[n_streams] bytes (a caller that knows a bound for the op count can skip the first call: a call with
 * d_ops fills d_op_off too, and writes no op past ops_cap).  sym_width 1, 2 or 4 bytes per symbol; 8-bit symbols take the
 * table kernels (bin counts and op strings by table, one table load and a word-wise append per symbol).
 * d_scratch: cabac_binarize_scratch_bytes(n_symbols) bytes. */
size_t cabac_binarize_scratch_bytes(uint64_t n_symbols, uint32_t n_streams);
int cabac_binarize_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                           const void* d_symbols, int sym_width, uint64_t n_symbols,
                           uint64_t* d_op_off, uint8_t* d_ops, uint64_t ops_cap,
                           void* d_scratch, void* stream);
/* fused symbols -> bytes (binarize + select + encode in one kernel, no op array in HBM);
 * d_bits_after_symbol (optional, u32 per symbol) = getNumBits() after the symbol
 * (CABAC_BitstreamFile.h:70 semantics; feeds the heat map H of cabacEncode.m:49,67). */
int cabac_encode_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                         const void* d_symbols, int sym_width,
                         const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                         uint8_t* d_slab, uint64_t slab_stride, uint32_t* d_lengths,
                         uint32_t* d_bits_after_symbol, uint32_t* d_overflow, void* stream);
/* bytes -> symbols: decode loop with on-device finish detector, context selection and
 * debinarizer (cabacDecode.m:29-55,62-66). */
int cabac_decode_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_byte_off,
                         const uint8_t* d_bytes, const uint64_t* d_sym_off,
                         const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                         void* d_symbols, int sym_width, uint8_t* d_finish_ok, void* stream);

/* ---- ISS context-init statistics (ISS/+coder/cabacInitContextModel.m:15-129) -- */
/* Device-side reduction of the per-context zero counts over the binarised symbols of every
 * stream (ISS layout: column-major with cfg->rows rows; the up neighbour of a symbol is the
 * previous one of the same column).  Streams are pooled in groups of streams_per_group
 * consecutive streams (1 = statistics per stream; the 20 column streams of one matrix = 20).
 * d_counters: n_groups * cabac_iss_num_counters(Nlbp) u64, zeroed by the call. */
int cabac_iss_num_counters(int Nlbp);
int cabac_iss_ctx_stats(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                        const void* d_symbols, int sym_width, uint64_t n_symbols, uint32_t streams_per_group,
                        uint64_t* d_counters, void* stream);
/* Host: counters -> p(0) per context (7*Nlbp+2 per group), the uint8 side information
 * uint8(p*255) of cabacEncode.m:30 and the context state bytes initByProb derives from
 * q/255 (cabacEncode.m:31, CABAC_ContextModelsInit.cpp:124-148).  equal_prob mirrors
 * param.equalProb (cabacEncode.m:25-27).  Output pointers may be NULL. */
int cabac_iss_ctx_from_counters(const isscabac_symcfg* cfg, const uint64_t* h_counters, uint32_t n_groups,
                                int equal_prob, double* h_p0, uint8_t* h_ctx_quant, uint8_t* h_ctx_state);
/* The same on the device, stream-ordered (no host round trip between the statistics and the encoder): d_counters as
 * cabac_iss_ctx_stats leaves them, outputs are device arrays [n_groups][7*Nlbp+2].  Bit-identical to the host function:
 * the divisions are IEEE double on both sides and the probability -> state step is a 256-entry table filled from
 * cabac_ctx_from_prob(q / 255) (the side information has only 256 values). */
int cabac_iss_ctx_from_counters_device(const isscabac_symcfg* cfg, const uint64_t* d_counters, uint32_t n_groups,
                                       int equal_prob, double* d_p0, uint8_t* d_ctx_quant, uint8_t* d_ctx_state, void* stream);

/* ---- quantiser in front of the coder (ISS/quantizeWrapper.m, ISS/quantize.m) -- */
/* Replaces quantizeWrapper(x, qParam) (quantizeWrapper.m:1-88) for a batch of matrices: dead zone
 * below the data's deadzone_quant quantile (:22-36; negative = none), then Lloyd-Max (quantizeLloyd,
 * :91-176; qParam.GMM = 1), uniformly spaced (quantize.m:59-75 between the q_lo / q_hi quantiles) or
 * caller-given centroids (qParam.fixedCentroids) on the rest.  Matrix m is d_x[elem_off[m] ..
 * elem_off[m+1]) in double precision (the transformed values, ISS.m:104-105).  Outputs: d_groups =
 * group index - 1 per element (u8: the coder's symbols, ISS.m:108-110), d_centroids[m * N ..] = the N
 * centroids (dead-zone mean first), d_iters (optional) = Lloyd iterations run.  Sums are taken over
 * the sorted data, so centroids agree with the reference formulation to rounding, not bit for bit. */
enum { ISSCABAC_QUANT_UNIFORM = 0, ISSCABAC_QUANT_LLOYD = 1, ISSCABAC_QUANT_FIXED = 2 };
typedef struct isscabac_quantcfg {
  int32_t N;               /* number of centroids including the dead zone's (qParam.N, ISS.m:42), 1..32 */
  int32_t mode;            /* ISSCABAC_QUANT_* */
  double deadzone_quant;   /* qParam.deadzoneQuant (ISS.m:44: 0.7); < 0 = no dead zone */
  double q_lo, q_hi;       /* qParam.quantileprob for the uniform mode (default [0 1]) */
  double tol;              /* Lloyd stop: mean squared centroid change (quantizeWrapper.m:102: eps) */
  int32_t max_iter;        /* quantizeWrapper.m:103: 100 */
  int32_t reserved;
} isscabac_quantcfg;
/* d_scratch: cabac_quantize_scratch_bytes(n_matrices, max_elems) bytes (only used when a matrix has
 * more than 8192 elements); max_elems = the largest matrix of the batch. */
size_t cabac_quantize_scratch_bytes(uint32_t n_matrices, uint64_t max_elems);
int cabac_quantize_matrices(const isscabac_quantcfg* cfg, uint32_t n_matrices, const uint64_t* d_elem_off,
                            uint64_t max_elems, const double* d_x, const double* d_fixed_centroids,
                            uint8_t* d_groups, double* d_centroids, uint32_t* d_iters, void* d_scratch, void* stream);

/* ---- host-buffer API (what a reference-side caller binds) ------------------ */
/* Same semantics with HOST pointers; all host<->device copies happen inside the call.
 * Encode returns the compacted payload and the offset table: h_payload (capacity
 * payload_cap), h_byte_off[n_streams+1].  Pinned host memory makes the copies
 * asynchronous/pipelined but is not required.  These calls synchronise. */
int cabac_encode_ops_host(uint32_t n_streams, const uint64_t* h_op_off, const void* h_ops, int op_width,
                          const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                          uint8_t* h_payload, uint64_t payload_cap, uint64_t* h_byte_off);
int cabac_decode_ops_host(uint32_t n_streams, const uint64_t* h_byte_off, const uint8_t* h_bytes,
                          const uint64_t* h_op_off, const void* h_ops, int op_width,
                          const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                          uint8_t* h_bins, uint8_t* h_finish_ok);
/* The same decode with the bins returned BIT-PACKED: bit (i & 7) of h_bins_packed[i >> 3] = bin of op i (ops counted over
 * the whole op array, (n_ops + 7) / 8 bytes).  A decoded bin is one bit of information; one byte per bin is what the
 * op-level contract above carries, and on a PCIe-bound call the device-to-host leg shrinks eightfold. */
int cabac_decode_ops_host_packed(uint32_t n_streams, const uint64_t* h_byte_off, const uint8_t* h_bytes,
                                 const uint64_t* h_op_off, const void* h_ops, int op_width,
                                 const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                                 uint8_t* h_bins_packed, uint8_t* h_finish_ok);
/* device: d_packed[(bit_begin >> 3) ...] = bins [bit_begin, bit_end) packed as above; bit_begin must be a multiple of 8 */
int cabac_pack_bins(const uint8_t* d_bins, uint64_t bit_begin, uint64_t bit_end, uint8_t* d_packed, void* stream);
int cabac_encode_symbols_host(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* h_sym_off,
                              const void* h_symbols, int sym_width,
                              const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                              uint8_t* h_payload, uint64_t payload_cap, uint64_t* h_byte_off,
                              uint32_t* h_bits_after_symbol);
int cabac_decode_symbols_host(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* h_byte_off,
                              const uint8_t* h_bytes, const uint64_t* h_sym_off,
                              const uint8_t* h_ctx_init, uint32_t n_ctx, int per_stream_init,
                              void* h_symbols, int sym_width, uint8_t* h_finish_ok);
/* pinned host allocations for the callers above */
int cabac_host_alloc(void** p, size_t bytes);
int cabac_host_free(void* p);

/* ---- statistics outputs (device) -------------------------------------------- */
/* What the reference collects with one MEX call per bin: ctxHist / ctxCost of
 * ISS/+coder/cabacEncode.m:40-41,61-65, and -- under RWTH_TRACE_CABAC_STATES (CommonDef.h:39-43,
 * Windows builds) -- per-context trace-state histograms, 128x128 transition counts and step logs
 * (CABAC/ContextModel.cpp:97-134, SimpleCABACMex.cpp:231-241,356-466).  Computed here in one pass
 * over the op arrays (for decoder-side statistics pass the decoded bins in bit 0 of the ops).
 * Trace-state index of a state byte: mps == 0 ? 63 - state : state + 64 (ContextModel.cpp:99-101).
 * Streams are pooled in groups of streams_per_group consecutive streams.  Outputs (device,
 * zeroed by the call; all but d_state_hist optional):
 *   d_state_hist  u64[groups][n_ctx][128]       visits per trace state BEFORE each update; its sum over
 *                                               the 128 states is the bins coded per context (ctxHist)
 *   d_trans       u32[groups][n_ctx][128][128]  [state before][state after] transition counts
 *   d_cost_bits   u64[groups][n_ctx + 1]        bits getNumBits() advanced by while a bin of the context
 *                                               was coded (ctxCost); slot n_ctx = bypass + terminate bins
 *   d_step_states u8[n_ops][2]                  context state byte before / after every op (0xFF, 0xFF
 *                                               for bypass and terminate ops)
 *   d_final_ctx   u8[n_streams][n_ctx]          context state bytes after the last op of every stream */
int cabac_ctx_trace_ops(uint32_t n_streams, const uint64_t* d_op_off, const void* d_ops, int op_width,
                        const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                        uint32_t streams_per_group, uint64_t* d_state_hist, uint32_t* d_trans,
                        uint64_t* d_cost_bits, uint8_t* d_step_states, uint8_t* d_final_ctx, void* stream);

/* ---- container: wire format for many streams (host only) ------------------- */
/* The reference has one file per stream (CABAC/SimpleCABACMex.cpp:195,288) and ships the
 * context initialisation as uint8 side information in a .mat file (ISS/ISS.m:197-201,
 * ISS/+coder/cabacEncode.m:30).  The container holds payload + u64 offset table + per-stream
 * unit (symbol / op) counts + context-init bytes in one buffer; the bytes of stream s are
 * exactly the file the reference would have written, so a stream cut out with
 * cabac_container_stream() is decodable by the reference's decodeStart ... decodeFinish.
 * Layout: see isscabac_b200/csrc/container.cpp.  All pointers are HOST pointers. */
typedef struct {
  uint32_t n_streams, n_ctx;
  int32_t per_stream_init;   /* ctx_init holds n_streams*n_ctx bytes instead of n_ctx */
  int32_t ctx_is_prob;       /* ctx_init bytes are uint8(p0*255) side info (cabacEncode.m:30), not state bytes */
  int32_t has_cfg;           /* cfg / sym_width are meaningful (symbol-level streams) */
  int32_t sym_width;
  isscabac_symcfg cfg;
  uint64_t payload_bytes;
  const uint64_t* byte_off;  /* n_streams + 1 */
  const uint64_t* unit_off;  /* n_streams + 1 exclusive offsets of symbols (or ops) per stream; may be NULL */
  const uint8_t* ctx_init;
  const uint8_t* payload;
} isscabac_container_view;
uint32_t cabac_crc32(const uint8_t* p, uint64_t n);                 /* IEEE CRC-32 as used by the container */
uint64_t cabac_container_size(const isscabac_container_view* v);    /* 0 on invalid view */
int cabac_container_write(const isscabac_container_view* v, uint8_t* out, uint64_t cap, uint64_t* written);
/* Validates magic, version, checksums (payload CRC only when verify_payload_crc != 0), section
 * table and monotone offsets; the view then points INTO buf (8-byte aligned). */
int cabac_container_parse(const uint8_t* buf, uint64_t n, int verify_payload_crc, isscabac_container_view* v);
int cabac_container_stream(const isscabac_container_view* v, uint32_t s, const uint8_t** bytes, uint64_t* n_bytes,
                           const uint8_t** ctx_init, uint64_t* n_units);

/* ---- single-stream engine handle (backs the SimpleCABAC C++ facade) -------- */
/* One handle = the reference's `class CABAC` aggregate (SimpleCABACMex.cpp:69-80): a
 * bitstream name or memory buffer, an encoder context set and a decoder context set.
 * The coder state lives on the GPU; encode ops are queued on the host and executed by
 * the encode kernel at finish() / getNumBits(); decode calls run the decode kernel. */
typedef struct simplecabac simplecabac;
int simplecabac_create(simplecabac** h, const char* filename /* may be NULL: memory sink */);
int simplecabac_destroy(simplecabac* h);
int simplecabac_init_by_prob(simplecabac* h, const double* p0, uint32_t n);          /* initByProb  */
int simplecabac_init_by_state(simplecabac* h, const double* triples, uint32_t n);    /* initByState */
int simplecabac_encode_start(simplecabac* h);
int simplecabac_encode_bin(simplecabac* h, unsigned bin, unsigned ctx_idx);
int simplecabac_encode_bin_ep(simplecabac* h, unsigned bin);
int simplecabac_encode_bins_ep(simplecabac* h, unsigned bins, int n);
int simplecabac_encode_bin_trm(simplecabac* h, unsigned bin);
int simplecabac_get_num_bits(simplecabac* h, uint64_t* bits);
int simplecabac_get_bins_coded(simplecabac* h, uint64_t* bins);
int simplecabac_encode_finish(simplecabac* h);
/* memory sink access after encode_finish (valid until the next encode_start/destroy) */
int simplecabac_get_bytes(simplecabac* h, const uint8_t** bytes, uint64_t* n);
/* memory source for decoding (instead of the file) */
int simplecabac_set_bytes(simplecabac* h, const uint8_t* bytes, uint64_t n);
int simplecabac_decode_start(simplecabac* h);
int simplecabac_decode_bin(simplecabac* h, unsigned ctx_idx, unsigned* bin);
int simplecabac_decode_bin_ep(simplecabac* h, unsigned* bin);
int simplecabac_decode_bins_ep(simplecabac* h, int n, unsigned* bins);
int simplecabac_decode_bin_trm(simplecabac* h, unsigned* bin);
/* batched form of the three calls above: kinds as u16 ops (bit 0 ignored) */
int simplecabac_decode_ops(simplecabac* h, const uint16_t* ops, uint32_t n, uint8_t* bins);
int simplecabac_decode_finish(simplecabac* h);   /* ISSCABAC_ERR_CORRUPT if the checks fail */
int simplecabac_get_ctx_state(simplecabac* h, int decoder_set, unsigned ctx_idx, unsigned* state, unsigned* mps);
/* Trace statistics of one context (the reference's RWTH_TRACE_CABAC_STATES members,
 * ContextModel.cpp:97-134).  Tracing is off by default; enable it before the first bin.  Like the
 * reference, the logs start at initByProb / initByState and run across streams.
 * steps5: 5 bytes per step [bin, trace state before, mps before, trace state after, mps after]
 * (capacity cap_steps steps; *n_steps = steps available); for the decoder set the bin byte is 0, as
 * in the reference (SimpleCABACMex.cpp:318-320 samples it before decodeBin).  trans: 128*128 u32,
 * index state_before*128 + state_after, or NULL. */
int simplecabac_set_trace(simplecabac* h, int on);
int simplecabac_get_stats(simplecabac* h, int decoder_set, unsigned ctx_idx, uint8_t* steps5, uint64_t cap_steps,
                          uint64_t* n_steps, uint32_t* trans);

/* ---- MEX command protocol (CABAC/SimpleCABACMex.cpp:100-472) ---------------- */
/* One call = one mexFunction invocation.  args[0] is the command string; numeric
 * arguments are MATLAB doubles (column-major m x n).  Returns 0, or 1 when the
 * reference would have raised mexErrMsgTxt -- the same message text is written to err. */
typedef struct {
  int32_t is_char;
  const char* s;      /* is_char != 0 */
  const double* d;    /* is_char == 0 */
  int32_t m, n;
} isscabac_mxarg;
int simplecabac_dispatch(int nlhs, double* out, int out_cap, int* out_n,
                         int nrhs, const isscabac_mxarg* args, char* err, int errcap);

/* getEncoderStats / getDecoderStats (SimpleCABACMex.cpp:356-466, compiled into Windows builds of the
 * reference only): args = {command, handle, ctxIdx}; same arity checks and error texts.  The two
 * outputs are typed arrays there (uint8 5 x M, uint32 128 x 128), hence a separate entry point.
 * The handle must have tracing on (command "setTrace", handle, 1 -- or simplecabac_set_trace). */
int simplecabac_dispatch_stats(int nlhs, int nrhs, const isscabac_mxarg* args, uint8_t* steps5, uint64_t cap_steps,
                               uint64_t* n_steps, uint32_t* trans, char* err, int errcap);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
