// wide_common.cuh -- what the wide-window kernels (kernels.cu: op arrays, symbols.cu: fused
// binarizer + coder) share: the per-lane replicated state table, the per-warp context block and
// the launch geometry (one warp per tile of 32 streams, NW warps per CTA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cabac_wide.cuh"
#include "internal.h"

namespace cabac {

constexpr int WIDE_MAX_WARPS = 16;
// Columns of the table: an LDS.128 is served one QUARTER-warp (8 lanes x 16 B = all 32 banks) at a time, so 8 copies of a row
// already keep 32 lanes with 32 unrelated states conflict-free (ncu: exactly 4 wavefronts per LDS.128 and 0 bank conflicts,
// profiles/r2_v1_lone_tile_lat_decode_stalls.txt): 16.5 KB instead of the 66 KB of a copy per lane (round 1) -- which is what
// lets two CTAs of the symbol kernels share an SM, and a quarter of the table fill for small jobs.
#ifndef WIDE_COLS
#define WIDE_COLS 8
#endif
constexpr uint32_t WIDE_ROW_STRIDE = WIDE_COLS * sizeof(WRow);   // bytes between consecutive states of one column
constexpr size_t WIDE_TAB_BYTES = (size_t)kNumRows * WIDE_ROW_STRIDE;

struct WideRowTable {
  WRow r[kNumRows];
  constexpr WideRowTable() : r{} {
    for (uint32_t i = 0; i < kNumRows; ++i) r[i] = wide_row(i);
  }
};
static __constant__ WideRowTable c_wide_rows = WideRowTable();

// What a context slot holds.  WIDE_CTX_ROWS = 0: a TOKEN (4 bytes) for its state's row -- per bin LDS token -> LDS.128 row,
// update = one word store.  WIDE_CTX_ROWS = 1 (compile-time experiment, profiles/r2_ctx_rows_experiment.txt): the ROW itself
// (16 bytes) -- per bin ONE LDS.128 on the dependent chain, update = LDS.128 of the successor row + STS.128 (off the chain).
#ifndef WIDE_CTX_ROWS
#define WIDE_CTX_ROWS 0
#endif
constexpr uint32_t WIDE_SLOT_BYTES = WIDE_CTX_ROWS ? 16 : 4;                 // per (context, lane)
constexpr uint32_t WIDE_CTX_STRIDE = 32 * WIDE_SLOT_BYTES;                   // bytes between consecutive contexts of one lane

#if WIDE_CTX_ROWS
struct WCtx {
  uint32_t base;  // shared-window address of this lane's slot 0
  // "loading" a slot yields what WTab::row() turns into the row: here the slot's address
  __device__ __forceinline__ uint32_t load(uint32_t c) const { return base + c * WIDE_CTX_STRIDE; }
  __device__ __forceinline__ void store(uint32_t c, uint32_t tok) const {
    uint32_t a, b, d, e;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(d), "=r"(e) : "r"(tok) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(base + c * WIDE_CTX_STRIDE), "r"(a), "r"(b), "r"(d), "r"(e) : "memory");
  }
  __device__ __forceinline__ void store_sel(uint32_t c, uint32_t sel, uint32_t a, uint32_t b) const { store(c, sel ? a : b); }
};
struct WTab {
  uint32_t base;  // shared-window address of row 0 of this lane's column
  __device__ __forceinline__ uint32_t token(uint32_t st) const { return base + st * WIDE_ROW_STRIDE; }
  __device__ __forceinline__ WRow row(uint32_t addr) const {     // addr: a slot (mutable) -- ordered against the slot stores
    WRow r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.lps4), "=r"(r.next_mps), "=r"(r.next_lps), "=r"(r.mps4) : "r"(addr) : "memory");
    return r;
  }
};
#else
struct WCtx {
  uint32_t* p;  // this lane's column of this warp's context block
  __device__ __forceinline__ uint32_t load(uint32_t c) const { return p[c * 32]; }
  __device__ __forceinline__ void store(uint32_t c, uint32_t v) const { p[c * 32] = v; }
  // slot c = sel ? a : b.  (As two predicated stores -- the select would leave the binding ALU pipe for the LSU pipe --
  // it measured slower on B200: encode 590 -> 559, decode 556 -> 541 Gbins/s at C3.)
  __device__ __forceinline__ void store_sel(uint32_t c, uint32_t sel, uint32_t a, uint32_t b) const { p[c * 32] = sel ? a : b; }
};
// Table [state][column] of 16-byte rows in shared memory: lane l always reads column l % WIDE_COLS, so 32
// lanes with 32 unrelated states never conflict (each quarter-warp of an LDS.128 covers all 32
// banks once).  A token is the shared-window address of a row of this lane's column; the next_*
// fields of the rows hold tokens of the same column.
struct WTab {
  uint32_t base;  // shared-window address of row 0 of this lane's column
  __device__ __forceinline__ uint32_t token(uint32_t st) const { return base + st * WIDE_ROW_STRIDE; }
  __device__ __forceinline__ WRow row(uint32_t tok) const {
    WRow r;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.lps4), "=r"(r.next_mps), "=r"(r.next_lps), "=r"(r.mps4) : "r"(tok));
    return r;
  }
};
#endif

// Fills the table, initialises this warp's context block (slot n_ctx = bypass slot); returns
// false for lanes without a stream.  The per-lane offsets are made opaque so that they stay in
// registers (the optimiser otherwise recomputes them from threadIdx at every table access).
__device__ __forceinline__ bool wide_setup(uint32_t n_streams, uint32_t n_ctx_in, const uint8_t* ctx_init,
                                           int per_stream_init, uint8_t* smem, uint32_t& s, WCtx& ctx, WTab& tab,
                                           uint32_t& n_ctx, uint32_t* vmask = nullptr) {
  WRow* t = reinterpret_cast<WRow*>(smem);
  const uint32_t tab0 = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t i = threadIdx.x; i < kNumRows * WIDE_COLS; i += blockDim.x) {
    WRow r = c_wide_rows.r[i / WIDE_COLS];
    const uint32_t col = tab0 + (i % WIDE_COLS) * (uint32_t)sizeof(WRow);
    r.next_mps = col + r.next_mps * WIDE_ROW_STRIDE;
    r.next_lps = col + r.next_lps * WIDE_ROW_STRIDE;
    t[i] = r;
  }
  const uint32_t warp = threadIdx.x >> 5, lane = cb_keep32(threadIdx.x & 31);
  const uint32_t nw = blockDim.x >> 5;
  n_ctx = cb_keep32(n_ctx_in);
  s = (blockIdx.x * nw + warp) * 32 + lane;
  const bool valid = s < n_streams;
  const uint8_t* init = ctx_init + (per_stream_init && valid ? (uint64_t)s * n_ctx : 0);
  tab.base = cb_keep32(tab0 + (lane % WIDE_COLS) * (uint32_t)sizeof(WRow));
#if WIDE_CTX_ROWS
  ctx.base = cb_keep32(tab0 + (uint32_t)WIDE_TAB_BYTES + (warp * (n_ctx + 1) * 32 + lane) * WIDE_SLOT_BYTES);
  __syncthreads();                       // the slots are filled with rows read from the table
#else
  const uint32_t coff = cb_keep32(warp * (n_ctx + 1) * 32 + lane);
  ctx.p = reinterpret_cast<uint32_t*>(smem + WIDE_TAB_BYTES) + coff;
#endif
  for (uint32_t c = 0; c < n_ctx; ++c) ctx.store(c, tab.token(init[c] & 127u));
  ctx.store(n_ctx, tab.token(kEpState));
  __syncthreads();
  if (vmask) *vmask = __ballot_sync(0xffffffffu, valid);   // the lanes of this warp that hold a stream
  return valid;
}

// Host: warps per CTA / grid / dynamic shared memory for n_streams streams with n_ctx contexts.
// Tiles are spread evenly over the SMs in whole CTAs (65,536 streams = 2,048 tiles = 14 warps on
// each of 147 SMs).  Returns false when one warp's context block does not fit shared memory.
inline bool wide_geometry(uint32_t n_streams, uint32_t n_ctx, uint32_t& nw, uint32_t& grid, size_t& smem) {
  const size_t lim = isscabac_internal::smem_limit();
  const size_t warp_ctx = ((size_t)n_ctx + 1) * WIDE_CTX_STRIDE;
  if (!lim || n_ctx > 125 || WIDE_TAB_BYTES + warp_ctx > lim) return false;
  const uint32_t sms = (uint32_t)isscabac_internal::sm_count();
  const uint32_t tiles = (n_streams + 31) / 32;
  const uint32_t ctas_per_sm = (tiles + sms * WIDE_MAX_WARPS - 1) / (sms * WIDE_MAX_WARPS);
  nw = (tiles + sms * ctas_per_sm - 1) / (sms * (ctas_per_sm ? ctas_per_sm : 1));
  const uint32_t nw_smem = (uint32_t)((lim - WIDE_TAB_BYTES) / warp_ctx);
  if (nw > nw_smem) nw = nw_smem;
  if (nw > (uint32_t)WIDE_MAX_WARPS) nw = WIDE_MAX_WARPS;
  if (nw < 1) nw = 1;
  grid = (tiles + nw - 1) / nw;
  smem = WIDE_TAB_BYTES + warp_ctx * nw;
  return true;
}

}  // namespace cabac
