// symbols.cu -- symbol-level kernels: the binarizer + context selection of the MATLAB
// layer (cabacBinarizer.m, cabacContextSelection.m, cabacDemo.m:113-121,
// cabacDecodeSymbolFinished.m, cabacDebinarizer.m) moved onto the device.
//
//   cabac_binarize_symbols   symbol-parallel: every symbol's bin count is known in closed
//                            form, so ops are produced by count -> device-wide scan -> emit;
//                            context selection only looks at the symbol itself and its up
//                            neighbour (cabacContextSelection.m:24-67 never uses g_lft/g_up2).
//   cabac_encode_symbols     lane per stream, binarize + select + encode fused (no op array
//                            in HBM); optional getNumBits() trace per symbol.
//   cabac_decode_symbols     lane per stream: decodeBin with the finish detector and the
//                            context rule in the loop (inherently serial per stream), then
//                            the debinarized symbol is stored.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cstdio>
#include <mutex>
#include <type_traits>
#include <vector>

#include "../../include/isscabac.h"
#include "bin_emit.cuh"
#include "cabac_lane.cuh"
#include "internal.h"
#include "wide_common.cuh"

using namespace cabac;
using namespace isscabac_internal;

namespace {

constexpr int NT = 128;
constexpr int TAB_WORDS = 128 * 32;

struct RowTable {
  uint2 r[128];
  constexpr RowTable() : r{} {
    for (uint32_t i = 0; i < 128; ++i) r[i] = fused_row(i);
  }
};
__constant__ RowTable c_rows_sym = RowTable();

__device__ __forceinline__ uint32_t load_sym(const void* p, int width, uint64_t i) {
  if (width == 1) return static_cast<const uint8_t*>(p)[i];
  if (width == 2) return static_cast<const uint16_t*>(p)[i];
  return static_cast<const uint32_t*>(p)[i];
}
__device__ __forceinline__ void store_sym(void* p, int width, uint64_t i, uint32_t v) {
  if (width == 1) static_cast<uint8_t*>(p)[i] = (uint8_t)v;
  else if (width == 2) static_cast<uint16_t*>(p)[i] = (uint16_t)v;
  else static_cast<uint32_t*>(p)[i] = v;
}

__device__ __forceinline__ SymCfg to_cfg(const isscabac_symcfg& c) {
  return SymCfg{c.profile, c.method, c.Nq, c.Nlbp, c.types, c.rows};
}

// ---- symbol-parallel binarizer ------------------------------------------------
// Tiles of 2,048 consecutive symbols, 8 per thread (one 8/16/32-byte vector load per thread).
// Pass 1 sums the (closed-form) bin counts of every tile; a device-wide scan over the tile sums
// gives each tile's first op position; pass 2 recomputes the counts, scans them inside the block
// and produces the ops op-parallel, one aligned 16-byte piece of the op array per thread (see
// k_bin_emit).  Only pass 2 needs to know which stream a symbol belongs to (context selection
// looks at the position inside the stream, and the thread that owns a stream's first symbol
// records op_off[s]): one binary search per tile, a short one per thread inside the tile's range.
// HBM traffic: 2 x symbols in, ops out, 12 B per tile.
constexpr int BIN_THREADS = 256, BIN_ITEMS = 8, BIN_TILE = BIN_THREADS * BIN_ITEMS;
constexpr uint32_t BIN_CHUNKS = BIN_TILE / 64;      // k_bin_count8 / k_bin_offsets8: runs of 64 symbols per tile

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& block_total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < BIN_THREADS / 32; ++w) {
    const uint32_t t = s_warp[w];
    if (w < wid) off += t;
    tot += t;
  }
  block_total = tot;
  return off + inc - v;
}

// the 8 symbols of this thread (0 past the end): one vector load when the run is whole and aligned
template <int W>
__device__ __forceinline__ void load_syms8(const void* sym, uint64_t i0, uint64_t n, uint32_t v[BIN_ITEMS]) {
  const uint8_t* base = static_cast<const uint8_t*>(sym) + i0 * W;
  if (i0 + BIN_ITEMS <= n && (reinterpret_cast<uintptr_t>(base) & (W == 1 ? 7u : 15u)) == 0) {
    if (W == 1) {
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(base));
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[k] = (q.x >> (8 * k)) & 0xffu; v[4 + k] = (q.y >> (8 * k)) & 0xffu; }
    } else if (W == 2) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(base));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[2 * k] = w[k] & 0xffffu; v[2 * k + 1] = w[k] >> 16; }
    } else {
      const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(base)), q1 = __ldg(reinterpret_cast<const uint4*>(base) + 1);
      v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < BIN_ITEMS; ++k) v[k] = i0 + k < n ? load_sym(sym, W, i0 + k) : 0u;
  }
}

// PROF / METH: profile and binarization fixed at compile time for the combinations the
// applications use (the closed-form helpers then fold to a few instructions); -1 = read from cfg.
template <int PROF, int METH>
__device__ __forceinline__ SymCfg fixed_cfg(const isscabac_symcfg& c) {
  SymCfg cfg = to_cfg(c);
  if (PROF >= 0) cfg.profile = PROF;
  if (METH >= 0) cfg.method = METH;
  return cfg;
}

template <int W, int METH>
__global__ void __launch_bounds__(BIN_THREADS) k_bin_count(isscabac_symcfg c, const void* sym, uint64_t n, uint32_t* tile_sums) {
  __shared__ uint32_t s_warp[BIN_THREADS / 32];
  const SymCfg cfg = fixed_cfg<-1, METH>(c);
  const uint64_t i0 = (uint64_t)blockIdx.x * BIN_TILE + (uint64_t)threadIdx.x * BIN_ITEMS;
  uint32_t v[BIN_ITEMS];
  load_syms8<W>(sym, i0, n, v);
  uint32_t tot = 0;
#pragma unroll
  for (int k = 0; k < BIN_ITEMS; ++k)
    if (i0 + k < n) tot += sym_code(v[k], cfg.Nq, cfg.method).len;
  uint32_t block_total;
  block_exclusive_scan(tot, s_warp, block_total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = block_total;
}

// last s in [lo, hi) with sym_off[s] <= i  (sym_off[lo] <= i is known)
__device__ __forceinline__ uint32_t find_stream(const uint64_t* sym_off, uint32_t lo, uint32_t hi, uint64_t i) {
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sym_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// stream of the first symbol of every tile (and of the last symbol overall in slot n_tiles): one
// binary search per thread, all tiles in parallel -- inside k_bin_emit the same search would be one
// thread walking 20 dependent loads while the other 255 of its CTA wait
__global__ void k_bin_tile_streams(const uint64_t* sym_off, uint32_t n_streams, uint64_t n, uint32_t n_tiles,
                                   uint32_t* tile_stream, uint32_t* tile_first) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  const uint64_t i = t < n_tiles ? (uint64_t)t * BIN_TILE : n - 1;
  tile_stream[t] = find_stream(sym_off, 0, n_streams, i);
  // tile_first[t] = the first entry of sym_off[0 .. n_streams] that is >= the tile's first symbol (>= n for slot n_tiles):
  // the streams -- empty ones too -- that START in tile t are [tile_first[t], tile_first[t + 1]), and the entries from
  // tile_first[n_tiles] on are the empty streams at the very end and the total
  const uint64_t x = t < n_tiles ? (uint64_t)t * BIN_TILE : n;
  uint32_t lo = 0, hi = n_streams;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sym_off[mid] >= x) hi = mid; else lo = mid + 1;
  }
  tile_first[t] = lo;
}

// Op strings by table.  For the alphabets the applications use, the ops of a symbol -- bins AND
// context ids -- are a function of a small key: the symbol itself (FLAT profiles), the symbol and
// the first bin of its predecessor (DEMO, cabacDemo.m:113-121), the symbol and its up neighbour
// (ISS, cabacContextSelection.m:24-67 looks at nothing else).  k_bin_lut computes the string of
// every key once per call with the same closed-form code as everywhere else (sym_code / select_ctx /
// sym_bin); the emit kernel then fetches ops with one byte load each.  Entry = 15 op bytes + the
// length; symbols outside the table (value beyond its domain, more than 15 bins) take the
// closed-form route op by op.
// (LUT_MAX, lut_geom, lut_index, lut_entry: bin_emit.cuh, shared with the host emulation)
__global__ void k_bin_lut(isscabac_symcfg c, uint4* lut) {
  const SymCfg cfg = to_cfg(c);
  const LutGeom g = lut_geom(cfg.profile, cfg.method, cfg.Nq);
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.entries) return;
  uint32_t w[4];
  lut_entry(cfg, e, w);
  lut[e] = make_uint4(w[0], w[1], w[2], w[3]);
}
// bin count of every 8-bit symbol value (the u8 kernels sum and scan table entries instead of evaluating the closed form)
// (one byte per value: the host takes the closed-form kernels when a count does not fit)
__global__ void k_bin_lentab(isscabac_symcfg c, uint8_t* len_tab) {
  const SymCfg cfg = to_cfg(c);
  const uint32_t l = sym_code(threadIdx.x, cfg.Nq, cfg.method).len;
  len_tab[threadIdx.x] = (uint8_t)(l < 0xffu ? l : 0xffu);
}

// Pass 2.  Phase A (symbol-parallel, 8 symbols per thread): codes, block scan, the op_off entries of the streams that
// start in this tile.  Phase B: the threads write the op bytes of their symbols into a stage in shared memory at their
// tile positions -- one table byte per op -- and the stage leaves for HBM as the aligned 16-byte pieces of the op array.
constexpr uint32_t BIN_STAGE = 16 * 1024; // bytes of ops staged per round (a tile of 2,048 EG0 symbols has ~4,800)

template <int W, int PROF, int METH>
__global__ void __launch_bounds__(BIN_THREADS) k_bin_emit(isscabac_symcfg c, const void* sym, uint64_t n,
                                                           const uint64_t* sym_off, uint32_t n_streams,
                                                           const uint32_t* tile_stream,
                                                           const uint64_t* tile_prefix, uint64_t* op_off, uint8_t* ops,
                                                           uint64_t cap, const uint4* lut) {
  __shared__ uint32_t s_warp[BIN_THREADS / 32];
  __shared__ __align__(16) uint8_t s_stage[BIN_STAGE];
  __shared__ __align__(16) uint4 s_lut[LUT_MAX];
  const SymCfg cfg = fixed_cfg<PROF, METH>(c);
  const LutGeom geom = lut_geom(cfg.profile, cfg.method, cfg.Nq);
  const uint64_t t0 = (uint64_t)blockIdx.x * BIN_TILE;
  const uint64_t i0 = t0 + (uint64_t)threadIdx.x * BIN_ITEMS;
  // streams this tile touches: [s_lo, s_hi] (the next tile's first stream bounds this tile's last)
  const uint32_t s_lo = tile_stream[blockIdx.x], s_hi = tile_stream[blockIdx.x + 1];
  if (ops)
    for (uint32_t e = threadIdx.x; e < geom.entries; e += BIN_THREADS) s_lut[e] = __ldg(lut + e);
  uint32_t v[BIN_ITEMS];
  load_syms8<W>(sym, i0, n, v);
  SymCode code[BIN_ITEMS];
  uint32_t tot = 0;
#pragma unroll
  for (int k = 0; k < BIN_ITEMS; ++k) {
    code[k] = SymCode{0, 0, 0};
    if (i0 + k < n) code[k] = sym_code(v[k], cfg.Nq, cfg.method);
    tot += code[k].len;
  }
  uint32_t block_total;
  const uint32_t lo = block_exclusive_scan(tot, s_warp, block_total);
  const uint64_t tile_base = tile_prefix[blockIdx.x];
  uint32_t up_mask = 0;
  if (i0 < n) {
    uint32_t s = find_stream(sym_off, s_lo, s_hi + 1, i0);
    uint64_t start = sym_off[s], next = sym_off[s + 1];
    if (i0 > start && i0 + BIN_ITEMS <= next && i0 + BIN_ITEMS < n) {
      // the whole run lies inside one stream, away from its first symbol and from the end of the input:
      // no offset to record, the neighbour flags follow from the row of the run's first symbol
      if (cfg.profile == PROFILE_DEMO || (cfg.profile == PROFILE_ISS && cfg.rows == 0)) {
        up_mask = 0xffu;
      } else if (cfg.profile == PROFILE_ISS) {
        const uint64_t rel = i0 - start;
        uint32_t r = rel >> 32 ? (uint32_t)(rel % cfg.rows) : (uint32_t)rel % cfg.rows;
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; ++k) {
          if (r != 0) up_mask |= 1u << k;
          if (++r == cfg.rows) r = 0;
        }
      }
    } else {
      uint64_t o = tile_base + lo;
#pragma unroll
      for (int k = 0; k < BIN_ITEMS; ++k) {
        const uint64_t i = i0 + k;
        if (i >= n) break;
        if (i == next) {   // next non-empty stream
          ++s;
          while (s + 1 < n_streams && sym_off[s + 1] == i) ++s;
          start = i;
          next = sym_off[s + 1];
        }
        if (i == start) {  // first symbol of stream s: record where its ops begin (also for empty streams in front of it)
          for (uint32_t e = s;; --e) {
            op_off[e] = o;
            if (e == 0 || sym_off[e - 1] != i) break;
          }
        }
        if (sym_has_up(cfg, i - start)) up_mask |= 1u << k;
        o += code[k].len;
        if (i == n - 1) {  // streams that start at the very end are empty; op_off[n_streams] = total
          for (uint32_t e = s + 1; e <= n_streams; ++e) op_off[e] = o;
        }
      }
    }
  }
  if (!ops) return;      // offsets only (first of the two calls)
  // Phase B.  Every thread writes the op bytes of its 8 symbols into a STAGE in shared memory at their tile positions
  // (one table byte load + one byte store per op; symbols outside the table take the closed form op by op), then the
  // stage leaves for HBM in the aligned 16-byte pieces of the op array, one LDS.128 + STG.128 per piece.  (Round 1
  // produced each piece by one thread walking the symbols that overlap it: 43 thread-instructions per op, 63 % of
  // the shared-memory wavefronts bank conflicts.)  A tile whose ops exceed the stage is done in rounds.
  const uint32_t skew = (uint32_t)((reinterpret_cast<uintptr_t>(ops) + tile_base) & 15u);   // stage byte = tile position + skew
  // ops the caller's buffer has room for (a too small buffer is filled up to its capacity)
  const uint32_t room = tile_base >= cap ? 0u : (cap - tile_base < block_total ? (uint32_t)(cap - tile_base) : block_total);
  const uint32_t lut0 = (uint32_t)__cvta_generic_to_shared(s_lut);
  const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(s_stage);
  // the neighbour of this thread's first symbol
  uint32_t before = 0;
  if (up_mask & 1u) before = i0 > 0 && i0 <= n ? load_sym(sym, W, i0 - 1) : 0u;
  for (uint32_t w0 = 0; w0 < block_total + skew; w0 += BIN_STAGE) {     // stage window = stage bytes [w0, w0 + BIN_STAGE)
    if (w0) __syncthreads();                                            // the previous round has left the stage
    const uint32_t w1 = w0 + BIN_STAGE;
    uint32_t pos = lo + skew;
#pragma unroll
    for (int k = 0; k < BIN_ITEMS; ++k) {
      const uint32_t len = code[k].len;
      if (len && pos < w1 && pos + len > w0) {
        const uint32_t u = k ? v[k - 1] : before;
        const bool up = (up_mask >> k) & 1u;
        const uint32_t idx = len > 15u ? LUT_ESC : lut_index(cfg, geom.dom, v[k], u, up);
        const uint32_t j0 = pos < w0 ? w0 - pos : 0u, j1 = pos + len > w1 ? w1 - pos : len;   // ops of this symbol inside the window
        if (idx != LUT_ESC) {
          const uint32_t src = lut0 + (idx << 4), dst = stage0 + pos - w0;
          for (uint32_t j = j0; j < j1; ++j) {
            uint32_t byte;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(byte) : "r"(src + j));
            asm volatile("st.shared.u8 [%0], %1;" :: "r"(dst + j), "r"(byte) : "memory");
          }
        } else {
          const SymCode prev = sym_code(u, cfg.Nq, cfg.method);
          for (uint32_t j = j0; j < j1; ++j) {
            const int cx = select_ctx(cfg, j + 1u, code[k].np, prev, up);
            const uint32_t byte = ((cx < 0 ? ISSCABAC_OP8_EP : (uint32_t)cx) << 1) | sym_bin(code[k], j + 1u);
            asm volatile("st.shared.u8 [%0], %1;" :: "r"(stage0 + pos - w0 + j), "r"(byte) : "memory");
          }
        }
      }
      pos += len;
    }
    __syncthreads();
    // pieces of this window: stage bytes [16 p, 16 p + 16) <-> tile positions [w0 + 16 p - skew, ...)
    const uint32_t wend = w1 < block_total + skew ? w1 : block_total + skew;
    const uint32_t npieces = (wend - w0 + 15u) >> 4;
    for (uint32_t pc = threadIdx.x; pc < npieces; pc += BIN_THREADS) {
      const int32_t pbeg = (int32_t)(w0 + 16u * pc) - (int32_t)skew;           // tile position of byte 0 of the piece
      const uint32_t e0 = pbeg < 0 ? (uint32_t)(-pbeg) : 0u;                   // valid bytes: [e0, e1)
      const int32_t left = (int32_t)room - pbeg;
      const uint32_t e1 = left <= 0 ? 0u : (left < 16 ? (uint32_t)left : 16u);
      if (e0 >= e1) continue;
      uint4 q;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(stage0 + 16u * pc) : "memory");
      uint8_t* dst = ops + tile_base + pbeg;       // 16-byte aligned by the choice of skew
      if (e0 == 0 && e1 == 16) {
        *reinterpret_cast<uint4*>(dst) = q;
      } else {
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (uint32_t e = 0; e < 16; ++e)
          if (e >= e0 && e < e1) dst[e] = (uint8_t)(w[e >> 2] >> (8u * (e & 3u)));
      }
    }
  }
}

// ---- u8 symbols: counts by table, ops appended word-wise (round 2) -------------------------------
// Round 2's ncu on the kernels above: instruction-bound, 64 thread-instructions per symbol in phase A (the closed form per
// symbol, the codes kept in 24 registers) and 27 per op in phase B (a byte load, a byte store and the loop around them).
// For 8-bit symbols -- what every application alphabet is -- both halves are table work:
//   k_bin_count8   a warp per tile, 16 symbols per 16-byte load, bin counts out of a 256-entry table in shared memory;
//   k_bin_emit8    phase A sums table entries; phase B appends each symbol's op string (one 8-byte table load: 7 op
//                  bytes + length) to a running word in a register and stores every completed word with one aligned
//                  store -- bin_emit.cuh has the scheme, its proof obligations and the host emulation's entry points.
//                  A CTA walks `tpc` consecutive tiles, so the tables are set up once per 8 K symbols or more.

__global__ void __launch_bounds__(BIN_THREADS) k_bin_count8(const uint8_t* __restrict__ sym, uint64_t n, const uint8_t* __restrict__ len_tab,
                                                             uint32_t n_tiles, uint32_t* __restrict__ tile_sums, uint16_t* __restrict__ chunks) {
  __shared__ uint8_t s_len[256];
  s_len[threadIdx.x] = len_tab[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t tile = blockIdx.x * (BIN_THREADS / 32) + (threadIdx.x >> 5);
  if (tile >= n_tiles) return;
  const uint64_t t0 = (uint64_t)tile * BIN_TILE;
  uint32_t tot = 0;
#pragma unroll
  for (int j = 0; j < BIN_TILE / (32 * 16); ++j) {
    const uint64_t b = t0 + (uint64_t)(lane + 32u * j) * 16u;
    uint32_t c = 0;
    if (b + 16u <= n) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(sym + b));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) c += s_len[(w[k >> 2] >> (8 * (k & 3))) & 0xffu];
    } else {
      for (uint64_t i = b; i < n && i < b + 16u; ++i) c += s_len[sym[i]];
    }
    tot += c;
    if (chunks) {      // the sizing call: bin counts of the tile's 32 runs of 64 symbols (four neighbouring lanes each) for k_bin_offsets8
      c += __shfl_xor_sync(0xffffffffu, c, 1);
      c += __shfl_xor_sync(0xffffffffu, c, 2);
      if ((lane & 3u) == 0u) chunks[(size_t)tile * BIN_CHUNKS + 8u * j + (lane >> 2)] = (uint16_t)c;
    }
  }
  tot = __reduce_add_sync(0xffffffffu, tot);
  if (lane == 0) tile_sums[tile] = tot;
}

// The sizing call's offsets, one thread per stream: op_off[e] = the op position of symbol sym_off[e] = its tile's first op
// (the scan over the tile sums) + the counts of the 64-symbol runs in front of it inside the tile (k_bin_count8) + the counts
// of the at most 63 symbols in front of it inside its run.  No second walk over the symbols (the emit kernel's phase A took
// 0.55 ms of the 0.81 ms sizing call at C4).
__global__ void __launch_bounds__(256) k_bin_offsets8(const uint8_t* __restrict__ sym, uint64_t n, const uint64_t* __restrict__ sym_off,
                                                       uint32_t n_streams, const uint8_t* __restrict__ len_tab, uint32_t n_tiles,
                                                       const uint64_t* __restrict__ tile_prefix, const uint16_t* __restrict__ chunks,
                                                       uint64_t* __restrict__ op_off) {
  __shared__ uint8_t s_len[256];
  s_len[threadIdx.x] = len_tab[threadIdx.x];
  __syncthreads();
  const uint64_t e = (uint64_t)blockIdx.x * 256u + threadIdx.x;
  if (e > n_streams) return;
  const uint64_t so = sym_off[e];
  if (so >= n) {           // the empty streams at the very end, and the total
    op_off[e] = tile_prefix[n_tiles];
    return;
  }
  const uint32_t tile = (uint32_t)(so / BIN_TILE), p = (uint32_t)(so % BIN_TILE), run = p / 64u;
  uint64_t o = tile_prefix[tile];
  const uint16_t* c = chunks + (size_t)tile * BIN_CHUNKS;
  uint32_t acc = 0;
  for (uint32_t r = 0; r < run; ++r) acc += c[r];
  const uint8_t* s0 = sym + (so - p % 64u);
  for (uint32_t j = 0; j < p % 64u; ++j) acc += s_len[s0[j]];
  op_off[e] = o + acc;
}

struct StageStore {            // the whole tile is in the stage; `stage` = its shared-window address
  uint32_t stage;
  __device__ __forceinline__ void word(uint32_t a, uint32_t w) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(stage + a), "r"(w) : "memory"); }
  __device__ __forceinline__ void byte(uint32_t a, uint32_t b) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(stage + a), "r"(b) : "memory"); }
  __device__ __forceinline__ void words(BinAcc& A, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t nfb) {
    asm volatile("{\n\t.reg .pred p1, p2;\n\t"
                 "setp.ge.u32 p1, %6, 32;\n\t"
                 "setp.ge.u32 p2, %6, 64;\n\t"
                 "@p1 st.shared.u32 [%2], %3;\n\t"
                 "@p2 st.shared.u32 [%2+4], %4;\n\t"
                 "selp.u32 %0, %4, %3, p1;\n\t"
                 "selp.u32 %0, %5, %0, p2;\n\t"
                 "@p1 add.u32 %1, %1, 4;\n\t"
                 "@p2 add.u32 %1, %1, 4;\n\t}"
                 : "=&r"(A.a0), "+r"(A.wp) : "r"(stage + A.wp), "r"(w0), "r"(w1), "r"(w2), "r"(nfb) : "memory");
  }
};
struct StageWindow {           // only stage positions [w0, w0 + BIN_STAGE) are in the stage this round
  uint32_t stage, w0;
  __device__ __forceinline__ void word(uint32_t a, uint32_t w) { if (a - w0 < BIN_STAGE) asm volatile("st.shared.u32 [%0], %1;" :: "r"(stage + (a - w0)), "r"(w) : "memory"); }
  __device__ __forceinline__ void byte(uint32_t a, uint32_t b) { if (a - w0 < BIN_STAGE) asm volatile("st.shared.u8 [%0], %1;" :: "r"(stage + (a - w0)), "r"(b) : "memory"); }
  __device__ __forceinline__ void words(BinAcc& A, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t nfb) { bin_words_plain(*this, A, x0, x1, x2, nfb); }
};

// fast-table entry of a symbol: v = its value, u = its neighbour's, up = the neighbour exists
__device__ __forceinline__ uint2 bin8_entry(const SymCfg& cfg, uint32_t dom, uint32_t esc, uint32_t lut8, uint32_t v, uint32_t u, bool up, uint32_t& idx) {
  idx = lut_index(cfg, dom, v, u, up);
  if (idx == LUT_ESC) idx = esc;
  uint2 e;       // volatile: in program order with the appends (a movable load is hoisted a whole run ahead and spilled)
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(lut8 + idx * 8u));
  return e;
}
// one symbol, whatever its string
template <class St>
__device__ __forceinline__ void bin8_symbol(const SymCfg& cfg, uint32_t dom, uint32_t entries, uint32_t lut8, const uint4* lut16,
                                            BinAcc& A, St& st, uint32_t v, uint32_t u, bool up) {
  uint32_t idx;
  const uint2 e = bin8_entry(cfg, dom, entries, lut8, v, u, up, idx);
  if (e.y >> 24) {
    bin_append(A, st, e.x, e.y & 0x00ffffffu, e.y >> 24);
  } else {
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (idx < entries) q = __ldg(lut16 + idx);
    A = bin_append_long(A, st, cfg, q.x, q.y, q.z, q.w, v, u, up);
  }
}
__device__ __forceinline__ uint32_t pinned(uint32_t x) {      // a value the compiler must keep instead of recomputing
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// 128 threads x 16 symbols per tile of 2,048: what a thread does once per tile (scan, tail bytes, its share of the pieces
// and of the stream pass, the look-ahead loads) is a third of its instructions with 8 symbols per thread
constexpr int B8_THREADS = 128, B8_ITEMS = 16, B8_MIN_CTAS = 8;
static_assert(B8_THREADS * B8_ITEMS == BIN_TILE, "the emit kernel's tile is the count kernel's");
constexpr uint32_t BIN_SOFF = 258;      // stream offsets of a tile kept in shared memory (profiles with neighbours)

// symbol k (0..15) of a run held in four words; k need not be a constant
__device__ __forceinline__ uint32_t run_sym(const uint32_t q[4], uint32_t k) {
  const uint32_t w = k < 8u ? (k < 4u ? q[0] : q[1]) : (k < 12u ? q[2] : q[3]);
  return (w >> (8u * (k & 3u))) & 0xffu;
}

template <int PROF, int METH>
__global__ void __launch_bounds__(B8_THREADS, B8_MIN_CTAS) k_bin_emit8(isscabac_symcfg c, const uint8_t* __restrict__ sym, uint64_t n,
                                                           const uint64_t* __restrict__ sym_off, uint32_t n_streams,
                                                           const uint32_t* __restrict__ tile_stream, const uint32_t* __restrict__ tile_first,
                                                           const uint64_t* __restrict__ tile_prefix, uint64_t* op_off, uint8_t* ops,
                                                           uint64_t cap, const uint4* __restrict__ lut, const uint8_t* __restrict__ len_tab,
                                                           uint32_t n_tiles, uint32_t tpc) {
  __shared__ uint32_t s_warp2[2][B8_THREADS / 32];     // (two copies, by tile parity: the stream pass of a tile reads them while
  __shared__ uint32_t s_pre2[2][B8_THREADS];           //  the next tile's scan is already being written -- no barrier at the tile's end)
  __shared__ __align__(16) uint8_t s_stage[BIN_STAGE];
  extern __shared__ __align__(8) uint2 s_lut8[];      // the fast table: geom.entries + 1 slots (the last one: "not in the table")
  __shared__ __align__(8) uint64_t s_soff[BIN_SOFF];
  __shared__ uint8_t s_len[256];
  const SymCfg cfg = fixed_cfg<PROF, METH>(c);
  const LutGeom geom = lut_geom(cfg.profile, cfg.method, cfg.Nq);
  const uint32_t esc = geom.entries;
  const bool need_up = cfg.profile == PROFILE_ISS || cfg.profile == PROFILE_DEMO;   // the other profiles' ops do not depend on the position
  for (uint32_t v = threadIdx.x; v < 256u; v += B8_THREADS) s_len[v] = len_tab[v];
  if (ops) {
    for (uint32_t e = threadIdx.x; e < geom.entries; e += B8_THREADS) {
      const uint4 q = __ldg(lut + e);
      uint2 f;
      lut8_from16(q.x, q.y, q.w, f.x, f.y);
      s_lut8[e] = f;
    }
    if (threadIdx.x == 0) s_lut8[esc] = make_uint2(0u, 0u);
  }
  __syncthreads();
  const uint32_t lut8 = pinned((uint32_t)__cvta_generic_to_shared(s_lut8)), stage0 = pinned((uint32_t)__cvta_generic_to_shared(s_stage));
  const uint32_t len0 = s_len[0];
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const bool vec = (reinterpret_cast<uintptr_t>(sym) & 15u) == 0;
  // the 16 symbols of this thread in tile t (zero past the end), as four words (scalars: an array would live in local memory)
  auto load_run = [&](uint32_t t, uint32_t& q0, uint32_t& q1, uint32_t& q2, uint32_t& q3) {
    const uint64_t i0 = (uint64_t)t * BIN_TILE + (uint64_t)threadIdx.x * B8_ITEMS;
    if (vec && i0 + B8_ITEMS <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(sym + i0));
      q0 = v.x; q1 = v.y; q2 = v.z; q3 = v.w;
    } else {
      q0 = q1 = q2 = q3 = 0u;
#pragma unroll 1
      for (uint32_t k = 0; k < B8_ITEMS && i0 + k < n; ++k) {
        const uint32_t b = (uint32_t)sym[i0 + k] << (8u * (k & 3u));
        if (k < 4u) q0 |= b; else if (k < 8u) q1 |= b; else if (k < 12u) q2 |= b; else q3 |= b;
      }
    }
  };
  const uint32_t tile_begin = blockIdx.x * tpc;
  const uint32_t tile_end = tile_begin + tpc < n_tiles ? tile_begin + tpc : n_tiles;
  if (tile_begin >= tile_end) return;
  // A tile's inputs are asked for one tile ahead: its symbols, its first op position, the streams it touches
  // ([s_lo, s_hi], tile_stream) and the streams that start in it ([f0, f1), tile_first).  Without this the kernel is a chain
  // of dependent load latencies per tile (first measurement: phase A alone 2.25 ms at C4, phase B 1.25).
  uint32_t nq0, nq1, nq2, nq3;
  load_run(tile_begin, nq0, nq1, nq2, nq3);
  uint64_t nbase = tile_prefix[tile_begin];
  uint32_t nf0 = tile_first[tile_begin], nf1 = tile_first[tile_begin + 1];
  uint32_t ns_lo = need_up ? tile_stream[tile_begin] : 0u, ns_hi = need_up ? tile_stream[tile_begin + 1] : 0u;
  for (uint32_t tile = tile_begin; tile < tile_end; ++tile) {
    uint32_t* s_warp = s_warp2[tile & 1u];
    uint32_t* s_pre = s_pre2[tile & 1u];
    const uint64_t t0 = (uint64_t)tile * BIN_TILE;
    const uint32_t in_tile = n - t0 < BIN_TILE ? (uint32_t)(n - t0) : (uint32_t)BIN_TILE;       // symbols of this tile (>= 1)
    const uint32_t r0 = threadIdx.x * B8_ITEMS;                                                // this thread's first, tile-relative
    const uint32_t nvalid = r0 >= in_tile ? 0u : (in_tile - r0 < B8_ITEMS ? in_tile - r0 : (uint32_t)B8_ITEMS);
    const uint64_t i0 = t0 + r0;
    const uint32_t qa[4] = {nq0, nq1, nq2, nq3};
    const uint32_t f0 = nf0, f1 = nf1, s_lo = ns_lo, s_hi = ns_hi;
    const uint64_t tile_base = nbase;
    if (tile + 1 < tile_end) {
      load_run(tile + 1, nq0, nq1, nq2, nq3);
      nbase = tile_prefix[tile + 1];
      nf0 = f1;
      nf1 = tile_first[tile + 2];
      if (need_up) { ns_lo = s_hi; ns_hi = tile_stream[tile + 2]; }
    }
    // ---- phase A: bin counts by table (the symbols past the end read as zeros: their counts are taken off again), the
    // block scan (per-thread part in s_pre, per-warp totals in s_warp: the stream pass at the end of the tile reads other
    // threads' positions out of them)
    uint32_t tot = 0;
#pragma unroll
    for (int k = 0; k < B8_ITEMS; ++k) tot += s_len[(qa[k >> 2] >> (8 * (k & 3))) & 0xffu];
    tot -= (B8_ITEMS - nvalid) * len0;
    // the offsets of the streams this tile touches, in shared memory: every thread looks its stream up in them
    const uint32_t n_soff = s_hi - s_lo + 2u;
    const bool soff_shared = need_up && n_soff <= BIN_SOFF;
    if (soff_shared)
      for (uint32_t j = threadIdx.x; j < n_soff; j += B8_THREADS) s_soff[j] = sym_off[s_lo + j];
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (uint32_t)d) inc += t;
    }
    s_pre[threadIdx.x] = inc - tot;
    if (lane == 31u) s_warp[wid] = inc;
    __syncthreads();
    uint32_t lo = inc - tot, block_total = 0;
#pragma unroll
    for (uint32_t w = 0; w < B8_THREADS / 32; ++w) {
      const uint32_t t = s_warp[w];
      if (w < wid) lo += t;
      block_total += t;
    }
    // the first stream that starts in this tile, for this thread (asked for now, used at the end of the tile)
    const uint64_t my_soff = f0 + threadIdx.x < f1 ? sym_off[f0 + threadIdx.x] : 0u;
    uint32_t up_mask = 0;
    if (need_up && nvalid) {
      // generic pointer: the tile's slice in shared memory, re-based to stream indices, or the whole array
      const uint64_t* so = soff_shared ? s_soff - s_lo : sym_off;
      uint32_t s = find_stream(so, s_lo, s_hi + 1, i0);
      uint64_t start = so[s], next = so[s + 1];
      if (i0 > start && i0 + B8_ITEMS <= next) {
        // the whole run lies inside one stream, away from its first symbol: the neighbour flags follow from the row of
        // the run's first symbol
        if (cfg.profile == PROFILE_DEMO || cfg.rows == 0) {
          up_mask = 0xffffu;
        } else {
          const uint64_t rel = i0 - start;
          uint32_t r = rel >> 32 ? (uint32_t)(rel % cfg.rows) : (uint32_t)rel % cfg.rows;
#pragma unroll
          for (int k = 0; k < B8_ITEMS; ++k) {
            if (r != 0) up_mask |= 1u << k;
            if (++r == cfg.rows) r = 0;
          }
        }
      } else {
        for (uint32_t k = 0; k < nvalid; ++k) {
          const uint64_t i = i0 + k;
          if (i == next) {   // next non-empty stream
            ++s;
            while (s + 1 < n_streams && so[s + 1] == i) ++s;
            start = i;
            next = so[s + 1];
          }
          if (sym_has_up(cfg, i - start)) up_mask |= 1u << k;
        }
      }
    }
    if (ops) {
      // ---- phase B (bin_emit.cuh): stage byte = tile position + skew, so that the stage's 16-byte pieces are the op
      // array's aligned ones
      const uint32_t skew = (uint32_t)((reinterpret_cast<uintptr_t>(ops) + tile_base) & 15u);
      const uint32_t room = tile_base >= cap ? 0u : (cap - tile_base < block_total ? (uint32_t)(cap - tile_base) : block_total);
      uint32_t before = 0;
      if (up_mask & 1u) before = i0 > 0 ? sym[i0 - 1] : 0u;
      const uint32_t pos = lo + skew, span = block_total + skew;
      BinAcc A;
      for (uint32_t w0 = 0; w0 < span; w0 += BIN_STAGE) {
        bin_acc_start(A, pos);
        const uint32_t wp_first = A.wp, fb_first = A.fb;
        StageWindow sw{stage0, w0};
        if (pos < w0 + BIN_STAGE && pos + tot + 4u > w0) {    // (span and w0 are the block's: the barriers below are met by all)
          // the common case -- the whole tile in the stage, 8 symbols in a row whose strings are all in the fast table --
          // is straight-line: 8 table loads, then 8 appends.  Otherwise symbol by symbol: a tile larger than the stage
          // leaves in rounds (every thread appends all its symbols in every round it has a byte in, the stores outside
          // the round's window are dropped), and the last tile's short runs come this way too.
#pragma unroll
          for (int h = 0; h < B8_ITEMS; h += 8) {
            bool fast = span <= BIN_STAGE && nvalid == B8_ITEMS;
            uint2 e[8];
            if (fast) {
              // (this half's two words again, opaque to the compiler and in program order behind the previous half's
              // appends: it would otherwise form all 16 table addresses ahead -- from phase A's byte extraction, across the
              // barrier -- and keep them in local memory)
              const uint32_t qh[2] = {pinned(qa[h >> 2]), pinned(qa[(h >> 2) + 1])};
              const uint32_t ub = h == 0 ? before : qa[(h >> 2) - 1] >> 24;
              uint32_t all = 0xffffffffu;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t vk = (qh[k >> 2] >> (8 * (k & 3))) & 0xffu;
                const uint32_t uk = k == 0 ? ub : ((qh[(k - 1) >> 2] >> (8 * ((k - 1) & 3))) & 0xffu);
                uint32_t idx;
                e[k] = bin8_entry(cfg, geom.dom, esc, lut8, vk, uk, (up_mask >> (h + k)) & 1u, idx);
                all = min(all, e[k].y);
              }
              fast = all >> 24;        // every length is at least 1
            }
            if (fast) {
              StageStore st{stage0};
#pragma unroll
              for (int k = 0; k < 8; ++k) bin_append(A, st, e[k].x, e[k].y & 0x00ffffffu, e[k].y >> 24);
            } else {
              const uint32_t kend = nvalid < (uint32_t)h + 8u ? nvalid : (uint32_t)h + 8u;
#pragma unroll 1
              for (uint32_t k = h; k < kend; ++k)
                bin8_symbol(cfg, geom.dom, geom.entries, lut8, lut, A, sw, run_sym(qa, k), k == 0u ? before : run_sym(qa, k - 1u), (up_mask >> k) & 1u);
            }
          }
        } else {
          A.a0 = 0u;                     // nothing of this thread in the window: neither words nor tail bytes
          A.fb = 0u;
        }
        __syncthreads();       // every word is in the stage: now the incomplete last words, byte by byte
        bin_tail(A, sw, wp_first, fb_first);
        __syncthreads();
        // The window leaves: stage byte b <-> op array byte ops + tile_base - skew + b, so the stage's 16-byte pieces are
        // the op array's aligned ones.  Whole pieces with one LDS.128 + STG.128 each; the (at most two) incomplete ones --
        // in front of the tile's first op, behind its last or behind the caller's capacity -- byte by byte by warp 0.
        const uint32_t w1 = w0 + BIN_STAGE;
        const uint32_t wend = w1 < span ? w1 : span;
        const uint32_t vb = w0 > skew ? w0 : skew, ve = wend < room + skew ? wend : room + skew;     // bytes of the window to write
        if (vb < ve) {
          uint8_t* gbase = ops + tile_base - skew;
          const uint32_t fb0 = (vb + 15u) & ~15u, fe0 = ve & ~15u;
          for (uint32_t bb = fb0 + threadIdx.x * 16u; bb < fe0; bb += B8_THREADS * 16u) {
            uint4 qq;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(qq.x), "=r"(qq.y), "=r"(qq.z), "=r"(qq.w) : "r"(stage0 + bb - w0) : "memory");
            *reinterpret_cast<uint4*>(gbase + bb) = qq;
          }
          if (threadIdx.x < 32u) {
            const bool tail = threadIdx.x >= 16u;
            const uint32_t hb = tail ? (fe0 > fb0 ? fe0 : fb0) + (threadIdx.x - 16u) : vb + threadIdx.x;
            const uint32_t he = tail ? (fe0 >= fb0 ? ve : 0u) : (fb0 < ve ? fb0 : ve);
            if (hb < he) {
              uint32_t byte;
              asm volatile("ld.shared.u8 %0, [%1];" : "=r"(byte) : "r"(stage0 + hb - w0) : "memory");
              gbase[hb] = (uint8_t)byte;
            }
          }
        }
        if (w1 < span) __syncthreads();          // the next round overwrites the stage
      }
    }
    // ---- the stream pass: op_off of the streams that start in this tile, one stream per thread.  A stream that starts
    // at tile position p begins with symbol p % 16 of thread p / 16: that thread's first op position is in s_pre / s_warp,
    // the counts of the symbols in front are looked up again (at most 15 bytes, in L1 since the tile was loaded).
    {
      uint64_t so_e = my_soff;
      for (uint32_t e = f0 + threadIdx.x; e < f1; e += B8_THREADS) {
        if (e != f0 + threadIdx.x) so_e = sym_off[e];
        const uint32_t p = (uint32_t)(so_e - t0), owner = p / B8_ITEMS;
        uint32_t o = s_pre[owner];
        for (uint32_t w = 0; w < (owner >> 5); ++w) o += s_warp[w];
        for (uint32_t j = 0; j < p % B8_ITEMS; ++j) o += s_len[sym[t0 + (p - p % B8_ITEMS) + j]];
        op_off[e] = tile_base + o;
      }
      if (tile + 1 == n_tiles)      // the streams at the very end are empty; op_off[n_streams] = the total
        for (uint32_t e = f1 + threadIdx.x; e <= n_streams; e += B8_THREADS) op_off[e] = tile_base + block_total;
    }
    // no barrier here: the next tile writes the other copy of s_pre / s_warp, s_soff was last read in front of this tile's
    // word barrier, and the stage is not written before the next tile's scan barrier, which every thread reaches only after
    // its pieces and its streams of this tile
  }
}

// ---- fused symbol encode / decode ----------------------------------------------
struct SymParams {
  isscabac_symcfg cfg;
  uint32_t n_streams, n_ctx;
  int per_stream_init, sym_width;
  const uint64_t* sym_off;
  const void* symbols;
  const uint8_t* ctx_init;
  uint8_t* slab;
  uint64_t slab_stride;
  uint32_t* lengths;
  uint32_t* bits_after;
  uint32_t* overflow;
  const uint64_t* byte_off;
  const uint8_t* bytes;
  void* out_symbols;
  uint8_t* finish_ok;
  const uint4* lut;        // op strings by table (k_bin_lut) for the fused encoder, or NULL
  uint32_t lut_entries, lut_dom;
  uint32_t pair_dom;            // ring encoder: op strings of symbol PAIRS for values below it (0: no pair table)
};

__device__ __forceinline__ uint32_t* setup_smem(const SymParams& P, uint8_t* smem, uint32_t s, bool valid) {
  uint2* tab = reinterpret_cast<uint2*>(smem);
  for (int i = threadIdx.x; i < TAB_WORDS; i += blockDim.x) tab[i] = c_rows_sym.r[i >> 5];
  uint32_t* ctx = reinterpret_cast<uint32_t*>(smem + TAB_WORDS * sizeof(uint2));
  const uint8_t* init = P.ctx_init + (P.per_stream_init && valid ? (uint64_t)s * P.n_ctx : 0);
  for (uint32_t c = 0; c < P.n_ctx; ++c) ctx[c * NT + threadIdx.x] = init[c];
  __syncthreads();
  return ctx + threadIdx.x;
}

template <bool TRACK>
__global__ void __launch_bounds__(NT) k_encode_symbols(SymParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  const bool valid = s < P.n_streams;
  uint32_t* myctx = setup_smem(P, smem, valid ? s : 0, valid);
  if (!valid) return;
  const uint2* mytab = reinterpret_cast<const uint2*>(smem) + (threadIdx.x & 31);
  const SymCfg cfg = to_cfg(P.cfg);
  const uint32_t ctx_max = P.n_ctx ? P.n_ctx - 1 : 0;
  const uint64_t s0 = P.sym_off[s], s1 = P.sym_off[s + 1];
  EncLane L;
  uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  enc_start(L, P.slab + (uint64_t)s * P.slab_stride, cap);
  SymCode prev = {0, 0, 0};
  for (uint64_t i = s0; i < s1; ++i) {
    const SymCode code = sym_code(load_sym(P.symbols, P.sym_width, i), cfg.Nq, cfg.method);
    const bool up = sym_has_up(cfg, i - s0);
    for (uint32_t b = 1; b <= code.len; ++b) {
      const int cx = select_ctx(cfg, b, code.np, prev, up);
      const uint32_t bin = sym_bin(code, b);
      if (cx < 0) {
        enc_bin_ep<TRACK>(L, bin);
      } else {
        const uint32_t c = min((uint32_t)cx, ctx_max);
        uint32_t st = myctx[c * NT];
        enc_bin_ctx<TRACK>(L, bin, st, mytab[st * 32]);
        myctx[c * NT] = st;
      }
    }
    if (TRACK) P.bits_after[i] = enc_bits_written(L);
    prev = code;
  }
  enc_finish<TRACK>(L);
  enc_flush_pending(L);
  P.lengths[s] = L.nbytes;
  if ((L.overflow || L.nbytes > cap) && P.overflow) atomicOr(P.overflow, 1u);
}

__global__ void __launch_bounds__(NT) k_decode_symbols(SymParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t s = blockIdx.x * NT + threadIdx.x;
  const bool valid = s < P.n_streams;
  uint32_t* myctx = setup_smem(P, smem, valid ? s : 0, valid);
  if (!valid) return;
  const uint2* mytab = reinterpret_cast<const uint2*>(smem) + (threadIdx.x & 31);
  const SymCfg cfg = to_cfg(P.cfg);
  const uint32_t ctx_max = P.n_ctx ? P.n_ctx - 1 : 0;
  const uint64_t s0 = P.sym_off[s], s1 = P.sym_off[s + 1];
  const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];
  DecLane D;
  dec_start(D, P.bytes + b0, (uint32_t)(b1 - b0));
  SymCode prev = {0, 0, 0};
  for (uint64_t i = s0; i < s1; ++i) {
    const bool up = sym_has_up(cfg, i - s0);
    SymDec sd;
    symdec_reset(sd);
    uint32_t v = 0;
    for (;;) {
      const int cx = select_ctx(cfg, sd.n + 1, sd.np, prev, up);
      uint32_t bin;
      if (cx < 0) {
        bin = dec_bin_ep(D);
      } else {
        const uint32_t c = min((uint32_t)cx, ctx_max);
        uint32_t st = myctx[c * NT];
        bin = dec_bin_ctx(D, st, mytab[st * 32]);
        myctx[c * NT] = st;
      }
      if (symdec_push(sd, bin, cfg, v)) break;
    }
    store_sym(P.out_symbols, P.sym_width, i, v);
    prev = sym_code(v, cfg.Nq, cfg.method);
  }
  if (P.finish_ok) P.finish_ok[s] = (uint8_t)dec_finish(D);
}

// ---- fused symbol encode / decode, wide-window formulation ------------------------
// Persistent warps (wide_common.cuh: per-lane replicated table, per-warp context block).  Every
// LANE runs a small state machine over the symbols of its current stream -- binarizer, context
// selection and (decode) finish detector + debinarizer in closed form -- that yields or consumes
// ONE bin per step, so the 32 lanes stay in lockstep bin by bin whatever their symbols' lengths
// are; the coder step is the same branch-free encw_op / decw_op as in the op-array kernels, with
// one word emission / refill after every 4th step.  No op array touches HBM.
// Load balance for ragged streams (lognormal lengths): a lane that finishes
// its stream takes the next unclaimed stream id from a global counter and carries on, so a warp
// never idles behind its longest stream; only the last few streams of the job run alone.
// Longest-first processing order for ragged streams: streams are bucketed by the magnitude of
// their symbol count (bucket = bit length), buckets laid out in descending order.  Within a
// bucket the order is arrival order -- results do not depend on it, every stream is independent.
constexpr int ORDER_BUCKETS = 65;
__global__ void k_order_hist(const uint64_t* off, uint32_t n, uint32_t* hist) {
  __shared__ uint32_t h[ORDER_BUCKETS];
  for (int i = threadIdx.x; i < ORDER_BUCKETS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) atomicAdd(&h[64 - __clzll((long long)(off[s + 1] - off[s]))], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < ORDER_BUCKETS; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}
__global__ void k_order_scatter(const uint64_t* off, uint32_t n, const uint32_t* hist, uint32_t* cursor, uint32_t* order) {
  __shared__ uint32_t base[ORDER_BUCKETS];
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    for (int b = ORDER_BUCKETS - 1; b >= 0; --b) { base[b] = acc; acc += hist[b]; }
  }
  __syncthreads();
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = s < n;
  const int b = live ? 64 - __clzll((long long)(off[s + 1] - off[s])) : -1;
  // one atomic per (warp, bucket) instead of one per stream: with equal-length streams every stream lands in the same
  // bucket and a million atomics on one counter took 0.68 ms
  const uint32_t peers = __match_any_sync(0xffffffffu, b);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t leader = __ffs(peers) - 1;
  uint32_t first = 0;
  if (live && lane == leader) first = atomicAdd(&cursor[b], (uint32_t)__popc(peers));
  first = __shfl_sync(0xffffffffu, first, leader);
  if (live) order[base[b] + first + __popc(peers & ((1u << lane) - 1u))] = s;
}

struct LaneCtx {
  WCtx ctx;
  WTab tab;
  uint32_t n_ctx;
  const uint8_t* init;
  int per_stream;
  __device__ __forceinline__ void reset(uint32_t s) const {
    const uint8_t* p = init + (per_stream ? (uint64_t)s * n_ctx : 0);
    for (uint32_t c = 0; c < n_ctx; ++c) ctx.store(c, tab.token(p[c] & 127u));
  }
};

// "has an up / previous neighbour" without a division: r = row of the symbol inside its column
__device__ __forceinline__ bool has_up_row(const SymCfg& cfg, uint32_t i_rel, uint32_t r) {
  if (cfg.profile == PROFILE_ISS) return cfg.rows ? r != 0 : i_rel > 0;
  if (cfg.profile == PROFILE_DEMO) return i_rel > 0;
  return false;
}

template <int PROF, int METH>
__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32) k_encode_symbols_wide(SymParams P, uint32_t* next_stream, const uint32_t* order) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s, n_ctx;
  WCtx ctx;
  WTab tab;
  wide_setup(P.n_streams, P.n_ctx, P.ctx_init, 0, smem, s, ctx, tab, n_ctx);   // shared init: reset() below does the rest
  const LaneCtx lc{ctx, tab, n_ctx, P.ctx_init, P.per_stream_init};
  const SymCfg cfg = fixed_cfg<PROF, METH>(P.cfg);
  const uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  // op strings by table (see k_bin_lut): behind the context blocks in shared memory; a symbol whose key is
  // in the table yields its ops with one byte load each, any other symbol takes the closed-form route
  // (compiled in only where it pays: the neighbour-conditioned profiles; the FLAT rules fold to a few instructions anyway)
  constexpr bool kLutProfile = PROF == PROFILE_ISS || PROF == PROFILE_DEMO || PROF < 0;
  const bool lut_on = kLutProfile && (cfg.profile == PROFILE_ISS || cfg.profile == PROFILE_DEMO) && P.lut != nullptr;
  uint32_t lut0 = 0;
  if (lut_on) {
    uint8_t* lp = smem + WIDE_TAB_BYTES + (size_t)(blockDim.x >> 5) * (P.n_ctx + 1) * WIDE_CTX_STRIDE;
    for (uint32_t e = threadIdx.x; e < P.lut_entries; e += blockDim.x) reinterpret_cast<uint4*>(lp)[e] = __ldg(P.lut + e);
    lut0 = (uint32_t)__cvta_generic_to_shared(lp);
    __syncthreads();
  }
  EncWide E;
  encw_start(E, nullptr, 0);
  const uint8_t* src = nullptr;     // first symbol of the stream
  uint32_t i = 0, cnt = 0, row = 0; // next symbol to fetch (stream-relative), symbols in the stream, row in the column
  SymCode cur = {0, 0, 0}, prev = {0, 0, 0};
  uint32_t b = 1, len = 0;          // next bin of the current symbol and its bin count; b > len: fetch the next symbol first
  uint32_t abase = 0;               // table route: shared-window address of the current symbol's op string (0 = closed form)
  uint32_t curv = 0, prevv = 0;     // values of the current symbol and of the one before it
  uint32_t nextv = 0;               // value of symbol i, loaded one symbol ahead
  bool up = false, active = false, have = s < P.n_streams;
  if (have && order) s = order[s];
  for (;;) {
    if (!active && have) {          // claim the stream: (re)initialise contexts and coder
      lc.reset(s);
      const uint64_t s0 = P.sym_off[s];
      cnt = (uint32_t)(P.sym_off[s + 1] - s0);
      src = static_cast<const uint8_t*>(P.symbols) + s0 * (uint64_t)P.sym_width;
      encw_start(E, P.slab + (uint64_t)s * P.slab_stride, cap);
      i = 0; row = 0; cur = SymCode{0, 0, 0}; prev = cur; b = 1; len = 0; abase = 0; curv = 0; prevv = 0; up = false;
      nextv = cnt ? load_sym(src, P.sym_width, 0) : 0u;
      active = true;
    }
    if (!__any_sync(0xffffffffu, active)) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (active && (b <= len || i < cnt)) {
        if (j == 2 && E.n >= kLazy) encw_emit(E);   // window budget of a 4-bin group, see kLazy
        if (b > len) {
          prevv = curv;
          curv = nextv;
          // the symbol after this one is requested now and used a symbol (several bins) later: a lane
          // that runs alone at the end of a ragged job must not wait for HBM once per symbol
          if (i + 1 < cnt) nextv = load_sym(src, P.sym_width, i + 1);
          up = has_up_row(cfg, i, row);
          ++i;
          if (++row == cfg.rows) row = 0;
          b = 1;
          const bool cur_is_code = abase == 0;   // `cur` still describes the symbol before this one
          abase = 0;
          if (lut_on) {
            const uint32_t key = lut_index(cfg, P.lut_dom, curv, prevv, up);
            if (key != LUT_ESC) {
              const uint32_t e = lut0 + (key << 4);
              asm volatile("ld.shared.u8 %0, [%1];" : "=r"(len) : "r"(e + 15u));
              if (len) abase = e;   // 0 = the string has more than 15 ops
            }
          }
          if (!abase) {
            prev = cur_is_code ? cur : sym_code(prevv, cfg.Nq, cfg.method);
            cur = sym_code(curv, cfg.Nq, cfg.method);
            len = cur.len;
          }
        }
        uint32_t op;
        if (abase) {
          asm volatile("ld.shared.u8 %0, [%1];" : "=r"(op) : "r"(abase + b - 1u));
        } else {
          const int cx = select_ctx(cfg, b, cur.np, prev, up);
          op = ((cx < 0 ? ISSCABAC_OP8_EP : (uint32_t)cx) << 1) | sym_bin(cur, b);
        }
        encw_op<0>(E, op >> 1, op, ctx, tab, n_ctx);
        ++b;
        if (cfg.profile == PROFILE_FLAT_EPSUF && b > cur.np && b <= cur.len) {
          // the prefix is out and everything behind it is bypass-coded: the whole suffix in one step per 16 bins
          // (encodeBinsEP); a word leaves before and after, so the run never meets a full window and the
          // step ends with fewer than 32 pending bits like any other
          uint32_t left = cur.len + 1u - b;
          do {
            const uint32_t c = left < kEpRunMax ? left : kEpRunMax;
            encw_emit(E);
            encw_ep_run(E, (cur.suf >> (left - c)) & ((1u << c) - 1u), c);
            left -= c;
          } while (left);
          encw_emit(E);
          b = cur.len + 1u;
        }
      }
    }
    if (__any_sync(0xffffffffu, E.n >= kLazy)) encw_emit(E);   // lazy, voted: all 32 lanes are here
    if (active && !(b <= len || i < cnt)) {   // stream complete: finish(), then take the next unclaimed one
      const uint32_t len = encw_finish(E);
      P.lengths[s] = len;
      if (len > cap && P.overflow) atomicOr(P.overflow, 1u);
      encw_start(E, nullptr, 0);
      active = false;
      s = gridDim.x * blockDim.x + atomicAdd(next_stream, 1u);
      have = s < P.n_streams;
      if (have && order) s = order[s];
    }
  }
}

// ---------------------------------------------------------------------------
// Fused encoder, second formulation: the binarizer runs AHEAD of the coder through a per-lane ring of op bytes in
// shared memory, so the coder's step is the 16-op block of the op-array kernels (encw_block16, ~35 instructions per
// warp-bin) instead of a per-bin state machine whose divergent symbol fetch runs in nearly every step for some lane
// (112 instructions per warp-step, profiles/r1_v10_fused_c4.txt).
//   refill (warp-uniform loop, lanes vote): every lane short of 16 buffered ops expands ONE symbol per trip -- its op
//     string comes out of the per-call table of k_bin_lut as 7 op bytes + length in 8 bytes (one LDS.64), is shifted to
//     the ring's byte position in registers and stored with up to three aligned word stores; strings longer than 7 ops
//     and symbols outside the table are produced 7 ops per trip by the closed form (sym_code / select_ctx / sym_bin);
//   code: a lane with 16 buffered ops loads them with one LDS.128 and runs encw_block16 (voted emission when all 32
//     lanes have a full block); a lane whose stream has run out of symbols codes its last ops one by one, finishes the
//     stream and claims the next one from the work queue.
// Used for u8 symbols when the profile has a table (every profile but FL32 alphabets); results are byte-identical to
// k_encode_symbols_wide (tests/test_gpu_symbols.py runs both against the oracle).
// ---------------------------------------------------------------------------
constexpr uint32_t RING_BYTES = 64;                 // per lane: four 16-op blocks
constexpr uint32_t RING_LANE_STRIDE = 80;           // + 16 B pad: the 8 lanes of a quarter-warp start in 8 different bank quads
constexpr uint32_t RING_WARP_BYTES = 32 * RING_LANE_STRIDE;

template <int PROF, int METH>
__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32) k_encode_symbols_ring(SymParams P, uint32_t* next_stream, const uint32_t* order) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s, n_ctx;
  WCtx ctx;
  WTab tab;
  wide_setup(P.n_streams, P.n_ctx, P.ctx_init, 0, smem, s, ctx, tab, n_ctx);
  const LaneCtx lc{ctx, tab, n_ctx, P.ctx_init, P.per_stream_init};
  const SymCfg cfg = fixed_cfg<PROF, METH>(P.cfg);
  const uint32_t cap = (uint32_t)(P.slab_stride > 0xfffffffcull ? 0xfffffffcull : P.slab_stride);
  const uint32_t nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // shared memory behind the context blocks: [entries] 8-byte op strings, then the rings
  uint8_t* lp = smem + WIDE_TAB_BYTES + (size_t)nwarps * (P.n_ctx + 1) * WIDE_CTX_STRIDE;
  uint2* lut8 = reinterpret_cast<uint2*>(lp);
  for (uint32_t e = threadIdx.x; e < P.lut_entries; e += blockDim.x) {
    const uint4 q = __ldg(P.lut + e);
    const uint32_t len = q.w >> 24;                 // 0: more than 15 ops
    lut8[e] = (len >= 1 && len <= 7) ? make_uint2(q.x, (q.y & 0x00ffffffu) | (len << 24)) : make_uint2(0u, 0u);
  }
  __syncthreads();
  // Where a symbol's string depends on nothing but its value (FLAT profiles) and the alphabet is small, the strings of
  // symbol PAIRS sit behind the single ones: 14 op bytes + length in 16 bytes, so that a refill trip expands two symbols
  // with one LDS.128 and one append (the trip, not the coding, was the larger half of the kernel: 66 instructions per
  // symbol against 29 per bin, profiles/r2_ring_encoder_pairs.txt).  Pairs with a string of more than 7 ops stay single.
  constexpr bool PAIRS = PROF == PROFILE_FLAT || PROF == PROFILE_FLAT_EPSUF;
  const uint32_t pd = PAIRS ? P.pair_dom : 0u;
  uint4* pairs = reinterpret_cast<uint4*>(lp + ((P.lut_entries * 8u + 15u) & ~15u));
  if (PAIRS) {
    for (uint32_t e = threadIdx.x; e < pd * pd; e += blockDim.x) {
      const uint2 a = lut8[e % pd], b = lut8[e / pd];
      const uint32_t la = a.y >> 24, lb = b.y >> 24;
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (la && lb) {
        const uint64_t A = (uint64_t)a.x | ((uint64_t)(a.y & 0x00ffffffu) << 32), B = (uint64_t)b.x | ((uint64_t)(b.y & 0x00ffffffu) << 32);
        const uint64_t lo = A | (B << (8u * la)), hi = B >> (64u - 8u * la);
        q = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, ((uint32_t)(hi >> 32) & 0xffffu) | ((la + lb) << 24));
      }
      pairs[e] = q;
    }
    __syncthreads();
  }
  const uint32_t lut0 = (uint32_t)__cvta_generic_to_shared(lp);
  const uint32_t pair0 = lut0 + ((P.lut_entries * 8u + 15u) & ~15u);
  const uint32_t ring0 = pair0 + pd * pd * 16u + warp * RING_WARP_BYTES + lane * RING_LANE_STRIDE;

  EncWide E;
  encw_start(E, nullptr, 0);
  const uint8_t* src = nullptr;
  uint32_t cnt = 0, si = 0, row = 0;      // symbols in the stream, symbols expanded, row of the next symbol in its column
  uint32_t curv = 0, nextv = 0;           // value of the symbol expanded last / of symbol si (loaded one symbol ahead)
  uint32_t next2 = 0, next3 = 0, next4 = 0;   // with a pair table: the values of symbols si + 1 .. si + 3, loaded two trips ahead
                                              // (one trip ahead the pair's table lookup waited for its symbols: 10 % of all stall samples)
  uint32_t wr = 0, buffered = 0, pw = 0;  // ring: bytes appended (mod 64 = write position), ops not yet coded, the partial word at wr
  uint32_t rd = 0;                        // ring read position (multiple of 16 until the tail)
  // a symbol whose string does not come out of the table in one piece: produced 7 ops per trip by the closed form
  SymCode cur = {0, 0, 0}, prevc = {0, 0, 0};
  uint32_t mb = 1, mlen = 0;              // next bin (1-based) and length of that symbol's string; mb > mlen: none in progress
  bool up = false, active = false, have = s < P.n_streams;
  if (have && order) s = order[s];

  auto sts_if = [&](uint32_t addr, uint32_t v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" :: "r"(addr), "r"(v), "r"((uint32_t)p) : "memory");
  };
  // a symbol byte loaded IN PLACE when p holds (the queue entries are read-write asm operands: no register move can follow the
  // load and wait for it, cabac_wide.cuh decw_refill_p)
  auto ldsym_if = [](uint32_t& dst, const uint8_t* a, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u8 %0, [%1];\n\t}" : "+r"(dst) : "l"(a), "r"((uint32_t)p));
  };
  // append `len` (0..7) op bytes held in (lo, hi) -- little-endian, zero above len -- to the ring; no branch, length 0
  // changes nothing
  auto append = [&](uint32_t lo, uint32_t hi, uint32_t len) {
    const uint32_t b8 = (wr & 3u) * 8u;
    const uint32_t t0 = pw | (lo << b8);
    const uint32_t t1 = cb_funnel_l(lo, hi, b8);
    const uint32_t t2 = cb_funnel_l(hi, 0u, b8);
    const uint32_t k0 = wr & 60u, end = (wr & 3u) + len;           // bytes of the touched words that are valid afterwards
    sts_if(ring0 + k0, t0, len != 0u);
    sts_if(ring0 + ((k0 + 4u) & 60u), t1, end > 4u);
    sts_if(ring0 + ((k0 + 8u) & 60u), t2, end > 8u);
    uint32_t tl = t0;
    tl = end >= 4u ? t1 : tl;
    tl = end >= 8u ? t2 : tl;
    pw = (end & 3u) == 0u ? 0u : tl;
    wr += len;
    buffered += len;
  };
  // the same for the `len` (2..14) op bytes of a pair held in q.x, q.y, q.z and the low half of q.w; no branch: the lanes of
  // a warp end on different words, so every arm would run anyway
  auto append_pair = [&](const uint4& q, uint32_t len) {
    const uint32_t b8 = (wr & 3u) * 8u, d3 = q.w & 0xffffu;
    const uint32_t t0 = pw | (q.x << b8);
    const uint32_t t1 = cb_funnel_l(q.x, q.y, b8);
    const uint32_t t2 = cb_funnel_l(q.y, q.z, b8);
    const uint32_t t3 = cb_funnel_l(q.z, d3, b8);
    const uint32_t t4 = cb_funnel_l(d3, 0u, b8);
    const uint32_t k0 = wr & 60u, end = (wr & 3u) + len;           // <= 17
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(ring0 + k0), "r"(t0) : "memory");
    sts_if(ring0 + ((k0 + 4u) & 60u), t1, end > 4u);
    sts_if(ring0 + ((k0 + 8u) & 60u), t2, end > 8u);
    sts_if(ring0 + ((k0 + 12u) & 60u), t3, end > 12u);
    sts_if(ring0 + ((k0 + 16u) & 60u), t4, end > 16u);
    uint32_t tl = t0;
    tl = end >= 4u ? t1 : tl;
    tl = end >= 8u ? t2 : tl;
    tl = end >= 12u ? t3 : tl;
    tl = end >= 16u ? t4 : tl;
    pw = (end & 3u) == 0u ? 0u : tl;
    wr += len;
    buffered += len;
  };

  for (;;) {
    if (!active && have) {               // claim the stream
      lc.reset(s);
      const uint64_t s0 = P.sym_off[s];
      cnt = (uint32_t)(P.sym_off[s + 1] - s0);
      src = static_cast<const uint8_t*>(P.symbols) + s0;
      encw_start(E, P.slab + (uint64_t)s * P.slab_stride, cap);
      si = 0; row = 0; curv = 0; wr = 0; rd = 0; buffered = 0; pw = 0; mb = 1; mlen = 0; up = false;
      cur = SymCode{0, 0, 0}; prevc = cur;
      nextv = cnt ? src[0] : 0u;
      next2 = cnt > 1u ? src[1] : 0u;
      if (PAIRS) {
        next3 = cnt > 2u ? src[2] : 0u;
        next4 = cnt > 3u ? src[3] : 0u;
      }
      active = true;
    }
    if (!__any_sync(0xffffffffu, active)) break;
    // ---- refill: one symbol (or one piece of a long symbol) per lane and trip
    // A trip is forced by the lanes that are short of a block (must); every lane with room works ahead in it (need), so the
    // number of trips per block follows the average symbols per block, not the maximum over the 32 lanes.
    for (;;) {
      const bool more = active && (mb <= mlen || si < cnt);
      const bool must = more && buffered < 16u;
      const bool need = more && buffered < 32u;
      if (!__any_sync(0xffffffffu, must)) break;
      bool single = need;
      if (PAIRS) {
        if (need && mb > mlen && si + 2u <= cnt && (nextv | next2) < 256u && nextv < pd && next2 < pd) {   // two symbols at once
          uint4 q;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(pair0 + (nextv + next2 * pd) * 16u));
          const uint32_t len = q.w >> 24;
          if (len) {
            append_pair(q, len);
            curv = next2;
            si += 2u;
            nextv = next3;
            next2 = next4;
            if (si + 2u < cnt) next3 = src[si + 2u];
            if (si + 3u < cnt) next4 = src[si + 3u];
            single = false;
          }
        }
      }
      // ---- one symbol, the common case without a branch: the queue moves up, the string comes out of the table, a string
      // of length 0 (lane not in this trip, symbol not in the table in one piece) appends nothing.  With a pair table the
      // block is skipped when every lane of the trip took a pair (nearly always).
      if (!PAIRS || __any_sync(0xffffffffu, single)) {
        const bool fetch = single && mb > mlen;
        const uint32_t prevv = curv;
        curv = fetch ? nextv : curv;
        nextv = fetch ? next2 : nextv;
        if (PAIRS) {
          next2 = fetch ? next3 : next2;
          next3 = fetch ? next4 : next3;
          ldsym_if(next4, src + si + 4u, fetch && si + 4u < cnt);
        } else {
          ldsym_if(next2, src + si + 2u, fetch && si + 2u < cnt);
        }
        const bool upn = has_up_row(cfg, si, row);
        up = fetch ? upn : up;
        const uint32_t key = lut_index(cfg, P.lut_dom, curv, prevv, upn);
        uint2 e = make_uint2(0u, 0u);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0, %1}, [%2];\n\t}"
                     : "+r"(e.x), "+r"(e.y) : "r"(lut0 + (key & 0x3ffu) * 8u), "r"((uint32_t)(fetch && key != LUT_ESC)));
        const uint32_t len = e.y >> 24;
        append(e.x, e.y & 0x00ffffffu, len);
        si += fetch ? 1u : 0u;
        if (cfg.rows) row = fetch ? (row + 1u == cfg.rows ? 0u : row + 1u) : row;
        if (__builtin_expect(fetch && len == 0u, 0)) {     // not in the table in one piece: closed form, 7 ops per trip
          prevc = sym_code(prevv, cfg.Nq, cfg.method);
          cur = sym_code(curv, cfg.Nq, cfg.method);
          mb = 1;
          mlen = cur.len;
        }
        if (__builtin_expect(single && mb <= mlen, 0)) {
          uint32_t lo = 0, hi = 0, k = 0;
          for (; k < 7u && mb <= mlen; ++k, ++mb) {
            const int cx = select_ctx(cfg, mb, cur.np, prevc, up);
            const uint32_t op = ((cx < 0 ? ISSCABAC_OP8_EP : (uint32_t)cx) << 1) | sym_bin(cur, mb);
            if (k < 4u) lo |= op << (8u * k); else hi |= op << (8u * (k - 4u));
          }
          append(lo, hi, k);
        }
      }
    }
    // ---- code: a full block of 16 ops, or the tail of the stream
    const bool full = active && buffered >= 16u;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (full) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(ring0 + (rd & 48u)) : "memory");
    const uint32_t cw[4] = {op_codes4(w[0]), op_codes4(w[1]), op_codes4(w[2]), op_codes4(w[3])};
    if (__all_sync(0xffffffffu, full)) {
      encw_block16<true>(E, w, cw, ctx, tab, n_ctx);
    } else if (full) {
      encw_block16<false>(E, w, cw, ctx, tab, n_ctx);
    }
    if (full) { rd += 16u; buffered -= 16u; }
    if (active && !full) {               // fewer than 16 ops left and no symbol to expand: the stream's tail
      for (uint32_t k = 0; k < buffered; ++k) {
        uint32_t op;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(op) : "r"(ring0 + ((rd + k) & 63u)) : "memory");
        encw_general(E, op, ctx, tab, n_ctx);
      }
      const uint32_t len = encw_finish(E);
      P.lengths[s] = len;
      if (len > cap && P.overflow) atomicOr(P.overflow, 1u);
      encw_start(E, nullptr, 0);
      active = false;
      buffered = 0;
      s = gridDim.x * blockDim.x + atomicAdd(next_stream, 1u);
      have = s < P.n_streams;
      if (have && order) s = order[s];
    }
  }
}

template <int PROF, int METH>
__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32) k_decode_symbols_wide(SymParams P, uint32_t* next_stream, const uint32_t* order) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t s, n_ctx;
  WCtx ctx;
  WTab tab;
  wide_setup(P.n_streams, P.n_ctx, P.ctx_init, 0, smem, s, ctx, tab, n_ctx);
  const LaneCtx lc{ctx, tab, n_ctx, P.ctx_init, P.per_stream_init};
  const SymCfg cfg = fixed_cfg<PROF, METH>(P.cfg);
  DecWide D;
  decw_start(D, P.bytes, 0);
  uint8_t* dst = nullptr;           // first symbol of the stream in the output
  uint32_t i = 0, cnt = 0, row = 0; // symbol being decoded (stream-relative), symbols in the stream, row in the column
  SymDec sd;
  symdec_reset(sd);
  SymCode prev = {0, 0, 0};
  bool up = false, active = false, have = s < P.n_streams;
  if (have && order) s = order[s];
  for (;;) {
    if (!active && have) {
      lc.reset(s);
      const uint64_t s0 = P.sym_off[s];
      cnt = (uint32_t)(P.sym_off[s + 1] - s0);
      dst = static_cast<uint8_t*>(P.out_symbols) + s0 * (uint64_t)P.sym_width;
      const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];
      decw_start(D, P.bytes + b0, (uint32_t)(b1 - b0));
      i = 0; row = 0; symdec_reset(sd); prev = SymCode{0, 0, 0}; up = false;
      active = true;
    }
    if (!__any_sync(0xffffffffu, active)) break;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (active && i < cnt) {
        const int cx = select_ctx(cfg, sd.n + 1, sd.np, prev, up);
        const uint32_t bin = decw_op<0>(D, cx < 0 ? ISSCABAC_OP8_EP : (uint32_t)cx, ctx, tab, n_ctx);
        uint32_t v = 0;
        bool done = symdec_push(sd, bin, cfg, v);
        if (cfg.profile == PROFILE_FLAT_EPSUF && cfg.method >= BIN_EG0 && cfg.method <= BIN_EG2 && !done && sd.np != 0xffffffffu) {
          // the prefix just ended and its suffix is bypass-coded: all of it in one step per 16 bins (decodeBinsEP),
          // the window topped up before (the run reads up to 26 bits) and after (the next decisions need theirs)
          uint32_t left = sd.ns_left;
          do {
            const uint32_t c = left < kEpRunMax ? left : kEpRunMax;
            decw_refill(D);
            sd.suf = (sd.suf << c) | decw_ep_run(D, c);
            left -= c;
          } while (left);
          decw_refill(D);
          sd.n += sd.ns_left;
          sd.ns_left = 0;
          v = (uint32_t)((((uint64_t)1 << (cfg.method - BIN_EG0)) * ((((uint64_t)1) << (sd.np - 1)) - 1u)) + sd.suf);
          done = true;
        }
        if (done) {
          store_sym(dst, P.sym_width, i, v);
          // the finished symbol's bin string as the next symbol's neighbour: its length, the
          // position of its first zero (length + 1 when it has none) and its suffix bits
          prev = SymCode{sd.n, sd.np != 0xffffffffu ? sd.np : (cfg.method == BIN_TU && bin == 0 ? sd.n : sd.n + 1u), sd.suf};
          ++i;
          if (++row == cfg.rows) row = 0;
          symdec_reset(sd);
          up = has_up_row(cfg, i, row);
        }
      }
    }
    if (__any_sync(0xffffffffu, D.f >= kLazyDec)) decw_refill(D);
    if (active && i >= cnt) {
      if (P.finish_ok) P.finish_ok[s] = (uint8_t)decw_finish(D);
      decw_start(D, P.bytes, 0);
      active = false;
      s = gridDim.x * blockDim.x + atomicAdd(next_stream, 1u);
      have = s < P.n_streams;
      if (have && order) s = order[s];
    }
  }
}

// ---------------------------------------------------------------------------
// Fused decoder, second formulation: the binarization is a prefix code, so "finish detector + context selection +
// debinarizer" (cabacDecodeSymbolFinished.m:10-32, cabacContextSelection.m:24-67, cabacDebinarizer.m:28-57) is a walk down
// the CODE TREE: every internal node names the context of its bin (or bypass) and its two children, a leaf is a symbol value.
// The tree is built per call on the host from the same closed-form helpers as everything else (sym_code / select_ctx /
// sym_bin), once per neighbour variant (ISS: the up neighbour's value; DEMO: the first bin of the previous symbol), and sits
// in shared memory.  A decode step is then: node (LDS.64) -> decw_op -> child by the bin -> leaf? store the symbol and
// restart at the root of the next symbol's variant.  No per-profile code in the kernel, ~50 instead of 112 instructions
// per warp-step.  A bin string that is no codeword of the alphabet (possible only in a corrupt stream) ends in an escape
// leaf: the stream is flagged (finish_ok = 0) and decoding goes on at the root.
// ---------------------------------------------------------------------------
struct TreeNode {            // 8 bytes
  uint16_t child[2];         // by bin: node index, or kLeaf | next variant << 8 | value, or kEscape
  uint16_t code;             // op code of the node's bin: context index, ISSCABAC_OP8_EP for a bypass bin
  uint16_t pad;
};
constexpr uint16_t kLeaf = 0x8000, kEscape = 0xffff;     // bit 14 is set in kEscape only: leaves carry at most 6 variant bits
// A subtree of bypass nodes that is a COMPLETE binary tree of depth n whose leaves are base, base + 1, ... in bin order is
// n bypass bins read as one number (the suffix of an exp-Golomb code under FLAT_EPSUF): it collapses into one node with
// code = kTreeEpRun, child[0] = n, child[1] = the leaf entry of value 0 (kLeaf | next variant << 8), pad = base, and the
// decoder takes it in one step with decw_ep_bits (decodeBinsEP, Decoder.cpp:333-421 == n single decodeBinEP calls).
constexpr uint16_t kTreeEpRun = 0xfffe;
constexpr uint32_t kTreeRunMax = 13;        // bins per run node: decw_ep_recip is exact up to 13
#ifndef TREE_EP_RUNS
#define TREE_EP_RUNS 1
#endif
#ifndef TREE_VOTE_REFILL
#define TREE_VOTE_REFILL 1      // measured: without the vote C2 decode 0.27 -> 0.24 ms, C4 decode 4.83 -> 4.96 ms
#endif
constexpr uint32_t TREE_MAX_NODES = 4096;
#ifndef TREE_MIN_BLOCKS
#define TREE_MIN_BLOCKS 2
#endif
constexpr uint32_t TREE_STAGE_STRIDE = 48;     // 32-byte output stage per lane + 16 B (bank spread, keeps 16-byte alignment)

struct TreeInfo {
  uint32_t n_nodes, var_stride;   // var_stride: nodes per variant (variant v's root = node v * var_stride)
  uint32_t first0;                // lane g of this launch starts with the stream at position first0 + g of the order
  uint32_t claim_base;            // the work counter hands out positions claim_base, claim_base + 1, ...
  uint32_t* started;              // SOLO: raised by the lone CTA when it starts (k_hold waits for it)
};

// host: is the subtree below node `at` a complete tree of bypass nodes over consecutive values with one next variant?
static bool tree_ep_complete(const std::vector<TreeNode>& t, uint32_t at, uint32_t& depth, uint32_t& base, uint32_t& nv) {
  if (t[at].code != ISSCABAC_OP8_EP) return false;
  uint32_t d[2], b[2], v[2];
  for (int k = 0; k < 2; ++k) {
    const uint16_t c = t[at].child[k];
    if (c == kEscape) return false;
    if (c & kLeaf) { d[k] = 0; b[k] = c & 0xffu; v[k] = (c >> 8) & 0x3fu; }
    else if (!tree_ep_complete(t, c, d[k], b[k], v[k])) return false;
  }
  if (d[0] != d[1] || v[0] != v[1] || b[1] != b[0] + (1u << d[0])) return false;
  depth = d[0] + 1; base = b[0]; nv = v[0];
  return true;
}
static void tree_collapse_ep(std::vector<TreeNode>& t, uint32_t at, bool is_root) {
  uint32_t depth, base, nv;
  // the root stays a plain node: the kernel looks for a run only behind a decoded bin
  if (!is_root && tree_ep_complete(t, at, depth, base, nv) && depth <= kTreeRunMax) {
    t[at] = TreeNode{{(uint16_t)depth, (uint16_t)(kLeaf | (nv << 8))}, kTreeEpRun, (uint16_t)base};
    return;
  }
  for (int k = 0; k < 2; ++k)
    if (!(t[at].child[k] & kLeaf)) tree_collapse_ep(t, t[at].child[k], false);
}

// host: the code trees of a configuration; false when the configuration is not covered (large alphabets, FL32, TR)
static bool build_code_tree(const isscabac_symcfg& c, std::vector<TreeNode>& nodes, TreeInfo& info) {
  const SymCfg cfg = SymCfg{c.profile, c.method, c.Nq, c.Nlbp, c.types, c.rows};
  if (!(cfg.method == BIN_TU || (cfg.method >= BIN_EG0 && cfg.method <= BIN_EG2))) return false;
  const uint32_t nq = cfg.Nq;
  if (nq < 2 || nq > 256 || (cfg.profile == PROFILE_ISS && nq > 31)) return false;
  const uint32_t n_var = cfg.profile == PROFILE_ISS ? nq + 1 : (cfg.profile == PROFILE_DEMO ? 3u : 1u);
  std::vector<std::vector<TreeNode>> trees(n_var);
  uint32_t stride = 0;
  for (uint32_t var = 0; var < n_var; ++var) {
    bool has_up = false;
    SymCode uc = {0, 0, 0};
    if (cfg.profile == PROFILE_ISS) { has_up = var != 0; if (has_up) uc = sym_code(var - 1, nq, cfg.method); }
    else if (cfg.profile == PROFILE_DEMO) { has_up = var != 0; uc = SymCode{1u, var == 1 ? 2u : 1u, 0u}; }   // first bin 1 / 0, as k_bin_lut
    std::vector<TreeNode>& t = trees[var];
    t.push_back(TreeNode{{kEscape, kEscape}, 0, 0});
    std::vector<bool> set(1, false);
    for (uint32_t v = 0; v < nq; ++v) {
      const SymCode code = sym_code(v, nq, cfg.method);
      if (code.len == 0 || code.len > 64) return false;
      uint32_t at = 0;
      for (uint32_t b = 1; b <= code.len; ++b) {
        const int cx = select_ctx(cfg, b, code.np, uc, has_up);
        const uint16_t opc = (uint16_t)(cx < 0 ? ISSCABAC_OP8_EP : cx);
        if (set[at] && t[at].code != opc) return false;      // a prefix shared by two symbols must name one context
        t[at].code = opc;
        set[at] = true;
        const uint32_t bin = sym_bin(code, b);
        if (b == code.len) {
          // the variant of the NEXT symbol when this one is its neighbour
          const uint32_t nv = cfg.profile == PROFILE_ISS ? v + 1u : (cfg.profile == PROFILE_DEMO ? (code.np > 1u ? 1u : 2u) : 0u);
          t[at].child[bin] = (uint16_t)(kLeaf | (nv << 8) | v);
        } else {
          if (t[at].child[bin] == kEscape) {
            t[at].child[bin] = (uint16_t)t.size();
            t.push_back(TreeNode{{kEscape, kEscape}, 0, 0});
            set.push_back(false);
          } else if (t[at].child[bin] & kLeaf) {
            return false;                                     // not a prefix code
          }
          at = t[at].child[bin];
        }
      }
    }
    if (TREE_EP_RUNS) tree_collapse_ep(t, 0, true);
    if (t.size() > stride) stride = (uint32_t)t.size();
  }
  if ((uint64_t)stride * n_var > TREE_MAX_NODES) return false;
  nodes.assign((size_t)stride * n_var, TreeNode{{kEscape, kEscape}, 0, 0});
  for (uint32_t var = 0; var < n_var; ++var)
    for (size_t k = 0; k < trees[var].size(); ++k) {
      TreeNode nd = trees[var][k];
      for (int b = 0; b < 2; ++b)
        if (nd.code != kTreeEpRun && !(nd.child[b] & kLeaf)) nd.child[b] = (uint16_t)(nd.child[b] + var * stride);   // variant-local -> global node index
      nodes[(size_t)var * stride + k] = nd;
    }
  info.n_nodes = stride * n_var;
  info.var_stride = stride;
  info.first0 = 0;
  info.claim_base = 0;
  info.started = nullptr;
  return true;
}

// ---------------------------------------------------------------------------
// The tree as the kernel walks it: 16-byte nodes {entry by bin 0, entry by bin 1, byte offset of the context slot,
// bypass flag}.  An entry is what follows the bin:
//   inner child     bits 0-15 = byte offset of the child node
//   symbol complete bit 31 set; bits 16-23 = value (the base value when a run follows), bits 24-27 = n, the number of
//                   bypass bins still to read as one number (a collapsed run node; 0 = none), bit 30 = escape (no such
//                   codeword), bits 0-15 = byte offset of the root the NEXT symbol starts at (its variant's)
// so a run node is never loaded: the step that decodes the bin in front of it has n and the base value in the entry,
// and "n = 0" makes the run code a no-op for every other entry.
// ---------------------------------------------------------------------------
constexpr uint32_t kEntLeaf = 0x80000000u, kEntEscape = 0x40000000u;

static void flatten_tree(const std::vector<TreeNode>& nodes, const TreeInfo& ti, uint32_t n_ctx, std::vector<uint4>& out, bool& has_runs) {
  out.assign(nodes.size(), make_uint4(kEntLeaf | kEntEscape, kEntLeaf | kEntEscape, 0u, 0u));
  has_runs = false;
  for (size_t k = 0; k < nodes.size(); ++k) {
    const TreeNode& nd = nodes[k];
    if (nd.code == kTreeEpRun) continue;                    // reached through its parent's entry only
    uint32_t e[2];
    for (int b = 0; b < 2; ++b) {
      const uint16_t c = nd.child[b];
      if (c == kEscape) e[b] = kEntLeaf | kEntEscape;       // restart at variant 0
      else if (c & kLeaf) e[b] = kEntLeaf | ((uint32_t)(c & 0xffu) << 16) | (((c >> 8) & 0x3fu) * ti.var_stride * 16u);
      else if (nodes[c].code == kTreeEpRun) {
        const TreeNode& r = nodes[c];
        e[b] = kEntLeaf | ((uint32_t)r.child[0] << 24) | ((uint32_t)r.pad << 16) | (((r.child[1] >> 8) & 0x3fu) * ti.var_stride * 16u);
        has_runs = true;
      } else e[b] = (uint32_t)c * 16u;
    }
    const bool ep = nd.code >= n_ctx;                       // ISSCABAC_OP8_EP, or a context this call does not have
    out[k] = make_uint4(e[0], e[1], (ep ? n_ctx : nd.code) * WIDE_CTX_STRIDE, ep ? 1u : 0u);
  }
}

// MODE: how the next symbol's variant follows from the one just decoded -- 0: one variant (FLAT profiles), 1: always the
// decoded symbol's (DEMO), 2: the decoded symbol's unless the next symbol starts a column (ISS).  RUNS: the tree holds run
// entries (collapsed bypass subtrees).
//
// A step is written WITHOUT branches on the lane's own state: every lane of the warp decodes one bin per step whether or
// not it still has symbols (a lane past its last symbol works on a window nobody reads any more: all its loads stay
// inside the stream or the tables), what distinguishes lanes -- symbol complete or not, run behind the bin or not,
// symbols left or not -- is predication.  Measured for a warp of long streams of C5's profile
// (profiles/r2_tree_decoder_latency.txt): the branchy step spent more cycles in BSSY / BSYNC / taken branches than in
// the arithmetic (560 - 680 cycles per step at 0.18 instructions per cycle).
// The stream's finish() checks are evaluated at the moment its last symbol is complete, while the window is the
// reference decoder's.
// SOLO: the same code as a second function, for the lone CTA of launch_sym_tree (a function's shared-memory attribute -- and
// with it the L1 / shared-memory split of the SMs it runs on -- belongs to the function: raising it for the lone CTA on the one
// function cost the main launch 9 % at C4)
template <int MODE, bool RUNS, bool SOLO = false>
__global__ void __launch_bounds__(WIDE_MAX_WARPS * 32, TREE_MIN_BLOCKS) k_decode_symbols_tree(SymParams P, uint32_t* next_stream, const uint32_t* order,
                                                                                               const uint4* tree, TreeInfo ti) {
  static_assert(WIDE_CTX_ROWS == 0, "the tree decoder addresses token slots");
  extern __shared__ __align__(16) uint8_t smem[];
  if (SOLO && threadIdx.x == 0 && ti.started) atomicExch(ti.started, 1u);      // this CTA has its SM: the main launch may go
  uint32_t s, n_ctx;
  WCtx ctx;
  WTab tab;
  wide_setup(P.n_streams, P.n_ctx, P.ctx_init, 0, smem, s, ctx, tab, n_ctx);
  const LaneCtx lc{ctx, tab, n_ctx, P.ctx_init, P.per_stream_init};
  uint8_t* tp = smem + WIDE_TAB_BYTES + (size_t)(blockDim.x >> 5) * (P.n_ctx + 1) * WIDE_CTX_STRIDE;
  for (uint32_t e = threadIdx.x; e < ti.n_nodes; e += blockDim.x) reinterpret_cast<uint4*>(tp)[e] = __ldg(tree + e);
  // 2^32 / range for range = 256 .. 511 (decw_ep_recip), behind the tree
  uint32_t* rcp = reinterpret_cast<uint32_t*>(tp + (size_t)ti.n_nodes * 16u);
  if (RUNS) for (uint32_t e = threadIdx.x; e < 256u; e += blockDim.x) rcp[e] = (uint32_t)(0x100000000ull / (256u + e)) + 1u;
  __syncthreads();
  const uint32_t tree0 = (uint32_t)__cvta_generic_to_shared(tp);
  const uint32_t rcp0 = tree0 + ti.n_nodes * 16u;
  const uint32_t ctxs = cb_keep32((uint32_t)__cvta_generic_to_shared(ctx.p));   // this lane's slot 0
  const uint32_t rows = P.cfg.rows;
  // Decoded symbols leave through a 32-byte stage per lane: one byte store to shared memory per symbol, one 16-byte store
  // to global memory per 16 symbols (a byte store per symbol and lane is a 32-byte sector write each: at C4 that was
  // the bound of the launch, not the instruction count).  Stage position = low address bits of the symbol's place in
  // the output, so the 16-byte pieces are the aligned pieces of the output array.
  const uint32_t stage0 = rcp0 + (RUNS ? 1024u : 0u) + ((threadIdx.x >> 5) * 32u + (threadIdx.x & 31u)) * TREE_STAGE_STRIDE;
  DecWide D;
  decw_start(D, P.bytes, 0);
  uint8_t* dst = nullptr;       // output position of symbol 0 of the stream
  uint32_t a0 = 0;              // low address bits of dst
  uint32_t fl = 0;              // symbols [0, fl) of the stream have been written to global memory
  uint32_t i = 0, cnt = 0, row = 0, seen = 0;   // seen: OR of every entry taken; bit 30 = an escape was hit
  uint32_t fin = 0;             // the stream's finish() verdict, taken when its last symbol was complete
  s += ti.first0;
  bool active = false, have = s < P.n_streams;
  if (have && order) s = order[s];
  auto ld_node = [](uint4& v, uint32_t a) { asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); };
  uint4 nd;                     // the node of the next bin
  ld_node(nd, tree0);
  // write symbols [fl, upto) out of the stage: the aligned 16-byte piece in one store, anything else byte by byte
  auto flush = [&](uint32_t upto) {
    while (fl < upto) {
      const uint32_t a = a0 + fl;
      if ((a & 15u) == 0u && fl + 16u <= upto) {
        uint4 q;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(stage0 + (a & 16u)) : "memory");
        *reinterpret_cast<uint4*>(dst + fl) = q;
        fl += 16u;
      } else {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(stage0 + (a & 31u)) : "memory");
        dst[fl] = (uint8_t)v;
        fl += 1u;
      }
    }
  };
  for (;;) {
    if (!active && have) {
      lc.reset(s);
      const uint64_t s0 = P.sym_off[s];
      cnt = (uint32_t)(P.sym_off[s + 1] - s0);
      dst = static_cast<uint8_t*>(P.out_symbols) + s0;
      a0 = (uint32_t)reinterpret_cast<uintptr_t>(dst);
      const uint64_t b0 = P.byte_off[s], b1 = P.byte_off[s + 1];
      decw_start(D, P.bytes + b0, (uint32_t)(b1 - b0));
      i = 0; fl = 0; row = 0; seen = 0;
      ld_node(nd, tree0);
      fin = 0;
      if (cnt == 0) { DecWide E = D; fin = decw_finish(E); }
      active = true;
    }
    if (!__any_sync(0xffffffffu, active)) break;
    // One step: the bin at node nd, the run behind it, the window top-up, the symbol if one is complete.  No branch in it;
    // LAST = this may be the stream's last symbol (then, and only then, its finish() verdict is taken -- a branch).
    auto step = [&](auto last_tag, int j) {
      constexpr bool LAST = decltype(last_tag)::value;
      const bool live = LAST ? active && i < cnt : active;
      uint32_t tok;
      const uint32_t slot = ctxs + nd.z;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tok) : "r"(slot) : "memory");
      const WRow rw = tab.row(tok);
      bool is_lps;
      const uint32_t bin = decw_bin<0>(D, nd.w != 0u, is_lps, rw);
      asm volatile("st.shared.u32 [%0], %1;" :: "r"(slot), "r"(is_lps ? rw.next_lps : rw.next_mps) : "memory");
      const uint32_t e = bin ? nd.y : nd.x;
      const bool leaf = (int32_t)e < 0;
      // the next node: the child, or the root of the next symbol's variant -- its neighbour is this symbol, unless it
      // starts a column (ISS, cabacEncode.m:52)
      uint32_t off = e & 0xffffu, next_row = 0;
      if (MODE == 2) {
        next_row = row + 1u == rows ? 0u : row + 1u;
        if (leaf && rows && next_row == 0u) off = 0u;
      }
      ld_node(nd, tree0 + off);
      // the run behind the bin (n = 0: none), then the window top-up.  With runs: every step, unconditionally -- a lane
      // enters a step with f <= 31, its bin takes <= 6 bits, its run <= 13 and finds f <= 37 <= 54 - 13.  Without: voted
      // every fourth step (cabac_wide.cuh, kLazyDec).
      uint32_t q = 0;
      if (RUNS) {
        q = decw_ep_recip(D, (e >> 24) & 0xfu, rcp0);
        decw_refill_p(D);
      } else {
#if TREE_VOTE_REFILL
        if (j == 3 && __any_sync(0xffffffffu, D.f >= kLazyDec)) decw_refill_p(D);
#else
        if (j == 3) decw_refill_p(D);      // no vote: with 32 lanes it is true in nearly every group (as in the wide op decoder)
#endif
      }
      // a complete symbol
      seen |= e;
      const bool sym = leaf && live;
      if (sym) asm volatile("st.shared.u8 [%0], %1;" :: "r"(stage0 + ((a0 + i) & 31u)), "r"((e >> 16) + q) : "memory");
      i += sym ? 1u : 0u;
      if (MODE == 2) row = sym ? next_row : row;
      if (LAST) {
        if (__builtin_expect(sym && i == cnt, 0)) {       // the last one: the window is the reference decoder's right now
          DecWide E = D;
          fin = decw_finish(E) & (((seen >> 30) & 1u) ^ 1u);
        }
      }
    };
    for (;;) {       // groups of four steps until some lane has decoded its last symbol
      // a group completes at most 4 symbols per lane: only when some lane is that close to its end do the steps look for it
      const bool careful = __any_sync(0xffffffffu, active && cnt - i <= 4u);
      if (careful) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) step(std::true_type{}, j);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) step(std::false_type{}, j);
      }
      // at most 4 symbols per group: the stage (two 16-byte pieces) never holds more than one complete aligned piece + 4
      if (active && (((a0 + i) ^ (a0 + fl)) & ~15u) != 0u) flush(i - ((a0 + i) & 15u));      // up to the last 16-byte boundary reached
      if (careful && __any_sync(0xffffffffu, active && i >= cnt)) break;
    }
    if (active && i >= cnt) {
      flush(cnt);
      if (P.finish_ok) P.finish_ok[s] = (uint8_t)fin;
      decw_start(D, P.bytes, 0);
      ld_node(nd, tree0);
      active = false;
      s = ti.claim_base + atomicAdd(next_stream, 1u);
      have = s < P.n_streams;
      if (have && order) s = order[s];
    }
  }
}

// The op-string table of a configuration (k_bin_lut), computed once per (configuration, device) and kept: like the code
// tree, a fixed cost of every small call otherwise.  own = true: the cache is full, the table belongs to this call (free it).
constexpr size_t kTreeCacheMax = 64;     // entries of the per-configuration caches (op-string tables, code trees); nothing is evicted
struct LutCacheEntry { isscabac_symcfg cfg; int dev; uint4* lut; };
static std::mutex g_lut_cache_mutex;
static std::vector<LutCacheEntry> g_lut_cache;
static int cached_lut(const isscabac_symcfg& cfg, uint32_t entries, cudaStream_t st, uint4** lut, bool* own) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  *own = false;
  std::lock_guard<std::mutex> lock(g_lut_cache_mutex);
  for (const LutCacheEntry& e : g_lut_cache)
    if (e.dev == dev && memcmp(&e.cfg, &cfg, sizeof cfg) == 0) { *lut = e.lut; return ISSCABAC_OK; }
  if (g_lut_cache.size() >= kTreeCacheMax) {
    CK(cudaMallocAsync(reinterpret_cast<void**>(lut), entries * sizeof(uint4), st));
    k_bin_lut<<<(entries + 127) / 128, 128, 0, st>>>(cfg, *lut);
    *own = true;
    return ISSCABAC_OK;
  }
  CK(cudaMalloc(reinterpret_cast<void**>(lut), entries * sizeof(uint4)));
  k_bin_lut<<<(entries + 127) / 128, 128, 0, st>>>(cfg, *lut);
  CK(cudaStreamSynchronize(st));      // once: any stream may read it from now on
  g_lut_cache.push_back(LutCacheEntry{cfg, dev, *lut});
  return ISSCABAC_OK;
}

// persistent launch: as many warps as the streams need, at most what is resident on the device
template <class K>
int launch_sym_wide(K kernel, const SymParams& P_in, cudaStream_t st, const char* name, bool& done, bool want_lut = false,
                    bool ring = false) {
  SymParams P = P_in;
  uint32_t nw, grid;
  size_t smem;
  done = false;
  if (!wide_geometry(P.n_streams, P.n_ctx, nw, grid, smem)) return ISSCABAC_OK;
  if (ring) {
    // the ring encoder: 8-byte op strings + one ring per lane behind the context blocks; fewer warps per CTA if needed
    const LutGeom rg = lut_geom(P.cfg.profile, P.cfg.method, P.cfg.Nq);
    if (!rg.entries) return ISSCABAC_OK;
    // strings of symbol pairs behind the single ones, for the profiles whose strings depend on the value alone
    const bool pairs_ok = (P.cfg.profile == ISSCABAC_PROFILE_FLAT || P.cfg.profile == ISSCABAC_PROFILE_FLAT_EPSUF) && rg.dom <= 32u && !getenv("ISSCABAC_RING_NOPAIRS");
    P.pair_dom = pairs_ok ? rg.dom : 0u;
    const size_t lut_b = (((size_t)rg.entries * 8 + 15) & ~(size_t)15) + (size_t)P.pair_dom * P.pair_dom * 16;
    const size_t per_warp = ((size_t)P.n_ctx + 1) * WIDE_CTX_STRIDE + RING_WARP_BYTES;
    const size_t lim = smem_limit();
    if (WIDE_TAB_BYTES + lut_b + per_warp > lim) return ISSCABAC_OK;
    const uint32_t nw_fit = (uint32_t)((lim - WIDE_TAB_BYTES - lut_b) / per_warp);
    if (nw > nw_fit) { nw = nw_fit; grid = ((P.n_streams + 31) / 32 + nw - 1) / nw; }
    smem = WIDE_TAB_BYTES + lut_b + per_warp * nw;
    int rc = keep_pool_cached();
    if (rc) return rc;
    uint4* lut = nullptr;
    bool lut_own = false;
    if ((rc = cached_lut(P.cfg, rg.entries, st, &lut, &lut_own))) return rc;
    P.lut = lut; P.lut_entries = rg.entries; P.lut_dom = rg.dom;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)(nw * 32), smem));
    const uint32_t resident = (uint32_t)(per_sm > 0 ? per_sm : 1) * (uint32_t)sm_count();
    if (grid > resident) grid = resident;
    uint32_t* counter = nullptr;
    if ((rc = work_counter(st, &counter))) return rc;
    uint32_t* order = nullptr;
    if (P.n_streams > grid * nw * 32) {
      CK(cudaMallocAsync(reinterpret_cast<void**>(&order), ((size_t)P.n_streams + 2 * ORDER_BUCKETS) * sizeof(uint32_t), st));
      uint32_t* hist = order + P.n_streams;
      CK(cudaMemsetAsync(hist, 0, 2 * ORDER_BUCKETS * sizeof(uint32_t), st));
      const uint32_t blocks = (P.n_streams + 255) / 256;
      k_order_hist<<<blocks, 256, 0, st>>>(P.sym_off, P.n_streams, hist);
      k_order_scatter<<<blocks, 256, 0, st>>>(P.sym_off, P.n_streams, hist, hist + ORDER_BUCKETS, order);
    }
    kernel<<<grid, nw * 32, smem, st>>>(P, counter, order);
    cudaError_t e = cudaGetLastError();
    if (order) cudaFreeAsync(order, st);
    if (lut_own) cudaFreeAsync(lut, st);
    done = true;
    return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, name);
  }
  // op strings by table for the encoder when they fit shared memory behind the context blocks
  const LutGeom geom = lut_geom(P.cfg.profile, P.cfg.method, P.cfg.Nq);
  uint4* lut = nullptr;
  if (want_lut && geom.entries && (P.cfg.profile == ISSCABAC_PROFILE_ISS || P.cfg.profile == ISSCABAC_PROFILE_DEMO) &&
      smem + geom.entries * sizeof(uint4) <= smem_limit()) {
    int rc = keep_pool_cached();
    if (rc) return rc;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&lut), geom.entries * sizeof(uint4), st));
    k_bin_lut<<<(geom.entries + 127) / 128, 128, 0, st>>>(P.cfg, lut);
    P.lut = lut; P.lut_entries = geom.entries; P.lut_dom = geom.dom;
    smem += geom.entries * sizeof(uint4);
  }
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)(nw * 32), smem));
  const uint32_t resident = (uint32_t)(per_sm > 0 ? per_sm : 1) * (uint32_t)sm_count();
  if (grid > resident) grid = resident;
  uint32_t* counter = nullptr;     // streams claimed beyond the first one of every lane (lane g starts with stream g)
  int rc = work_counter(st, &counter);
  if (rc) return rc;
  // more streams than lanes: lanes will claim further streams, so hand them out longest first
  uint32_t* order = nullptr;
  if (P.n_streams > grid * nw * 32) {
    if ((rc = keep_pool_cached())) return rc;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&order), ((size_t)P.n_streams + 2 * ORDER_BUCKETS) * sizeof(uint32_t), st));
    uint32_t* hist = order + P.n_streams;
    CK(cudaMemsetAsync(hist, 0, 2 * ORDER_BUCKETS * sizeof(uint32_t), st));
    const uint32_t blocks = (P.n_streams + 255) / 256;
    k_order_hist<<<blocks, 256, 0, st>>>(P.sym_off, P.n_streams, hist);
    k_order_scatter<<<blocks, 256, 0, st>>>(P.sym_off, P.n_streams, hist, hist + ORDER_BUCKETS, order);
  }
  kernel<<<grid, nw * 32, smem, st>>>(P, counter, order);
  cudaError_t e = cudaGetLastError();
  if (order) cudaFreeAsync(order, st);
  if (lut) cudaFreeAsync(lut, st);
  done = true;
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, name);
}

// holds the main stream back until the lone CTA has started (it raises *flag first thing), at most `ns`
constexpr unsigned kSoloHeadStartNs = 500000;
__global__ void k_hold(const uint32_t* flag, unsigned ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    if (*reinterpret_cast<const volatile uint32_t*>(flag)) return;
    __nanosleep(500);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

// a high-priority stream and a fork / join event pair per device for the lone CTA of launch_sym_tree
static std::mutex g_solo_mutex;
static int solo_stream(cudaStream_t* ss, cudaEvent_t* ev_fork, cudaEvent_t* ev_join) {
  constexpr int kDevs = 64;
  static cudaStream_t s_stream[kDevs] = {};
  static cudaEvent_t s_fork[kDevs] = {}, s_join[kDevs] = {};
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kDevs) { set_error("device index out of range"); return ISSCABAC_ERR_CUDA; }
  std::lock_guard<std::mutex> lock(g_solo_mutex);
  if (!s_stream[dev]) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&s_stream[dev], cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&s_fork[dev], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s_join[dev], cudaEventDisableTiming));
  }
  *ss = s_stream[dev]; *ev_fork = s_fork[dev]; *ev_join = s_join[dev];
  return ISSCABAC_OK;
}

// The flattened tree of a configuration, built and copied to the device once per (configuration, context count, device): a call
// on a small job is a few hundred microseconds, of which building the tree on the host, a stream-ordered allocation and a copy
// from pageable memory were a noticeable part.  A handful of configurations per process; beyond kTreeCacheMax a call uploads its own.
struct TreeCacheEntry {
  isscabac_symcfg cfg;
  uint32_t n_ctx;
  int dev;
  bool covered, has_runs;
  TreeInfo ti;
  uint4* d_tree;
};
static std::mutex g_tree_cache_mutex;
static std::vector<TreeCacheEntry> g_tree_cache;

// `spill`: filled instead of out.d_tree when the cache is full (the caller then uploads it for this call only)
static int cached_tree(const isscabac_symcfg& cfg, uint32_t n_ctx, TreeCacheEntry& out, std::vector<uint4>& spill) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tree_cache_mutex);
  for (const TreeCacheEntry& e : g_tree_cache)
    if (e.dev == dev && e.n_ctx == n_ctx && memcmp(&e.cfg, &cfg, sizeof cfg) == 0) { out = e; return ISSCABAC_OK; }
  TreeCacheEntry e;
  memset(&e, 0, sizeof e);
  e.cfg = cfg; e.n_ctx = n_ctx; e.dev = dev;
  std::vector<TreeNode> nodes;
  e.covered = build_code_tree(cfg, nodes, e.ti);
  const bool full = g_tree_cache.size() >= kTreeCacheMax;      // nothing is ever evicted: an entry may be in use by another thread's launch
  if (e.covered) {
    std::vector<uint4> flat;
    flatten_tree(nodes, e.ti, n_ctx, flat, e.has_runs);
    if (full) {
      spill.swap(flat);
    } else {
      CK(cudaMalloc(reinterpret_cast<void**>(&e.d_tree), flat.size() * sizeof(uint4)));
      CK(cudaMemcpy(e.d_tree, flat.data(), flat.size() * sizeof(uint4), cudaMemcpyHostToDevice));   // synchronous: any stream may use it next
    }
  }
  if (!full) g_tree_cache.push_back(e);
  out = e;
  return ISSCABAC_OK;
}

int launch_sym_tree(const SymParams& P, cudaStream_t st, bool& done) {
  done = false;
  TreeCacheEntry tc;
  std::vector<uint4> spill;
  int rc = cached_tree(P.cfg, P.n_ctx, tc, spill);
  if (rc) return rc;
  if (!tc.covered) return ISSCABAC_OK;
  TreeInfo ti = tc.ti;
  const bool has_runs = tc.has_runs;
  uint32_t nw, grid;
  size_t smem;
  if (!wide_geometry(P.n_streams, P.n_ctx, nw, grid, smem)) return ISSCABAC_OK;
  const size_t tree_b = (size_t)ti.n_nodes * sizeof(uint4) + (has_runs ? 1024 : 0);      // nodes + the reciprocal table
  const size_t lim = smem_limit();
  const size_t per_warp = ((size_t)P.n_ctx + 1) * WIDE_CTX_STRIDE + 32 * TREE_STAGE_STRIDE;
  if (WIDE_TAB_BYTES + tree_b + per_warp > lim) return ISSCABAC_OK;
  const uint32_t nw_fit = (uint32_t)((lim - WIDE_TAB_BYTES - tree_b) / per_warp);
  if (nw > nw_fit) { nw = nw_fit; grid = ((P.n_streams + 31) / 32 + nw - 1) / nw; }
  smem = WIDE_TAB_BYTES + tree_b + per_warp * nw;
  if ((rc = keep_pool_cached())) return rc;
  const uint4* d_tree = tc.d_tree;
  uint4* d_own = nullptr;
  if (!d_tree) {      // cache full: this call's own copy
    CK(cudaMallocAsync(reinterpret_cast<void**>(&d_own), spill.size() * sizeof(uint4), st));
    CK(cudaMemcpyAsync(d_own, spill.data(), spill.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
    d_tree = d_own;
  }
  auto kernel = P.cfg.profile == ISSCABAC_PROFILE_ISS ? (has_runs ? k_decode_symbols_tree<2, true> : k_decode_symbols_tree<2, false>)
              : P.cfg.profile == ISSCABAC_PROFILE_DEMO ? (has_runs ? k_decode_symbols_tree<1, true> : k_decode_symbols_tree<1, false>)
              : has_runs ? k_decode_symbols_tree<0, true> : k_decode_symbols_tree<0, false>;
  auto kernel_solo = P.cfg.profile == ISSCABAC_PROFILE_ISS ? (has_runs ? k_decode_symbols_tree<2, true, true> : k_decode_symbols_tree<2, false, true>)
                   : P.cfg.profile == ISSCABAC_PROFILE_DEMO ? (has_runs ? k_decode_symbols_tree<1, true, true> : k_decode_symbols_tree<1, false, true>)
                   : has_runs ? k_decode_symbols_tree<0, true, true> : k_decode_symbols_tree<0, false, true>;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)(nw * 32), smem));
  const uint32_t resident = (uint32_t)(per_sm > 0 ? per_sm : 1) * (uint32_t)sm_count();
  if (grid > resident) grid = resident;
  uint32_t* counter = nullptr;
  if ((rc = work_counter(st, &counter))) return rc;
  uint32_t* order = nullptr;
  if (P.n_streams > grid * nw * 32) {
    CK(cudaMallocAsync(reinterpret_cast<void**>(&order), ((size_t)P.n_streams + 2 * ORDER_BUCKETS) * sizeof(uint32_t), st));
    uint32_t* hist = order + P.n_streams;
    CK(cudaMemsetAsync(hist, 0, 2 * ORDER_BUCKETS * sizeof(uint32_t), st));
    const uint32_t blocks = (P.n_streams + 255) / 256;
    k_order_hist<<<blocks, 256, 0, st>>>(P.sym_off, P.n_streams, hist);
    k_order_scatter<<<blocks, 256, 0, st>>>(P.sym_off, P.n_streams, hist, hist + ORDER_BUCKETS, order);
  }
  // More streams than lanes, i.e. lanes will claim further streams, longest first: the launch ends with its longest streams'
  // serial chains, and while there is other work those chains share their SM with 31 busy warps and advance at less than half
  // their speed (C5 on one GPU: 3.5 ms of such sharing in front of a 6.9 ms chain).  So the 128 longest streams get an SM of
  // their own: a second launch of the same kernel, ONE CTA of four warps (one per scheduler) that asks for all of the SM's shared
  // memory, on a high-priority stream beside the main launch, which runs on the other SMs.  Both take further streams from the
  // same counter, so the lone CTA is not idle when its own streams are short.  C5 decode 9.06 -> 7.14 ms, C4 (equally long
  // streams: nothing to gain) 4.87 -> 4.88 ms (profiles/r2_tree_solo_experiment.txt).  ISSCABAC_TREE_SOLO=0 switches it off.
  const char* solo_env = getenv("ISSCABAC_TREE_SOLO");
  const bool solo = order && sm_count() > 8 && !(solo_env && solo_env[0] == '0');
  cudaError_t e;
  if (solo) {
    constexpr uint32_t kSoloWarps = 4, kSoloLanes = kSoloWarps * 32;
    cudaStream_t ss;
    cudaEvent_t ev_fork, ev_join;
    if ((rc = solo_stream(&ss, &ev_fork, &ev_join))) return rc;
    const uint32_t grid_b = (uint32_t)(per_sm > 0 ? per_sm : 1) * ((uint32_t)sm_count() - 1u);
    TreeInfo ta = ti, tb = ti;
    ta.first0 = 0;
    tb.first0 = kSoloLanes;
    ta.claim_base = tb.claim_base = kSoloLanes + grid_b * nw * 32;
    uint32_t* started = nullptr;
    if ((rc = work_counter(st, &started))) return rc;      // a zeroed word: the lone CTA's "I have started"
    ta.started = started;
    CK(cudaFuncSetAttribute(kernel_solo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
    {
      std::lock_guard<std::mutex> lock(g_solo_mutex);      // the two events are shared by the calls of a device
      // The lone CTA needs an EMPTY SM: it has to be placed before the main launch's CTAs spread over all of them, and the
      // main launch, straight behind the ordering kernels in its own stream, is pending a few microseconds before the side
      // stream has seen its event (then the lone CTA starts when the first SM drains and the call takes 9.7 instead of 7.1
      // ms; an event wait on the main stream did not change the order reliably).  A one-thread kernel holds the main stream
      // back until the lone CTA has raised its flag (at most kSoloHeadStartNs).
      CK(cudaEventRecord(ev_fork, st));
      CK(cudaStreamWaitEvent(ss, ev_fork, 0));
      kernel_solo<<<1, kSoloLanes, lim, ss>>>(P, counter, order, d_tree, ta);
      e = cudaGetLastError();
      CK(cudaEventRecord(ev_join, ss));
      k_hold<<<1, 1, 0, st>>>(started, kSoloHeadStartNs);
      kernel<<<grid_b, nw * 32, smem, st>>>(P, counter, order, d_tree, tb);
      if (e == cudaSuccess) e = cudaGetLastError();
      CK(cudaStreamWaitEvent(st, ev_join, 0));
    }
  } else {
    ti.first0 = 0;
    ti.claim_base = grid * nw * 32;
    kernel<<<grid, nw * 32, smem, st>>>(P, counter, order, d_tree, ti);
    e = cudaGetLastError();
  }
  if (order) cudaFreeAsync(order, st);
  if (d_own) cudaFreeAsync(d_own, st);
  done = true;
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_decode_symbols_tree");
}

int check_cfg(const isscabac_symcfg* cfg, uint32_t n_ctx, int sym_width, bool need_ctx, bool decode = false) {
  if (!cfg) { set_error("symcfg is NULL"); return ISSCABAC_ERR_INVALID; }
  if (cfg->profile < 0 || cfg->profile > ISSCABAC_PROFILE_FLAT_EPSUF) { set_error("unknown profile %d", cfg->profile); return ISSCABAC_ERR_INVALID; }
  if (cfg->method < 0 || cfg->method > ISSCABAC_BIN_TR2) { set_error("unknown binarization method %d", cfg->method); return ISSCABAC_ERR_INVALID; }
  if (decode && cfg->method >= ISSCABAC_BIN_TR0) {
    // truncated Rice is encode-only upstream: the escape code is a TODO (cabacBinarizer.m:47-50) and the decode loops
    // have no case for it (cabacDecodeSymbolFinished.m:10-32)
    set_error("truncated-Rice streams cannot be decoded (as in the reference)");
    return ISSCABAC_ERR_UNSUPPORTED;
  }
  if (cfg->Nlbp < 1 || cfg->Nlbp > 32) { set_error("Nlbp out of range"); return ISSCABAC_ERR_INVALID; }
  if (cfg->method == ISSCABAC_BIN_TU && cfg->Nq < 2) { set_error("TU needs Nq >= 2"); return ISSCABAC_ERR_INVALID; }
  if (sym_width != 1 && sym_width != 2 && sym_width != 4) { set_error("sym_width must be 1, 2 or 4"); return ISSCABAC_ERR_INVALID; }
  if (need_ctx) {
    int want = cabac_profile_num_ctx(cfg->profile, cfg->Nlbp);
    if ((int)n_ctx < want) { set_error("profile needs %d contexts, got %u", want, n_ctx); return ISSCABAC_ERR_INVALID; }
    if ((size_t)n_ctx * NT * 4 + TAB_WORDS * sizeof(uint2) > smem_limit()) { set_error("too many contexts for the symbol kernels"); return ISSCABAC_ERR_UNSUPPORTED; }
  }
  return ISSCABAC_OK;
}

template <class K>
int launch_sym(K kernel, const SymParams& P, cudaStream_t st, const char* name) {
  size_t smem = TAB_WORDS * sizeof(uint2) + (size_t)P.n_ctx * NT * 4;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<(P.n_streams + NT - 1) / NT, NT, smem, st>>>(P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, name);
}

}  // namespace

extern "C" {

size_t cabac_binarize_scratch_bytes(uint64_t n_symbols, uint32_t n_streams) {
  (void)n_streams;
  const uint64_t tiles = (n_symbols + BIN_TILE - 1) / BIN_TILE;
  const size_t sums = ((size_t)tiles * 4 + 255) & ~(size_t)255;
  const size_t pref = (((size_t)tiles + 1) * 8 + 255) & ~(size_t)255;
  const size_t tstr = (((size_t)tiles + 1) * 4 + 255) & ~(size_t)255;
  const size_t chk = ((size_t)tiles * BIN_CHUNKS * sizeof(uint16_t) + 255) & ~(size_t)255;
  return sums + pref + 2 * tstr + chk + ((cabac_compact_scratch_bytes((uint32_t)tiles) + 255) & ~(size_t)255) + LUT_MAX * sizeof(uint4) + 512 + 256;
}

int cabac_binarize_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                           const void* d_symbols, int sym_width, uint64_t n_symbols,
                           uint64_t* d_op_off, uint8_t* d_ops, uint64_t ops_cap,
                           void* d_scratch, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cfg(cfg, 0, sym_width, false);
  if (rc) return rc;
  if (!d_sym_off || !d_op_off || !d_scratch || (n_symbols && !d_symbols)) { set_error("cabac_binarize_symbols: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (n_symbols == 0 || n_streams == 0) {
    CK(cudaMemsetAsync(d_op_off, 0, ((size_t)n_streams + 1) * sizeof(uint64_t), st));
    return ISSCABAC_OK;
  }
  const uint64_t tiles64 = (n_symbols + BIN_TILE - 1) / BIN_TILE;
  if (tiles64 > 0x7fffffffull) { set_error("cabac_binarize_symbols: too many symbols for one call"); return ISSCABAC_ERR_UNSUPPORTED; }
  const uint32_t tiles = (uint32_t)tiles64;
  uint8_t* scr = static_cast<uint8_t*>(d_scratch);
  const size_t sums_b = ((size_t)tiles * 4 + 255) & ~(size_t)255;
  const size_t pref_b = (((size_t)tiles + 1) * 8 + 255) & ~(size_t)255;
  uint32_t* tile_sums = reinterpret_cast<uint32_t*>(scr);
  uint64_t* tile_prefix = reinterpret_cast<uint64_t*>(scr + sums_b);
  const size_t tstr_b = (((size_t)tiles + 1) * 4 + 255) & ~(size_t)255;
  uint32_t* tile_stream = reinterpret_cast<uint32_t*>(scr + sums_b + pref_b);
  uint32_t* tile_first = reinterpret_cast<uint32_t*>(scr + sums_b + pref_b + tstr_b);
  const size_t chk_b = ((size_t)tiles * BIN_CHUNKS * sizeof(uint16_t) + 255) & ~(size_t)255;
  uint16_t* chunks = reinterpret_cast<uint16_t*>(scr + sums_b + pref_b + 2 * tstr_b);
  void* scan_scr = scr + sums_b + pref_b + 2 * tstr_b + chk_b;
  uint4* lut = reinterpret_cast<uint4*>(static_cast<uint8_t*>(scan_scr) + ((cabac_compact_scratch_bytes(tiles) + 255) & ~(size_t)255));
  const LutGeom geom = lut_geom(cfg->profile, cfg->method, cfg->Nq);
  uint8_t* len_tab = reinterpret_cast<uint8_t*>(lut + LUT_MAX);
  // 8-bit symbols: counts by table, ops appended word-wise (k_bin_count8 / k_bin_emit8); ISSCABAC_BIN8=0 keeps the
  // closed-form kernels (both run against the oracle in the tests)
  const char* bin8_env = getenv("ISSCABAC_BIN8");
  bool bin8 = sym_width == 1 && !(bin8_env && bin8_env[0] == '0');
  if (bin8)      // the count table holds bytes (a string of 256 ops: unary codes of the value 255)
    for (uint32_t v = 0; v < 256u && bin8; ++v) bin8 = sym_code(v, cfg->Nq, cfg->method).len <= 255u;
  const bool count8 = bin8 && (reinterpret_cast<uintptr_t>(d_symbols) & 15u) == 0;
  if (bin8) k_bin_lentab<<<1, 256, 0, st>>>(*cfg, len_tab);
  if (count8 && !d_ops) {
    // the sizing call: the offsets come out of the count pass (k_bin_offsets8), the symbols are read once
    k_bin_count8<<<(tiles + BIN_THREADS / 32 - 1) / (BIN_THREADS / 32), BIN_THREADS, 0, st>>>(static_cast<const uint8_t*>(d_symbols), n_symbols, len_tab, tiles, tile_sums, chunks);
    if ((rc = exclusive_scan_u32_u64(tile_sums, tile_prefix, tiles, scan_scr, st))) return rc;
    k_bin_offsets8<<<(uint32_t)(((uint64_t)n_streams + 1 + 255) / 256), 256, 0, st>>>(static_cast<const uint8_t*>(d_symbols), n_symbols, d_sym_off, n_streams, len_tab, tiles,
                                                                                   tile_prefix, chunks, d_op_off);
    cudaError_t e8 = cudaGetLastError();
    return e8 == cudaSuccess ? ISSCABAC_OK : cuda_fail(e8, "cabac_binarize_symbols");
  }
  k_bin_tile_streams<<<(tiles + 1 + 255) / 256, 256, 0, st>>>(d_sym_off, n_streams, n_symbols, tiles, tile_stream, tile_first);
  if (d_ops && geom.entries) k_bin_lut<<<(geom.entries + 127) / 128, 128, 0, st>>>(*cfg, lut);
#define BIN_COUNT(WW, ME) k_bin_count<WW, ME><<<tiles, BIN_THREADS, 0, st>>>(*cfg, d_symbols, n_symbols, tile_sums)
  if (count8)
    k_bin_count8<<<(tiles + BIN_THREADS / 32 - 1) / (BIN_THREADS / 32), BIN_THREADS, 0, st>>>(static_cast<const uint8_t*>(d_symbols), n_symbols, len_tab, tiles, tile_sums, nullptr);
  else if (sym_width == 1 && cfg->method == ISSCABAC_BIN_EG0) BIN_COUNT(1, ISSCABAC_BIN_EG0);
  else if (sym_width == 1 && cfg->method == ISSCABAC_BIN_EG2) BIN_COUNT(1, ISSCABAC_BIN_EG2);
  else if (sym_width == 1 && cfg->method == ISSCABAC_BIN_TU) BIN_COUNT(1, ISSCABAC_BIN_TU);
  else if (sym_width == 1) BIN_COUNT(1, -1);
  else if (sym_width == 2) BIN_COUNT(2, -1);
  else BIN_COUNT(4, -1);
#undef BIN_COUNT
  if ((rc = exclusive_scan_u32_u64(tile_sums, tile_prefix, tiles, scan_scr, st))) return rc;
  // tiles per CTA of the u8 emit kernel: the tables are set up once per CTA; keep at least ~8 CTAs per SM
  uint32_t tpc = 1;
  while (tpc < 8u && tiles / (tpc * 2u) >= (uint32_t)sm_count() * 8u) tpc *= 2u;
#define BIN_EMIT8(PR, ME) \
  k_bin_emit8<PR, ME><<<(tiles + tpc - 1) / tpc, B8_THREADS, (geom.entries + 1) * sizeof(uint2), st>>>(*cfg, static_cast<const uint8_t*>(d_symbols), n_symbols, d_sym_off, n_streams, \
                                                                      tile_stream, tile_first, tile_prefix, d_op_off, d_ops, ops_cap, lut, len_tab, tiles, tpc)
#define BIN_EMIT(WW, PR, ME) \
  k_bin_emit<WW, PR, ME><<<tiles, BIN_THREADS, 0, st>>>(*cfg, d_symbols, n_symbols, d_sym_off, n_streams, tile_stream, tile_prefix, d_op_off, d_ops, ops_cap, lut)
#define BIN_EMIT_CASE(PR, ME) \
  if (sym_width == 1 && cfg->profile == PR && cfg->method == ME) { if (bin8) BIN_EMIT8(PR, ME); else BIN_EMIT(1, PR, ME); } else
  BIN_EMIT_CASE(ISSCABAC_PROFILE_ISS, ISSCABAC_BIN_EG0)
  BIN_EMIT_CASE(ISSCABAC_PROFILE_FLAT, ISSCABAC_BIN_EG0)
  BIN_EMIT_CASE(ISSCABAC_PROFILE_FLAT_EPSUF, ISSCABAC_BIN_EG2)
  BIN_EMIT_CASE(ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_TU)
  BIN_EMIT_CASE(ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_EG0)
  if (bin8) BIN_EMIT8(-1, -1);
  else if (sym_width == 1) BIN_EMIT(1, -1, -1);
  else if (sym_width == 2) BIN_EMIT(2, -1, -1);
  else BIN_EMIT(4, -1, -1);
#undef BIN_EMIT_CASE
#undef BIN_EMIT
#undef BIN_EMIT8
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "cabac_binarize_symbols");
}

int cabac_encode_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_sym_off,
                         const void* d_symbols, int sym_width,
                         const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                         uint8_t* d_slab, uint64_t slab_stride, uint32_t* d_lengths,
                         uint32_t* d_bits_after_symbol, uint32_t* d_overflow, void* stream) {
  int rc = check_cfg(cfg, n_ctx, sym_width, true);
  if (rc) return rc;
  if (n_streams == 0) return ISSCABAC_OK;
  if (!d_sym_off || !d_slab || !d_lengths || !d_ctx_init) { set_error("cabac_encode_symbols: null pointer"); return ISSCABAC_ERR_INVALID; }
  if ((slab_stride & 15u) || slab_stride < 16 || (reinterpret_cast<uintptr_t>(d_slab) & 15u)) {
    set_error("cabac_encode_symbols: slab and slab_stride must be 16-byte aligned");
    return ISSCABAC_ERR_INVALID;
  }
  SymParams P;
  memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.n_streams = n_streams; P.n_ctx = n_ctx; P.per_stream_init = per_stream_init; P.sym_width = sym_width;
  P.sym_off = d_sym_off; P.symbols = d_symbols; P.ctx_init = d_ctx_init;
  P.slab = d_slab; P.slab_stride = slab_stride; P.lengths = d_lengths; P.bits_after = d_bits_after_symbol; P.overflow = d_overflow;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_bits_after_symbol) return launch_sym(k_encode_symbols<true>, P, st, "k_encode_symbols");   // getNumBits() trace
  bool done = false;
  constexpr bool kWantLut = true;
  // the ring encoder (binarizer ahead of the coder) for u8 symbols of profiles with an op-string table;
  // ISSCABAC_SYM_RING=0 keeps the per-bin state machine (both are tested against the oracle)
  const char* ring_env = getenv("ISSCABAC_SYM_RING");
  if (sym_width == 1 && !(ring_env && ring_env[0] == '0')) {
#define SYM_RING_CASE(PR, ME) \
  if (cfg->profile == PR && cfg->method == ME) rc = launch_sym_wide(k_encode_symbols_ring<PR, ME>, P, st, "k_encode_symbols_ring", done, false, true); else
    SYM_RING_CASE(ISSCABAC_PROFILE_ISS, ISSCABAC_BIN_EG0)
    SYM_RING_CASE(ISSCABAC_PROFILE_FLAT, ISSCABAC_BIN_EG0)
    SYM_RING_CASE(ISSCABAC_PROFILE_FLAT_EPSUF, ISSCABAC_BIN_EG2)
    SYM_RING_CASE(ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_TU)
    SYM_RING_CASE(ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_EG0)
    rc = launch_sym_wide(k_encode_symbols_ring<-1, -1>, P, st, "k_encode_symbols_ring", done, false, true);
#undef SYM_RING_CASE
    if (rc || done) return rc;
  }
#define SYM_WIDE_CASE(K, PR, ME) \
  if (cfg->profile == PR && cfg->method == ME) rc = launch_sym_wide(K<PR, ME>, P, st, #K, done, kWantLut); else
  SYM_WIDE_CASE(k_encode_symbols_wide, ISSCABAC_PROFILE_ISS, ISSCABAC_BIN_EG0)
  SYM_WIDE_CASE(k_encode_symbols_wide, ISSCABAC_PROFILE_FLAT, ISSCABAC_BIN_EG0)
  SYM_WIDE_CASE(k_encode_symbols_wide, ISSCABAC_PROFILE_FLAT_EPSUF, ISSCABAC_BIN_EG2)
  SYM_WIDE_CASE(k_encode_symbols_wide, ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_TU)
  SYM_WIDE_CASE(k_encode_symbols_wide, ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_EG0)
  rc = launch_sym_wide(k_encode_symbols_wide<-1, -1>, P, st, "k_encode_symbols_wide", done, kWantLut);
  if (rc || done) return rc;
  return launch_sym(k_encode_symbols<false>, P, st, "k_encode_symbols");
}

int cabac_decode_symbols(const isscabac_symcfg* cfg, uint32_t n_streams, const uint64_t* d_byte_off,
                         const uint8_t* d_bytes, const uint64_t* d_sym_off,
                         const uint8_t* d_ctx_init, uint32_t n_ctx, int per_stream_init,
                         void* d_symbols, int sym_width, uint8_t* d_finish_ok, void* stream) {
  if (cfg && getenv("ISSCABAC_DEBUG_TREE")) {
    std::vector<TreeNode> nodes;
    TreeInfo ti{0, 0, 0, 0, nullptr};
    const bool ok = build_code_tree(*cfg, nodes, ti);
    uint32_t n_ep = 0;
    for (const TreeNode& nd : nodes) n_ep += nd.code == kTreeEpRun;
    fprintf(stderr, "code tree: built %d, %u nodes, stride %u, %u run nodes\n", (int)ok, ti.n_nodes, ti.var_stride, n_ep);
    for (size_t k = 0; k < nodes.size() && k < 24; ++k)
      fprintf(stderr, "  node %zu: child %04x %04x code %04x pad %u\n", k, nodes[k].child[0], nodes[k].child[1], nodes[k].code, nodes[k].pad);
  }
  int rc = check_cfg(cfg, n_ctx, sym_width, true, true);
  if (rc) return rc;
  if (n_streams == 0) return ISSCABAC_OK;
  if (!d_sym_off || !d_byte_off || !d_bytes || !d_symbols || !d_ctx_init) { set_error("cabac_decode_symbols: null pointer"); return ISSCABAC_ERR_INVALID; }
  SymParams P;
  memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.n_streams = n_streams; P.n_ctx = n_ctx; P.per_stream_init = per_stream_init; P.sym_width = sym_width;
  P.sym_off = d_sym_off; P.ctx_init = d_ctx_init; P.byte_off = d_byte_off; P.bytes = d_bytes;
  P.out_symbols = d_symbols; P.finish_ok = d_finish_ok;
  bool done = false;
  constexpr bool kWantLut = false;   // the decoder's contexts depend on the bins it decodes
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the tree decoder (code tree in shared memory, no per-profile code) for u8 symbols of small alphabets;
  // ISSCABAC_SYM_TREE=0 keeps the closed-form state machine (both are tested against the oracle)
  const char* tree_env = getenv("ISSCABAC_SYM_TREE");
  if (sym_width == 1 && !(tree_env && tree_env[0] == '0')) {
    rc = launch_sym_tree(P, st, done);
    if (rc || done) return rc;
  }
  SYM_WIDE_CASE(k_decode_symbols_wide, ISSCABAC_PROFILE_ISS, ISSCABAC_BIN_EG0)
  SYM_WIDE_CASE(k_decode_symbols_wide, ISSCABAC_PROFILE_FLAT, ISSCABAC_BIN_EG0)
  SYM_WIDE_CASE(k_decode_symbols_wide, ISSCABAC_PROFILE_FLAT_EPSUF, ISSCABAC_BIN_EG2)
  SYM_WIDE_CASE(k_decode_symbols_wide, ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_TU)
  SYM_WIDE_CASE(k_decode_symbols_wide, ISSCABAC_PROFILE_DEMO, ISSCABAC_BIN_EG0)
  rc = launch_sym_wide(k_decode_symbols_wide<-1, -1>, P, st, "k_decode_symbols_wide", done);
  if (rc || done) return rc;
  return launch_sym(k_decode_symbols, P, static_cast<cudaStream_t>(stream), "k_decode_symbols");
}

}  // extern "C"
