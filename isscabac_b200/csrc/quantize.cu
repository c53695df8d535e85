// quantize.cu -- the quantiser in front of the coder on the device:
// ISS/quantizeWrapper.m:1-88 (dead zone by a data quantile, then Lloyd-Max / uniform / fixed
// centroids on the rest) with quantizeLloyd (:91-176) and ISS/quantize.m:59-84 behind it.
// Output per matrix: the group index of every element minus one -- the symbols the coder takes
// (ISS.m:108-110: param.gW = miscW.group-1) -- and the N centroids.
//
// One CTA per matrix (W 400x20, H 109x20 in ISS.m; a batch of tracks = thousands of matrices).
// The reference walks masks over the unsorted data on every Lloyd iteration; here the data is sorted
// once (bitonic, shared memory), a prefix sum over the sorted values gives any group's mean as a
// difference of two entries, and a group is an interval of the sorted array found by binary search,
// so one iteration costs one warp a few dozen instructions whatever the matrix size.  Double
// precision throughout like the reference.  The sums are taken in sorted order, the reference's in
// storage order: centroids agree to rounding (1e-12 relative in tests/test_gpu_quantize.py), group
// indices are identical except for elements that sit within rounding of a decision threshold.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/isscabac.h"
#include "internal.h"

using namespace isscabac_internal;

namespace {

constexpr int QT = 512;            // threads per CTA
constexpr int QMAXN = 32;          // centroids: one lane per group
constexpr uint32_t QSMEM_ELEMS = 8192;   // largest padded matrix whose sort + prefix arrays fit shared memory

struct QParams {
  isscabac_quantcfg cfg;
  uint32_t n_mat, npad_max;
  const uint64_t* off;
  const double* x;
  const double* fixed;
  uint8_t* groups;
  double* centroids;
  uint32_t* iters;
  double* scratch;   // (2 * npad_max + 2) doubles per CTA when the arrays do not fit shared memory
};

// first index in [0, n) with a[idx] >= v (n when there is none)
__device__ __forceinline__ uint32_t lower_bound(const double* a, uint32_t n, double v) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// the same, starting from where the boundary was in the previous iteration (it moves by a few elements late in
// Lloyd's iteration): exponential bracket around the hint, then bisection inside it
__device__ __forceinline__ uint32_t lower_bound_from(const double* a, uint32_t n, double v, uint32_t hint) {
  if (hint > n) hint = n;
  uint32_t lo, hi;
  if (hint < n && a[hint] < v) {          // answer is to the right of the hint
    uint32_t step = 1;
    lo = hint + 1;
    hi = n;
    while (lo + step <= n && a[lo + step - 1] < v) { lo += step; step <<= 1; }
    if (lo + step < hi) hi = lo + step;
    if (hi > n) hi = n;
  } else {                                // a[hint] >= v (or hint == n): answer is at the hint or to its left
    uint32_t step = 1;
    hi = hint;
    lo = 0;
    while (hi >= step && !(a[hi - step] < v)) { hi -= step; step <<= 1; }
    if (hi >= step) lo = hi - step + 1;   // a[hi - step] < v
  }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// MATLAB quantile of a sorted vector (prctile: sample i is the (i - 0.5)/n quantile); quantizeWrapper.m:23, quantize.m:61
__device__ __forceinline__ double sorted_quantile(const double* xs, uint32_t n, double p) {
  double r = p * (double)n;
  long long k = (long long)floor(r + 0.5);
  long long kp1 = k + 1;
  r -= (double)k;
  if (k < 1) k = 1;
  if (k > (long long)n) k = n;
  if (kp1 > (long long)n) kp1 = n;
  return (0.5 + r) * xs[kp1 - 1] + (0.5 - r) * xs[k - 1];
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sorted copy of c[0..N) over the lanes (quantize.m:79 `centroids = sort(centroids)`): rank sort, N <= 32
__device__ __forceinline__ double warp_sort(double c, int g, int N) {
  int rank = 0;
  for (int h = 0; h < N; ++h) {
    const double o = __shfl_sync(0xffffffffu, c, h);
    if (g < N && (o < c || (o == c && h < g))) ++rank;
  }
  double out = 0.0;
  for (int h = 0; h < N; ++h) {
    const double o = __shfl_sync(0xffffffffu, c, h);
    const int r = __shfl_sync(0xffffffffu, rank, h);
    if (r == g) out = o;
  }
  return out;
}

__global__ void __launch_bounds__(QT) k_quantize(QParams P) {
  extern __shared__ __align__(16) uint8_t q_smem[];
  __shared__ double s_part[QT / 32];
  __shared__ double s_cent[QMAXN];       // final centroids of the part outside the dead zone (sorted)
  __shared__ uint32_t s_iters;
  const bool in_smem = P.npad_max <= QSMEM_ELEMS;
  double* xs = in_smem ? reinterpret_cast<double*>(q_smem) : P.scratch + (size_t)blockIdx.x * (2ull * P.npad_max + 2);
  double* ps = xs + P.npad_max;          // ps[j] = sum of the j smallest values
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int N = P.cfg.N;

  for (uint32_t m = blockIdx.x; m < P.n_mat; m += gridDim.x) {
    const uint64_t o0 = P.off[m];
    const uint32_t n = (uint32_t)(P.off[m + 1] - o0);
    const double* xin = P.x + o0;
    uint8_t* gout = P.groups + o0;
    double* cout = P.centroids + (size_t)m * N;
    __syncthreads();   // the previous matrix is done with the arrays
    if (n == 0) {
      if (tid < N) cout[tid] = 0.0;
      if (tid == 0 && P.iters) P.iters[m] = 0;
      continue;
    }
    uint32_t npad = 2;
    while (npad < n) npad <<= 1;
    if (in_smem && npad >= 16) {
      // ---- bitonic sort, ascending, register-blocked: a thread owns chunks of 16 consecutive elements, so the stages
      // whose partner distance is below 16 (four per phase, and all of the first four phases) run in registers without a
      // barrier, and only distances >= 16 -- conflict-free runs of consecutive doubles -- go through shared memory.
      // One double of padding per 16 keeps the chunk loads of a warp off a single bank; the workspace overlays xs and
      // the (not yet used) prefix array, the sorted values are moved to xs unpadded at the end.
      double* wsp = xs;
      auto PADX = [](uint32_t i) { return i + (i >> 4); };
      for (uint32_t i = tid; i < npad; i += QT) wsp[PADX(i)] = i < n ? xin[i] : INFINITY;
      __syncthreads();
      auto cx = [](double& a, double& b, bool asc) {
        if ((a > b) == asc) { const double t = a; a = b; b = t; }
      };
      const uint32_t nchunks = npad >> 4;
      for (uint32_t c = tid; c < nchunks; c += QT) {        // phases k = 2, 4, 8, 16 entirely in registers
        double v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = wsp[17u * c + e];
#pragma unroll
        for (int k = 2; k <= 16; k <<= 1)
#pragma unroll
          for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if ((e & j) == 0) cx(v[e], v[e | j], k < 16 ? (e & k) == 0 : (c & 1u) == 0);
#pragma unroll
        for (int e = 0; e < 16; ++e) wsp[17u * c + e] = v[e];
      }
      __syncthreads();
      for (uint32_t k = 32; k <= npad; k <<= 1) {
        for (uint32_t j = k >> 1; j >= 16; j >>= 1) {
          for (uint32_t t = tid; t < (npad >> 1); t += QT) {
            const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
            const double a = wsp[PADX(i)], b = wsp[PADX(l)];
            if ((a > b) == ((i & k) == 0)) { wsp[PADX(i)] = b; wsp[PADX(l)] = a; }
          }
          __syncthreads();
        }
        for (uint32_t c = tid; c < nchunks; c += QT) {      // distances 8, 4, 2, 1 of this phase
          double v[16];
          const bool asc = ((c << 4) & k) == 0;
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = wsp[17u * c + e];
#pragma unroll
          for (int j = 8; j > 0; j >>= 1)
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if ((e & j) == 0) cx(v[e], v[e | j], asc);
#pragma unroll
          for (int e = 0; e < 16; ++e) wsp[17u * c + e] = v[e];
        }
        __syncthreads();
      }
      // unpadded copy: through registers, the two layouts overlap
      double r[QSMEM_ELEMS / QT];
#pragma unroll
      for (uint32_t m = 0; m < QSMEM_ELEMS / QT; ++m) {
        const uint32_t i = tid + m * QT;
        r[m] = i < npad ? wsp[PADX(i)] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (uint32_t m = 0; m < QSMEM_ELEMS / QT; ++m) {
        const uint32_t i = tid + m * QT;
        if (i < npad) xs[i] = r[m];
      }
      __syncthreads();
    } else {
      for (uint32_t i = tid; i < npad; i += QT) xs[i] = i < n ? xin[i] : INFINITY;
      __syncthreads();
      // ---- bitonic sort, ascending (matrices beyond the shared-memory budget, in global scratch; tiny ones)
      for (uint32_t k = 2; k <= npad; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          for (uint32_t t = tid; t < (npad >> 1); t += QT) {
            const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
            const double a = xs[i], b = xs[l];
            if ((a > b) == ((i & k) == 0)) { xs[i] = b; xs[l] = a; }
          }
          __syncthreads();
        }
      }
    }
    // ---- prefix sums of the sorted values (ps[i] = sum of the i smallest).  Every warp takes a contiguous span and walks it
    // in rows of 32 consecutive elements -- lane = element, so shared memory is read and written without bank conflicts (a
    // contiguous run per THREAD put all 32 lanes of an access on one bank) -- with a shuffle scan per row and a running carry.
    {
      const uint32_t span = ((n + QT - 1) / QT) * 32u;        // elements per warp, a multiple of 32
      const uint32_t w0 = min((uint32_t)wid * span, n), w1 = min(w0 + span, n);
      double tot = 0.0;
      for (uint32_t i = w0 + lane; i < w1; i += 32) tot += xs[i];
      tot = warp_sum(tot);
      if (lane == 0) s_part[wid] = tot;
      __syncthreads();
      double carry = 0.0;
      for (int w = 0; w < wid; ++w) carry += s_part[w];
      for (uint32_t row = w0; row < w1; row += 32) {
        const uint32_t i = row + lane;
        const double v = i < w1 ? xs[i] : 0.0;
        double inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double t = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += t;
        }
        if (i < w1) ps[i] = carry + (inc - v);
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (tid == 0) {
        double all = 0.0;
        for (int w = 0; w < QT / 32; ++w) all += s_part[w];
        ps[n] = all;
      }
      __syncthreads();
    }

    // ---- dead zone (quantizeWrapper.m:22-36): everything below the data's deadzone_quant quantile
    double thr = -INFINITY;
    uint32_t m_dz = 0;
    if (P.cfg.deadzone_quant >= 0.0) {
      thr = sorted_quantile(xs, n, P.cfg.deadzone_quant);
      m_dz = lower_bound(xs, n, thr);
    }
    const bool has_dz = m_dz != 0;
    const int Nr = has_dz ? N - 1 : N;           // centroids for the rest
    const double* xr = xs + m_dz;
    const double* pr = ps + m_dz;
    const uint32_t nr = n - m_dz;                 // > 0: the quantile never exceeds the maximum

    // ---- centroids of the rest: one warp, lane g = group g
    if (wid == 0) {
      const int g = lane;
      const double mn = xr[0], mx = xr[nr - 1];
      double c = 0.0;
      uint32_t iters = 0;
      if (P.cfg.mode == ISSCABAC_QUANT_FIXED) {                       // quantizeWrapper.m:39-40
        c = g < Nr ? P.fixed[g + (has_dz ? 1 : 0)] : 0.0;
        c = warp_sort(c, g, Nr);
      } else {
        // quantize.m:61-75: linspace between the border quantiles ([0 1] = min and max for Lloyd's start)
        double a = mn, b = mx;
        if (P.cfg.mode == ISSCABAC_QUANT_UNIFORM) {
          a = sorted_quantile(xr, nr, P.cfg.q_lo);
          b = sorted_quantile(xr, nr, P.cfg.q_hi);
        }
        c = (Nr == 1 || g == Nr - 1) ? b : a + (double)g * (b - a) / (double)(Nr - 1);
        if (a > b) c = warp_sort(c, g, Nr);
      }
      if (P.cfg.mode == ISSCABAC_QUANT_LLOYD) {                       // quantizeLloyd, quantizeWrapper.m:121-171
        // group g = the sorted values in [lo, hi): edges(k) <= x < edges(k+1) (quantize.m:81)
        double up = __shfl_down_sync(0xffffffffu, c, 1);
        double mid = 0.5 * (c + up);                                   // edge between groups g and g+1
        uint32_t hi = g < Nr - 1 ? lower_bound(xr, nr, mid) : nr;
        uint32_t lo = __shfl_up_sync(0xffffffffu, hi, 1);
        if (g == 0) lo = 0;
        // edges = [min(min(x),min(c)), mids, max(max(x),max(c))] (:135): e_lo / e_hi = this group's two edges
        double cmin = warp_min(g < Nr ? c : INFINITY), cmax = warp_max(g < Nr ? c : -INFINITY);
        double e_hi = g < Nr - 1 ? mid : fmax(mx, cmax);
        double e_lo = __shfl_up_sync(0xffffffffu, e_hi, 1);
        if (g == 0) e_lo = fmin(mn, cmin);
        double old = c;
        for (int it = 0; it < P.cfg.max_iter; ++it) {
          ++iters;
          double cn = 0.5 * (e_lo + e_hi);                             // :143
          if (g < Nr && hi > lo) cn = (pr[hi] - pr[lo]) / (double)(hi - lo);   // :146-151, mean of the group
          // edges from the updated centroids in group order (:156), groups from the sorted ones (:159-161)
          up = __shfl_down_sync(0xffffffffu, cn, 1);
          // the updated centroids are ordered in practice (means of ordered groups, midpoints of their edges): then the
          // sort is the identity and min / max are the two ends
          const bool ordered = __all_sync(0xffffffffu, g >= Nr - 1 || !(up < cn));
          if (ordered) {
            cmin = __shfl_sync(0xffffffffu, cn, 0);
            cmax = __shfl_sync(0xffffffffu, cn, Nr - 1);
          } else {
            cmin = warp_min(g < Nr ? cn : INFINITY);
            cmax = warp_max(g < Nr ? cn : -INFINITY);
          }
          e_hi = g < Nr - 1 ? 0.5 * (cn + up) : fmax(mx, cmax);
          e_lo = __shfl_up_sync(0xffffffffu, e_hi, 1);
          if (g == 0) e_lo = fmin(mn, cmin);
          c = ordered ? cn : warp_sort(cn, g, Nr);
          up = __shfl_down_sync(0xffffffffu, c, 1);
          mid = 0.5 * (c + up);
          hi = g < Nr - 1 ? lower_bound_from(xr, nr, mid, hi) : nr;
          lo = __shfl_up_sync(0xffffffffu, hi, 1);
          if (g == 0) lo = 0;
          const double d = g < Nr ? (c - old) * (c - old) : 0.0;
          if (warp_sum(d) / (double)Nr < P.cfg.tol) break;            // :167
          old = c;
        }
      }
      if (g < Nr) s_cent[g] = c;
      if (g == 0) s_iters = iters;
    }
    __syncthreads();
    // ---- outputs: centroids (dead-zone mean first, :57) and the group of every element minus one
    if (tid < N) cout[tid] = has_dz ? (tid == 0 ? ps[m_dz] / (double)m_dz : s_cent[tid - 1]) : s_cent[tid];
    if (tid == 0 && P.iters) P.iters[m] = s_iters;
    for (uint32_t i = tid; i < n; i += QT) {
      const double v = xin[i];
      uint32_t grp;
      if (has_dz && v < thr) {
        grp = 0;
      } else {
        uint32_t cnt = 0;                                              // number of group edges <= v
        for (int k = 0; k + 1 < Nr; ++k) cnt += (0.5 * (s_cent[k] + s_cent[k + 1]) <= v) ? 1u : 0u;
        grp = cnt + (has_dz ? 1u : 0u);
      }
      gout[i] = (uint8_t)grp;
    }
  }
}

uint32_t pad_pow2(uint64_t n) {
  uint32_t p = 2;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

extern "C" {

size_t cabac_quantize_scratch_bytes(uint32_t n_matrices, uint64_t max_elems) {
  const uint32_t npad = pad_pow2(max_elems);
  if (npad <= QSMEM_ELEMS) return 16;
  const uint32_t grid = (uint32_t)sm_count() * 2u < n_matrices ? (uint32_t)sm_count() * 2u : n_matrices;
  return (size_t)(grid ? grid : 1) * (2ull * npad + 2) * sizeof(double);
}

int cabac_quantize_matrices(const isscabac_quantcfg* cfg, uint32_t n_matrices, const uint64_t* d_elem_off,
                            uint64_t max_elems, const double* d_x, const double* d_fixed_centroids,
                            uint8_t* d_groups, double* d_centroids, uint32_t* d_iters, void* d_scratch, void* stream) {
  if (!cfg) { set_error("quantcfg is NULL"); return ISSCABAC_ERR_INVALID; }
  if (cfg->N < 1 || cfg->N > QMAXN) { set_error("N must be 1..%d", QMAXN); return ISSCABAC_ERR_INVALID; }
  if (cfg->mode != ISSCABAC_QUANT_UNIFORM && cfg->mode != ISSCABAC_QUANT_LLOYD && cfg->mode != ISSCABAC_QUANT_FIXED) {
    set_error("unknown quantiser mode %d", cfg->mode);
    return ISSCABAC_ERR_INVALID;
  }
  if (cfg->deadzone_quant >= 0.0 && cfg->N < 2) { set_error("a dead zone needs N >= 2"); return ISSCABAC_ERR_INVALID; }
  if (cfg->deadzone_quant > 1.0) { set_error("deadzone_quant must be a probability (or negative for none)"); return ISSCABAC_ERR_INVALID; }
  if (cfg->mode == ISSCABAC_QUANT_FIXED && !d_fixed_centroids) { set_error("fixed mode needs centroids"); return ISSCABAC_ERR_INVALID; }
  if (n_matrices == 0) return ISSCABAC_OK;
  if (!d_elem_off || !d_x || !d_groups || !d_centroids) { set_error("cabac_quantize_matrices: null pointer"); return ISSCABAC_ERR_INVALID; }
  if (max_elems > (1ull << 26)) { set_error("matrix too large for one CTA"); return ISSCABAC_ERR_UNSUPPORTED; }
  QParams P;
  memset(&P, 0, sizeof P);
  P.cfg = *cfg; P.n_mat = n_matrices; P.npad_max = pad_pow2(max_elems);
  P.off = d_elem_off; P.x = d_x; P.fixed = d_fixed_centroids; P.groups = d_groups; P.centroids = d_centroids;
  P.iters = d_iters; P.scratch = static_cast<double*>(d_scratch);
  size_t smem = 0;
  uint32_t grid = n_matrices;
  if (P.npad_max <= QSMEM_ELEMS) {
    smem = (2ull * P.npad_max + 2) * sizeof(double);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_quantize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  } else {
    if (!d_scratch) { set_error("matrices of more than %u elements need the scratch buffer", QSMEM_ELEMS); return ISSCABAC_ERR_INVALID; }
    const uint32_t cap = (uint32_t)sm_count() * 2u;
    if (grid > cap) grid = cap;
  }
  k_quantize<<<grid, QT, smem, static_cast<cudaStream_t>(stream)>>>(P);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ISSCABAC_OK : cuda_fail(e, "k_quantize");
}

}  // extern "C"
