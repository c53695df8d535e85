// int_peak.cu -- measured int32 issue peaks of one B200 SM ("measure it").
//
// Every thread runs CH independent dependency chains of one instruction kind (or an interleaved mix of
// two kinds) for ITER iterations, so the warp schedulers always have an eligible instruction; each
// CTA times itself with clock64() and the host reports
//     lanes per clock per SM = resident threads x instructions per thread / SM cycles.
// The instruction kinds are the ones the CABAC kernels are made of (cabac_wide.cuh): IADD3, LOP3,
// SHF (funnel shift), PRMT, ISETP+SEL on the ALU pipe; IMAD on the FMA pipe; BFIND (FLO) on the
// transcendental pipe; LDS.32 / LDS.128 conflict-free.  Inline PTX with opaque operands so that ptxas
// can neither fold the chains nor move work between the pipes (checked with cuobjdump -sass).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_bin/int_peak tools/int_peak.cu
//   tools/_bin/int_peak > profiles/int_peak.json
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

enum Kind { K_IADD3, K_LOP3, K_SHF, K_PRMT, K_SEL, K_IMAD, K_BFIND, K_LDS32, K_LDS128, K_NKINDS };

template <int K>
__device__ __forceinline__ void step(uint32_t& x, uint32_t a, uint32_t b, uint32_t saddr) {
  if (K == K_IADD3) asm volatile("{\n\t.reg .u32 t;\n\tadd.u32 t, %0, %1;\n\tadd.u32 %0, t, %2;\n\t}" : "+r"(x) : "r"(a), "r"(b));
  else if (K == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
  else if (K == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
  else if (K == K_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
  else if (K == K_SEL) asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %0, %1;\n\tselp.u32 %0, %2, %0, q;\n\t}" : "+r"(x) : "r"(a), "r"(b));
  else if (K == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
  else if (K == K_BFIND) asm volatile("bfind.u32 %0, %0;" : "+r"(x));
  else if (K == K_LDS32) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(saddr + x) : "memory");   // the table holds zeros: a dependent chain of loads
  else if (K == K_LDS128) {
    uint32_t y, z, w;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(saddr + x) : "memory");
    x |= y | z | w;
  }
}

constexpr int CH = 8;        // independent chains per thread
constexpr int ITER = 4096;

// NA instructions of kind A followed by NB of kind B per chain and iteration
template <int A, int NA, int B, int NB>
__global__ void __launch_bounds__(1024) k_peak(uint32_t a, uint32_t b, uint32_t* sink, unsigned long long* cycles, uint32_t active_lanes) {
  __shared__ __align__(16) uint32_t sm[32 * 4 * 33];
  for (int i = threadIdx.x; i < 32 * 4 * 33; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 16u;
  uint32_t x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = (A == K_LDS32 || A == K_LDS128) ? 0u : threadIdx.x * 2654435761u + c;
  __syncthreads();
  if ((threadIdx.x & 31u) >= active_lanes) return;    // partial warps: does a half-empty warp issue faster?
  __syncwarp(__activemask());
  const long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int r = 0; r < NA; ++r)
#pragma unroll
      for (int c = 0; c < CH; ++c) step<A>(x[c], a, b, saddr);
#pragma unroll
    for (int r = 0; r < NB; ++r)
#pragma unroll
      for (int c = 0; c < CH; ++c) step<B>(x[c], a, b, saddr);
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc ^= x[c];
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

struct Result { std::string name; double lanes_per_clk_sm; double inst_per_clk_smsp; };

template <int A, int NA, int B, int NB>
Result run(const char* name, int sms, int threads, uint32_t* d_sink, unsigned long long* d_cyc, uint32_t lanes = 32) {
  // one CTA per SM (1024 threads = 8 warps per scheduler, 8 chains each: 64 independent instructions per scheduler)
  k_peak<A, NA, B, NB><<<sms, threads>>>(3u, 0x5410u, d_sink, d_cyc, lanes);
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    k_peak<A, NA, B, NB><<<sms, threads>>>(3u, 0x5410u, d_sink, d_cyc, lanes);
    if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", name); exit(1); }
    std::vector<unsigned long long> c(sms);
    cudaMemcpy(c.data(), d_cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    std::sort(c.begin(), c.end());
    const double med = (double)c[sms / 2];
    const double inst = (double)threads * CH * (NA + NB) * ITER;     // thread-instructions per SM (32 lanes per warp counted: warp-instruction rate)
    best = std::max(best, inst / med);
  }
  return Result{name, best, best / 32.0 / 4.0};
}

int main() {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  const int sms = p.multiProcessorCount;
  uint32_t* d_sink;
  unsigned long long* d_cyc;
  cudaMalloc(&d_sink, 64);
  cudaMalloc(&d_cyc, sms * sizeof(unsigned long long));
  std::vector<Result> r;
  const int T = 1024;
  r.push_back(run<K_IADD3, 1, K_IADD3, 0>("IADD3", sms, T, d_sink, d_cyc));
  r.push_back(run<K_LOP3, 1, K_LOP3, 0>("LOP3", sms, T, d_sink, d_cyc));
  r.push_back(run<K_SHF, 1, K_SHF, 0>("SHF", sms, T, d_sink, d_cyc));
  r.push_back(run<K_PRMT, 1, K_PRMT, 0>("PRMT", sms, T, d_sink, d_cyc));
  r.push_back(run<K_SEL, 1, K_SEL, 0>("ISETP+SEL (2 instr)", sms, T, d_sink, d_cyc));
  r.push_back(run<K_IMAD, 1, K_IMAD, 0>("IMAD", sms, T, d_sink, d_cyc));
  r.push_back(run<K_BFIND, 1, K_BFIND, 0>("BFIND", sms, T, d_sink, d_cyc));
  r.push_back(run<K_LDS32, 1, K_LDS32, 0>("LDS.32", sms, T, d_sink, d_cyc));
  r.push_back(run<K_LDS128, 1, K_LDS128, 0>("LDS.128", sms, T, d_sink, d_cyc));
  r.push_back(run<K_IADD3, 1, K_LOP3, 1>("ALU mix IADD3+LOP3", sms, T, d_sink, d_cyc));
  r.push_back(run<K_IADD3, 1, K_IMAD, 1>("IADD3:IMAD 1:1", sms, T, d_sink, d_cyc));
  r.push_back(run<K_LOP3, 2, K_IMAD, 1>("LOP3:IMAD 2:1", sms, T, d_sink, d_cyc));
  r.push_back(run<K_SHF, 3, K_IMAD, 1>("SHF:IMAD 3:1", sms, T, d_sink, d_cyc));
  // one warp per scheduler (the latency regime of the CABAC kernels), full and half-empty warps: warp-instructions per clock
  std::vector<Result> lone;
  lone.push_back(run<K_IADD3, 1, K_IADD3, 0>("1 warp/SMSP, 32 lanes, IADD3", sms, 128, d_sink, d_cyc, 32));
  lone.push_back(run<K_IADD3, 1, K_IADD3, 0>("1 warp/SMSP, 16 lanes, IADD3", sms, 128, d_sink, d_cyc, 16));
  lone.push_back(run<K_IADD3, 1, K_IADD3, 0>("1 warp/SMSP, 8 lanes, IADD3", sms, 128, d_sink, d_cyc, 8));
  lone.push_back(run<K_IADD3, 1, K_IMAD, 1>("1 warp/SMSP, 32 lanes, IADD3:IMAD 1:1", sms, 128, d_sink, d_cyc, 32));
  lone.push_back(run<K_IADD3, 1, K_IMAD, 1>("1 warp/SMSP, 16 lanes, IADD3:IMAD 1:1", sms, 128, d_sink, d_cyc, 16));
  lone.push_back(run<K_IMAD, 1, K_IMAD, 0>("1 warp/SMSP, 32 lanes, IMAD", sms, 128, d_sink, d_cyc, 32));
  lone.push_back(run<K_SEL, 1, K_SEL, 0>("1 warp/SMSP, 32 lanes, ISETP+SEL", sms, 128, d_sink, d_cyc, 32));
  lone.push_back(run<K_LDS128, 1, K_LDS128, 0>("1 warp/SMSP, 32 lanes, LDS.128 chains", sms, 128, d_sink, d_cyc, 32));
  // ISETP+SEL counts two instructions per step
  r[4].lanes_per_clk_sm *= 2; r[4].inst_per_clk_smsp *= 2;
  double alu = 0, both = 0;
  for (size_t i = 0; i < 4; ++i) alu = std::max(alu, r[i].lanes_per_clk_sm);
  both = r[10].lanes_per_clk_sm;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_domain\": \"SM cycles (clock64), independent of the SM frequency\",\n", p.name, sms);
  printf(" \"threads_per_sm\": %d, \"chains_per_thread\": %d,\n \"kinds\": {", T, CH);
  for (size_t i = 0; i < r.size(); ++i)
    printf("%s\n  \"%s\": {\"lanes_per_clk_per_sm\": %.2f, \"warp_instr_per_clk_per_smsp\": %.3f}", i ? "," : "", r[i].name.c_str(),
           r[i].lanes_per_clk_sm, r[i].inst_per_clk_smsp);
  printf("\n },\n \"one_warp_per_scheduler\": {");
  for (size_t i = 0; i < lone.size(); ++i)
    printf("%s\n  \"%s\": {\"warp_instr_per_clk_per_smsp\": %.3f}", i ? "," : "", lone[i].name.c_str(), lone[i].inst_per_clk_smsp * (i == 6 ? 2 : 1));
  printf("\n },\n \"alu_pipe_lanes_per_clk_per_sm\": %.2f,\n \"alu_plus_fma_lanes_per_clk_per_sm\": %.2f\n}\n", alu, both);
  return 0;
}
