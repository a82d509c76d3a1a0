// issue_mix.cu -- how many warp instructions per clock an sm_100a SM sub-partition issues for the
// instruction MIXES of the fused encode+CRC kernels (LOP3 on the ALU pipe, IDP.2A/IMAD on the FMA
// pipe, conflict-free LDS), as opposed to one pipe at a time (tools/microbench.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_mix tools/issue_mix.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 20000;
constexpr int CH = 8;

// per chain and iteration: NL lop3, ND dp2a, NM imad, NS lds (conflict-free, address from dp2a chain)
template <int NL, int ND, int NM, int NS>
__global__ void __launch_bounds__(1024, 1) mix_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  __shared__ uint32_t tab[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = ((((i * 2654435761u + seed) >> 7) & 0xff) * 128) | ((i & 31) * 4);   // next address stays in the reader's bank
  __syncthreads();
  const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(tab);
  uint32_t a[CH], b[CH], c[CH], s[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) {
    a[i] = seed + threadIdx.x * 7 + i;
    b[i] = seed * 3 + i * 11 + blockIdx.x;
    c[i] = seed ^ (i * 977);
    s[i] = ((threadIdx.x & 31) * 4 + i * 128) & 0x7ffc;
  }
  uint32_t k1 = seed | 1, k2 = (seed >> 3) | 0x10001;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
#pragma unroll
      for (int j = 0; j < NL; j++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
#pragma unroll
      for (int j = 0; j < ND; j++) asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %2;" : "+r"(b[i]) : "r"(k2), "r"(k1));
#pragma unroll
      for (int j = 0; j < NM; j++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c[i]) : "r"(k1), "r"(k2));
#pragma unroll
      for (int j = 0; j < NS; j++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(s[i]) : "r"(tab_base + s[i]) : "memory");
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) acc ^= a[i] ^ b[i] ^ c[i] ^ s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NL, int ND, int NM, int NS>
static int run(const char* name, int sms, int threads) {
  const int blocks = sms;
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)blocks * threads * 4));
  CK(cudaMalloc(&cyc, blocks * sizeof(long long)));
  mix_kernel<NL, ND, NM, NS><<<blocks, threads>>>(out, cyc, 12345u);
  CK(cudaDeviceSynchronize());
  mix_kernel<NL, ND, NM, NS><<<blocks, threads>>>(out, cyc, 777u);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(blocks);
  CK(cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  const double cyc_max = (double)*std::max_element(h.begin(), h.end());
  const double warps = threads / 32.0;
  const double per_iter = (double)CH * (NL + ND + NM + NS);
  const double ipc_smsp = warps * per_iter * ITERS / cyc_max / 4.0;
  printf("%-34s warps/SM %2.0f  IPC/SMSP %.3f   (ALU %.3f  FMA %.3f  LDS/clk/SM %.3f)\n", name, warps, ipc_smsp,
         ipc_smsp * NL / (NL + ND + NM + NS), ipc_smsp * (ND + NM) / (NL + ND + NM + NS), 4 * ipc_smsp * NS / (NL + ND + NM + NS));
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  for (int threads : {512, 768, 1024}) {
    run<1, 0, 0, 0>("LOP3", sms, threads);
    run<0, 1, 0, 0>("IDP.2A", sms, threads);
    run<0, 0, 1, 0>("IMAD", sms, threads);
    run<0, 0, 0, 1>("LDS (conflict-free, dependent)", sms, threads);
    run<1, 1, 0, 0>("LOP3:IDP 1:1", sms, threads);
    run<2, 1, 0, 0>("LOP3:IDP 2:1", sms, threads);
    run<1, 0, 1, 0>("LOP3:IMAD 1:1", sms, threads);
    run<2, 1, 0, 1>("LOP3:IDP:LDS 2:1:1", sms, threads);
    run<3, 1, 1, 1>("LOP3:IDP:IMAD:LDS 3:1:1:1", sms, threads);
    run<1, 1, 0, 1>("LOP3:IDP:LDS 1:1:1", sms, threads);
    run<0, 1, 0, 1>("IDP:LDS 1:1", sms, threads);
  }
  return 0;
}
