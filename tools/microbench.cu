// microbench.cu -- integer-pipe / shared-memory / streaming micro-benchmarks that size the
// instruction budget of the GF(2^8)+CRC kernels on sm_100a (see DESIGN.md "instruction budget").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int CH = 8;   // independent chains per thread

enum Op { LOP3, IMAD, IMADHI, SHF, PRMT, IADD3, MIX_LOP_IMAD, MIX_LOP_IMADHI, MIX_LOP_SHF, MIX_LOP_PRMT, POPC, IMAD_SHL, DP2A, DP4A, MIX_LOP_DP2A, MIX_LOP2_DP2A };

template <int OP>
__global__ void alu_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  uint32_t a[CH], b[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 11 + blockIdx.x; }
  uint32_t k1 = seed | 1, k2 = (seed >> 3) | 0x10001;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1)); }
      else if (OP == IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(b[i])); }
      else if (OP == IMADHI) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k2), "r"(b[i])); }
      else if (OP == SHF) { asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(k1)); }
      else if (OP == PRMT) { asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(k1)); }
      else if (OP == IADD3) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }
      else if (OP == POPC) { asm volatile("popc.b32 %0, %0;" : "+r"(a[i])); }
      else if (OP == IMAD_SHL) { asm volatile("mul.lo.u32 %0, %0, 16;" : "+r"(a[i])); }
      else if (OP == DP2A) { asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %2;" : "+r"(a[i]) : "r"(k2), "r"(b[i])); }
      else if (OP == DP4A) { asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k2), "r"(b[i])); }
      else if (OP == MIX_LOP_DP2A) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %0, %2;" : "+r"(b[i]) : "r"(k2), "r"(k1));
      } else if (OP == MIX_LOP2_DP2A) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k2));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %0, %2;" : "+r"(b[i]) : "r"(k2), "r"(k1));
      }
      else if (OP == MIX_LOP_IMAD) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(k1), "r"(k2));
      } else if (OP == MIX_LOP_IMADHI) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
        asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(k2), "r"(k1));
      } else if (OP == MIX_LOP_SHF) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
        asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(k2), "r"(k1));
      } else if (OP == MIX_LOP_PRMT) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(k1));
        asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(k2), "r"(k1));
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) acc ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory lookups: MODE 0 = conflict-free (own bank), 1 = random 256-entry table, 2 = R=16 replicated
template <int MODE>
__global__ void lds_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  __shared__ uint32_t tab[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = i * 2654435761u + seed;
  __syncthreads();
  uint32_t x[4];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = (threadIdx.x * 2654435761u + i * 40503u + seed);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t v = (x[i] >> 9) & 0xff;
      uint32_t idx;
      if (MODE == 0) idx = v * 32 + lane;              // lane-private bank
      else if (MODE == 1) idx = v;                      // plain 256-entry table
      else idx = v * 16 + (lane & 15);                  // 16 copies
      x[i] = x[i] * 1664525u + tab[idx];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] ^ x[1] ^ x[2] ^ x[3];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// streaming pattern of the coding kernel: read K shards, write M shards, 16 B per thread
template <int K, int M>
__global__ void stream_kernel(const uint4* __restrict__ in, uint4* __restrict__ outp, size_t shard_vec, size_t n_cols) {
  for (size_t col = blockIdx.x * (size_t)blockDim.x + threadIdx.x; col < n_cols; col += (size_t)gridDim.x * blockDim.x) {
    uint4 d[K];
#pragma unroll
    for (int c = 0; c < K; c++) {
      const uint4* p = in + c * shard_vec + col;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d[c].x), "=r"(d[c].y), "=r"(d[c].z), "=r"(d[c].w) : "l"(p));
    }
    uint4 a = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < K; c++) { a.x ^= d[c].x; a.y ^= d[c].y; a.z ^= d[c].z; a.w ^= d[c].w; }
#pragma unroll
    for (int r = 0; r < M; r++) {
      uint4 o = make_uint4(a.x + r, a.y, a.z, a.w);
      uint4* q = outp + r * shard_vec + col;
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(q), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
    }
  }
}

template <typename F>
static int run_alu(const char* name, F kernel, int ops_per_iter_per_chain, int sms) {
  const int blocks = sms * 2, threads = 512;
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)blocks * threads * 4));
  CK(cudaMalloc(&cyc, blocks * sizeof(long long)));
  kernel<<<blocks, threads>>>(out, cyc, 12345u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kernel<<<blocks, threads>>>(out, cyc, 777u);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(blocks);
  CK(cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : h) avg += v; avg /= blocks;
  double lane_ops_per_sm = 2.0 * threads * (double)ITERS * CH * ops_per_iter_per_chain;
  printf("%-18s %8.1f lane-ops/clk/SM   (%.0f cycles/block, %.3f ms, est clock %.0f MHz)\n", name, lane_ops_per_sm / avg, avg, ms, avg / (ms * 1e3));
  cudaFree(out); cudaFree(cyc);
  return 0;
}

template <typename F>
static int run_lds(const char* name, F kernel, int sms) {
  const int blocks = sms, threads = 1024;
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)blocks * threads * 4));
  CK(cudaMalloc(&cyc, blocks * sizeof(long long)));
  kernel<<<blocks, threads>>>(out, cyc, 1u);
  kernel<<<blocks, threads>>>(out, cyc, 2u);
  CK(cudaDeviceSynchronize());
  std::vector<long long> h(blocks);
  CK(cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (auto v : h) avg += v; avg /= blocks;
  double lookups = (double)threads * ITERS * 4;
  printf("%-18s %8.2f lookups/clk/SM\n", name, lookups / avg);
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clockRate %d kHz\n", prop.name, sms, prop.clockRate);
  run_alu("LOP3", alu_kernel<LOP3>, 1, sms);
  run_alu("IMAD", alu_kernel<IMAD>, 1, sms);
  run_alu("IMAD.HI", alu_kernel<IMADHI>, 1, sms);
  run_alu("IMAD(shl imm)", alu_kernel<IMAD_SHL>, 1, sms);
  run_alu("SHF", alu_kernel<SHF>, 1, sms);
  run_alu("PRMT", alu_kernel<PRMT>, 1, sms);
  run_alu("IADD", alu_kernel<IADD3>, 1, sms);
  run_alu("POPC", alu_kernel<POPC>, 1, sms);
  run_alu("DP2A", alu_kernel<DP2A>, 1, sms);
  run_alu("DP4A", alu_kernel<DP4A>, 1, sms);
  run_alu("LOP3+DP2A", alu_kernel<MIX_LOP_DP2A>, 2, sms);
  run_alu("2xLOP3+DP2A", alu_kernel<MIX_LOP2_DP2A>, 3, sms);
  run_alu("LOP3+IMAD", alu_kernel<MIX_LOP_IMAD>, 2, sms);
  run_alu("LOP3+IMAD.HI", alu_kernel<MIX_LOP_IMADHI>, 2, sms);
  run_alu("LOP3+SHF", alu_kernel<MIX_LOP_SHF>, 2, sms);
  run_alu("LOP3+PRMT", alu_kernel<MIX_LOP_PRMT>, 2, sms);
  run_lds("LDS conflict-free", lds_kernel<0>, sms);
  run_lds("LDS random-256", lds_kernel<1>, sms);
  run_lds("LDS 16-copies", lds_kernel<2>, sms);

  // streaming: 12 reads + 4 writes per column, ~3.7 GiB moved
  {
    const size_t shard_bytes = 224u << 20;   // per shard
    const size_t shard_vec = shard_bytes / 16;
    uint4 *in, *outp;
    CK(cudaMalloc(&in, shard_bytes * 12));
    CK(cudaMalloc(&outp, shard_bytes * 4));
    CK(cudaMemset(in, 1, shard_bytes * 12));
    for (int bs : {256, 512, 1024}) for (int mult : {1, 2, 4, 8}) {
      int blocks = sms * mult * (1024 / bs);
      if (blocks > sms * 32) continue;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      stream_kernel<12, 4><<<blocks, bs>>>(in, outp, shard_vec, shard_vec);
      cudaEventRecord(e0);
      for (int r = 0; r < 3; r++) stream_kernel<12, 4><<<blocks, bs>>>(in, outp, shard_vec, shard_vec);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("stream 12r+4w  block %4d grid %5d : %.1f GB/s\n", bs, blocks, 3.0 * 16.0 * shard_bytes / (ms * 1e6));
    }
    cudaFree(in); cudaFree(outp);
  }
  return 0;
}
