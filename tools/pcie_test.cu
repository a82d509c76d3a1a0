// pcie_test.cu -- how fast can pinned host <-> device copies go on this box, 1-D vs 2-D (pitch-converting)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
int main() {
  const size_t S = 349526, P = 349696, k = 12, m = 4, ns = 256;
  uint8_t *h, *d;
  CK(cudaMallocHost(&h, ns * (k + m) * S));
  CK(cudaMalloc(&d, ns * (k + m) * P));
  cudaStream_t s1, s2;
  CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; rep++) {
    // (a) one big 1-D H2D
    cudaEventRecord(e0, s1);
    CK(cudaMemcpyAsync(d, h, ns * k * S, cudaMemcpyHostToDevice, s1));
    cudaEventRecord(e1, s1); CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("1-D H2D one copy      : %.1f GB/s\n", ns * k * S / (ms * 1e6));
    // (b) per-stripe 1-D H2D (k*S each)
    cudaEventRecord(e0, s1);
    for (size_t s = 0; s < ns; s++) CK(cudaMemcpyAsync(d + s * (k + m) * P, h + s * (k + m) * S, k * S, cudaMemcpyHostToDevice, s1));
    cudaEventRecord(e1, s1); CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("1-D H2D per stripe    : %.1f GB/s\n", ns * k * S / (ms * 1e6));
    // (c) per-stripe 2-D H2D (12 rows, pitch S -> P)
    cudaEventRecord(e0, s1);
    for (size_t s = 0; s < ns; s++) CK(cudaMemcpy2DAsync(d + s * (k + m) * P, P, h + s * (k + m) * S, S, S, k, cudaMemcpyHostToDevice, s1));
    cudaEventRecord(e1, s1); CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("2-D H2D per stripe    : %.1f GB/s\n", ns * k * S / (ms * 1e6));
    // (d) 1-D D2H one copy
    cudaEventRecord(e0, s1);
    CK(cudaMemcpyAsync(h, d, ns * m * S, cudaMemcpyDeviceToHost, s1));
    cudaEventRecord(e1, s1); CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("1-D D2H one copy      : %.1f GB/s\n", ns * m * S / (ms * 1e6));
    // (e) concurrent: H2D per-stripe 1-D on s1, D2H per-stripe 2-D on s2
    cudaEventRecord(e0, s1);
    for (size_t s = 0; s < ns; s++) {
      CK(cudaMemcpyAsync(d + s * (k + m) * P, h + s * (k + m) * S, k * S, cudaMemcpyHostToDevice, s1));
      CK(cudaMemcpy2DAsync(h + s * (k + m) * S + k * S, S, d + s * (k + m) * P + k * P, P, S, m, cudaMemcpyDeviceToHost, s2));
    }
    cudaEventRecord(e1, s1); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); cudaEventElapsedTime(&ms, e0, e1);
    printf("duplex H2D(1-D)+D2H(2-D): H2D %.1f GB/s (D2H overlapped)\n", ns * k * S / (ms * 1e6));
    // (f) device repitch D2D 2-D
    cudaEventRecord(e0, s1);
    for (size_t s = 0; s < ns; s++) CK(cudaMemcpy2DAsync(d + s * (k + m) * P, P, d + (ns - 1 - s) * (k + m) * P, S, S, k, cudaMemcpyDeviceToDevice, s1));
    cudaEventRecord(e1, s1); CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("D2D 2-D repitch       : %.1f GB/s\n", ns * k * S / (ms * 1e6));
  }
  return 0;
}
