// bitslice_mp.cu -- instantiations of the bit-sliced encode / verify kernel (bs_kernel.cuh) for the
// code modes with more than 4 parity shards (EC15P12, EC6P6, EC16P20L2, EC6P10L2, EC12P9, EC24P8,
// EC6P8L10; blobstore/common/codemode/codemode.go:65-94).  8*M accumulator registers per thread cap a
// pass at 4 parity rows, so RS(k, MT) runs ceil(MT/4) passes, each with its own compile-time XOR
// network over the k data shards (rows [R0, R0+4) of the generator), writing slots k+R0 .. k+R0+3.
// The first pass also checksums the data shards; later passes checksum only what they write.
#include "bs_kernel.cuh"

namespace cbe {

int bs_mp_passes(int k, int m, const uint8_t* parity_rows) {
  int n = 0, ok = 1;
#define X(KK, MM, VV, MT, RR)                                          \
  if (k == KK && m == MT) {                                            \
    n++;                                                               \
    ok &= bs_rows_match<KK, MM, VV>(parity_rows) ? 1 : 0;              \
  }
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return ok ? n : 0;
}

cudaError_t launch_bs_mp(int k, int m, int pass, const BsParams& p, int crc, bool verify, int grid, cudaStream_t st) {
#define X(KK, MM, VV, MT, RR)                                                                          \
  if (k == KK && m == MT && pass * 4 == RR) {                                                          \
    return bs_launch_cfg<KK, MM, VV>(p, crc, verify, grid, st);                                        \
  }
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace cbe
