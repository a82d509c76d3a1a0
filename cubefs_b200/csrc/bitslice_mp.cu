// bitslice_mp.cu -- instantiations of the bit-sliced encode / verify kernel (bs_kernel.cuh) for the
// code modes with more than 4 parity shards (EC15P12, EC6P6, EC16P20L2, EC6P10L2, EC12P9, EC24P8,
// EC6P8L10; blobstore/common/codemode/codemode.go:65-94).  The 8*M accumulator registers of a thread
// cap the parity rows of a pass, so RS(k, MT) runs several passes, each with its own compile-time XOR
// network over the k data shards (rows [R0, R0+M) of the generator), writing slots k+R0 ...
//   plan 0 (fused CRC): 4 rows per pass; the first pass also checksums the data shards, later passes
//                       only what they write;
//   plan 1 (plain encode, verify): 6 rows per pass -- EC15P12 reads its data twice, EC6P6 once.
#include "bs_kernel.cuh"

namespace cbe {

int bs_mp_passes(int k, int m, const uint8_t* parity_rows, int plan) {
  int n = 0, ok = 1;
#define X(KK, MM, VV, MT, RR, PP, PL)                                  \
  if (k == KK && m == MT && plan == PL) {                              \
    n++;                                                               \
    ok &= bs_rows_match<KK, MM, VV>(parity_rows) ? 1 : 0;              \
  }
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return ok ? n : 0;
}

// rows [*r0, *r0 + *rows) of the generator that pass `pass` of plan `plan` computes; false if there is none
bool bs_mp_pass_rows(int k, int m, int plan, int pass, int* r0, int* rows) {
#define X(KK, MM, VV, MT, RR, PP, PL)                        \
  if (k == KK && m == MT && pass == PP && plan == PL) {      \
    *r0 = RR;                                                \
    *rows = MM;                                              \
    return true;                                             \
  }
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return false;
}

cudaError_t launch_bs_mp(int k, int m, int pass, const BsParams& p, int crc, bool verify, int grid, cudaStream_t st) {
  const int plan = crc ? 0 : 1;
#define X(KK, MM, VV, MT, RR, PP, PL)                                                                  \
  if (k == KK && m == MT && pass == PP && plan == PL) return bs_launch_cfg<KK, MM, VV, PL == 1 ? 2 : (RR == 0 ? 5 : 4)>(p, crc, verify, grid, st);
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace cbe
