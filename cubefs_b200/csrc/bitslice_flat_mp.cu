// bitslice_flat_mp.cu -- flat-split fused encode + CRC32 kernel (bs_flat.cuh): the passes (4 parity rows
// each, pass plan 0 of gen_bitslice.py) of the code modes with more than 4 parity shards.  The first pass
// checksums the data shards and its outputs (mode 1; mode 2 when the code is an LRC local stripe), later
// passes only what they write (mode 2).
#include "bs_flat.cuh"

namespace cbe {

bool bsf_mp_supported(int k, int m, int pass, int crc_mode) {
#define X(KK, MM, VV, MT, RR, PP, PL) \
  if (PL == 0 && k == KK && m == MT && pass == PP) return crc_mode == 2 || (crc_mode == 1 && RR == 0);
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return false;
}

cudaError_t launch_bsf_mp(int k, int m, int pass, int crc_mode, const BsfParams& p, int grid, cudaStream_t st, int flip) {
#define X(KK, MM, VV, MT, RR, PP, PL)                                              \
  if constexpr (PL == 0) {                                                         \
    if (k == KK && m == MT && pass == PP) {                                        \
      if (crc_mode == 2) return bsf_launch_one<KK, MM, VV, 2>(p, grid, st, flip);      \
      if constexpr (RR == 0) {                                                     \
        if (crc_mode == 1) return bsf_launch_one<KK, MM, VV, 1>(p, grid, st, flip);    \
      }                                                                            \
      return cudaErrorInvalidValue;                                                \
    }                                                                              \
  }
  CUBEEC_BS_PASS_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace cbe
