// bitslice_flat.cu -- instantiations of the flat-split fused encode + CRC32 kernel (bs_flat.cuh) for the
// single-pass code modes (m <= 4) and the LRC local stripes; the passes of the m > 4 codes are in
// bitslice_flat_mp.cu.  See bs_flat.cuh for the design.
#include "bs_flat.cuh"

namespace cbe {

bool bsf_mp_supported(int k, int m, int pass, int crc_mode);                                                   // bitslice_flat_mp.cu
cudaError_t launch_bsf_mp(int k, int m, int pass, int crc_mode, const BsfParams& p, int grid, cudaStream_t st, int flip);

// outputs-only CRC (mode 2) exists for the LRC local-stripe codes (same rule as bitslice.cu)
#define BSF_HAS_MODE2(KK, MM) ((MM) == 1 || ((KK) == 4 && (MM) == 3))

bool bsf_supported(int k, int m, int pass, int crc_mode) {
#define X(KK, MM) \
  if (k == KK && m == MM) return pass == 0 && (crc_mode == 1 || (crc_mode == 2 && BSF_HAS_MODE2(KK, MM)));
  CUBEEC_BS_CONFIGS(X)
#undef X
  return bsf_mp_supported(k, m, pass, crc_mode);
}

cudaError_t launch_bsf(int k, int m, int pass, int crc_mode, const BsfParams& p, int grid, cudaStream_t st, int flip) {
#define X(KK, MM)                                                                   \
  if (k == KK && m == MM) {                                                         \
    if (pass != 0) return cudaErrorInvalidValue;                                    \
    if (crc_mode == 1) return bsf_launch_one<KK, MM, 0, 1>(p, grid, st, flip);           \
    if constexpr (BSF_HAS_MODE2(KK, MM)) {                                          \
      if (crc_mode == 2) return bsf_launch_one<KK, MM, 0, 2>(p, grid, st, flip);         \
    }                                                                               \
    return cudaErrorInvalidValue;                                                   \
  }
  CUBEEC_BS_CONFIGS(X)
#undef X
  return launch_bsf_mp(k, m, pass, crc_mode, p, grid, st, flip);
}

cudaError_t launch_bsf_variant(int k, int variant, const BsfParams& p, int grid, cudaStream_t st) {
  if (k != 12) return cudaErrorInvalidValue;
  switch (variant) {
    case 512: return bsf_launch_one<12, 4, 0, 1, 512>(p, grid, st);
    case 384: return bsf_launch_one<12, 4, 0, 1, 384>(p, grid, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace cbe
