// bitslice_gen.cu -- generic bit-sliced GF(2^8) coding kernel: any inputs, any coefficients, up to 4 outputs
// per pass, flat work split.  It serves what has no compile-time XOR network: reconstruct (the decode rows
// depend on the erasure pattern, RS/reedsolomon.go:1407-1552 -> cubeec_dev_reconstruct / cubeec_reconstruct*),
// custom matrices, and verify of those.
//
// Why not tables: the table kernels (kernels.cu) do one shared-memory lookup per input byte and sit on the
// LDS pipe (profiles/r01_prof_tabk_rec.txt: 93 % of the wavefront peak, 2.1 wavefronts per lookup, 65 % of
// HBM).  Here a thread transposes 32 bytes of an input shard into 8 bit-planes (as rs_bs_kernel does) and
// multiplies by the run-time coefficient through a 256-way switch whose cases are the constant-multiplier
// networks BsMul<G> (bs_net_gen.cuh: 16 LOP3 on average, no memory traffic).  All lanes of a warp take the
// same case (coefficients are per stripe), a pattern touches n_in * n_out <= 96 of the 255 cases, so the
// code a CTA actually runs stays in the instruction cache.
//
// Work split: units of 1 KiB of every shard of one stripe (32 lanes x one 32-byte column; a warp's 256-bit
// request covers eight whole 128-byte lines) numbered stripe-major; warp g of the grid takes units
// [g*U/GW, (g+1)*U/GW): no wave quantisation, the ragged end of a shard costs one warp one masked unit.
#include <type_traits>

#include "bs_net_gen.cuh"
#include "kernels.cuh"
#include "bs_device.cuh"

namespace cbe {

using namespace bsdev;

namespace {

template <int LO, int HI>
struct MulSwitch {
  // binary dispatch tree would cost 8 compares; a dense switch compiles to one jump table (BRX)
};

__device__ __forceinline__ void mul_mac(const uint32_t g, const uint32_t (&p)[8], uint32_t (&a)[8]) {
  switch (g) {
#define C1(G) case G: BsMul<G>::mac(p, a); break;
#define C4(G) C1(G) C1(G + 1) C1(G + 2) C1(G + 3)
#define C16(G) C4(G) C4(G + 4) C4(G + 8) C4(G + 12)
#define C64(G) C16(G) C16(G + 16) C16(G + 32) C16(G + 48)
    C1(1) C1(2) C1(3) C4(4) C4(8) C4(12) C16(16) C16(32) C16(48) C64(64) C64(128) C64(192)
#undef C64
#undef C16
#undef C4
#undef C1
    default: break;   // coefficient 0: nothing to add
  }
}

constexpr int kBsgUnit = 1024;   // bytes of a shard per unit (one 32-byte column per lane)

// MODE 0: store the outputs; MODE 1: compare with the stored outputs (mismatch flag per stripe).
template <int NOUT, int MODE>
__global__ void __launch_bounds__(kBsgThreads, 1) rs_bsg_kernel(const BsgParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = kBsgThreads / 32;
  const uint64_t U = p.total_units, GW = (uint64_t)gridDim.x * NW, gw = (uint64_t)blockIdx.x * NW + warp;
  const uint64_t u_lo = gw * U / GW, u_hi = (gw + 1) * U / GW;
  if (u_lo >= u_hi) return;
  const uint32_t wt = p.units_per_shard;
  uint32_t s = (uint32_t)(u_lo / wt), t = (uint32_t)(u_lo - (uint64_t)s * wt);

  for (uint64_t u = u_lo; u < u_hi; u++) {
    const Pattern* pat = p.patterns + (p.pattern_of_stripe ? p.pattern_of_stripe[s] : 0u);
    const uint32_t n_in = pat->n_in, n_out = pat->n_out;
    if (n_out != 0) {
      uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
      const uint32_t col = t * kBsgUnit + (uint32_t)lane * 32u;
      const bool live = col < p.shard_len;
      const bool full = (t + 1) * kBsgUnit <= p.shard_len;   // warp-uniform
      uint32_t msk[8];
      if (!full) {
        const int tail = (live && col + 32 > p.shard_len) ? (int)(p.shard_len - col) : 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rem = tail - 4 * i;
          msk[i] = !live ? 0u : ((tail == 0 || rem >= 4) ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u)));
        }
      }
      uint32_t acc[NOUT][8];
#pragma unroll
      for (int r = 0; r < NOUT; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[r][i] = 0;
      // two 256-bit loads in flight ahead of the shard being coded
      uint32_t cur[8], n1[8], n2[8];
#pragma unroll
      for (int i = 0; i < 8; i++) cur[i] = n1[i] = n2[i] = 0;
      const uint8_t* src = sbase + col;
      if (live) {
        ldg256(src + (size_t)pat->in_slot[0] * p.shard_pitch, cur);
        if (n_in > 1) ldg256(src + (size_t)pat->in_slot[1] * p.shard_pitch, n1);
      }
      uint32_t coef[NOUT];
#pragma unroll
      for (int r = 0; r < NOUT; r++) coef[r] = (uint32_t)r < n_out ? pat->coef[r][0] : 0u;
#pragma unroll 1
      for (uint32_t c = 0; c < n_in; c++) {
        if (c + 2 < n_in && live) ldg256(src + (size_t)pat->in_slot[c + 2] * p.shard_pitch, n2);
        uint32_t cnext[NOUT];
#pragma unroll
        for (int r = 0; r < NOUT; r++) cnext[r] = ((uint32_t)r < n_out && c + 1 < n_in) ? pat->coef[r][c + 1] : 0u;
        if (!full) {
#pragma unroll
          for (int i = 0; i < 8; i++) cur[i] &= msk[i];
        }
        bit_transpose8(cur);
#pragma unroll
        for (int r = 0; r < NOUT; r++) mul_mac(coef[r], cur, acc[r]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          cur[i] = n1[i];
          n1[i] = n2[i];
        }
#pragma unroll
        for (int r = 0; r < NOUT; r++) coef[r] = cnext[r];
      }
      uint32_t vdiff = 0;
#pragma unroll
      for (int r = 0; r < NOUT; r++) {
        if ((uint32_t)r < n_out) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; i++) o[i] = acc[r][i];
          bit_transpose8(o);
          uint8_t* dst = sbase + (size_t)pat->out_slot[r] * p.shard_pitch + col;
          if (MODE == 0) {
            if (live) stg256(dst, o);
          } else if (live) {
            uint32_t e[8];
            ldg256(dst, e);
#pragma unroll
            for (int i = 0; i < 8; i++) vdiff |= o[i] ^ (full ? e[i] : (e[i] & msk[i]));
          }
        }
      }
      if (MODE == 1 && vdiff) p.mismatch[s] = 1;   // benign race: every writer stores the same value
    }
    if (++t == wt) {
      t = 0;
      s++;
    }
  }
}

template <int NOUT, int MODE>
cudaError_t launch_one(const BsgParams& p, int grid, cudaStream_t st) {
  rs_bsg_kernel<NOUT, MODE><<<grid, kBsgThreads, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

uint32_t bsg_units_per_shard(size_t shard_len) { return (uint32_t)((shard_len + kBsgUnit - 1) / kBsgUnit); }

cudaError_t launch_bsg(const BsgParams& p, int max_out, int mode, int grid, cudaStream_t st) {
  if (mode == 0) {
    if (max_out <= 1) return launch_one<1, 0>(p, grid, st);
    if (max_out <= 2) return launch_one<2, 0>(p, grid, st);
    if (max_out <= 4) return launch_one<4, 0>(p, grid, st);
  } else {
    if (max_out <= 4) return launch_one<4, 1>(p, grid, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace cbe
