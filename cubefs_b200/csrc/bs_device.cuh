// bs_device.cuh -- device helpers shared by the bit-sliced kernels (bitslice.cu, bitslice_ws.cu).
#pragma once
#include <cstdint>
#include "kernels.cuh"

namespace cbe {
namespace bsdev {


__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
#ifndef CUBEEC_LDG_HINT
#define CUBEEC_LDG_HINT ".L2::256B"
#endif
  asm volatile("ld.global.nc.L1::no_allocate" CUBEEC_LDG_HINT ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
#ifndef CUBEEC_STG_HINT
#define CUBEEC_STG_HINT ".L2::evict_first"   /* parity is written once and not read again: +7 % on the plain encode kernel */
#endif
  asm volatile("st.global.L1::no_allocate" CUBEEC_STG_HINT ".v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <int IMM>
__device__ __forceinline__ uint32_t lds32_off(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
  return v;
}
// base + mult * byte B of w in ONE instruction on the FMA pipe: IDP.2A (dp2a) multiplies the two 16-bit
// halves of `m` with two bytes of `w` (.lo: bytes 0,1; .hi: bytes 2,3) and adds `base`.  m = MULT selects
// the even byte, m = MULT << 16 the odd one.  This replaces PRMT (ALU pipe) for lookup addresses:
// the kernels are bound by the ALU pipe (LOP3/SHF/PRMT) while the FMA pipe idles.
template <int B>
__device__ __forceinline__ uint32_t byte_madd(uint32_t w, uint32_t m_even, uint32_t m_odd, uint32_t base) {
  uint32_t d;
  if (B == 0) asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(m_even), "r"(w), "r"(base));
  if (B == 1) asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(m_odd), "r"(w), "r"(base));
  if (B == 2) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(m_even), "r"(w), "r"(base));
  if (B == 3) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(m_odd), "r"(w), "r"(base));
  return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// 8x8 bit transpose across 8 words, independently in each of the 4 byte lanes (an involution):
// afterwards word j holds bit j of all 32 bytes.  12 delta swaps = 24 LOP3 + 12 right shifts
// (ALU pipe) + 12 left shifts written as multiplies (FMA pipe).
// d = (a & MASK) | (b & ~MASK) in ONE LOP3 (lut 0xE4 with the mask as the immediate operand)
template <uint32_t MASK>
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "n"(MASK));
  return d;
}
// (Measured: moving the right shifts or the byte-0 address to IMAD.HI on the FMA pipe does not
// pay -- IMAD.HI issues at a quarter of the IMAD rate and lengthens the serial CRC chain.)
template <int S, uint32_t MASK>
__device__ __forceinline__ void delta_swap(uint32_t& a, uint32_t& b) {
  const uint32_t na = bitsel<MASK>(a, b * (1u << S));
  const uint32_t nb = bitsel<MASK>(a >> S, b);
  a = na;
  b = nb;
}
__device__ __forceinline__ void bit_transpose8(uint32_t (&w)[8]) {
  delta_swap<4, 0x0f0f0f0fu>(w[0], w[4]);
  delta_swap<4, 0x0f0f0f0fu>(w[1], w[5]);
  delta_swap<4, 0x0f0f0f0fu>(w[2], w[6]);
  delta_swap<4, 0x0f0f0f0fu>(w[3], w[7]);
  delta_swap<2, 0x33333333u>(w[0], w[2]);
  delta_swap<2, 0x33333333u>(w[1], w[3]);
  delta_swap<2, 0x33333333u>(w[4], w[6]);
  delta_swap<2, 0x33333333u>(w[5], w[7]);
  delta_swap<1, 0x55555555u>(w[0], w[1]);
  delta_swap<1, 0x55555555u>(w[2], w[3]);
  delta_swap<1, 0x55555555u>(w[4], w[5]);
  delta_swap<1, 0x55555555u>(w[6], w[7]);
}

__device__ __forceinline__ uint32_t gf32_mul_dev(uint32_t a, uint32_t b, uint32_t poly) {
  uint32_t r = 0;
#pragma unroll 8
  for (int i = 0; i < 32; i++) {
    r ^= a & (uint32_t)((int32_t)b >> 31);
    b <<= 1;
    a = (a >> 1) ^ (poly & (0u - (a & 1u)));
  }
  return r;
}

// u[q] = u[q] * kt for q in [Q0, Q1): the 32 successive x-multiples of the COMMON factor kt are computed
// once and shared by all registers (3 + 3*(Q1-Q0) instructions per bit instead of 5*(Q1-Q0)).
template <int Q0, int Q1, int N>
__device__ __forceinline__ void gf32_mul_common(uint32_t (&u)[N], uint32_t kt, uint32_t poly) {
  uint32_t r[N];
#pragma unroll
  for (int q = Q0; q < Q1; q++) r[q] = 0;
  uint32_t a = kt;
#pragma unroll 4
  for (int i = 0; i < 32; i++) {
#pragma unroll
    for (int q = Q0; q < Q1; q++) {
      r[q] ^= a & (uint32_t)((int32_t)u[q] >> 31);
      u[q] <<= 1;
    }
    a = (a >> 1) ^ (poly & (0u - (a & 1u)));
  }
#pragma unroll
  for (int q = Q0; q < Q1; q++) u[q] = r[q];
}

// compile-time dispatch on the shard index
template <class Net, int C, int K>
struct ApplyAt {
  static __device__ __forceinline__ void run(int c, const uint32_t (&w)[8], uint32_t (&acc)[8 * Net::M]) {
    if (c == C) Net::template apply<C>(w, acc);
    else ApplyAt<Net, C + 1, K>::run(c, w, acc);
  }
};
template <class Net, int K>
struct ApplyAt<Net, K, K> {
  static __device__ __forceinline__ void run(int, const uint32_t (&)[8], uint32_t (&)[8 * Net::M]) {}
};

// run-time dispatch on the shard index (rolled variant): apply network C, park the CRC register of shard C
// and pick up the one of shard C+1.  Only indices of the parity of C0 are compared (step 2).
template <class Net, int C, int K>
struct RollAt {
  template <int NQ>
  static __device__ __forceinline__ void run(int c, const uint32_t (&w)[8], uint32_t (&acc)[8 * Net::M], uint32_t (&crc_u)[NQ],
                                             uint32_t& u) {
    if constexpr (C < K) {
      if (c == C) {
        Net::template apply<C>(w, acc);
        crc_u[C] = u;
        u = crc_u[C + 1];
      } else {
        RollAt<Net, C + 2, K>::run(c, w, acc, crc_u, u);
      }
    }
  }
};


}  // namespace bsdev
}  // namespace cbe
