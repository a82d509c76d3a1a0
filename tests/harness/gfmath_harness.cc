// Test harness (g++ only, no CUDA): exposes the product's HOST arithmetic (cubefs_b200/csrc/gfmath.h: GF(2^8)
// tables, generator matrix, Gauss-Jordan inverse, CRC32 polynomial constants and lookup tables) through a
// C ABI so that tests/test_host_math.py can compare it with the oracle.  Built by the test, never shipped.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../cubefs_b200/csrc/gfmath.h"

using namespace cbe;

extern "C" {
int hm_build_generator(int k, int total, uint8_t* out) {
  std::vector<uint8_t> m;
  if (!build_generator(k, total, m)) return 1;
  std::memcpy(out, m.data(), m.size());
  return 0;
}
int hm_invert(const uint8_t* in, int n, uint8_t* out) { return gf_invert(in, n, out) ? 0 : 1; }
uint8_t hm_gf_mul(uint8_t a, uint8_t b) { return gf().mul(a, b); }
uint32_t hm_crc_mul(uint32_t poly, int64_t ord, uint32_t a, uint32_t b) { return CrcPoly{poly, ord}.mul(a, b); }
uint32_t hm_crc_shift(uint32_t poly, int64_t ord, int64_t nbytes) { return CrcPoly{poly, ord}.shift_bytes_const(nbytes); }
void hm_crc_slice(uint32_t poly, uint32_t* out /* [4][256] */) {
  uint32_t t[4][256];
  crc_slice_tables(poly, t);
  std::memcpy(out, t, sizeof(t));
}
void hm_crc_constmul(uint32_t poly, int64_t ord, uint32_t constant, uint32_t* out /* [4][256] */) {
  uint32_t t[4][256];
  crc_const_mul_tables(CrcPoly{poly, ord}, constant, t);
  std::memcpy(out, t, sizeof(t));
}
}
