// gfmath.h -- host-side GF(2^8) and GF(2)[x]/CRC constant math for libcubeec.
//
// Product code (never includes anything from oracle/).  Builds the same field and the same
// generator matrix as klauspost/reedsolomon v1.11.7, which CubeFS constructs with
// reedsolomon.New(N, M) and no options (blobstore/common/ec/encoder.go:86,95):
//   field ......... polynomial 0x11D, generator 2   (RS/galois.go:13-26)
//   matrix ........ vandermonde(total,k) * inverse(top k x k)   (RS/reedsolomon.go:220-244)
//   decode rows ... inverse of the first k present generator rows (RS/reedsolomon.go:1453-1501)
// The inverse of a matrix is unique, so the pivoting order of RS/matrix.go:210-266 does not
// have to be replayed to be bit-exact; plain Gauss-Jordan is used.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace cbe {

struct Gf256 {
  uint8_t log[256];
  uint8_t exp[512];  // exp[i] for i < 510 valid (doubled so log a + log b needs no modulo)
  Gf256() {
    unsigned x = 1;
    std::memset(exp, 0, sizeof(exp));
    for (int i = 0; i < 255; i++) {
      exp[i] = exp[i + 255] = (uint8_t)x;
      log[x] = (uint8_t)i;
      x <<= 1;
      if (x & 0x100) x ^= 0x11D;
    }
    log[0] = 0;
  }
  uint8_t mul(uint8_t a, uint8_t b) const { return (a && b) ? exp[log[a] + log[b]] : 0; }
  uint8_t inv(uint8_t a) const { return exp[255 - log[a]]; }
  uint8_t pow(uint8_t a, int n) const {  // galExp
    if (n == 0) return 1;
    if (a == 0) return 0;
    return exp[(log[a] * n) % 255];
  }
};

inline const Gf256& gf() {
  static const Gf256 g;
  return g;
}

// n x n inverse over GF(2^8); false when singular.
inline bool gf_invert(const uint8_t* in, int n, uint8_t* out) {
  const Gf256& g = gf();
  std::vector<uint8_t> w((size_t)n * 2 * n, 0);
  const int cols = 2 * n;
  for (int r = 0; r < n; r++) {
    std::memcpy(&w[(size_t)r * cols], in + (size_t)r * n, (size_t)n);
    w[(size_t)r * cols + n + r] = 1;
  }
  for (int col = 0; col < n; col++) {
    int piv = col;
    while (piv < n && w[(size_t)piv * cols + col] == 0) piv++;
    if (piv == n) return false;
    if (piv != col)
      for (int c = 0; c < cols; c++) std::swap(w[(size_t)piv * cols + c], w[(size_t)col * cols + c]);
    uint8_t s = g.inv(w[(size_t)col * cols + col]);
    for (int c = 0; c < cols; c++) w[(size_t)col * cols + c] = g.mul(w[(size_t)col * cols + c], s);
    for (int r = 0; r < n; r++) {
      if (r == col) continue;
      uint8_t f = w[(size_t)r * cols + col];
      if (!f) continue;
      for (int c = 0; c < cols; c++) w[(size_t)r * cols + c] ^= g.mul(f, w[(size_t)col * cols + c]);
    }
  }
  for (int r = 0; r < n; r++) std::memcpy(out + (size_t)r * n, &w[(size_t)r * cols + n], (size_t)n);
  return true;
}

// (k+m) x k systematic generator of reedsolomon.New(k, m) (default options).
inline bool build_generator(int k, int total, std::vector<uint8_t>& out) {
  const Gf256& g = gf();
  std::vector<uint8_t> vm((size_t)total * k), top_inv((size_t)k * k);
  for (int r = 0; r < total; r++)
    for (int c = 0; c < k; c++) vm[(size_t)r * k + c] = g.pow((uint8_t)r, c);
  if (!gf_invert(vm.data(), k, top_inv.data())) return false;
  out.assign((size_t)total * k, 0);
  for (int r = 0; r < total; r++)
    for (int c = 0; c < k; c++) {
      uint8_t v = 0;
      for (int i = 0; i < k; i++) v ^= g.mul(vm[(size_t)r * k + i], top_inv[(size_t)i * k + c]);
      out[(size_t)r * k + c] = v;
    }
  return true;
}

// ---------------------------------------------------------------------------------------
// CRC32 constant math.  Reflected representation (Go hash/crc32, zlib): bit 31 of a word is
// the coefficient of x^0, bit 0 that of x^31.  P = reflected polynomial.
// ---------------------------------------------------------------------------------------
struct CrcPoly {
  uint32_t poly;  // 0xEDB88320 (IEEE, what BlobStore uses) or 0x82F63B78 (Castagnoli)
  // multiplicative order of x modulo the polynomial: 2^32-1 for IEEE (primitive); the Castagnoli
  // polynomial is (x+1) * (primitive of degree 31), so x has order 2^31-1 there.
  int64_t ord;
  // a(x) * b(x) mod P
  uint32_t mul(uint32_t a, uint32_t b) const {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) {
      if (b & 0x80000000u) r ^= a;
      b <<= 1;
      a = (a >> 1) ^ ((a & 1) ? poly : 0);
    }
    return r;
  }
  // x^n mod P, n may be any integer (negative exponents via the order of x).
  uint32_t xpow(int64_t n) const {
    n %= ord;
    if (n < 0) n += ord;
    return xpow_raw((uint64_t)n);
  }
  // x^n mod P for n >= 0 with no use of the order (self-check of `ord`)
  uint32_t xpow_raw(uint64_t n) const {
    uint32_t result = 0x80000000u;  // x^0
    uint32_t base = 0x40000000u;    // x^1
    while (n) {
      if (n & 1) result = mul(result, base);
      base = mul(base, base);
      n >>= 1;
    }
    return result;
  }
  // register after n zero BYTES = reg * x^(8n)
  uint32_t shift_bytes_const(int64_t nbytes) const { return xpow(8 * nbytes); }
};

// Byte-at-a-time slicing tables: slice[j][v] = register after processing byte v then j zero bytes.
inline void crc_slice_tables(uint32_t poly, uint32_t slice[4][256]) {
  for (int i = 0; i < 256; i++) {
    uint32_t c = (uint32_t)i;
    for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ poly : (c >> 1);
    slice[0][i] = c;
  }
  for (int s = 1; s < 4; s++)
    for (int i = 0; i < 256; i++) {
      uint32_t c = slice[s - 1][i];
      slice[s][i] = (c >> 8) ^ slice[0][c & 0xff];
    }
}

// mult[j][v] = (v << 8j) * constant  -- a 32-bit register times a fixed field element by 4 lookups.
inline void crc_const_mul_tables(const CrcPoly& p, uint32_t constant, uint32_t mult[4][256]) {
  for (int j = 0; j < 4; j++)
    for (int v = 0; v < 256; v++) mult[j][v] = p.mul((uint32_t)v << (8 * j), constant);
}

}  // namespace cbe
