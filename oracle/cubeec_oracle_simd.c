/*
 * cubeec_oracle_simd.c -- CPU ORACLE, multi-threaded SIMD leg (test infrastructure).
 *
 * This is the TIMED CPU BASELINE ("kind": "port"): the same arithmetic as
 * cubeec_oracle.c, vectorised the way the reference's assembly is:
 *   - AVX2: two 16-entry nibble tables per coefficient, VPSHUFB lo/hi + XOR
 *     (algorithm comment RS/galois_amd64.go:35-52, tables RS/galois.go:340,596,
 *     kernel family mulAvxTwo_* RS/galois_gen_amd64.s);
 *   - GFNI+AVX512: one 8x8 GF(2) matrix per coefficient, VGF2P8AFFINEQB
 *     (RS/galois.go:937-953, mulGFNI_* RS/galois_gen_amd64.s:4480-4537), chosen
 *     only when inputs <= 10 and outputs <= 10 as RS/reedsolomon.go:786-796 does;
 *     k > 10 (RS(12,4), RS(20,4)) runs the AVX2 plan even on GFNI CPUs (:911-916).
 *   - CRC32-IEEE: PCLMULQDQ folding like Go's hash/crc32 ieeeCLMUL on amd64.
 * ISA is selected at run time (the .so is built on one box and timed on another).
 * Stripes are distributed one per thread over all host cores ("all-cores batch"
 * figure of BASELINE.md section 3).
 */
#define _GNU_SOURCE
#include "cubeec_oracle.h"

#include <immintrin.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int oracle_online_cores(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

static int have_avx2(void) { return __builtin_cpu_supports("avx2"); }
static int have_gfni512(void) {
  return __builtin_cpu_supports("gfni") && __builtin_cpu_supports("avx512f") &&
         __builtin_cpu_supports("avx512bw");
}
static int have_pclmul(void) { return __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1"); }

static int use_gfni(int nin, int nout) { return have_gfni512() && nin <= 10 && nout <= 10; }

const char* oracle_simd_kind(const oracle_rs_t* r) {
  if (use_gfni(oracle_rs_k(r), oracle_rs_m(r))) return "gfni";
  if (have_avx2()) return "avx2";
  return "scalar";
}

/* ---- coded multiply: out[r] = XOR_c rows[r][c] * in[c] over [0, bytes) ------ */

static void code_scalar(const uint8_t* rows, int nin, int nout, const uint8_t* const* in,
                        uint8_t* const* out, size_t from, size_t to) {
  const uint8_t* mul = oracle_gf_mul_table();
  for (int r = 0; r < nout; r++)
    for (int c = 0; c < nin; c++) {
      const uint8_t* mt = mul + (size_t)rows[r * nin + c] * 256;
      const uint8_t* s = in[c];
      uint8_t* o = out[r];
      if (c == 0) for (size_t i = from; i < to; i++) o[i] = mt[s[i]];
      else        for (size_t i = from; i < to; i++) o[i] ^= mt[s[i]];
    }
}

/* AVX2 nibble-table kernel, NR (<= 4) outputs per pass, 64 bytes per iteration, accumulators in
 * registers (the shape of klauspost's mulAvxTwo_KxM_64 kernels). */
#define DEF_AVX2_PASS(NR)                                                                                     \
  __attribute__((target("avx2"))) static void avx2_pass_##NR(const __m256i* tlo, const __m256i* thi, int nin,  \
                                                             const uint8_t* const* in, uint8_t* const* out,    \
                                                             size_t vec_bytes) {                               \
    const __m256i mask = _mm256_set1_epi8(0x0f);                                                               \
    for (size_t i = 0; i < vec_bytes; i += 64) {                                                               \
      __m256i a0[NR], a1[NR];                                                                                  \
      for (int r = 0; r < NR; r++) { a0[r] = _mm256_setzero_si256(); a1[r] = _mm256_setzero_si256(); }         \
      for (int c = 0; c < nin; c++) {                                                                          \
        const __m256i d0 = _mm256_loadu_si256((const __m256i*)(in[c] + i));                                    \
        const __m256i d1 = _mm256_loadu_si256((const __m256i*)(in[c] + i + 32));                               \
        const __m256i l0 = _mm256_and_si256(d0, mask), h0 = _mm256_and_si256(_mm256_srli_epi64(d0, 4), mask);  \
        const __m256i l1 = _mm256_and_si256(d1, mask), h1 = _mm256_and_si256(_mm256_srli_epi64(d1, 4), mask);  \
        for (int r = 0; r < NR; r++) {                                                                         \
          const __m256i tl = tlo[c * 4 + r], th = thi[c * 4 + r];                                              \
          a0[r] = _mm256_xor_si256(a0[r], _mm256_xor_si256(_mm256_shuffle_epi8(tl, l0), _mm256_shuffle_epi8(th, h0))); \
          a1[r] = _mm256_xor_si256(a1[r], _mm256_xor_si256(_mm256_shuffle_epi8(tl, l1), _mm256_shuffle_epi8(th, h1))); \
        }                                                                                                      \
      }                                                                                                        \
      for (int r = 0; r < NR; r++) {                                                                           \
        _mm256_storeu_si256((__m256i*)(out[r] + i), a0[r]);                                                    \
        _mm256_storeu_si256((__m256i*)(out[r] + i + 32), a1[r]);                                               \
      }                                                                                                        \
    }                                                                                                          \
  }
DEF_AVX2_PASS(1)
DEF_AVX2_PASS(2)
DEF_AVX2_PASS(3)
DEF_AVX2_PASS(4)

__attribute__((target("avx2")))
static void code_avx2(const uint8_t* rows, int nin, int nout, const uint8_t* const* in,
                      uint8_t* const* out, size_t bytes) {
  const uint8_t* mul = oracle_gf_mul_table();
  size_t vec_bytes = bytes & ~(size_t)63;
  __m256i* tlo = (__m256i*)aligned_alloc(32, (size_t)nin * 4 * 32);
  __m256i* thi = (__m256i*)aligned_alloc(32, (size_t)nin * 4 * 32);
  for (int r0 = 0; r0 < nout; r0 += 4) {
    int nr = nout - r0 < 4 ? nout - r0 : 4;
    /* tables: [c][r][lo|hi] 16 bytes each, broadcast to both lanes */
    for (int c = 0; c < nin; c++)
      for (int r = 0; r < nr; r++) {
        uint8_t lo[16], hi[16];
        const uint8_t* mt = mul + (size_t)rows[(r0 + r) * nin + c] * 256;
        for (int v = 0; v < 16; v++) { lo[v] = mt[v]; hi[v] = mt[v << 4]; }
        tlo[c * 4 + r] = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i*)lo));
        thi[c * 4 + r] = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i*)hi));
      }
    switch (nr) {
      case 1: avx2_pass_1(tlo, thi, nin, in, out + r0, vec_bytes); break;
      case 2: avx2_pass_2(tlo, thi, nin, in, out + r0, vec_bytes); break;
      case 3: avx2_pass_3(tlo, thi, nin, in, out + r0, vec_bytes); break;
      default: avx2_pass_4(tlo, thi, nin, in, out + r0, vec_bytes); break;
    }
  }
  free(tlo);
  free(thi);
  if (vec_bytes < bytes) code_scalar(rows, nin, nout, in, out, vec_bytes, bytes);
}

/* 8x8 bit matrix for VGF2P8AFFINEQB: result bit i = parity(A.byte[7-i] & x);
 * A.byte[7-i] bit j = bit i of (c * 2^j).  Equals gf2p811dMulMatrices[c], RS/galois.go:937. */
static uint64_t gfni_matrix(uint8_t c) {
  uint64_t a = 0;
  for (int i = 0; i < 8; i++) {
    uint8_t row = 0;
    for (int j = 0; j < 8; j++)
      if ((oracle_gf_mul(c, (uint8_t)(1u << j)) >> i) & 1) row |= (uint8_t)(1u << j);
    a |= (uint64_t)row << (8 * (7 - i));
  }
  return a;
}

__attribute__((target("avx512f,avx512bw,gfni")))
static void code_gfni(const uint8_t* rows, int nin, int nout, const uint8_t* const* in,
                      uint8_t* const* out, size_t bytes) {
  size_t vec_bytes = bytes & ~(size_t)63;
  uint64_t mats[10 * 10];
  for (int r = 0; r < nout; r++)
    for (int c = 0; c < nin; c++) mats[c * nout + r] = gfni_matrix(rows[r * nin + c]);
  for (size_t i = 0; i < vec_bytes; i += 64) {
    __m512i acc[10];
    for (int c = 0; c < nin; c++) {
      __m512i d = _mm512_loadu_si512((const void*)(in[c] + i));
      for (int r = 0; r < nout; r++) {
        __m512i p = _mm512_gf2p8affine_epi64_epi8(d, _mm512_set1_epi64((long long)mats[c * nout + r]), 0);
        acc[r] = c == 0 ? p : _mm512_xor_si512(acc[r], p);
      }
    }
    for (int r = 0; r < nout; r++) _mm512_storeu_si512((void*)(out[r] + i), acc[r]);
  }
  if (vec_bytes < bytes) code_scalar(rows, nin, nout, in, out, vec_bytes, bytes);
}

static void code_best(const uint8_t* rows, int nin, int nout, const uint8_t* const* in,
                      uint8_t* const* out, size_t bytes) {
  if (nout == 0) return;
  if (use_gfni(nin, nout)) code_gfni(rows, nin, nout, in, out, bytes);
  else if (have_avx2()) code_avx2(rows, nin, nout, in, out, bytes);
  else code_scalar(rows, nin, nout, in, out, 0, bytes);
}

/* ---- CRC32-IEEE with PCLMULQDQ folding (state in = ~crc convention handled here) ---- */
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_ieee_clmul_raw(const uint8_t* buf, size_t len, uint32_t crc /* raw register */) {
  /* len >= 64 and len % 16 == 0 */
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596LL, 0x0154442bd4LL);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009eLL, 0x01751997d0LL);
  const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124LL);
  const __m128i poly = _mm_set_epi64x(0x01f7011641LL, 0x01db710641LL);
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = k1k2;
  buf += 64; len -= 64;
  while (len >= 64) {
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64; len -= 64;
  }
  x0 = k3k4;
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16; len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = k5k0;
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = poly;
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}

static uint32_t crc32_ieee_fast(const uint8_t* p, size_t n) {
  if (have_pclmul() && n >= 64) {
    size_t body = n & ~(size_t)15;
    uint32_t raw = crc32_ieee_clmul_raw(p, body, 0xFFFFFFFFu);
    uint32_t crc = ~raw;
    return oracle_crc32(ORACLE_CRC_IEEE, crc, p + body, n - body);
  }
  return oracle_crc32(ORACLE_CRC_IEEE, 0, p, n);
}

/* exported for the oracle self-test (tests compare it with the table CRC) */
uint32_t oracle_crc32_ieee_fast(const void* p, size_t n) { return crc32_ieee_fast((const uint8_t*)p, n); }

/* ---- batch drivers ------------------------------------------------------------ */
typedef struct {
  const oracle_rs_t* rs;
  uint8_t* base;
  size_t shard_len, shard_pitch, stripe_pitch, n_stripes;
  const uint8_t* present; /* reconstruct only */
  int with_crc;
  uint32_t* crc_out;
  size_t repeat; /* passes over the batch inside one call (amortises thread start-up when timing) */
  atomic_size_t next;
  atomic_int err;
} batch_t;

static void* encode_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  int k = oracle_rs_k(b->rs), m = oracle_rs_m(b->rs);
  const uint8_t* rows = oracle_rs_matrix(b->rs) + (size_t)k * k;
  for (;;) {
    size_t s = atomic_fetch_add(&b->next, 1);
    if (s >= b->n_stripes * b->repeat) break;
    s %= b->n_stripes;
    uint8_t* sp = b->base + s * b->stripe_pitch;
    const uint8_t* in[256];
    uint8_t* out[256];
    for (int c = 0; c < k; c++) in[c] = sp + (size_t)c * b->shard_pitch;
    for (int r = 0; r < m; r++) out[r] = sp + (size_t)(k + r) * b->shard_pitch;
    code_best(rows, k, m, in, out, b->shard_len);
    if (b->with_crc && b->crc_out)
      for (int i = 0; i < k + m; i++)
        b->crc_out[s * (size_t)(k + m) + i] = crc32_ieee_fast(sp + (size_t)i * b->shard_pitch, b->shard_len);
  }
  return NULL;
}

static void* reconstruct_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  int k = oracle_rs_k(b->rs), m = oracle_rs_m(b->rs), n = k + m;
  const uint8_t* gen = oracle_rs_matrix(b->rs);
  uint8_t* dec = (uint8_t*)malloc((size_t)k * k);
  uint8_t* rows = (uint8_t*)malloc((size_t)n * k);
  for (;;) {
    size_t s = atomic_fetch_add(&b->next, 1);
    if (s >= b->n_stripes * b->repeat) break;
    s %= b->n_stripes;
    uint8_t* sp = b->base + s * b->stripe_pitch;
    const uint8_t* pres = b->present + s * (size_t)n;
    int valid[256];
    int rc = oracle_rs_decode_matrix(b->rs, pres, valid, dec);
    if (rc) { atomic_store(&b->err, rc); continue; }
    const uint8_t* in[256];
    uint8_t* out[256];
    int nout = 0;
    for (int i = 0; i < k; i++) in[i] = sp + (size_t)valid[i] * b->shard_pitch;
    for (int i = 0; i < k; i++)
      if (!pres[i]) { memcpy(rows + (size_t)nout * k, dec + (size_t)i * k, (size_t)k); out[nout++] = sp + (size_t)i * b->shard_pitch; }
    code_best(rows, k, nout, in, out, b->shard_len);      /* pass 1: data, RS/reedsolomon.go:1524 */
    nout = 0;
    for (int i = 0; i < k; i++) in[i] = sp + (size_t)i * b->shard_pitch;
    for (int i = k; i < n; i++)
      if (!pres[i]) { memcpy(rows + (size_t)nout * k, gen + (size_t)i * k, (size_t)k); out[nout++] = sp + (size_t)i * b->shard_pitch; }
    code_best(rows, k, nout, in, out, b->shard_len);      /* pass 2: parity, RS/reedsolomon.go:1550 */
  }
  free(dec);
  free(rows);
  return NULL;
}

static int run_batch(batch_t* b, int threads, void* (*fn)(void*)) {
  if (threads <= 0) threads = oracle_online_cores();
  if (b->repeat == 0) b->repeat = 1;
  if ((size_t)threads > b->n_stripes * b->repeat) threads = (int)(b->n_stripes ? b->n_stripes * b->repeat : 1);
  oracle_gf_mul_table(); /* build tables before the threads race to */
  oracle_crc32(0, 0, "", 0);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int i = 1; i < threads; i++) pthread_create(&th[i], NULL, fn, b);
  fn(b);
  for (int i = 1; i < threads; i++) pthread_join(th[i], NULL);
  free(th);
  return atomic_load(&b->err);
}

int oracle_rs_encode_batch_simd(const oracle_rs_t* rs, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                size_t stripe_pitch, size_t n_stripes, int threads, int with_crc,
                                uint32_t* crc_out) {
  if (!rs || !base || shard_len == 0) return ORACLE_ERR_INVALID_ARG;
  batch_t b;
  memset(&b, 0, sizeof(b));
  b.rs = rs; b.base = base; b.shard_len = shard_len; b.shard_pitch = shard_pitch;
  b.stripe_pitch = stripe_pitch; b.n_stripes = n_stripes; b.with_crc = with_crc; b.crc_out = crc_out;
  atomic_init(&b.next, 0); atomic_init(&b.err, 0);
  return run_batch(&b, threads, encode_worker);
}

int oracle_rs_reconstruct_batch_simd(const oracle_rs_t* rs, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                     size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int threads) {
  if (!rs || !base || !present || shard_len == 0) return ORACLE_ERR_INVALID_ARG;
  batch_t b;
  memset(&b, 0, sizeof(b));
  b.rs = rs; b.base = base; b.shard_len = shard_len; b.shard_pitch = shard_pitch;
  b.stripe_pitch = stripe_pitch; b.n_stripes = n_stripes; b.present = present;
  atomic_init(&b.next, 0); atomic_init(&b.err, 0);
  return run_batch(&b, threads, reconstruct_worker);
}

/* Timing variants: `repeat` passes over the batch inside one call. */
int oracle_rs_encode_batch_simd_rep(const oracle_rs_t* rs, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                    size_t stripe_pitch, size_t n_stripes, int threads, int with_crc, uint32_t* crc_out,
                                    size_t repeat) {
  if (!rs || !base || shard_len == 0) return ORACLE_ERR_INVALID_ARG;
  batch_t b;
  memset(&b, 0, sizeof(b));
  b.rs = rs; b.base = base; b.shard_len = shard_len; b.shard_pitch = shard_pitch;
  b.stripe_pitch = stripe_pitch; b.n_stripes = n_stripes; b.with_crc = with_crc; b.crc_out = crc_out; b.repeat = repeat;
  atomic_init(&b.next, 0); atomic_init(&b.err, 0);
  return run_batch(&b, threads, encode_worker);
}
int oracle_rs_reconstruct_batch_simd_rep(const oracle_rs_t* rs, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                         size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int threads,
                                         size_t repeat) {
  if (!rs || !base || !present || shard_len == 0) return ORACLE_ERR_INVALID_ARG;
  batch_t b;
  memset(&b, 0, sizeof(b));
  b.rs = rs; b.base = base; b.shard_len = shard_len; b.shard_pitch = shard_pitch;
  b.stripe_pitch = stripe_pitch; b.n_stripes = n_stripes; b.present = present; b.repeat = repeat;
  atomic_init(&b.next, 0); atomic_init(&b.err, 0);
  return run_batch(&b, threads, reconstruct_worker);
}
