/*
 * cubeec_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that CubeFS BlobStore runs on its
 * erasure-coding + shard-checksum hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libcubeec.so) never links, loads or calls anything in oracle/.
 *
 * Reference followed (paths relative to /root/reference; RS/ =
 * vendor/github.com/klauspost/reedsolomon v1.11.7, BS/ = blobstore):
 *   GF(2^8) tables ........ RS/galois.go:13-26,28,70,855-905
 *   matrix ................ RS/matrix.go:103-118,193-266,271-282
 *   buildMatrix / New ..... RS/reedsolomon.go:220-244,413-472,568-571
 *   Encode / Verify ....... RS/reedsolomon.go:609-625,770-784,1287-1301
 *   checkShards ........... RS/reedsolomon.go:1314-1339
 *   reconstruct ........... RS/reedsolomon.go:1407-1552
 *   Split / Join sizes .... RS/reedsolomon.go:1574-1684
 *   ec.Buffer sizes ....... BS/common/ec/buf.go:67-133
 *   crc32block sizes ...... BS/common/crc32block/util.go:44-71, block.go:38-49
 *   shard CRC (IEEE) ...... BS/access/stream/stream_put.go:265-269 (Go hash/crc32)
 *
 * Parity pinning status: the reference holds NO golden parity bytes for this
 * path (its tests are round-trip tests and the vendored module ships without
 * _test.go).  What IS pinned against the reference: the GF log/exp/mul tables
 * (hashes of the literals in RS/galois.go, tests/golden/klauspost_tables.json),
 * and CRC32-IEEE against the 7 golden values in
 * BS/blobnode/core/storage/datafile_test.go.  The matrix / reconstruct logic is a
 * restatement anchored on those tables and on SURVEY.md section 8c KATs.
 */
#ifndef CUBEEC_ORACLE_H
#define CUBEEC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: numerically identical to include/cubeec.h (the product ABI). */
enum {
  ORACLE_OK = 0,
  ORACLE_ERR_INV_SHARD_NUM = 1,      /* reedsolomon.ErrInvShardNum   RS/reedsolomon.go:204 */
  ORACLE_ERR_MAX_SHARD_NUM = 2,      /* reedsolomon.ErrMaxShardNum   RS/reedsolomon.go:209 */
  ORACLE_ERR_TOO_FEW_SHARDS = 3,     /* reedsolomon.ErrTooFewShards  RS/reedsolomon.go:601 */
  ORACLE_ERR_SHARD_NO_DATA = 4,      /* reedsolomon.ErrShardNoData   RS/reedsolomon.go:1305 */
  ORACLE_ERR_SHARD_SIZE = 5,         /* reedsolomon.ErrShardSize     RS/reedsolomon.go:1309 */
  ORACLE_ERR_SHORT_DATA = 6,         /* reedsolomon.ErrShortData     RS/reedsolomon.go:1556 */
  ORACLE_ERR_RECONSTRUCT_REQUIRED = 7, /* RS/reedsolomon.go:1636 */
  ORACLE_ERR_SINGULAR = 8,           /* errSingular RS/matrix.go:184 */
  ORACLE_ERR_INVALID_ARG = 9
};

/* ---- GF(2^8), polynomial 0x11D ---------------------------------------- */
const uint8_t* oracle_gf_log_table(void);      /* 256 entries, log[0] = 0 like RS/galois.go:28 */
const uint8_t* oracle_gf_exp_table(void);      /* 510 entries like RS/galois.go:70 */
const uint8_t* oracle_gf_mul_table(void);      /* 256*256, mul[a*256+b] like RS/galois.go:83 */
uint8_t oracle_gf_mul(uint8_t a, uint8_t b);   /* galMultiply */
uint8_t oracle_gf_div(uint8_t a, uint8_t b);   /* galDivide (b != 0) */
uint8_t oracle_gf_exp(uint8_t a, int n);       /* galExp */

/* ---- matrices (row-major bytes) ----------------------------------------- */
/* (k+m) x k systematic generator, identity on top. buildMatrix. */
int oracle_build_matrix(int k, int total, uint8_t* out /* total*k */);
/* n x n inverse by the reference's Gauss-Jordan. */
int oracle_matrix_invert(const uint8_t* in, int n, uint8_t* out);

/* ---- encoder object ----------------------------------------------------- */
typedef struct oracle_rs oracle_rs_t;
int oracle_rs_new(int k, int m, oracle_rs_t** out);
void oracle_rs_free(oracle_rs_t*);
int oracle_rs_k(const oracle_rs_t*);
int oracle_rs_m(const oracle_rs_t*);
const uint8_t* oracle_rs_matrix(const oracle_rs_t*);     /* (k+m) x k */

/* shards[i] / lens[i]: Go's [][]byte.  len 0 == missing (nil or [:0]). */
int oracle_rs_encode(const oracle_rs_t*, uint8_t* const* shards, const size_t* lens, int n);
int oracle_rs_verify(const oracle_rs_t*, uint8_t* const* shards, const size_t* lens, int n, int* ok);
/* Missing shards (len 0) are written into shards[i] (caller provides room for
 * the shard size; that is the cap>=shardSize branch of RS/reedsolomon.go:1514).
 * filled[i] is set to 1 for every shard that was regenerated. */
int oracle_rs_reconstruct(const oracle_rs_t*, uint8_t* const* shards, const size_t* lens, int n,
                          int data_only, uint8_t* filled /* n or NULL */);
/* Decode rows the reference would use for the given presence pattern:
 * valid[k] = first k present indices, rows[k*k] = inverse of those generator rows. */
int oracle_rs_decode_matrix(const oracle_rs_t*, const uint8_t* present /* k+m */, int* valid, uint8_t* rows);

/* Multi-threaded SIMD encode used ONLY as the timed CPU baseline (AVX2 nibble
 * tables as RS/galois_amd64.go:35-52 describes; GFNI affine when k,m <= 10 and the
 * CPU has it, as RS/reedsolomon.go:786-796 selects).  Same results as
 * oracle_rs_encode.  threads <= 0 -> all online cores. */
int oracle_rs_encode_batch_simd(const oracle_rs_t*, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                size_t stripe_pitch, size_t n_stripes, int threads, int with_crc,
                                uint32_t* crc_out /* n_stripes*(k+m) or NULL */);
int oracle_rs_reconstruct_batch_simd(const oracle_rs_t*, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                     size_t stripe_pitch, size_t n_stripes,
                                     const uint8_t* present /* n_stripes*(k+m) */, int threads);
/* same, `repeat` passes inside one call (thread start-up amortised; used when timing) */
int oracle_rs_encode_batch_simd_rep(const oracle_rs_t*, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                    size_t stripe_pitch, size_t n_stripes, int threads, int with_crc,
                                    uint32_t* crc_out, size_t repeat);
int oracle_rs_reconstruct_batch_simd_rep(const oracle_rs_t*, uint8_t* base, size_t shard_len, size_t shard_pitch,
                                         size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int threads,
                                         size_t repeat);
const char* oracle_simd_kind(const oracle_rs_t*);   /* "gfni" | "avx2" | "scalar" */
int oracle_online_cores(void);

/* ---- sizes ---------------------------------------------------------------- */
/* reedSolomon.Split: perShard = ceil(len/k). */
size_t oracle_split_shard_size(size_t data_len, int k);
/* ec.GetBufferSizes (BS/common/ec/buf.go:67-84). Returns 0 or ORACLE_ERR_SHORT_DATA. */
int oracle_ec_buffer_sizes(size_t data_size, int n, int m, int l, size_t min_shard_size,
                           size_t* shard_size, size_t* ec_data_size, size_t* ec_size);

/* ---- CRC32 ---------------------------------------------------------------- */
enum { ORACLE_CRC_IEEE = 0, ORACLE_CRC_CASTAGNOLI = 1 };
uint32_t oracle_crc32(int poly, uint32_t crc /* running value, 0 to start */, const void* p, size_t n);
/* zlib crc32_combine semantics: crc(A||B) from crc(A), crc(B), len(B). */
uint32_t oracle_crc32_combine(int poly, uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* crc32block (BS/common/crc32block): */
int64_t oracle_crc32block_encode_size(int64_t size, int64_t block_len);
int64_t oracle_crc32block_decode_size(int64_t total, int64_t block_len);
/* Frame src[0..n) into [crc LE][payload] blocks; returns bytes written. */
int64_t oracle_crc32block_encode(const uint8_t* src, int64_t n, int64_t block_len, uint8_t* dst);
/* Verify + strip; returns payload bytes or -1 on the first mismatched block. */
int64_t oracle_crc32block_decode(const uint8_t* src, int64_t total, int64_t block_len, uint8_t* dst);
/* The on-disk image datafile.Write produces for one shard (header | framed body | footer); `out` must
 * hold oracle_shard_phys_size(size) bytes.  Returns bytes written, -1 on error. */
int64_t oracle_shard_image(uint64_t bid, uint64_t vuid, const uint8_t* data, uint32_t size, uint8_t* out, uint32_t* crc_out);
/* core.Alignphysize (BS/blobnode/core/shard.go:419-422). */
int64_t oracle_shard_phys_size(int64_t shard_size);

#ifdef __cplusplus
}
#endif
#endif
