/*
 * cubeec_oracle.c -- CPU ORACLE, scalar core (test infrastructure, NOT product code).
 * See cubeec_oracle.h for scope, provenance and the parity-pinning statement.
 * Every function cites the reference lines it restates (relative to /root/reference;
 * RS/ = vendor/github.com/klauspost/reedsolomon, BS/ = blobstore).
 */
#include "cubeec_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* GF(2^8), generating polynomial 29 (x^8+x^4+x^3+x^2+1 = 0x11D)              */
/* RS/galois.go:13-26 (constants), :28 logTable, :70 expTable, :83 mulTable    */
/* ------------------------------------------------------------------------ */
static uint8_t g_log[256];
static uint8_t g_exp[510];
static uint8_t g_mul[256 * 256];
static int g_tables_ready = 0;

static void gf_init(void) {
  if (g_tables_ready) return;
  unsigned x = 1;
  for (int i = 0; i < 255; i++) {
    g_exp[i] = (uint8_t)x;
    g_exp[i + 255] = (uint8_t)x;
    g_log[x] = (uint8_t)i;
    x <<= 1;
    if (x & 0x100) x ^= 0x11D;
  }
  g_log[0] = 0; /* the reference table stores 0 for log(0) */
  for (int a = 0; a < 256; a++)
    for (int b = 0; b < 256; b++)
      g_mul[a * 256 + b] = (a == 0 || b == 0) ? 0 : g_exp[(int)g_log[a] + (int)g_log[b]];
  g_tables_ready = 1;
}

const uint8_t* oracle_gf_log_table(void) { gf_init(); return g_log; }
const uint8_t* oracle_gf_exp_table(void) { gf_init(); return g_exp; }
const uint8_t* oracle_gf_mul_table(void) { gf_init(); return g_mul; }

/* galMultiply, RS/galois.go:855-857 */
uint8_t oracle_gf_mul(uint8_t a, uint8_t b) { gf_init(); return g_mul[a * 256 + b]; }

/* galDivide, RS/galois.go:873-887 */
uint8_t oracle_gf_div(uint8_t a, uint8_t b) {
  gf_init();
  if (a == 0) return 0;
  if (b == 0) abort(); /* the reference panics */
  int r = (int)g_log[a] - (int)g_log[b];
  if (r < 0) r += 255;
  return g_exp[r];
}

/* galExp, RS/galois.go:892-905 */
uint8_t oracle_gf_exp(uint8_t a, int n) {
  gf_init();
  if (n == 0) return 1;
  if (a == 0) return 0;
  int r = (int)g_log[a] * n;
  while (r >= 255) r -= 255;
  return g_exp[r];
}

/* ------------------------------------------------------------------------ */
/* matrices                                                                   */
/* ------------------------------------------------------------------------ */

/* matrix.gaussianElimination, RS/matrix.go:210-266, on an n x cols work matrix. */
static int gaussian_elimination(uint8_t* w, int rows, int cols) {
  for (int r = 0; r < rows; r++) {
    if (w[r * cols + r] == 0) {
      for (int below = r + 1; below < rows; below++) {
        if (w[below * cols + r] != 0) {
          for (int c = 0; c < cols; c++) {
            uint8_t t = w[r * cols + c];
            w[r * cols + c] = w[below * cols + c];
            w[below * cols + c] = t;
          }
          break;
        }
      }
    }
    if (w[r * cols + r] == 0) return ORACLE_ERR_SINGULAR;
    if (w[r * cols + r] != 1) {
      uint8_t scale = oracle_gf_div(1, w[r * cols + r]);
      for (int c = 0; c < cols; c++) w[r * cols + c] = oracle_gf_mul(w[r * cols + c], scale);
    }
    for (int below = r + 1; below < rows; below++) {
      uint8_t scale = w[below * cols + r];
      if (scale != 0)
        for (int c = 0; c < cols; c++) w[below * cols + c] ^= oracle_gf_mul(scale, w[r * cols + c]);
    }
  }
  for (int d = 0; d < rows; d++) {
    for (int above = 0; above < d; above++) {
      uint8_t scale = w[above * cols + d];
      if (scale != 0)
        for (int c = 0; c < cols; c++) w[above * cols + c] ^= oracle_gf_mul(scale, w[d * cols + c]);
    }
  }
  return ORACLE_OK;
}

/* matrix.Invert, RS/matrix.go:193-208: augment with identity, eliminate, take right half. */
int oracle_matrix_invert(const uint8_t* in, int n, uint8_t* out) {
  gf_init();
  if (n <= 0) return ORACLE_ERR_INVALID_ARG;
  int cols = 2 * n;
  uint8_t* w = (uint8_t*)calloc((size_t)n * cols, 1);
  if (!w) return ORACLE_ERR_INVALID_ARG;
  for (int r = 0; r < n; r++) {
    memcpy(w + r * cols, in + r * n, (size_t)n);
    w[r * cols + n + r] = 1;
  }
  int rc = gaussian_elimination(w, n, cols);
  if (rc == ORACLE_OK)
    for (int r = 0; r < n; r++) memcpy(out + r * n, w + r * cols + n, (size_t)n);
  free(w);
  return rc;
}

/* buildMatrix, RS/reedsolomon.go:220-244: vandermonde(total,k) (RS/matrix.go:271-282,
 * entry = galExp(r, c)) times the inverse of its top k x k square. */
int oracle_build_matrix(int k, int total, uint8_t* out) {
  gf_init();
  if (k <= 0 || total < k || total > 256) return ORACLE_ERR_INVALID_ARG;
  uint8_t* vm = (uint8_t*)malloc((size_t)total * k);
  uint8_t* top_inv = (uint8_t*)malloc((size_t)k * k);
  if (!vm || !top_inv) { free(vm); free(top_inv); return ORACLE_ERR_INVALID_ARG; }
  for (int r = 0; r < total; r++)
    for (int c = 0; c < k; c++) vm[r * k + c] = oracle_gf_exp((uint8_t)r, c);
  int rc = oracle_matrix_invert(vm, k, top_inv); /* top square = first k rows */
  if (rc == ORACLE_OK) {
    /* matrix.Multiply, RS/matrix.go:103-118 */
    for (int r = 0; r < total; r++)
      for (int c = 0; c < k; c++) {
        uint8_t v = 0;
        for (int i = 0; i < k; i++) v ^= oracle_gf_mul(vm[r * k + i], top_inv[i * k + c]);
        out[r * k + c] = v;
      }
  }
  free(vm);
  free(top_inv);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* encoder                                                                    */
/* ------------------------------------------------------------------------ */
struct oracle_rs {
  int k, m, total;
  uint8_t* matrix; /* total x k; parity rows are matrix + k*k (RS/reedsolomon.go:568-571) */
};

/* reedsolomon.New, RS/reedsolomon.go:413-472 (default options only: CubeFS passes none,
 * BS/common/ec/encoder.go:86,95). */
int oracle_rs_new(int k, int m, oracle_rs_t** out) {
  gf_init();
  if (!out) return ORACLE_ERR_INVALID_ARG;
  *out = NULL;
  if (k + m > 256) return ORACLE_ERR_MAX_SHARD_NUM;
  if (k <= 0 || m < 0) return ORACLE_ERR_INV_SHARD_NUM;
  oracle_rs_t* r = (oracle_rs_t*)calloc(1, sizeof(*r));
  r->k = k; r->m = m; r->total = k + m;
  r->matrix = (uint8_t*)calloc((size_t)r->total * k, 1);
  if (m == 0) {
    for (int i = 0; i < k; i++) r->matrix[i * k + i] = 1;
  } else {
    int rc = oracle_build_matrix(k, r->total, r->matrix);
    if (rc != ORACLE_OK) { oracle_rs_free(r); return rc; }
  }
  *out = r;
  return ORACLE_OK;
}

void oracle_rs_free(oracle_rs_t* r) {
  if (!r) return;
  free(r->matrix);
  free(r);
}
int oracle_rs_k(const oracle_rs_t* r) { return r->k; }
int oracle_rs_m(const oracle_rs_t* r) { return r->m; }
const uint8_t* oracle_rs_matrix(const oracle_rs_t* r) { return r->matrix; }

/* shardSize, RS/reedsolomon.go:1332-1339: first non-zero length. */
static size_t shard_size(const size_t* lens, int n) {
  for (int i = 0; i < n; i++)
    if (lens[i] != 0) return lens[i];
  return 0;
}

/* checkShards, RS/reedsolomon.go:1314-1327 */
static int check_shards(const size_t* lens, int n, int nilok) {
  size_t size = shard_size(lens, n);
  if (size == 0) return ORACLE_ERR_SHARD_NO_DATA;
  for (int i = 0; i < n; i++)
    if (lens[i] != size && (lens[i] != 0 || !nilok)) return ORACLE_ERR_SHARD_SIZE;
  return ORACLE_OK;
}

/* codeSomeShards, RS/reedsolomon.go:807-893 reduced to its arithmetic
 * (the pure-Go tail, RS/galois_noasm.go:10-32): out[r][i] = XOR_c rows[r][c] * in[c][i]. */
static void code_some_shards(const uint8_t* const* rows, int nin, const uint8_t* const* in,
                             uint8_t* const* outp, int nout, size_t bytes) {
  for (int r = 0; r < nout; r++) {
    uint8_t* o = outp[r];
    for (int c = 0; c < nin; c++) {
      const uint8_t* mt = g_mul + (size_t)rows[r][c] * 256;
      const uint8_t* src = in[c];
      if (c == 0) for (size_t i = 0; i < bytes; i++) o[i] = mt[src[i]];
      else        for (size_t i = 0; i < bytes; i++) o[i] ^= mt[src[i]];
    }
  }
}

/* reedSolomon.Encode, RS/reedsolomon.go:609-625 */
int oracle_rs_encode(const oracle_rs_t* r, uint8_t* const* shards, const size_t* lens, int n) {
  if (n != r->total) return ORACLE_ERR_TOO_FEW_SHARDS;
  int rc = check_shards(lens, n, 0);
  if (rc) return rc;
  if (r->m == 0) return ORACLE_OK;
  const uint8_t* rows[256];
  for (int i = 0; i < r->m; i++) rows[i] = r->matrix + (size_t)(r->k + i) * r->k;
  code_some_shards(rows, r->k, (const uint8_t* const*)shards, shards + r->k, r->m, lens[0]);
  return ORACLE_OK;
}

/* reedSolomon.Verify + checkSomeShards, RS/reedsolomon.go:770-784,1287-1301 */
int oracle_rs_verify(const oracle_rs_t* r, uint8_t* const* shards, const size_t* lens, int n, int* ok) {
  *ok = 0;
  if (n != r->total) return ORACLE_ERR_TOO_FEW_SHARDS;
  int rc = check_shards(lens, n, 0);
  if (rc) return rc;
  if (r->m == 0) { *ok = 1; return ORACLE_OK; }
  size_t bytes = lens[0];
  uint8_t* tmp = (uint8_t*)malloc(bytes);
  int good = 1;
  for (int i = 0; i < r->m && good; i++) {
    const uint8_t* row = r->matrix + (size_t)(r->k + i) * r->k;
    uint8_t* outp = tmp;
    code_some_shards(&row, r->k, (const uint8_t* const*)shards, &outp, 1, bytes);
    if (memcmp(tmp, shards[r->k + i], bytes) != 0) good = 0;
  }
  free(tmp);
  *ok = good;
  return ORACLE_OK;
}

/* Row selection + inversion of reedSolomon.reconstruct, RS/reedsolomon.go:1453-1501. */
int oracle_rs_decode_matrix(const oracle_rs_t* r, const uint8_t* present, int* valid, uint8_t* rows) {
  int sub = 0;
  for (int row = 0; row < r->total && sub < r->k; row++)
    if (present[row]) valid[sub++] = row;
  if (sub < r->k) return ORACLE_ERR_TOO_FEW_SHARDS;
  uint8_t* subm = (uint8_t*)malloc((size_t)r->k * r->k);
  for (int i = 0; i < r->k; i++) memcpy(subm + (size_t)i * r->k, r->matrix + (size_t)valid[i] * r->k, (size_t)r->k);
  int rc = oracle_matrix_invert(subm, r->k, rows);
  free(subm);
  return rc;
}

/* reedSolomon.reconstruct, RS/reedsolomon.go:1407-1552 (required == nil). */
int oracle_rs_reconstruct(const oracle_rs_t* r, uint8_t* const* shards, const size_t* lens, int n,
                          int data_only, uint8_t* filled) {
  if (filled) memset(filled, 0, (size_t)(n > 0 ? n : 0));
  if (n != r->total) return ORACLE_ERR_TOO_FEW_SHARDS;
  int rc = check_shards(lens, n, 1);
  if (rc) return rc;
  size_t size = shard_size(lens, n);
  int number_present = 0, data_present = 0;
  uint8_t present[256];
  for (int i = 0; i < r->total; i++) {
    present[i] = lens[i] != 0;
    if (present[i]) { number_present++; if (i < r->k) data_present++; }
  }
  if (number_present == r->total || (data_only && data_present == r->k)) return ORACLE_OK;
  if (number_present < r->k) return ORACLE_ERR_TOO_FEW_SHARDS;

  int valid[256];
  uint8_t* dec = (uint8_t*)malloc((size_t)r->k * r->k);
  rc = oracle_rs_decode_matrix(r, present, valid, dec);
  if (rc) { free(dec); return rc; }

  const uint8_t* sub_shards[256];
  for (int i = 0; i < r->k; i++) sub_shards[i] = shards[valid[i]];

  /* missing data shards from the decode rows (RS/reedsolomon.go:1503-1524) */
  const uint8_t* rows[256];
  uint8_t* outs[256];
  int nout = 0;
  for (int i = 0; i < r->k; i++)
    if (!present[i]) { rows[nout] = dec + (size_t)i * r->k; outs[nout] = shards[i]; nout++; if (filled) filled[i] = 1; }
  code_some_shards(rows, r->k, sub_shards, outs, nout, size);
  free(dec);
  if (data_only) return ORACLE_OK;

  /* missing parity from ALL data shards (RS/reedsolomon.go:1531-1550) */
  nout = 0;
  for (int i = r->k; i < r->total; i++)
    if (!present[i]) { rows[nout] = r->matrix + (size_t)i * r->k; outs[nout] = shards[i]; nout++; if (filled) filled[i] = 1; }
  code_some_shards(rows, r->k, (const uint8_t* const*)shards, outs, nout, size);
  return ORACLE_OK;
}

/* ------------------------------------------------------------------------ */
/* sizes                                                                      */
/* ------------------------------------------------------------------------ */

/* reedSolomon.Split, RS/reedsolomon.go:1583-1584 */
size_t oracle_split_shard_size(size_t data_len, int k) { return (data_len + (size_t)k - 1) / (size_t)k; }

/* newBuffer, BS/common/ec/buf.go:67-84,120-133 */
int oracle_ec_buffer_sizes(size_t data_size, int n, int m, int l, size_t min_shard_size,
                           size_t* shard_size_out, size_t* ec_data_size, size_t* ec_size) {
  if (data_size == 0) return ORACLE_ERR_SHORT_DATA; /* isOutOfRange: dataSize <= 0 */
  if (n <= 0) return ORACLE_ERR_INVALID_ARG;        /* ErrInvalidCodeMode */
  size_t s = (data_size + (size_t)n - 1) / (size_t)n;
  if (s < min_shard_size) s = min_shard_size;
  *shard_size_out = s;
  *ec_data_size = s * (size_t)n;
  *ec_size = s * (size_t)(n + m + l);
  return ORACLE_OK;
}

/* ------------------------------------------------------------------------ */
/* CRC32 (Go hash/crc32 semantics: reflected, init/xorout 0xFFFFFFFF)          */
/* IEEE 0xEDB88320 everywhere in BlobStore (stream_put.go:268, block.go:39,   */
/* datafile.go:337); Castagnoli 0x82F63B78 offered for the north-star wording. */
/* ------------------------------------------------------------------------ */
static uint32_t g_crc_tab[2][8][256];
static int g_crc_ready = 0;

static void crc_init(void) {
  if (g_crc_ready) return;
  const uint32_t polys[2] = {0xEDB88320u, 0x82F63B78u};
  for (int p = 0; p < 2; p++) {
    for (int i = 0; i < 256; i++) {
      uint32_t c = (uint32_t)i;
      for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ polys[p] : (c >> 1);
      g_crc_tab[p][0][i] = c;
    }
    for (int s = 1; s < 8; s++)
      for (int i = 0; i < 256; i++) {
        uint32_t c = g_crc_tab[p][s - 1][i];
        g_crc_tab[p][s][i] = (c >> 8) ^ g_crc_tab[p][0][c & 0xff];
      }
  }
  g_crc_ready = 1;
}

uint32_t oracle_crc32(int poly, uint32_t crc, const void* data, size_t n) {
  crc_init();
  const uint8_t* p = (const uint8_t*)data;
  const uint32_t(*t)[256] = g_crc_tab[poly ? 1 : 0];
  uint32_t c = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = t[7][lo & 0xff] ^ t[6][(lo >> 8) & 0xff] ^ t[5][(lo >> 16) & 0xff] ^ t[4][lo >> 24] ^
        t[3][hi & 0xff] ^ t[2][(hi >> 8) & 0xff] ^ t[1][(hi >> 16) & 0xff] ^ t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ t[0][(c ^ *p++) & 0xff];
  return ~c;
}

/* GF(2) 32x32 helpers for combine (same construction as zlib's crc32_combine). */
static uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  while (vec) {
    if (vec & 1) sum ^= *mat;
    vec >>= 1;
    mat++;
  }
  return sum;
}
static void gf2_square(uint32_t* sq, const uint32_t* mat) {
  for (int n = 0; n < 32; n++) sq[n] = gf2_times(mat, mat[n]);
}

uint32_t oracle_crc32_combine(int poly, uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = poly ? 0x82F63B78u : 0xEDB88320u;
  uint32_t row = 1;
  for (int n = 1; n < 32; n++) { odd[n] = row; row <<= 1; }
  gf2_square(even, odd);
  gf2_square(odd, even);
  do {
    gf2_square(even, odd);
    if (len2 & 1) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (len2 == 0) break;
    gf2_square(odd, even);
    if (len2 & 1) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2 != 0);
  return crc1 ^ crc2;
}

/* ------------------------------------------------------------------------ */
/* crc32block framing, BS/common/crc32block                                    */
/* ------------------------------------------------------------------------ */
static int valid_block_len(int64_t bl) { return bl > 0 && (bl % 4096) == 0; } /* util.go:33-35 */

/* EncodeSize, util.go:50-57 */
int64_t oracle_crc32block_encode_size(int64_t size, int64_t block_len) {
  if (!valid_block_len(block_len)) return -1;
  int64_t payload = block_len - 4;
  int64_t cnt = (size + payload - 1) / payload;
  return size + 4 * cnt;
}
/* DecodeSize, util.go:59-65 */
int64_t oracle_crc32block_decode_size(int64_t total, int64_t block_len) {
  if (!valid_block_len(block_len)) return -1;
  int64_t cnt = (total + block_len - 1) / block_len;
  return total - 4 * cnt;
}

/* encodeBlock loop, sized_coder_block.go:43-67 + blockUnit.writeCrc block.go:45-48:
 * [crc32-IEEE(payload) little-endian][payload <= block_len-4] ... */
int64_t oracle_crc32block_encode(const uint8_t* src, int64_t n, int64_t block_len, uint8_t* dst) {
  if (!valid_block_len(block_len)) return -1;
  int64_t payload = block_len - 4, w = 0;
  while (n > 0) {
    int64_t take = n < payload ? n : payload;
    uint32_t c = oracle_crc32(ORACLE_CRC_IEEE, 0, src, (size_t)take);
    dst[w + 0] = (uint8_t)c; dst[w + 1] = (uint8_t)(c >> 8); dst[w + 2] = (uint8_t)(c >> 16); dst[w + 3] = (uint8_t)(c >> 24);
    memcpy(dst + w + 4, src, (size_t)take);
    w += 4 + take; src += take; n -= take;
  }
  return w;
}

/* decodeBlock loop, sized_coder_block.go:72-103 + blockUnit.check block.go:37-43 */
int64_t oracle_crc32block_decode(const uint8_t* src, int64_t total, int64_t block_len, uint8_t* dst) {
  if (!valid_block_len(block_len)) return -1;
  int64_t w = 0;
  while (total > 0) {
    int64_t blk = total < block_len ? total : block_len;
    if (blk <= 4) return -1;
    uint32_t want = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    uint32_t got = oracle_crc32(ORACLE_CRC_IEEE, 0, src + 4, (size_t)(blk - 4));
    if (want != got) return -1;
    if (dst) memcpy(dst + w, src + 4, (size_t)(blk - 4));
    w += blk - 4; src += blk; total -= blk;
  }
  return w;
}

/* Alignphysize, BS/blobnode/core/shard.go:419-422: header(32) + framed body + footer(8) */
int64_t oracle_shard_phys_size(int64_t shard_size) {
  return 32 + oracle_crc32block_encode_size(shard_size, 64 * 1024) + 8;
}

/* ------------------------------------------------------------------------ */
/* On-disk shard image written by blobnode (BS/blobnode/core/storage/datafile.go:304-408 with     */
/* core.Shard.WriterHeader / WriterFooter, BS/blobnode/core/shard.go:241-274):                    */
/*   header(32) = crc32(header[4:32]) BE | magic ab cd ef cc | bid BE64 | vuid BE64 | size BE32 | 0 */
/*   body       = crc32block framing of the shard data (64 KiB blocks)                             */
/*   footer(8)  = magic cc ef cd ab | crc32-IEEE(shard data) BE                                    */
/* ------------------------------------------------------------------------ */
static void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static void put_be64(uint8_t* p, uint64_t v) { put_be32(p, (uint32_t)(v >> 32)); put_be32(p + 4, (uint32_t)v); }

int64_t oracle_shard_image(uint64_t bid, uint64_t vuid, const uint8_t* data, uint32_t size, uint8_t* out, uint32_t* crc_out) {
  static const uint8_t hmagic[4] = {0xab, 0xcd, 0xef, 0xcc}, fmagic[4] = {0xcc, 0xef, 0xcd, 0xab};
  memcpy(out + 4, hmagic, 4);
  put_be64(out + 8, bid);
  put_be64(out + 16, vuid);
  put_be32(out + 24, size);
  put_be32(out + 28, 0);
  put_be32(out, oracle_crc32(ORACLE_CRC_IEEE, 0, out + 4, 28));
  int64_t body = oracle_crc32block_encode(data, size, 64 * 1024, out + 32);
  if (body < 0) return -1;
  uint32_t crc = oracle_crc32(ORACLE_CRC_IEEE, 0, data, size);
  memcpy(out + 32 + body, fmagic, 4);
  put_be32(out + 32 + body + 4, crc);
  if (crc_out) *crc_out = crc;
  return 32 + body + 8;
}
