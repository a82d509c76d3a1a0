/*
 * cubeec.h -- C-ABI of libcubeec: the B200-native Reed-Solomon + CRC32 engine that
 * drops in behind CubeFS BlobStore's erasure-coding seam.
 *
 * What it replaces (paths relative to the cubefs tree; RS/ =
 * vendor/github.com/klauspost/reedsolomon v1.11.7, BS/ = blobstore):
 *
 *   The `engine reedsolomon.Encoder` field of ec.encoder / ec.lrcEncoder
 *   (BS/common/ec/encoder.go:71-75, lrcencoder.go:28-33), i.e. the six methods CubeFS
 *   calls on it -- Encode, Verify, Reconstruct, ReconstructData (device work, below)
 *   and Split, Join (pure slicing, stay in Go).  Plus the whole-shard and
 *   per-64KiB-block CRC32-IEEE passes of the shard write/read path
 *   (BS/access/stream/stream_put.go:265-269, BS/blobnode/core/storage/datafile.go:337-342,
 *   BS/common/crc32block/block.go:38-49).
 *
 * Conventions
 *   - Plain C types only; no pointer is retained after a call returns (cgo rule).
 *   - Every function returns 0 on success or one of the CUBEEC_ERR_* codes; the Go shim
 *     maps them onto the existing error VALUES (reedsolomon.ErrTooFewShards, ...), see
 *     INTEGRATION.md.  Validation order mirrors the reference (shard count, then
 *     checkShards RS/reedsolomon.go:1314-1327).
 *   - A shard list is Go's [][]byte flattened: `shards[i]` + `lens[i]`; len 0 == missing
 *     (nil or [:0]).  The caller owns all memory; parity / regenerated shards are written in
 *     place.  For a missing shard the caller passes a buffer with room for the shard size
 *     (the Go shim allocates when cap < shardSize, as RS/reedsolomon.go:1514-1518 does).
 *   - All entry points are thread-safe; a handle is immutable after create.
 *   - There is NO CPU fallback: every compute entry point fails with CUBEEC_ERR_NO_DEVICE
 *     if no CUDA device is usable.
 */
#ifndef CUBEEC_H
#define CUBEEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  CUBEEC_OK = 0,
  CUBEEC_ERR_INV_SHARD_NUM = 1,        /* reedsolomon.ErrInvShardNum        RS/reedsolomon.go:204 */
  CUBEEC_ERR_MAX_SHARD_NUM = 2,        /* reedsolomon.ErrMaxShardNum        RS/reedsolomon.go:209 */
  CUBEEC_ERR_TOO_FEW_SHARDS = 3,       /* reedsolomon.ErrTooFewShards       RS/reedsolomon.go:601 */
  CUBEEC_ERR_SHARD_NO_DATA = 4,        /* reedsolomon.ErrShardNoData        RS/reedsolomon.go:1305 */
  CUBEEC_ERR_SHARD_SIZE = 5,           /* reedsolomon.ErrShardSize          RS/reedsolomon.go:1309 */
  CUBEEC_ERR_SHORT_DATA = 6,           /* reedsolomon.ErrShortData          RS/reedsolomon.go:1556 */
  CUBEEC_ERR_RECONSTRUCT_REQUIRED = 7, /* reedsolomon.ErrReconstructRequired RS/reedsolomon.go:1636 */
  CUBEEC_ERR_SINGULAR = 8,             /* errSingular                       RS/matrix.go:184 */
  CUBEEC_ERR_INVALID_ARG = 9,
  CUBEEC_ERR_NO_DEVICE = 10,           /* no usable CUDA device / extension: fail loudly */
  CUBEEC_ERR_CUDA = 11,                /* a CUDA call failed; see cubeec_last_error() */
  CUBEEC_ERR_UNSUPPORTED = 12          /* geometry outside what the engine supports */
};

enum { CUBEEC_CRC_IEEE = 0, CUBEEC_CRC_CASTAGNOLI = 1 };

typedef struct cubeec cubeec_t;

/* ---- process-wide ---------------------------------------------------------------- */

/* Select the CUDA devices the engine may use (default: device 0 only).  With n > 1 the
 * batched entry points partition stripes contiguously over the devices;
 * the coding matrices are the only shared state.  May be called once, before any handle
 * is created.  devices == NULL -> 0..n-1. */
int cubeec_init(const int* devices, int n_devices);
int cubeec_device_count(void);
const char* cubeec_strerror(int code);
/* Thread-local detail string of the last CUBEEC_ERR_CUDA on this thread. */
const char* cubeec_last_error(void);

/* Pinned host memory for zero-staging H2D/D2H (optional; any host pointer is accepted
 * by the host entry points, pinned ones are DMA'd directly). */
int cubeec_host_alloc(size_t bytes, void** out);
int cubeec_host_free(void* p);
int cubeec_host_register(void* p, size_t bytes);
int cubeec_host_unregister(void* p);

/* ---- handle ---------------------------------------------------------------------- */

/* reedsolomon.New(k, m) with default options (BS/common/ec/encoder.go:86,95):
 * parity_rows == NULL -> the systematic Vandermonde-derived matrix of
 * RS/reedsolomon.go:220-244 (bit-exact with klauspost); otherwise m x k row-major bytes
 * (the WithCustomMatrix case, RS/reedsolomon.go:440-456). */
int cubeec_create(int k, int m, const uint8_t* parity_rows, cubeec_t** out);
void cubeec_destroy(cubeec_t* h);
int cubeec_k(const cubeec_t* h);
int cubeec_m(const cubeec_t* h);
/* Full (k+m) x k generator, identity on top. */
int cubeec_matrix(const cubeec_t* h, uint8_t* out);
/* The decode rows the engine uses for a presence pattern: valid[k] = the first k present
 * indices, rows = k x k inverse of those generator rows (RS/reedsolomon.go:1453-1501). */
int cubeec_decode_matrix(const cubeec_t* h, const uint8_t* present /* k+m */, int* valid, uint8_t* rows);

/* ---- single stripe, host scatter pointers (the reedsolomon.Encoder subset) ---------- */

/* reedSolomon.Encode (RS/reedsolomon.go:609-625) = ec.encoder.Encode's engine call
 * (BS/common/ec/encoder.go:118).  n must be k+m.  Parity shards are overwritten, data
 * shards are only read.  crc_out (optional, k+m entries) receives CRC32 of every shard
 * computed in the same pass (replaces BS/access/stream/stream_put.go:265-269). */
int cubeec_encode(cubeec_t* h, uint8_t* const* shards, const size_t* lens, int n,
                  uint32_t* crc_out, int crc_poly);

/* Coalescing of concurrent single-stripe cubeec_encode / cubeec_reconstruct calls (the shape access uses: one blob per call
 * from up to 1000 goroutines, BS/common/ec/encoder.go:114-131, BS/access/stream/config_defaulter.go:24).  Callers
 * block in the call; the library gathers up to max_batch stripes of the same code / shard size that arrive within
 * delay_us of the first one into ONE H2D copy, ONE device encode (+ fused CRC) and ONE D2H copy, staged through
 * pinned memory that the callers fill and drain in parallel (cubeec_reconstruct without crc_out: survivors in, one
 * batched device reconstruct -- one pattern per batch runs the kernel compiled for it -- regenerated shards out).
 * Defaults: 32 stripes, 100 us.  max_batch <= 1
 * disables the queue: every call then performs its own round trip. */
int cubeec_set_coalescing(int max_batch, int delay_us);

/* reedSolomon.Verify (RS/reedsolomon.go:770-784): *ok = 1 iff parity matches. */
int cubeec_verify(cubeec_t* h, uint8_t* const* shards, const size_t* lens, int n, int* ok);

/* reedSolomon.Reconstruct / ReconstructData (RS/reedsolomon.go:1368-1388,1407-1552).
 * Missing = lens[i] == 0.  data_only != 0 leaves missing parity missing.
 * filled (optional, n bytes) is set to 1 for every shard that was regenerated.
 * crc_out (optional, n entries): CRC32 of the regenerated shards (others untouched). */
int cubeec_reconstruct(cubeec_t* h, uint8_t* const* shards, const size_t* lens, int n,
                       int data_only, uint8_t* filled, uint32_t* crc_out, int crc_poly);

/* ---- batched, host memory ------------------------------------------------------------ */

/* ec.Buffer layout (BS/common/ec/buf.go:23-35 + RS/reedsolomon.go:1618-1624): shard i of
 * stripe s lives at base + s*stripe_pitch + i*shard_len, i < k+m.  Encodes every stripe.
 * crc_out (optional): n_stripes*(k+m) whole-shard CRCs.  blockcrc_out (optional):
 * n_stripes*(k+m)*ceil(shard_len/block_payload) per-block CRCs in crc32block order
 * (block_payload = 65532 for the default 64 KiB block, BS/common/crc32block/util.go:44-46). */
int cubeec_encode_contig(cubeec_t* h, uint8_t* base, size_t shard_len, size_t n_stripes,
                         size_t stripe_pitch, uint32_t* crc_out, uint32_t* blockcrc_out,
                         size_t block_payload, int crc_poly);

/* LRC code modes in one call.  Replaces lrcEncoder.Encode (blobstore/common/ec/lrcencoder.go:35-80): the
 * global reedsolomon.Encode over the first N+M shards followed by one localEngine.Encode per AZ over
 * GetShardsInIdc (codemode.GetECLayoutByAZ, blobstore/common/codemode/codemode.go:301-318).
 *   global = handle of RS(N, M); local = handle of RS((N+M)/az_count, L/az_count)  (both as NewEncoder
 *   builds them, encoder.go:86,95)
 *   layout = the LRC ec.Buffer: shard i (< N+M+L) of stripe s at base + s*stripe_pitch + i*shard_len;
 *   data shards are read, the M global and L local parity shards are written.
 *   crc_out: NULL or n_stripes*(N+M+L) checksums (stream_put.go:265-269 computes all of them).
 * The stripe stays in HBM between the passes: N shards cross PCIe once, every shard is checksummed once.
 * CUBEEC_ERR_UNSUPPORTED when one of the codes has no generated network (custom matrices): compose
 * cubeec_encode calls as the reference does. */
int cubeec_lrc_encode_contig(cubeec_t* global, cubeec_t* local, int az_count, uint8_t* base,
                             size_t shard_len, size_t n_stripes, size_t stripe_pitch,
                             uint32_t* crc_out, int crc_poly);

/* One stripe of a batched reconstruct: the repair loop of
 * BS/blobnode/worker_slice_recover.go:822-885 (one entry per bid). */
typedef struct cubeec_stripe {
  uint8_t* const* shards;   /* k+m pointers */
  const uint8_t* present;   /* k+m flags, 0 = regenerate this shard */
  size_t shard_len;
} cubeec_stripe_t;

/* Reconstruct (+ optionally Verify, as the repair loop does at :871) a batch of stripes
 * with independent erasure patterns and lengths.  verify_ok (optional, n entries). */
int cubeec_reconstruct_batch(cubeec_t* h, const cubeec_stripe_t* stripes, size_t n_stripes,
                             int data_only, int* verify_ok);
/* The same with the checksums of the repaired shards: the repair worker re-reads every rebuilt
 * shard on the CPU to checksum it before writing it out (BS/blobnode/worker_slice_recover.go:367-373 after
 * :865-871).  crc_out: n_stripes*(k+m) entries; only the entries of shards this call regenerated are written
 * (fused into the reconstruct pass on the device), the others are left untouched. */
int cubeec_reconstruct_batch_crc(cubeec_t* h, const cubeec_stripe_t* stripes, size_t n_stripes,
                                 int data_only, int* verify_ok, uint32_t* crc_out, int crc_poly);

/* ---- batched, device-resident (what bench.py's `value` times) ------------------------- */

/* A stripe batch already in HBM: shard i of stripe s at d_base + s*stripe_pitch +
 * i*shard_pitch; d_base 16-byte aligned, shard_pitch and stripe_pitch multiples of 16 and
 * shard_pitch >= shard_len.  Output bytes from shard_len up to the next 16-byte boundary (32-byte
 * for 32-aligned layouts) are written as zeros when they fit the pitch; the rest of the pitch is left untouched.
 * d_crc_out (optional, device): n_stripes*(k+m) CRCs.  stream: a cudaStream_t (NULL = the
 * engine's own stream; the call then returns after the work completed). */
int cubeec_dev_encode(cubeec_t* h, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                      size_t stripe_pitch, size_t n_stripes, uint32_t* d_crc_out, int crc_poly,
                      void* stream);
/* Device-resident form of cubeec_lrc_encode_contig: stripes of N+M+L shards at shard_pitch (multiple of
 * 32), d_crc_out = NULL or n_stripes*(N+M+L) device words. */
int cubeec_dev_lrc_encode(cubeec_t* global, cubeec_t* local, int az_count, int device, void* d_base,
                          size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes,
                          uint32_t* d_crc_out, int crc_poly, void* stream);
/* lrcEncoder.Verify / Reconstruct / ReconstructData (BS/common/ec/lrcencoder.go:87-200) on device-resident LRC stripes
 * (N+M+L shards at shard_pitch, a multiple of 32): the global code over the first N+M shards plus one local code per AZ
 * over codemode.GetECLayoutByAZ's shard list, without leaving HBM between the codes.
 * verify: d_ok[s] = 1 iff the global parity AND every AZ's local parity of stripe s match.
 * reconstruct: present = HOST array n_stripes*(N+M+L); global shards come from the global code only (as the reference:
 * local parity never helps the global decode), missing local parity is rebuilt from its AZ afterwards; data_only =
 * ReconstructData (global data shards only).  CUBEEC_ERR_UNSUPPORTED: a code without generated network (verify). */
int cubeec_dev_lrc_verify(cubeec_t* global, cubeec_t* local, int az_count, int device, const void* d_base,
                          size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes,
                          int32_t* d_ok, void* stream);
int cubeec_dev_lrc_reconstruct(cubeec_t* global, cubeec_t* local, int az_count, int device, void* d_base,
                               size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes,
                               const uint8_t* present, int data_only, void* stream);
/* present: HOST array n_stripes*(k+m).  Regenerates every missing shard of every stripe in
 * one fused pass (missing parity is produced directly from the survivors). */
int cubeec_dev_reconstruct(cubeec_t* h, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                           size_t stripe_pitch, size_t n_stripes, const uint8_t* present,
                           int data_only, void* stream);
/* d_ok (device, n_stripes int32): 1 iff the stripe's parity matches. */
int cubeec_dev_verify(cubeec_t* h, int device, const void* d_base, size_t shard_len, size_t shard_pitch,
                      size_t stripe_pitch, size_t n_stripes, int32_t* d_ok, void* stream);

/* ---- CRC32 ------------------------------------------------------------------------------ */

/* crc32.ChecksumIEEE(p[:n]) on the GPU (host pointer). */
int cubeec_crc32(const uint8_t* p, size_t n, int crc_poly, uint32_t* out);
/* Per-block CRCs of the crc32block framing plus the whole-buffer CRC in one pass
 * (datafile.Write computes both, BS/blobnode/core/storage/datafile.go:337-342).
 * per_block: ceil(n/block_payload) entries (optional); whole: optional. */
int cubeec_crc32_blocks(const uint8_t* p, size_t n, size_t block_payload, int crc_poly,
                        uint32_t* per_block, uint32_t* whole);
/* Device-resident variant over many equal-length buffers: buffer b at d_base + b*pitch. */
int cubeec_dev_crc32(int device, const void* d_base, size_t len, size_t pitch, size_t n_buffers,
                     size_t block_payload /* 0 = whole only */, int crc_poly,
                     uint32_t* d_whole /* n_buffers or NULL */, uint32_t* d_blocks /* or NULL */,
                     void* stream);

/* ---- crc32block framing (BS/common/crc32block) --------------------------------------------- */

/* A body of n bytes is framed into blocks of block_len bytes (a multiple of 4096, default 64 KiB), each
 * [crc32 of the payload, little endian | up to block_len-4 payload bytes] (block.go:38-49,
 * sized_coder_block.go:43-103).  This is the body of a shard on disk (datafile.Write, BS/blobnode/core/storage/
 * datafile.go:337-386) and of an rpc body sent with rpc.WithCrcEncode (BS/common/rpc/client.go:54-65 ->
 * crc32block.NewBodyEncoder, request_body.go:81-130; block boundaries restart with every body).
 * EncodeSize / DecodeSize of util.go:56-71; 0 for an invalid block_len. */
size_t cubeec_crc32block_encode_size(size_t n, size_t block_len);
size_t cubeec_crc32block_decode_size(size_t total, size_t block_len);
/* Host pointers, one body: dst receives cubeec_crc32block_encode_size(n, block_len) bytes, byte-identical to
 * crc32block.NewBodyEncoder / Encoder.Encode. */
int cubeec_crc32block_encode(const uint8_t* src, size_t n, size_t block_len, uint8_t* dst, int crc_poly);
/* Verify every block and (dst != NULL) strip the framing.  *first_bad = -1 when all checksums match, else the index
 * of the first mismatching block (the reference's ErrMismatchedCrc, block.go:42-49 / decode.go:84-107). */
int cubeec_crc32block_decode(const uint8_t* framed, size_t n_framed, size_t block_len, uint8_t* dst,
                             int64_t* first_bad, int crc_poly);
/* Device-resident, many equal-length buffers (buffer b at base + b*pitch; bases and pitches 16-byte aligned):
 * frame shards that are already in HBM (e.g. right after cubeec_dev_encode), or verify framed shard images the
 * way datainspect does (BS/blobnode/datainspect.go:246-283): d_first_bad[b] = -1 or the first bad block,
 * d_block_ok (optional) = one byte per block.  d_dst == NULL: verify only. */
int cubeec_dev_crc32block_encode(int device, const void* d_src, size_t len, size_t src_pitch, size_t n_buffers,
                                 size_t block_len, void* d_dst, size_t dst_pitch, int crc_poly, void* stream);
int cubeec_dev_crc32block_decode(int device, const void* d_framed, size_t framed_len, size_t src_pitch,
                                 size_t n_buffers, size_t block_len, void* d_dst, size_t dst_pitch,
                                 int64_t* d_first_bad, uint8_t* d_block_ok, int crc_poly, void* stream);

/* ---- introspection (tests / bench) -------------------------------------------------------- */
/* Kernels launched by this process so far (all devices). */
uint64_t cubeec_kernel_launches(void);
/* Name of the kernel variant the last dev_* call on this thread dispatched to. */
const char* cubeec_last_kernel(void);
/* Measurement aid: 0 = automatic kernel choice (default), 1 = no bit-sliced kernels, 2 = bit-sliced
 * syndrome kernel for cubeec_dev_reconstruct, 3 = generic (runtime-arity) table kernel only,
 * 5 = warp-specialised fused encode+CRC kernel (rs_bsw_kernel, RS(12,4) only) instead of rs_bs_kernel<crc>,
 * 6 = rolled-loop fused encode+CRC kernel (smaller hot loop; RS(12,4), (10,4), (6,2)),
 * 7 = tile-split fused kernel rs_bs_kernel<crc> for every shard size (default: the flat-split rs_bsf_kernel from
 * 16 KiB), 1000+T = flat-split fused kernel with T threads per CTA (RS(12,4) only),
 * 3000+b = flat-split fused kernel with its measured picks flipped: bit 0 the entry point (parameters by value /
 * __grid_constant__), bit 1 the per-unit block barrier (bs_flat.cuh),
 * 4 = no run-time compiled (NVRTC) reconstruct kernels: single-pattern batches take the table kernels too,
 * 9 = flat-split bit-sliced syndrome kernel (rs_bssyn_kernel) for cubeec_dev_reconstruct,
 * 10 = first-generation stand-alone CRC kernel (crc_range_kernel) instead of crc_flat_kernel. */
void cubeec_debug_force_kernel(int which);
/* Tests: generate and NVRTC-compile (no device needed) the run-time specialised reconstruct kernel of RS(k, m) for
 * a presence pattern.  0 = compiled, 101 = NVRTC not installed, 102 = compile error (log), else CUBEEC_ERR_*. */
int cubeec_debug_jit_check(int k, int m, const uint8_t* present, int data_only, char* log, size_t log_cap);

#ifdef __cplusplus
}
#endif
#endif
