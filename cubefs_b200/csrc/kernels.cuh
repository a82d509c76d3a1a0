// kernels.cuh -- device-side data structures and launch wrappers of libcubeec (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace cbe {

constexpr int kMaxIn = 128;   // inputs (survivor / data shards) one pass can read
constexpr int kMaxOut = 4;    // outputs one pass produces (one packed u32 table entry)
constexpr int kTabThreads = 512;
constexpr int kPiece = 16;    // bytes per thread per tile: one 128-bit access
constexpr int kTabTile = kTabThreads * kPiece;

// One coding "pattern": which shard slots are read, which are produced, and the coefficients.
// Encode has one pattern per pass; reconstruct has one per distinct erasure pattern.
struct alignas(16) Pattern {
  uint8_t n_in, n_out;
  uint8_t crc_in;   // also checksum the inputs (first pass of encode)
  uint8_t pad;
  uint8_t out_slot[kMaxOut];
  uint8_t in_slot[kMaxIn];
  uint8_t coef[kMaxOut][kMaxIn];   // coef[r][c]: out r += coef * in c
  uint8_t pad2[8];
};
static_assert(sizeof(Pattern) % 16 == 0, "Pattern must be bulk-copyable");

// Field tables staged into shared memory by cp.async.bulk (TMA 1-D).
struct alignas(16) GfDeviceTables {
  uint8_t log[256];
  uint8_t exp[512];
};

// CRC tables for a fixed thread count / tile stride.
struct alignas(16) CrcDeviceTables {
  uint32_t slice[4][256];      // slicing-by-4 byte tables
  uint32_t shift_tile[4][256]; // register * x^(8*tile bytes): Horner step between a thread's pieces
  uint32_t kthread[1024];      // x^(8*(tile - 16*(tid+1))): aligns a thread's partial to the tile end
  uint32_t poly;
  uint32_t ord;                // multiplicative order of x mod poly (for negative shifts)
  uint32_t pad[2];
};

struct TabParams {
  uint8_t* base;
  size_t stripe_pitch;
  size_t shard_pitch;
  uint32_t shard_len;
  uint32_t n_stripes;
  uint32_t n_seg;            // segments per shard
  uint32_t tiles_per_seg;    // tiles in a full segment
  uint32_t tiles_last;       // tiles in the last segment
  uint32_t n_slots;          // shard slots per stripe (k+m[+l])
  const Pattern* patterns;
  const uint32_t* pattern_of_stripe;   // nullptr -> pattern 0 for every stripe
  int mode;                  // 0 = store outputs, 1 = compare with stored outputs
  int32_t* mismatch;         // compare mode: [n_stripes], set to 1 on any difference
  uint32_t* crc_part;        // [n_stripes][n_slots][n_seg] raw segment remainders, or nullptr
  const GfDeviceTables* gf;
  const CrcDeviceTables* crc;
};

struct CrcFinalizeParams {
  const uint32_t* crc_part;  // [n_units][n_seg]
  uint32_t n_units;          // n_stripes * n_slots
  uint32_t n_slots;
  uint32_t n_seg;
  uint32_t x_full;           // x^(8*full segment bytes)
  uint32_t x_last;           // x^(8*virtual bytes of the last segment)
  uint32_t fix;              // x^(-8*zero padding after the real end)
  uint32_t init_term;        // 0xFFFFFFFF * x^(8*len)
  uint32_t poly;
  const uint8_t* slot_enable;   // [n_slots] or nullptr (all)
  uint32_t* out;             // [n_units]
};

// Smallest table replication (copies per entry, power of two <= 16) that fits.
int tab_pick_replication(int n_in, bool with_crc, int n_crc_slots, size_t smem_limit, size_t* smem_bytes);
cudaError_t launch_tab(const TabParams& p, int replication, bool with_crc, int n_crc_slots,
                       size_t smem_bytes, int grid, cudaStream_t stream);
cudaError_t launch_crc_finalize(const CrcFinalizeParams& p, cudaStream_t stream);
// fixed-arity fast path of the table kernel (no CRC); every pattern of the launch must have n_in inputs
bool tabk_supported(int n_in);
cudaError_t launch_tabk(const TabParams& p, int n_in, int grid, cudaStream_t stream);
cudaError_t tab_configure(size_t* smem_limit_out);   // sets max dynamic smem attributes

cudaError_t launch_invert_flags(int32_t* flags, size_t n, cudaStream_t stream);   // flags[i] = !flags[i]

// Stand-alone CRC over byte ranges of flat buffers (crc32 / crc32block surface).
struct CrcRangeParams {
  const uint8_t* base;       // buffer b at base + b*pitch
  size_t pitch;
  uint32_t n_buffers;
  uint32_t len;              // bytes per buffer
  uint32_t block;            // bytes per CRC unit (block payload); len if whole only
  uint32_t units_per_buffer; // ceil(len/block)
  const CrcDeviceTables* crc;
  uint32_t* out;             // [n_buffers][units_per_buffer] final CRCs of each unit
  uint32_t stride;           // distance between unit starts (0 = block: units back to back)
  uint32_t offset;           // start of unit 0 in the buffer
};
// crc_flat.cu: the same ranges, lane-private slicing tables and a flat tile split (see there)
constexpr int kCrcFlatThreads = 1024;
struct CrcFlatParams {
  const uint8_t* base;
  size_t pitch;
  uint32_t n_buffers, len, block, units_per_buffer, stride, offset;   // as CrcRangeParams
  uint32_t tiles_per_range;           // 2 KiB windows of the buffer a range can touch (uniform upper bound)
  uint64_t total_tiles;               // n_buffers * units_per_buffer * tiles_per_range
  uint32_t poly;
  uint32_t init_full;                 // 0xFFFFFFFF * x^(8 * block)
  uint32_t* out;                      // [n_buffers][units_per_buffer]; zero on entry
  const uint32_t* slice_image;        // BsfParams' tables
  const uint32_t* fold_tables;
  const uint32_t* klane;
  uint32_t x_unit_pow[24];            // x^(8 * 2048 * 2^i)
  uint32_t x_neg_pow[33];             // x^(-8 * 2^i)
};
cudaError_t launch_crc_flat(const CrcFlatParams& p, int sm_count, cudaStream_t stream);
// crc32block framing kernels (kernels.cu)
struct Crc32BlockParams {
  const uint8_t* plain;      // plain image: buffer b at plain + b*plain_pitch, plain_len bytes
  const uint8_t* framed;     // framed image: buffer b at framed + b*framed_pitch, plain_len + 4*n_blocks bytes
  size_t plain_pitch, framed_pitch;
  uint64_t plain_len;
  uint32_t n_buffers, n_blocks, block_len;
  int mode;                  // frame kernel: 0 plain -> framed, 1 framed -> plain
  const uint32_t* crcs;      // [n_buffers][n_blocks] checksums of the payloads
  uint8_t* block_ok;         // check kernel: optional [n_buffers][n_blocks]
  int64_t* first_bad;        // check kernel: optional [n_buffers], preset to LLONG_MAX-like "none" by the caller
};
cudaError_t launch_crc32block_frame(const Crc32BlockParams& p, int grid, cudaStream_t stream);
cudaError_t launch_crc32block_check(const Crc32BlockParams& p, cudaStream_t stream);
cudaError_t launch_crc_ranges(const CrcRangeParams& p, int grid, cudaStream_t stream);
// whole[b] = combine(units of buffer b); device-side Horner over the per-unit CRCs.
cudaError_t launch_crc_combine(const uint32_t* unit_crc, uint32_t n_buffers, uint32_t units_per_buffer,
                               uint32_t len, uint32_t block, uint32_t poly, uint32_t* whole,
                               cudaStream_t stream);

// ---- bit-sliced encode kernel (bitslice.cu) ----------------------------------------------
// Shared-memory image of the lane-private slicing tables (R = 32 copies):
//   table j, byte v, lane l  ->  byte offset (j>>1)*65536 + v*256 + (j&1)*128 + l*4
// placed at a 64 KiB aligned shared address so that one PRMT builds the whole LDS address:
//   addr = (hi16 of table base) | v << 8 | lane*4      (+ immediate for the table number)
constexpr int kBsThreads = 512;
constexpr int kBsGroups = 2;                                  // 32-byte groups per thread per tile
constexpr int kBsPiece = 32 * kBsGroups;                      // contiguous bytes per thread per tile
constexpr int kBsTile = kBsThreads * kBsPiece;                // bytes of every shard per tile
constexpr int kBsFoldCopies = 4;
constexpr int kBsPackedMaxStripes = 72;                       // stripes a packed tile can touch (pieces/shard >= 8)
constexpr size_t kBsSliceImageBytes = 2 * 65536;
constexpr size_t kBsSmemBytes = 65536 + kBsSliceImageBytes + 1024;
// warp-specialised fused encode+CRC kernel (bitslice_ws.cu): kBswE coder threads (XOR network, parity
// CRC) and kBswC checksum threads (data-shard CRC) per CTA; the coder threads define the tile.
#ifndef CUBEEC_BSW_E
#define CUBEEC_BSW_E 256
#define CUBEEC_BSW_C 384
#define CUBEEC_BSW_RE 176
#define CUBEEC_BSW_RC 40
#endif
constexpr int kBswE = CUBEEC_BSW_E;
constexpr int kBswC = CUBEEC_BSW_C;
constexpr int kBswTile = kBswE * kBsPiece;
// setmaxnreg budgets: kBswRegsE * kBswE + kBswRegsC * kBswC <= registers the launch allocates
constexpr int kBswRegsE = CUBEEC_BSW_RE, kBswRegsC = CUBEEC_BSW_RC;

struct BsParams {
  uint8_t* base;
  size_t stripe_pitch, shard_pitch;
  uint32_t shard_len, n_stripes;
  uint32_t n_seg, tiles_per_seg, tiles_last;
  uint32_t n_slots;
  uint32_t* crc_part;                 // [n_stripes][n_slots][n_seg] or nullptr
  int32_t* mismatch;                  // verify variant: [n_stripes], set to 1 on any parity difference
  uint8_t in_slot[24];                // shard of the stripe that is input c of the network (identity for a plain code)
  uint8_t out_slot[8];                // shard that output r of this pass is written to
  const uint32_t* slice_image;        // global: 128 KiB replicated slicing tables (kBsSliceImageBytes)
  const uint32_t* fold_tables;        // global: [4][256] register * x^(8*(tile - piece))
  const uint32_t* kthread;            // global: [2][kBsThreads] x^(8*(tile - piece*(tid+1))), and the same times x^(8*tile)
  uint32_t poly;
  uint32_t packed_pps;                // 0 = one shard spans >= 1 tile; else 64-byte pieces per shard (< kBsThreads): packed mode
};


// ---- fused encode + CRC32, flat work split (bs_flat.cuh, bitslice_flat.cu) -----------------------------
// Units (2 KiB of every shard of one stripe) are numbered stripe-major; warp g of the grid takes units
// [g*U/GW, (g+1)*U/GW).  crc_part[((stripe * n_slots + slot) * max_parts + j)] = remainder of the j-th run that
// touches the stripe (aligned to the end of its last unit there); crc_parts_finalize_kernel chains them.
// Copies of the 4 x 256-entry fold tables in shared memory (bank conflicts of the Horner step vs shared-memory size).
// Build-time knobs for A/B builds (tools/ab_fold_copies.sh); the defaults are the measured picks.  B200, fraction of the
// measured HBM peak, 16 copies / 8 copies: RS(12,4) C2 0.550 / 0.542, 383 stripes 0.518 / 0.512, RS(6,3) 0.635 / 0.627,
// RS(10,4) 0.555 / 0.552, RS(16,4) 0.515 / 0.511, RS(20,4) 0.498 / 0.493, crc_flat_kernel 0.699 / 0.691.
#ifndef CUBEEC_FC_LO
#define CUBEEC_FC_LO 16   /* fused kernel, k < 16 */
#endif
#ifndef CUBEEC_FC_HI
#define CUBEEC_FC_HI 16   /* fused kernel, k >= 16 */
#endif
#ifndef CUBEEC_FC_CRC
#define CUBEEC_FC_CRC 16  /* crc_flat_kernel */
#endif
constexpr int bsf_fold_copies(int k) { return k >= 16 ? CUBEEC_FC_HI : CUBEEC_FC_LO; }
constexpr int kCrcFlatFoldCopies = CUBEEC_FC_CRC;
#ifndef CUBEEC_BSF_THREADS
#define CUBEEC_BSF_THREADS 384
#endif
constexpr int kBsfThreads = CUBEEC_BSF_THREADS;   // 12 warps x 168 registers: room for the interleaved schedule (bs_flat.cuh)
constexpr int kBsfUnitBytes = 32 * kBsPiece;   // bytes of a shard per unit
constexpr size_t bsf_smem_bytes(int fold_copies) { return 128 + (size_t)4 * 256 * fold_copies * 4 + kBsSliceImageBytes; }
constexpr size_t kBsfSmemBytes = bsf_smem_bytes(16);   // upper bound
constexpr size_t kCrcFlatSmemBytes = bsf_smem_bytes(kCrcFlatFoldCopies);
struct BsfParams {
  uint8_t* base;
  size_t stripe_pitch, shard_pitch;
  uint32_t shard_len, n_stripes;
  uint32_t n_slots;
  uint32_t units_per_shard;           // ceil(shard_len / kBsfUnitBytes)
  uint64_t total_units;               // n_stripes * units_per_shard
  uint32_t max_parts;
  uint32_t poly;
  uint32_t sync_units;                // set by the launcher (bs_flat.cuh): block-wide barrier at the top of every unit
  uint32_t* crc_part;
  uint8_t in_slot[24];
  uint8_t out_slot[8];
  const uint32_t* slice_image;        // global: 128 KiB lane-private slicing tables
  const uint32_t* fold_tables;        // global: [4][256] register * x^(8*(unit - piece))
  const uint32_t* klane;              // global: [32] x^(8 * piece * (31 - lane))
};
struct CrcPartsFinalizeParams {
  const uint32_t* crc_part;           // [n_stripes][n_slots][max_parts]
  uint32_t n_stripes, n_slots, max_parts;
  uint32_t units_per_shard;
  uint64_t total_units, total_warps;  // U and GW of the launch that produced the parts
  uint32_t shard_len;
  uint32_t first_slot, n_out;         // slots [first_slot, first_slot + n_out) are finalized (others untouched)
  uint32_t poly;
  uint32_t x_unit_pow[24];            // x^(8 * kBsfUnitBytes * 2^i)
  uint32_t fix;                       // x^(-8 * (units_per_shard * kBsfUnitBytes - shard_len))
  uint32_t init_term;                 // 0xFFFFFFFF * x^(8 * shard_len)
  uint32_t* out;                      // [n_stripes][n_slots]
};
cudaError_t launch_crc_parts_finalize(const CrcPartsFinalizeParams& p, cudaStream_t stream);
// does bitslice_flat.cu have RS(k, m) pass `pass` of plan 0 with this CRC mode (1 all shards, 2 outputs only)?
bool bsf_supported(int k, int m, int pass, int crc_mode);
cudaError_t launch_bsf(int k, int m, int pass, int crc_mode, const BsfParams& p, int grid, cudaStream_t st, int flip = 0);
// A/B measurement aid: RS(12,4) mode 1 with another CTA size (threads in {512, 448, 384, 320, 256})
cudaError_t launch_bsf_variant(int k, int variant, const BsfParams& p, int grid, cudaStream_t st);


// ---- generic bit-sliced coding kernel (bitslice_gen.cu): run-time coefficients, flat work split -------
constexpr int kBsgThreads = 512;
struct BsgParams {
  uint8_t* base;
  size_t stripe_pitch, shard_pitch;
  uint32_t shard_len, n_stripes;
  uint32_t units_per_shard;           // bsg_units_per_shard(shard_len)
  uint64_t total_units;               // n_stripes * units_per_shard
  const Pattern* patterns;
  const uint32_t* pattern_of_stripe;  // nullptr -> pattern 0 for every stripe
  int32_t* mismatch;                  // compare mode: [n_stripes], set to 1 on any difference
};
uint32_t bsg_units_per_shard(size_t shard_len);
// max_out: largest n_out of any pattern of the launch (1..4); mode 0 = store outputs, 1 = compare with stored
cudaError_t launch_bsg(const BsgParams& p, int max_out, int mode, int grid, cudaStream_t st);

// ---- run-time compiled pattern kernels (jit.cu) --------------------------------------------------------
struct JitParams {   // layout mirrored in the generated source (jit.cu: kPrologue)
  uint8_t* base;
  uint64_t stripe_pitch, shard_pitch;
  uint32_t shard_len, n_stripes;
  uint32_t units_per_shard, pad;     // units of 1 KiB
  uint64_t total_units;
};
bool jit_available();
const void* jit_kernel(int device, const std::vector<uint8_t>& in_slots, const std::vector<uint8_t>& out_slots,
                       const std::vector<uint8_t>& rows, std::string* err);
int jit_compile_check(const std::vector<uint8_t>& in_slots, const std::vector<uint8_t>& out_slots, const std::vector<uint8_t>& rows,
                      std::string* source, std::string* err);
cudaError_t jit_launch(const void* kern, const JitParams& p, int grid, cudaStream_t st);

// ---- bit-sliced syndrome reconstruct (bitslice.cu) ---------------------------------------------
// The reference decodes from the first k present shards (RS/reedsolomon.go:1453-1465).  Data indices
// precede parity indices, so that set is: every present data shard + the first e_d present parity
// shards (e_d = missing data shards).  With the FIXED encode network, syndromes
//   S_r = P_r ^ sum_{c present} M[r][c] D_c = sum_{c missing} M[r][c] D_c        (r in those e_d parity rows)
// determine the missing data through the tiny e_d x e_d inverse; missing parity p follows as
//   P_p = T_p ^ sum_{c missing} M[p][c] D_c,   T_p = sum_{c present} M[p][c] D_c  (same network pass).
// Same k shards in, so results are identical to the reference even on inconsistent input.
struct alignas(16) RecPattern {
  uint32_t data_mask;      // bit c set: data shard c is present (read it)
  uint8_t n_syn;           // e_d: syndromes used
  uint8_t n_out;           // shards regenerated (<= 4)
  uint8_t syn_mask;        // bit r set: parity row r supplies a syndrome
  uint8_t t_mask;          // bit p set: parity row p is missing and wanted (needs T_p)
  uint8_t out_slot[4];     // shard slot of output j
  uint8_t out_prow[4];     // parity row of output j, 0xff for a data shard
  uint8_t coef[4][4];      // coef[j][i]: output j += coef * (i-th syndrome, in increasing row order)
  uint8_t pad[12];
};
static_assert(sizeof(RecPattern) == 48, "RecPattern layout");

struct BsRecParams {
  uint8_t* base;
  size_t stripe_pitch, shard_pitch;
  uint32_t shard_len, n_stripes;
  uint32_t n_seg, tiles_per_seg, tiles_last;
  const RecPattern* patterns;
  const uint32_t* pattern_of_stripe;
  const GfDeviceTables* gf;
  // flat split of the second-generation kernel (bitslice_syn.cu): units of 1 KiB of every shard of a stripe
  uint32_t units_per_shard;
  uint64_t total_units;
};

// Number of bit-sliced passes RS(k, m) with these parity rows takes: 1 (m <= 4); for the generated
// m > 4 codes ceil(m/4) with fused CRC (plan 0) or ceil(m/6) without (plan 1); 0 = no specialised
// network (the table kernels serve it).
int bs_passes(int k, int m, const uint8_t* parity_rows, int plan);
int bs_mp_passes(int k, int m, const uint8_t* parity_rows, int plan);           // bitslice_mp.cu
bool bs_mp_pass_rows(int k, int m, int plan, int pass, int* r0, int* rows);     // m > 4 codes: rows of a pass
bool bs_rec_supported(int k, int m);
bool bs_syn_supported(int k, int m);                                             // bitslice_syn.cu
uint32_t bs_syn_units_per_shard(size_t shard_len);
cudaError_t launch_bs_syn(int k, int m, const BsRecParams& p, int grid, cudaStream_t st);
cudaError_t launch_bs_rec(int k, int m, const BsRecParams& p, int grid, cudaStream_t st);
// crc: 0 none, 1 all shards (pass 0), 2 the pass's outputs only (pass > 0)
cudaError_t launch_bs(int k, int m, int pass, const BsParams& p, int crc, bool verify, int grid, cudaStream_t st);
cudaError_t launch_bs_mp(int k, int m, int pass, const BsParams& p, int crc, bool verify, int grid, cudaStream_t st);
bool bs_rolled_supported(int k, int m);                                        // experiment: rolled fused kernel
cudaError_t launch_bs_rolled(int k, int m, const BsParams& p, int grid, cudaStream_t st);
bool bsw_supported(int k, int m);                              // bitslice_ws.cu has this configuration (and its register split is safe)
cudaError_t launch_bsw(int k, int m, const BsParams& p, int grid, cudaStream_t st);

}  // namespace cbe
