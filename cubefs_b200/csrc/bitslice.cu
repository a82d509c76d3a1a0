// bitslice.cu -- the bit-sliced RS encode kernel with fused CRC32 (sm_100a).
//
// The hot kernel of BASELINE config C2 (RS(12,4) encode + CRC32 of all shards in one HBM pass).
// Replaces reedSolomon.Encode's SIMD loop (RS/reedsolomon.go:609-625,897-1134; kernels
// RS/galois_gen_amd64.s) and the per-shard crc32.ChecksumIEEE pass of
// blobstore/access/stream/stream_put.go:265-269.
//
// Why bit-sliced: B200 has no byte-shuffle (PSHUFB) or GF affine instruction, and a table lookup
// per input byte saturates the shared-memory pipe at about half of HBM speed (kernels.cu).  Here a
// thread turns 32 bytes of a shard (one 256-bit load) into 8 bit-planes with a SWAR 8x8 bit
// transpose (12 delta swaps), and multiplication by the (compile-time) matrix coefficients becomes
// a fixed XOR network on planes: 3-input XORs = one LOP3 each, no memory traffic.  The CRC runs
// on the same registers before the transpose: slicing-by-4 over the 8 words with lane-private
// (conflict-free) copies of the tables in shared memory, one PRMT to form each lookup address.
// Per thread the CRC registers advance Horner-style over the tiles; partial remainders are
// aligned and XOR-reduced once per work item and finished by crc_finalize_kernel.
#include <type_traits>

#include "bs_net_gen.cuh"
#include "kernels.cuh"
#include "bs_device.cuh"

namespace cbe {

using namespace bsdev;

// VERIFY (reedSolomon.Verify / checkSomeShards, RS/reedsolomon.go:770-784,1287-1301): the computed
// parity is compared with the stored parity instead of being written; any difference raises the
// stripe's flag in p.mismatch.  The reference allocates m scratch shards and runs bytes.Equal.
template <int K, int M, bool CRC, bool PACKED, bool VERIFY = false>
__global__ void __launch_bounds__(kBsThreads, 1) rs_bs_kernel(const BsParams p) {
  static_assert(!(CRC && VERIFY), "verify does not checksum");
  using Net = BsNet<K, M>;
  constexpr int NT = kBsThreads, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- shared memory: [misc: mbarrier | fold tables (4 copies) | kthread | reduction] ... [64K-aligned slice image]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* fold_s = reinterpret_cast<uint32_t*>(smem + 64);                       // [4][256][kBsFoldCopies]
  uint32_t* kth_s = fold_s + 4 * 256 * kBsFoldCopies;                              // [2][NT]: [1] = [0] * x^(8*tile)
  uint32_t* red_s = kth_s + 2 * NT;                                                    // [(K+M)][NW]
  uint32_t* red2_s = red_s + (K + M) * NW;                                         // packed mode: [kBsPackedMaxStripes][(K+M)]
  const uint32_t base_addr = smem_addr(smem);
  uint32_t tab_addr = 0;   // shared address of the slice image
  if (CRC) {
    tab_addr = (base_addr + (uint32_t)(64 + 4 * 256 * kBsFoldCopies * 4 + 2 * NT * 4 + (K + M) * NW * 4 + kBsPackedMaxStripes * (K + M) * 4) + 65535u) & ~65535u;
    uint8_t* tab_ptr = smem + (tab_addr - base_addr);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      // TMA 1-D bulk copies: the replicated slicing tables, two 64 KiB halves
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                   "r"((uint32_t)kBsSliceImageBytes)
                   : "memory");
      for (int h = 0; h < 2; h++)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_addr(tab_ptr + h * 65536)),
            "l"(reinterpret_cast<const uint8_t*>(p.slice_image) + h * 65536), "r"(65536u), "r"(smem_addr(bar))
            : "memory");
    }
    for (int i = tid; i < 4 * 256; i += NT) {
      const uint32_t v = p.fold_tables[i];
#pragma unroll
      for (int q = 0; q < kBsFoldCopies; q++) fold_s[i * kBsFoldCopies + q] = v;
    }
    kth_s[tid] = p.kthread[tid];
    kth_s[NT + tid] = p.kthread[NT + tid];
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_addr(bar))
          : "memory");
    }
    __syncthreads();
  }
  // lane-private lookup base: byte0 = lane*4, byte2 = bits 16..23 of the table address
  const uint32_t lane_base = tab_addr | (uint32_t)(lane * 4);
  const uint32_t fold_lane = smem_addr(fold_s) + (uint32_t)((lane & (kBsFoldCopies - 1)) * 4);

  // one slicing-by-4 step: register after absorbing the 4 bytes of y (= state ^ data word)
  auto slice4 = [&](uint32_t y) -> uint32_t {
    // lookup address = table base + lane*4 + byte * 256: one IDP.2A per byte (FMA pipe)
    const uint32_t a0 = byte_madd<0>(y, 256u, 256u << 16, lane_base);
    const uint32_t a1 = byte_madd<1>(y, 256u, 256u << 16, lane_base);
    const uint32_t a2 = byte_madd<2>(y, 256u, 256u << 16, lane_base);
    const uint32_t a3 = byte_madd<3>(y, 256u, 256u << 16, lane_base);
    // byte0 -> table 3, byte1 -> table 2, byte2 -> table 1, byte3 -> table 0
    const uint32_t t3 = lds32_off<65536 + 128>(a0);
    const uint32_t t2 = lds32_off<65536>(a1);
    const uint32_t t1 = lds32_off<128>(a2);
    const uint32_t t0 = lds32_off<0>(a3);
    return t3 ^ t2 ^ t1 ^ t0;
  };
  auto fold = [&](uint32_t u) -> uint32_t {
    const uint32_t* f = fold_s;
    (void)f;
    constexpr uint32_t ST = kBsFoldCopies * 4;   // bytes per table entry group
    return lds32(byte_madd<0>(u, ST, ST << 16, fold_lane + 0 * 256 * ST)) ^ lds32(byte_madd<1>(u, ST, ST << 16, fold_lane + 1 * 256 * ST)) ^
           lds32(byte_madd<2>(u, ST, ST << 16, fold_lane + 2 * 256 * ST)) ^ lds32(byte_madd<3>(u, ST, ST << 16, fold_lane + 3 * 256 * ST));
  };

  uint32_t crc_u[K + M];
#pragma unroll
  for (int i = 0; i < K + M; i++) crc_u[i] = 0;
  uint32_t vdiff = 0;   // VERIFY: OR of (computed ^ stored) parity words of this thread

  // (an explicit prefetch.global.L2 of the next column group was measured: 30 % slower -- not used)
  // one 32-byte group of every shard.  FULL = the whole tile lies inside [0, shard_len): no
  // predicates, no tail masks (every tile but the last of a shard).
  auto group = [&](auto full_tag, uint8_t* sbase, const size_t col) {
    constexpr bool FULL = decltype(full_tag)::value;
    const bool live = FULL || col < p.shard_len;
    const int tail = (!FULL && live && col + 32 > p.shard_len) ? (int)(p.shard_len - col) : 0;
    uint32_t msk[8];
    if (!FULL) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int rem = tail - 4 * i;
        msk[i] = (tail == 0 || rem >= 4) ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u));
      }
    }
    uint32_t acc[8 * M];
#pragma unroll
    for (int i = 0; i < 8 * M; i++) acc[i] = 0;
    // Look-ahead ring: DEPTH 256-bit loads per thread in flight ahead of the shard being coded.
    // The fused-CRC variant is ALU/issue bound and out of registers (DEPTH 1); the plain variant is
    // latency bound and uses its spare registers for a deep ring.
    constexpr int DEPTH = CRC ? 1 : 4;
    uint32_t ring[DEPTH + 1][8];
#pragma unroll
    for (int b = 0; b <= DEPTH; b++)
#pragma unroll
      for (int i = 0; i < 8; i++) ring[b][i] = 0;
    const uint8_t* src = sbase + col;
    auto shard = [&](const int c, uint32_t (&w)[8]) {
      if (!FULL) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] &= msk[i];
      }
      if (CRC) {
        uint32_t u = crc_u[c];
#pragma unroll
        for (int i = 0; i < 8; i++) u = slice4(u ^ w[i]);
        crc_u[c] = u;
      }
      bit_transpose8(w);
      ApplyAt<Net, 0, K>::run(c, w, acc);   // c is a compile-time constant after unrolling
    };
#pragma unroll
    for (int c = 0; c < DEPTH && c < K; c++)
      if (live) ldg256(src + (size_t)c * p.shard_pitch, ring[c % (DEPTH + 1)]);
#pragma unroll
    for (int c = 0; c < K; c++) {
      if (c + DEPTH < K && live) ldg256(src + (size_t)(c + DEPTH) * p.shard_pitch, ring[(c + DEPTH) % (DEPTH + 1)]);
      shard(c, ring[c % (DEPTH + 1)]);
    }
#pragma unroll
    for (int r = 0; r < M; r++) {
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = acc[r * 8 + i];
      bit_transpose8(o);
      if (VERIFY) {
        if (live) {
          uint32_t e[8];
          ldg256(sbase + (size_t)(K + r) * p.shard_pitch + col, e);
#pragma unroll
          for (int i = 0; i < 8; i++) vdiff |= o[i] ^ (FULL ? e[i] : (e[i] & msk[i]));
        }
      } else if (live) {
        stg256(sbase + (size_t)(K + r) * p.shard_pitch + col, o);
      }
      if (CRC) {
        uint32_t u = crc_u[K + r];
#pragma unroll
        for (int i = 0; i < 8; i++) u = slice4(u ^ o[i]);
        crc_u[K + r] = u;
      }
    }
  };


  if constexpr (PACKED) {
    // ---- packed mode: shards shorter than a tile.  The (stripe, 64-byte piece) pairs of the whole
    // batch are laid end to end and a tile takes 512 of them, so small shards (2 KiB is the
    // production minimum) still fill the CTA.  One piece per thread: no Horner step; the per-thread
    // CRC remainder is aligned to the end of its own shard and XOR-reduced per stripe.
    const uint32_t PPS = p.packed_pps;
    const uint64_t F = (uint64_t)p.n_stripes * PPS;
    const uint32_t n_tiles = (uint32_t)((F + NT - 1) / NT);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const uint64_t f = (uint64_t)tile * NT + tid;
      const bool active = f < F;
      const uint32_t stripe = active ? (uint32_t)(f / PPS) : 0xFFFFFFFFu;
      const uint32_t pos = active ? (uint32_t)(f - (uint64_t)stripe * PPS) : 0u;
      if constexpr (!CRC) {
        // no Horner register to keep contiguous: in group g lane L takes 32-byte half-piece g*32+L of
        // the warp's 32 pieces, so one 256-bit request of a warp is a contiguous 1 KiB (whole lines)
#pragma unroll 1
        for (int g = 0; g < kBsGroups; g++) {
          const uint32_t hp = (uint32_t)g * 32u + (uint32_t)lane;
          const uint64_t fh = (uint64_t)tile * NT + (uint64_t)warp * 32u + (hp >> 1);
          if (fh < F) {
            const uint32_t st = (uint32_t)(fh / PPS);
            const uint32_t ps = (uint32_t)(fh - (uint64_t)st * PPS);
            group(std::false_type{}, p.base + (size_t)st * p.stripe_pitch, (size_t)ps * kBsPiece + (size_t)(hp & 1u) * 32);
            if (VERIFY && vdiff) p.mismatch[st] = 1;
            vdiff = 0;
          }
        }
      } else if (active) {
        uint8_t* sb = p.base + (size_t)stripe * p.stripe_pitch;
#pragma unroll 1
        for (int g = 0; g < kBsGroups; g++) group(std::false_type{}, sb, (size_t)pos * kBsPiece + (size_t)g * 32);
      }
      if (CRC && p.crc_part) {
        const uint32_t stripe0 = (uint32_t)(((uint64_t)tile * NT) / PPS);
        uint64_t last_f = (uint64_t)(tile + 1) * NT - 1;
        if (last_f >= F) last_f = F - 1;
        const uint32_t nstr = (uint32_t)(last_f / PPS) - stripe0 + 1;
        for (uint32_t i = tid; i < nstr * (K + M); i += NT) red2_s[i] = 0;
        __syncthreads();
        const uint32_t kt = active ? kth_s[pos + NT - PPS] : 0u;
        const uint32_t peers = __match_any_sync(0xffffffffu, stripe);
        const bool leader = active && (lane == __ffs(peers) - 1);
#pragma unroll
        for (int q = 0; q < K + M; q++) {
          uint32_t u = active ? gf32_mul_dev(crc_u[q], kt, p.poly) : 0u;
          crc_u[q] = 0;
          u = __reduce_xor_sync(peers, u);
          if (leader) atomicXor(&red2_s[(stripe - stripe0) * (K + M) + q], u);
        }
        __syncthreads();
        for (uint32_t i = tid; i < nstr * (K + M); i += NT) {
          const uint32_t sj = stripe0 + i / (K + M), q = i % (K + M);
          const uint32_t segidx = tile - (uint32_t)(((uint64_t)sj * PPS) / NT);   // 0 or 1: a shard spans at most two tiles
          p.crc_part[((size_t)sj * p.n_slots + q) * 2 + segidx] = red2_s[i];
        }
        __syncthreads();
      }
    }
  } else {

  const uint32_t n_items = p.n_stripes * p.n_seg;
  const size_t seg_bytes = (size_t)p.tiles_per_seg * kBsTile;

  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t s = item / p.n_seg, seg = item - s * p.n_seg;
    uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
    const uint32_t T = (seg == p.n_seg - 1) ? p.tiles_last : p.tiles_per_seg;
    const size_t seg_start = (size_t)seg * seg_bytes;

    bool skipped_tile = false;   // warp-uniform: this warp's 2 KiB of the shard's final tile lie beyond shard_len
    for (uint32_t t = 0; t < T; t++) {
      const size_t tile_start = seg_start + (size_t)t * kBsTile;
      // Without CRC, FULL / ragged / nothing is decided per WARP (2 KiB of every shard), not per tile: in
      // the final tile of a shard the warps inside run the unmasked path and the warps past the end do
      // nothing.  (A CRC register that sits a tile out is realigned through kthread[NT + tid].)
      // (fused-CRC variant: measured 1.6 % SLOWER with the per-warp decision -- it keeps the per-tile one)
      const size_t warp_lo = tile_start + (size_t)warp * (32 * kBsPiece);
      const bool warp_full = CRC ? tile_start + kBsTile <= p.shard_len : warp_lo + 32 * kBsPiece <= p.shard_len;
      if (!CRC && warp_lo >= p.shard_len) {
        skipped_tile = true;
        continue;
      }
      if (CRC) {
        // Horner step: skip the gap between the end of this thread's previous piece and this one
#pragma unroll
        for (int i = 0; i < K + M; i++) crc_u[i] = fold(crc_u[i]);
      }
      // CRC: a thread owns 64 contiguous bytes per tile (its Horner register needs contiguity), so a
      // warp's 256-bit load covers every other 32-byte sector of 2 KiB.  Without CRC the two groups of a
      // warp are two contiguous 1 KiB runs: whole 128-byte lines per request.
      const size_t col0 = CRC ? tile_start + (size_t)tid * kBsPiece : warp_lo + (size_t)lane * 32;
      constexpr size_t GSTRIDE = CRC ? 32 : 1024;
      if (warp_full) {
#pragma unroll 1
        for (int g = 0; g < kBsGroups; g++) group(std::true_type{}, sbase, col0 + (size_t)g * GSTRIDE);
      } else {
#pragma unroll 1
        for (int g = 0; g < kBsGroups; g++) group(std::false_type{}, sbase, col0 + (size_t)g * GSTRIDE);
      }
    }

    if (VERIFY) {
      if (vdiff) p.mismatch[s] = 1;   // benign race: every writer stores the same value
      vdiff = 0;
    }
    if (CRC && p.crc_part) {
      const uint32_t kt = kth_s[skipped_tile ? NT + tid : tid];
#pragma unroll
      for (int q = 0; q < K + M; q++) {
        uint32_t u = gf32_mul_dev(crc_u[q], kt, p.poly);
        crc_u[q] = 0;
        u = __reduce_xor_sync(0xffffffffu, u);
        if (lane == 0) red_s[q * NW + warp] = u;
      }
      __syncthreads();
      if (tid < K + M) {
        uint32_t u = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) u ^= red_s[tid * NW + w2];
        p.crc_part[((size_t)s * p.n_slots + tid) * p.n_seg + seg] = u;
      }
      __syncthreads();
    }
  }
  }   // !PACKED
}


// ------------------------------------------------------------------------------------------
// Bit-sliced syndrome reconstruct (see RecPattern in kernels.cuh).  Per 32-byte column:
//   present data shards  -> transpose -> fixed XOR network (all M rows at once)
//   syndrome parity rows -> transpose -> XOR into their accumulator rows
//   needed rows (syndromes, T_p) transposed back to bytes
//   outputs = per-pattern 4x4 coefficient product of the syndromes through packed shared-memory
//   tables (one lookup per syndrome byte), T_p XORed into regenerated parity.
// ------------------------------------------------------------------------------------------
constexpr int kRecCopies = 16;                       // table copies (copy = lane % 16)
constexpr size_t kRecSmemBytes = 65536 + 1024 + 1024;

template <int K, int M>
__global__ void __launch_bounds__(kBsThreads, 1) rs_bsrec_kernel(const BsRecParams p) {
  using Net = BsNet<K, M>;
  constexpr int NT = kBsThreads;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tab = smem;                                              // 256 rows x (4 syndromes x 16 copies) u32 = 64 KiB
  GfDeviceTables* gf_s = reinterpret_cast<GfDeviceTables*>(smem + 65536);
  RecPattern* pat_s = reinterpret_cast<RecPattern*>(smem + 65536 + 1024);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t tab_lane = smem_addr(tab) + (uint32_t)((lane & (kRecCopies - 1)) * 4);

  for (int i = tid; i < (int)(sizeof(GfDeviceTables) / 4); i += NT)
    reinterpret_cast<uint32_t*>(gf_s)[i] = reinterpret_cast<const uint32_t*>(p.gf)[i];

  uint32_t cur_pattern = 0xFFFFFFFFu;
  const uint32_t n_items = p.n_stripes * p.n_seg;
  const size_t seg_bytes = (size_t)p.tiles_per_seg * kBsTile;

  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t s = item / p.n_seg, seg = item - s * p.n_seg;
    const uint32_t pat_id = p.pattern_of_stripe ? p.pattern_of_stripe[s] : 0u;
    if (pat_id != cur_pattern) {
      __syncthreads();
      cur_pattern = pat_id;
      if (tid < (int)(sizeof(RecPattern) / 4)) reinterpret_cast<uint32_t*>(pat_s)[tid] = reinterpret_cast<const uint32_t*>(p.patterns + pat_id)[tid];
      __syncthreads();
      // packed product tables: entry(i, v) = sum_j (coef[j][i] * v) << 8j
      for (int idx = tid; idx < 4 * 256; idx += NT) {
        const int i = idx >> 8, v = idx & 255;
        uint32_t e = 0;
        if (v && i < pat_s->n_syn) {
          const int lv = gf_s->log[v];
          for (int j = 0; j < pat_s->n_out; j++) {
            const int co = pat_s->coef[j][i];
            if (co) e |= (uint32_t)gf_s->exp[gf_s->log[co] + lv] << (8 * j);
          }
        }
        uint32_t* row = reinterpret_cast<uint32_t*>(tab + (size_t)v * 256 + (size_t)i * (4 * kRecCopies));
#pragma unroll
        for (int q = 0; q < kRecCopies; q++) row[q] = e;
      }
      __syncthreads();
    }
    const uint32_t data_mask = pat_s->data_mask;
    const uint32_t syn_mask = pat_s->syn_mask, t_mask = pat_s->t_mask;
    const int n_out = pat_s->n_out;
    if (n_out == 0) continue;
    uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
    const uint32_t T = (seg == p.n_seg - 1) ? p.tiles_last : p.tiles_per_seg;
    const size_t seg_start = (size_t)seg * seg_bytes;

    // one 32-byte column group; FULL = whole tile inside the shard (no predicates / tail masks)
    auto group = [&](auto full_tag, const size_t col) {
      constexpr bool FULL = decltype(full_tag)::value;
      const bool live = FULL || col < p.shard_len;
      const int tail = (!FULL && live && col + 32 > p.shard_len) ? (int)(p.shard_len - col) : 0;
      uint32_t acc[8 * M];
#pragma unroll
      for (int i = 0; i < 8 * M; i++) acc[i] = 0;
      constexpr int DEPTH = 3;
      uint32_t ring[DEPTH + 1][8];
#pragma unroll
      for (int bq = 0; bq <= DEPTH; bq++)
#pragma unroll
        for (int i = 0; i < 8; i++) ring[bq][i] = 0;
      const uint8_t* src = sbase + col;
      // inputs 0..K-1 are the data shards, K..K+M-1 the parity rows that supply syndromes; all go
      // through one look-ahead ring (DEPTH 256-bit loads in flight ahead of the one being consumed)
      auto wanted = [&](const int c) -> bool {
        return c < K ? ((data_mask >> c) & 1u) != 0 : (c < K + M && ((syn_mask >> (c - K)) & 1u) != 0);
      };
      auto fetch = [&](const int c, uint32_t (&w)[8]) {
        if (c < K + M && live && wanted(c)) ldg256(src + (size_t)c * p.shard_pitch, w);
      };
      auto consume = [&](const int c, uint32_t (&w)[8]) {
        if (wanted(c)) {   // uniform over the stripe
          if (!FULL && tail) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int rem = tail - 4 * i;
              if (rem <= 0) w[i] = 0;
              else if (rem < 4) w[i] &= (1u << (8 * rem)) - 1u;
            }
          }
          bit_transpose8(w);
          if (c < K) {
            ApplyAt<Net, 0, K>::run(c, w, acc);
          } else {
            // S_r = P_r ^ (network row r)
#pragma unroll
            for (int i = 0; i < 8; i++) acc[(c - K) * 8 + i] ^= w[i];
          }
        }
      };
#pragma unroll
      for (int c = 0; c < DEPTH; c++) fetch(c, ring[c % (DEPTH + 1)]);
#pragma unroll
      for (int c = 0; c < K + M; c++) {
        fetch(c + DEPTH, ring[(c + DEPTH) % (DEPTH + 1)]);
        consume(c, ring[c % (DEPTH + 1)]);
      }
      // back to bytes, in place, for the rows that are used
#pragma unroll
      for (int r = 0; r < M; r++) {
        if (((syn_mask | t_mask) >> r) & 1u) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; i++) w[i] = acc[r * 8 + i];
          bit_transpose8(w);
#pragma unroll
          for (int i = 0; i < 8; i++) acc[r * 8 + i] = w[i];
        }
      }
      // table stage, 16 bytes at a time: packed[q*4+b] = sum over syndromes of entry(i, byte b of word q)
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; i++) packed[i] = 0;
        int si = 0;
#pragma unroll
        for (int r = 0; r < M; r++) {
          if ((syn_mask >> r) & 1u) {
            const uint32_t tb = tab_lane + (uint32_t)(si * 4 * kRecCopies);
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const uint32_t w = acc[r * 8 + hh * 4 + q];
              packed[q * 4 + 0] ^= lds32(byte_madd<0>(w, 256u, 256u << 16, tb));
              packed[q * 4 + 1] ^= lds32(byte_madd<1>(w, 256u, 256u << 16, tb));
              packed[q * 4 + 2] ^= lds32(byte_madd<2>(w, 256u, 256u << 16, tb));
              packed[q * 4 + 3] ^= lds32(byte_madd<3>(w, 256u, 256u << 16, tb));
            }
            si++;
          }
        }
        // unpack output j, add T_p for regenerated parity, store 16 bytes
        for (int j = 0; j < n_out; j++) {
          uint32_t o[4];
          const uint32_t sel = 0x0040u | (uint32_t)j | ((uint32_t)j << 4);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const uint32_t lo = prmt(packed[q * 4 + 0], packed[q * 4 + 1], sel);
            const uint32_t hi = prmt(packed[q * 4 + 2], packed[q * 4 + 3], sel);
            o[q] = prmt(lo, hi, 0x5410u);
          }
          const int prow = pat_s->out_prow[j];
#pragma unroll
          for (int r = 0; r < M; r++) {
            if (prow == r) {
#pragma unroll
              for (int q = 0; q < 4; q++) o[q] ^= acc[r * 8 + hh * 4 + q];
            }
          }
          if (live) {
            uint8_t* dst = sbase + (size_t)pat_s->out_slot[j] * p.shard_pitch + col + hh * 16;
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]),
                         "r"(o[3])
                         : "memory");
          }
        }
      }
    };

    for (uint32_t t = 0; t < T; t++) {
      const size_t tile_start = seg_start + (size_t)t * kBsTile;
      // whole-line requests: the two groups of a warp are two contiguous 1 KiB runs (see rs_bs_kernel)
      const size_t warp_lo = tile_start + (size_t)(tid >> 5) * (32 * kBsPiece);
      if (warp_lo >= p.shard_len) continue;   // per-warp decision, as in rs_bs_kernel
      const size_t col0 = warp_lo + (size_t)(tid & 31) * 32;
      if (warp_lo + 32 * kBsPiece <= p.shard_len) {
#pragma unroll 1
        for (int g = 0; g < kBsGroups; g++) group(std::true_type{}, col0 + (size_t)g * 1024);
      } else {
#pragma unroll 1
        for (int g = 0; g < kBsGroups; g++) group(std::false_type{}, col0 + (size_t)g * 1024);
      }
    }
  }
}

template <int K, int M>
static cudaError_t launch_rec_cfg(const BsRecParams& p, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rs_bsrec_kernel<K, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRecSmemBytes);
  if (e != cudaSuccess) return e;
  rs_bsrec_kernel<K, M><<<grid, kBsThreads, kRecSmemBytes, st>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
template <int K, int M>
static bool rows_match(const uint8_t* rows) {
  for (int i = 0; i < K * M; i++)
    if (rows[i] != BsNet<K, M>::kRows[i]) return false;
  return true;
}

template <int K, int M>
static cudaError_t launch_cfg(const BsParams& p, bool crc, bool verify, int grid, cudaStream_t st) {
  static bool configured = false;   // per process; attribute is per device function, set for every device lazily
  cudaError_t e;
  (void)configured;
  const bool packed = p.packed_pps != 0;
  if (crc) {
    auto kern = packed ? rs_bs_kernel<K, M, true, true> : rs_bs_kernel<K, M, true, false>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBsSmemBytes)) != cudaSuccess) return e;
    kern<<<grid, kBsThreads, kBsSmemBytes, st>>>(p);
  } else if (verify) {
    if (packed) rs_bs_kernel<K, M, false, true, true><<<grid, kBsThreads, 4096, st>>>(p);
    else rs_bs_kernel<K, M, false, false, true><<<grid, kBsThreads, 4096, st>>>(p);
  } else if (packed) {
    rs_bs_kernel<K, M, false, true><<<grid, kBsThreads, 4096, st>>>(p);
  } else {
    rs_bs_kernel<K, M, false, false><<<grid, kBsThreads, 4096, st>>>(p);
  }
  return cudaGetLastError();
}

#define CUBEEC_BS_CONFIGS(X) X(4, 2) X(6, 3) X(12, 4) X(20, 4) X(16, 4) X(10, 4) X(3, 3) X(4, 4) X(8, 4) X(6, 2) X(10, 2) X(5, 2)

bool bs_supported(int k, int m, const uint8_t* parity_rows) {
#define X(KK, MM) \
  if (k == KK && m == MM) return rows_match<KK, MM>(parity_rows);
  CUBEEC_BS_CONFIGS(X)
#undef X
  return false;
}

cudaError_t launch_bs_rec(int k, int m, const BsRecParams& p, int grid, cudaStream_t st) {
#define X(KK, MM) \
  if (k == KK && m == MM) return launch_rec_cfg<KK, MM>(p, grid, st);
  CUBEEC_BS_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t launch_bs(int k, int m, const BsParams& p, bool crc, bool verify, int grid, cudaStream_t st) {
  if (crc && verify) return cudaErrorInvalidValue;
#define X(KK, MM) \
  if (k == KK && m == MM) return launch_cfg<KK, MM>(p, crc, verify, grid, st);
  CUBEEC_BS_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace cbe
