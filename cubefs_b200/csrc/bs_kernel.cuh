// bs_kernel.cuh -- the bit-sliced RS encode / verify kernel template (see bitslice.cu for the design
// notes).  Included by the translation units that instantiate it: bitslice.cu (single-pass codes) and
// bitslice_mp.cu (passes of the codes with more than 4 parity shards).
#pragma once
#include <type_traits>

#include "bs_net_gen.cuh"
#include "kernels.cuh"
#include "bs_device.cuh"

namespace cbe {

using namespace bsdev;

// Stripe slot of the CRC register with local index q (inputs 0..K-1, outputs K..K+M-1).  The data shards
// are checksummed (CRC mode 1) only by a code whose inputs are slots 0..K-1 -- LRC local stripes use
// mode 2 -- and the outputs of a pass are consecutive slots, so no table lookup is needed.  (A run-time
// index into the parameter arrays made the compiler copy the parameter block to local memory: the fused
// kernel fell from 52 % to 42 %.)
__device__ __forceinline__ uint32_t bs_slot(const BsParams& p, uint32_t q, uint32_t k) {
  return q < k ? q : (uint32_t)p.out_slot[0] + (q - k);
}

// VERIFY (reedSolomon.Verify / checkSomeShards, RS/reedsolomon.go:770-784,1287-1301): the computed
// parity is compared with the stored parity instead of being written; any difference raises the
// stripe's flag in p.mismatch.  The reference allocates m scratch shards and runs bytes.Equal.
// CRC: 0 = none, 1 = every shard, 2 = the M outputs only (later passes of an MT > 4 code and LRC local
// stripes: the inputs were checksummed by an earlier pass).  V selects the network (bs_net_gen.cuh):
// 0 = RS(K, M), else rows [kRow0, kRow0+M) of RS(K, kTotalM).  Which shard of the stripe input c / output
// r is comes from p.in_slot / p.out_slot (identity + kRow0 for a plain code; the AZ's shard list for an
// LRC local stripe, codemode.GetECLayoutByAZ).
//
// ROLLED (experiment, cubeec_debug_force_kernel(6)): the fully unrolled column loop of the fused-CRC
// variant is 48 KB of straight-line code, more than the 32 KB instruction cache, and anything that lets
// warps drift apart in it costs throughput.  ROLLED keeps ONE copy of the CRC absorption
// and the bit transpose in a loop over shard pairs and dispatches only the per-shard XOR network through
// a switch: about 23 KB of hot code, no calls.
template <int K, int M, int V, int CRC, bool PACKED, bool VERIFY = false, bool ROLLED = false>
__global__ void __launch_bounds__(kBsThreads, 1) rs_bs_kernel(const BsParams p) {
  static_assert(!(CRC && VERIFY), "verify does not checksum");
  using Net = BsNet<K, M, V>;
  constexpr int C0 = CRC == 2 ? K : 0;      // first checksummed shard (local index)
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
    constexpr int DEPTH = CRC ? 1 : (M > 6 ? 1 : (M > 4 ? 2 : 4));   // wide passes (M > 4) need the registers for accumulators
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
      if (CRC == 1) {
        uint32_t u = crc_u[c];
#pragma unroll
        for (int i = 0; i < 8; i++) u = slice4(u ^ w[i]);
        crc_u[c] = u;
      }
      bit_transpose8(w);
      ApplyAt<Net, 0, K>::run(c, w, acc);   // c is a compile-time constant after unrolling
    };
    if constexpr (ROLLED && FULL && CRC == 1) {
      static_assert(!ROLLED || (K % 2 == 0 && !VERIFY), "rolled variant: even k, encode only");
      // two shard buffers alternate; the running CRC register u is swapped with crc_u[] inside the
      // switch cases, where the shard index is a constant (inputs are slots 0..K-1 in CRC mode 1)
      uint32_t (&a)[8] = ring[0];
      uint32_t (&b)[8] = ring[1];
      ldg256(src, a);
      uint32_t u = crc_u[0];
#pragma unroll 1
      for (int c = 0; c < K; c += 2) {
        ldg256(src + (size_t)(c + 1) * p.shard_pitch, b);
#pragma unroll
        for (int i = 0; i < 8; i++) u = slice4(u ^ a[i]);
        bit_transpose8(a);
        RollAt<Net, 0, K>::run(c, a, acc, crc_u, u);
        if (c + 2 < K) ldg256(src + (size_t)(c + 2) * p.shard_pitch, a);
#pragma unroll
        for (int i = 0; i < 8; i++) u = slice4(u ^ b[i]);
        bit_transpose8(b);
        RollAt<Net, 1, K>::run(c + 1, b, acc, crc_u, u);
      }
    } else {
#pragma unroll
      for (int c = 0; c < DEPTH && c < K; c++)
        if (live) ldg256(src + (size_t)p.in_slot[c] * p.shard_pitch, ring[c % (DEPTH + 1)]);
#pragma unroll
      for (int c = 0; c < K; c++) {
        if (c + DEPTH < K && live) ldg256(src + (size_t)p.in_slot[c + DEPTH] * p.shard_pitch, ring[(c + DEPTH) % (DEPTH + 1)]);
        shard(c, ring[c % (DEPTH + 1)]);
      }
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
          ldg256(sbase + (size_t)p.out_slot[r] * p.shard_pitch + col, e);
#pragma unroll
          for (int i = 0; i < 8; i++) vdiff |= o[i] ^ (FULL ? e[i] : (e[i] & msk[i]));
        }
      } else if (live) {
        stg256(sbase + (size_t)p.out_slot[r] * p.shard_pitch + col, o);
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
        gf32_mul_common<C0, K + M>(crc_u, kt, p.poly);   // kt = 0 for inactive threads
#pragma unroll
        for (int q = C0; q < K + M; q++) {
          uint32_t u = crc_u[q];
          crc_u[q] = 0;
          u = __reduce_xor_sync(peers, u);
          if (leader) atomicXor(&red2_s[(stripe - stripe0) * (K + M) + q], u);
        }
        __syncthreads();
        for (uint32_t i = tid; i < nstr * (K + M); i += NT) {
          const uint32_t sj = stripe0 + i / (K + M), q = i % (K + M);
          const uint32_t segidx = tile - (uint32_t)(((uint64_t)sj * PPS) / NT);   // 0 or 1: a shard spans at most two tiles
          if (q >= (uint32_t)C0) p.crc_part[((size_t)sj * p.n_slots + bs_slot(p, q, (uint32_t)K)) * 2 + segidx] = red2_s[i];
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
        for (int i = C0; i < K + M; i++) crc_u[i] = fold(crc_u[i]);
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
      gf32_mul_common<C0, K + M>(crc_u, kt, p.poly);
#pragma unroll
      for (int q = C0; q < K + M; q++) {
        uint32_t u = crc_u[q];
        crc_u[q] = 0;
        u = __reduce_xor_sync(0xffffffffu, u);
        if (lane == 0) red_s[q * NW + warp] = u;
      }
      __syncthreads();
      if (tid >= C0 && tid < K + M) {
        uint32_t u = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) u ^= red_s[tid * NW + w2];
        p.crc_part[((size_t)s * p.n_slots + bs_slot(p, (uint32_t)tid, (uint32_t)K)) * p.n_seg + seg] = u;
      }
      __syncthreads();
    }
  }
  }   // !PACKED
}


template <int K, int M, int V>
static bool bs_rows_match(const uint8_t* rows /* [kTotalM][K] of the handle */) {
  using Net = BsNet<K, M, V>;
  for (int i = 0; i < K * M; i++)
    if (rows[Net::kRow0 * K + i] != Net::kRows[i]) return false;
  return true;
}

// crc: 0 none, 1 all shards, 2 outputs only.  Only the variants a network can be asked for are
// instantiated, WHAT = OR of: 1 fused CRC of every shard, 2 plain encode + verify, 4 fused CRC of the
// outputs only.
template <int K, int M, int V, int MODE>
static cudaError_t bs_launch_crc(const BsParams& p, int grid, cudaStream_t st) {
  auto kern = p.packed_pps != 0 ? rs_bs_kernel<K, M, V, MODE, true> : rs_bs_kernel<K, M, V, MODE, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBsSmemBytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, kBsThreads, kBsSmemBytes, st>>>(p);
  return cudaGetLastError();
}

template <int K, int M, int V, int WHAT>
static cudaError_t bs_launch_cfg(const BsParams& p, int crc, bool verify, int grid, cudaStream_t st) {
  const bool packed = p.packed_pps != 0;
  if (crc && verify) return cudaErrorInvalidValue;
  if (crc == 1) {
    if constexpr ((WHAT & 1) != 0) return bs_launch_crc<K, M, V, 1>(p, grid, st);
    return cudaErrorInvalidValue;
  }
  if (crc == 2) {
    if constexpr ((WHAT & 4) != 0) return bs_launch_crc<K, M, V, 2>(p, grid, st);
    return cudaErrorInvalidValue;
  }
  if constexpr ((WHAT & 2) != 0) {
    if (verify) {
      if (packed) rs_bs_kernel<K, M, V, 0, true, true><<<grid, kBsThreads, 4096, st>>>(p);
      else rs_bs_kernel<K, M, V, 0, false, true><<<grid, kBsThreads, 4096, st>>>(p);
    } else if (packed) {
      rs_bs_kernel<K, M, V, 0, true><<<grid, kBsThreads, 4096, st>>>(p);
    } else {
      rs_bs_kernel<K, M, V, 0, false><<<grid, kBsThreads, 4096, st>>>(p);
    }
    return cudaGetLastError();
  }
  return cudaErrorInvalidValue;
}

}  // namespace cbe
