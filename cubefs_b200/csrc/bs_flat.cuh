// bs_flat.cuh -- the fused RS encode + CRC32 kernel, second generation ("flat" work split, software-pipelined
// column body).  Same arithmetic as rs_bs_kernel<.., CRC != 0> (bs_kernel.cuh): bit-sliced XOR networks for
// the GF(2^8) coding (replaces reedSolomon.Encode's SIMD loop, RS/reedsolomon.go:609-625,897-1134) and
// slicing-by-4 table lookups for the per-shard CRC32 of blobstore/access/stream/stream_put.go:265-269 in
// the same HBM pass.  What changed, and why (profiles/r01_prof_r1_bs_crc.txt: issue-bound, 59.8 % issue
// slots, phases of pure LOP3 and phases of pure lookups in the SASS):
//
//  * Work split.  The unit of work is a WARP-tile: 2 KiB of every shard of one stripe (32 lanes x 64
//    contiguous bytes).  The units of the whole batch are numbered stripe-major and cut into one contiguous
//    run per warp of the grid, runs differing by at most one unit.  No CTA-wide tile loop, no barrier in
//    the main loop, no wave quantisation: 383 stripes load the 148 SMs as evenly as 1024 do, the ragged
//    end of a shard costs one warp one masked unit (not a CTA a masked tile), and a warp aligns/reduces
//    its CRC registers once per stripe it touches (every ~75 units for 4 MiB blobs) instead of every 11.
//    A run that crosses a stripe boundary flushes its partial remainders; crc_parts_finalize_kernel
//    (kernels.cu) chains the parts of a shard by their lengths.
//  * Column body.  ptxas scheduled the old body as "all lookups of a few shards, then a long stretch of
//    LOP3": with 4 warps per scheduler the ALU pipe and the LDS latency were exposed in turn.  Here the
//    body is written as two tracks that the source interleaves statement by statement: the CRC chain of
//    shard c+1 (8 dependent lookup steps) runs between the pieces of the bit transpose and XOR network of
//    shard c; in the parity phase the CRC of parity r runs inside the back-transpose of parity r+1.  Loads
//    are two shards ahead (three 256-bit buffers), and the first two loads of the NEXT column are issued
//    before the parity phase of this one, so no column starts with an exposed DRAM round trip.
#pragma once
#include <type_traits>
#include <utility>

#include "bs_net_gen.cuh"
#include "kernels.cuh"
#include "bs_device.cuh"

namespace cbe {

using namespace bsdev;

template <int I, int N>
struct StaticFor {
  template <class F>
  static __device__ __forceinline__ void run(F&& f) {
    if constexpr (I < N) {
      f(std::integral_constant<int, I>{});
      StaticFor<I + 1, N>::run(f);
    }
  }
};

// one third of the 8x8 bit transpose (bs_device.cuh: bit_transpose8 = stages 0, 1, 2)
template <int J>
__device__ __forceinline__ void transpose_stage(uint32_t (&w)[8]) {
  if constexpr (J == 0) {
    delta_swap<4, 0x0f0f0f0fu>(w[0], w[4]);
    delta_swap<4, 0x0f0f0f0fu>(w[1], w[5]);
    delta_swap<4, 0x0f0f0f0fu>(w[2], w[6]);
    delta_swap<4, 0x0f0f0f0fu>(w[3], w[7]);
  }
  if constexpr (J == 1) {
    delta_swap<2, 0x33333333u>(w[0], w[2]);
    delta_swap<2, 0x33333333u>(w[1], w[3]);
    delta_swap<2, 0x33333333u>(w[4], w[6]);
    delta_swap<2, 0x33333333u>(w[5], w[7]);
  }
  if constexpr (J == 2) {
    delta_swap<1, 0x55555555u>(w[0], w[1]);
    delta_swap<1, 0x55555555u>(w[2], w[3]);
    delta_swap<1, 0x55555555u>(w[4], w[5]);
    delta_swap<1, 0x55555555u>(w[6], w[7]);
  }
}

// owner of unit u: the warp g with lo(g) <= u < lo(g+1), lo(g) = floor(g * U / GW)
__host__ __device__ __forceinline__ uint64_t bsf_run_lo(uint64_t g, uint64_t U, uint64_t GW) { return g * U / GW; }
__host__ __device__ __forceinline__ uint64_t bsf_owner(uint64_t u, uint64_t U, uint64_t GW) {
  uint64_t g = u * GW / U;
  while (g + 1 < GW && bsf_run_lo(g + 1, U, GW) <= u) g++;
  while (g > 0 && bsf_run_lo(g, U, GW) > u) g--;
  return g;
}

// CRC: 1 = every shard, 2 = the M outputs only (later passes of an m > 4 code, LRC local stripes).
// RD = number of 256-bit load buffers (RD - 1 shards in flight ahead of the one being coded).
// PF: also prefetch into L2 (no registers) shard c of the NEXT column while shard c of this one is coded, so that the
// 256-bit loads of the ring find their lines in L2.  Measured (B200, C2): 0.519 with, 0.535 without -- off.
// A variant that staged the data shards through shared memory with cp.async (counted waits, no LDG scoreboards: ptxas puts
// the ring's LDGs on scoreboards 2-5 and the CRC lookups on 0-4, so lookup waits can inherit DRAM latency) was measured at
// 0.48-0.49 (384 threads, 4 / 6 slots) and 0.505 (512 threads x 128 registers) against 0.536: the extra shared-memory
// traffic costs more than the aliasing.  Not kept.
// (Warp counts between 12 and 16 do not exist for this kernel: registers are per scheduler, 16 K each, so 4 warps per
// scheduler cap a thread at 128 registers and 3 at 168; ptxas picks exactly those two.)
template <int K, int M, int V, int CRC, int NT, int RD, bool PF = false, int ENTRY = 0>
__device__ __forceinline__ void bsf_body(const BsfParams& p) {
  static_assert(CRC == 1 || CRC == 2, "fused-CRC kernel");
  static_assert(K >= 2 && K + M <= 32, "lane q publishes the remainder of shard q");
  static_assert(RD >= 2 && K >= RD - 1, "load ring");
  using Net = BsNet<K, M, V>;
  constexpr int C0 = CRC == 2 ? K : 0;   // first checksummed shard (local index)
  constexpr int NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // copies of the fold tables (Horner step between units), see bsf_fold_copies()
  constexpr int FC = K >= 16 ? CUBEEC_FC_HI : CUBEEC_FC_LO;   // (= bsf_fold_copies(K), kernels.cuh)
  constexpr size_t FTB = 256 * FC * 4;

  // ---- shared memory: [mbarrier | fold tables (FC copies) | slice image].  Nothing here needs more than 128-byte
  // alignment (a lookup address is base + 256 * byte + 4 * lane, all additions), so the image follows the fold tables
  // directly: 192 KB with 16 fold copies (160 KB with 8) instead of the 209 KB (193 KB) a 64K-aligned image costs.
  // The size matters beyond occupancy: 209 KB selects the 228 KB shared-memory carve-out, and what is left of the SM's
  // 256 KB is the L1 that holds the loads in flight -- measured with that layout: crc_flat_kernel 0.69 -> 0.53 of the
  // HBM peak, RS(20,4) 0.495 -> 0.476.  (Below the 196 KB carve-out nothing more is gained: 160 KB = 192 KB.)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* fold_s = reinterpret_cast<uint32_t*>(smem + 128);   // [4][256][FC]
  const uint32_t base_addr = smem_addr(smem);
  const uint32_t tab_addr = base_addr + (uint32_t)(128 + 4 * FTB);
  {
    uint8_t* tab_ptr = smem + (tab_addr - base_addr);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      // TMA 1-D bulk copies: the lane-private slicing tables, two 64 KiB halves
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
      for (int q = 0; q < FC; q++) fold_s[i * FC + q] = v;
    }
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
  const uint32_t lane_base = tab_addr + (uint32_t)(lane * 4);
  const uint32_t fold_lane = smem_addr(fold_s) + (uint32_t)((lane & (FC - 1)) * 4);
  const uint32_t klane = p.klane[lane];   // x^(8 * 64 * (31 - lane)): aligns a lane's remainder to the end of the unit

  auto slice4 = [&](uint32_t y) -> uint32_t {
    const uint32_t a0 = byte_madd<0>(y, 256u, 256u << 16, lane_base);
    const uint32_t a1 = byte_madd<1>(y, 256u, 256u << 16, lane_base);
    const uint32_t a2 = byte_madd<2>(y, 256u, 256u << 16, lane_base);
    const uint32_t a3 = byte_madd<3>(y, 256u, 256u << 16, lane_base);
    const uint32_t t3 = lds32_off<65536 + 128>(a0);
    const uint32_t t2 = lds32_off<65536>(a1);
    const uint32_t t1 = lds32_off<128>(a2);
    const uint32_t t0 = lds32_off<0>(a3);
    return t3 ^ t2 ^ t1 ^ t0;
  };
  auto fold = [&](uint32_t u) -> uint32_t {
    constexpr uint32_t ST = FC * 4;
    return lds32(byte_madd<0>(u, ST, ST << 16, fold_lane + 0 * 256 * ST)) ^ lds32(byte_madd<1>(u, ST, ST << 16, fold_lane + 1 * 256 * ST)) ^
           lds32(byte_madd<2>(u, ST, ST << 16, fold_lane + 2 * 256 * ST)) ^ lds32(byte_madd<3>(u, ST, ST << 16, fold_lane + 3 * 256 * ST));
  };

  uint32_t crc_u[K + M];
#pragma unroll
  for (int i = 0; i < K + M; i++) crc_u[i] = 0;
  // RD 256-bit buffers: shard c is coded from ring[c % RD] while shard c+1 is checksummed in ring[(c+1) % RD]
  // and shards c+2 .. c+RD-1 are in flight
  uint32_t ring[RD][8];
#pragma unroll
  for (int b = 0; b < RD; b++)
#pragma unroll
    for (int i = 0; i < 8; i++) ring[b][i] = 0;

  const uint64_t U = p.total_units;
  const uint64_t GW = (uint64_t)gridDim.x * NW, gw = (uint64_t)blockIdx.x * NW + warp;
  const uint64_t u_lo = bsf_run_lo(gw, U, GW), u_hi = bsf_run_lo(gw + 1, U, GW);
  const uint32_t wt = p.units_per_shard;

  // a 32-byte column of this lane: stripe base, byte offset in the shard, liveness
  struct Col {
    const uint8_t* sbase;
    uint32_t col;
    bool live;   // col < shard_len (lane-level)
  };
  auto locate = [&](uint32_t s, uint32_t t, uint32_t g) -> Col {
    Col c;
    c.sbase = p.base + (size_t)s * p.stripe_pitch;
    c.col = t * (32u * kBsPiece) + (uint32_t)lane * kBsPiece + g * 32u;
    c.live = c.col < p.shard_len;
    return c;
  };

  // ---- one 32-byte column of every shard.  ring[0 .. RD-2] already hold (or are receiving) data shards
  // 0 .. RD-2 of this column; on return they hold those of column `nx` (if nx.live).
  auto column = [&](auto full_tag, const Col cc, const Col nx, const bool nx_valid) {
    constexpr bool FULL = decltype(full_tag)::value;
    uint32_t msk[8];
    if constexpr (!FULL) {
      const int tail = (cc.live && cc.col + 32 > p.shard_len) ? (int)(p.shard_len - cc.col) : 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int rem = tail - 4 * i;
        msk[i] = !cc.live ? 0u : ((tail == 0 || rem >= 4) ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u)));
      }
    }
    uint32_t acc[8 * M];
#pragma unroll
    for (int i = 0; i < 8 * M; i++) acc[i] = 0;
    uint32_t t[32];   // XOR combinations of the network part functions (registers; most entries never exist)
    const uint8_t* src = cc.sbase + cc.col;
    const uint8_t* nsrc = nx.sbase + nx.col;
    const bool nlive = nx_valid && nx.live;

    auto mask_buf = [&](uint32_t (&w)[8]) {
      if constexpr (!FULL) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] &= msk[i];
      }
    };
    // prologue: checksum data shard 0 (nothing of this column to overlap it with)
    mask_buf(ring[0]);
    if constexpr (CRC == 1) {
      uint32_t u = crc_u[0];
#pragma unroll
      for (int i = 0; i < 8; i++) u = slice4(u ^ ring[0][i]);
      crc_u[0] = u;
    }
    StaticFor<0, K>::run([&](auto cconst) {
      constexpr int c = decltype(cconst)::value;
      // load RD-1 shards ahead; past the last data shard: shards 0 .. RD-2 of the next column
      if constexpr (c + RD - 1 < K) {
        if (FULL || cc.live) ldg256(src + (size_t)p.in_slot[c + RD - 1] * p.shard_pitch, ring[(c + RD - 1) % RD]);
      } else {
        if (nlive) ldg256(nsrc + (size_t)p.in_slot[c + RD - 1 - K] * p.shard_pitch, ring[(c + RD - 1) % RD]);
      }
      if constexpr (PF) {
        if (nlive) asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + (size_t)p.in_slot[c] * p.shard_pitch));
      }
      uint32_t (&w)[8] = ring[c % RD];
      if constexpr (c + 1 < K) {
        uint32_t (&wn)[8] = ring[(c + 1) % RD];
        mask_buf(wn);
        uint32_t u = CRC == 1 ? crc_u[c + 1] : 0u;
        StaticFor<0, 8>::run([&](auto jc) {
          constexpr int j = decltype(jc)::value;
          if constexpr (CRC == 1) u = slice4(u ^ wn[j]);
          if constexpr (j < 3) transpose_stage<j>(w);
          else Net::template part<c, j - 3>(w, acc, t);
        });
        if constexpr (CRC == 1) crc_u[c + 1] = u;
      } else {
        transpose_stage<0>(w);
        transpose_stage<1>(w);
        transpose_stage<2>(w);
        StaticFor<0, Net::kParts>::run([&](auto jc) { Net::template part<c, decltype(jc)::value>(w, acc, t); });
      }
    });
    // parity phase: back to bytes, store, checksum; the CRC of parity r runs inside the transpose of r+1
    const bool do_store = FULL || cc.live;
    uint8_t* dst = const_cast<uint8_t*>(src);
    StaticFor<0, M + 1>::run([&](auto rconst) {
      constexpr int r = decltype(rconst)::value;
      uint32_t u = 0;
      if constexpr (r > 0) u = crc_u[K + r - 1];
      StaticFor<0, 8>::run([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (r > 0) u = slice4(u ^ acc[(r - 1) * 8 + j]);
        if constexpr (r < M && j < 3) {
          uint32_t o[8];
#pragma unroll
          for (int i = 0; i < 8; i++) o[i] = acc[r * 8 + i];
          transpose_stage<j>(o);
#pragma unroll
          for (int i = 0; i < 8; i++) acc[r * 8 + i] = o[i];
          if constexpr (j == 2) {
            if (do_store) stg256(dst + (size_t)p.out_slot[r] * p.shard_pitch, o);
          }
        }
      });
      if constexpr (r > 0) crc_u[K + r - 1] = u;
    });
    // (ring[(K + j) % RD] now holds the next column's shard j, j < RD - 1: move them to ring[j])
    if constexpr (K % RD != 0) {
      uint32_t tmp[RD - 1][8];
#pragma unroll
      for (int j = 0; j < RD - 1; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) tmp[j][i] = ring[(K + j) % RD][i];
#pragma unroll
      for (int j = 0; j < RD - 1; j++)
#pragma unroll
        for (int i = 0; i < 8; i++) ring[j][i] = tmp[j][i];
    }
  };

  auto flush = [&](uint32_t s) {
    gf32_mul_common<C0, K + M>(crc_u, klane, p.poly);
    uint32_t mine = 0;
#pragma unroll
    for (int q = C0; q < K + M; q++) {
      const uint32_t v = __reduce_xor_sync(0xffffffffu, crc_u[q]);
      crc_u[q] = 0;
      if (lane == q) mine = v;
    }
    // part number = position of this warp among the warps whose runs touch stripe s
    const uint32_t part = (uint32_t)(gw - bsf_owner((uint64_t)s * wt, U, GW));
    if (lane >= C0 && lane < K + M) {
      const uint32_t slot = lane < K ? (uint32_t)lane : (uint32_t)p.out_slot[0] + (uint32_t)(lane - K);
      p.crc_part[((size_t)s * p.n_slots + slot) * p.max_parts + part] = mine;
    }
  };

  if (u_lo < u_hi) {
    uint32_t s = (uint32_t)(u_lo / wt), t = (uint32_t)(u_lo - (uint64_t)s * wt);
    {
      const Col c0 = locate(s, t, 0);
      if (c0.live) {
#pragma unroll
        for (int j = 0; j < RD - 1; j++) ldg256(c0.sbase + c0.col + (size_t)p.in_slot[j] * p.shard_pitch, ring[j]);
      }
    }
    for (uint64_t u = u_lo; u < u_hi; u++) {
      // the unit after this one (stripe-major numbering)
      const bool wrap = t + 1 == wt;
      const uint32_t s2 = wrap ? s + 1 : s, t2 = wrap ? 0u : t + 1;
      const bool more = u + 1 < u_hi;
      const bool full = (t + 1) * (32u * kBsPiece) <= p.shard_len;   // warp-uniform
      // p.sync_units: the warps of the block start every unit together.  The straight-line body of a long code is
      // several times the SM's 32 KB instruction cache (124 KB for RS(12,4), 182 KB for RS(20,4)); in step, one fetch
      // from L2 can serve all 12 warps instead of one each.  Measured on B200 (with the grid-constant entry):
      // RS(20,4) 0.437 -> 0.494, RS(16,4) 0.48 -> 0.51, RS(12,4) C2 0.528 -> 0.543; k <= 10 loses 10-15 %.  Finer
      // barriers (per column, per shard), warp groups on named barriers, and a rolled shard loop whose body fits the
      // cache (networks behind a switch) were all slower.  Only the grid-constant entry carries the test.
      if constexpr (ENTRY == 1) {
        if (p.sync_units) __syncthreads();
      }
#pragma unroll 1
      for (uint32_t g = 0; g < 2; g++) {
        const Col cc = locate(s, t, g);
        const Col nx = g == 0 ? locate(s, t, 1) : locate(s2, t2, 0);
        const bool nx_valid = g == 0 || more;
        if (full) column(std::true_type{}, cc, nx, nx_valid);
        else column(std::false_type{}, cc, nx, nx_valid);
      }
      if (more) {
        if (wrap) {
          flush(s);
        } else {
          // Horner step over the gap between this lane's pieces of consecutive units
#pragma unroll
          for (int i = C0; i < K + M; i++) crc_u[i] = fold(crc_u[i]);
        }
      }
      s = s2;
      t = t2;
    }
    flush(t == 0 ? s - 1 : s);   // (s, t) is one unit past the run: the last unit's stripe
  }
  if constexpr (ENTRY == 1) {
    if (p.sync_units) {
      // runs are floor(U / GW) or one more units long: the shorter ones still owe the block their barriers
      const uint64_t longest = (U + GW - 1) / GW;
      for (uint64_t i = u_hi - u_lo; i < longest; i++) __syncthreads();
    }
  }
}
// Two entry points over the same body: parameters by value (ENTRY 0), or __grid_constant__ with the per-unit barrier
// (ENTRY 1).  How the parameter block is addressed changes nothing semantically, but the schedule ptxas finds for this
// register-bound kernel follows it.  Measured on B200 (fraction of the measured HBM peak, fused encode + CRC,
// profiles/r02_sweep_bsf_entry_sync.jsonl): RS(12,4) C2 0.531 by value / 0.528 grid-constant / 0.543 + barrier;
// RS(16,4) 0.43 / 0.48 / 0.51; RS(20,4) 0.396 / 0.437 / 0.494; RS(24,8) passes 0.26 / - / 0.28; k <= 10: by value.
template <int K, int M, int V, int CRC, int NT, int RD = (NT <= 384 ? 4 : 3)>
__global__ void __launch_bounds__(NT, 1) rs_bsf_kernel(const BsfParams p) {
  bsf_body<K, M, V, CRC, NT, RD, false, 0>(p);
}
template <int K, int M, int V, int CRC, int NT, int RD = (NT <= 384 ? 4 : 3)>
__global__ void __launch_bounds__(NT, 1) rs_bsf_kernel_gc(const __grid_constant__ BsfParams p) {
  bsf_body<K, M, V, CRC, NT, RD, false, 1>(p);   // (its own instantiation of the body, so that both entries inline theirs)
}
constexpr bool bsf_has_gc_entry(int k, int m, int v, int mode) { return (m == 4 && v == 0 && mode == 1) || k >= 15; }
// flip (cubeec_debug_force_kernel(3000 + bits), A/B aid): bit 0 the entry, bit 1 the barrier
template <int K, int M, int V, int MODE, int NT = kBsfThreads>
static cudaError_t bsf_launch_one(const BsfParams& p0, int grid, cudaStream_t st, int flip = 0) {
  BsfParams p = p0;
  p.sync_units = 0;
  void (*kern)(const BsfParams) = rs_bsf_kernel<K, M, V, MODE, NT>;
  if constexpr (bsf_has_gc_entry(K, M, V, MODE)) {
    if ((K >= 12) != ((flip & 1) != 0)) {
      kern = rs_bsf_kernel_gc<K, M, V, MODE, NT>;
      p.sync_units = ((K >= 12) != ((flip & 2) != 0)) ? 1u : 0u;
    }
  }
  constexpr size_t smem_bytes = bsf_smem_bytes(bsf_fold_copies(K));
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return e;
  kern<<<grid, NT, smem_bytes, st>>>(p);
  return cudaGetLastError();
}

}  // namespace cbe
