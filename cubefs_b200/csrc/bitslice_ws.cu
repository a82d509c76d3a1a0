// bitslice_ws.cu -- warp-specialised variant of the fused RS encode + CRC32 kernel (sm_100a).
//
// Same arithmetic and the same reference boundary as bitslice.cu (reedSolomon.Encode,
// RS/reedsolomon.go:609-625, plus crc32.ChecksumIEEE per shard, blobstore/access/stream/
// stream_put.go:265-269); what changes is how the work is laid on the SM.  rs_bs_kernel<crc> makes
// every thread do both jobs and ends up issue/latency bound at 16 warps per SM (all 64 K registers
// gone, the serial slicing-by-4 chains of 16 shards interleaved by the compiler with the XOR
// network).  The two jobs want different pipes:
//
//   coder warps    (kBswE threads): 256-bit loads, 8x8 bit transposes, XOR network, parity stores,
//                  CRC of the 4 parity shards from registers        -> ALU pipe (LOP3/SHF)
//   checksum warps (kBswC threads): CRC of the K data shards, slicing-by-4 on lane-private tables
//                                                                   -> FMA pipe (IDP.2A) + LDS
//
// so the CTA is split by role and the register file is re-divided with setmaxnreg (coder 120,
// checksum 40 registers per thread): 24 warps per SM instead of 16, two short instruction streams
// instead of one 50 KB one, and the warp scheduler -- not the compiler -- interleaves the CRC
// chains with the network.  The checksum warps read the data shards a second time; both roles
// walk the same tile between two CTA barriers, so the second read is an L2 hit (HBM traffic stays
// 12 reads + 4 writes per column).
//
// Thread <-> byte mapping is the one of rs_bs_kernel with NT = kBswE: in tile t coder thread j owns
// bytes [t*TILE + 64 j, +64) of every shard, and checksum thread j owns the same piece of the data
// shards (slot i = shard i when kBswC == kBswE), so Horner stepping (fold tables), the per-thread
// alignment constants and crc_finalize_kernel are shared with bitslice.cu.
#include <type_traits>

#include "bs_net_gen.cuh"
#include "kernels.cuh"
#include "bs_device.cuh"

namespace cbe {

using namespace bsdev;

namespace {

template <int N>
__device__ __forceinline__ void role_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}

}  // namespace

template <int K, int M>
__global__ void __launch_bounds__(kBswE + kBswC, 1) rs_bsw_kernel(const BsParams p) {
  using Net = BsNet<K, M>;
  constexpr int NE = kBswE, NC = kBswC, NT = NE + NC, NWE = NE / 32;
  constexpr int SLOTS = (K * NE + NC - 1) / NC;      // data-shard pieces per checksum thread per tile
  constexpr int TILE = kBswTile;
  static_assert(NE % 128 == 0 && NC % 128 == 0, "setmaxnreg works on warpgroups");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- shared memory: [mbarrier | fold tables | kthread | reductions | checksum states] ... [64K-aligned slice image]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* fold_s = reinterpret_cast<uint32_t*>(smem + 64);   // [4][256][kBsFoldCopies]
  uint32_t* kth_s = fold_s + 4 * 256 * kBsFoldCopies;          // [NE]
  uint32_t* red_s = kth_s + NE;                                // [M][NWE]   parity remainders per coder warp
  uint32_t* red_c = red_s + M * NWE;                           // [K]        data remainders
  uint32_t* state_s = red_c + ((K + 31) & ~31);                // [SLOTS][NC] Horner registers of the checksum threads
  constexpr uint32_t kMisc = 64 + 4 * 256 * kBsFoldCopies * 4 + NE * 4 + M * NWE * 4 + ((K + 31) & ~31) * 4 + SLOTS * NC * 4;
  static_assert(kMisc + 2048 <= 65536, "misc area must end before the 64 KiB aligned table image");
  const uint32_t base_addr = smem_addr(smem);
  const uint32_t tab_addr = (base_addr + kMisc + 65535u) & ~65535u;
  {
    uint8_t* tab_ptr = smem + (tab_addr - base_addr);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
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
    if (tid < NE) kth_s[tid] = p.kthread[tid];
    for (int i = tid; i < SLOTS * NC; i += NT) state_s[i] = 0;
    if (tid < K) red_c[tid] = 0;
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
  const uint32_t lane_base = tab_addr | (uint32_t)(lane * 4);
  const uint32_t fold_lane = smem_addr(fold_s) + (uint32_t)((lane & (kBsFoldCopies - 1)) * 4);

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
    constexpr uint32_t ST = kBsFoldCopies * 4;
    return lds32(byte_madd<0>(u, ST, ST << 16, fold_lane + 0 * 256 * ST)) ^ lds32(byte_madd<1>(u, ST, ST << 16, fold_lane + 1 * 256 * ST)) ^
           lds32(byte_madd<2>(u, ST, ST << 16, fold_lane + 2 * 256 * ST)) ^ lds32(byte_madd<3>(u, ST, ST << 16, fold_lane + 3 * 256 * ST));
  };

  const uint32_t n_items = p.n_stripes * p.n_seg;
  const size_t seg_bytes = (size_t)p.tiles_per_seg * TILE;

  if (warp < NWE) {
    // =========================== coder role ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kBswRegsE));
    uint32_t crc_u[M];
#pragma unroll
    for (int i = 0; i < M; i++) crc_u[i] = 0;

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
#ifndef CUBEEC_BSW_DEPTH
#define CUBEEC_BSW_DEPTH 2
#endif
      constexpr int DEPTH = CUBEEC_BSW_DEPTH;   // 256-bit loads in flight ahead of the shard being coded
      uint32_t ring[DEPTH + 1][8];
#pragma unroll
      for (int b = 0; b <= DEPTH; b++)
#pragma unroll
        for (int i = 0; i < 8; i++) ring[b][i] = 0;
      const uint8_t* src = sbase + col;
#pragma unroll
      for (int c = 0; c < DEPTH && c < K; c++)
        if (live) ldg256(src + (size_t)c * p.shard_pitch, ring[c % (DEPTH + 1)]);
#pragma unroll
      for (int c = 0; c < K; c++) {
        if (c + DEPTH < K && live) ldg256(src + (size_t)(c + DEPTH) * p.shard_pitch, ring[(c + DEPTH) % (DEPTH + 1)]);
        uint32_t(&w)[8] = ring[c % (DEPTH + 1)];
        if (!FULL) {
#pragma unroll
          for (int i = 0; i < 8; i++) w[i] &= msk[i];
        }
        bit_transpose8(w);
        ApplyAt<Net, 0, K>::run(c, w, acc);
      }
#pragma unroll
      for (int r = 0; r < M; r++) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = acc[r * 8 + i];
        bit_transpose8(o);
        if (live) stg256(sbase + (size_t)(K + r) * p.shard_pitch + col, o);
        uint32_t u = crc_u[r];
#pragma unroll
        for (int i = 0; i < 8; i++) u = slice4(u ^ o[i]);
        crc_u[r] = u;
      }
    };

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t s = item / p.n_seg, seg = item - s * p.n_seg;
      uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
      const uint32_t T = (seg == p.n_seg - 1) ? p.tiles_last : p.tiles_per_seg;
      const size_t seg_start = (size_t)seg * seg_bytes;
      for (uint32_t t = 0; t < T; t++) {
#pragma unroll
        for (int i = 0; i < M; i++) crc_u[i] = fold(crc_u[i]);
        const size_t tile_start = seg_start + (size_t)t * TILE;
        const size_t col0 = tile_start + (size_t)tid * kBsPiece;
        if (tile_start + TILE <= p.shard_len) {
#pragma unroll 1
          for (int g = 0; g < kBsGroups; g++) group(std::true_type{}, sbase, col0 + (size_t)g * 32);
        } else {
#pragma unroll 1
          for (int g = 0; g < kBsGroups; g++) group(std::false_type{}, sbase, col0 + (size_t)g * 32);
        }
#ifndef CUBEEC_BSW_NOTILEBAR
        role_barrier<NT>();   // keeps both roles inside one tile: the second read of a data column hits L2
#endif
      }
      const uint32_t kt = kth_s[tid];
#pragma unroll
      for (int q = 0; q < M; q++) {
        uint32_t u = gf32_mul_dev(crc_u[q], kt, p.poly);
        crc_u[q] = 0;
        u = __reduce_xor_sync(0xffffffffu, u);
        if (lane == 0) red_s[q * NWE + warp] = u;
      }
      role_barrier<NT>();
      if (tid < M) {
        uint32_t u = 0;
#pragma unroll
        for (int w2 = 0; w2 < NWE; w2++) u ^= red_s[tid * NWE + w2];
        p.crc_part[((size_t)s * p.n_slots + K + tid) * p.n_seg + seg] = u;
      }
      role_barrier<NT>();
    }
  } else {
    // =========================== checksum role ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kBswRegsC));
    const int ctid = tid - NE;
    // the 64-byte piece of slot i: q = ctid + NC*i -> (shard q / NE, piece q % NE)
    auto half = [&](auto full_tag, const uint8_t* sbase, size_t tile_start, int hidx, uint32_t (&w)[8]) {
      constexpr bool FULL = decltype(full_tag)::value;
      const int q = ctid + NC * (hidx >> 1);
      const int c = q / NE, pc = q - c * NE;
      const size_t col = tile_start + (size_t)pc * kBsPiece + (size_t)(hidx & 1) * 32;
      const bool live = (NE == NC || c < K) && (FULL || col < p.shard_len);
      if (live) {
        ldg256(sbase + (size_t)c * p.shard_pitch + col, w);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = 0;
      }
      if (!FULL && live && col + 32 > p.shard_len) {
        const int tail = (int)(p.shard_len - col);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rem = tail - 4 * i;
          w[i] &= rem >= 4 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u));
        }
      }
    };
    auto tile = [&](auto full_tag, const uint8_t* sbase, size_t tile_start) {
      uint32_t a[8], b[8];
      half(full_tag, sbase, tile_start, 0, a);
#pragma unroll 1
      for (int i = 0; i < SLOTS; i++) {
        half(full_tag, sbase, tile_start, 2 * i + 1, b);
        uint32_t u = fold(state_s[i * NC + ctid]);
#pragma unroll
        for (int j = 0; j < 8; j++) u = slice4(u ^ a[j]);
        if (i + 1 < SLOTS) half(full_tag, sbase, tile_start, 2 * i + 2, a);
#pragma unroll
        for (int j = 0; j < 8; j++) u = slice4(u ^ b[j]);
        state_s[i * NC + ctid] = u;
      }
    };

    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t s = item / p.n_seg, seg = item - s * p.n_seg;
      const uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
      const uint32_t T = (seg == p.n_seg - 1) ? p.tiles_last : p.tiles_per_seg;
      const size_t seg_start = (size_t)seg * seg_bytes;
      for (uint32_t t = 0; t < T; t++) {
        const size_t tile_start = seg_start + (size_t)t * TILE;
        if (tile_start + TILE <= p.shard_len) tile(std::true_type{}, sbase, tile_start);
        else tile(std::false_type{}, sbase, tile_start);
#ifndef CUBEEC_BSW_NOTILEBAR
        role_barrier<NT>();
#endif
      }
#pragma unroll 1
      for (int i = 0; i < SLOTS; i++) {
        const int q = ctid + NC * i;
        const int c = q / NE, pc = q - c * NE;
        uint32_t u = gf32_mul_dev(state_s[i * NC + ctid], kth_s[pc], p.poly);
        state_s[i * NC + ctid] = 0;
        u = __reduce_xor_sync(0xffffffffu, u);
        if (lane == 0 && c < K) atomicXor(&red_c[c], u);
      }
      role_barrier<NT>();
      if (ctid < K) {
        p.crc_part[((size_t)s * p.n_slots + ctid) * p.n_seg + seg] = red_c[ctid];
        red_c[ctid] = 0;
      }
      role_barrier<NT>();
    }
  }
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
#ifndef CUBEEC_BSW_CONFIGS
#define CUBEEC_BSW_CONFIGS(X) X(12, 4)
#endif

template <int K, int M>
static cudaError_t bsw_prepare(bool* ok) {
  // setmaxnreg can only re-divide what the launch allocated: registers/thread (as compiled) * threads.
  // If ptxas ever reports fewer registers than the two budgets need, the .inc would wait for ever --
  // refuse the kernel instead (the caller falls back to rs_bs_kernel<crc>).
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, rs_bsw_kernel<K, M>);
  if (e != cudaSuccess) return e;
  *ok = (long)fa.numRegs * (kBswE + kBswC) >= (long)kBswRegsE * kBswE + (long)kBswRegsC * kBswC;
  if (!*ok) return cudaSuccess;
  return cudaFuncSetAttribute(rs_bsw_kernel<K, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBsSmemBytes);
}

bool bsw_supported(int k, int m) {
#define X(KK, MM)                                                   \
  if (k == KK && m == MM) {                                         \
    bool ok = false;                                                \
    return bsw_prepare<KK, MM>(&ok) == cudaSuccess && ok;           \
  }
  CUBEEC_BSW_CONFIGS(X)
#undef X
  return false;
}

cudaError_t launch_bsw(int k, int m, const BsParams& p, int grid, cudaStream_t st) {
#define X(KK, MM)                                                                  \
  if (k == KK && m == MM) {                                                        \
    rs_bsw_kernel<KK, MM><<<grid, kBswE + kBswC, kBsSmemBytes, st>>>(p);          \
    return cudaGetLastError();                                                     \
  }
  CUBEEC_BSW_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace cbe
