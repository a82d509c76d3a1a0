// crc_flat.cu -- stand-alone CRC32 of byte ranges at (close to) HBM speed.
//
// The ranges are the units of cubeec_dev_crc32 / the crc32block payload blocks (blobstore/common/crc32block/block.go:38-49:
// 64 KiB - 4 bytes, starting 4 bytes into every block of the framed image), whole shards (blobnode shard CRC,
// blobstore/blobnode/core/storage/datafile.go, and the replica mode of the access layer, stream_put.go:265-269 with no parity),
// i.e. every CRC the path takes that is NOT fused into a coding pass.  The first version (crc_range_kernel, kernels.cu:
// one CTA per range, 16-byte pieces, two table lookups chains from ONE table copy) reached 0.19 of the measured HBM peak:
// bank conflicts on the single copy and 8 lookups per 16 bytes.  This kernel takes the lookup scheme of the fused
// encode + CRC kernel (bs_flat.cuh): slicing-by-4 from 32 lane-private table copies (conflict-free, 4 lookups per word),
// a lane checksums 64 contiguous bytes of a 2 KiB warp-tile, Horner across the tiles of a run through the fold tables,
// lanes aligned and reduced once per range.  Work split: the tiles of all ranges are numbered range-major and cut into
// one contiguous run per warp (runs differ by at most one tile), so 32768 blocks of 64 KiB load the SMs as evenly as
// 2048 shards of 1 MiB.  Tiles are 2 KiB windows of the BUFFER (not of the range), so every load is an aligned 256-bit
// (or 128-bit) load whatever the alignment of the range; bytes outside [start, end) are masked to zero, which the linear
// (zero-init) remainder ignores in front and which crc_flat_finish_kernel divides out behind.
//
// Bound: 68 lookups per 64 bytes and lane, 148 SMs x 128 B/clk of shared-memory bandwidth = 8.5 TB/s at 1.9 GHz, above
// the 6.6 TB/s HBM peak.  Measured on B200 (tools/crc_speed.py, device-resident, fraction of the measured HBM peak):
// 1 MiB shards 0.69, C2 shards (349,526 B) 0.65, the same with their 65,532-byte block checksums 0.52, 4 KiB buffers
// 0.58; crc_range_kernel on the same inputs 0.15 / 0.12 / 0.12 / 0.013.
#include "bs_device.cuh"
#include "kernels.cuh"

namespace cbe {

using namespace bsdev;

namespace {

constexpr int kCfThreads = kCrcFlatThreads;
constexpr uint32_t kCfTile = kBsfUnitBytes;   // 2 KiB: the unit the fold tables and klane are built for
constexpr uint32_t kCfPiece = kBsPiece;       // 64 bytes per lane

__device__ __forceinline__ uint4 ldg128_stream(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

struct Range {
  const uint8_t* buf;
  uint32_t a, e;      // byte range [a, e) of the buffer
  uint32_t t0;        // first tile (2 KiB window of the buffer) that the range touches
};

__device__ __forceinline__ Range locate_range(const CrcFlatParams& p, uint32_t r) {
  Range x;
  const uint32_t b = r / p.units_per_buffer, u = r - b * p.units_per_buffer;
  x.buf = p.base + (size_t)b * p.pitch;
  const uint64_t a = (uint64_t)p.offset + (uint64_t)u * (p.stride ? p.stride : p.block);
  const uint64_t e = a + p.block < p.len ? a + p.block : p.len;
  x.a = (uint32_t)(a < e ? a : e);
  x.e = (uint32_t)e;
  x.t0 = x.a / kCfTile;
  return x;
}

}  // namespace

// V8: buffers start on 32-byte boundaries -> 256-bit loads (whole 32-byte sectors per lane and request; with 128-bit loads
// two requests share every sector and L2 serves it twice: measured 0.45 of the HBM peak against 0.69 with them).
template <bool V8>
__global__ void __launch_bounds__(kCfThreads, 1) crc_flat_kernel(const __grid_constant__ CrcFlatParams p) {
  constexpr int NT = kCfThreads, NW = NT / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- shared memory: [mbarrier | fold tables (FC copies) | slice image]   (as bs_flat.cuh)
  constexpr int FC = kCrcFlatFoldCopies;
  constexpr size_t FTB = 256 * FC * 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* fold_s = reinterpret_cast<uint32_t*>(smem + 128);
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
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((uint32_t)kBsSliceImageBytes) : "memory");
      for (int h = 0; h < 2; h++)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
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
  const uint32_t klane = p.klane[lane];

  auto slice4 = [&](uint32_t y) -> uint32_t {
    const uint32_t a0 = byte_madd<0>(y, 256u, 256u << 16, lane_base);
    const uint32_t a1 = byte_madd<1>(y, 256u, 256u << 16, lane_base);
    const uint32_t a2 = byte_madd<2>(y, 256u, 256u << 16, lane_base);
    const uint32_t a3 = byte_madd<3>(y, 256u, 256u << 16, lane_base);
    return lds32_off<65536 + 128>(a0) ^ lds32_off<65536>(a1) ^ lds32_off<128>(a2) ^ lds32_off<0>(a3);
  };
  auto fold = [&](uint32_t u) -> uint32_t {
    constexpr uint32_t ST = FC * 4;
    return lds32(byte_madd<0>(u, ST, ST << 16, fold_lane + 0 * 256 * ST)) ^ lds32(byte_madd<1>(u, ST, ST << 16, fold_lane + 1 * 256 * ST)) ^
           lds32(byte_madd<2>(u, ST, ST << 16, fold_lane + 2 * 256 * ST)) ^ lds32(byte_madd<3>(u, ST, ST << 16, fold_lane + 3 * 256 * ST));
  };

  const uint64_t G = p.total_tiles, GW = (uint64_t)gridDim.x * NW, gw = (uint64_t)blockIdx.x * NW + warp;
  const uint64_t lo = gw * G / GW, hi = (gw + 1) * G / GW;
  if (lo >= hi) return;
  const uint32_t T = p.tiles_per_range;

  // the 64 bytes of this lane in tile t of range x: four 128-bit loads, chunks wholly outside [a, e) not touched
  auto load_tile = [&](const Range& x, uint32_t t, uint32_t (&w)[16]) {
    const uint32_t col = (x.t0 + t) * kCfTile + (uint32_t)lane * kCfPiece;   // (a range ends below 4 GiB - 16)
    if constexpr (V8) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const uint32_t cj = col + 32u * j;
        uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cj < x.e && cj + 32u > x.a) ldg256(x.buf + cj, v);   // (reads up to 31 bytes past e: inside the pitch, see launch_crc_flat)
#pragma unroll
        for (int i = 0; i < 8; i++) w[8 * j + i] = v[i];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t cj = col + 16u * j;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (cj < x.e && cj + 16u > x.a) v = ldg128_stream(x.buf + cj);
        w[4 * j + 0] = v.x;
        w[4 * j + 1] = v.y;
        w[4 * j + 2] = v.z;
        w[4 * j + 3] = v.w;
      }
    }
  };

  uint32_t r = (uint32_t)(lo / T), t = (uint32_t)(lo - (uint64_t)r * T);
  Range x = locate_range(p, r);
  uint32_t cur[16], nxt[16];
  load_tile(x, t, cur);
  uint32_t u = 0;
  for (uint64_t g = lo; g < hi; g++) {
    const bool last_of_range = t + 1 == T, more = g + 1 < hi;
    const uint32_t r2 = last_of_range ? r + 1 : r, t2 = last_of_range ? 0u : t + 1;
    Range x2 = x;
    if (more) {
      if (last_of_range) x2 = locate_range(p, r2);
      load_tile(x2, t2, nxt);
    }
    // bytes of this tile outside [a, e) count as zero (warp-uniform test: the whole tile inside the range?)
    const uint32_t tile_lo = (x.t0 + t) * kCfTile;
    if (tile_lo < x.a || (uint64_t)tile_lo + kCfTile > x.e) {
      const int64_t w0 = (int64_t)tile_lo + (int64_t)lane * kCfPiece;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        int64_t lo_cut = (int64_t)x.a - (w0 + 4 * i), hi_cut = (int64_t)x.e - (w0 + 4 * i);
        lo_cut = lo_cut < 0 ? 0 : (lo_cut > 4 ? 4 : lo_cut);
        hi_cut = hi_cut < 0 ? 0 : (hi_cut > 4 ? 4 : hi_cut);
        const uint32_t keep_hi = hi_cut >= 4 ? 0xffffffffu : ((1u << (8 * (int)hi_cut)) - 1u);
        const uint32_t drop_lo = lo_cut >= 4 ? 0xffffffffu : ((1u << (8 * (int)lo_cut)) - 1u);
        cur[i] &= keep_hi & ~drop_lo;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) u = slice4(u ^ cur[i]);
    if (last_of_range || !more) {
      // align the lanes' remainders to the end of this tile, reduce, move behind the remaining tiles of the range
      uint32_t v = gf32_mul_dev(u, klane, p.poly);
      v = __reduce_xor_sync(0xffffffffu, v);
      if (lane == 0) {
        uint32_t n = T - 1 - t;
        for (int i = 0; n; i++, n >>= 1)
          if (n & 1) v = gf32_mul_dev(v, p.x_unit_pow[i], p.poly);
        atomicXor(&p.out[r], v);
      }
      u = 0;
    } else {
      u = fold(u);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) cur[i] = nxt[i];
    x = x2;
    r = r2;
    t = t2;
  }
}

// out[r] holds the zero-init remainder of range r as if it ended at the end of its last tile:
//   crc = ~( out * x^(-8 z) ^ 0xFFFFFFFF * x^(8 n) ),  z = bytes between the end of the range and that tile end, n = its length
__global__ void crc_flat_finish_kernel(const __grid_constant__ CrcFlatParams p) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.n_buffers * p.units_per_buffer) return;
  const Range x = locate_range(p, r);
  uint64_t z = ((uint64_t)x.t0 + p.tiles_per_range) * kCfTile - x.e;
  uint32_t v = p.out[r];
  for (int i = 0; z; i++, z >>= 1)
    if (z & 1) v = gf32_mul_dev(v, p.x_neg_pow[i], p.poly);
  const uint32_t n = x.e - x.a;
  uint32_t init_term = p.init_full;
  if (n != p.block) {
    // 0xFFFFFFFF * x^(8 n) by square and multiply (the short last unit of a buffer)
    uint32_t result = 0x80000000u, base = 0x40000000u;
    for (uint64_t q = 8ull * n; q; q >>= 1) {
      if (q & 1) result = gf32_mul_dev(result, base, p.poly);
      base = gf32_mul_dev(base, base, p.poly);
    }
    init_term = gf32_mul_dev(0xFFFFFFFFu, result, p.poly);
  }
  p.out[r] = ~(v ^ init_term);
}

// p.out must be zero when crc_flat_kernel starts (the engine clears it on the same stream).
cudaError_t launch_crc_flat(const CrcFlatParams& p, int sm_count, cudaStream_t stream) {
  // 256-bit loads need 32-byte aligned buffers, and a pitch that covers the 32-byte chunk holding the last byte
  const bool v8 = ((uintptr_t)p.base % 32 == 0) && (p.pitch % 32 == 0) && (((size_t)p.len + 31) / 32 * 32 <= p.pitch);
  auto kern = v8 ? crc_flat_kernel<true> : crc_flat_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCrcFlatSmemBytes);
  if (e != cudaSuccess) return e;
  const uint64_t nw = kCfThreads / 32;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((p.total_tiles + nw - 1) / nw, (uint64_t)sm_count));
  kern<<<grid, kCfThreads, kCrcFlatSmemBytes, stream>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const uint32_t n = p.n_buffers * p.units_per_buffer;
  crc_flat_finish_kernel<<<(n + 127) / 128, 128, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace cbe
