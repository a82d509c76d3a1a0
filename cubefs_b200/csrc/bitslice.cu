// bitslice.cu -- the bit-sliced RS encode kernel with fused CRC32 (sm_100a): design notes, the syndrome
// reconstruct kernel and the single-pass instantiations.  The kernel template itself is bs_kernel.cuh.
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
// (conflict-free) copies of the tables in shared memory, one IDP.2A (FMA pipe) to form each lookup address.
// Per thread the CRC registers advance Horner-style over the tiles; partial remainders are
// aligned and XOR-reduced once per work item and finished by crc_finalize_kernel.
#include "bs_kernel.cuh"

namespace cbe {

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
// syndrome-reconstruct kernels exist for the code modes with 2 <= m <= 4
#define CUBEEC_BSREC_CONFIGS(X) X(4, 2) X(6, 3) X(12, 4) X(20, 4) X(16, 4) X(10, 4) X(3, 3) X(4, 4) X(8, 4) X(6, 2) X(10, 2) X(5, 2)

int bs_passes(int k, int m, const uint8_t* parity_rows, int plan) {
#define X(KK, MM) \
  if (k == KK && m == MM) return bs_rows_match<KK, MM, 0>(parity_rows) ? 1 : 0;
  CUBEEC_BS_CONFIGS(X)
#undef X
  return bs_mp_passes(k, m, parity_rows, plan);
}

bool bs_rec_supported(int k, int m) {
#define X(KK, MM) \
  if (k == KK && m == MM) return true;
  CUBEEC_BSREC_CONFIGS(X)
#undef X
  return false;
}

cudaError_t launch_bs_rec(int k, int m, const BsRecParams& p, int grid, cudaStream_t st) {
#define X(KK, MM) \
  if (k == KK && m == MM) return launch_rec_cfg<KK, MM>(p, grid, st);
  CUBEEC_BSREC_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

// experiment: rolled fused kernel (see bs_kernel.cuh), non-packed, CRC of every shard
#define CUBEEC_BS_ROLLED_CONFIGS(X) X(12, 4) X(6, 2) X(10, 4)
bool bs_rolled_supported(int k, int m) {
#define X(KK, MM) \
  if (k == KK && m == MM) return true;
  CUBEEC_BS_ROLLED_CONFIGS(X)
#undef X
  return false;
}
cudaError_t launch_bs_rolled(int k, int m, const BsParams& p, int grid, cudaStream_t st) {
  if (p.packed_pps) return cudaErrorInvalidValue;
#define X(KK, MM)                                                                                             \
  if (k == KK && m == MM) {                                                                                   \
    auto kern = rs_bs_kernel<KK, MM, 0, 1, false, false, true>;                                               \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBsSmemBytes); \
    if (e != cudaSuccess) return e;                                                                           \
    kern<<<grid, kBsThreads, kBsSmemBytes, st>>>(p);                                                          \
    return cudaGetLastError();                                                                                \
  }
  CUBEEC_BS_ROLLED_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t launch_bs(int k, int m, int pass, const BsParams& p, int crc, bool verify, int grid, cudaStream_t st) {
  // single-pass codes serve every variant; the m > 4 codes have a pass plan per variant (bitslice_mp.cu)
  // outputs-only CRC (4) exists for the LRC local-stripe codes
#define X(KK, MM) \
  if (k == KK && m == MM)   \
    return pass == 0 ? bs_launch_cfg<KK, MM, 0, (MM == 1 || (KK == 4 && MM == 3)) ? 7 : 3>(p, crc, verify, grid, st) : cudaErrorInvalidValue;
  CUBEEC_BS_CONFIGS(X)
#undef X
  return launch_bs_mp(k, m, pass, p, crc, verify, grid, st);
}

}  // namespace cbe
