// bitslice_syn.cu -- bit-sliced *syndrome* reconstruct, second generation: flat work split, explicit load ring,
// per-warp solve tables.  Serves reconstruct batches that mix erasure patterns (BASELINE config C3: 3 random
// erasures per stripe), where no kernel can be specialised per pattern (csrc/jit.cu does that for single-pattern
// batches).
//
// Algebra (see RecPattern in kernels.cuh; identical results to the reference's inverse of the first k present rows,
// RS/reedsolomon.go:1453-1552): with e_d data shards missing, the k shards the reference decodes from are every present
// data shard plus the first e_d present parity rows R.  With the FIXED encode network of the code,
//   S_r = P_r ^ sum_{c present} M[r][c] D_c = sum_{c missing} M[r][c] D_c                (r in R)
// are the syndromes, the missing data follows through the e_d x e_d inverse, a missing parity row p as
//   P_p = T_p ^ sum_{c missing} M[p][c] D_c,  T_p = sum_{c present} M[p][c] D_c           (same network pass).
// Per 32-byte column: k loads, k bit transposes, the generated XOR networks of the present data shards, e
// back-transposes, then the small solve as ONE shared-memory lookup per syndrome byte (entries pack the products for
// up to 4 outputs).
//
// Why a second generation (profiles/r02_prof_bsrec.txt, the round-1 kernel at 0.49 of HBM): issue slots 38 % busy,
// long_scoreboard 5.7 warps per issue, DRAM 40 % busy -- its conditional loads were not running ahead of the coding, and
// every stripe cost a CTA-wide barrier + table rebuild.  Here: flat split (a warp owns a contiguous run of 1 KiB units,
// no barrier in the main loop), shards staged in shared memory by cp.async three slots ahead (and for the next unit
// before the solve stage of this one) with counted waits, and solve tables private to the warp (2 copies, rebuilt by the
// warp alone when its stripe changes).
#include <type_traits>

#include "bs_net_gen.cuh"
#include "kernels.cuh"
#include "bs_device.cuh"

namespace cbe {

using namespace bsdev;

namespace {

constexpr int kSynThreads = 512;
constexpr int kSynUnit = 1024;                 // bytes of a shard per unit: one 32-byte column per lane
constexpr int kSynRing = 4;                    // staged shards per thread (power of two)
constexpr int kSynTabBytes = 4 * 256 * 2 * 4;  // per warp: [4 syndromes][256][2 copies] u32
constexpr size_t kSynStageBytes = (size_t)kSynRing * 2 * kSynThreads * 16;   // [slot][16-byte half][thread]
constexpr size_t kSynSmemBytes = 1024 + (size_t)(kSynThreads / 32) * kSynTabBytes + kSynStageBytes;

// The loads of the ring do not go through registers: ptxas puts every 256-bit LDG of the ring on ONE scoreboard and
// the first use of the oldest load then also waits for the load issued a few instructions earlier (decoded from the
// SASS control words; profiles/r02_prof_bssyn_ring4.txt: 46 % of all warp samples sit on those first uses).  cp.async
// (LDGSTS) copies global -> shared without a register scoreboard, and cp.async.wait_group N is a COUNTED wait: "all
// but the N youngest groups have landed" -- exactly the semantics a software pipeline needs.  Every thread copies and
// later reads only its own 32 bytes, so no barrier is involved.
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}

template <int I, int N>
struct SFor {
  template <class F>
  static __device__ __forceinline__ void run(F&& f) {
    if constexpr (I < N) {
      f(std::integral_constant<int, I>{});
      SFor<I + 1, N>::run(f);
    }
  }
};

// (A block-wide barrier at the top of every unit, which gives the long fused-CRC kernels 4-24 % by sharing instruction
// fetches -- bs_flat.cuh -- does nothing here: 0.548 without, 0.551 with; the warps of a block work on different
// erasure patterns and leave the common path at once.)
template <int K, int M>
__global__ void __launch_bounds__(kSynThreads, 1) rs_bssyn_kernel(const BsRecParams p) {
  using Net = BsNet<K, M>;
  constexpr int N = K + M, RD = kSynRing, NW = kSynThreads / 32;
  extern __shared__ __align__(1024) uint8_t smem[];
  GfDeviceTables* gf_s = reinterpret_cast<GfDeviceTables*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* tab = reinterpret_cast<uint32_t*>(smem + 1024 + (size_t)warp * kSynTabBytes);
  const uint32_t tab_lane = smem_addr(tab) + (uint32_t)((lane & 1) * 4);
  for (int i = tid; i < (int)(sizeof(GfDeviceTables) / 4); i += kSynThreads)
    reinterpret_cast<uint32_t*>(gf_s)[i] = reinterpret_cast<const uint32_t*>(p.gf)[i];
  __syncthreads();

  const uint64_t U = p.total_units, GW = (uint64_t)gridDim.x * NW, gw = (uint64_t)blockIdx.x * NW + warp;
  const uint64_t u_lo = gw * U / GW, u_hi = (gw + 1) * U / GW;
  if (u_lo >= u_hi) return;
  const uint32_t wt = p.units_per_shard;
  uint32_t s = (uint32_t)(u_lo / wt), t = (uint32_t)(u_lo - (uint64_t)s * wt);

  static_assert((RD & (RD - 1)) == 0, "ring size must be a power of two");
  // staging area of this thread: slot r, half h at stage_base + ((r * 2 + h) * threads + tid) * 16
  const uint32_t stage_base = smem_addr(smem + 1024 + (size_t)NW * kSynTabBytes) + (uint32_t)tid * 16u;
  auto stage_addr = [&](uint32_t r, uint32_t h) -> uint32_t { return stage_base + ((r * 2u + h) * (uint32_t)kSynThreads) * 16u; };
  uint32_t rbase = 0;   // ring position of slot 0 of the current unit (advances by N per unit, modulo RD)

  // pattern of the current stripe (warp-uniform values)
  uint32_t cur_pat = 0xFFFFFFFFu, want = 0, syn_mask = 0, t_mask = 0, n_out = 0, out_slots = 0, out_prows = 0;
  bool primed = false;   // ring[0 .. RD-2] already hold (or are receiving) slots 0 .. RD-2 of this unit

  uint32_t pat_id = 0, pat_s = 0xFFFFFFFFu;
  for (uint64_t u = u_lo; u < u_hi; u++) {
    if (s != pat_s) {   // one (dependent) global load per stripe, not per unit
      pat_s = s;
      pat_id = p.pattern_of_stripe ? p.pattern_of_stripe[s] : 0u;
    }
    if (pat_id != cur_pat) {
      cur_pat = pat_id;
      primed = false;
      const RecPattern* rp = p.patterns + pat_id;
      const uint32_t dm = rp->data_mask;
      syn_mask = rp->syn_mask;
      t_mask = rp->t_mask;
      n_out = rp->n_out;
      want = (dm & ((1u << K) - 1u)) | (syn_mask << K);   // slots read: present data + syndrome parity rows
      out_slots = (uint32_t)rp->out_slot[0] | ((uint32_t)rp->out_slot[1] << 8) | ((uint32_t)rp->out_slot[2] << 16) | ((uint32_t)rp->out_slot[3] << 24);
      out_prows = (uint32_t)rp->out_prow[0] | ((uint32_t)rp->out_prow[1] << 8) | ((uint32_t)rp->out_prow[2] << 16) | ((uint32_t)rp->out_prow[3] << 24);
      // solve tables of this warp: entry(i, v) = sum_j (coef[j][i] * v) << 8j, two copies (lane parity)
      __syncwarp();
      const uint32_t n_syn = rp->n_syn;
      for (int idx = lane; idx < 4 * 256; idx += 32) {
        const int i = idx >> 8, v = idx & 255;
        uint32_t e = 0;
        if (v && (uint32_t)i < n_syn) {
          const int lv = gf_s->log[v];
          for (uint32_t j = 0; j < n_out; j++) {
            const int co = rp->coef[j][i];
            if (co) e |= (uint32_t)gf_s->exp[gf_s->log[co] + lv] << (8 * j);
          }
        }
        tab[idx * 2] = e;
        tab[idx * 2 + 1] = e;
      }
      __syncwarp();
    }
    // the unit after this one
    const bool wrap = t + 1 == wt;
    const uint32_t s2 = wrap ? s + 1 : s, t2 = wrap ? 0u : t + 1;
    const bool next_same = (u + 1 < u_hi) && !wrap;   // same stripe, hence same pattern: its first loads can be issued early

    if (n_out != 0) {
      uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
      const uint32_t col = t * kSynUnit + (uint32_t)lane * 32u;
      const bool live = col < p.shard_len;
      const bool full = (t + 1) * kSynUnit <= p.shard_len;   // warp-uniform
      uint32_t msk[8];
#pragma unroll
      for (int i = 0; i < 8; i++) msk[i] = 0xffffffffu;
      if (!full) {
        const int tail = (live && col + 32 > p.shard_len) ? (int)(p.shard_len - col) : 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rem = tail - 4 * i;
          msk[i] = !live ? 0u : ((tail == 0 || rem >= 4) ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << (8 * rem)) - 1u)));
        }
      }
      const uint8_t* src = sbase + col;
      const uint32_t ncol = t2 * kSynUnit + (uint32_t)lane * 32u;
      const uint8_t* nsrc = p.base + (size_t)s2 * p.stripe_pitch + ncol;
      const bool nlive = next_same && ncol < p.shard_len;
      // issue(c, base, live): copy slot c of the unit at `base` into its ring position, as its own cp.async group
      auto issue = [&](const int c, const uint8_t* base, const bool on, const uint32_t rb) {
        if (((want >> c) & 1u) && on) {
          const uint8_t* g = base + (size_t)c * p.shard_pitch;
          const uint32_t r = (rb + (uint32_t)c) & (RD - 1);
          cp_async16(stage_addr(r, 0), g);
          cp_async16(stage_addr(r, 1), g + 16);
        }
        cp_async_commit();   // a group per slot, empty or not: the wait below counts groups
      };
      if (!primed) {
#pragma unroll
        for (int c = 0; c < RD - 1; c++) issue(c, src, live, rbase);
      }
      uint32_t acc[8 * M];
#pragma unroll
      for (int i = 0; i < 8 * M; i++) acc[i] = 0;
      uint32_t tmp[32];   // XOR combinations of the network part functions

      SFor<0, N>::run([&](auto cc) {
        constexpr int c = decltype(cc)::value;
        // copy RD-1 slots ahead; past the last slot: slots 0 .. RD-2 of the next unit
        if constexpr (c + RD - 1 < N) issue(c + RD - 1, src, live, rbase);
        else issue(c + RD - 1 - N, nsrc, nlive, rbase + (uint32_t)N);
        cp_async_wait<RD - 1>();   // everything but the RD-1 youngest groups has landed: slot c is in shared memory
        if ((want >> c) & 1u) {    // warp-uniform
          uint32_t w[8];
          const uint32_t r = (rbase + (uint32_t)c) & (RD - 1);
          lds128(stage_addr(r, 0), w[0], w[1], w[2], w[3]);
          lds128(stage_addr(r, 1), w[4], w[5], w[6], w[7]);
          if (!full) {
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] &= msk[i];
          }
          bit_transpose8(w);
          if constexpr (c < K) {
            SFor<0, Net::kParts>::run([&](auto jc) { Net::template part<c, decltype(jc)::value>(w, acc, tmp); });
          } else {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[(c - K) * 8 + i] ^= w[i];   // S_r = P_r ^ (network row r)
          }
        }
      });
      rbase = (rbase + (uint32_t)N) & (RD - 1);
      primed = next_same;

      // back to bytes, in place, for the rows that are used (syndromes and T_p of missing parity)
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
      // solve: packed[b] = sum over syndromes of entry(i, byte b of the syndrome); byte j of packed[b] = output j
      uint32_t packed[32];
#pragma unroll
      for (int i = 0; i < 32; i++) packed[i] = 0;
      uint32_t si = 0;
#pragma unroll
      for (int r = 0; r < M; r++) {
        if ((syn_mask >> r) & 1u) {
          const uint32_t tb = tab_lane + si * (256u * 8u);
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const uint32_t w = acc[r * 8 + q];
            packed[q * 4 + 0] ^= lds32(byte_madd<0>(w, 8u, 8u << 16, tb));
            packed[q * 4 + 1] ^= lds32(byte_madd<1>(w, 8u, 8u << 16, tb));
            packed[q * 4 + 2] ^= lds32(byte_madd<2>(w, 8u, 8u << 16, tb));
            packed[q * 4 + 3] ^= lds32(byte_madd<3>(w, 8u, 8u << 16, tb));
          }
          si++;
        }
      }
      // unpack output j, add T_p for regenerated parity, store 32 bytes
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if ((uint32_t)j < n_out) {
          uint32_t o[8];
          constexpr uint32_t kSel[4] = {0x0040u, 0x0051u, 0x0062u, 0x0073u};   // byte j of (a, b) -> bytes 0, 1
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const uint32_t lo = prmt(packed[q * 4 + 0], packed[q * 4 + 1], kSel[j]);
            const uint32_t hi = prmt(packed[q * 4 + 2], packed[q * 4 + 3], kSel[j]);
            o[q] = prmt(lo, hi, 0x5410u);
          }
          const uint32_t prow = (out_prows >> (8 * j)) & 255u;
#pragma unroll
          for (int r = 0; r < M; r++) {
            if (prow == (uint32_t)r) {
#pragma unroll
              for (int q = 0; q < 8; q++) o[q] ^= acc[r * 8 + q];
            }
          }
          if (live) stg256(sbase + (size_t)((out_slots >> (8 * j)) & 255u) * p.shard_pitch + col, o);
        }
      }
    }
    else {
      primed = false;
    }
    if (!primed) cp_async_wait<0>();   // nothing of this warp stays in flight across a pattern change / the end of its run
    s = s2;
    t = t2;
  }
}

template <int K, int M>
cudaError_t launch_syn_cfg(const BsRecParams& p, int grid, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rs_bssyn_kernel<K, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSynSmemBytes);
  if (e != cudaSuccess) return e;
  rs_bssyn_kernel<K, M><<<grid, kSynThreads, kSynSmemBytes, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

// syndrome-reconstruct kernels exist for the single-pass code modes with 2 <= m <= 4 (same list as bitslice.cu)
#define CUBEEC_BSSYN_CONFIGS(X) X(4, 2) X(6, 3) X(12, 4) X(20, 4) X(16, 4) X(10, 4) X(3, 3) X(4, 4) X(8, 4) X(6, 2) X(10, 2) X(5, 2)

bool bs_syn_supported(int k, int m) {
#define X(KK, MM) \
  if (k == KK && m == MM) return true;
  CUBEEC_BSSYN_CONFIGS(X)
#undef X
  return false;
}

uint32_t bs_syn_units_per_shard(size_t shard_len) { return (uint32_t)((shard_len + kSynUnit - 1) / kSynUnit); }

cudaError_t launch_bs_syn(int k, int m, const BsRecParams& p, int grid, cudaStream_t st) {
#define X(KK, MM) \
  if (k == KK && m == MM) return launch_syn_cfg<KK, MM>(p, grid, st);
  CUBEEC_BSSYN_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

}  // namespace cbe
