// kernels.cu -- sm_100a kernels of libcubeec: generic table-driven GF(2^8) coding kernel
// (encode / reconstruct / verify for any k, any coefficients) with an optional fused CRC32 pass,
// plus the CRC finalize / range kernels.
//
// Replaces the CPU SIMD kernels of klauspost/reedsolomon (mulAvxTwo_* / mulGFNI_*,
// RS/galois_gen_amd64.s) and Go's hash/crc32 on the BlobStore shard path.  No tensor cores:
// this is byte-field arithmetic, bounded by HBM bandwidth, LSU (shared-memory lookups) and the
// integer ALU pipe.  See DESIGN.md for the roofline of each kernel.
#include <algorithm>

#include "kernels.cuh"

namespace cbe {

// ------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// table lookup address = base + STRIDE * (byte B of w) in ONE instruction on the FMA pipe: IDP.2A (dp2a)
// multiplies the 16-bit halves of its first operand with two bytes of w (.lo: bytes 0,1; .hi: 2,3).
//
// Product-table layout: entry (input c, byte value v, copy g) at  c*256*4R + v*4R + g*4  (R copies, copy
// = lane % R).  The bank is then (v*R + g) mod 32 = g + R*(v mod 32/R): lanes that share a copy collide
// only when their bytes agree modulo 32/R (R = 16: the low bit) -- 1.5 wavefronts per lookup instead of
// the 2.0 of a layout whose bank does not depend on v.
template <int B, uint32_t STRIDE>
__device__ __forceinline__ uint32_t row_addr_b(uint32_t w, uint32_t base) {
  uint32_t d;
  if (B == 0) asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(STRIDE), "r"(w), "r"(base));
  if (B == 1) asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(STRIDE << 16), "r"(w), "r"(base));
  if (B == 2) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(STRIDE), "r"(w), "r"(base));
  if (B == 3) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(STRIDE << 16), "r"(w), "r"(base));
  return d;
}

// zero the bytes at positions >= n (0 < n < 16) of a 16-byte piece
__device__ __forceinline__ uint4 keep_head(uint4 d, int n) {
  uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int rem = n - 4 * i;
    if (rem <= 0) w[i] = 0;
    else if (rem < 4) w[i] &= (1u << (8 * rem)) - 1u;
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
// zero the bytes at positions < n (0 < n < 16)
__device__ __forceinline__ uint4 drop_head(uint4 d, int n) {
  uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int rem = n - 4 * i;
    if (rem >= 4) w[i] = 0;
    else if (rem > 0) w[i] &= ~((1u << (8 * rem)) - 1u);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// a(x)*b(x) mod P in the reflected representation (bit 31 = x^0)
__device__ __forceinline__ uint32_t gf32_mul(uint32_t a, uint32_t b, uint32_t poly) {
  uint32_t r = 0;
#pragma unroll 8
  for (int i = 0; i < 32; i++) {
    r ^= a & (uint32_t)((int32_t)b >> 31);
    b <<= 1;
    a = (a >> 1) ^ (poly & (0u - (a & 1u)));
  }
  return r;
}
// x^n mod P for n >= 0
__device__ uint32_t gf32_xpow(uint64_t n, uint32_t poly) {
  uint32_t result = 0x80000000u, base = 0x40000000u;
  while (n) {
    if (n & 1) result = gf32_mul(result, base, poly);
    base = gf32_mul(base, base, poly);
    n >>= 1;
  }
  return result;
}

// register after the 16 bytes of `d`, starting from register 0 (slicing-by-4, 4 dependent steps)
__device__ __forceinline__ uint32_t crc_piece(uint4 d, const uint32_t* __restrict__ sl) {
  uint32_t c = d.x;
  c = sl[768 + (c & 0xff)] ^ sl[512 + ((c >> 8) & 0xff)] ^ sl[256 + ((c >> 16) & 0xff)] ^ sl[c >> 24];
  c ^= d.y;
  c = sl[768 + (c & 0xff)] ^ sl[512 + ((c >> 8) & 0xff)] ^ sl[256 + ((c >> 16) & 0xff)] ^ sl[c >> 24];
  c ^= d.z;
  c = sl[768 + (c & 0xff)] ^ sl[512 + ((c >> 8) & 0xff)] ^ sl[256 + ((c >> 16) & 0xff)] ^ sl[c >> 24];
  c ^= d.w;
  c = sl[768 + (c & 0xff)] ^ sl[512 + ((c >> 8) & 0xff)] ^ sl[256 + ((c >> 16) & 0xff)] ^ sl[c >> 24];
  return c;
}
// register * constant by 4 byte-indexed lookups
__device__ __forceinline__ uint32_t crc_constmul(uint32_t u, const uint32_t* __restrict__ mt) {
  return mt[u & 0xff] ^ mt[256 + ((u >> 8) & 0xff)] ^ mt[512 + ((u >> 16) & 0xff)] ^ mt[768 + (u >> 24)];
}

// ------------------------------------------------------------------------------------------
// Generic table kernel.
//
// Work item = (stripe, segment); a segment is tiles_per_seg tiles of NT*16 bytes of every shard.
// Persistent CTAs walk the items; per item, thread `tid` owns the 16-byte column
// [seg_start + t*TILE + 16*tid, +16) of every shard for t = 0..T-1:
//   * loads the column of each input shard with one coalesced 128-bit access,
//   * looks every byte up in a per-input 256-entry shared-memory table whose 32-bit entries
//     pack the products for up to 4 outputs (one lookup per input byte, XOR accumulate),
//   * unpacks the 16 accumulators into 4 output pieces and stores (or compares) them,
//   * optionally advances a per-shard CRC register: piece remainder by slicing-by-4, Horner
//     step (multiply by x^(8*TILE)) between pieces, so bytes are checksummed while in registers.
// Tables are replicated R times (copy = lane % R) to cut bank conflicts; R is the largest
// power of two <= 16 that fits next to the CRC state in shared memory.  Layout: one 256-byte row
// per byte value v holding 64/R shards x R copies, 64 KiB per group of 64/R shards, so a lookup
// address is  v*256 + (per-shard, per-lane base): one PRMT (byte extract, ALU pipe), one
// multiply-add (FMA pipe), one LDS.
// ------------------------------------------------------------------------------------------
template <int R, bool CRC>
__global__ void __launch_bounds__(kTabThreads, 1) rs_tab_kernel(const TabParams p, const int n_crc_slots) {
  constexpr int NT = kTabThreads;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(128) uint8_t smem[];

  // ---- shared memory carve-up ----
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);                       // 16 B
  GfDeviceTables* gf_s = reinterpret_cast<GfDeviceTables*>(smem + 16);     // 768 B
  Pattern* pat_s = reinterpret_cast<Pattern*>(smem + 16 + sizeof(GfDeviceTables));
  size_t off = 16 + sizeof(GfDeviceTables) + sizeof(Pattern);
  uint32_t* crc_sl = nullptr;   // slice[4][256]
  uint32_t* crc_sh = nullptr;   // shift_tile[4][256]
  uint32_t* crc_kt = nullptr;   // kthread[NT]
  uint32_t* crc_st = nullptr;   // [n_crc_slots][NT]
  uint32_t* crc_red = nullptr;  // [n_crc_slots][NW]
  if (CRC) {
    crc_sl = reinterpret_cast<uint32_t*>(smem + off);
    crc_sh = crc_sl + 1024;
    crc_kt = crc_sh + 1024;
    off += (2048 + 1024) * 4;   // slice + shift + kthread[1024] are contiguous in CrcDeviceTables
    crc_st = reinterpret_cast<uint32_t*>(smem + off);
    off += (size_t)n_crc_slots * NT * 4;
    crc_red = reinterpret_cast<uint32_t*>(smem + off);
    off += (size_t)n_crc_slots * NW * 4;
    off = (off + 127) & ~(size_t)127;
  }
  uint8_t* tab = smem + off;                                   // [n_in][256 values][R copies] u32 (see row_addr_b)
  const uint32_t tab_s = smem_u32(tab);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane % R;

  // ---- stage the constant tables with TMA bulk copies ----
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    uint32_t bytes = sizeof(GfDeviceTables) + (CRC ? (2048 + 1024) * 4 : 0);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(gf_s, p.gf, sizeof(GfDeviceTables), bar);
    if (CRC) bulk_g2s(crc_sl, p.crc, (2048 + 1024) * 4, bar);
  }
  mbar_wait(bar, 0);
  const uint32_t poly = CRC ? p.crc->poly : 0;

  if (CRC)
    for (int i = tid; i < n_crc_slots * NT; i += NT) crc_st[i] = 0;

  uint32_t cur_pattern = 0xFFFFFFFFu;
  const uint32_t n_items = p.n_stripes * p.n_seg;
  const size_t seg_bytes = (size_t)p.tiles_per_seg * kTabTile;

  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t s = item / p.n_seg, seg = item - s * p.n_seg;
    const uint32_t pat_id = p.pattern_of_stripe ? p.pattern_of_stripe[s] : 0u;
    if (pat_id != cur_pattern) {
      __syncthreads();   // everyone finished with the old tables / pattern
      cur_pattern = pat_id;
      const uint4* src = reinterpret_cast<const uint4*>(p.patterns + pat_id);
      uint4* dst = reinterpret_cast<uint4*>(pat_s);
      for (int i = tid; i < (int)(sizeof(Pattern) / 16); i += NT) dst[i] = src[i];
      __syncthreads();
      const int n_in = pat_s->n_in, n_out = pat_s->n_out;
      for (int idx = tid; idx < n_in * 256; idx += NT) {
        const int c = idx >> 8, v = idx & 255;
        uint32_t e = 0;
        if (v) {
          const int lv = gf_s->log[v];
          for (int r = 0; r < n_out; r++) {
            const int co = pat_s->coef[r][c];
            if (co) e |= (uint32_t)gf_s->exp[gf_s->log[co] + lv] << (8 * r);
          }
        }
        uint32_t* row = reinterpret_cast<uint32_t*>(tab + (size_t)c * (1024 * R) + (size_t)v * (4 * R));
#pragma unroll
        for (int q = 0; q < R; q++) row[q] = e;
      }
      __syncthreads();
    }
    const int n_in = pat_s->n_in, n_out = pat_s->n_out;
    const bool crc_in = CRC && pat_s->crc_in;
    uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
    const uint32_t T = (seg == p.n_seg - 1) ? p.tiles_last : p.tiles_per_seg;
    const size_t seg_start = (size_t)seg * seg_bytes;
    bool bad = false;

    for (uint32_t t = 0; t < T; t++) {
      const size_t col = seg_start + (size_t)t * kTabTile + (size_t)tid * kPiece;
      const bool live = col < p.shard_len;
      const int tail = (live && col + kPiece > p.shard_len) ? (int)(p.shard_len - col) : 0;
      uint32_t acc[16];
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = 0;

      // Inputs are consumed in chunks of CH shards; the next chunk's loads are issued before the
      // current chunk is coded (2*CH 16-byte loads per thread in flight).
      constexpr int CH = 6;
      uint4 cur[CH], nxt[CH];
      auto fetch = [&](uint4 (&d)[CH], const int c0) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
          d[i] = make_uint4(0, 0, 0, 0);
          if (c0 + i < n_in && live) d[i] = ldg_stream(sbase + (size_t)pat_s->in_slot[c0 + i] * p.shard_pitch + col);
        }
      };
      fetch(cur, 0);
      for (int c0 = 0; c0 < n_in; c0 += CH) {
        if (c0 + CH < n_in) fetch(nxt, c0 + CH);
#pragma unroll
        for (int i = 0; i < CH; i++) {
          if (c0 + i < n_in) {
            uint4 dd = cur[i];
            if (tail) dd = keep_head(dd, tail);
            const int cc = c0 + i;
            const uint32_t tb = tab_s + (uint32_t)cc * (1024u * R) + (uint32_t)(g * 4);
            const uint32_t w[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
              acc[q * 4 + 0] ^= lds_u32(row_addr_b<0, 4 * R>(w[q], tb));
              acc[q * 4 + 1] ^= lds_u32(row_addr_b<1, 4 * R>(w[q], tb));
              acc[q * 4 + 2] ^= lds_u32(row_addr_b<2, 4 * R>(w[q], tb));
              acc[q * 4 + 3] ^= lds_u32(row_addr_b<3, 4 * R>(w[q], tb));
            }
            if (crc_in) {
              uint32_t* st = crc_st + (size_t)cc * NT + tid;
              *st = crc_constmul(*st, crc_sh) ^ crc_piece(dd, crc_sl);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < CH; i++) cur[i] = nxt[i];
      }
      // unpack: output r, word q = byte r of acc[q*4 + 0..3]
      for (int r = 0; r < n_out; r++) {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t lo = __byte_perm(acc[q * 4 + 0], acc[q * 4 + 1], 0x0040 | r | (r << 4));
          const uint32_t hi = __byte_perm(acc[q * 4 + 2], acc[q * 4 + 3], 0x0040 | r | (r << 4));
          o[q] = __byte_perm(lo, hi, 0x5410);
        }
        const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
        uint8_t* optr = sbase + (size_t)pat_s->out_slot[r] * p.shard_pitch + col;
        if (p.mode == 0) {
          if (live) stg_stream(optr, ov);
        } else if (live) {
          uint4 e = ldg_stream(optr);
          if (tail) e = keep_head(e, tail);
          bad |= (e.x != ov.x) | (e.y != ov.y) | (e.z != ov.z) | (e.w != ov.w);
        }
        if (CRC) {
          uint32_t* st = crc_st + (size_t)((crc_in ? n_in : 0) + r) * NT + tid;
          *st = crc_constmul(*st, crc_sh) ^ crc_piece(ov, crc_sl);
        }
      }
    }

    if (p.mode == 1 && bad) atomicExch(&p.mismatch[s], 1);

    if (CRC && p.crc_part) {
      // Align every thread's partial to the end of the (virtual) segment, XOR-reduce over the CTA.
      const int nq = (crc_in ? n_in : 0) + n_out;
      const uint32_t kt = crc_kt[tid];
      for (int q = 0; q < nq; q++) {
        uint32_t u = gf32_mul(crc_st[(size_t)q * NT + tid], kt, poly);
        crc_st[(size_t)q * NT + tid] = 0;
        u = __reduce_xor_sync(0xffffffffu, u);
        if (lane == 0) crc_red[q * NW + warp] = u;
      }
      __syncthreads();
      if (tid < nq) {
        uint32_t u = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) u ^= crc_red[tid * NW + w2];
        const int slot = (crc_in && tid < n_in) ? pat_s->in_slot[tid] : pat_s->out_slot[tid - (crc_in ? n_in : 0)];
        p.crc_part[((size_t)s * p.n_slots + slot) * p.n_seg + seg] = u;
      }
      __syncthreads();
    }
  }
}


// ------------------------------------------------------------------------------------------
// Fixed-arity variant of the table kernel (no CRC): the number of inputs is a template parameter,
// so the shard loop is fully unrolled with no predicates, all loads of a column are in flight at
// once, and pairs of lookups are folded into one 3-input XOR.  Used for reconstruct / verify /
// unaligned encode of the common code modes; anything else takes rs_tab_kernel.
// ------------------------------------------------------------------------------------------
template <int R, int KF>
__global__ void __launch_bounds__(kTabThreads, 1) rs_tabk_kernel(const TabParams p) {
  constexpr int NT = kTabThreads;
  constexpr int CH = KF <= 12 ? KF : (KF + 1) / 2;     // inputs loaded per chunk
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  GfDeviceTables* gf_s = reinterpret_cast<GfDeviceTables*>(smem + 16);
  Pattern* pat_s = reinterpret_cast<Pattern*>(smem + 16 + sizeof(GfDeviceTables));
  uint8_t* tab = smem + ((16 + sizeof(GfDeviceTables) + sizeof(Pattern) + 127) & ~(size_t)127);
  const uint32_t tab_s = smem_u32(tab);
  const int tid = threadIdx.x;
  const int g = (tid & 31) % R;

  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, sizeof(GfDeviceTables));
    bulk_g2s(gf_s, p.gf, sizeof(GfDeviceTables), bar);
  }
  mbar_wait(bar, 0);

  uint32_t cur_pattern = 0xFFFFFFFFu;
  const uint32_t n_items = p.n_stripes * p.n_seg;
  const size_t seg_bytes = (size_t)p.tiles_per_seg * kTabTile;

  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t s = item / p.n_seg, seg = item - s * p.n_seg;
    const uint32_t pat_id = p.pattern_of_stripe ? p.pattern_of_stripe[s] : 0u;
    if (pat_id != cur_pattern) {
      __syncthreads();
      cur_pattern = pat_id;
      const uint4* src = reinterpret_cast<const uint4*>(p.patterns + pat_id);
      uint4* dst = reinterpret_cast<uint4*>(pat_s);
      for (int i = tid; i < (int)(sizeof(Pattern) / 16); i += NT) dst[i] = src[i];
      __syncthreads();
      const int n_out = pat_s->n_out;
      for (int idx = tid; idx < KF * 256; idx += NT) {
        const int c = idx >> 8, v = idx & 255;
        uint32_t e = 0;
        if (v) {
          const int lv = gf_s->log[v];
          for (int r = 0; r < n_out; r++) {
            const int co = pat_s->coef[r][c];
            if (co) e |= (uint32_t)gf_s->exp[gf_s->log[co] + lv] << (8 * r);
          }
        }
        uint32_t* row = reinterpret_cast<uint32_t*>(tab + (size_t)c * (1024 * R) + (size_t)v * (4 * R));
#pragma unroll
        for (int q = 0; q < R; q++) row[q] = e;
      }
      __syncthreads();
    }
    const int n_out = pat_s->n_out;
    if (n_out == 0) continue;
    uint8_t* sbase = p.base + (size_t)s * p.stripe_pitch;
    const uint32_t T = (seg == p.n_seg - 1) ? p.tiles_last : p.tiles_per_seg;
    const size_t seg_start = (size_t)seg * seg_bytes;
    // per-input source offsets are stripe constants
    size_t in_off[KF];
#pragma unroll
    for (int c = 0; c < KF; c++) in_off[c] = (size_t)pat_s->in_slot[c] * p.shard_pitch;
    bool bad = false;

    for (uint32_t t = 0; t < T; t++) {
      const size_t col = seg_start + (size_t)t * kTabTile + (size_t)tid * kPiece;
      const bool live = col < p.shard_len;
      const int tail = (live && col + kPiece > p.shard_len) ? (int)(p.shard_len - col) : 0;
      uint32_t acc[16];
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = 0;
#pragma unroll
      for (int c0 = 0; c0 < KF; c0 += CH) {
        uint4 d[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
          d[i] = make_uint4(0, 0, 0, 0);
          if (c0 + i < KF && live) d[i] = ldg_stream(sbase + in_off[c0 + i] + col);
        }
        if (tail) {
#pragma unroll
          for (int i = 0; i < CH; i++) d[i] = keep_head(d[i], tail);
        }
        // two inputs per step: acc ^= t_a ^ t_b is one LOP3
#pragma unroll
        for (int i = 0; i < CH; i += 2) {
          const int ca = c0 + i, cb = c0 + i + 1;
          if (ca < KF) {
            const bool two = (i + 1 < CH) && (cb < KF);
            const uint32_t ta = tab_s + (uint32_t)ca * (1024u * R) + (uint32_t)(g * 4);
            const uint32_t tb = tab_s + (uint32_t)(two ? cb : ca) * (1024u * R) + (uint32_t)(g * 4);
            const uint32_t wa[4] = {d[i].x, d[i].y, d[i].z, d[i].w};
            const uint4 db = d[two ? i + 1 : i];
            const uint32_t wb[4] = {db.x, db.y, db.z, db.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
              if (two) {
                acc[q * 4 + 0] ^= lds_u32(row_addr_b<0, 4 * R>(wa[q], ta)) ^ lds_u32(row_addr_b<0, 4 * R>(wb[q], tb));
                acc[q * 4 + 1] ^= lds_u32(row_addr_b<1, 4 * R>(wa[q], ta)) ^ lds_u32(row_addr_b<1, 4 * R>(wb[q], tb));
                acc[q * 4 + 2] ^= lds_u32(row_addr_b<2, 4 * R>(wa[q], ta)) ^ lds_u32(row_addr_b<2, 4 * R>(wb[q], tb));
                acc[q * 4 + 3] ^= lds_u32(row_addr_b<3, 4 * R>(wa[q], ta)) ^ lds_u32(row_addr_b<3, 4 * R>(wb[q], tb));
              } else {
                acc[q * 4 + 0] ^= lds_u32(row_addr_b<0, 4 * R>(wa[q], ta));
                acc[q * 4 + 1] ^= lds_u32(row_addr_b<1, 4 * R>(wa[q], ta));
                acc[q * 4 + 2] ^= lds_u32(row_addr_b<2, 4 * R>(wa[q], ta));
                acc[q * 4 + 3] ^= lds_u32(row_addr_b<3, 4 * R>(wa[q], ta));
              }
            }
          }
        }
      }
      for (int r = 0; r < n_out; r++) {
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const uint32_t lo = __byte_perm(acc[q * 4 + 0], acc[q * 4 + 1], 0x0040 | r | (r << 4));
          const uint32_t hi = __byte_perm(acc[q * 4 + 2], acc[q * 4 + 3], 0x0040 | r | (r << 4));
          o[q] = __byte_perm(lo, hi, 0x5410);
        }
        const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
        uint8_t* optr = sbase + (size_t)pat_s->out_slot[r] * p.shard_pitch + col;
        if (p.mode == 0) {
          if (live) stg_stream(optr, ov);
        } else if (live) {
          uint4 e = ldg_stream(optr);
          if (tail) e = keep_head(e, tail);
          bad |= (e.x != ov.x) | (e.y != ov.y) | (e.z != ov.z) | (e.w != ov.w);
        }
      }
    }
    if (p.mode == 1 && bad) atomicExch(&p.mismatch[s], 1);
  }
}

static size_t tabk_smem_bytes(int kf, int R) {
  return ((16 + sizeof(GfDeviceTables) + sizeof(Pattern) + 127) & ~(size_t)127) + (size_t)kf * 1024 * R;
}

template <int R, int KF>
static cudaError_t launch_tabk_cfg(const TabParams& p, int grid, cudaStream_t stream) {
  const size_t smem = tabk_smem_bytes(KF, R);
  cudaError_t e = cudaFuncSetAttribute(rs_tabk_kernel<R, KF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  rs_tabk_kernel<R, KF><<<grid, kTabThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

// (n_in -> replication) pairs with a fixed-arity instantiation
#define CUBEEC_TABK_CONFIGS(X) X(16, 3) X(16, 4) X(16, 5) X(16, 6) X(16, 8) X(16, 10) X(16, 12) X(8, 15) X(8, 16) X(8, 20) X(8, 24)

bool tabk_supported(int n_in) {
#define X(RR, KK) if (n_in == KK) return true;
  CUBEEC_TABK_CONFIGS(X)
#undef X
  return false;
}

cudaError_t launch_tabk(const TabParams& p, int n_in, int grid, cudaStream_t stream) {
#define X(RR, KK) if (n_in == KK) return launch_tabk_cfg<RR, KK>(p, grid, stream);
  CUBEEC_TABK_CONFIGS(X)
#undef X
  return cudaErrorInvalidValue;
}

static size_t tab_smem_bytes(int n_in, int R, bool with_crc, int n_crc_slots) {
  size_t off = 16 + sizeof(GfDeviceTables) + sizeof(Pattern);
  if (with_crc) {
    off += (2048 + 1024) * 4;
    off += (size_t)n_crc_slots * kTabThreads * 4;
    off += (size_t)n_crc_slots * (kTabThreads / 32) * 4;
    off = (off + 127) & ~(size_t)127;
  }
  off += (size_t)n_in * 1024 * R;
  return off;
}

int tab_pick_replication(int n_in, bool with_crc, int n_crc_slots, size_t smem_limit, size_t* smem_bytes) {
  for (int R = 16; R >= 1; R >>= 1) {
    size_t need = tab_smem_bytes(n_in, R, with_crc, n_crc_slots);
    if (need <= smem_limit) {
      if (smem_bytes) *smem_bytes = need;
      return R;
    }
  }
  return 0;
}

template <int R, bool CRC>
static cudaError_t tab_set_attr(size_t limit) {
  return cudaFuncSetAttribute(rs_tab_kernel<R, CRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
}

cudaError_t tab_configure(size_t* smem_limit_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int optin = 0;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  size_t limit = (size_t)optin;
#define CUBEEC_SET(R)                                                      \
  if ((e = tab_set_attr<R, false>(limit)) != cudaSuccess) return e;        \
  if ((e = tab_set_attr<R, true>(limit)) != cudaSuccess) return e;
  CUBEEC_SET(16) CUBEEC_SET(8) CUBEEC_SET(4) CUBEEC_SET(2) CUBEEC_SET(1)
#undef CUBEEC_SET
  if (smem_limit_out) *smem_limit_out = limit;
  return cudaSuccess;
}

cudaError_t launch_tab(const TabParams& p, int R, bool with_crc, int n_crc_slots, size_t smem_bytes, int grid,
                       cudaStream_t stream) {
#define CUBEEC_LAUNCH(RR)                                                                              \
  case RR:                                                                                             \
    if (with_crc) rs_tab_kernel<RR, true><<<grid, kTabThreads, smem_bytes, stream>>>(p, n_crc_slots);  \
    else rs_tab_kernel<RR, false><<<grid, kTabThreads, smem_bytes, stream>>>(p, n_crc_slots);          \
    break;
  switch (R) {
    CUBEEC_LAUNCH(16) CUBEEC_LAUNCH(8) CUBEEC_LAUNCH(4) CUBEEC_LAUNCH(2) CUBEEC_LAUNCH(1)
    default: return cudaErrorInvalidValue;
  }
#undef CUBEEC_LAUNCH
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// CRC finalize: Horner over the per-segment remainders of one (stripe, slot), then the
// init / xorout terms:  crc = ~( R(data) ^ 0xFFFFFFFF * x^(8 len) ).
// ------------------------------------------------------------------------------------------
__global__ void crc_finalize_kernel(const CrcFinalizeParams p) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= p.n_units) return;
  if (p.slot_enable && !p.slot_enable[u % p.n_slots]) return;
  const uint32_t* part = p.crc_part + (size_t)u * p.n_seg;
  uint32_t r = 0;
  for (uint32_t j = 0; j < p.n_seg; j++) {
    if (j) r = gf32_mul(r, (j == p.n_seg - 1) ? p.x_last : p.x_full, p.poly);
    r ^= part[j];
  }
  r = gf32_mul(r, p.fix, p.poly);
  p.out[u] = ~(r ^ p.init_term);
}

cudaError_t launch_crc_finalize(const CrcFinalizeParams& p, cudaStream_t stream) {
  const int nt = 128;
  crc_finalize_kernel<<<(p.n_units + nt - 1) / nt, nt, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// CRC finalize for the flat work split (bs_flat.cuh): a shard's remainder is chained from the parts the
// warps left, part j covering the units of stripe s that run (first owner + j) holds:
//   R = sum_j part_j * x^(8 * unit * units after part j);  crc = ~( R * fix ^ 0xFFFFFFFF * x^(8 len) ).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t cpf_run_lo(uint64_t g, uint64_t U, uint64_t GW) { return g * U / GW; }
// One thread per (stripe, slot, part): a batch of few, long stripes has hundreds of parts per shard (the number of
// warp runs that touch it), so the chain is not walked serially -- every part is moved to the end of its shard on its
// own (x^(8 * unit * units after it), square-and-multiply over the precomputed x_unit_pow table) and XOR-accumulated.
__global__ void crc_parts_scatter_kernel(const CrcPartsFinalizeParams p) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = (uint64_t)p.n_stripes * p.n_out * p.max_parts;
  if (idx >= total) return;
  const uint32_t j = (uint32_t)(idx % p.max_parts);
  const uint64_t sq = idx / p.max_parts;
  const uint32_t s = (uint32_t)(sq / p.n_out), slot = p.first_slot + (uint32_t)(sq % p.n_out);
  const uint64_t U = p.total_units, GW = p.total_warps, wt = p.units_per_shard;
  const uint64_t u0 = (uint64_t)s * wt, u1 = u0 + wt;
  uint64_t g = u0 * GW / U;
  while (g + 1 < GW && cpf_run_lo(g + 1, U, GW) <= u0) g++;
  while (g > 0 && cpf_run_lo(g, U, GW) > u0) g--;
  g += j;   // part j of the stripe belongs to run (first owner + j)
  if (g >= GW) return;
  uint64_t a = cpf_run_lo(g, U, GW), b = cpf_run_lo(g + 1, U, GW);
  if (a < u0) a = u0;
  if (b > u1) b = u1;
  if (a >= b) return;   // an empty run, or a run beyond this stripe
  uint32_t r = p.crc_part[((size_t)s * p.n_slots + slot) * p.max_parts + j];
  uint64_t n = u1 - b;   // units behind this part
  for (int i = 0; n; i++, n >>= 1)
    if (n & 1) r = gf32_mul(r, p.x_unit_pow[i], p.poly);
  atomicXor(&p.out[(size_t)s * p.n_slots + slot], r);
}
// crc = ~( R * fix ^ 0xFFFFFFFF * x^(8 len) )
__global__ void crc_parts_finish_kernel(const CrcPartsFinalizeParams p) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.n_stripes * p.n_out) return;
  const uint32_t s = idx / p.n_out, slot = p.first_slot + idx % p.n_out;
  uint32_t* o = p.out + (size_t)s * p.n_slots + slot;
  *o = ~(gf32_mul(*o, p.fix, p.poly) ^ p.init_term);
}
__global__ void crc_parts_zero_kernel(const CrcPartsFinalizeParams p) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.n_stripes * p.n_out) return;
  p.out[(size_t)(idx / p.n_out) * p.n_slots + p.first_slot + idx % p.n_out] = 0;
}

cudaError_t launch_crc_parts_finalize(const CrcPartsFinalizeParams& p, cudaStream_t stream) {
  const int nt = 128;
  const uint32_t n = p.n_stripes * p.n_out;
  const uint64_t total = (uint64_t)n * p.max_parts;
  crc_parts_zero_kernel<<<(n + nt - 1) / nt, nt, 0, stream>>>(p);
  crc_parts_scatter_kernel<<<(unsigned)((total + nt - 1) / nt), nt, 0, stream>>>(p);
  crc_parts_finish_kernel<<<(n + nt - 1) / nt, nt, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Stand-alone CRC over arbitrary byte ranges ("units") of flat buffers: the crc32block
// payload blocks (65532 B, aligned to nothing) and whole buffers.  One CTA per unit at a time.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTabThreads) crc_range_kernel(const CrcRangeParams p) {
  constexpr int NT = kTabThreads, NW = NT / 32;
  __shared__ __align__(16) uint32_t sl[1024];
  __shared__ __align__(16) uint32_t sh[1024];
  __shared__ uint32_t red[NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += NT) {
    sl[i] = (&p.crc->slice[0][0])[i];
    sh[i] = (&p.crc->shift_tile[0][0])[i];
  }
  const uint32_t kt = p.crc->kthread[tid];
  const uint32_t poly = p.crc->poly;
  __syncthreads();
  const uint64_t n_units = (uint64_t)p.n_buffers * p.units_per_buffer;
  for (uint64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const uint32_t b = (uint32_t)(unit / p.units_per_buffer), u = (uint32_t)(unit - (uint64_t)b * p.units_per_buffer);
    const uint8_t* buf = p.base + (size_t)b * p.pitch;
    // unit u of a buffer: block bytes from offset + u * stride (crc32block framing: stride = block + 4, offset = 4)
    const size_t a = (size_t)p.offset + (size_t)u * (p.stride ? p.stride : p.block);
    const size_t e = (a + p.block < p.len) ? a + p.block : p.len;
    // pieces are 16-byte aligned in the address space (base and pitch are 16-aligned)
    const size_t a_al = a & ~(size_t)15;
    const uint32_t T = (uint32_t)((e - a_al + kTabTile - 1) / kTabTile);
    uint32_t st = 0;
    for (uint32_t t = 0; t < T; t++) {
      const size_t col = a_al + (size_t)t * kTabTile + (size_t)tid * kPiece;
      uint4 d = make_uint4(0, 0, 0, 0);
      if (col < e) {
        d = ldg_stream(buf + col);
        if (col < a) d = drop_head(d, (int)(a - col));
        if (col + kPiece > e) d = keep_head(d, (int)(e - col));
      }
      st = crc_constmul(st, sh) ^ crc_piece(d, sl);
    }
    uint32_t v = gf32_mul(st, kt, poly);
    v = __reduce_xor_sync(0xffffffffu, v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      uint32_t r = 0;
      for (int w = 0; w < NW; w++) r ^= red[w];
      // virtual end = a_al + T*TILE; remove the z zero bytes behind the real end e
      const uint64_t z = (uint64_t)a_al + (uint64_t)T * kTabTile - e;
      const uint64_t ord = p.crc->ord;
      const uint64_t neg = (ord - (8 * z) % ord) % ord;
      r = gf32_mul(r, gf32_xpow(neg, poly), poly);
      const uint32_t init_term = gf32_mul(0xFFFFFFFFu, gf32_xpow(8ull * (e - a), poly), poly);
      p.out[unit] = ~(r ^ init_term);
    }
    __syncthreads();
  }
}

cudaError_t launch_crc_ranges(const CrcRangeParams& p, int grid, cudaStream_t stream) {
  crc_range_kernel<<<grid, kTabThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// crc32block framing (blobstore/common/crc32block/block.go:38-49, sized_coder_block.go:43-103, request_body.go:81-130):
// a body of n bytes becomes blocks of block_len bytes, each [crc32(payload) little endian | payload of up to
// block_len - 4 bytes].  Payloads are word multiples (block_len is a multiple of 4096), so the copy between the
// plain and the framed image moves whole 32-bit words; only the final partial word is copied byte by byte.
//   mode 0 (encode): plain -> framed, block checksums taken from `crcs` (crc_range_kernel over the plain image)
//   mode 1 (decode): framed -> plain (the checksums are verified separately by crc32block_check_kernel)
// ------------------------------------------------------------------------------------------
__global__ void crc32block_frame_kernel(const Crc32BlockParams p) {
  const uint64_t words_per_buf = (p.plain_len + 3) / 4;
  const uint64_t total = (uint64_t)p.n_buffers * words_per_buf;
  const uint32_t pw = (p.block_len - 4) / 4;   // payload words per block
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = (uint32_t)(i / words_per_buf);
    const uint64_t w = i - (uint64_t)b * words_per_buf;
    const uint32_t u = (uint32_t)(w / pw);
    const uint64_t plain_off = 4 * w, framed_off = 4 * (w + u + 1);
    const uint8_t* src = (p.mode == 0 ? p.plain : p.framed) + (size_t)b * (p.mode == 0 ? p.plain_pitch : p.framed_pitch) +
                         (p.mode == 0 ? plain_off : framed_off);
    uint8_t* dst = const_cast<uint8_t*>(p.mode == 0 ? p.framed : p.plain) + (size_t)b * (p.mode == 0 ? p.framed_pitch : p.plain_pitch) +
                   (p.mode == 0 ? framed_off : plain_off);
    if (plain_off + 4 <= p.plain_len) {
      *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
    } else {
      for (uint64_t q = plain_off; q < p.plain_len; q++) dst[q - plain_off] = src[q - plain_off];
    }
    if (p.mode == 0 && w == (uint64_t)u * pw) {
      // first payload word of block u: this thread also writes the block's checksum in front of it
      *reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(p.framed) + (size_t)b * p.framed_pitch + (size_t)u * p.block_len) =
          p.crcs[(size_t)b * p.n_blocks + u];
    }
  }
}

// blockUnit.check (block.go:42-49) for every block of every framed buffer: ok[b][u] = (stored == computed);
// first_bad[b] = index of the first failing block, or -1.
__global__ void crc32block_check_kernel(const Crc32BlockParams p) {
  const uint64_t total = (uint64_t)p.n_buffers * p.n_blocks;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = (uint32_t)(i / p.n_blocks), u = (uint32_t)(i - (uint64_t)b * p.n_blocks);
    const uint32_t stored = *reinterpret_cast<const uint32_t*>(p.framed + (size_t)b * p.framed_pitch + (size_t)u * p.block_len);
    const bool ok = stored == p.crcs[i];
    if (p.block_ok) p.block_ok[i] = ok ? 1 : 0;
    if (!ok && p.first_bad) atomicMin(reinterpret_cast<unsigned long long*>(p.first_bad + b), (unsigned long long)u);
  }
}

cudaError_t launch_crc32block_frame(const Crc32BlockParams& p, int grid, cudaStream_t stream) {
  crc32block_frame_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_crc32block_check(const Crc32BlockParams& p, cudaStream_t stream) {
  const uint64_t total = (uint64_t)p.n_buffers * p.n_blocks;
  crc32block_check_kernel<<<(unsigned)std::min<uint64_t>((total + 255) / 256, 65535), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// whole = crc32_combine over the units of a buffer: crc(A||B) = crc(A)*x^(8|B|) ^ crc(B)
__global__ void crc_combine_kernel(const uint32_t* unit_crc, uint32_t n_buffers, uint32_t units, uint32_t len,
                                   uint32_t block, uint32_t poly, uint32_t* whole) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_buffers) return;
  const uint32_t x_block = gf32_xpow(8ull * block, poly);
  const uint32_t last_len = len - (units - 1) * block;
  const uint32_t x_last = gf32_xpow(8ull * last_len, poly);
  uint32_t acc = 0;
  for (uint32_t u = 0; u < units; u++) {
    if (u) acc = gf32_mul(acc, (u == units - 1) ? x_last : x_block, poly);
    acc ^= unit_crc[(size_t)b * units + u];
  }
  whole[b] = acc;
}

cudaError_t launch_crc_combine(const uint32_t* unit_crc, uint32_t n_buffers, uint32_t units, uint32_t len,
                               uint32_t block, uint32_t poly, uint32_t* whole, cudaStream_t stream) {
  const int nt = 64;
  crc_combine_kernel<<<(n_buffers + nt - 1) / nt, nt, 0, stream>>>(unit_crc, n_buffers, units, len, block, poly,
                                                                   whole);
  return cudaGetLastError();
}

__global__ void invert_flags_kernel(int32_t* flags, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = flags[i] ? 0 : 1;
}
cudaError_t launch_invert_flags(int32_t* flags, size_t n, cudaStream_t stream) {
  const int nt = 256;
  invert_flags_kernel<<<(unsigned)((n + nt - 1) / nt), nt, 0, stream>>>(flags, n);
  return cudaGetLastError();
}

}  // namespace cbe
