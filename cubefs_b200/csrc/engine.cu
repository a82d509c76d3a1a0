// engine.cu -- host side of libcubeec behind the C-ABI in include/cubeec.h.
//
// Mirrors the call contract of klauspost/reedsolomon's Encoder as CubeFS uses it
// (blobstore/common/ec/encoder.go:118,122,136,143,150) and adds batched / device-resident
// entry points.  No CPU compute path exists here: every coding or checksum byte is produced by
// the CUDA kernels in kernels.cu / bitslice.cu; without a device the calls fail.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cubeec.h"
#include "gfmath.h"
#include "kernels.cuh"

using namespace cbe;

// ------------------------------------------------------------------------------------------
// errors / counters
// ------------------------------------------------------------------------------------------
static thread_local std::string t_last_error;
static thread_local const char* t_last_kernel = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_force_kernel{0};   // cubeec_debug_force_kernel: A/B measurement aid

static int cuda_fail(cudaError_t e, const char* what) {
  t_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CUBEEC_ERR_NO_DEVICE : CUBEEC_ERR_CUDA;
}
#define CU(call)                                        \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

extern "C" const char* cubeec_strerror(int code) {
  switch (code) {
    case CUBEEC_OK: return "ok";
    case CUBEEC_ERR_INV_SHARD_NUM: return "cannot create Encoder with less than one data shard or less than zero parity shards";
    case CUBEEC_ERR_MAX_SHARD_NUM: return "cannot create Encoder with more than 256 data+parity shards";
    case CUBEEC_ERR_TOO_FEW_SHARDS: return "too few shards given";
    case CUBEEC_ERR_SHARD_NO_DATA: return "no shard data";
    case CUBEEC_ERR_SHARD_SIZE: return "shard sizes do not match";
    case CUBEEC_ERR_SHORT_DATA: return "not enough data to fill the number of requested shards";
    case CUBEEC_ERR_RECONSTRUCT_REQUIRED: return "reconstruction required as one or more required data shards are nil";
    case CUBEEC_ERR_SINGULAR: return "matrix is singular";
    case CUBEEC_ERR_INVALID_ARG: return "invalid argument";
    case CUBEEC_ERR_NO_DEVICE: return "no usable CUDA device (libcubeec has no CPU fallback)";
    case CUBEEC_ERR_CUDA: return "CUDA error";
    case CUBEEC_ERR_UNSUPPORTED: return "unsupported geometry";
  }
  return "unknown error";
}
extern "C" const char* cubeec_last_error(void) { return t_last_error.c_str(); }
extern "C" uint64_t cubeec_kernel_launches(void) { return g_launches.load(); }
extern "C" const char* cubeec_last_kernel(void) { return t_last_kernel; }
extern "C" void cubeec_debug_force_kernel(int which);

// ------------------------------------------------------------------------------------------
// per-device context
// ------------------------------------------------------------------------------------------
namespace {

constexpr size_t kAlign = 256;
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Lane {
  cudaStream_t stream = nullptr;
  uint8_t* d_buf = nullptr;
  size_t d_cap = 0;
  uint8_t* d_aux = nullptr;   // patterns, crc parts, flags
  size_t aux_cap = 0;
  uint8_t* h_aux = nullptr;   // pinned mirror of small results
  size_t h_aux_cap = 0;
};

// Stream-ordered scratch that is released on EVERY exit path of the function that allocated it (the
// free is enqueued behind the kernels already launched on the stream).
struct AsyncScratch {
  cudaStream_t stream;
  void* ptr = nullptr;
  explicit AsyncScratch(cudaStream_t s) : stream(s) {}
  AsyncScratch(const AsyncScratch&) = delete;
  AsyncScratch& operator=(const AsyncScratch&) = delete;
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&ptr, bytes, stream); }
  ~AsyncScratch() {
    if (ptr) cudaFreeAsync(ptr, stream);
  }
};

struct DevCtx {
  int device = 0;
  int sm_count = 0;
  size_t smem_limit = 0;
  GfDeviceTables* d_gf = nullptr;
  CrcDeviceTables* d_crc[2] = {nullptr, nullptr};
  // bit-sliced kernel: lane-private slicing-table image, Horner (fold) tables, per-thread alignment constants
  uint32_t* d_bs_slice[2] = {nullptr, nullptr};
  uint32_t* d_bs_fold[2] = {nullptr, nullptr};
  uint32_t* d_bs_kthread[2] = {nullptr, nullptr};
  // the same two for the tile of the warp-specialised kernel (bitslice_ws.cu)
  uint32_t* d_bsw_fold[2] = {nullptr, nullptr};
  uint32_t* d_bsw_kthread[2] = {nullptr, nullptr};
  // flat-split fused kernel (bs_flat.cuh): Horner step between a lane's pieces of consecutive units, lane alignment
  uint32_t* d_bsf_fold[2] = {nullptr, nullptr};
  uint32_t* d_bsf_klane[2] = {nullptr, nullptr};
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Lane*> free_lanes;
  int lanes_created = 0;
  // Multi-lane (batched) operations take one of kBulkSlots slots before their lanes: bounded staging memory, and
  // no caller ever waits for a lane while holding some (kBulkSlots * 3 + coalescing workers <= kLanesPerDevice).
  int bulk_free = 4;
};

constexpr int kLanesPerDevice = 16;

struct BulkSlot {
  DevCtx* c;
  explicit BulkSlot(DevCtx* ctx) : c(ctx) {
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv.wait(lk, [&] { return c->bulk_free > 0; });
    c->bulk_free--;
  }
  BulkSlot(const BulkSlot&) = delete;
  BulkSlot& operator=(const BulkSlot&) = delete;
  ~BulkSlot() {
    std::lock_guard<std::mutex> lk(c->mu);
    c->bulk_free++;
    c->cv.notify_all();
  }
};

struct Global {
  std::mutex mu;
  bool ready = false;
  int init_rc = CUBEEC_OK;
  std::vector<int> devices;
  std::vector<std::unique_ptr<DevCtx>> ctx;
  CrcPoly poly[2] = {{0xEDB88320u, 0xFFFFFFFFLL}, {0x82F63B78u, 0x7FFFFFFFLL}};
};
Global g;

int setup_device(DevCtx& c) {
  CU(cudaSetDevice(c.device));
  CU(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, c.device));
  CU(tab_configure(&c.smem_limit));
  GfDeviceTables gt;
  std::memcpy(gt.log, gf().log, 256);
  std::memcpy(gt.exp, gf().exp, 512);
  CU(cudaMalloc(&c.d_gf, sizeof(gt)));
  CU(cudaMemcpy(c.d_gf, &gt, sizeof(gt), cudaMemcpyHostToDevice));
  for (int pi = 0; pi < 2; pi++) {
    auto ct = std::make_unique<CrcDeviceTables>();
    const CrcPoly& P = g.poly[pi];
    crc_slice_tables(P.poly, ct->slice);
    crc_const_mul_tables(P, P.shift_bytes_const(kTabTile), ct->shift_tile);
    for (int t = 0; t < 1024; t++)
      ct->kthread[t] = t < kTabThreads ? P.shift_bytes_const((int64_t)kTabTile - 16 * (t + 1)) : 0;
    ct->poly = P.poly;
    ct->ord = (uint32_t)P.ord;
    if (P.xpow_raw((uint64_t)P.ord) != 0x80000000u || P.mul(P.xpow(-12345), P.xpow(12345)) != 0x80000000u) {
      t_last_error = "CRC polynomial order self-check failed";
      return CUBEEC_ERR_CUDA;
    }
    CU(cudaMalloc(&c.d_crc[pi], sizeof(CrcDeviceTables)));
    CU(cudaMemcpy(c.d_crc[pi], ct.get(), sizeof(CrcDeviceTables), cudaMemcpyHostToDevice));
    // bitslice.cu tables: image offset (j>>1)*65536 + v*256 + (j&1)*128 + lane*4 = slice[j][v]
    std::vector<uint32_t> img(kBsSliceImageBytes / 4);
    for (int j = 0; j < 4; j++)
      for (int v = 0; v < 256; v++)
        for (int l = 0; l < 32; l++) img[((size_t)(j >> 1) * 65536 + (size_t)v * 256 + (size_t)(j & 1) * 128) / 4 + l] = ct->slice[j][v];
    uint32_t fold[4][256];
    crc_const_mul_tables(P, P.shift_bytes_const((int64_t)kBsTile - kBsPiece), fold);
    std::vector<uint32_t> kth(2 * kBsThreads);   // second half: the thread sat out the shard's final tile
    for (int t = 0; t < kBsThreads; t++) {
      kth[t] = P.shift_bytes_const((int64_t)kBsTile - (int64_t)kBsPiece * (t + 1));
      kth[kBsThreads + t] = P.shift_bytes_const(2 * (int64_t)kBsTile - (int64_t)kBsPiece * (t + 1));
    }
    CU(cudaMalloc(&c.d_bs_slice[pi], kBsSliceImageBytes));
    CU(cudaMemcpy(c.d_bs_slice[pi], img.data(), kBsSliceImageBytes, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.d_bs_fold[pi], sizeof(fold)));
    CU(cudaMemcpy(c.d_bs_fold[pi], fold, sizeof(fold), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.d_bs_kthread[pi], kth.size() * 4));
    CU(cudaMemcpy(c.d_bs_kthread[pi], kth.data(), kth.size() * 4, cudaMemcpyHostToDevice));
    crc_const_mul_tables(P, P.shift_bytes_const((int64_t)kBswTile - kBsPiece), fold);
    kth.assign(kBswE, 0);
    for (int t = 0; t < kBswE; t++) kth[t] = P.shift_bytes_const((int64_t)kBswTile - (int64_t)kBsPiece * (t + 1));
    CU(cudaMalloc(&c.d_bsw_fold[pi], sizeof(fold)));
    CU(cudaMemcpy(c.d_bsw_fold[pi], fold, sizeof(fold), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.d_bsw_kthread[pi], kth.size() * 4));
    CU(cudaMemcpy(c.d_bsw_kthread[pi], kth.data(), kth.size() * 4, cudaMemcpyHostToDevice));
    crc_const_mul_tables(P, P.shift_bytes_const((int64_t)kBsfUnitBytes - kBsPiece), fold);
    uint32_t klane[32];
    for (int l = 0; l < 32; l++) klane[l] = P.shift_bytes_const((int64_t)kBsPiece * (31 - l));
    CU(cudaMalloc(&c.d_bsf_fold[pi], sizeof(fold)));
    CU(cudaMemcpy(c.d_bsf_fold[pi], fold, sizeof(fold), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.d_bsf_klane[pi], sizeof(klane)));
    CU(cudaMemcpy(c.d_bsf_klane[pi], klane, sizeof(klane), cudaMemcpyHostToDevice));
  }
  return CUBEEC_OK;
}

int ensure_init() {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.ready) return g.init_rc;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    t_last_error = e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA devices";
    return CUBEEC_ERR_NO_DEVICE;   // not cached: a later call may find a device
  }
  if (g.devices.empty()) g.devices.push_back(0);
  // contexts are built aside and committed only when every device came up (a failed attempt leaves no
  // half-initialised state behind; its device tables are released with the process)
  std::vector<std::unique_ptr<DevCtx>> fresh;
  for (int d : g.devices) {
    if (d < 0 || d >= n) return CUBEEC_ERR_INVALID_ARG;
    auto c = std::make_unique<DevCtx>();
    c->device = d;
    int rc = setup_device(*c);
    if (rc) return rc;
    fresh.push_back(std::move(c));
  }
  g.ctx = std::move(fresh);
  g.ready = true;
  g.init_rc = CUBEEC_OK;
  return CUBEEC_OK;
}

DevCtx* ctx_for_device(int device) {
  for (auto& c : g.ctx)
    if (c->device == device) return c.get();
  return nullptr;
}

// Borrow a lane (stream + scratch).  Blocks when all lanes of the device are in use.
struct LaneLease {
  DevCtx* c = nullptr;
  Lane* lane = nullptr;
  int acquire(DevCtx* ctx) {
    c = ctx;
    std::unique_lock<std::mutex> lk(c->mu);
    for (;;) {
      if (!c->free_lanes.empty()) {
        lane = c->free_lanes.back();
        c->free_lanes.pop_back();
        break;
      }
      if (c->lanes_created < kLanesPerDevice) {
        c->lanes_created++;
        lk.unlock();
        lane = new Lane();
        cudaSetDevice(c->device);
        cudaError_t e = cudaStreamCreateWithFlags(&lane->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
          delete lane;
          lane = nullptr;   // the destructor must not pool a lane without a stream
          lk.lock();
          c->lanes_created--;
          c->cv.notify_one();
          return cuda_fail(e, "cudaStreamCreate");
        }
        break;
      }
      c->cv.wait(lk);
    }
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return CUBEEC_OK;
  }
  ~LaneLease() {
    if (c && lane) {
      std::lock_guard<std::mutex> lk(c->mu);
      c->free_lanes.push_back(lane);
      c->cv.notify_all();
    }
  }
};

int lane_reserve(Lane& l, size_t d_bytes, size_t aux_bytes) {
  if (d_bytes > l.d_cap) {
    CU(cudaStreamSynchronize(l.stream));
    if (l.d_buf) CU(cudaFree(l.d_buf));
    l.d_buf = nullptr;
    l.d_cap = 0;
    size_t cap = round_up(d_bytes + d_bytes / 4, 1 << 20);
    CU(cudaMalloc(&l.d_buf, cap));
    l.d_cap = cap;
  }
  if (aux_bytes > l.aux_cap) {
    CU(cudaStreamSynchronize(l.stream));
    if (l.d_aux) CU(cudaFree(l.d_aux));
    if (l.h_aux) CU(cudaFreeHost(l.h_aux));
    l.d_aux = nullptr;
    l.h_aux = nullptr;
    size_t cap = round_up(aux_bytes * 2, 1 << 16);
    CU(cudaMalloc(&l.d_aux, cap));
    CU(cudaMallocHost(&l.h_aux, cap));
    l.aux_cap = l.h_aux_cap = cap;
  }
  return CUBEEC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct cubeec {
  int k = 0, m = 0;
  std::vector<uint8_t> gen;   // (k+m) x k
  // encode passes: each <= kMaxOut outputs
  std::vector<Pattern> enc_passes;
  std::vector<Pattern> verify_passes;
  std::vector<Pattern*> d_enc;      // per ctx: device copy [n_passes]
  std::vector<Pattern*> d_verify;   // same, crc_in = 0
  int bs_passes = 0;                // bit-sliced passes with fused CRC (bs_passes() plan 0); 0 = table kernels only
  int bs_passes_plain = 0;          // ... of a plain encode / verify (plan 1: wider passes)
  // Batched-reconstruct plans: the pattern tables of a whole presence array, resident on a device,
  // so that repeating a repair batch costs no host-side matrix work or uploads.
  struct Plan {
    int device = 0;
    bool data_only = false;
    std::vector<uint8_t> present;
    Pattern* d_pat = nullptr;
    RecPattern* d_rec = nullptr;      // bit-sliced syndrome reconstruct patterns (instead of d_pat)
    uint32_t* d_pos = nullptr;
    size_t n_pass = 0, n_pat = 0;
    std::vector<int> nin;
    std::vector<uint8_t> slot_map;    // empty = identity; else shard i of the code is slot_map[i] of the stripe (LRC local stripes)
    std::vector<Pattern> h_pat;       // host copy of d_pat ([n_pass][n_pat]): single-pattern plans feed the JIT (jit.cu)
    std::vector<const void*> jit;     // per pass: run-time compiled kernel of the (only) pattern (nullptr: pass without outputs)
    int jit_state = 0;                // 0 not tried, 1 ready, 2 unavailable (no NVRTC / compile error): table kernels
  };
  std::vector<std::unique_ptr<Plan>> plans;
  std::mutex mu;
  // decode pattern cache: presence string (+data_only) -> passes
  std::map<std::string, std::vector<Pattern>> dec_cache;
};

namespace {

// Pattern passes for an arbitrary (inputs, outputs, rows) job.
void make_passes(const std::vector<int>& in_slots, const std::vector<int>& out_slots,
                 const std::vector<uint8_t>& rows /* n_out x n_in */, bool crc_in_first, std::vector<Pattern>& out) {
  const int n_in = (int)in_slots.size(), n_out = (int)out_slots.size();
  out.clear();
  for (int o0 = 0; o0 < n_out || (o0 == 0 && crc_in_first); o0 += kMaxOut) {
    Pattern p;
    std::memset(&p, 0, sizeof(p));
    p.n_in = (uint8_t)n_in;
    p.n_out = (uint8_t)std::min(kMaxOut, n_out - o0);
    p.crc_in = (o0 == 0 && crc_in_first) ? 1 : 0;
    for (int c = 0; c < n_in; c++) p.in_slot[c] = (uint8_t)in_slots[c];
    for (int r = 0; r < p.n_out; r++) {
      p.out_slot[r] = (uint8_t)out_slots[o0 + r];
      for (int c = 0; c < n_in; c++) p.coef[r][c] = rows[(size_t)(o0 + r) * n_in + c];
    }
    out.push_back(p);
    if (n_out == 0) break;
  }
}

// The rows that regenerate the missing shards of a presence pattern from the first k present shards
// (RS/reedsolomon.go:1453-1524): inverse rows for data, parity_row * inverse for parity.
int decode_rows(int k, int m, const std::vector<uint8_t>& gen, const uint8_t* present, bool data_only, std::vector<int>& valid,
                std::vector<int>& outs, std::vector<uint8_t>& rows) {
  const int n = k + m;
  for (int i = 0; i < n && (int)valid.size() < k; i++)
    if (present[i]) valid.push_back(i);
  if ((int)valid.size() < k) return CUBEEC_ERR_TOO_FEW_SHARDS;
  std::vector<uint8_t> sub((size_t)k * k), dec((size_t)k * k);
  for (int r = 0; r < k; r++) std::memcpy(&sub[(size_t)r * k], &gen[(size_t)valid[r] * k], (size_t)k);
  if (!gf_invert(sub.data(), k, dec.data())) return CUBEEC_ERR_SINGULAR;
  const Gf256& G = gf();
  for (int i = 0; i < n; i++) {
    if (present[i]) continue;
    if (i >= k && data_only) continue;
    outs.push_back(i);
    if (i < k) {
      rows.insert(rows.end(), dec.begin() + (size_t)i * k, dec.begin() + (size_t)(i + 1) * k);
    } else {
      for (int c = 0; c < k; c++) {
        uint8_t v = 0;
        for (int j = 0; j < k; j++) v ^= G.mul(gen[(size_t)i * k + j], dec[(size_t)j * k + c]);
        rows.push_back(v);
      }
    }
  }
  return CUBEEC_OK;
}

// Presence pattern -> fused decode passes: every missing shard expressed directly over the
// first k present shards (data rows from the inverse, RS/reedsolomon.go:1469-1524; parity rows
// pre-multiplied: parity_row * decode, same values as the reference's second pass :1531-1550).
int decode_passes(cubeec* h, const uint8_t* present, bool data_only, std::vector<Pattern>& passes) {
  const int k = h->k, n = h->k + h->m;
  std::string key((const char*)present, (size_t)n);
  for (auto& ch : key) ch = ch ? 1 : 0;
  key.push_back(data_only ? 1 : 0);
  {
    std::lock_guard<std::mutex> lk(h->mu);
    auto it = h->dec_cache.find(key);
    if (it != h->dec_cache.end()) { passes = it->second; return CUBEEC_OK; }
  }
  std::vector<int> valid, outs;
  std::vector<uint8_t> rows;
  int rc = decode_rows(k, h->m, h->gen, present, data_only, valid, outs, rows);
  if (rc) return rc;
  make_passes(valid, outs, rows, false, passes);
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->dec_cache.size() > 8192) h->dec_cache.clear();
  h->dec_cache[key] = passes;
  return CUBEEC_OK;
}

// Presence pattern -> syndrome-decode pattern for rs_bsrec_kernel (see RecPattern in kernels.cuh).
int rec_pattern(cubeec* h, const uint8_t* present, bool data_only, RecPattern& rp) {
  const int k = h->k, m = h->m;
  std::memset(&rp, 0, sizeof(rp));
  std::memset(rp.out_prow, 0xff, sizeof(rp.out_prow));
  std::vector<int> Ed, R, Ep;
  int n_present = 0;
  for (int i = 0; i < k + m; i++) n_present += present[i] ? 1 : 0;
  for (int c = 0; c < k; c++) {
    if (present[c]) rp.data_mask |= 1u << c;
    else Ed.push_back(c);
  }
  if (n_present == k + m || (data_only && Ed.empty())) return CUBEEC_OK;   // nothing to do (n_out = 0)
  if (n_present < k) return CUBEEC_ERR_TOO_FEW_SHARDS;
  for (int r = 0; r < m; r++) {
    if (present[k + r]) { if (R.size() < Ed.size()) R.push_back(r); }
    else if (!data_only) Ep.push_back(r);
  }
  if (R.size() < Ed.size()) return CUBEEC_ERR_TOO_FEW_SHARDS;
  const int ed = (int)Ed.size();
  if (ed + (int)Ep.size() > 4 || ed > 4) return CUBEEC_ERR_UNSUPPORTED;
  const Gf256& G = gf();
  std::vector<uint8_t> A((size_t)ed * ed), Ainv((size_t)ed * ed);
  for (int i = 0; i < ed; i++)
    for (int j = 0; j < ed; j++) A[(size_t)i * ed + j] = h->gen[(size_t)(k + R[i]) * k + Ed[j]];
  if (ed && !gf_invert(A.data(), ed, Ainv.data())) return CUBEEC_ERR_SINGULAR;
  rp.n_syn = (uint8_t)ed;
  for (int r : R) rp.syn_mask |= (uint8_t)(1u << r);
  int j = 0;
  for (int a = 0; a < ed; a++, j++) {   // missing data shard Ed[a] = sum_i Ainv[a][i] * S_{R[i]}
    rp.out_slot[j] = (uint8_t)Ed[a];
    for (int i = 0; i < ed; i++) rp.coef[j][i] = Ainv[(size_t)a * ed + i];
  }
  for (int pr : Ep) {                   // missing parity row pr = T_pr + sum_a M[pr][Ed[a]] * D_{Ed[a]}
    rp.out_slot[j] = (uint8_t)(k + pr);
    rp.out_prow[j] = (uint8_t)pr;
    rp.t_mask |= (uint8_t)(1u << pr);
    for (int i = 0; i < ed; i++) {
      uint8_t v = 0;
      for (int a = 0; a < ed; a++) v ^= G.mul(h->gen[(size_t)(k + pr) * k + Ed[a]], Ainv[(size_t)a * ed + i]);
      rp.coef[j][i] = v;
    }
    j++;
  }
  rp.n_out = (uint8_t)j;
  return CUBEEC_OK;
}

struct Geometry {
  uint32_t n_seg, tiles_per_seg, tiles_last;
  int grid;
  uint32_t tile;   // bytes of a shard per tile
  uint32_t packed_pps = 0;   // bit-sliced kernel packed mode: 64-byte pieces per shard (see BsParams)
};

Geometry pick_geometry(const DevCtx& c, size_t shard_len, size_t n_stripes, bool per_stripe_patterns,
                       uint32_t tile = kTabTile) {
  Geometry gm;
  gm.tile = tile;
  const uint64_t tiles_total = (shard_len + tile - 1) / tile;
  // enough work items to balance the persistent CTAs; whole stripes once the batch is large
  // (per-item costs: table rebuild for per-stripe patterns, CRC alignment multiply)
  uint64_t want_items = (uint64_t)c.sm_count * (per_stripe_patterns ? 2 : 8);
  if (n_stripes >= (uint64_t)c.sm_count * 4) want_items = n_stripes;
  uint64_t n_seg = (want_items + n_stripes - 1) / std::max<size_t>(n_stripes, 1);
  const uint64_t max_seg = std::max<uint64_t>(1, tiles_total / (tile >= 16384 ? 1 : 4));
  n_seg = std::max<uint64_t>(1, std::min(n_seg, max_seg));
  uint64_t tps = (tiles_total + n_seg - 1) / n_seg;
  n_seg = (tiles_total + tps - 1) / tps;
  gm.n_seg = (uint32_t)n_seg;
  gm.tiles_per_seg = (uint32_t)tps;
  gm.tiles_last = (uint32_t)(tiles_total - (n_seg - 1) * tps);
  uint64_t items = n_stripes * n_seg;
  gm.grid = (int)std::min<uint64_t>(items, (uint64_t)c.sm_count);
  if (gm.grid < 1) gm.grid = 1;
  return gm;
}

// Run `n_pass` pattern passes over a device-resident batch.  d_pass_patterns[j] points at the
// pattern array of pass j (indexed by pattern_of_stripe, or a single pattern when that is null).
int run_passes(DevCtx& c, cudaStream_t stream, uint8_t* d_base, size_t shard_len, size_t shard_pitch,
               size_t stripe_pitch, size_t n_stripes, int n_slots, const std::vector<const Pattern*>& d_pass_patterns,
               const std::vector<int>& pass_n_in, const std::vector<int>& pass_crc_slots,
               const std::vector<char>& pass_n_in_exact /* every pattern of the pass has exactly pass_n_in inputs */,
               const uint32_t* d_pattern_of_stripe, int mode, int32_t* d_mismatch, uint32_t* d_crc_part,
               int crc_poly, const Geometry& gm) {
  for (size_t j = 0; j < d_pass_patterns.size(); j++) {
    TabParams p;
    std::memset(&p, 0, sizeof(p));
    p.base = d_base;
    p.stripe_pitch = stripe_pitch;
    p.shard_pitch = shard_pitch;
    p.shard_len = (uint32_t)shard_len;
    p.n_stripes = (uint32_t)n_stripes;
    p.n_seg = gm.n_seg;
    p.tiles_per_seg = gm.tiles_per_seg;
    p.tiles_last = gm.tiles_last;
    p.n_slots = (uint32_t)n_slots;
    p.patterns = d_pass_patterns[j];
    p.pattern_of_stripe = d_pattern_of_stripe;
    p.mode = mode;
    p.mismatch = d_mismatch;
    p.crc_part = d_crc_part;
    p.gf = c.d_gf;
    p.crc = c.d_crc[crc_poly ? 1 : 0];
    const bool with_crc = d_crc_part != nullptr;
    if (!with_crc && pass_n_in_exact[j] && tabk_supported(pass_n_in[j]) && g_force_kernel.load() != 3) {
      CU(launch_tabk(p, pass_n_in[j], gm.grid, stream));
      g_launches++;
      t_last_kernel = "rs_tabk_kernel";
      continue;
    }
    size_t smem = 0;
    int R = tab_pick_replication(pass_n_in[j], with_crc, pass_crc_slots[j], c.smem_limit, &smem);
    if (R == 0) return CUBEEC_ERR_UNSUPPORTED;
    CU(launch_tab(p, R, with_crc, pass_crc_slots[j], smem, gm.grid, stream));
    g_launches++;
    t_last_kernel = with_crc ? "rs_tab_kernel<crc>" : "rs_tab_kernel";
  }
  return CUBEEC_OK;
}

int finalize_crc(DevCtx& c, cudaStream_t stream, const uint32_t* d_crc_part, size_t n_stripes, int n_slots,
                 size_t shard_len, const Geometry& gm, int crc_poly, const uint8_t* d_slot_enable, uint32_t* d_out) {
  (void)c;
  const CrcPoly& P = g.poly[crc_poly ? 1 : 0];
  CrcFinalizeParams f;
  std::memset(&f, 0, sizeof(f));
  f.crc_part = d_crc_part;
  f.n_units = (uint32_t)(n_stripes * n_slots);
  f.n_slots = (uint32_t)n_slots;
  f.n_seg = gm.n_seg;
  const int64_t seg_bytes = (int64_t)gm.tiles_per_seg * gm.tile;
  const int64_t last_bytes = (int64_t)gm.tiles_last * gm.tile;
  int64_t virt = (int64_t)(gm.n_seg - 1) * seg_bytes + last_bytes;
  f.x_full = P.shift_bytes_const(seg_bytes);
  f.x_last = P.shift_bytes_const(last_bytes);
  if (gm.packed_pps) {
    // packed mode: the (at most two) partial remainders of a shard are already aligned to the end of
    // the shard's last 64-byte piece, so they are simply XORed (multiplier x^0)
    f.x_full = f.x_last = 0x80000000u;
    virt = (int64_t)gm.packed_pps * kBsPiece;
  }
  f.fix = P.shift_bytes_const(-(virt - (int64_t)shard_len));
  f.init_term = P.mul(0xFFFFFFFFu, P.shift_bytes_const((int64_t)shard_len));
  f.poly = P.poly;
  f.slot_enable = d_slot_enable;
  f.out = d_out;
  CU(launch_crc_finalize(f, stream));
  g_launches++;
  return CUBEEC_OK;
}

// checkShards (RS/reedsolomon.go:1314-1327)
int check_shards(const size_t* lens, int n, bool nilok, size_t* size_out) {
  size_t size = 0;
  for (int i = 0; i < n; i++)
    if (lens[i]) { size = lens[i]; break; }
  if (size == 0) return CUBEEC_ERR_SHARD_NO_DATA;
  for (int i = 0; i < n; i++)
    if (lens[i] != size && (lens[i] != 0 || !nilok)) return CUBEEC_ERR_SHARD_SIZE;
  *size_out = size;
  return CUBEEC_OK;
}

int upload_handle_patterns(cubeec* h) {
  h->d_enc.assign(g.ctx.size(), nullptr);
  h->d_verify.assign(g.ctx.size(), nullptr);
  for (size_t ci = 0; ci < g.ctx.size(); ci++) {
    CU(cudaSetDevice(g.ctx[ci]->device));
    const size_t bytes = sizeof(Pattern) * h->enc_passes.size();
    if (!bytes) continue;
    CU(cudaMalloc(&h->d_enc[ci], bytes));
    CU(cudaMemcpy(h->d_enc[ci], h->enc_passes.data(), bytes, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&h->d_verify[ci], bytes));
    CU(cudaMemcpy(h->d_verify[ci], h->verify_passes.data(), bytes, cudaMemcpyHostToDevice));
  }
  return CUBEEC_OK;
}

size_t ctx_index(const DevCtx* c) {
  for (size_t i = 0; i < g.ctx.size(); i++)
    if (g.ctx[i].get() == c) return i;
  return 0;
}

// The bit-sliced kernel needs 32-byte columns: base and pitches 32-aligned, room for whole groups.
bool bs_layout_ok(const uint8_t* d_base, size_t shard_len, size_t shard_pitch, size_t stripe_pitch) {
  return (((uintptr_t)d_base | shard_pitch | stripe_pitch) & 31) == 0 && shard_pitch >= round_up(shard_len, 32);
}



bool bs_is_packed(size_t shard_len) {
  const uint32_t pps = (uint32_t)((shard_len + kBsPiece - 1) / kBsPiece);
  return pps < (uint32_t)kBsThreads && pps >= 8;
}

// Geometry of the bit-sliced kernels; shards shorter than a tile use packed mode (pieces of many
// stripes share a tile).
// balanced = a kernel without per-item cost (plain encode, verify: no CRC alignment per segment): the number of
// segments per shard is chosen to minimise ceil(items / SMs) * tiles per item, so that e.g. 383 stripes of 11 tiles
// do not run 11 rounds of 3-tile items (33 tile steps, 86 % of the 28.5 the work needs) but 29 rounds of 1-tile items.
Geometry bs_geometry(const DevCtx& c, size_t shard_len, size_t n_stripes, bool balanced = false) {
  Geometry gm = pick_geometry(c, shard_len, n_stripes, false, kBsTile);
  if (balanced && !bs_is_packed(shard_len) && n_stripes > 0) {
    const uint64_t tiles_total = (shard_len + kBsTile - 1) / kBsTile, sms = (uint64_t)c.sm_count;
    uint64_t best_seg = gm.n_seg, best_cost = ~0ull;
    for (uint64_t ns = 1; ns <= tiles_total; ns++) {
      const uint64_t tps = (tiles_total + ns - 1) / ns, nseg = (tiles_total + tps - 1) / tps;
      const uint64_t items = n_stripes * nseg;
      const uint64_t cost = ((items + sms - 1) / sms) * tps * 64 + (items + sms - 1) / sms;   // tile steps (x64) + a little per item
      if (cost < best_cost) {
        best_cost = cost;
        best_seg = nseg;
      }
    }
    const uint64_t tps = (tiles_total + best_seg - 1) / best_seg;
    gm.n_seg = (uint32_t)((tiles_total + tps - 1) / tps);
    gm.tiles_per_seg = (uint32_t)tps;
    gm.tiles_last = (uint32_t)(tiles_total - (uint64_t)(gm.n_seg - 1) * tps);
    gm.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_stripes * gm.n_seg, sms));
  }
  if (bs_is_packed(shard_len)) {
    const uint32_t pps = (uint32_t)((shard_len + kBsPiece - 1) / kBsPiece);
    const uint64_t tiles = ((uint64_t)n_stripes * pps + kBsThreads - 1) / kBsThreads;
    gm.n_seg = 2;               // a shard's pieces fall into at most two tiles
    gm.packed_pps = pps;
    gm.grid = (int)std::min<uint64_t>(tiles, (uint64_t)c.sm_count);
  }
  return gm;
}

// All bit-sliced passes of handle h over a device-resident batch whose stripes have n_slots shards.
// in_slots (h->k entries, nullptr = 0..k-1) says which shard of the stripe each input is, outputs go to
// out_first, out_first+1, ...  crc: 0 none, 1 every shard of this code, 2 only its outputs (the caller
// finalises the partial remainders in d_part once all codes of the stripe have run).
int bs_run(cubeec* h, DevCtx& c, cudaStream_t stream, uint8_t* d_base, size_t shard_len, size_t shard_pitch,
           size_t stripe_pitch, size_t n_stripes, const Geometry& gm, int n_slots, const uint8_t* in_slots, int out_first,
           int crc, uint32_t* d_part, int crc_poly, bool verify, int32_t* d_mismatch, bool ws = false, bool rolled = false) {
  BsParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.base = d_base;
  bp.stripe_pitch = stripe_pitch;
  bp.shard_pitch = shard_pitch;
  bp.shard_len = (uint32_t)shard_len;
  bp.n_stripes = (uint32_t)n_stripes;
  bp.n_seg = gm.n_seg;
  bp.tiles_per_seg = gm.tiles_per_seg;
  bp.tiles_last = gm.tiles_last;
  bp.n_slots = (uint32_t)n_slots;
  bp.crc_part = crc ? d_part : nullptr;
  bp.mismatch = verify ? d_mismatch : nullptr;
  const int pi = crc_poly ? 1 : 0;
  bp.slice_image = c.d_bs_slice[pi];
  bp.fold_tables = ws ? c.d_bsw_fold[pi] : c.d_bs_fold[pi];
  bp.kthread = ws ? c.d_bsw_kthread[pi] : c.d_bs_kthread[pi];
  bp.poly = g.poly[pi].poly;
  bp.packed_pps = gm.packed_pps;
  if (h->k > (int)sizeof(bp.in_slot)) return CUBEEC_ERR_UNSUPPORTED;
  if (crc == 1 && in_slots) return CUBEEC_ERR_INVALID_ARG;   // data CRCs only with the identity input map (bs_slot)
  for (int i = 0; i < h->k; i++) bp.in_slot[i] = in_slots ? in_slots[i] : (uint8_t)i;
  if (ws) {
    CU(launch_bsw(h->k, h->m, bp, gm.grid, stream));
    g_launches++;
    t_last_kernel = "rs_bsw_kernel";
    return CUBEEC_OK;
  }
  if (rolled) {
    for (int r = 0; r < h->m; r++) bp.out_slot[r] = (uint8_t)(out_first + r);
    CU(launch_bs_rolled(h->k, h->m, bp, gm.grid, stream));
    g_launches++;
    t_last_kernel = "rs_bs_kernel<crc,rolled>";
    return CUBEEC_OK;
  }
  // m > 4: one pass per group of parity rows (4 per pass with fused CRC, <= 6 without)
  const int n_pass = crc ? h->bs_passes : h->bs_passes_plain;
  for (int pass = 0; pass < n_pass; pass++) {
    int r0 = 0, rows = h->m;
    if (n_pass > 1 || h->m > 4) {
      if (!bs_mp_pass_rows(h->k, h->m, crc ? 0 : 1, pass, &r0, &rows)) return CUBEEC_ERR_UNSUPPORTED;
    }
    for (int r = 0; r < rows && r < (int)sizeof(bp.out_slot); r++) bp.out_slot[r] = (uint8_t)(out_first + r0 + r);
    const int mode = crc == 0 ? 0 : (crc == 1 && pass == 0 ? 1 : 2);
    CU(launch_bs(h->k, h->m, pass, bp, mode, verify, gm.grid, stream));
    g_launches++;
  }
  t_last_kernel = crc ? "rs_bs_kernel<crc>" : verify ? "rs_bs_kernel<verify>" : "rs_bs_kernel";
  return CUBEEC_OK;
}

int dev_crc32_impl(DevCtx& c, cudaStream_t st, const uint8_t* d_base, size_t len, size_t pitch, size_t n_buffers,
                   size_t block, int crc_poly, uint32_t* d_whole, uint32_t* d_blocks);

// ---- flat work split of the fused encode + CRC kernel (bs_flat.cuh) ----------------------------------
struct FlatGeometry {
  uint32_t units_per_shard = 0;
  uint64_t total_units = 0, total_warps = 0;
  uint32_t max_parts = 0;
  int grid = 0, threads = kBsfThreads, variant = kBsfThreads;
};

int bsf_threads() {
  const int f = g_force_kernel.load();
  return (f >= 1000 && f < 2000) ? f - 1000 : kBsfThreads;   // cubeec_debug_force_kernel(1000 + variant): A/B aid
}

// cubeec_debug_force_kernel(3000 + bits): bit 0 flips the parameter-passing entry, bit 1 the per-unit barrier (A/B aid)
int bsf_flip() {
  const int f = g_force_kernel.load();
  return (f >= 3000 && f < 3004) ? f - 3000 : 0;
}

FlatGeometry flat_geometry(const DevCtx& c, size_t shard_len, size_t n_stripes) {
  FlatGeometry fg;
  fg.variant = bsf_threads();
  fg.threads = fg.variant & ~31;
  const uint64_t nw = (uint64_t)fg.threads / 32;
  fg.units_per_shard = (uint32_t)((shard_len + kBsfUnitBytes - 1) / kBsfUnitBytes);
  fg.total_units = (uint64_t)n_stripes * fg.units_per_shard;
  fg.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((fg.total_units + nw - 1) / nw, (uint64_t)c.sm_count));
  fg.total_warps = (uint64_t)fg.grid * nw;
  // runs that can touch one stripe (part slots per shard): runs are floor(U/GW) or one more units long
  const uint64_t U = fg.total_units, GW = fg.total_warps, wt = fg.units_per_shard;
  if (U >= GW) fg.max_parts = (uint32_t)((wt - 1) / (U / GW) + 2);
  else fg.max_parts = (uint32_t)((wt * GW + U - 1) / U + 2);   // more warps than units: empty runs in between
  return fg;
}
size_t flat_part_bytes(const FlatGeometry& fg, size_t n_stripes, int n_slots) {
  return n_stripes * (size_t)n_slots * fg.max_parts * sizeof(uint32_t);
}
// Is the flat kernel the path for this (handle, layout)?  Shards of at least 16 KiB (smaller ones: packed mode
// of rs_bs_kernel), every pass of the code instantiated.
bool flat_usable(const cubeec* h, size_t shard_len, int first_pass_mode) {
  const int f = g_force_kernel.load();
  if (f == 5 || f == 6 || f == 7 || !h->bs_passes || shard_len < 16384) return false;   // 5/6/7: the tile-split variants
  for (int pass = 0; pass < h->bs_passes; pass++)
    if (!bsf_supported(h->k, h->m, pass, pass == 0 ? first_pass_mode : 2)) return false;
  return true;
}

// All fused-CRC passes of handle h (crc: 1 = data + outputs in the first pass, 2 = outputs only) with the flat split.
int bsf_run(cubeec* h, DevCtx& c, cudaStream_t stream, uint8_t* d_base, size_t shard_len, size_t shard_pitch,
            size_t stripe_pitch, size_t n_stripes, const FlatGeometry& fg, int n_slots, const uint8_t* in_slots, int out_first,
            int crc, uint32_t* d_part, int crc_poly) {
  BsfParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.base = d_base;
  bp.stripe_pitch = stripe_pitch;
  bp.shard_pitch = shard_pitch;
  bp.shard_len = (uint32_t)shard_len;
  bp.n_stripes = (uint32_t)n_stripes;
  bp.n_slots = (uint32_t)n_slots;
  bp.units_per_shard = fg.units_per_shard;
  bp.total_units = fg.total_units;
  bp.max_parts = fg.max_parts;
  bp.crc_part = d_part;
  const int pi = crc_poly ? 1 : 0;
  bp.poly = g.poly[pi].poly;
  bp.slice_image = c.d_bs_slice[pi];
  bp.fold_tables = c.d_bsf_fold[pi];
  bp.klane = c.d_bsf_klane[pi];
  if (h->k > (int)sizeof(bp.in_slot)) return CUBEEC_ERR_UNSUPPORTED;
  if (crc == 1 && in_slots) return CUBEEC_ERR_INVALID_ARG;   // data CRCs only with the identity input map
  for (int i = 0; i < h->k; i++) bp.in_slot[i] = in_slots ? in_slots[i] : (uint8_t)i;
  for (int pass = 0; pass < h->bs_passes; pass++) {
    int r0 = 0, rows = h->m;
    if (h->bs_passes > 1 || h->m > 4) {
      if (!bs_mp_pass_rows(h->k, h->m, 0, pass, &r0, &rows)) return CUBEEC_ERR_UNSUPPORTED;
    }
    for (int r = 0; r < rows && r < (int)sizeof(bp.out_slot); r++) bp.out_slot[r] = (uint8_t)(out_first + r0 + r);
    const int mode = (crc == 1 && pass == 0) ? 1 : 2;
    if (fg.variant != kBsfThreads && h->k == 12 && h->m == 4 && mode == 1) CU(launch_bsf_variant(h->k, fg.variant, bp, fg.grid, stream));
    else CU(launch_bsf(h->k, h->m, pass, mode, bp, fg.grid, stream, bsf_flip()));
    g_launches++;
  }
  t_last_kernel = "rs_bsf_kernel<crc>";
  return CUBEEC_OK;
}

int finalize_flat_crc(cudaStream_t stream, const uint32_t* d_part, size_t n_stripes, int n_slots, size_t shard_len,
                      const FlatGeometry& fg, int crc_poly, uint32_t* d_out) {
  const CrcPoly& P = g.poly[crc_poly ? 1 : 0];
  CrcPartsFinalizeParams f;
  std::memset(&f, 0, sizeof(f));
  f.crc_part = d_part;
  f.n_stripes = (uint32_t)n_stripes;
  f.n_slots = (uint32_t)n_slots;
  f.max_parts = fg.max_parts;
  f.units_per_shard = fg.units_per_shard;
  f.total_units = fg.total_units;
  f.total_warps = fg.total_warps;
  f.shard_len = (uint32_t)shard_len;
  f.first_slot = 0;
  f.n_out = (uint32_t)n_slots;
  f.poly = P.poly;
  for (int i = 0; i < 24; i++) f.x_unit_pow[i] = P.shift_bytes_const((int64_t)kBsfUnitBytes << i);
  f.fix = P.shift_bytes_const(-((int64_t)fg.units_per_shard * kBsfUnitBytes - (int64_t)shard_len));
  f.init_term = P.mul(0xFFFFFFFFu, P.shift_bytes_const((int64_t)shard_len));
  f.out = d_out;
  CU(launch_crc_parts_finalize(f, stream));
  g_launches += 3;   // zero, scatter, finish
  return CUBEEC_OK;
}

// Bytes of CRC scratch (per-segment remainders) an encode of this geometry needs.
size_t crc_part_bytes(const DevCtx& c, size_t shard_len, size_t n_stripes, int n_slots) {
  // upper bound over both kernels' geometries
  const Geometry g1 = pick_geometry(c, shard_len, n_stripes, false);
  const Geometry g2 = pick_geometry(c, shard_len, n_stripes, false, kBsTile);
  const Geometry g3 = pick_geometry(c, shard_len, n_stripes, false, kBswTile);
  const size_t tiled = n_stripes * (size_t)n_slots * std::max<uint32_t>(std::max(std::max(g1.n_seg, g2.n_seg), g3.n_seg), 2u) * sizeof(uint32_t);
  size_t flat = 0;
  if (shard_len >= 16384) flat = flat_part_bytes(flat_geometry(c, shard_len, n_stripes), n_stripes, n_slots);
  return std::max(tiled, flat);
}

// Encode (mode 0) or verify (mode 1) a device-resident batch.  d_part: caller scratch of
// crc_part_bytes() when CRCs are wanted, or nullptr to use the stream-ordered allocator.
int dev_encode_impl(cubeec* h, DevCtx& c, cudaStream_t stream, uint8_t* d_base, size_t shard_len,
                    size_t shard_pitch, size_t stripe_pitch, size_t n_stripes, uint32_t* d_crc_out, int crc_poly,
                    int mode, int32_t* d_mismatch, uint32_t* d_part) {
  if (h->m == 0 && !d_crc_out) return CUBEEC_OK;
  if (h->m == 0) {
    // reedsolomon.New(N, 0) (the Replica code modes): nothing to code, the checksums are plain buffer CRCs
    if (mode != 0) return CUBEEC_OK;
    if (stripe_pitch == (size_t)h->k * shard_pitch)
      return dev_crc32_impl(c, stream, d_base, shard_len, shard_pitch, n_stripes * (size_t)h->k, 0, crc_poly, d_crc_out, nullptr);
    for (size_t s = 0; s < n_stripes; s++) {
      int rc = dev_crc32_impl(c, stream, d_base + s * stripe_pitch, shard_len, shard_pitch, (size_t)h->k, 0, crc_poly,
                              d_crc_out + s * (size_t)h->k, nullptr);
      if (rc) return rc;
    }
    return CUBEEC_OK;
  }
  const int n = h->k + h->m;
  const size_t ci = ctx_index(&c);
  const bool want_crc = mode == 0 && d_crc_out;
  if (h->bs_passes && g_force_kernel.load() != 1 && bs_layout_ok(d_base, shard_len, shard_pitch, stripe_pitch)) {
    if (want_crc && flat_usable(h, shard_len, 1)) {
      // fused encode + CRC32 with the flat work split (bs_flat.cuh)
      const FlatGeometry fg = flat_geometry(c, shard_len, n_stripes);
      AsyncScratch fscratch(stream);
      if (!d_part) {
        CU(fscratch.alloc(flat_part_bytes(fg, n_stripes, n)));
        d_part = static_cast<uint32_t*>(fscratch.ptr);
      }
      int rc = bsf_run(h, c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, fg, n, nullptr, h->k, 1, d_part,
                       crc_poly);
      if (rc) return rc;
      return finalize_flat_crc(stream, d_part, n_stripes, n, shard_len, fg, crc_poly, d_crc_out);
    }
    // bit-sliced XOR-network kernel with the CTA-tile split (bitslice.cu): plain encode, verify, packed mode
    const bool ws = want_crc && g_force_kernel.load() == 5 && h->bs_passes == 1 && bsw_supported(h->k, h->m) &&
                    !bs_is_packed(shard_len);
    const Geometry gm = ws ? pick_geometry(c, shard_len, n_stripes, false, kBswTile) : bs_geometry(c, shard_len, n_stripes, !want_crc);
    AsyncScratch scratch(stream);
    if (want_crc && !d_part) {
      CU(scratch.alloc(n_stripes * n * gm.n_seg * sizeof(uint32_t)));
      d_part = static_cast<uint32_t*>(scratch.ptr);
    }
    if (want_crc && gm.packed_pps) CU(cudaMemsetAsync(d_part, 0, n_stripes * n * gm.n_seg * sizeof(uint32_t), stream));
    const bool rolled = want_crc && g_force_kernel.load() == 6 && !gm.packed_pps && bs_rolled_supported(h->k, h->m);
    int rc = bs_run(h, c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, gm, n, nullptr, h->k,
                    want_crc ? 1 : 0, want_crc ? d_part : nullptr, crc_poly, mode == 1, d_mismatch, ws, rolled);
    if (rc) return rc;
    if (want_crc) {
      rc = finalize_crc(c, stream, d_part, n_stripes, n, shard_len, gm, crc_poly, nullptr, d_crc_out);
      if (rc) return rc;
    }
    return CUBEEC_OK;
  }
  const Geometry gm = pick_geometry(c, shard_len, n_stripes, false);
  const auto& passes = want_crc ? h->enc_passes : h->verify_passes;
  const Pattern* dp = want_crc ? h->d_enc[ci] : h->d_verify[ci];
  std::vector<const Pattern*> pp;
  std::vector<int> nin, ncrc;
  for (size_t j = 0; j < passes.size(); j++) {
    pp.push_back(dp + j);
    nin.push_back(passes[j].n_in);
    ncrc.push_back((passes[j].crc_in ? passes[j].n_in : 0) + passes[j].n_out);
  }
  AsyncScratch scratch(stream);
  if (want_crc && !d_part) {
    CU(scratch.alloc(n_stripes * n * gm.n_seg * sizeof(uint32_t)));
    d_part = static_cast<uint32_t*>(scratch.ptr);
  }
  const std::vector<char> exact(pp.size(), 1);
  int rc = run_passes(c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, n, pp, nin, ncrc, exact, nullptr,
                      mode, d_mismatch, want_crc ? d_part : nullptr, crc_poly, gm);
  if (rc) return rc;
  if (want_crc) {
    rc = finalize_crc(c, stream, d_part, n_stripes, n, shard_len, gm, crc_poly, nullptr, d_crc_out);
    if (rc) return rc;
  }
  return CUBEEC_OK;
}

}  // namespace

extern "C" void cubeec_debug_force_kernel(int which) { g_force_kernel.store(which); }

// CPU-side check of the run-time code generator (tests, no device needed): builds the decode rows of RS(k, m) for
// the presence pattern the way cubeec_dev_reconstruct does and compiles the specialised kernel with NVRTC.
// 0 = compiled, 1 = NVRTC not installed, 2 = compile error, else a CUBEEC_ERR_* of the pattern itself.
extern "C" int cubeec_debug_jit_check(int k, int m, const uint8_t* present, int data_only, char* log, size_t log_cap) {
  if (k <= 0 || m < 0 || k + m > 256 || !present) return CUBEEC_ERR_INVALID_ARG;
  std::vector<uint8_t> gen;
  if (!build_generator(k, k + m, gen)) return CUBEEC_ERR_SINGULAR;
  std::vector<int> valid, outs;
  std::vector<uint8_t> rows;
  int rc = decode_rows(k, m, gen, present, data_only != 0, valid, outs, rows);
  if (rc) return rc;
  std::string err;
  int result = 0;
  for (size_t o0 = 0; o0 < outs.size() && result == 0; o0 += kMaxOut) {
    const size_t no = std::min<size_t>(kMaxOut, outs.size() - o0);
    std::vector<uint8_t> ins(valid.begin(), valid.end()), os(outs.begin() + (long)o0, outs.begin() + (long)(o0 + no));
    std::vector<uint8_t> rr(rows.begin() + (long)(o0 * k), rows.begin() + (long)((o0 + no) * k));
    result = jit_compile_check(ins, os, rr, nullptr, &err);
  }
  if (log && log_cap) {
    std::strncpy(log, err.c_str(), log_cap - 1);
    log[log_cap - 1] = 0;
  }
  return result ? result + 100 : 0;   // 101 = NVRTC missing, 102 = compile error
}

// ------------------------------------------------------------------------------------------
// process-wide API
// ------------------------------------------------------------------------------------------
extern "C" int cubeec_init(const int* devices, int n_devices) {
  {
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.ready) return CUBEEC_ERR_INVALID_ARG;
    if (n_devices <= 0) return CUBEEC_ERR_INVALID_ARG;
    g.devices.clear();
    for (int i = 0; i < n_devices; i++) g.devices.push_back(devices ? devices[i] : i);
  }
  return ensure_init();
}

extern "C" int cubeec_device_count(void) {
  int rc = ensure_init();
  if (rc) return 0;
  return (int)g.ctx.size();
}

extern "C" int cubeec_host_alloc(size_t bytes, void** out) {
  int rc = ensure_init();
  if (rc) return rc;
  CU(cudaMallocHost(out, bytes));
  return CUBEEC_OK;
}
extern "C" int cubeec_host_free(void* p) {
  CU(cudaFreeHost(p));
  return CUBEEC_OK;
}
extern "C" int cubeec_host_register(void* p, size_t bytes) {
  int rc = ensure_init();
  if (rc) return rc;
  CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return CUBEEC_OK;
}
extern "C" int cubeec_host_unregister(void* p) {
  CU(cudaHostUnregister(p));
  return CUBEEC_OK;
}

// ------------------------------------------------------------------------------------------
// handle API
// ------------------------------------------------------------------------------------------
extern "C" int cubeec_create(int k, int m, const uint8_t* parity_rows, cubeec_t** out) {
  if (!out) return CUBEEC_ERR_INVALID_ARG;
  *out = nullptr;
  // argument checks first, in the reference's order (RS/reedsolomon.go:419-441)
  if (k + m > 256) return CUBEEC_ERR_MAX_SHARD_NUM;
  if (k <= 0 || m < 0) return CUBEEC_ERR_INV_SHARD_NUM;
  if (k > kMaxIn) return CUBEEC_ERR_UNSUPPORTED;
  int rc = ensure_init();
  if (rc) return rc;
  auto h = std::make_unique<cubeec>();
  h->k = k;
  h->m = m;
  if (m == 0 || parity_rows) {
    h->gen.assign((size_t)(k + m) * k, 0);
    for (int i = 0; i < k; i++) h->gen[(size_t)i * k + i] = 1;
    if (m) std::memcpy(&h->gen[(size_t)k * k], parity_rows, (size_t)m * k);
  } else if (!build_generator(k, k + m, h->gen)) {
    return CUBEEC_ERR_SINGULAR;
  }
  if (m > 0) {
    std::vector<int> ins, outs;
    for (int i = 0; i < k; i++) ins.push_back(i);
    for (int i = 0; i < m; i++) outs.push_back(k + i);
    std::vector<uint8_t> rows(h->gen.begin() + (size_t)k * k, h->gen.end());
    make_passes(ins, outs, rows, true, h->enc_passes);
    make_passes(ins, outs, rows, false, h->verify_passes);
    rc = upload_handle_patterns(h.get());
    if (rc) return rc;
    h->bs_passes = bs_passes(k, m, rows.data(), 0);
    h->bs_passes_plain = bs_passes(k, m, rows.data(), 1);
    if (!h->bs_passes || !h->bs_passes_plain) h->bs_passes = h->bs_passes_plain = 0;
  }
  *out = h.release();
  return CUBEEC_OK;
}

extern "C" void cubeec_destroy(cubeec_t* h) {
  if (!h) return;
  for (auto& pl : h->plans) {
    cudaSetDevice(pl->device);
    if (pl->d_pat) cudaFree(pl->d_pat);
    if (pl->d_rec) cudaFree(pl->d_rec);
    if (pl->d_pos) cudaFree(pl->d_pos);
  }
  for (size_t ci = 0; ci < h->d_enc.size() && ci < g.ctx.size(); ci++) {
    cudaSetDevice(g.ctx[ci]->device);
    if (h->d_enc[ci]) cudaFree(h->d_enc[ci]);
    if (h->d_verify[ci]) cudaFree(h->d_verify[ci]);
  }
  delete h;
}
extern "C" int cubeec_k(const cubeec_t* h) { return h ? h->k : 0; }
extern "C" int cubeec_m(const cubeec_t* h) { return h ? h->m : 0; }
extern "C" int cubeec_matrix(const cubeec_t* h, uint8_t* out) {
  if (!h || !out) return CUBEEC_ERR_INVALID_ARG;
  std::memcpy(out, h->gen.data(), h->gen.size());
  return CUBEEC_OK;
}
extern "C" int cubeec_decode_matrix(const cubeec_t* h, const uint8_t* present, int* valid, uint8_t* rows) {
  if (!h || !present || !valid || !rows) return CUBEEC_ERR_INVALID_ARG;
  const int k = h->k, n = h->k + h->m;
  int cnt = 0;
  for (int i = 0; i < n && cnt < k; i++)
    if (present[i]) valid[cnt++] = i;
  if (cnt < k) return CUBEEC_ERR_TOO_FEW_SHARDS;
  std::vector<uint8_t> sub((size_t)k * k);
  for (int r = 0; r < k; r++) std::memcpy(&sub[(size_t)r * k], &h->gen[(size_t)valid[r] * k], (size_t)k);
  return gf_invert(sub.data(), k, rows) ? CUBEEC_OK : CUBEEC_ERR_SINGULAR;
}

// ------------------------------------------------------------------------------------------
// device-resident API
// ------------------------------------------------------------------------------------------
static int check_dev_layout(const void* d_base, size_t shard_len, size_t shard_pitch, size_t stripe_pitch, int n_slots = 0,
                            size_t n_stripes = 0) {
  if (!d_base || shard_len == 0) return CUBEEC_ERR_INVALID_ARG;
  if (((uintptr_t)d_base & 15) || (shard_pitch & 15) || (stripe_pitch & 15) || shard_pitch < shard_len)
    return CUBEEC_ERR_INVALID_ARG;
  if (shard_len > 0xFFFFFFF0ull) return CUBEEC_ERR_UNSUPPORTED;
  // stripes must not overlap, and the kernels count stripes / work items in 32 bits
  if (n_slots > 0 && stripe_pitch < (size_t)n_slots * shard_pitch) return CUBEEC_ERR_INVALID_ARG;
  if (n_stripes > 0) {
    const uint64_t units = (shard_len + 1023) / 1024;   // upper bound of work items per stripe over all kernels
    if (n_stripes >= (1ull << 31) || n_stripes * (uint64_t)std::max(n_slots, 1) >= (1ull << 32) || n_stripes * units >= (1ull << 40))
      return CUBEEC_ERR_UNSUPPORTED;
  }
  return CUBEEC_OK;
}

extern "C" int cubeec_dev_encode(cubeec_t* h, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                                 size_t stripe_pitch, size_t n_stripes, uint32_t* d_crc_out, int crc_poly,
                                 void* stream) {
  if (!h) return CUBEEC_ERR_INVALID_ARG;
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_dev_layout(d_base, shard_len, shard_pitch, stripe_pitch, h->k + h->m, n_stripes))) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  if (stream) {
    return dev_encode_impl(h, *c, (cudaStream_t)stream, (uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch,
                           n_stripes, d_crc_out, crc_poly, 0, nullptr, nullptr);
  }
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  uint32_t* d_part = nullptr;
  if (d_crc_out) {
    if ((rc = lane_reserve(*lease.lane, 0, crc_part_bytes(*c, shard_len, n_stripes, h->k + h->m)))) return rc;
    d_part = reinterpret_cast<uint32_t*>(lease.lane->d_aux);
  }
  rc = dev_encode_impl(h, *c, lease.lane->stream, (uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch, n_stripes,
                       d_crc_out, crc_poly, 0, nullptr, d_part);
  if (rc) return rc;
  CU(cudaStreamSynchronize(lease.lane->stream));
  return CUBEEC_OK;
}

// ------------------------------------------------------------------------------------------
// LRC code modes (lrcEncoder.Encode, blobstore/common/ec/lrcencoder.go:35-80): global RS(N, M) over the
// stripe, then per AZ a local RS((N+M)/AZ, L/AZ) over that AZ's data + global-parity shards
// (codemode.GetECLayoutByAZ, codemode.go:301-318; e.g. EC6P10L2: AZ 0 = shards [0,1,2, 6..10] -> 16).
// All passes run on the device-resident stripe: the data crosses PCIe once, every shard is
// checksummed once (the local passes checksum only what they write).
// ------------------------------------------------------------------------------------------
namespace {
struct LrcLayout {
  int N, M, L, az, kl, ml;
};

int lrc_layout(const cubeec* hg, const cubeec* hl, int az, LrcLayout* y) {
  if (!hg || !hl || az <= 0) return CUBEEC_ERR_INVALID_ARG;
  y->N = hg->k;
  y->M = hg->m;
  y->az = az;
  y->kl = hl->k;
  y->ml = hl->m;
  y->L = hl->m * az;
  if (y->N % az || y->M % az || (y->N + y->M) / az != y->kl || y->ml <= 0) return CUBEEC_ERR_INVALID_ARG;
  return CUBEEC_OK;
}

int dev_lrc_encode_impl(cubeec* hg, cubeec* hl, const LrcLayout& y, DevCtx& c, cudaStream_t stream, uint8_t* d_base,
                        size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes,
                        uint32_t* d_crc_out, int crc_poly, uint32_t* d_part) {
  const int n_slots = y.N + y.M + y.L;
  if (!hg->bs_passes || !hl->bs_passes || !bs_layout_ok(d_base, shard_len, shard_pitch, stripe_pitch)) {
    t_last_error = "LRC device path needs the generated networks of both codes and a 32-byte aligned layout";
    return CUBEEC_ERR_UNSUPPORTED;
  }
  const bool want_crc = d_crc_out != nullptr;
  if (want_crc && flat_usable(hg, shard_len, 1) && flat_usable(hl, shard_len, 2)) {
    // every code of the stripe with the flat split: the global passes checksum data + global parity, each
    // AZ's local pass the local parity it writes; one finalize over all N+M+L slots
    const FlatGeometry fg = flat_geometry(c, shard_len, n_stripes);
    AsyncScratch fscratch(stream);
    if (!d_part) {
      CU(fscratch.alloc(flat_part_bytes(fg, n_stripes, n_slots)));
      d_part = static_cast<uint32_t*>(fscratch.ptr);
    }
    int rc = bsf_run(hg, c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, fg, n_slots, nullptr, y.N, 1, d_part,
                     crc_poly);
    if (rc) return rc;
    for (int a = 0; a < y.az; a++) {
      uint8_t in_slots[64];
      int q = 0;
      for (int i = 0; i < y.N / y.az; i++) in_slots[q++] = (uint8_t)(a * (y.N / y.az) + i);
      for (int i = 0; i < y.M / y.az; i++) in_slots[q++] = (uint8_t)(y.N + a * (y.M / y.az) + i);
      rc = bsf_run(hl, c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, fg, n_slots, in_slots,
                   y.N + y.M + a * y.ml, 2, d_part, crc_poly);
      if (rc) return rc;
    }
    return finalize_flat_crc(stream, d_part, n_stripes, n_slots, shard_len, fg, crc_poly, d_crc_out);
  }
  const Geometry gm = bs_geometry(c, shard_len, n_stripes);
  AsyncScratch scratch(stream);
  if (want_crc && !d_part) {
    CU(scratch.alloc(n_stripes * n_slots * gm.n_seg * sizeof(uint32_t)));
    d_part = static_cast<uint32_t*>(scratch.ptr);
  }
  if (want_crc && gm.packed_pps) CU(cudaMemsetAsync(d_part, 0, n_stripes * n_slots * gm.n_seg * sizeof(uint32_t), stream));
  int rc = bs_run(hg, c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, gm, n_slots, nullptr, y.N,
                  want_crc ? 1 : 0, d_part, crc_poly, false, nullptr);
  if (rc) return rc;
  for (int a = 0; a < y.az; a++) {
    uint8_t in_slots[64];
    int q = 0;
    for (int i = 0; i < y.N / y.az; i++) in_slots[q++] = (uint8_t)(a * (y.N / y.az) + i);
    for (int i = 0; i < y.M / y.az; i++) in_slots[q++] = (uint8_t)(y.N + a * (y.M / y.az) + i);
    rc = bs_run(hl, c, stream, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, gm, n_slots, in_slots,
                y.N + y.M + a * y.ml, want_crc ? 2 : 0, d_part, crc_poly, false, nullptr);
    if (rc) return rc;
  }
  if (want_crc) {
    rc = finalize_crc(c, stream, d_part, n_stripes, n_slots, shard_len, gm, crc_poly, nullptr, d_crc_out);
    if (rc) return rc;
  }
  return CUBEEC_OK;
}
}  // namespace

extern "C" int cubeec_dev_lrc_encode(cubeec_t* global, cubeec_t* local, int az_count, int device, void* d_base,
                                     size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes,
                                     uint32_t* d_crc_out, int crc_poly, void* stream) {
  LrcLayout y;
  int rc = lrc_layout(global, local, az_count, &y);
  if (rc) return rc;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_dev_layout(d_base, shard_len, shard_pitch, stripe_pitch, y.N + y.M + y.L, n_stripes))) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  if (stream)
    return dev_lrc_encode_impl(global, local, y, *c, (cudaStream_t)stream, (uint8_t*)d_base, shard_len, shard_pitch,
                               stripe_pitch, n_stripes, d_crc_out, crc_poly, nullptr);
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  uint32_t* d_part = nullptr;
  if (d_crc_out) {
    if ((rc = lane_reserve(*lease.lane, 0, crc_part_bytes(*c, shard_len, n_stripes, y.N + y.M + y.L)))) return rc;
    d_part = reinterpret_cast<uint32_t*>(lease.lane->d_aux);
  }
  rc = dev_lrc_encode_impl(global, local, y, *c, lease.lane->stream, (uint8_t*)d_base, shard_len, shard_pitch,
                           stripe_pitch, n_stripes, d_crc_out, crc_poly, d_part);
  if (rc) return rc;
  CU(cudaStreamSynchronize(lease.lane->stream));
  return CUBEEC_OK;
}

static int dev_reconstruct_core(cubeec_t* h, DevCtx* c, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                                size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int data_only, void* stream,
                                const uint8_t* slot_map);

// ------------------------------------------------------------------------------------------
// LRC Verify / Reconstruct on the device-resident stripe (lrcEncoder.Verify / Reconstruct / ReconstructData,
// blobstore/common/ec/lrcencoder.go:87-200): the global RS(N, M) over the first N+M shards, then per AZ the local
// RS((N+M)/AZ, L/AZ) over that AZ's shard list (codemode.GetECLayoutByAZ) -- the composition the reference does with
// 1 + AZCount engine calls, each of which would stage its shards over PCIe again through the plain ABI.
// ------------------------------------------------------------------------------------------
extern "C" int cubeec_dev_lrc_verify(cubeec_t* global, cubeec_t* local, int az_count, int device, const void* d_base,
                                     size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes, int32_t* d_ok,
                                     void* stream) {
  LrcLayout y;
  int rc = lrc_layout(global, local, az_count, &y);
  if (rc) return rc;
  if (!d_ok) return CUBEEC_ERR_INVALID_ARG;
  if ((rc = ensure_init())) return rc;
  if ((rc = check_dev_layout(d_base, shard_len, shard_pitch, stripe_pitch, y.N + y.M + y.L, n_stripes))) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  if (!global->bs_passes_plain || !local->bs_passes_plain || !bs_layout_ok((const uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch)) {
    t_last_error = "LRC device path needs the generated networks of both codes and a 32-byte aligned layout";
    return CUBEEC_ERR_UNSUPPORTED;
  }
  LaneLease lease;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    if ((rc = lease.acquire(c))) return rc;
    st = lease.lane->stream;
  }
  const int n_slots = y.N + y.M + y.L;
  const Geometry gm = bs_geometry(*c, shard_len, n_stripes, true);
  // d_ok doubles as the mismatch flag array: every code of the stripe raises the same flag, then it is inverted
  CU(cudaMemsetAsync(d_ok, 0, n_stripes * sizeof(int32_t), st));
  rc = bs_run(global, *c, st, (uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, gm, n_slots, nullptr, y.N, 0, nullptr, 0,
              true, d_ok);
  if (rc) return rc;
  for (int a = 0; a < y.az; a++) {
    uint8_t in_slots[64];
    int q = 0;
    for (int i = 0; i < y.N / y.az; i++) in_slots[q++] = (uint8_t)(a * (y.N / y.az) + i);
    for (int i = 0; i < y.M / y.az; i++) in_slots[q++] = (uint8_t)(y.N + a * (y.M / y.az) + i);
    rc = bs_run(local, *c, st, (uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, gm, n_slots, in_slots,
                y.N + y.M + a * y.ml, 0, nullptr, 0, true, d_ok);
    if (rc) return rc;
  }
  CU(launch_invert_flags(d_ok, n_stripes, st));
  g_launches++;
  if (!stream) CU(cudaStreamSynchronize(st));
  return CUBEEC_OK;
}

// present: HOST array n_stripes*(N+M+L).  Global shards are regenerated from the global code (as lrcEncoder.Reconstruct
// does first, lrcencoder.go:152-157: local parity never helps the global decode), then every missing local parity shard
// from its AZ's (now complete) shards (:159-183).  data_only = lrcEncoder.ReconstructData (:185-200): global data only.
extern "C" int cubeec_dev_lrc_reconstruct(cubeec_t* global, cubeec_t* local, int az_count, int device, void* d_base,
                                          size_t shard_len, size_t shard_pitch, size_t stripe_pitch, size_t n_stripes,
                                          const uint8_t* present, int data_only, void* stream) {
  LrcLayout y;
  int rc = lrc_layout(global, local, az_count, &y);
  if (rc) return rc;
  if (!present) return CUBEEC_ERR_INVALID_ARG;
  if ((rc = ensure_init())) return rc;
  const int n_slots = y.N + y.M + y.L, ng = y.N + y.M, nl = y.kl + y.ml;
  if ((rc = check_dev_layout(d_base, shard_len, shard_pitch, stripe_pitch, n_slots, n_stripes))) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  std::vector<uint8_t> pg(n_stripes * (size_t)ng);
  for (size_t s = 0; s < n_stripes; s++)
    for (int i = 0; i < ng; i++) pg[s * ng + i] = present[s * n_slots + i] ? 1 : 0;
  rc = dev_reconstruct_core(global, c, device, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, pg.data(), data_only, stream, nullptr);
  if (rc || data_only) return rc;
  for (int a = 0; a < y.az; a++) {
    std::vector<uint8_t> map(nl), pl(n_stripes * (size_t)nl, 1);
    int q = 0;
    for (int i = 0; i < y.N / y.az; i++) map[q++] = (uint8_t)(a * (y.N / y.az) + i);
    for (int i = 0; i < y.M / y.az; i++) map[q++] = (uint8_t)(y.N + a * (y.M / y.az) + i);
    for (int i = 0; i < y.ml; i++) map[q++] = (uint8_t)(ng + a * y.ml + i);
    bool any = false;
    for (size_t s = 0; s < n_stripes; s++)
      for (int i = 0; i < y.ml; i++)
        if (!present[s * n_slots + ng + a * y.ml + i]) { pl[s * nl + y.kl + i] = 0; any = true; }
    if (!any) continue;
    rc = dev_reconstruct_core(local, c, device, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, pl.data(), 0, stream, map.data());
    if (rc) return rc;
  }
  return CUBEEC_OK;
}

extern "C" int cubeec_dev_verify(cubeec_t* h, int device, const void* d_base, size_t shard_len, size_t shard_pitch,
                                 size_t stripe_pitch, size_t n_stripes, int32_t* d_ok, void* stream) {
  if (!h || !d_ok) return CUBEEC_ERR_INVALID_ARG;
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_dev_layout(d_base, shard_len, shard_pitch, stripe_pitch, h->k + h->m, n_stripes))) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  LaneLease lease;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    if ((rc = lease.acquire(c))) return rc;
    st = lease.lane->stream;
  }
  // d_ok doubles as the mismatch flag array: kernel sets 1 on mismatch, then it is inverted.
  CU(cudaMemsetAsync(d_ok, 0, n_stripes * sizeof(int32_t), st));
  rc = dev_encode_impl(h, *c, st, (uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, nullptr, 0, 1,
                       d_ok, nullptr);
  if (rc) return rc;
  CU(launch_invert_flags(d_ok, n_stripes, st));   // ok = !mismatch
  g_launches++;
  if (!stream) CU(cudaStreamSynchronize(st));
  return CUBEEC_OK;
}

// slot_map (k+m entries, or nullptr = identity): shard i of the code is shard slot_map[i] of the stripe -- the local
// stripes of the LRC code modes (codemode.GetECLayoutByAZ) reconstruct through it.
static int dev_reconstruct_core(cubeec_t* h, DevCtx* c, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                                size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int data_only, void* stream,
                                const uint8_t* slot_map);

extern "C" int cubeec_dev_reconstruct(cubeec_t* h, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                                      size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int data_only,
                                      void* stream) {
  if (!h || !present) return CUBEEC_ERR_INVALID_ARG;
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_dev_layout(d_base, shard_len, shard_pitch, stripe_pitch, h->k + h->m, n_stripes))) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  return dev_reconstruct_core(h, c, device, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, present, data_only, stream, nullptr);
}

static int dev_reconstruct_core(cubeec_t* h, DevCtx* c, int device, void* d_base, size_t shard_len, size_t shard_pitch,
                                size_t stripe_pitch, size_t n_stripes, const uint8_t* present, int data_only, void* stream,
                                const uint8_t* slot_map) {
  int rc = CUBEEC_OK;
  const int n = h->k + h->m;
  const std::vector<uint8_t> smap = slot_map ? std::vector<uint8_t>(slot_map, slot_map + n) : std::vector<uint8_t>();
  // plan lookup (same presence array as an earlier call on this device?)
  cubeec::Plan* plan = nullptr;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    for (auto& pl : h->plans)
      if (pl->device == device && pl->data_only == (data_only != 0) && pl->present.size() == n_stripes * (size_t)n &&
          pl->slot_map == smap && std::memcmp(pl->present.data(), present, pl->present.size()) == 0) {
        plan = pl.get();
        break;
      }
  }
  // Kernel choice.  One pattern for the whole batch -> the run-time compiled kernel of that pattern (jit.cu), below.
  // Mixed patterns -> the table kernels.  A/B: 9 = flat-split bit-sliced syndrome kernel (bitslice_syn.cu), 2 = the round-1
  // syndrome kernel, 1 / 3 = table kernels for everything.
  const int fkr = g_force_kernel.load();
  static const bool jit_off = getenv("CUBEEC_NO_JIT") != nullptr;
  bool single = true;
  for (size_t s = 1; s < n_stripes && single; s++)
    for (int i = 0; i < n; i++)
      if ((present[s * n + i] != 0) != (present[i] != 0)) { single = false; break; }
  const bool jit_ok = single && !jit_off && fkr != 1 && fkr != 3 && fkr != 4 && fkr != 2 && fkr != 9 && jit_available();
  const bool layout32 = bs_layout_ok((const uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch);
  const bool use_old_rec = fkr == 2 && !slot_map && h->bs_passes == 1 && bs_rec_supported(h->k, h->m) && layout32;
  // (the flat-split syndrome kernel is measured at 0.45 of HBM on C3 against 0.65 for the table kernel: opt-in, force 9)
  const bool use_syn = fkr == 9 && !slot_map && h->bs_passes == 1 && bs_syn_supported(h->k, h->m) && layout32;
  const bool use_rec = use_old_rec || use_syn;
  if (plan && use_rec != (plan->d_rec != nullptr)) plan = nullptr;
  std::unique_ptr<cubeec::Plan> fresh;
  if (!plan && use_rec) {
    std::map<std::string, uint32_t> ids;
    std::vector<RecPattern> pats;
    std::vector<uint32_t> pos(n_stripes);
    for (size_t s = 0; s < n_stripes; s++) {
      std::string key((const char*)present + s * n, (size_t)n);
      for (auto& ch : key) ch = ch ? 1 : 0;
      auto it = ids.find(key);
      if (it == ids.end()) {
        RecPattern rp;
        rc = rec_pattern(h, (const uint8_t*)key.data(), data_only != 0, rp);
        if (rc) return rc;
        it = ids.emplace(key, (uint32_t)pats.size()).first;
        pats.push_back(rp);
      }
      pos[s] = it->second;
    }
    fresh = std::make_unique<cubeec::Plan>();
    fresh->device = device;
    fresh->data_only = data_only != 0;
    fresh->present.assign(present, present + n_stripes * (size_t)n);
    fresh->n_pat = pats.size();
    CU(cudaMalloc(&fresh->d_rec, pats.size() * sizeof(RecPattern)));
    CU(cudaMalloc(&fresh->d_pos, n_stripes * sizeof(uint32_t)));
    CU(cudaMemcpy(fresh->d_rec, pats.data(), pats.size() * sizeof(RecPattern), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(fresh->d_pos, pos.data(), n_stripes * sizeof(uint32_t), cudaMemcpyHostToDevice));
    plan = fresh.get();
  }
  if (plan && plan->d_rec) {
    LaneLease lease2;
    cudaStream_t st2 = (cudaStream_t)stream;
    if (!st2) {
      if ((rc = lease2.acquire(c))) return rc;
      st2 = lease2.lane->stream;
    }
    const Geometry gm = pick_geometry(*c, shard_len, n_stripes, plan->n_pat > 1, kBsTile);
    BsRecParams rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.base = (uint8_t*)d_base;
    rp.stripe_pitch = stripe_pitch;
    rp.shard_pitch = shard_pitch;
    rp.shard_len = (uint32_t)shard_len;
    rp.n_stripes = (uint32_t)n_stripes;
    rp.n_seg = gm.n_seg;
    rp.tiles_per_seg = gm.tiles_per_seg;
    rp.tiles_last = gm.tiles_last;
    rp.patterns = plan->d_rec;
    rp.pattern_of_stripe = plan->d_pos;
    rp.gf = c->d_gf;
    if (use_syn) {
      rp.units_per_shard = bs_syn_units_per_shard(shard_len);
      rp.total_units = (uint64_t)n_stripes * rp.units_per_shard;
      const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((rp.total_units + 15) / 16, (uint64_t)c->sm_count));
      CU(launch_bs_syn(h->k, h->m, rp, grid, st2));
      t_last_kernel = "rs_bssyn_kernel";
    } else {
      CU(launch_bs_rec(h->k, h->m, rp, gm.grid, st2));
      t_last_kernel = "rs_bsrec_kernel";
    }
    g_launches++;
    if (fresh) {
      std::lock_guard<std::mutex> lk(h->mu);
      if (h->plans.size() < 16) {
        h->plans.push_back(std::move(fresh));
      } else {
        CU(cudaStreamSynchronize(st2));
        cudaFree(fresh->d_rec);
        cudaFree(fresh->d_pos);
      }
    }
    if (!stream) CU(cudaStreamSynchronize(st2));
    return CUBEEC_OK;
  }
  if (!plan) {
    // distinct presence patterns -> pattern ids
    std::map<std::string, uint32_t> ids;
    std::vector<std::vector<Pattern>> pat_passes;
    std::vector<uint32_t> pos(n_stripes);
    size_t n_pass = 0;
    for (size_t s = 0; s < n_stripes; s++) {
      std::string key((const char*)present + s * n, (size_t)n);
      for (auto& ch : key) ch = ch ? 1 : 0;
      auto it = ids.find(key);
      if (it == ids.end()) {
        std::vector<Pattern> passes;
        rc = decode_passes(h, (const uint8_t*)key.data(), data_only != 0, passes);
        if (rc) return rc;
        n_pass = std::max(n_pass, passes.size());
        it = ids.emplace(key, (uint32_t)pat_passes.size()).first;
        pat_passes.push_back(std::move(passes));
      }
      pos[s] = it->second;
    }
    // drop passes that do nothing for every pattern
    while (n_pass > 0) {
      bool any = false;
      for (auto& pp : pat_passes)
        if (pp.size() >= n_pass && pp[n_pass - 1].n_out) any = true;
      if (any) break;
      n_pass--;
    }
    if (n_pass == 0) return CUBEEC_OK;
    const size_t n_pat = pat_passes.size();
    std::vector<Pattern> flat(n_pass * n_pat);
    std::memset(flat.data(), 0, flat.size() * sizeof(Pattern));
    fresh = std::make_unique<cubeec::Plan>();
    fresh->device = device;
    fresh->data_only = data_only != 0;
    fresh->present.assign(present, present + n_stripes * (size_t)n);
    fresh->n_pass = n_pass;
    fresh->n_pat = n_pat;
    fresh->nin.assign(n_pass, 0);
    for (size_t j = 0; j < n_pass; j++)
      for (size_t q = 0; q < n_pat; q++)
        if (j < pat_passes[q].size()) {
          flat[j * n_pat + q] = pat_passes[q][j];
          if (slot_map) {
            Pattern& fp = flat[j * n_pat + q];
            for (int ci = 0; ci < fp.n_in; ci++) fp.in_slot[ci] = slot_map[fp.in_slot[ci]];
            for (int ri = 0; ri < fp.n_out; ri++) fp.out_slot[ri] = slot_map[fp.out_slot[ri]];
          }
          fresh->nin[j] = std::max<int>(fresh->nin[j], pat_passes[q][j].n_in);
        }
    CU(cudaMalloc(&fresh->d_pat, flat.size() * sizeof(Pattern)));
    CU(cudaMalloc(&fresh->d_pos, n_stripes * sizeof(uint32_t)));
    CU(cudaMemcpy(fresh->d_pat, flat.data(), flat.size() * sizeof(Pattern), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(fresh->d_pos, pos.data(), n_stripes * sizeof(uint32_t), cudaMemcpyHostToDevice));
    fresh->h_pat = flat;
    fresh->slot_map = smap;
    plan = fresh.get();
  }
  // One erasure pattern for the whole batch (a repair task: one broken vuid, worker_slice_recover.go:822-871):
  // run the kernel compiled for exactly these decode rows (jit.cu); ~1 s once per pattern, cached.
  if (plan->n_pat == 1 && jit_ok && layout32) {
    if (plan->jit_state == 0) {
      std::vector<const void*> ks;
      bool ok = true;
      for (size_t j = 0; j < plan->n_pass && ok; j++) {
        const Pattern& pt = plan->h_pat[j];
        if (pt.n_out == 0) { ks.push_back(nullptr); continue; }
        std::vector<uint8_t> ins(pt.in_slot, pt.in_slot + pt.n_in), outs(pt.out_slot, pt.out_slot + pt.n_out), rows;
        for (int r = 0; r < pt.n_out; r++) rows.insert(rows.end(), pt.coef[r], pt.coef[r] + pt.n_in);
        std::string err;
        const void* kf = jit_kernel(device, ins, outs, rows, &err);
        if (!kf) { ok = false; t_last_error = err; }
        ks.push_back(kf);
      }
      std::lock_guard<std::mutex> lk(h->mu);
      plan->jit = ks;
      plan->jit_state = ok ? 1 : 2;
    }
    if (plan->jit_state == 1) {
      LaneLease jl;
      cudaStream_t js = (cudaStream_t)stream;
      if (!js) {
        if ((rc = jl.acquire(c))) return rc;
        js = jl.lane->stream;
      }
      JitParams jp;
      std::memset(&jp, 0, sizeof(jp));
      jp.base = (uint8_t*)d_base;
      jp.stripe_pitch = stripe_pitch;
      jp.shard_pitch = shard_pitch;
      jp.shard_len = (uint32_t)shard_len;
      jp.n_stripes = (uint32_t)n_stripes;
      jp.units_per_shard = (uint32_t)((shard_len + 1023) / 1024);
      jp.total_units = (uint64_t)n_stripes * jp.units_per_shard;
      const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((jp.total_units + 15) / 16, (uint64_t)c->sm_count));
      for (size_t j = 0; j < plan->n_pass; j++) {
        if (!plan->jit[j]) continue;
        CU(jit_launch(plan->jit[j], jp, grid, js));
        g_launches++;
      }
      t_last_kernel = "rs_jit_kernel";
      if (fresh) {
        std::lock_guard<std::mutex> lk(h->mu);
        if (h->plans.size() < 16) {
          h->plans.push_back(std::move(fresh));
        } else {
          CU(cudaStreamSynchronize(js));
          cudaFree(fresh->d_pat);
          cudaFree(fresh->d_pos);
        }
      }
      if (!stream) CU(cudaStreamSynchronize(js));
      return CUBEEC_OK;
    }
  }
  LaneLease lease;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    if ((rc = lease.acquire(c))) return rc;
    st = lease.lane->stream;
  }
  std::vector<const Pattern*> pp;
  for (size_t j = 0; j < plan->n_pass; j++) pp.push_back(plan->d_pat + j * plan->n_pat);
  std::vector<int> ncrc(plan->n_pass, 0);
  const Geometry gm = pick_geometry(*c, shard_len, n_stripes, plan->n_pat > 1);
  const std::vector<char> exact(plan->n_pass, 1);   // no-op patterns (n_out = 0) are skipped by both kernels
  rc = run_passes(*c, st, (uint8_t*)d_base, shard_len, shard_pitch, stripe_pitch, n_stripes, n, pp, plan->nin, ncrc, exact,
                  plan->d_pos, 0, nullptr, nullptr, 0, gm);
  if (rc) return rc;
  if (fresh) {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->plans.size() < 16) {
      h->plans.push_back(std::move(fresh));
    } else {
      // cache full: this plan is single-use; release it once the stream has consumed it
      CU(cudaStreamSynchronize(st));
      cudaFree(fresh->d_pat);
      cudaFree(fresh->d_pos);
    }
  }
  if (!stream) CU(cudaStreamSynchronize(st));
  return CUBEEC_OK;
}

// ------------------------------------------------------------------------------------------
// Coalescing submit queue for the single-stripe host calls.
//
// How access calls the codec (blobstore/common/ec/encoder.go:114-131): ONE stripe per Encode call, at most 4
// blobs of a request in flight (access/stream/stream_put.go:105-110), up to defaultEncoderConcurrency = 1000
// goroutines in the call at once (access/stream/config_defaulter.go:24).  A cgo call pins an OS thread for its
// duration, so the natural shape is: the caller threads block, the library forms the batches.
//
//   caller thread   claim a slot of the OPEN batch of its (handle, shard size, checksum) key  -> copy its k data
//                   shards from pageable memory into the slot of the batch's PINNED staging buffer (the copies of
//                   all callers run in parallel, on the callers' own cores) -> wait for the batch -> copy its m
//                   parity shards (+ CRCs) out of the pinned buffer.
//   worker threads  (kCoWorkers per device, one lane/stream each) take a batch once it is full or its deadline
//                   (first claim + delay) has passed and every claimed slot is filled: ONE 2-D H2D copy, the same
//                   device encode as cubeec_dev_encode (fused CRC), ONE 2-D D2H copy of the parity columns.
//                   Three workers keep the H2D of one batch, the kernels of the next and the D2H of a third in flight.
//
// cubeec_set_coalescing(max_batch, delay_us): max_batch <= 1 switches the queue off (every call then does its own
// H2D / kernel / D2H round trip, the round-1 behaviour).
// ------------------------------------------------------------------------------------------
namespace {

constexpr int kCoWorkers = 3;
constexpr size_t kCoBatchBytes = 160u << 20;   // pinned staging of one batch

struct CoBatch {
  cubeec* h = nullptr;
  size_t S = 0, P = 0;
  int n = 0, poly = 0;
  bool want_crc = false;
  int op = 0;                   // 0 = encode, 1 = reconstruct
  bool data_only = false;       // reconstruct: ReconstructData
  std::vector<uint8_t> present; // reconstruct: [cap][n] presence flags of the slots
  int cap = 0, claimed = 0, filled = 0, left = 0;   // left: callers that still have to copy their results out
  enum { OPEN, CLOSED, RUNNING, DONE } state = OPEN;
  std::chrono::steady_clock::time_point deadline;
  uint8_t* h_buf = nullptr;     // pinned: [cap][n][P]
  size_t h_cap = 0;
  uint32_t* h_crc = nullptr;    // pinned: [cap][n]
  size_t crc_cap = 0;
  int rc = CUBEEC_OK;
  std::string err;
  std::condition_variable done_cv;
};

struct CoQueue {
  std::mutex mu;
  std::condition_variable work_cv;
  std::vector<CoBatch*> open;       // batches that still accept claims
  std::vector<CoBatch*> pending;    // closed or open-with-deadline batches waiting for a worker (FIFO)
  std::vector<CoBatch*> pool;       // idle batch objects (their pinned buffers are kept)
  std::vector<std::thread> workers;
  bool started = false, stop = false;
  DevCtx* ctx = nullptr;
};

std::atomic<int> g_co_max_batch{32};
std::atomic<int> g_co_delay_us{100};
CoQueue g_co;   // device 0 of the engine (the single-stripe host calls always use the first configured device)

// a reconstruct batch: survivors in (runs of consecutive shards merged into one copy), one batched device reconstruct
// (one pattern for all slots -> the run-time compiled kernel), regenerated shards out
int co_run_reconstruct(CoBatch& b, DevCtx& c, Lane& l) {
  const size_t dstripe = b.P * b.n;
  const int n = b.n, k = b.h->k;
  int rc = lane_reserve(l, dstripe * b.claimed, 256);
  if (rc) return rc;
  CU(cudaSetDevice(c.device));
  auto copy_runs = [&](bool to_device) -> int {
    for (int s = 0; s < b.claimed; s++) {
      const uint8_t* pr = &b.present[(size_t)s * n];
      for (int i = 0; i < n;) {
        const bool pick = to_device ? pr[i] != 0 : (pr[i] == 0 && !(b.data_only && i >= k));
        if (!pick) { i++; continue; }
        int j = i + 1;
        while (j < n && (to_device ? pr[j] != 0 : (pr[j] == 0 && !(b.data_only && j >= k)))) j++;
        uint8_t* hp = b.h_buf + (size_t)s * dstripe + (size_t)i * b.P;
        uint8_t* dp = l.d_buf + (size_t)s * dstripe + (size_t)i * b.P;
        if (to_device) CU(cudaMemcpyAsync(dp, hp, (size_t)(j - i) * b.P, cudaMemcpyHostToDevice, l.stream));
        else CU(cudaMemcpyAsync(hp, dp, (size_t)(j - i) * b.P, cudaMemcpyDeviceToHost, l.stream));
        i = j;
      }
    }
    return CUBEEC_OK;
  };
  if ((rc = copy_runs(true))) return rc;
  rc = dev_reconstruct_core(b.h, &c, c.device, l.d_buf, b.S, b.P, dstripe, (size_t)b.claimed, b.present.data(), b.data_only ? 1 : 0,
                            (void*)l.stream, nullptr);
  if (rc) return rc;
  if ((rc = copy_runs(false))) return rc;
  CU(cudaStreamSynchronize(l.stream));
  return CUBEEC_OK;
}

int co_run_batch(CoBatch& b, DevCtx& c, Lane& l) {
  if (b.op == 1) return co_run_reconstruct(b, c, l);
  const size_t dstripe = b.P * b.n;
  const int k = b.h->k, m = b.h->m;
  const size_t part_bytes = b.want_crc ? round_up(crc_part_bytes(c, b.S, (size_t)b.claimed, b.n), 256) : 0;
  int rc = lane_reserve(l, dstripe * b.claimed, part_bytes + (size_t)b.claimed * b.n * 4 + 256);
  if (rc) return rc;
  CU(cudaSetDevice(c.device));
  // data columns of every slot in one 2-D copy (row = one slot: k*P of its n*P bytes)
  CU(cudaMemcpy2DAsync(l.d_buf, dstripe, b.h_buf, dstripe, (size_t)k * b.P, (size_t)b.claimed, cudaMemcpyHostToDevice, l.stream));
  uint32_t* d_part = b.want_crc ? reinterpret_cast<uint32_t*>(l.d_aux) : nullptr;
  uint32_t* d_crc = b.want_crc ? reinterpret_cast<uint32_t*>(l.d_aux + part_bytes) : nullptr;
  rc = dev_encode_impl(b.h, c, l.stream, l.d_buf, b.S, b.P, dstripe, (size_t)b.claimed, d_crc, b.poly, 0, nullptr, d_part);
  if (rc) return rc;
  if (m > 0)
    CU(cudaMemcpy2DAsync(b.h_buf + (size_t)k * b.P, dstripe, l.d_buf + (size_t)k * b.P, dstripe, (size_t)m * b.P, (size_t)b.claimed,
                         cudaMemcpyDeviceToHost, l.stream));
  if (b.want_crc) CU(cudaMemcpyAsync(b.h_crc, d_crc, (size_t)b.claimed * b.n * 4, cudaMemcpyDeviceToHost, l.stream));
  CU(cudaStreamSynchronize(l.stream));
  return CUBEEC_OK;
}

void co_worker() {
  CoQueue& q = g_co;
  LaneLease lease;
  const int lrc = lease.acquire(q.ctx);
  std::unique_lock<std::mutex> lk(q.mu);
  for (;;) {
    // first batch that can start: all claimed slots filled and (closed or past its deadline)
    CoBatch* b = nullptr;
    auto next_deadline = std::chrono::steady_clock::time_point::max();
    const auto now = std::chrono::steady_clock::now();
    for (size_t i = 0; i < q.pending.size(); i++) {
      CoBatch* x = q.pending[i];
      if (x->state == CoBatch::OPEN && now >= x->deadline) {
        x->state = CoBatch::CLOSED;   // no more claims
        q.open.erase(std::remove(q.open.begin(), q.open.end(), x), q.open.end());
      }
      if (x->state == CoBatch::CLOSED && x->filled == x->claimed) {
        b = x;
        q.pending.erase(q.pending.begin() + (long)i);
        break;
      }
      if (x->state == CoBatch::OPEN) next_deadline = std::min(next_deadline, x->deadline);
    }
    if (!b) {
      if (q.stop) return;
      if (next_deadline == std::chrono::steady_clock::time_point::max()) q.work_cv.wait(lk);
      else q.work_cv.wait_until(lk, next_deadline);
      continue;
    }
    b->state = CoBatch::RUNNING;
    lk.unlock();
    int rc = lrc ? lrc : co_run_batch(*b, *q.ctx, *lease.lane);
    std::string err = rc ? t_last_error : std::string();
    lk.lock();
    b->rc = rc;
    b->err = err;
    b->state = CoBatch::DONE;
    b->done_cv.notify_all();
  }
}

void co_start_locked(CoQueue& q) {
  if (q.started) return;
  q.started = true;
  q.ctx = g.ctx[0].get();
  for (int i = 0; i < kCoWorkers; i++) q.workers.emplace_back(co_worker);
  // the workers are joined at process exit: a detached worker would race with CUDA's own teardown
  std::atexit([] {
    CoQueue& qq = g_co;
    {
      std::lock_guard<std::mutex> lk(qq.mu);
      qq.stop = true;
      qq.work_cv.notify_all();
    }
    for (auto& t : qq.workers)
      if (t.joinable()) t.join();
  });
}

// Claim a slot in the open batch of the key (handle, shard size, operation, flags), creating the batch if needed.
// Returns the batch and the slot index; nullptr + rc on failure.
CoBatch* co_claim(cubeec* h, size_t S, int n, int op, bool want_crc, int crc_poly, bool data_only, int cap, int* idx_out, int* rc_out) {
  const size_t P = round_up(S, kAlign), dstripe = P * n;
  CoQueue& q = g_co;
  CoBatch* b = nullptr;
  std::unique_lock<std::mutex> lk(q.mu);
  co_start_locked(q);
  for (CoBatch* x : q.open)
    if (x->h == h && x->S == S && x->op == op && x->poly == crc_poly && x->want_crc == want_crc && x->data_only == data_only &&
        x->state == CoBatch::OPEN && x->claimed < x->cap) {
      b = x;
      break;
    }
  if (!b) {
    if (!q.pool.empty()) {
      b = q.pool.back();
      q.pool.pop_back();
    } else {
      b = new CoBatch();
    }
    b->h = h;
    b->S = S;
    b->P = P;
    b->n = n;
    b->op = op;
    b->poly = crc_poly;
    b->want_crc = want_crc;
    b->data_only = data_only;
    b->cap = cap;
    b->claimed = b->filled = b->left = 0;
    b->state = CoBatch::OPEN;
    b->rc = CUBEEC_OK;
    if (op == 1) b->present.assign((size_t)cap * n, 1);
    b->deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(std::max(0, g_co_delay_us.load()));
    const size_t need = dstripe * (size_t)cap, need_crc = (size_t)cap * n * 4;
    if (need > b->h_cap || need_crc > b->crc_cap) {
      lk.unlock();   // pinned allocation is slow: not under the queue lock
      if (b->h_buf) cudaFreeHost(b->h_buf);
      if (b->h_crc) cudaFreeHost(b->h_crc);
      b->h_buf = nullptr;
      b->h_crc = nullptr;
      b->h_cap = b->crc_cap = 0;
      cudaSetDevice(q.ctx->device);
      cudaError_t e = cudaMallocHost(&b->h_buf, need);
      if (e == cudaSuccess) e = cudaMallocHost(&b->h_crc, need_crc);
      lk.lock();
      if (e != cudaSuccess) {
        if (b->h_buf) cudaFreeHost(b->h_buf);
        b->h_buf = nullptr;
        q.pool.push_back(b);
        *rc_out = cuda_fail(e, "cudaMallocHost(coalescing batch)");
        return nullptr;
      }
      b->h_cap = need;
      b->crc_cap = need_crc;
    }
    q.open.push_back(b);
    q.pending.push_back(b);
  }
  const int idx = b->claimed++;
  b->left++;
  if (b->claimed == b->cap) {
    b->state = CoBatch::CLOSED;
    q.open.erase(std::remove(q.open.begin(), q.open.end(), b), q.open.end());
  }
  if (idx == 0) q.work_cv.notify_one();   // a worker has to watch the new deadline
  *idx_out = idx;
  return b;
}

// After the slot has been filled: wait for the batch, return its result.
int co_wait(CoBatch* b) {
  CoQueue& q = g_co;
  std::unique_lock<std::mutex> lk(q.mu);
  b->filled++;
  if (b->filled == b->claimed) q.work_cv.notify_all();
  b->done_cv.wait(lk, [&] { return b->state == CoBatch::DONE; });
  if (b->rc) t_last_error = b->err;
  return b->rc;
}
void co_release(CoBatch* b) {
  CoQueue& q = g_co;
  std::lock_guard<std::mutex> lk(q.mu);
  if (--b->left == 0) q.pool.push_back(b);   // last caller out recycles the batch (pinned buffers stay allocated)
}
int co_cap(size_t S, int n) {
  const int max_batch = g_co_max_batch.load();
  if (max_batch <= 1) return 0;
  return (int)std::min<size_t>((size_t)max_batch, std::max<size_t>(1, kCoBatchBytes / (round_up(S, kAlign) * (size_t)n)));
}

// cubeec_encode through the queue.  Returns -1 when the call is not eligible (queue off, m == 0, ...).
int co_encode(cubeec* h, uint8_t* const* shards, size_t S, int n, uint32_t* crc_out, int crc_poly) {
  const int cap = co_cap(S, n);
  if (cap <= 1 || h->m == 0) return -1;   // (stripes this large are a batch by themselves)
  int idx = 0, rc = CUBEEC_OK;
  CoBatch* b = co_claim(h, S, n, 0, crc_out != nullptr, crc_poly, false, cap, &idx, &rc);
  if (!b) return rc;
  // fill the slot from the caller's (pageable) memory, in the caller's thread
  const size_t P = b->P;
  uint8_t* slot = b->h_buf + (size_t)idx * P * n;
  for (int i = 0; i < h->k; i++) std::memcpy(slot + (size_t)i * P, shards[i], S);
  rc = co_wait(b);
  if (rc == CUBEEC_OK) {
    for (int i = h->k; i < n; i++) std::memcpy(shards[i], slot + (size_t)i * P, S);
    if (crc_out) std::memcpy(crc_out, b->h_crc + (size_t)idx * n, (size_t)n * 4);
  }
  co_release(b);
  return rc;
}

// cubeec_reconstruct through the queue (degraded reads: stream_get.go:454-465 calls ReconstructData per blob, from
// many goroutines, and one dead node gives every call the same erasure pattern -> one pattern per batch -> the run-time
// compiled kernel).  Returns -1 when not eligible (queue off, checksums requested).
int co_reconstruct(cubeec* h, uint8_t* const* shards, const uint8_t* present, size_t S, int n, bool data_only, uint8_t* filled) {
  const int cap = co_cap(S, n);
  if (cap <= 1) return -1;
  int idx = 0, rc = CUBEEC_OK;
  CoBatch* b = co_claim(h, S, n, 1, false, 0, data_only, cap, &idx, &rc);
  if (!b) return rc;
  const size_t P = b->P;
  uint8_t* slot = b->h_buf + (size_t)idx * P * n;
  for (int i = 0; i < n; i++) {
    b->present[(size_t)idx * n + i] = present[i] ? 1 : 0;   // this slot's flags: nobody else touches them
    if (present[i]) std::memcpy(slot + (size_t)i * P, shards[i], S);
  }
  rc = co_wait(b);
  if (rc == CUBEEC_OK) {
    for (int i = 0; i < n; i++)
      if (!present[i] && !(data_only && i >= h->k)) {
        std::memcpy(shards[i], slot + (size_t)i * P, S);
        if (filled) filled[i] = 1;
      }
  }
  co_release(b);
  return rc;
}

}  // namespace

extern "C" int cubeec_set_coalescing(int max_batch, int delay_us) {
  if (max_batch < 0 || delay_us < 0) return CUBEEC_ERR_INVALID_ARG;
  g_co_max_batch.store(max_batch);
  g_co_delay_us.store(delay_us);
  return CUBEEC_OK;
}

// ------------------------------------------------------------------------------------------
// single stripe, host scatter pointers
// ------------------------------------------------------------------------------------------
extern "C" int cubeec_encode(cubeec_t* h, uint8_t* const* shards, const size_t* lens, int n, uint32_t* crc_out,
                             int crc_poly) {
  if (!h || !shards || !lens) return CUBEEC_ERR_INVALID_ARG;
  if (n != h->k + h->m) return CUBEEC_ERR_TOO_FEW_SHARDS;
  size_t S = 0;
  int rc = check_shards(lens, n, false, &S);
  if (rc) return rc;
  if ((rc = ensure_init())) return rc;
  rc = co_encode(h, shards, S, n, crc_out, crc_poly);   // coalescing queue (-1: not eligible, direct path below)
  if (rc >= 0) return rc;
  DevCtx* c = g.ctx[0].get();
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  Lane& l = *lease.lane;
  const size_t P = round_up(S, kAlign);
  const size_t part_bytes = round_up(crc_part_bytes(*c, S, 1, n), 256);
  if ((rc = lane_reserve(l, P * n, part_bytes + (size_t)n * 4 + 256))) return rc;
  for (int i = 0; i < h->k; i++) CU(cudaMemcpyAsync(l.d_buf + i * P, shards[i], S, cudaMemcpyHostToDevice, l.stream));
  uint32_t* d_part = reinterpret_cast<uint32_t*>(l.d_aux);
  uint32_t* d_crc = crc_out ? reinterpret_cast<uint32_t*>(l.d_aux + part_bytes) : nullptr;
  rc = dev_encode_impl(h, *c, l.stream, l.d_buf, S, P, P * n, 1, d_crc, crc_poly, 0, nullptr, d_part);
  if (rc) return rc;
  for (int i = h->k; i < n; i++) CU(cudaMemcpyAsync(shards[i], l.d_buf + i * P, S, cudaMemcpyDeviceToHost, l.stream));
  if (crc_out) CU(cudaMemcpyAsync(crc_out, d_crc, (size_t)n * 4, cudaMemcpyDeviceToHost, l.stream));
  CU(cudaStreamSynchronize(l.stream));
  return CUBEEC_OK;
}

extern "C" int cubeec_verify(cubeec_t* h, uint8_t* const* shards, const size_t* lens, int n, int* ok) {
  if (!h || !shards || !lens || !ok) return CUBEEC_ERR_INVALID_ARG;
  *ok = 0;
  if (n != h->k + h->m) return CUBEEC_ERR_TOO_FEW_SHARDS;
  size_t S = 0;
  int rc = check_shards(lens, n, false, &S);
  if (rc) return rc;
  if (h->m == 0) { *ok = 1; return CUBEEC_OK; }
  if ((rc = ensure_init())) return rc;
  DevCtx* c = g.ctx[0].get();
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  Lane& l = *lease.lane;
  const size_t P = round_up(S, kAlign);
  if ((rc = lane_reserve(l, P * n, 4096))) return rc;
  for (int i = 0; i < n; i++) CU(cudaMemcpyAsync(l.d_buf + i * P, shards[i], S, cudaMemcpyHostToDevice, l.stream));
  int32_t* d_flag = reinterpret_cast<int32_t*>(l.d_aux);
  CU(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), l.stream));
  rc = dev_encode_impl(h, *c, l.stream, l.d_buf, S, P, P * n, 1, nullptr, 0, 1, d_flag, nullptr);
  if (rc) return rc;
  int32_t flag = 1;
  CU(cudaMemcpyAsync(&flag, d_flag, sizeof(flag), cudaMemcpyDeviceToHost, l.stream));
  CU(cudaStreamSynchronize(l.stream));
  *ok = flag ? 0 : 1;
  return CUBEEC_OK;
}

namespace {
// Shared by cubeec_reconstruct and the batch variant: issues all work for one stripe on a lane.
// Does not synchronize unless crc_out is given.  Small host->device control data (patterns,
// enable flags) is copied from pageable memory, which the runtime stages before returning, so a
// lane can be reused for the next stripe without a host-side wait.
// crc_async (pinned, n words): batched form of crc_out -- the checksums of the regenerated shards are copied there
// without a host-side wait (entries of shards that were not regenerated are undefined).
int reconstruct_issue(cubeec* h, DevCtx& c, Lane& l, uint8_t* const* shards, const uint8_t* present, size_t S,
                      bool data_only, uint8_t* filled, uint32_t* crc_out, int crc_poly, bool* did_work,
                      int32_t* verify_flag_host, uint32_t* crc_async = nullptr) {
  const int k = h->k, n = h->k + h->m;
  *did_work = false;
  int number_present = 0, data_present = 0;
  for (int i = 0; i < n; i++)
    if (present[i]) { number_present++; if (i < k) data_present++; }
  // RS/reedsolomon.go:1434-1444: nothing to do / too few
  const bool nothing = number_present == n || (data_only && data_present == k);
  if (nothing && !verify_flag_host) return CUBEEC_OK;
  if (!nothing && number_present < k) return CUBEEC_ERR_TOO_FEW_SHARDS;
  std::vector<Pattern> passes;
  if (!nothing) {
    int rc = decode_passes(h, present, data_only, passes);
    if (rc) return rc;
  }
  const size_t P = round_up(S, kAlign);
  const Geometry gm = pick_geometry(c, S, 1, false);
  // aux layout
  const size_t o_pat = 0;
  const size_t o_flag = o_pat + round_up(sizeof(Pattern) * std::max<size_t>(passes.size(), 1), 256);
  const size_t o_crc = o_flag + 256;
  const size_t o_en = o_crc + round_up((size_t)n * 4, 256);
  const size_t o_part = o_en + round_up((size_t)n, 256);
  const size_t aux_total = o_part + round_up((size_t)n * gm.n_seg * 4, 256);
  int rc = lane_reserve(l, P * n, aux_total);
  if (rc) return rc;
  // which shards have to travel: the k survivors used as inputs (all present ones when verifying)
  std::vector<uint8_t> need(n, 0);
  if (verify_flag_host) {
    for (int i = 0; i < n; i++) need[i] = present[i];
  } else if (!passes.empty()) {
    for (int c2 = 0; c2 < passes[0].n_in; c2++) need[passes[0].in_slot[c2]] = 1;
  }
  for (int i = 0; i < n; i++)
    if (need[i]) CU(cudaMemcpyAsync(l.d_buf + i * P, shards[i], S, cudaMemcpyHostToDevice, l.stream));
  Pattern* d_pat = reinterpret_cast<Pattern*>(l.d_aux + o_pat);
  int32_t* d_flag = reinterpret_cast<int32_t*>(l.d_aux + o_flag);
  uint32_t* d_crc = reinterpret_cast<uint32_t*>(l.d_aux + o_crc);
  uint8_t* d_enable = l.d_aux + o_en;
  const bool want_crc = crc_out || crc_async;
  uint32_t* d_part = want_crc ? reinterpret_cast<uint32_t*>(l.d_aux + o_part) : nullptr;
  if (!passes.empty()) {
    CU(cudaMemcpyAsync(d_pat, passes.data(), sizeof(Pattern) * passes.size(), cudaMemcpyHostToDevice, l.stream));
    std::vector<const Pattern*> pp;
    std::vector<int> nin, ncrc;
    for (size_t j = 0; j < passes.size(); j++) {
      pp.push_back(d_pat + j);
      nin.push_back(passes[j].n_in);
      ncrc.push_back(passes[j].n_out);
    }
    const std::vector<char> exact(pp.size(), 1);
    rc = run_passes(c, l.stream, l.d_buf, S, P, P * n, 1, n, pp, nin, ncrc, exact, nullptr, 0, nullptr, d_part, crc_poly, gm);
    if (rc) return rc;
    std::vector<uint8_t> enable(n, 0);
    for (auto& ps : passes)
      for (int r = 0; r < ps.n_out; r++) {
        const int slot = ps.out_slot[r];
        enable[slot] = 1;
        if (filled) filled[slot] = 1;
        CU(cudaMemcpyAsync(shards[slot], l.d_buf + slot * P, S, cudaMemcpyDeviceToHost, l.stream));
      }
    if (want_crc) {
      CU(cudaMemcpyAsync(d_enable, enable.data(), (size_t)n, cudaMemcpyHostToDevice, l.stream));
      rc = finalize_crc(c, l.stream, d_part, 1, n, S, gm, crc_poly, d_enable, d_crc);
      if (rc) return rc;
    }
    if (crc_async) CU(cudaMemcpyAsync(crc_async, d_crc, (size_t)n * 4, cudaMemcpyDeviceToHost, l.stream));
    if (crc_out) {
      std::vector<uint32_t> h_crc(n);
      CU(cudaMemcpyAsync(h_crc.data(), d_crc, (size_t)n * 4, cudaMemcpyDeviceToHost, l.stream));
      CU(cudaStreamSynchronize(l.stream));
      for (int i = 0; i < n; i++)
        if (enable[i]) crc_out[i] = h_crc[i];
    }
    *did_work = true;
  }
  if (verify_flag_host) {
    // the repair loop's Verify right after Reconstruct (worker_slice_recover.go:871).  Verify needs every shard:
    // with data_only a missing parity shard stays missing (reedSolomon.Verify would fail with ErrShardSize on the
    // empty shard, RS/reedsolomon.go:770-776)
    if (data_only)
      for (int i = k; i < n; i++)
        if (!present[i]) {
          cudaStreamSynchronize(l.stream);
          return CUBEEC_ERR_SHARD_SIZE;
        }
    CU(cudaMemsetAsync(d_flag, 0, sizeof(int32_t), l.stream));
    rc = dev_encode_impl(h, c, l.stream, l.d_buf, S, P, P * n, 1, nullptr, 0, 1, d_flag, nullptr);
    if (rc) return rc;
    CU(cudaMemcpyAsync(verify_flag_host, d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, l.stream));
    *did_work = true;
  }
  return CUBEEC_OK;
}
}  // namespace

extern "C" int cubeec_reconstruct(cubeec_t* h, uint8_t* const* shards, const size_t* lens, int n, int data_only,
                                  uint8_t* filled, uint32_t* crc_out, int crc_poly) {
  if (!h || !shards || !lens) return CUBEEC_ERR_INVALID_ARG;
  if (filled && n > 0) std::memset(filled, 0, (size_t)n);
  if (n != h->k + h->m) return CUBEEC_ERR_TOO_FEW_SHARDS;
  size_t S = 0;
  int rc = check_shards(lens, n, true, &S);
  if (rc) return rc;
  std::vector<uint8_t> present(n);
  for (int i = 0; i < n; i++) present[i] = lens[i] != 0;
  if ((rc = ensure_init())) return rc;
  if (!crc_out) {
    // the cases that need no device work, decided as RS/reedsolomon.go:1434-1444 does, then the coalescing queue
    int number_present = 0, data_present = 0;
    for (int i = 0; i < n; i++)
      if (present[i]) { number_present++; if (i < h->k) data_present++; }
    if (number_present == n || (data_only && data_present == h->k)) return CUBEEC_OK;
    if (number_present < h->k) return CUBEEC_ERR_TOO_FEW_SHARDS;
    rc = co_reconstruct(h, shards, present.data(), S, n, data_only != 0, filled);
    if (rc >= 0) return rc;
  }
  DevCtx* c = g.ctx[0].get();
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  bool did = false;
  rc = reconstruct_issue(h, *c, *lease.lane, shards, present.data(), S, data_only != 0, filled, crc_out, crc_poly, &did,
                         nullptr);
  if (rc) {
    cudaStreamSynchronize(lease.lane->stream);   // copies from / to the caller's buffers may already be queued
    return rc;
  }
  if (did) CU(cudaStreamSynchronize(lease.lane->stream));
  return CUBEEC_OK;
}

extern "C" int cubeec_reconstruct_batch_crc(cubeec_t* h, const cubeec_stripe_t* stripes, size_t n_stripes, int data_only,
                                            int* verify_ok, uint32_t* crc_out, int crc_poly);
extern "C" int cubeec_reconstruct_batch(cubeec_t* h, const cubeec_stripe_t* stripes, size_t n_stripes, int data_only,
                                        int* verify_ok) {
  return cubeec_reconstruct_batch_crc(h, stripes, n_stripes, data_only, verify_ok, nullptr, 0);
}

extern "C" int cubeec_reconstruct_batch_crc(cubeec_t* h, const cubeec_stripe_t* stripes, size_t n_stripes, int data_only,
                                            int* verify_ok, uint32_t* crc_out, int crc_poly) {
  if (!h || (!stripes && n_stripes)) return CUBEEC_ERR_INVALID_ARG;
  int rc = ensure_init();
  if (rc) return rc;
  const int n = h->k + h->m;
  // Stripes are dealt round-robin to (device, lane) pairs; stream order keeps each lane's
  // staging buffer safe for reuse, so the host never waits between stripes.
  const size_t n_ctx = g.ctx.size();
  std::vector<std::unique_ptr<BulkSlot>> bulk;   // context order: no cycle between concurrent batch calls
  for (size_t ci = 0; ci < n_ctx; ci++) bulk.push_back(std::make_unique<BulkSlot>(g.ctx[ci].get()));
  std::vector<std::unique_ptr<LaneLease>> leases;
  for (size_t ci = 0; ci < n_ctx; ci++)
    for (int q = 0; q < 2; q++) {
      auto ls = std::make_unique<LaneLease>();
      if ((rc = ls->acquire(g.ctx[ci].get()))) return rc;
      leases.push_back(std::move(ls));
    }
  std::vector<int32_t*> flags(leases.size(), nullptr);
  int32_t* h_flags = nullptr;
  if (verify_ok) {
    CU(cudaMallocHost(&h_flags, sizeof(int32_t) * std::max<size_t>(n_stripes, 1)));
    for (size_t s = 0; s < n_stripes; s++) h_flags[s] = 0;
  }
  uint32_t* h_crc = nullptr;   // pinned [n_stripes][n]: checksums travel back asynchronously
  if (crc_out) CU(cudaMallocHost(&h_crc, sizeof(uint32_t) * std::max<size_t>(n_stripes * (size_t)n, 1)));
  int result = CUBEEC_OK;
  for (size_t s = 0; s < n_stripes && result == CUBEEC_OK; s++) {
    LaneLease& ls = *leases[s % leases.size()];
    cudaSetDevice(ls.c->device);
    const cubeec_stripe_t& sp = stripes[s];
    if (!sp.shards || !sp.present || sp.shard_len == 0) { result = CUBEEC_ERR_INVALID_ARG; break; }
    bool did = false;
    result = reconstruct_issue(h, *ls.c, *ls.lane, sp.shards, sp.present, sp.shard_len, data_only != 0, nullptr, nullptr,
                               crc_poly, &did, verify_ok ? &h_flags[s] : nullptr, crc_out ? h_crc + s * (size_t)n : nullptr);
  }
  for (auto& ls : leases) {
    cudaSetDevice(ls->c->device);
    cudaError_t e = cudaStreamSynchronize(ls->lane->stream);
    if (e != cudaSuccess && result == CUBEEC_OK) result = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (verify_ok) {
    for (size_t s = 0; s < n_stripes; s++) verify_ok[s] = h_flags[s] ? 0 : 1;
    cudaFreeHost(h_flags);
  }
  if (crc_out) {
    // only the shards this call regenerated carry a checksum: missing, and (data_only) not a parity shard
    if (result == CUBEEC_OK)
      for (size_t s = 0; s < n_stripes; s++)
        for (int i = 0; i < n; i++)
          if (!stripes[s].present[i] && !(data_only && i >= h->k)) crc_out[s * (size_t)n + i] = h_crc[s * (size_t)n + i];
    cudaFreeHost(h_crc);
  }
  return result;
}

// ------------------------------------------------------------------------------------------
// batched host-contiguous encode (ec.Buffer layout), pipelined H2D / kernel / D2H over lanes
// and partitioned over the configured devices.
// ------------------------------------------------------------------------------------------
namespace {
bool crc_flat_usable(size_t len) { return g_force_kernel.load() != 10 && len < 0xFFFFE000ull; }

// CRC32 of `units` ranges per buffer ([offset + u * stride, + block) clipped to len): crc_flat_kernel unless forced off
// (cubeec_debug_force_kernel(10): the first-generation crc_range_kernel) or the buffer is within 8 KiB of 4 GiB.
int run_crc_ranges(DevCtx& c, cudaStream_t st, const uint8_t* d_base, size_t pitch, size_t n_buffers, size_t len, size_t block,
                   size_t stride, size_t offset, size_t units, int crc_poly, uint32_t* d_out) {
  const int pi = crc_poly ? 1 : 0;
  const uint64_t total = (uint64_t)n_buffers * units;
  if (crc_flat_usable(len) && total < 0xFFFFFFFFull) {
    const CrcPoly& P = g.poly[pi];
    CrcFlatParams p;
    std::memset(&p, 0, sizeof(p));
    p.base = d_base;
    p.pitch = pitch;
    p.n_buffers = (uint32_t)n_buffers;
    p.len = (uint32_t)len;
    p.block = (uint32_t)block;
    p.units_per_buffer = (uint32_t)units;
    p.stride = (uint32_t)stride;
    p.offset = (uint32_t)offset;
    const size_t step = stride ? stride : block;
    const bool aligned = offset % kBsfUnitBytes == 0 && (units == 1 || step % kBsfUnitBytes == 0);
    p.tiles_per_range = (uint32_t)((std::min(block, len) + kBsfUnitBytes - 1) / kBsfUnitBytes + (aligned ? 0 : 1));
    p.total_tiles = total * p.tiles_per_range;
    p.poly = P.poly;
    p.init_full = P.mul(0xFFFFFFFFu, P.shift_bytes_const((int64_t)block));
    p.out = d_out;
    p.slice_image = c.d_bs_slice[pi];
    p.fold_tables = c.d_bsf_fold[pi];
    p.klane = c.d_bsf_klane[pi];
    for (int i = 0; i < 24; i++) p.x_unit_pow[i] = P.shift_bytes_const((int64_t)kBsfUnitBytes << i);
    for (int i = 0; i < 33; i++) p.x_neg_pow[i] = P.shift_bytes_const(-((int64_t)1 << i));
    CU(cudaMemsetAsync(d_out, 0, total * 4, st));
    CU(launch_crc_flat(p, c.sm_count, st));
    g_launches += 2;
    t_last_kernel = "crc_flat_kernel";
    return CUBEEC_OK;
  }
  CrcRangeParams p;
  std::memset(&p, 0, sizeof(p));
  p.base = d_base;
  p.pitch = pitch;
  p.n_buffers = (uint32_t)n_buffers;
  p.len = (uint32_t)len;
  p.block = (uint32_t)block;
  p.units_per_buffer = (uint32_t)units;
  p.crc = c.d_crc[pi];
  p.out = d_out;
  p.stride = (uint32_t)stride;
  p.offset = (uint32_t)offset;
  CU(launch_crc_ranges(p, (int)std::min<uint64_t>(total, (uint64_t)c.sm_count * 4), st));
  g_launches++;
  t_last_kernel = "crc_range_kernel";
  return CUBEEC_OK;
}

int dev_crc32_impl(DevCtx& c, cudaStream_t st, const uint8_t* d_base, size_t len, size_t pitch, size_t n_buffers,
                   size_t block, int crc_poly, uint32_t* d_whole, uint32_t* d_blocks);

// hl != nullptr: LRC (h = global code, hl = local code, y = layout): stripes have N+M+L shards.
// blockcrc_out: per-block CRCs (crc32block payloads of block_payload bytes) of every shard, computed on the
// device-resident stripes before they are copied back (one more HBM pass, no extra PCIe traffic).
int encode_contig_device(cubeec* h, DevCtx* c, uint8_t* base, size_t S, size_t first, size_t count, size_t stripe_pitch,
                         uint32_t* crc_out, int crc_poly, cubeec* hl = nullptr, const LrcLayout* y = nullptr,
                         uint32_t* blockcrc_out = nullptr, size_t block_payload = 0) {
  const int k = h->k, n = hl ? y->N + y->M + y->L : h->k + h->m, m = n - k;
  const size_t P = round_up(S, kAlign);
  const size_t dstripe = P * n;
  BulkSlot bulk(c);   // up to 4 batched calls per device stage and compute concurrently
  // chunk: ~192 MiB of device staging per lane (measured 32: 37.4, 96: 39.7, 192: 40.2 GiB/s end to end) -- small enough that the fill / drain of the
  // H2D -> kernel -> D2H pipeline is a small part of a batch, large enough to amortise launches
  // (CUBEEC_CONTIG_CHUNK_MB / CUBEEC_CONTIG_LANES: measurement knobs, defaults are the tuned values)
  static const size_t chunk_mb = [] { const char* s = getenv("CUBEEC_CONTIG_CHUNK_MB"); return s ? (size_t)atoi(s) : (size_t)192; }();
  static const int lanes_env = [] { const char* s = getenv("CUBEEC_CONTIG_LANES"); return s ? atoi(s) : 3; }();
  size_t chunk = std::max<size_t>(1, (std::max<size_t>(chunk_mb, 1) << 20) / dstripe);
  chunk = std::min(chunk, count);
  const int n_lanes = std::max(1, std::min(lanes_env, 8));
  // CRC scratch: only two batch sizes occur, `chunk` and the remainder of the last chunk
  size_t part_cap = crc_part_bytes(*c, S, chunk, n);
  if (count % chunk) part_cap = std::max(part_cap, crc_part_bytes(*c, S, count % chunk, n));
  part_cap = round_up(part_cap, 256);
  const size_t crc_cap = round_up(chunk * n * 4, 256);
  const size_t units = blockcrc_out ? (S + block_payload - 1) / block_payload : 0;
  const size_t blk_cap = round_up(chunk * n * units * 4, 256);
  std::vector<std::unique_ptr<LaneLease>> leases;
  for (int q = 0; q < n_lanes; q++) {
    auto ls = std::make_unique<LaneLease>();
    int rc = ls->acquire(c);
    if (rc) return rc;
    rc = lane_reserve(*ls->lane, dstripe * chunk, (crc_out ? part_cap + crc_cap : 256) + blk_cap);
    if (rc) return rc;
    leases.push_back(std::move(ls));
  }
  size_t done = 0, ci = 0;
  while (done < count) {
    const size_t nb = std::min(chunk, count - done);
    Lane& l = *leases[ci % n_lanes]->lane;
    for (size_t s = 0; s < nb; s++) {
      const uint8_t* src = base + (first + done + s) * stripe_pitch;
      CU(cudaMemcpy2DAsync(l.d_buf + s * dstripe, P, src, S, S, (size_t)k, cudaMemcpyHostToDevice, l.stream));
    }
    uint32_t* d_part = crc_out ? reinterpret_cast<uint32_t*>(l.d_aux) : nullptr;
    uint32_t* d_crc = crc_out ? reinterpret_cast<uint32_t*>(l.d_aux + part_cap) : nullptr;
    if (crc_out && crc_part_bytes(*c, S, nb, n) > part_cap) return CUBEEC_ERR_UNSUPPORTED;
    int rc = hl ? dev_lrc_encode_impl(h, hl, *y, *c, l.stream, l.d_buf, S, P, dstripe, nb, d_crc, crc_poly, d_part)
                : dev_encode_impl(h, *c, l.stream, l.d_buf, S, P, dstripe, nb, d_crc, crc_poly, 0, nullptr, d_part);
    if (rc) return rc;
    if (crc_out)
      CU(cudaMemcpyAsync(crc_out + (first + done) * n, d_crc, nb * n * 4, cudaMemcpyDeviceToHost, l.stream));
    if (blockcrc_out) {
      // shard i of stripe s is buffer s*n + i of the lane (pitch P): same order as blockcrc_out
      uint32_t* d_blk = reinterpret_cast<uint32_t*>(l.d_aux + (crc_out ? part_cap + crc_cap : 256));
      rc = dev_crc32_impl(*c, l.stream, l.d_buf, S, P, nb * (size_t)n, block_payload, crc_poly, nullptr, d_blk);
      if (rc) return rc;
      CU(cudaMemcpyAsync(blockcrc_out + (first + done) * n * units, d_blk, nb * n * units * 4, cudaMemcpyDeviceToHost,
                         l.stream));
    }
    for (size_t s = 0; s < nb && m > 0; s++) {
      uint8_t* dst = base + (first + done + s) * stripe_pitch + (size_t)k * S;
      CU(cudaMemcpy2DAsync(dst, S, l.d_buf + s * dstripe + (size_t)k * P, P, S, (size_t)m, cudaMemcpyDeviceToHost,
                           l.stream));
    }
    done += nb;
    ci++;
  }
  for (auto& ls : leases) CU(cudaStreamSynchronize(ls->lane->stream));
  return CUBEEC_OK;
}
}  // namespace

extern "C" int cubeec_encode_contig(cubeec_t* h, uint8_t* base, size_t shard_len, size_t n_stripes, size_t stripe_pitch,
                                    uint32_t* crc_out, uint32_t* blockcrc_out, size_t block_payload, int crc_poly) {
  if (!h || !base || shard_len == 0) return CUBEEC_ERR_INVALID_ARG;
  if (stripe_pitch < shard_len * (size_t)(h->k + h->m)) return CUBEEC_ERR_INVALID_ARG;
  if (blockcrc_out && block_payload == 0) return CUBEEC_ERR_INVALID_ARG;
  int rc = ensure_init();
  if (rc) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  if (h->m == 0 && !crc_out && !blockcrc_out) return CUBEEC_OK;
  if (blockcrc_out && block_payload > 0xFFFFFFF0ull) return CUBEEC_ERR_INVALID_ARG;
  const size_t n_ctx = g.ctx.size();
  std::vector<int> rcs(n_ctx, CUBEEC_OK);
  std::vector<std::string> errs(n_ctx);
  auto work = [&](size_t ci) {
    const size_t first = n_stripes * ci / n_ctx, last = n_stripes * (ci + 1) / n_ctx;
    if (last > first) {
      cudaSetDevice(g.ctx[ci]->device);
      rcs[ci] = encode_contig_device(h, g.ctx[ci].get(), base, shard_len, first, last - first, stripe_pitch, crc_out,
                                     crc_poly, nullptr, nullptr, blockcrc_out, block_payload);
      if (rcs[ci]) errs[ci] = t_last_error;
    }
  };
  if (n_ctx == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (size_t ci = 0; ci < n_ctx; ci++) th.emplace_back(work, ci);
    for (auto& t : th) t.join();
  }
  for (size_t ci = 0; ci < n_ctx; ci++)
    if (rcs[ci]) { t_last_error = errs[ci]; return rcs[ci]; }
  return CUBEEC_OK;
}

extern "C" int cubeec_lrc_encode_contig(cubeec_t* global, cubeec_t* local, int az_count, uint8_t* base, size_t shard_len,
                                        size_t n_stripes, size_t stripe_pitch, uint32_t* crc_out, int crc_poly) {
  LrcLayout y;
  int rc = lrc_layout(global, local, az_count, &y);
  if (rc) return rc;
  if (!base || shard_len == 0) return CUBEEC_ERR_INVALID_ARG;
  if (stripe_pitch < shard_len * (size_t)(y.N + y.M + y.L)) return CUBEEC_ERR_INVALID_ARG;
  if ((rc = ensure_init())) return rc;
  if (n_stripes == 0) return CUBEEC_OK;
  const size_t n_ctx = g.ctx.size();
  std::vector<int> rcs(n_ctx, CUBEEC_OK);
  std::vector<std::string> errs(n_ctx);
  auto work = [&](size_t ci) {
    const size_t first = n_stripes * ci / n_ctx, last = n_stripes * (ci + 1) / n_ctx;
    if (last > first) {
      cudaSetDevice(g.ctx[ci]->device);
      rcs[ci] = encode_contig_device(global, g.ctx[ci].get(), base, shard_len, first, last - first, stripe_pitch, crc_out,
                                     crc_poly, local, &y);
      if (rcs[ci]) errs[ci] = t_last_error;
    }
  };
  if (n_ctx == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (size_t ci = 0; ci < n_ctx; ci++) th.emplace_back(work, ci);
    for (auto& t : th) t.join();
  }
  for (size_t ci = 0; ci < n_ctx; ci++)
    if (rcs[ci]) { t_last_error = errs[ci]; return rcs[ci]; }
  return CUBEEC_OK;
}

// ------------------------------------------------------------------------------------------
// CRC32 surface
// ------------------------------------------------------------------------------------------
namespace {
int dev_crc32_impl(DevCtx& c, cudaStream_t st, const uint8_t* d_base, size_t len, size_t pitch, size_t n_buffers,
                   size_t block, int crc_poly, uint32_t* d_whole, uint32_t* d_blocks) {
  // internal unit: the caller's block; when only the whole CRC is wanted, the buffer itself (crc_flat_kernel splits it
  // into tiles on its own) or 64 KiB slices (crc_range_kernel: one CTA per unit)
  const size_t unit = block ? block : (crc_flat_usable(len) ? len : std::min<size_t>(len, 1u << 16));
  const size_t units = (len + unit - 1) / unit;
  if (!block && units == 1 && d_whole) return run_crc_ranges(c, st, d_base, pitch, n_buffers, len, unit, 0, 0, 1, crc_poly, d_whole);
  uint32_t* d_units = d_blocks;
  AsyncScratch scratch(st);
  if (!d_units || !block) {
    CU(scratch.alloc(n_buffers * units * 4));
    d_units = static_cast<uint32_t*>(scratch.ptr);
  }
  int rc = run_crc_ranges(c, st, d_base, pitch, n_buffers, len, unit, 0, 0, units, crc_poly, d_units);
  if (rc) return rc;
  if (d_whole) {
    CU(launch_crc_combine(d_units, (uint32_t)n_buffers, (uint32_t)units, (uint32_t)len, (uint32_t)unit,
                          g.poly[crc_poly ? 1 : 0].poly, d_whole, st));
    g_launches++;
  }
  return CUBEEC_OK;
}
}  // namespace

extern "C" int cubeec_dev_crc32(int device, const void* d_base, size_t len, size_t pitch, size_t n_buffers,
                                size_t block_payload, int crc_poly, uint32_t* d_whole, uint32_t* d_blocks,
                                void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!d_base || len == 0 || ((uintptr_t)d_base & 15) || (pitch & 15) || len > 0xFFFFFFF0ull)
    return CUBEEC_ERR_INVALID_ARG;
  if (n_buffers == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  LaneLease lease;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    if ((rc = lease.acquire(c))) return rc;
    st = lease.lane->stream;
  }
  rc = dev_crc32_impl(*c, st, (const uint8_t*)d_base, len, pitch, n_buffers, block_payload, crc_poly, d_whole, d_blocks);
  if (rc) return rc;
  if (!stream) CU(cudaStreamSynchronize(st));
  return CUBEEC_OK;
}

extern "C" int cubeec_crc32_blocks(const uint8_t* p, size_t n, size_t block_payload, int crc_poly, uint32_t* per_block,
                                   uint32_t* whole) {
  int rc = ensure_init();
  if (rc) return rc;
  if (n == 0) {
    if (whole) *whole = 0;   // crc32 of the empty string
    return CUBEEC_OK;
  }
  if (!p || n > 0xFFFFFFF0ull) return CUBEEC_ERR_INVALID_ARG;
  DevCtx* c = g.ctx[0].get();
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  Lane& l = *lease.lane;
  const size_t units = block_payload ? (n + block_payload - 1) / block_payload : 0;
  const size_t aux = round_up(units * 4, 256) + 512;
  if ((rc = lane_reserve(l, round_up(n, kAlign), aux))) return rc;
  CU(cudaMemcpyAsync(l.d_buf, p, n, cudaMemcpyHostToDevice, l.stream));
  uint32_t* d_blocks = units ? reinterpret_cast<uint32_t*>(l.d_aux) : nullptr;
  uint32_t* d_whole = reinterpret_cast<uint32_t*>(l.d_aux + round_up(units * 4, 256));
  rc = dev_crc32_impl(*c, l.stream, l.d_buf, n, round_up(n, kAlign), 1, block_payload, crc_poly, whole ? d_whole : nullptr,
                      (per_block && units) ? d_blocks : nullptr);
  if (rc) return rc;
  if (per_block && units) CU(cudaMemcpyAsync(per_block, d_blocks, units * 4, cudaMemcpyDeviceToHost, l.stream));
  if (whole) CU(cudaMemcpyAsync(whole, d_whole, 4, cudaMemcpyDeviceToHost, l.stream));
  CU(cudaStreamSynchronize(l.stream));
  return CUBEEC_OK;
}

// ------------------------------------------------------------------------------------------
// crc32block framing on the device: the blobnode shard image body, the rpc body
// (rpc.WithCrcEncode, BS/common/rpc/client.go:54-65 -> crc32block.NewBodyEncoder, request_body.go:81-130) and the
// verification that datainspect / the block decoder do on every read (BS/blobnode/datainspect.go:246-283,
// BS/common/crc32block/decode.go:84-107).
// ------------------------------------------------------------------------------------------
namespace {
bool crc32block_valid_len(size_t block_len) { return block_len > 0 && block_len % 4096 == 0 && block_len <= (1u << 30); }

int dev_crc_units(DevCtx& c, cudaStream_t st, const uint8_t* d_base, size_t pitch, size_t n_buffers, size_t len, size_t block,
                  size_t stride, size_t offset, size_t units, int crc_poly, uint32_t* d_out) {
  return run_crc_ranges(c, st, d_base, pitch, n_buffers, len, block, stride, offset, units, crc_poly, d_out);
}

int dev_crc32block_encode_impl(DevCtx& c, cudaStream_t st, const uint8_t* d_src, size_t len, size_t src_pitch, size_t n_buffers,
                               size_t block_len, uint8_t* d_dst, size_t dst_pitch, int crc_poly) {
  const size_t payload = block_len - 4, blocks = (len + payload - 1) / payload;
  AsyncScratch scratch(st);
  CU(scratch.alloc(n_buffers * blocks * 4));
  uint32_t* d_crcs = static_cast<uint32_t*>(scratch.ptr);
  int rc = dev_crc_units(c, st, d_src, src_pitch, n_buffers, len, payload, payload, 0, blocks, crc_poly, d_crcs);
  if (rc) return rc;
  Crc32BlockParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.plain = d_src;
  fp.framed = d_dst;
  fp.plain_pitch = src_pitch;
  fp.framed_pitch = dst_pitch;
  fp.plain_len = len;
  fp.n_buffers = (uint32_t)n_buffers;
  fp.n_blocks = (uint32_t)blocks;
  fp.block_len = (uint32_t)block_len;
  fp.mode = 0;
  fp.crcs = d_crcs;
  CU(launch_crc32block_frame(fp, c.sm_count * 8, st));
  g_launches++;
  t_last_kernel = "crc32block_frame_kernel";
  return CUBEEC_OK;
}

int dev_crc32block_decode_impl(DevCtx& c, cudaStream_t st, const uint8_t* d_framed, size_t framed_len, size_t src_pitch,
                               size_t n_buffers, size_t block_len, uint8_t* d_dst, size_t dst_pitch, int64_t* d_first_bad,
                               uint8_t* d_block_ok, int crc_poly) {
  const size_t blocks = (framed_len + block_len - 1) / block_len;
  const size_t plain_len = framed_len - 4 * blocks;
  AsyncScratch scratch(st);
  CU(scratch.alloc(n_buffers * blocks * 4));
  uint32_t* d_crcs = static_cast<uint32_t*>(scratch.ptr);
  int rc = dev_crc_units(c, st, d_framed, src_pitch, n_buffers, framed_len, block_len - 4, block_len, 4, blocks, crc_poly, d_crcs);
  if (rc) return rc;
  Crc32BlockParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.plain = d_dst;
  fp.framed = d_framed;
  fp.plain_pitch = dst_pitch;
  fp.framed_pitch = src_pitch;
  fp.plain_len = plain_len;
  fp.n_buffers = (uint32_t)n_buffers;
  fp.n_blocks = (uint32_t)blocks;
  fp.block_len = (uint32_t)block_len;
  fp.crcs = d_crcs;
  fp.block_ok = d_block_ok;
  fp.first_bad = d_first_bad;
  if (d_first_bad) CU(cudaMemsetAsync(d_first_bad, 0xFF, n_buffers * sizeof(int64_t), st));   // -1 = every block good
  if (d_first_bad || d_block_ok) {
    CU(launch_crc32block_check(fp, st));
    g_launches++;
    t_last_kernel = "crc32block_check_kernel";
  }
  if (d_dst) {
    fp.mode = 1;
    CU(launch_crc32block_frame(fp, c.sm_count * 8, st));
    g_launches++;
  }
  return CUBEEC_OK;
}

int crc32block_args(const void* src, size_t len, size_t pitch, size_t block_len, bool framed) {
  if (!src || len == 0 || ((uintptr_t)src & 15) || (pitch & 15) || pitch < len || len > 0xFFFFFFF0ull) return CUBEEC_ERR_INVALID_ARG;
  if (!crc32block_valid_len(block_len)) return CUBEEC_ERR_INVALID_ARG;   // crc32block.ErrInvalidBlock (util.go:41-43)
  if (framed) {
    const size_t tail = len % block_len;
    if (tail != 0 && tail <= 4) return CUBEEC_ERR_INVALID_ARG;   // a block without payload: the decoder's ErrMismatchedCrc (request_body.go:118-120)
  }
  return CUBEEC_OK;
}
}  // namespace

extern "C" size_t cubeec_crc32block_encode_size(size_t n, size_t block_len) {
  if (!crc32block_valid_len(block_len)) return 0;
  const size_t payload = block_len - 4;
  return n + 4 * ((n + payload - 1) / payload);   // crc32block.EncodeSize (util.go:56-63)
}
extern "C" size_t cubeec_crc32block_decode_size(size_t total, size_t block_len) {
  if (!crc32block_valid_len(block_len)) return 0;
  return total - 4 * ((total + block_len - 1) / block_len);   // crc32block.DecodeSize (util.go:65-71)
}

extern "C" int cubeec_dev_crc32block_encode(int device, const void* d_src, size_t len, size_t src_pitch, size_t n_buffers,
                                            size_t block_len, void* d_dst, size_t dst_pitch, int crc_poly, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = crc32block_args(d_src, len, src_pitch, block_len, false))) return rc;
  if (!d_dst || ((uintptr_t)d_dst & 15) || (dst_pitch & 15) || dst_pitch < cubeec_crc32block_encode_size(len, block_len))
    return CUBEEC_ERR_INVALID_ARG;
  if (n_buffers == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  LaneLease lease;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    if ((rc = lease.acquire(c))) return rc;
    st = lease.lane->stream;
  }
  rc = dev_crc32block_encode_impl(*c, st, (const uint8_t*)d_src, len, src_pitch, n_buffers, block_len, (uint8_t*)d_dst, dst_pitch, crc_poly);
  if (rc) return rc;
  if (!stream) CU(cudaStreamSynchronize(st));
  return CUBEEC_OK;
}

extern "C" int cubeec_dev_crc32block_decode(int device, const void* d_framed, size_t framed_len, size_t src_pitch, size_t n_buffers,
                                            size_t block_len, void* d_dst, size_t dst_pitch, int64_t* d_first_bad,
                                            uint8_t* d_block_ok, int crc_poly, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = crc32block_args(d_framed, framed_len, src_pitch, block_len, true))) return rc;
  if (d_dst && (((uintptr_t)d_dst & 15) || (dst_pitch & 15) || dst_pitch < cubeec_crc32block_decode_size(framed_len, block_len)))
    return CUBEEC_ERR_INVALID_ARG;
  if (n_buffers == 0) return CUBEEC_OK;
  DevCtx* c = ctx_for_device(device);
  if (!c) return CUBEEC_ERR_INVALID_ARG;
  CU(cudaSetDevice(device));
  LaneLease lease;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    if ((rc = lease.acquire(c))) return rc;
    st = lease.lane->stream;
  }
  rc = dev_crc32block_decode_impl(*c, st, (const uint8_t*)d_framed, framed_len, src_pitch, n_buffers, block_len, (uint8_t*)d_dst,
                                  dst_pitch, d_first_bad, d_block_ok, crc_poly);
  if (rc) return rc;
  if (!stream) CU(cudaStreamSynchronize(st));
  return CUBEEC_OK;
}

// Host-pointer forms: one body.  dst needs cubeec_crc32block_encode_size(n) bytes.
extern "C" int cubeec_crc32block_encode(const uint8_t* src, size_t n, size_t block_len, uint8_t* dst, int crc_poly) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!crc32block_valid_len(block_len)) return CUBEEC_ERR_INVALID_ARG;
  if (n == 0) return CUBEEC_OK;   // an empty body frames to an empty body
  if (!src || !dst || n > 0xFFFFFFF0ull) return CUBEEC_ERR_INVALID_ARG;
  DevCtx* c = g.ctx[0].get();
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  Lane& l = *lease.lane;
  const size_t enc = cubeec_crc32block_encode_size(n, block_len);
  const size_t o_dst = round_up(n, kAlign);
  if ((rc = lane_reserve(l, o_dst + round_up(enc, kAlign), 256))) return rc;
  CU(cudaMemcpyAsync(l.d_buf, src, n, cudaMemcpyHostToDevice, l.stream));
  rc = dev_crc32block_encode_impl(*c, l.stream, l.d_buf, n, o_dst, 1, block_len, l.d_buf + o_dst, round_up(enc, kAlign), crc_poly);
  if (rc) return rc;
  CU(cudaMemcpyAsync(dst, l.d_buf + o_dst, enc, cudaMemcpyDeviceToHost, l.stream));
  CU(cudaStreamSynchronize(l.stream));
  return CUBEEC_OK;
}

// first_bad: -1 when every block checksum matches, else the index of the first bad block (dst is then still filled:
// the caller decides, as blockReader.Read stops with ErrMismatchedCrc at that block).  dst may be NULL (verify only).
extern "C" int cubeec_crc32block_decode(const uint8_t* framed, size_t n_framed, size_t block_len, uint8_t* dst, int64_t* first_bad,
                                        int crc_poly) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!crc32block_valid_len(block_len) || !first_bad) return CUBEEC_ERR_INVALID_ARG;
  *first_bad = -1;
  if (n_framed == 0) return CUBEEC_OK;
  if (!framed || n_framed > 0xFFFFFFF0ull) return CUBEEC_ERR_INVALID_ARG;
  const size_t tail = n_framed % block_len;
  if (tail != 0 && tail <= 4) {   // last block has no payload: ErrMismatchedCrc in the reference (request_body.go:118-120)
    *first_bad = (int64_t)(n_framed / block_len);
    return CUBEEC_OK;
  }
  DevCtx* c = g.ctx[0].get();
  LaneLease lease;
  if ((rc = lease.acquire(c))) return rc;
  Lane& l = *lease.lane;
  const size_t dec = cubeec_crc32block_decode_size(n_framed, block_len);
  const size_t o_dst = round_up(n_framed, kAlign);
  if ((rc = lane_reserve(l, o_dst + round_up(dec, kAlign) + kAlign, 256))) return rc;
  CU(cudaMemcpyAsync(l.d_buf, framed, n_framed, cudaMemcpyHostToDevice, l.stream));
  int64_t* d_bad = reinterpret_cast<int64_t*>(l.d_aux);
  rc = dev_crc32block_decode_impl(*c, l.stream, l.d_buf, n_framed, o_dst, 1, block_len, dst ? l.d_buf + o_dst : nullptr,
                                  round_up(dec, kAlign) + kAlign, d_bad, nullptr, crc_poly);
  if (rc) return rc;
  if (dst) CU(cudaMemcpyAsync(dst, l.d_buf + o_dst, dec, cudaMemcpyDeviceToHost, l.stream));
  CU(cudaMemcpyAsync(first_bad, d_bad, sizeof(int64_t), cudaMemcpyDeviceToHost, l.stream));
  CU(cudaStreamSynchronize(l.stream));
  return CUBEEC_OK;
}

extern "C" int cubeec_crc32(const uint8_t* p, size_t n, int crc_poly, uint32_t* out) {
  if (!out) return CUBEEC_ERR_INVALID_ARG;
  return cubeec_crc32_blocks(p, n, 0, crc_poly, nullptr, out);
}
