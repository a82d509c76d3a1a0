// jit.cu -- pattern-specialised bit-sliced coding kernels, compiled at run time with NVRTC.
//
// Reconstruct multiplies the survivors by decode rows that depend on the erasure pattern
// (RS/reedsolomon.go:1407-1552), so no XOR network can be generated when the library is built.  But a repair
// task regenerates the SAME shard index in every stripe of a chunk (one broken vuid per task,
// blobstore/blobnode/worker_slice_recover.go:822-871), and a degraded read hits the same missing node for
// every blob of the volume: one pattern, millions of stripes.  For such single-pattern batches the engine
// emits the CUDA source of an encode-like kernel whose networks are the pattern's decode rows (four-Russians
// form, exactly what gen_bitslice.py does at build time for the encode matrices), compiles it for sm_100a
// with NVRTC (dlopen'ed: libcubeec has no link-time dependency on it), loads the cubin through the runtime's
// library API and caches it per (inputs, outputs, coefficients).  The table kernels remain the path for
// batches that mix many patterns, and the fallback when NVRTC is not installed.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "gfmath.h"
#include "kernels.cuh"

namespace cbe {

namespace {

// ---- NVRTC, resolved at first use -------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* so = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  int major = 0, minor = 0;
  bool ok = false;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    // the toolkit's own NVRTC first: a process that imported torch already holds torch's bundled (older) libnvrtc
    // under the same soname, and 256-bit vector loads need the 12.9 ptxas
    const char* env = getenv("CUBEEC_NVRTC");
    for (const char* name : {env ? env : "", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"}) {
      if (!name[0]) continue;
      n.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (n.so) break;
    }
    if (!n.so) return;
#define SYM(field, name) *(void**)(&n.field) = dlsym(n.so, name)
    SYM(CreateProgram, "nvrtcCreateProgram");
    SYM(CompileProgram, "nvrtcCompileProgram");
    SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    SYM(GetCUBIN, "nvrtcGetCUBIN");
    SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    SYM(GetProgramLog, "nvrtcGetProgramLog");
    SYM(DestroyProgram, "nvrtcDestroyProgram");
    SYM(Version, "nvrtcVersion");
#undef SYM
    if (n.Version) n.Version(&n.major, &n.minor);
    n.ok = n.CreateProgram && n.CompileProgram && n.GetCUBINSize && n.GetCUBIN && n.GetProgramLogSize && n.GetProgramLog &&
           n.DestroyProgram;
  });
  return n;
}

// ---- network generation (the run-time twin of gen_bitslice.py: emit_shard_parts) -----------------------------
// masks[i] = set of input planes j feeding output plane i of (c * x)
void row_masks(uint8_t c, int masks[8]) {
  const Gf256& G = gf();
  for (int i = 0; i < 8; i++) {
    int mk = 0;
    for (int j = 0; j < 8; j++)
      if ((G.mul(c, (uint8_t)(1u << j)) >> i) & 1) mk |= 1 << j;
    masks[i] = mk;
  }
}

// statements of: acc[r*8+i] ^= plane i of (coefs[r] * shard), shard given as planes p[0..7]; XOR combinations of
// planes 0..3 / 4..7 are created right before their first use (t[x], t[16+x])
void emit_network(std::ostringstream& o, const std::vector<uint8_t>& coefs) {
  struct Use { int lo, hi, idx; };
  std::vector<Use> uses;
  for (size_t r = 0; r < coefs.size(); r++) {
    int mk[8];
    row_masks(coefs[r], mk);
    for (int i = 0; i < 8; i++)
      if (mk[i]) uses.push_back({mk[i] & 15, mk[i] >> 4, (int)r * 8 + i});
  }
  std::sort(uses.begin(), uses.end(), [](const Use& a, const Use& b) { return a.lo != b.lo ? a.lo < b.lo : (a.hi != b.hi ? a.hi < b.hi : a.idx < b.idx); });
  std::set<int> built;
  auto name = [](int x, int base) -> std::string {
    if ((x & (x - 1)) == 0) {
      int bit = 0;
      while (!((x >> bit) & 1)) bit++;
      return "p[" + std::to_string(base + bit) + "]";
    }
    return "t[" + std::to_string((base ? 16 : 0) + x) + "]";
  };
  std::function<void(int, int)> build = [&](int x, int base) {
    if ((x & (x - 1)) == 0 || built.count(x + (base ? 16 : 0))) return;
    const int parent = x & (x - 1), low = x & -x;
    build(parent, base);
    built.insert(x + (base ? 16 : 0));
    o << "      " << name(x, base) << " = " << name(parent, base) << " ^ " << name(low, base) << ";\n";
  };
  for (const Use& u : uses) {
    std::string terms;
    if (u.lo) {
      build(u.lo, 0);
      terms += name(u.lo, 0);
    }
    if (u.hi) {
      build(u.hi, 4);
      if (!terms.empty()) terms += " ^ ";
      terms += name(u.hi, 4);
    }
    o << "      acc[" << u.idx << "] ^= " << terms << ";\n";
  }
}

const char* kPrologue = R"SRC(
typedef unsigned int u32;
typedef unsigned long long u64;
typedef unsigned char u8;
struct JitParams {
  u8* base;
  u64 stripe_pitch, shard_pitch;
  u32 shard_len, n_stripes;
  u32 units_per_shard, pad;
  u64 total_units;
};
#if CUBEEC_JIT_V8
__device__ __forceinline__ void ldg256(const void* p, u32 (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const u32 (&r)[8]) {
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
#else   // NVRTC older than 12.9: no 256-bit vector accesses in its ptxas, two 128-bit ones instead
__device__ __forceinline__ void ldg256(const void* p, u32 (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const u32 (&r)[8]) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0+16], {%1,%2,%3,%4};" ::"l"(p), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
#endif
template <u32 MASK>
__device__ __forceinline__ u32 bitsel(u32 a, u32 b) {
  u32 d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "n"(MASK));
  return d;
}
template <int S, u32 MASK>
__device__ __forceinline__ void delta_swap(u32& a, u32& b) {
  const u32 na = bitsel<MASK>(a, b * (1u << S));
  const u32 nb = bitsel<MASK>(a >> S, b);
  a = na;
  b = nb;
}
// 8x8 bit transpose across 8 words, independently in each byte lane (an involution): word j = bit j of 32 bytes
__device__ __forceinline__ void bit_transpose8(u32 (&w)[8]) {
  delta_swap<4, 0x0f0f0f0fu>(w[0], w[4]);
  delta_swap<4, 0x0f0f0f0fu>(w[1], w[5]);
  delta_swap<4, 0x0f0f0f0fu>(w[2], w[6]);
  delta_swap<4, 0x0f0f0f0fu>(w[3], w[7]);
  delta_swap<2, 0x33333333u>(w[0], w[2]);
  delta_swap<2, 0x33333333u>(w[1], w[3]);
  delta_swap<2, 0x33333333u>(w[4], w[6]);
  delta_swap<2, 0x33333333u>(w[5], w[7]);
  delta_swap<1, 0x55555555u>(w[0], w[1]);
  delta_swap<1, 0x55555555u>(w[2], w[3]);
  delta_swap<1, 0x55555555u>(w[4], w[5]);
  delta_swap<1, 0x55555555u>(w[6], w[7]);
}
)SRC";

std::string make_source(const std::vector<uint8_t>& in_slots, const std::vector<uint8_t>& out_slots, const std::vector<uint8_t>& rows) {
  const int K = (int)in_slots.size(), M = (int)out_slots.size();
  const int RD = std::min(4, K);   // load ring: RD - 1 shards in flight ahead of the one being coded
  std::ostringstream o;
  const Nvrtc& nv = nvrtc();
  o << "#define CUBEEC_JIT_V8 " << ((nv.major > 12 || (nv.major == 12 && nv.minor >= 9)) ? 1 : 0) << "\n";
  o << kPrologue;
  o << "#define K " << K << "\n#define M " << M << "\n#define RD " << RD << "\n";
  o << "// network of input c: acc[r*8+i] ^= plane i of (rows[r][c] * shard c)\n";
  o << "template <int C> __device__ __forceinline__ void net(const u32 (&p)[8], u32 (&acc)[8 * M], u32 (&t)[32]);\n";
  for (int c = 0; c < K; c++) {
    std::vector<uint8_t> coefs(M);
    for (int r = 0; r < M; r++) coefs[r] = rows[(size_t)r * K + c];
    o << "template <> __device__ __forceinline__ void net<" << c << ">(const u32 (&p)[8], u32 (&acc)[8 * M], u32 (&t)[32]) {\n";
    emit_network(o, coefs);
    o << "}\n";
  }
  o << "__device__ __forceinline__ u32 in_slot(int c) {\n  switch (c) {\n";
  for (int c = 0; c < K; c++) o << "    case " << c << ": return " << (int)in_slots[c] << ";\n";
  o << "  }\n  return 0;\n}\n__device__ __forceinline__ u32 out_slot(int r) {\n  switch (r) {\n";
  for (int r = 0; r < M; r++) o << "    case " << r << ": return " << (int)out_slots[r] << ";\n";
  o << "  }\n  return 0;\n}\n";
  o << R"SRC(
template <int C> struct Steps {
  // code input C from ring[C % RD]; the load of input C + RD - 1 is issued first
  static __device__ __forceinline__ void run(const u8* src, const u64 shard_pitch, const bool live, const bool full, const u32 (&msk)[8],
                                             u32 (&ring)[RD][8], u32 (&acc)[8 * M], u32 (&t)[32]) {
    if constexpr (C + RD - 1 < K) {
      if (live) ldg256(src + (u64)in_slot(C + RD - 1) * shard_pitch, ring[(C + RD - 1) % RD]);
    }
    u32 (&w)[8] = ring[C % RD];
    if (!full) {
#pragma unroll
      for (int i = 0; i < 8; i++) w[i] &= msk[i];
    }
    bit_transpose8(w);
    net<C>(w, acc, t);
    if constexpr (C + 1 < K) Steps<C + 1>::run(src, shard_pitch, live, full, msk, ring, acc, t);
  }
};

// Flat work split: units of 1 KiB of every shard of one stripe (32 lanes x one 32-byte column; a warp's 256-bit
// request covers eight whole 128-byte lines), numbered stripe-major; warp g takes units [g*U/GW, (g+1)*U/GW).
extern "C" __global__ void __launch_bounds__(512, 1) rs_jit_kernel(const JitParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u64 U = p.total_units, GW = (u64)gridDim.x * (blockDim.x >> 5), gw = (u64)blockIdx.x * (blockDim.x >> 5) + warp;
  const u64 u_lo = gw * U / GW, u_hi = (gw + 1) * U / GW;
  if (u_lo >= u_hi) return;
  const u32 wt = p.units_per_shard;
  u32 s = (u32)(u_lo / wt), tt = (u32)(u_lo - (u64)s * wt);
  u32 ring[RD][8];
#pragma unroll
  for (int b = 0; b < RD; b++)
#pragma unroll
    for (int i = 0; i < 8; i++) ring[b][i] = 0;
  for (u64 u = u_lo; u < u_hi; u++) {
    u8* sbase = p.base + (u64)s * p.stripe_pitch;
    const u32 col = tt * 1024u + (u32)lane * 32u;
    const bool live = col < p.shard_len;
    const bool full = (tt + 1) * 1024u <= p.shard_len;   // warp-uniform
    u32 msk[8];
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
    u32 acc[8 * M];
#pragma unroll
    for (int i = 0; i < 8 * M; i++) acc[i] = 0;
    u32 t[32];
    const u8* src = sbase + col;
    if (live) {
#pragma unroll
      for (int j = 0; j < RD - 1; j++) ldg256(src + (u64)in_slot(j) * p.shard_pitch, ring[j]);
    }
    Steps<0>::run(src, p.shard_pitch, live, full, msk, ring, acc, t);
#pragma unroll
    for (int r = 0; r < M; r++) {
      u32 o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = acc[r * 8 + i];
      bit_transpose8(o);
      if (live) stg256(sbase + (u64)out_slot(r) * p.shard_pitch + col, o);
    }
    if (++tt == wt) {
      tt = 0;
      s++;
    }
  }
}
)SRC";
  return o.str();
}

struct Entry {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kern = nullptr;
  bool failed = false;
};
std::mutex g_mu;
std::map<std::string, Entry> g_cache;   // key: device | inputs | outputs | coefficients

}  // namespace

bool jit_available() { return nvrtc().ok; }

// The kernel for (inputs, outputs, rows) on `device`, compiling it on first use (about a second); nullptr when NVRTC
// is missing or the compilation failed (err then says why) -- the caller falls back to the table kernels.
const void* jit_kernel(int device, const std::vector<uint8_t>& in_slots, const std::vector<uint8_t>& out_slots,
                       const std::vector<uint8_t>& rows, std::string* err) {
  Nvrtc& n = nvrtc();
  if (!n.ok) {
    if (err) *err = "libnvrtc.so.12 not found";
    return nullptr;
  }
  if (in_slots.empty() || out_slots.empty() || out_slots.size() > 6 || rows.size() != in_slots.size() * out_slots.size()) return nullptr;
  std::string key;
  key.push_back((char)device);
  key.append((const char*)in_slots.data(), in_slots.size());
  key.push_back((char)0xff);
  key.append((const char*)out_slots.data(), out_slots.size());
  key.push_back((char)0xff);
  key.append((const char*)rows.data(), rows.size());
  std::lock_guard<std::mutex> lk(g_mu);   // one compilation at a time: callers with the same pattern wait for the first
  auto it = g_cache.find(key);
  if (it != g_cache.end()) return it->second.failed ? nullptr : (const void*)it->second.kern;
  Entry e;
  e.failed = true;
  const std::string src = make_source(in_slots, out_slots, rows);
  nvrtcProgram prog = nullptr;
  std::string log;
  if (n.CreateProgram(&prog, src.c_str(), "rs_jit_kernel.cu", 0, nullptr, nullptr) == 0) {
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
    const int rc = n.CompileProgram(prog, 3, opts);
    size_t ls = 0;
    if (n.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
      log.resize(ls);
      n.GetProgramLog(prog, &log[0]);
    }
    size_t cs = 0;
    if (rc == 0 && n.GetCUBINSize(prog, &cs) == 0 && cs > 0) {
      std::vector<char> cubin(cs);
      if (n.GetCUBIN(prog, cubin.data()) == 0) {
        cudaSetDevice(device);
        cudaError_t ce = cudaLibraryLoadData(&e.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
        if (ce == cudaSuccess) ce = cudaLibraryGetKernel(&e.kern, e.lib, "rs_jit_kernel");
        if (ce == cudaSuccess) e.failed = false;
        else log += std::string(" | load: ") + cudaGetErrorString(ce);
      }
    } else if (rc != 0) {
      log += " | nvrtcCompileProgram failed";
    }
    n.DestroyProgram(&prog);
  }
  if (e.failed && err) *err = "rs_jit_kernel: " + log;
  if (g_cache.size() > 256) {   // bounded: a process that cycles through more patterns recompiles
    for (auto& kv : g_cache)
      if (kv.second.lib) cudaLibraryUnload(kv.second.lib);
    g_cache.clear();
  }
  g_cache[key] = e;
  return e.failed ? nullptr : (const void*)e.kern;
}

// Compile-only check (no device needed): generates the source for the pattern and runs NVRTC on it.  Returns 0 on
// success, 1 when NVRTC is not installed, 2 on a compilation error (log in *err).  Used by the CPU test-suite.
int jit_compile_check(const std::vector<uint8_t>& in_slots, const std::vector<uint8_t>& out_slots, const std::vector<uint8_t>& rows,
                      std::string* source, std::string* err) {
  const std::string src = make_source(in_slots, out_slots, rows);
  if (source) *source = src;
  Nvrtc& n = nvrtc();
  if (!n.ok) return 1;
  nvrtcProgram prog = nullptr;
  if (n.CreateProgram(&prog, src.c_str(), "rs_jit_kernel.cu", 0, nullptr, nullptr) != 0) return 2;
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17"};
  const int rc = n.CompileProgram(prog, 2, opts);
  size_t ls = 0;
  if (err && n.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
    err->resize(ls);
    n.GetProgramLog(prog, &(*err)[0]);
  }
  size_t cs = 0;
  const bool ok = rc == 0 && n.GetCUBINSize(prog, &cs) == 0 && cs > 0;
  n.DestroyProgram(&prog);
  return ok ? 0 : 2;
}

cudaError_t jit_launch(const void* kern, const JitParams& p, int grid, cudaStream_t st) {
  JitParams pp = p;
  void* args[] = {&pp};
  return cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(512), args, 0, st);
}

}  // namespace cbe
