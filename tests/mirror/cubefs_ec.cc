// cubefs_ec.cc -- see cubefs_ec.hpp.  Every function cites the reference lines it mirrors
// (BS/ = blobstore/, RS/ = vendor/github.com/klauspost/reedsolomon/).
#include "cubefs_ec.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

namespace cubefs {

Slice Slice::make(size_t n) {
  Slice s;
  s.owner = std::shared_ptr<uint8_t>(static_cast<uint8_t*>(std::calloc(n ? n : 1, 1)), std::free);
  s.ptr = s.owner.get();
  s.len = s.cap = n;
  return s;
}
Slice Slice::sub(size_t from, size_t to, size_t cap_to) const {
  Slice s;
  s.ptr = ptr + from;
  s.len = to - from;
  s.cap = cap_to - from;
  s.owner = owner;
  return s;
}

// =============================================================================================
// codemode
// =============================================================================================
namespace codemode {
namespace {
struct Entry { CodeMode mode; const char* name; Tactic t; };
// constCodeModeTactic, BS/common/codemode/codemode.go:65-94
const Entry kModes[] = {
    {EC15P12, "EC15P12", {15, 12, 0, 3, 24, 0, 2048}},
    {EC6P6, "EC6P6", {6, 6, 0, 3, 11, 0, 2048}},
    {EC12P9, "EC12P9", {12, 9, 0, 3, 20, 0, 2048}},
    {EC16P20L2, "EC16P20L2", {16, 20, 2, 2, 34, 0, 2048}},
    {EC6P10L2, "EC6P10L2", {6, 10, 2, 2, 14, 0, 2048}},
    {EC12P4, "EC12P4", {12, 4, 0, 1, 15, 0, 2048}},
    {EC16P4, "EC16P4", {16, 4, 0, 1, 19, 0, 2048}},
    {EC3P3, "EC3P3", {3, 3, 0, 1, 5, 0, 2048}},
    {EC10P4, "EC10P4", {10, 4, 0, 1, 13, 0, 2048}},
    {EC6P3, "EC6P3", {6, 3, 0, 1, 8, 0, 2048}},
    {EC24P8, "EC24P8", {24, 8, 0, 1, 30, 0, 2048}},
    {EC6P3L3, "EC6P3L3", {6, 3, 3, 3, 9, 0, 2048}},
    {EC6P6Align0, "EC6P6Align0", {6, 6, 0, 3, 11, 0, 0}},
    {EC6P6Align512, "EC6P6Align512", {6, 6, 0, 3, 11, 0, 512}},
    {EC4P4L2, "EC4P4L2", {4, 4, 2, 2, 6, 0, 2048}},
    {EC6P6L9, "EC6P6L9", {6, 6, 9, 3, 11, 0, 2048}},
    {EC6P8L10, "EC6P8L10", {6, 8, 10, 2, 13, 0, 0}},
    {Replica4TwoAZ, "Replica4TwoAZ", {4, 0, 0, 2, 3, 0, 0}},
    {Replica3, "Replica3", {3, 0, 0, 3, 3, 0, 0}},
    {Replica3OneAZ, "Replica3OneAZ", {3, 0, 0, 1, 3, 0, 0}},
};
const Entry* find(CodeMode m) {
  for (const auto& e : kModes)
    if (e.mode == m) return &e;
  return nullptr;
}
}  // namespace

// Tactic.IsValid, codemode.go:290-298
bool Tactic::IsValid() const {
  if (IsReplicateMode()) return N > 0 && AZCount > 0 && N % AZCount == 0 && PutQuorum > 0 && GetQuorum >= 0;
  return N > 0 && M > 0 && L >= 0 && AZCount > 0 && PutQuorum > 0 && GetQuorum >= 0 && MinShardSize >= 0 &&
         N % AZCount == 0 && M % AZCount == 0 && L % AZCount == 0;
}

// GetECLayoutByAZ, codemode.go:301-318
std::vector<std::vector<int>> Tactic::GetECLayoutByAZ() const {
  std::vector<std::vector<int>> az(AZCount);
  const int n = N / AZCount, m = M / AZCount, l = L / AZCount;
  for (int idx = 0; idx < AZCount; idx++) {
    auto& st = az[idx];
    for (int i = 0; i < n; i++) st.push_back(idx * n + i);
    for (int i = 0; i < m; i++) st.push_back(N + idx * m + i);
    for (int i = 0; i < l; i++) st.push_back(N + M + idx * l + i);
  }
  return az;
}

// LocalStripeInAZ, codemode.go:357-372
bool Tactic::LocalStripeInAZ(int azIndex, std::vector<int>& idx, int& n, int& m) const {
  idx.clear();
  n = m = 0;
  if (L == 0) return false;
  auto az = GetECLayoutByAZ();
  if (azIndex < 0 || azIndex >= (int)az.size()) return false;
  idx = az[azIndex];
  n = N / AZCount + M / AZCount;
  m = L / AZCount;
  return true;
}

// LocalStripe, codemode.go:337-355
bool Tactic::LocalStripe(int index, std::vector<int>& idx, int& n, int& m) const {
  idx.clear();
  n = m = 0;
  if (L == 0) return false;
  const int dn = N / AZCount, dm = M / AZCount, dl = L / AZCount;
  int az;
  if (index < N) az = index / dn;
  else if (index < N + M) az = (index - N) / dm;
  else if (index < N + M + L) az = (index - N - M) / dl;
  else return false;
  return LocalStripeInAZ(az, idx, n, m);
}

bool IsValid(CodeMode m) { return find(m) != nullptr; }
Tactic TacticOf(CodeMode m) {
  const Entry* e = find(m);
  if (!e) std::abort();   // Go: panic("Invalid codemode")
  return e->t;
}
const char* Name(CodeMode m) {
  const Entry* e = find(m);
  return e ? e->name : "";
}
int GetShardNum(CodeMode m) {
  Tactic t = TacticOf(m);
  return t.N + t.M + t.L;
}
std::vector<CodeMode> GetAllCodeModes() {
  std::vector<CodeMode> v;
  for (const auto& e : kModes) v.push_back(e.mode);
  return v;
}
std::vector<CodeMode> GetECCodeModes() {
  std::vector<CodeMode> v;
  for (const auto& e : kModes)
    if (!e.t.IsReplicateMode()) v.push_back(e.mode);
  return v;
}
}  // namespace codemode

// =============================================================================================
// ec
// =============================================================================================
namespace ec {
namespace {

// The `engine reedsolomon.Encoder` seam (encoder.go:71-75): the four compute methods go to the GPU
// through the C-ABI, Split/Join are the reference's pure slicing logic.
class Engine {
 public:
  Engine(cubeec_t* h, int k, int m) : h_(h), k_(k), m_(m) {}
  ~Engine() { cubeec_destroy(h_); }
  int total() const { return k_ + m_; }
  cubeec_t* handle() const { return h_; }

  static size_t shard_size(const Shards& s) {   // RS/reedsolomon.go:1332-1339
    for (const auto& x : s)
      if (x.len) return x.len;
    return 0;
  }

  int Encode(Shards& s) {
    std::vector<uint8_t*> p;
    std::vector<size_t> l;
    marshal(s, p, l);
    return cubeec_encode(h_, p.data(), l.data(), (int)s.size(), nullptr, CUBEEC_CRC_IEEE);
  }
  int Verify(Shards& s, bool& ok) {
    std::vector<uint8_t*> p;
    std::vector<size_t> l;
    marshal(s, p, l);
    int okc = 0;
    int rc = cubeec_verify(h_, p.data(), l.data(), (int)s.size(), &okc);
    ok = okc != 0;
    return rc;
  }
  int Reconstruct(Shards& s, bool data_only) {
    const size_t size = shard_size(s);
    // missing shards: reuse capacity or allocate (RS/reedsolomon.go:1514-1518) -- before the ABI
    for (auto& x : s)
      if (x.len == 0 && size && x.cap < size) {
        Slice n = Slice::make(size);
        n.len = 0;
        x = n;
      }
    std::vector<uint8_t*> p;
    std::vector<size_t> l;
    marshal(s, p, l);
    std::vector<uint8_t> filled(s.size(), 0);
    int rc = cubeec_reconstruct(h_, p.data(), l.data(), (int)s.size(), data_only ? 1 : 0, filled.data(), nullptr,
                                CUBEEC_CRC_IEEE);
    if (rc) return rc;
    for (size_t i = 0; i < s.size(); i++)
      if (filled[i]) s[i].len = size;
    return 0;
  }

  // reedSolomon.Split, RS/reedsolomon.go:1574-1632
  int Split(const Slice& data_in, Shards& dst) {
    Slice data = data_in;
    if (data.len == 0) return CUBEEC_ERR_SHORT_DATA;
    const int total = k_ + m_;
    if (total == 1) { dst = {data}; return 0; }
    const size_t dataLen = data.len;
    const size_t perShard = (data.len + k_ - 1) / k_;
    const size_t needTotal = (size_t)total * perShard;
    if (data.cap > data.len) {
      data.len = data.cap > needTotal ? needTotal : data.cap;
      std::memset(data.ptr + dataLen, 0, data.len - dataLen);
    }
    Shards padding;
    if (data.len < needTotal) {
      const size_t fullShards = data.len / perShard;
      for (size_t i = 0; i < total - fullShards; i++) padding.push_back(Slice::make(perShard));
      if (dataLen > perShard * fullShards) {
        size_t from = perShard * fullShards, left = dataLen - from;
        for (auto& pd : padding) {
          if (!left) break;
          size_t n = std::min(left, perShard);
          std::memcpy(pd.ptr, data.ptr + from, n);
          from += n;
          left -= n;
        }
      }
    }
    dst.assign(total, Slice());
    int i = 0;
    size_t off = 0;
    for (; i < total && data.len - off >= perShard; i++) {
      dst[i] = data.sub(off, off + perShard, off + perShard);
      off += perShard;
    }
    for (int j = 0; i + j < total; j++) dst[i + j] = padding[j];
    return 0;
  }

  // reedSolomon.Join, RS/reedsolomon.go:1646-1684
  int Join(std::vector<uint8_t>& out, const Shards& shards_in, int outSize) {
    if ((int)shards_in.size() < k_) return CUBEEC_ERR_TOO_FEW_SHARDS;
    size_t size = 0;
    for (int i = 0; i < k_; i++) {
      if (shards_in[i].ptr == nullptr) return CUBEEC_ERR_RECONSTRUCT_REQUIRED;
      size += shards_in[i].len;
      if ((int64_t)size >= outSize) break;
    }
    if ((int64_t)size < outSize) return CUBEEC_ERR_SHORT_DATA;
    size_t write = (size_t)outSize;
    for (int i = 0; i < k_; i++) {
      const Slice& s = shards_in[i];
      if (write < s.len) { out.insert(out.end(), s.ptr, s.ptr + write); return 0; }
      out.insert(out.end(), s.ptr, s.ptr + s.len);
      write -= s.len;
    }
    return 0;
  }

 private:
  static void marshal(Shards& s, std::vector<uint8_t*>& p, std::vector<size_t>& l) {
    p.resize(s.size());
    l.resize(s.size());
    for (size_t i = 0; i < s.size(); i++) { p[i] = s[i].ptr; l[i] = s[i].len; }
  }
  cubeec_t* h_;
  int k_, m_;
};

int new_engine(int k, int m, std::unique_ptr<Engine>& out) {
  cubeec_t* h = nullptr;
  int rc = cubeec_create(k, m, nullptr, &h);   // reedsolomon.New(N, M), no options (encoder.go:86)
  if (rc) return rc;
  out = std::make_unique<Engine>(h, k, m);
  return 0;
}

// initBadShards, encoder.go:182-188
void initBadShards(Shards& shards, const std::vector<int>& badIdx) {
  for (int i : badIdx)
    if (i >= 0 && i < (int)shards.size() && shards[i].ptr != nullptr && shards[i].len != 0 && shards[i].cap > 0) shards[i].len = 0;
}
// fillFullShards, encoder.go:199-210
void fillFullShards(Shards& shards, size_t begin, size_t end) {
  size_t size = 0;
  for (size_t i = begin; i < end; i++)
    if (shards[i].len) { size = shards[i].len; break; }
  for (size_t i = begin; i < end; i++)
    if (shards[i].len == 0) {
      if (shards[i].cap >= size) shards[i].len = size;
      else shards[i] = Slice::make(size);
    }
}

// ---- encoder, encoder.go:71-180 ----
class PlainEncoder : public Encoder {
 public:
  PlainEncoder(const Config& c, std::unique_ptr<Engine> e) : cfg_(c), engine_(std::move(e)) {}
  int Encode(Shards& shards) override {
    int rc = engine_->Encode(shards);
    if (rc) return rc;
    if (cfg_.EnableVerify) {
      bool ok = false;
      if ((rc = engine_->Verify(shards, ok))) return rc;
      if (!ok) return ErrVerify;
    }
    return 0;
  }
  int Verify(Shards& shards, bool& ok) override { return engine_->Verify(shards, ok); }
  int Reconstruct(Shards& shards, const std::vector<int>& bad) override {
    initBadShards(shards, bad);
    return engine_->Reconstruct(shards, false);
  }
  int ReconstructData(Shards& shards, const std::vector<int>& bad) override {
    initBadShards(shards, bad);
    return engine_->Reconstruct(shards, true);
  }
  int Split(const Slice& data, Shards& out) override { return engine_->Split(data, out); }
  Shards GetDataShards(const Shards& s) const override { return Shards(s.begin(), s.begin() + cfg_.CodeMode.N); }
  Shards GetParityShards(const Shards& s) const override { return Shards(s.begin() + cfg_.CodeMode.N, s.end()); }
  Shards GetLocalShards(const Shards&) const override { return {}; }
  // encoder.go:169-176.  The Go version appends into a sub-slice of the caller's [][]byte and so
  // clobbers later pointer slots of the argument; only the returned list is mirrored.
  Shards GetShardsInIdc(const Shards& s, int idx) const override {
    const int n = cfg_.CodeMode.N, m = cfg_.CodeMode.M, az = cfg_.CodeMode.AZCount;
    const int ln = n / az, lm = m / az;
    Shards r(s.begin() + idx * ln, s.begin() + (idx + 1) * ln);
    r.insert(r.end(), s.begin() + n + lm * idx, s.begin() + n + lm * (idx + 1));
    return r;
  }
  int Join(std::vector<uint8_t>& dst, const Shards& shards, int outSize) override { return engine_->Join(dst, shards, outSize); }

 private:
  Config cfg_;
  std::unique_ptr<Engine> engine_;
};

// ---- lrcEncoder, lrcencoder.go:28-247 ----
class LrcEncoder : public Encoder {
 public:
  LrcEncoder(const Config& c, std::unique_ptr<Engine> g, std::unique_ptr<Engine> l)
      : cfg_(c), engine_(std::move(g)), local_(std::move(l)) {}

  int Encode(Shards& shards) override {   // lrcencoder.go:35-80
    const auto& t = cfg_.CodeMode;
    if ((int)shards.size() != t.N + t.M + t.L) return ErrInvalidShards;
    fillFullShards(shards, 0, shards.size());
    // The PUT path hands over the shards of ONE ec.Buffer (buf.go:23-35): back to back, equal length.
    // Then global + per-AZ local encodes are a single engine call that keeps the stripe in HBM
    // (cubeec_lrc_encode_contig); otherwise compose the calls exactly as lrcencoder.go does.
    if (!cfg_.EnableVerify) {
      bool contig = true;
      const size_t len = shards[0].len;
      for (size_t i = 0; i < shards.size() && contig; i++)
        contig = shards[i].len == len && shards[i].ptr == shards[0].ptr + i * len;
      if (contig && len) {
        int rc = cubeec_lrc_encode_contig(engine_->handle(), local_->handle(), t.AZCount, shards[0].ptr, len, 1,
                                          len * shards.size(), nullptr, CUBEEC_CRC_IEEE);
        if (rc != CUBEEC_ERR_UNSUPPORTED) return rc;
      }
    }
    Shards global(shards.begin(), shards.begin() + t.N + t.M);
    int rc = engine_->Encode(global);
    if (rc) return rc;
    if (cfg_.EnableVerify) {
      bool ok = false;
      if ((rc = engine_->Verify(global, ok))) return rc;
      if (!ok) return ErrVerify;
    }
    copy_back(shards, global, 0);
    for (int i = 0; i < t.AZCount; i++) {
      std::vector<int> idx;
      Shards local = gather(shards, i, idx);
      if ((rc = local_->Encode(local))) return rc;
      if (cfg_.EnableVerify) {
        bool ok = false;
        if ((rc = local_->Verify(local, ok))) return rc;
        if (!ok) return ErrVerify;
      }
      scatter(shards, local, idx);
    }
    return 0;
  }

  int Verify(Shards& shards, bool& ok) override {   // lrcencoder.go:87-128
    const auto& t = cfg_.CodeMode;
    ok = false;
    if ((int)shards.size() == (t.N + t.M + t.L) / t.AZCount) return local_->Verify(shards, ok);
    Shards global(shards.begin(), shards.begin() + std::min<size_t>(shards.size(), t.N + t.M));
    int rc = engine_->Verify(global, ok);
    if (!ok || rc) return rc;
    for (int i = 0; i < t.AZCount; i++) {
      std::vector<int> idx;
      Shards local = gather(shards, i, idx);
      rc = local_->Verify(local, ok);
      if (!ok || rc) return rc;
    }
    ok = true;
    return 0;
  }

  int Reconstruct(Shards& shards, const std::vector<int>& badIdx) override {   // lrcencoder.go:130-185
    const auto& t = cfg_.CodeMode;
    const int n = t.N, m = t.M, l = t.L, az = t.AZCount;
    fillFullShards(shards, 0, shards.size());
    std::vector<int> globalBad;
    for (int i : badIdx)
      if (i < n + m) globalBad.push_back(i);
    initBadShards(shards, globalBad);
    if ((int)shards.size() == (n + m + l) / az) return local_->Reconstruct(shards, false);   // local stripe only
    Shards global(shards.begin(), shards.begin() + n + m);
    int rc = engine_->Reconstruct(global, false);
    if (rc) return rc;
    copy_back(shards, global, 0);
    std::map<int, std::vector<int>> localRestructs;
    for (int i : badIdx)
      if (i >= n + m) {
        const int idc = (i - n - m) * az / l;
        const int localBad = i - n - m - l / az * idc + (n + m) / az;
        localRestructs[idc].push_back(localBad);
      }
    for (auto& kv : localRestructs) {
      std::vector<int> idx;
      Shards local = gather(shards, kv.first, idx);
      initBadShards(local, kv.second);
      if ((rc = local_->Reconstruct(local, false))) return rc;
      scatter(shards, local, idx);
    }
    return 0;
  }

  int ReconstructData(Shards& shards, const std::vector<int>& badIdx) override {   // lrcencoder.go:187-200
    const auto& t = cfg_.CodeMode;
    fillFullShards(shards, 0, t.N + t.M);
    std::vector<int> globalBad;
    for (int i : badIdx)
      if (i < t.N + t.M) globalBad.push_back(i);
    initBadShards(shards, globalBad);
    Shards global(shards.begin(), shards.begin() + t.N + t.M);
    int rc = engine_->Reconstruct(global, true);
    if (rc) return rc;
    copy_back(shards, global, 0);
    return 0;
  }

  int Split(const Slice& data_in, Shards& out) override {   // lrcencoder.go:202-222
    int rc = engine_->Split(data_in, out);
    if (rc) return rc;
    const int L = cfg_.CodeMode.L;
    const size_t shardN = out.size(), shardLen = out[0].len;
    if (data_in.cap >= (L + shardN) * shardLen) {
      for (int i = 0; i < L; i++) {
        Slice s = data_in.sub((shardN + i) * shardLen, (shardN + i + 1) * shardLen, data_in.cap);
        out.push_back(s);
      }
    } else {
      for (int i = 0; i < L; i++) out.push_back(Slice::make(shardLen));
    }
    return 0;
  }
  Shards GetDataShards(const Shards& s) const override { return Shards(s.begin(), s.begin() + cfg_.CodeMode.N); }
  Shards GetParityShards(const Shards& s) const override {
    return Shards(s.begin() + cfg_.CodeMode.N, s.begin() + cfg_.CodeMode.N + cfg_.CodeMode.M);
  }
  Shards GetLocalShards(const Shards& s) const override { return Shards(s.begin() + cfg_.CodeMode.N + cfg_.CodeMode.M, s.end()); }
  Shards GetShardsInIdc(const Shards& s, int idx) const override {   // lrcencoder.go:236-243
    std::vector<int> ids;
    return gather(s, idx, ids);
  }
  int Join(std::vector<uint8_t>& dst, const Shards& shards, int outSize) override {
    Shards g(shards.begin(), shards.begin() + std::min<size_t>(shards.size(), cfg_.CodeMode.N + cfg_.CodeMode.M));
    return engine_->Join(dst, g, outSize);
  }

 private:
  Shards gather(const Shards& shards, int az, std::vector<int>& idx) const {
    int n, m;
    cfg_.CodeMode.LocalStripeInAZ(az, idx, n, m);
    Shards local;
    for (int gi : idx) local.push_back(shards[gi]);
    return local;
  }
  static void scatter(Shards& shards, const Shards& local, const std::vector<int>& idx) {
    for (size_t i = 0; i < idx.size(); i++) shards[idx[i]] = local[i];
  }
  static void copy_back(Shards& shards, const Shards& part, size_t at) {
    for (size_t i = 0; i < part.size(); i++) shards[at + i] = part[i];
  }
  Config cfg_;
  std::unique_ptr<Engine> engine_, local_;
};

}  // namespace

// newBuffer size rules, buf.go:67-84
int GetBufferSizes(int dataSize, const codemode::Tactic& t, BufferSizes& out) {
  if (dataSize <= 0) return ErrShortData;   // isOutOfRange(dataSize, 0, dataSize)
  if (t.N <= 0) return ErrInvalidCodeMode;
  int shardSize = (dataSize + t.N - 1) / t.N;
  if (shardSize < t.MinShardSize) shardSize = t.MinShardSize;
  out.ShardSize = shardSize;
  out.DataSize = dataSize;
  out.ECDataSize = shardSize * t.N;
  out.ECSize = shardSize * (t.N + t.M + t.L);
  out.From = 0;
  out.To = dataSize;
  return 0;
}

// NewEncoder, encoder.go:78-112
int NewEncoder(const Config& cfg_in, std::unique_ptr<Encoder>& out) {
  Config cfg = cfg_in;
  if (!cfg.CodeMode.IsValid()) return ErrInvalidCodeMode;
  if (cfg.Concurrency <= 0) cfg.Concurrency = 100;
  std::unique_ptr<Engine> g;
  int rc = new_engine(cfg.CodeMode.N, cfg.CodeMode.M, g);
  if (rc) return rc;
  if (cfg.CodeMode.L != 0) {
    const int localN = (cfg.CodeMode.N + cfg.CodeMode.M) / cfg.CodeMode.AZCount;
    const int localM = cfg.CodeMode.L / cfg.CodeMode.AZCount;
    std::unique_ptr<Engine> l;
    if ((rc = new_engine(localN, localM, l))) return rc;
    out = std::make_unique<LrcEncoder>(cfg, std::move(g), std::move(l));
  } else {
    out = std::make_unique<PlainEncoder>(cfg, std::move(g));
  }
  return 0;
}

}  // namespace ec

// =============================================================================================
// crc32block
// =============================================================================================
namespace crc32block {
static bool valid_block(int64_t b) { return b > 0 && b % 4096 == 0; }   // util.go:33-35
int64_t BlockPayload(int64_t blockLen) { return blockLen - 4; }
int64_t EncodeSize(int64_t size, int64_t blockLen) {
  if (!valid_block(blockLen)) return -1;
  const int64_t payload = BlockPayload(blockLen);
  return size + 4 * ((size + payload - 1) / payload);
}
int64_t DecodeSize(int64_t total, int64_t blockLen) {
  if (!valid_block(blockLen)) return -1;
  return total - 4 * ((total + blockLen - 1) / blockLen);
}

// encodeBlock loop (sized_coder_block.go:43-67) with the block CRCs computed on the GPU in one pass
// together with the whole-buffer CRC that datafile.Write keeps for the shard footer (datafile.go:337-342).
int Encode(const uint8_t* src, int64_t n, int64_t blockLen, std::vector<uint8_t>& dst, uint32_t* whole_crc) {
  if (!valid_block(blockLen)) return ErrInvalidBlock;
  dst.clear();
  if (n <= 0) { if (whole_crc) *whole_crc = 0; return 0; }
  const int64_t payload = BlockPayload(blockLen);
  const int64_t blocks = (n + payload - 1) / payload;
  std::vector<uint32_t> crcs((size_t)blocks);
  uint32_t whole = 0;
  int rc = cubeec_crc32_blocks(src, (size_t)n, (size_t)payload, CUBEEC_CRC_IEEE, crcs.data(), &whole);
  if (rc) return rc;
  dst.reserve((size_t)(n + 4 * blocks));
  for (int64_t b = 0; b < blocks; b++) {
    const int64_t off = b * payload, take = std::min(payload, n - off);
    const uint32_t c = crcs[(size_t)b];
    const uint8_t le[4] = {(uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16), (uint8_t)(c >> 24)};
    dst.insert(dst.end(), le, le + 4);
    dst.insert(dst.end(), src + off, src + off + take);
  }
  if (whole_crc) *whole_crc = whole;
  return 0;
}

// decodeBlock loop (sized_coder_block.go:72-103, block.go:37-43): strip, then check every block CRC on the GPU.
int Decode(const uint8_t* src, int64_t total, int64_t blockLen, std::vector<uint8_t>& dst) {
  if (!valid_block(blockLen)) return ErrInvalidBlock;
  dst.clear();
  if (total <= 0) return 0;
  const int64_t payload = BlockPayload(blockLen);
  std::vector<uint32_t> want;
  for (int64_t off = 0; off < total; off += blockLen) {
    const int64_t blk = std::min(blockLen, total - off);
    if (blk <= 4) return ErrMismatchedCrc;
    want.push_back((uint32_t)src[off] | ((uint32_t)src[off + 1] << 8) | ((uint32_t)src[off + 2] << 16) | ((uint32_t)src[off + 3] << 24));
    dst.insert(dst.end(), src + off + 4, src + off + blk);
  }
  std::vector<uint32_t> got(want.size());
  int rc = cubeec_crc32_blocks(dst.data(), dst.size(), (size_t)payload, CUBEEC_CRC_IEEE, got.data(), nullptr);
  if (rc) return rc;
  for (size_t i = 0; i < want.size(); i++)
    if (want[i] != got[i]) return ErrMismatchedCrc;
  return 0;
}
}  // namespace crc32block

// =============================================================================================
// blobnode shard image
// =============================================================================================
namespace blobnode {
namespace {
const uint8_t kHeaderMagic[4] = {0xab, 0xcd, 0xef, 0xcc};   // core/shard.go:_shardHeaderMagic
const uint8_t kFooterMagic[4] = {0xcc, 0xef, 0xcd, 0xab};
void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
void put_be64(uint8_t* p, uint64_t v) { put_be32(p, (uint32_t)(v >> 32)); put_be32(p + 4, (uint32_t)v); }
uint32_t get_be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t get_be64(const uint8_t* p) { return ((uint64_t)get_be32(p) << 32) | get_be32(p + 4); }
}  // namespace

int64_t Alignphysize(int64_t shardSize) { return kHeaderSize + crc32block::EncodeSize(shardSize, crc32block::kDefaultBlock) + kFooterSize; }
int64_t AlignSize(int64_t p, int64_t bound) { return (p + bound - 1) & ~(bound - 1); }

// Shard.WriterHeader (core/shard.go:241-261) + datafile.Write (storage/datafile.go:304-408) + WriterFooter (:263-274)
int WriteShard(ShardMeta& meta, const uint8_t* data, std::vector<uint8_t>& image) {
  std::vector<uint8_t> body;
  uint32_t whole = 0;
  int rc = crc32block::Encode(data, meta.Size, crc32block::kDefaultBlock, body, &whole);
  if (rc) return rc;
  meta.Crc = whole;
  image.assign((size_t)Alignphysize(meta.Size), 0);
  uint8_t* h = image.data();
  std::memcpy(h + 4, kHeaderMagic, 4);
  put_be64(h + 8, meta.Bid);
  put_be64(h + 16, meta.Vuid);
  put_be32(h + 24, meta.Size);
  put_be32(h + 28, 0);
  uint32_t hcrc = 0;
  if ((rc = cubeec_crc32(h + 4, 28, CUBEEC_CRC_IEEE, &hcrc))) return rc;   // headerCrc := crc32.ChecksumIEEE(buf[4:])
  put_be32(h, hcrc);
  std::memcpy(h + kHeaderSize, body.data(), body.size());
  uint8_t* f = h + kHeaderSize + body.size();
  std::memcpy(f, kFooterMagic, 4);
  put_be32(f + 4, whole);
  return 0;
}

// Shard.ParseHeader (core/shard.go:276-297) + block decode (blockUnit.check) + footer check
int ReadShard(const uint8_t* image, int64_t n, ShardMeta& meta, std::vector<uint8_t>& data) {
  if (n < kHeaderSize + kFooterSize) return ErrShardSize;
  if (std::memcmp(image + 4, kHeaderMagic, 4) != 0) return ErrShardHeaderMagic;
  uint32_t hcrc = 0;
  int rc = cubeec_crc32(image + 4, 28, CUBEEC_CRC_IEEE, &hcrc);
  if (rc) return rc;
  if (hcrc != get_be32(image)) return ErrShardHeaderCrc;
  meta.Bid = get_be64(image + 8);
  meta.Vuid = get_be64(image + 16);
  meta.Size = get_be32(image + 24);
  if (Alignphysize(meta.Size) != n) return ErrShardSize;
  const int64_t body = n - kHeaderSize - kFooterSize;
  if ((rc = crc32block::Decode(image + kHeaderSize, body, crc32block::kDefaultBlock, data))) return rc;
  const uint8_t* f = image + kHeaderSize + body;
  if (std::memcmp(f, kFooterMagic, 4) != 0) return ErrShardFooterMagic;
  meta.Crc = get_be32(f + 4);
  uint32_t whole = 0;
  if ((rc = cubeec_crc32(data.data(), data.size(), CUBEEC_CRC_IEEE, &whole))) return rc;
  return whole == meta.Crc ? 0 : ErrShardCrc;
}
}  // namespace blobnode

}  // namespace cubefs
