// cubefs_ec_capi.cc -- flat C view of the C++ host mirror, for the ctypes-driven parity tests
// (tests/test_host_mirror.py).  Not part of the drop-in ABI (that is include/cubeec.h).
#include <cstring>
#include <memory>
#include <vector>

#include "cubefs_ec.hpp"

using namespace cubefs;

extern "C" {

typedef struct { uint8_t* ptr; size_t len; size_t cap; } cubefs_slice_t;

struct cubefs_encoder {
  std::unique_ptr<ec::Encoder> enc;
  std::vector<std::shared_ptr<uint8_t>> keep;   // buffers the mirror had to allocate
};

static Shards to_shards(const cubefs_slice_t* s, int n) {
  Shards v((size_t)n);
  for (int i = 0; i < n; i++) { v[i].ptr = s[i].ptr; v[i].len = s[i].len; v[i].cap = s[i].cap; }
  return v;
}
static void from_shards(cubefs_encoder* e, const Shards& v, cubefs_slice_t* s, int n) {
  for (int i = 0; i < n && i < (int)v.size(); i++) {
    s[i].ptr = v[i].ptr; s[i].len = v[i].len; s[i].cap = v[i].cap;
    if (v[i].owner) e->keep.push_back(v[i].owner);
  }
}

int cubefs_tactic(int mode, int* out7) {
  auto m = (codemode::CodeMode)mode;
  if (!codemode::IsValid(m)) return -1;
  codemode::Tactic t = codemode::TacticOf(m);
  int v[7] = {t.N, t.M, t.L, t.AZCount, t.PutQuorum, t.GetQuorum, t.MinShardSize};
  std::memcpy(out7, v, sizeof(v));
  return 0;
}
const char* cubefs_codemode_name(int mode) { return codemode::Name((codemode::CodeMode)mode); }
int cubefs_all_codemodes(int* out, int cap, int ec_only) {
  auto v = ec_only ? codemode::GetECCodeModes() : codemode::GetAllCodeModes();
  for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
  return (int)v.size();
}
static codemode::Tactic tactic_of(const int* t7) {
  codemode::Tactic t;
  t.N = t7[0]; t.M = t7[1]; t.L = t7[2]; t.AZCount = t7[3]; t.PutQuorum = t7[4]; t.GetQuorum = t7[5]; t.MinShardSize = t7[6];
  return t;
}
int cubefs_tactic_valid(const int* t7) { return tactic_of(t7).IsValid() ? 1 : 0; }
// out: AZCount rows of (N+M+L)/AZCount indices
int cubefs_layout_by_az(const int* t7, int* out, int cap) {
  auto az = tactic_of(t7).GetECLayoutByAZ();
  int w = 0;
  for (auto& row : az)
    for (int v : row) { if (w < cap) out[w] = v; w++; }
  return w;
}
int cubefs_local_stripe(const int* t7, int index, int in_az, int* out, int cap, int* n, int* m) {
  std::vector<int> idx;
  bool ok = in_az ? tactic_of(t7).LocalStripeInAZ(index, idx, *n, *m) : tactic_of(t7).LocalStripe(index, idx, *n, *m);
  if (!ok) return 0;
  for (size_t i = 0; i < idx.size() && (int)i < cap; i++) out[i] = idx[i];
  return (int)idx.size();
}
int cubefs_buffer_sizes(int data_size, const int* t7, int* out4) {
  ec::BufferSizes b;
  int rc = ec::GetBufferSizes(data_size, tactic_of(t7), b);
  if (rc) return rc;
  out4[0] = b.ShardSize; out4[1] = b.DataSize; out4[2] = b.ECDataSize; out4[3] = b.ECSize;
  return 0;
}

cubefs_encoder* cubefs_new_encoder(const int* t7, int enable_verify, int concurrency, int* err) {
  ec::Config cfg;
  cfg.CodeMode = tactic_of(t7);
  cfg.EnableVerify = enable_verify != 0;
  cfg.Concurrency = concurrency;
  auto e = std::make_unique<cubefs_encoder>();
  *err = ec::NewEncoder(cfg, e->enc);
  if (*err) return nullptr;
  return e.release();
}
void cubefs_free_encoder(cubefs_encoder* e) { delete e; }

int cubefs_encode(cubefs_encoder* e, cubefs_slice_t* s, int n) {
  Shards v = to_shards(s, n);
  int rc = e->enc->Encode(v);
  from_shards(e, v, s, n);
  return rc;
}
int cubefs_verify(cubefs_encoder* e, cubefs_slice_t* s, int n, int* ok) {
  Shards v = to_shards(s, n);
  bool b = false;
  int rc = e->enc->Verify(v, b);
  *ok = b ? 1 : 0;
  return rc;
}
int cubefs_reconstruct(cubefs_encoder* e, cubefs_slice_t* s, int n, const int* bad, int n_bad, int data_only) {
  Shards v = to_shards(s, n);
  std::vector<int> b(bad, bad + n_bad);
  int rc = data_only ? e->enc->ReconstructData(v, b) : e->enc->Reconstruct(v, b);
  from_shards(e, v, s, n);
  return rc;
}
// Split: returns the number of shards written to out (<= cap) or a negative error
int cubefs_split(cubefs_encoder* e, uint8_t* data, size_t len, size_t cap_bytes, cubefs_slice_t* out, int cap) {
  Slice d;
  d.ptr = data; d.len = len; d.cap = cap_bytes;
  Shards v;
  int rc = e->enc->Split(d, v);
  if (rc) return -rc;
  from_shards(e, v, out, cap);
  return (int)v.size();
}
int cubefs_join(cubefs_encoder* e, const cubefs_slice_t* s, int n, int out_size, uint8_t* dst) {
  Shards v = to_shards(s, n);
  std::vector<uint8_t> out;
  int rc = e->enc->Join(out, v, out_size);
  if (rc) return rc;
  std::memcpy(dst, out.data(), out.size());
  return 0;
}
// which: 0 data, 1 parity, 2 local, 3 in-idc(idx); writes indices into the caller's shard list order
int cubefs_select(cubefs_encoder* e, const cubefs_slice_t* s, int n, int which, int idx, cubefs_slice_t* out, int cap) {
  Shards v = to_shards(s, n), r;
  if (which == 0) r = e->enc->GetDataShards(v);
  else if (which == 1) r = e->enc->GetParityShards(v);
  else if (which == 2) r = e->enc->GetLocalShards(v);
  else r = e->enc->GetShardsInIdc(v, idx);
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) { out[i].ptr = r[i].ptr; out[i].len = r[i].len; out[i].cap = r[i].cap; }
  return (int)r.size();
}

long long cubefs_crc32block_encode_size(long long size, long long block) { return crc32block::EncodeSize(size, block); }
long long cubefs_crc32block_decode_size(long long total, long long block) { return crc32block::DecodeSize(total, block); }
long long cubefs_crc32block_encode(const uint8_t* src, long long n, long long block, uint8_t* dst, uint32_t* whole) {
  std::vector<uint8_t> out;
  int rc = crc32block::Encode(src, n, block, out, whole);
  if (rc) return -rc;
  std::memcpy(dst, out.data(), out.size());
  return (long long)out.size();
}
long long cubefs_crc32block_decode(const uint8_t* src, long long total, long long block, uint8_t* dst) {
  std::vector<uint8_t> out;
  int rc = crc32block::Decode(src, total, block, out);
  if (rc) return -rc;
  std::memcpy(dst, out.data(), out.size());
  return (long long)out.size();
}

long long cubefs_shard_physize(long long size) { return blobnode::Alignphysize(size); }
long long cubefs_shard_write(unsigned long long bid, unsigned long long vuid, const uint8_t* data, uint32_t size, uint8_t* out, uint32_t* crc) {
  blobnode::ShardMeta m;
  m.Bid = bid; m.Vuid = vuid; m.Size = size;
  std::vector<uint8_t> img;
  int rc = blobnode::WriteShard(m, data, img);
  if (rc) return -rc;
  std::memcpy(out, img.data(), img.size());
  *crc = m.Crc;
  return (long long)img.size();
}
long long cubefs_shard_read(const uint8_t* image, long long n, unsigned long long* bid, unsigned long long* vuid, uint32_t* crc, uint8_t* out) {
  blobnode::ShardMeta m;
  std::vector<uint8_t> data;
  int rc = blobnode::ReadShard(image, n, m, data);
  if (rc) return -rc;
  *bid = m.Bid; *vuid = m.Vuid; *crc = m.Crc;
  std::memcpy(out, data.data(), data.size());
  return (long long)data.size();
}

}  // extern "C"
