// cubefs_ec.hpp -- host-side mirror (C++) of the BlobStore packages that sit directly above the
// libcubeec C-ABI.  The reference is Go and no Go toolchain exists in the build image, so the layer
// the cgo shim would leave untouched is restated here with the reference's names, argument meaning
// and error behaviour, calling the engine only through include/cubeec.h:
//
//   cubefs::codemode   <- blobstore/common/codemode/codemode.go   (tactics, LRC layout)
//   cubefs::ec         <- blobstore/common/ec/{encoder,lrcencoder,buf}.go
//   cubefs::crc32block <- blobstore/common/crc32block/{block,util,sized_coder_block}.go
//
// Go's []byte is modelled by Slice{ptr,len,cap} so that the len==0 / cap>=size rules of
// reedsolomon.Reconstruct (RS/reedsolomon.go:1514-1518) and Split (:1574-1632) carry over.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#include "../../include/cubeec.h"

namespace cubefs {

struct Slice {
  uint8_t* ptr = nullptr;
  size_t len = 0;
  size_t cap = 0;
  std::shared_ptr<uint8_t> owner;   // set when the mirror had to allocate (Go: make([]byte, n))
  static Slice make(size_t n);
  Slice sub(size_t from, size_t to, size_t cap_to) const;   // s[from:to:cap_to]
};
using Shards = std::vector<Slice>;

// ---- error values ---------------------------------------------------------------------------
// engine errors are the CUBEEC_ERR_* codes (== reedsolomon.Err*); ec-level errors below
enum : int {
  ErrShortData = 100,        // ec.ErrShortData        encoder.go:34
  ErrInvalidCodeMode = 101,  // ec.ErrInvalidCodeMode  encoder.go:35
  ErrVerify = 102,           // ec.ErrVerify           encoder.go:36
  ErrInvalidShards = 103,    // ec.ErrInvalidShards    encoder.go:37
  ErrMismatchedCrc = 110,    // crc32block.ErrMismatchedCrc util.go:35
  ErrInvalidBlock = 111      // crc32block.ErrInvalidBlock  util.go:34
};

namespace codemode {

enum CodeMode : uint8_t {   // codemode.go:28-53
  EC15P12 = 1, EC6P6 = 2, EC16P20L2 = 3, EC6P10L2 = 4, EC6P3L3 = 5, EC6P6Align0 = 6, EC6P6Align512 = 7,
  EC4P4L2 = 8, EC12P4 = 9, EC16P4 = 10, EC3P3 = 11, EC10P4 = 12, EC6P3 = 13, EC12P9 = 14, EC24P8 = 15,
  Replica3 = 100, Replica3OneAZ = 101, EC6P6L9 = 200, EC6P8L10 = 201, Replica4TwoAZ = 202
};

struct Tactic {   // codemode.go:156-190
  int N = 0, M = 0, L = 0, AZCount = 0, PutQuorum = 0, GetQuorum = 0, MinShardSize = 0;
  bool IsValid() const;                                            // :290-298
  bool IsReplicateMode() const { return M == 0 && L == 0; }        // :374-377
  std::vector<std::vector<int>> GetECLayoutByAZ() const;           // :301-318
  // local stripe of an AZ: indexes, n = data+parity count, m = local parity count  (:357-372)
  bool LocalStripeInAZ(int az, std::vector<int>& idx, int& n, int& m) const;
  bool LocalStripe(int index, std::vector<int>& idx, int& n, int& m) const;   // :337-355
};

bool IsValid(CodeMode m);
Tactic TacticOf(CodeMode m);                  // CodeMode.Tactic(); aborts on an invalid mode like the Go panic
const char* Name(CodeMode m);
int GetShardNum(CodeMode m);                  // N+M+L
std::vector<CodeMode> GetAllCodeModes();
std::vector<CodeMode> GetECCodeModes();       // non-replica modes (:379-388)

}  // namespace codemode

namespace ec {

struct Config {   // encoder.go:65-69
  codemode::Tactic CodeMode;
  bool EnableVerify = false;
  int Concurrency = 0;
};

struct BufferSizes {   // buf.go:47-55
  int ShardSize = 0, DataSize = 0, ECDataSize = 0, ECSize = 0, From = 0, To = 0;
};
int GetBufferSizes(int dataSize, const codemode::Tactic& t, BufferSizes& out);   // buf.go:67-133,143-150

class Encoder {   // encoder.go:41-62
 public:
  virtual ~Encoder() = default;
  virtual int Encode(Shards& shards) = 0;
  virtual int Reconstruct(Shards& shards, const std::vector<int>& badIdx) = 0;
  virtual int ReconstructData(Shards& shards, const std::vector<int>& badIdx) = 0;
  virtual int Split(const Slice& data, Shards& out) = 0;
  virtual Shards GetDataShards(const Shards& s) const = 0;
  virtual Shards GetParityShards(const Shards& s) const = 0;
  virtual Shards GetLocalShards(const Shards& s) const = 0;
  virtual Shards GetShardsInIdc(const Shards& s, int idx) const = 0;
  virtual int Join(std::vector<uint8_t>& dst, const Shards& shards, int outSize) = 0;
  virtual int Verify(Shards& shards, bool& ok) = 0;
};

int NewEncoder(const Config& cfg, std::unique_ptr<Encoder>& out);   // encoder.go:78-112

}  // namespace ec

namespace crc32block {
constexpr int64_t kDefaultBlock = 64 * 1024;
int64_t BlockPayload(int64_t blockLen);                         // util.go:44-46
int64_t EncodeSize(int64_t size, int64_t blockLen);             // util.go:50-57, -1 on an invalid block
int64_t DecodeSize(int64_t total, int64_t blockLen);            // util.go:59-65
// [crc32-IEEE LE][payload] framing; block CRCs come from the GPU (cubeec_crc32_blocks).
int Encode(const uint8_t* src, int64_t n, int64_t blockLen, std::vector<uint8_t>& dst, uint32_t* whole_crc);
// blockUnit.check for every block; ErrMismatchedCrc on the first bad one.
int Decode(const uint8_t* src, int64_t total, int64_t blockLen, std::vector<uint8_t>& dst);
}  // namespace crc32block

// blobstore/blobnode/core: the on-disk shard image (core/shard.go:74-112,241-297; storage/datafile.go:304-445)
namespace blobnode {
constexpr int kHeaderSize = 32, kFooterSize = 8;
enum : int { ErrShardHeaderMagic = 120, ErrShardHeaderCrc = 121, ErrShardFooterMagic = 122, ErrShardCrc = 123, ErrShardSize = 124 };
int64_t Alignphysize(int64_t shardSize);                       // core/shard.go:419-422
int64_t AlignSize(int64_t p, int64_t bound);                   // page alignment of chunk offsets
struct ShardMeta { uint64_t Bid = 0, Vuid = 0; uint32_t Size = 0, Crc = 0; };
// datafile.Write: header | crc32block-framed body | footer.  Both CRC passes of the reference (whole
// shard + per block) are ONE GPU pass here (cubeec_crc32_blocks).  meta.Crc is filled in.
int WriteShard(ShardMeta& meta, const uint8_t* data, std::vector<uint8_t>& image);
// datafile.Read / data inspect: parse header, verify every block CRC and the footer CRC, return data.
int ReadShard(const uint8_t* image, int64_t n, ShardMeta& meta, std::vector<uint8_t>& data);
}  // namespace blobnode

}  // namespace cubefs
