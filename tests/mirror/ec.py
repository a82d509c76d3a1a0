"""TEST INFRASTRUCTURE, not product: Python view of the C++ host mirror (tests/mirror/cubefs_ec.{hpp,cc}): the BlobStore
packages directly above the libcubeec C-ABI -- blobstore/common/{codemode,ec,crc32block} -- with the
reference's names and error behaviour.  Used by tests/test_host_mirror.py.

Go's []byte is a GoSlice(buf, len): `buf` is the backing numpy array (its size is the capacity)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from cubefs_b200.engine import load as _load_engine

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libcubefs_host.so")

# ec-level errors (encoder.go:33-38) next to the engine's reedsolomon.Err* codes
ERR = {0: "ok", 1: "ErrInvShardNum", 2: "ErrMaxShardNum", 3: "ErrTooFewShards", 4: "ErrShardNoData", 5: "ErrShardSize",
       6: "reedsolomon.ErrShortData", 7: "ErrReconstructRequired", 8: "errSingular", 9: "invalid argument",
       10: "no CUDA device", 11: "CUDA error", 12: "unsupported",
       100: "ErrShortData", 101: "ErrInvalidCodeMode", 102: "ErrVerify", 103: "ErrInvalidShards",
       110: "ErrMismatchedCrc", 111: "ErrInvalidBlock", 120: "ErrShardHeaderMagic", 121: "ErrShardHeaderCrc",
       122: "ErrShardFooterMagic", 123: "ErrShardCrc", 124: "ErrShardSize"}


class EcError(RuntimeError):
    def __init__(self, code: int):
        self.code = abs(code)
        self.name = ERR.get(self.code, str(self.code))
        super().__init__(f"{self.name} ({self.code})")


class _CSlice(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


_lib = None


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        _load_engine()   # libcubeec.so first (fails loudly if missing)
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: run __graft_entry__.build()")
        L = C.CDLL(_SO)
        L.cubefs_codemode_name.restype = C.c_char_p
        L.cubefs_new_encoder.restype = C.c_void_p
        L.cubefs_new_encoder.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.cubefs_free_encoder.argtypes = [C.c_void_p]
        L.cubefs_encode.argtypes = [C.c_void_p, C.POINTER(_CSlice), C.c_int]
        L.cubefs_verify.argtypes = [C.c_void_p, C.POINTER(_CSlice), C.c_int, C.POINTER(C.c_int)]
        L.cubefs_reconstruct.argtypes = [C.c_void_p, C.POINTER(_CSlice), C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.cubefs_split.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(_CSlice), C.c_int]
        L.cubefs_join.argtypes = [C.c_void_p, C.POINTER(_CSlice), C.c_int, C.c_int, C.c_void_p]
        L.cubefs_select.argtypes = [C.c_void_p, C.POINTER(_CSlice), C.c_int, C.c_int, C.c_int, C.POINTER(_CSlice), C.c_int]
        for fn in ("cubefs_crc32block_encode_size", "cubefs_crc32block_decode_size"):
            getattr(L, fn).restype = C.c_longlong
            getattr(L, fn).argtypes = [C.c_longlong, C.c_longlong]
        L.cubefs_crc32block_encode.restype = C.c_longlong
        L.cubefs_crc32block_encode.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.POINTER(C.c_uint32)]
        L.cubefs_crc32block_decode.restype = C.c_longlong
        L.cubefs_crc32block_decode.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p]
        L.cubefs_shard_physize.restype = C.c_longlong
        L.cubefs_shard_physize.argtypes = [C.c_longlong]
        L.cubefs_shard_write.restype = C.c_longlong
        L.cubefs_shard_write.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
        L.cubefs_shard_read.restype = C.c_longlong
        L.cubefs_shard_read.argtypes = [C.c_void_p, C.c_longlong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                        C.POINTER(C.c_uint32), C.c_void_p]
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# codemode
# ---------------------------------------------------------------------------------------------
class Tactic:
    FIELDS = ("N", "M", "L", "AZCount", "PutQuorum", "GetQuorum", "MinShardSize")

    def __init__(self, N=0, M=0, L=0, AZCount=0, PutQuorum=0, GetQuorum=0, MinShardSize=0):
        self.N, self.M, self.L, self.AZCount = N, M, L, AZCount
        self.PutQuorum, self.GetQuorum, self.MinShardSize = PutQuorum, GetQuorum, MinShardSize

    def _c(self):
        return (C.c_int * 7)(*[getattr(self, f) for f in self.FIELDS])

    def __eq__(self, o):
        return all(getattr(self, f) == getattr(o, f) for f in self.FIELDS)

    def IsValid(self) -> bool:
        return bool(_load().cubefs_tactic_valid(self._c()))

    def IsReplicateMode(self) -> bool:
        return self.M == 0 and self.L == 0

    def GetECLayoutByAZ(self) -> List[List[int]]:
        total = self.N + self.M + self.L
        out = (C.c_int * max(total, 1))()
        n = _load().cubefs_layout_by_az(self._c(), out, total)
        per = n // self.AZCount
        return [list(out[i * per:(i + 1) * per]) for i in range(self.AZCount)]

    def GlobalStripe(self):
        return list(range(self.N + self.M)), self.N, self.M

    def _local(self, index, in_az):
        out = (C.c_int * 64)()
        n, m = C.c_int(0), C.c_int(0)
        cnt = _load().cubefs_local_stripe(self._c(), index, in_az, out, 64, C.byref(n), C.byref(m))
        return (list(out[:cnt]) if cnt else None), n.value, m.value

    def LocalStripeInAZ(self, az):
        return self._local(az, 1)

    def LocalStripe(self, index):
        return self._local(index, 0)

    def AllLocalStripe(self):
        if self.L == 0:
            return None, 0, 0
        return self.GetECLayoutByAZ(), (self.N + self.M) // self.AZCount, self.L // self.AZCount


class CodeMode(int):
    def IsValid(self) -> bool:
        out = (C.c_int * 7)()
        return _load().cubefs_tactic(int(self), out) == 0

    def Tactic(self) -> Tactic:
        out = (C.c_int * 7)()
        if _load().cubefs_tactic(int(self), out) != 0:
            raise ValueError(f"Invalid codemode:{int(self)}")   # the Go code panics
        return Tactic(*out)

    T = Tactic

    def Name(self) -> str:
        if not self.IsValid():
            raise ValueError(f"codemode: {int(self)} is invalid")
        return _load().cubefs_codemode_name(int(self)).decode()

    def String(self) -> str:
        return _load().cubefs_codemode_name(int(self)).decode()

    def GetShardNum(self) -> int:
        t = self.Tactic()
        return t.N + t.M + t.L


def GetAllCodeModes() -> List[CodeMode]:
    out = (C.c_int * 64)()
    n = _load().cubefs_all_codemodes(out, 64, 0)
    return [CodeMode(v) for v in out[:n]]


def GetECCodeModes() -> List[CodeMode]:
    out = (C.c_int * 64)()
    n = _load().cubefs_all_codemodes(out, 64, 1)
    return [CodeMode(v) for v in out[:n]]


(EC15P12, EC6P6, EC16P20L2, EC6P10L2, EC6P3L3, EC6P6Align0, EC6P6Align512, EC4P4L2, EC12P4, EC16P4, EC3P3, EC10P4,
 EC6P3, EC12P9, EC24P8) = [CodeMode(i) for i in range(1, 16)]
Replica3, Replica3OneAZ, EC6P6L9, EC6P8L10, Replica4TwoAZ = CodeMode(100), CodeMode(101), CodeMode(200), CodeMode(201), CodeMode(202)


# ---------------------------------------------------------------------------------------------
# ec
# ---------------------------------------------------------------------------------------------
class GoSlice:
    """[]byte: `buf` backs it (capacity = buf.size), `len` bytes are visible."""

    def __init__(self, buf: Optional[np.ndarray], length: Optional[int] = None, _addr: int = 0, _cap: int = 0):
        self.buf = buf
        self.addr = buf.ctypes.data if buf is not None else _addr
        self.cap = buf.size if buf is not None else _cap
        self.len = self.cap if length is None else length

    @property
    def data(self) -> np.ndarray:
        if self.buf is not None:
            return self.buf[:self.len]
        return np.ctypeslib.as_array(C.cast(self.addr, C.POINTER(C.c_uint8)), (max(self.len, 1),))[:self.len]

    def truncate(self):            # s = s[:0]
        self.len = 0
        return self

    def copy(self) -> "GoSlice":
        return GoSlice(self.data.copy())


def _marshal(shards: Sequence[Optional[GoSlice]]):
    arr = (_CSlice * max(len(shards), 1))()
    for i, s in enumerate(shards):
        if s is None:
            arr[i].ptr, arr[i].len, arr[i].cap = None, 0, 0
        else:
            arr[i].ptr, arr[i].len, arr[i].cap = s.addr, s.len, s.cap
    return arr


def _unmarshal(arr, shards: list, keep):
    for i in range(len(shards)):
        s = shards[i]
        addr = arr[i].ptr or 0
        if s is not None and s.addr == addr:
            s.len = arr[i].len
        else:   # the mirror allocated a fresh buffer (cap < shardSize); it is owned by the encoder object
            shards[i] = GoSlice(None, arr[i].len, _addr=addr, _cap=arr[i].cap)
            keep.append(shards[i])


class BufferSizes:
    def __init__(self, ShardSize, DataSize, ECDataSize, ECSize):
        self.ShardSize, self.DataSize, self.ECDataSize, self.ECSize = ShardSize, DataSize, ECDataSize, ECSize


def GetBufferSizes(dataSize: int, tactic: Tactic) -> BufferSizes:
    out = (C.c_int * 4)()
    rc = _load().cubefs_buffer_sizes(dataSize, tactic._c(), out)
    if rc:
        raise EcError(rc)
    return BufferSizes(*out)


class Config:
    def __init__(self, CodeMode: Tactic, EnableVerify: bool = False, Concurrency: int = 0):
        self.CodeMode, self.EnableVerify, self.Concurrency = CodeMode, EnableVerify, Concurrency


class Encoder:
    """ec.Encoder (encoder.go:41-62).  NewEncoder(cfg) picks encoder / lrcEncoder."""

    def __init__(self, cfg: Config):
        err = C.c_int(0)
        self._h = _load().cubefs_new_encoder(cfg.CodeMode._c(), int(cfg.EnableVerify), cfg.Concurrency, C.byref(err))
        if not self._h:
            raise EcError(err.value)
        self.cfg = cfg
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            _load().cubefs_free_encoder(self._h)
            self._h = None

    def Encode(self, shards):
        n = len(shards) if shards is not None else 0
        shards = shards if shards is not None else []
        arr = _marshal(shards)
        rc = _load().cubefs_encode(self._h, arr, n)
        if rc:
            raise EcError(rc)
        _unmarshal(arr, shards, self._keep)

    def Verify(self, shards) -> bool:
        arr = _marshal(shards)
        ok = C.c_int(0)
        rc = _load().cubefs_verify(self._h, arr, len(shards), C.byref(ok))
        if rc:
            raise EcError(rc)
        return bool(ok.value)

    def _rec(self, shards, badIdx, data_only):
        arr = _marshal(shards)
        bad = (C.c_int * max(len(badIdx), 1))(*badIdx)
        rc = _load().cubefs_reconstruct(self._h, arr, len(shards), bad, len(badIdx), data_only)
        if rc:
            raise EcError(rc)
        _unmarshal(arr, shards, self._keep)

    def Reconstruct(self, shards, badIdx):
        self._rec(shards, list(badIdx), 0)

    def ReconstructData(self, shards, badIdx):
        self._rec(shards, list(badIdx), 1)

    def Split(self, data: GoSlice) -> List[GoSlice]:
        out = (_CSlice * 64)()
        n = _load().cubefs_split(self._h, data.addr, data.len, data.cap, out, 64)
        if n < 0:
            raise EcError(n)
        res = []
        for i in range(n):
            a, ln, cp = out[i].ptr or 0, out[i].len, out[i].cap
            if data.buf is not None and data.addr <= a < data.addr + data.cap:
                off = a - data.addr
                res.append(GoSlice(data.buf[off:off + cp], ln))
            else:
                s = GoSlice(None, ln, _addr=a, _cap=cp)
                self._keep.append(s)
                res.append(s)
        return res

    def Join(self, shards, outSize: int) -> bytes:
        arr = _marshal(shards)
        dst = np.zeros(max(outSize, 1), dtype=np.uint8)
        rc = _load().cubefs_join(self._h, arr, len(shards), outSize, dst.ctypes.data)
        if rc:
            raise EcError(rc)
        return dst[:outSize].tobytes()

    def _select(self, shards, which, idx=0):
        arr = _marshal(shards)
        out = (_CSlice * 64)()
        n = _load().cubefs_select(self._h, arr, len(shards), which, idx, out, 64)
        by_addr = {s.addr: s for s in shards if s is not None}
        return [by_addr[out[i].ptr] for i in range(n)]

    def GetDataShards(self, shards): return self._select(shards, 0)
    def GetParityShards(self, shards): return self._select(shards, 1)
    def GetLocalShards(self, shards): return self._select(shards, 2)
    def GetShardsInIdc(self, shards, idx): return self._select(shards, 3, idx)


def NewEncoder(cfg: Config) -> Encoder:
    return Encoder(cfg)


# ---------------------------------------------------------------------------------------------
# crc32block
# ---------------------------------------------------------------------------------------------
def EncodeSize(size: int, blockLen: int = 65536) -> int:
    return _load().cubefs_crc32block_encode_size(size, blockLen)


def DecodeSize(total: int, blockLen: int = 65536) -> int:
    return _load().cubefs_crc32block_decode_size(total, blockLen)


def BlockEncode(src: bytes, blockLen: int = 65536):
    """Framed body ([crc LE][payload]...) and the whole-buffer CRC, block CRCs computed on the GPU."""
    a = np.frombuffer(src, dtype=np.uint8)
    dst = np.zeros(max(EncodeSize(len(a), blockLen), 1), dtype=np.uint8)
    whole = C.c_uint32(0)
    n = _load().cubefs_crc32block_encode(a.ctypes.data if a.size else None, a.size, blockLen, dst.ctypes.data, C.byref(whole))
    if n < 0:
        raise EcError(n)
    return dst[:n].tobytes(), int(whole.value)


def BlockDecode(src: bytes, blockLen: int = 65536) -> bytes:
    a = np.frombuffer(src, dtype=np.uint8)
    dst = np.zeros(max(len(a), 1), dtype=np.uint8)
    n = _load().cubefs_crc32block_decode(a.ctypes.data if a.size else None, a.size, blockLen, dst.ctypes.data)
    if n < 0:
        raise EcError(n)
    return dst[:n].tobytes()


# ---------------------------------------------------------------------------------------------
# blobnode shard image (core/shard.go, core/storage/datafile.go)
# ---------------------------------------------------------------------------------------------
def Alignphysize(shardSize: int) -> int:
    return _load().cubefs_shard_physize(shardSize)


def AlignSize(p: int, bound: int) -> int:
    return (p + bound - 1) & ~(bound - 1)


def WriteShard(bid: int, vuid: int, data: bytes):
    """datafile.Write: (on-disk image, shard.Crc); whole-shard and per-block CRCs from one GPU pass."""
    a = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(Alignphysize(len(a)), dtype=np.uint8)
    crc = C.c_uint32(0)
    n = _load().cubefs_shard_write(bid, vuid, a.ctypes.data if a.size else None, a.size, out.ctypes.data, C.byref(crc))
    if n < 0:
        raise EcError(n)
    return out[:n].tobytes(), int(crc.value)


def ReadShard(image: bytes):
    """datafile.Read / data inspect: verify header, every block CRC and the footer; -> (bid, vuid, crc, data)."""
    a = np.frombuffer(image, dtype=np.uint8)
    out = np.zeros(max(len(a), 1), dtype=np.uint8)
    bid, vuid, crc = C.c_ulonglong(0), C.c_ulonglong(0), C.c_uint32(0)
    n = _load().cubefs_shard_read(a.ctypes.data, a.size, C.byref(bid), C.byref(vuid), C.byref(crc), out.ctypes.data)
    if n < 0:
        raise EcError(n)
    return int(bid.value), int(vuid.value), int(crc.value), out[:n].tobytes()
