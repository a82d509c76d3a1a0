"""ctypes binding of the CPU oracle (oracle/libcubeec_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(cubefs_b200) must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcubeec_oracle.so")

ERR_NAMES = {
    0: "ok", 1: "ErrInvShardNum", 2: "ErrMaxShardNum", 3: "ErrTooFewShards", 4: "ErrShardNoData",
    5: "ErrShardSize", 6: "ErrShortData", 7: "ErrReconstructRequired", 8: "errSingular", 9: "invalid argument",
}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("cubeec_oracle.c", "cubeec_oracle_simd.c", "cubeec_oracle.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, szp, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_size_t), C.c_void_p
        L.oracle_gf_log_table.restype = u8p
        L.oracle_gf_exp_table.restype = u8p
        L.oracle_gf_mul_table.restype = u8p
        L.oracle_gf_mul.restype = C.c_uint8
        L.oracle_gf_mul.argtypes = [C.c_uint8, C.c_uint8]
        L.oracle_gf_div.restype = C.c_uint8
        L.oracle_gf_div.argtypes = [C.c_uint8, C.c_uint8]
        L.oracle_gf_exp.restype = C.c_uint8
        L.oracle_gf_exp.argtypes = [C.c_uint8, C.c_int]
        L.oracle_build_matrix.argtypes = [C.c_int, C.c_int, vp]
        L.oracle_matrix_invert.argtypes = [vp, C.c_int, vp]
        L.oracle_rs_new.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.oracle_rs_free.argtypes = [vp]
        L.oracle_rs_k.argtypes = [vp]
        L.oracle_rs_m.argtypes = [vp]
        L.oracle_rs_matrix.argtypes = [vp]
        L.oracle_rs_matrix.restype = u8p
        L.oracle_rs_encode.argtypes = [vp, vp, szp, C.c_int]
        L.oracle_rs_verify.argtypes = [vp, vp, szp, C.c_int, C.POINTER(C.c_int)]
        L.oracle_rs_reconstruct.argtypes = [vp, vp, szp, C.c_int, C.c_int, vp]
        L.oracle_rs_decode_matrix.argtypes = [vp, vp, C.POINTER(C.c_int), vp]
        L.oracle_rs_encode_batch_simd.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                                  C.c_int, C.c_int, vp]
        L.oracle_rs_reconstruct_batch_simd.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                                       vp, C.c_int]
        L.oracle_rs_encode_batch_simd_rep.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                                      C.c_int, C.c_int, vp, C.c_size_t]
        L.oracle_rs_reconstruct_batch_simd_rep.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                                           vp, C.c_int, C.c_size_t]
        L.oracle_simd_kind.argtypes = [vp]
        L.oracle_simd_kind.restype = C.c_char_p
        L.oracle_split_shard_size.argtypes = [C.c_size_t, C.c_int]
        L.oracle_split_shard_size.restype = C.c_size_t
        L.oracle_ec_buffer_sizes.argtypes = [C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, szp, szp, szp]
        L.oracle_crc32.argtypes = [C.c_int, C.c_uint32, vp, C.c_size_t]
        L.oracle_crc32.restype = C.c_uint32
        L.oracle_crc32_ieee_fast.argtypes = [vp, C.c_size_t]
        L.oracle_crc32_ieee_fast.restype = C.c_uint32
        L.oracle_crc32_combine.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64]
        L.oracle_crc32_combine.restype = C.c_uint32
        for fn in ("oracle_crc32block_encode_size", "oracle_crc32block_decode_size"):
            getattr(L, fn).argtypes = [C.c_int64, C.c_int64]
            getattr(L, fn).restype = C.c_int64
        L.oracle_crc32block_encode.argtypes = [vp, C.c_int64, C.c_int64, vp]
        L.oracle_crc32block_encode.restype = C.c_int64
        L.oracle_crc32block_decode.argtypes = [vp, C.c_int64, C.c_int64, vp]
        L.oracle_crc32block_decode.restype = C.c_int64
        L.oracle_shard_image.argtypes = [C.c_uint64, C.c_uint64, vp, C.c_uint32, vp, C.POINTER(C.c_uint32)]
        L.oracle_shard_image.restype = C.c_int64
        L.oracle_shard_phys_size.argtypes = [C.c_int64]
        L.oracle_shard_phys_size.restype = C.c_int64
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int):
        super().__init__(ERR_NAMES.get(code, str(code)))
        self.code = code


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def gf_tables():
    L = lib()
    log = np.ctypeslib.as_array(L.oracle_gf_log_table(), (256,)).copy()
    exp = np.ctypeslib.as_array(L.oracle_gf_exp_table(), (510,)).copy()
    mul = np.ctypeslib.as_array(L.oracle_gf_mul_table(), (256 * 256,)).copy().reshape(256, 256)
    return log, exp, mul


def build_matrix(k: int, total: int) -> np.ndarray:
    out = np.zeros((total, k), dtype=np.uint8)
    rc = lib().oracle_build_matrix(k, total, _ptr(out))
    if rc:
        raise OracleError(rc)
    return out


def matrix_invert(m: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(m, dtype=np.uint8)
    out = np.zeros_like(m)
    rc = lib().oracle_matrix_invert(_ptr(m), m.shape[0], _ptr(out))
    if rc:
        raise OracleError(rc)
    return out


class RS:
    """reedsolomon.New(k, m) with default options, as CubeFS calls it."""

    def __init__(self, k: int, m: int):
        h = C.c_void_p()
        rc = lib().oracle_rs_new(k, m, C.byref(h))
        if rc:
            raise OracleError(rc)
        self._h = h
        self.k, self.m = k, m

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_rs_free(self._h)
            self._h = None

    @property
    def matrix(self) -> np.ndarray:
        p = lib().oracle_rs_matrix(self._h)
        return np.ctypeslib.as_array(p, ((self.k + self.m) * self.k,)).copy().reshape(self.k + self.m, self.k)

    @property
    def parity_rows(self) -> np.ndarray:
        return self.matrix[self.k:]

    def simd_kind(self) -> str:
        return lib().oracle_simd_kind(self._h).decode()

    @staticmethod
    def _args(shards, size_hint=None):
        """shards: list of np.uint8 arrays or None (missing). Missing shards get a scratch buffer."""
        n = len(shards)
        size = size_hint
        if size is None:
            for s in shards:
                if s is not None and len(s):
                    size = len(s)
                    break
        bufs, lens = [], (C.c_size_t * n)()
        ptrs = (C.c_void_p * n)()
        for i, s in enumerate(shards):
            if s is None or len(s) == 0:
                b = np.zeros(size or 1, dtype=np.uint8)
                lens[i] = 0
            else:
                b = s if (isinstance(s, np.ndarray) and s.dtype == np.uint8 and s.flags.c_contiguous) \
                    else np.ascontiguousarray(s, dtype=np.uint8)
                lens[i] = len(b)
            bufs.append(b)
            ptrs[i] = b.ctypes.data
        return bufs, ptrs, lens

    def encode(self, shards):
        """In place: shards[k:] are overwritten (must be writable np arrays)."""
        bufs, ptrs, lens = self._args(shards)
        rc = lib().oracle_rs_encode(self._h, ptrs, lens, len(shards))
        if rc:
            raise OracleError(rc)
        return bufs

    def verify(self, shards) -> bool:
        bufs, ptrs, lens = self._args(shards)
        ok = C.c_int(0)
        rc = lib().oracle_rs_verify(self._h, ptrs, lens, len(shards), C.byref(ok))
        if rc:
            raise OracleError(rc)
        return bool(ok.value)

    def reconstruct(self, shards, data_only=False):
        """Returns the full shard list; missing entries (None) replaced by regenerated arrays
        (missing parity stays None when data_only)."""
        bufs, ptrs, lens = self._args(shards)
        filled = np.zeros(len(shards), dtype=np.uint8)
        rc = lib().oracle_rs_reconstruct(self._h, ptrs, lens, len(shards), int(data_only), _ptr(filled))
        if rc:
            raise OracleError(rc)
        out = []
        for i, s in enumerate(shards):
            if s is not None and len(s):
                out.append(s)
            else:
                out.append(bufs[i] if filled[i] else None)
        return out

    def decode_matrix(self, present):
        present = np.ascontiguousarray(present, dtype=np.uint8)
        valid = (C.c_int * self.k)()
        rows = np.zeros((self.k, self.k), dtype=np.uint8)
        rc = lib().oracle_rs_decode_matrix(self._h, _ptr(present), valid, _ptr(rows))
        if rc:
            raise OracleError(rc)
        return list(valid), rows

    def encode_batch_simd(self, buf: np.ndarray, shard_len, shard_pitch, stripe_pitch, n_stripes,
                          threads=0, crc_out=None, repeat=1):
        rc = lib().oracle_rs_encode_batch_simd_rep(self._h, _ptr(buf), shard_len, shard_pitch, stripe_pitch,
                                                   n_stripes, threads, int(crc_out is not None),
                                                   _ptr(crc_out) if crc_out is not None else None, repeat)
        if rc:
            raise OracleError(rc)

    def reconstruct_batch_simd(self, buf, shard_len, shard_pitch, stripe_pitch, n_stripes, present, threads=0, repeat=1):
        present = np.ascontiguousarray(present, dtype=np.uint8)
        rc = lib().oracle_rs_reconstruct_batch_simd_rep(self._h, _ptr(buf), shard_len, shard_pitch, stripe_pitch,
                                                        n_stripes, _ptr(present), threads, repeat)
        if rc:
            raise OracleError(rc)


def crc32(data, poly: int = 0, crc: int = 0) -> int:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    return int(lib().oracle_crc32(poly, crc, _ptr(a) if a.size else None, a.size))


def crc32_ieee_fast(data) -> int:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    return int(lib().oracle_crc32_ieee_fast(_ptr(a) if a.size else None, a.size))


def crc32_combine(crc_a: int, crc_b: int, len_b: int, poly: int = 0) -> int:
    return int(lib().oracle_crc32_combine(poly, crc_a, crc_b, len_b))


def ec_buffer_sizes(data_size, n, m, l, min_shard_size):
    a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = lib().oracle_ec_buffer_sizes(data_size, n, m, l, min_shard_size, C.byref(a), C.byref(b), C.byref(c))
    if rc:
        raise OracleError(rc)
    return a.value, b.value, c.value


def crc32block_encode(src: bytes, block_len: int = 65536) -> bytes:
    a = np.frombuffer(src, dtype=np.uint8)
    n = lib().oracle_crc32block_encode_size(len(a), block_len)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    w = lib().oracle_crc32block_encode(_ptr(a) if a.size else None, a.size, block_len, _ptr(out))
    return out[:w].tobytes()


def crc32block_decode(src: bytes, block_len: int = 65536):
    a = np.frombuffer(src, dtype=np.uint8)
    out = np.zeros(max(len(a), 1), dtype=np.uint8)
    w = lib().oracle_crc32block_decode(_ptr(a) if a.size else None, a.size, block_len, _ptr(out))
    return None if w < 0 else out[:w].tobytes()


def shard_image(bid: int, vuid: int, data: bytes):
    """(image bytes, shard crc) as blobnode's datafile.Write lays a shard out on disk."""
    a = np.frombuffer(data, dtype=np.uint8)
    out = np.zeros(int(lib().oracle_shard_phys_size(len(a))), dtype=np.uint8)
    crc = C.c_uint32(0)
    n = lib().oracle_shard_image(bid, vuid, _ptr(a) if a.size else None, a.size, _ptr(out), C.byref(crc))
    assert n == out.size
    return out.tobytes(), int(crc.value)
