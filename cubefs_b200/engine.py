"""ctypes binding of libcubeec (include/cubeec.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CUBEEC_LIB: load an alternative build of the same ABI (kernel A/B experiments); default = the in-tree library
_SO = os.environ.get("CUBEEC_LIB") or os.path.join(_HERE, "lib", "libcubeec.so")

ERR = {
    0: "ok", 1: "ErrInvShardNum", 2: "ErrMaxShardNum", 3: "ErrTooFewShards", 4: "ErrShardNoData",
    5: "ErrShardSize", 6: "ErrShortData", 7: "ErrReconstructRequired", 8: "errSingular",
    9: "invalid argument", 10: "no CUDA device", 11: "CUDA error", 12: "unsupported",
}
CRC_IEEE, CRC_CASTAGNOLI = 0, 1


class CubeecError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        self.name = ERR.get(code, str(code))
        super().__init__(f"cubeec error {code} ({self.name}) {detail}".strip())


_lib = None


def lib_path() -> str:
    return _SO


def load() -> C.CDLL:
    """Load libcubeec.so.  No fallback of any kind: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise ImportError(f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). cubefs_b200 has no CPU fallback.")
    L = C.CDLL(_SO)
    vp, u8p, szp, u32p, ip = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_int)
    L.cubeec_init.argtypes = [ip, C.c_int]
    L.cubeec_device_count.restype = C.c_int
    L.cubeec_strerror.restype = C.c_char_p
    L.cubeec_strerror.argtypes = [C.c_int]
    L.cubeec_last_error.restype = C.c_char_p
    L.cubeec_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.cubeec_host_free.argtypes = [vp]
    L.cubeec_host_register.argtypes = [vp, C.c_size_t]
    L.cubeec_host_unregister.argtypes = [vp]
    L.cubeec_create.argtypes = [C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.cubeec_destroy.argtypes = [vp]
    L.cubeec_destroy.restype = None
    L.cubeec_k.argtypes = [vp]
    L.cubeec_m.argtypes = [vp]
    L.cubeec_matrix.argtypes = [vp, vp]
    L.cubeec_decode_matrix.argtypes = [vp, vp, ip, vp]
    L.cubeec_encode.argtypes = [vp, vp, szp, C.c_int, vp, C.c_int]
    L.cubeec_set_coalescing.argtypes = [C.c_int, C.c_int]
    L.cubeec_verify.argtypes = [vp, vp, szp, C.c_int, ip]
    L.cubeec_reconstruct.argtypes = [vp, vp, szp, C.c_int, C.c_int, vp, vp, C.c_int]
    L.cubeec_encode_contig.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp, C.c_size_t, C.c_int]
    L.cubeec_reconstruct_batch.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
    L.cubeec_reconstruct_batch_crc.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, vp, C.c_int]
    L.cubeec_dev_encode.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_int, vp]
    L.cubeec_lrc_encode_contig.argtypes = [vp, vp, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_int]
    L.cubeec_dev_lrc_encode.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp,
                                        C.c_int, vp]
    L.cubeec_dev_lrc_verify.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp]
    L.cubeec_dev_lrc_reconstruct.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp,
                                             C.c_int, vp]
    L.cubeec_dev_reconstruct.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_int, vp]
    L.cubeec_dev_verify.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, vp]
    L.cubeec_crc32.argtypes = [vp, C.c_size_t, C.c_int, u32p]
    L.cubeec_crc32_blocks.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_int, vp, vp]
    L.cubeec_dev_crc32.argtypes = [C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, vp, vp, vp]
    L.cubeec_kernel_launches.restype = C.c_uint64
    L.cubeec_last_kernel.restype = C.c_char_p
    L.cubeec_debug_force_kernel.argtypes = [C.c_int]
    L.cubeec_debug_force_kernel.restype = None
    _lib = L
    return L


def _check(rc: int):
    if rc:
        detail = load().cubeec_last_error().decode() if rc == 11 else ""
        raise CubeecError(rc, detail)


def init(devices: Optional[Sequence[int]] = None):
    L = load()
    if devices is None:
        devices = [0]
    arr = (C.c_int * len(devices))(*devices)
    _check(L.cubeec_init(arr, len(devices)))


def device_count() -> int:
    return load().cubeec_device_count()


def kernel_launches() -> int:
    return int(load().cubeec_kernel_launches())


def last_kernel() -> str:
    return load().cubeec_last_kernel().decode()


def force_kernel(which: int) -> None:
    """Measurement aid: 0 = automatic choice, 1 = generic table kernel only."""
    load().cubeec_debug_force_kernel(which)


def set_coalescing(max_batch: int = 32, delay_us: int = 100) -> None:
    """cubeec_set_coalescing: batching of concurrent single-stripe encode calls (max_batch <= 1: off)."""
    _check(load().cubeec_set_coalescing(max_batch, delay_us))


def lrc_encode_contig(global_eng: "RSEngine", local_eng: "RSEngine", az_count: int, buf: np.ndarray, shard_len: int,
                      n_stripes: int, stripe_pitch: int, crc: bool = False, poly: int = 0, ptr: Optional[int] = None):
    """cubeec_lrc_encode_contig: global RS(N,M) + per-AZ local RS in one device pass sequence (lrcencoder.go:35-80).
    Returns the (n_stripes, N+M+L) checksums when crc is set."""
    n = global_eng.k + global_eng.m + local_eng.m * az_count
    crc_out = np.zeros((n_stripes, n), dtype=np.uint32) if crc else None
    base = ptr if ptr is not None else buf.ctypes.data
    _check(load().cubeec_lrc_encode_contig(global_eng._h, local_eng._h, az_count, base, shard_len, n_stripes, stripe_pitch,
                                           crc_out.ctypes.data if crc else None, poly))
    return crc_out


def dev_lrc_encode(global_eng: "RSEngine", local_eng: "RSEngine", az_count: int, d_base: int, shard_len: int,
                   shard_pitch: int, stripe_pitch: int, n_stripes: int, d_crc: int = 0, poly: int = 0, stream: int = 0,
                   device: int = 0):
    _check(load().cubeec_dev_lrc_encode(global_eng._h, local_eng._h, az_count, device, d_base, shard_len, shard_pitch,
                                        stripe_pitch, n_stripes, d_crc or None, poly, stream or None))


def dev_lrc_verify(global_eng: "RSEngine", local_eng: "RSEngine", az_count: int, d_base: int, shard_len: int, shard_pitch: int,
                   stripe_pitch: int, n_stripes: int, d_ok: int, stream: int = 0, device: int = 0):
    """lrcEncoder.Verify on device-resident LRC stripes: d_ok[s] = 1 iff global and every local parity match."""
    _check(load().cubeec_dev_lrc_verify(global_eng._h, local_eng._h, az_count, device, d_base, shard_len, shard_pitch, stripe_pitch,
                                        n_stripes, d_ok, stream or None))


def dev_lrc_reconstruct(global_eng: "RSEngine", local_eng: "RSEngine", az_count: int, d_base: int, shard_len: int, shard_pitch: int,
                        stripe_pitch: int, n_stripes: int, present: np.ndarray, data_only: bool = False, stream: int = 0,
                        device: int = 0):
    """lrcEncoder.Reconstruct / ReconstructData on device-resident LRC stripes (present: n_stripes x (N+M+L))."""
    present = np.ascontiguousarray(present, dtype=np.uint8)
    _check(load().cubeec_dev_lrc_reconstruct(global_eng._h, local_eng._h, az_count, device, d_base, shard_len, shard_pitch,
                                             stripe_pitch, n_stripes, present.ctypes.data, int(data_only), stream or None))


class StripeDesc(C.Structure):
    _fields_ = [("shards", C.c_void_p), ("present", C.c_void_p), ("shard_len", C.c_size_t)]


def _as_u8(a) -> np.ndarray:
    if isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if not isinstance(a, np.ndarray) else a, dtype=np.uint8)


class RSEngine:
    """One reedsolomon.New(k, m) handle (cubeec_create)."""

    def __init__(self, k: int, m: int, parity_rows: Optional[np.ndarray] = None):
        L = load()
        h = C.c_void_p()
        pr = None
        if parity_rows is not None:
            pr = np.ascontiguousarray(parity_rows, dtype=np.uint8)
            assert pr.shape == (m, k)
        _check(L.cubeec_create(k, m, pr.ctypes.data if pr is not None else None, C.byref(h)))
        self._h, self.k, self.m = h, k, m

    def close(self):
        if getattr(self, "_h", None):
            load().cubeec_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def matrix(self) -> np.ndarray:
        out = np.zeros((self.k + self.m, self.k), dtype=np.uint8)
        _check(load().cubeec_matrix(self._h, out.ctypes.data))
        return out

    def decode_matrix(self, present):
        present = np.ascontiguousarray(present, dtype=np.uint8)
        valid = (C.c_int * self.k)()
        rows = np.zeros((self.k, self.k), dtype=np.uint8)
        _check(load().cubeec_decode_matrix(self._h, present.ctypes.data, valid, rows.ctypes.data))
        return list(valid), rows

    # ---- host scatter API (Go [][]byte semantics: None / empty == missing) ----
    @staticmethod
    def _marshal(shards, size_hint=None):
        n = len(shards)
        size = size_hint
        if size is None:
            for s in shards:
                if s is not None and len(s):
                    size = len(s)
                    break
        bufs, keep = [], []
        ptrs = (C.c_void_p * n)()
        lens = (C.c_size_t * n)()
        for i, s in enumerate(shards):
            if s is None or len(s) == 0:
                b = np.zeros(max(size or 1, 1), dtype=np.uint8)   # the cap >= shardSize buffer the shim provides
                lens[i] = 0
            else:
                b = _as_u8(s)
                lens[i] = len(b)
            bufs.append(b)
            ptrs[i] = b.ctypes.data
        return bufs, ptrs, lens

    def encode(self, shards, crc: bool = False, poly: int = CRC_IEEE):
        """Parity shards (np.uint8 arrays) are overwritten in place.  Returns CRCs if crc."""
        bufs, ptrs, lens = self._marshal(shards)
        crc_out = np.zeros(len(shards), dtype=np.uint32) if crc else None
        _check(load().cubeec_encode(self._h, ptrs, lens, len(shards), crc_out.ctypes.data if crc else None, poly))
        return crc_out

    def verify(self, shards) -> bool:
        bufs, ptrs, lens = self._marshal(shards)
        ok = C.c_int(0)
        _check(load().cubeec_verify(self._h, ptrs, lens, len(shards), C.byref(ok)))
        return bool(ok.value)

    def reconstruct(self, shards, data_only: bool = False, crc: bool = False, poly: int = CRC_IEEE):
        """Returns (shards_out, crcs): missing entries replaced by regenerated arrays
        (missing parity stays None when data_only)."""
        bufs, ptrs, lens = self._marshal(shards)
        n = len(shards)
        filled = np.zeros(n, dtype=np.uint8)
        crc_out = np.zeros(n, dtype=np.uint32) if crc else None
        _check(load().cubeec_reconstruct(self._h, ptrs, lens, n, int(data_only), filled.ctypes.data,
                                         crc_out.ctypes.data if crc else None, poly))
        out = []
        for i, s in enumerate(shards):
            if s is not None and len(s):
                out.append(s)
            else:
                out.append(bufs[i] if filled[i] else None)
        return (out, crc_out) if crc else out

    def encode_contig(self, buf: np.ndarray, shard_len: int, n_stripes: int, stripe_pitch: int,
                      crc: bool = False, block_payload: int = 0, poly: int = CRC_IEEE, ptr: Optional[int] = None):
        n = self.k + self.m
        crc_out = np.zeros((n_stripes, n), dtype=np.uint32) if crc else None
        blk = None
        if block_payload:
            units = (shard_len + block_payload - 1) // block_payload
            blk = np.zeros((n_stripes, n, units), dtype=np.uint32)
        base = ptr if ptr is not None else buf.ctypes.data
        _check(load().cubeec_encode_contig(self._h, base, shard_len, n_stripes, stripe_pitch,
                                           crc_out.ctypes.data if crc else None,
                                           blk.ctypes.data if blk is not None else None, block_payload, poly))
        return crc_out, blk

    def encode_single_call_bench(self, blob: int, thread_counts, seconds: float = 3.0, check: bool = True):
        """The call shape of access (blobstore/common/ec/encoder.go:114-131): T host threads, each encoding ONE blob
        per cubeec_encode call (pageable memory, scatter pointers, CRCs requested); the engine's coalescing queue forms
        the batches.  Returns per T: stripes/s, data GiB/s, p50 / p99 call latency; with the queue switched off for
        the middle T as a comparison."""
        import threading
        import time
        import zlib
        k, m = self.k, self.m
        n = k + m
        S = max((blob + k - 1) // k, 2048)
        rng = np.random.default_rng(0xC0BEF5)
        data = [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(k)]   # shared, read-only (Encode never writes data)
        L = load()

        def run(T, secs):
            lat = [[] for _ in range(T)]
            bufs = [[np.zeros(S, np.uint8) for _ in range(m)] for _ in range(T)]
            crcs = [np.zeros(n, np.uint32) for _ in range(T)]
            ptrs, lens = [], []
            for t in range(T):
                pa = (C.c_void_p * n)(*([d.ctypes.data for d in data] + [b.ctypes.data for b in bufs[t]]))
                ptrs.append(pa)
                lens.append((C.c_size_t * n)(*([S] * n)))
            stop = [False]
            errs = []
            go = threading.Barrier(T + 1)

            def worker(t):
                go.wait()
                while not stop[0]:
                    t0 = time.perf_counter()
                    rc = L.cubeec_encode(self._h, ptrs[t], lens[t], n, crcs[t].ctypes.data, 0)
                    lat[t].append(time.perf_counter() - t0)
                    if rc:
                        errs.append(rc)
                        return

            th = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(T)]
            for x in th:
                x.start()
            go.wait()
            t0 = time.perf_counter()
            time.sleep(secs)
            stop[0] = True
            for x in th:
                x.join()
            el = time.perf_counter() - t0
            if errs:
                raise CubeecError(errs[0])
            allv = np.sort(np.concatenate([np.asarray(v) for v in lat if v]))
            calls = int(allv.size)
            out = {"threads": T, "calls": calls, "stripes_per_s": round(calls / el, 1),
                   "data_GiB_per_s": round(calls * k * S / el / (1 << 30), 3),
                   "p50_ms": round(float(allv[calls // 2]) * 1e3, 3), "p99_ms": round(float(allv[min(calls - 1, int(calls * 0.99))]) * 1e3, 3)}
            if check:
                want = [zlib.crc32(d.tobytes()) for d in data]
                for t in (0, T - 1):
                    assert [int(c) for c in crcs[t][:k]] == want, "single-call CRC mismatch"
                    assert all(int(crcs[t][k + r]) == zlib.crc32(bufs[t][r].tobytes()) for r in range(m)), "single-call parity CRC mismatch"
                    assert all(np.array_equal(bufs[t][r], bufs[0][r]) for r in range(m)), "single-call parity mismatch between threads"
                out["checked"] = True
            return out

        res = {"api": "cubeec_encode: one 4 MiB blob per call, pageable host memory, CRCs requested; callers block, the "
                      "library coalesces (cubeec_set_coalescing default 32 stripes / 100 us)",
               "shard_bytes": S, "runs": []}
        for T in thread_counts:
            run(T, 0.7)   # warm-up at this concurrency: pinned staging buffers of the batches, lanes, thread start-up
            res["runs"].append(run(T, seconds))
        mid = thread_counts[len(thread_counts) // 2]
        set_coalescing(1, 0)
        try:
            r = run(mid, seconds)
            r["coalescing"] = "off (every call its own H2D / kernel / D2H round trip)"
            res["runs"].append(r)
        finally:
            set_coalescing(32, 100)
        return res

    def reconstruct_batch(self, stripes, data_only: bool = False, verify: bool = False, crc: bool = False, poly: int = CRC_IEEE):
        """stripes: list of (shards list of np arrays (all allocated), present flags)."""
        n = self.k + self.m
        descs = (StripeDesc * len(stripes))()
        keep = []
        for i, (shards, present) in enumerate(stripes):
            ptrs = (C.c_void_p * n)(*[s.ctypes.data for s in shards])
            pres = np.ascontiguousarray(present, dtype=np.uint8)
            keep.append((ptrs, pres))
            descs[i].shards = C.cast(ptrs, C.c_void_p)
            descs[i].present = pres.ctypes.data
            descs[i].shard_len = len(shards[0])
        ok = (C.c_int * len(stripes))() if verify else None
        if crc:
            crcs = np.zeros((len(stripes), n), dtype=np.uint32)
            _check(load().cubeec_reconstruct_batch_crc(self._h, descs, len(stripes), int(data_only), ok, crcs.ctypes.data, poly))
            return ([bool(v) for v in ok] if verify else None), crcs
        _check(load().cubeec_reconstruct_batch(self._h, descs, len(stripes), int(data_only), ok))
        return [bool(v) for v in ok] if verify else None

    # ---- device-resident API (raw device pointers, e.g. torch tensor.data_ptr()) ----
    def dev_encode(self, d_base: int, shard_len: int, shard_pitch: int, stripe_pitch: int, n_stripes: int,
                   d_crc: int = 0, poly: int = CRC_IEEE, stream: int = 0, device: int = 0):
        _check(load().cubeec_dev_encode(self._h, device, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes,
                                        d_crc or None, poly, stream or None))

    def dev_reconstruct(self, d_base: int, shard_len: int, shard_pitch: int, stripe_pitch: int, n_stripes: int,
                        present: np.ndarray, data_only: bool = False, stream: int = 0, device: int = 0):
        present = np.ascontiguousarray(present, dtype=np.uint8)
        assert present.size == n_stripes * (self.k + self.m)
        _check(load().cubeec_dev_reconstruct(self._h, device, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes,
                                             present.ctypes.data, int(data_only), stream or None))

    def dev_verify(self, d_base: int, shard_len: int, shard_pitch: int, stripe_pitch: int, n_stripes: int,
                   d_ok: int, stream: int = 0, device: int = 0):
        _check(load().cubeec_dev_verify(self._h, device, d_base, shard_len, shard_pitch, stripe_pitch, n_stripes,
                                        d_ok, stream or None))


def crc32(data, poly: int = CRC_IEEE) -> int:
    a = _as_u8(data)
    out = C.c_uint32(0)
    _check(load().cubeec_crc32(a.ctypes.data if a.size else None, a.size, poly, C.byref(out)))
    return int(out.value)


def crc32_blocks(data, block_payload: int = 65532, poly: int = CRC_IEEE):
    a = _as_u8(data)
    units = (a.size + block_payload - 1) // block_payload
    per = np.zeros(max(units, 1), dtype=np.uint32)
    whole = C.c_uint32(0)
    _check(load().cubeec_crc32_blocks(a.ctypes.data if a.size else None, a.size, block_payload, poly,
                                      per.ctypes.data, C.byref(whole)))
    return per[:units], int(whole.value)


def dev_crc32(d_base: int, length: int, pitch: int, n_buffers: int, block_payload: int = 0, poly: int = CRC_IEEE,
              d_whole: int = 0, d_blocks: int = 0, stream: int = 0, device: int = 0):
    _check(load().cubeec_dev_crc32(device, d_base, length, pitch, n_buffers, block_payload, poly,
                                   d_whole or None, d_blocks or None, stream or None))
