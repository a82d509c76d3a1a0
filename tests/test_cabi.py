"""CPU tests of the drop-in boundary: libcubeec.so loads and exports every symbol that
include/cubeec.h declares; argument validation that needs no GPU behaves like the reference."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cubeec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cubeec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cb):
    lib = ctypes.CDLL(cb.lib_path())
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/cubeec.h but not exported"


def test_create_argument_errors_without_gpu(cb):
    """reedsolomon.New argument checks (RS/reedsolomon.go:419-441) run before any device work."""
    with pytest.raises(cb.CubeecError) as e:
        cb.RSEngine(0, 2)
    assert e.value.name == "ErrInvShardNum"
    with pytest.raises(cb.CubeecError) as e:
        cb.RSEngine(3, -1)
    assert e.value.name == "ErrInvShardNum"
    with pytest.raises(cb.CubeecError) as e:
        cb.RSEngine(200, 100)
    assert e.value.name == "ErrMaxShardNum"


def test_no_cpu_fallback(cb):
    """Without a CUDA device every compute entry point must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CubeecError) as e:
        cb.RSEngine(4, 2)
    assert e.value.code == 10
    with pytest.raises(cb.CubeecError) as e:
        cb.crc32(b"123456789")
    assert e.value.code == 10


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cubefs_b200/ or include/ may reference it."""
    bad = []
    for base in ("cubefs_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    for line in txt.splitlines():
                        if re.search(r"^\s*(#\s*include|import|from)\b.*oracle|(dlopen|CDLL)\(.*oracle|-l\S*oracle", line):
                            bad.append((f, line.strip()))
    assert not bad, bad


def test_runtime_codegen_compiles_for_every_erasure_class(cb):
    """The NVRTC-specialised reconstruct kernels (csrc/jit.cu): generate + compile the source for single, double,
    triple and quadruple erasures, data-only, and the m > 4 codes (several passes).  No device is needed to compile."""
    import ctypes as C

    import numpy as np
    L = cb.load()
    L.cubeec_debug_jit_check.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]
    cases = [(12, 4, [3], 0), (12, 4, [1, 7, 13], 0), (12, 4, [0, 1, 2, 3], 0), (12, 4, [0, 13], 1), (6, 3, [8], 0),
             (4, 2, [0, 5], 0), (6, 10, [0, 1, 2, 3, 4, 7, 9], 0), (3, 1, [1], 0)]
    for k, m, miss, data_only in cases:
        pres = np.ones(k + m, np.uint8)
        pres[miss] = 0
        log = C.create_string_buffer(8000)
        rc = L.cubeec_debug_jit_check(k, m, pres.ctypes.data, data_only, log, 8000)
        if rc == 101:
            pytest.skip("libnvrtc is not installed here: the engine falls back to the table kernels")
        assert rc == 0, (k, m, miss, rc, log.value[:2000])
    # too few survivors is reported as such, before any code generation
    pres = np.ones(6, np.uint8)
    pres[[0, 1, 2]] = 0
    assert L.cubeec_debug_jit_check(4, 2, pres.ctypes.data, 0, None, 0) == 3


def test_crc32block_size_math_without_gpu(cb):
    """crc32block.EncodeSize / DecodeSize (blobstore/common/crc32block/util.go:56-71) are host arithmetic in the ABI:
    size + 4 * ceil(size / (blockLen - 4)) and back; blockLen must be a positive multiple of 4096 (isValidBlockLen,
    util.go:40-42; the reference panics with ErrInvalidBlock), else 0.  The on-disk sizes the reference's tests pin (datafile_test.go:194-195,237-241: a
    9-byte shard occupies 32 + 9 + 4 + 8 bytes) follow from it."""
    L = cb.load()
    L.cubeec_crc32block_encode_size.restype = ctypes.c_size_t
    L.cubeec_crc32block_encode_size.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
    L.cubeec_crc32block_decode_size.restype = ctypes.c_size_t
    L.cubeec_crc32block_decode_size.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
    for block in (4096, 8192, 65536, 1 << 20):
        payload = block - 4
        for n in (0, 1, 9, payload - 1, payload, payload + 1, 2 * payload, 2 * payload + 1, 349526, (1 << 22) + 5, 1 << 30):
            enc = L.cubeec_crc32block_encode_size(n, block)
            assert enc == n + 4 * ((n + payload - 1) // payload), (block, n)
            assert L.cubeec_crc32block_decode_size(enc, block) == n, (block, n)
    assert L.cubeec_crc32block_encode_size(9, 65536) == 13          # 32-byte header + 13 + 8-byte footer = the 53 bytes on disk
    for bad in (0, 1000, 4095, 4097, 65536 + 4):
        assert L.cubeec_crc32block_encode_size(100, bad) == 0 and L.cubeec_crc32block_decode_size(100, bad) == 0
