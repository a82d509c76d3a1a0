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
