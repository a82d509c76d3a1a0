"""CPU tests of the product's host arithmetic (cubefs_b200/csrc/gfmath.h) against the oracle.

engine.cu builds the coding / decoding matrices and every CRC constant the kernels use (Horner fold
tables, per-thread alignment constants, finalize terms) on the host with gfmath.h.  A small g++ harness
(tests/harness/gfmath_harness.cc) exposes those functions; here they are compared with the oracle's
restatement of the reference (buildMatrix RS/reedsolomon.go:220-244, matrix.Invert RS/matrix.go:193-266,
hash/crc32) -- on CPU, so a wrong constant is caught without a GPU."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POLYS = ((0xEDB88320, (1 << 32) - 1, 0), (0x82F63B78, (1 << 31) - 1, 1))   # (poly, order of x, oracle poly id)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "libgfmath_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "harness", "gfmath_harness.cc")])
    L = C.CDLL(so)
    L.hm_gf_mul.restype = C.c_uint8
    L.hm_gf_mul.argtypes = [C.c_uint8, C.c_uint8]
    L.hm_crc_mul.restype = C.c_uint32
    L.hm_crc_mul.argtypes = [C.c_uint32, C.c_int64, C.c_uint32, C.c_uint32]
    L.hm_crc_shift.restype = C.c_uint32
    L.hm_crc_shift.argtypes = [C.c_uint32, C.c_int64, C.c_int64]
    L.hm_crc_slice.argtypes = [C.c_uint32, C.c_void_p]
    L.hm_crc_constmul.argtypes = [C.c_uint32, C.c_int64, C.c_uint32, C.c_void_p]
    L.hm_build_generator.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.hm_invert.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return L


def test_gf_multiply_table(hm, oracle):
    mul = oracle.gf_tables()[2]
    got = np.array([[hm.hm_gf_mul(a, b) for b in range(256)] for a in range(0, 256, 5)], dtype=np.uint8)
    assert (got == mul[0:256:5]).all()


@pytest.mark.parametrize("km", [(4, 2), (6, 3), (12, 4), (20, 4), (15, 12), (16, 20), (24, 8), (6, 10), (18, 1), (3, 3)])
def test_generator_matrix_equals_reference_build_matrix(hm, oracle, km):
    k, m = km
    out = np.zeros((k + m, k), dtype=np.uint8)
    assert hm.hm_build_generator(k, k + m, out.ctypes.data) == 0
    assert (out == oracle.build_matrix(k, k + m)).all()


def test_inverse_equals_reference_and_rejects_singular(hm, oracle):
    rng = np.random.default_rng(5)
    for (k, m) in ((4, 2), (12, 4), (20, 4), (15, 12)):
        gen = oracle.build_matrix(k, k + m)
        for _ in range(8):
            rows = np.sort(rng.choice(k + m, size=k, replace=False))   # any k rows of the generator are invertible
            sub = np.ascontiguousarray(gen[rows])
            out = np.zeros_like(sub)
            assert hm.hm_invert(sub.ctypes.data, k, out.ctypes.data) == 0
            assert (out == oracle.matrix_invert(sub)).all()
    sing = np.zeros((3, 3), dtype=np.uint8)
    sing[0] = sing[1] = [1, 2, 3]
    sing[2] = [4, 5, 6]
    assert hm.hm_invert(sing.ctypes.data, 3, np.zeros_like(sing).ctypes.data) != 0


@pytest.mark.parametrize("poly,order,pid", POLYS)
def test_crc_slicing_tables_compute_the_reference_crc(hm, oracle, poly, order, pid):
    t = np.zeros((4, 256), dtype=np.uint32)
    hm.hm_crc_slice(poly, t.ctypes.data)
    data = np.random.default_rng(pid).integers(0, 256, 1000, dtype=np.uint8)
    reg = 0xFFFFFFFF
    i = 0
    while i + 4 <= len(data):                      # slicing-by-4 exactly as the kernels use the tables
        y = reg ^ int.from_bytes(data[i:i + 4].tobytes(), "little")
        reg = int(t[3][y & 255]) ^ int(t[2][(y >> 8) & 255]) ^ int(t[1][(y >> 16) & 255]) ^ int(t[0][y >> 24])
        i += 4
    assert (reg ^ 0xFFFFFFFF) == oracle.crc32(data.tobytes(), pid)
    if pid == 0:
        assert (reg ^ 0xFFFFFFFF) == zlib.crc32(data.tobytes())


@pytest.mark.parametrize("poly,order,pid", POLYS)
def test_crc_shift_constants_and_combine(hm, oracle, poly, order, pid):
    """crc(A || B) from crc(A), crc(B): the raw remainder of A times x^(8|B|) -- the Horner / alignment /
    finalize algebra of the kernels -- including NEGATIVE shifts through the order of x."""
    rng = np.random.default_rng(7 + pid)
    X0 = 0x80000000
    for na, nb in ((1, 1), (64, 32704), (349526, 349526 % 32768), (5, 1 << 20), (24576, 8)):
        a = rng.integers(0, 256, na, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, nb, dtype=np.uint8).tobytes()
        ca, cb, cab = oracle.crc32(a, pid), oracle.crc32(b, pid), oracle.crc32(a + b, pid)
        sh = hm.hm_crc_shift(poly, order, nb)
        # crc(A||B) = crc(A) * x^(8|B|) ^ crc(B): the init and xorout terms cancel (crc_combine_kernel)
        got = hm.hm_crc_mul(poly, order, ca, sh) ^ cb
        assert got == cab, (na, nb)
        assert got == oracle.crc32_combine(ca, cb, nb, pid)
        # x^(8n) * x^(-8n) = 1
        assert hm.hm_crc_mul(poly, order, sh, hm.hm_crc_shift(poly, order, -nb)) == X0
    # the order of x really is `order`
    assert hm.hm_crc_shift(poly, order, 0) == X0
    # raw remainder <-> CRC: crc(M) = ~(R(M) ^ 0xFFFFFFFF * x^(8|M|))  (what crc_finalize_kernel applies)
    m = rng.integers(0, 256, 777, dtype=np.uint8).tobytes()
    zeros_crc = oracle.crc32(bytes(len(m)), pid)
    init_term = hm.hm_crc_mul(poly, order, 0xFFFFFFFF, hm.hm_crc_shift(poly, order, len(m)))
    assert zeros_crc == (init_term ^ 0xFFFFFFFF)        # R(zeros) = 0


@pytest.mark.parametrize("poly,order,pid", POLYS)
def test_const_multiply_tables(hm, poly, order, pid):
    rng = np.random.default_rng(11 + pid)
    for const in (hm.hm_crc_shift(poly, order, 32768 - 64), hm.hm_crc_shift(poly, order, 24576 - 64), 0x12345678):
        t = np.zeros((4, 256), dtype=np.uint32)
        hm.hm_crc_constmul(poly, order, const, t.ctypes.data)
        for u in [int(x) for x in rng.integers(0, 1 << 32, 50, dtype=np.uint64)]:
            got = int(t[0][u & 255]) ^ int(t[1][(u >> 8) & 255]) ^ int(t[2][(u >> 16) & 255]) ^ int(t[3][u >> 24])
            assert got == hm.hm_crc_mul(poly, order, u, const)
