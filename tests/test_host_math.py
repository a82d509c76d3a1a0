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


@pytest.mark.parametrize("poly,order,pid", POLYS)
def test_crc_flat_kernel_decomposition(hm, oracle, poly, order, pid):
    """The algebra of crc_flat_kernel / crc_flat_finish_kernel (cubefs_b200/csrc/crc_flat.cu) replayed with the constants
    the engine builds on the host (engine.cu: fold tables x^(8*(2048-64)), klane x^(8*64*(31-lane)), x_unit_pow, x_neg_pow,
    init term): 2 KiB tiles of the BUFFER, 32 lanes x 64 contiguous bytes, bytes outside [a, e) masked to zero, Horner
    across the tiles of a run, runs combined by XOR behind x^(8*2048*tiles after), x^(-8z) for the zero bytes behind the
    range.  Ranges as the crc32block payloads have them: starting 4 bytes into a block, aligned to nothing."""
    TILE, PIECE = 2048, 64
    sl = np.zeros((4, 256), dtype=np.uint32)
    hm.hm_crc_slice(poly, sl.ctypes.data)
    fold_t = np.zeros((4, 256), dtype=np.uint32)
    hm.hm_crc_constmul(poly, order, hm.hm_crc_shift(poly, order, TILE - PIECE), fold_t.ctypes.data)
    klane = [hm.hm_crc_shift(poly, order, PIECE * (31 - l)) for l in range(32)]
    x_unit_pow = [hm.hm_crc_shift(poly, order, TILE << i) for i in range(24)]
    x_neg_pow = [hm.hm_crc_shift(poly, order, -(1 << i)) for i in range(33)]
    mul = lambda a, b: hm.hm_crc_mul(poly, order, a, b)   # noqa: E731

    def slice4(y):
        return int(sl[3][y & 255]) ^ int(sl[2][(y >> 8) & 255]) ^ int(sl[1][(y >> 16) & 255]) ^ int(sl[0][y >> 24])

    def fold(u):
        return int(fold_t[0][u & 255]) ^ int(fold_t[1][(u >> 8) & 255]) ^ int(fold_t[2][(u >> 16) & 255]) ^ int(fold_t[3][u >> 24])

    def model(buf, a, e, T, cuts):
        t0 = a // TILE
        out = 0
        bounds = [0] + sorted(cuts) + [T]
        for lo, hi in zip(bounds[:-1], bounds[1:]):          # one warp run each
            u = [0] * 32
            for t in range(lo, hi):
                for lane in range(32):
                    col = (t0 + t) * TILE + lane * PIECE
                    piece = bytes(buf[i] if a <= i < e else 0 for i in range(col, col + PIECE))   # masked; past the buffer: zero
                    for w in range(0, PIECE, 4):
                        u[lane] = slice4(u[lane] ^ int.from_bytes(piece[w:w + 4], "little"))
                if t + 1 == hi:
                    v = 0
                    for lane in range(32):
                        v ^= mul(u[lane], klane[lane])
                    n, i = T - 1 - t, 0
                    while n:
                        if n & 1:
                            v = mul(v, x_unit_pow[i])
                        n >>= 1
                        i += 1
                    out ^= v                                     # atomicXor into the range's slot
                else:
                    u = [fold(x) for x in u]
        z, i = (t0 + T) * TILE - e, 0
        while z:
            if z & 1:
                out = mul(out, x_neg_pow[i])
            z >>= 1
            i += 1
        init_term = mul(0xFFFFFFFF, hm.hm_crc_shift(poly, order, e - a))
        return (out ^ init_term) ^ 0xFFFFFFFF

    class Padded(bytes):
        def __getitem__(self, i):
            return bytes.__getitem__(self, i) if i < len(self) else 0

    rng = np.random.default_rng(99 + pid)
    raw = Padded(rng.integers(0, 256, 9000, dtype=np.uint8).tobytes())
    block, stride, offset = 2996, 3000, 4                       # payload / block / header of a scaled-down crc32block image
    T = (block + TILE - 1) // TILE + 1                          # the engine's uniform upper bound for unaligned ranges
    for u in range(3):
        a = offset + u * stride
        e = min(a + block, len(raw))
        want = oracle.crc32(bytes.__getitem__(raw, slice(a, e)), pid)
        for cuts in ([], [1], [1, 2]):                          # whole range in one run; runs that end inside the range
            assert model(raw, a, e, T, cuts) == want, (u, cuts)
    # a whole 2 KiB-aligned buffer (tiles_per_range exact) and a 1-byte range
    assert model(raw, 0, 4096, 2, [1]) == oracle.crc32(bytes.__getitem__(raw, slice(0, 4096)), pid)
    assert model(raw, 4099, 4100, 1, []) == oracle.crc32(bytes.__getitem__(raw, slice(4099, 4100)), pid)


def test_flat_split_part_slots_bound():
    """The flat work split of the fused kernel (bs_flat.cuh: warp g owns units [g*U/GW, (g+1)*U/GW); a run writes its CRC
    remainder of stripe s into part slot  g - owner(first unit of s))  and the number of part slots engine.cu reserves per
    shard (flat_geometry: (wt-1)/(U/GW) + 2, or ceil(wt*GW/U) + 2 when there are more warps than units).  Property: every
    part index any run produces is below that bound, for batches from one stripe to thousands, ragged or not."""
    def run_lo(g, U, GW):
        return g * U // GW

    def owner(u, U, GW):            # bsf_owner: the warp g with lo(g) <= u < lo(g+1)
        g = u * GW // U
        while g + 1 < GW and run_lo(g + 1, U, GW) <= u:
            g += 1
        while g > 0 and run_lo(g, U, GW) > u:
            g -= 1
        return g

    rng = np.random.default_rng(2024)
    cases = [(1, 8), (1, 171), (16, 4096), (383, 171), (1024, 171), (85, 512), (3, 1), (5000, 8), (149, 171), (2, 100000)]
    cases += [(int(rng.integers(1, 3000)), int(rng.integers(1, 600))) for _ in range(40)]
    for n_stripes, wt in cases:
        nw = 12                                                   # 384 threads
        U = n_stripes * wt
        grid = max(1, min((U + nw - 1) // nw, 148))
        GW = grid * nw
        max_parts = (wt - 1) // (U // GW) + 2 if U >= GW else (wt * GW + U - 1) // U + 2
        worst = 0
        for g in range(GW):
            lo, hi = run_lo(g, U, GW), run_lo(g + 1, U, GW)
            assert 0 <= hi - lo <= U // GW + 1
            if lo == hi:
                continue
            assert owner(lo, U, GW) == g and owner(hi - 1, U, GW) == g
            for s in range(lo // wt, (hi - 1) // wt + 1):        # every stripe this run touches
                part = g - owner(s * wt, U, GW)
                assert 0 <= part < max_parts, (n_stripes, wt, g, s, part, max_parts)
                worst = max(worst, part + 1)
        assert worst >= 1
