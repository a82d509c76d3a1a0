"""crc32block framing on the GPU against the oracle's restatement of
blobstore/common/crc32block (block.go:38-49, util.go:56-71, request_body.go:81-130, decode.go:84-107)."""
import ctypes as C
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BLOCK = 65536
PAYLOAD = BLOCK - 4


def _lib(cb):
    L = cb.load()
    vp = C.c_void_p
    L.cubeec_crc32block_encode_size.restype = C.c_size_t
    L.cubeec_crc32block_encode_size.argtypes = [C.c_size_t, C.c_size_t]
    L.cubeec_crc32block_decode_size.restype = C.c_size_t
    L.cubeec_crc32block_decode_size.argtypes = [C.c_size_t, C.c_size_t]
    L.cubeec_crc32block_encode.argtypes = [vp, C.c_size_t, C.c_size_t, vp, C.c_int]
    L.cubeec_crc32block_decode.argtypes = [vp, C.c_size_t, C.c_size_t, vp, C.POINTER(C.c_int64), C.c_int]
    L.cubeec_dev_crc32block_encode.argtypes = [C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_size_t, C.c_int, vp]
    L.cubeec_dev_crc32block_decode.argtypes = [C.c_int, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_size_t, vp, vp,
                                               C.c_int, vp]
    return L


SIZES = [1, 3, 4, 5, 4095, PAYLOAD - 1, PAYLOAD, PAYLOAD + 1, BLOCK - 1, BLOCK, BLOCK + 1, 2 * PAYLOAD - 1, 2 * PAYLOAD, 2 * PAYLOAD + 1,
         2 * PAYLOAD + 3, 349526, 1 << 20, (1 << 20) + 2]


@pytest.mark.parametrize("n", SIZES)
def test_body_encode_decode_matches_reference_framing(cb, oracle, n):
    L = _lib(cb)
    rng = np.random.default_rng(n)
    src = rng.integers(0, 256, n, dtype=np.uint8)
    want = np.frombuffer(oracle.crc32block_encode(src.tobytes(), BLOCK), dtype=np.uint8)
    assert L.cubeec_crc32block_encode_size(n, BLOCK) == want.size
    dst = np.zeros(want.size, dtype=np.uint8)
    assert L.cubeec_crc32block_encode(src.ctypes.data, n, BLOCK, dst.ctypes.data, 0) == 0
    assert np.array_equal(dst, want), "framed body differs from crc32block.NewBodyEncoder"
    # decode: all good
    assert L.cubeec_crc32block_decode_size(want.size, BLOCK) == n
    back = np.zeros(n, dtype=np.uint8)
    bad = C.c_int64(7)
    assert L.cubeec_crc32block_decode(dst.ctypes.data, dst.size, BLOCK, back.ctypes.data, C.byref(bad), 0) == 0
    assert bad.value == -1 and np.array_equal(back, src)
    # one flipped bit in the last block's payload and, separately, in a stored checksum
    blocks = (n + PAYLOAD - 1) // PAYLOAD
    for pos, blk in ((dst.size - 1, blocks - 1), (0, 0)):
        cor = dst.copy()
        cor[pos] ^= 0x10
        bad = C.c_int64(0)
        assert L.cubeec_crc32block_decode(cor.ctypes.data, cor.size, BLOCK, None, C.byref(bad), 0) == 0
        assert bad.value == blk


def test_block_len_rules(cb):
    L = _lib(cb)
    assert L.cubeec_crc32block_encode_size(100, 1000) == 0          # not a multiple of 4096: ErrInvalidBlock
    assert L.cubeec_crc32block_encode_size(100, 4096) == 104
    assert L.cubeec_crc32block_encode_size(4092 * 2 + 1, 4096) == 4092 * 2 + 1 + 12
    src = np.zeros(16, np.uint8)
    assert L.cubeec_crc32block_encode(src.ctypes.data, 16, 1000, src.ctypes.data, 0) == 9   # invalid argument
    # a trailing block with no payload cannot come from the encoder: the decoder reports it as bad
    bad = C.c_int64(0)
    framed = np.zeros(BLOCK + 3, np.uint8)
    assert L.cubeec_crc32block_decode(framed.ctypes.data, framed.size, BLOCK, None, C.byref(bad), 0) == 0
    assert bad.value == 1


def test_device_resident_shard_images_inspect(cb, oracle):
    """datainspect shape: many framed shard images in HBM, verify every block, report the first bad block of each
    (blobstore/blobnode/datainspect.go:246-283); and frame shards that are already device-resident."""
    import torch
    L = _lib(cb)
    n_buf, n = 37, 349526
    rng = np.random.default_rng(3)
    plain = rng.integers(0, 256, (n_buf, n), dtype=np.uint8)
    enc = L.cubeec_crc32block_encode_size(n, BLOCK)
    blocks = (n + PAYLOAD - 1) // PAYLOAD
    sp, dp = (n + 255) // 256 * 256, (enc + 255) // 256 * 256
    d_plain = torch.zeros((n_buf, sp), dtype=torch.uint8, device="cuda")
    d_plain[:, :n] = torch.from_numpy(plain).cuda()
    d_framed = torch.zeros((n_buf, dp), dtype=torch.uint8, device="cuda")
    assert L.cubeec_dev_crc32block_encode(0, d_plain.data_ptr(), n, sp, n_buf, BLOCK, d_framed.data_ptr(), dp, 0, None) == 0
    framed = d_framed.cpu().numpy()
    for b in (0, 17, n_buf - 1):
        want = np.frombuffer(oracle.crc32block_encode(plain[b].tobytes(), BLOCK), dtype=np.uint8)
        assert np.array_equal(framed[b, :enc], want)
    # corrupt block 3 of image 5 and block 0 of image 20
    d_framed[5, 3 * BLOCK + 100] ^= 1
    d_framed[20, 2] ^= 0x80
    d_bad = torch.zeros(n_buf, dtype=torch.int64, device="cuda")
    d_ok = torch.zeros((n_buf, blocks), dtype=torch.uint8, device="cuda")
    d_back = torch.zeros((n_buf, sp), dtype=torch.uint8, device="cuda")
    assert L.cubeec_dev_crc32block_decode(0, d_framed.data_ptr(), enc, dp, n_buf, BLOCK, d_back.data_ptr(), sp, d_bad.data_ptr(),
                                          d_ok.data_ptr(), 0, None) == 0
    bad = d_bad.cpu().numpy()
    ok = d_ok.cpu().numpy()
    want_bad = np.full(n_buf, -1, dtype=np.int64)
    want_bad[5], want_bad[20] = 3, 0
    assert np.array_equal(bad, want_bad)
    assert ok.sum() == n_buf * blocks - 2 and ok[5, 3] == 0 and ok[20, 0] == 0
    back = d_back.cpu().numpy()[:, :n]
    good = [b for b in range(n_buf) if b != 5]
    assert np.array_equal(back[good], plain[good])
    # whole-shard CRC of a frame-checked shard equals zlib (what blobnode stores next to the body)
    assert zlib.crc32(back[0].tobytes()) == zlib.crc32(plain[0].tobytes())
