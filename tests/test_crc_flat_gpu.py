"""Stand-alone CRC32 on the device (crc_flat_kernel, cubefs_b200/csrc/crc_flat.cu) against zlib / the oracle:
batched device buffers with a pitch, block payloads that are aligned to nothing (crc32block: 65,532 B, block.go:38-49),
both polynomials (IEEE: hash/crc32 in stream_put.go:265-269; Castagnoli: the oracle's bitwise restatement), and the
first-generation kernel (force 10) as a second opinion on the same inputs."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(cb, torch, host, length, block, poly=0):
    """host: (n_buffers, pitch) uint8.  Returns (whole[n], blocks[n][units] or None, kernel name)."""
    n, pitch = host.shape
    dev = torch.from_numpy(host).cuda()
    units = (length + block - 1) // block if block else 0
    whole = torch.zeros(n, dtype=torch.int32, device="cuda")
    blocks = torch.zeros(max(1, n * units), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    from cubefs_b200.engine import dev_crc32
    dev_crc32(dev.data_ptr(), length, pitch, n, block_payload=block, poly=poly, d_whole=whole.data_ptr(),
                 d_blocks=blocks.data_ptr() if block else 0, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    w = whole.cpu().numpy().view(np.uint32)
    b = blocks.cpu().numpy().view(np.uint32)[:n * units].reshape(n, units) if block else None
    return w, b, cb.last_kernel()


CASES = [
    # (buffers, length, pitch, block)
    (1, 1, 16, 0), (3, 15, 16, 0), (3, 16, 16, 0), (5, 17, 32, 0), (7, 63, 64, 0), (7, 64, 64, 0), (7, 65, 80, 0),
    (4, 2047, 2048, 0), (4, 2048, 2048, 0), (4, 2049, 2064, 0), (9, 4100, 4112, 0),
    (5, 65532, 65536, 0), (5, 349526, 349568, 0), (2, (1 << 20) + 3, (1 << 20) + 16, 0),
    (6, 349526, 349568, 65532),      # shard of the C2 blob in crc32block payloads: blocks start 65532 * u, aligned to 4 only
    (3, 200000, 200000, 1000),       # block aligned to nothing at all, 200 per buffer
    (11, 5000, 5008, 4092),
    (2, 1 << 20, 1 << 20, 1 << 16),
    (700, 300, 304, 0),              # many small buffers: most warps own several whole ranges
    (40, 3 * 2048, 3 * 2048 + 16, 2048),
]


@pytest.mark.parametrize("n,length,pitch,block", CASES)
def test_dev_crc32_matches_zlib(cb, n, length, pitch, block):
    import torch
    rng = np.random.default_rng(n * 1000003 + length)
    host = rng.integers(0, 256, (n, pitch), dtype=np.uint8)
    w, b, kern = _run(cb, torch, host, length, block)
    assert kern == "crc_flat_kernel"
    for i in range(n):
        raw = host[i, :length].tobytes()
        assert int(w[i]) == zlib.crc32(raw), (i, length)
        if block:
            assert [int(x) for x in b[i]] == [zlib.crc32(raw[o:o + block]) for o in range(0, length, block)], i
    # the first-generation kernel on the same device image
    cb.force_kernel(10)
    try:
        w2, b2, kern2 = _run(cb, torch, host, length, block)
    finally:
        cb.force_kernel(0)
    assert kern2 == "crc_range_kernel" and (w2 == w).all() and (b is None or (b2 == b).all())


def test_dev_crc32_castagnoli_and_zero_runs(cb, oracle):
    import torch
    rng = np.random.default_rng(77)
    host = rng.integers(0, 256, (4, 70000), dtype=np.uint8)
    host[1, :] = 0            # all-zero buffer: the init term alone
    host[2, 100:69000] = 0    # long zero run in the middle
    length = 69999
    host = np.ascontiguousarray(np.pad(host, ((0, 0), (0, 16 - 70000 % 16))))
    for poly in (0, 1):
        w, b, _ = _run(cb, torch, host, length, 65532, poly)
        for i in range(4):
            raw = host[i, :length].tobytes()
            assert int(w[i]) == oracle.crc32(raw, poly), (poly, i)
            assert [int(x) for x in b[i]] == [oracle.crc32(raw[o:o + 65532], poly) for o in range(0, length, 65532)]


def test_replica_mode_crc_uses_flat_kernel(cb):
    """m == 0 handles (replica code modes: the access layer only checksums, stream_put.go:265-269)."""
    import torch
    k, S, ns = 3, 100_003, 17
    P = (S + 127) // 128 * 128
    g = torch.Generator(device="cuda").manual_seed(5)
    batch = torch.randint(0, 256, (ns, k, P), dtype=torch.uint8, device="cuda", generator=g)
    dcrc = torch.zeros(ns * k, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    eng = cb.RSEngine(k, 0)
    eng.dev_encode(batch.data_ptr(), S, P, k * P, ns, d_crc=dcrc.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, device=0)
    torch.cuda.synchronize()
    got = dcrc.cpu().numpy().view(np.uint32).reshape(ns, k)
    h = batch.cpu().numpy()
    for s in range(ns):
        for i in range(k):
            assert int(got[s, i]) == zlib.crc32(h[s, i, :S].tobytes()), (s, i)
    assert cb.last_kernel() == "crc_flat_kernel"
