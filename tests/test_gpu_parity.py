"""GPU parity tests: the CUDA path through the C-ABI vs the CPU oracle, bit-exact.

Mirrors blobstore/common/ec/encoder_test.go (split/encode/corrupt/reconstruct/verify round trips,
all code modes with cumulative erasures) and adds direct comparisons with the oracle on the same
seeded inputs, the reference's golden CRC vectors, ragged / tiny / unaligned sizes."""
import json
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CODE_MODES = [(15, 12), (6, 6), (16, 20), (6, 10), (6, 3), (4, 4), (12, 4), (16, 4), (3, 3), (10, 4), (12, 9), (24, 8)]


def _rand_shards(rng, k, m, S):
    return [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(k)] + [np.zeros(S, np.uint8) for _ in range(m)]


@pytest.mark.parametrize("km", [(4, 2), (6, 3), (12, 4), (20, 4)] + CODE_MODES)
@pytest.mark.parametrize("S", [1, 17, 2048, 65536 + 6])
def test_encode_matches_oracle(cb, oracle, km, S):
    k, m = km
    rng = np.random.default_rng(S * 1000 + k * 10 + m)
    sh = _rand_shards(rng, k, m, S)
    want = [s.copy() for s in sh]
    oracle.RS(k, m).encode(want)
    eng = cb.RSEngine(k, m)
    assert (eng.matrix == oracle.RS(k, m).matrix).all()
    crc = eng.encode(sh, crc=True)
    for i in range(k + m):
        assert (sh[i] == want[i]).all(), f"shard {i}"
        assert crc[i] == zlib.crc32(want[i].tobytes()), f"crc {i}"
    assert eng.verify(sh)


def test_config_c1_rs42_64k(cb, oracle):
    """BASELINE config 1: RS(4,2), 64 KiB shards, single stripe."""
    rng = np.random.default_rng(0xC0BEF5)
    sh = _rand_shards(rng, 4, 2, 65536)
    want = [s.copy() for s in sh]
    oracle.RS(4, 2).encode(want)
    crc = cb.RSEngine(4, 2).encode(sh, crc=True)
    assert all((a == b).all() for a, b in zip(sh, want))
    assert [int(c) for c in crc] == [zlib.crc32(w.tobytes()) for w in want]


def test_config_c2_shape_single_stripe(cb, oracle):
    """EC12P4 with the production shard size 349526 (4 MiB blob, pad 8 bytes): parity + CRC."""
    rng = np.random.default_rng(7)
    S = 349526
    sh = _rand_shards(rng, 12, 4, S)
    sh[11][-8:] = 0
    want = [s.copy() for s in sh]
    oracle.RS(12, 4).encode(want)
    eng = cb.RSEngine(12, 4)
    crc = eng.encode(sh, crc=True)
    assert all((a == b).all() for a, b in zip(sh, want))
    assert [int(c) for c in crc] == [zlib.crc32(w.tobytes()) for w in want]
    crc_c = eng.encode(sh, crc=True, poly=1)
    assert [int(c) for c in crc_c] == [oracle.crc32(w, 1) for w in want]


def test_small_kats(cb):
    """SURVEY 8c known answers straight through the C-ABI."""
    eng = cb.RSEngine(4, 2)
    sh = [np.array([16 * c + i for i in range(8)], dtype=np.uint8) for c in range(4)] + [np.zeros(8, np.uint8) for _ in range(2)]
    crc = eng.encode(sh, crc=True)
    assert sh[4].tobytes().hex() == "4041424344454647" and sh[5].tobytes().hex() == "5051525354555657"
    assert [int(c) for c in crc] == [2292869279, 3954419385, 1318712531, 763378421, 3753662022, 3164969056]
    valid, rows = eng.decode_matrix([0, 0, 1, 1, 1, 1])
    assert valid == [2, 3, 4, 5] and rows[0].tobytes().hex() == "d06b68d2"
    eng = cb.RSEngine(12, 4)
    assert eng.matrix[12].tobytes().hex() == "afb4968cf5e8c4d81b1c1214"
    sh = [np.array([(7 * c + 13 * i + 1) & 255 for i in range(4)], dtype=np.uint8) for c in range(12)] + [np.zeros(4, np.uint8) for _ in range(4)]
    eng.encode(sh)
    assert [s.tobytes().hex() for s in sh[12:]] == ["e82088e5", "8671d66a", "e0dba9fe", "ceb6d765"]


@pytest.mark.parametrize("km", CODE_MODES)
def test_reconstruct_all_modes_cumulative_erasures(cb, oracle, km):
    """encoder_test.go:249-307: for every code mode, erase 1..M shards, reconstruct, compare."""
    k, m = km
    rng = np.random.default_rng(k * 131 + m)
    S = int(rng.integers(64 << 10, 128 << 10))
    sh = _rand_shards(rng, k, m, S)
    eng = cb.RSEngine(k, m)
    eng.encode(sh)
    orig = [s.copy() for s in sh]
    ora = oracle.RS(k, m)
    assert ora.verify(orig)
    order = rng.permutation(k + m)
    for e in sorted({1, 2, max(1, m // 2), m}):
        broken = [None if i in order[:e] else orig[i].copy() for i in range(k + m)]
        out, crc = eng.reconstruct(broken, crc=True)
        for i in range(k + m):
            assert (out[i] == orig[i]).all(), (e, i)
        for i in order[:e]:
            assert crc[i] == zlib.crc32(orig[i].tobytes())
        out = eng.reconstruct(broken, data_only=True)
        for i in range(k + m):
            if i in order[:e] and i >= k:
                assert out[i] is None
            else:
                assert (out[i] == orig[i]).all()
    broken = [None if i in order[:m + 1] else orig[i] for i in range(k + m)]
    with pytest.raises(cb.CubeecError) as e:
        eng.reconstruct(broken)
    assert e.value.name == "ErrTooFewShards"
    # all present: no-op
    out = eng.reconstruct([s.copy() for s in orig])
    assert all((a == b).all() for a, b in zip(out, orig))


def test_encoder_test_flow_ec15p12(cb):
    """encoder_test.go:53-106: corrupt a data shard -> ReconstructData; corrupt parity -> Reconstruct + Verify."""
    rng = np.random.default_rng(15)
    S = 4096 + 3
    sh = _rand_shards(rng, 15, 12, S)
    eng = cb.RSEngine(15, 12)
    eng.encode(sh)
    assert eng.verify(sh)
    orig = [s.copy() for s in sh]
    sh[0][:] = 11
    assert not eng.verify(sh)
    sh[0] = None
    sh = eng.reconstruct(sh, data_only=True)
    assert (sh[0] == orig[0]).all()
    sh[16][:] = 11
    sh[16] = None
    sh = eng.reconstruct(sh)
    assert eng.verify(sh)
    assert all((a == b).all() for a, b in zip(sh, orig))


def test_error_cases(cb):
    eng = cb.RSEngine(4, 2)
    with pytest.raises(cb.CubeecError) as e:
        eng.encode([np.zeros(8, np.uint8)] * 5)
    assert e.value.name == "ErrTooFewShards"
    sh = [np.zeros(8, np.uint8) for _ in range(6)]
    sh[2] = np.zeros(9, np.uint8)
    with pytest.raises(cb.CubeecError) as e:
        eng.encode(sh)
    assert e.value.name == "ErrShardSize"
    with pytest.raises(cb.CubeecError) as e:
        eng.encode([None] * 6)
    assert e.value.name == "ErrShardNoData"
    with pytest.raises(cb.CubeecError) as e:
        eng.verify([np.zeros(8, np.uint8)] * 7)
    assert e.value.name == "ErrTooFewShards"


def _crc_inputs():
    def z12(n):
        b = bytearray(n)
        b[0], b[-1] = ord("1"), ord("2")
        return bytes(b)
    big = bytearray(1 << 20)
    for pos, ch in ((0, "1"), (65531, "2"), (65532, "3"), (131063, "4"), (131064, "5"), (196595, "6"), (1048575, "0")):
        big[pos] = ord(ch)
    return [b"test data", bytes(ord("0") + i % 10 for i in range(32768)), z12(65492), z12(65532), z12(65536),
            z12(65530), bytes(big)]


def test_crc32_reference_golden_vectors(cb):
    gold = json.load(open(os.path.join(HERE, "golden", "crc_golden.json")))["cases"]
    for case, data in zip(gold, _crc_inputs()):
        assert cb.crc32(data) == case["crc32_ieee"], case["input"]


def test_crc32_sizes_and_blocks(cb, oracle):
    assert cb.crc32(b"123456789") == 0xCBF43926
    assert cb.crc32(b"123456789", 1) == 0xE3069283
    assert cb.crc32(b"") == 0
    assert cb.crc32(bytes(2048)) == 4058561182 and cb.crc32(bytes(349526)) == 2136928390
    rng = np.random.default_rng(5)
    for n in (1, 3, 15, 16, 17, 8191, 8192, 8193, 65531, 65532, 65533, 131064, 349526, 1 << 20, (1 << 22) + 5):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert cb.crc32(d) == zlib.crc32(d.tobytes()), n
        per, whole = cb.crc32_blocks(d, 65532)
        assert whole == zlib.crc32(d.tobytes())
        raw = d.tobytes()
        assert [int(x) for x in per] == [zlib.crc32(raw[o:o + 65532]) for o in range(0, n, 65532)], n
        # the framed body the blobnode writes (datafile.go:342) rebuilt from GPU block CRCs == oracle framing
        framed = b"".join(int(c).to_bytes(4, "little") + raw[o:o + 65532] for c, o in zip(per, range(0, n, 65532)))
        assert framed == oracle.crc32block_encode(raw)


def test_encode_contig_batch(cb, oracle):
    """ec.Buffer layout, many stripes, unaligned shard length (6 mod 16 like 349526)."""
    k, m, S, ns = 12, 4, 21846, 37
    rng = np.random.default_rng(9)
    buf = rng.integers(0, 256, (ns, (k + m) * S), dtype=np.uint8)
    ref = buf.copy()
    eng = cb.RSEngine(k, m)
    crc, blk = eng.encode_contig(buf, S, ns, (k + m) * S, crc=True, block_payload=65532)
    ora = oracle.RS(k, m)
    for s in range(ns):
        sh = [ref[s, i * S:(i + 1) * S].copy() for i in range(k + m)]
        ora.encode(sh)
        for i in range(k + m):
            assert (buf[s, i * S:(i + 1) * S] == sh[i]).all(), (s, i)
            assert crc[s, i] == zlib.crc32(sh[i].tobytes())
            assert blk[s, i, 0] == zlib.crc32(sh[i].tobytes())
    # shards spanning several crc32block payloads (65,532 B; the last block is short): per-block CRCs come
    # from the device-resident stripes, in blobnode framing order (crc32block/block.go:38-49)
    k, m, S, ns = 6, 3, 65532 * 2 + 4711, 5
    buf = rng.integers(0, 256, (ns, (k + m) * S), dtype=np.uint8)
    ref = buf.copy()
    eng = cb.RSEngine(k, m)
    crc, blk = eng.encode_contig(buf, S, ns, (k + m) * S, crc=True, block_payload=65532)
    ora = oracle.RS(k, m)
    assert blk.shape == (ns, k + m, 3)
    for s in range(ns):
        sh = [ref[s, i * S:(i + 1) * S].copy() for i in range(k + m)]
        ora.encode(sh)
        for i in range(k + m):
            raw = sh[i].tobytes()
            assert (buf[s, i * S:(i + 1) * S] == sh[i]).all(), (s, i)
            assert crc[s, i] == zlib.crc32(raw)
            assert [int(x) for x in blk[s, i]] == [zlib.crc32(raw[o:o + 65532]) for o in range(0, S, 65532)], (s, i)


def test_reconstruct_batch_with_verify(cb, oracle):
    """The repair loop (worker_slice_recover.go:822-885): variable sizes, independent patterns."""
    k, m = 6, 3
    eng = cb.RSEngine(k, m)
    rng = np.random.default_rng(11)
    stripes, origs = [], []
    for s in range(25):
        S = int(rng.integers(1, 70000))
        sh = _rand_shards(rng, k, m, S)
        eng.encode(sh)
        origs.append([x.copy() for x in sh])
        present = np.ones(k + m, dtype=np.uint8)
        miss = rng.choice(k + m, size=int(rng.integers(0, m + 1)), replace=False)
        present[miss] = 0
        for i in miss:
            sh[i][:] = 0x5A
        stripes.append((sh, present))
    ok = eng.reconstruct_batch(stripes, verify=True)
    assert all(ok)
    for (sh, _), orig in zip(stripes, origs):
        assert all((a == b).all() for a, b in zip(sh, orig))
    # a corrupted survivor must make Verify fail for that stripe only
    sh, present = stripes[3]
    bad = [x.copy() for x in sh]
    bad[0][0] ^= 1
    ok = eng.reconstruct_batch([(bad, np.ones(k + m, np.uint8)), stripes[4]], verify=True)
    assert ok == [False, True]


def test_device_resident_batch(cb, oracle):
    """cubeec_dev_encode / dev_reconstruct / dev_verify on a pitched HBM layout, fused CRC."""
    import torch
    k, m, S, P, ns = 12, 4, 349526, 349568, 9
    n = k + m
    rng = np.random.default_rng(21)
    host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
    dev = torch.from_numpy(host).cuda()
    dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
    eng = cb.RSEngine(k, m)
    eng.dev_encode(dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr())
    torch.cuda.synchronize()
    out = dev.cpu().numpy()
    crc = dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)
    ora = oracle.RS(k, m)
    for s in range(ns):
        sh = [host[s, i, :S].copy() for i in range(n)]
        ora.encode(sh)
        for i in range(n):
            assert (out[s, i, :S] == sh[i]).all(), (s, i)
            assert crc[s, i] == zlib.crc32(sh[i].tobytes()), (s, i)
        assert (out[s, k:, S:(S + 31) // 32 * 32] == 0).all()   # output pad bytes up to the 32-byte boundary are zeroed
    # verify
    dok = torch.zeros(ns, dtype=torch.int32, device="cuda")
    eng.dev_verify(dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
    assert dok.cpu().tolist() == [1] * ns
    dev[2, 13, 100] ^= 1
    eng.dev_verify(dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
    assert dok.cpu().tolist() == [1, 1, 0] + [1] * (ns - 3)
    dev[2, 13, 100] ^= 1
    # reconstruct: 3 random erasures per stripe (BASELINE config 3)
    good = dev.clone()
    present = np.ones((ns, n), dtype=np.uint8)
    for s in range(ns):
        miss = rng.choice(n, size=3, replace=False)
        present[s, miss] = 0
        for i in miss:
            dev[s, int(i)].fill_(0xEE)
    eng.dev_reconstruct(dev.data_ptr(), S, P, n * P, ns, present)
    torch.cuda.synchronize()
    assert torch.equal(dev[:, :, :S], good[:, :, :S])


def test_linearity_and_roundtrip_at_scale(cb):
    """Size-independent properties at a large batch (oracle-free): encode -> erase -> decode round trip,
    GF(2)-linearity of parity, and verify == 1."""
    import torch
    k, m, S, P, ns = 12, 4, 349526, 349568, 64
    n = k + m
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(0, 256, (ns, n, P), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(0, 256, (ns, n, P), dtype=torch.uint8, device="cuda", generator=g)
    eng = cb.RSEngine(k, m)
    c = a ^ b
    for t in (a, b, c):
        eng.dev_encode(t.data_ptr(), S, P, n * P, ns)
    torch.cuda.synchronize()
    assert torch.equal(a[:, k:, :S] ^ b[:, k:, :S], c[:, k:, :S])
    dok = torch.zeros(ns, dtype=torch.int32, device="cuda")
    eng.dev_verify(c.data_ptr(), S, P, n * P, ns, dok.data_ptr())
    assert int(dok.sum()) == ns
    good = c.clone()
    rng = np.random.default_rng(8)
    present = np.ones((ns, n), dtype=np.uint8)
    for s in range(ns):
        miss = rng.choice(n, size=int(rng.integers(1, m + 1)), replace=False)
        present[s, miss] = 0
        for i in miss:
            c[s, int(i)].zero_()
    eng.dev_reconstruct(c.data_ptr(), S, P, n * P, ns, present)
    torch.cuda.synchronize()
    assert torch.equal(c[:, :, :S], good[:, :, :S])


def test_lrc_encode_fused(cb, oracle):
    """cubeec_lrc_encode_contig / cubeec_dev_lrc_encode == lrcEncoder.Encode (lrcencoder.go:35-80): global
    RS(N,M), then per AZ the local RS over [AZ data | AZ global parity] (codemode.GetECLayoutByAZ); parity
    bytes and the CRC of every one of the N+M+L shards against the oracle.  All LRC code modes of
    codemode.go (EC16P20L2, EC6P10L2, EC6P3L3, EC4P4L2, EC6P8L10) + a packed-size and a ragged-size case."""
    import torch
    modes = ((16, 20, 2, 2, 70001), (6, 10, 2, 2, 9000), (6, 3, 3, 3, 2048), (4, 4, 2, 2, 33000), (6, 8, 10, 2, 4097),
             (6, 10, 2, 2, 349526 // 4))
    for (N, M, L, az, S) in modes:
        n, kl, ml = N + M + L, (N + M) // az, L // az
        ns = 5
        g_eng, l_eng = cb.RSEngine(N, M), cb.RSEngine(kl, ml)
        g_ora, l_ora = oracle.RS(N, M), oracle.RS(kl, ml)
        rng = np.random.default_rng(N * 100 + M)
        # ---- host ec.Buffer layout (shards back to back, pitch = S) ----
        host = rng.integers(0, 256, (ns, n, S), dtype=np.uint8)
        want = host.copy()
        for s in range(ns):
            sh = [want[s, i].copy() for i in range(N + M)]
            g_ora.encode(sh)
            for i in range(N + M):
                want[s, i] = sh[i]
            for a in range(az):
                idx = list(range(a * N // az, (a + 1) * N // az)) + list(range(N + a * M // az, N + (a + 1) * M // az))
                loc = [want[s, i].copy() for i in idx] + [np.zeros(S, np.uint8) for _ in range(ml)]
                l_ora.encode(loc)
                for r in range(ml):
                    want[s, N + M + a * ml + r] = loc[kl + r]
        buf = host.copy()
        buf[:, N:, :] = 0x5A   # parity areas hold garbage before the call
        crc = cb.lrc_encode_contig(g_eng, l_eng, az, buf, S, ns, n * S, crc=True)
        assert (buf == want).all(), (N, M, L, S)
        for s in range(ns):
            for i in range(n):
                assert crc[s, i] == zlib.crc32(want[s, i].tobytes()), (N, M, L, S, s, i)
        # ---- device-resident pitched layout, Castagnoli ----
        P = (S + 127) // 128 * 128
        dev = torch.zeros((ns, n, P), dtype=torch.uint8, device="cuda")
        dev[:, :N, :S] = torch.from_numpy(host[:, :N, :]).cuda()
        dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
        cb.dev_lrc_encode(g_eng, l_eng, az, dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), poly=1)
        torch.cuda.synchronize()
        assert (dev[:, :, :S].cpu().numpy() == want).all(), (N, M, L, S, "dev")
        got = dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)
        for s in (0, ns - 1):
            for i in range(n):
                assert got[s, i] == oracle.crc32(want[s, i].tobytes(), 1), (N, M, L, S, s, i, "crc32c")
    # layout errors: local code that does not match the AZ split
    with pytest.raises(cb.CubeecError):
        cb.lrc_encode_contig(cb.RSEngine(6, 10), cb.RSEngine(4, 1), 2, np.zeros(18 * 64, np.uint8), 64, 1, 18 * 64)


def test_rolled_fused_kernel(cb, oracle):
    """Opt-in rolled-loop variant of the fused encode+CRC kernel (cubeec_debug_force_kernel(6)): same parity
    bytes and CRCs (both polynomials) as the oracle on whole-stripe, segmented and ragged geometries."""
    import torch
    for (k, m) in ((12, 4), (10, 4), (6, 2)):
        n = k + m
        eng = cb.RSEngine(k, m)
        ora = oracle.RS(k, m)
        for (S, ns, poly) in ((349526, 3, 0), (32768 * 2, 2, 1), (32768 * 3 + 1, 2, 0), (100001, 700, 1)):
            P = (S + 127) // 128 * 128
            rng = np.random.default_rng(S + k)
            host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
            cb.force_kernel(6)
            try:
                dev = torch.from_numpy(host).cuda()
                dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
                eng.dev_encode(dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), poly=poly)
                assert cb.last_kernel() == "rs_bs_kernel<crc,rolled>"
                torch.cuda.synchronize()
            finally:
                cb.force_kernel(0)
            out = dev.cpu().numpy()
            crc = dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)
            for s in list(range(min(ns, 3))) + [ns - 1]:
                sh = [host[s, i, :S].copy() for i in range(n)]
                ora.encode(sh)
                for i in range(n):
                    assert (out[s, i, :S] == sh[i]).all(), (k, m, S, s, i)
                    assert crc[s, i] == oracle.crc32(sh[i].tobytes(), poly), (k, m, S, s, i, poly)


def test_bitsliced_verify_kernel(cb):
    """rs_bs_kernel<verify> (reedSolomon.Verify, RS/reedsolomon.go:770-784): ok on encoded stripes;
    one flipped bit in any parity OR data shard -- first byte, last byte, middle -- fails exactly that
    stripe; bytes in the pitch padding beyond shard_len never matter.  Packed (small shard),
    whole-stripe and segmented geometries."""
    import torch
    for (k, m, S, ns) in ((12, 4, 349526, 6), (6, 3, 4096 + 7, 40), (4, 2, 2048, 700), (20, 4, 70001, 3), (12, 4, 33, 5)):
        n, P = k + m, (S + 127) // 128 * 128
        rng = np.random.default_rng(S + k)
        host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
        eng = cb.RSEngine(k, m)
        dev = torch.from_numpy(host).cuda()
        eng.dev_encode(dev.data_ptr(), S, P, n * P, ns)
        dok = torch.zeros(ns, dtype=torch.int32, device="cuda")
        eng.dev_verify(dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
        assert cb.last_kernel() == "rs_bs_kernel<verify>"
        torch.cuda.synchronize()
        assert bool((dok == 1).all()), (k, m, S)
        # garbage in the pitch padding of every shard: still ok
        if P > S:
            dev[:, :, S:] = 0xA5
            eng.dev_verify(dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
            torch.cuda.synchronize()
            assert bool((dok == 1).all()), (k, m, S, "pad")
        picks = [(0, 0, 0), (ns - 1, n - 1, S - 1), (ns // 2, k, S // 2), (ns // 3, k - 1, S - 1), (1 % ns, n - 1, 0)]
        for (s, shard, pos) in picks:
            dev[s, shard, pos] ^= 0x10
            eng.dev_verify(dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
            torch.cuda.synchronize()
            want = torch.ones(ns, dtype=torch.int32)
            want[s] = 0
            assert torch.equal(dok.cpu(), want), (k, m, S, s, shard, pos)
            dev[s, shard, pos] ^= 0x10


def test_warp_specialised_fused_kernel(cb, oracle):
    """Opt-in rs_bsw_kernel (coder warps + checksum warps, bitslice_ws.cu): same parity bytes and the
    same CRCs (both polynomials) as the oracle, on whole-stripe and segmented geometries and on
    shard lengths that end inside a 32-byte group / inside a tile."""
    import torch
    k, m = 12, 4
    n = k + m
    eng = cb.RSEngine(k, m)
    ora = oracle.RS(k, m)
    for (S, ns, poly) in ((349526, 3, 0), (24576 * 2, 2, 0), (24576 * 3 + 1, 2, 1), (33, 700, 0), (100001, 700, 1)):
        P = (S + 127) // 128 * 128
        rng = np.random.default_rng(S)
        host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
        cb.force_kernel(5)
        try:
            dev = torch.from_numpy(host).cuda()
            dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
            eng.dev_encode(dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), poly=poly)
            assert cb.last_kernel() == "rs_bsw_kernel"
            torch.cuda.synchronize()
        finally:
            cb.force_kernel(0)
        out = dev.cpu().numpy()
        crc = dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)
        for s in list(range(min(ns, 3))) + [ns - 1]:
            sh = [host[s, i, :S].copy() for i in range(n)]
            ora.encode(sh)
            for i in range(n):
                assert (out[s, i, :S] == sh[i]).all(), (S, s, i)
                assert crc[s, i] == oracle.crc32(sh[i].tobytes(), poly), (S, s, i, poly)


def test_kernel_selection_and_ab_equivalence(cb, oracle):
    """The bit-sliced kernel serves the specialised matrices on 32-byte-aligned layouts; the generic
    table kernel must produce the same bytes and CRCs (A/B through cubeec_debug_force_kernel)."""
    import torch
    # the last five are codes with m > 4 (EC15P12, EC6P6, EC16P20L2, EC12P9, EC24P8): ceil(m/4) passes of the
    # bit-sliced kernel, the first one also checksums the data shards; small S = packed mode
    for (k, m, S) in ((12, 4, 349526), (4, 2, 65536), (6, 3, 4096 + 7), (20, 4, 70001), (10, 4, 33),
                      (15, 12, 70001), (6, 6, 4096 + 7), (16, 20, 40000), (12, 9, 2048), (24, 8, 33000), (6, 10, 9000),
                      (18, 1, 5000), (4, 3, 66000)):
        n, P, ns = k + m, (S + 127) // 128 * 128, 5
        rng = np.random.default_rng(k * 7 + m)
        host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
        eng = cb.RSEngine(k, m)
        outs = []
        # default: the flat-split fused kernel for shards of 16 KiB and more, the tile-split one (packed mode) below;
        # 7 = tile-split kernel for every size; 1 = generic table kernel
        default = "rs_bsf_kernel<crc>" if S >= 16384 else "rs_bs_kernel<crc>"
        for force, want in ((0, default), (7, "rs_bs_kernel<crc>"), (1, "rs_tab_kernel<crc>")):
            cb.force_kernel(force)
            try:
                dev = torch.from_numpy(host).cuda()
                dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
                eng.dev_encode(dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr())
                assert cb.last_kernel() == want
                torch.cuda.synchronize()
                outs.append((dev.cpu().numpy(), dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)))
            finally:
                cb.force_kernel(0)
        for o in outs[1:]:
            assert (outs[0][0][:, :, :S] == o[0][:, :, :S]).all()
            assert (outs[0][1] == o[1]).all()
        ora = oracle.RS(k, m)
        for s in range(ns):
            sh = [host[s, i, :S].copy() for i in range(n)]
            ora.encode(sh)
            for i in range(n):
                assert (outs[0][0][s, i, :S] == sh[i]).all(), (k, m, s, i)
                assert outs[0][1][s, i] == zlib.crc32(sh[i].tobytes()), (k, m, s, i)
        # a 16-byte-aligned (not 32) layout must still work, through the table kernel
        dev = torch.from_numpy(host).cuda()
        flat = torch.zeros(dev.numel() + 64, dtype=torch.uint8, device="cuda")
        off = 16 - (flat.data_ptr() % 32) if flat.data_ptr() % 32 != 16 else 0
        view = flat[off:off + dev.numel()]
        assert view.data_ptr() % 32 == 16
        view.copy_(dev.reshape(-1))
        eng.dev_encode(view.data_ptr(), S, P, n * P, ns)
        assert cb.last_kernel() in ("rs_tabk_kernel", "rs_tab_kernel")   # fixed-arity variant when k has one
        torch.cuda.synchronize()
        got = view.cpu().numpy().reshape(ns, n, P)
        assert (got[:, :, :S] == outs[0][0][:, :, :S]).all()


@pytest.mark.parametrize("km", [(12, 4), (6, 3), (4, 2), (20, 4), (10, 4), (3, 3), (4, 4)])
def test_device_reconstruct_syndrome_kernel(cb, oracle, km):
    """cubeec_dev_reconstruct on specialised matrices runs the bit-sliced syndrome kernel; it must agree
    with the oracle for every erasure shape (data only, parity only, mixed, max), honour data_only, and
    agree with the generic table kernel even on INCONSISTENT survivors (same k shards are used)."""
    import torch
    k, m = km
    n = k + m
    rng = np.random.default_rng(k * 17 + m)
    S = int(rng.integers(3000, 90000))
    P = (S + 127) // 128 * 128
    shapes = [[0], [k], [k - 1, n - 1], list(range(m)), list(range(k, n)), [1, k] + ([k + 1] if m > 2 else []), []]
    shapes += [sorted(rng.choice(n, size=int(rng.integers(1, m + 1)), replace=False).tolist()) for _ in range(12)]
    ns = len(shapes)
    host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
    ora = oracle.RS(k, m)
    for s in range(ns):
        sh = [host[s, i, :S].copy() for i in range(n)]
        ora.encode(sh)
        for i in range(k, n):
            host[s, i, :S] = sh[i]
    eng = cb.RSEngine(k, m)
    present = np.ones((ns, n), dtype=np.uint8)
    for s, miss in enumerate(shapes):
        present[s, miss] = 0
    for data_only in (False, True):
        broken = host.copy()
        for s, miss in enumerate(shapes):
            broken[s, miss, :] = 0xA5
        dev = torch.from_numpy(broken).cuda()
        cb.force_kernel(2)
        try:
            eng.dev_reconstruct(dev.data_ptr(), S, P, n * P, ns, present, data_only=data_only)
        finally:
            cb.force_kernel(0)
        assert cb.last_kernel() == "rs_bsrec_kernel"
        torch.cuda.synchronize()
        got = dev.cpu().numpy()
        for s, miss in enumerate(shapes):
            for i in range(n):
                if data_only and i >= k and i in miss:
                    assert (got[s, i, :S] == 0xA5).all()          # missing parity stays missing
                else:
                    assert (got[s, i, :S] == host[s, i, :S]).all(), (km, s, i, miss, data_only)
    # inconsistent survivors: corrupt one present shard per stripe, both kernels must produce the same bytes
    bad = host.copy()
    for s, miss in enumerate(shapes):
        bad[s, miss, :] = 0
        keep = [i for i in range(n) if i not in miss]
        bad[s, keep[s % len(keep)], 7] ^= 0x5A
    outs = []
    for force in (2, 0, 3):   # bit-sliced syndrome kernel, fixed-arity table kernel, generic table kernel
        cb.force_kernel(force)
        try:
            dev = torch.from_numpy(bad).cuda()
            eng.dev_reconstruct(dev.data_ptr(), S, P, n * P, ns, present)
            torch.cuda.synchronize()
            outs.append(dev.cpu().numpy())
        finally:
            cb.force_kernel(0)
    assert (outs[0][:, :, :S] == outs[1][:, :, :S]).all() and (outs[1][:, :, :S] == outs[2][:, :, :S]).all()
    # too few shards
    present2 = np.ones((1, n), dtype=np.uint8)
    present2[0, :m + 1] = 0
    dev = torch.from_numpy(host[:1].copy()).cuda()
    with pytest.raises(cb.CubeecError) as e:
        eng.dev_reconstruct(dev.data_ptr(), S, P, n * P, 1, present2)
    assert e.value.name == "ErrTooFewShards"


@pytest.mark.parametrize("km", [(12, 4), (6, 3), (4, 2)])
def test_small_shards_packed_mode(cb, oracle, km):
    """Shards shorter than a tile (2 KiB is the production MinShardSize): the bit-sliced kernel packs the
    pieces of many stripes into one tile.  Parity and CRCs must still match the oracle for every size,
    including sizes that are not multiples of 64 and batches whose last tile is partly empty."""
    import torch
    k, m = km
    n = k + m
    rng = np.random.default_rng(k + 100 * m)
    for S, ns in ((2048, 300), (4096, 77), (2048 + 6, 41), (511, 9), (16384, 33), (32767, 5), (20000, 130), (640, 1000)):
        P = (S + 127) // 128 * 128
        host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
        dev = torch.from_numpy(host).cuda()
        dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
        eng = cb.RSEngine(k, m)
        eng.dev_encode(dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr())
        assert cb.last_kernel() == ("rs_bsf_kernel<crc>" if S >= 16384 else "rs_bs_kernel<crc>")   # packed mode below 16 KiB
        torch.cuda.synchronize()
        out = dev.cpu().numpy()
        crc = dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)
        dev2 = torch.from_numpy(host).cuda()
        eng.dev_encode(dev2.data_ptr(), S, P, n * P, ns)          # without CRC
        torch.cuda.synchronize()
        assert torch.equal(dev2[:, :, :S].cpu(), torch.from_numpy(out[:, :, :S]))
        ora = oracle.RS(k, m)
        for s in list(range(min(ns, 6))) + [ns // 2, ns - 1]:
            sh = [host[s, i, :S].copy() for i in range(n)]
            ora.encode(sh)
            for i in range(n):
                assert (out[s, i, :S] == sh[i]).all(), (km, S, s, i)
                assert crc[s, i] == zlib.crc32(sh[i].tobytes()), (km, S, s, i)
        # all stripes: parity by linear algebra on the whole batch (oracle SIMD path) and CRC of every shard
        want = host[:, :, :S].copy()
        ora.encode_batch_simd(want, S, S, n * S, ns)
        assert (out[:, :, :S] == want).all(), (km, S)
        assert all(crc[s, i] == zlib.crc32(want[s, i].tobytes()) for s in range(ns) for i in range(n)), (km, S)


def test_concurrent_callers(cb, oracle):
    """ec.Encoder methods are called from many goroutines at once (encoder.go:115, Concurrency up to 1000):
    the C-ABI must be thread-safe.  16 threads x mixed encode / verify / reconstruct / crc32 on one handle."""
    import threading
    k, m = 12, 4
    eng = cb.RSEngine(k, m)
    ora = oracle.RS(k, m)
    errors = []

    def worker(seed):
        try:
            rng = np.random.default_rng(seed)
            for it in range(6):
                S = int(rng.integers(1, 200000))
                sh = [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(k)] + [np.zeros(S, np.uint8) for _ in range(m)]
                want = [x.copy() for x in sh]
                ora.encode(want)
                crc = eng.encode(sh, crc=True)
                assert all((a == b).all() for a, b in zip(sh, want))
                assert [int(c) for c in crc] == [zlib.crc32(w.tobytes()) for w in want]
                assert eng.verify(sh)
                miss = rng.choice(k + m, size=int(rng.integers(1, m + 1)), replace=False)
                out = eng.reconstruct([None if i in miss else want[i] for i in range(k + m)])
                assert all((a == b).all() for a, b in zip(out, want))
                assert cb.crc32(want[0]) == zlib.crc32(want[0].tobytes())
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(16)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors[:3]
