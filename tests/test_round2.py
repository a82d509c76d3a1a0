"""Round-2 GPU tests: the coalescing submit queue under concurrent callers, the ADVICE r1 fixes (replica handles
with CRC, large batches of small shards through cubeec_encode_contig), big-batch and C4-size oracle checks."""
import threading
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_shards(rng, k, m, S):
    return [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(k)] + [np.zeros(S, np.uint8) for _ in range(m)]


@pytest.mark.parametrize("coalesce", [(32, 100), (8, 2000), (1, 0)])
def test_64_concurrent_single_stripe_callers(cb, oracle, coalesce):
    """64 threads, each calling cubeec_encode on its own stripes (the access call shape, encoder.go:114-131):
    parity and all CRCs must equal the oracle whatever batches the queue forms; mixed shard sizes and two code
    modes interleave, so batches of different keys are open at the same time."""
    cb.set_coalescing(*coalesce)
    try:
        engs = {(12, 4): cb.RSEngine(12, 4), (6, 3): cb.RSEngine(6, 3)}
        errors = []

        def worker(i):
            try:
                rng = np.random.default_rng(1000 + i)
                for it in range(6):
                    k, m = (12, 4) if (i + it) % 3 else (6, 3)
                    S = [349526, 65536 + 6, 2048, 21846][(i + it) % 4]
                    sh = _rand_shards(rng, k, m, S)
                    want = [s.copy() for s in sh]
                    oracle.RS(k, m).encode(want)
                    crc = engs[(k, m)].encode(sh, crc=(it % 2 == 0))
                    for j in range(k + m):
                        assert (sh[j] == want[j]).all(), (i, it, j)
                        if crc is not None:
                            assert int(crc[j]) == zlib.crc32(want[j].tobytes()), (i, it, j)
            except Exception as e:   # noqa: BLE001
                errors.append(repr(e))

        th = [threading.Thread(target=worker, args=(i,)) for i in range(64)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errors, errors[:3]
    finally:
        cb.set_coalescing(32, 100)


def test_replica_handle_crc(cb):
    """reedsolomon.New(N, 0) handles (the Replica code modes) with crc_out used to return garbage."""
    rng = np.random.default_rng(5)
    for S in (2048, 70001):
        sh = [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(3)]
        crc = cb.RSEngine(3, 0).encode(sh, crc=True)
        assert [int(c) for c in crc] == [zlib.crc32(s.tobytes()) for s in sh]
        buf = np.concatenate(sh)
        crc2, _ = cb.RSEngine(3, 0).encode_contig(buf, S, 1, 3 * S, crc=True)
        assert [int(c) for c in crc2[0]] == [zlib.crc32(s.tobytes()) for s in sh]


def test_contig_many_small_stripes_with_crc(cb, oracle):
    """cubeec_encode_contig + crc_out on a large batch of small shards (packed mode, chunk > 8*SMs)."""
    k, m, S, ns = 12, 4, 2048, 6000
    n = k + m
    rng = np.random.default_rng(11)
    buf = rng.integers(0, 256, (ns, n * S), dtype=np.uint8)
    buf[:, k * S:] = 0
    crc, _ = cb.RSEngine(k, m).encode_contig(buf, S, ns, n * S, crc=True)
    ora = oracle.RS(k, m)
    for s in (0, 1, 2999, ns - 1):
        sh = [buf[s, i * S:(i + 1) * S].copy() for i in range(n)]
        want = [x.copy() for x in sh[:k]] + [np.zeros(S, np.uint8) for _ in range(m)]
        ora.encode(want)
        for i in range(n):
            assert (sh[i] == want[i]).all(), (s, i)
            assert int(crc[s, i]) == zlib.crc32(want[i].tobytes()), (s, i)


def test_c4_size_vs_oracle(cb, oracle):
    """BASELINE C4 at its real size: RS(20,4), 1 MiB shards, fused CRC, vs the oracle."""
    k, m, S = 20, 4, 1 << 20
    rng = np.random.default_rng(20)
    sh = _rand_shards(rng, k, m, S)
    want = [s.copy() for s in sh]
    oracle.RS(k, m).encode(want)
    crc = cb.RSEngine(k, m).encode(sh, crc=True)
    for i in range(k + m):
        assert (sh[i] == want[i]).all(), i
        assert int(crc[i]) == zlib.crc32(want[i].tobytes()), i


@pytest.mark.parametrize("ns", [1024, 383, 149])
def test_c2_big_batch_vs_oracle_simd(cb, oracle, ns):
    """Device-resident C2 batches (the bench shape and the awkward stripe counts) against encode_batch_simd of the
    oracle: every parity byte and every CRC of every stripe."""
    import torch
    k, m, S = 12, 4, 349526
    n = k + m
    P = (S + 127) // 128 * 128
    g = torch.Generator(device="cuda").manual_seed(ns)
    dev = torch.randint(0, 256, (ns, n, P), dtype=torch.uint8, device="cuda", generator=g)
    dcrc = torch.zeros(ns * n, dtype=torch.int32, device="cuda")
    eng = cb.RSEngine(k, m)
    eng.dev_encode(dev.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr())
    torch.cuda.synchronize()
    host = dev.cpu().numpy()
    crc = dcrc.cpu().numpy().view(np.uint32).reshape(ns, n)
    ref = host.copy()
    ref[:, k:, :] = 0
    rcrc = np.zeros((ns, n), dtype=np.uint32)
    oracle.RS(k, m).encode_batch_simd(ref, S, P, n * P, ns, threads=8, crc_out=rcrc)
    assert np.array_equal(host[:, :, :S], ref[:, :, :S])
    assert np.array_equal(crc, rcrc)
    # spot-check the oracle's own CRCs against zlib
    for s in (0, ns - 1):
        assert int(rcrc[s, 0]) == zlib.crc32(host[s, 0, :S].tobytes())


@pytest.mark.parametrize("case", [(12, 4, [3], 349526, 40), (12, 4, [0, 5, 13], 70001, 25), (6, 3, [8], 4096 + 7, 300),
                                  (20, 4, [1, 2, 21, 23], 65536, 9), (6, 10, [0, 1, 2, 3, 4, 7, 9], 20000, 12), (4, 2, [0, 1], 33, 50)])
def test_single_pattern_reconstruct_runtime_compiled_kernel(cb, oracle, case):
    """One erasure pattern for the whole batch (a repair task): cubeec_dev_reconstruct runs the NVRTC-specialised kernel
    (csrc/jit.cu).  Bit-exact against the oracle's encode, and identical to the table kernels (force 4)."""
    import torch
    k, m, miss, S, ns = case
    n = k + m
    P = (S + 127) // 128 * 128
    rng = np.random.default_rng(k * 100 + m + S)
    host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
    ora = oracle.RS(k, m)
    for s in range(ns):
        sh = [host[s, i, :S] for i in range(n)]
        ora.encode(sh)
    present = np.ones((ns, n), dtype=np.uint8)
    present[:, miss] = 0
    eng = cb.RSEngine(k, m)
    outs = []
    for force in (0, 4):
        cb.force_kernel(force)
        try:
            broken = host.copy()
            broken[:, miss, :] = 0x5A
            dev = torch.from_numpy(broken).cuda()
            eng.dev_reconstruct(dev.data_ptr(), S, P, n * P, ns, present)
            name = cb.last_kernel()
            torch.cuda.synchronize()
            outs.append(dev.cpu().numpy())
        finally:
            cb.force_kernel(0)
        if force == 0:
            assert name == "rs_jit_kernel", name
        else:
            assert name in ("rs_tabk_kernel", "rs_tab_kernel"), name
    assert np.array_equal(outs[0][:, :, :S], host[:, :, :S])
    assert np.array_equal(outs[1][:, :, :S], host[:, :, :S])
    # the survivors were not touched, bytes of the regenerated shards beyond S stay inside the 32-byte column rule
    for i in range(n):
        if i not in miss:
            assert np.array_equal(outs[0][:, i, :], host[:, i, :])


@pytest.mark.parametrize("km", [(12, 4), (6, 3), (4, 2), (20, 4), (10, 4), (3, 3), (8, 4)])
def test_mixed_pattern_reconstruct_syndrome_kernel_v2(cb, oracle, km):
    """Batches that mix erasure patterns (config C3) through the opt-in flat-split syndrome kernel (bitslice_syn.cu, force 9): every count of
    missing data / parity shards up to m, ragged and multi-unit sizes, warps whose runs cross stripe (= pattern)
    boundaries.  Bit-exact against the originals and against the table kernels (force 8); data_only leaves parity alone."""
    import torch
    k, m = km
    n = k + m
    ora = oracle.RS(k, m)
    eng = cb.RSEngine(k, m)
    rng = np.random.default_rng(k * 31 + m)
    for S, ns in ((349526 if k == 12 else 70001, 37), (1024 * 3, 300), (33, 64), (2048 + 6, 200)):
        P = (S + 127) // 128 * 128
        host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
        for s in range(ns):
            ora.encode([host[s, i, :S] for i in range(n)])
        present = np.ones((ns, n), dtype=np.uint8)
        for s in range(ns):
            e = int(rng.integers(0, m + 1))                    # 0 .. m erasures, anywhere
            present[s, rng.choice(n, size=e, replace=False)] = 0
        for data_only in (False, True):
            outs = []
            for force in (9, 8):
                cb.force_kernel(force)
                try:
                    broken = host.copy()
                    for s in range(ns):
                        broken[s, present[s] == 0, :] = 0xC3
                    dev = torch.from_numpy(broken).cuda()
                    eng.dev_reconstruct(dev.data_ptr(), S, P, n * P, ns, present, data_only=data_only)
                    name = cb.last_kernel()
                    torch.cuda.synchronize()
                    outs.append(dev.cpu().numpy())
                finally:
                    cb.force_kernel(0)
                assert name == ("rs_bssyn_kernel" if force == 9 else name), name
                if force == 8:
                    assert name in ("rs_tabk_kernel", "rs_tab_kernel"), name
            for s in range(ns):
                for i in range(n):
                    regenerated = present[s, i] == 0 and not (data_only and i >= k)
                    want = host[s, i, :S] if (present[s, i] or regenerated) else None
                    if want is not None:
                        assert np.array_equal(outs[0][s, i, :S], want), (km, S, s, i, data_only)
                        assert np.array_equal(outs[1][s, i, :S], want), (km, S, s, i, data_only)


@pytest.mark.parametrize("mode", [(6, 10, 2, 2), (16, 20, 2, 2), (6, 3, 3, 3), (4, 4, 2, 2)])
def test_lrc_verify_and_reconstruct_on_device(cb, oracle, mode):
    """lrcEncoder.Verify / Reconstruct / ReconstructData (lrcencoder.go:87-200) on device-resident LRC stripes: global
    code + per-AZ local codes through slot maps, against the oracle's composition of the same codes."""
    import torch
    N, M, L, az = mode
    n, ng = N + M + L, N + M
    kl, ml = ng // az, L // az
    ge, le = cb.RSEngine(N, M), cb.RSEngine(kl, ml)
    S, ns = 20000 + 6, 40
    P = (S + 127) // 128 * 128
    rng = np.random.default_rng(N * 100 + M)
    host = rng.integers(0, 256, (ns, n, P), dtype=np.uint8)
    og, ol = oracle.RS(N, M), oracle.RS(kl, ml)
    for s in range(ns):
        og.encode([host[s, i, :S] for i in range(ng)])
        for a in range(az):
            members = list(range(a * N // az, (a + 1) * N // az)) + list(range(N + a * M // az, N + (a + 1) * M // az))
            ol.encode([host[s, i, :S] for i in members] + [host[s, ng + a * ml + i, :S] for i in range(ml)])
    dev = torch.from_numpy(host).cuda()
    dok = torch.zeros(ns, dtype=torch.int32, device="cuda")
    # the device encode produces the same stripes (fused LRC path)
    enc = torch.from_numpy(host).cuda()
    enc[:, N:, :] = 0
    cb.dev_lrc_encode(ge, le, az, enc.data_ptr(), S, P, n * P, ns)
    torch.cuda.synchronize()
    assert torch.equal(enc[:, :, :S], dev[:, :, :S])
    cb.dev_lrc_verify(ge, le, az, dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
    assert dok.cpu().numpy().tolist() == [1] * ns
    # one flipped bit: in a global parity shard of stripe 3, in a local parity shard of stripe 7, in data of stripe 11
    dev[3, N + 1, 5] ^= 1
    dev[7, ng + L - 1, S - 1] ^= 0x80
    dev[11, 0, 100] ^= 4
    cb.dev_lrc_verify(ge, le, az, dev.data_ptr(), S, P, n * P, ns, dok.data_ptr())
    want = [1] * ns
    want[3] = want[7] = want[11] = 0
    assert dok.cpu().numpy().tolist() == want
    # reconstruct: per stripe a random set of missing global shards (<= M) and missing local parity
    present = np.ones((ns, n), dtype=np.uint8)
    for s in range(ns):
        e = int(rng.integers(0, min(M, 4) + 1))
        present[s, rng.choice(ng, size=e, replace=False)] = 0
        if s % 3 == 0:
            present[s, ng + int(rng.integers(0, L))] = 0
    for data_only in (False, True):
        broken = host.copy()
        for s in range(ns):
            broken[s, present[s] == 0, :] = 0x3C
        dev = torch.from_numpy(broken).cuda()
        cb.dev_lrc_reconstruct(ge, le, az, dev.data_ptr(), S, P, n * P, ns, present, data_only=data_only)
        torch.cuda.synchronize()
        out = dev.cpu().numpy()
        for s in range(ns):
            for i in range(n):
                if present[s, i] or (not data_only) or i < N:
                    assert np.array_equal(out[s, i, :S], host[s, i, :S]), (mode, s, i, data_only)
                else:
                    assert np.array_equal(out[s, i, :S], broken[s, i, :S]), (mode, s, i, "left missing")


@pytest.mark.parametrize("coalesce", [(32, 100), (1, 0)])
def test_concurrent_degraded_read_callers(cb, oracle, coalesce):
    """48 threads calling cubeec_reconstruct (ReconstructData and Reconstruct, stream_get.go:454-465) at once: the queue
    batches calls of the same code / size / flavour; most share one erasure pattern (a dead node), some do not; a few ask
    for checksums (direct path).  Every regenerated shard must equal the original, `filled` semantics included."""
    cb.set_coalescing(*coalesce)
    try:
        k, m = 12, 4
        n = k + m
        eng = cb.RSEngine(k, m)
        ora = oracle.RS(k, m)
        errors = []

        def worker(i):
            try:
                rng = np.random.default_rng(7000 + i)
                for it in range(5):
                    S = [349526, 21846, 2048 + 6][(i + it) % 3]
                    sh = [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(k)] + [np.zeros(S, np.uint8) for _ in range(m)]
                    ora.encode(sh)
                    miss = [3] if (i + it) % 4 else sorted(rng.choice(n, size=int(rng.integers(1, m + 1)), replace=False).tolist())
                    data_only = (it % 2 == 1)
                    broken = [None if j in miss else sh[j].copy() for j in range(n)]
                    if it == 4 and i % 5 == 0:
                        out, crc = eng.reconstruct(broken, data_only=data_only, crc=True)
                    else:
                        out, crc = eng.reconstruct(broken, data_only=data_only), None
                    for j in range(n):
                        if j in miss and data_only and j >= k:
                            assert out[j] is None, (i, it, j, "parity must stay missing with ReconstructData")
                        else:
                            assert out[j] is not None and (out[j] == sh[j]).all(), (i, it, j)
                            if crc is not None and j in miss:
                                assert int(crc[j]) == zlib.crc32(sh[j].tobytes()), (i, it, j)
            except Exception as e:   # noqa: BLE001
                errors.append(repr(e))

        th = [threading.Thread(target=worker, args=(i,)) for i in range(48)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errors, errors[:3]
        # too few shards is still reported before anything is queued
        sh = [None] * 5 + [np.zeros(64, np.uint8)] * 11
        with pytest.raises(cb.CubeecError) as e:
            eng.reconstruct(sh)
        assert e.value.name == "ErrTooFewShards"
    finally:
        cb.set_coalescing(32, 100)
