"""The host-side mirror of blobstore/common/{codemode,ec,crc32block} above the C-ABI.

CPU part: codemode / buffer-size arithmetic (codemode_test.go, buf_test.go).
GPU part: a replay of blobstore/common/ec/encoder_test.go -- TestEncoderNew, TestEncoder (EC15P12),
TestLrcEncoder (EC6P10L2), TestLrcReconstruct (every EC code mode) -- and crc32block round trips."""
import zlib

import numpy as np
import pytest

from mirror import ec as cm   # codemode + ec + crc32block mirror

SRC = bytes(range(256)) * 3 + b"cubefs-blobstore-ec-src-data" * 5   # encoder_test.go's srcData stand-in

EC6P10L2_STRIPES = [[0, 1, 2, 6, 7, 8, 9, 10, 16], [3, 4, 5, 11, 12, 13, 14, 15, 17]]
EC16P20L2_STRIPES = [[0, 1, 2, 3, 4, 5, 6, 7, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 36],
                     [8, 9, 10, 11, 12, 13, 14, 15, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 37]]


# ---------------------------------------------------------------- CPU: codemode_test.go / buf_test.go
def test_codemode_base():
    for mode in (cm.EC15P12, cm.EC6P10L2, cm.EC10P4):
        t = mode.Tactic()
        assert t.MinShardSize == 2048 and t.IsValid()
        assert t.N + t.M + t.L == mode.GetShardNum()
        idx, n, m = t.GlobalStripe()
        assert (n, m) == (t.N, t.M) and idx == list(range(t.N + t.M))
    assert len(cm.GetAllCodeModes()) == 20
    for mode in cm.GetAllCodeModes():
        assert mode.IsValid() and mode.Tactic().IsValid() and mode.Name() == mode.String()
    for mode in cm.GetECCodeModes():
        assert not mode.Tactic().IsReplicateMode()
    for bad in (0, 255, 127):
        assert not cm.CodeMode(bad).IsValid() and cm.CodeMode(bad).String() == ""
        with pytest.raises(ValueError):
            cm.CodeMode(bad).Tactic()


def test_layout_by_az_and_local_stripes():
    idx = cm.EC15P12.Tactic().GetECLayoutByAZ()
    assert len(idx) == 3 and all(len(r) == 9 for r in idx)
    assert cm.EC6P10L2.Tactic().GetECLayoutByAZ() == EC6P10L2_STRIPES
    assert cm.EC16P20L2.Tactic().GetECLayoutByAZ() == EC16P20L2_STRIPES
    t = cm.EC12P4.Tactic()
    assert [len(r) for r in t.GetECLayoutByAZ()] == [16]
    assert t.LocalStripe(3) == (None, 0, 0) and t.AllLocalStripe() == (None, 0, 0)
    t = cm.EC6P10L2.Tactic()
    assert t.AllLocalStripe() == (EC6P10L2_STRIPES, 8, 1)
    for i in (0, 1, 2, 6, 10, 16):
        assert t.LocalStripe(i) == (EC6P10L2_STRIPES[0], 8, 1)
    for i in (3, 5, 11, 15, 17):
        assert t.LocalStripe(i) == (EC6P10L2_STRIPES[1], 8, 1)
    assert t.LocalStripe(18) == (None, 0, 0)
    assert t.LocalStripeInAZ(1) == (EC6P10L2_STRIPES[1], 8, 1) and t.LocalStripeInAZ(2) == (None, 0, 0)
    assert not cm.Tactic().IsValid() and not cm.Tactic(N=6, M=3, L=0, AZCount=2, PutQuorum=8).IsValid()


def test_get_buffer_sizes():
    kb, kb512 = 1 << 10, 1 << 19
    for mode, size in ((cm.EC6P6, kb), (cm.EC16P20L2, kb512), (cm.EC12P4, 4 << 20)):
        t = mode.Tactic()
        s = cm.GetBufferSizes(size, t)
        shard = max((size + t.N - 1) // t.N, t.MinShardSize)
        assert (s.ShardSize, s.DataSize, s.ECDataSize, s.ECSize) == (shard, size, shard * t.N, shard * (t.N + t.M + t.L))
    assert cm.GetBufferSizes(4 << 20, cm.EC12P4.Tactic()).ShardSize == 349526
    for bad in (0, -1):
        with pytest.raises(cm.EcError) as e:
            cm.GetBufferSizes(bad, cm.EC6P6.Tactic())
        assert e.value.name == "ErrShortData"


def test_crc32block_sizes():
    assert cm.EncodeSize(65532) == 65536 and cm.EncodeSize(65533) == 65541 and cm.DecodeSize(65536) == 65532
    assert cm.EncodeSize(10, 1000) == -1   # invalid block length (the Go code panics)


# ---------------------------------------------------------------- GPU: encoder_test.go
def _data_slice(data: bytes, cap=None):
    buf = np.zeros(cap or len(data), dtype=np.uint8)
    buf[:len(data)] = np.frombuffer(data, dtype=np.uint8)
    return cm.GoSlice(buf, len(data))


def _ec_buffer(data: bytes, t: "cm.Tactic"):
    """ec.NewBuffer: one buffer of ECSize, data at the front (buf.go:23-35)."""
    sizes = cm.GetBufferSizes(len(data), t)
    return _data_slice(data, sizes.ECSize), sizes


@pytest.mark.gpu
def test_encoder_new():
    with pytest.raises(cm.EcError) as e:
        cm.NewEncoder(cm.Config(CodeMode=cm.Tactic()))
    assert e.value.name == "ErrInvalidCodeMode"
    cm.NewEncoder(cm.Config(CodeMode=cm.EC15P12.Tactic()))
    cm.NewEncoder(cm.Config(CodeMode=cm.EC16P20L2.Tactic()))


@pytest.mark.gpu
def test_encoder_ec15p12():
    cfg = cm.Config(CodeMode=cm.EC15P12.Tactic(), EnableVerify=True, Concurrency=10)
    enc = cm.NewEncoder(cfg)
    shards = enc.Split(_data_slice(SRC))
    assert len(shards) == 27
    enc.Encode(shards)
    assert enc.Join(shards, len(SRC)) == SRC
    enc.GetDataShards(shards)[0].data[:] = 222
    enc.ReconstructData(shards, [0])
    assert enc.Join(shards, len(SRC)) == SRC
    enc.GetParityShards(shards)[1].data[:] = 11
    enc.Reconstruct(shards, [cfg.CodeMode.N + 1])
    assert enc.Verify(shards)
    assert enc.Join(shards, len(SRC)) == SRC
    assert len(enc.GetLocalShards(shards)) == 0
    assert len(enc.GetShardsInIdc(shards, 0)) == (cfg.CodeMode.N + cfg.CodeMode.M) // 3


@pytest.mark.gpu
def test_lrc_encoder_ec6p10l2():
    cfg = cm.Config(CodeMode=cm.EC6P10L2.Tactic(), EnableVerify=True)
    t = cfg.CodeMode
    enc = cm.NewEncoder(cfg)
    with pytest.raises(cm.EcError):
        enc.Split(cm.GoSlice(np.zeros(0, np.uint8)))
    shards = enc.Split(_data_slice(SRC))
    assert len(shards) == 18
    enc.Split(_data_slice(SRC, cap=1 << 10))
    with pytest.raises(cm.EcError) as e:
        enc.Encode(shards[:-1])
    assert e.value.name == "ErrInvalidShards"
    with pytest.raises(cm.EcError) as e:
        enc.Encode(None)
    assert e.value.name == "ErrInvalidShards"
    enc.Encode(shards)
    assert enc.Join(shards, len(SRC)) == SRC
    enc.GetDataShards(shards)[0].data[:] = 222
    assert not enc.Verify(shards)
    enc.ReconstructData(shards, [0])
    assert enc.Join(shards, len(SRC)) == SRC
    # local reconstruct inside AZ 0, one shard at a time
    local = enc.GetShardsInIdc(shards, 0)
    assert len(local) == 9
    for idx in range(len(local)):
        local[idx].data[:] = 11
        assert not enc.Verify(shards)
        enc.Reconstruct(local, [idx])
        assert enc.Verify(shards)
    bad = [t.N + t.M + 1]
    shards[t.N + t.M + 1].data[:] = 222
    assert not enc.Verify(shards)
    data, parity = enc.GetDataShards(shards), enc.GetParityShards(shards)
    for i in range(t.M):
        if i % 2 == 0:
            bad.append(i)
            if i < len(data):
                data[i].data[:] = 222
        else:
            bad.append(t.N + i)
            parity[i].data[:] = 222
    assert not enc.Verify(shards)
    enc.Reconstruct(shards, bad)
    assert enc.Verify(shards)
    assert enc.Join(shards, len(SRC)) == SRC
    assert len(enc.GetLocalShards(shards)) == t.L
    assert len(enc.GetShardsInIdc(shards, 0)) == (t.N + t.M + t.L) // t.AZCount
    # length errors: Verify reports an error (not just false) when a shard is empty
    shards[bad[0]].truncate()
    with pytest.raises(cm.EcError):
        enc.Verify(shards)
    enc.Reconstruct(shards, bad)
    assert enc.Verify(shards)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [m for m in range(1, 16)] + [200, 201])
def test_lrc_reconstruct_all_modes(mode):
    """testLrcReconstruct (encoder_test.go:249-307) for every EC code mode."""
    mode = cm.CodeMode(mode)
    t = mode.Tactic()
    enc = cm.NewEncoder(cm.Config(CodeMode=t, EnableVerify=True))
    rng = np.random.default_rng(int(mode))
    data = rng.integers(0, 256, (1 << 16) + int(rng.integers(0, 1 << 16)), dtype=np.uint8).tobytes()
    buf, sizes = _ec_buffer(data, t)
    ecdata = cm.GoSlice(buf.buf, sizes.ECDataSize)      # ECDataBuf: Split fills parity/local from spare capacity
    shards = enc.Split(ecdata)
    assert len(shards) == mode.GetShardNum()
    enc.Encode(shards)
    origin = [s.data.copy() for s in shards]
    bads = []
    for bad in range(t.N + t.M, mode.GetShardNum()):
        bads.append(bad)
        for i in bads:
            shards[i].data[:] = 0
            shards[i].truncate()
        enc.Reconstruct(shards, bads)
        assert all((s.data == o).all() for s, o in zip(shards, origin))
    with pytest.raises(cm.EcError):
        enc.Reconstruct([s.copy() for s in shards], bads + list(range(t.N + t.M)))
    for az in range(t.AZCount if t.L else 0):
        locals_, n, m = t.LocalStripeInAZ(az)
        local = [shards[i] for i in locals_]
        local_origin = [s.data.copy() for s in local]
        bads = []
        for bad in range(n, n + m):
            bads.append(bad)
            for i in bads:
                local[i].data[:] = 0
                local[i].truncate()
            enc.Reconstruct(local, bads)
            assert all((s.data == o).all() for s, o in zip(local, local_origin))
        if n > 0:
            bads.append(n - 1)
            for i in bads:
                local[i].truncate()
            with pytest.raises(cm.EcError):
                enc.Reconstruct(local, bads)
            for s, o in zip(local, local_origin):   # restore for the next AZ
                s.len = len(o)
                s.data[:] = o


@pytest.mark.gpu
def test_lrc_parity_matches_oracle(oracle):
    """Global parity = RS(N,M); each local parity = RS((N+M)/AZ, L/AZ) over the AZ's shards (lrcencoder.go:35-80)."""
    t = cm.EC6P10L2.Tactic()
    enc = cm.NewEncoder(cm.Config(CodeMode=t))
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, 50000, dtype=np.uint8).tobytes()
    buf, sizes = _ec_buffer(data, t)
    shards = enc.Split(cm.GoSlice(buf.buf, sizes.ECDataSize))
    enc.Encode(shards)
    glob = [s.data.copy() for s in shards[:t.N + t.M]]
    want = [g.copy() for g in glob]
    oracle.RS(t.N, t.M).encode(want)
    assert all((a == b).all() for a, b in zip(glob, want))
    for az, stripe in enumerate(EC6P10L2_STRIPES):
        loc = [shards[i].data.copy() for i in stripe]
        want = [x.copy() for x in loc]
        oracle.RS(8, 1).encode(want)
        assert (loc[8] == want[8]).all(), az


@pytest.mark.gpu
def test_crc32block_round_trip(oracle):
    rng = np.random.default_rng(4)
    for n in (1, 9, 65531, 65532, 65533, 131064, 300000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        framed, whole = cm.BlockEncode(d)
        assert framed == oracle.crc32block_encode(d) and whole == zlib.crc32(d)
        assert len(framed) == cm.EncodeSize(n)
        assert cm.BlockDecode(framed) == d
        bad = bytearray(framed)
        bad[len(bad) // 2] ^= 4
        with pytest.raises(cm.EcError) as e:
            cm.BlockDecode(bytes(bad))
        assert e.value.name == "ErrMismatchedCrc"


def test_shard_image_layout_cpu(oracle):
    """datafile_test.go:175-260: a 9-byte shard occupies 32+(9+4)+8 bytes, page-aligned to 4096; 32 KiB -> 4096+32768."""
    assert cm.Alignphysize(9) == 53 and cm.AlignSize(4096 + 53, 4096) == 8192
    assert cm.AlignSize(65536 + cm.Alignphysize(32768), 4096) == 65536 + 4096 + 32768
    img, crc = oracle.shard_image(1024, 10, b"test data")
    assert crc == 3540561586 and len(img) == 53
    assert img[4:8] == bytes([0xab, 0xcd, 0xef, 0xcc]) and img[-8:-4] == bytes([0xcc, 0xef, 0xcd, 0xab])
    assert int.from_bytes(img[8:16], "big") == 1024 and int.from_bytes(img[16:24], "big") == 10
    assert int.from_bytes(img[24:28], "big") == 9 and int.from_bytes(img[-4:], "big") == 3540561586
    assert int.from_bytes(img[:4], "big") == zlib.crc32(img[4:32])


@pytest.mark.gpu
def test_shard_write_read_matches_oracle(oracle):
    """blobnode shard write (datafile.Write) with GPU CRCs == the oracle's restatement, byte for byte;
    read-back verifies; the reference's golden CRCs hold (datafile_test.go:198-396)."""
    rng = np.random.default_rng(6)
    cases = [b"test data", bytes(ord("0") + i % 10 for i in range(32768))]
    cases += [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (1, 65531, 65532, 65533, 349526, 1 << 20)]
    golden = {0: 3540561586, 1: 629387998}
    for i, data in enumerate(cases):
        img, crc = cm.WriteShard(1024 + i, 10, data)
        want_img, want_crc = oracle.shard_image(1024 + i, 10, data)
        assert img == want_img and crc == want_crc == zlib.crc32(data)
        if i in golden:
            assert crc == golden[i]
        bid, vuid, rcrc, back = cm.ReadShard(img)
        assert (bid, vuid, rcrc, back) == (1024 + i, 10, crc, data)
        for pos, err in ((5, "ErrShardHeaderMagic"), (20, "ErrShardHeaderCrc"), (40, "ErrMismatchedCrc"), (len(img) - 1, "ErrShardCrc")):
            if pos >= len(img) - 8 and pos < len(img) - 4:
                continue
            bad = bytearray(img)
            bad[pos] ^= 0x10
            with pytest.raises(cm.EcError) as e:
                cm.ReadShard(bytes(bad))
            assert e.value.name == err, (i, pos)
