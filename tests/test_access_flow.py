"""Integration replay of the callers either side of the hot path:

PUT  (access/stream/stream_put.go:114-166,265-299): NewBuffer sizes -> Split -> Encode -> per-shard CRC ->
     blobnode ShardPut (datafile.Write image) -> the CRC blobnode returns must equal the client's crcOrigin.
GET  (access/stream/stream_get.go:214-219,454-465): read N+X shards, ReconstructData on whole shards and on a
     byte SEGMENT of every shard (valid because RS is byte-wise).
REPAIR (blobnode/worker_slice_recover.go:804-888): per-bid Reconstruct + Verify over a task's bids in one batch.
"""
import zlib

import numpy as np
import pytest

from mirror import ec as cm

pytestmark = pytest.mark.gpu


def _put_blob(cb, enc, eng, t, blob: bytes, bid: int):
    sizes = cm.GetBufferSizes(len(blob), t)
    buf = np.zeros(sizes.ECSize, dtype=np.uint8)                       # ec.NewBuffer: data | zero pad | parity
    buf[:len(blob)] = np.frombuffer(blob, dtype=np.uint8)
    shards = enc.Split(cm.GoSlice(buf, sizes.ECDataSize))              # stream_put.go:128
    # fused engine call: parity + crcOrigin of every shard from one pass (replaces :146 and :265-269)
    arrs = [s.data for s in shards]
    crcs = eng.encode(arrs, crc=True)
    stored = []
    for i, s in enumerate(shards):
        image, blobnode_crc = cm.WriteShard(bid, 1000 + i, s.data.tobytes())   # blobnode ShardPut -> datafile.Write
        assert blobnode_crc == int(crcs[i]) == zlib.crc32(s.data.tobytes())    # stream_put.go:295-299
        stored.append(image)
    return sizes, stored


def test_put_get_repair_flow(cb, oracle):
    t = cm.EC12P4.Tactic()
    enc = cm.NewEncoder(cm.Config(CodeMode=t))
    eng = cb.RSEngine(t.N, t.M)
    rng = np.random.default_rng(2024)
    blobs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (1, 2048 * 12, 100_000, 4 << 20)]
    for bid, blob in enumerate(blobs):
        sizes, stored = _put_blob(cb, enc, eng, t, blob, bid)
        assert sizes.ShardSize == max((len(blob) + 11) // 12, 2048)
        # ---- GET with 3 broken shards: blobnode read verifies the framing, access reconstructs the data
        bad = [0, 5, 13]
        shards = []
        for i, image in enumerate(stored):
            if i in bad:
                shards.append(cm.GoSlice(np.zeros(sizes.ShardSize, np.uint8)).truncate())   # cap kept, len 0
            else:
                _, _, _, data = cm.ReadShard(image)
                shards.append(cm.GoSlice(np.frombuffer(data, dtype=np.uint8).copy()))
        enc.ReconstructData(shards, bad)
        assert enc.Join(shards, len(blob)) == blob
        assert shards[13].len == 0                                       # missing parity stays missing
        # ---- segment reconstruct (stream_get.go:454-461): any byte range of the shards decodes independently
        full = [cm.ReadShard(image)[3] for image in stored]
        off, ln = sizes.ShardSize // 3, max(1, sizes.ShardSize // 5)
        seg = [None if i in bad else np.frombuffer(full[i][off:off + ln], dtype=np.uint8).copy() for i in range(16)]
        out = eng.reconstruct(seg, data_only=True)
        for i in (0, 5):
            assert out[i].tobytes() == full[i][off:off + ln]


def test_repair_batch_many_bids(cb, oracle):
    """One repair task = many bids of different sizes with the same bad chunk indexes (worker_slice_recover.go:822-885)."""
    k, m = 6, 6
    eng = cb.RSEngine(k, m)
    ora = oracle.RS(k, m)
    rng = np.random.default_rng(9)
    bad = [1, 4, 8, 11]
    stripes, originals = [], []
    for bid in range(40):
        S = int(rng.integers(2048, 50000))
        sh = [rng.integers(0, 256, S, dtype=np.uint8) for _ in range(k)] + [np.zeros(S, np.uint8) for _ in range(m)]
        ora.encode(sh)
        originals.append([x.copy() for x in sh])
        present = np.ones(k + m, np.uint8)
        present[bad] = 0
        for i in bad:
            sh[i][:] = 0
        stripes.append((sh, present))
    ok, crcs = eng.reconstruct_batch(stripes, verify=True, crc=True)
    assert all(ok)
    for s, ((sh, _), orig) in enumerate(zip(stripes, originals)):
        assert all((a == b).all() for a, b in zip(sh, orig))
        # the checksums the repair worker needs for the rebuilt shards (worker_slice_recover.go:367-373), from the same call
        for i in bad:
            assert int(crcs[s, i]) == zlib.crc32(orig[i].tobytes()), (s, i)
        assert all(int(crcs[s, i]) == 0 for i in range(k + m) if i not in bad)   # untouched entries


def test_in_process_multi_device_partition(oracle):
    """cubeec_init(devices) + cubeec_encode_contig: one process, stripes split over 2 devices.
    Runs in a subprocess because the device list is fixed at first use."""
    import subprocess
    import sys
    import torch
    # on a single-GPU box the same partition code runs over two engine contexts of device 0 (cubeec_init([0, 0])):
    # two sets of tables, lanes and worker threads, stripes split between them
    devs = "[0, 1]" if torch.cuda.device_count() >= 2 else "[0, 0]"
    code = r'''
import sys, zlib, numpy as np
sys.path.insert(0, ".")
import cubefs_b200 as cb
from oracle import pyoracle
cb.init(DEVS)
assert cb.device_count() == 2
k, m, S, ns = 12, 4, 21846, 301
rng = np.random.default_rng(1)
buf = rng.integers(0, 256, (ns, (k + m) * S), dtype=np.uint8)
ref = buf.copy()
crc, _ = cb.RSEngine(k, m).encode_contig(buf, S, ns, (k + m) * S, crc=True)
ora = pyoracle.RS(k, m)
for s in (0, 1, 149, 150, 151, 299, 300):
    sh = [ref[s, i * S:(i + 1) * S].copy() for i in range(k + m)]
    ora.encode(sh)
    for i in range(k + m):
        assert (buf[s, i * S:(i + 1) * S] == sh[i]).all() and crc[s, i] == zlib.crc32(sh[i].tobytes())
print("ok")
'''.replace("DEVS", devs)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
