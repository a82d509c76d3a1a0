"""world_size-2 gloo test (CPU) of the N>1 host path: contiguous stripe partition, broadcast of the
coding matrix from rank 0 (the only shared state), max-over-ranks timing, and rank-local encoding of
the owned slice reproducing the single-process result (encode runs through the oracle here: no GPU)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from cubefs_b200 import parallel
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k, m, S, ns = 12, 4, 1024 + 6, 11
    rows = parallel.parity_rows_host(k, m) if rank == 0 else None
    rows = parallel.broadcast_matrix(rows, k, m, src=0)
    assert (rows == pyoracle.RS(k, m).parity_rows).all()          # product-side matrix == oracle == klauspost
    first, last = parallel.partition(ns, world, rank)
    rng = np.random.default_rng(1234)                              # same batch on every rank
    batch = rng.integers(0, 256, (ns, k + m, S), dtype=np.uint8)
    ora = pyoracle.RS(k, m)
    for s in range(first, last):
        sh = [batch[s, i] for i in range(k + m)]
        ora.encode(sh)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), batch[first:last])
    slow = parallel.max_over_ranks(1.0 + rank)
    assert slow == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_broadcast(tmp_path, oracle):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    k, m, S, ns = 12, 4, 1024 + 6, 11
    rng = np.random.default_rng(1234)
    batch = rng.integers(0, 256, (ns, k + m, S), dtype=np.uint8)
    ora = oracle.RS(k, m)
    for s in range(ns):
        ora.encode([batch[s, i] for i in range(k + m)])
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert got.shape == batch.shape and (got == batch).all()


def test_partition_covers_everything():
    from cubefs_b200 import parallel
    for ns in (0, 1, 7, 1024):
        for world in (1, 2, 3, 8):
            spans = [parallel.partition(ns, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == ns
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
