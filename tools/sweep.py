#!/usr/bin/env python3
"""BASELINE configs C4/C5: device-resident encode throughput over shard sizes and code modes.

python tools/sweep.py [--gpu 0] [--crc 1]   -> one JSON line per (k, m, S) with GiB/s of data and
the fraction of the measured HBM peak ((k+m)*S per stripe of algorithmic traffic).

8 GPUs: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1
--master-port 29511 tools/sweep.py   -> every rank runs the same cases on its own GPU (stripes are independent: the batch
is partitioned, no data-path collective), a barrier brackets each timing, the time of a case is the MAX over ranks and
rank 0 prints the aggregate (sum of all ranks' bytes / that time) next to the per-GPU roofline fraction."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cubefs_b200 as cb  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))


def sync_max_ms(ms, dev):
    """a multi-GPU case takes as long as its slowest rank"""
    if WORLD == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dev):
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize(dev)


def run(eng, k, m, S, crc, total_bytes, dev, steps=5):
    n = k + m
    P = (S + 127) // 128 * 128
    ns = max(1, int(total_bytes // (n * P)))
    batch = torch.randint(0, 256, (ns, n, P), dtype=torch.uint8, device=dev)
    dcrc = torch.zeros(ns * n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream

    def step():
        eng.dev_encode(batch.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr() if crc else 0, stream=st, device=dev.index)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(dev)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier(dev)
    ms = sync_max_ms(e0.elapsed_time(e1) / steps, dev)
    return ns, ms, cb.last_kernel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--crc", type=int, default=1)
    ap.add_argument("--gib", type=float, default=2.0, help="batch size per case (GiB in HBM)")
    ap.add_argument("--forces", default="", help="comma list of cubeec_debug_force_kernel values to run every case under")
    ap.add_argument("--modes", action="store_true",
                    help="instead of the C4/C5 shard-size sweep: every predefined EC code mode (codemode.go:65-94) at "
                         "1 MiB shards, bit-sliced path and (A/B) the table kernels")
    args = ap.parse_args()
    if WORLD > 1:
        import torch.distributed as dist
        args.gpu = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(args.gpu)
        dist.init_process_group("nccl", device_id=torch.device("cuda", args.gpu))
    dev = torch.device("cuda", args.gpu)
    torch.cuda.set_device(dev)
    cb.init([args.gpu])
    pk = peak()
    cases = [(6, 3, s) for s in (4096, 16384, 65536, 262144, 1 << 20, 4 << 20, 8 << 20)]
    cases += [(12, 4, s) for s in (4096, 16384, 65536, 262144, 349526, 1 << 20, 4 << 20, 8 << 20)]
    cases += [(20, 4, 1 << 20), (4, 2, 65536)]
    forces = [0]
    if args.modes:
        # global RS(N, M) of every predefined mode; m > 4 runs ceil(m/4) passes
        cases = [(k, m, 1 << 20) for (k, m) in ((15, 12), (6, 6), (16, 20), (6, 10), (6, 3), (4, 4), (12, 4), (16, 4), (3, 3),
                                                 (10, 4), (12, 9), (24, 8), (6, 8))]
        forces = [0, 1]
    if args.forces:
        forces = [int(x) for x in args.forces.split(",")]
    if args.modes and not args.forces:
        # LRC modes: global + per-AZ local passes fused on the device (cubeec_dev_lrc_encode)
        for (N, M, L, az) in ((16, 20, 2, 2), (6, 10, 2, 2), (6, 3, 3, 3), (4, 4, 2, 2)):
            S = 1 << 20
            n, P = N + M + L, S
            ns = max(1, int(args.gib * (1 << 30) // (n * P)))
            ge, le = cb.RSEngine(N, M), cb.RSEngine((N + M) // az, L // az)
            batch = torch.randint(0, 256, (ns, n, P), dtype=torch.uint8, device=dev)
            dcrc = torch.zeros(ns * n, dtype=torch.int32, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            for crc in ([0, 1] if args.crc else [0]):
                f = lambda: cb.dev_lrc_encode(ge, le, az, batch.data_ptr(), S, P, n * P, ns,
                                              d_crc=dcrc.data_ptr() if crc else 0, stream=st, device=dev.index)
                for _ in range(3):
                    f()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(dev)
                e0.record()
                for _ in range(5):
                    f()
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / 5
                moved = n * S * ns / (ms * 1e-3) / 1e9
                if RANK == 0:
                  print(json.dumps({"lrc": [N, M, L, az], "shard_bytes": S, "stripes": ns, "crc": bool(crc), "kernel": cb.last_kernel(),
                                  "ms": round(ms, 4), "data_GiB_s": round(N * S * ns / (ms * 1e-3) / 2**30, 1),
                                  "moved_GB_s": round(moved, 1), "frac_of_measured_hbm": round(moved / pk, 4)}), flush=True)
    engines = {}
    for (k, m, S), force in [(c, f) for c in cases for f in forces]:
        eng = engines.setdefault((k, m), cb.RSEngine(k, m))
        for crc in ([0, 1] if args.crc else [0]):
            cb.force_kernel(force)
            ns, ms, kern = run(eng, k, m, S, crc, args.gib * (1 << 30), dev)
            cb.force_kernel(0)
            moved = (k + m) * S * ns / (ms * 1e-3) / 1e9
            if RANK == 0:
                print(json.dumps({"k": k, "m": m, "shard_bytes": S, "stripes_per_gpu": ns, "n_gpus": WORLD, "crc": bool(crc), "kernel": kern, "force": force,
                                  "ms": round(ms, 4), "data_GiB_s_all_gpus": round(WORLD * k * S * ns / (ms * 1e-3) / 2**30, 1),
                                  "moved_GB_s_per_gpu": round(moved, 1), "frac_of_measured_hbm": round(moved / pk, 4),
                                  "frac_nominal_8TBs": round(moved / 8000.0, 4)}), flush=True)
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
