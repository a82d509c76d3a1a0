#!/usr/bin/env python3
"""Throughput of the stand-alone CRC32 kernels on device-resident buffers (not a bench line: A/B aid).

For whole-buffer CRCs of 1 MiB shards, crc32block payload blocks (65,532 B) of the C2 shard, and 4 KiB buffers it
times cubeec_dev_crc32 with crc_flat_kernel (default) and the first-generation crc_range_kernel (force 10) and prints
bytes checksummed / CUDA-event time as a fraction of the measured HBM peak.      python tools/crc_speed.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cubefs_b200 as cb  # noqa: E402
from cubefs_b200.engine import dev_crc32  # noqa: E402


def main():
    cb.init([0])
    peak = 6570.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for (nbuf, L, block) in ((2040, 1 << 20, 0), (6144, 349526, 65532), (6144, 349526, 0), (500000, 4096, 0)):
        P = (L + 127) // 128 * 128
        buf = torch.randint(0, 256, (nbuf, P), dtype=torch.uint8, device="cuda")
        units = (L + block - 1) // block if block else 1
        whole = torch.zeros(nbuf, dtype=torch.int32, device="cuda")
        blocks = torch.zeros(nbuf * units, dtype=torch.int32, device="cuda")
        ref = None
        for force in (0, 10):
            cb.force_kernel(force)
            f = lambda: dev_crc32(buf.data_ptr(), L, P, nbuf, block_payload=block, d_whole=whole.data_ptr(),
                                  d_blocks=blocks.data_ptr() if block else 0, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                f()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            got = (whole.clone(), blocks.clone())
            same = True if ref is None else bool((got[0] == ref[0]).all() and (got[1] == ref[1]).all())
            ref = ref or got
            print(json.dumps({"buffers": nbuf, "len": L, "block": block, "force": force, "kernel": cb.last_kernel(), "ms": round(ms, 4),
                              "GB_s": round(nbuf * L / ms / 1e6, 1), "frac_of_measured_hbm": round(nbuf * L / ms / 1e6 / peak, 3),
                              "same_as_first": same}), flush=True)
        cb.force_kernel(0)
        del buf


if __name__ == "__main__":
    main()
