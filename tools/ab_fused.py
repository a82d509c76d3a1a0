#!/usr/bin/env python3
"""A/B of the fused encode+CRC kernel variants on a device-resident C2 batch (RS(12,4), S = 349,526).

For every cubeec_debug_force_kernel value given (7 = tile-split rs_bs_kernel<crc>, 0 = default flat-split
rs_bsf_kernel, 1000+T = flat split with T threads per CTA) it checks parity + CRCs against the first variant
(bit-exact) and prints the CUDA-event time of `steps` back-to-back launches.  Not a bench line: A/B only.

python tools/ab_fused.py [--stripes 1024,383] [--force 7,0,1512,1448,1384,1320,1256] [--k 12 --m 4 --shard 349526]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stripes", default="1024")
    ap.add_argument("--force", default="7,0,1512,1448,1384,1320,1256")
    ap.add_argument("--k", type=int, default=12)
    ap.add_argument("--m", type=int, default=4)
    ap.add_argument("--shard", type=int, default=349526)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch

    import cubefs_b200 as cb
    dev = torch.device("cuda", 0)
    cb.init([0])
    k, m, S = args.k, args.m, args.shard
    n = k + m
    P = (S + 127) // 128 * 128
    eng = cb.RSEngine(k, m)
    peak = 6570.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for ns in [int(x) for x in args.stripes.split(",")]:
        g = torch.Generator(device=dev).manual_seed(0xC0BEF5)
        batch = torch.empty((ns, n, P), dtype=torch.uint8, device=dev)
        for s0 in range(0, ns, 64):
            batch[s0:s0 + 64] = torch.randint(0, 256, batch[s0:s0 + 64].shape, dtype=torch.uint8, device=dev, generator=g)
        dcrc = torch.zeros(ns * n, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ref_par = ref_crc = None
        for f in [int(x) for x in args.force.split(",")]:
            cb.force_kernel(f)
            batch[:, k:, :].zero_()
            dcrc.zero_()
            torch.cuda.synchronize()   # the engine runs on its own non-blocking stream when torch's is the legacy default stream
            try:
                eng.dev_encode(batch.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), stream=stream, device=0)
                torch.cuda.synchronize()
            except Exception as e:   # noqa: BLE001
                print(json.dumps({"force": f, "stripes": ns, "error": str(e)}), flush=True)
                continue
            par = batch[:, k:, :S].clone()
            crc = dcrc.clone()
            # independent check of two stripes against the CPU oracle / zlib, and that the data shards were not touched
            import zlib
            from oracle import pyoracle
            orc = None
            for s_ in (0, ns - 1):
                h = batch[s_, :, :S].cpu().numpy()
                want = [h[i].copy() for i in range(k)] + [np.zeros(S, np.uint8) for _ in range(m)]
                pyoracle.RS(k, m).encode(want)
                c_ = crc.cpu().numpy().view(np.uint32).reshape(ns, n)[s_]
                good = all(np.array_equal(h[i], want[i]) for i in range(n)) and all(int(c_[i]) == zlib.crc32(want[i].tobytes()) for i in range(n))
                orc = good if orc is None else (orc and good)
            dsum = int(batch[:, :k, :S].to(torch.int64).sum().item())
            ok = True
            detail = None
            if ref_par is None:
                ref_par, ref_crc = par, crc
            else:
                ok = bool(torch.equal(par, ref_par) and torch.equal(crc, ref_crc))
                if not ok:
                    bad_par = (par != ref_par).flatten(2).any(dim=2)          # [ns, m]
                    bad_crc = (crc != ref_crc).reshape(ns, n)
                    ps, cs = bad_par.nonzero().tolist(), bad_crc.nonzero().tolist()
                    first = None
                    if ps:
                        s0, r0 = ps[0]
                        first = int((par[s0, r0] != ref_par[s0, r0]).nonzero()[0].item())
                    detail = {"bad_parity_shards": len(ps), "first_parity": ps[:6], "first_byte": first,
                              "bad_crcs": len(cs), "first_crcs": cs[:12]}
                    # is the run reproducible?
                    batch[:, k:, :].zero_()
                    dcrc.zero_()
                    torch.cuda.synchronize()
                    eng.dev_encode(batch.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), stream=stream, device=0)
                    torch.cuda.synchronize()
                    detail["repeat_identical"] = bool(torch.equal(batch[:, k:, :S], par) and torch.equal(dcrc, crc))
            for _ in range(3):
                eng.dev_encode(batch.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), stream=stream, device=0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                eng.dev_encode(batch.data_ptr(), S, P, n * P, ns, d_crc=dcrc.data_ptr(), stream=stream, device=0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            frac = n * S * ns / (ms * 1e-3) / 1e9 / peak
            print(json.dumps({"force": f, "stripes": ns, "kernel": cb.last_kernel(), "ms": round(ms, 4), "frac": round(frac, 4),
                              "same_as_first": ok, "oracle_ok": orc, "data_sum": dsum, "detail": detail}), flush=True)
        cb.force_kernel(0)
        del batch


if __name__ == "__main__":
    main()
