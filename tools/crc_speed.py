import sys, json, torch
sys.path.insert(0, '.')
import cubefs_b200 as cb
cb.init([0])
from cubefs_b200.engine import dev_crc32
for (nbuf, L) in ((2040, 1 << 20), (16384, 349568 // 1 and 349526), (500000, 4096)):
    P = (L + 127) // 128 * 128
    buf = torch.randint(0, 256, (nbuf, P), dtype=torch.uint8, device='cuda')
    out = torch.zeros(nbuf, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        dev_crc32(buf.data_ptr(), L, P, nbuf, d_whole=out.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        dev_crc32(buf.data_ptr(), L, P, nbuf, d_whole=out.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"buffers": nbuf, "len": L, "ms": round(ms, 4), "GBps": round(nbuf * L / ms / 1e6, 1), "frac": round(nbuf * L / ms / 1e6 / 6570.6, 3), "kernel": cb.last_kernel()}))
    del buf
