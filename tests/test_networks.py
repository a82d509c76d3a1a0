"""CPU check of the GENERATED XOR networks (cubefs_b200/csrc/bs_net_gen.cuh) against the oracle.

The bit-sliced kernels multiply by the coding coefficients with compile-time XOR networks on bit-planes
(gen_bitslice.py).  Here the emitted C statements of every network -- all single-pass codes, every pass of
both pass plans of the m > 4 codes, the LRC local stripes -- are executed in Python on random bit-planes
and compared with the oracle's GF(2^8) arithmetic (klauspost mulTable, RS/galois.go:83) and its
buildMatrix rows (RS/reedsolomon.go:220-244).  No GPU needed: a wrong network is caught before a kernel
ever runs."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "cubefs_b200", "csrc", "bs_net_gen.cuh")


def parse_nets():
    txt = open(HDR).read()
    nets = []
    for m in re.finditer(r"template <> struct BsNet<(\d+), (\d+), (\d+)> \{(.*?)\n\};", txt, re.S):
        k, mm, v, body = int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4)
        mt, r0 = map(int, re.search(r"kTotalM = (\d+), kRow0 = (\d+);", body).groups())
        rows = list(map(int, re.search(r"kRows\[\d+\] = \{([^}]*)\}", body).group(1).split(",")))
        shards = {}
        for c, blk in re.findall(r"if constexpr \(C == (\d+)\) \{(.*?)\n    \}", body, re.S):
            stmts = []
            for ln in blk.strip().splitlines():
                ln = ln.strip().rstrip(";")
                if ln.startswith("const uint32_t "):
                    ln = ln[len("const uint32_t "):]
                stmts.append(ln)
            shards[int(c)] = stmts
        nets.append(dict(k=k, m=mm, v=v, mt=mt, r0=r0, rows=rows, shards=shards))
    return nets


def macro_list(name):
    txt = open(HDR).read()
    line = next(l for l in txt.splitlines() if l.startswith("#define " + name))
    return [tuple(int(x) for x in t.split(",")) for t in re.findall(r"X\(([^)]*)\)", line)]


NETS = parse_nets()


def test_every_configuration_has_its_network():
    full = macro_list("CUBEEC_BS_CONFIGS(X)")
    passes = macro_list("CUBEEC_BS_PASS_CONFIGS(X)")
    have = {(n["k"], n["m"], n["v"]) for n in NETS}
    assert len(NETS) == len(full) + len(passes)
    for (k, m) in full:
        assert (k, m, 0) in have
    for (k, m, v, mt, r0, pi, plan) in passes:
        assert (k, m, v) in have and v == plan * 10000 + mt * 100 + r0
    # every plan of every m > 4 code covers each parity row exactly once, passes numbered in row order
    for plan in (0, 1):
        codes = {(k, mt) for (k, m, v, mt, r0, pi, pl) in passes if pl == plan}
        assert codes, plan
        for (k, mt) in codes:
            ps = sorted((pi, r0, m) for (kk, m, v, mtt, r0, pi, pl) in passes if (kk, mtt, pl) == (k, mt, plan))
            assert [p[0] for p in ps] == list(range(len(ps)))
            nxt = 0
            for (_, r0, m) in ps:
                assert r0 == nxt and 1 <= m <= (4 if plan == 0 else 6)
                nxt += m
            assert nxt == mt


@pytest.mark.parametrize("net", NETS, ids=lambda n: f"rs{n['k']}_{n['mt']}_rows{n['r0']}+{n['m']}_v{n['v']}")
def test_network_equals_oracle_gf_arithmetic(net, oracle):
    k, m, mt, r0 = net["k"], net["m"], net["mt"], net["r0"]
    # the baked-in rows are the reference's generator rows
    want_rows = oracle.build_matrix(k, k + mt)[k + r0:k + r0 + m]
    assert net["rows"] == [int(x) for x in want_rows.reshape(-1)]
    mul = oracle.gf_tables()[2]
    rng = np.random.default_rng(k * 1000 + mt * 10 + r0)
    data = rng.integers(0, 256, (k, 32), dtype=np.uint8)       # 32 byte columns, one per bit of a plane word
    acc = [0] * (8 * m)
    for c in range(k):
        p = [int(sum(((int(data[c, j]) >> b) & 1) << j for j in range(32))) for b in range(8)]
        env = {"p": p, "acc": acc}
        for st in net["shards"][c]:
            exec(st, {}, env)   # the generated C statements are valid Python ('x = a ^ b', 'acc[i] ^= a ^ b')
    got = np.zeros((m, 32), dtype=np.uint8)
    for r in range(m):
        for b in range(8):
            for j in range(32):
                got[r, j] |= ((acc[r * 8 + b] >> j) & 1) << b
    want = np.zeros((m, 32), dtype=np.uint8)
    for r in range(m):
        for c in range(k):
            want[r] ^= mul[want_rows[r, c], data[c]]
    assert (got == want).all()
