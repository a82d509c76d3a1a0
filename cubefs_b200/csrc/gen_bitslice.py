#!/usr/bin/env python3
"""Generate bs_net_gen.cuh: straight-line GF(2) XOR networks for the bit-sliced encode kernel.

For every (k, m) configuration listed in CONFIGS the parity rows of the generator matrix are
the ones klauspost/reedsolomon v1.11.7 builds for reedsolomon.New(k, m) with default options
(vandermonde(total,k) * inverse(top k x k), vendor/github.com/klauspost/reedsolomon/
reedsolomon.go:220-244; field polynomial 0x11D, galois.go:13-26) -- the matrix CubeFS uses
(blobstore/common/ec/encoder.go:86).  Multiplying a byte by a constant c is a GF(2)-linear map;
with 32 bytes held as 8 bit-planes (plane j = bit j of each byte) output plane i of `c * x` is the
XOR of the input planes j for which bit i of (c * 2^j) is set.  Per input shard the generator
emits a "four Russians" network: all XOR combinations of planes 0..3 and of planes 4..7 that
are needed (<= 11 + 11 XORs), then one 3-input XOR per output plane
(acc ^= lo_combo ^ hi_combo), which nvcc maps to a single LOP3.

This file is part of the product build (run by the Makefile); it does not use oracle/.
The generated header also carries the parity rows so that the engine can check at run time that
the handle's matrix equals the one the network was specialised for.
"""
import sys

# (k, m) with m <= 4: one pass, 8*m accumulator registers.  The second group are the local stripes of
# the LRC code modes (codemode.go:72-73,83,86-88: RS((N+M)/AZ, L/AZ)).
CONFIGS = [(4, 2), (6, 3), (12, 4), (20, 4), (16, 4), (10, 4), (3, 3), (4, 4), (8, 4), (6, 2), (10, 2), (5, 2),
           (18, 1), (8, 1), (3, 1), (4, 1), (4, 3)]
# (k, m) with m > 4 (EC15P12, EC6P6, EC16P20L2, EC6P10L2, EC12P9, EC24P8, EC6P8L10 and its local stripe):
# passes of 4 parity rows, each pass re-reads the k data shards
PASS_CONFIGS = [(15, 12), (6, 6), (16, 20), (6, 10), (12, 9), (24, 8), (6, 8), (7, 5)]


def gf_tables():
    exp, log = [0] * 512, [0] * 256
    x = 1
    for i in range(255):
        exp[i] = exp[i + 255] = x
        log[x] = i
        x <<= 1
        if x & 0x100:
            x ^= 0x11D
    return exp, log


EXP, LOG = gf_tables()


def mul(a, b):
    return 0 if a == 0 or b == 0 else EXP[LOG[a] + LOG[b]]


def inv(a):
    return EXP[255 - LOG[a]]


def gexp(a, n):
    if n == 0:
        return 1
    if a == 0:
        return 0
    return EXP[(LOG[a] * n) % 255]


def invert(m):
    n = len(m)
    w = [row[:] + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(m)]
    for col in range(n):
        piv = next(r for r in range(col, n) if w[r][col])
        w[col], w[piv] = w[piv], w[col]
        s = inv(w[col][col])
        w[col] = [mul(v, s) for v in w[col]]
        for r in range(n):
            if r != col and w[r][col]:
                f = w[r][col]
                w[r] = [a ^ mul(f, b) for a, b in zip(w[r], w[col])]
    return [row[n:] for row in w]


def parity_rows(k, m):
    total = k + m
    vm = [[gexp(r, c) for c in range(k)] for r in range(total)]
    top_inv = invert([row[:] for row in vm[:k]])
    full = [[0] * k for _ in range(total)]
    for r in range(total):
        for c in range(k):
            v = 0
            for i in range(k):
                v ^= mul(vm[r][i], top_inv[i][c])
            full[r][c] = v
    for i in range(k):
        assert full[i] == [1 if j == i else 0 for j in range(k)]
    return full[k:]


def row_masks(c):
    """masks[i] = set of input planes j feeding output plane i of (c * x)."""
    masks = []
    for i in range(8):
        mk = 0
        for j in range(8):
            if (mul(c, 1 << j) >> i) & 1:
                mk |= 1 << j
        masks.append(mk)
    return masks


def emit_shard(lines, coefs, var_p="p", var_acc="acc"):
    """coefs[r] = coefficient of this shard for output r."""
    need_lo, need_hi = set(), set()
    uses = []
    for r, c in enumerate(coefs):
        for i, mk in enumerate(row_masks(c)):
            lo, hi = mk & 15, mk >> 4
            uses.append((r * 8 + i, lo, hi))
            if lo:
                need_lo.add(lo)
            if hi:
                need_hi.add(hi)

    def combos(need, base, prefix):
        # close the needed set under "drop lowest set bit" so each combo costs one XOR
        todo = set(need)
        closed = set()
        while todo:
            x = todo.pop()
            if x in closed or x & (x - 1) == 0:
                continue
            closed.add(x)
            todo.add(x & (x - 1))
        names = {}
        for x in range(1, 16):
            if x & (x - 1) == 0:
                names[x] = f"{var_p}[{base + x.bit_length() - 1}]"
        for x in sorted(closed):
            parent, low = x & (x - 1), x & -x
            names[x] = f"{prefix}{x}"
            lines.append(f"    const uint32_t {prefix}{x} = {names[parent]} ^ {names[low]};")
        return names

    lo_names = combos(need_lo, 0, "l")
    hi_names = combos(need_hi, 4, "h")
    n_xor3 = 0
    for idx, lo, hi in uses:
        terms = []
        if lo:
            terms.append(lo_names[lo])
        if hi:
            terms.append(hi_names[hi])
        if terms:
            lines.append(f"    {var_acc}[{idx}] ^= {' ^ '.join(terms)};")
            n_xor3 += 1
    return n_xor3


N_PARTS = 5


def emit_shard_parts(coefs):
    """The same network as emit_shard, as N_PARTS statement lists for the software-pipelined kernel
    (bs_flat.cuh): every XOR combination is created right before its first use (short live ranges) and the
    outputs are visited grouped by their low-nibble combination.  Combinations live in t[]: t[x] = XOR of
    the planes 0..3 selected by x, t[16 + x] = the same for planes 4..7."""
    uses = []
    for r, c in enumerate(coefs):
        for i, mk in enumerate(row_masks(c)):
            if mk:
                uses.append((mk & 15, mk >> 4, r * 8 + i))
    uses.sort()
    built = set()
    stmts = []

    def name(x, base):
        if x & (x - 1) == 0:
            return f"p[{base + x.bit_length() - 1}]"
        return f"t[{(16 if base else 0) + x}]"

    def build(x, base):
        if x & (x - 1) == 0 or (x, base) in built:
            return
        parent, low = x & (x - 1), x & -x
        build(parent, base)
        built.add((x, base))
        stmts.append(f"{name(x, base)} = {name(parent, base)} ^ {name(low, base)};")

    for lo, hi, idx in uses:
        terms = []
        if lo:
            build(lo, 0)
            terms.append(name(lo, 0))
        if hi:
            build(hi, 4)
            terms.append(name(hi, 4))
        stmts.append(f"acc[{idx}] ^= {' ^ '.join(terms)};")
    n = len(stmts)
    return [stmts[n * j // N_PARTS: n * (j + 1) // N_PARTS] for j in range(N_PARTS)]


def emit_net(L, k, m, v, rows, mt, r0):
    L.append(f"template <> struct BsNet<{k}, {m}, {v}> {{")
    L.append(f"  static constexpr int K = {k}, M = {m}, kTotalM = {mt}, kRow0 = {r0};")
    flat = ", ".join(str(x) for row in rows for x in row)
    L.append(f"  static constexpr uint8_t kRows[{k * m}] = {{{flat}}};")
    L.append("  // acc[r*8+i] ^= plane i of (rows[r][C] * shard C), shard C given as 8 bit-planes p[0..7]")
    L.append("  template <int C> static __device__ __forceinline__ void apply(const uint32_t (&p)[8], uint32_t (&acc)[8 * M]) {")
    for c in range(k):
        L.append(f"    if constexpr (C == {c}) {{")
        body = []
        emit_shard(body, [rows[r][c] for r in range(m)])
        L.extend("  " + ln for ln in body)
        L.append("    }")
    L.append("  }")
    L.append(f"  static constexpr int kParts = {N_PARTS};")
    L.append("  // the same network in kParts pieces (combinations created just before first use, kept in t[])")
    L.append("  template <int C, int J> static __device__ __forceinline__ void part(const uint32_t (&p)[8], uint32_t (&acc)[8 * M], uint32_t (&t)[32]) {")
    for c in range(k):
        parts = emit_shard_parts([rows[r][c] for r in range(m)])
        for j, st in enumerate(parts):
            if st:
                L.append(f"    if constexpr (C == {c} && J == {j}) {{ " + " ".join(st) + " }")
    L.append("  }")
    L.append("};")


def const_mul_statements(g):
    """Statements of a[i] ^= plane i of (g * x): greedy common-pair elimination (Paar) over the 8x8 bit
    matrix of the constant, then every row folded two signals per 3-input XOR."""
    rows = [set(f"p[{j}]" for j in range(8) if (mk >> j) & 1) for mk in row_masks(g)]
    stmts, ntmp = [], 0
    while True:
        best, cnt = None, 1
        sigs = sorted(set().union(*rows))
        for a in range(len(sigs)):
            for b in range(a + 1, len(sigs)):
                c = sum(1 for r in rows if sigs[a] in r and sigs[b] in r)
                if c > cnt:
                    best, cnt = (sigs[a], sigs[b]), c
        if best is None:
            break
        # a shared pair pays when it saves LOP3s: rows with an odd number of signals lose nothing by keeping
        # one signal unpaired, so only take pairs used by at least 3 rows (1 LOP3 to build, >= 1.5 saved)
        if cnt < 3:
            break
        name = f"t{ntmp}"
        ntmp += 1
        stmts.append(f"const uint32_t {name} = {best[0]} ^ {best[1]};")
        for r in rows:
            if best[0] in r and best[1] in r:
                r.discard(best[0])
                r.discard(best[1])
                r.add(name)
    for i, r in enumerate(rows):
        terms = sorted(r)
        while terms:
            take, terms = terms[:2], terms[2:]
            stmts.append(f"a[{i}] ^= {' ^ '.join(take)};")
    return stmts


def emit_const_multipliers(L):
    """BsMul<G>::mac(p, a): a[i] ^= plane i of (G * x) for every field constant G -- the networks of the
    generic bit-sliced kernel (bitslice_gen.cu), which picks one per (input shard, output) at run time
    through a 256-way switch on the coefficient (reconstruct with data-dependent decode rows, custom
    matrices)."""
    L.append("// ---- constant multipliers: a[i] ^= plane i of (G * x), x given as 8 bit-planes p[0..7]")
    L.append("template <int G> struct BsMul;")
    for g in range(1, 256):
        L.append(f"template <> struct BsMul<{g}> {{ static __device__ __forceinline__ void mac(const uint32_t (&p)[8], uint32_t (&a)[8]) {{ "
                 + " ".join(const_mul_statements(g)) + " } };")


def pass_split(mt, width):
    """(first row, rows) of every pass of a code with mt > 4 parity shards: ceil(mt/width) passes of
    balanced size (8 rows -> 4+4 rather than 6+2: a 2-row pass re-reads all the data for little)"""
    n = -(-mt // width)
    base, extra = divmod(mt, n)
    out, r0 = [], 0
    for i in range(n):
        m = base + (1 if i < extra else 0)
        out.append((r0, m))
        r0 += m
    return out


# Two pass plans per code.  Plan 0 (fused CRC): 4 rows per pass -- the first pass also carries the k CRC
# registers of the data shards.  Plan 1 (plain encode / verify): 6 rows per pass (48 accumulator
# registers, no spills at 128 registers/thread), e.g. EC15P12 reads the data twice instead of 3 times.
PLAN_WIDTH = (4, 6)


def pass_id(mt, r0, plan):
    """template tag of the pass that computes parity rows [r0, ...) of a code with mt > 4 parity shards"""
    return plan * 10000 + mt * 100 + r0


def main(out_path):
    L = []
    L.append("// GENERATED by gen_bitslice.py -- do not edit.  GF(2) XOR networks of the bit-sliced encode kernel.")
    L.append("#pragma once")
    L.append("#include <cstdint>")
    L.append("namespace cbe {")
    L.append("// BsNet<K, M, 0>: all M <= 4 parity rows of RS(K, M).  BsNet<K, M, V>, V = 10000*PLAN + 100*MT + R0:")
    L.append("// rows [R0, R0+M) of RS(K, MT) with MT > 4 -- such codes run several passes over the data shards.")
    L.append("template <int K, int M, int V = 0> struct BsNet;")
    full, passes = [], []
    for (k, m) in CONFIGS:
        emit_net(L, k, m, 0, parity_rows(k, m), m, 0)
        full.append(f"X({k}, {m})")
    for (k, mt) in PASS_CONFIGS:
        rows = parity_rows(k, mt)
        for plan, width in enumerate(PLAN_WIDTH):
            for pi, (r0, m) in enumerate(pass_split(mt, width)):
                emit_net(L, k, m, pass_id(mt, r0, plan), rows[r0:r0 + m], mt, r0)
                passes.append(f"X({k}, {m}, {pass_id(mt, r0, plan)}, {mt}, {r0}, {pi}, {plan})")
    emit_const_multipliers(L)
    L.append("}  // namespace cbe")
    L.append("// X(K, M): single-pass configurations;  X(K, M, V, MT, R0, PASS, PLAN): passes of the MT > 4 codes (plan 0: fused CRC, plan 1: plain/verify)")
    L.append("#define CUBEEC_BS_CONFIGS(X) " + " ".join(full))
    L.append("#define CUBEEC_BS_PASS_CONFIGS(X) " + " ".join(passes))
    open(out_path, "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "bs_net_gen.cuh")
