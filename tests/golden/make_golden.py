#!/usr/bin/env python3
"""Generate tests/golden/klauspost_tables.json and crc_golden.json FROM THE REFERENCE TREE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

What it extracts (no reference code is copied, only digests / constants the reference's
own sources and tests pin):
  * SHA-256 + length + first 16 entries of the GF(2^8) literal tables in
    vendor/github.com/klauspost/reedsolomon/galois.go (logTable :28, expTable :70,
    invTable :81, mulTable :83, mulTableLow :340, mulTableHigh :596,
    gf2p811dMulMatrices :937).  The oracle must regenerate byte-identical tables.
  * The 7 golden CRC32-IEEE values asserted in
    blobstore/blobnode/core/storage/datafile_test.go (:225..:396) with a description of
    the input each one checksums, re-verified here with zlib.
"""
import hashlib
import json
import os
import re
import zlib

REF = "/root/reference"
GALOIS = os.path.join(REF, "vendor/github.com/klauspost/reedsolomon/galois.go")
DATAFILE_TEST = os.path.join(REF, "blobstore/blobnode/core/storage/datafile_test.go")
HERE = os.path.dirname(os.path.abspath(__file__))


def go_literal(src: str, name: str):
    m = re.search(r"var %s = [^{]*\{" % re.escape(name), src)
    assert m, name
    i, depth = m.end(), 1
    while depth:
        ch = src[i]
        depth += ch == "{"
        depth -= ch == "}"
        i += 1
    body = src[m.end():i - 1]
    return [int(t, 0) for t in re.findall(r"0x[0-9a-fA-F]+|\b\d+\b", body)]


def main():
    src = open(GALOIS).read()
    tables = {}
    for name, width in (("logTable", 1), ("expTable", 1), ("invTable", 1), ("mulTable", 1),
                        ("mulTableLow", 1), ("mulTableHigh", 1), ("gf2p811dMulMatrices", 8)):
        vals = go_literal(src, name)
        raw = b"".join(v.to_bytes(width, "little") for v in vals)
        tables[name] = {"count": len(vals), "width": width,
                        "sha256": hashlib.sha256(raw).hexdigest(), "head": vals[:16]}
    poly = int(re.search(r"generatingPolynomial = (\d+)", src).group(1))
    json.dump({"source": "vendor/github.com/klauspost/reedsolomon/galois.go (v1.11.7)",
               "generatingPolynomial": poly, "tables": tables},
              open(os.path.join(HERE, "klauspost_tables.json"), "w"), indent=1)

    test_src = open(DATAFILE_TEST).read()
    asserted = [int(x) for x in re.findall(r"require\.Equal\(t, uint32\((\d+)\), shard\.Crc\)", test_src)]

    def zeros_12(n):
        b = bytearray(n)
        b[0] = ord("1")
        b[-1] = ord("2")
        return bytes(b)

    big = bytearray(1 << 20)
    for pos, ch in ((0, "1"), (65531, "2"), (65532, "3"), (131063, "4"), (131064, "5"), (196595, "6"), (1048575, "0")):
        big[pos] = ord(ch)
    cases = [
        ("ascii 'test data'", b"test data"),
        ("32768 B, b[i] = '0' + i%10", bytes(ord("0") + i % 10 for i in range(32768))),
        ("65492 zero bytes, first='1', last='2'", zeros_12(65492)),
        ("65532 zero bytes, first='1', last='2'", zeros_12(65532)),
        ("65536 zero bytes, first='1', last='2'", zeros_12(65536)),
        ("65530 zero bytes, first='1', last='2'", zeros_12(65530)),
        ("1 MiB zeros with marks at 0,65531,65532,131063,131064,196595,1048575", bytes(big)),
    ]
    out = []
    for desc, data in cases:
        crc = zlib.crc32(data)
        assert crc in asserted, (desc, crc)
        out.append({"input": desc, "crc32_ieee": crc})
    assert sorted(set(asserted)) == sorted(set(c["crc32_ieee"] for c in out))
    json.dump({"source": "blobstore/blobnode/core/storage/datafile_test.go:225-396", "cases": out},
              open(os.path.join(HERE, "crc_golden.json"), "w"), indent=1)
    print("wrote klauspost_tables.json, crc_golden.json")


if __name__ == "__main__":
    main()
