"""Helpers for the reference-run golden vectors (tests/golden/rs_golden.json, written by
tools/ref_harness/run.sh on a box with Go): the input generator of tools/ref_harness/main.go restated in
numpy, and the shard layout the harness builds (ec.Buffer + Split for code modes, plain slices for RS(k, m))."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rs_golden.json")
UNPINNED = ("parity unpinned: tests/golden/rs_golden.json is absent -- no Go toolchain in the build image; run "
            "tools/ref_harness/run.sh on a box with Go >= 1.18 and the cubefs tree to generate it")


def load():
    if not os.path.exists(GOLDEN):
        return None
    return json.load(open(GOLDEN))["vectors"]


def splitmix_bytes(seed: int, n: int) -> np.ndarray:
    """byte i = byte (i % 8) of the (i // 8)-th splitmix64 output (little endian), as main.go's fill()."""
    words = (n + 7) // 8
    M64 = (1 << 64) - 1
    with np.errstate(over="ignore"):
        s = (np.uint64(seed & M64) + np.arange(1, words + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = s
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8)[:n].copy()


def data_shards(v):
    """The N data shards of vector v as the harness built them (zero padded to shard_size)."""
    n, size = v["n"], v["shard_size"]
    flat = np.zeros(n * size, dtype=np.uint8)
    flat[:v["data_len"]] = splitmix_bytes(v["seed"], v["data_len"])
    return [flat[i * size:(i + 1) * size].copy() for i in range(n)]


def local_stripes(v):
    """codemode.GetECLayoutByAZ (blobstore/common/codemode/codemode.go:301-318): per AZ the indices of its
    data shards, global parity shards and local parity shards."""
    n, m, l, az = v["n"], v["m"], v["l"], max(v["az_count"], 1)
    out = []
    for a in range(az):
        d = list(range(a * n // az, (a + 1) * n // az))
        p = list(range(n + a * m // az, n + (a + 1) * m // az))
        lp = list(range(n + m + a * l // az, n + m + (a + 1) * l // az))
        out.append((d + p, lp))
    return out


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
