"""cubefs_b200 -- B200-native erasure-coding + shard-checksum engine for CubeFS BlobStore.

The product is the C-ABI library ``cubefs_b200/lib/libcubeec.so`` (include/cubeec.h); this
package is the Python host-side binding used by tests and bench.py:

  cubefs_b200.engine    ctypes binding of the C-ABI (RSEngine, crc32, device-resident calls)
  cubefs_b200.parallel  stripe partition + coding-matrix broadcast helpers for the one-process-per-GPU runs
  cubefs_b200.csrc      the CUDA / C++ sources of the library and the XOR-network generator

(The C++/Python mirror of the Go host layer that the reference's tests are replayed through lives in
tests/mirror/ -- it is test infrastructure, not product.)

There is no CPU compute path: importing works anywhere, calling needs a CUDA device.
"""
from .engine import (CubeecError, RSEngine, crc32, crc32_blocks, dev_lrc_encode, dev_lrc_reconstruct, dev_lrc_verify, device_count, force_kernel,  # noqa: F401
                     init, kernel_launches, last_kernel, lib_path, load, lrc_encode_contig, set_coalescing)
