"""Test-side mirror of the Go host layer that stays Go in production (blobstore/common/{ec,codemode,crc32block}):
C++ above the C-ABI + a Python view, so that the reference's encoder_test.go can be replayed here without a Go toolchain.
Nothing in the product package (cubefs_b200/) depends on it."""
