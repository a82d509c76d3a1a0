#!/bin/bash
# A/B of the fold-table copies (kernels.cuh: CUBEEC_FC_LO / _HI / _CRC) on a GPU box.  Build the alternative library first:
#   cd cubefs_b200/csrc && for f in bitslice_flat bitslice_flat_mp crc_flat; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a \
#     -Xcompiler -fPIC -DCUBEEC_FC_LO=8 -DCUBEEC_FC_HI=8 -DCUBEEC_FC_CRC=8 -c $f.cu -o /tmp/alt_$f.o; done   (then link like the Makefile
#   does, with these three objects, into cubefs_b200/lib/libcubeec_alt.so)
run() {
  echo "== $1"
  timeout 100 python tools/ab_fused.py --stripes 1024,383 --force 0 2>&1 | cut -c1-90
  timeout 100 python tools/ab_fused.py --k 20 --m 4 --shard 1048576 --stripes 85 --force 0 2>&1 | tail -1 | cut -c1-90
  timeout 100 python tools/ab_fused.py --k 16 --m 4 --shard 1048576 --stripes 102 --force 0 2>&1 | tail -1 | cut -c1-90
  timeout 100 python tools/ab_fused.py --k 6 --m 3 --shard 1048576 --stripes 227 --force 0 2>&1 | tail -1 | cut -c1-90
  timeout 100 python tools/ab_fused.py --k 10 --m 4 --shard 1048576 --stripes 146 --force 0 2>&1 | tail -1 | cut -c1-90
  timeout 100 python tools/crc_speed.py 2>&1 | grep flat | cut -c1-170
}
run "libcubeec.so as built"
[ -f cubefs_b200/lib/libcubeec_alt.so ] || exit 0
cp cubefs_b200/lib/libcubeec_alt.so cubefs_b200/lib/libcubeec.so   # (on the GPU box's scratch copy of the repo)
run "libcubeec_alt.so"
