#!/bin/bash
# Regenerates everything under profiles/ for one round in ONE GPU call:  bash tools/round_profiles.sh r02
# (launch list + ncu --set full summaries of the dominant kernels + bench JSONs + sweeps + GPU test log)
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
set -x
timeout 600 python -m pytest tests -q -m gpu > $O/${R}_pytest_gpu.log 2>&1; tail -3 $O/${R}_pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/launches_run.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-check"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_bsf_kernel -s 3 -c 1 -o /tmp/bsf_crc $B --no-extra > $O/ncu_bsf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rs_bs_kernel" -s 3 -c 1 -o /tmp/bs_nocrc $B --no-extra --crc 0 > $O/ncu_bs.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_tabk -s 3 -c 1 -o /tmp/tabk_rec $B --no-extra --workload reconstruct > $O/ncu_tabk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_jit -s 3 -c 1 -o /tmp/jit_rec $B > $O/ncu_jit.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_bssyn -s 3 -c 1 -o /tmp/syn_rec $B --no-extra --workload reconstruct --force 9 > $O/ncu_syn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crc_flat_kernel -s 3 -c 1 -o /tmp/crc_flat python tools/crc_speed.py > $O/ncu_crc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_bsf_kernel -s 3 -c 1 -o /tmp/bsf_c4 python tools/ab_fused.py --k 20 --m 4 --shard 1048576 --stripes 85 --force 0 --steps 3 > $O/ncu_bsf_c4.log 2>&1
python tools/ncu_summary.py /tmp/bsf_crc.ncu-rep > $O/${R}_prof_bsf_crc.txt 2>&1
python tools/ncu_summary.py /tmp/bs_nocrc.ncu-rep > $O/${R}_prof_bs_nocrc.txt 2>&1
python tools/ncu_summary.py /tmp/tabk_rec.ncu-rep > $O/${R}_prof_tabk_rec.txt 2>&1
python tools/ncu_summary.py /tmp/jit_rec.ncu-rep > $O/${R}_prof_jit_rec.txt 2>&1
python tools/ncu_summary.py /tmp/syn_rec.ncu-rep > $O/${R}_prof_bssyn_cpasync.txt 2>&1
python tools/ncu_summary.py /tmp/crc_flat.ncu-rep > $O/${R}_prof_crc_flat.txt 2>&1
python tools/ncu_summary.py /tmp/bsf_c4.ncu-rep > $O/${R}_prof_bsf_crc_rs20_4.txt 2>&1
timeout 200 python tools/crc_speed.py > $O/${R}_crc_speed.jsonl 2>/dev/null
timeout 300 python tools/ab_fused.py --stripes 1024,383 --force 0,3001,3002,7 > $O/${R}_ab_fused_c2.jsonl 2>/dev/null
timeout 300 python tools/sweep.py --modes --forces 0,3001,3002 2>/dev/null | grep -v '"crc": false' > $O/${R}_sweep_bsf_entry_sync.jsonl
timeout 400 python bench.py > $O/${R}_bench_encode_crc.json 2> $O/bench.err; tail -2 $O/bench.err
timeout 300 python bench.py --workload reconstruct --no-e2e > $O/${R}_bench_reconstruct.json 2>/dev/null
timeout 300 python bench.py --workload reconstruct --no-e2e --no-cpu --no-extra --force 9 > $O/${R}_bench_reconstruct_bssyn.json 2>/dev/null
timeout 300 python bench.py --impl reference > $O/${R}_bench_reference_arm.json 2>/dev/null
timeout 400 python tools/sweep.py > $O/${R}_sweep_c4_c5.jsonl 2>/dev/null
timeout 400 python tools/sweep.py --modes > $O/${R}_sweep_code_modes.jsonl 2>/dev/null
ls -la $O
