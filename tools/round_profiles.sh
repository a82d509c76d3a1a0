set -x
mkdir -p gpurun_out/p
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/p/pytest_gpu.log 2>&1; tail -2 gpurun_out/p/pytest_gpu.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p/launches_encode_crc.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/p/launches_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_bs_kernel -s 3 -c 1 -o /tmp/bs_crc python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/p/bs_crc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_bs_kernel -s 3 -c 1 -o /tmp/bs_nocrc python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --crc 0 > gpurun_out/p/bs_nocrc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rs_tabk -s 3 -c 1 -o /tmp/tabk_rec python bench.py --workload reconstruct --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/p/tabk.log 2>&1
for n in bs_crc bs_nocrc tabk_rec; do python tools/ncu_summary.py /tmp/$n.ncu-rep > gpurun_out/p/prof_$n.txt 2>&1; done
timeout 200 python bench.py > gpurun_out/p/bench_encode_crc.json 2> gpurun_out/p/bench_encode_crc.err
timeout 200 python bench.py --crc 0 --no-cpu > gpurun_out/p/bench_encode_nocrc.json 2>/dev/null
timeout 200 python bench.py --workload reconstruct > gpurun_out/p/bench_reconstruct.json 2>/dev/null
timeout 300 python bench.py --impl reference > gpurun_out/p/bench_reference_arm.json 2>/dev/null
timeout 300 python tools/sweep.py > gpurun_out/p/sweep.jsonl 2>/dev/null
timeout 300 python tools/sweep.py --modes > gpurun_out/p/sweep_modes.jsonl 2>/dev/null
timeout 100 ./tools/issue_mix > gpurun_out/p/issue_mix.log 2>&1
ls -la gpurun_out/p
