# usage: bash tools/ab_kernel.sh <bench --kernel value>   (A/B of an opt-in kernel variant against the default, same build)
mkdir -p gpurun_out
P='import sys,json; j=json.loads(sys.stdin.read()); print(j.get("kernel"), j["value"], j["roofline"]["frac"], j["ms_per_step"])'
for r in 1 2; do
  for x in auto $1; do
    echo -n "$x: "; timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --kernel $x 2>gpurun_out/ab_err.log | tail -1 | python -c "$P" || tail -3 gpurun_out/ab_err.log
  done
done
