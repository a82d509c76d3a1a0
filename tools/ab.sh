# usage: bash tools/ab.sh <variant lib suffix> [bench args...]   (A/B of cubefs_b200/lib/exp/libcubeec_<v>.so vs the in-tree build)
v=$1; shift
mkdir -p gpurun_out
P='import sys,json; j=json.loads(sys.stdin.read()); print(j.get("kernel"), j["value"], j["roofline"]["frac"], j["ms_per_step"], j["clocks"])'
for r in 1 2 3; do
  for x in base $v; do
    L=$PWD/cubefs_b200/lib/exp/libcubeec_$x.so; [ $x = base ] && L=$PWD/cubefs_b200/lib/libcubeec.so
    echo -n "$x: "; CUBEEC_LIB=$L timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu "$@" 2>gpurun_out/ab_err.log | tail -1 | python -c "$P" || tail -3 gpurun_out/ab_err.log
  done
done
