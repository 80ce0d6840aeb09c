# usage: bash scripts/gpu_ncu.sh <out-name> <regex> <count> <skip> [<regex> <count> <skip> ...]
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --genome-mbp 16 --no-e2e --no-cpu-baseline --no-random-bench"
name=$1; shift
$B > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
i=0
while [ $# -ge 3 ]; do
  ncu --set full --clock-control none --import-source on -k "regex:$1" -s $3 -c $2 -f -o gpurun_out/${name}_$i $B > gpurun_out/ncu_${name}_$i.log 2>&1
  echo "ncu rc=$?"; tail -2 gpurun_out/ncu_${name}_$i.log
  shift 3; i=$((i+1))
done
