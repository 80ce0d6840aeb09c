# usage: bash scripts/gpu_ncu.sh <kernel-regex> <out-name> [count]
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --genome-mbp 16 --no-e2e --no-cpu-baseline --no-random-bench"
$B > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${4:-0} -c ${3:-2} -o gpurun_out/$2 $B > gpurun_out/ncu_$2.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_$2.log
