# round 2j: the state the round ends with — GPU tests, smoke(), the default bench line, the reference
# arm, launch list with DRAM bytes, ncu of the hot kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,driver_version --format=csv > gpurun_out/r2j_box.txt; nproc >> gpurun_out/r2j_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_gpu.txt
tail -4 gpurun_out/r2j_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2j_smoke.txt
timeout 1200 python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2j_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; echo "ref rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
$B > gpurun_out/r2j_bench_plain_under_profile_cmd.json 2> gpurun_out/r2j_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r2j_launches.csv $B > gpurun_out/r2j_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_packed_keys -s 70 -c 1 -f -o gpurun_out/r2j_k_packed_keys $B > gpurun_out/r2j_ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 5 -c 2 -f -o gpurun_out/r2j_k_stream $B > gpurun_out/r2j_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bin_stream -s 2 -c 1 -f -o gpurun_out/r2j_k_bin_stream $B > gpurun_out/r2j_ncu_c.log 2>&1; echo "ncu c rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_emit_packed -s 40 -c 1 -f -o gpurun_out/r2j_k_emit_packed $B > gpurun_out/r2j_ncu_d.log 2>&1; echo "ncu d rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['parity_checked']['ok'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('l2_mixed_rate_fraction'), d['roofline']['traffic'])
print(d['random_access'])
print(json.dumps(d['discovery_wall'])[:1500])
r=json.loads(open('gpurun_out/r2j_bench_reference.json').read().strip().splitlines()[-1])
print("reference", r['value']/1e9, r['ran'], r['cpu_baseline']['cores'])
PY
