# round 2a: new binning kernel (8192-key rounds, flat copy-out, pass filter), chunked
# extraction, split-phase CAS ring in the packed count, sort-based reduce_hits.
# GPU tests, bench line, A/B of the switches, launch list, ncu of the two child-count kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > gpurun_out/r2a_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -15 gpurun_out/r2a_pytest_gpu.log
Q="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-random-bench"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $Q > gpurun_out/r2a_$name.json 2> gpurun_out/r2a_$name.err || tail -3 gpurun_out/r2a_$name.err
  python - $name <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2a_%s.json'%v).read().strip().splitlines()[-1])
    print("%-10s %.2f G/s %.2f ms | "%(v,d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:12]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
except Exception as e: print(v,"ERR",e)
PY
}
run base KDF_X=0
run nosplit KDF_PQ_SPLIT=0
run div6 KDF_COUNT_SLOTS_DIV=6
run div5 KDF_COUNT_SLOTS_DIV=5
run b1024w8 KDF_BIN_THREADS=1024 KDF_BIN_WPR=8
run b512w8 KDF_BIN_THREADS=512 KDF_BIN_WPR=8
run b256w16 KDF_BIN_THREADS=256 KDF_BIN_WPR=16
timeout 300 python bench.py $Q --k 47 > gpurun_out/r2a_k47.json 2> gpurun_out/r2a_k47.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2a_k47.json').read().strip().splitlines()[-1])
    print("k47 %.2f G/s %.2f ms | "%(d['value']/1e9,d['ms_per_step'])+" ".join("%s=%.2f"%(k.split('/')[0][:12]+k[-4:],x['ms_total']/d['steps']) for k,x in d['kernels'].items() if x['ms_total']/d['steps']>0.5))
except Exception as e: print("k47 ERR",e)
PY
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
tail -2 gpurun_out/r2a_bench_n1.err
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r2a_launches.csv $B > gpurun_out/r2a_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bin_stream -s 2 -c 1 -f -o gpurun_out/r2a_k_bin_stream $B > gpurun_out/r2a_ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_packed_keys -s 70 -c 1 -f -o gpurun_out/r2a_k_packed_keys $B > gpurun_out/r2a_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 5 -c 1 -f -o gpurun_out/r2a_k_stream_scan $B > gpurun_out/r2a_ncu_c.log 2>&1; echo "ncu c rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
