# round 2b: single push site in the packed count, batched filter loads (chunk 8), the new
# bench line (parity_checked, k sweep), traffic of every kernel, ncu of the three hot kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
tail -5 gpurun_out/r2b_pytest_gpu.log
Q="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $Q > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err || tail -3 gpurun_out/r2b_$name.err; }
run base KDF_X=0
run filt4 KDF_LIB=$PWD/build/libkdf_filt4.so
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench_n1.err
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r2b_launches.csv $B > gpurun_out/r2b_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_packed_keys -s 70 -c 1 -f -o gpurun_out/r2b_k_packed_keys $B > gpurun_out/r2b_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 5 -c 2 -f -o gpurun_out/r2b_k_stream $B > gpurun_out/r2b_ncu_c.log 2>&1; echo "ncu c rc=$?"
