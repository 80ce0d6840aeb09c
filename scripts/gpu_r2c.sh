# round 2c: lean packed count (no software pipeline, higher occupancy) in four shapes vs the
# pipelined kernel; predicated batched filter loads; bin kernel specialisations
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log
tail -5 gpurun_out/r2c_pytest_gpu.log
Q="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
run() { name=$1; shift; env "$@" timeout 300 python bench.py $Q > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err || tail -3 gpurun_out/r2c_$name.err; }
run lean23 KDF_COUNT_LEAN=23
run lean24 KDF_COUNT_LEAN=24
run lean43 KDF_COUNT_LEAN=43
run lean42 KDF_COUNT_LEAN=42
run lean0 KDF_COUNT_LEAN=0
run smem2 KDF_LIB=$PWD/build/libkdf_smem2.so
timeout 300 python bench.py $Q --k 47 > gpurun_out/r2c_k47.json 2> gpurun_out/r2c_k47.err
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench --no-parity --no-k-sweep --no-wall"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_packed_count_lean -s 70 -c 1 -f -o gpurun_out/r2c_k_count_lean $B > gpurun_out/r2c_ncu_a.log 2>&1; echo "ncu a rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 5 -c 1 -f -o gpurun_out/r2c_k_stream $B > gpurun_out/r2c_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bin_stream -s 2 -c 1 -f -o gpurun_out/r2c_k_bin_stream $B > gpurun_out/r2c_ncu_c.log 2>&1; echo "ncu c rc=$?"
