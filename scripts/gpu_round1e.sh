# round 1e (after the table filter): GPU tests, default bench line, launch list, one ncu capture of the dominant kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1e_pytest_gpu.log
tail -3 gpurun_out/r1e_pytest_gpu.log
python bench.py > gpurun_out/r1e_bench_n1.json 2> gpurun_out/r1e_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r1e_bench_n1.err
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench"
$B > gpurun_out/r1e_bench_plain_under_profile_cmd.json 2> gpurun_out/r1e_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r1e_launches.csv $B > gpurun_out/r1e_ncu_launch.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 4 -c 1 -f -o gpurun_out/r1e_k_stream $B > gpurun_out/r1e_ncu_a.log 2>&1; echo "ncu a rc=$?"
