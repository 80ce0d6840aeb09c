# round 1d measurement set on one B200: full GPU test suite, bench lines, launch list, ncu captures
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1d_box.txt
nproc >> gpurun_out/r1d_box.txt; lscpu | grep "Model name" >> gpurun_out/r1d_box.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest_gpu.log
tail -3 gpurun_out/r1d_pytest_gpu.log
python bench.py > gpurun_out/r1d_bench_n1.json 2> gpurun_out/r1d_bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r1d_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1d_bench_reference.json 2> gpurun_out/r1d_bench_reference.err; echo "ref rc=$?"
for k in 21 47 63; do
  python bench.py --k $k --steps 5 --warmup 3 --no-cpu-baseline --no-random-bench > gpurun_out/r1d_bench_n1_k$k.json 2> gpurun_out/r1d_bench_n1_k$k.err; echo "k$k rc=$?"
done
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench"
$B > gpurun_out/r1d_bench_plain_under_profile_cmd.json 2> gpurun_out/r1d_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/r1d_launches.csv $B > gpurun_out/r1d_ncu_launch.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 4 -c 1 -f -o gpurun_out/r1d_k_stream $B > gpurun_out/r1d_ncu_a.log 2>&1; echo "ncu a rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_packed_keys -s 40 -c 2 -f -o gpurun_out/r1d_k_packed_keys $B > gpurun_out/r1d_ncu_b.log 2>&1; echo "ncu b rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_emit_packed -s 20 -c 1 -f -o gpurun_out/r1d_k_emit_packed $B > gpurun_out/r1d_ncu_c.log 2>&1; echo "ncu c rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_stream -s 6 -c 1 -f -o gpurun_out/r1d_k_stream_scan $B > gpurun_out/r1d_ncu_d.log 2>&1; echo "ncu d rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_bin_stream -s 2 -c 1 -f -o gpurun_out/r1d_k_bin_stream $B > gpurun_out/r1d_ncu_e.log 2>&1; echo "ncu e rc=$?"
ls -la gpurun_out | grep r1d
