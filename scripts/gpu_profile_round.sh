# Round profile set: plain bench, ncu launch list of the same command, ncu --set full of the top kernels.
set -x
mkdir -p gpurun_out
R=${1:-r1b}
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench"
$B > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || { tail -5 gpurun_out/${R}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 12000 --csv --log-file gpurun_out/${R}_launches.csv $B > gpurun_out/${R}_ncu_launch.log 2>&1
echo "launch list rc=$?"
# top kernels at the bench's own size: probe (parents), insert into the L2 slice, binning, smem scan
ncu --set full --clock-control none --import-source on -k regex:k_stream -c 3 -f -o gpurun_out/${R}_k_stream $B > gpurun_out/${R}_ncu_a.log 2>&1; echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_update_keys -s 100 -c 2 -f -o gpurun_out/${R}_k_update_keys $B > gpurun_out/${R}_ncu_b.log 2>&1; echo "rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_bin_stream -c 1 -f -o gpurun_out/${R}_k_bin_stream $B > gpurun_out/${R}_ncu_c.log 2>&1; echo "rc=$?"
ls -la gpurun_out | grep ${R}
