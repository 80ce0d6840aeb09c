# per-launch kernel times of one bench step (ncu, time metric only) -> gpurun_out/<name>_launches.csv
name=${1:-cur}; shift
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-random-bench $@"
$B > gpurun_out/${name}_plain.json 2> gpurun_out/${name}_plain.err || { tail -5 gpurun_out/${name}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 3000 --csv --log-file gpurun_out/${name}_launches.csv $B > gpurun_out/${name}_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/launch_shares.py gpurun_out/${name}_launches.csv | head -20
