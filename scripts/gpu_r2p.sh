# round 2p: last check of the committed state — GPU tests, smoke(), the default bench line
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_gpu.txt
tail -3 gpurun_out/r2p_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.txt
KDF_BAM_TIMING=1 timeout 1200 python bench.py > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['parity_checked']['ok'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('l2_mixed_rate_fraction'), d['roofline']['traffic'])
w=d['discovery_wall']; print(w['wall_s'], w['first_run_wall_s'], w['stages_s'], w['vcf_mode']['wall_s'], w['vcf_mode']['first_run_wall_s'])
PY
