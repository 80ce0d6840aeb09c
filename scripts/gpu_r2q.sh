# round 2q: GPU tests + smoke on the final commit
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_gpu.txt
tail -3 gpurun_out/r2q_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2q_smoke.txt
