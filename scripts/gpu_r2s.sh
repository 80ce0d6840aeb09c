# round 2s: the BAM-consuming GPU tests (golden pipelines) on the final decoder
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
timeout 100 python -m pytest tests/test_gpu_discovery.py tests/test_gpu_vcf.py -m gpu -x -q > gpurun_out/r2s_pytest_gpu_pipelines.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_gpu_pipelines.txt
tail -3 gpurun_out/r2s_pytest_gpu_pipelines.txt
