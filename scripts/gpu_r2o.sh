# round 2o: host-side sweep of the discovery wall on the bench host (chunk size of the decoder
# pipeline, huge pages for its buffers, OpenMP wait policy) on one set of BAMs
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
D=/tmp/kdf_sweep
python scripts/wall_sweep.py --dir $D --make 2>/dev/null | tail -1
O=gpurun_out/r2o_wall_sweep.jsonl
: > $O
python scripts/wall_sweep.py --dir $D --label default 2>/dev/null | tail -1 >> $O
KDF_BAM_CHUNK_KB=32768 python scripts/wall_sweep.py --dir $D --label chunk32m 2>/dev/null | tail -1 >> $O
KDF_BAM_CHUNK_KB=16384 python scripts/wall_sweep.py --dir $D --label chunk16m 2>/dev/null | tail -1 >> $O
KDF_BAM_CHUNK_KB=8192 python scripts/wall_sweep.py --dir $D --label chunk8m 2>/dev/null | tail -1 >> $O
KDF_BAM_THP=1 python scripts/wall_sweep.py --dir $D --label thp 2>/dev/null | tail -1 >> $O
KDF_BAM_CHUNK_KB=16384 KDF_BAM_THP=1 python scripts/wall_sweep.py --dir $D --label chunk16m_thp 2>/dev/null | tail -1 >> $O
OMP_WAIT_POLICY=passive python scripts/wall_sweep.py --dir $D --label omp_passive 2>/dev/null | tail -1 >> $O
KDF_PACK_SCALAR=1 python scripts/wall_sweep.py --dir $D --label pack_scalar 2>/dev/null | tail -1 >> $O
KDF_BAM_ZLIB=1 KDF_CRC_ZLIB=1 python scripts/wall_sweep.py --dir $D --label zlib 2>/dev/null | tail -1 >> $O
python scripts/wall_sweep.py --dir $D --label default_again 2>/dev/null | tail -1 >> $O
python - <<'PY'
import json
for l in open('gpurun_out/r2o_wall_sweep.jsonl'):
    try:
        d=json.loads(l); print("%-16s %s  child %.3f" % (d['label'], d['wall_s'], d['best_stages'].get('child_decode_and_count_s', -1)))
    except Exception as e: print("ERR", l[:200])
PY
