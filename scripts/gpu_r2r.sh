# round 2r: discovery wall with the headers classified by the walker (default) and after the barrier
set -x
mkdir -p gpurun_out
python -c "from kmer_denovo_filter_b200 import engine; engine.load_library(); print('lib ok')" || exit 1
D=/tmp/kdf_sweep
python scripts/wall_sweep.py --dir $D --make 2>/dev/null | tail -1
O=gpurun_out/r2r_wall_sweep.jsonl
: > $O
python scripts/wall_sweep.py --dir $D --label default --reps 4 2>/dev/null | tail -1 >> $O
KDF_BAM_WALK_CLASSIFY=0 python scripts/wall_sweep.py --dir $D --label cold_classify --reps 4 2>/dev/null | tail -1 >> $O
KDF_BAM_TIMING=1 python scripts/wall_sweep.py --dir $D --label default_again --reps 4 2> gpurun_out/r2r_timing.err | tail -1 >> $O
grep kdf_bam gpurun_out/r2r_timing.err | tail -3 | cut -c1-240
python - <<'PY'
import json
for l in open('gpurun_out/r2r_wall_sweep.jsonl'):
    d=json.loads(l); print("%-16s %s  child %.3f" % (d['label'], d['wall_s'], d['best_stages'].get('child_decode_and_count_s', -1)))
PY
