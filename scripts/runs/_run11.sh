set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -x -q -m gpu -k "walk or monte or fora_got or scale21 or shards" > $O/pytest_walk.log 2>&1; echo "rc=$?" >> $O/pytest_walk.log
tail -5 $O/pytest_walk.log
for v in 1 2; do PPR_WALK_VARIANT=$v timeout 300 python scripts/push_ab.py 22 12 >> $O/walk_ab22.log 2>&1; done
for v in 0 2; do PPR_WALK_VARIANT=$v timeout 300 python scripts/push_ab.py 20 12 >> $O/walk_ab20.log 2>&1; done
python - <<'PY'
import json
for f in ('gpurun_out/r2i/walk_ab22.log','gpurun_out/r2i/walk_ab20.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f[-8:], 'ms_walk', round(d['ms_walk'],4), 'walks/s %.3g' % d['walks_per_s'], 'ms_push', round(d['ms_push'],3))
PY
