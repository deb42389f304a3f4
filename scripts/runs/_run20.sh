O=gpurun_out/r2s
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 3 --queries-per-step 10 --cpu-baseline 0 --extras 0 > $O/ncu_launches.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s/bench.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print({k:(round(v['seconds'],4),round(v['device_seconds'],4)) for k,v in d['allpairs_s']['thresholds'].items()})
r=json.load(open('gpurun_out/r2s/bench_reference.json')); print('REF', r['value'], r['wall_s'], r['seconds_per_query_with_out_edges'])
PY
