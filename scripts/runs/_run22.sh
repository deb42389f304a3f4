O=gpurun_out/r2u
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u/bench.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print({k:(round(v['seconds'],4),[round(x,4) for x in v['seconds_runs']],round(v['device_seconds'],4)) for k,v in d['allpairs_s']['thresholds'].items()}); print(d.get('extras_error'))
PY
