O=gpurun_out/r2t
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu -k "all_pairs or store or Base" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t/bench.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
print({k:(round(v['seconds'],4),round(v['device_seconds'],4)) for k,v in d['allpairs_s']['thresholds'].items()})
r=json.load(open('gpurun_out/r2t/bench_reference.json')); print('REF', r['value'], r['value_on_this_sample_mix'], r['wall_s'], r['seconds_per_query_with_out_edges'], r['share_of_sources_with_out_edges'])
PY
