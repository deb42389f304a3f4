O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu -k "batch or topk or top_k or metrics or harness" 2>&1 | tail -3 | cut -c1-300
timeout 600 python scripts/bench_configs.py topk --cpu 0 > $O/config3_topk_1gpu.json 2> $O/config3.err
timeout 900 python bench.py --cpu-baseline 0 --extras 0 > $O/bench_noextras.json 2> $O/bench.err
python - <<'PY'
import json
t=json.load(open('gpurun_out/r2r/config3_topk_1gpu.json')); print('topk', {k:t[k] for k in ('queries_per_s','ms_per_query','ms_push','ms_walk','ms_select','rounds')})
d=json.load(open('gpurun_out/r2r/bench_noextras.json')); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
PY
