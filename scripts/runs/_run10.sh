set -x
V=$PWD/personalized-pagerank-algorithms-on-neo4j_b200/variants
O=gpurun_out/r2g
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_parity.py -x -q -m gpu -k "metrics or topk or top_k or harness or cpp_host" > $O/pytest_metrics.log 2>&1; echo "rc=$?" >> $O/pytest_metrics.log
tail -25 $O/pytest_metrics.log | cut -c1-250
for lib in hot8k hot16k pa8 pa2; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 22 12 >> $O/push_ab.log 2>&1; done
timeout 300 python scripts/push_ab.py 22 12 >> $O/push_ab.log 2>&1
for lib in hot8k hot16k; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 20 12 >> $O/push_ab20.log 2>&1; done
timeout 300 python scripts/push_ab.py 20 12 >> $O/push_ab20.log 2>&1
PPR_DEBUG_TIMING=1 timeout 900 python bench.py --steps 4 --warmup 3 --cpu-baseline 0 > $O/bench_quick.json 2> $O/bench_quick.err
cat $O/push_ab.log $O/push_ab20.log | cut -c1-200; tail -40 $O/bench_quick.err; python -c "
import json; d=json.load(open('$O/bench_quick.json')); print(d.get('accuracy')); print(d['allpairs_s']); print(d.get('extras_error'))"
