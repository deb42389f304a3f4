set -x
V=$PWD/personalized-pagerank-algorithms-on-neo4j_b200/variants
O=gpurun_out/r2c
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu -k "all_pairs or backward or store or Base" > $O/pytest_bwd.log 2>&1; echo "rc=$?" >> $O/pytest_bwd.log
tail -15 $O/pytest_bwd.log
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "config4" > $O/pytest_cfg4.log 2>&1; echo "rc=$?" >> $O/pytest_cfg4.log
tail -5 $O/pytest_cfg4.log
PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_default.log 2>&1
for s in 37 74 111; do PPR_BWD_SLOTS=$s PPR_BWD_TIMING=1 timeout 600 python scripts/bwd_ab.py >> $O/bwd_slots.log 2>&1; done
PPR_BWD_BD=512 PPR_BWD_CPS=2 PPR_BWD_TIMING=1 timeout 600 python scripts/bwd_ab.py >> $O/bwd_slots.log 2>&1
for lib in c1 t768c1 t640c1 t1024c1 t384c2; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 22 12 >> $O/push_ab.log 2>&1; done
timeout 300 python scripts/push_ab.py 22 12 >> $O/push_ab.log 2>&1
for lib in c1 t768c1 t640c1 t1024c1 t384c2; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 20 12 >> $O/push_ab20.log 2>&1; done
timeout 300 python scripts/push_ab.py 20 12 >> $O/push_ab20.log 2>&1
PPRB200_LIB=$V/libpprb200_c1.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_c1.log 2>&1; echo "rc=$?" >> $O/pytest_c1.log
tail -3 $O/pytest_c1.log
grep -h "per target\|^{" $O/bwd_default.log $O/bwd_slots.log | cut -c1-330; cat $O/push_ab.log $O/push_ab20.log | cut -c1-200
