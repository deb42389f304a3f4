set -x
V=$PWD/personalized-pagerank-algorithms-on-neo4j_b200/variants
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu -k "all_pairs or backward or store or Base" > $O/pytest_bwd.log 2>&1; echo "rc=$?" >> $O/pytest_bwd.log
tail -15 $O/pytest_bwd.log
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "config4" > $O/pytest_cfg4.log 2>&1; echo "rc=$?" >> $O/pytest_cfg4.log
tail -5 $O/pytest_cfg4.log
PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_default.log 2>&1
for cfg in "128 12" "128 8" "256 4" "256 6" "512 2" "512 3" "1024 1" "64 16"; do set -- $cfg; PPR_BWD_BD=$1 PPR_BWD_CPS=$2 timeout 600 python scripts/bwd_ab.py >> $O/bwd_ab.log 2>&1; done
for lib in plain4 plain8 plain4c1 c1; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 22 12 >> $O/push_ab.log 2>&1; done
timeout 300 python scripts/push_ab.py 22 12 >> $O/push_ab.log 2>&1
PPRB200_LIB=$V/libpprb200_plain4.so timeout 300 python scripts/push_ab.py 22 4 show > $O/tl22_plain4.log 2>&1
for lib in plain4 plain4c1; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 20 12 >> $O/push_ab20.log 2>&1; done
timeout 300 python scripts/push_ab.py 20 12 >> $O/push_ab20.log 2>&1
cat $O/bwd_default.log $O/bwd_ab.log | cut -c1-400; cat $O/push_ab.log $O/push_ab20.log | cut -c1-330
