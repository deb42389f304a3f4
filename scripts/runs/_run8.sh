set -x
O=gpurun_out/r2e
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pull_direction" > $O/pytest_pull.log 2>&1; echo "rc=$?" >> $O/pytest_pull.log
tail -25 $O/pytest_pull.log
for cfg in "-1 0" "0.3 0.2" "0.2 0.12" "0.4 0.3" "0.12 0.06" "0.6 0.45"; do set -- $cfg; PPR_PULL_ENTER=$1 PPR_PULL_EXIT=$2 timeout 300 python scripts/push_ab.py 22 12 >> $O/pull_ab22.log 2>&1; done
PPR_PULL_ENTER=0.2 PPR_PULL_EXIT=0.12 timeout 300 python scripts/push_ab.py 22 4 show > $O/tl22_pull.log 2>&1
for cfg in "-1 0" "0.3 0.2" "0.2 0.12"; do set -- $cfg; PPR_PULL_ENTER=$1 PPR_PULL_EXIT=$2 timeout 300 python scripts/push_ab.py 20 12 >> $O/pull_ab20.log 2>&1; done
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
cat $O/pull_ab22.log $O/pull_ab20.log | cut -c1-300
