set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4" > $O/pytest_bwd.log 2>&1; echo "rc=$?" >> $O/pytest_bwd.log
tail -5 $O/pytest_bwd.log
PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_order.log 2>&1
PPR_BWD_NO_ORDER=1 PPR_BWD_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_noorder.log 2>&1
timeout 600 python scripts/bwd_ab.py > $O/bwd_plain.log 2>&1
grep -h "per target\|^{" $O/bwd_order.log $O/bwd_noorder.log $O/bwd_plain.log | cut -c1-400
