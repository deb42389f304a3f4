set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4" > $O/pytest_bwd.log 2>&1; echo "rc=$?" >> $O/pytest_bwd.log
tail -15 $O/pytest_bwd.log | cut -c1-250
PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_team8.log 2>&1
PPR_BWD_TEAM=4 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_team4.log 2>&1
PPR_BWD_TEAM=2 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_team2.log 2>&1
PPR_BWD_TEAM=8 PPR_BWD_HEAVY=512 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 1e-3 5e-4 > $O/bwd_team8_h512.log 2>&1
PPR_BWD_TEAM=8 PPR_BWD_HEAVY=32 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 1e-3 5e-4 > $O/bwd_team8_h32.log 2>&1
PPR_BWD_TEAM=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_team1.log 2>&1
for f in team8 team4 team2 team8_h512 team8_h32 team1; do echo == $f; grep -h "per target\|^{\|\[pprb200\] reverse" $O/bwd_$f.log | cut -c1-300; done
