O=gpurun_out/r2o
mkdir -p $O
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4" 2>&1 | tail -4 | cut -c1-300; done
PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_default.log 2>&1
PPR_BWD_NO_EDGE_INV=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_noinv.log 2>&1
PPR_BWD_TEAM_REST=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 5e-5 > $O/bwd_rest1.log 2>&1
PPR_BWD_TEAM_REST=4 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 5e-5 > $O/bwd_rest4.log 2>&1
PPR_BWD_HEAVY=512 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 1e-3 5e-4 > $O/bwd_h512.log 2>&1
PPR_BWD_HEAVY=32 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 1e-3 5e-4 > $O/bwd_h32.log 2>&1
timeout 600 python scripts/bwd_ab.py > $O/bwd_plain.log 2>&1
for f in default noinv rest1 rest4 h512 h32 plain; do echo == $f; grep -h "per target\|^{\|\[pprb200\] reverse" $O/bwd_$f.log | cut -c1-330; done
