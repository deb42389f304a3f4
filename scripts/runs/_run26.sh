O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4" 2>&1 | tail -6 | cut -c1-300
PPR_DEBUG_TIMING=1 PPR_BWD_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_default.log 2>&1
PPR_BWD_SCAN=1000000000000 timeout 600 python scripts/bwd_ab.py > $O/bwd_noscan.log 2>&1
PPR_BWD_SCAN=2048 timeout 600 python scripts/bwd_ab.py > $O/bwd_scan2k.log 2>&1
PPR_BWD_SCAN=50000 timeout 600 python scripts/bwd_ab.py > $O/bwd_scan50k.log 2>&1
for f in default noscan scan2k scan50k; do echo == $f; grep -h "per target\|^{" $O/bwd_$f.log | cut -c1-300; done
