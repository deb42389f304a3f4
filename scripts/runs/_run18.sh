O=gpurun_out/r2p
mkdir -p $O
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4" 2>&1 | tail -3 | cut -c1-300; done
PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_default.log 2>&1
grep -h "^{\|\[pprb200\] reverse" $O/bwd_default.log | cut -c1-330
