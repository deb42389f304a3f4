timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4" 2>&1 | tail -3 | cut -c1-300
timeout 600 python scripts/bwd_ab.py 2>&1 | grep "^{" | cut -c1-140
