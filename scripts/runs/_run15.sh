O=gpurun_out/r2m
mkdir -p $O
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_pairs or backward" 2>&1 | tail -3 | cut -c1-300; done
timeout 300 python scripts/bwd_ab.py 1e-3 2>&1 | tail -3 | cut -c1-300
timeout 300 python scripts/bwd_dbg.py big 2>&1 | tail -2
