#!/bin/bash
# Round-2 single-GPU evidence: GPU tests, both bench arms, launch list and one full ncu capture (run under gpurun).
set -x
O=gpurun_out/r2q
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python scripts/push_ab.py 22 12 show > $O/push_timeline_scale22.txt 2>&1
timeout 300 python scripts/push_ab.py 20 12 show > $O/push_timeline_scale20.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --queries-per-step 10 --cpu-baseline 0 --extras 0 > $O/ncu_launches.log 2>&1
python scripts/one_query.py 22 4 > $O/oq22.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_push_query|k_walk" -s 4 -c 4 -o $O/full_s22 -f python scripts/one_query.py 22 4 > $O/ncu_s22.log 2>&1
ls -la $O
tail -3 $O/pytest_gpu.log; cat $O/bench.json | cut -c1-3000
timeout 600 python scripts/bench_configs.py topk --cpu 0 > $O/config3_topk_1gpu.json 2> $O/config3.err
PPR_DEBUG_TIMING=1 PPR_BWD_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_ab_final.log 2>&1
python scripts/bwd_ncu.py 5e-5 6000 > $O/bn2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bwd_cta -s 2 -c 2 -o $O/bwd_5e-5_final -f python scripts/bwd_ncu.py 5e-5 6000 > $O/ncu_bwd.log 2>&1
ls -la $O
