#!/bin/bash
# Regenerates the single-GPU evidence under profiles/ (run on the GPU box through gpurun; outputs land in gpurun_out/).
set -x
O=gpurun_out
python bench.py > $O/bench_scale20.json 2> $O/bench_scale20.err || exit 1
python scripts/timeline.py 20 1 > $O/push_timeline_scale20.txt 2>&1
python scripts/timeline.py 22 1 > $O/push_timeline_scale22.txt 2>&1
python scripts/bench_configs.py topk > $O/config3_topk_1gpu.json 2> $O/config3.err
python scripts/bench_configs.py allpairs > $O/config4_allpairs_1gpu.json 2> $O/config4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
# ncu passes last, each after the plain command above has exited 0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_scale20.csv \
    python bench.py --steps 2 --warmup 3 --also-scale 0 --cpu-baseline 0 > $O/ncu_launches.log 2>&1
python scripts/one_query.py 22 4 > $O/oq22.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_push_query|k_walk" -s 4 -c 4 -o $O/r01b_full_s22 -f python scripts/one_query.py 22 4 > $O/ncu_s22.log 2>&1
ls -la $O
