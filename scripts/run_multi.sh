#!/bin/bash
# multi-GPU evidence run: NCCL parity check, walk sharding at a large scale, all-pair partition, bench scaling line
N=${1:-8}; SCALE=${2:-25}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
free -g | head -2; nproc
timeout -s KILL 600 $TR --master-port 29611 scripts/dist_check.py > gpurun_out/dist$N.log 2>&1; echo dist=$?; grep -E "bit-exact|merge|walks/s|DIST_CHECK" gpurun_out/dist$N.log | tail -9
timeout -s KILL 900 $TR --master-port 29612 scripts/bench_configs.py walks --scale $SCALE --cpu 0 > gpurun_out/walks${SCALE}_$N.json 2> gpurun_out/walks${SCALE}_$N.err; echo walks=$?; tail -2 gpurun_out/walks${SCALE}_$N.err | cut -c1-300; cat gpurun_out/walks${SCALE}_$N.json
timeout -s KILL 600 $TR --master-port 29613 scripts/bench_configs.py allpairs --cpu 0 > gpurun_out/allpairs_$N.json 2> gpurun_out/allpairs_$N.err; echo ap=$?; cat gpurun_out/allpairs_$N.json
timeout -s KILL 600 $TR --master-port 29615 scripts/bench_configs.py topk --cpu 0 > gpurun_out/topk_$N.json 2> gpurun_out/topk_$N.err; echo topk=$?; cat gpurun_out/topk_$N.json
timeout -s KILL 600 $TR --master-port 29614 bench.py --gpus $N --cpu-baseline 0 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo bench=$?; grep "^{" gpurun_out/bench_$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])"
