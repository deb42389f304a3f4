set -x
O=gpurun_out/r2mg2
mkdir -p $O
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --master-port 29611 scripts/dist_check.py > $O/dist2.log 2>&1; echo dist=$?
grep -E "bit-exact|merge|walks/s|DIST_CHECK|communicator|Error|error" $O/dist2.log | tail -20
NCCL_DEBUG=WARN timeout -s KILL 900 $TR --master-port 29614 bench.py --gpus 2 --steps 6 --warmup 3 --cpu-baseline 0 > $O/bench_2.json 2> $O/bench_2.err; echo bench=$?
tail -5 $O/bench_2.err | cut -c1-300
python -c "
import json; d=json.load(open('$O/bench_2.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value']); print(json.dumps(d.get('sharded'))[:3000]); print(d.get('extras_error'))"
