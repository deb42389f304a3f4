set -x
N=${1:-8}
O=gpurun_out/r2mg$N
mkdir -p $O
nvidia-smi -L | head -8; free -g | head -2; nproc
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29611 scripts/dist_check.py > $O/dist.log 2>&1; echo dist=$?
grep -E "bit-exact|merge|walks/s|DIST_CHECK|communicator" $O/dist.log | tail -20
timeout -s KILL 600 $TR --master-port 29614 bench.py --gpus $N --cpu-baseline 0 > $O/bench.json 2> $O/bench.err; echo bench=$?
timeout -s KILL 300 $TR --master-port 29615 scripts/bench_configs.py sharded --scale 24 --edge-factor 16 --queries 4 > $O/sharded_s24.json 2> $O/sharded_s24.err; echo s24=$?
grep "^{" $O/sharded_s24.json | cut -c1-2500
timeout -s KILL 540 $TR --master-port 29616 scripts/bench_configs.py sharded --scale 25 --edge-factor 32 --queries 3 > $O/sharded_s25.json 2> $O/sharded_s25.err; echo s25=$?
grep "^{" $O/sharded_s25.json | cut -c1-2500; tail -3 $O/sharded_s25.err | cut -c1-300
grep "^{" $O/bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value']); print(json.dumps(d.get('sharded'))[:2500])"
