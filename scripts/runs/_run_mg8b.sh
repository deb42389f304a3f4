set -x
N=8
O=gpurun_out/r2mg8b
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --master-port 29611 scripts/dist_check.py > $O/dist.log 2>&1; echo dist=$?
grep -E "bit-exact|merge|walks/s|DIST_CHECK|communicator" $O/dist.log | tail -8
timeout -s KILL 600 $TR --master-port 29614 bench.py --gpus $N --cpu-baseline 0 > $O/bench.json 2> $O/bench.err; echo bench=$?
grep "^{" $O/bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value']); s=d['sharded']; print(s['fora']); print({k:s['monte_carlo'][k] for k in ('value','ms_per_query')}); print({k:(v['seconds'],v['device_seconds'],v['gather_ms']) for k,v in s['allpairs']['thresholds'].items()}); print(d.get('extras_error'))"
