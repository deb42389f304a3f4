N=$1
O=gpurun_out/r2mg${N}b
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --master-port 29614 bench.py --gpus $N --cpu-baseline 0 > $O/bench.json 2> $O/bench.err; echo bench=$?
grep "^{" $O/bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value']); s=d['sharded']; print(s['fora']['value'], s['monte_carlo']['value'], {k:(round(v['seconds'],3),round(v['device_seconds'],3)) for k,v in s['allpairs']['thresholds'].items()}); print(d.get('extras_error'))"
