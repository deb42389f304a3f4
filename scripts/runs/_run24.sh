for mb in 0 16 24 32; do echo "== PPR_WALK_PERSIST=$mb"; PPR_WALK_PERSIST=$mb timeout 300 python scripts/push_ab.py 22 12 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ms_walk', round(d['ms_walk'],4), 'walks/s %.3g' % d['walks_per_s'], 'ms_push', round(d['ms_push'],3))"; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch" 2>&1 | tail -2
