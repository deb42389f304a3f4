set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "all_pairs or backward or store or Base or config4 or pull" > $O/pytest_bwd.log 2>&1; echo "rc=$?" >> $O/pytest_bwd.log
tail -5 $O/pytest_bwd.log
PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_default.log 2>&1
PPR_BWD_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_timing.log 2>&1
for cfg in "32 4096 65536" "64 16384 262144" "96 65536 1048576" "148 131072 2097152" "64 4096 65536" "16 4096 65536"; do set -- $cfg; PPR_MIDI_CTAS=$1 PPR_MIDI_V=$2 PPR_MIDI_E=$3 timeout 300 python scripts/push_ab.py 22 12 >> $O/midi22.log 2>&1; done
for cfg in "32 4096 65536" "64 16384 262144" "96 65536 1048576"; do set -- $cfg; PPR_MIDI_CTAS=$1 PPR_MIDI_V=$2 PPR_MIDI_E=$3 timeout 300 python scripts/push_ab.py 20 12 >> $O/midi20.log 2>&1; done
timeout 600 python bench.py --steps 6 --warmup 3 --cpu-baseline 0 > $O/bench_quick.json 2> $O/bench_quick.err
grep -h "per target\|^{" $O/bwd_default.log $O/bwd_timing.log | cut -c1-330; cat $O/midi22.log $O/midi20.log | cut -c1-230; cut -c1-1500 $O/bench_quick.json
