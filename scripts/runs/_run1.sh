set -x
V=personalized-pagerank-algorithms-on-neo4j_b200/variants
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2/parity_default.log 2>&1; echo "rc=$?" >> gpurun_out/r2/parity_default.log
PPRB200_LIB=$PWD/$V/libpprb200_c1s4.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2/parity_c1s4.log 2>&1; echo "rc=$?" >> gpurun_out/r2/parity_c1s4.log
for s in 22 20; do
for lib in r01 c1s4; do PPRB200_LIB=$PWD/$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py $s 12 >> gpurun_out/r2/ab.log 2>&1; done
timeout 300 python scripts/push_ab.py $s 12 show >> gpurun_out/r2/ab_default_$s.log 2>&1
for k in 0 16 64; do PPR_MIDI_CTAS=$k timeout 300 python scripts/push_ab.py $s 12 >> gpurun_out/r2/ab.log 2>&1; done
PPR_MIDI_CTAS=32 PPR_MIDI_V=8192 PPR_MIDI_E=131072 timeout 300 python scripts/push_ab.py $s 12 >> gpurun_out/r2/ab.log 2>&1
PPR_MIDI_CTAS=64 PPR_MIDI_V=131072 PPR_MIDI_E=2097152 timeout 300 python scripts/push_ab.py $s 12 >> gpurun_out/r2/ab.log 2>&1
done
tail -3 gpurun_out/r2/parity_default.log gpurun_out/r2/parity_c1s4.log; cat gpurun_out/r2/ab.log | cut -c1-400
