set -x
V=$PWD/personalized-pagerank-algorithms-on-neo4j_b200/variants
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2/parity3.log 2>&1; echo "rc=$?" >> gpurun_out/r2/parity3.log
timeout 300 python scripts/push_ab.py 22 6 show > gpurun_out/r2/tl22_v1.log 2>&1
for lib in t256c3 t384c2 t512c1 tma; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab3.log 2>&1; done
PPR_MIDI_CTAS=0 timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab3.log 2>&1
PPR_MIDI_CTAS=64 timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab3.log 2>&1
PPR_MIDI_CTAS=64 PPR_MIDI_V=16384 PPR_MIDI_E=262144 timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab3.log 2>&1
timeout 300 python scripts/push_ab.py 20 6 show > gpurun_out/r2/tl20_v1.log 2>&1
for lib in t256c3 t384c2 t512c1; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 20 6 >> gpurun_out/r2/ab3.log 2>&1; done
