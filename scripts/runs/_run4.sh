set -x
V=$PWD/personalized-pagerank-algorithms-on-neo4j_b200/variants
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2/parity4.log 2>&1; echo "rc=$?" >> gpurun_out/r2/parity4.log
for lib in v768 v384; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2/parity4_$lib.log 2>&1; echo "rc=$?" >> gpurun_out/r2/parity4_$lib.log; done
rm -f gpurun_out/r2/ab4.log
timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab4.log 2>&1
for lib in v768 v384 v256 v1024; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab4.log 2>&1; done
PPRB200_LIB=$V/libpprb200_v768.so PPR_MIDI_CTAS=0 timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab4.log 2>&1
PPRB200_LIB=$V/libpprb200_v768.so PPR_MIDI_V=1024 PPR_MIDI_E=32768 timeout 300 python scripts/push_ab.py 22 6 >> gpurun_out/r2/ab4.log 2>&1
PPRB200_LIB=$V/libpprb200_v768.so timeout 300 python scripts/push_ab.py 22 6 show > gpurun_out/r2/tl22_v768.log 2>&1
timeout 300 python scripts/push_ab.py 20 6 >> gpurun_out/r2/ab4.log 2>&1
for lib in v768 v384 v256; do PPRB200_LIB=$V/libpprb200_$lib.so timeout 300 python scripts/push_ab.py 20 6 >> gpurun_out/r2/ab4.log 2>&1; done
