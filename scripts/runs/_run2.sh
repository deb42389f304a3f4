set -x
mkdir -p gpurun_out/r2
python scripts/one_query.py 22 3 > gpurun_out/r2/oneq22.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_push_query --launch-skip 2 --launch-count 1 -o gpurun_out/r2/push22_new python scripts/one_query.py 22 3 > gpurun_out/r2/ncu22.log 2>&1
ls -la gpurun_out/r2/
timeout 300 python scripts/push_ab.py 22 4 show > gpurun_out/r2/tl22_new.log 2>&1
PPRB200_LIB=$PWD/personalized-pagerank-algorithms-on-neo4j_b200/variants/libpprb200_r01.so timeout 300 python scripts/push_ab.py 22 4 show > gpurun_out/r2/tl22_r01.log 2>&1
