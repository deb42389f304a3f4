set -x
O=gpurun_out/r2d
mkdir -p $O
python scripts/bwd_ncu.py 1e-3 20000 > $O/bn1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bwd_cta -s 1 -c 1 -o $O/bwd_1e-3 -f python scripts/bwd_ncu.py 1e-3 20000 > $O/ncu1.log 2>&1
python scripts/bwd_ncu.py 5e-5 4000 > $O/bn2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bwd_cta -s 1 -c 1 -o $O/bwd_5e-5 -f python scripts/bwd_ncu.py 5e-5 4000 > $O/ncu2.log 2>&1
cat $O/bn1.log $O/bn2.log; ls -la $O
