O=gpurun_out/r2z
mkdir -p $O
python scripts/bwd_ncu.py 5e-5 8000 > $O/bn.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bwd_cta -s 2 -c 2 -o $O/bwd_5e-5_dense -f python scripts/bwd_ncu.py 5e-5 8000 > $O/ncu_bwd.log 2>&1
cat $O/bn.log; ls -la $O
