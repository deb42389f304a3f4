O=gpurun_out/r2n
mkdir -p $O
PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py > $O/bwd_team8.log 2>&1
PPR_BWD_TEAM=8 PPR_BWD_HEAVY=100000 PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 5e-4 5e-5 > $O/bwd_team8_all.log 2>&1
PPR_BWD_TEAM=4 PPR_BWD_HEAVY=100000 PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 5e-5 > $O/bwd_team4_all.log 2>&1
PPR_BWD_TEAM=2 PPR_BWD_HEAVY=100000 PPR_BWD_TIMING=1 PPR_DEBUG_TIMING=1 timeout 600 python scripts/bwd_ab.py 5e-5 > $O/bwd_team2_all.log 2>&1
for f in team8 team8_all team4_all team2_all; do echo == $f; grep -h "per target\|^{\|\[pprb200\] reverse" $O/bwd_$f.log | cut -c1-330; done
