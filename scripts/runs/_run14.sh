O=gpurun_out/r2l
mkdir -p $O
for mode in r12 big; do for dbg in 2 3 0; do echo "== auto plan $mode dbg $dbg"; PPR_BWD_DBG=$dbg CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/bwd_dbg.py $mode 2>&1 | tail -2 | cut -c1-400; done; done
for team in 2 8; do echo "== forced team $team r12deep"; PPR_BWD_FORCE_TEAM=1 PPR_BWD_TEAM=$team CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/bwd_dbg.py r12deep 2>&1 | tail -2 | cut -c1-400; done
echo "== auto plan r12, team 2"; PPR_BWD_TEAM=2 timeout 120 python scripts/bwd_dbg.py r12 2>&1 | tail -2 | cut -c1-400
echo "== auto plan r12, heavy 4"; PPR_BWD_HEAVY=4 timeout 120 python scripts/bwd_dbg.py r12 2>&1 | tail -2 | cut -c1-400
