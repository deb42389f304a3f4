O=gpurun_out/r2x
mkdir -p $O
for cfg in "1 8" "4 8" "2 4" "2 2"; do set -- $cfg; echo "== TEAM_REST=$1 TEAM=$2"; PPR_BWD_TEAM_REST=$1 PPR_BWD_TEAM=$2 timeout 600 python scripts/bwd_ab.py 2>&1 | grep "^{" | cut -c1-150; done
for h in 0 32 512; do echo "== HEAVY=$h"; PPR_BWD_HEAVY=$h timeout 600 python scripts/bwd_ab.py 1e-3 5e-4 2>&1 | grep "^{" | cut -c1-150; done
for sc in 6000 25000; do echo "== SCAN=$sc"; PPR_BWD_SCAN=$sc timeout 600 python scripts/bwd_ab.py 2>&1 | grep "^{" | cut -c1-150; done
