#!/bin/bash
# Sweep the experiment knobs of the fused kernel (read by launch_sim_fused from the environment) with bench.py:
#   CABM_FUSED_SPLIT_DAY   day after which sweep 0 parks the replicates that are still active (default 60)
#   CABM_FUSED_TAIL_CTAS   CTAs that copy finished curves to the host at any one time (default 8)
for sd in 20 35 60 90; do CABM_FUSED_SPLIT_DAY=$sd timeout 100 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('split_day', $sd, 'device', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; done
for tc in 4 8 16 32; do CABM_FUSED_TAIL_CTAS=$tc timeout 100 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('copiers', $tc, 'device', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"; done
