for sd in 20 30 45 60 90; do echo split $sd; CABM_FUSED_SPLIT_DAY=$sd python tools/time_workloads.py B 2>&1 | tail -1 | cut -c1-120; done
