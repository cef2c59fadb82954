set -x
mkdir -p gpurun_out/ncu
timeout 200 python tools/one_launch.py B 1000 hm || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_fused -s 1 -c 1 -f -o gpurun_out/ncu/r02f_B python tools/one_launch.py B 1000 hm > gpurun_out/ncu/B.log 2>&1
timeout 200 python tools/one_launch.py E 100 nohm exact || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sim_fused -s 1 -c 1 -f -o gpurun_out/ncu/r02f_E python tools/one_launch.py E 100 nohm exact > gpurun_out/ncu/E.log 2>&1
ls -la gpurun_out/ncu
