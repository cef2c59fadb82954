# End-of-round measurement batch (one GPU): bench lines of every workload, workload timings, per-phase diagnostics, then the ncu capture
# of the dominant kernel (after the same command has run plain), then the GPU test suite. Outputs under gpurun_out/fin/.
set -x
mkdir -p gpurun_out/fin
timeout 400 python bench.py > gpurun_out/fin/B_n1.json 2> gpurun_out/fin/B_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/fin/reference_arm.json 2> gpurun_out/fin/reference_arm.err
for c in REF C1 C3 D; do timeout 300 python bench.py --config $c --no-cpu > gpurun_out/fin/${c}_n1.json 2> gpurun_out/fin/${c}_n1.err; done
timeout 300 python bench.py --config E --no-cpu > gpurun_out/fin/E_exact_n1.json 2> gpurun_out/fin/E_exact_n1.err
CABM_BENCH_E_MODE=native timeout 300 python bench.py --config E --no-cpu > gpurun_out/fin/E_native_n1.json 2> gpurun_out/fin/E_native_n1.err
timeout 200 python bench.py --path fused-global --no-cpu --steps 10 > gpurun_out/fin/B_fused_global_n1.json 2> gpurun_out/fin/B_fg.err
timeout 300 python tools/time_workloads.py B REF HOT C1 CB1 > gpurun_out/fin/workload_times.jsonl 2>&1
for w in B REF C1; do timeout 200 python tools/diag_fused.py $w > gpurun_out/fin/diag_$w.txt 2>&1; done
timeout 200 python tools/diag_fused.py E 100 > gpurun_out/fin/diag_E.txt 2>&1
mkdir -p gpurun_out/ncu
timeout 200 python tools/one_launch.py B 1000 hm && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sim_fused -s 1 -c 1 -f -o gpurun_out/ncu/r02f_B python tools/one_launch.py B 1000 hm > gpurun_out/ncu/B.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/fin/gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin/smoke.txt 2>&1
ls -la gpurun_out/fin gpurun_out/ncu
