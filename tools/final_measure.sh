set -x
mkdir -p gpurun_out/fin
timeout 400 python bench.py > gpurun_out/fin/B_n1.json 2> gpurun_out/fin/B_n1.err
for c in REF C1 C3 D; do timeout 300 python bench.py --config $c --no-cpu > gpurun_out/fin/${c}_n1.json 2> gpurun_out/fin/${c}_n1.err; done
timeout 300 python bench.py --config E --no-cpu > gpurun_out/fin/E_exact_n1.json 2> gpurun_out/fin/E_exact_n1.err
CABM_BENCH_E_MODE=native timeout 300 python bench.py --config E --no-cpu > gpurun_out/fin/E_native_n1.json 2> gpurun_out/fin/E_native_n1.err
timeout 200 python bench.py --path fused-global --no-cpu --steps 10 > gpurun_out/fin/B_fused_global_n1.json 2> gpurun_out/fin/B_fg.err
timeout 300 python tools/time_workloads.py B REF HOT C1 CB1 > gpurun_out/fin/workload_times.jsonl 2>&1
timeout 200 python tools/diag_fused.py B > gpurun_out/fin/diag_B.txt 2>&1
timeout 200 python tools/diag_fused.py REF > gpurun_out/fin/diag_REF.txt 2>&1
timeout 200 python tools/diag_fused.py C1 > gpurun_out/fin/diag_C1.txt 2>&1
timeout 200 python tools/diag_fused.py E 100 > gpurun_out/fin/diag_E.txt 2>&1
ls -la gpurun_out/fin
