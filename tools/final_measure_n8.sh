# End-of-round 8-GPU lines of config E (exact and native schedule). Outputs under gpurun_out/fin8/.
set -x
mkdir -p gpurun_out/fin8
P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $P bench.py --gpus 8 --config E --no-cpu > gpurun_out/fin8/E_exact_n8.json 2> gpurun_out/fin8/E_exact_n8.err
CABM_BENCH_E_MODE=native timeout 300 $P bench.py --gpus 8 --config E --no-cpu > gpurun_out/fin8/E_native_n8.json 2> gpurun_out/fin8/E_native_n8.err
ls -la gpurun_out/fin8
