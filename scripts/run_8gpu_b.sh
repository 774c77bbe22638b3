# Round-1 refresh after the K1 v7/v8 work (run under `gpurun --gpus 8`): multi-GPU tests, weak-scaling sweeps/s,
# nsig 2..20 x 4 chains model selection on 8 GPUs.  The 1-GPU reference of the latter: scripts/run_1gpu_nsig4.sh.
set -x
python -m pytest tests/test_gpu_batch.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 1000 --warmup 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_8gpu_b.json 2> gpurun_out/bench_8gpu_b.err
$TR --nproc-per-node 8 --master-port 29615 bench.py --gpus 8 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 --chains 4 > gpurun_out/nsig20c3x4_8gpu_b.json 2>/dev/null
for f in bench_8gpu_b nsig20c3x4_8gpu_b; do cut -c1-400 gpurun_out/$f.json; done
