# 8-GPU measurements (run under `gpurun --gpus 8`): weak-scaling sweeps/s, nsig 2..20 model selection at 1 and 8 GPUs,
# column-sharded config 4 at 8 GPUs, and the multi-GPU tests.
set -x
python -m pytest tests/test_gpu_batch.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 1000 --warmup 100 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python bench.py --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20_1gpu.json 2> gpurun_out/nsig20_1gpu.err
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20_8gpu.json 2> gpurun_out/nsig20_8gpu.err
python bench.py --mode sharded --gpus 8 --steps 100 2> gpurun_out/sharded_8gpu.err | grep metric > gpurun_out/sharded_8gpu.json
python bench.py --mode sharded --gpus 4 --steps 100 2> /dev/null | grep metric > gpurun_out/sharded_4gpu.json
for f in bench_8gpu nsig20_1gpu nsig20_8gpu sharded_8gpu sharded_4gpu; do echo $f; cut -c1-260 gpurun_out/$f.json; done
