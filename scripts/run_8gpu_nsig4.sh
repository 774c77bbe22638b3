set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20c3_8gpu_b.json 2>/dev/null
$TR --nproc-per-node 8 --master-port 29615 bench.py --gpus 8 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 --chains 4 > gpurun_out/nsig20c3x4_8gpu.json 2>/dev/null
python bench.py --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 --chains 4 > gpurun_out/nsig20c3x4_1gpu.json 2>/dev/null
for f in nsig20c3_8gpu_b nsig20c3x4_1gpu nsig20c3x4_8gpu; do cut -c1-300 gpurun_out/$f.json; done
