set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20c3_1gpu.json 2> gpurun_out/nsig20c3_1gpu.err
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20c3_8gpu.json 2> gpurun_out/nsig20c3_8gpu.err
$TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20c3_2gpu.json 2>/dev/null
$TR --nproc-per-node 4 --master-port 29614 bench.py --gpus 4 --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 > gpurun_out/nsig20c3_4gpu.json 2>/dev/null
for f in nsig20c3_1gpu nsig20c3_2gpu nsig20c3_4gpu nsig20c3_8gpu; do cut -c1-330 gpurun_out/$f.json; done
