set -x
python bench.py --mode nsig --workload cfg3_96x10000_K20 --nsig-range 2,20 --chains 4 > gpurun_out/nsig20c3x4_1gpu_b.json 2>/dev/null
python bench.py --mode sharded --gpus 1 --steps 100 2> /dev/null | grep metric > gpurun_out/sharded_1gpu_b.json
for f in nsig20c3x4_1gpu_b sharded_1gpu_b; do cut -c1-400 gpurun_out/$f.json; done
