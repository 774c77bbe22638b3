# Round-2 final multi-GPU measurements (run under `gpurun --gpus 8`): the driver's line at N = 8; ONE chain column-sharded
# over 8, 4, 2 GPUs with the per-shard breakdown; model selection of the small cohort on 8 GPUs.
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
O=gpurun_out
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02y_bench_n8.json 2> $O/r02y_bench_n8.err
for n in 8 4 2; do SIGNER_TRACE=1 timeout 200 python bench.py --mode sharded --gpus $n --steps 100 > $O/r02y_sharded_n$n.json 2> $O/r02y_sharded_n$n.err; done
$TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --mode nsig --workload cfg2_96x560 --nsig-range 2,10 --chains 4 > $O/r02y_nsig_cfg2_n8.json 2> $O/r02y_nsig_cfg2_n8.err
for f in bench_n8 sharded_n8 sharded_n4 sharded_n2 nsig_cfg2_n8; do echo == $f; tail -1 $O/r02y_$f.json | cut -c1-300; done
grep -h "shard\|loop" $O/r02y_sharded_n8.err | tail -12
