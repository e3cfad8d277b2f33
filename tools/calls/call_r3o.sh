export DKD_BENCH_KERNEL_TABLE=gpurun_out/r3o_kernels
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3o_bench_n8.json 2> gpurun_out/r3o_bench_n8.err
echo "rc=$?"
python tools/bench_table.py gpurun_out/r3o_bench_n8.json
tail -c 800 gpurun_out/r3o_bench_n8.err
ls gpurun_out/r3o_kernels*
