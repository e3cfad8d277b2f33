export PYTHONPATH=.
N=${DKD_N:-2}
export DKD_BENCH_KERNEL_TABLE=gpurun_out/r4q_kernels
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r4q_bench_n$N.json 2> gpurun_out/r4q_bench_n$N.err
echo "rc=$?"
python tools/bench_table.py gpurun_out/r4q_bench_n$N.json
tail -c 500 gpurun_out/r4q_bench_n$N.err
