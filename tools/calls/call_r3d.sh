set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3d_bench_n2.json 2> gpurun_out/r3d_bench_n2.err
echo "rc=$?"
python tools/bench_table.py gpurun_out/r3d_bench_n2.json
tail -c 1500 gpurun_out/r3d_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 | tail -c 400
