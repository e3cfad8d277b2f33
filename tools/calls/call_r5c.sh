export PYTHONPATH=.
timeout 900 python -m pytest tests/test_sinkhorn_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "sinkhorn" 2>&1 | tail -2
timeout 300 python bench.py --workload wasskd_sinkhorn_b512_f32 --no-cpu-baseline --steps 10 > gpurun_out/r5c_sink.json 2> gpurun_out/r5c_sink.err
python tools/bench_table.py gpurun_out/r5c_sink.json | tail -1 | cut -c1-160
