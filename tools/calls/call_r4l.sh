export PYTHONPATH=.
timeout 900 python -m pytest tests/test_saliency_gpu.py -m gpu -q -x > gpurun_out/r4l_sal.log 2>&1; tail -5 gpurun_out/r4l_sal.log
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "saliency" > gpurun_out/r4l_sal_big.log 2>&1; tail -3 gpurun_out/r4l_sal_big.log
timeout 300 python bench.py --workload saliency_mgd_m1_b512_f32 --no-cpu-baseline --steps 10 > gpurun_out/r4l_sal.json 2> gpurun_out/r4l_sal.err
python tools/bench_table.py gpurun_out/r4l_sal.json | tail -1; tail -c 300 gpurun_out/r4l_sal.err
