timeout 900 python -m pytest tests/test_wass_gpu.py tests/test_saliency_gpu.py tests/test_curkd_gpu.py tests/test_sinkhorn_gpu.py tests/test_lrkd_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "wass or saliency" 2>&1 | tail -5
for W in wasskd_l1_b512_f32 saliency_mgd_m1_b512_f32; do
  timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 10 > gpurun_out/r3n_${W}.json 2> gpurun_out/r3n_${W}.err
  python tools/bench_table.py gpurun_out/r3n_${W}.json | tail -1
done
