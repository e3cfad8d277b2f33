timeout 900 python -m pytest tests/test_curkd_gpu.py tests/test_wass_gpu.py tests/test_diffkd_gpu.py tests/test_lrkd_gpu.py tests/test_sinkhorn_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "curkd_hidden or wass" 2>&1 | tail -4
for W in curkd_early_3layers_b512_f32 curkd_early_3layers_b512_bf16 wasskd_l1_b512_f32; do
  timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 20 > gpurun_out/r3j_${W}.json 2> gpurun_out/r3j_${W}.err
  python tools/bench_table.py gpurun_out/r3j_${W}.json | tail -1
done
