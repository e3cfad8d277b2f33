set -x
timeout 600 python -m pytest tests/test_curkd_gpu.py tests/test_wass_gpu.py tests/test_diffkd_gpu.py tests/test_lrkd_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "curkd_hidden or wass" 2>&1 | tail -6
for C in 1 3; do
  for W in curkd_early_3layers_b512_f32 curkd_early_3layers_b512_bf16; do
    DKD_ALIGN_WGRAD_CLUSTER=$C timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 20 > gpurun_out/r3h_${W}_c$C.json 2> gpurun_out/r3h_${W}_c$C.err
    python tools/bench_table.py gpurun_out/r3h_${W}_c$C.json; tail -c 300 gpurun_out/r3h_${W}_c$C.err | grep -v Warn | grep -v run_backward
  done
done
