set -x
timeout 600 python -m pytest tests/test_curkd_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k curkd_hidden 2>&1 | tail -6
timeout 300 python -m pytest tests/test_mgd_gpu.py -m gpu -q -x -k vitkd 2>&1 | tail -3
for W in curkd_early_3layers_b512_f32 curkd_early_3layers_b512_bf16; do
  timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 20 > gpurun_out/r3f_${W}.json 2> gpurun_out/r3f_${W}.err
  python tools/bench_table.py gpurun_out/r3f_${W}.json; tail -c 300 gpurun_out/r3f_${W}.err | grep -v Warn
done
NCU_WORKLOADS="curkd_early_3layers_b512_f32 curkd_early_3layers_b512_bf16" bash tools/gpu_r3.sh r3f o
