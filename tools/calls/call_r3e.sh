set -x
timeout 300 python -m pytest tests/test_step_gpu.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_curkd_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k curkd_hidden 2>&1 | tail -8
timeout 300 python -m pytest tests/test_mgd_gpu.py -m gpu -q -x -k vitkd 2>&1 | tail -3
for F in 0 1; do
  for W in curkd_early_3layers_b512_f32 curkd_early_3layers_b512_bf16; do
    DKD_ALIGN_FUSED=$F timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 20 > gpurun_out/r3e_${W}_f$F.json 2> gpurun_out/r3e_${W}_f$F.err
    python tools/bench_table.py gpurun_out/r3e_${W}_f$F.json; tail -c 300 gpurun_out/r3e_${W}_f$F.err | grep -v Warn
  done
done
