# LRKD: cluster-resident Jacobi vs the cooperative one (A/B), parity tests on both
timeout 900 python -m pytest tests/test_lrkd_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "lrkd" 2>&1 | tail -5
for C in 1 0; do
  DKD_LRKD_CLUSTER=$C timeout 300 python bench.py --workload lrkd_r64_b512_f32 --no-cpu-baseline --steps 10 > gpurun_out/r4a_lrkd_c$C.json 2> gpurun_out/r4a_lrkd_c$C.err
  python tools/bench_table.py gpurun_out/r4a_lrkd_c$C.json | tail -1
  tail -c 300 gpurun_out/r4a_lrkd_c$C.err
done
