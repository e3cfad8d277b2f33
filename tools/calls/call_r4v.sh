export PYTHONPATH=.
timeout 900 python -m pytest tests/test_mgd_gpu.py tests/test_saliency_gpu.py tests/test_vitkd_gpu.py -m gpu -q -x > gpurun_out/r4v_tests.log 2>&1; tail -3 gpurun_out/r4v_tests.log
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "mgd or saliency or late" > gpurun_out/r4v_big.log 2>&1; tail -2 gpurun_out/r4v_big.log
for SS in 1 0; do for W in mgd_b512_bf16 mgd_b512_f32 saliency_mgd_m1_b512_f32; do
  DKD_SIDE_STREAM=$SS timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 10 > gpurun_out/r4v_${W}_ss$SS.json 2> gpurun_out/r4v_${W}_ss$SS.err
  echo "SS=$SS $(python tools/bench_table.py gpurun_out/r4v_${W}_ss$SS.json | grep "^$W" | cut -c1-150)"
done; done
true
