export PYTHONPATH=.
timeout 900 python -m pytest tests/test_curkd_gpu.py tests/test_wass_gpu.py tests/test_sinkhorn_gpu.py tests/test_host_contract_gpu.py -m gpu -q -x > gpurun_out/r4s_tests.log 2>&1; tail -3 gpurun_out/r4s_tests.log
for LS in 1 0; do for W in curkd_early_3layers_b512_f32 curkd_mid_4layers_b512_f32 curkd_early_3layers_b512_bf16 wasskd_l1_b512_f32 wasskd_sinkhorn_b512_f32; do
  DKD_LAYER_STREAMS=$LS timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 10 > gpurun_out/r4s_${W}_ls$LS.json 2> gpurun_out/r4s_${W}_ls$LS.err
  echo "LS=$LS $(python tools/bench_table.py gpurun_out/r4s_${W}_ls$LS.json | grep $W | cut -c1-140)"; tail -c 200 gpurun_out/r4s_${W}_ls$LS.err | grep -i "error\|Traceback" 
done; done
