set -x
timeout 300 python -m pytest tests/test_logit_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --workload soft_kd_logits_b16384_c1000_bf16 --no-cpu-baseline --steps 20 > gpurun_out/r3c_logit.json 2> gpurun_out/r3c_logit.err; python tools/bench_table.py gpurun_out/r3c_logit.json
for CFG in "1 1" "2 3" "4 3"; do
  set -- $CFG
  export DKD_CONV_CLUSTER=$1 DKD_WGRAD_CLUSTER=$2
  echo "=== conv cluster $1 wgrad cluster $2"
  timeout 600 python -m pytest tests/test_mgd_gpu.py tests/test_saliency_gpu.py -m gpu -q -x 2>&1 | tail -4
  for W in mgd_b512_bf16 mgd_b512_f32; do
    timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 10 > gpurun_out/r3c_${W}_c$1_w$2.json 2> gpurun_out/r3c_${W}_c$1_w$2.err
    python tools/bench_table.py gpurun_out/r3c_${W}_c$1_w$2.json; tail -c 300 gpurun_out/r3c_${W}_c$1_w$2.err
  done
done
export DKD_CONV_CLUSTER=2 DKD_WGRAD_CLUSTER=3
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -x -k "masked_generation or wass" 2>&1 | tail -12
timeout 200 python bench.py --workload wasskd_l1_b512_f32 --no-cpu-baseline --steps 10 > gpurun_out/r3c_wass.json 2> gpurun_out/r3c_wass.err; python tools/bench_table.py gpurun_out/r3c_wass.json
