set -x
timeout 600 python -m pytest tests/test_logit_gpu.py tests/test_host_contract_gpu.py -m gpu -q -x 2>&1 | tail -15
for R in 0 1; do
  DKD_LOGIT_RING=$R timeout 300 python bench.py --workload soft_kd_logits_b16384_c1000_bf16 --no-cpu-baseline --steps 20 > gpurun_out/r3b_logit_ring$R.json 2> gpurun_out/r3b_logit_ring$R.err
  python tools/bench_table.py gpurun_out/r3b_logit_ring$R.json
done
bash tools/gpu_r3.sh r3b o
