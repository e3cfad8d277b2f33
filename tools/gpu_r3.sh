#!/bin/bash
# Round-2 GPU pass (run on the box through gpurun):  bash tools/gpu_r3.sh TAG STAGES
#   t  pytest -m gpu                      s  smoke            b  bench (all workloads) + reference arm
#   q  pytest: baseline-size parity only  o  per-op ncu passes (time + DRAM bytes of ONE fused-op call per workload)
TAG=${1:-r3a}
STAGES=${2:-tsb}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1
if [[ $STAGES == *q* ]]; then
  timeout 1500 python -m pytest tests/test_baseline_sizes_gpu.py -m gpu -q -s 2>&1 | tail -80 > $OUT/${TAG}_pytest_big.log
  tail -40 $OUT/${TAG}_pytest_big.log
fi
if [[ $STAGES == *t* ]]; then
  timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -60 > $OUT/${TAG}_pytest.log
  tail -25 $OUT/${TAG}_pytest.log
fi
if [[ $STAGES == *s* ]]; then
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
fi
if [[ $STAGES == *b* ]]; then
  timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "bench rc=$?"; python tools/bench_table.py $OUT/${TAG}_bench.json; tail -c 1500 $OUT/${TAG}_bench.err
  timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
  tail -c 600 $OUT/${TAG}_bench_ref.json
fi
if [[ $STAGES == *o* ]]; then
  for W in ${NCU_WORKLOADS:-soft_kd_logits_b256_c1000_bf16 soft_kd_logits_b16384_c1000_bf16 curkd_early_3layers_b512_f32 curkd_early_3layers_b512_bf16 mgd_b512_f32 mgd_b512_bf16 saliency_mgd_m1_b512_f32 lrkd_r64_b512_f32 wasskd_l1_b512_f32 wasskd_sinkhorn_b512_f32}; do
    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --profile-from-start off --csv --log-file $OUT/ncuop_${W}.csv python bench.py --workload $W --ncu-op > $OUT/${TAG}_ncuop_${W}.log 2>&1
    echo "ncu-op $W rc=$?"
  done
fi
