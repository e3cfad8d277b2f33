#!/bin/bash
# Trimmed profiling pass (round 1, final build): ncu launch lists of three workloads and full captures of the kernels
# changed since the r1m pass.  Each ncu run follows a plain run of the same command; nothing printed under ncu is a bench value.
TAG=${1:-r2p}
OUT=gpurun_out
mkdir -p $OUT
for W in soft_kd_logits_b256_c1000_bf16 curkd_early_3layers_b512_f32 deit_tiny_kd_step_soft_b256_bf16; do
  CMD="python bench.py --workload $W --no-extras --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > $OUT/${TAG}_plain_$W.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_$W.csv $CMD > /dev/null 2>&1
  echo "launch list $W rc=$?"
done
full() {  # name, kernel regex, count, workload
  CMD="python bench.py --workload $4 --no-extras --steps 1 --warmup 3 --no-cpu-baseline"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -c $3 -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
full logit_kd logit_kd 2 soft_kd_logits_b256_c1000_bf16
full logit_big logit_kd 1 soft_kd_logits_b16384_c1000_bf16
DKD_BENCH_STEP_GRAPH=0 full rowops "layernorm|colsum_kernel|head_copy" 8 deit_tiny_kd_step_soft_b256_bf16
ls -la $OUT/${TAG}_*.ncu-rep
