#!/bin/bash
# Profiling pass on a GPU box: per-workload ncu launch lists (every launch with its device time) and one
# `ncu --set full` capture of each workload's dominant kernels.  Usage: bash tools/profile_round.sh <tag>
# Each ncu run follows a plain run of the same command (B200_PROFILING.md); nothing printed under ncu is a bench value.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
LIST="soft_kd_logits_b256_c1000_bf16 curkd_early_3layers_b512_f32 mgd_b512_f32 lrkd_r64_b512_f32 wasskd_l1_b512_f32 wasskd_sinkhorn_b512_f32 saliency_mgd_m1_b512_f32"
for W in $LIST; do
  CMD="python bench.py --workload $W --no-extras --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > $OUT/${TAG}_plain_$W.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_$W.csv $CMD > /dev/null 2>&1
  echo "launch list $W rc=$?"
done
full() {  # name, kernel regex, count, workload
  CMD="python bench.py --workload $4 --no-extras --steps 1 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > /dev/null 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -c $3 -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
full logit_kd logit_kd 2 soft_kd_logits_b256_c1000_bf16
full curkd gemm_ 3 curkd_early_3layers_b512_f32
full mgd gemm_ 10 mgd_b512_f32
full wass_sort wass_sort 1 wasskd_l1_b512_f32
full sinkhorn "sinkhorn_kernel|gemm_" 6 wasskd_sinkhorn_b512_f32
full jacobi "jacobi|select" 2 lrkd_r64_b512_f32
ls -la $OUT/${TAG}_*.ncu-rep
