#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, then the ncu launch list and full captures.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag] [stages]
#   stages: any of t (tests) s (smoke) b (bench) l (launch list) n (ncu full captures); default "tsbln"
TAG=${1:-r1}
STAGES=${2:-tsbln}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1

if [[ $STAGES == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $OUT/${TAG}_pytest.log
  tail -15 $OUT/${TAG}_pytest.log
fi
if [[ $STAGES == *s* ]]; then
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
fi
if [[ $STAGES == *b* ]]; then
  timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  echo "bench rc=$?"; tail -c 6000 $OUT/${TAG}_bench.json; tail -c 1500 $OUT/${TAG}_bench.err
  timeout 300 python bench.py --impl reference --steps 200 --warmup 5 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
  tail -c 1000 $OUT/${TAG}_bench_ref.json
fi
if [[ $STAGES == *l* ]]; then
  # launch list of the same bench command (short): per-launch durations, cold cache, serialised
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "launch list rc=$?"; wc -l $OUT/${TAG}_launches.csv
fi
if [[ $STAGES == *n* ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:logit_kd -c 2 -f -o $OUT/${TAG}_logit_kd \
    python bench.py --no-extras --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_logit.log 2>&1
  echo "ncu logit rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 12 -f -o $OUT/${TAG}_curkd \
    python bench.py --workload curkd_early_3layers_b512_f32 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_curkd.log 2>&1
  echo "ncu curkd rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_ -c 12 -f -o $OUT/${TAG}_mgd \
    python bench.py --workload mgd_b512_f32 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_mgd.log 2>&1
  echo "ncu mgd rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:wass_sort -c 1 -f -o $OUT/${TAG}_wass_sort \
    python bench.py --workload wasskd_l1_b512_f32 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_wass.log 2>&1
  echo "ncu wass rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -c 2 -f -o $OUT/${TAG}_sinkhorn \
    python bench.py --workload wasskd_sinkhorn_b512_f32 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_sink.log 2>&1
  echo "ncu sinkhorn rc=$?"
  ls -la $OUT/*.ncu-rep
fi
