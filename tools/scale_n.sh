#!/bin/bash
# N-GPU pass on one box (torchrun, one rank per GPU): headline loss path and the end-to-end DeiT-Tiny KD step.
#   bash tools/scale_n.sh <N> <tag>
N=$1; TAG=$2; OUT=gpurun_out; mkdir -p $OUT
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --no-extras --no-cpu-baseline > $OUT/${TAG}_n${N}_headline.json 2> $OUT/${TAG}_n${N}.err; echo "headline rc=$?"
run --workload deit_tiny_kd_step_soft_b256_bf16 --no-extras --no-cpu-baseline --steps 20 --warmup 3 > $OUT/${TAG}_n${N}_deit_soft.json 2>> $OUT/${TAG}_n${N}.err; echo "deit soft rc=$?"
run --workload deit_tiny_kd_step_curkd_b256_bf16 --no-extras --no-cpu-baseline --steps 20 --warmup 3 > $OUT/${TAG}_n${N}_deit_curkd.json 2>> $OUT/${TAG}_n${N}.err; echo "deit curkd rc=$?"
for f in headline deit_soft deit_curkd; do python - $OUT/${TAG}_n${N}_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"], "n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -3 $OUT/${TAG}_n${N}.err | cut -c1-300
