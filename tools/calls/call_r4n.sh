export PYTHONPATH=.
for W in soft_kd_logits_b256_c1000_bf16 mgd_b512_bf16; do
  timeout 600 python bench.py --workload $W --no-cpu-baseline --steps 20 > gpurun_out/r4n_$W.json 2> gpurun_out/r4n_$W.err
  python tools/bench_table.py gpurun_out/r4n_$W.json | grep "$W"; tail -c 300 gpurun_out/r4n_$W.err
  python -c "
import json;d=json.loads(open('gpurun_out/r4n_$W.json').read().strip().splitlines()[-1]);print(d['e2e']['ms_per_step_by_mode'])"
done
