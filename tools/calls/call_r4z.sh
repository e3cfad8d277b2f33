export PYTHONPATH=.
for W in lrkd_r64_b512_f32 curkd_early_3layers_b512_bf16; do
  timeout 300 python bench.py --workload $W --no-cpu-baseline --steps 10 > gpurun_out/r4z_$W.json 2> gpurun_out/r4z_$W.err; echo "rc=$?"
  python tools/bench_table.py gpurun_out/r4z_$W.json | grep "^$W" | cut -c1-170
  python -c "
import json;d=json.loads(open('gpurun_out/r4z_$W.json').read().strip().splitlines()[-1]);r=d['roofline'];print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k.startswith('frac') or k.startswith('eigensolve_us') or k=='bound'})"
done
