for D in 0 1 2 4 8 16 24 3 7 31; do
  DKD_BENCH_PARITY=0 DKD_FUSED_DEBUG=$D timeout 300 python bench.py --workload curkd_early_3layers_b512_f32 --no-cpu-baseline --steps 20 > gpurun_out/r3i_d$D.json 2> gpurun_out/r3i_d$D.err
  echo "debug=$D"; python tools/bench_table.py gpurun_out/r3i_d$D.json | tail -1
done
