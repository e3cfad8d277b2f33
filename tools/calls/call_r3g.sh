set -x
timeout 300 python bench.py --workload curkd_early_3layers_b512_f32 --ncu-op > gpurun_out/r3g_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:align_fused -c 1 -f -o gpurun_out/r3g_fused \
  python bench.py --workload curkd_early_3layers_b512_f32 --ncu-op > gpurun_out/r3g_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python tools/e2e_probe.py cprofile > gpurun_out/r3g_e2e_probe.log 2>&1
head -12 gpurun_out/r3g_e2e_probe.log
