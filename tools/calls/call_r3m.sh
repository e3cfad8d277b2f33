timeout 300 python bench.py --workload wasskd_l1_b512_f32 --ncu-op > gpurun_out/r3m_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wass_sort -c 1 -f -o gpurun_out/r3m_sort \
  python bench.py --workload wasskd_l1_b512_f32 --ncu-op > gpurun_out/r3m_ncu.log 2>&1
echo "ncu rc=$?"
