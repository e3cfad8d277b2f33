export PYTHONPATH=.
timeout 600 python bench.py --workload wasskd_sinkhorn_b512_f32 --ncu-op > gpurun_out/r4j_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r4j_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -c 2 -o gpurun_out/r4j_sinkhorn python bench.py --workload wasskd_sinkhorn_b512_f32 --ncu-op > gpurun_out/r4j_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r4j_sinkhorn.ncu-rep
