export PYTHONPATH=.
timeout 900 python -m pytest tests/test_lrkd_eigensolve_gpu.py -m gpu -q -s > gpurun_out/r4f_eig.log 2>&1
grep -n "LRKD eigensolve\|passed\|failed\|FAILED\|Error" gpurun_out/r4f_eig.log | head -30
timeout 900 python -m pytest tests/test_lrkd_gpu.py -m gpu -q > gpurun_out/r4f_lrkd.log 2>&1; tail -3 gpurun_out/r4f_lrkd.log
timeout 300 python bench.py --workload lrkd_r64_b512_f32 --no-cpu-baseline --steps 10 > gpurun_out/r4f_lrkd_c1.json 2> gpurun_out/r4f_lrkd_c1.err
python tools/bench_table.py gpurun_out/r4f_lrkd_c1.json | tail -1
