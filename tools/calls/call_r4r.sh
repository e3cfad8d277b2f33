export PYTHONPATH=.
CMD='python -m pytest tests/test_lrkd_eigensolve_gpu.py tests/test_saliency_gpu.py tests/test_sinkhorn_gpu.py tests/test_lrkd_gpu.py -m gpu -q -x -k "wide-32-1 or zero-384-1 or decay-64-1 or scores_shapes or sinkhorn or lrkd_r32"'
timeout 900 bash -c "$CMD" > gpurun_out/r4r_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r4r_plain.log; exit 1; }
tail -2 gpurun_out/r4r_plain.log
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/r4r_memcheck.log bash -c "$CMD" > gpurun_out/r4r_memcheck_run.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r4r_memcheck_run.log; tail -5 gpurun_out/r4r_memcheck.log
