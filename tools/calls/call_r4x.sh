export PYTHONPATH=.
bash tools/gpu_r3.sh r4x tsb
NCU_WORKLOADS="wasskd_sinkhorn_b512_f32" bash tools/gpu_r3.sh r4x o
