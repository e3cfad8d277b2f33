export PYTHONPATH=.
NCU_WORKLOADS="saliency_mgd_m1_b512_f32" bash tools/gpu_r3.sh r4m o
python tools/launch_summary.py gpurun_out/ncuop_saliency_mgd_m1_b512_f32.csv 2>/dev/null | head -30
