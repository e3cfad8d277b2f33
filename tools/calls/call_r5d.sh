export PYTHONPATH=.
bash tools/gpu_r3.sh r5d ts
