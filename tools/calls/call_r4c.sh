export PYTHONPATH=.
python tools/debug_lrkd.py 2>&1 | grep -v Warn | tail -20
DKD_LRKD_CLUSTER=0 python tools/debug_lrkd.py 2>&1 | grep -v Warn | tail -20
timeout 900 python -m pytest tests/test_lrkd_eigensolve_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -30
