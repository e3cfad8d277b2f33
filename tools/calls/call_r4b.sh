timeout 900 python -m pytest tests/test_lrkd_eigensolve_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -40
timeout 900 python -m pytest tests/test_lrkd_gpu.py -m gpu -q 2>&1 | grep -v Warning | tail -15
