export PYTHONPATH=.
timeout 600 python -m pytest tests/test_host_contract_gpu.py -m gpu -q 2>&1 | grep -v Warn | tail -15
