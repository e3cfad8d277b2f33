export PYTHONPATH=.
python tools/debug_jacobi_prof.py 2>&1 | grep -v Warn | tail -20
