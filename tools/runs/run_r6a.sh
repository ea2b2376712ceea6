( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "porous_equilibration" ) 2>&1 | tail -6
