timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/probe_c5.py 2>&1 | tail -2
