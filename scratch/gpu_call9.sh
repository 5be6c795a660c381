timeout 900 python -m pytest tests -m gpu -x -q -k "segmented or large_grid" 2>&1 | tail -2
python scratch/probe_orient.py 2>&1 | grep "1024"
python scratch/probe_sizes.py diff 2>&1 | grep "1024\|2048"
