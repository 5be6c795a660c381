echo "== default"; python scratch/probe_sizes.py diff 2>&1 | grep diffusion
echo "== inplace=1 ns=2"; QPB_PIPE_INPLACE=1 QPB_PIPE_NS=2 python scratch/probe_sizes.py diff 2>&1 | grep diffusion
echo "== inplace=0 ns<=4"; QPB_PIPE_INPLACE=0 python scratch/probe_sizes.py diff 2>&1 | grep diffusion
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
