timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/probe_sizes.py diff 2>&1 | grep diffusion
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','time_shares')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_sweeps']['frac'], d['roofline_sweeps']['ms_per_launch'])
PY
