# usage: bash scripts/gpu_quick.sh <tag> [bench args]  -- KWT parity tests + a short device-resident bench (development loop)
tag=${1:-q}; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -x -m gpu 2>&1 | tail -2
python bench.py --steps 2 --no-e2e "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], {k:round(v["ms_per_step"],1) for k,v in d["kernels"].items()}, d["cpu_baseline"]["max_rel_err_gpu_vs_cpu"] if d.get("cpu_baseline") else None)
PY
tail -3 gpurun_out/${tag}_bench.err
