# usage: bash scripts/gpu_check_bench.sh <tag>   -- GPU test suite, KWT task-class profile, default bench line
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -4 gpurun_out/${tag}_pytest_gpu.log
MR_KWT_PROFILE=1 python bench.py --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_prof.json 2> gpurun_out/${tag}_prof.err; grep "kwt tasks" gpurun_out/${tag}_prof.err
python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_c4.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()}, d["cpu_baseline"]["max_rel_err_gpu_vs_cpu"] if d.get("cpu_baseline") else None)
PY
tail -3 gpurun_out/${tag}_bench_c4.err
