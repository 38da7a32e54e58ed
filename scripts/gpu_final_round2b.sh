# Final evidence of round 2 with the final library (1 GPU): GPU suite, smoke, bench lines
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu > gpurun_out/r2y_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_pytest_gpu.log; tail -3 gpurun_out/r2y_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2y_bench_c4.json 2> gpurun_out/r2y_bench_c4.err; tail -2 gpurun_out/r2y_bench_c4.err
for wl in C2 C3 C5; do python bench.py --workload $wl > gpurun_out/r2y_bench_$wl.json 2> gpurun_out/r2y_bench_$wl.err; done
python - <<PY
import json
for f in ("c4","C2","C3","C5"):
    try:
        d=json.loads([l for l in open("gpurun_out/r2y_bench_%s.json"%f) if l.startswith("{")][-1])
        print(f, "value %.3e e2e %.3e ms %.1f frac %.3f k1 %s cpu %.3e err %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("k1_ms_per_step"), d["cpu_baseline"]["value"], d["cpu_baseline"]["max_rel_err_gpu_vs_cpu"]))
    except Exception as e: print(f, "failed", e)
PY
