# Round-2 evidence run (1 GPU): default bench line, the other single-GPU workloads, an ncu window with DRAM bytes, two full captures
mkdir -p gpurun_out
python bench.py > gpurun_out/r2z_bench_c4.json 2> gpurun_out/r2z_bench_c4.err; tail -2 gpurun_out/r2z_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2z_ref_c4.json 2> gpurun_out/r2z_ref_c4.err
for wl in C2 C3 C5; do python bench.py --workload $wl > gpurun_out/r2z_bench_$wl.json 2> gpurun_out/r2z_bench_$wl.err; done
python - <<PY
import json
for f in ("c4","C2","C3","C5"):
    try:
        d=json.loads([l for l in open("gpurun_out/r2z_bench_%s.json"%f) if l.startswith("{")][-1])
        print(f, "value %.3e e2e %.3e ms %.1f frac %.3f cpu %.3e err %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["max_rel_err_gpu_vs_cpu"]))
    except Exception as e: print(f, "failed", e)
PY
# ncu window: 400 launches from the middle of the profiled C4 batch, device time + DRAM bytes
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -s 5000 -c 400 --csv \
    --log-file gpurun_out/r2z_window.csv python scripts/prof_kwt.py 3000000 384 1 12 > gpurun_out/r2z_window.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2z_window.csv | tee gpurun_out/r2z_window.txt
# full captures of one large light launch and one team launch of the same region
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_route_kwt_light -s 1700 -c 1 -f -o gpurun_out/r2z_light \
    python scripts/prof_kwt.py 3000000 384 1 2 > gpurun_out/r2z_light.log 2>&1; tail -1 gpurun_out/r2z_light.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_route_kwt_team -s 1700 -c 1 -f -o gpurun_out/r2z_team \
    python scripts/prof_kwt.py 3000000 384 1 2 > gpurun_out/r2z_team.log 2>&1; tail -1 gpurun_out/r2z_team.log
