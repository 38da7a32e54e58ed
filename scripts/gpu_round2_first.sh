mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -15 gpurun_out/r2a_pytest_gpu.log
python scripts/repeat_gpu_test.py tests.test_schemes_gpu:test_water_management_in_kwt 40 > gpurun_out/r2a_wm_repeat.log 2>&1; tail -3 gpurun_out/r2a_wm_repeat.log
MR_KWT_PROFILE=1 python bench.py --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_prof.json 2> gpurun_out/r2a_prof.err; grep "kwt tasks" gpurun_out/r2a_prof.err
python bench.py > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; cat gpurun_out/r2a_bench_c4.json | cut -c1-1500
