# usage: bash scripts/gpu_ncu_list.sh <tag> [prof_kwt.py args] -- device time of every launch of ONE profiled batch (scripts/prof_kwt.py)
tag=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv \
    python scripts/prof_kwt.py "$@" > gpurun_out/${tag}_list.log 2>&1
tail -2 gpurun_out/${tag}_list.log
python scripts/ncu_launch_summary.py gpurun_out/${tag}_launches.csv | tee gpurun_out/${tag}_launches.txt
