# usage: bash scripts/gpu_ncu_launches.sh <tag> <skip> <count> [bench args] -- ncu launch list (device time + DRAM bytes per launch) of one batch
tag=$1; skip=$2; cnt=$3; shift 3
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $skip -c $cnt --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu_bench.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/${tag}_launches.csv | tee gpurun_out/${tag}_launches.txt
