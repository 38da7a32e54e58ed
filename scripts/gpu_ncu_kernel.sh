# usage: bash scripts/gpu_ncu_kernel.sh <tag> <kernel regex> <skip> [prof_kwt.py args] -- ncu --set full of ONE launch of the profiled batch
tag=$1; rx=$2; skip=$3; shift 3
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${tag} \
    python scripts/prof_kwt.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
