# usage: bash scripts/gpu_sweep.sh -- bench with a few settings of the size thresholds (development)
for cfg in "49152 131072" "0 1000000000" "32768 262144" "65536 131072" "49152 65536"; do
  set -- $cfg
  MR_KWT_SMALL=$1 MR_KWT_LARGE=$2 python bench.py --steps 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('small $1 large $2:', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
done
