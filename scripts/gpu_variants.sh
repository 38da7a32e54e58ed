# usage: bash scripts/gpu_variants.sh <tag> -- short device-resident bench of every library build under variants/ (development)
tag=${1:-v}
for so in variants/*.so mizuroute_b200/libmizuroute_b200.so; do
  MR_LIB_PATH=$PWD/$so python bench.py --steps 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$so', round(d['ms_per_step'],1), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
done
