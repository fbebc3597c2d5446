# tuning sweep for the fused env kernel: warps per CTA x resident CTAs per SM (28 warps per SM), at one wave and at 65 536 envs
for cfg in "4 7" "7 4" "14 2" "2 14"; do
  set -- $cfg
  B200_NVCC_EXTRA="-DENV_WARPS_PER_BLOCK=$1 -DENV_MIN_BLOCKS=$2" python -m hcr_genesis_lr_cl_b200.build --force 2>&1 | grep -E "Used" | tail -1
  for n in 4096 65536; do
    python bench.py --envs $n --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 100 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('env kernel warps/CTA x CTAs/SM = $1 x $2, envs $n:', round(d['value']/1e6,2), 'M', {k:round(v.get('avg_ms', 0),4) for k,v in d['kernels'].items()})"
  done
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
