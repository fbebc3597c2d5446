# tuning sweep for the dynamics kernel: warps per CTA x resident CTAs per SM (28 warps/SM in every case), at one wave (4096 envs)
# and in the multi-wave regime (8192 / 65536 envs), where CTAs of different phases share an SM's instruction cache
for cfg in "7 4" "14 2" "28 1"; do
  set -- $cfg
  B200_NVCC_EXTRA="-DDYN_WARPS_PER_BLOCK=$1 -DDYN_MIN_BLOCKS=$2" python -m hcr_genesis_lr_cl_b200.build --force 2>&1 | grep -E "Used" | head -1
  for n in 4096 8192 65536; do
    python bench.py --envs $n --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 100 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warps/CTA x CTAs/SM = $1 x $2, envs $n:', round(d['value']/1e6,2), 'M', {k:round(v.get('avg_ms', 0),4) for k,v in d['kernels'].items()})"
  done
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
