# which of the dynamics kernel's five phase barriers to keep (A top of the substep, B after FK, C before collision, D before the rows,
# E after the solve): bit mask DYN_SYNC_MASK, at 4096 and 65 536 envs
for m in 17 1 16 21 25 31; do
  B200_NVCC_EXTRA="-DDYN_SYNC_MASK=$m" python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
  for n in 4096 65536; do
    python bench.py --envs $n --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 100 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sync mask $m, envs $n:', round(d['value']/1e6,2), 'M', {k:round(v.get('avg_ms', 0),4) for k,v in d['kernels'].items()})"
  done
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
