for mb in 5 7; do
  B200_NVCC_EXTRA="-DENV_MIN_BLOCKS=$mb" python -m hcr_genesis_lr_cl_b200.build --force 2>&1 | grep -E "Used" | tail -1
  for task in go2_wtw tron1_pf; do
  python bench.py --task $task --steps 100 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minblocks=$mb $task', round(d['value']/1e6,2), {k:(round(v['avg_ms'],4), v.get('blocks_per_sm')) for k,v in d['kernels'].items()})"
  done
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
