# tuning sweep for the fused env kernel: warps per CTA, resident CTAs per SM, section barriers
for cfg in "4 5 1" "4 5 0" "8 2 1" "2 10 1" "4 4 1"; do
  set -- $cfg
  ex="-DENV_WARPS_PER_BLOCK=$1 -DENV_MIN_BLOCKS=$2"; [ "$3" = "0" ] && ex="$ex -DENV_NO_SECTION_SYNC"
  B200_NVCC_EXTRA="$ex" python -m hcr_genesis_lr_cl_b200.build --force 2>&1 | grep -E "Used" | tail -1
  python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg=$cfg', round(d['value']/1e6,2), {k:(round(v['avg_ms'],4), v.get('blocks_per_sm')) for k,v in d['kernels'].items()})"
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
