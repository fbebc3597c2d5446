# tuning sweep for the dynamics kernel: warps per CTA x resident CTAs per SM (28 warps/SM in every case), phase barriers
for cfg in "4 7 1" "7 4 1" "14 2 1" "28 1 1" "4 7 0"; do
  set -- $cfg
  ex="-DDYN_WARPS_PER_BLOCK=$1 -DDYN_MIN_BLOCKS=$2"; [ "$3" = "0" ] && ex="$ex -DDYN_NO_PHASE_SYNC"
  B200_NVCC_EXTRA="$ex" python -m hcr_genesis_lr_cl_b200.build --force 2>&1 | grep -E "Used" | head -1
  python bench.py --steps 100 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg=$cfg', round(d['value']/1e6,2), {k:(round(v.get('avg_ms', v.get('avg_ms_alone', 0)),4), v.get('blocks_per_sm')) for k,v in d['kernels'].items()})"
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
