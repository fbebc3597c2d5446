# tuning sweep for history_shift_kernel: loads in flight per thread, block size, resident blocks, grid cap per SM
for cfg in "4 256 1 8" "4 256 4 8" "8 256 2 8" "8 256 4 4" "4 512 2 4" "8 512 1 2" "2 256 6 8" "4 256 4 3" "8 256 3 3"; do
  set -- $cfg
  B200_NVCC_EXTRA="-DHIST_SHIFT_UNROLL=$1 -DHIST_SHIFT_BLOCK=$2 -DHIST_SHIFT_MIN_BLOCKS=$3 -DHIST_SHIFT_GRID_PER_SM=$4" python -m hcr_genesis_lr_cl_b200.build --force 2>&1 | grep -A3 "history_shift" | grep Used | head -1
  python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']['history_shift_kernel']; print('cfg=$cfg', 'shift_ms', round(k['avg_ms_alone'],4), 'GB/s', round(k['hbm_gbs']), 'step', round(d['ms_per_step'],4))"
done
python -m hcr_genesis_lr_cl_b200.build --force > /dev/null 2>&1
