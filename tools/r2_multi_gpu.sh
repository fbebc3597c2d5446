# bash tools/r2_multi_gpu.sh N tag : the driver's N-GPU bench command (+ C3 go2_wtw at 2048 envs per GPU when N = 8)
N=$1; tag=$2
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_n$N.err
tail -c 600 gpurun_out/${tag}_n$N.err
if [ "$N" = "8" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --task go2_wtw --envs 2048 > gpurun_out/${tag}_bench_wtw_n8.json 2>> gpurun_out/${tag}_n$N.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*n$N.json")):
    j=json.loads([l for l in open(f) if l.startswith("{")][0]); print(f, j["n_gpus"], round(j["value"]/1e6,1), round(j["ms_per_step"],4), round(j["e2e"]["value"]/1e6,1))
PY
