# bash tools/r2_multi_gpu.sh N tag : the driver's N-GPU bench command, + (N = 8) BASELINE config C3 (go2_wtw, 2048 envs per GPU)
# and C5 (go2_cat sweep with the policy, a PPO-shaped update and the bucketed gradient all-reduce)
N=$1; tag=$2
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/${tag}_n$N.err | grep "^{" > gpurun_out/${tag}_bench_n$N.json
if [ "$N" = "8" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --task go2_wtw --envs 2048 2>> gpurun_out/${tag}_n$N.err | grep "^{" > gpurun_out/${tag}_bench_wtw_n8.json
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_cat_sweep.py --envs 8192 65536 2>> gpurun_out/${tag}_n$N.err | grep "^{" > gpurun_out/${tag}_cat_sweep_n$N.log
tail -c 400 gpurun_out/${tag}_n$N.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*n$N.json")):
    j=json.loads(open(f).readline()); print(f, j["n_gpus"], round(j["value"]/1e6,1), round(j["ms_per_step"],4), round(j["e2e"]["value"]/1e6,1))
for l in open("gpurun_out/${tag}_cat_sweep_n$N.log"):
    j=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in j.items() if k in ("envs_total","collection_env_substeps_per_s","iteration_env_substeps_per_s","allreduce_us_per_call","allreduce_back_to_back_us","bucket_bytes")})
PY
