set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2ac_pdl1.json 2> gpurun_out/r2ac.err
B200_ENV_PDL=0 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2ac_pdl0.json 2>> gpurun_out/r2ac.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2ac_pdl1b.json 2>> gpurun_out/r2ac.err
python - <<'PY'
import json
for n in ("pdl1","pdl0","pdl1b"):
    j=json.load(open(f"gpurun_out/r2ac_{n}.json"))
    print(n, j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j["e2e"]["ms_per_step"], j["kernels"]["dynamics_step_kernel"]["avg_ms"], j["kernels"]["env_post_step_kernel"]["avg_ms"])
PY
