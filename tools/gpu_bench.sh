#!/bin/bash
# one short steady-state bench line per call: bash tools/gpu_bench.sh <name> [bench.py args...]
name=$1; shift
python bench.py --steps 100 --warmup 10 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
python -c "
import json; d=json.load(open('gpurun_out/$name.json')); print('RESULT $name', round(d['value']/1e6,2), round(d['ms_per_step'],4), round(d['e2e']['value']/1e6,2), {k:round(v.get('avg_ms',v.get('avg_ms_alone')),4) for k,v in d['kernels'].items()}, round(d['workload_state']['solver_work_per_env']))"
