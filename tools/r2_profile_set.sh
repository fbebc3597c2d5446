set -x
python bench.py > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench.err
python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/r2p_bench_reference.json 2>> gpurun_out/r2p_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 300 > gpurun_out/r2p_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dynamics_step_kernel|env_post_step" -s 640 -c 4 -o gpurun_out/r2p_kernels_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 300 > gpurun_out/r2p_ncu_full.log 2>&1
tail -3 gpurun_out/r2p_ncu_full.log
cat gpurun_out/r2p_bench_n1.json | cut -c1-400
