set -x
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench.err
python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/r2z_bench_reference.json 2>> gpurun_out/r2z_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 300 > gpurun_out/r2z_ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dynamics_step_kernel|env_post_step" -s 640 -c 4 -o gpurun_out/r2z_kernels_full -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pre-roll 300 > gpurun_out/r2z_ncu_full.log 2>&1
tail -3 gpurun_out/r2z_ncu_full.log
cat gpurun_out/r2z_bench_n1.json | cut -c1-400
for t in go2 go2_cat go2_wtw go2_cts go2_ee go2_dreamwaq tron1_pf tron1_pf_ee; do
  python bench.py --steps 100 --warmup 10 --no-cpu-baseline --task $t > gpurun_out/r2z_bench_$t.json 2>> gpurun_out/r2z_bench.err
done
python tools/bench_collection.py > gpurun_out/r2z_collection.json 2>> gpurun_out/r2z_bench.err
python tools/bench_collection.py --task go2_cat --envs 8192 > gpurun_out/r2z_collection_cat8192.json 2>> gpurun_out/r2z_bench.err
