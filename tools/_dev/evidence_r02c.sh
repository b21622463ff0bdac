set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02c_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ipm_kernel|evaluate_kernel|linearize_kernel" -s 6 -c 3 -o gpurun_out/r02c_step python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r02c_ncu_full.log 2>&1
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference.json 2> gpurun_out/r02c_bench_reference.err
python tools/config_sweep.py > gpurun_out/r02c_configs.md 2> gpurun_out/r02c_configs.err
tail -c 300 gpurun_out/r02c_bench_n1.json; tail -3 gpurun_out/r02c_configs.md | cut -c1-200
