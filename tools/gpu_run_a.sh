mkdir -p gpurun_out
set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -3 gpurun_out/bench_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_a.csv python bench.py --steps 2 --warmup 1 --clips-per-band 296 --no-cpu > gpurun_out/bench_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_a -f python bench.py --steps 1 --warmup 1 --clips-per-band 296 --no-cpu > gpurun_out/bench_ncu_b.log 2>&1
ls -la gpurun_out
