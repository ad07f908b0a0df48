mkdir -p gpurun_out
TAG=${1:-b}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --clips-per-band 888 --no-cpu > gpurun_out/bench_ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 888 --no-cpu > gpurun_out/bench_ncu2_$TAG.log 2>&1
ls -la gpurun_out
