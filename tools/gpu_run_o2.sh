mkdir -p gpurun_out
TAG=${1:-o2}
MICLOC_FUSED_FIR=ffa ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_ffa_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu > gpurun_out/bench_ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_ffa_$TAG.ncu-rep
