# ncu --set full captures of the fused kernel: all roles, and FIR warps only (role ablation)
mkdir -p gpurun_out
TAG=${1:-u}
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu > gpurun_out/bench_ncu_$TAG.log 2>&1
MICLOC_FUSED_SKIP=0xF0 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_firo_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu > gpurun_out/bench_ncu2_$TAG.log 2>&1
ls -la gpurun_out/
