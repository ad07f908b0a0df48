# round-2 evidence on one B200: ncu launch list of a bench run, one full ncu capture of k_fused_tc, role timing
# (instrumented build), parity sweep against the oracle
mkdir -p gpurun_out
TAG=${1:-ev}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --clips-per-band 1184 --no-cpu --no-extras > gpurun_out/bench_ncu1_$TAG.log 2>&1
tail -2 gpurun_out/bench_ncu1_$TAG.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_tc -s 1 -c 1 -o gpurun_out/prof_tc_$TAG -f python bench.py --steps 1 --warmup 1 --clips-per-band 1184 --no-cpu --no-extras > gpurun_out/bench_ncu2_$TAG.log 2>&1
ls -la gpurun_out/prof_tc_$TAG.ncu-rep
MICLOC_B200_LIB=$PWD/tools/libmicloc_b200_rt.so ROLE_NAMES=front0,front1,front2,front3,bandpass,rzcc,neuron,gram FIR_ROLES=1 timeout 300 python tools/role_timing.py 1776 > gpurun_out/roles_$TAG.log 2>&1; grep -E "^rep|phase|mean busy" gpurun_out/roles_$TAG.log
( time python tools/parity_sweep.py 2002 > gpurun_out/parity_sweep_$TAG.json ) 2>&1 | tail -12
