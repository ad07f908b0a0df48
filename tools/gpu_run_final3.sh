# round-2 final evidence (session 4): full GPU suite, smoke, memcheck over the new wide-array kernels, bench + CPU arm,
# launch list of a bench run, ncu captures of the config-5 tensor-core kernels
mkdir -p gpurun_out
TAG=${1:-fin3}
python -m pytest tests -m gpu -q -rs --durations=5 2>&1 | tail -16 > gpurun_out/tests_$TAG.log; tail -4 gpurun_out/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
SAN_TIMEOUT=200 bash tools/gpu_run_sanitize.sh $TAG memcheck "wide_array or tensor_core_stht or segmented or staged_taps"
bash tools/gpu_run_benchonly.sh $TAG
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --clips-per-band 1184 --no-cpu --no-extras > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_stht_tc|k_gram_tc" -s 2 -c 2 -o gpurun_out/prof_c5tc_$TAG -f python tools/c5_probe.py 1 > gpurun_out/ncu_c5tc_$TAG.log 2>&1
tail -1 gpurun_out/ncu_c5tc_$TAG.log
bash tools/gpu_run_c5.sh
