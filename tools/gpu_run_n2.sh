N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_r2.json 2> gpurun_out/bench_${N}gpu_r2.err
tail -3 gpurun_out/bench_${N}gpu_r2.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu_r2.json'))
print('N=$N', round(d['value']), 'clips/s', 'e2e', round(d['e2e']['value']), 'pcie', round(d['e2e']['pcie_h2d_gbs'],1), 'f32', round(d['e2e']['f32']['value']), d['e2e']['numa'], d['clocks'])
PY
nvidia-smi topo -m 2>/dev/null | head -12; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" | head; free -g | head -2
