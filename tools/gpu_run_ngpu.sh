# bench.py on N GPUs of one box (torchrun, one rank per GPU): usage  gpu_run_ngpu.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -2 gpurun_out/bench_${N}gpu.err
python -c "
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json')); print(d['n_gpus'], 'GPUs', round(d['value']), 'clips/s', round(d['ms_per_step'],2), 'ms/step e2e', round(d['e2e']['value']), d['clocks'])"
