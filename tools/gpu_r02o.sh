# round 2, call o (N GPUs of one box): the scaling bench with the P2P row exchange and with the NCCL all-gather (A/B), as the driver launches it
set -x
N=${1:-2}
nvidia-smi topo -m 2>&1 | head -12
for x in ${3:-auto nccl}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${2:-128} --warmup 4 --exchange $x > gpurun_out/bench_n${N}_${x}_r02o.json 2> gpurun_out/bench_n${N}_${x}_r02o.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_${x}_r02o.json').read().strip().splitlines()[-1]); print('N=$N $x', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'kernel', d['config']['kernel_ms_over_ranks'], 'e2e', round(d['e2e']['value']), d['config']['collective'][:60])
except Exception as e: print('ERR', e); print(open('gpurun_out/bench_n${N}_${x}_r02o.err').read()[-1500:])
PY
done
