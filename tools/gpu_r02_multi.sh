#!/bin/bash
# N-GPU session: the multi-process tests, then the bench lines the driver's scaling run produces (render + train)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1; echo "pytest multi rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|skipped|E   " gpurun_out/m_pytest.log | tail -12 | cut -c1-600
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m_bench_${N}gpu.json 2> gpurun_out/m_bench_${N}gpu.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/m_bench_${N}gpu.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d.get('strong',{}).items(): print('strong', k, v['ms_per_frame'], v['value'], v['rays_per_rank'])
print('train', json.dumps(d.get('train')))
PY
tail -5 gpurun_out/m_bench_${N}gpu.err
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --workload train > gpurun_out/m_bench_train_${N}gpu.json 2> gpurun_out/m_bench_train_${N}gpu.err; echo "bench train rc=$?"
head -c 700 gpurun_out/m_bench_train_${N}gpu.json; echo; tail -3 gpurun_out/m_bench_train_${N}gpu.err
