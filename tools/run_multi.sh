#!/bin/bash
# bash tools/run_multi.sh N [tag]: the driver's launch line for N GPUs of one node + the data-parallel equivalence check
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/ddp_check.py > gpurun_out/ddp_check_${TAG}_n$N.log 2>&1
tail -3 gpurun_out/ddp_check_${TAG}_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
tail -c 600 gpurun_out/bench_${TAG}_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${TAG}_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('allreduce'), d['timed_repeats'])
PY
