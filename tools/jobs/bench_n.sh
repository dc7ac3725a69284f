#!/bin/bash
# GPU-box job: bench.py at N ranks (one per GPU), launched as the driver does.  $1 = N, $2 = tag
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}; TAG=${2:-x}
nvidia-smi topo -m 2>/dev/null | head -14; numactl -H 2>/dev/null | head -6; lscpu | grep -i "numa\|socket\|^CPU(s)"
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
fi
echo rc=$?; tail -3 gpurun_out/bench_n${N}_$TAG.err; cat gpurun_out/bench_n${N}_$TAG.json
