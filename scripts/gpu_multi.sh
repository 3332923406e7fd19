#!/bin/bash
# multi-GPU check of bench.py (layer-group pipeline over NCCL p2p)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
for WL in ${WORKLOADS:-7b-4b-32k 7b-3b-128k}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 10 --warmup 3 --workload $WL --no-cpu-baseline ${BENCH_EXTRA} > gpurun_out/bench_n${N}_${WL}.log 2> gpurun_out/bench_n${N}_${WL}.err
  echo "rc=$? $WL"; tail -2 gpurun_out/bench_n${N}_${WL}.log | cut -c1-600; tail -3 gpurun_out/bench_n${N}_${WL}.err
done
