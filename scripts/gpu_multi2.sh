#!/bin/bash
# multi-GPU bench: sequence-sharded (sp, default) and layer-pipeline (pp) decode on N GPUs of this box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
for par in ${PARS:-sp pp}; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --parallelism $par > gpurun_out/bench_n${N}_${par}.log 2> gpurun_out/bench_n${N}_${par}.err
  echo "rc=$? $par"; cut -c1-330 gpurun_out/bench_n${N}_${par}.log; tail -2 gpurun_out/bench_n${N}_${par}.err
done
