#!/bin/bash
# sequence-sharded vs pipeline bench at N GPUs + single-GPU tests of the new pieces + --impl reference smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "shard or 13b or qnorm" > gpurun_out/pytest_sp.log 2>&1; tail -3 gpurun_out/pytest_sp.log
for PAR in ${PARS:-sp pp}; do
  timeout ${BENCH_TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --parallelism $PAR > gpurun_out/bench_n${N}_$PAR.log 2> gpurun_out/bench_n${N}_$PAR.err
  echo "rc=$? $PAR"; grep '^{' gpurun_out/bench_n${N}_$PAR.log | cut -c1-400; grep -v "Warning\|warn" gpurun_out/bench_n${N}_$PAR.err | tail -5
done
if [ "${RUN_REF:-1}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-600; tail -3 gpurun_out/bench_reference.err
fi
