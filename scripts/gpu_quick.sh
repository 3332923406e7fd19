#!/bin/bash
# quick loop: the K/attend/decode parity tests, GEMV micro-benchmark, probe timings at 128K, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode.py -m gpu -q --tb=short -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_gemv.py > gpurun_out/gemv.jsonl 2> gpurun_out/gemv.err; cut -c1-200 gpurun_out/gemv.jsonl; tail -3 gpurun_out/gemv.err
rm -f gpurun_out/probe.jsonl
PROBE_TAG=default PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_default.log 2>&1; tail -2 gpurun_out/probe_default.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-profile gpurun_out/step_kernels.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
