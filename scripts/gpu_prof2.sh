#!/bin/bash
# parity tests, GEMV micro-benchmark, probe timings, default bench with kernel table
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_gemv.py > gpurun_out/gemv.jsonl 2> gpurun_out/gemv.err; cat gpurun_out/gemv.jsonl | cut -c1-300; tail -3 gpurun_out/gemv.err
rm -f gpurun_out/probe.jsonl
PROBE_TAG=default PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=32768,131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_default.log 2>&1; tail -2 gpurun_out/probe_default.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-profile gpurun_out/step_kernels.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
