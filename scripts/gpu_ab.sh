#!/bin/bash
# A/B runs of kernel variants selected by environment switches (KVQ_K_IMPL, KVQ_KOUT_IMPL): parity tests with the
# defaults, then probe timings per variant, then the default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/probe.jsonl
PROBE_TAG=default PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_default.log 2>&1; tail -2 gpurun_out/probe_default.log | cut -c1-300
KVQ_KOUT_IMPL=table KVQ_K_IMPL=generic PROBE_TAG=old PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_old.log 2>&1; tail -2 gpurun_out/probe_old.log | cut -c1-300
PROBE_TAG=default32k PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=32768 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_default32k.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench.log
