#!/bin/bash
# A/B of the dense K-score kernels: parity tests with the pair-table form (default), then probe timings pair vs lds64
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/probe.jsonl
PROBE_TAG=pair PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_pair.log 2>&1; tail -4 gpurun_out/probe_pair.log | cut -c1-700
KVQ_K_IMPL=lds64 PROBE_TAG=lds64 PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_lds64.log 2>&1; tail -4 gpurun_out/probe_lds64.log | cut -c1-700
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench.log
