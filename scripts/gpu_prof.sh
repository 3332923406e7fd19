#!/bin/bash
# parity tests, probe timings (default vs gather-form outlier pass), ncu full capture of the attend kernels, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/probe.jsonl
PROBE_TAG=default PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=32768,131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_default.log 2>&1; tail -2 gpurun_out/probe_default.log | cut -c1-300
KVQ_KOUT_IMPL=table PROBE_TAG=kout_table PROBE_SKIP_REF=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_old.log 2>&1; tail -2 gpurun_out/probe_old.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench.log
PROBE_QUICK=1 PROBE_BITS=3,4 PROBE_L=131072 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scores|v_native|k_outlier' -s 4 -c 10 -o gpurun_out/prof_kv2 python scripts/gpu_probe.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
