#!/bin/bash
# parity tests (incl. device-resident length), bench with the dynamic and the static graph, ncu full capture of the attend kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --torch-profile gpurun_out/step_kernels.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-420 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --graph static > gpurun_out/bench_static.log 2> gpurun_out/bench_static.err; echo "bench static rc=$?"; cut -c1-420 gpurun_out/bench_static.log
PROBE_QUICK=1 PROBE_BITS=3,4 PROBE_L=131072 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scores|v_native|k_outlier|attend_' -c 60 -o gpurun_out/prof_kv3 python scripts/gpu_probe.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
