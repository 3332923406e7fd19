#!/bin/bash
# final evidence pass: full GPU test suite, bench (default + 4-bit target), ncu launch list + full capture of the hot kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
rm -f gpurun_out/probe.jsonl
PROBE_BITS=4,3 PROBE_L=32768,131072 timeout 600 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; grep -c '"event"' gpurun_out/probe.log
timeout 900 python bench.py --steps 20 --warmup 3 --torch-profile gpurun_out/step_kernels.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.log
timeout 900 python bench.py --steps 20 --warmup 3 --workload 7b-4b-128k --no-cpu-baseline > gpurun_out/bench_4b128k.log 2> gpurun_out/bench_4b128k.err; echo "bench4 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --workload 7b-4b-32k --no-cpu-baseline > gpurun_out/bench_4b32k.log 2> gpurun_out/bench_4b32k.err; echo "bench32k rc=$?"
PROBE_QUICK=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_scores|v_accum|v_native|k_outlier|attend_|append_kv' -c 400 --csv --log-file gpurun_out/launches_probe.csv python scripts/gpu_probe.py > gpurun_out/ncu_list.log 2>&1
PROBE_QUICK=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scores|v_native|k_outlier' -s 4 -c 10 -o gpurun_out/prof_kv python scripts/gpu_probe.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
