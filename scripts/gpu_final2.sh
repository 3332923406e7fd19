#!/bin/bash
# round-1 evidence pass: full GPU test suite, smoke, probe (ours vs the reference's own kernels), GEMV micro-benchmark,
# bench (default + 4-bit targets + reference arm), ncu launch list of the bench command, ncu full capture of the attend kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/nvidia_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
rm -f gpurun_out/probe.jsonl
PROBE_BITS=4,3 PROBE_L=32768,131072 timeout 900 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; grep -c '"event"' gpurun_out/probe.log
timeout 300 python scripts/gpu_gemv.py > gpurun_out/gemv.jsonl 2> gpurun_out/gemv.err
timeout 900 python bench.py --steps 20 --warmup 3 --torch-profile gpurun_out/step_kernels.txt > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 700 gpurun_out/bench.log
timeout 900 python bench.py --steps 20 --warmup 3 --workload 7b-4b-128k --no-cpu-baseline > gpurun_out/bench_4b128k.log 2> gpurun_out/bench_4b128k.err; echo "bench4 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 --workload 7b-4b-32k --no-cpu-baseline > gpurun_out/bench_4b32k.log 2> gpurun_out/bench_4b32k.err; echo "bench32k rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_scores|v_native|k_outlier|attend_|append_kv|dec_' -c 1400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
PROBE_QUICK=1 PROBE_BITS=3,4 PROBE_L=131072 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scores|v_native|k_outlier|attend_' -c 28 -o gpurun_out/prof_attend python scripts/gpu_probe.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -25
