#!/bin/bash
# GPU pass: parity tests, probe, bench, ncu launch list + full capture of the two streaming kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tests/golden/gen_refcuda_golden.py > gpurun_out/gen_refcuda.log 2>&1; tail -3 gpurun_out/gen_refcuda.log
rm -f gpurun_out/probe.jsonl
PROBE_BITS=${PROBE_BITS:-4,3} PROBE_L=${PROBE_L:-131072} timeout 600 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
grep -E '"impl": "ours"|rc=|Error|error' gpurun_out/probe.log | tail -20
KVQ_K_IMPL=kappa PROBE_TAG=kappa PROBE_SKIP_REF=1 PROBE_BITS=${PROBE_BITS:-4,3} PROBE_L=${PROBE_L:-131072} timeout 600 python scripts/gpu_probe.py > gpurun_out/probe_kappa.log 2>&1
grep -E '"op": "k_opt' gpurun_out/probe_kappa.log | tail -8
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 --torch-profile gpurun_out/step_kernels.txt ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
  tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
if [ "${RUN_NCU:-1}" = "1" ]; then
  PROBE_QUICK=1 PROBE_BITS=4,3 PROBE_L=131072 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_scores|v_accum|v_native|k_outlier|attend_|k_qrot' -c 400 --csv --log-file gpurun_out/launches_probe.csv python scripts/gpu_probe.py > gpurun_out/ncu_list.log 2>&1
  PROBE_QUICK=1 PROBE_BITS=4 PROBE_L=131072 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_scores|v_native|k_outlier' -s 4 -c 8 -o gpurun_out/prof_kv python scripts/gpu_probe.py > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out | tail -12
fi
