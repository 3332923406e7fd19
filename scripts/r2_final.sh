#!/bin/bash
# round-2 final evidence pass (run under gpurun, one GPU): full GPU test suite, bench lines, ncu launch list of the bench
# command, ncu --set full of one fused attend per width (DRAM traffic), all under gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_7b3b128k.json 2> gpurun_out/r02_bench_7b3b128k.err; echo "bench3 rc=$?"
timeout 600 python bench.py --workload 7b-4b-128k --no-anchors > gpurun_out/r02_bench_7b4b128k.json 2> gpurun_out/r02_bench_7b4b128k.err; echo "bench4 rc=$?"
timeout 600 python bench.py --workload 7b-4b-32k --no-anchors --no-cpu-baseline > gpurun_out/r02_bench_7b4b32k.json 2>/dev/null; echo "bench32k rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"
# launch list of the bench command (cold-cache, serialised: shares of the step, not absolute times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 600 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-anchors > gpurun_out/r02_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
for b in 3 4; do
  PROBE_BITS=$b PROBE_L=131072 PROBE_NL=2 PROBE_PREC=fp32 timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:'k_ratio|v_native_kernel|k_outlier_pers|attend_init|attend_combine' --launch-skip 12 --launch-count 12 \
      -o gpurun_out/r02_attend_${b}b -f python scripts/r2_attend_probe.py > gpurun_out/r02_ncu_attend_${b}b.log 2>&1
  ncu -i gpurun_out/r02_attend_${b}b.ncu-rep --page raw --csv > gpurun_out/r02_attend_${b}b_raw.csv 2>/dev/null
  echo "ncu full ${b}b rc=$?"
done
grep -h '"value"' gpurun_out/r02_bench_7b3b128k.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('north_star_kernel',{}).get('frac'))"
