#!/bin/bash
# ncu --set full of the fused attend's kernels at 128K (one launch each), raw CSV back under gpurun_out/
mkdir -p gpurun_out
export PROBE_BITS=${PROBE_BITS:-4} PROBE_L=131072 PROBE_NL=2 PROBE_PREC=fp16
ncu --set full --clock-control none --import-source on -k regex:'k_fast_kernel|v_fast_kernel|k_outlier_pers' \
    --launch-skip 6 --launch-count 3 -o gpurun_out/r2_attend_${PROBE_BITS}b -f python scripts/r2_attend_probe.py > gpurun_out/ncu_run.log 2>&1
ncu -i gpurun_out/r2_attend_${PROBE_BITS}b.ncu-rep --page raw --csv > gpurun_out/r2_attend_${PROBE_BITS}b_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_run.log
