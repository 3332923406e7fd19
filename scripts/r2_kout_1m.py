#!/usr/bin/env python
"""probe: fused attend at 1M tokens (one layer), K outlier scatter by rope-table gather vs direct sincos; and equality
of the two at a small size (run under gpurun)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for npos_thr, tag in (("100000000", "gather"), ("0", "direct")):
    env = dict(os.environ, KVQ_KOUT_DIRECT_NPOS=npos_thr, PROBE_BITS=os.environ.get("PROBE_BITS", "3"), PROBE_L=os.environ.get("PROBE_L", "1048576"), PROBE_NL="1", PROBE_PREC="fp32", PROBE_TAG="kout_" + tag)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "r2_attend_probe.py")], env=env, capture_output=True, text=True)
    for l in r.stdout.splitlines():
        if '"time"' in l or '"kernels"' in l:
            print(tag, l[:420])
    if r.returncode:
        print(r.stderr[-1500:])
