#!/usr/bin/env python
"""GPU probe, round 2 (run under gpurun): the fused attend at full size in both table precisions -- time (CUDA
events, cycling over NL distinct layer caches so that nothing is L2-resident), per-kernel breakdown (torch profiler)
and the fp16-vs-fp32 difference.  Writes gpurun_out/r2_probe.jsonl.  Diagnostic, not a bench value."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kvquant_b200 import synth, cache as kc  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
PEAK = 6501.9


def main():
    bits_list = [int(b) for b in os.environ.get("PROBE_BITS", "4,3").split(",")]
    Ls = [int(x) for x in os.environ.get("PROBE_L", "131072").split(",")]
    NL = int(os.environ.get("PROBE_NL", "4"))
    sparse = os.environ.get("PROBE_DENSE", "0") in ("", "0")
    dev = torch.device("cuda:0")
    H = int(os.environ.get("PROBE_H", "32"))
    sp = synth.SynthSpec(H, 128, seed=0)
    log = open(os.path.join(OUT, "r2_probe.jsonl"), "a")

    def emit(**kw):
        if os.environ.get("PROBE_TAG"):
            kw["tag"] = os.environ["PROBE_TAG"]
        print(json.dumps(kw), flush=True)
        log.write(json.dumps(kw) + "\n"); log.flush()

    emit(event="start", gpu=torch.cuda.get_device_name(0))
    for bits in bits_list:
        cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
        klut = kc.build_k_lookup_table(cal["k"][0], cal["k"][1], cal["k"][2][0], H, device=dev)
        for L in Ls:
            caches = []
            t0 = time.time()
            for i in range(NL):
                lc = kc.LayerCache.from_luts(bits, H, L + 64, dict(lut=klut["lut"], lut2=None, thr_lower=klut["thr_lower"],
                                                                  thr_upper=klut["thr_upper"]), cal["v"][2][0], device=dev,
                                             include_sparse=sparse)
                synth.fill_layer_cache_gpu(lc, sp, L, seed=bits + 10 * i)
                caches.append(lc)
            torch.cuda.synchronize()
            emit(event="filled", bits=bits, L=L, n=NL, secs=round(time.time() - t0, 2))
            q = torch.randn((H, 128), device=dev).half().float().contiguous()
            fbytes = L * caches[0].bytes_per_token()
            outs = {}
            precs = os.environ.get("PROBE_PREC", "fp32,fp16").split(",")
            for prec in precs:
                for lc in caches:
                    lc.precision = prec
                for _ in range(2):
                    for lc in caches:
                        lc.attend(q)
                torch.cuda.synchronize()
                iters = 10
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    for lc in caches:
                        lc.attend(q)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / (iters * NL)
                emit(event="time", op="attend_fused", precision=prec, bits=bits, L=L, sparse=sparse, ms=ms,
                     gbs=fbytes / ms / 1e6, frac=fbytes / ms / 1e6 / PEAK)
                outs[prec] = caches[0].attend(q).clone()
                # per-kernel breakdown
                try:
                    from torch.profiler import profile, ProfilerActivity
                    with profile(activities=[ProfilerActivity.CUDA]) as prof:
                        for _ in range(3):
                            for lc in caches:
                                lc.attend(q)
                        torch.cuda.synchronize()
                    rows = []
                    for e in prof.key_averages():
                        tot = getattr(e, "device_time_total", None)
                        if tot is None:
                            tot = getattr(e, "cuda_time_total", 0)
                        if tot > 0:
                            rows.append((e.key[:70], e.count, tot / max(e.count, 1)))
                    rows.sort(key=lambda r: -r[1] * r[2])
                    emit(event="kernels", precision=prec, bits=bits, L=L,
                         kernels=[dict(name=n, count=c, avg_us=round(u, 2)) for n, c, u in rows[:10]])
                except Exception as ex:  # noqa: BLE001
                    emit(event="profiler_failed", err=repr(ex)[:200])
            if len(outs) < 2:
                del caches
                torch.cuda.empty_cache()
                continue
            d = (outs["fp16"] - outs["fp32"]).abs()
            ref = outs["fp32"].abs()
            emit(event="fp16_vs_fp32", bits=bits, L=L, max_rel=(d.max() / ref.max()).item(),
                 per_head_max_rel=(d.amax(dim=1) / ref.amax(dim=1)).max().item())
            del caches
            torch.cuda.empty_cache()
    emit(event="done")


if __name__ == "__main__":
    main()
