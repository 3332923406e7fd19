#!/usr/bin/env python
"""GPU probe (run under gpurun): per-kernel timings at full sizes + full-size agreement with the reference's own
CUDA kernels (oracle/_ref/quant_cuda_ref.so, built from /root/reference by oracle/build_ref.py).
Writes gpurun_out/probe.jsonl.  Test/diagnostic infrastructure, not a bench value."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from kvquant_b200 import synth, cache as kc, quant_cuda as qc  # noqa: E402
import build_ref  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
PEAK = 6501.9


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    bits_list = [int(b) for b in os.environ.get("PROBE_BITS", "4,3").split(",")]
    Ls = [int(x) for x in os.environ.get("PROBE_L", "32768,131072").split(",")]
    ref = build_ref.load()
    dev = torch.device("cuda:0")
    H = 32
    sp = synth.SynthSpec(H, 128, seed=0)
    log = open(os.path.join(OUT, "probe.jsonl"), "a")

    def emit(**kw):
        if os.environ.get("PROBE_TAG"):
            kw["tag"] = os.environ["PROBE_TAG"]
        print(json.dumps(kw), flush=True)
        log.write(json.dumps(kw) + "\n"); log.flush()

    emit(event="start", gpu=torch.cuda.get_device_name(0), ref_so=ref is not None)
    for bits in bits_list:
        cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
        klut = kc.build_k_lookup_table(cal["k"][0], cal["k"][1], cal["k"][2][0], H, device=dev)
        for L in Ls:
            Lmax = L + 64
            lc = kc.LayerCache.from_luts(bits, H, Lmax, dict(lut=klut["lut"], lut2=None, thr_lower=klut["thr_lower"],
                                                            thr_upper=klut["thr_upper"]), cal["v"][2][0], device=dev)
            t0 = time.time()
            synth.fill_layer_cache_gpu(lc, sp, L, seed=bits)
            torch.cuda.synchronize()
            emit(event="filled", bits=bits, L=L, secs=round(time.time() - t0, 2))
            q = torch.randn((1, H, 128), device=dev).half().float()
            lutK = lc.klut.view(H, 128, -1)
            mulK = torch.zeros((1, H, L), device=dev)
            p = torch.softmax(torch.randn((1, H, L), device=dev) * 2, -1).half().float()
            mulV = torch.zeros((1, H, 128), device=dev)
            kname = "vecquant%dmatmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2" % bits
            vname = "vecquant%dmatmul_nuq_perchannel_transposed_mha_batched_fused_opt2" % bits
            kd = "vecquant%dmatmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt" % bits
            vd = "vecquant%dmatmul_nuq_perchannel_transposed_mha_batched_fused_opt" % bits
            nout = lc.n_out
            kbytes = L * (H * 128 * bits // 8 + 8 * nout)
            vbytes = L * (H * 128 * bits // 8 + 8 * nout + 4 * 2 ** bits)
            fbytes = L * lc.bytes_per_token()

            def run_k(mod):
                getattr(mod, kname)(q, lc.kcache, mulK, lutK, L, lc.k_outliers, lc.k_outlier_idx, 10000.0, 0)

            def run_v(mod):
                getattr(mod, vname)(p, lc.vcache, mulV, lc.vlut, L, lc.v_outliers, lc.v_outlier_idx)

            def run_kd(mod):
                getattr(mod, kd)(q, lc.kcache, mulK, lutK, L, 10000.0, 0)

            def run_vd(mod):
                getattr(mod, vd)(p, lc.vcache, mulV, lc.vlut, L)

            if os.environ.get("PROBE_QUICK"):   # under ncu: just a few launches of each op
                for _ in range(2):
                    run_k(qc); run_v(qc)
                    lc.attend(q[0].contiguous())
                torch.cuda.synchronize()
                continue
            # ---- correctness at full size against the reference kernels --------------------------------------
            if ref is not None and L <= 1 << 20 and not os.environ.get("PROBE_SKIP_REF"):
                for nm, runner, buf in (("k_opt2", run_k, mulK), ("v_opt2", run_v, mulV), ("k_opt", run_kd, mulK), ("v_opt", run_vd, mulV)):
                    buf.zero_(); runner(qc); ours = buf.clone()
                    buf.zero_(); runner(ref); theirs = buf.clone()
                    d = (ours - theirs).abs().max().item() / max(theirs.abs().max().item(), 1e-30)
                    l2 = ((ours - theirs).norm() / theirs.norm()).item()
                    emit(event="vs_reference_kernel", op=nm, bits=bits, L=L, max_rel=d, rel_l2=l2)
            # ---- timings ------------------------------------------------------------------------------------
            for nm, runner, nbytes in (("k_opt2", run_k, kbytes), ("v_opt2", run_v, vbytes),
                                       ("k_opt", run_kd, L * H * 128 * bits // 8), ("v_opt", run_vd, L * (H * 128 * bits // 8 + 4 * 2 ** bits))):
                med, best = timeit(lambda: runner(qc))
                emit(event="time", impl="ours", op=nm, bits=bits, L=L, ms=med, best_ms=best, gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / PEAK)
                if ref is not None and L <= 1 << 18 and not os.environ.get("PROBE_SKIP_REF"):
                    med, best = timeit(lambda: runner(ref), iters=5, warm=1)
                    emit(event="time", impl="reference_cuda", op=nm, bits=bits, L=L, ms=med, best_ms=best, gbs=nbytes / med / 1e6, frac=nbytes / med / 1e6 / PEAK)
            qa = q[0].contiguous()
            med, best = timeit(lambda: lc.attend(qa))
            emit(event="time", impl="ours", op="attend_fused", bits=bits, L=L, ms=med, best_ms=best, gbs=fbytes / med / 1e6, frac=fbytes / med / 1e6 / PEAK)
            # fused attend vs the legacy two-op chain (same kernels + torch softmax) for a sanity number
            mulK.zero_(); run_k(qc)
            pr = torch.softmax(mulK[0] / np.sqrt(128), -1)
            mulV.zero_()
            getattr(qc, vname)(pr[None].contiguous(), lc.vcache, mulV, lc.vlut, L, lc.v_outliers, lc.v_outlier_idx)
            o = lc.attend(qa)
            emit(event="attend_vs_chain", bits=bits, L=L, max_rel=((o - mulV[0]).abs().max() / mulV[0].abs().max()).item())
            del lc, mulK, p
            torch.cuda.empty_cache()
    emit(event="done")


if __name__ == "__main__":
    main()
