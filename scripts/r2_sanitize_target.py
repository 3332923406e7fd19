#!/usr/bin/env python
"""Workload for compute-sanitizer (scripts/r2_sanitize.sh): fused append + fused attend (both table precisions, sinks,
device-resident length) + the legacy K / V ops, at 4 K tokens for 4 / 3 / 2 bits, and a dense-only cache."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kvquant_b200 import synth, cache as kc, quant_cuda as qc  # noqa: E402

DEV = "cuda:0"
H = 32
L = int(os.environ.get("SAN_L", "4096"))
for bits, sparse in ((4, True), (3, True), (2, True), (4, False)):
    sp = synth.SynthSpec(H, 128, seed=0)
    cal = synth.calibrate(sp, bits, calib_tokens=256, seed=7)
    klut = kc.build_k_lookup_table(cal["k"][0], cal["k"][1], cal["k"][2][0], H, device=DEV)
    lc = kc.LayerCache.from_luts(bits, H, L + 64, dict(lut=klut["lut"], lut2=None, thr_lower=klut["thr_lower"],
                                                      thr_upper=klut["thr_upper"]), cal["v"][2][0], device=DEV,
                                 include_sparse=sparse, n_sink=3)
    lc.set_sinks(torch.randn((H, 128, 3), device=DEV).half(), torch.randn((H, 3, 128), device=DEV).half())
    synth.fill_layer_cache_gpu(lc, sp, L, seed=bits, chunk=2048)
    q = torch.randn((H, 128), device=DEV).half().float()
    for t in range(3):                                       # fused device append (radix select, pack, outlier rows)
        lc.append(torch.randn(H * 128, device=DEV), torch.randn(H * 128, device=DEV))
    len_dev = torch.full((1,), lc.len - 1, dtype=torch.int64, device=DEV)
    for prec in ("fp32", "fp16"):
        lc.precision = prec
        a = lc.attend(q).clone()
        b = lc.attend_dyn(q, len_dev, 1).clone()
        torch.cuda.synchronize()
        assert torch.isfinite(a).all() and (a - b).abs().max() < 1e-3 * a.abs().max()
    if sparse:
        mulK = torch.zeros((1, H, lc.len), device=DEV)
        getattr(qc, "vecquant%dmatmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2" % bits)(
            q[None].contiguous(), lc.kcache, mulK, lc.klut.view(H, 128, -1), lc.len, lc.k_outliers, lc.k_outlier_idx, 10000.0, 3)
        p = torch.softmax(mulK / 11.3, -1)
        mulV = torch.zeros((1, H, 128), device=DEV)
        getattr(qc, "vecquant%dmatmul_nuq_perchannel_transposed_mha_batched_fused_opt2" % bits)(
            p, lc.vcache, mulV, lc.vlut, lc.len, lc.v_outliers, lc.v_outlier_idx)
    torch.cuda.synchronize()
    print("ok", bits, sparse, flush=True)
print("SANITIZE_TARGET_DONE")
