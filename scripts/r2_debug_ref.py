#!/usr/bin/env python
"""debug: K prefill packer, shim vs reference extension (same inputs)"""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import build_ref
from _util import spec, synth
from kvquant_b200 import quant_cuda as shim, cache as kc
ref = build_ref.load()
H, T = 32, 96
for bits in (4, 3):
    sp = spec()
    cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
    t = kc.build_k_lookup_table(cal["k"][0], cal["k"][1], cal["k"][2][0], H, device="cuda:0")
    k = torch.from_numpy(sp.k_tokens(T, seed=21)).cuda().half().float()
    kin = k.view(1, T, H, 128).transpose(1, 2)[0].transpose(1, 2).float().contiguous()   # [H,128,T]
    W = 128 * bits // 32
    outs = {}
    for name, mod in (("shim", shim), ("ref", ref)):
        cache = torch.zeros((H, W, 256), dtype=torch.int32, device="cuda:0")
        resc = kin.clone()
        getattr(mod, "vecquant%dappendvecKsparseParallel" % bits)(cache, t["lut"], kin, resc, t["thr_lower"], t["thr_upper"])
        torch.cuda.synchronize()
        outs[name] = (cache.clone(), resc.clone())
    ca, ra = outs["shim"]; cb, rb = outs["ref"]
    d = (ca != cb)
    print(bits, "cache words differing:", int(d.sum()), "of", d.numel(), "cols with diffs:", torch.nonzero(d.any(0).any(0)).flatten()[:20].tolist())
    dr = (ra != rb)
    print(bits, "rescaled differing:", int(dr.sum()), "max abs diff", float((ra - rb).abs().max()))
    if dr.any():
        i = torch.nonzero(dr)[0].tolist(); print("  first", i, float(ra[tuple(i)]), float(rb[tuple(i)]))
    if d.any():
        i = torch.nonzero(d)[0].tolist(); print("  first word", i, hex(int(ca[tuple(i)]) & 0xffffffff), hex(int(cb[tuple(i)]) & 0xffffffff))
    # run twice with the reference to see if it is deterministic
    cache2 = torch.zeros((H, W, 256), dtype=torch.int32, device="cuda:0"); resc2 = kin.clone()
    getattr(ref, "vecquant%dappendvecKsparseParallel" % bits)(cache2, t["lut"], kin, resc2, t["thr_lower"], t["thr_upper"])
    print(bits, "reference run-to-run differences:", int((cache2 != cb).sum()), int((resc2 != rb).sum()))

# ---- the full caller flow of tests/test_gpu_reference_callers.py, locating the first mismatch -------------------
import warnings
import build_ref_py
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_reference_callers as tc
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    on_shim = build_ref_py.load(shim, "m_shim"); on_ref = build_ref_py.load(ref, "m_ref")
for bits in (4, 3):
    sp = spec()
    cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
    T, ND, Lmax = 96, 64, 256
    k_all = torch.from_numpy(sp.k_tokens(T + ND, seed=21)).cuda().half().float()
    v_all = torch.from_numpy(sp.v_tokens(T + ND, seed=22)).cuda().half().float()
    g = torch.Generator(device="cuda:0").manual_seed(5)
    q_dec = torch.randn((ND, H, 128), generator=g, device="cuda:0")
    args = (bits, cal, k_all[:T], v_all[:T], k_all[T:], v_all[T:], q_dec, Lmax)
    ka, va, sa, oa = tc._drive(on_shim.QuantK, on_shim.QuantV, *args)
    kb, vb, sb, ob = tc._drive(on_ref.QuantK, on_ref.QuantV, *args)
    L = T + ND
    for name, x, y in (("kcache", ka.kcache[:, :, :L], kb.kcache[:, :, :L]), ("vcache", va.vcache[:, :, :L], vb.vcache[:, :, :L]),
                       ("k_idx", ka.outlier_indices[:L], kb.outlier_indices[:L]), ("k_out", ka.outliers[:L], kb.outliers[:L]),
                       ("v_idx", va.outlier_indices[:L], vb.outlier_indices[:L]), ("v_out", va.outliers[:L], vb.outliers[:L]),
                       ("vlut", va.lookup_table[:L], vb.lookup_table[:L]), ("klut", ka.lookup_table, kb.lookup_table)):
        d = x != y
        if d.any():
            nz = torch.nonzero(d)
            dim = -1 if name.endswith("cache") else 0
            slots = sorted(set(nz[:, dim].tolist()))
            print(bits, name, "differs at", int(d.sum()), "places; slots", slots[:12], "first", nz[0].tolist(),
                  x[tuple(nz[0].tolist())].item(), y[tuple(nz[0].tolist())].item())
        else:
            print(bits, name, "equal")
