"""GPU: our ops against the REFERENCE's own CUDA extension on the same inputs.

oracle/_ref/quant_cuda_ref.so is the unmodified reference (deployment/kvquant/quant_cuda.cpp + quant_cuda_kernel.cu)
compiled for sm_100a by oracle/build_ref.py in the build container; it travels with the gpurun snapshot.  When it is
absent these tests are skipped (the oracle-based tests in test_gpu_parity.py still run).

Bars: packed codes / returned index arrays bit-exact; fp32 element-wise outputs bit-exact; matvecs within 2e-5
norm-wise (both sides accumulate in fp32, in different orders; the reference's own order is non-deterministic).
Reference defects are avoided, not reproduced: the 3-bit V prefill packer (quant_cuda_kernel.cu:2574-2579) and the
racy K prefill packer (1857-1883) are compared through the single-token ops instead.
"""
import os
import sys

import numpy as np
import pytest
import torch

from _util import O, ROOT, quantizer, rel_err, spec

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def ref():
    import build_ref
    m = build_ref.load()
    if m is None:
        pytest.skip("oracle/_ref/quant_cuda_ref.so not present")
    return m


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _both(ref, name):
    import quant_cuda
    return getattr(quant_cuda, name), getattr(ref, name)


@pytest.mark.parametrize("bits", [4, 3, 2])
def test_single_token_appends_bit_exact_vs_reference(ref, bits):
    klut, vcent = quantizer(bits)
    sp = spec()
    H, W, Lmax, T = 32, 128 * bits // 32, 64, 9
    k, v = sp.k_tokens(T, 41), sp.v_tokens(T, 42)
    lut = cu(klut["lut"].reshape(H, 128, -1))
    lo, hi = cu(klut["thr_lower"]), cu(klut["thr_upper"])
    caches = [torch.zeros((H, W, Lmax), dtype=torch.int32, device=DEV) for _ in range(8)]
    vlut = torch.zeros((Lmax, 2 ** bits), dtype=torch.float32, device=DEV)
    for t in range(T):
        kv, vv = cu(k[t]), cu(v[t])
        a, b = _both(ref, "vecquant%dappendvecK" % bits)
        a(caches[0], lut, kv, t); b(caches[1], lut, kv, t)
        a, b = _both(ref, "vecquant%dappendvecKsparse" % bits)
        r1, r2 = kv.clone(), kv.clone()
        a(caches[2], lut, kv, r1, lo, hi, t); b(caches[3], lut, kv, r2, lo, hi, t)
        assert torch.equal(r1, r2)
        thi, tlo, _, _ = O.v_thresholds(v[t], 21)
        lt = O.v_token_lut(vcent, thi, tlo)
        vlut[t] = cu(lt)
        a, b = _both(ref, "vecquant%dappendvecV" % bits)
        a(caches[4], vlut, vv, t); b(caches[5], vlut, vv, t)
        a, b = _both(ref, "vecquant%dappendvecVsparse" % bits)
        zp = float(lt[O.zero_point_code(bits)])
        a(caches[6], vlut, vv, zp, float(tlo), float(thi), t); b(caches[7], vlut, vv, zp, float(tlo), float(thi), t)
    for i in range(0, 8, 2):
        assert torch.equal(caches[i], caches[i + 1]), i


@pytest.mark.parametrize("bits", [4, 2])
def test_prefill_v_packer_bit_exact_vs_reference(ref, bits):
    # (3-bit is excluded: the reference kernel indexes the LUT by channel there, quant_cuda_kernel.cu:2574-2579)
    klut, vcent = quantizer(bits)
    sp = spec()
    H, W, Lmax, T = 32, 128 * bits // 32, 320, 300
    v = sp.v_tokens(T, 43)
    lo = np.zeros(T, np.float32); hi = np.zeros(T, np.float32)
    vlut = np.zeros((Lmax, 2 ** bits), np.float32)
    for t in range(T):
        hi[t], lo[t], _, _ = O.v_thresholds(v[t], 21)
        vlut[t] = O.v_token_lut(vcent, hi[t], lo[t])
    a, b = _both(ref, "vecquant%dappendvecVsparseParallel" % bits)
    c1 = torch.zeros((H, W, Lmax), dtype=torch.int32, device=DEV)
    c2 = torch.zeros_like(c1)
    vin = cu(v.T.reshape(H, 128, T))
    a(c1, cu(vlut), vin, cu(lo), cu(hi)); b(c2, cu(vlut), vin, cu(lo), cu(hi))
    assert torch.equal(c1, c2)


@pytest.mark.parametrize("bits", [4, 3, 2])
def test_matvecs_vs_reference_kernels(ref, bits):
    from _util import oracle_cache
    L = 700
    c, k, v = oracle_cache(bits, L)
    H, W = 32, 128 * bits // 32
    kc, vc = cu(c.kwords.reshape(H, W, c.Lmax)), cu(c.vwords.reshape(H, W, c.Lmax))
    lut = cu(c.klut["lut"].reshape(H, 128, -1))
    q = cu(O.rope_rotate_q(spec().q_vec(7), L + 3, 10000.0)[None])
    p = torch.softmax(torch.randn((1, H, L), device=DEV) * 2, -1).half().float()
    for sparse in (False, True):
        sfx = "2" if sparse else ""
        a, b = _both(ref, "vecquant%dmatmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt%s" % (bits, sfx))
        m1 = torch.zeros((1, H, L), device=DEV); m2 = torch.zeros_like(m1)
        if sparse:
            a(q, kc, m1, lut, L, cu(c.k_out), cu(c.k_idx), 10000.0, 3); b(q, kc, m2, lut, L, cu(c.k_out), cu(c.k_idx), 10000.0, 3)
        else:
            a(q, kc, m1, lut, L, 10000.0, 3); b(q, kc, m2, lut, L, 10000.0, 3)
        assert rel_err(m1.cpu().numpy(), m2.cpu().numpy())[0] < 2e-5
        a, b = _both(ref, "vecquant%dmatmul_nuq_perchannel_transposed_mha_batched_fused_opt%s" % (bits, sfx))
        o1 = torch.zeros((1, H, 128), device=DEV); o2 = torch.zeros_like(o1)
        if sparse:
            a(p, vc, o1, cu(c.vlut), L, cu(c.v_out), cu(c.v_idx)); b(p, vc, o2, cu(c.vlut), L, cu(c.v_out), cu(c.v_idx))
        else:
            a(p, vc, o1, cu(c.vlut), L); b(p, vc, o2, cu(c.vlut), L)
        assert rel_err(o1.cpu().numpy(), o2.cpu().numpy())[0] < 2e-5


def test_uncapped_orig_path_vs_reference(ref):
    """vecquant4appendvec{K,V}sparseorig + ..._opt2_orig: CSR/CSC arrays and results identical to the reference."""
    import quant_cuda
    klut, vcent = quantizer(4)
    sp = spec()
    H, W, Lmax, T = 32, 16, 64, 12
    k, v = sp.k_tokens(T, 51), sp.v_tokens(T, 52)
    lut = cu(klut["lut"].reshape(H, 128, -1))
    lo, hi, zp = cu(klut["thr_lower"]), cu(klut["thr_upper"]), cu(klut["zeropoint"])
    st = {}
    for name, mod in (("ours", quant_cuda), ("ref", ref)):
        kc = torch.zeros((H, W, Lmax), dtype=torch.int32, device=DEV)
        vc = torch.zeros_like(kc)
        vlut = torch.zeros((Lmax, 16), dtype=torch.float32, device=DEV)
        e = lambda: torch.tensor([]).to(DEV)
        rows, cols, vals, start = e(), e(), e(), e()
        vrows, vcols, vvals, vstart = e(), e(), e(), e()
        for t in range(T):
            rows, cols, vals, start, nth, cnt = mod.vecquant4appendvecKsparseorig(kc, lut, cu(k[t]), zp, rows, cols, vals, start, lo, hi, t)
            thi, tlo, _, _ = O.v_thresholds(v[t], 21)
            lt = O.v_token_lut(vcent, thi, tlo)
            vlut[t] = cu(lt)
            vrows, vcols, vvals, vstart, vnth, vcnt = mod.vecquant4appendvecVsparseorig(
                vc, vlut, cu(v[t]), float(lt[7]), vrows, vcols, vvals, vstart, float(tlo), float(thi), t)
        L = T
        q = cu(O.rope_rotate_q(sp.q_vec(3), L, 10000.0)[None])
        mul = torch.zeros((1, H, L), device=DEV)
        mod.vecquant4matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2_orig(
            q, kc, mul, lut, L, rows, cols, start, vals, L, int(nth[0]), int(vals.shape[0]), 10000.0, 0)
        p = torch.softmax(torch.arange(H * L, device=DEV).float().view(1, H, L).sin(), -1)
        out = torch.zeros((1, H, 128), device=DEV)
        mod.vecquant4matmul_nuq_perchannel_transposed_mha_batched_fused_opt2_orig(
            p, vc, out, vlut, L, vrows, vcols, vstart, vvals, L, int(vnth[0]), int(vvals.shape[0]))
        st[name] = dict(kc=kc, vc=vc, rows=rows, cols=cols, vals=vals, start=start, nth=int(nth[0]), vrows=vrows,
                        vcols=vcols, vvals=vvals, vstart=vstart, vnth=int(vnth[0]), mul=mul, out=out)
    a, b = st["ours"], st["ref"]
    for key in ("kc", "vc", "rows", "cols", "vals", "start", "vrows", "vcols", "vvals", "vstart"):
        assert torch.equal(a[key].cpu(), b[key].cpu()), key
    assert a["nth"] == b["nth"] and a["vnth"] == b["vnth"]
    assert rel_err(a["mul"].cpu().numpy(), b["mul"].cpu().numpy())[0] < 2e-5
    assert rel_err(a["out"].cpu().numpy(), b["out"].cpu().numpy())[0] < 2e-5
