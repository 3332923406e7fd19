"""GPU: the reference's REAL caller code runs on the drop-in boundary.

`QuantK` / `QuantV` are cut verbatim out of the reference's modeling_llama.py (lines 352-975, 978-1385) by
oracle/build_ref_py.py into the git-ignored oracle/_ref/ref_cache_managers.py and loaded twice: once with
`import quant_cuda` resolving to this repository's shim (kvquant_b200.quant_cuda over the C ABI) and once resolving to
the reference's own CUDA extension (oracle/_ref/quant_cuda_ref.so).  Both copies are driven through the reference's
own sequence -- load_lookup_table, one prefill `parallel_pack`, then decode steps of `forward_fused_sparse` chained as
modeling_llama.py:1803-1820 and 1963-1999 do (CPU top-k of v, scores.half()/sqrt(d), fp32 softmax -> fp16, V op) --
and compared: cache words, per-token LUT rows, outlier rows and indices bit for bit; the fp16 results to fp16 accuracy.
The repository's own mirror classes (kvquant_b200.cache.QuantK / QuantV) are driven through the same sequence too.
Skipped when the two oracle/_ref artefacts are absent."""
import math
import os
import sys
import warnings

import numpy as np
import pytest
import torch

from _util import spec, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
DEV = "cuda:0"
H, HID = 32, 4096


@pytest.fixture(scope="module")
def mods():
    import build_ref
    import build_ref_py
    ref_ext = build_ref.load()
    if ref_ext is None or not os.path.exists(build_ref_py.OUT):
        pytest.skip("oracle/_ref/quant_cuda_ref.so or ref_cache_managers.py not present")
    from kvquant_b200 import quant_cuda as shim
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        on_shim = build_ref_py.load(shim, "ref_managers_on_shim")
        on_ref = build_ref_py.load(ref_ext, "ref_managers_on_ref")
    assert on_shim.quant_cuda is shim and on_ref.quant_cuda is ref_ext
    return on_shim, on_ref


def _topk_v(v_flat_gpu, kk):
    """modeling_llama.py:1813-1820: top-k of the new value vector on the CPU, results back on the device."""
    v = v_flat_gpu.cpu()
    uv, ui = torch.topk(v, kk)
    lv, li = torch.topk(v, kk, largest=False)
    return uv.cuda(), ui.cuda(), lv.cuda(), li.cuda()


def _drive(QK, QV, bits, cal, k_pre, v_pre, k_dec, v_dec, q_dec, Lmax):
    """One manager pair through prefill + decode; returns (kmgr, vmgr, scores list, outputs list)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")   # the reference wraps tensors in torch.tensor(...)
        kmgr = QK(bits=bits, hidden_size=HID, num_heads=H, max_position_embeddings=Lmax, include_sparse=True,
                  sparsity_threshold=0.99, rope_theta=10000)
        vmgr = QV(bits=bits, hidden_size=HID, num_heads=H, max_position_embeddings=Lmax, include_sparse=True,
                  sparsity_threshold=0.99)
        kmgr.load_lookup_table(cal["k"], include_sparse=True, sparsity_threshold=0.99)
        vmgr.load_lookup_table(cal["v"], include_sparse=True, sparsity_threshold=0.99)
        kk = int(((1 - 0.99) / 2) * HID) + 2
        T = k_pre.shape[0]
        # prefill (modeling_llama.py:1829-1832, 1907-1927): key_states[0].transpose(1, 2) is [H, 128, T]
        ks = k_pre.view(1, T, H, 128).transpose(1, 2).half()
        vs = v_pre.view(1, T, H, 128).transpose(1, 2).half()
        vf = v_pre.float()
        uv, ui = torch.topk(vf, kk, dim=-1)
        lv, li = torch.topk(vf, kk, dim=-1, largest=False)
        kmgr.parallel_pack(ks[0].transpose(1, 2))
        vmgr.parallel_pack(vs[0].transpose(1, 2), uv, ui, lv, li)
        scores, outs = [], []
        for i in range(k_dec.shape[0]):
            q = q_dec[i].view(H, 1, 128).half()
            kn = k_dec[i].view(1, H, 1, 128).half()
            vn = v_dec[i].view(1, H, 1, 128).half()
            tk = _topk_v(vn.flatten().float(), kk)
            s = kmgr.forward_fused_sparse(q, kn)                               # [H, 1, L] fp16
            scores.append(s.clone())
            a = s.unsqueeze(0) / math.sqrt(128)
            p = torch.nn.functional.softmax(a, dim=-1, dtype=torch.float32).to(torch.float16).squeeze(0)
            o = vmgr.forward_fused_sparse(p, vn, *tk)                          # [H, 1, 128] fp16
            outs.append(o.clone())
    return kmgr, vmgr, scores, outs


@pytest.mark.parametrize("bits", [4, 3])
def test_reference_quantk_quantv_run_unmodified_on_the_shim(mods, bits):
    on_shim, on_ref = mods
    sp = spec()
    cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
    # prefill of 32 tokens: the reference's K prefill packer stages its tables in shared memory without a barrier
    # (quant_cuda_kernel.cu:1857-1883; 128 threads per block); with more than one warp of tokens its output is not
    # reproducible run to run (observed on the B200: 256-320 differing values between two identical calls), so a
    # bit-for-bit comparison against it is only meaningful while one warp owns all the tokens
    T, ND = 32, 64
    Lmax = 256
    k_all = torch.from_numpy(sp.k_tokens(T + ND, seed=21)).to(DEV)
    v_all = torch.from_numpy(sp.v_tokens(T + ND, seed=22)).to(DEV)
    # fp16-representable inputs: the reference feeds .half() activations to both managers
    k_all, v_all = k_all.half().float(), v_all.half().float()
    g = torch.Generator(device=DEV).manual_seed(5)
    q_dec = torch.randn((ND, H, 128), generator=g, device=DEV)
    args = (bits, cal, k_all[:T], v_all[:T], k_all[T:], v_all[T:], q_dec, Lmax)
    ka, va, sa, oa = _drive(on_shim.QuantK, on_shim.QuantV, *args)
    kb, vb, sb, ob = _drive(on_ref.QuantK, on_ref.QuantV, *args)
    L = T + ND
    assert ka.klen == kb.klen == L and va.vlen == vb.vlen == L
    # 3-bit V prefill: the reference kernel indexes the LUT by channel for entries 1..7 (quant_cuda_kernel.cu:2574-2579,
    # a defect our packer does not reproduce) -> compare V codes from the decode-time slots only
    v_lo = T if bits == 3 else 0
    same_k = (torch.equal(ka.kcache[:, :, :L], kb.kcache[:, :, :L]) and torch.equal(ka.outlier_indices[:L], kb.outlier_indices[:L])
              and torch.equal(ka.outliers[:L], kb.outliers[:L]))
    if not same_k:
        # the decode-time slots never depend on the prefill packer: they must agree in any case
        assert torch.equal(ka.kcache[:, :, T:L], kb.kcache[:, :, T:L]), "K cache words of the decode steps differ"
        assert torch.equal(ka.outlier_indices[T:L], kb.outlier_indices[T:L]) and torch.equal(ka.outliers[T:L], kb.outliers[T:L])
        # prefill slots: is the reference reproducible here?  (its packer races, see above)
        kb2, _, _, _ = _drive(on_ref.QuantK, on_ref.QuantV, *args)
        if not (torch.equal(kb.kcache[:, :, :T], kb2.kcache[:, :, :T]) and torch.equal(kb.outliers[:T], kb2.outliers[:T])):
            pytest.skip("the reference's K prefill packer disagreed with itself on this run (shared-memory race, "
                        "quant_cuda_kernel.cu:1857-1883); decode-time state verified bit-identical")
    assert torch.equal(ka.kcache[:, :, :L], kb.kcache[:, :, :L]), "K cache words differ"
    assert torch.equal(va.vcache[:, :, v_lo:L], vb.vcache[:, :, v_lo:L]), "V cache words differ"
    assert torch.equal(va.lookup_table[:L], vb.lookup_table[:L]), "per-token V LUT rows differ"
    assert torch.equal(ka.outlier_indices[:L], kb.outlier_indices[:L]) and torch.equal(ka.outliers[:L], kb.outliers[:L])
    assert torch.equal(va.outlier_indices[:L], vb.outlier_indices[:L]) and torch.equal(va.outliers[:L], vb.outliers[:L])
    if bits == 3:
        return   # the 3-bit V prefill codes differ by the reference defect, so the V outputs do too
    for i in range(ND):
        s1, s2 = sa[i].float(), sb[i].float()
        # fp16 scores: the fp32 results agree to ~1e-6, so at most an occasional one-ulp rounding flip
        assert (s1 - s2).abs().max().item() <= 2.0 ** -10 * max(1.0, s2.abs().max().item()), i
        assert (s1 == s2).float().mean().item() > 0.98, i
        o1, o2 = oa[i].float(), ob[i].float()
        assert (o1 - o2).abs().max().item() <= 2e-3 * max(o2.abs().max().item(), 1e-3), i


@pytest.mark.parametrize("bits", [4, 3, 2])
def test_mirror_classes_equal_the_reference_classes_on_the_shim(mods, bits):
    """kvquant_b200.cache.QuantK / QuantV (device-side top-k, vectorised LUT build) against the reference's own classes,
    both on the shim: same caches bit for bit, same fp16 results (same kernels, same inputs)."""
    on_shim, _ = mods
    from kvquant_b200 import cache as kc
    sp = spec()
    cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
    T, ND = 32, 24
    Lmax = 128
    k_all = torch.from_numpy(sp.k_tokens(T + ND, seed=31)).to(DEV).half().float()
    v_all = torch.from_numpy(sp.v_tokens(T + ND, seed=32)).to(DEV).half().float()
    g = torch.Generator(device=DEV).manual_seed(6)
    q_dec = torch.randn((ND, H, 128), generator=g, device=DEV)
    args = (bits, cal, k_all[:T], v_all[:T], k_all[T:], v_all[T:], q_dec, Lmax)
    ka, va, sa, oa = _drive(on_shim.QuantK, on_shim.QuantV, *args)
    kb, vb, sb, ob = _drive(kc.QuantK, kc.QuantV, *args)
    L = T + ND
    assert torch.equal(ka.lookup_table.view(-1), kb.lookup_table.view(-1)), "K LUT differs"
    assert torch.equal(ka.kcache[:, :, :L], kb.kcache[:, :, :L]) and torch.equal(va.vcache[:, :, :L], vb.vcache[:, :, :L])
    assert torch.equal(va.lookup_table[:L], vb.lookup_table[:L])
    assert torch.equal(ka.outlier_indices[:L], kb.outlier_indices[:L]) and torch.equal(ka.outliers[:L], kb.outliers[:L])
    assert torch.equal(va.outlier_indices[:L], vb.outlier_indices[:L]) and torch.equal(va.outliers[:L], vb.outliers[:L])
    for i in range(ND):
        assert (sa[i].float() - sb[i].float()).abs().max().item() <= 2.0 ** -10 * max(1.0, sb[i].float().abs().max().item())
        assert (oa[i].float() - ob[i].float()).abs().max().item() <= 2e-3 * max(ob[i].float().abs().max().item(), 1e-3)
